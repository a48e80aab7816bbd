"""Static evidence of the shipped library, produced WITHOUT a GPU: per kernel the registers / spills / static shared memory ptxas
reports (`csrc/build/*.ptxas.log`, written by `make`) and the Blackwell SASS mnemonics found in `cuobjdump -sass libmarsb200.so`
(UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor loads, SYNCS = mbarrier, POPC, REDG / ATOMG).

    python profiles/static_build_report.py > profiles/r2_static_build_report.txt
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mars-multimodal-alignment-and-ranking-system-for-few-shot-segmentation_b200")
WATCH = ("UTCHMMA", "UTCIMMA", "UTCOMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "POPC",
         "REDG", "ATOMG", "ATOMS", "SHFL", "LDGSTS", "HMMA", "IMMA")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return {m: re.sub(r"\(.*", "", d).replace("marsb200::", "") for m, d in zip(names, out)}


def ptxas_rows():
    rows = {}
    for log in sorted(glob.glob(os.path.join(PKG, "csrc", "build", "*.ptxas.log"))):
        src = os.path.basename(log).replace(".ptxas.log", ".cu")
        name = None
        for line in open(log):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                name = m.group(1)
                rows[name] = {"src": src, "regs": None, "spill_st": 0, "spill_ld": 0, "smem": 0}
                continue
            if name is None:
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m:
                rows[name]["spill_st"], rows[name]["spill_ld"] = int(m.group(2)), int(m.group(3))
            m = re.search(r"Used (\d+) registers", line)
            if m:
                rows[name]["regs"] = int(m.group(1))
                s = re.search(r"(\d+) bytes smem", line)
                rows[name]["smem"] = int(s.group(1)) if s else 0
    return rows


def sass_counts():
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(PKG, "libmarsb200.so")], capture_output=True, text=True, check=True).stdout
    counts, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            counts[name]["_total"] += 1
            for w in WATCH:
                if op.startswith(w):
                    counts[name][w] += 1
                    break
    return counts


def main():
    rows, counts = ptxas_rows(), sass_counts()
    names = demangle(sorted(rows))
    print("# libmarsb200.so, sm_100a: ptxas resources and SASS mnemonics per kernel (static, no GPU)")
    print("# kernel | source | registers | spill st/ld bytes | static smem | SASS instructions | watched mnemonics")
    for mangled in sorted(rows, key=lambda k: (rows[k]["src"], names[k])):
        r, c = rows[mangled], counts.get(mangled, {})
        seen = ", ".join(f"{w} x{c[w]}" for w in WATCH if c.get(w))
        print(f"{names[mangled]} | {r['src']} | {r['regs']} | {r['spill_st']}/{r['spill_ld']} | {r['smem']} | {c.get('_total', 0)} | {seen}")
    total = collections.Counter()
    for c in counts.values():
        total.update(c)
    print("# whole library: " + ", ".join(f"{w} x{total[w]}" for w in WATCH if total.get(w)))
    spilled = [names[k] for k in rows if rows[k]["spill_st"] or rows[k]["spill_ld"]]
    print("# kernels with register spills: " + (", ".join(sorted(spilled)) or "none"))


if __name__ == "__main__":
    main()
