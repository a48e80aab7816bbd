"""ncu --set full report -> the per-kernel summary table under profiles/ (+ the dram traffic row bench.py reads).

    ncu -i REP.ncu-rep --page raw --csv > raw.csv
    python profiles/extract_ncu_full.py raw.csv profiles/rN_ncu_full_top_kernels.csv "header comment" [--traffic c2 16 f32]
"""
import csv
import hashlib
import json
import os
import sys

COLS = ["Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    raw, out, comment = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(l for l in open(raw) if not l.startswith("==")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(c) for c in COLS if c in hdr]
    with open(out, "w", newline="") as f:
        f.write(f'"# {comment}"\n')
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in data:
            w.writerow([r[i] for i in idx])
    if "--traffic" in sys.argv:
        k = sys.argv.index("--traffic")
        workload, episodes, dtype = sys.argv[k + 1], int(sys.argv[k + 2]), sys.argv[k + 3]
        kern = "pack_f32_vec_kernel" if dtype == "f32" else "pack_u8_vec_kernel"
        ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        best = None
        for r in data:
            if r[hdr.index("Kernel Name")].startswith(kern):
                best = r  # the last captured launch of the kernel
        if best is not None:
            sys.path.insert(0, ROOT)
            import bench  # the same hash bench.py checks

            sha = bench.kernel_source_sha16(kern)
            path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            table = [row for row in json.load(open(path))
                     if (row["kernel"], row["workload"], row["episodes_per_launch"], row["mask_dtype"]) != (kern, workload, episodes, dtype)]
            table.insert(0, {"kernel": kern, "workload": workload, "episodes_per_launch": episodes, "mask_dtype": dtype,
                             "dram_read_bytes": int(to_bytes(best[ir], units[ir])), "dram_write_bytes": int(to_bytes(best[iw], units[iw])),
                             "duration_ms_under_ncu": float(best[it]) * {"ms": 1, "us": 1e-3, "ns": 1e-6}[units[it]],
                             "source": f"{os.path.basename(out)} ({comment})", "source_sha16": sha, "source_file": "csrc/masks.cu"})
            json.dump(table, open(path, "w"), indent=1)
            print("traffic row:", table[0])


if __name__ == "__main__":
    main()
