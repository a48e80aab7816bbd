"""End-to-end episodes/s from float32 proposals in pinned HOST memory (c2, every input of the step crosses PCIe inside
   the timed loop) for different shares of the proposals sent raw over PCIe (the rest is packed by host threads first).
   raw_fraction = 1.0 is the plain path (everything raw, packed on the device)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marsb200

dev = torch.device("cuda:0")
shape = marsb200.CONFIGS["c2"]
Ee = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfg = marsb200.RankingConfig(nms_iou_threshold=0.7)
batch = marsb200.stack_episodes([marsb200.make_episode(shape, i, dev) for i in range(Ee)])
host = {k: v.cpu().pin_memory() for k, v in batch.items()}
ref_eng = marsb200.RankingEngine(shape, Ee, cfg, dev)
ref = ref_eng.run(batch)
ref_rec = ref_eng.records().clone()
torch.cuda.synchronize()
del batch

for frac, threads in ((1.0, 0), (0.0, 0), (0.1, 0), (0.15, 0), (0.2, 0), (0.25, 0), (0.3, 0), (0.35, 0), (0.2, 15), (0.2, 24)):
    engs = [marsb200.RankingEngine(shape, Ee, cfg, dev) for _ in range(2)]
    ings = [marsb200.HostMaskIngest(Ee, shape.P, shape.H, shape.W, dev, raw_fraction=frac, threads=threads) for _ in range(2)]
    others = [k for k in host if k != "masks"]
    dev_in = [{k: torch.empty_like(host[k], device=dev) for k in others} for _ in range(2)]
    rec_host = [torch.empty((Ee, engs[0].record_bytes()), dtype=torch.uint8).pin_memory() for _ in range(2)]
    main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    ev_in, ev_done, ev_out = ([torch.cuda.Event() for _ in range(2)] for _ in range(3))

    def upload(i):
        b = i % 2
        s_in.wait_event(ev_done[b])
        with torch.cuda.stream(s_in):
            for k in others:
                dev_in[b][k].copy_(host[k], non_blocking=True)
        dev_in[b]["mask_bits"] = ings[b].upload(host["masks"], s_in)
        ev_in[b].record(s_in)

    def loop(n):
        for b in range(2):
            ev_done[b].record(main)
        upload(0)
        for i in range(n):
            b = i % 2
            main.wait_event(ev_in[b])
            main.wait_event(ev_out[b])
            engs[b].run(dev_in[b])
            rec = engs[b].records()
            ev_done[b].record(main)
            s_out.wait_event(ev_done[b])
            with torch.cuda.stream(s_out):
                rec_host[b].copy_(rec, non_blocking=True)
                ev_out[b].record(s_out)
            if i + 1 < n:
                upload(i + 1)  # the host packs step i + 1 while the device ranks step i
            if i >= 1:
                ev_out[(i - 1) % 2].synchronize()
        ev_out[(n - 1) % 2].synchronize()

    for b in range(2):
        ev_out[b].record(s_out)
    loop(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loop(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = torch.equal(rec_host[(steps - 1) % 2], ref_rec.cpu())
    h2d = ings[0].h2d_bytes() + sum(host[k].numel() * host[k].element_size() for k in others)
    print(f"raw_fraction {frac:.2f} threads {threads or 'all'}: {Ee * steps / dt:7.1f} episodes/s, {dt / steps * 1e3:6.2f} ms per step of {Ee}, "
          f"PCIe {h2d * steps / dt / 1e9:5.1f} GB/s, records equal the device-resident run: {ok}", flush=True)
    del engs, ings, dev_in
