"""Timeline of the data-parallel step inside the captured graph (no nsys in this image: torch.profiler / CUPTI on graph
replays).  Run under torchrun with N >= 2; rank 0 prints where the NCCL kernels sit relative to the compute kernels and how
much the compute kernels that overlap them are slowed down.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/profile_ddp.py
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import bench
from open_flamingo_b200.testing import build_flamingo, synthetic_batch
from open_flamingo_b200.train import FlatTrainer, GraphedTrainStep


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        from open_flamingo_b200.train import configure_nccl_for_overlap
        configure_nccl_for_overlap()
        dist.init_process_group("nccl", device_id=dev)
    vit_cfg, mpt_kw, every = bench.model_dims("of3b")
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model, _, tok = build_flamingo(vit_cfg, mpt_kw, cross_attn_every_n_layers=every, device=dev, gate_init=1.0)
    model.train()
    media_id, eoc_id = tok.encode("<image>")[-1], tok.encode("<|endofchunk|>")[-1]
    trainer = FlatTrainer(model, num_chunks=int(os.environ.get("CHUNKS", "6")))
    batch = {k: v.to(dev) for k, v in synthetic_batch(32, 2, 256, media_id, eoc_id, mpt_kw["vocab_size"], image_size=224,
                                                      seed=100 + rank).items()}
    g = GraphedTrainStep(model, trainer, batch, warmup=3)
    assert g.ok, g.error
    for _ in range(3):
        g(batch)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        g(batch)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank != 0:
        torch.cuda.synchronize()
        os._exit(0)
    evs = []
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and ev.time_range is not None:
            dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
            evs.append((ev.time_range.start, ev.time_range.start + dur, dur, ev.name))
    evs.sort()
    t0 = evs[0][0]
    t_end = max(e[1] for e in evs)
    nccl = [e for e in evs if "nccl" in e[3].lower()]
    comp = [e for e in evs if "nccl" not in e[3].lower() and not e[3].startswith("Mem")]
    print(f"step wall (first kernel start -> last kernel end): {(t_end - t0) / 1e3:.2f} ms; {len(comp)} compute kernels "
          f"({sum(e[2] for e in comp) / 1e3:.2f} ms summed), {len(nccl)} NCCL kernels ({sum(e[2] for e in nccl) / 1e3:.2f} ms summed)")
    for s, e, d, n in nccl:
        inside = [c for c in comp if c[0] < e and c[1] > s]
        busy = sum(min(c[1], e) - max(c[0], s) for c in inside)
        print(f"  NCCL {n[:44]:44s} start {(s - t0) / 1e3:8.2f} ms  dur {d / 1e3:6.2f} ms  compute kernels overlapping: {len(inside):4d} "
              f"covering {100 * busy / max(d, 1):5.1f} % of it")
    # slowdown of the compute kernels that run beside an all-reduce: same kernel name (template instance), inside vs outside
    def in_window(c):
        return any(c[0] < e and c[1] > s for s, e, _, _ in nccl)
    by = collections.defaultdict(lambda: [[], []])
    for c in comp:
        by[c[3][:60]][1 if in_window(c) else 0].append(c[2])
    print("kernel (first 60 chars)                                       outside: n, mean us | beside NCCL: n, mean us | ratio")
    rows = []
    for name, (out, ins) in by.items():
        if out and ins and sum(ins) > 200:
            rows.append((sum(ins), name, len(out), sum(out) / len(out), len(ins), sum(ins) / len(ins)))
    extra = 0.0
    for tot, name, no, mo, ni, mi in sorted(rows, reverse=True)[:14]:
        print(f"  {name:60s} {no:4d} {mo:8.1f} | {ni:4d} {mi:8.1f} | {mi / mo:5.2f}")
    for tot, name, no, mo, ni, mi in rows:
        extra += ni * (mi - mo)
    print(f"estimated extra compute time from running beside NCCL (same-kernel comparison): {extra / 1e3:.2f} ms")
    # idle gaps on the compute side
    gaps = []
    last_end = comp[0][1]
    for c in comp[1:]:
        if c[0] - last_end > 20:
            gaps.append((c[0] - last_end, (last_end - t0) / 1e3, c[3][:40]))
        last_end = max(last_end, c[1])
    print(f"compute-side idle gaps > 20 us: {len(gaps)}, total {sum(g_[0] for g_ in gaps) / 1e3:.2f} ms; largest:")
    for gap, at, nxt in sorted(gaps, reverse=True)[:8]:
        print(f"  {gap / 1e3:6.2f} ms at {at:8.2f} ms before {nxt}")
    os._exit(0)


if __name__ == "__main__":
    main()
