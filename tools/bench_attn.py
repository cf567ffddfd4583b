"""Per-shape timing of the attention cores: tcgen05 path vs the mma.sync path on the same inputs (CUDA events, inputs
rotated through > L2).  Usage: python tools/bench_attn.py [--iters N] [--only fwd|bwd] [--ncu]  (--ncu: 2 iterations of
the tensor-core path only, for `ncu -k regex:attn_.*_tc_kernel`)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from open_flamingo_b200 import _lib as L
from open_flamingo_b200 import ops

bf16 = torch.bfloat16

# name, kind, B, heads, hd, nq, nk, (mask_mode, kpm) | causal
SHAPES = [
    ("xattn C2 (32x256 q, 2x64 keys, eq mask)", "media", 32, 8, 64, 256, 128, 1),
    ("xattn C4 (8x512 q, 5x64 keys, eq mask)", "media", 8, 8, 64, 512, 320, 1),
    ("perceiver C2 (64 img: 64 q, 320 keys)", "media", 64, 8, 64, 64, 320, 0),
    ("perceiver C5 (64 img: 64 q, 4160 keys)", "media", 64, 8, 64, 64, 4160, 0),
    ("ViT-L/14 (64 img, 16 heads, 257 tokens)", "media", 64, 16, 64, 257, 257, 0),
    ("LM MPT-1B (32 x 256, 16 heads, hd 128)", "dense", 32, 16, 128, 256, 256, 1),
    ("LM MPT-7B (8 x 512, 32 heads, hd 128)", "dense", 8, 32, 128, 512, 512, 1),
]


def make(kind, B, heads, hd, nq, nk, flag, nbuf):
    D = heads * hd
    bufs = []
    for i in range(nbuf):
        torch.manual_seed(i)
        if kind == "dense":
            qkv = torch.randn(B, nq, 3 * D, device="cuda").to(bf16)
            q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
        else:
            q = torch.randn(B, nq, D, device="cuda").to(bf16)
            kv = torch.randn(B, nk, 2 * D, device="cuda").to(bf16)
            k, v = kv[..., :D], kv[..., D:]
        d_o = torch.randn(B, nq, D, device="cuda").to(bf16)
        bufs.append((q, k, v, d_o))
    tt = None
    if kind == "media" and flag:
        n_media = nk // 64
        loc = torch.zeros(B, nq, dtype=torch.bool, device="cuda")
        for m in range(n_media):
            loc[:, (m * nq) // n_media] = True
        tt = ops.text_time(media_locations=loc)
    slopes = (2.0 ** (-8.0 * torch.arange(1, heads + 1, device="cuda", dtype=torch.float32) / heads)) if kind == "dense" else None
    return bufs, tt, slopes


def run_fwd(kind, heads, hd, buf, tt, slopes, flag):
    q, k, v, _ = buf
    if kind == "dense":
        return ops.attn_dense_fwd(q, k, v, heads, hd, hd ** -0.5, causal=True, slopes=slopes)
    return ops.attn_fwd(q, k, v, heads, hd ** -0.5, mask_mode=L.MASK_MEDIA_EQ if flag else L.MASK_NONE, text_time=tt)


def run_bwd(kind, heads, hd, buf, o, lse, tt, slopes, flag, grads):
    q, k, v, d_o = buf
    if kind == "dense":
        D = heads * hd
        return ops.attn_dense_bwd(q, k, v, o, d_o, lse, heads, hd, hd ** -0.5, causal=True, slopes=slopes,
                                  dq=grads[..., :D], dk=grads[..., D:2 * D], dv=grads[..., 2 * D:])
    return ops.attn_bwd(q, k, v, o, d_o, lse, heads, hd ** -0.5, mask_mode=L.MASK_MEDIA_EQ if flag else L.MASK_NONE,
                        text_time=tt, dq=grads[0], dk=grads[1], dv=grads[2])


def timed(fn, iters):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--ncu", action="store_true")
    ap.add_argument("--shapes", default="")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rows = []
    for si, (name, kind, B, heads, hd, nq, nk, flag) in enumerate(SHAPES):
        if args.shapes and str(si) not in args.shapes.split(","):
            continue
        D = heads * hd
        nbuf = 1 if args.ncu else 4
        bufs, tt, slopes = make(kind, B, heads, hd, nq, nk, flag, nbuf)
        if kind == "dense":
            grads = torch.empty(B, nq, 3 * D, device="cuda", dtype=bf16)
        else:
            grads = (torch.empty(B, nq, D, device="cuda", dtype=bf16), torch.empty(B, nk, D, device="cuda", dtype=bf16),
                     torch.empty(B, nk, D, device="cuda", dtype=bf16))
        causal_frac = 0.5 * (1 + 1 / max(1, nq // 128)) if kind == "dense" else 1.0
        keys_seen = 64 if (kind == "media" and flag) else nk
        flops_f = 4.0 * B * heads * nq * keys_seen * hd * (causal_frac if kind == "dense" else 1.0)
        bytes_f = 2.0 * (B * nq * D * 2 + 2 * B * nk * D)                       # q, o + k, v (bf16)
        res = {"shape": name}
        for impl in (("tc",) if args.ncu else ("tc", "legacy")):
            prev = ops.attn_force_legacy(impl == "legacy")
            try:
                outs = [run_fwd(kind, heads, hd, b, tt, slopes, flag) for b in bufs]
                if args.ncu:
                    run_bwd(kind, heads, hd, bufs[0], outs[0][0], outs[0][1], tt, slopes, flag, grads)
                    torch.cuda.synchronize()
                    continue
                tf = timed(lambda i: run_fwd(kind, heads, hd, bufs[i % nbuf], tt, slopes, flag), args.iters)
                tb = timed(lambda i: run_bwd(kind, heads, hd, bufs[i % nbuf], outs[i % nbuf][0], outs[i % nbuf][1], tt, slopes,
                                             flag, grads), args.iters)
            finally:
                ops.attn_force_legacy(prev)
            res[impl] = {"fwd_us": round(tf, 1), "bwd_us": round(tb, 1), "fwd_TFLOPs": round(flops_f / tf / 1e6, 1),
                         "bwd_TFLOPs": round(2.5 * flops_f / tb / 1e6, 1), "fwd_GBs": round(bytes_f / tf / 1e3, 0)}
        if not args.ncu:
            print(f"{name:44s} fwd tc {res['tc']['fwd_us']:7.1f} us ({res['tc']['fwd_TFLOPs']:6.1f} TF/s, {res['tc']['fwd_GBs']:5.0f} GB/s)"
                  f"  legacy {res['legacy']['fwd_us']:7.1f} | bwd tc {res['tc']['bwd_us']:7.1f} us ({res['tc']['bwd_TFLOPs']:6.1f} TF/s)"
                  f"  legacy {res['legacy']['bwd_us']:7.1f}", flush=True)
        rows.append(res)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
