"""K-sweep of the tcgen05 GEMM at M=N=8192: separates the per-tile fixed cost (intercept) from the per-k-block
cost (slope) for a few epilogues and both kernels (block_n=256: 1-CTA, 512: 2-CTA)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from open_flamingo_b200 import _lib as L
from open_flamingo_b200 import ops

dev, bf16 = "cuda", torch.bfloat16
M = N = 8192


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


out16 = torch.empty(M, N, device=dev, dtype=bf16)
out16b = torch.empty(M, N, device=dev, dtype=bf16)
out32 = torch.zeros(M, N, device=dev)
print("K, epi, bn, us, TFLOP/s, us_per_tile_wave")
KS = [int(k) for k in sys.argv[1].split(',')] if len(sys.argv) > 1 else [64, 256, 512, 1024, 2048, 4096, 8192]
for K in KS:
    a = torch.randn(M, K, device=dev, dtype=bf16)
    b = torch.randn(N, K, device=dev, dtype=bf16)
    for bn in (256, 512):
        tiles = (M // 128) * (N // 256) if bn == 256 else (M // 256) * (N // 256)
        waves = -(-tiles // (148 if bn == 256 else 74))
        for name, kw in (("store_bf16", dict(epi=L.EPI_STORE_BF16, out=out16)),
                         ("gelu_dual", dict(epi=L.EPI_GELU_DUAL, out=out16, out2=out16b)),
                         ("store_f32", dict(epi=L.EPI_STORE_F32, out=out32))):
            us = timeit(lambda: ops.gemm(a, b, block_n=bn, **kw))
            print(f"{K:5d}, {name:10s}, {bn}, {us:8.1f}, {2*M*N*K/us/1e6:7.1f}, {us/waves:6.2f}", flush=True)
    us = timeit(lambda: torch.matmul(a, b.t()))
    print(f"{K:5d}, cublas    ,   0, {us:8.1f}, {2*M*N*K/us/1e6:7.1f}", flush=True)
