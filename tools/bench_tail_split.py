"""A/B of the GEMM tail split (ofk_gemm_bf16_ws) at the OF-3B shapes whose 256 x 256 tile count leaves the last round
of the 74-pair persistent grid at most half full.  Run twice: OFK_GEMM_TAIL_SPLIT=0 / 1 (read once per process)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from open_flamingo_b200 import _lib as L
from open_flamingo_b200 import ops

bf16 = torch.bfloat16


def timeit(fn, iters=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


print("tail_split =", os.environ.get("OFK_GEMM_TAIL_SPLIT", "1"))
for (M, N, K) in [(8192, 2048, 8192), (8192, 2048, 6144), (8192, 2048, 4096), (8192, 2048, 2048), (16448, 1024, 4096),
                  (8192, 8192, 2048)]:
    a = torch.randn(M, K, device="cuda", dtype=bf16)
    b = torch.randn(N, K, device="cuda", dtype=bf16)
    bt = b.t().contiguous()
    resid = torch.randn(M, N, device="cuda")
    gate = torch.tensor([0.5], device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=bf16)
    o32 = torch.empty(M, N, device="cuda")
    tiles = ((M + 255) // 256) * ((N + 255) // 256)
    for name, fn in (("store_bf16", lambda: ops.gemm(a, b, out=o16)),
                     ("store_bf16 b_mn", lambda: ops.gemm(a, bt, b_mn=True, out=o16)),
                     ("gate_resid_f32", lambda: ops.gemm(a, b, epi=L.EPI_GATE_RESID_F32, aux=resid, gate=gate, out=o32))):
        us = timeit(fn)
        print(f"M={M} N={N} K={K} tiles={tiles} rounds={tiles/74:.2f} {name:16s} {us:8.1f} us {2*M*N*K/us/1e6:7.1f} TF/s", flush=True)
    us = timeit(lambda: torch.matmul(a, b.t()))
    print(f"M={M} N={N} K={K} cublas                                   {us:8.1f} us {2*M*N*K/us/1e6:7.1f} TF/s", flush=True)
