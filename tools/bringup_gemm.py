"""GPU bring-up for the tcgen05 GEMM: every operand-major combination, tails, epilogues, split-K, timing.
Run on the B200 box:  timeout 300 python tools/bringup_gemm.py
"""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from open_flamingo_b200 import _lib as L
from open_flamingo_b200 import ops

torch.manual_seed(0)
dev = "cuda"
bf16, f32 = torch.bfloat16, torch.float32
fails = 0


def report(name, got, ref, tol):
    global fails
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item() + 1e-6
    ok = err <= tol * scale
    if not ok:
        fails += 1
    print(f"{'OK ' if ok else 'BAD'} {name:60s} max_abs_err={err:.4e} ref_max={scale:.3e}", flush=True)


def ref_mm(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    B = b.float().t() if b_mn else b.float()
    return A @ B.t()


def run_major(M, N, K, a_mn, b_mn, bn):
    a = torch.randn((K, M) if a_mn else (M, K), device=dev, dtype=bf16)
    b = torch.randn((K, N) if b_mn else (N, K), device=dev, dtype=bf16)
    out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, epi=L.EPI_STORE_F32, block_n=bn)
    torch.cuda.synchronize()
    report(f"f32 M{M} N{N} K{K} a_mn{int(a_mn)} b_mn{int(b_mn)} bn{bn}", out, ref_mm(a, b, a_mn, b_mn), 2e-3)


print("device:", torch.cuda.get_device_name(0), flush=True)
# 1. smallest sanity first
run_major(128, 128, 64, False, False, 128)
run_major(128, 256, 64, False, False, 256)
run_major(128, 256, 256, False, False, 256)
for a_mn in (False, True):
    for b_mn in (False, True):
        for bn in (128, 256, 512):
            run_major(256, 512, 512, a_mn, b_mn, bn)
# tails: M not multiple of 128, K not multiple of 64, N multiple of 16 only
run_major(200, 272, 328, False, False, 128)
run_major(200, 272, 328, False, False, 256)
run_major(200, 272, 328, False, True, 256)
run_major(200, 272, 328, True, True, 128)
run_major(32, 512, 2048, False, False, 256)
run_major(16448, 1024, 640, False, False, 256)
# many tiles per CTA (persistence + phase wrap)
run_major(4096, 4096, 1024, False, False, 256)
run_major(4096, 4096, 1024, False, True, 256)
run_major(2048, 4096, 4096, True, True, 256)
for a_mn_, b_mn_ in ((False, False), (False, True), (True, True)):
    run_major(4096, 4096, 1024, a_mn_, b_mn_, 512)
    run_major(1000, 768, 1096, a_mn_, b_mn_, 512)

# epilogues
M, N, K = 512, 1024, 512
a = torch.randn(M, K, device=dev, dtype=bf16)
b = torch.randn(N, K, device=dev, dtype=bf16) * 0.05
acc = a.float() @ b.float().t()
report("epi STORE_BF16", ops.gemm(a, b), acc.to(bf16), 1e-2)
bias = torch.randn(N, device=dev)
report("epi BIAS_BF16", ops.gemm(a, b, epi=L.EPI_BIAS_BF16, bias=bias), (acc + bias).to(bf16), 1e-2)
t = (acc + bias).to(bf16).float()
report("epi BIAS_QGELU_BF16", ops.gemm(a, b, epi=L.EPI_BIAS_QGELU_BF16, bias=bias), t * torch.sigmoid(1.702 * t), 1e-2)
z = torch.empty(M, N, device=dev, dtype=bf16)
h = torch.empty(M, N, device=dev, dtype=bf16)
ops.gemm(a, b, epi=L.EPI_GELU_DUAL, out=z, out2=h)
report("epi GELU_DUAL z", z, acc.to(bf16), 1e-2)
report("epi GELU_DUAL h", h, torch.nn.functional.gelu(acc.to(bf16).float()), 1e-2)
resid = torch.randn(M, N, device=dev)
gate = torch.tensor([0.7], device=dev)
br = torch.empty(M, N, device=dev, dtype=bf16)
o = ops.gemm(a, b, epi=L.EPI_GATE_RESID_F32, aux=resid, gate=gate, out2=br)
report("epi GATE_RESID_F32", o, acc.to(bf16).float() * math.tanh(0.7) + resid, 1e-2)
report("epi GATE_RESID_F32 branch", br, acc.to(bf16), 1e-2)
o = ops.gemm(a, b, epi=L.EPI_GATE_RESID_F32, aux=resid)
report("epi RESID_F32 (no gate)", o, acc.to(bf16).float() + resid, 1e-2)
o = ops.gemm(a, b, epi=L.EPI_BIAS_RESID_F32, aux=resid, bias=bias)
report("epi BIAS_RESID_F32", o, (acc + bias).to(bf16).float() + resid, 1e-2)
zz = torch.randn(M, N, device=dev, dtype=bf16)
zf = zz.float().requires_grad_(True)
torch.nn.functional.gelu(zf).sum().backward()
report("epi DGELU_BF16", ops.gemm(a, b, epi=L.EPI_DGELU_BF16, aux=zz), acc.to(bf16).float() * zf.grad, 1e-2)
# atomic / split-K
for splits in (1, 2, 4, 8):
    o = torch.ones(M, N, device=dev)
    ops.gemm(a, b, epi=L.EPI_ATOMIC_F32, out=o, splits=splits)
    report(f"epi ATOMIC_F32 splits={splits}", o, acc + 1.0, 2e-3)
# wgrad-shaped: dW[N_out, K_in] += dY[R, N_out]^T X[R, K_in]
R, NO, KI = 8192, 512, 2048
dy = torch.randn(R, NO, device=dev, dtype=bf16) * 0.1
x = torch.randn(R, KI, device=dev, dtype=bf16)
dw = torch.zeros(NO, KI, device=dev)
ops.gemm(dy, x, a_mn=True, b_mn=True, epi=L.EPI_ATOMIC_F32, out=dw, splits=8)
report("wgrad split8 [512,2048] K=8192", dw, dy.float().t() @ x.float(), 2e-3)

# timing (L2-cold is not attempted here; big shapes exceed nothing -- this is bring-up only)
def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for (M, N, K, a_mn, b_mn, epi, name) in [
    (8192, 8192, 2048, False, False, L.EPI_GELU_DUAL, "ffn1 fwd gelu_dual"),
    (8192, 2048, 8192, False, False, L.EPI_GATE_RESID_F32, "ffn2 fwd gate_resid"),
    (8192, 8192, 2048, False, True, L.EPI_DGELU_BF16, "ffn2 dgrad dgelu"),
    (8192, 2048, 8192, False, True, L.EPI_STORE_BF16, "ffn1 dgrad"),
    (8192, 2048, 8192, True, True, L.EPI_ATOMIC_F32, "ffn1 wgrad (M=4D,N=D,K=R)"),
    (2048, 8192, 8192, True, True, L.EPI_ATOMIC_F32, "ffn2 wgrad (M=D,N=4D,K=R)"),
    (8192, 8192, 8192, False, False, L.EPI_STORE_BF16, "square 8192 bf16"),
]:
    a = torch.randn((K, M) if a_mn else (M, K), device=dev, dtype=bf16)
    b = torch.randn((K, N) if b_mn else (N, K), device=dev, dtype=bf16)
    odt = f32 if epi in (L.EPI_ATOMIC_F32, L.EPI_GATE_RESID_F32) else bf16
    out = torch.zeros(M, N, device=dev, dtype=odt)
    out2 = torch.empty(M, N, device=dev, dtype=bf16) if epi in (L.EPI_GELU_DUAL, L.EPI_GATE_RESID_F32) else None
    aux = None
    if epi == L.EPI_GATE_RESID_F32:
        aux = torch.randn(M, N, device=dev)
    if epi == L.EPI_DGELU_BF16:
        aux = torch.randn(M, N, device=dev, dtype=bf16)
    for bn in (256, 512):
        ms = timeit(lambda: ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, epi=epi, out=out, out2=out2, aux=aux, block_n=bn))
        print(f"TIME {name:32s} bn={bn} {ms*1e3:9.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s", flush=True)
    A2 = a.t().contiguous() if a_mn else a
    B2 = b.t().contiguous() if b_mn else b
    ms = timeit(lambda: torch.matmul(A2, B2.t()))
    print(f"TIME {name:32s} cublas {ms*1e3:9.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s", flush=True)

print("launches:", L.launch_count())
print("FAILS:", fails)
sys.exit(1 if fails else 0)
