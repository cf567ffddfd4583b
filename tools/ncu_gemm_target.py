"""Target for `ncu --set full -k regex:gemm2_kernel`: the six FFN GEMM shapes of one gated block at C2
(R = 8192 rows, D = 2048), each launched a few times through the C ABI."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from open_flamingo_b200 import _lib as L
from open_flamingo_b200 import ops

dev, bf16 = "cuda", torch.bfloat16
R, D = 8192, 2048
x = torch.randn(R, D, device=dev, dtype=bf16)
w1 = torch.randn(4 * D, D, device=dev, dtype=bf16) * 0.02
w2 = torch.randn(D, 4 * D, device=dev, dtype=bf16) * 0.02
z = torch.empty(R, 4 * D, device=dev, dtype=bf16)
h = torch.empty(R, 4 * D, device=dev, dtype=bf16)
resid = torch.randn(R, D, device=dev)
gate = torch.tensor([0.5], device=dev)
br = torch.empty(R, D, device=dev, dtype=bf16)
dy = torch.randn(R, D, device=dev, dtype=bf16)
dw1 = torch.zeros(4 * D, D, device=dev)
dw2 = torch.zeros(D, 4 * D, device=dev)
for it in range(3):
    ops.gemm(x, w1, epi=L.EPI_GELU_DUAL, out=z, out2=h)                                  # ffn1 fwd
    ops.gemm(h, w2, epi=L.EPI_GATE_RESID_F32, aux=resid, gate=gate, out2=br)             # ffn2 fwd
    dz = ops.gemm(dy, w2, b_mn=True, epi=L.EPI_DGELU_BF16, aux=z)                        # ffn2 dgrad (+gelu')
    ops.gemm(dy, h, a_mn=True, b_mn=True, epi=L.EPI_ATOMIC_F32, out=dw2)                 # ffn2 wgrad
    ops.gemm(dz, x, a_mn=True, b_mn=True, epi=L.EPI_ATOMIC_F32, out=dw1)                 # ffn1 wgrad
    ops.gemm(dz, w1, b_mn=True)                                                          # ffn1 dgrad
torch.cuda.synchronize()
print("done", L.launch_count())
