"""EXPERIMENTAL -- skipped unless OFK_EXPERIMENTAL=1.  The tcgen05 attention forward core (csrc/attention_tc.cu,
ofk_attn_fwd_tc) was written at the end of round 1 without GPU time left to validate it; this file is its bring-up
harness for the next round: same cases as the mma.sync kernel's test, checked against the fp32 torch reference AND
against the validated mma.sync kernel (outputs within bf16 rounding, LSE within 1e-3)."""
import os

import pytest
import torch

from test_kernels_gpu import CASES, close, make_tt, ref_attention

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("OFK_EXPERIMENTAL") != "1",
                                 reason="experimental tcgen05 attention core: set OFK_EXPERIMENTAL=1 to bring it up")]
bf16 = torch.bfloat16


def _call(fn, q, k, v, heads, scale, mode, tt, kpm):
    from open_flamingo_b200 import _lib as L
    B, nq, inner = q.shape
    nk = k.shape[1]
    out = torch.empty((B, nq, inner), device="cuda", dtype=bf16)
    lse = torch.empty((B, heads, nq), device="cuda", dtype=torch.float32)
    L.check(fn(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(), B, heads, nq, nk,
               q.stride(0), q.stride(1), k.stride(0), k.stride(1), v.stride(0), v.stride(1), out.stride(0), out.stride(1),
               scale, mode, L.ptr(tt), kpm, L.stream_ptr()))
    torch.cuda.synchronize()
    return out, lse


@pytest.mark.parametrize("case", CASES)
def test_attention_tc_forward(case):
    from open_flamingo_b200 import _lib as L
    B, heads, nq, nk, mode, kpm = case
    torch.manual_seed(5)
    inner = heads * 64
    q = torch.randn(B, nq, inner, device="cuda").to(bf16)
    kv = torch.randn(B, nk, 2 * inner, device="cuda").to(bf16)
    k, v = kv[..., :inner], kv[..., inner:]
    tt = None
    if mode:
        tt = make_tt(B, nq, nk // kpm, mode, 7).cuda()
        if nq == 1:
            tt[:] = nk // kpm
    scale = 64 ** -0.5
    o_ref, lse_ref = _call(L.lib().ofk_attn_fwd, q, k, v, heads, scale, mode, tt, kpm)
    o_tc, lse_tc = _call(L.lib().ofk_attn_fwd_tc, q, k, v, heads, scale, mode, tt, kpm)
    close(o_tc, ref_attention(q, k, v, heads, scale, mode, tt, kpm), 2e-2, "tc attn fwd vs fp32 reference")
    close(o_tc, o_ref, 1e-2, "tc attn fwd vs mma.sync kernel")
    assert (lse_tc - lse_ref).abs().max().item() <= 1e-3 * (1 + lse_ref.abs().max().item())
    if mode == 1 and (tt == 0).any():
        assert o_tc[(tt == 0)].abs().max().item() == 0.0
