"""Thin, shape-checked Python wrappers over the libofk.so C ABI (one function per entry point).

These do no math of their own.  Argument-shape violations raise ValueError/AssertionError *before* the
call (mirroring the reference's conventions, e.g. helpers.py:175-178); nonzero return codes from the
library become RuntimeError.
"""
import torch

from . import _lib as L

bf16 = torch.bfloat16
f32 = torch.float32


def _rowmajor_2d(t, name):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name} must be 2-D with unit inner stride, got shape {tuple(t.shape)} stride {t.stride()}")


def gemm(a, b, *, a_mn=False, b_mn=False, epi=L.EPI_STORE_BF16, out=None, out2=None, aux=None, bias=None,
         gate=None, splits=1, block_n=0, M=None, N=None, K=None):
    """out[m,n] = epi(sum_k A(m,k) B(n,k)).

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True); b: [N,K] (b_mn=False) or [K,N] (b_mn=True); both bf16.
    """
    L.require_cuda(a, b)
    _rowmajor_2d(a, "a")
    _rowmajor_2d(b, "b")
    if a.dtype != bf16 or b.dtype != bf16:
        raise ValueError("gemm operands must be bfloat16")
    m, ka = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    n, kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if ka != kb:
        raise ValueError(f"gemm reduction dims differ: {ka} vs {kb}")
    M = m if M is None else M
    N = n if N is None else N
    K = ka if K is None else K
    out_dtype = f32 if epi in (L.EPI_STORE_F32, L.EPI_ATOMIC_F32, L.EPI_GATE_RESID_F32,
                               L.EPI_BIAS_RESID_F32) else bf16
    if out is None:
        if epi == L.EPI_ATOMIC_F32:
            raise ValueError("atomic epilogue accumulates into an existing `out`")
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    _rowmajor_2d(out, "out")
    if out.dtype != out_dtype or out.shape[0] < M or out.shape[1] < N:
        raise ValueError(f"bad out tensor {tuple(out.shape)} {out.dtype} for ({M},{N}) {out_dtype}")
    if out2 is not None:
        _rowmajor_2d(out2, "out2")
    if aux is not None:
        _rowmajor_2d(aux, "aux")
    L.check(L.lib().ofk_gemm_bf16(
        epi, int(a_mn), int(b_mn), a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), M, N, K, splits, block_n,
        out.data_ptr(), out.stride(0), L.ptr(out2), 0 if out2 is None else out2.stride(0),
        L.ptr(aux), 0 if aux is None else aux.stride(0), L.ptr(bias), L.ptr(gate), L.stream_ptr()))
    return out
