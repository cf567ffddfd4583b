"""All tcgen05 GEMM launches of one OF-3B training step (shape list written by `bench.py --gemm-shapes`), replayed
back-to-back from one CUDA graph: total time and TFLOP/s per epilogue family.  Used for same-box A/B of GEMM build
variants:  OFK_LIB_VARIANT=<name> python tools/bench_gemm_step.py profiles/r02_gemm_by_shape.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from open_flamingo_b200 import _lib as L
from open_flamingo_b200 import ops

bf16, f32 = torch.bfloat16, torch.float32


def main():
    shapes = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "profiles/r02_gemm_by_shape.json"))
    reps = int(os.environ.get("REPS", "5"))
    dev = "cuda"
    torch.manual_seed(0)
    pool = {}

    def buf(key, shape, dtype):
        k = (key, tuple(shape), dtype)
        if k not in pool:
            pool[k] = (torch.randn(shape, device=dev) * 0.05).to(dtype) if dtype != f32 else torch.randn(shape, device=dev)
        return pool[k]

    calls = []
    for s in shapes:
        M, N, K, epi, a_mn, b_mn, splits = s["M"], s["N"], s["K"], s["epi"], bool(s["a_mn"]), bool(s["b_mn"]), s["splits"]
        a = buf("a", (K, M) if a_mn else (M, K), bf16)
        b = buf("b", (K, N) if b_mn else (N, K), bf16)
        kw = dict(a_mn=a_mn, b_mn=b_mn, epi=epi, splits=splits)
        if epi in (L.EPI_STORE_F32, L.EPI_ATOMIC_F32, L.EPI_GATE_RESID_F32, L.EPI_BIAS_RESID_F32):
            kw["out"] = buf("of", (M, N), f32)
        else:
            kw["out"] = buf("ob", (M, N), bf16)
        if epi in (L.EPI_BIAS_BF16, L.EPI_BIAS_QGELU_BF16, L.EPI_BIAS_GELU_BF16, L.EPI_BIAS_RESID_F32):
            kw["bias"] = buf("bias", (N,), f32)
        if epi == L.EPI_GELU_DUAL:
            kw["out2"] = buf("o2", (M, N), bf16)
        if epi in (L.EPI_GATE_RESID_F32, L.EPI_BIAS_RESID_F32):
            kw["aux"] = buf("auxf", (M, N), f32)
        if epi == L.EPI_GATE_RESID_F32:
            kw["gate"] = buf("gate", (1,), f32)
            kw["out2"] = buf("o2", (M, N), bf16)
        if epi == L.EPI_DGELU_BF16:
            kw["aux"] = buf("auxb", (M, N), bf16)
        calls.append((a, b, kw, int(round(s["launches_per_step"])), 2.0 * M * N * K, epi))

    def run_all():
        for a, b, kw, n, _f, _e in calls:
            for _ in range(n):
                ops.gemm(a, b, **kw)

    run_all()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run_all()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    # per-family graphs (so families can be timed separately) + one graph of everything
    fam = {}
    for c in calls:
        fam.setdefault(c[5], []).append(c)
    out = {}
    total_flop = sum(c[3] * c[4] for c in calls)
    g_all = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_all):
        run_all()
    def timed(g):
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    ms_all = timed(g_all)
    for epi, cs in sorted(fam.items()):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for a, b, kw, n, _f, _e in cs:
                for _ in range(n):
                    ops.gemm(a, b, **kw)
        ms = timed(g)
        fl = sum(c[3] * c[4] for c in cs)
        out[f"epi{epi}"] = (round(ms, 3), round(fl / ms / 1e9))
    print(json.dumps({"variant": os.environ.get("OFK_LIB_VARIANT", "default"), "launches": sum(c[3] for c in calls),
                      "ms_all": round(ms_all, 3), "TFLOPs_all": round(total_flop / ms_all / 1e9), "by_epi (ms, TF/s)": out}))


if __name__ == "__main__":
    main()
