"""Shared by tests/golden/make_golden.py (authoring container, runs the real reference) and the tests (anywhere).

To keep the committed fixtures small, weights and inputs are NOT stored: they are regenerated from seeds by the
deterministic CPU procedure below (torch's CPU generator is bit-reproducible for a given torch version; the
fixture records the torch version that produced it).  Fixtures hold only shapes, seeds, full outputs, and for
large gradient tensors a digest (Frobenius norm + 4 seeded random projections) instead of the tensor.
"""
import zlib

import torch

DIGEST_LIMIT = 4096  # tensors up to this many elements are stored in full


def _seed_for(name, seed):
    return (zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31 - 1)


def seeded_tensor(name, shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(_seed_for(name, seed))
    return torch.randn(tuple(shape), generator=g) * scale


def seeded_param(name, shape, seed):
    """Deterministic, well-conditioned parameter values by role."""
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]
    if leaf in ("attn_gate", "ff_gate"):
        g = torch.Generator().manual_seed(_seed_for(name, seed))
        return torch.rand(shape, generator=g) * 2 - 1                    # gates init to 0 in the reference: use U(-1,1)
    if len(shape) == 1:
        if leaf == "weight" or leaf.endswith("norm") or "ln" in leaf:
            return 1.0 + 0.1 * seeded_tensor(name, shape, seed)          # LayerNorm gains
        return 0.05 * seeded_tensor(name, shape, seed)                   # biases, class embedding
    fan_in = shape[-1] if len(shape) == 2 else int(torch.tensor(shape[1:]).prod())
    if leaf in ("latents", "positional_embedding", "frame_embs", "media_time_embs"):
        return seeded_tensor(name, shape, seed, 0.5)
    return seeded_tensor(name, shape, seed, fan_in ** -0.5)


def seeded_state_dict(shapes, seed):
    return {k: seeded_param(k, s, seed) for k, s in shapes.items()}


def shapes_of(module_or_sd):
    sd = module_or_sd if isinstance(module_or_sd, dict) else module_or_sd.state_dict()
    return {k: tuple(v.shape) for k, v in sd.items() if v.dtype.is_floating_point}


def load_seeded(module, seed):
    """Overwrite every floating-point entry of module.state_dict() with its seeded value; returns the dict."""
    sd = seeded_state_dict(shapes_of(module), seed)
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    return sd


def digest(name, t):
    t = t.detach().float().cpu()
    if t.numel() <= DIGEST_LIMIT:
        return {"full": t.clone()}
    probes = torch.stack([seeded_tensor(f"{name}/probe{i}", t.shape, 99) for i in range(4)])
    return {"norm": t.norm().double(), "proj": (probes * t).flatten(1).sum(1).double(), "numel": t.numel()}


def check_digest(name, t, d, rtol, atol_scale=1.0):
    """Assert tensor `t` matches digest `d` (returns the worst relative error seen)."""
    t = t.detach().float().cpu()
    if "full" in d:
        ref = d["full"]
        err = (t - ref).abs().max().item()
        scale = ref.abs().max().item() + 1e-12
        assert err <= rtol * scale * atol_scale + 1e-7, f"{name}: max_abs_err {err:.3e} (ref max {scale:.3e})"
        return err / scale
    probes = torch.stack([seeded_tensor(f"{name}/probe{i}", t.shape, 99) for i in range(4)])
    proj = (probes * t).flatten(1).sum(1).double()
    nrm = d["norm"].item()
    # a random projection of an error e has std ~ |e|; compare against the tensor norm
    perr = (proj - d["proj"]).abs().max().item() / (nrm + 1e-12)
    nerr = abs(t.norm().item() - nrm) / (nrm + 1e-12)
    assert perr <= rtol * atol_scale and nerr <= rtol * atol_scale, f"{name}: proj err {perr:.3e}, norm err {nerr:.3e}"
    return max(perr, nerr)
