"""Fixture loading + model reconstruction shared by the CPU (oracle) and GPU (kernel) parity tests."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)

from golden_utils import check_digest, seeded_state_dict, seeded_tensor  # noqa: E402,F401


def load(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def build_mpt(mpt_kw, seed, shapes):
    from transformers import MptConfig, MptForCausalLM
    lm = MptForCausalLM(MptConfig(**mpt_kw)).eval()
    sd = seeded_state_dict(shapes, seed)
    missing, unexpected = lm.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    return lm


def flamingo_state(fx, device="cpu"):
    """Full trainable + ViT state dict of a flamingo_every* fixture under the reference's key names."""
    sd = {}
    for k, v in seeded_state_dict(fx["vit_shapes"], fx["vit_seed"]).items():
        sd["vision_encoder." + k] = v
    sd["vision_encoder.proj"] = torch.eye(fx["vit_cfg"]["width"])
    for k, v in seeded_state_dict(fx["perceiver_shapes"], fx["perceiver_seed"]).items():
        sd["perceiver." + k] = v
    for k, v in seeded_state_dict(fx["xattn_shapes"], fx["xattn_seed"]).items():
        sd["lang_encoder.gated_cross_attn_layers." + k] = v
    return {k: v.to(device) for k, v in sd.items()}


def greedy_generate(step_logits_fn, prompt, max_new_tokens, eos_id, pad_id):
    """HF-style greedy decoding (finished rows are padded) over a `full sequence -> logits` function."""
    cur = prompt.clone()
    finished = torch.zeros(cur.shape[0], dtype=torch.bool, device=cur.device)
    for _ in range(max_new_tokens):
        nxt = step_logits_fn(cur)[:, -1].argmax(-1)
        nxt = torch.where(finished, torch.full_like(nxt, pad_id), nxt)
        cur = torch.cat([cur, nxt[:, None]], dim=1)
        finished |= nxt == eos_id
        if finished.all():
            break
    return cur
