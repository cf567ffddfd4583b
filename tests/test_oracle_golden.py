"""CPU: the oracle (oracle/flamingo_oracle.py) against the golden fixtures produced by the REAL reference
(tests/golden/make_golden.py).  fp32 on both sides: tolerance 2e-5 relative to the tensor's max (summation-order
noise only)."""
import pytest
import torch

from helpers_golden import (build_mpt, check_digest, flamingo_state, greedy_generate, load, seeded_state_dict,
                            seeded_tensor)
from oracle import flamingo_oracle as O

TOL = 2e-5


def rel_err(a, b):
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


def test_fixture_torch_version_noted():
    fx = load("perceiver")
    assert "torch" in fx["meta"]


def test_perceiver_resampler_fwd_bwd():
    fx = load("perceiver")
    sd = {k: v.requires_grad_(True) for k, v in seeded_state_dict(fx["shapes"], fx["seed"]).items()}
    x = seeded_tensor("perceiver/x", fx["x_shape"], 1)
    y = O.perceiver_resampler(x, sd)
    assert rel_err(y.detach(), fx["y"]) < TOL
    w = seeded_tensor("perceiver/w", y.shape, 1)
    (y * w).sum().backward()
    for k, d in fx["grads"].items():
        check_digest(k, sd[k].grad, d, 1e-4)


def test_perceiver_frame_and_time_embeddings():
    fx = load("perceiver_embs")
    sd = seeded_state_dict(fx["shapes"], fx["seed"])
    x = seeded_tensor("perceiver_embs/x", fx["x_shape"], 2)
    assert rel_err(O.perceiver_resampler(x, sd), fx["y"]) < TOL


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_gated_xattn_block_cases(idx):
    c = load("xattn")["cases"][idx]
    sd = {k: v.requires_grad_(True) for k, v in seeded_state_dict(c["shapes"], c["seed"]).items()}
    x = seeded_tensor(f"xattn/{c['name']}/x", c["x_shape"], 3).requires_grad_(True)
    media = seeded_tensor(f"xattn/{c['name']}/media", c["media_shape"], 3).requires_grad_(True)
    y = O.gated_cross_attention_block(x, media, sd, "", c["loc"], c["cached"],
                                      only_attend_immediate_media=c["immediate"])
    assert rel_err(y.detach(), c["y"]) < TOL, c["name"]
    w = seeded_tensor(f"xattn/{c['name']}/w", y.shape, 3)
    (y * w).sum().backward()
    assert rel_err(x.grad, c["dx"]) < 1e-4
    check_digest("dmedia", media.grad, c["dmedia"], 1e-4)
    for k, d in c["grads"].items():
        check_digest(k, sd[k].grad, d, 1e-4)
    if c["name"] == "eq":
        # text before the first <image> gets exactly zero cross-attention contribution (helpers.py:223-229)
        tt = c["loc"].cumsum(-1)
        a = O.masked_cross_attention(x.detach(), media.detach(), {k: v.detach() for k, v in sd.items()}, "attn",
                                     c["loc"])
        assert (tt == 0).any() and a[tt == 0].abs().max().item() == 0.0


def test_vit_tokens_match_independent_clip_implementation():
    fx = load("vit")
    sd = seeded_state_dict(fx["shapes"], fx["seed"])
    sd["proj"] = torch.eye(fx["cfg"]["width"])
    imgs = seeded_tensor("vit/images", fx["images_shape"], 20)
    _, tokens = O.vit_forward(imgs, sd, heads=fx["cfg"]["heads"], patch=fx["cfg"]["patch_size"], quick_gelu=True)
    assert rel_err(tokens, fx["tokens"]) < 5e-5


@pytest.mark.parametrize("every", [1, 2])
def test_full_flamingo_logits_loss_grads_generate(every):
    fx = load(f"flamingo_every{every}")
    lm = build_mpt(fx["mpt"], fx["lm_seed"], fx["lm_shapes"])
    sd = flamingo_state(fx)
    for k, v in sd.items():
        if not k.startswith("vision_encoder."):
            v.requires_grad_(True)
    orc = O.OracleFlamingo(lm, lm.transformer.blocks, sd, fx["media_id"], xattn_every=every,
                           vit_heads=fx["vit_cfg"]["heads"], vit_patch=fx["vit_cfg"]["patch_size"])
    vision_x = seeded_tensor("flamingo/vision_x", fx["vision_x_shape"], 33)
    lang_x, labels = fx["lang_x"], fx["labels"]
    out = orc.forward(vision_x, lang_x, attention_mask=torch.ones_like(lang_x), labels=labels)
    assert rel_err(out.logits.detach(), fx["logits"]) < 5e-5
    assert abs(out.loss.item() - fx["loss"].item()) < 1e-5
    lm.zero_grad()
    out.loss.backward()
    for k, d in fx["grads"].items():
        assert sd[k].grad is not None, k
        check_digest(k, sd[k].grad, d, 2e-4)
    with torch.no_grad():
        media = orc.encode_vision(vision_x)
        gen = greedy_generate(lambda ids: orc.forward(None, ids, attention_mask=torch.ones_like(ids), media=media).logits,
                              lang_x[:, :12], 6, fx["eoc_id"], 0)
        assert torch.equal(gen, fx["generated"][:, :gen.shape[1]])
        cached = orc.forward(None, lang_x[:, 12:15], media=media, use_cached_media=True,
                             media_locations=lang_x[:, :12] == fx["media_id"]).logits
        assert rel_err(cached, fx["cached_logits"]) < 5e-5
