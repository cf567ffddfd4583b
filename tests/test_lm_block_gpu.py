"""GPU: the frozen-LM fast path (lm_blocks.FastMptBlock on libofk kernels) against HF's own MptBlock in fp32
(the LM is third-party code on both sides; HF's eager module IS the reference for it).  Tolerances as for the
other bf16 paths: outputs 2e-2 of max, input gradients 4e-2 of max, cosine >= 0.999."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def cmp(got, ref, tol, what):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-9
    cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
    assert err <= tol * scale and cos >= 0.999, f"{what}: err {err:.3e} / max {scale:.3e}, cos {cos:.5f}"


def make_block(d_model, n_heads, seed):
    from transformers import MptConfig
    from transformers.models.mpt.modeling_mpt import MptBlock, build_mpt_alibi_tensor
    torch.manual_seed(seed)
    cfg = MptConfig(d_model=d_model, n_heads=n_heads, n_layers=1, vocab_size=32, max_seq_len=512, expansion_ratio=4)
    blk = MptBlock(cfg, 0).cuda().eval().requires_grad_(False)
    with torch.no_grad():
        for n, p in blk.named_parameters():
            if p.dim() == 1:
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
            else:
                p.copy_(torch.randn_like(p) * p.shape[1] ** -0.5)
    alibi = build_mpt_alibi_tensor(n_heads, cfg.max_seq_len, device="cuda")
    return blk, alibi


def masks(B, T, kind):
    causal = torch.triu(torch.ones(T, T, dtype=torch.bool, device="cuda"), 1)  # True = masked
    m = causal.view(1, 1, T, T).expand(B, 1, T, T).clone()
    att = torch.ones(B, T, dtype=torch.long, device="cuda")
    if kind == "right_pad":
        att[0, T - 7:] = 0
    elif kind == "left_pad":
        att[1, :5] = 0
    m = m | (att == 0).view(B, 1, 1, T)
    return m, att


@pytest.mark.parametrize("d_model,n_heads", [(256, 2), (256, 4)])   # head_dim 128 (MPT-1B/7B) and 64
@pytest.mark.parametrize("kind", ["causal", "right_pad", "left_pad"])
@pytest.mark.parametrize("T", [96, 130])
def test_fast_mpt_block_matches_hf(d_model, n_heads, kind, T):
    from open_flamingo_b200 import lm_blocks
    blk, alibi = make_block(d_model, n_heads, 3)
    fast = lm_blocks.accelerate(blk)
    assert fast is not None
    B = 2
    torch.manual_seed(4)
    x = torch.randn(B, T, d_model, device="cuda")
    mask, att = masks(B, T, kind)
    flag = att.all().to(torch.int32).reshape(1)
    xr = x.clone().requires_grad_(True)
    ref, _ = blk(xr, position_bias=alibi, attention_mask=mask)
    xg = x.clone().requires_grad_(True)
    res = fast(xg, position_bias=alibi, attention_mask=mask, pure_causal_flag=flag)
    assert res is not None, "fast path declined"
    got = res[0]
    w = torch.randn_like(ref)
    valid = (att == 1).view(B, T, 1).float()     # gradients only flow from real (non-padded) positions
    (ref * w * valid).sum().backward()
    (got * w * valid).sum().backward()
    cmp(got, ref, 2e-2, f"block out [{kind}]")   # includes padded query rows: uniform attention, as in HF
    cmp(xg.grad, xr.grad, 4e-2, f"block dx [{kind}]")


def test_fast_path_declines_when_not_applicable():
    from open_flamingo_b200 import lm_blocks
    blk, alibi = make_block(256, 2, 5)
    fast = lm_blocks.accelerate(blk)
    x = torch.randn(1, 8, 256, device="cuda")
    mask, _ = masks(1, 8, "causal")
    assert fast(x, position_bias=alibi, attention_mask=mask, output_attentions=True) is None
    blk.ffn.up_proj.weight.requires_grad_(True)      # not frozen -> needs wgrad -> PyTorch path
    assert fast(x, position_bias=alibi, attention_mask=mask) is None
    assert lm_blocks.accelerate(torch.nn.Linear(4, 4)) is None


def test_full_model_same_logits_with_and_without_fast_lm():
    """Flamingo.forward with the LM fast path on/off (off = HF eager blocks, the reference behaviour)."""
    from helpers_golden import load, seeded_tensor
    from test_blocks_gpu import build_product_model
    from open_flamingo_b200 import lm_blocks
    fx = load("flamingo_every1")
    model = build_product_model(fx, 1)
    vision_x = seeded_tensor("flamingo/vision_x", fx["vision_x_shape"], 33).cuda()
    lang_x = fx["lang_x"].cuda()
    outs = {}
    for on in (True, False):
        lm_blocks.ENABLED = on
        try:
            outs[on] = model(vision_x=vision_x, lang_x=lang_x, attention_mask=torch.ones_like(lang_x)).logits
        finally:
            lm_blocks.ENABLED = True
    cmp(outs[True], outs[False], 2e-2, "fast-LM logits vs eager-LM logits")
    cmp(outs[True], fx["logits"], 2e-2, "fast-LM logits vs golden")
