"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference
(/root/reference/open_flamingo/src/*.py) in the authoring container (CPU, fp32).

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so these fixtures are what pins
the oracle (oracle/flamingo_oracle.py) and, through it, the CUDA kernels.  The fixtures travel to the GPU box;
/root/reference does not, so nothing else may import it.

Import shims (the only two the reference needs here):
  * einops_exts.rearrange_many  -- not installed; one-liner over einops.rearrange
  * open_clip                   -- not installed; only imported by factory.py, which is not exercised
Third-party pieces:
  * frozen LM   = transformers.MptForCausalLM (random init, tiny config)
  * ViT         = transformers.CLIPVisionModel (random init, tiny config, quick_gelu) as an INDEPENDENT
                  implementation of the open_clip ViT architecture; weights are exported under open_clip names.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    from einops import rearrange
    ee = types.ModuleType("einops_exts")
    ee.rearrange_many = lambda tensors, pattern, **kw: [rearrange(t, pattern, **kw) for t in tensors]
    sys.modules["einops_exts"] = ee
    sys.modules.setdefault("open_clip", types.ModuleType("open_clip"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from open_flamingo.src import helpers, flamingo, flamingo_lm, utils  # noqa: F401
    return helpers, flamingo, flamingo_lm, utils


sys.path.insert(0, HERE)
from golden_utils import digest, load_seeded, seeded_tensor, shapes_of  # noqa: E402


def hf_clip_to_openclip(hf_sd, width):
    """HF CLIPVisionModel state dict -> open_clip `visual.*` names."""
    out = {}
    g = lambda k: hf_sd["vision_model." + k]
    out["conv1.weight"] = g("embeddings.patch_embedding.weight")
    out["class_embedding"] = g("embeddings.class_embedding")
    out["positional_embedding"] = g("embeddings.position_embedding.weight")
    out["ln_pre.weight"], out["ln_pre.bias"] = g("pre_layrnorm.weight"), g("pre_layrnorm.bias")
    out["ln_post.weight"], out["ln_post.bias"] = g("post_layernorm.weight"), g("post_layernorm.bias")
    i = 0
    while f"vision_model.encoder.layers.{i}.layer_norm1.weight" in hf_sd:
        s, d = f"encoder.layers.{i}.", f"transformer.resblocks.{i}."
        out[d + "ln_1.weight"], out[d + "ln_1.bias"] = g(s + "layer_norm1.weight"), g(s + "layer_norm1.bias")
        out[d + "ln_2.weight"], out[d + "ln_2.bias"] = g(s + "layer_norm2.weight"), g(s + "layer_norm2.bias")
        out[d + "attn.in_proj_weight"] = torch.cat([g(s + f"self_attn.{n}_proj.weight") for n in "qkv"], 0)
        out[d + "attn.in_proj_bias"] = torch.cat([g(s + f"self_attn.{n}_proj.bias") for n in "qkv"], 0)
        out[d + "attn.out_proj.weight"], out[d + "attn.out_proj.bias"] = g(s + "self_attn.out_proj.weight"), g(s + "self_attn.out_proj.bias")
        out[d + "mlp.c_fc.weight"], out[d + "mlp.c_fc.bias"] = g(s + "mlp.fc1.weight"), g(s + "mlp.fc1.bias")
        out[d + "mlp.c_proj.weight"], out[d + "mlp.c_proj.bias"] = g(s + "mlp.fc2.weight"), g(s + "mlp.fc2.bias")
        i += 1
    out["proj"] = torch.eye(width)  # unused by Flamingo (tokens path); identity keeps pooled well-defined
    return {k: v.detach().clone() for k, v in out.items()}


def load_openclip_into_hf(vit, oc):
    """Inverse of hf_clip_to_openclip: write open_clip-named weights into the HF CLIPVisionModel."""
    hf = {}
    p = "vision_model."
    hf[p + "embeddings.patch_embedding.weight"] = oc["conv1.weight"]
    hf[p + "embeddings.class_embedding"] = oc["class_embedding"]
    hf[p + "embeddings.position_embedding.weight"] = oc["positional_embedding"]
    hf[p + "pre_layrnorm.weight"], hf[p + "pre_layrnorm.bias"] = oc["ln_pre.weight"], oc["ln_pre.bias"]
    hf[p + "post_layernorm.weight"], hf[p + "post_layernorm.bias"] = oc["ln_post.weight"], oc["ln_post.bias"]
    i = 0
    while f"transformer.resblocks.{i}.ln_1.weight" in oc:
        d, s_ = f"{p}encoder.layers.{i}.", f"transformer.resblocks.{i}."
        hf[d + "layer_norm1.weight"], hf[d + "layer_norm1.bias"] = oc[s_ + "ln_1.weight"], oc[s_ + "ln_1.bias"]
        hf[d + "layer_norm2.weight"], hf[d + "layer_norm2.bias"] = oc[s_ + "ln_2.weight"], oc[s_ + "ln_2.bias"]
        wq, wk, wv = oc[s_ + "attn.in_proj_weight"].chunk(3, 0)
        bq, bk, bv = oc[s_ + "attn.in_proj_bias"].chunk(3, 0)
        for n_, w_, b_ in (("q", wq, bq), ("k", wk, bk), ("v", wv, bv)):
            hf[d + f"self_attn.{n_}_proj.weight"], hf[d + f"self_attn.{n_}_proj.bias"] = w_, b_
        hf[d + "self_attn.out_proj.weight"], hf[d + "self_attn.out_proj.bias"] = oc[s_ + "attn.out_proj.weight"], oc[s_ + "attn.out_proj.bias"]
        hf[d + "mlp.fc1.weight"], hf[d + "mlp.fc1.bias"] = oc[s_ + "mlp.c_fc.weight"], oc[s_ + "mlp.c_fc.bias"]
        hf[d + "mlp.fc2.weight"], hf[d + "mlp.fc2.bias"] = oc[s_ + "mlp.c_proj.weight"], oc[s_ + "mlp.c_proj.bias"]
        i += 1
    missing, unexpected = vit.load_state_dict(hf, strict=False)
    assert not unexpected, unexpected


def media_locations_case(B, T_txt, T_img, gen, first_at=None):
    loc = torch.zeros(B, T_txt, dtype=torch.bool)
    for b in range(B):
        pos = sorted(torch.randperm(T_txt, generator=gen)[:T_img].tolist())
        if first_at is not None and b == 0:
            pos = [max(p, first_at) for p in pos]
            pos = sorted(set(pos))
            while len(pos) < T_img:
                pos.append(pos[-1] + 1)
        loc[b, pos] = True
    return loc


def main():
    torch.set_grad_enabled(True)
    helpers, flamingo, flamingo_lm, utils = import_reference()
    gen = torch.Generator().manual_seed(1234)
    out = {}
    meta = dict(torch=torch.__version__)

    def grads_of(module, skip_prefix=None):
        return {k: digest(k, p.grad) for k, p in module.named_parameters()
                if p.grad is not None and not (skip_prefix and k.startswith(skip_prefix))}

    # ---------------------------------------------------------------- PerceiverResampler (fwd + grads)
    dim = 128
    pr = helpers.PerceiverResampler(dim=dim, depth=2)
    load_seeded(pr, 1)
    x = seeded_tensor("perceiver/x", (2, 2, 1, 20, dim), 1)
    y = pr(x)
    w = seeded_tensor("perceiver/w", y.shape, 1)
    (y * w).sum().backward()
    out["perceiver"] = dict(meta=meta, dim=dim, depth=2, seed=1, shapes=shapes_of(pr), x_shape=tuple(x.shape),
                            y=y.detach(), grads=grads_of(pr))

    # with frame / media-time embeddings (constructor options Flamingo does not use, helpers.py:77-93)
    pr2 = helpers.PerceiverResampler(dim=dim, depth=1, max_num_media=3, max_num_frames=2)
    load_seeded(pr2, 2)
    x2 = seeded_tensor("perceiver_embs/x", (1, 2, 2, 9, dim), 2)
    out["perceiver_embs"] = dict(meta=meta, dim=dim, depth=1, seed=2, shapes=shapes_of(pr2), x_shape=tuple(x2.shape),
                                 y=pr2(x2).detach())

    # ---------------------------------------------------------------- GatedCrossAttentionBlock cases
    D, Dv, B, T_txt, T_img, n = 128, 128, 3, 24, 2, 64
    cases = []
    for ci, (name, immediate, cached, first_at, with_loc) in enumerate([
        ("eq", True, False, 5, True),          # text before the first image -> zero rows (helpers.py:223-229)
        ("ge", False, False, 5, True),         # attend to all previous media; text_time==0 rows are uniform
        ("cached", True, True, None, True),    # use_cached_media: everything attends to the last image
        ("nomask", True, True, None, False),   # media_locations=None (legal only with use_cached_media, helpers.py:196)
    ]):
        blk = helpers.GatedCrossAttentionBlock(dim=D, dim_visual=Dv, only_attend_immediate_media=immediate)
        load_seeded(blk, 10 + ci)
        xx = seeded_tensor(f"xattn/{name}/x", (B, T_txt if not cached else 8, D), 3).requires_grad_(True)
        media = seeded_tensor(f"xattn/{name}/media", (B, T_img, n, Dv), 3).requires_grad_(True)
        loc = media_locations_case(B, T_txt, T_img, gen, first_at) if with_loc else None
        yy = blk(xx, media, media_locations=loc, use_cached_media=cached)
        ww = seeded_tensor(f"xattn/{name}/w", yy.shape, 3)
        (yy * ww).sum().backward()
        cases.append(dict(name=name, immediate=immediate, cached=cached, seed=10 + ci, shapes=shapes_of(blk),
                          x_shape=tuple(xx.shape), media_shape=tuple(media.shape), loc=loc, y=yy.detach(),
                          dx=xx.grad.clone(), dmedia=digest("dmedia", media.grad), grads=grads_of(blk)))
    out["xattn"] = dict(meta=meta, D=D, Dv=Dv, cases=cases)

    # ---------------------------------------------------------------- ViT stand-in (HF CLIPVisionModel, quick_gelu)
    from transformers import CLIPVisionConfig, CLIPVisionModel
    vcfg = CLIPVisionConfig(hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                            image_size=56, patch_size=14, hidden_act="quick_gelu", projection_dim=128)
    vit = CLIPVisionModel(vcfg).eval()
    vit_cfg = dict(image_size=56, patch_size=14, width=128, layers=2, heads=2, output_dim=128)
    # seed the weights under their open_clip names, then push them into the HF module
    oc_shapes = shapes_of(hf_clip_to_openclip(vit.state_dict(), 128))
    from golden_utils import seeded_state_dict
    oc_sd = seeded_state_dict(oc_shapes, 20)
    oc_sd["proj"] = torch.eye(128)
    load_openclip_into_hf(vit, oc_sd)
    imgs = seeded_tensor("vit/images", (3, 3, 56, 56), 20)
    with torch.no_grad():
        hid = vit(pixel_values=imgs).last_hidden_state  # encoder output, before post_layernorm
    out["vit"] = dict(meta=meta, cfg=vit_cfg, seed=20, shapes=oc_shapes, images_shape=tuple(imgs.shape),
                      tokens=hid[:, 1:].clone())

    # ---------------------------------------------------------------- full Flamingo (reference classes, tiny MPT)
    from transformers import MptConfig, MptForCausalLM
    vocab = 64
    mpt_kw = dict(d_model=128, n_heads=2, n_layers=4, vocab_size=vocab, max_seq_len=64, expansion_ratio=2)
    media_id, eoc_id = vocab - 2, vocab - 3

    class RefVisual(torch.nn.Module):  # the reference only needs `.visual(x)[1]` (flamingo.py:47,195)
        def __init__(self, hf):
            super().__init__()
            self.hf = hf

        def forward(self, x):
            h = self.hf(pixel_values=x).last_hidden_state
            return h[:, 0], h[:, 1:]

    class RefClip(torch.nn.Module):
        def __init__(self, visual):
            super().__init__()
            self.visual = visual

    for every in (1, 2):
        lm_i = MptForCausalLM(MptConfig(**mpt_kw)).eval()
        load_seeded(lm_i, 30)
        lm_shapes = shapes_of(lm_i)
        utils.extend_instance(lm_i, flamingo_lm.FlamingoLMMixin)
        lm_i.set_decoder_layers_attr_name("transformer.blocks")
        model = flamingo.Flamingo(RefClip(RefVisual(vit)), lm_i, eoc_id, media_id, vis_dim=128,
                                  cross_attn_every_n_layers=every)
        model.eval()
        load_seeded(model.perceiver, 31)
        load_seeded(model.lang_encoder.gated_cross_attn_layers, 32)
        model.requires_grad_(False)
        model.perceiver.requires_grad_(True)
        model.lang_encoder.gated_cross_attn_layers.requires_grad_(True)
        Bf, Tt, Ti = 2, 20, 2
        vision_x = seeded_tensor("flamingo/vision_x", (Bf, Ti, 1, 3, 56, 56), 33)
        g2 = torch.Generator().manual_seed(34)
        lang_x = torch.randint(0, vocab - 3, (Bf, Tt), generator=g2)
        lang_x[0, 3], lang_x[0, 11] = media_id, media_id       # row 0: text before the first image
        lang_x[1, 0], lang_x[1, 9] = media_id, media_id
        lang_x[0, 10], lang_x[1, 8] = eoc_id, eoc_id
        labels = lang_x.clone()
        labels[labels == media_id] = -100
        attn_mask = torch.ones_like(lang_x)
        o = model(vision_x=vision_x, lang_x=lang_x, attention_mask=attn_mask, labels=labels)
        o.loss.backward()
        trainable = {k: digest(k, p.grad) for k, p in model.named_parameters()
                     if p.requires_grad and p.grad is not None and not k.startswith("lang_encoder.transformer.blocks")}
        with torch.no_grad():
            gen_out = model.generate(vision_x=vision_x, lang_x=lang_x[:, :12], attention_mask=attn_mask[:, :12],
                                     max_new_tokens=6, do_sample=False, pad_token_id=0)
            # cached-media scoring path (eval/models/open_flamingo.py:155-254)
            model.cache_media(input_ids=lang_x[:, :12], vision_x=vision_x)
            cached_logits = model(vision_x=None, lang_x=lang_x[:, 12:15], attention_mask=None,
                                  clear_conditioned_layers=False).logits
            model.uncache_media()
        out[f"flamingo_every{every}"] = dict(
            meta=meta, mpt=mpt_kw, lm_seed=30, perceiver_seed=31, xattn_seed=32, vit_seed=20, vit_cfg=vit_cfg,
            vit_shapes=oc_shapes, lm_shapes=lm_shapes, perceiver_shapes=shapes_of(model.perceiver),
            xattn_shapes=shapes_of(model.lang_encoder.gated_cross_attn_layers),
            media_id=media_id, eoc_id=eoc_id, every=every, vision_x_shape=tuple(vision_x.shape), lang_x=lang_x,
            labels=labels, logits=o.logits.detach(), loss=o.loss.detach(), grads=trainable, generated=gen_out,
            cached_logits=cached_logits)

    for key, val in out.items():
        path = os.path.join(HERE, f"{key}.pt")
        torch.save(val, path)
        print(f"{key:20s} {os.path.getsize(path)/1024:8.1f} KiB")


if __name__ == "__main__":
    main()
