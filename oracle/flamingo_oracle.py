"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU (torch, fp32) restatement of the reference's algorithm for the OpenFlamingo dense hot path, written as pure
functions over a state dict that uses the reference's parameter names.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` leg may import this.

Parity status: PINNED AGAINST THE REFERENCE ITSELF.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so tests/golden/make_golden.py imports the unmodified reference modules from
/root/reference in the authoring container, runs them on seeded inputs and commits inputs+weights+outputs as
fixtures; tests/test_oracle_golden.py checks every function below against those fixtures (fp32, atol 1e-5).
The ViT is third-party code absent from /root/reference (open_clip_torch>=2.16.0, requirements.txt:6): its
published algorithm is restated here and pinned against HF transformers' CLIPVisionModel (same architecture,
independent implementation) through the same fixture mechanism.

Each function cites the reference lines it follows (paths relative to /root/reference/open_flamingo/src/).
"""
import math

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------- labels
def make_labels(input_ids, pad_token_id, media_token_id, endofchunk_token_id=None, interleaved=False):
    """Training labels exactly as the reference's training loop builds them (open_flamingo/train/train_utils.py;
    NOT under src/): :102-106 for image-text pairs, :126-149 for interleaved rows -- a literal, loop-for-loop
    restatement (pure-Python, small cases only).  Pinned by tests/golden/labels.pt, which holds the tensors the
    unmodified `train_one_epoch` handed to the model (tests/golden/make_golden_labels.py)."""
    labels = input_ids.clone()
    labels[labels == pad_token_id] = -100                                   # :104 / :127
    if interleaved:
        B, T = labels.shape
        for i in range(B):
            j = 0
            while j < T and labels[i, j] != media_token_id:                 # :129-136 nothing before the first <image>
                labels[i, j] = -100
                j += 1
            for e in torch.where(labels[i] == endofchunk_token_id)[0].tolist():   # :139 (positions taken up front)
                j = e + 1
                while j < T and labels[i, j] != media_token_id:             # :141-147 nothing between eoc and <image>
                    labels[i, j] = -100
                    j += 1
    labels[labels == media_token_id] = -100                                 # :105 / :149
    return labels


# ---------------------------------------------------------------------------------------------- helpers
def _ln(x, sd, prefix, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def _split_heads(t, heads):
    """(..., n, h*d) -> (..., h, n, d): the `b t n (h d) -> b h t n d` rearranges of helpers.py:55,190."""
    *lead, n, hd = t.shape
    return t.reshape(*lead, n, heads, hd // heads).transpose(-3, -2)


def _merge_heads(t):
    """(..., h, n, d) -> (..., n, h*d): helpers.py:64,232."""
    t = t.transpose(-3, -2)
    *lead, n, h, d = t.shape
    return t.reshape(*lead, n, h * d)


def feed_forward(x, sd, prefix):
    """helpers.py:15-22 -- Sequential(LayerNorm, Linear(no bias), GELU(erf), Linear(no bias))."""
    y = _ln(x, sd, prefix + ".0")
    y = F.gelu(y @ sd[prefix + ".1.weight"].t())
    return y @ sd[prefix + ".3.weight"].t()


# ---------------------------------------------------------------------------------------------- perceiver
def perceiver_attention(x, latents, sd, prefix, heads=8):
    """helpers.py:39-65.  x: (b, T, n1, D) media tokens, latents: (b, T, n2, D)."""
    x = _ln(x, sd, prefix + ".norm_media")
    latents = _ln(latents, sd, prefix + ".norm_latents")
    q = latents @ sd[prefix + ".to_q.weight"].t()                                    # :52
    kv = torch.cat((x, latents), dim=-2) @ sd[prefix + ".to_kv.weight"].t()          # :53-54
    k, v = kv.chunk(2, dim=-1)
    dim_head = q.shape[-1] // heads
    # the reference's pattern is "b t n (h d) -> b h t n d"; heads-before-T vs after does not change the math
    q, k, v = _split_heads(q, heads), _split_heads(k, heads), _split_heads(v, heads)
    q = q * dim_head ** -0.5                                                         # :56
    sim = q @ k.transpose(-1, -2)                                                    # :59
    sim = sim - sim.amax(dim=-1, keepdim=True).detach()                              # :60
    attn = sim.softmax(dim=-1)                                                       # :61
    out = _merge_heads(attn @ v)                                                     # :63-64
    return out @ sd[prefix + ".to_out.weight"].t()                                   # :65


def perceiver_resampler(x, sd, prefix="", depth=None, heads=8):
    """helpers.py:107-132.  x: (b, T, F, v, D) -> (b, T, n, D).  frame / media-time embeddings are applied when
    present in the state dict (they are None in Flamingo, flamingo.py:48)."""
    b, T, Fr, v, D = x.shape
    if prefix + "frame_embs" in sd:                                                  # :118-120
        x = x + sd[prefix + "frame_embs"][:Fr].view(1, 1, Fr, 1, D)
    x = x.reshape(b, T, Fr * v, D)                                                   # :121-123
    if prefix + "media_time_embs" in sd:                                             # :124-125
        x = x + sd[prefix + "media_time_embs"][:T]
    lat = sd[prefix + "latents"]
    latents = lat.view(1, 1, *lat.shape).expand(b, T, *lat.shape)                    # :128
    if depth is None:
        depth = 1 + max(int(k[len(prefix) + 7:].split(".")[0]) for k in sd if k.startswith(prefix + "layers."))
    for i in range(depth):                                                           # :129-131
        latents = perceiver_attention(x, latents, sd, f"{prefix}layers.{i}.0", heads) + latents
        latents = feed_forward(latents, sd, f"{prefix}layers.{i}.1") + latents
    return _ln(latents, sd, prefix + "norm")                                         # :132


# ---------------------------------------------------------------------------------------------- gated xattn
def text_time_of(media_locations, use_cached_media, t_txt):
    """helpers.py:199-208."""
    if use_cached_media:
        return media_locations.count_nonzero(dim=1).view(-1, 1).expand(-1, t_txt)
    return media_locations.cumsum(dim=-1)


def masked_cross_attention(x, media, sd, prefix, media_locations=None, use_cached_media=False, heads=8,
                           only_attend_immediate_media=True):
    """helpers.py:160-233.  x: (B, T_txt, D); media: (B, T_img, n, Dv); media_locations: (B, T_txt) bool."""
    if not use_cached_media and media_locations is not None:
        assert media_locations.shape[1] == x.shape[1]                                # :175-178
    T_txt = x.shape[1]
    _, T_img, n = media.shape[:3]
    x = _ln(x, sd, prefix + ".norm")                                                 # :184
    q = x @ sd[prefix + ".to_q.weight"].t()                                          # :186
    media = media.reshape(media.shape[0], T_img * n, media.shape[-1])                # :187
    k, v = (media @ sd[prefix + ".to_kv.weight"].t()).chunk(2, dim=-1)               # :189
    dim_head = q.shape[-1] // heads
    q, k, v = _split_heads(q, heads), _split_heads(k, heads), _split_heads(v, heads)
    q = q * dim_head ** -0.5                                                         # :192
    sim = q @ k.transpose(-1, -2)                                                    # :194
    text_time = None
    if media_locations is not None:                                                  # :196-218
        media_time = torch.arange(T_img, device=x.device) + 1
        text_time = text_time_of(media_locations, use_cached_media, T_txt)
        key_time = media_time.repeat_interleave(n)                                   # "j -> (j n)"
        tt = text_time[:, None, :, None]
        allowed = (tt == key_time) if only_attend_immediate_media else (tt >= key_time)
        sim = sim.masked_fill(~allowed, -torch.finfo(sim.dtype).max)
    sim = sim - sim.amax(dim=-1, keepdim=True).detach()                              # :220
    attn = sim.softmax(dim=-1)                                                       # :221
    if media_locations is not None and only_attend_immediate_media:                  # :223-229
        attn = attn.masked_fill((text_time == 0)[:, None, :, None], 0.0)
    out = _merge_heads(attn @ v)                                                     # :231-232
    return out @ sd[prefix + ".to_out.weight"].t()                                   # :233


def gated_cross_attention_block(x, media, sd, prefix, media_locations=None, use_cached_media=False, heads=8,
                                only_attend_immediate_media=True):
    """helpers.py:260-279."""
    p = prefix + "." if prefix else ""
    a = masked_cross_attention(x, media, sd, p + "attn", media_locations, use_cached_media, heads,
                               only_attend_immediate_media)
    x = a * sd[p + "attn_gate"].tanh() + x                                           # :267-276
    x = feed_forward(x, sd, p + "ff") * sd[p + "ff_gate"].tanh() + x                 # :277
    return x


# ---------------------------------------------------------------------------------------------- ViT (third party)
def vit_forward(images, sd, prefix="", heads=16, patch=14, quick_gelu=True):
    """open_clip VisionTransformer.forward with output_tokens=True (open_clip_torch 2.16-2.20; called at
    flamingo.py:195).  Returns (pooled, tokens); tokens exclude the class token and are NOT ln_post-normalised."""
    w = sd[prefix + "conv1.weight"]
    x = F.conv2d(images, w, stride=patch)                                            # (N, width, g, g)
    x = x.flatten(2).transpose(1, 2)                                                 # (N, g*g, width)
    cls = sd[prefix + "class_embedding"].view(1, 1, -1).expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1) + sd[prefix + "positional_embedding"]
    x = _ln(x, sd, prefix + "ln_pre")
    width = x.shape[-1]
    i = 0
    while f"{prefix}transformer.resblocks.{i}.ln_1.weight" in sd:
        p = f"{prefix}transformer.resblocks.{i}"
        h = _ln(x, sd, p + ".ln_1")
        qkv = h @ sd[p + ".attn.in_proj_weight"].t() + sd[p + ".attn.in_proj_bias"]
        q, k, v = (_split_heads(t, heads) for t in qkv.chunk(3, dim=-1))
        att = (q @ k.transpose(-1, -2)) / math.sqrt(width // heads)
        o = _merge_heads(att.softmax(dim=-1) @ v)
        x = x + o @ sd[p + ".attn.out_proj.weight"].t() + sd[p + ".attn.out_proj.bias"]
        h = _ln(x, sd, p + ".ln_2")
        f = h @ sd[p + ".mlp.c_fc.weight"].t() + sd[p + ".mlp.c_fc.bias"]
        f = f * torch.sigmoid(1.702 * f) if quick_gelu else F.gelu(f)
        x = x + f @ sd[p + ".mlp.c_proj.weight"].t() + sd[p + ".mlp.c_proj.bias"]
        i += 1
    pooled = _ln(x[:, 0], sd, prefix + "ln_post") @ sd[prefix + "proj"]
    return pooled, x[:, 1:]


# ---------------------------------------------------------------------------------------------- whole model
class OracleFlamingo:
    """Functional restatement of Flamingo.forward (flamingo.py:60-122) + FlamingoLMMixin/FlamingoLayer
    (flamingo_lm.py:39-66,128-157) around an arbitrary HF causal LM (the frozen LM is common to both sides of
    every parity test, so it is used as-is through forward pre-hooks on its decoder blocks)."""

    def __init__(self, lang_model, decoder_blocks, state_dict, media_token_id, xattn_every=1, vit_heads=16,
                 vit_patch=14, quick_gelu=True, only_attend_immediate_media=True):
        self.lm = lang_model
        self.blocks = list(decoder_blocks)
        self.sd = state_dict
        self.media_token_id = media_token_id
        self.every = xattn_every
        self.vit_heads, self.vit_patch, self.quick_gelu = vit_heads, vit_patch, quick_gelu
        self.immediate = only_attend_immediate_media

    def encode_vision(self, vision_x):
        """flamingo.py:177-200."""
        assert vision_x.ndim == 6 and vision_x.shape[2] == 1
        b, T, Fr = vision_x.shape[:3]
        with torch.no_grad():
            tokens = vit_forward(vision_x.reshape(b * T * Fr, *vision_x.shape[3:]), self.sd, "vision_encoder.",
                                 self.vit_heads, self.vit_patch, self.quick_gelu)[1]
        tokens = tokens.reshape(b, T, Fr, tokens.shape[1], tokens.shape[2])
        return perceiver_resampler(tokens, self.sd, "perceiver.")

    def forward(self, vision_x, lang_x, attention_mask=None, labels=None, media=None, use_cached_media=False,
                media_locations=None, **lm_kwargs):
        if media is None:
            media = self.encode_vision(vision_x)
        if media_locations is None:
            media_locations = lang_x == self.media_token_id                          # flamingo.py:310
        hooks = []
        for i, blk in enumerate(self.blocks):
            if (i + 1) % self.every != 0:                                            # flamingo_lm.py:100
                continue
            prefix = f"lang_encoder.gated_cross_attn_layers.{i}"

            def pre(module, args, kwargs, prefix=prefix):
                h = args[0] if args else kwargs["hidden_states"]
                h = gated_cross_attention_block(h, media, self.sd, prefix, media_locations, use_cached_media,
                                                only_attend_immediate_media=self.immediate)
                if args:
                    return (h,) + tuple(args[1:]), kwargs
                kwargs = dict(kwargs)
                kwargs["hidden_states"] = h
                return args, kwargs

            hooks.append(blk.register_forward_pre_hook(pre, with_kwargs=True))
        try:
            return self.lm(input_ids=lang_x, attention_mask=attention_mask, labels=labels, **lm_kwargs)
        finally:
            for h in hooks:
                h.remove()
