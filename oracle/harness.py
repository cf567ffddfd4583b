"""ORACLE HARNESS -- TEST INFRASTRUCTURE ONLY (same rules as flamingo_oracle.py: imported by tests/, smoke() and the
baseline legs of bench.py, never by the product package).

oracle_from_model(model): the oracle (functional restatement of the reference, flamingo_oracle.OracleFlamingo) built
around a PRIVATE copy of a product model's frozen LM and of every hot-path parameter, so the two can be evaluated on
identical weights: used by the full-size parity tests (fp32 and torch.autocast(bf16) -- the reference's own training
numerics, train_utils.py:34-43) and by bench.py's `gpu_eager_reference` leg (the reference's eager-PyTorch path timed
on the same GPU)."""
import copy

import torch

from . import flamingo_oracle as O

PREFIXES = ("vision_encoder.", "perceiver.", "lang_encoder.gated_cross_attn_layers.")


def oracle_from_model(model, every, vit_heads=16, vit_patch=14):
    lm = copy.deepcopy(model.lang_encoder)
    blocks = [layer.decoder_layer for layer in lm._get_decoder_layers()]
    lm._set_decoder_layers(torch.nn.ModuleList(blocks))
    lm.__class__ = lm.__class__.__mro__[2]   # drop the FlamingoLMMixin: plain HF LM + the oracle's pre-hooks
    lm.gated_cross_attn_layers = None
    lm.old_decoder_blocks = None
    lm.__dict__.pop("loss_function", None)   # the product installs its fused loss on the instance: the oracle uses HF's
    for p in lm.parameters():
        p.requires_grad_(False)
    sd = {}
    for k, v in model.state_dict().items():
        if k.startswith(PREFIXES):
            sd[k] = v.detach().clone().float()
    # every gated block is reachable under two names (flamingo_lm.py:94-126); keep the reference's checkpoint names
    trainable = [k for k, p in model.named_parameters(remove_duplicate=False) if p.requires_grad and k in sd]
    for k in trainable:
        sd[k].requires_grad_(True)
    orc = O.OracleFlamingo(lm, blocks, sd, model.media_token_id, xattn_every=every, vit_heads=vit_heads, vit_patch=vit_patch)
    return orc, sd, trainable


def oracle_train_step(orc, sd, trainable, batch, amp_dtype=None):
    """One forward + backward of the oracle (the reference's step without DDP / optimizer): returns the HF output."""
    for k in trainable:
        sd[k].grad = None
    if amp_dtype is not None:
        with torch.autocast(batch["lang_x"].device.type, dtype=amp_dtype):
            out = orc.forward(batch["vision_x"], batch["lang_x"], attention_mask=batch["attention_mask"], labels=batch["labels"])
    else:
        out = orc.forward(batch["vision_x"], batch["lang_x"], attention_mask=batch["attention_mask"], labels=batch["labels"])
    out.loss.float().backward()
    return out
