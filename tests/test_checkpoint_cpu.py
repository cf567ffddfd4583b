"""Checkpoint format (train_utils.py:299-375, train.py:297-308): the kept key set equals what the unmodified
reference filter keeps (tests/golden/checkpoint_keys.json), and a save -> load round trip through the reference's
resume rule restores every trainable tensor."""
import json
import os

import torch

from helpers_golden import GOLDEN, flamingo_state, load
from test_host_logic_cpu import _product


def test_kept_keys_match_the_reference_filter():
    from open_flamingo_b200.checkpoint import trainable_state_dict
    cases = json.load(open(os.path.join(GOLDEN, "checkpoint_keys.json")))
    assert len(cases) == 3
    for c in cases:
        model, _, _ = _product(load(c["fixture"]), c["every"], freeze_lm_embeddings=c["freeze_lm_embeddings"])
        got = trainable_state_dict(model)
        assert sorted(got) == c["keys"], (c["every"], c["freeze_lm_embeddings"])
        assert not any(k.startswith("vision_encoder") or "old_decoder_blocks" in k or "gated_cross_attn_layers" in k for k in got)
        trainable = {id(p) for p in model.parameters() if p.requires_grad}
        kept_ids = {id(v) for v in got.values()}
        sd = model.state_dict(keep_vars=True)
        assert all(any(sd[k] is p for k in got) for p in model.parameters() if p.requires_grad), "a trainable tensor was dropped"
        assert len(trainable) <= len(kept_ids)


def test_save_load_round_trip(tmp_path):
    from open_flamingo_b200.checkpoint import load_checkpoint, save_checkpoint
    fx = load("flamingo_every2")
    src, _, _ = _product(fx, 2)
    src.load_state_dict(flamingo_state(fx), strict=False)
    path = os.path.join(tmp_path, "run", "checkpoint_3.pt")
    ckpt = save_checkpoint(path, src, trainer=None, epoch=3)
    assert set(ckpt) == {"epoch", "model_state_dict", "optimizer_state_dict", "lr_scheduler_state_dict"}   # train_utils.py:359-364
    dst, _, _ = _product(fx, 2)
    # a DDP-saved checkpoint prefixes every key with "module." (train.py:303)
    on_disk = torch.load(path, map_location="cpu", weights_only=False)
    on_disk["model_state_dict"] = {"module." + k: v for k, v in on_disk["model_state_dict"].items()}
    assert load_checkpoint(on_disk, dst) == 4
    a, b = dict(src.named_parameters()), dict(dst.named_parameters())
    for n, p in a.items():
        if p.requires_grad:
            assert torch.equal(p.detach(), b[n].detach()), n
    # the per-layer alias and the ModuleList view of a gated block are the same storage after loading
    lm = dst.lang_encoder
    blk = lm.gated_cross_attn_layers[1]
    assert blk is lm._get_decoder_layers()[1].gated_cross_attn_layer
    try:
        load_checkpoint({"epoch": 0, "model_state_dict": {"no.such.key": torch.zeros(1)}}, dst)
        raise AssertionError("unknown keys must be reported")
    except KeyError:
        pass
