"""Host-side logic that needs no GPU: the drop-in surface (`create_model_and_transforms`, `Flamingo`,
`FlamingoLMMixin`) must build on CPU, keep the reference's parameter names / shapes / freezing rule / special
tokens / placement rule, and raise the reference's exceptions before any kernel is reached.

The expected names and shapes are the ones the UNMODIFIED reference produced when the fixtures were authored
(tests/golden/make_golden.py records `state_dict()` shapes of open_flamingo/src/helpers.py's modules)."""
import pytest
import torch

from helpers_golden import build_mpt, flamingo_state, load
from golden_utils import shapes_of


def _product(fx, every, freeze_lm_embeddings=True):
    from open_flamingo_b200 import create_model_and_transforms
    from open_flamingo_b200.src.vit import CLIPVisionStandIn, VisionTransformer
    from open_flamingo_b200.testing import SimpleTokenizer
    lm = build_mpt(fx["mpt"], fx["lm_seed"], fx["lm_shapes"])
    vit = VisionTransformer(**fx["vit_cfg"])
    tok = SimpleTokenizer(fx["mpt"]["vocab_size"] - 3)
    return create_model_and_transforms(CLIPVisionStandIn(vit), None, lm, tok, cross_attn_every_n_layers=every,
                                       freeze_lm_embeddings=freeze_lm_embeddings)


def test_perceiver_state_dict_names_and_shapes_match_reference():
    from open_flamingo_b200.src.helpers import PerceiverResampler
    for name in ("perceiver", "perceiver_embs"):
        fx = load(name)
        extra = {}
        if "frame_embs" in fx["shapes"]:
            extra["max_num_frames"] = fx["shapes"]["frame_embs"][0]
        if "media_time_embs" in fx["shapes"]:
            extra["max_num_media"] = fx["shapes"]["media_time_embs"][0]
        m = PerceiverResampler(dim=fx["dim"], depth=fx["depth"], **extra)
        assert shapes_of(m) == {k: tuple(v) for k, v in fx["shapes"].items()}, name


def test_gated_block_state_dict_names_and_shapes_match_reference():
    from open_flamingo_b200.src.helpers import GatedCrossAttentionBlock
    fx = load("xattn")
    m = GatedCrossAttentionBlock(dim=fx["D"], dim_visual=fx["Dv"])
    for c in fx["cases"]:
        assert shapes_of(m) == {k: tuple(v) for k, v in c["shapes"].items()}, c["name"]
    # the reference initialises both gates to zero (helpers.py:252,257): a fresh block is the identity
    assert float(m.attn_gate.detach()) == 0.0 and float(m.ff_gate.detach()) == 0.0


def test_vit_state_dict_names_match_open_clip_layout():
    from open_flamingo_b200.src.vit import VisionTransformer
    fx = load("vit")
    m = VisionTransformer(**fx["cfg"], output_tokens=True)
    got = shapes_of(m)
    for k, s in fx["shapes"].items():
        assert got.get(k) == tuple(s), k


@pytest.mark.parametrize("every", [1, 2])
def test_factory_wiring_freezing_and_placement(every):
    fx = load(f"flamingo_every{every}")
    model, image_processor, tok = _product(fx, every)
    # special tokens appended in the reference's order (factory.py:58-67), embeddings resized (factory.py:98)
    assert tok.encode("<image>")[-1] == fx["media_id"] and tok.encode("<|endofchunk|>")[-1] == fx["eoc_id"]
    assert model.media_token_id == fx["media_id"] and model.eoc_token_id == fx["eoc_id"]
    assert model.lang_encoder.get_input_embeddings().weight.shape[0] == len(tok)
    # checkpoint compatibility: every tensor of a reference checkpoint finds its key
    missing, unexpected = model.load_state_dict(flamingo_state(fx), strict=False)
    assert not unexpected, unexpected
    # placement rule (flamingo_lm.py:101-108): a gated block before decoder layer i iff (i + 1) % every == 0
    layers = model.lang_encoder._get_decoder_layers()
    n_layers = fx["mpt"]["n_layers"]
    assert len(layers) == n_layers
    for i, layer in enumerate(layers):
        assert (layer.gated_cross_attn_layer is not None) == ((i + 1) % every == 0), i
    assert len(model.lang_encoder.gated_cross_attn_layers) == n_layers
    # freezing rule (factory.py:101-111): everything frozen except perceiver + gated blocks (+ embeddings if asked)
    trainable = {n for n, p in model.named_parameters() if p.requires_grad}
    assert trainable and all(n.startswith(("perceiver.", "lang_encoder.gated_cross_attn_layers.",
                                           "lang_encoder.transformer.blocks.")) for n in trainable), trainable
    assert all("gated_cross_attn_layer" in n for n in trainable if n.startswith("lang_encoder.transformer.blocks."))
    assert not any(p.requires_grad for p in model.vision_encoder.parameters())
    assert not model.lang_encoder.get_input_embeddings().weight.requires_grad


def test_factory_unfreezes_embeddings_by_default_like_the_reference():
    fx = load("flamingo_every1")
    model, _, _ = _product(fx, 1, freeze_lm_embeddings=False)
    assert model.lang_encoder.get_input_embeddings().weight.requires_grad      # factory.py:108-110


def test_hot_path_parameters_are_exactly_the_trainable_resampler_and_gated_blocks():
    from open_flamingo_b200.train import hot_path_parameters
    fx = load("flamingo_every2")
    model, _, _ = _product(fx, 2)
    groups = hot_path_parameters(model)
    # groups come in backward-completion order: last gated block first, resampler last (its grads are final last)
    kinds = [(kind, idx) for kind, idx, _ in groups]
    xattn_idx = [idx for kind, idx in kinds if kind == "xattn"]
    assert xattn_idx == sorted(xattn_idx, reverse=True) and kinds[-1] == ("perceiver", -1), kinds
    named = [np for _, _, ps in groups for np in ps]
    names = [n for n, _ in named]
    assert len(names) == len(set(names))
    assert len({id(p) for _, p in named}) == len(named)          # aliases (per-layer attribute) appear once
    assert all(n.startswith(("perceiver.", "lang_encoder.gated_cross_attn_layers.")) for n in names), names
    # with the LM embeddings frozen these are ALL the trainable tensors of the model (named_parameters() reports a
    # gated block under its per-layer alias `transformer.blocks.{i}.gated_cross_attn_layer`, hence ids not names)
    want = {id(p) for p in model.parameters() if p.requires_grad}
    assert {id(p) for _, p in named} == want


def test_errors_raised_before_any_kernel():
    fx = load("flamingo_every1")
    model, _, _ = _product(fx, 1)
    lang_x = fx["lang_x"]
    with pytest.raises(AssertionError):                                       # flamingo.py:94-96
        model(vision_x=None, lang_x=lang_x)
    with pytest.raises(AssertionError):                                       # flamingo.py:189 (ndim != 6)
        model(vision_x=torch.zeros(2, 2, 3, 56, 56), lang_x=lang_x)
    with pytest.raises(AssertionError):                                       # flamingo.py:191 (F != 1)
        model(vision_x=torch.zeros(2, 2, 2, 3, 56, 56), lang_x=lang_x)
    layer = model.lang_encoder._get_decoder_layers()[0]
    with pytest.raises(ValueError):                                           # flamingo_lm.py:47-53
        layer(torch.zeros(1, 4, fx["mpt"]["d_model"]))
    with pytest.raises(RuntimeError):                                         # CPU tensor: there is no fallback
        model.perceiver(torch.zeros(1, 1, 1, 4, fx["mpt"]["d_model"]))


def test_conditioning_plumbing_on_cpu():
    """condition / is_conditioned / clear_conditioned_layers bookkeeping (flamingo_lm.py:35-45,145-167)."""
    fx = load("flamingo_every2")
    model, _, _ = _product(fx, 2)
    lm = model.lang_encoder
    assert not lm.is_conditioned()
    model._condition_media_locations(fx["lang_x"])
    assert not lm.is_conditioned()                                            # vis_x still missing
    media = torch.zeros(2, 2, 64, fx["mpt"]["d_model"])
    for layer in lm._get_decoder_layers():
        layer.condition_vis_x(media)
    assert lm.is_conditioned()
    for layer in lm._get_decoder_layers():
        assert layer.media_locations.dtype == torch.bool
        assert torch.equal(layer.media_locations, fx["lang_x"] == fx["media_id"])
    lm.clear_conditioned_layers()
    assert not lm.is_conditioned()
    assert all(layer.vis_x is None and layer.media_locations is None for layer in lm._get_decoder_layers())


def test_gradient_bucket_layout_splits_the_resampler_in_backward_order():
    """train.hot_path_parameters / GradBucket: gated blocks last-to-first, then the resampler in the order its gradients
    become final (final norm, layers depth-1..0, then latents), the resampler cut into two chunks, every trainable
    hot-path parameter exactly once, chunks contiguous and ending on group boundaries."""
    from open_flamingo_b200.testing import build_flamingo
    from open_flamingo_b200.train import GradBucket, hot_path_parameters
    vit = dict(image_size=56, patch_size=14, width=128, layers=1, heads=2, output_dim=128)
    mpt = dict(d_model=128, n_heads=2, n_layers=4, vocab_size=61, max_seq_len=64, expansion_ratio=2)
    model, _, _ = build_flamingo(vit, mpt, cross_attn_every_n_layers=2, device="cpu", seed=0)
    groups = hot_path_parameters(model)
    kinds = [(k, i) for k, i, _ in groups]
    depth = len(model.perceiver.layers)
    assert kinds == [("xattn", 3), ("xattn", 1), ("perceiver_norm", -1)] + [("perceiver_layer", j) for j in reversed(range(depth))] + \
        [("perceiver", -1)]
    assert [n for n, _ in groups[-1][2]] == ["perceiver.latents"]
    names = [n for _, _, ps in groups for n, _ in ps]
    want = {n for n, p in model.named_parameters(remove_duplicate=False)
            if p.requires_grad and n.startswith(("perceiver.", "lang_encoder.gated_cross_attn_layers."))}
    assert len(names) == len(set(names)) and set(names) == want
    bucket = GradBucket(groups, num_chunks=3, flatten_params=False)
    assert bucket.chunks[0][0] == 0 and bucket.chunks[-1][1] == bucket.total
    assert all(a[1] == b[0] for a, b in zip(bucket.chunks, bucket.chunks[1:]))
    # 2 xattn chunks (num_chunks - 1) + the resampler in two: [norm, layers depth-1 .. 1] closes when layer 1's backward is
    # done and overlaps layer 0's; [layer 0, latents] is what remains for after the backward
    assert len(bucket.chunks) == 2 + 2
    n_groups = len(groups)
    assert bucket.group_to_chunk[n_groups - 1] == bucket.group_to_chunk[n_groups - 2] == len(bucket.chunks) - 1
    assert all(bucket.group_to_chunk[g] == len(bucket.chunks) - 2 for g in range(2, n_groups - 2))
    assert bucket._chunk_last_group[len(bucket.chunks) - 2] == n_groups - 3


def test_sm_reservation_contract_and_nccl_cta_cap(monkeypatch):
    """Host-side state only (no launch): ofk_gemm_reserve_sms clamps to whole SM pairs in [0, 64] and returns the previous
    value; ops.set_comm_in_flight toggles it around the all-reduce window; configure_nccl_for_overlap caps NCCL's CTAs at
    the same number unless the launcher already chose one."""
    import __graft_entry__ as g
    g.build()
    from open_flamingo_b200 import _lib, ops, train
    lib = _lib.lib()
    lib.ofk_gemm_reserve_sms(0)
    assert lib.ofk_gemm_reserve_sms(17) == 0
    assert lib.ofk_gemm_reserve_sms(1000) == 16      # 17 -> 16: whole pairs
    assert lib.ofk_gemm_reserve_sms(-5) == 64        # clamped
    assert lib.ofk_gemm_reserve_sms(0) == 0
    assert ops.COMM_RESERVED_SMS > 0, "default build reserves SMs while gradient chunks are in flight"
    ops.set_comm_in_flight(True)
    assert lib.ofk_gemm_reserve_sms(ops.COMM_RESERVED_SMS) == (ops.COMM_RESERVED_SMS & ~1)
    ops.set_comm_in_flight(False)
    assert lib.ofk_gemm_reserve_sms(0) == 0
    monkeypatch.setenv("NCCL_MAX_CTAS", "0")   # so that teardown restores whatever the environment had
    monkeypatch.delenv("NCCL_MAX_CTAS")
    train.configure_nccl_for_overlap()
    import os
    assert os.environ["NCCL_MAX_CTAS"] == str(ops.COMM_RESERVED_SMS)
    monkeypatch.setenv("NCCL_MAX_CTAS", "4")
    train.configure_nccl_for_overlap()
    assert os.environ["NCCL_MAX_CTAS"] == "4"
