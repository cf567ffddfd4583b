"""Training-label rule (train_utils.py:102-106, :126-149): the oracle's literal loop restatement against the tensors
the UNMODIFIED reference `train_one_epoch` handed to the model (tests/golden/labels.pt, made by
tests/golden/make_golden_labels.py), and against an independent closed form (two-state scan) on random rows --
the same scan the CUDA kernel implements."""
import torch

from helpers_golden import load
from oracle import flamingo_oracle as O


def closed_form(ids, pad, media, eoc, interleaved):
    """label kept iff not pad / <image> and (not interleaved or the latest marker strictly before it is an <image>)."""
    B, T = ids.shape
    out = ids.clone()
    for b in range(B):
        is_open = False
        for t in range(T):
            tok = int(ids[b, t])
            if tok == pad or tok == media or (interleaved and not is_open):
                out[b, t] = -100
            if tok != pad:
                if tok == media:
                    is_open = True
                elif tok == eoc:
                    is_open = False
    return out


def test_oracle_labels_match_the_reference_training_loop():
    fx = load("labels")
    assert len(fx["cases"]) == 8
    for c in fx["cases"]:
        ids = c["input_ids"].long()
        got = O.make_labels(ids, fx["pad_id"], fx["media_id"], fx["eoc_id"], interleaved=c["interleaved"])
        assert torch.equal(got, c["labels"].long()), (c["interleaved"], tuple(ids.shape))


def test_closed_form_scan_equals_the_loop_form():
    fx = load("labels")
    for c in fx["cases"]:
        ids = c["input_ids"].long()
        assert torch.equal(closed_form(ids, fx["pad_id"], fx["media_id"], fx["eoc_id"], c["interleaved"]), c["labels"].long())
    g = torch.Generator().manual_seed(5)
    for T in (1, 7, 33, 64):
        ids = torch.randint(0, 8, (40, T), generator=g)       # tiny vocabulary: markers and pads are dense
        for inter in (False, True):
            assert torch.equal(closed_form(ids, 7, 5, 6, inter), O.make_labels(ids, 7, 5, 6, interleaved=inter))
    # a marker id equal to the pad id is never seen as a marker (the reference masks pads first, :127)
    ids = torch.tensor([[5, 1, 2, 6, 3, 5, 4]])
    assert torch.equal(closed_form(ids, 6, 5, 6, True), O.make_labels(ids, 6, 5, 6, interleaved=True))


def test_embedding_grad_mask_matches_the_reference_training_loop():
    """train_utils.py:172-194: after the reference loop ran with trainable LM embeddings, only the <image> and
    <|endofchunk|> rows of the embedding gradient are non-zero (golden: tests/golden/make_golden_labels.py)."""
    import types
    from open_flamingo_b200.train import mask_embedding_grad
    fx = load("labels")
    coeff, golden = fx["embed_coeff"], fx["embed_grad_after"]
    emb = torch.nn.Embedding(*coeff.shape)
    emb.weight.grad = coeff.clone()                                            # raw gradient of the golden run's loss
    model = types.SimpleNamespace(lang_encoder=types.SimpleNamespace(get_input_embeddings=lambda: emb),
                                  media_token_id=fx["media_id"], eoc_token_id=fx["eoc_id"])
    g = mask_embedding_grad(model)
    assert g is emb.weight.grad
    support = lambda t: (t != 0).any(1).nonzero().flatten().tolist()
    assert support(g) == support(golden) == sorted([fx["media_id"], fx["eoc_id"]])
    rows = [fx["media_id"], fx["eoc_id"]]
    assert torch.equal(g[rows], coeff[rows])
    ratio = golden[rows] / coeff[rows]                                         # the loop also clips: one global scale
    assert torch.allclose(ratio, ratio.flatten()[0].expand_as(ratio), rtol=1e-5)
    emb.weight.grad = None
    assert mask_embedding_grad(model) is None
