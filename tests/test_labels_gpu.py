"""GPU: ofk_make_labels (through ops.make_labels -> C ABI) is bit-exact against the golden labels of the unmodified
reference training loop and against the oracle on larger / ragged shapes."""
import pytest
import torch

from helpers_golden import load

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from open_flamingo_b200 import ops as _ops
    return _ops


def test_labels_match_reference_golden(ops):
    fx = load("labels")
    for c in fx["cases"]:
        ids = c["input_ids"].long().cuda()
        got = ops.make_labels(ids, fx["pad_id"], fx["media_id"], fx["eoc_id"], interleaved=c["interleaved"])
        assert got.dtype == torch.int64 and torch.equal(got.cpu(), c["labels"].long()), (c["interleaved"], tuple(ids.shape))


@pytest.mark.parametrize("B,T", [(32, 256), (3, 2048), (5, 1), (2, 255), (2, 257), (64, 513)])
def test_labels_match_oracle_on_random_rows(ops, B, T):
    from oracle import flamingo_oracle as O
    g = torch.Generator().manual_seed(B * 1000 + T)
    ids = torch.randint(0, 50280, (B, T), generator=g)
    u = torch.rand((B, T), generator=g)
    media, eoc, pad = 50278, 50277, 50279
    ids[u < 0.02] = media
    ids[(u >= 0.02) & (u < 0.05)] = eoc
    ids[0, T // 2:] = pad                                                    # right-padded row
    ids[B - 1, : T // 3] = pad                                               # left-padded row
    for inter in (False, True):
        want = O.make_labels(ids, pad, media, eoc, interleaved=inter)
        assert torch.equal(ops.make_labels(ids.cuda(), pad, media, eoc, interleaved=inter).cpu(), want)
    # strided rows (a column slice of a wider buffer) and a caller-provided output
    wide = torch.full((B, T + 9), 7, dtype=torch.int64)
    wide[:, 3:3 + T] = ids
    out = torch.empty((B, T), dtype=torch.int64, device="cuda")
    ops.make_labels(wide.cuda()[:, 3:3 + T], pad, media, eoc, interleaved=True, out=out)
    assert torch.equal(out.cpu(), O.make_labels(ids, pad, media, eoc, interleaved=True))


def test_labels_edge_cases(ops):
    empty = torch.empty((0, 16), dtype=torch.int64, device="cuda")
    assert ops.make_labels(empty, 1, 2, 3, interleaved=True).shape == (0, 16)
    with pytest.raises(ValueError):
        ops.make_labels(torch.zeros(2, 4, dtype=torch.int32, device="cuda"), 1, 2, 3)
    with pytest.raises(ValueError):
        ops.make_labels(torch.zeros(2, 4, dtype=torch.int64, device="cuda"), 1, 2, None, interleaved=True)
    with pytest.raises(RuntimeError):
        ops.make_labels(torch.zeros(2, 4, dtype=torch.int64), 1, 2, 3)       # CPU tensor: no fallback
