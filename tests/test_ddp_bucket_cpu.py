"""CPU / gloo, world_size 2: the host-side logic of the data-parallel path (train.GradBucket) -- flat layout in
backward-completion order, chunking on layer boundaries, asynchronous chunk all-reduce -- reproduces the
gradient of a single-process run over the concatenated batch (what DDP guarantees, train.py:366)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


class ToyBlock(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.lin = nn.Linear(d, d, bias=False)
        self.attn_gate = nn.Parameter(torch.tensor([0.3]))

    def forward(self, x):
        return x + self.lin(x) * self.attn_gate.tanh()


class ToyLM(nn.Module):
    def __init__(self, d, n):
        super().__init__()
        self.gated_cross_attn_layers = nn.ModuleList([ToyBlock(d) if i % 2 == 0 else None for i in range(n)])


class ToyModel(nn.Module):
    def __init__(self, d=8, n=6):
        super().__init__()
        self.perceiver = nn.Linear(d, d)
        self.lang_encoder = ToyLM(d, n)

    def forward(self, x):
        x = self.perceiver(x)
        for blk in self.lang_encoder.gated_cross_attn_layers:
            if blk is not None:
                x = blk(x)
        return x.pow(2).mean()


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from open_flamingo_b200.train import GradBucket, hot_path_parameters
    torch.manual_seed(0)
    model = ToyModel()
    groups = hot_path_parameters(model)
    bucket = GradBucket(groups, num_chunks=3, flatten_params=True)
    # layout: last gated block first, resampler last; chunks end on group boundaries
    assert groups[0][1] == 4 and groups[-1][0] == "perceiver"
    assert bucket.chunks[0][0] == 0 and bucket.chunks[-1][1] == bucket.total
    assert all(a[1] == b[0] for a, b in zip(bucket.chunks, bucket.chunks[1:]))
    torch.manual_seed(100)
    full = torch.randn(8, 8)
    shard = full[rank * 4:(rank + 1) * 4]
    bucket.zero()
    loss = model(shard)
    loss.backward()
    _fire_block_hooks(bucket, groups)
    bucket.finish()
    mean_grads = bucket.grads / world
    # single-process reference over the whole batch (mean of per-shard means == DDP average)
    torch.manual_seed(0)
    ref = ToyModel()
    (0.5 * (ref(full[:4]) + ref(full[4:]))).backward()
    ref_named = dict(ref.named_parameters())
    for name, p, o, n in bucket.entries:
        got = mean_grads[o:o + n].view_as(p)
        assert torch.allclose(got, ref_named[name].grad, atol=1e-6), name
        assert p.grad.data_ptr() == bucket.grads[o:o + n].data_ptr()      # gradients live in the bucket
        assert p.data.data_ptr() == bucket.params[o:o + n].data_ptr()     # parameters live in the flat buffer
    if rank == 0:
        open(tmp, "w").write("ok")
    dist.destroy_process_group()


def _fire_block_hooks(bucket, groups):
    """emulate the fused backward's per-block callback (blocks finish in reverse order)"""
    for kind, idx, ps in groups:
        if kind == "xattn":
            params = [None] * 11
            params[5] = dict(ps)[f"lang_encoder.gated_cross_attn_layers.{idx}.attn_gate"]
            bucket.on_block_backward_done(params)


def _worker_accum(rank, world, port, tmp):
    """The reference's optimizer step is TWO backward passes (LAION + MMC4, train_utils.py:118,172): the first runs
    under no_sync() (accumulate only), chunk all-reduces start during the second; result == single-process gradient
    of the summed losses over the concatenated batches.  A second backward WITHOUT no_sync must raise, not corrupt."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from open_flamingo_b200.train import GradBucket, hot_path_parameters
    torch.manual_seed(0)
    model = ToyModel()
    groups = hot_path_parameters(model)
    bucket = GradBucket(groups, num_chunks=3, flatten_params=True)
    torch.manual_seed(100)
    full_a, full_b = torch.randn(8, 8), torch.randn(8, 8)
    sl = slice(rank * 4, (rank + 1) * 4)
    bucket.zero()
    with bucket.no_sync():
        model(full_a[sl]).backward()
        _fire_block_hooks(bucket, groups)
        assert not bucket._launched and not bucket._pending      # nothing reduced while accumulating
    model(full_b[sl]).backward()
    _fire_block_hooks(bucket, groups)
    assert bucket._launched                                        # chunks went out during the LAST backward
    bucket.finish()
    mean_grads = bucket.grads / world
    torch.manual_seed(0)
    ref = ToyModel()
    (0.5 * (ref(full_a[:4]) + ref(full_a[4:])) + 0.5 * (ref(full_b[:4]) + ref(full_b[4:]))).backward()
    ref_named = dict(ref.named_parameters())
    for name, p, o, n in bucket.entries:
        got = mean_grads[o:o + n].view_as(p)
        assert torch.allclose(got, ref_named[name].grad, atol=1e-6), name
    # misuse: a further backward after the chunks were launched must be refused loudly
    bucket.zero()
    model(full_a[sl]).backward()
    _fire_block_hooks(bucket, groups)
    raised = False
    try:
        model(full_b[sl]).backward()
        _fire_block_hooks(bucket, groups)
    except RuntimeError as e:
        raised = "no_sync" in str(e)
    bucket.finish()
    assert raised
    if rank == 0:
        open(tmp, "w").write("ok")
    dist.destroy_process_group()


def _spawn(fn, tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    flag = str(tmp_path / "ok")
    mp.spawn(fn, args=(2, port, flag), nprocs=2, join=True)
    assert open(flag).read() == "ok"


def test_grad_bucket_world2_two_micro_batches(tmp_path):
    _spawn(_worker_accum, tmp_path)


def test_grad_bucket_world2(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    flag = str(tmp_path / "ok")
    mp.spawn(_worker, args=(2, port, flag), nprocs=2, join=True)
    assert open(flag).read() == "ok"
