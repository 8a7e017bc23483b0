"""World-size-2 gloo tests (CPU) of the batch-sharded data-parallel host logic
(emoasr_b200/sharding.py): shard ranges, bucketed gradient all-reduce, mean-of-replica-means
normalisation (asr/train_asr.py:67-71).  The per-rank loss is the oracle's CPU restatement of the
path -- on the GPU box the same host logic drives the CUDA kernels (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from emoasr_b200 import sharding


def test_shard_ranges_cover_batch_exactly_once():
    for B in (1, 7, 8, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_batch_slices_only_per_utterance_tensors():
    batch = {"xs": torch.arange(24.).view(4, 3, 2), "xlens": torch.tensor([3, 3, 2, 1]), "utt_ids": ["a", "b", "c", "d"],
             "scalar": torch.tensor(5)}
    s = sharding.shard_batch(batch, 1, 2)
    assert s["xs"].shape == (2, 3, 2) and torch.equal(s["xlens"], torch.tensor([2, 1]))
    assert s["utt_ids"] == ["a", "b", "c", "d"] and int(s["scalar"]) == 5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_problem():
    g = torch.Generator().manual_seed(0)
    B, T, U, V, He, Hd, J = 4, 9, 4, 13, 6, 5, 8
    eouts = torch.randn(B, T, He, generator=g)
    douts = torch.tanh(torch.randn(B, U + 1, Hd, generator=g))
    ys = torch.randint(1, V, (B, U), generator=g)
    tl = torch.tensor([9, 8, 9, 7])
    ul = torch.tensor([4, 3, 4, 2])
    torch.manual_seed(7)
    lin = torch.nn.ModuleList([torch.nn.Linear(He, J), torch.nn.Linear(Hd, J), torch.nn.Linear(J, V)])
    return eouts, douts, ys, tl, ul, lin


def _loss(lin, eouts, douts, ys, tl, ul):
    from oracle import torch_path
    return torch_path.rnnt_joint_loss(eouts, douts, lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias,
                                      lin[2].weight, lin[2].bias, ys, tl, ul, blank=0)


def _worker(rank, world, port, out, own_grads=False, hooks=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    eouts, douts, ys, tl, ul, lin = _make_problem()
    batch = sharding.shard_batch({"eouts": eouts, "douts": douts, "ys": ys, "tl": tl, "ul": ul}, rank, world)
    if hooks == "reducer":   # bench.py's N > 1 mode: per-gradient all-reduce from hooks, gradients assigned by autograd
        red = sharding.GradReducer(lin.parameters(), small_numel=16)
        extra = torch.full((5,), float(rank + 1))
        for step in range(2):                      # two steps: the hooks must re-arm
            for p in lin.parameters():
                p.grad = None
            loss = _loss(lin, batch["eouts"], batch["douts"], batch["ys"], batch["tl"], batch["ul"])
            loss.backward()
            assert len(red._pending) >= 3          # launched during backward()
            if step == 1:
                red.start_extra(extra)
            red.finish()
        assert torch.allclose(extra, torch.full((5,), 1.5))
        mean_loss = sharding.mean_over_replicas(loss)
        if rank == 0:
            torch.save({"loss": mean_loss, "grads": [p.grad.clone() for p in lin.parameters()]}, out)
        dist.barrier()
        dist.destroy_process_group()
        return
    buckets = None
    if own_grads:   # gradients accumulate straight into the flat all-reduce buffers
        buckets = sharding.GradBuckets(lin.parameters(), bucket_bytes=256, own_grads=True)
        for p in lin.parameters():
            p.grad.fill_(123.0)     # stale values from a previous step ...
        buckets.zero()              # ... are cleared in place
        if hooks:                   # all-reduces launched from autograd hooks while backward() is still running
            buckets.attach_hooks()
    extra = torch.full((5,), float(rank + 1)) if hooks else None
    loss = _loss(lin, batch["eouts"], batch["douts"], batch["ys"], batch["tl"], batch["ul"])
    loss.backward()
    if buckets is None:
        buckets = sharding.GradBuckets(lin.parameters(), bucket_bytes=256)   # tiny buckets: several all-reduces
    assert len(buckets.buckets) > 1
    if hooks:
        assert len(buckets._pending) == len(buckets.buckets)   # every bucket was launched during backward()
        buckets.start_extra(extra)
    else:
        buckets.start()
    buckets.finish()
    if hooks:
        assert torch.allclose(extra, torch.full((5,), 1.5))    # mean of 1 and 2
    mean_loss = sharding.mean_over_replicas(loss)
    if rank == 0:
        torch.save({"loss": mean_loss, "grads": [p.grad.clone() for p in lin.parameters()]}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("own_grads,hooks", [(False, False), (True, False), (True, True), (False, "reducer")],
                         ids=["copy_buckets", "grads_in_buckets", "overlap_hooks", "per_gradient_reducer"])
def test_two_rank_sharded_step_matches_single_process(tmp_path, own_grads, hooks):
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, own_grads, hooks), nprocs=2, join=True)
    got = torch.load(out)
    eouts, douts, ys, tl, ul, lin = _make_problem()
    loss = _loss(lin, eouts, douts, ys, tl, ul)      # global mean over the 4 utterances
    loss.backward()
    assert abs(float(got["loss"]) - float(loss)) < 1e-5 * abs(float(loss))
    for p, g in zip(lin.parameters(), got["grads"]):
        assert np.allclose(g.numpy(), p.grad.numpy(), rtol=1e-4, atol=1e-6)


class _FlatLinear(torch.autograd.Function):
    """y = x W^T + b whose backward writes d_W and d_b side by side into one flat buffer, the way the folded joint's
    backward does on the GPU (functional.flat_views)."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return x @ w.t() + b

    @staticmethod
    def backward(ctx, gy):
        from emoasr_b200.functional import flat_views
        x, w = ctx.saved_tensors
        d_w, d_b = flat_views(x.device, tuple(w.shape), (w.size(0),))
        d_w.copy_(gy.t() @ x)
        d_b.copy_(gy.sum(0))
        return gy @ w, d_w, d_b


def _flat_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(3)
    w = torch.randn(5, 7, requires_grad=True)      # 35 elements: the bias piece starts at the padded offset 36
    b = torch.randn(5, requires_grad=True)
    other = torch.randn(3, requires_grad=True)     # a gradient that is NOT part of a flat buffer
    x = torch.randn(4, 7, generator=torch.Generator().manual_seed(10 + rank))
    red = sharding.GradReducer([w, b, other], small_numel=0)
    for step in range(2):
        for p in (w, b, other):
            p.grad = None
        ((_FlatLinear.apply(x, w, b) ** 2).sum() + (other * (rank + 1.0)).sum()).backward()
        assert w.grad.untyped_storage().data_ptr() == b.grad.untyped_storage().data_ptr()
        assert len(red._pending) == 2              # one collective for the flat buffer, one for `other`
        red.finish()
    if rank == 0:
        torch.save({"w": w.grad.clone(), "b": b.grad.clone(), "other": other.grad.clone()}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_reducer_allreduces_a_flat_gradient_buffer_once(tmp_path):
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_flat_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(3)
    w = torch.randn(5, 7, requires_grad=True)
    b = torch.randn(5, requires_grad=True)
    gw, gb = torch.zeros_like(w), torch.zeros_like(b)
    for rank in range(2):
        x = torch.randn(4, 7, generator=torch.Generator().manual_seed(10 + rank))
        y = x @ w.t() + b
        a, c = torch.autograd.grad((y ** 2).sum(), (w, b))
        gw += a / 2
        gb += c / 2
    assert torch.allclose(got["w"], gw, rtol=1e-5, atol=1e-6) and torch.allclose(got["b"], gb, rtol=1e-5, atol=1e-6)
    assert torch.allclose(got["other"], torch.full((3,), 1.5))
