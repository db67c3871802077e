"""Host-side logic of the multi-GPU path on CPU (gloo, world_size 2): sharding helpers, result
gathering, and the data-parallel training rule (local step normalised by the GLOBAL position
count + SUM all-reduce + replicated update == single-process training on the whole batch).
The CUDA training step itself is covered on the GPU by tests/test_train_gpu.py
(test_shard_gradients_sum_to_the_full_batch_gradient)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from image_captioning_b200 import parallel


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 8, 4096, 100000):
        for world in (1, 2, 3, 4, 8):
            b = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(5, 2, 2)
    assert parallel.shard(list(range(10)), 1, 3) == [4, 5, 6]


class _StubModel(object):
    """Stand-in with the training surface of RoiCaptionModel: a linear softmax-free regressor whose
    'positions' are the (row, time) cells of gt; loss = sum((f.w - gt)^2) * inv_count."""

    def __init__(self, dim, P):
        g = torch.Generator().manual_seed(7)
        self.w = torch.randn(dim, P, generator=g, dtype=torch.float64)
        self.g = torch.zeros(dim * P, dtype=torch.float64)
        self.steps = 0

    def train_step_device(self, features, gt, targets, inv_count):
        f = torch.as_tensor(features, dtype=torch.float64)
        y = torch.as_tensor(gt, dtype=torch.float64)
        r = f @ self.w - y
        self.g.copy_((2.0 * inv_count * (f.t() @ r)).reshape(-1))
        return (r * r).sum() * inv_count

    def grad_buffer(self):
        return self.g

    def param_buffer(self):
        return self.w.view(-1)

    def apply_gradients(self, grad_scale=1.0):
        self.steps += 1
        self.w -= 0.05 * grad_scale * self.g.view_as(self.w)

    # sharded-optimiser surface: the update on ranges of the flat buffers, then a refresh hook
    def apply_gradients_ranges(self, ranges, grad_scale=1.0):
        self.steps += 1
        flat = self.w.view(-1)
        for off, n in ranges:
            flat[off:off + n] -= 0.05 * grad_scale * self.g[off:off + n]

    def params_updated(self):
        self.refreshed = getattr(self, "refreshed", 0) + 1

    def grad_buckets(self):
        n = self.g.numel()
        return [(n // 2, n - n // 2), (0, n // 2)]          # two buckets in backward-completion order


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, feats, gt, bucket_mb, out, shard_optimizer=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _StubModel(feats.shape[1], gt.shape[1])
        if rank == 1:
            model.w += 1.0                                   # replicas start apart: broadcast must fix it
        tr = parallel.DataParallelTrainer(model, bucket_mb=bucket_mb, shard_optimizer=shard_optimizer)
        assert tr.shard_optimizer == bool(shard_optimizer)
        tr.broadcast_parameters(0)
        lo, hi = parallel.shard_bounds(feats.shape[0], rank, world)
        losses = [tr.train_on_batch([feats[lo:hi], gt[lo:hi]]) for _ in range(3)]
        rows = parallel.gather_rows(torch.arange(lo, hi, dtype=torch.int32)[:, None].repeat(1, 2), feats.shape[0])
        out[rank] = (losses, model.w.clone(), rows)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bucket_mb", [0, 1e-4])
def test_data_parallel_training_equals_single_process(bucket_mb):
    torch.manual_seed(3)
    feats = torch.randn(11, 6, dtype=torch.float64)          # 11 rows over 2 ranks: unequal shards (6 + 5)
    gt = torch.randn(11, 4, dtype=torch.float64)
    ref = _StubModel(6, 4)
    ref_tr = parallel.DataParallelTrainer(ref)
    ref_losses = [ref_tr.train_on_batch([feats, gt]) for _ in range(3)]

    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, feats, gt, bucket_mb, out), nprocs=world, join=True)
    for r in range(world):
        losses, w, rows = out[r]
        np.testing.assert_allclose(losses, ref_losses, rtol=1e-12)
        torch.testing.assert_close(w, ref.w, rtol=1e-12, atol=1e-12)
        assert rows[:, 0].tolist() == list(range(11))
    torch.testing.assert_close(out[0][1], out[1][1], rtol=0, atol=0)     # replicas stay identical


def test_sharded_optimizer_equals_single_process():
    """ZeRO-1 style: reduce-scatter of every gradient bucket (emulated under gloo), range update on the owning rank,
    all-gather of the updated parameter slices == single-process training; replicas stay bit-identical; the remainder
    that does not divide by 4 * world floats is updated redundantly on every rank."""
    torch.manual_seed(4)
    feats = torch.randn(11, 7, dtype=torch.float64)          # 7 x 5 = 35 parameters: chunks of 4 per rank + remainders
    gt = torch.randn(11, 5, dtype=torch.float64)
    ref = _StubModel(7, 5)
    ref_tr = parallel.DataParallelTrainer(ref)
    ref_losses = [ref_tr.train_on_batch([feats, gt]) for _ in range(3)]

    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, feats, gt, 0, out, True), nprocs=world, join=True)
    for r in range(world):
        losses, w, rows = out[r]
        np.testing.assert_allclose(losses, ref_losses, rtol=1e-12)
        torch.testing.assert_close(w, ref.w, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(out[0][1], out[1][1], rtol=0, atol=0)


def test_host_batch_prefetcher_order_and_slot_reuse():
    """put / get / done bookkeeping on a CPU device (synchronous copies): batches come back in the order they were put,
    a slot is reused only after its batch was fetched, shapes may change between batches."""
    from image_captioning_b200.parallel import HostBatchPrefetcher
    pf = HostBatchPrefetcher("cpu", depth=2)
    with pytest.raises(RuntimeError):
        pf.get()
    with pytest.raises(RuntimeError):
        pf.done()
    batches = [(torch.full((3, 4), float(i)), torch.arange(5, dtype=torch.int32) + i) for i in range(5)]
    batches.append((torch.full((2, 7), 9.0), torch.arange(3, dtype=torch.int32)))         # a last, smaller batch
    pf.put(*batches[0])
    for k in range(len(batches)):
        if k + 1 < len(batches):
            pf.put(*batches[k + 1])
            assert pf.pending() == 2
            with pytest.raises(RuntimeError):
                pf.put(*batches[k + 1])                     # both slots hold unfetched batches
        a, b = pf.get()
        assert torch.equal(a, batches[k][0]) and torch.equal(b, batches[k][1])
        pf.done()
    assert pf.pending() == 0
    with pytest.raises(ValueError):
        HostBatchPrefetcher("cpu", depth=0)


def _global_batches(n_batches, rows, dim, P, seed=11):
    g = torch.Generator().manual_seed(seed)
    for _ in range(n_batches):
        yield ([torch.randn(rows, dim, generator=g, dtype=torch.float64), torch.randn(rows, P, generator=g, dtype=torch.float64)], None)


def _fit_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _StubModel(6, 4)
        tr = parallel.DataParallelTrainer(model)
        seen = []

        class Cb(object):
            def on_epoch_end(self, epoch, logs):
                seen.append((epoch, logs["loss"]))
        hist = tr.fit_generator(_global_batches(6, 9, 6, 4), steps_per_epoch=3, epochs=2, verbose=0, callbacks=[Cb()],
                                max_queue_size=2, workers=1)
        out[rank] = (hist.history["loss"], model.w.clone(), seen)
    finally:
        dist.destroy_process_group()


def test_fit_generator_deals_global_batches_over_the_ranks():
    """DataParallelTrainer.fit_generator: every rank runs the same generator of GLOBAL batches (9 rows: shards of 5 + 4) and
    trains on its rows; losses and weights equal single-process training on the whole batches; callbacks on rank 0 only."""
    ref = _StubModel(6, 4)
    ref_hist = parallel.DataParallelTrainer(ref).fit_generator(_global_batches(6, 9, 6, 4), steps_per_epoch=3, epochs=2,
                                                                verbose=0, workers=0)
    assert len(ref_hist.history["loss"]) == 2 and ref.steps == 6
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_fit_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        losses, w, seen = out[r]
        np.testing.assert_allclose(losses, ref_hist.history["loss"], rtol=1e-12)
        torch.testing.assert_close(w, ref.w, rtol=1e-12, atol=1e-12)
        assert len(seen) == (2 if r == 0 else 0)
    with pytest.raises(ValueError):
        parallel.DataParallelTrainer(_StubModel(6, 4)).fit_generator(_global_batches(1, 9, 6, 4))
