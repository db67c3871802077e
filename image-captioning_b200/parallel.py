"""One process per GPU: the replacement of the reference's in-process ``ParallelModel``
(/root/reference/dense_img_cap_separate_models/parallel_model.py:22-102: tf.split of the inputs
over towers that share variables, outputs concatenated on the CPU).

The path shards by construction: RoIs are independent and a RoI needs only
its own image's pyramid, so inference shards IMAGES (ROIAlign, captions) or RoI ROWS (beam over
pre-extracted features) across ranks with no data-path collective; weights are replicated.  The
only exchange step is the decoder-training gradient: every rank runs the training step on its
slice of the global batch with the loss normalised by the GLOBAL position count, the flat fp32
gradient buffer is summed with ONE NCCL all-reduce over NVLink/NVSwitch, and every rank applies
the same optimiser update (replicas stay bit-identical: the all-reduce result is identical on all
ranks and the update is deterministic).

``torch.distributed`` is plumbing only (rendezvous + the NCCL communicator); under ``gloo`` on CPU
the same host logic is exercised by tests/test_parallel.py with a stand-in model.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous, balanced split of n items: the first n % world ranks get one extra item."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    q, r = divmod(int(n), world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard(seq, rank, world):
    """The slice of ``seq`` (images, RoI rows, captions ...) this rank owns."""
    lo, hi = shard_bounds(len(seq), rank, world)
    return seq[lo:hi]


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def gather_rows(local, total_rows, group=None):
    """Concatenate per-rank result rows (token ids, scores) in rank order on every rank.  Used only
    to assemble the final report -- never on the data path.  ``local`` holds this rank's
    shard_bounds(total_rows) rows."""
    rank, world = _world(group)
    if world == 1:
        return local
    sizes = [hi - lo for lo, hi in (shard_bounds(total_rows, r, world) for r in range(world))]
    pad = max(sizes)                                        # equal-sized messages (shards differ by <= 1 row)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], 0)


class HostBatchPrefetcher(object):
    """Uploads training batches from host memory one step ahead of the step that consumes them -- the device-side half of
    what ``fit_generator(..., max_queue_size=, workers=)`` does for the reference (text_generation_model.py:470-472: Keras
    keeps a queue of generator batches filled while the previous batch trains).

    ``put(*host_tensors)`` copies a batch into one of ``depth`` device slots on a private copy stream (pinned host tensors
    make the copies asynchronous; the host tensors must stay unmodified until the matching ``get``); ``get()`` makes the
    CURRENT stream wait for the oldest uploaded batch and returns its device tensors; ``done()`` marks them consumed by
    everything enqueued on the current stream so far, which is what lets ``put`` reuse the slot.  Loop::

        pf.put(feats0, gt0)
        for k in range(steps):
            if k + 1 < steps: pf.put(*batch(k + 1))     # rides under step k's kernels
            d_feats, d_gt = pf.get()
            loss = trainer.train_step(d_feats, d_gt, None, npos); pf.done()

    Measured on cfg3 (205 MB of fp32 RoI features per 4096-RoI step, 3.7 ms over PCIe against a 7.7 ms step): see
    bench.py --workload train, ``e2e``.  On a CPU device the copies are plain synchronous copies (gloo tests)."""

    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.depth = int(depth)
        if self.depth < 1:
            raise ValueError("depth must be >= 1")
        self._cuda = self.device.type == "cuda"
        self._copy = torch.cuda.Stream(device=self.device) if self._cuda else None
        self._bufs = [None] * self.depth
        self._ready = [None] * self.depth
        self._consumed = [None] * self.depth
        self._n_put = self._n_get = 0
        self._last = None

    def pending(self):
        """Batches uploaded (or uploading) and not yet handed out."""
        return self._n_put - self._n_get

    def put(self, *host_tensors):
        if self.pending() >= self.depth:
            raise RuntimeError("all %d slots hold batches that were not fetched yet: call get() first" % self.depth)
        slot = self._n_put % self.depth
        hts = [torch.as_tensor(h) for h in host_tensors]
        bufs = self._bufs[slot]
        if bufs is None or len(bufs) != len(hts) or any(b.shape != h.shape or b.dtype != h.dtype for b, h in zip(bufs, hts)):
            if self._cuda and self._consumed[slot] is not None:
                self._consumed[slot].synchronize()          # the old buffers may still be read: do not free them under it
            bufs = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in hts]
            self._bufs[slot] = bufs
            if self._cuda:                                  # the allocator may hand out memory the current stream still uses
                self._copy.wait_stream(torch.cuda.current_stream(self.device))
        if self._cuda:
            with torch.cuda.stream(self._copy):
                if self._consumed[slot] is not None:
                    self._copy.wait_event(self._consumed[slot])
                for b, h in zip(bufs, hts):
                    b.copy_(h, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy)
            self._ready[slot] = ev
        else:
            for b, h in zip(bufs, hts):
                b.copy_(h)
        self._n_put += 1

    def get(self):
        if self.pending() <= 0:
            raise RuntimeError("get() without a batch in flight: call put() first")
        slot = self._n_get % self.depth
        if self._cuda:
            torch.cuda.current_stream(self.device).wait_event(self._ready[slot])
        self._n_get += 1
        self._last = slot
        return tuple(self._bufs[slot])

    def done(self):
        if self._last is None:
            raise RuntimeError("done() before get()")
        if self._cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._consumed[self._last] = ev


class DataParallelTrainer(object):
    """Data-parallel ``train_on_batch`` over the ranks of ``group``.

    ``model`` needs the training surface of RoiCaptionModel: ``train_step_device(features, gt,
    targets, inv_count) -> device scalar``, ``grad_buffer() -> flat fp32 tensor`` and
    ``apply_gradients(grad_scale)``.  Semantics equal ONE process training on the concatenated
    global batch: loss = mean over all positions of all ranks."""

    def __init__(self, model, group=None, bucket_mb=0, overlap=True, shard_optimizer=False):
        self.model, self.group = model, group
        self.overlap, self._comm, self._buckets = overlap, None, None
        self.rank, self.world = _world(group)
        self.bucket_elems = int(bucket_mb * (1 << 20) // 4)
        # Sharded optimiser (ZeRO-1 style): every gradient bucket is reduce-SCATTERED, each rank applies the update to
        # the slice it received, and the updated parameter slices are all-gathered -- the same bytes on the wire as the
        # all-reduce, but the optimiser pass (31.6 M parameters, 0.33 ms on a B200) shrinks by the world size instead of
        # being repeated on every rank.  Replicas stay bit-identical: every rank ends with the same gathered buffer.
        # Measured on B200s (cfg3, global batch 4096 x 16): 2 ranks 4.96-5.02 ms sharded vs 4.80 replicated (the all-gather of
        # the updated parameters is not overlapped with anything, the all-reduce was hidden under the backward pass), 8 ranks
        # 2.755 vs 2.81 ms -- "auto" shards from 4 ranks on.
        if shard_optimizer == "auto":
            shard_optimizer = self.world >= 4
        self.shard_optimizer = bool(shard_optimizer) and self.world > 1 and hasattr(model, "apply_gradients_ranges")

    def _slices(self, off, n):
        """Split [off, off + n) into world equal chunks (multiples of 4 floats) + a replicated remainder."""
        chunk = (n // (self.world * 4)) * 4
        return chunk, off + chunk * self.world, n - chunk * self.world

    def _reduce_scatter(self, g, off, n):
        """Sum over ranks of g[off:off+n]; afterwards this rank's chunk (and the replicated remainder) hold the totals."""
        chunk, rem_off, rem = self._slices(off, n)
        nccl = dist.get_backend(self.group) == "nccl"
        if chunk > 0:
            whole = g[off:off + chunk * self.world]
            if nccl:                                        # in place: the output is this rank's chunk of the input
                dist.reduce_scatter_tensor(whole[self.rank * chunk:(self.rank + 1) * chunk], whole, group=self.group)
            else:                                           # gloo has no reduce-scatter: same result through an all-reduce
                dist.all_reduce(whole, group=self.group)
        if rem > 0:
            dist.all_reduce(g[rem_off:rem_off + rem], group=self.group)

    def _all_gather(self, p, off, n):
        chunk, _, _ = self._slices(off, n)
        if chunk <= 0:
            return
        whole = p[off:off + chunk * self.world]
        mine = whole[self.rank * chunk:(self.rank + 1) * chunk]
        if dist.get_backend(self.group) == "nccl":
            dist.all_gather_into_tensor(whole, mine, group=self.group)       # in place
        else:
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine.clone(), group=self.group)
            for r, part in enumerate(parts):
                whole[r * chunk:(r + 1) * chunk].copy_(part)

    def broadcast_parameters(self, src=0):
        """Make every replica start from rank ``src``'s trainable weights (Keras towers share variables)."""
        if self.world > 1:
            dist.broadcast(self.model.param_buffer(), src, group=self.group)

    def _count_positions(self, gt, targets):
        if targets is None:
            n = float(gt.shape[0] * gt.shape[1])
        else:
            t = torch.as_tensor(targets)
            # one-hot rows [N,P,V]: a position counts when its row is not all-zero; ids [N,P]: when >= 0
            n = float((t.sum(-1) > 0).sum()) if t.dim() == 3 else float((t >= 0).sum())
        return torch.tensor([n], dtype=torch.float64)

    def train_step(self, features, gt, targets=None, global_positions=None, **step_options):
        """Local forward/backward + gradient all-reduce + update.  Returns the GLOBAL mean loss as a
        device scalar (no host sync).  ``global_positions``: number of loss positions over all
        ranks; derived with one tiny all-reduce if not given.  ``step_options`` go to
        ``model.train_step_device`` (recurrent_dropout / dropout_seed / dropout_step / row_offset: pass this rank's
        first global row as row_offset and the masks equal those of the unsharded batch)."""
        if global_positions is None:
            cnt = self._count_positions(gt, targets)
            if self.world > 1:
                g = self.model.grad_buffer()
                cnt = cnt.to(g.device)
                dist.all_reduce(cnt, group=self.group)
            global_positions = float(cnt.item())
        if global_positions <= 0:
            return torch.zeros(())
        loss = self.model.train_step_device(features, gt, targets, 1.0 / global_positions, **step_options)
        if self.shard_optimizer:
            g, p = self.model.grad_buffer(), self.model.param_buffer()
            buckets = self.model.grad_buckets() if hasattr(self.model, "grad_buckets") else [(0, g.numel())]
            overlap = self.overlap and hasattr(self.model, "wait_grad_bucket") and g.is_cuda
            if overlap:
                cur = torch.cuda.current_stream(g.device)
                if self._comm is None:
                    self._comm = torch.cuda.Stream(device=g.device)
                for i, (off, n) in enumerate(buckets):
                    self.model.wait_grad_bucket(i, self._comm)
                    with torch.cuda.stream(self._comm):
                        self._reduce_scatter(g, off, n)
                cur.wait_stream(self._comm)
            else:
                for off, n in buckets:
                    self._reduce_scatter(g, off, n)
            dist.all_reduce(loss, group=self.group)
            ranges = []
            for off, n in buckets:
                chunk, rem_off, rem = self._slices(off, n)
                ranges.append((off + self.rank * chunk, chunk))
                ranges.append((rem_off, rem))              # the remainder (< 4 * world floats) is updated on every rank
            self.model.apply_gradients_ranges(ranges, 1.0)
            for off, n in buckets:
                self._all_gather(p, off, n)
            self.model.params_updated()
            return loss
        if self.world > 1:
            g = self.model.grad_buffer()
            if self.overlap and hasattr(self.model, "wait_grad_bucket"):
                # The step above is only ENQUEUED on the compute stream.  Each gradient bucket is
                # all-reduced on the communication stream as soon as the backward pass has produced
                # it (vocabulary projection first), i.e. under the remaining BPTT / head kernels.
                cur = torch.cuda.current_stream(g.device)
                if self._comm is None:
                    self._comm = torch.cuda.Stream(device=g.device)
                    self._buckets = self.model.grad_buckets()
                for i, (off, n) in enumerate(self._buckets):
                    self.model.wait_grad_bucket(i, self._comm)
                    with torch.cuda.stream(self._comm):
                        dist.all_reduce(g[off:off + n], group=self.group)
                cur.wait_stream(self._comm)
            elif self.bucket_elems > 0:
                # reverse-layer order (vocabulary projection first) so the last-produced head
                # gradients are the last bucket on the wire
                hs = [dist.all_reduce(g[max(0, e - self.bucket_elems):e], group=self.group, async_op=True)
                      for e in range(g.numel(), 0, -self.bucket_elems)]
                for h in hs:
                    h.wait()
            else:
                dist.all_reduce(g, group=self.group)
            dist.all_reduce(loss, group=self.group)
        self.model.apply_gradients(1.0)
        return loss

    def train_on_batch(self, x, y=None):
        feats, gt = x
        return float(self.train_step(feats, gt, y).item())

    def fit_generator(self, generator, steps_per_epoch=None, epochs=1, verbose=1, callbacks=None, max_queue_size=10,
                      workers=1, initial_epoch=0, global_batches=True, **kwargs):
        """``ParallelModel(model, gpu_count).fit_generator(...)`` of the reference (parallel_model.py:22-102 wraps the
        Keras model, so the training script's fit_generator call -- text_generation_model.py:470-472 -- feeds ONE
        generator whose batches tf.split deals over the towers).  Here every rank runs the same generator (same seed)
        and, with ``global_batches=True``, takes its ``shard_bounds`` rows of every batch -- features, captions and
        targets alike; ``global_batches=False`` means the generator already yields this rank's shard.  The loss reported
        per epoch is the global mean (identical on all ranks); callbacks run on rank 0 only (checkpoint / CSV writers).
        Batches are pulled through the same worker-thread queue as the single-GPU fit_generator."""
        from .text_model import GeneratorQueue, History
        if steps_per_epoch is None:
            raise ValueError("steps_per_epoch is required for a generator")
        hist = History()
        cbs = (callbacks or []) if self.rank == 0 else []
        for cb in cbs:
            if hasattr(cb, "set_model"):
                cb.set_model(self.model)
        for cb in cbs:
            if hasattr(cb, "on_train_begin"):
                cb.on_train_begin()

        def mine(a):
            if a is None or not global_batches:
                return a
            lo, hi = shard_bounds(len(a), self.rank, self.world)
            return a[lo:hi]
        with GeneratorQueue(generator, max_queue_size, workers) as batches:
            for epoch in range(initial_epoch, epochs):
                losses = []
                for _ in range(steps_per_epoch):
                    x, y = batches.get()[:2]
                    feats, gt = x
                    losses.append(float(self.train_step(mine(feats), mine(gt), mine(y)).item()))
                logs = {"loss": float(sum(losses) / max(1, len(losses)))}
                hist.epoch.append(epoch)
                for k, v in logs.items():
                    hist.history.setdefault(k, []).append(v)
                if verbose and self.rank == 0:
                    print("Epoch %d/%d - " % (epoch + 1, epochs) + " - ".join("%s: %.4f" % kv for kv in logs.items()))
                for cb in cbs:
                    if hasattr(cb, "on_epoch_end"):
                        cb.on_epoch_end(epoch, logs)
        for cb in cbs:
            if hasattr(cb, "on_train_end"):
                cb.on_train_end()
        return hist
