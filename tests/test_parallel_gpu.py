"""Data-parallel decoder training over NCCL with the REAL model (BASELINE configs[2]): two ranks, each
with its own GPU, train on halves of a global batch through DataParallelTrainer (bucketed all-reduce overlapped with
the backward pass) and must land on the weights a single process reaches on the whole batch.  Needs >= 2 GPUs
(`gpurun --gpus 2`); on the 1-GPU box the test skips -- the host logic is covered on CPU by tests/test_parallel.py."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPE = dict(V=1000, E=48, U=128, C=64)
P, B, STEPS = 6, 64, 3


def _data():
    from image_captioning_b200 import synth
    rng = np.random.default_rng(77)
    w = synth.synth_weights_v1(rng, trained_like=False, **SHAPE)
    feat = rng.standard_normal((B, 7, 7, SHAPE["C"])).astype(np.float32)
    gt = synth.synth_captions(rng, B, P, SHAPE["V"])
    return w, feat, gt


def _model(w, batch, device):
    import image_captioning_b200 as pkg
    cfg = pkg.DenseCapConfig(SHAPE["V"], w["imgcap_embedding_layer/embeddings"], batch, P)
    m = pkg.build_lstm_model([7, 7, SHAPE["C"]], cfg, SHAPE["U"], "training", dtype="bfloat16", device=device)
    m.set_weights(w)
    m.compile(optimizer=pkg.Adam(amsgrad=True), loss=pkg.roi_caption_loss)
    return m


def _worker(rank, world, port, dropout, out_path, shard_optimizer=False):
    import torch.distributed as dist
    from image_captioning_b200 import parallel
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    w, feat, gt = _data()
    lo, hi = parallel.shard_bounds(B, rank, world)
    m = _model(w, hi - lo, dev)
    tr = parallel.DataParallelTrainer(m, overlap=True, shard_optimizer=shard_optimizer)
    assert tr.shard_optimizer == bool(shard_optimizer)
    tr.broadcast_parameters()
    losses = []
    for it in range(STEPS):
        opts = dict(recurrent_dropout=dropout, dropout_seed=99, dropout_step=it, row_offset=lo) if dropout else {}
        losses.append(float(tr.train_step(feat[lo:hi], gt[lo:hi], None, float(B * P), **opts).item()))
        if it == 0:
            p_first = m.param_buffer().clone()
    p = m.param_buffer().clone()
    ps = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(ps, p)
    if rank == 0:
        np.savez(out_path, losses=np.array(losses), params=p.cpu().numpy(), params_first=p_first.cpu().numpy(),
                 replicas_identical=np.array([bool(torch.equal(ps[0], q)) for q in ps]))
    dist.destroy_process_group()


@pytest.mark.parametrize("dropout,shard_optimizer", [(0.0, False), (0.2, False), (0.0, True)])
def test_two_rank_nccl_training_equals_single_process(tmp_path, dropout, shard_optimizer):
    """shard_optimizer: ZeRO-1 style -- NCCL reduce-scatter of every gradient bucket (in place), the AMSGrad update on the
    rank's own ranges (dc_adam_step_range), NCCL all-gather of the updated parameter ranges, dc_decoder_params_updated."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "dp.npz")
    mp.spawn(_worker, args=(2, port, dropout, out, shard_optimizer), nprocs=2, join=True)
    got = np.load(out)
    assert got["replicas_identical"].all()
    # single process, whole batch, same steps
    w, feat, gt = _data()
    m = _model(w, B, torch.device("cuda", 0))
    losses = []
    for it in range(STEPS):
        opts = dict(recurrent_dropout=dropout, dropout_seed=99, dropout_step=it, row_offset=0) if dropout else {}
        losses.append(float(m.train_step_device(feat, gt, None, 1.0 / (B * P), **opts).item()))
        m.apply_gradients()
        if it == 0:
            want_first = m.param_buffer().cpu().numpy()
    # The first loss sees identical weights: equal to fp32 summation order.  Later ones follow two AMSGrad steps that move
    # EVERY weight by ~lr whatever its gradient's size, so a gradient element that cancels to rounding noise (the shard sum
    # and the full-batch gradient agree to 2e-7 relative L2 per tensor, 2e-6 for the vocabulary bias whose column sums are
    # atomics over many CTAs: tools/shard_grad_check.py) may flip its sign between the two runs and move that weight by
    # 2 lr: measured |d loss| up to 3e-6 after one step, 3.4e-4 after two.  A wrong normalisation or a lost bucket would
    # be O(1).
    np.testing.assert_allclose(got["losses"][:1], losses[:1], rtol=1e-5)
    np.testing.assert_allclose(got["losses"], losses, rtol=float(os.environ.get("DCAP_TEST_LOSS_RTOL", "2e-3")))
    assert losses[-1] < losses[0]
    want = m.param_buffer().cpu().numpy()
    # After ONE step the two runs have seen identical weights: their gradients differ by summation order only, and
    # Keras-AMSGrad's first update is lr * g / (|g| + 3.2e-6), so only elements whose gradient cancels to rounding noise
    # can move differently (by up to 2 lr): a quantile bar plus a hard cap on the update scale.
    d1 = np.abs(got["params_first"] - want_first)
    assert np.quantile(d1, 0.9999) <= 0.05 * 1e-3 and d1.max() <= 2.5 * 1e-3, (np.quantile(d1, 0.9999), d1.max())
    # Later steps: one such element is enough to decorrelate every bf16 rounding of the next forward pass, after which
    # all gradients differ at bf16 level (0.4 %) and the unsaturated updates of small-gradient weights (|g| ~ 3e-6, most
    # of the head's first kernel) drift apart smoothly -- measured median 6e-6, 99.99 % quantile 2.1e-4 after three
    # steps.  The bars below still catch what this test is for (a lost bucket, a wrong normalisation, replicas that
    # diverge): those move weights by O(lr) everywhere.
    step = 1e-3 * STEPS
    diff = np.abs(got["params"] - want)
    assert np.median(diff) <= 0.01 * step and np.quantile(diff, 0.9999) <= 0.25 * step and diff.max() <= 2.5 * step, \
        (np.median(diff), np.quantile(diff, 0.9999), diff.max())
