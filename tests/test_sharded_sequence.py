"""One sequence sharded over several ranks (BASELINE.json configs[4]) and the bench's headline shape.

CPU: cartb200_sequence_parameters against the oracle-side restatement of the reference's bookkeeping
     (planeseg.cu:379-403, sp_planeseg.cu:352-388).
GPU: shards from plan_shards run through phase 1 / parameter schedule / phase 2 (one context per shard, as one rank per
     GPU would) equal the unsharded runner bit for bit - both pipelines, both providers, start_id != 1 included;
     the headline configuration (1242x375, 128 disparities, batch 64, three chunks of the 64-frame reset period,
     histogram_peak) equals the composition of the per-stage C-ABI calls on every frame and the oracle on sampled
     frames."""
import numpy as np
import pytest

import cart_slam_b200 as cb
import pyoracle as po
from cart_slam_b200.parallel import plan_shards
from cart_slam_b200.synth import SyntheticSequence


def _schedule_oracle(pipeline, provider, static, update, reset, start_id, hists):
    """ref_pipeline.py's bookkeeping, histograms given."""
    running = None if pipeline == 1 else np.zeros(256, np.int64)
    params = [0, 0, 0, 0, 0, 0] if provider == 1 else [0, 0] + list(static)
    out = []
    for i, hv in enumerate(hists):
        fid = start_id + i
        hv = hv.astype(np.int64)
        if pipeline == 0:
            running += hv
            if provider == 1 and fid % update == 1:
                snap = running.astype(np.int32).copy()
                if fid % (update * reset) == 1:
                    running[:] = 0
                _, params = po.histogram_peak_update(snap, params)
        else:
            if running is None:
                running = np.zeros(256, np.int64)
                hist = hv.copy()
            else:
                running += hv
                hist = running.copy()
            if fid % (update * reset) == 1:
                running[:] = 0
            if provider == 1 and fid % update == 1:
                _, params = po.histogram_peak_update(hist.astype(np.int32), params)
        out.append(params[2:6])
    return np.array(out, np.int32).reshape(len(hists), 4)


@pytest.mark.parametrize("pipeline", [0, 1])
@pytest.mark.parametrize("provider", [0, 1])
@pytest.mark.parametrize("start_id", [1, 7, 64])
def test_sequence_parameters_match_the_reference_bookkeeping(pipeline, provider, start_id):
    rng = np.random.default_rng(pipeline * 10 + provider + start_id)
    n = 70
    # derivative-like histograms: a dominant near-zero peak plus a moving secondary peak, so that findPeaks fires
    x = np.arange(256)
    hists = []
    for i in range(n):
        h = 4000 * np.exp(-0.5 * ((x - 128) / 2.0) ** 2) + 900 * np.exp(-0.5 * ((x - (140 + (i % 9))) / 3.0) ** 2)
        hists.append((h + rng.integers(0, 30, 256)).astype(np.int32))
    hists = np.stack(hists)
    opts = cb.SequenceOptions(pipeline=pipeline, provider=provider, static_params=(1, 30, -3, 1), update_interval=5,
                              reset_interval=3, start_id=start_id)
    got = cb.sequence_parameters(opts, hists)
    want = _schedule_oracle(pipeline, provider, (1, 30, -3, 1), 5, 3, start_id, hists)
    assert np.array_equal(got, want)
    if provider == 1:
        assert len({tuple(r) for r in got.tolist()}) > 1, "the peak provider never changed the ranges on the test data"


@pytest.mark.parametrize("pipeline", [0, 1])
def test_sequence_parameters_long_table_threaded_path(pipeline):
    """From 2048 frames on, the table takes the snapshots first, runs the peak updates on a few threads and applies them in
    frame order (the 10,000-frame sharded run computes it on every rank): same numbers as the frame-by-frame oracle,
    failing updates (flat histograms, which keep the previous ranges) included."""
    rng = np.random.default_rng(99 + pipeline)
    n = 2600
    x = np.arange(256)
    hists = np.empty((n, 256), np.int32)
    for i in range(n):
        if rng.random() < 0.15:
            hists[i] = rng.integers(0, 2, 256)  # (almost) nothing: whole update windows of these make the update fail
        else:
            a, b = 128 + rng.integers(-3, 4), 140 + rng.integers(0, 25)
            h = 4000 * np.exp(-0.5 * ((x - a) / 2.0) ** 2) + 900 * np.exp(-0.5 * ((x - b) / 3.0) ** 2)
            hists[i] = (h + rng.integers(0, 30, 256)).astype(np.int32)
    hists[300:420] = 0  # several consecutive update windows without any data
    for start_id in (1, 64):
        opts = cb.SequenceOptions(pipeline=pipeline, provider=1, static_params=(1, 30, -3, 1), update_interval=7,
                                  reset_interval=4, start_id=start_id)
        got = cb.sequence_parameters(opts, hists)
        want = _schedule_oracle(pipeline, 1, (1, 30, -3, 1), 7, 4, start_id, hists)
        assert np.array_equal(got, want), start_id
        assert np.array_equal(got[:2000], cb.sequence_parameters(opts, hists[:2000]))  # the sequential path on the prefix


# ---- GPU -----------------------------------------------------------------------------------------------------------
torch = pytest.importorskip("torch")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch.device("cuda:0")


def _frames(W, H, D, n, tint=True):
    seq = SyntheticSequence(W, H, D, n_frames=n, tint=tint)
    fr = [seq.frame(i + 1)[:2] for i in range(n)]
    return np.stack([f[0] for f in fr]), np.stack([f[1] for f in fr])


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("provider", [0, 1])
@pytest.mark.parametrize("pipeline", [0, 1])
def test_sharded_run_equals_unsharded(gpu, pipeline, provider, world):
    W, H, D, n, reset = 160, 64, 64, 37, 8
    L, R = _frames(W, H, D, n)
    cfg = cb.Config(W, H, max_batch=4, num_disparities=D, smoothing_radius=2, smoothing_iterations=1, sp_block_size=8,
                    enable_superpixels=pipeline == 1)

    def opts(start_id):
        return cb.SequenceOptions(pipeline=pipeline, provider=provider, static_params=(1, 30, -3, 1), update_interval=5,
                                  reset_interval=2, sp_initial_iterations=6, sp_iterations=3, sp_reset_iterations=reset,
                                  start_id=start_id)
    with cb.Context(cfg) as ctx:
        whole = host(ctx.run_sequence_device(opts(1), dev(L), dev(R)))
        # the two-phase path on the whole sequence is the same computation
        hist = ctx.run_sequence_phase1(opts(1), dev(L), dev(R))
        two_phase = host(ctx.run_sequence_phase2(opts(1), cb.sequence_parameters(opts(1), hist)))
        hist_h = ctx.run_sequence_phase1_host(opts(1), L, R)  # the same with host buffers (the bench's e2e arm)
        two_phase_h = ctx.run_sequence_phase2_host(opts(1), cb.sequence_parameters(opts(1), hist_h))
    assert np.array_equal(two_phase, whole)
    assert np.array_equal(hist_h, hist) and np.array_equal(two_phase_h, whole)
    shards = plan_shards(n, world, reset)
    assert sum(s.count for s in shards) == n and any(s.first_id != 1 and s.count for s in shards)
    ctxs, hists = [], []
    try:
        for s in shards:  # one context per shard, like one rank per GPU
            if s.count == 0:
                ctxs.append(None)
                hists.append(np.zeros((0, 256), np.int32))
                continue
            c = cb.Context(cfg)
            ctxs.append(c)
            hists.append(c.run_sequence_phase1(opts(s.first_id), dev(L[s.frame_slice]), dev(R[s.frame_slice])))
        # "all-gather" + the schedule over the whole sequence (every rank computes the same table)
        params = cb.sequence_parameters(opts(1), np.concatenate(hists))
        assert np.array_equal(np.concatenate(hists), hist)
        parts = [host(c.run_sequence_phase2(opts(s.first_id), params[s.frame_slice])) for c, s in zip(ctxs, shards) if c]
    finally:
        for c in ctxs:
            if c:
                c.close()
    assert np.array_equal(np.concatenate(parts), whole)
    # a shard that starts in the middle of a chunk is refused for the superpixel pipeline, not silently wrong
    if pipeline == 1:
        with cb.Context(cfg) as ctx, pytest.raises(cb.CartB200Error, match="reset frame"):
            ctx.run_sequence_device(opts(5), dev(L[:8]), dev(R[:8]))


@pytest.mark.gpu
def test_phase2_needs_its_phase1(gpu):
    W, H, D = 160, 64, 64
    L, R = _frames(W, H, D, 6)
    cfg = cb.Config(W, H, max_batch=4, num_disparities=D, sp_block_size=8)
    o = cb.SequenceOptions(pipeline=1, provider=1, sp_reset_iterations=8)
    with cb.Context(cfg) as ctx:
        with pytest.raises(cb.CartB200Error, match="phase 1"):
            ctx.run_sequence_phase2(o, np.zeros((6, 4), np.int32))
        ctx.run_sequence_phase1(o, dev(L), dev(R))
        with pytest.raises(cb.CartB200Error, match="phase 1"):
            ctx.run_sequence_phase2(o, np.zeros((5, 4), np.int32))
        ctx.run_sequence_phase2(o, np.zeros((6, 4), np.int32))


@pytest.mark.gpu
@pytest.mark.parametrize("n,reset", [(134, 64), (70, 16)])
def test_headline_shape_sequence_equals_per_stage_composition(gpu, n, reset):
    """BENCH's shape: 1242x375, 128 disparities, batch 64, reset 64, 24 / 8 iterations, histogram_peak; 134 frames =
    chunks [1..63], [64..127], [128..134] advanced in lock step (and 70 frames with reset 16: five chunks, which the
    runner advances as two groups on two streams).  Every frame's planes equal the composition of the
    per-stage entry points (each oracle-checked in test_gpu_parity.py) driven by the reference's schedule; three
    sampled frames are also checked against the oracle itself."""
    W, H, D = 1242, 375, 128
    seq = SyntheticSequence(W, H, D, n_frames=8, tint=True)
    base = [seq.frame(1 + i)[:2] for i in range(8)]
    rng = np.random.default_rng(5)
    L = np.empty((n, H, W, 3), np.uint8)
    R = np.empty((n, H, W, 3), np.uint8)
    for i in range(n):  # distinct frames: a base pair with a few perturbed rows (keeps host-side generation short)
        l, r = base[i % 8]
        L[i], R[i] = l, r
        rows = rng.integers(0, H, 6)
        L[i, rows] = np.roll(l[rows], i % 5, axis=1)
    cfg = cb.Config(W, H, max_batch=64, num_disparities=D, smoothing_radius=2, smoothing_iterations=1, sp_block_size=12)
    opts = cb.SequenceOptions(pipeline=1, provider=1, static_params=(1, 30, -3, 1), sp_initial_iterations=24,
                              sp_iterations=8, sp_reset_iterations=reset)
    dL, dR = dev(L), dev(R)
    with cb.Context(cfg) as ctx:
        planes, disp = ctx.run_sequence_device(opts, dL, dR, want_disparity=True)
        planes, disp = host(planes), host(disp)
    sample = {1: None, reset: None, n - 4: None}
    cfg1 = cb.Config(W, H, max_batch=1, num_disparities=D, smoothing_radius=2, smoothing_iterations=1, sp_block_size=12)
    hists = np.zeros((n, 256), np.int32)
    keep = []
    with cb.Context(cfg1) as ctx:
        ctx.superpixels_reset(1)
        for i in range(n):
            fid = i + 1
            d = ctx.disparity(dL[i:i + 1], dR[i:i + 1])
            assert np.array_equal(host(d)[0], disp[i]), fid
            deriv, hist = ctx.derivative(d)
            if fid % reset == 0:
                ctx.superpixels_reset(1)
            its = 24 if (fid == 1 or fid % reset == 0) else 8
            if fid in sample:
                before = host(ctx.superpixels_relax(dL[i:i + 1], deriv, 0))[0] if fid % reset else None
            labels = ctx.superpixels_relax(dL[i:i + 1], deriv, its)
            if fid in sample:
                sample[fid] = (before, host(deriv)[0], host(labels)[0], its)
            hists[i] = host(hist)[0][:, 0]
            keep.append((deriv, labels))
        params = cb.sequence_parameters(opts, hists)
        assert len({tuple(p) for p in params.tolist()}) > 1, "the peak provider never fired"
        for i, (deriv, labels) in enumerate(keep):
            _, pl = ctx.sp_planeseg(deriv, labels, [list(map(int, params[i]))])
            assert np.array_equal(host(pl)[0], planes[i]), i + 1
    # the oracle on the sampled frames: disparity, derivative, one warm-started relaxation, vote
    for fid, (before, deriv, labels, its) in sample.items():
        i = fid - 1
        o_disp = po.interpolate(po.sgm_compute(L[i], R[i], D), 2, 1, 64, W)
        assert np.array_equal(o_disp, disp[i]), fid
        o_deriv, _ = po.derivative(o_disp)
        assert np.array_equal(o_deriv, deriv), fid
        lab0, nlab = po.block_init(W, H, 12, 12)
        start = lab0 if before is None else before
        o_lab, _, _ = po.sp_relax(start.copy(), nlab, po.ycrcb(L[i]), o_deriv, its)
        assert np.array_equal(o_lab, labels), fid
        _, o_planes = po.sp_planeseg(o_deriv, o_lab, nlab, *[int(v) for v in params[i]])
        assert np.array_equal(o_planes, planes[i]), fid
