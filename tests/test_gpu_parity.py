"""Parity of the CUDA path (through the C ABI) against the CPU oracle, on a real GPU.
Integer stages: bit-exact.  Superpixel labels: >= 99.9 % per-pixel agreement (fp64 costs, CUDA log vs libm)."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import cart_slam_b200 as cb
import pyoracle as po
import ref_pipeline as rp
from cart_slam_b200.synth import SyntheticSequence

LABEL_AGREEMENT = 0.999  # tolerance stated by BASELINE.json north_star


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch.device("cuda:0")


SGM_CASES = [
    # W, H, D, min_disp, paths, n
    (200, 60, 64, 4, 4, 2),
    (333, 77, 128, 4, 4, 1),     # ragged sizes
    (160, 48, 64, 0, 8, 2),      # 8 paths, min_disp 0
    (300, 50, 256, 7, 4, 1),     # D = 256, odd min_disp
    (131, 37, 128, 4, 8, 1),     # image narrower than D + margins
]


@pytest.mark.parametrize("W,H,D,md,paths,n", SGM_CASES)
def test_sgm_stages_bit_exact(gpu, W, H, D, md, paths, n):
    seq = SyntheticSequence(W, H, D, min_disp=md, n_frames=4)
    L, R = seq.batch(1, n)
    rng = np.random.default_rng(W)
    L[0, 5:20, 10:40] = rng.integers(0, 256, (15, 30, 3))  # some real colour
    cfg = cb.Config(W, H, max_batch=n, num_disparities=D, min_disparity=md, paths=paths, enable_superpixels=False)
    with cb.Context(cfg) as ctx:
        ctx.sgm_gray_census(dev(L), dev(R))
        ctx.sgm_aggregate(n)
        disp = host(ctx.sgm_wta_post(n))
        cl, cr, gl = (host(ctx.sgm_intermediate(k, n)) for k in (0, 1, 2))
        vols = [host(ctx.sgm_intermediate(10 + p, n)) for p in range(paths)]
        wl, wr = host(ctx.sgm_intermediate(3, n)), host(ctx.sgm_intermediate(4, n))
        for f in range(n):
            o_disp, inter = po.sgm_compute(L[f], R[f], D, md, paths=paths, intermediates=True)
            assert np.array_equal(gl[f], po.gray(L[f])), "gray"
            assert np.array_equal(cl[f], inter["census_l"]), "census left"
            assert np.array_equal(cr[f], inter["census_r"]), "census right"
            for p in range(paths):
                bad = np.argwhere(vols[p][f] != inter["volumes"][p])
                assert bad.size == 0, f"path {p} volume differs first at {bad[:3]}"
            assert np.array_equal(wl[f], inter["left_raw"]), "WTA left"
            assert np.array_equal(wr[f], inter["right_raw"]), "WTA right"
            assert np.array_equal(disp[f], o_disp), "disparity"


def test_disparity_batch_equals_single_and_is_repeatable(gpu):
    W, H, D = 256, 64, 64
    seq = SyntheticSequence(W, H, D, n_frames=6)
    L, R = seq.batch(1, 5)
    cfg = cb.Config(W, H, max_batch=5, num_disparities=D, smoothing_radius=2, smoothing_iterations=1, enable_superpixels=False)
    with cb.Context(cfg) as ctx:
        a = host(ctx.disparity(dev(L), dev(R)))
        b = host(ctx.disparity(dev(L), dev(R)))
        assert np.array_equal(a, b)
        for f in range(5):
            one = host(ctx.disparity(dev(L[f:f + 1]), dev(R[f:f + 1])))
            assert np.array_equal(one[0], a[f])
        o = po.interpolate(po.sgm_compute(L[2], R[2], D), 2, 1, 64, W)
        assert np.array_equal(a[2], o)
        # sliced SGM (tuning aid CARTB200_SGM_SLICE: WTA of slice i beside the aggregation of slice i + 1 on a second
        # stream, frame windows into the context's scratch): same bits for every slice size, ragged last slice included
        try:
            for sl in ("1", "2", "3"):
                os.environ["CARTB200_SGM_SLICE"] = sl
                assert np.array_equal(host(ctx.disparity(dev(L), dev(R))), a), f"slice {sl}"
        finally:
            os.environ.pop("CARTB200_SGM_SLICE", None)


@pytest.mark.parametrize("radius,iters", [(2, 1), (3, 1), (3, 3), (4, 2)])
@pytest.mark.parametrize("W,H", [(200, 150), (129, 65), (64, 64)])
def test_interpolate(gpu, W, H, radius, iters):
    rng = np.random.default_rng(radius * 10 + iters)
    d = rng.integers(60, 900, (1, H, W)).astype(np.int16)
    d[0][rng.random((H, W)) < 0.3] = 48
    d[0][rng.random((H, W)) < 0.05] = -32768
    with cb.Context(cb.Config(W, H, num_disparities=64, enable_superpixels=False)) as ctx:
        got = host(ctx.interpolate(dev(d), radius, iters, 64, W))
    assert np.array_equal(got[0], po.interpolate(d[0], radius, iters, 64, W))


@pytest.mark.parametrize("W,H", [(300, 290), (128, 128), (131, 259), (50, 20)])
def test_derivative_and_naive(gpu, W, H):
    rng = np.random.default_rng(W + H)
    base = (np.arange(H)[:, None] * 3 + np.arange(W)[None, :] // 7).astype(np.int16) + 64
    d = np.stack([base + rng.integers(-3, 4, (H, W)).astype(np.int16) for _ in range(2)])
    d[:, rng.random((H, W)) < 0.1] = -32768
    with cb.Context(cb.Config(W, H, max_batch=2, num_disparities=64, enable_superpixels=False)) as ctx:
        deriv, hist = ctx.derivative(dev(d))
        nd, nh = ctx.naive_derivative(dev(d))
        planes = ctx.classify(nd, [[1, 30, -3, 1], [2, 9, -8, 2]])
        planes2 = ctx.classify(deriv, [[1, 30, -3, 1]])
        for f in range(2):
            o_d, o_h = po.derivative(d[f])
            assert np.array_equal(host(deriv)[f], o_d)
            assert np.array_equal(host(hist)[f], o_h)
            o_nd, o_nh = po.naive_derivative(d[f])
            assert np.array_equal(host(nd)[f], o_nd)
            assert np.array_equal(host(nh)[f], o_nh)
            pr = [1, 30, -3, 1] if f == 0 else [2, 9, -8, 2]
            assert np.array_equal(host(planes)[f], po.classify(o_nd, *pr))
            assert np.array_equal(host(planes2)[f], po.classify(o_d, 1, 30, -3, 1, channel_stride=2))


@pytest.mark.parametrize("W,H,block", [(150, 140, 12), (128, 128, 16), (70, 50, 8), (333, 97, 10)])
def test_border_map_and_block_init(gpu, W, H, block):
    with cb.Context(cb.Config(W, H, num_disparities=64, sp_block_size=block)) as ctx:
        n = ctx.superpixels_reset(1)
        lab0, n0 = po.block_init(W, H, block, block)
        assert n == n0
        rng = np.random.default_rng(0)
        lab = lab0.copy()
        ys, xs = rng.integers(0, H, 400), rng.integers(0, W, 400)
        lab[ys, xs] = rng.integers(0, n0, 400)
        got = host(ctx.superpixels_border_map(dev(lab)))
        assert np.array_equal(got, po.border_map(lab))


SP_CASES = [
    dict(W=192, H=96, block=12, its=8, kw={}),
    dict(W=170, H=70, block=10, its=5, kw=dict(w_compact=0.03, progressive=1.0)),   # kitti-superpixels.json
    dict(W=160, H=64, block=8, its=6, kw=dict(w_disp=0.0)),                          # disparity feature off
]


@pytest.mark.parametrize("case", SP_CASES, ids=lambda c: f"{c['W']}x{c['H']}b{c['block']}")
def test_superpixel_relax_agreement(gpu, case):
    """One frame at a time from the oracle's state (the chain is re-synchronised after every frame): bit equality.
    `sp_exact = False` is accepted and ignored (the approximate mode is gone) - the result must not depend on it."""
    W, H, block, its, kw = case["W"], case["H"], case["block"], case["its"], case["kw"]
    seq = SyntheticSequence(W, H, 64, n_frames=4, tint=True)
    cfg = cb.Config(W, H, max_batch=2, num_disparities=64, smoothing_radius=2, smoothing_iterations=1,
                    sp_block_size=block, sp_compactness_weight=kw.get("w_compact", 0.1),
                    sp_progressive_compactness_cost=kw.get("progressive", 0.0),
                    sp_disparity_weight=kw.get("w_disp", 1.0), sp_exact=False)
    with cb.Context(cfg) as ctx:
        lab_o, nlab = po.block_init(W, H, block, block)
        for fid in (1, 2, 3):
            l, r, _ = seq.frame(fid)
            disp = ctx.disparity(dev(l[None]), dev(r[None]))
            deriv, _ = ctx.derivative(disp)
            use_d = kw.get("w_disp", 1.0) > 0
            got = host(ctx.superpixels_relax(dev(l[None]), deriv if use_d else None, its, slots=[1]))[0]
            lab_o, bc, mv = po.sp_relax(lab_o, nlab, po.ycrcb(l), host(deriv)[0] if use_d else None, its, **kw)
            assert np.array_equal(got, lab_o), (fid, float((got == lab_o).mean()))
            assert mv.sum() > 0
            ctx.superpixels_set_labels(1, dev(lab_o))  # exercises set_labels between frames (a no-op for the values)


@pytest.mark.parametrize("case", SP_CASES, ids=lambda c: f"{c['W']}x{c['H']}b{c['block']}")
def test_superpixel_exact_mode_is_bit_identical_over_a_chain(gpu, case):
    """Label costs in the reference's operation order with the fully specified logarithm - the labels
    equal the oracle's bit for bit, frame after frame of a warm-started chain (no re-synchronisation of the state)."""
    W, H, block, its, kw = case["W"], case["H"], case["block"], case["its"], case["kw"]
    n_frames = 12
    seq = SyntheticSequence(W, H, 64, n_frames=n_frames + 1, tint=True)
    cfg = cb.Config(W, H, max_batch=1, num_disparities=64, smoothing_radius=2, smoothing_iterations=1,
                    sp_block_size=block, sp_compactness_weight=kw.get("w_compact", 0.1),
                    sp_progressive_compactness_cost=kw.get("progressive", 0.0),
                    sp_disparity_weight=kw.get("w_disp", 1.0), sp_exact=True)
    with cb.Context(cfg) as ctx:
        lab_o, nlab = po.block_init(W, H, block, block)
        ctx.superpixels_reset(1)
        for fid in range(1, n_frames + 1):
            l, r, _ = seq.frame(fid)
            disp = ctx.disparity(dev(l[None]), dev(r[None]))
            deriv, _ = ctx.derivative(disp)
            use_d = kw.get("w_disp", 1.0) > 0
            got = host(ctx.superpixels_relax(dev(l[None]), deriv if use_d else None, its))[0]
            lab_o, _, mv = po.sp_relax(lab_o, nlab, po.ycrcb(l), host(deriv)[0] if use_d else None, its, **kw)
            assert np.array_equal(got, lab_o), (fid, float((got == lab_o).mean()))
            assert mv.sum() > 0


def test_sp_planeseg_exact(gpu):
    W, H = 200, 120
    rng = np.random.default_rng(4)
    lab, n = po.block_init(W, H, 12, 12)
    lab = np.stack([lab, np.roll(lab, 5, axis=1)])
    deriv = rng.integers(-40, 40, (2, H, W, 2)).astype(np.int16)
    deriv[rng.random((2, H, W)) < 0.2] = -32768
    with cb.Context(cb.Config(W, H, max_batch=2, num_disparities=64)) as ctx:
        unsm, planes = ctx.sp_planeseg(dev(deriv), dev(lab), [[1, 30, -3, 1], [0, 5, -20, 0]])
        for f, pr in enumerate([[1, 30, -3, 1], [0, 5, -20, 0]]):
            ou, op = po.sp_planeseg(deriv[f], lab[f], n, *pr)
            assert np.array_equal(host(unsm)[f], ou) and np.array_equal(host(planes)[f], op)
    # the reference refuses more than 5461 superpixels (sp_planeseg.cu:327-331)
    with cb.Context(cb.Config(640, 480, num_disparities=64, sp_block_size=6)) as ctx:
        with pytest.raises(cb.CartB200Error) as e:
            ctx.sp_planeseg(dev(np.zeros((1, 480, 640, 2), np.int16)), dev(np.zeros((1, 480, 640), np.uint16)), [[0, 0, 0, 0]])
        assert e.value.code == cb.E_UNSUPPORTED


def _frames(W, H, D, n, tint=False):
    seq = SyntheticSequence(W, H, D, n_frames=n, tint=tint)
    return seq, [seq.frame(i + 1)[:2] for i in range(n)]


@pytest.mark.parametrize("provider", ["static", "histogram_peak"])
def test_sequence_naive_pipeline(gpu, provider):
    W, H, D, n = 192, 96, 64, 23
    seq, frames = _frames(W, H, D, n)
    cfgd = dict(D=D, radius=2, iters=1)
    ref = rp.naive_sequence(frames, cfgd, provider=provider, update=5, reset=2)
    L = np.stack([f[0] for f in frames])
    R = np.stack([f[1] for f in frames])
    cfg = cb.Config(W, H, max_batch=4, num_disparities=D, smoothing_radius=2, smoothing_iterations=1, enable_superpixels=False)
    opts = cb.SequenceOptions(pipeline=0, provider=0 if provider == "static" else 1, update_interval=5, reset_interval=2)
    with cb.Context(cfg) as ctx:
        planes, disp = ctx.run_sequence_host(opts, L, R, want_disparity=True)
        pd = host(ctx.run_sequence_device(opts, dev(L), dev(R)))
    for i in range(n):
        assert np.array_equal(disp[i], ref[i]["disparity"]), i
        assert np.array_equal(planes[i], ref[i]["planes"]), (i, ref[i]["params"])
    assert np.array_equal(pd, planes)
    if provider == "histogram_peak":
        assert any(r["params"][2:] != [0, 0, 0, 0] for r in ref), "peak provider never fired on the test data"


@pytest.mark.parametrize("max_batch", [2, 8])
@pytest.mark.parametrize("provider", ["static", "histogram_peak"])
def test_sequence_superpixel_pipeline(gpu, provider, max_batch):
    # reset every 8 frames so that 21 frames span chunks [1..7], [8..15], [16..21]; 2 slots force two groups,
    # 8 slots put all three chunks in one group whose SGM/derivative stages run two steps per launch
    W, H, D, n = 160, 64, 64, 21
    seq, frames = _frames(W, H, D, n, tint=True)
    cfgd = dict(D=D, radius=2, iters=1)
    ref = rp.sp_sequence(frames, cfgd, provider=provider, update=5, reset=2, initial=6, steady=3, sp_reset=8, block=8)
    L = np.stack([f[0] for f in frames])
    R = np.stack([f[1] for f in frames])
    cfg = cb.Config(W, H, max_batch=max_batch, num_disparities=D, smoothing_radius=2, smoothing_iterations=1, sp_block_size=8)
    opts = cb.SequenceOptions(pipeline=1, provider=0 if provider == "static" else 1, update_interval=5, reset_interval=2,
                              sp_initial_iterations=6, sp_iterations=3, sp_reset_iterations=8)
    with cb.Context(cfg) as ctx:
        planes, disp = ctx.run_sequence_host(opts, L, R, want_disparity=True)
        pd = host(ctx.run_sequence_device(opts, dev(L), dev(R)))
    agree = []
    for i in range(n):
        assert np.array_equal(disp[i], ref[i]["disparity"]), i
        agree.append((planes[i] == ref[i]["planes"]).mean())
    # superpixel labels are bit-identical to the oracle over the whole warm-started chain, so are the planes
    assert min(agree) == 1.0, agree
    assert np.array_equal(pd, planes)


@pytest.mark.parametrize("provider", ["static", "histogram_peak"])
def test_sequence_superpixel_chunk_groups(gpu, provider):
    """Eight lock-stepped chunks (reset every 4 frames, 30 frames) in one slot group: the runner advances them as two
    groups on two streams by default (CARTB200_SP_SPLIT groups; disjoint scratch ranges).  Planes and disparities equal
    the oracle pipeline bit for bit, and 1 / 2 / 3 / 4 groups give the same bytes."""
    W, H, D, n = 160, 64, 64, 30
    seq, frames = _frames(W, H, D, n, tint=True)
    cfgd = dict(D=D, radius=2, iters=1)
    ref = rp.sp_sequence(frames, cfgd, provider=provider, update=5, reset=2, initial=5, steady=2, sp_reset=4, block=8)
    L = np.stack([f[0] for f in frames])
    R = np.stack([f[1] for f in frames])
    cfg = cb.Config(W, H, max_batch=16, num_disparities=D, smoothing_radius=2, smoothing_iterations=1, sp_block_size=8)
    opts = cb.SequenceOptions(pipeline=1, provider=0 if provider == "static" else 1, update_interval=5, reset_interval=2,
                              sp_initial_iterations=5, sp_iterations=2, sp_reset_iterations=4)
    got = {}
    try:
        with cb.Context(cfg) as ctx:
            for g in ("2", "1", "3", "4"):
                os.environ["CARTB200_SP_SPLIT"] = g
                planes, disp = ctx.run_sequence_host(opts, L, R, want_disparity=True)
                got[g] = (planes.copy(), disp.copy(), host(ctx.run_sequence_device(opts, dev(L), dev(R))))
    finally:
        os.environ.pop("CARTB200_SP_SPLIT", None)
    planes, disp, pd = got["2"]
    for i in range(n):
        assert np.array_equal(disp[i], ref[i]["disparity"]), i
        assert np.array_equal(planes[i], ref[i]["planes"]), i
    for g, (p2, d2, pd2) in got.items():
        assert np.array_equal(p2, planes) and np.array_equal(d2, disp) and np.array_equal(pd2, planes), f"{g} groups"


def test_full_size_kitti_properties(gpu):
    """BASELINE.json full size (1242x375, D=128): size-independent properties instead of the slow oracle."""
    W, H, D = 1242, 375, 128
    seq = SyntheticSequence(W, H, D, n_frames=4)
    L, R = seq.batch(1, 3)
    gts = [seq.frame(i)[2] for i in (1, 2, 3)]
    cfg = cb.Config(W, H, max_batch=3, num_disparities=D, smoothing_radius=2, smoothing_iterations=1)
    with cb.Context(cfg) as ctx:
        n = ctx.sgm_gray_census(dev(L), dev(R))
        ctx.sgm_aggregate(n)
        disp = ctx.sgm_wta_post(n)
        vols = [ctx.sgm_intermediate(10 + p, n) for p in range(4)]
        cl = host(ctx.sgm_intermediate(0, n))
        # path volumes: L >= C on the first pixel of every path means equality; bounded by 31 + P2
        for v in vols:
            assert int(v.max()) <= 31 + 120
        # L->R path at x = 0 and T->B path at y = 0 equal the raw matching cost there: census border is 0
        assert int(vols[2][:, 0].max()) == 0 and int(vols[3][:, H - 1].max()) == 0
        # left/right symmetric checksum: sum over d of the horizontal volumes is direction independent at row ends
        d = host(disp)
        for f in range(3):
            valid = d[f] != 48
            valid &= d[f] != -32768
            err = np.abs(d[f].astype(np.float32) / 16 - gts[f])
            assert valid.mean() > 0.6 and (err[valid] <= 1).mean() > 0.9, (f, valid.mean(), (err[valid] <= 1).mean())
        deriv, hist = ctx.derivative(disp)
        hv = host(hist)
        dv = host(deriv)
        for f in range(3):
            for ch in range(2):
                ok = (dv[f, :, :, ch] != -32768) & (dv[f, :, :, ch] >= -128) & (dv[f, :, :, ch] <= 127)
                assert hv[f, :, ch].sum() == ok.sum()            # histogram mass = in-range valid pixels
                assert np.array_equal(np.bincount(dv[f, :, :, ch][ok].astype(np.int64) + 128, minlength=256), hv[f, :, ch])
        # one frame end-to-end against the oracle at full size (a few seconds of CPU)
        o = po.interpolate(po.sgm_compute(L[1], R[1], D), 2, 1, 64, W)
        assert np.array_equal(d[1], o)
        labels = ctx.superpixels_relax(dev(L[:1]), deriv[:1], 8)
        lab0, nlab = po.block_init(W, H, 12, 12)
        o_lab, _, _ = po.sp_relax(lab0, nlab, po.ycrcb(L[0]), dv[0], 8)
        assert np.array_equal(host(labels)[0], o_lab)  # full-size frame: bit-identical superpixel labels


def test_depth_matches_oracle(gpu):
    """DepthModule replacement: single-precision arithmetic in the oracle's order (no FMA contraction) - tolerance 1e-6
    relative, and in practice bit-identical."""
    W, H = 333, 97
    rng = np.random.default_rng(3)
    disp = rng.integers(-40, 120 * 16, (2, H, W)).astype(np.int16)
    disp[0, 5:9, 7:30] = -32768  # invalid marker: no special handling, like the reference's GPU path
    Q = np.eye(4, dtype=np.float32)
    Q[0, 3], Q[1, 3], Q[2, 2], Q[2, 3], Q[3, 2], Q[3, 3] = -166.0, -48.5, 0, 718.9, -1 / 0.537, 0.25
    cfg = cb.Config(W, H, max_batch=2, enable_sgm=False, enable_superpixels=False)
    with cb.Context(cfg) as ctx:
        xyz = host(ctx.depth(dev(disp), Q))
    for f in range(2):
        o = po.depth(disp[f], Q)
        fin = np.isfinite(o)
        assert np.array_equal(np.isfinite(xyz[f]), fin)
        assert np.allclose(xyz[f][fin], o[fin], rtol=1e-6, atol=0)


def _check_sgm_full_size(gpu, W, H, D, paths, rows_sampled, seed):
    """BASELINE.json full sizes where the scalar oracle is too slow for whole volumes: every stage is still checked
    exactly, through properties that do not need the oracle's full run -
      census       = oracle census of the oracle gray image (cheap, whole image);
      path volumes = the recurrence D4 re-evaluated in numpy on sampled pixels FROM THE GPU'S OWN predecessor
                     vectors and the census words (self-consistency of every path at full size);
      raw WTA      = oracle WTA (D5, D6) on sampled full rows of the GPU's volumes (rows are independent);
      disparity    = oracle medians + L/R check + range correction (D7-D9) on the GPU's raw WTA images (whole image)."""
    md, P1, P2 = 4, 10, 120
    seq = SyntheticSequence(W, H, D, min_disp=md, n_frames=2)
    l, r, _ = seq.frame(1)
    cfg = cb.Config(W, H, max_batch=1, num_disparities=D, min_disparity=md, paths=paths, enable_superpixels=False)
    rng = np.random.default_rng(seed)
    with cb.Context(cfg) as ctx:
        ctx.sgm_gray_census(dev(l[None]), dev(r[None]))
        ctx.sgm_aggregate(1)
        disp = host(ctx.sgm_wta_post(1))[0]
        gl = host(ctx.sgm_intermediate(2, 1))[0]
        cl, cr = host(ctx.sgm_intermediate(0, 1))[0], host(ctx.sgm_intermediate(1, 1))[0]
        assert np.array_equal(gl, po.gray(l))
        assert np.array_equal(cl, po.census(po.gray(l))) and np.array_equal(cr, po.census(po.gray(r)))
        wl, wr = host(ctx.sgm_intermediate(3, 1))[0], host(ctx.sgm_intermediate(4, 1))[0]
        ys = np.sort(rng.choice(H, rows_sampled, replace=False))
        row_vols = []
        dirs = po.sgm_dirs(paths)
        d = np.arange(D)
        for p in range(paths):
            vol = ctx.sgm_intermediate(10 + p, 1)[0]  # [H, W, D] on the device
            assert int(vol.max()) <= 31 + P2
            row_vols.append(vol[torch.from_numpy(ys).to(vol.device)].cpu().numpy())
            dx, dy = int(dirs[p][0]), int(dirs[p][1])
            # recurrence on sampled pixels (vectorised over the samples and over d)
            n = 4000
            sx, sy = rng.integers(0, W, n), rng.integers(0, H, n)
            px, py = sx - dx, sy - dy
            inside = (px >= 0) & (px < W) & (py >= 0) & (py < H)
            cur = vol[torch.from_numpy(sy).to(vol.device), torch.from_numpy(sx).to(vol.device)].cpu().numpy().astype(np.int64)
            prev = vol[torch.from_numpy(np.clip(py, 0, H - 1)).to(vol.device),
                       torch.from_numpy(np.clip(px, 0, W - 1)).to(vol.device)].cpu().numpy().astype(np.int64)
            prev[~inside] = 0  # a path starts from the all-zero state where it enters the image
            xr = sx[:, None] - d[None, :] - md
            rword = np.where((xr >= 0) & (xr < W), cr[sy[:, None], np.clip(xr, 0, W - 1)], 0).astype(np.uint32)
            x = (cl[sy, sx][:, None] ^ rword).astype(np.uint32)
            cost = np.zeros(x.shape, np.int64)
            for b in range(32):
                cost += (x >> np.uint32(b)) & np.uint32(1)
            m = prev.min(1, keepdims=True)
            big = 1 << 20
            lo = np.concatenate([np.full((n, 1), big), prev[:, :-1] + P1], 1)
            hi = np.concatenate([prev[:, 1:] + P1, np.full((n, 1), big)], 1)
            exp = cost + np.minimum(np.minimum(prev, m + P2), np.minimum(lo, hi)) - m
            bad = np.argwhere(exp != cur)
            assert bad.size == 0, (p, bad[:3], sx[bad[:3, 0]], sy[bad[:3, 0]])
            del vol
        for i, y in enumerate(ys):
            o_l, o_r = po.sgm_wta([rv[i][None] for rv in row_vols])
            assert np.array_equal(wl[y], o_l[0]), ("WTA left row", y)
            assert np.array_equal(wr[y], o_r[0]), ("WTA right row", y)
        o_disp = po.lr_check_range(po.median3(wl), po.median3(wr), gl, md)
        assert np.array_equal(disp, o_disp)
        # determinism: a second run gives identical bits
        ctx.sgm_gray_census(dev(l[None]), dev(r[None]))
        ctx.sgm_aggregate(1)
        assert np.array_equal(host(ctx.sgm_wta_post(1))[0], disp)


def test_full_size_zed_properties(gpu):
    """BASELINE.json configs[2]: 1280x720, 256 disparities (4 paths)."""
    _check_sgm_full_size(gpu, 1280, 720, 256, 4, rows_sampled=6, seed=11)


def test_full_size_4k_8path_properties(gpu):
    """BASELINE.json configs[3]: 3840x2160, 256 disparities, 8-path aggregation."""
    _check_sgm_full_size(gpu, 3840, 2160, 256, 8, rows_sampled=4, seed=12)


def test_c_abi_error_paths(gpu):
    """Error behaviour at the boundary: return codes + messages, nothing throws inside the library, nothing exits
    (the reference logs and exit()s on CUDA errors and throws std::runtime_error on type guards)."""
    W, H = 96, 48
    with pytest.raises(cb.CartB200Error, match="num_disparities must be 64, 128 or 256"):
        cb.Context(cb.Config(W, H, num_disparities=100))
    with pytest.raises(cb.CartB200Error, match="paths must be 4"):
        cb.Context(cb.Config(W, H, num_disparities=64, paths=5))
    with pytest.raises(cb.CartB200Error, match="31 \\+ p2 <= 255"):
        cb.Context(cb.Config(W, H, num_disparities=64, p2=240))
    with pytest.raises(cb.CartB200Error, match="below 16384"):
        cb.Context(cb.Config(2048, 1024, num_disparities=64, enable_sgm=False, sp_block_size=8))
    cfg = cb.Config(W, H, max_batch=2, num_disparities=64, enable_superpixels=False)
    with cb.Context(cfg) as ctx:
        L = torch.zeros((3, H, W, 3), dtype=torch.uint8, device="cuda")
        with pytest.raises(cb.CartB200Error, match="batch size 3 outside 1..2"):
            ctx.disparity(L, L)
        with pytest.raises(cb.CartB200Error, match="path index out of range"):
            ctx.sgm_aggregate_path(1, 4)
        d = torch.zeros((1, H, W), dtype=torch.int16, device="cuda")
        with pytest.raises(cb.CartB200Error, match="without superpixels"):
            ctx.superpixels_reset(1)
        with pytest.raises(cb.CartB200Error, match="superpixel pipeline needs enable_superpixels"):
            ctx.run_sequence_device(cb.SequenceOptions(pipeline=1), L[:2], L[:2])
        # the context stays usable after an error
        out = ctx.disparity(L[:2], L[:2])
        assert out.shape == (2, H, W)
    sgm_off = cb.Config(W, H, max_batch=1, enable_sgm=False, enable_superpixels=False)
    with cb.Context(sgm_off) as ctx:
        L = torch.zeros((1, H, W, 3), dtype=torch.uint8, device="cuda")
        with pytest.raises(cb.CartB200Error, match="enable_sgm = 0"):
            ctx.disparity(L, L)
