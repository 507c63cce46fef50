"""SURVEY 8(f) rows f3 (temporal smoothing vote) and f4 (plane-fit superpixel consumers).

CPU: the oracle against an independent vectorised numpy restatement of the reference loops, and against the outputs of
the REFERENCE'S OWN kernels (tests/golden/ref_kernels_f34.npz, generated on a B200 by tests/golden/make_golden_f34.py).
GPU (-m gpu): the CUDA path through the C ABI against the oracle, bit-exact, on the golden inputs and at KITTI size."""
import os
import sys

import numpy as np
import pytest

import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import f34_inputs  # noqa: E402

BASE = os.path.join(HERE, "golden", "ref_kernels.npz")
GOLD = os.path.join(HERE, "golden", "ref_kernels_f34.npz")


@pytest.fixture(scope="module")
def inp():
    if not os.path.exists(BASE):
        pytest.skip("golden vectors missing")
    return f34_inputs.make(np.load(BASE))


@pytest.fixture(scope="module")
def gold(inp):
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/ref_kernels_f34.npz not generated yet (needs a GPU box)")
    g = np.load(GOLD)
    assert int(g["crc"][0]) == inp["crc"], "golden inputs drifted: regenerate with tests/golden/make_golden_f34.py"
    return g


def np_temporal(plane, prev_planes, prev_flow, mode):
    """Vectorised restatement of planeseg.cu:199-240 (mode 0) / sp_planeseg.cu:79-117 (mode 1)."""
    H, W = plane.shape
    votes = np.zeros((3, H, W), np.int32)
    ys, xs = np.mgrid[0:H, 0:W]
    np.add.at(votes, (plane, ys, xs), 2 if mode else 1)
    x, y = xs.copy(), ys.copy()
    for pl, fl in zip(prev_planes, prev_flow):
        x = x - (fl[:, :, 0].astype(np.int32) >> 5)
        y = y - (fl[:, :, 1].astype(np.int32) >> 5)
        ok = (x >= 0) & (y >= 0) & (x < W) & (y < H)
        v = pl[np.clip(y, 0, H - 1), np.clip(x, 0, W - 1)]
        np.add.at(votes, (v[ok], ys[ok], xs[ok]), 1)
    best = np.where(votes[0] > votes[1], 0, 1)
    vb = np.maximum(votes[0], votes[1])
    if mode == 0:
        return np.where(vb == 0, 2, best).astype(np.uint8)
    return np.where(vb < votes[2], 2, best).astype(np.uint8)


def np_sp_assign(labels, per_pixel, n_labels):
    votes = np.zeros((n_labels, 3), np.int64)
    np.add.at(votes, (labels.ravel(), per_pixel.ravel()), 1)
    best = np.full(n_labels, 2)
    mx = votes[:, 2].copy()
    v = votes[:, 1] > mx
    best[v] = 1
    mx[v] = votes[v, 1]
    best[votes[:, 0] > mx] = 0
    return best[labels].astype(np.uint8)


@pytest.mark.parametrize("count", [0, 1, 2, 4])
def test_oracle_temporal_vote_matches_numpy_restatement(inp, count):
    pp, pf = inp["prev_planes"][:count], inp["prev_flow"][:count]
    unsm, sm = po.classify_temporal(inp["naive"], *inp["params"], pp, pf)
    assert np.array_equal(unsm, po.classify(inp["naive"], *inp["params"]))
    assert np.array_equal(sm, np_temporal(unsm, pp, pf, 0) if count else unsm)
    u2, planes = po.sp_planeseg_temporal(inp["deriv2"], inp["labels"], inp["n_labels"], *inp["params"], pp, pf)
    assert np.array_equal(u2, po.classify(inp["deriv2"], *inp["params"], channel_stride=2))
    voted = np_temporal(u2, pp, pf, 1) if count else u2
    assert np.array_equal(planes, np_sp_assign(inp["labels"], voted, inp["n_labels"]))
    if count == 0:  # no history: identical to the plain stage
        assert np.array_equal(planes, po.sp_planeseg(inp["deriv2"], inp["labels"], inp["n_labels"], *inp["params"])[1])
    else:
        assert (sm != unsm).mean() > 0.01  # the vote does something on these inputs


def test_oracle_temporal_vote_edge_cases():
    H, W = 6, 9
    d = np.full((H, W), 5, np.int16)  # horizontal everywhere with (1, 30, -3, 1)
    d[:, 4:] = -32768                 # unknown on the right
    zero = np.zeros((H, W, 2), np.int16)
    allv = np.full((H, W), 1, np.uint8)
    # naive: one previous vertical vote ties 1:1 with the current horizontal -> "H > V" fails -> vertical
    _, sm = po.classify_temporal(d, 1, 30, -3, 1, [allv], [zero])
    assert (sm[:, :4] == 1).all() and (sm[:, 4:] == 1).all()  # unknown pixels: the only H/V vote is vertical
    # sp weighting: current counts twice -> horizontal survives one vertical vote, falls to two... (2 vs 2 -> V)
    lab = np.zeros((H, W), np.uint16)
    d2 = np.stack([d, d], axis=2)
    ref1 = np_temporal(po.classify(d2, 1, 30, -3, 1, channel_stride=2), [allv], [zero], 1)
    assert (ref1[:, :4] == 0).all() and (ref1[:, 4:] == 2).all()  # unknown (2 votes) beats one vertical vote
    # flow that leaves the image is skipped but stays accumulated: -31 >> 5 == -1 (arithmetic shift)
    fl = np.zeros((H, W, 2), np.int16)
    fl[:, :, 0] = -31  # x - (-1) = x + 1
    mark = np.full((H, W), 2, np.uint8)
    mark[:, 1] = 1  # only column 1 is vertical
    dd = np.full((H, W), -32768, np.int16)
    _, sm = po.classify_temporal(dd, 1, 30, -3, 1, [mark], [fl])
    assert (sm[:, 0] == 1).all() and (sm[:, 1:] == 2).all()
    far = np.zeros((H, W, 2), np.int16)
    far[:, :, 0] = 32 * 100  # first hop leaves the image; the second (flow -100 px) comes back to x
    back = np.zeros((H, W, 2), np.int16)
    back[:, :, 0] = -32 * 100
    _, sm = po.classify_temporal(dd, 1, 30, -3, 1, [allv, np.zeros((H, W), np.uint8)], [far, back])
    assert (sm == 0).all()  # only the second frame (horizontal) voted
    with pytest.raises(RuntimeError):
        po.classify_temporal(dd, 1, 30, -3, 1, [np.full((H, W), 3, np.uint8)], [zero])
    assert lab.sum() == 0


def test_oracle_label_statistics_and_inliers_match_numpy(inp):
    lab, xyz, n = inp["labels"], inp["xyz"], inp["n_labels"]
    cnt, inv = po.label_statistics(lab, xyz, n)
    z = xyz[:, :, 2]
    valid = np.isfinite(z) & (z <= 40.0) & (z > 0.0)
    assert np.array_equal(cnt, np.bincount(lab.ravel(), minlength=n))
    assert np.array_equal(inv, np.bincount(lab[~valid], minlength=n))
    assert 0 < inv.sum() < cnt.sum()
    for thr in inp["thresholds"]:
        got = po.region_inliers(lab, xyz, n, inp["planes"], thr)
        for i, (a, b, c, d) in enumerate(inp["planes"]):
            q = xyz.astype(np.float64)
            with np.errstate(invalid="ignore"):
                dist = np.abs(a * q[:, :, 0] + b * q[:, :, 1] + c * q[:, :, 2] + d) / np.sqrt(a * a + b * b + c * c)
            assert np.array_equal(got[i], np.bincount(lab[valid & (dist < thr)], minlength=n)), i
    with pytest.raises(RuntimeError):
        po.label_statistics(lab, xyz, n - 1)  # label out of range


# ---- pinned against the reference's own kernels -------------------------------------------------------------
@pytest.mark.parametrize("count", [0, 1, 2, 3, 4])
def test_golden_temporal_vote(inp, gold, count):
    pp, pf = inp["prev_planes"][:count], inp["prev_flow"][:count]
    unsm, sm = po.classify_temporal(inp["naive"], *inp["params"], pp, pf)
    assert np.array_equal(unsm, gold[f"naive_unsm_{count}"])
    if count:  # with no history the reference kernel leaves the smoothed image untouched
        assert np.array_equal(sm, gold[f"naive_smoothed_{count}"])
    else:
        assert (gold["naive_smoothed_0"] == 0).all()  # the harness's zero-filled allocation, never written
    u2, planes = po.sp_planeseg_temporal(inp["deriv2"], inp["labels"], inp["n_labels"], *inp["params"], pp, pf)
    assert np.array_equal(u2, gold[f"sp_unsm_{count}"])
    assert np.array_equal(planes, gold[f"sp_planes_{count}"])


def test_golden_planefit_consumers(inp, gold):
    cnt, inv = po.label_statistics(inp["labels"], inp["xyz"], inp["n_labels"])
    assert np.array_equal(cnt, gold["label_stats"][:, 0]) and np.array_equal(inv, gold["label_stats"][:, 1])
    for i, thr in enumerate(inp["thresholds"]):
        got = po.region_inliers(inp["labels"], inp["xyz"], inp["n_labels"], inp["planes"], thr)
        ref = gold[f"inliers_{i}"]
        # the reference's distance is compiled with FMA contraction: a pixel within rounding of the threshold may
        # fall on the other side.  None does on these inputs; the bound documents the tolerance.
        assert np.abs(got.astype(np.int64) - ref.astype(np.int64)).sum() <= 2, i


def test_oracle_overlays_and_golden(inp):
    base = np.load(BASE)
    bgr, planes, labels = base["left_bgr"], base["sp_planes"], inp["labels"]
    ov = po.overlay_planes(bgr, planes)
    col = np.zeros(planes.shape + (3,), np.int32)
    for pl in range(3):
        col[..., pl][planes == pl] = 127  # H -> blue, V -> green, UNKNOWN -> red (BGR order)
    assert np.array_equal(ov, (bgr // 2 + col).astype(np.uint8))
    ob = po.overlay_boundaries(bgr, labels, out=np.full(bgr.shape, 77, np.uint8))
    edge = np.zeros(labels.shape, bool)
    edge[:-1, :-1] = (labels[:-1, :-1] != labels[:-1, 1:]) | (labels[:-1, :-1] != labels[1:, :-1])
    exp = bgr.copy()
    exp[edge] = (0, 0, 255)
    exp[-1, :] = 77
    exp[:, -1] = 77
    assert np.array_equal(ob, exp)
    with pytest.raises(RuntimeError):
        po.overlay_planes(bgr, np.full(planes.shape, 3, np.uint8))
    if os.path.exists(GOLD) and "overlay_planes" in np.load(GOLD).files:  # the reference's own kernels
        g = np.load(GOLD)
        assert np.array_equal(ov, g["overlay_planes"]) and np.array_equal(ob, g["overlay_boundaries"])


# ---- CUDA path through the C ABI --------------------------------------------------------------------------------
def _ctx(W, H, superpixels=True, block=12):
    import cart_slam_b200 as cb

    return cb.Context(cb.Config(W, H, max_batch=1, enable_sgm=False, enable_superpixels=superpixels, sp_block_size=block))


@pytest.mark.gpu
@pytest.mark.parametrize("count", [0, 1, 3, 4])
def test_gpu_temporal_vote_bit_exact(inp, count):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    H, W = inp["labels"].shape
    pp, pf = inp["prev_planes"][:count], inp["prev_flow"][:count]
    with _ctx(W, H) as ctx:
        unsm, sm = ctx.classify_temporal(dev(inp["naive"]), inp["params"], [dev(a) for a in pp], [dev(a) for a in pf])
        o_unsm, o_sm = po.classify_temporal(inp["naive"], *inp["params"], pp, pf)
        assert np.array_equal(unsm.cpu().numpy(), o_unsm) and np.array_equal(sm.cpu().numpy(), o_sm)
        # two-channel derivative, channel 0
        unsm, sm = ctx.classify_temporal(dev(inp["deriv2"]), inp["params"], [dev(a) for a in pp], [dev(a) for a in pf])
        o_unsm, o_sm = po.classify_temporal(inp["deriv2"], *inp["params"], pp, pf, channel_stride=2)
        assert np.array_equal(unsm.cpu().numpy(), o_unsm) and np.array_equal(sm.cpu().numpy(), o_sm)
        u2, planes = ctx.sp_planeseg_temporal(dev(inp["deriv2"]), dev(inp["labels"]), inp["params"], [dev(a) for a in pp],
                                              [dev(a) for a in pf], max_label=inp["n_labels"])
        o_u2, o_planes = po.sp_planeseg_temporal(inp["deriv2"], inp["labels"], inp["n_labels"], *inp["params"], pp, pf)
        assert np.array_equal(u2.cpu().numpy(), o_u2) and np.array_equal(planes.cpu().numpy(), o_planes)


@pytest.mark.gpu
def test_gpu_temporal_vote_kitti_size_and_errors():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import cart_slam_b200 as cb

    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    W, H = 1242, 375
    rng = np.random.default_rng(7)
    deriv2 = rng.integers(-8, 40, (H, W, 2)).astype(np.int16)
    deriv2[rng.random((H, W)) < 0.1] = -32768
    labels, n = po.block_init(W, H, 12, 12)
    pp = [rng.integers(0, 3, (H, W)).astype(np.uint8) for _ in range(3)]
    pf = [rng.integers(-4 * 32, 4 * 32, (H, W, 2)).astype(np.int16) for _ in range(3)]
    with _ctx(W, H) as ctx:
        u, planes = ctx.sp_planeseg_temporal(dev(deriv2), dev(labels), (1, 30, -3, 1), [dev(a) for a in pp], [dev(a) for a in pf],
                                             max_label=n)
        o_u, o_planes = po.sp_planeseg_temporal(deriv2, labels, n, 1, 30, -3, 1, pp, pf)
        assert np.array_equal(u.cpu().numpy(), o_u) and np.array_equal(planes.cpu().numpy(), o_planes)
        unsm, sm = ctx.classify_temporal(dev(deriv2[:, :, 0].copy()), (1, 30, -3, 1), [dev(a) for a in pp], [dev(a) for a in pf])
        o = po.classify_temporal(deriv2[:, :, 0].copy(), 1, 30, -3, 1, pp, pf)
        assert np.array_equal(unsm.cpu().numpy(), o[0]) and np.array_equal(sm.cpu().numpy(), o[1])
        with pytest.raises(cb.CartB200Error):  # more history than CARTB200_MAX_TEMPORAL_DISTANCE
            ctx.classify_temporal(dev(deriv2), (1, 30, -3, 1), [dev(pp[0])] * 9, [dev(pf[0])] * 9)


@pytest.mark.gpu
def test_gpu_planefit_consumers_bit_exact(inp):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    H, W = inp["labels"].shape
    n = inp["n_labels"]
    with _ctx(W, H, superpixels=False) as ctx:
        cnt, inv = ctx.label_statistics(dev(inp["labels"]), dev(inp["xyz"]), n)
        o_cnt, o_inv = po.label_statistics(inp["labels"], inp["xyz"], n)
        assert np.array_equal(cnt.cpu().numpy().astype(np.uint32), o_cnt)
        assert np.array_equal(inv.cpu().numpy().astype(np.uint32), o_inv)
        planes = np.concatenate([inp["planes"]] * 3)[:20]  # more than one plane set per launch (16)
        for thr in inp["thresholds"]:
            got = ctx.region_inliers(dev(inp["labels"]), dev(inp["xyz"]), n, planes, thr).cpu().numpy().astype(np.uint32)
            assert np.array_equal(got, po.region_inliers(inp["labels"], inp["xyz"], n, planes, thr))


@pytest.mark.gpu
def test_gpu_overlays_bit_exact(inp):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    base = np.load(BASE)
    bgr, planes, labels = base["left_bgr"], base["sp_planes"], inp["labels"]
    H, W = labels.shape
    with _ctx(W, H, superpixels=False) as ctx:
        assert np.array_equal(ctx.overlay_planes(dev(bgr), dev(planes)).cpu().numpy(), po.overlay_planes(bgr, planes))
        seed = np.full(bgr.shape, 77, np.uint8)
        got = ctx.overlay_superpixel_boundaries(dev(bgr), dev(labels), out=dev(seed)).cpu().numpy()
        assert np.array_equal(got, po.overlay_boundaries(bgr, labels, out=seed))


# ---- randomised properties of the oracle's temporal vote (hypothesis) -------------------------------------------
def test_temporal_vote_properties_randomised():
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=40, deadline=None)
    @hyp.given(st.integers(0, 2 ** 31 - 1), st.integers(1, 4), st.integers(3, 24), st.integers(3, 24))
    def prop(seed, count, H, W):
        rng = np.random.default_rng(seed)
        d = rng.integers(-6, 36, (H, W)).astype(np.int16)
        d[rng.random((H, W)) < 0.2] = -32768
        pp = [rng.integers(0, 3, (H, W)).astype(np.uint8) for _ in range(count)]
        pf = [rng.integers(-32768, 32768, (H, W, 2)).astype(np.int16) if rng.random() < 0.3 else
              rng.integers(-96, 97, (H, W, 2)).astype(np.int16) for _ in range(count)]
        unsm, sm = po.classify_temporal(d, 1, 30, -3, 1, pp, pf)
        assert np.array_equal(sm, np_temporal(unsm, pp, pf, 0))          # independent restatement, any flow
        assert sm.max() <= 2
        # a pixel can only be UNKNOWN after the naive vote if neither H nor V received a vote
        assert ((sm == 2) <= (unsm == 2)).all()
        # zero flow and a history equal to the current classes changes nothing for decided pixels
        zero = [np.zeros((H, W, 2), np.int16)] * count
        _, same = po.classify_temporal(d, 1, 30, -3, 1, [unsm] * count, zero)
        assert np.array_equal(same[unsm != 2], unsm[unsm != 2])
        # superpixel weighting: one label for the whole image -> the image-wide majority of the voted planes
        d2 = np.stack([d, d], axis=2)
        lab = np.zeros((H, W), np.uint16)
        _, planes = po.sp_planeseg_temporal(d2, lab, 1, 1, 30, -3, 1, pp, pf)
        voted = np_temporal(unsm, pp, pf, 1)
        assert np.array_equal(planes, np_sp_assign(lab, voted, 1))
        assert len(np.unique(planes)) == 1

    prop()
