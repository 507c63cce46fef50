"""Pins the CPU oracle against the REFERENCE'S OWN CUDA kernels.

tests/golden/ref_kernels.npz holds outputs of the reference kernels (compiled unmodified from
/root/reference by oracle/ref/build_ref.sh, executed on a B200 by tests/golden/make_golden.py).
Deterministic reference stages must match the oracle bit-for-bit on the oracle's "defined" mask (the
pixels whose inputs the reference does not read from uninitialised / out-of-image memory).
Stages where the reference races against itself (Q10 low-pass, Q11 interpolation, Q13 stored feature
costs) cannot be bit-exact by construction: their agreement with the canonical schedule is bounded below."""
import os

import numpy as np
import pytest

import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_kernels.npz")


@pytest.fixture(scope="module")
def g():
    if not os.path.exists(GOLD):
        pytest.skip("golden vectors missing")
    return np.load(GOLD)


@pytest.mark.parametrize("tag", ["smooth", "noisy"])
def test_derivative_bit_exact_on_defined_region(g, tag):
    disp = g[f"{tag}_disparity"]
    o, h, m = po.derivative(disp, want_mask=True)
    ref = g[f"deriv_{tag}"]
    for ch in range(2):
        mm = m[:, :, ch] == 1
        assert mm.mean() > 0.98
        assert np.array_equal(o[:, :, ch][mm], ref[:, :, ch][mm]), f"channel {ch}"
    # SURVEY Q17 confirmed on hardware: the reference's own histogram is polluted by its shared-memory
    # overflow (bins go negative), so the oracle's histogram is a recount of the oracle's derivative image
    rh = g[f"deriv_hist_{tag}"]
    assert rh.min() < 0 or not np.array_equal(rh, h)
    for ch in range(2):
        ok = (o[:, :, ch] != -32768) & (o[:, :, ch] >= -128) & (o[:, :, ch] <= 127)
        assert h[:, ch].sum() == ok.sum()


def test_naive_race_free_inputs_bit_exact(g):
    """Inputs on which the 5-tap vertical mean is the identity (no invalid pixels, linear in y): the
    reference's in-place race (Q10) cannot change anything, so tile quirks + derivative must match exactly."""
    for name in ("rows", "columns"):
        key = f"naive_deriv_{name}"
        if key not in g.files:
            pytest.skip("golden file predates the race-free cases")
        o, h, m = po.naive_derivative(g[f"{name}_disparity"], want_mask=True)
        mm = m == 1
        assert np.array_equal(o[mm], g[key][mm]), name
        assert np.array_equal(po.classify(g[key], 1, 30, -3, 1), g[f"naive_planes_{name}"])


@pytest.mark.parametrize("tag,floor", [("smooth", 0.90), ("noisy", 0.45)])
def test_naive_agreement_with_racy_reference(g, tag, floor):
    o, h, m = po.naive_derivative(g[f"{tag}_disparity"], want_mask=True)
    ref = g[f"naive_deriv_{tag}"]
    assert (o == ref)[m == 1].mean() >= floor
    # classification of the reference's own derivative image is deterministic: exact
    assert np.array_equal(po.classify(ref, 1, 30, -3, 1), g[f"naive_planes_{tag}"])


@pytest.mark.parametrize("name,radius,iters,floor", [("interp_r2_i1", 2, 1, 0.90), ("interp_r3_i2", 3, 2, 0.80)])
def test_interpolate_agreement_with_racy_reference(g, name, radius, iters, floor):
    W = g["sgm_disparity"].shape[1]
    o, m = po.interpolate(g["sgm_disparity"], radius, iters, 64, W, want_mask=True)
    ref = g[name]
    assert (o == ref)[m == 1].mean() >= floor
    # invalid marker handling (Q15): SGM's 48 never survives, everything valid lies in (64, W)
    valid = ref != -32768
    assert ((ref[valid] > 0) & (ref[valid] < W)).all()


def test_border_map_bit_exact(g):
    b, m = po.border_map(g["labels_jitter"], want_mask=True)
    assert np.array_equal(b[m == 1], g["border_jitter"][m == 1])


def test_sp_planeseg_bit_exact(g):
    lab0, n = po.block_init(300, 290, 12, 12)
    u, p = po.sp_planeseg(g["oracle_deriv_smooth"], g["relax_4"], n, 1, 30, -3, 1)
    assert np.array_equal(u, g["sp_unsm"]) and np.array_equal(p, g["sp_planes"])


@pytest.mark.parametrize("name,its,kw", [
    ("relax_4", 4, dict(w_compact=0.1, progressive=0.0, w_disp=1.0)),
    ("relax_prog_3", 3, dict(w_compact=0.03, progressive=1.0, w_disp=1.0)),
    ("relax_nodisp_3", 3, dict(w_compact=0.1, progressive=0.0, w_disp=0.0)),
])
def test_relax_agreement_with_reference(g, name, its, kw):
    lab0, n = po.block_init(300, 290, 12, 12)
    ycc = po.ycrcb(g["left_bgr"])
    lab, bc, mv = po.sp_relax(lab0, n, ycc, g["oracle_deriv_smooth"] if kw["w_disp"] > 0 else None, its, **kw)
    ref = g[name]
    agree = (lab == ref).mean()
    # the reference itself is not reproducible here (stored featureCost race, Q13); its run-to-run
    # self-agreement is recorded in the golden file when available
    floor = 0.985
    assert agree >= floor, agree
    moved_ref = (ref != lab0).mean()
    assert 0.5 * moved_ref <= (lab != lab0).mean() <= 2.0 * moved_ref
