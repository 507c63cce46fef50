"""Generates tests/golden/ref_kernels.npz on a GPU box: outputs of the REFERENCE'S OWN kernels
(oracle/_ref/libref.so, built by oracle/ref/build_ref.sh from /root/reference) on seeded inputs.
Run through gpurun:   gpurun -- python tests/golden/make_golden.py
The CPU test-suite (tests/test_golden.py) then pins the oracle against these vectors on the masks where
the reference is deterministic.  Also prints the reference kernels' GPU times."""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import pyoracle as po  # noqa: E402
from cart_slam_b200.synth import SyntheticSequence  # noqa: E402


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref.so"))
    out = {}
    times = {}
    ms = C.c_float()
    # --- inputs: a real oracle disparity (with the SGM invalid marker 48) and a noisy random one ---
    W, H, D = 300, 290, 64
    seq = SyntheticSequence(W, H, D, n_frames=2, tint=True)
    l, r, _ = seq.frame(1)
    sgm = po.sgm_compute(l, r, D)
    rng = np.random.default_rng(2024)
    noisy = (np.arange(H)[:, None] * 2 + np.arange(W)[None, :] // 5 + 64 + rng.integers(-4, 5, (H, W))).astype(np.int16)
    noisy[rng.random((H, W)) < 0.08] = -32768
    out["left_bgr"], out["sgm_disparity"], out["noisy_disparity"] = l, sgm, noisy

    # interpolate r=2/it=1 and r=3/it=2 on the SGM output
    for name, (rad, it) in {"interp_r2_i1": (2, 1), "interp_r3_i2": (3, 2)}.items():
        d = sgm.copy()
        assert lib.ref_interpolate(p(d), W, H, rad, it, 4 * 16, W, 5, C.byref(ms)) == 0
        out[name] = d
        times[name] = ms.value
    smooth = po.interpolate(sgm, 2, 1, 64, W)  # canonical input for the next stages
    out["smooth_disparity"] = smooth

    # race-free inputs for the naive low-pass (the 5-tap vertical mean is the identity on them)
    rows = np.ascontiguousarray(np.broadcast_to((np.arange(H)[:, None] * 3 + 70).astype(np.int16), (H, W)))
    cols = np.ascontiguousarray(np.broadcast_to((np.arange(W)[None, :] * 2 + 70).astype(np.int16), (H, W)))
    out["rows_disparity"], out["columns_disparity"] = rows, cols

    for tag, disp in (("smooth", smooth), ("noisy", noisy), ("rows", rows), ("columns", cols)):
        deriv = np.zeros((H, W, 2), np.int16)
        hist = np.zeros((256, 2), np.int32)
        assert lib.ref_derivative(p(disp), W, H, p(deriv), p(hist), 5, C.byref(ms)) == 0
        out[f"deriv_{tag}"], out[f"deriv_hist_{tag}"] = deriv, hist
        times[f"derivative_{tag}"] = ms.value
        nd = np.zeros((H, W), np.int16)
        nh = np.zeros(256, np.int32)
        pl = np.zeros((H, W), np.uint8)
        params = np.array([1, 30, -3, 1], np.int32)
        assert lib.ref_naive(p(disp), W, H, p(params), p(nd), p(nh), p(pl), 5, C.byref(ms)) == 0
        out[f"naive_deriv_{tag}"], out[f"naive_hist_{tag}"], out[f"naive_planes_{tag}"] = nd, nh, pl
        times[f"naive_{tag}"] = ms.value

    # superpixels: border map + relaxation on the oracle's derivative
    o_deriv, _ = po.derivative(smooth)
    out["oracle_deriv_smooth"] = o_deriv
    lab0, nlab = po.block_init(W, H, 12, 12)
    jitter = lab0.copy()
    ys, xs = rng.integers(0, H, 600), rng.integers(0, W, 600)
    jitter[ys, xs] = rng.integers(0, nlab, 600)
    out["labels_jitter"] = jitter
    b = np.zeros((H, W), np.uint8)
    assert lib.ref_border_map(p(jitter), W, H, p(b)) >= 0
    out["border_jitter"] = b
    ycc = po.ycrcb(l)
    for name, its, kw in (("relax_4", 4, dict(wc=0.1, pr=0.0, wd=1.0, wi=1.5)),
                          ("relax_prog_3", 3, dict(wc=0.03, pr=1.0, wd=1.0, wi=1.5)),
                          ("relax_nodisp_3", 3, dict(wc=0.1, pr=0.0, wd=0.0, wi=1.5))):
        lab = lab0.copy()
        f = lib.ref_relax
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_double] * 6 + [C.c_void_p]
        dv = p(o_deriv) if kw["wd"] > 0 else None
        assert f(p(lab), W, H, nlab, p(ycc), dv, its, 0.5, 0.5 / np.sqrt(2), kw["wc"], kw["pr"], kw["wd"], kw["wi"], C.byref(ms)) == 0
        out[name] = lab
        times[name] = ms.value
        # run-to-run self-agreement of the reference (its stored feature costs race)
        lab2 = lab0.copy()
        assert f(p(lab2), W, H, nlab, p(ycc), dv, its, 0.5, 0.5 / np.sqrt(2), kw["wc"], kw["pr"], kw["wd"], kw["wi"], C.byref(ms)) == 0
        out[name + "_rerun"] = lab2
        times[name + "_self_agreement"] = float((lab == lab2).mean())
    # SP planeseg on the relaxed labels
    unsm = np.zeros((H, W), np.uint8)
    pls = np.zeros((H, W), np.uint8)
    params = np.array([1, 30, -3, 1], np.int32)
    assert lib.ref_sp_planeseg(p(o_deriv), p(out["relax_4"]), W, H, nlab, p(params), p(unsm), p(pls), 5, C.byref(ms)) == 0
    out["sp_unsm"], out["sp_planes"] = unsm, pls
    times["sp_planeseg"] = ms.value

    # KITTI-size timings of the reference kernels (the "reference GPU path" for the non-SGM stages)
    Wk, Hk = 1242, 375
    seqk = SyntheticSequence(Wk, Hk, 128, n_frames=2, tint=True)
    lk, rk, _ = seqk.frame(1)
    dk = np.ascontiguousarray(np.tile(smooth, (2, 5))[:Hk, :Wk])
    lib.ref_interpolate(p(dk.copy()), Wk, Hk, 2, 1, 64, Wk, 10, C.byref(ms)); times["K_interpolate_r2"] = ms.value
    derk = np.zeros((Hk, Wk, 2), np.int16); hk = np.zeros((256, 2), np.int32)
    lib.ref_derivative(p(dk), Wk, Hk, p(derk), p(hk), 10, C.byref(ms)); times["K_derivative"] = ms.value
    ndk = np.zeros((Hk, Wk), np.int16); nhk = np.zeros(256, np.int32); plk = np.zeros((Hk, Wk), np.uint8)
    lib.ref_naive(p(dk), Wk, Hk, p(params), p(ndk), p(nhk), p(plk), 10, C.byref(ms)); times["K_naive_planeseg"] = ms.value
    labk, nlk = po.block_init(Wk, Hk, 12, 12)
    yk = po.ycrcb(lk)
    for its in (8, 24):
        lab = labk.copy()
        lib.ref_relax(p(lab), Wk, Hk, nlk, p(yk), p(derk), its, 0.5, 0.5 / np.sqrt(2), 0.1, 0.0, 1.0, 1.5, C.byref(ms))
        lab = labk.copy()
        lib.ref_relax(p(lab), Wk, Hk, nlk, p(yk), p(derk), its, 0.5, 0.5 / np.sqrt(2), 0.1, 0.0, 1.0, 1.5, C.byref(ms))
        times[f"K_relax_{its}it"] = ms.value
    uk = np.zeros((Hk, Wk), np.uint8); pk = np.zeros((Hk, Wk), np.uint8)
    lib.ref_sp_planeseg(p(derk), p(lab), Wk, Hk, nlk, p(params), p(uk), p(pk), 10, C.byref(ms)); times["K_sp_planeseg"] = ms.value

    np.savez_compressed(os.path.join(HERE, "ref_kernels.npz"), **out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for dst in (os.path.join(ROOT, "gpurun_out", "ref_kernels.npz"),):
        np.savez_compressed(dst, **out)
    json.dump(times, open(os.path.join(ROOT, "gpurun_out", "ref_kernel_times_ms.json"), "w"), indent=1)
    print(json.dumps(times, indent=1))


if __name__ == "__main__":
    main()
