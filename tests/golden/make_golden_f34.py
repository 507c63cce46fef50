"""Generates tests/golden/ref_kernels_f34.npz on a GPU box: outputs of the REFERENCE'S OWN kernels for the temporal
smoothing vote (classifyPlanes / performSuperPixelClassifications with previousPlanesCount > 0) and the plane-fit
superpixel consumers (countPixels, calculateRegionDistance), from oracle/_ref/libref.so (oracle/ref/build_ref.sh).
Run through gpurun:   gpurun -- python tests/golden/make_golden_f34.py
Inputs are regenerated from tests/golden/f34_inputs.py; only the reference outputs and an input checksum are stored."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)

import f34_inputs  # noqa: E402


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def ptrs(arrs):
    a = (C.c_void_p * max(1, len(arrs)))()
    for i, x in enumerate(arrs):
        a[i] = x.ctypes.data
    return a


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref.so"))
    base = np.load(os.path.join(HERE, "ref_kernels.npz"))
    d = f34_inputs.make(base)
    H, W = d["labels"].shape
    params = np.array(d["params"], np.int32)
    out = {"crc": np.array([d["crc"]], np.int64)}
    for count in (0, 1, 2, 3, 4):
        pp, pf = d["prev_planes"][:count], d["prev_flow"][:count]
        pl = np.zeros((H, W), np.uint8)
        sm = np.full((H, W), 255, np.uint8)
        assert lib.ref_classify_temporal(p(d["naive"]), W, H, p(params), count, ptrs(pp), ptrs(pf), p(pl), p(sm)) == 0
        out[f"naive_unsm_{count}"], out[f"naive_smoothed_{count}"] = pl, sm
        pu = np.zeros((H, W), np.uint8)
        ps = np.zeros((H, W), np.uint8)
        assert lib.ref_sp_planeseg_temporal(p(d["deriv2"]), p(d["labels"]), W, H, d["n_labels"], p(params), count, ptrs(pp),
                                            ptrs(pf), p(pu), p(ps)) == 0
        out[f"sp_unsm_{count}"], out[f"sp_planes_{count}"] = pu, ps
    stats = np.zeros((d["n_labels"], 2), np.uint16)
    assert lib.ref_label_statistics(p(d["labels"]), p(d["xyz"]), W, H, d["n_labels"], p(stats)) == 0
    out["label_stats"] = stats
    f = lib.ref_region_inliers
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
    for i, thr in enumerate(d["thresholds"]):
        inl = np.zeros((len(d["planes"]), d["n_labels"]), np.uint32)
        assert f(p(d["labels"]), p(d["xyz"]), W, H, d["n_labels"], p(d["planes"]), len(d["planes"]), thr, p(inl)) == 0
        out[f"inliers_{i}"] = inl
    # overlay kernels (visual QA): plane overlay on the reference's own superpixel planes, boundary overlay on its labels
    bgr = np.ascontiguousarray(base["left_bgr"])
    ov = np.zeros((H, W, 3), np.uint8)
    assert lib.ref_overlay_planes(p(bgr), p(np.ascontiguousarray(base["sp_planes"])), W, H, p(ov)) == 0
    out["overlay_planes"] = ov
    ob = np.full((H, W, 3), 77, np.uint8)
    assert lib.ref_overlay_boundaries(p(bgr), p(d["labels"]), W, H, p(ob)) == 0
    out["overlay_boundaries"] = ob
    np.savez_compressed(os.path.join(HERE, "ref_kernels_f34.npz"), **out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_kernels_f34.npz"), **out)
    print({k: (v.shape, int(v.astype(np.int64).sum())) for k, v in out.items()})


if __name__ == "__main__":
    main()
