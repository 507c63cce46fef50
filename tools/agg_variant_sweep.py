"""Sweep of the horizontal aggregation kernel shapes (CARTB200_HORIZ_VARIANT): time per path of a 64-frame batch and
bit-identity of the path volumes with the default shape.   python tools/agg_variant_sweep.py [--variants 0,3,4,5,6]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import cart_slam_b200 as cb  # noqa: E402
from cart_slam_b200.synth import SyntheticSequence  # noqa: E402
from stage_bench import timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--variants", default="0,3,4,5,6,0")
    ap.add_argument("--width", type=int, default=1242)
    ap.add_argument("--height", type=int, default=375)
    ap.add_argument("--disp", type=int, default=128)
    args = ap.parse_args()
    W, H, D, B = args.width, args.height, args.disp, args.batch
    seq = SyntheticSequence(W, H, D, n_frames=4, tint=True)
    fr = [seq.frame(1 + (i % 4)) for i in range(B)]
    L = torch.from_numpy(np.stack([f[0] for f in fr])).cuda()
    R = torch.from_numpy(np.stack([f[1] for f in fr])).cuda()
    cfg = cb.Config(W, H, max_batch=B, num_disparities=D, paths=4, enable_superpixels=0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    alg = B * (2 * 4 * W * H + W * H * D)
    out = {}
    with cb.Context(cfg) as ctx:
        ctx.sgm_gray_census(L, R)
        os.environ["CARTB200_HORIZ_VARIANT"] = "0"
        ctx.sgm_aggregate(B)
        nchk = min(B, 4)
        ref = [ctx.sgm_intermediate(10 + p, nchk).clone() for p in (0, 1)]
        for v in [int(x) for x in args.variants.split(",")]:
            os.environ["CARTB200_HORIZ_VARIANT"] = str(v)
            ctx.sgm_aggregate(B)
            same = all(bool(torch.equal(ctx.sgm_intermediate(10 + p, nchk), ref[p])) for p in (0, 1))
            t0 = timeit(lambda: ctx.sgm_aggregate_path(B, 0), reps=8)
            t1 = timeit(lambda: ctx.sgm_aggregate_path(B, 1), reps=8)
            tall = timeit(lambda: ctx.sgm_aggregate(B), reps=8)
            out[str(v)] = {"path0_ms": t0, "path1_ms": t1, "all_paths_ms": tall, "identical": same}
            print(f"variant {v}: path0 {t0:6.3f} ms ({alg / t0 / 1e6 / peak:5.3f})  path1 {t1:6.3f} ms  all four paths {tall:6.3f} ms "
                  f"({4 * alg / tall / 1e6 / peak:5.3f} of HBM peak)  volumes identical: {same}", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "agg_variant_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
