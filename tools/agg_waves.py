"""Wave-quantisation experiment for the aggregation kernels: time per frame of a single direction and of the full
4-path call as a function of the batch size n (grid = CTAs per frame x n).  python tools/agg_waves.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cart_slam_b200 as cb  # noqa: E402
from cart_slam_b200.synth import SyntheticSequence  # noqa: E402
from stage_bench import timeit  # noqa: E402


def main():
    W, H, D, B = 1242, 375, 128, 64
    seq = SyntheticSequence(W, H, D, n_frames=4)
    fr = [seq.frame(1 + (i % 4)) for i in range(B)]
    L = torch.from_numpy(np.stack([f[0] for f in fr])).cuda()
    R = torch.from_numpy(np.stack([f[1] for f in fr])).cuda()
    cfg = cb.Config(W, H, max_batch=B, num_disparities=D, paths=4, enable_superpixels=False)
    out = {}
    bytes_per_frame_path = W * H * D + 2 * 4 * W * H
    with cb.Context(cfg) as ctx:
        ctx.sgm_gray_census(L, R)
        for n in (16, 32, 48, 56, 60, 61, 62, 63, 64):
            row = {}
            for p in (0, 2):
                ms = timeit(lambda: ctx.sgm_aggregate_path(n, p), reps=10)
                row[f"path{p}_us_per_frame"] = ms / n * 1e3
            ms = timeit(lambda: ctx.sgm_aggregate(n), reps=10)
            row["all_us_per_frame_path"] = ms / n / 4 * 1e3
            row["all_GBps"] = 4 * n * bytes_per_frame_path / ms / 1e6
            out[n] = row
            print(n, json.dumps(row))
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "agg_waves.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
