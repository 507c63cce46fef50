"""Small end-to-end invocation for compute-sanitizer (memcheck / racecheck / synccheck) on the GPU box:
   compute-sanitizer --tool racecheck python tools/sanitize_relax.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cart_slam_b200 as cb  # noqa: E402
from cart_slam_b200.synth import SyntheticSequence  # noqa: E402

W, H, D = 200, 136, 64  # ragged: edge tiles on every side, a partial tile row and column
seq = SyntheticSequence(W, H, D, n_frames=3, tint=True)
fr = [seq.frame(i + 1)[:2] for i in range(3)]
L = torch.from_numpy(np.stack([f[0] for f in fr])).cuda()
R = torch.from_numpy(np.stack([f[1] for f in fr])).cuda()
cfg = cb.Config(W, H, max_batch=3, num_disparities=D, smoothing_radius=2, smoothing_iterations=1, sp_block_size=9)
with cb.Context(cfg) as ctx:
    disp = ctx.disparity(L, R)
    deriv, hist = ctx.derivative(disp)
    ctx.superpixels_reset(3)
    labels = ctx.superpixels_relax(L, deriv, 3)
    unsm, planes = ctx.sp_planeseg(deriv, labels, [[1, 30, -3, 1]] * 3)
    opts = cb.SequenceOptions(pipeline=1, provider=1, sp_initial_iterations=3, sp_iterations=2, sp_reset_iterations=2)
    out = ctx.run_sequence_device(opts, L, R)
    torch.cuda.synchronize()
print("ok", int(labels.sum()), int(planes.sum()), int(out.sum()))
