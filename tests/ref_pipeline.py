"""Oracle-side restatement of the reference's per-sequence module logic (test infrastructure).

naive pipeline  = config/modules/kitti-naive-segmentation.json:
    ImageDisparityModule (disparity.cu:49-80) -> DisparityPlaneSegmentationModule (planeseg.cu:246-403)
superpixel pipeline = config/modules/kitti-planeseg.json minus optflow/depth/vis/temporal smoothing:
    ImageDisparityModule -> ImageDisparityDerivativeModule (derivative.cu:151-184)
    -> SuperPixelModule (superpixels.cu:71-121) -> SuperPixelDisparityPlaneSegmentationModule (sp_planeseg.cu:224-388)
Frames are processed strictly in id order (ids start at 1)."""
import numpy as np

import pyoracle as po


def disparity(l, r, cfg):
    d = po.sgm_compute(l, r, cfg["D"], cfg.get("min_disp", 4), cfg.get("p1", 10), cfg.get("p2", 120),
                       cfg.get("ur", 12), cfg.get("paths", 4))
    if cfg.get("radius", -1) > 0:
        d = po.interpolate(d, cfg["radius"], cfg.get("iters", 5), cfg.get("min_disp", 4) * 16, l.shape[1])
    return d


def naive_sequence(frames, cfg, provider="histogram_peak", static=(1, 30, -3, 1), update=30, reset=10):
    running = np.zeros(256, np.int64)
    params = [0, 0, 0, 0, 0, 0] if provider == "histogram_peak" else [0, 0] + list(static)
    out = []
    for i, (l, r) in enumerate(frames):
        fid = i + 1
        d = disparity(l, r, cfg)
        deriv, hist = po.naive_derivative(d)
        running += hist                                   # mergeHistogram, planeseg.cu:144-158
        if provider == "histogram_peak" and fid % update == 1:   # planeseg.cu:381
            snap = running.astype(np.int32).copy()
            if fid % (update * reset) == 1:               # planeseg.cu:391-394
                running[:] = 0
            _, params = po.histogram_peak_update(snap, params)
        planes = po.classify(deriv, *params[2:6])
        out.append(dict(disparity=d, derivative=deriv, hist=hist, planes=planes, params=list(params)))
    return out


def sp_sequence(frames, cfg, provider="histogram_peak", static=(1, 30, -3, 1), update=30, reset=10,
                initial=18, steady=6, sp_reset=64, block=12, sp_kwargs=None):
    sp_kwargs = sp_kwargs or {}
    H, W = frames[0][0].shape[:2]
    labels, nlab = po.block_init(W, H, block, block)
    running = None
    params = [0, 0, 0, 0, 0, 0] if provider == "histogram_peak" else [0, 0] + list(static)
    out = []
    for i, (l, r) in enumerate(frames):
        fid = i + 1
        d = disparity(l, r, cfg)
        deriv, hist2 = po.derivative(d)
        its = initial if (fid == 1 or fid % sp_reset == 0) else steady      # superpixels.cu:93
        if fid % sp_reset == 0:                                               # superpixels.cu:105-113
            labels, nlab = po.block_init(W, H, block, block)
        labels, _, _ = po.sp_relax(labels, nlab, po.ycrcb(l), deriv, its, **sp_kwargs)
        hv = hist2[:, 0].astype(np.int64)                                     # sp_planeseg.cu:358-359
        if running is None:                                                   # :364-366
            running = np.zeros(256, np.int64)
            hist = hv.copy()
        else:
            running += hv
            hist = running.copy()
        if fid % (update * reset) == 1:                                       # :372-375
            running[:] = 0
        if provider == "histogram_peak" and fid % update == 1:               # :378-382
            _, params = po.histogram_peak_update(hist.astype(np.int32), params)
        unsm, planes = po.sp_planeseg(deriv, labels, nlab, *params[2:6])
        out.append(dict(disparity=d, derivative=deriv, labels=labels.copy(), planes=planes, unsm=unsm,
                        params=list(params)))
    return out
