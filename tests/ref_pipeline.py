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


def temporal_history(fid, distance, unsm_by_id, flow_by_id):
    """previousPlanes / previousOpticalFlow lists of planeseg.cu:300-337 (same in sp_planeseg.cu:256-310):
    entry k = planes_unsmoothed of frame fid-(k+1) and optflow of frame fid-k."""
    pp, pf = [], []
    if fid <= 1:
        return pp, pf
    flow = flow_by_id[fid]
    for i in range(1, distance + 1):
        if fid - i <= 0:
            break
        pp.append(unsm_by_id[fid - i])
        pf.append(flow)
        flow = flow_by_id[fid - i] if (fid - i > 1 and len(pp) < distance) else None
    return pp, pf


def constant_flow(H, W, fx, fy):
    """ExternalOpticalFlowModule(flowX, flowY): S10.5, round to nearest."""
    f = np.empty((H, W, 2), np.int16)
    f[:, :, 0] = int(np.rint(fx * 32.0))
    f[:, :, 1] = int(np.rint(fy * 32.0))
    return f


def naive_sequence(frames, cfg, provider="histogram_peak", static=(1, 30, -3, 1), update=30, reset=10, temporal=None):
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
        if temporal:  # planeseg.cu:300-376; temporal = dict(distance=, flow={frame id: [H, W, 2] int16})
            unsm_by_id = {j + 1: o["unsm"] for j, o in enumerate(out)}
            pp, pf = temporal_history(fid, temporal["distance"], unsm_by_id, temporal["flow"])
            unsm, planes = po.classify_temporal(deriv, *params[2:6], pp, pf)
        else:
            planes = po.classify(deriv, *params[2:6])
            unsm = planes
        out.append(dict(disparity=d, derivative=deriv, hist=hist, planes=planes, unsm=unsm, params=list(params)))
    return out


def sp_sequence(frames, cfg, provider="histogram_peak", static=(1, 30, -3, 1), update=30, reset=10,
                initial=18, steady=6, sp_reset=64, block=12, sp_kwargs=None, temporal=None):
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
        if temporal:  # sp_planeseg.cu:256-345
            unsm_by_id = {j + 1: o["unsm"] for j, o in enumerate(out)}
            pp, pf = temporal_history(fid, temporal["distance"], unsm_by_id, temporal["flow"])
            unsm, planes = po.sp_planeseg_temporal(deriv, labels, nlab, *params[2:6], pp, pf)
        else:
            unsm, planes = po.sp_planeseg(deriv, labels, nlab, *params[2:6])
        out.append(dict(disparity=d, derivative=deriv, labels=labels.copy(), planes=planes, unsm=unsm,
                        params=list(params)))
    return out
