"""The C++ module layer (libcartb200_host.so): JSON schema of the reference, error behaviour (CPU),
and - on a GPU - the same results as the batched sequence runner and the oracle pipeline."""
import json
import os

import numpy as np
import pytest

import cart_slam_b200 as cb
from cart_slam_b200 import host
from cart_slam_b200.synth import SyntheticSequence

REF_CONFIGS = "/root/reference/config/modules"


def _frames(W, H, D, n, tint=False):
    seq = SyntheticSequence(W, H, D, n_frames=n, tint=tint)
    fr = [seq.frame(i + 1)[:2] for i in range(n)]
    return np.stack([f[0] for f in fr]), np.stack([f[1] for f in fr]), fr


def test_unknown_and_out_of_scope_modules_are_rejected():
    L = np.zeros((1, 32, 64, 3), np.uint8)
    with pytest.raises(host.HostError, match="Unknown module type nonsense"):
        host.run_config([{"type": "nonsense"}], L, L)
    with pytest.raises(host.HostError, match="outside the scope"):
        host.run_config([{"type": "optflow"}], L, L)
    with pytest.raises(host.HostError, match="not an array"):
        host.run_config({"type": "disparity"}, L, L)
    with pytest.raises(host.HostError, match="No modules"):
        host.run_config([{"type": "optflow"}, {"type": "disparity_visualization"}], L, L, skip_out_of_scope=True)


def test_static_provider_requires_its_keys():
    L = np.zeros((1, 32, 64, 3), np.uint8)
    with pytest.raises(host.HostError, match="horizontal_range_m"):
        host.run_config([{"type": "disparity_planeseg", "parameter_provider": {"type": "static"}}], L, L)
    with pytest.raises(host.HostError, match="Unknown parameter provider"):
        host.run_config([{"type": "disparity_planeseg", "parameter_provider": {"type": "magic"}}], L, L)


@pytest.mark.skipif(not os.path.isdir(REF_CONFIGS), reason="reference tree not present")
def test_reference_config_files_parse_with_the_same_schema():
    """Every module type named by the reference's shipped pipelines is known to the parser (in or out of scope)."""
    src = open(os.path.join(os.path.dirname(cb.library_path()), "host", "src", "config.cpp")).read()
    for fn in sorted(os.listdir(REF_CONFIGS)):
        if not fn.endswith(".json"):
            continue
        for m in json.load(open(os.path.join(REF_CONFIGS, fn))):
            assert f'"{m["type"]}"' in src, (fn, m["type"])


@pytest.mark.gpu
def test_missing_dependency_is_reported():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    L = np.zeros((1, 32, 64, 3), np.uint8)
    with pytest.raises(host.HostError, match="requires data disparity"):
        host.run_config([{"type": "disparity_derivative"}], L, L)


@pytest.mark.gpu
@pytest.mark.parametrize("sequential", [True, False])
def test_naive_json_pipeline_matches_sequence_runner(sequential):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    W, H, D, n = 192, 96, 64, 14
    L, R, _ = _frames(W, H, D, n)
    modules = [
        {"type": "disparity", "num_disparities": D, "smoothing_radius": 2, "smoothing_iterations": 1},
        {"type": "disparity_planeseg", "parameter_provider": {"type": "histogram_peak"}, "update_interval": 5, "reset_interval": 2},
        {"type": "disparity_planeseg_visualization", "show_histogram": True},
    ]
    out = host.run_config(modules, L, R, skip_out_of_scope=True, sequential=sequential, want_disparity=True)
    cfg = cb.Config(W, H, max_batch=4, num_disparities=D, smoothing_radius=2, smoothing_iterations=1, enable_superpixels=False)
    opts = cb.SequenceOptions(pipeline=0, provider=1, update_interval=5, reset_interval=2)
    with cb.Context(cfg) as ctx:
        planes, disp = ctx.run_sequence_host(opts, L, R, want_disparity=True)
    assert np.array_equal(out["disparity"], disp)
    if sequential:  # the running histogram is order dependent; only the in-order schedule is canonical
        assert np.array_equal(out["planes"], planes)
    else:
        assert (out["planes"] != 255).all()


@pytest.mark.gpu
@pytest.mark.parametrize("workers,sequential", [(None, True), (1, False), (2, False), (3, True)])
def test_superpixel_json_pipeline_matches_sequence_runner(workers, sequential, monkeypatch):
    """Also the scheduler's liveness: the shipped module order is not a dependency order (superpixels first), and with
    1-3 pool workers and up to 12 frames in flight every worker used to end up waiting for a module that could not start."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    if workers:
        monkeypatch.setenv("CARTB200_HOST_WORKERS", str(workers))
    W, H, D, n = 160, 64, 64, 19
    L, R, _ = _frames(W, H, D, n, tint=True)
    # kitti-planeseg.json shape (module order as shipped: superpixels first, dependencies resolve the order)
    modules = [
        {"type": "superpixels", "initial_iterations": 6, "iterations": 3, "block_size": 8, "reset_iterations": 8},
        {"type": "optflow"},
        {"type": "disparity", "num_disparities": D, "smoothing_radius": 2, "smoothing_iterations": 1},
        {"type": "disparity_derivative"},
        {"type": "depth"},
        {"type": "superpixel_disparity_planeseg", "parameter_provider": {"type": "histogram_peak"},
         "update_interval": 5, "reset_interval": 2, "use_temporal_smoothing": True},
        {"type": "bev_planeseg_visualization"},
    ]
    out = host.run_config(modules, L, R, skip_out_of_scope=True, sequential=sequential, want_labels=True, want_disparity=True)
    cfg = cb.Config(W, H, max_batch=3, num_disparities=D, smoothing_radius=2, smoothing_iterations=1, sp_block_size=8)
    opts = cb.SequenceOptions(pipeline=1, provider=1, update_interval=5, reset_interval=2, sp_initial_iterations=6,
                              sp_iterations=3, sp_reset_iterations=8)
    with cb.Context(cfg) as ctx:
        planes, disp = ctx.run_sequence_host(opts, L, R, want_disparity=True)
    assert np.array_equal(out["disparity"], disp)
    # same kernels, same order of operations per frame: the chunked runner and the frame-by-frame modules agree exactly
    assert np.array_equal(out["planes"], planes)


@pytest.mark.gpu
def test_depth_module_through_the_json_pipeline():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import pyoracle as po

    W, H, D, n = 192, 96, 64, 3
    L, R, fr = _frames(W, H, D, n)
    Q = np.eye(4, dtype=np.float32)
    Q[0, 3], Q[1, 3], Q[2, 2], Q[2, 3], Q[3, 2], Q[3, 3] = -96.0, -48.0, 0, 400.0, -2.0, 0.1
    out = host.run_config([{"type": "disparity", "num_disparities": D, "smoothing_radius": 2, "smoothing_iterations": 1},
                           {"type": "depth"}], L, R, Q=Q, want_depth=True, want_disparity=True)
    for i in range(n):
        o = po.depth(out["disparity"][i], Q)
        fin = np.isfinite(o)
        assert np.allclose(out["depth"][i][fin], o[fin], rtol=1e-6, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("sequential", [True, False])
def test_temporal_smoothing_through_the_json_pipeline(sequential):
    """use_temporal_smoothing with an externally supplied flow ("external_optflow" stands in for the
    NVOFA module), naive and superpixel planeseg, against the oracle's restatement of the module logic."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import ref_pipeline as rp

    W, H, D, n = 160, 64, 64, 9
    L, R, fr = _frames(W, H, D, n, tint=True)
    flow = {i + 1: rp.constant_flow(H, W, 3.0, -0.5) for i in range(n)}
    static = {"type": "static", "horizontal_range_min": 1, "horizontal_range_max": 30, "vertical_range_min": -3,
              "vertical_range_max": 1}
    cfg = dict(D=D, radius=2, iters=1)
    # naive planeseg, distance 2
    modules = [
        {"type": "external_optflow", "flow_x": 3.0, "flow_y": -0.5},
        {"type": "disparity", "num_disparities": D, "smoothing_radius": 2, "smoothing_iterations": 1},
        {"type": "disparity_planeseg", "parameter_provider": static, "use_temporal_smoothing": True, "temporal_smoothing_distance": 2},
    ]
    out = host.run_config(modules, L, R, sequential=sequential, want_disparity=True)
    ref = rp.naive_sequence(fr, cfg, provider="static", static=(1, 30, -3, 1), temporal=dict(distance=2, flow=flow))
    for i in range(n):
        assert np.array_equal(out["disparity"][i], ref[i]["disparity"]), i
        assert np.array_equal(out["planes"][i], ref[i]["planes"]), i
    assert any((ref[i]["planes"] != ref[i]["unsm"]).any() for i in range(1, n))
    # superpixel planeseg, default distance 3 (kitti-planeseg.json shape)
    modules = [
        {"type": "superpixels", "initial_iterations": 5, "iterations": 2, "block_size": 8, "reset_iterations": 8},
        {"type": "external_optflow", "flow_x": 3.0, "flow_y": -0.5},
        {"type": "disparity", "num_disparities": D, "smoothing_radius": 2, "smoothing_iterations": 1},
        {"type": "disparity_derivative"},
        {"type": "superpixel_disparity_planeseg", "parameter_provider": static, "use_temporal_smoothing": True},
    ]
    out = host.run_config(modules, L, R, sequential=sequential, want_labels=True)
    ref = rp.sp_sequence(fr, cfg, provider="static", static=(1, 30, -3, 1), initial=5, steady=2, sp_reset=8, block=8,
                         temporal=dict(distance=3, flow=flow))
    if not sequential:  # the superpixel warm start depends on the order frames reach the module (as in the reference)
        assert (out["planes"] <= 2).all()
        return
    for i in range(n):
        assert np.array_equal(out["labels"][i], ref[i]["labels"]), i
        assert np.array_equal(out["planes"][i], ref[i]["planes"]), i


def test_temporal_smoothing_needs_a_flow_provider_and_a_sane_distance():
    L = np.zeros((1, 32, 64, 3), np.uint8)
    static = {"type": "static", "horizontal_range_min": 1, "horizontal_range_max": 30, "vertical_range_min": -3,
              "vertical_range_max": 1}
    with pytest.raises(host.HostError, match="temporal_smoothing_distance"):
        host.run_config([{"type": "external_optflow"}, {"type": "disparity_planeseg", "parameter_provider": static,
                                                        "use_temporal_smoothing": True, "temporal_smoothing_distance": 9}], L, L)
