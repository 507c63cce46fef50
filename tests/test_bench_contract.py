"""bench.py pieces that can be checked without a GPU: the committed ncu summary the roofline's `traffic` figure comes
from, the CPU reference arm's JSON line (bounded sample), and the workload table against BASELINE.json."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_ncu_traffic_matches_the_algorithmic_bytes_of_one_path():
    """roofline.traffic comes from the newest committed `ncu --set full` summary of the aggregation kernels - and only
    while that summary was captured from the csrc/sgm.cu that is in the tree (hash stamped by tools/ncu_raw_summary.py)."""
    t, note = bench.ncu_traffic_per_launch()
    assert note
    if t is None:
        # csrc/sgm.cu changed after the last capture: bench.py reports `traffic: null` with this reason instead of a
        # number that may no longer describe the kernels (test_a_stale_ncu_summary_is_not_reported)
        assert "stale" in note or "no source hash" in note or "no ncu summary" in note, note
        pytest.skip(f"roofline.traffic not reported: {note}")
    W, H, D, B = 1242, 375, 128, 64
    algorithmic = B * (W * H * D + 2 * 4 * W * H)  # read both census images, write one u8 volume
    assert abs(t / algorithmic - 1.0) < 0.03, (t, algorithmic)  # no wasted re-reads


def test_a_stale_ncu_summary_is_not_reported(tmp_path, monkeypatch):
    import shutil
    root = tmp_path / "repo"
    (root / "profiles").mkdir(parents=True)
    (root / "cart_slam_b200" / "csrc").mkdir(parents=True)
    shutil.copy(os.path.join(ROOT, "cart_slam_b200", "csrc", "sgm.cu"), root / "cart_slam_b200" / "csrc" / "sgm.cu")
    launch = {"kernel": "aggregate_vertical_kernel", "launch__grid_size []": 10.0, "dram__bytes_read.sum [Mbyte]": 100.0,
              "dram__bytes_write.sum [Gbyte]": 3.9}
    monkeypatch.setattr(bench, "ROOT", str(root))
    json.dump({"source_sha16": "0" * 16, "launches": [launch]}, open(root / "profiles" / "r09_ncu_aggregate_batch64.json", "w"))
    t, note = bench.ncu_traffic_per_launch()
    assert t is None and "stale" in note
    json.dump({"launches": [launch]}, open(root / "profiles" / "r09_ncu_aggregate_batch64.json", "w"))
    assert bench.ncu_traffic_per_launch()[0] is None
    json.dump({"source_sha16": bench._sha16(str(root / "cart_slam_b200" / "csrc" / "sgm.cu")), "launches": [launch]},
              open(root / "profiles" / "r09_ncu_aggregate_batch64.json", "w"))
    t, note = bench.ncu_traffic_per_launch()
    assert t == pytest.approx(4.0e9)


def test_workloads_are_the_baseline_configs():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "1242" in base["metric"] and "128 disp" in base["metric"]
    k = bench.WORKLOADS["kitti"]
    assert (k["W"], k["H"], k["D"], k["frames"]) == (1242, 375, 128, 1000)
    z, u = bench.WORKLOADS["zed"], bench.WORKLOADS["4k"]
    assert (z["W"], z["H"], z["D"]) == (1280, 720, 256) and (u["W"], u["H"], u["D"], u["paths"]) == (3840, 2160, 256, 8)


@pytest.mark.timeout(600)
def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "1"], capture_output=True, text=True, timeout=580, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["metric"] == bench.METRIC and d["higher_is_better"] is True


def test_committed_bench_line_carries_the_contract_keys():
    """The newest committed default bench line (profiles/r*_bench_default_n1.json): base-contract keys, the roofline of
    the aggregation kernels with live timing and matching ncu traffic, the aggregation + WTA group of SURVEY 8(d) with
    its algorithmic bytes 2 P W H D + (8 P + 27) W H per frame, the CPU baseline and the end-to-end object."""
    import glob
    lines = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_default_n1.json")))
    d = json.loads(open(lines[-1]).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["vs_baseline"] is None
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["e2e"]["h2d_bytes_per_step"] == 2 * 1000 * 375 * 1242 * 3 and d["e2e"]["d2h_bytes_per_step"] == 1000 * 375 * 1242
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    W, H, D, B, P = 1242, 375, 128, 64, 4
    assert r["algorithmic_bytes_per_launch"] == B * (2 * 4 * W * H + W * H * D)
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["launch_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    if r["traffic"] is not None:
        assert abs(r["traffic"] / r["algorithmic_bytes_per_launch"] - 1.0) < 0.03
    g = r["aggregation_plus_wta"]
    assert g["algorithmic_bytes"] == B * (2 * P * W * H * D + (8 * P + 27) * W * H)
    assert abs(g["frac"] - g["algorithmic_bytes"] / (g["ms"] * 1e-3) / 1e9 / r["peak"]) < 1e-9
    assert abs(g["ms"] - (g["aggregation_ms"] + g["wta_post_interp_ms"])) < 1e-9
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["sample"]
