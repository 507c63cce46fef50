"""Disk -> plane labels through the C++ module layer with the on-disk KITTI source.
Writes a KITTI-shaped synthetic sequence as PNG files (image_2 / image_3 / calib.txt), then times
cartb200_host_run_source over it: PNG decode (pool of threads, pinned ring) -> upload -> kitti-planeseg.json modules.
Also times the decoder alone.  Run on the GPU box:  python tools/disk_bench.py [--frames 96]"""
import argparse
import json
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cart_slam_b200 import host  # noqa: E402
from cart_slam_b200.synth import SyntheticSequence  # noqa: E402


def main():
    import cv2
    from test_sources import CALIB

    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=96)
    ap.add_argument("--distinct", type=int, default=8)
    ap.add_argument("--repeat", type=int, default=1, help="repetitions of every disk -> planes measurement (all values are reported)")
    ap.add_argument("--quick", action="store_true", help="decoder timing + disk -> planes with the default decode threads only")
    args = ap.parse_args()
    W, H, D = 1242, 375, 128
    root = "/tmp/cartb200_disk_bench"
    shutil.rmtree(root, ignore_errors=True)
    d = os.path.join(root, "sequences", "00")
    os.makedirs(os.path.join(d, "image_2"))
    os.makedirs(os.path.join(d, "image_3"))
    open(os.path.join(d, "calib.txt"), "w").write(CALIB)
    seq = SyntheticSequence(W, H, D, n_frames=args.distinct, tint=True)
    enc = [[cv2.imencode(".png", im)[1].tobytes() for im in seq.frame(1 + i)[:2]] for i in range(args.distinct)]
    for i in range(args.frames):
        for cam, data in zip((2, 3), enc[i % args.distinct]):
            open(os.path.join(d, f"image_{cam}", f"{i:06d}.png"), "wb").write(data)
    png_bytes = sum(len(b) for pair in enc for b in pair) / (2 * args.distinct)
    # the decoder alone, one thread
    p0 = os.path.join(d, "image_2", "000000.png")
    t0 = time.perf_counter()
    for _ in range(10):
        host.decode_png(p0)
    decode_ms = (time.perf_counter() - t0) / 10 / 2 * 1000  # decode_png decodes twice (size query + data)
    modules = [
        {"type": "superpixels", "initial_iterations": 24, "iterations": 8, "block_size": 12, "reset_iterations": 64},
        {"type": "disparity", "num_disparities": D, "smoothing_radius": 2, "smoothing_iterations": 1},
        {"type": "disparity_derivative"},
        {"type": "superpixel_disparity_planeseg", "parameter_provider": {"type": "histogram_peak"}},
    ]
    source = {"type": "kitti", "path": root, "sequence": 0}
    out = {"frames": args.frames, "png_bytes_per_image": png_bytes, "decode_ms_per_image_one_thread": decode_ms,
           "host_cores": os.cpu_count()}
    # the same module list fed from host memory (no decode): what the module layer itself sustains
    frames = [seq.frame(1 + (i % args.distinct))[:2] for i in range(args.distinct)]
    Lm = np.stack([frames[i % args.distinct][0] for i in range(args.frames)])
    Rm = np.stack([frames[i % args.distinct][1] for i in range(args.frames)])
    host.run_config(modules, Lm[:8], Rm[:8], sequential=False)
    for name, sq in (("memory_to_planes_fps_in_flight_12", False),) if args.quick else (("memory_to_planes_fps_in_flight_12", False), ("memory_to_planes_fps_sequential", True)):
        t0 = time.perf_counter()
        host.run_config(modules, Lm, Rm, sequential=sq)
        out[name] = args.frames / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    host.run_config(modules, Lm[:1], Rm[:1], sequential=True)
    out["setup_plus_one_frame_s"] = time.perf_counter() - t0
    out["png_inflate"] = "zlib" if os.environ.get("CARTB200_PNG_ZLIB", "0") not in ("", "0") else "own (cart/inflate.hpp)"
    for threads in ((None,) if args.quick else (1, 4, None)):
        if threads:
            os.environ["CARTB200_KITTI_DECODE_THREADS"] = str(threads)
        else:
            os.environ.pop("CARTB200_KITTI_DECODE_THREADS", None)
        host.run_source(source, modules, 8)  # warm-up
        vals = []
        for _ in range(max(1, args.repeat)):
            t0 = time.perf_counter()
            res = host.run_source(source, modules, args.frames)
            dt = time.perf_counter() - t0
            n = res["planes"].shape[0] if isinstance(res, dict) and "planes" in res else args.frames
            vals.append(n / dt)
        out[f"disk_to_planes_fps_decode_threads_{threads or 'default'}"] = vals[0] if args.repeat <= 1 else sorted(vals)
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "disk_bench.json"), "w"), indent=1)
    shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
