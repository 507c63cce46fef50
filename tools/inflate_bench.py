"""Decode time of the PNG reader's own zlib-stream decoder against zlib's inflate on KITTI-sized PNG files (CPU only).
   python tools/inflate_bench.py"""
import os
import struct
import sys
import time
import zlib

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cart_slam_b200 import host  # noqa: E402
from cart_slam_b200.synth import SyntheticSequence  # noqa: E402


def idat_of(png: bytes) -> bytes:
    pos, out = 8, b""
    while pos < len(png):
        n, typ = struct.unpack(">I4s", png[pos:pos + 8])
        if typ == b"IDAT":
            out += png[pos + 8:pos + 8 + n]
        pos += 12 + n
    return out


def main():
    W, H = 1242, 375
    seq = SyntheticSequence(W, H, 128, n_frames=2, tint=True)
    synth = seq.frame(1)[0]
    rng = np.random.default_rng(0)
    # a smooth "photo-like" image: low-pass noise + gradients (compresses like camera images do, 2-3x)
    base = cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3)).astype(np.float32), (0, 0), 3.0)
    photo = np.clip((base - base.mean()) * 6 + 128 + rng.normal(0, 2.0, (H, W, 3)), 0, 255).astype(np.uint8)
    rows = []
    for name, img in (("synthetic frame (noise texture)", synth), ("photo-like", photo)):
        for level in (1, 3, 9):
            ok, buf = cv2.imencode(".png", img, [cv2.IMWRITE_PNG_COMPRESSION, level])
            stream = idat_of(buf.tobytes())
            raw = (W * 3 + 1) * H
            reps = 30
            assert host.inflate(stream, raw).tobytes() == zlib.decompress(stream)
            t = time.perf_counter()
            host.inflate(stream, raw, repeat=reps)
            own = (time.perf_counter() - t) / reps
            t = time.perf_counter()
            for _ in range(reps):
                zlib.decompress(stream)
            ref = (time.perf_counter() - t) / reps
            rows.append((name, level, len(stream), own * 1e3, ref * 1e3))
            print(f"{name:32s} png level {level}: {len(stream) / 1e6:5.2f} MB -> {raw / 1e6:4.2f} MB   own {own * 1e3:6.2f} ms "
                  f"({raw / own / 1e6:6.0f} MB/s)   zlib {zlib.ZLIB_RUNTIME_VERSION} {ref * 1e3:6.2f} ms ({raw / ref / 1e6:6.0f} MB/s)   x{ref / own:4.2f}")
    return rows


if __name__ == "__main__":
    main()
