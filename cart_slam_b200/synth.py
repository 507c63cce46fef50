"""Deterministic synthetic stereo generator.

Integer-only (splitmix64 hashing + integer bilinear value noise), so the same frames come out on any
host.  A frame is a rectified BGR stereo pair of a piecewise-planar scene: a road plane below the
horizon, fronto-parallel slabs above it and a few constant-disparity boxes.  The right image of
frame `f` is a window into a "world" texture that advances 3 px per frame; the left image is the
right image warped by the ground-truth disparity: left(x, y) = right(x - d(x, y), y).

This is input generation only - it contains nothing of the algorithm under test.
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _hash2(seed: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        k = (np.asarray(a, dtype=np.int64).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ (
            np.asarray(b, dtype=np.int64).astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
        )
        return splitmix64(k ^ np.uint64(seed & 0xFFFFFFFFFFFFFFFF))


def _value_noise(seed: int, xs: np.ndarray, ys: np.ndarray, cell: int) -> np.ndarray:
    """Integer bilinear lattice noise in 0..255 at integer coordinates (xs, ys) (broadcastable)."""
    xs = np.asarray(xs, dtype=np.int64)
    ys = np.asarray(ys, dtype=np.int64)
    ix, fx = np.floor_divide(xs, cell), np.mod(xs, cell)
    iy, fy = np.floor_divide(ys, cell), np.mod(ys, cell)
    v00 = (_hash2(seed, ix, iy) & np.uint64(0xFF)).astype(np.int64)
    v10 = (_hash2(seed, ix + 1, iy) & np.uint64(0xFF)).astype(np.int64)
    v01 = (_hash2(seed, ix, iy + 1) & np.uint64(0xFF)).astype(np.int64)
    v11 = (_hash2(seed, ix + 1, iy + 1) & np.uint64(0xFF)).astype(np.int64)
    top = v00 * (cell - fx) + v10 * fx
    bot = v01 * (cell - fx) + v11 * fx
    return (top * (cell - fy) + bot * fy) // (cell * cell)


def world_texture(seed: int, x0: int, width: int, height: int) -> np.ndarray:
    """3-octave value noise (cells 4/16/64), contrast-stretched to 0..255. Returns uint8 [height, width]."""
    xs = np.arange(x0, x0 + width, dtype=np.int64)[None, :]
    ys = np.arange(height, dtype=np.int64)[:, None]
    n = (
        _value_noise(seed + 1, xs, ys, 4) * 2
        + _value_noise(seed + 2, xs, ys, 16) * 3
        + _value_noise(seed + 3, xs, ys, 64) * 3
    ) // 8
    n = (n - 128) * 2 + 128  # stretch
    return np.clip(n, 0, 255).astype(np.uint8)


def ground_truth_disparity(W: int, H: int, D: int, min_disp: int, seed: int, frame: int) -> np.ndarray:
    """Integer-pixel ground-truth disparity on the left image grid, int32 [H, W]."""
    h0 = int(0.45 * H)
    y = np.arange(H, dtype=np.int64)[:, None]
    x = np.arange(W, dtype=np.int64)[None, :]
    span = min(D - 16, (45 * max(1, H - h0)) // 100)  # at most 0.45 px of disparity per row
    road = min_disp + (y - h0) * span // max(1, H - h0)
    slabs = min_disp + 8 + 16 * ((x // 160) % 4)
    d = np.where(y >= h0, np.broadcast_to(road, (H, W)), np.broadcast_to(slabs, (H, W))).astype(np.int64)
    # 6 constant-disparity boxes, re-seeded every 10 frames
    bs = splitmix64(np.arange(6 * 5, dtype=np.uint64) + np.uint64((seed * 1000003 + frame // 10) & 0xFFFFFFFF))
    bs = bs.reshape(6, 5)
    for b in bs:
        bw = 40 + int(b[0] % np.uint64(max(1, W // 6)))
        bh = 30 + int(b[1] % np.uint64(max(1, H // 4)))
        bx = int(b[2] % np.uint64(max(1, W - bw)))
        by = int(b[3] % np.uint64(max(1, H - bh)))
        bd = min_disp + 4 + int(b[4] % np.uint64(max(1, D - 24)))
        d[by : by + bh, bx : bx + bw] = np.maximum(d[by : by + bh, bx : bx + bw], bd)
    return np.clip(d, min_disp, min_disp + D - 1).astype(np.int32)


class SyntheticSequence:
    """A KITTI/ZED-shaped synthetic stereo sequence. frame ids start at 1 like the reference's run ids."""

    def __init__(self, W: int, H: int, D: int, min_disp: int = 4, sequence_id: int = 0, n_frames: int = 16,
                 tint: bool = False, zero_patch: bool = True):
        self.W, self.H, self.D, self.min_disp = W, H, D, min_disp
        self.seed = 0xCA27 + sequence_id
        self.n_frames = n_frames
        self.tint = tint
        self.zero_patch = zero_patch
        self.margin = D + min_disp + 16
        self._world = world_texture(self.seed, -self.margin, W + 3 * (n_frames + 1) + self.margin, H)

    def _noise(self, frame: int, which: int) -> np.ndarray:
        xs = np.arange(self.W, dtype=np.int64)[None, :]
        ys = np.arange(self.H, dtype=np.int64)[:, None]
        h = _hash2(self.seed ^ (frame * 0x9E3779B97F4A7C15 + which), xs, ys)
        return (h % np.uint64(5)).astype(np.int64) - 2

    def frame(self, frame: int):
        """Returns (left_bgr, right_bgr, gt_disparity) for 1-based frame id."""
        W, H = self.W, self.H
        f = (frame - 1) % max(1, self.n_frames)
        off = self.margin + 3 * f
        gt = ground_truth_disparity(W, H, self.D, self.min_disp, self.seed, f)
        right = self._world[:, off : off + W].astype(np.int64)
        xs = np.arange(W, dtype=np.int64)[None, :] + off - gt
        left = np.take_along_axis(self._world, xs, axis=1).astype(np.int64)
        right = np.clip(right + self._noise(f, 1), 0, 255)
        left = np.clip(left + self._noise(f, 2), 1, 255)  # 0 is reserved for the L/R-check mask patch
        if self.zero_patch and (int(splitmix64(np.uint64(self.seed * 7919 + f)) % np.uint64(50)) == 0 or f == 3):
            py = int(splitmix64(np.uint64(self.seed + 31 * f)) % np.uint64(max(1, H - 32)))
            px = int(splitmix64(np.uint64(self.seed + 17 * f + 5)) % np.uint64(max(1, W - 32)))
            left[py : py + 32, px : px + 32] = 0
        lb = np.repeat(left[:, :, None], 3, axis=2)
        rb = np.repeat(right[:, :, None], 3, axis=2)
        if self.tint:
            xs0 = np.arange(W, dtype=np.int64)[None, :] + 3 * f
            ys0 = np.arange(H, dtype=np.int64)[:, None]
            tg = (_value_noise(self.seed + 11, xs0, ys0, 64) - 128) // 5
            tr = (_value_noise(self.seed + 12, xs0, ys0, 64) - 128) // 5
            for img in (lb, rb):
                img[:, :, 1] = np.clip(img[:, :, 1] + tg, 0, 255)
                img[:, :, 2] = np.clip(img[:, :, 2] + tr, 0, 255)
        return lb.astype(np.uint8), rb.astype(np.uint8), gt

    def batch(self, first: int, count: int):
        ls, rs = [], []
        for k in range(count):
            l, r, _ = self.frame(first + k)
            ls.append(l)
            rs.append(r)
        return np.stack(ls), np.stack(rs)
