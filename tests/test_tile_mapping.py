"""The kernels fetch reference-tile elements through a closed-form inverse of copyToShared
(cart_slam_b200/csrc/tile_ref.cuh).  Here it is evaluated on the CPU through the C ABI's debug entry
point and compared, element by element, with the oracle's literal simulation of
/root/reference/include/utils/cuda.cuh:59-191 - the automated version of the reference's own
(never called) sanity_check.cu:58-65 idea: an image filled with y*W+x."""
import numpy as np
import pytest

import cart_slam_b200 as cb
import pyoracle as po

CASES = [
    # (W, H, bdx, bdy, XB, YB, yPad, xPad, interp, alloc)  -- the five call sites of the reference
    ("derivative", 300, 290, 32, 32, 4, 4, 2, 2, True, 132 * 132),
    ("naive", 300, 290, 32, 32, 4, 4, 2, 0, True, 128 * 144),
    ("interpolate_r2", 150, 140, 16, 16, 4, 4, 1, 1, True, 66 * 66),
    ("interpolate_r3", 150, 140, 16, 16, 4, 4, 2, 2, True, 68 * 68),
    ("border", 150, 140, 16, 16, 4, 4, 1, 1, False, 72 * 72),
    ("derivative_small", 100, 60, 32, 32, 4, 4, 2, 2, True, 132 * 132),
    ("border_small", 50, 30, 16, 16, 4, 4, 1, 1, False, 72 * 72),
    ("border_exact", 128, 128, 16, 16, 4, 4, 1, 1, False, 72 * 72),
    ("derivative_exact", 256, 256, 32, 32, 4, 4, 2, 2, True, 132 * 132),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_closed_form_matches_literal_simulation(case):
    _, W, H, bdx, bdy, XB, YB, yPad, xPad, interp, alloc = case
    img = (np.arange(H, dtype=np.int32)[:, None] * W + np.arange(W, dtype=np.int32)[None, :]) % 30011 + 7
    tw, th = bdx * XB, bdy * YB
    nbx, nby = -(-W // tw), -(-H // th)
    undef = -12345
    checked = 0
    for by in range(nby):
        for bx in range(nbx):
            ref, _ = po.tile_i32(img, bx, by, bdx, bdy, XB, YB, yPad, xPad, interp, alloc_elems=alloc, undef=undef)
            R, S = ref.shape
            # sample every cell of the halo ring and a sparse grid inside (full check is O(10^5) ctypes calls)
            for r in range(R):
                ly = r - yPad
                dense_row = ly < 4 or ly >= th - 3 or ly in (th // 2,)
                for c in range(S):
                    lx = c - xPad
                    if not (dense_row or lx < 4 or lx >= tw - 3 or (lx % 17 == 0 and ly % 13 == 0)):
                        continue
                    got = cb.debug_ref_tile_i32(img, bx, by, bdx, bdy, XB, YB, yPad, xPad, interp, lx, ly,
                                                alloc_elems=alloc, undef=undef)
                    assert got == ref[r, c], (bx, by, lx, ly, got, int(ref[r, c]))
                    checked += 1
    assert checked > 1000


def test_row_shift_quirk_is_present():
    """for block rows >= 1 the tile holds the image shifted up by yPad rows."""
    W = H = 300
    img = np.arange(H, dtype=np.int32)[:, None] * 1000 + np.arange(W, dtype=np.int32)[None, :]
    ref, d = po.tile_i32(img, 0, 1, 32, 32, 4, 4, 2, 2, True, alloc_elems=132 * 132)
    # local (lx=5, ly=0) of block row 1 -> image row 128 + 0 + 2
    assert ref[0 + 2, 5 + 2] == (128 + 2) * 1000 + 5
    ref0, _ = po.tile_i32(img, 1, 0, 32, 32, 4, 4, 2, 2, True, alloc_elems=132 * 132)
    assert ref0[0 + 2, 5 + 2] == 0 * 1000 + 128 + 5  # block row 0 is not shifted
    # Q4: top halo row -1, column c <- tile row 1, column c mod 4
    assert ref0[1, 9 + 2] == 1 * 1000 + 128 + (9 % 4)
