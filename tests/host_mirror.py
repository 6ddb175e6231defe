"""Test-only host mirror of the block-wise output plan of cholinv_local.cu (`block_done` calls): used by test_emission_tiling.py to
check on CPU that the emitted blocks tile both packed triangles exactly once.  Not part of the product."""
from __future__ import annotations

LEAF_MAX, BASECASE_MAX = 64, 512  # capital_b200/csrc/common.cuh


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def choose_split(o: int, n: int, bc: int, split: int, complete: bool) -> int:
    """cholinv_local.cu::choose_split for aligned buffers: the reference's rule s1 = n >> split above its base case
    (cholinv.hpp:92,107), 64-aligned halves below it, 0 when one kernel (cluster base case / leaf) takes the block."""
    if n > bc and (n >> split) >= split and (n >> split) > 0 and (n > LEAF_MAX or not complete):
        return n >> split
    if n <= LEAF_MAX:
        return 0
    if n <= BASECASE_MAX and n % 64 == 0 and complete and o % 2 == 0:
        return 0
    s1 = n >> 1
    s1 = _round_up(s1, LEAF_MAX) if n > 2 * LEAF_MAX else _round_up(s1, 2)
    return n >> 1 if s1 >= n else s1


def emission_blocks(n: int, bc: int, split: int, complete_inv: bool, zc_depth: int = 3) -> list:
    """Order and extent of the `block_done` calls of the experimental block-wise output (CAPITAL_ZC_OUT): tuples
    (which, r0, r1, c0, c1), which = 0 for R and 1 for Rinv, rows/columns half-open, clipped to the upper triangle by the kernel."""
    out = []

    def tri(o, m):
        out.append((0, o, o + m, o, o + m))
        out.append((1, o, o + m, o, o + m))

    def rec(o, m, complete, depth):
        s1 = choose_split(o, m, bc, split, complete)
        if s1 == 0:
            if depth <= zc_depth:
                tri(o, m)
            return
        if depth < zc_depth and not complete:
            out.append((1, o, o + s1, o + s1, o + m))  # the skipped inverse block: zeros
        rec(o, s1, True, depth + 1)
        if depth < zc_depth:
            out.append((0, o, o + s1, o + s1, o + m))
        rec(o + s1, m - s1, True, depth + 1)
        if complete and depth < zc_depth:
            out.append((1, o, o + s1, o + s1, o + m))
        if depth == zc_depth:
            tri(o, m)

    rec(0, n, bool(complete_inv), 0)
    return out
