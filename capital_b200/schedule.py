"""Host-side description of the multi-GPU communication schedule (mirrors capital_b200/csrc/dist.cu).

Pure functions, no GPU: used by the CPU (gloo) tests to run the same rank arithmetic with numpy operands, and as
documentation of who talks to whom.  Rank <-> (x, y, z) is topo::square layout 0 (topology.h:81-83).
"""
from __future__ import annotations


def rank_of(c: int, d: int, x: int, y: int, z: int) -> int:
    return y * c * d + x * c + z


def coords(c: int, d: int, rank: int):
    return (rank % (c * d)) // c, rank // (c * d), rank % c  # x, y, z


def summa_plan(c: int, d: int, rank: int) -> dict:
    """One distributed product C += X^T Y on the cubic grid (c == d), as executed by dist.cu::product():
    the X block multiplied on (x,y,z) is the local X window of (y,z,z), the Y block the local window of (x,z,z)
    (util::transpose + row/column MPI_Bcast in the reference, summa.hpp:185,193); partial products are summed over
    the depth communicator (summa.hpp:236)."""
    assert c == d
    x, y, z = coords(c, d, rank)
    plan = {
        "src_x": rank_of(c, d, y, z, z),
        "src_y": rank_of(c, d, x, z, z),
        "send_x_to": [rank_of(c, d, xx, x, z) for xx in range(d)] if y == z else [],
        "send_y_to": [rank_of(c, d, x, yy, z) for yy in range(d)] if y == z else [],
        "depth_group": [rank_of(c, d, x, y, zz) for zz in range(c)],
        "slice_group": [rank_of(c, d, xx, yy, z) for yy in range(d) for xx in range(d)],  # slice rank = x + d*y
        "transpose_partner": rank_of(c, d, y, x, z),
    }
    plan["send_x_to"] = [r for r in plan["send_x_to"] if r != rank]
    plan["send_y_to"] = [r for r in plan["send_y_to"] if r != rank]
    return plan


def product_slices(c: int, d: int, rank: int, k_local: int) -> list:
    """Generalised schedule of dist.cu::product() for grids with c | d or d | c (the reference has c == d only).
    The contraction index splits into d owner classes (k mod d = kb); layer z takes the classes kb = z (mod c) when c <= d,
    and when c > d (2x1x1) the single class is cut into c/d row chunks of the local window.  Returns, for this rank, a list of
    dicts {kb, rows=(r0, r1), src_x, src_y, send_x_to, send_y_to}; partial products are then summed over the depth group."""
    assert c % d == 0 or d % c == 0
    x, y, z = coords(c, d, rank)
    nslices = max(c, d)
    nchunk = c // d if c > d else 1
    out = []
    for sl in range(z, nslices, c):
        kb, chunk = sl % d, sl // d
        r0, r1 = 0, k_local
        if nchunk > 1:
            r0 = (k_local * chunk // nchunk) & ~1
            r1 = k_local if chunk + 1 == nchunk else (k_local * (chunk + 1) // nchunk) & ~1
        src = y == kb
        out.append({
            "kb": kb, "rows": (r0, r1),
            "src_x": rank_of(c, d, y, kb, z), "src_y": rank_of(c, d, x, kb, z),
            "send_x_to": [r for r in (rank_of(c, d, t, x, z) for t in range(d)) if r != rank] if src else [],
            "send_y_to": [r for r in (rank_of(c, d, x, t, z) for t in range(d)) if r != rank] if src else [],
        })
    return out


# ---- single-GPU recursion: node splits and the block-wise output tiling (mirrors cholinv_local.cu) ----------------------
LEAF_MAX, BASECASE_MAX = 64, 512


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
