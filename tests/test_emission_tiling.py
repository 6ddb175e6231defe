"""The block-wise output of the experimental CAPITAL_ZC_OUT path must write every slot of the packed triangles exactly once
(host mirror of the `block_done` calls in cholinv_local.cu: tests/host_mirror.py::emission_blocks)."""
import numpy as np
import pytest
import host_mirror as sch


@pytest.mark.parametrize("n,bc,split,ci,depth", [(16384, 512, 1, 0, 3), (16384, 512, 1, 1, 3), (8192, 512, 1, 1, 2), (9088, 568, 1, 1, 3),
                                                (2048, 128, 1, 0, 0), (4096, 4096, 1, 1, 3), (777, 97, 1, 1, 3), (6000, 750, 2, 0, 2), (640, 160, 1, 1, 5)])
def test_blocks_tile_both_triangles_exactly_once(n, bc, split, ci, depth):
    blocks = sch.emission_blocks(n, bc, split, ci, depth)
    step = max(1, n // 512)  # sample a grid of slots (plus the last row/column) -- the blocks are axis-aligned rectangles
    idx = np.unique(np.concatenate([np.arange(0, n, step), [n - 1]]))
    cover = np.zeros((2, idx.size, idx.size), dtype=np.int32)
    rows, cols = idx[:, None], idx[None, :]
    for which, r0, r1, c0, c1 in blocks:
        assert 0 <= r0 < r1 <= n and 0 <= c0 < c1 <= n and r0 <= c0  # nothing strictly below the diagonal blocks
        cover[which] += ((rows >= r0) & (rows < r1) & (cols >= c0) & (cols < c1) & (rows <= cols))
    upper = rows <= cols
    assert np.array_equal(cover[0][upper], np.ones(upper.sum(), dtype=np.int32))
    assert np.array_equal(cover[1][upper], np.ones(upper.sum(), dtype=np.int32))
    assert cover[:, ~upper].sum() == 0
    # exact element count as well: the sampled grid could miss a one-column gap
    total = [0, 0]
    for which, r0, r1, c0, c1 in blocks:
        c = np.arange(c0, c1)
        total[which] += int(np.clip(np.minimum(r1, c + 1) - r0, 0, None).sum())
    assert total == [n * (n + 1) // 2] * 2


def test_split_rule_matches_the_reference_above_the_base_case():
    # cholinv.hpp:92,107: s1 = n >> split while n > bc; below it 64-aligned halves down to the cluster / leaf kernels
    assert sch.choose_split(0, 16384, 512, 1, False) == 8192
    assert sch.choose_split(0, 1024, 512, 1, True) == 512
    assert sch.choose_split(0, 512, 512, 1, True) == 0      # cluster base-case kernel
    assert sch.choose_split(0, 64, 512, 1, True) == 0       # leaf kernel
    assert sch.choose_split(0, 568, 568, 1, True) == 320    # not a multiple of 64: split into 64-aligned halves
    assert sch.choose_split(0, 6000, 750, 2, False) == 1500
