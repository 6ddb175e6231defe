"""Selection logic of the base-case sweep (capital_b200/autotune.py; reference: autotune/cholesky/cholinv/tune.cpp:239-253) on CPU."""
import os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from capital_b200 import autotune as at
from oracle import capital_oracle as co


def test_configurations_follow_the_reference_base_case_rule():
    # config 2 of BASELINE.json: n = 16384 on one GPU, bcMultiplier -7 .. -3  ->  base cases 128 .. 2048 (cholinv.hpp:15-18)
    cfgs = at.configurations(16384, 1, 1, -7, 5)
    assert [c["bc_dim"] for c in cfgs] == [128, 256, 512, 1024, 2048]
    assert all(c["bc_dim"] == co.bc_dimension(16384, 1, 1, c["bc_mult_dim"]) for c in cfgs)
    # 2x2x2 grid, local 32768: multiplier -4 is BASELINE's b = 1024
    cfgs = at.configurations(32768, 2, 2, -5, 3)
    assert [(c["bc_mult_dim"], c["bc_dim"]) for c in cfgs] == [(-5, 512), (-4, 1024), (-3, 2048)]
    # multipliers that clamp to the same dimension are run once (bc <= 1 -> bc = 1 -> dimension d * local)
    cfgs = at.configurations(64, 1, 1, 0, 4)
    assert [c["bc_dim"] for c in cfgs] == [64] and cfgs[0]["k"] == 0


def test_sweep_and_best_pick_the_fastest_valid_configuration():
    cfgs = at.configurations(16384, 1, 1, -7, 5)
    model = {128: 80.0, 256: 70.0, 512: 64.0, 1024: 64.0, 2048: 75.0}
    calls = []

    def time_ms(cfg):
        calls.append(cfg["bc_dim"])
        return model[cfg["bc_dim"]] + 0.25 * (len(calls) % 3)  # jitter: the median must absorb it

    rows = at.sweep(time_ms, cfgs, num_iter=3, warmup=1)
    assert len(calls) == 5 * 4 and [r["samples"] for r in rows] == [3] * 5
    assert all(r["ms_min"] <= r["ms_median"] for r in rows)
    assert at.best(rows, "ms_min")["bc_dim"] == 1024  # tie on time -> the larger base case
    rows[3]["ms_median"] = rows[3]["ms_min"] = float("inf")  # what tune_cholinv does to a configuration whose residual fails
    assert at.best(rows)["bc_dim"] == 512


def test_grid_depth_matches_the_library_grids():
    assert [at.grid_depth(w) for w in (1, 2, 4, 8, 27)] == [1, 2, 1, 2, 3]
    assert at.grid_depth(8, 2) == 1
