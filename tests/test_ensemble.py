"""Ensemble layer: realization -> rank sharding, the final all-gather of the per-realization statistic rows
(world_size-2 gloo run on CPU), the reference's log-log regression, and — on the GPU — concurrent realizations giving exactly the
results of the same realizations run one by one."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from mcac_b200 import ensemble as ens

ROOT = Path(__file__).resolve().parent.parent


def test_round_robin_shards_cover_every_realization_once():
    for n, world in [(1024, 8), (1024, 1), (10, 4), (3, 8), (0, 2)]:
        seen = sorted(k for r in range(world) for k in ens.shard(n, r, world))
        assert seen == list(range(n))
        assert max(len(ens.shard(n, r, world)) for r in range(world)) - min(len(ens.shard(n, r, world)) for r in range(world)) <= 1
    assert ens.seeds(1000, ens.shard(10, 1, 4)) == [1001, 1005, 1009]


def test_linreg_follows_the_reference_formulas():
    """src/tools/tools.cpp:126-157 restated with numpy on a known power law: Np = kf (Dg/Dp)^Df."""
    rng = np.random.default_rng(0)
    x = rng.uniform(1.0, 30.0, 200)
    y = 1.3 * x ** 1.78
    ok, a, b, r = ens.linreg(x, y)
    assert ok and abs(a - 1.78) < 1e-12 and abs(np.exp(b) - 1.3) < 1e-12
    lx, ly = np.log(x), np.log(y)
    n = len(x)
    want_r = (np.sum(lx * ly) - lx.sum() * ly.sum() / n) / ((np.sum(lx * lx) - lx.sum() ** 2 / n) * (np.sum(ly * ly) - ly.sum() ** 2 / n)) ** 2
    assert np.isclose(r, want_r, rtol=1e-12)  # the reference's pow(..., 2), not a square root
    assert ens.linreg(np.ones(5), np.ones(5))[0] is False  # singular: every x equal
    assert ens.linreg(np.array([]), np.array([]))[0] is False


def test_fractal_law_from_statistic_rows():
    nb = ens.N_BINS
    x = np.array([2.0, 3.0, 5.0, 9.0]); y = 1.4 * x ** 1.8
    lx, ly = np.log(x), np.log(y)
    row = np.zeros(2 * nb + 8)
    row[2 * nb:] = [len(x), y.sum(), lx.sum(), (lx * lx).sum(), (lx * ly).sum(), ly.sum(), (ly * ly).sum(), 0.0]
    (df, kf), = ens.fractal_law(row[None, :])
    assert abs(df - 1.8) < 1e-10 and abs(kf - 1.4) < 1e-10
    s = ens.summarize(np.stack([row, row]))
    assert s["realizations"] == 2 and s["n_agg_total"] == 8 and abs(s["Df_mean"] - 1.8) < 1e-10


WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from mcac_b200 import ensemble as ens
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 7
idx = ens.shard(n, rank, world)
rows = np.array([[k + 0.5] * (2 * ens.N_BINS + 8) for k in idx]).reshape(len(idx), 2 * ens.N_BINS + 8)
full = ens.gather_rows(rows, idx, n, dist=dist, device="cpu")
assert full.shape == (n, 2 * ens.N_BINS + 8)
assert np.array_equal(full[:, 0], np.arange(n) + 0.5), full[:, 0]
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_all_gather_of_statistic_rows_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert p.stdout.count("ok") == 2


def test_single_process_gather_is_a_scatter_into_realization_order():
    rows = np.arange(3 * (2 * ens.N_BINS + 8), dtype=float).reshape(3, -1)
    full = ens.gather_rows(rows, [1, 3, 5], 6)
    assert np.array_equal(full[[1, 3, 5]], rows) and not full[[0, 2, 4]].any()


@pytest.mark.gpu
def test_concurrent_realizations_equal_the_same_realizations_run_alone(tmp_path):
    """Four realizations of examples/classic.ini (seeds 1000..1003) advanced concurrently by 3 host threads on one device give
    bit-identical states to the same four run one after the other: handles share nothing (own stream, own RNG stream)."""
    from golden_lib import write_interpotential_file
    from mcac_b200 import Ensemble, Simulation, ini_text
    from oracle.run_ref import merged_config

    table = write_interpotential_file(tmp_path / "Interpotential_input.dat")
    texts = [ini_text(merged_config("classic", {"numerics": {"random_seed": s}, "inter_potential": {"interpotential_file": table}}))
             for s in ens.seeds(1000, [0, 1, 2, 3])]
    steps = 600
    e = Ensemble(texts)
    reps = e.run(steps, threads=3)
    assert [r["steps"] for r in reps] == [steps] * 4
    rows = e.morphology_stats()
    for k, t in enumerate(texts):
        solo = Simulation(t)
        r, _ = solo.run(steps)
        assert r["events"] == reps[k]["events"] and r["n_aggregates"] == reps[k]["n_aggregates"]
        a, b = solo.state(), e.sims[k].state()
        np.testing.assert_array_equal(a["sphere_label"], b["sphere_label"])
        np.testing.assert_array_equal(a["spheres"]["x"], b["spheres"]["x"])
        np.testing.assert_array_equal(a["aggregates"]["rg"], b["aggregates"]["rg"])
        np.testing.assert_array_equal(solo.morphology_stats(), rows[k])
    assert len({r["events"] for r in reps}) > 1 or len({r["n_aggregates"] for r in reps}) > 1  # different seeds, different histories
