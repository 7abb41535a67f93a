"""Probe of the CTA that builds the exact cumulative table (csrc/seq_cumsum.cuh, k_event): synthetic tie-dominated pick tables of N
entries with x lighter ones, one sort each; MCAC_B200_K9_DEBUG=1 prints the builder's cycles by part, the report gives the sparse
simulation's cycles to compare with.  usage: MCAC_B200_K9_DEBUG=1 python profiles/cum_builder_probe.py [N] [x ...]"""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("MCAC_B200_TIE_MIN_N", "1000")
import mcac_b200
from mcac_b200 import HostModel, Simulation, ini_text
from mcac_b200.configs import merged_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
xs_list = [int(v) for v in sys.argv[2:]] or [1500, 4000, 7000]
text = ini_text(merged_config("monodisperse", {"numerics": {"random_seed": 5, "n_verlet_divisions": 40}, "monomers": {"number": n},
                                               "environment": {"volume_fraction": "10e-6"}}))
hm = HostModel(text).state()
rng = np.random.default_rng(1)
for xs in xs_list:
    ts = np.full(n, 0.5)
    sparse = rng.choice(n, xs, replace=False)
    ts[sparse] = 0.5 + 4.0 * rng.random(xs)
    st = dict(hm)
    af = hm["agg_fields"].copy()
    af[3] = ts  # TIME_STEP column
    st["agg_fields"] = af
    sim = Simulation(text)
    sim.upload(st)
    for _ in range(3):
        sim.sort_time_steps(2.0)
    idx, cum = sim.pick_table()
    keys = 2.0 / ts
    exact = bool(np.array_equal(cum, np.cumsum(np.sort(keys))))
    rep, _ = sim.run(0)
    print(f"x = {xs}: cumulative table == sequential sum: {exact}; tie sorts {rep['tie_sorts']}, sparse simulation cycles per sort "
          f"{[c // max(1, rep['tie_sorts']) for c in rep['tie_sim_cycles']]}, event phase cycles per sort "
          f"{[c // max(1, rep['tie_sorts']) for c in rep['event_phase_cycles']]}", flush=True)
    del sim
