"""Tuning probe (B200): the one-CTA sparse simulation of tie_sort.cuh alone, on synthetic tie-dominated tables of the bench size."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import mcac_b200
from bench import workload_config
from mcac_b200.configs import merged_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
base, ov = workload_config(n, 42)
sim = mcac_b200.Simulation(mcac_b200.ini_text(merged_config(base, ov)))
sim.run(256, batch=256)
xs = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [250, 500, 1000, 2000, 3915, 8000]
for x in xs:
    os.environ["MCAC_B200_PROBE_X"] = str(x)
    r = sim.kernel_bench("plan_probe", reps=20)
    print(x, r, flush=True)
