"""Per-kernel timings of the path on the N=1e6 bench state (after `steps` MC steps so that some aggregates exist):
prints one JSON line per kernel with ms per launch, units, algorithmic bytes per unit and GB/s."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import mcac_b200
from bench import workload_config
from mcac_b200.configs import merged_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
base, ov = workload_config(n, 42)
sim = mcac_b200.Simulation(mcac_b200.ini_text(merged_config(base, ov)))
rep, _ = sim.run(steps, batch=256)
print(json.dumps({k: rep[k] for k in ("steps", "events", "sorts", "sort_levels", "sort_span_elements", "device_ms", "n_aggregates")}))
ph = ["reduce", "labels", "sort_init", "grid_levels", "local_levels", "leaves", "cumulative", "pick_table"]
print(json.dumps({"event_kernel_us_per_sort": {p: round(c / 1965.0 / max(1, rep["sorts"]), 1) for p, c in zip(ph, rep["event_phase_cycles"])}}))
# algorithmic bytes per unit: DESIGN.md §5 / SURVEY.md §8(d)
BYTES = {"cells": 12 + 4 + 32 + 32 + 8, "grow": 32, "update_partial": 40 + 168, "update_full": 40 + 168, "event_sort": 60, "event_nosort": 4 + 24 + 8,
         "grid_barriers_x100": 0, "rng_fill": 4, "morphology_stats": 28}
for k in ["cells", "grow", "update_partial", "update_full", "event_sort", "event_nosort", "grid_barriers_x100", "rng_fill", "morphology_stats"]:
    r = sim.kernel_bench(k, reps=10)
    r["bytes_per_unit"] = BYTES[k]
    r["GBps"] = r["units"] * BYTES[k] / (r["ms"] * 1e-3) / 1e9 if r["ms"] > 0 else None
    print(json.dumps(r))
