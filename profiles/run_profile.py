"""Short profiling driver (run under ncu on the B200): the N=1e6 bench workload, a few thousand MC steps + one sweep."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import mcac_b200
from bench import workload_config
from mcac_b200.configs import merged_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
base, ov = workload_config(n, 42)
sim = mcac_b200.Simulation(mcac_b200.ini_text(merged_config(base, ov)))
if steps > 0:
    rep, _ = sim.run(steps, batch=256)
    print({k: rep[k] for k in ("steps", "events", "batches", "kernel_launches", "device_ms")})
print(sim.search_sweep(int(sys.argv[3]) if len(sys.argv) > 3 else 100000, repeats=1))
