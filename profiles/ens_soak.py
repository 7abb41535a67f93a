"""Soak of the ensemble runner: argv = realizations threads steps_per_call calls"""
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import mcac_b200
from golden_lib import write_interpotential_file
from mcac_b200.configs import merged_config

R, T, M, K = (int(x) for x in sys.argv[1:5])
FIRST = int(sys.argv[5]) if len(sys.argv) > 5 else 0
tmp = tempfile.mkdtemp()
table = write_interpotential_file(Path(tmp) / "Interpotential_input.dat")
texts = [mcac_b200.ini_text(merged_config("classic", {"numerics": {"random_seed": 1000 + k}, "inter_potential": {"interpotential_file": table}}))
         for k in range(FIRST, FIRST + R)]
e = mcac_b200.Ensemble(texts)
for c in range(K):
    t0 = time.perf_counter()
    try:
        reps = e.run(M, threads=T)
    except mcac_b200.McacError as err:
        import numpy as np
        print("ERROR", err, flush=True)
        for k, sim in enumerate(e.sims):
            msg = sim.L.mcac_gpu_last_error(sim.h).decode()
            if not msg:
                continue
            st = sim.state()
            idx, cum = sim.pick_table()
            out = ROOT / "gpurun_out" / f"sortfail_{FIRST + k}.npz"
            np.savez(out, time_step=st["aggregates"]["time_step"], max_time_step=st["max_time_step"], idx=idx, cum=cum, n_agg=st["n_agg"])
            print("dumped", out, msg, "n_agg", st["n_agg"], flush=True)
        raise SystemExit(1)
    dt = time.perf_counter() - t0
    import torch
    free_b, total_b = torch.cuda.mem_get_info()
    print(c, "used_GB %.1f" % ((total_b - free_b) / 1e9), "steps/s %.0f" % (sum(r["steps"] for r in reps) / dt), "n_sph", sorted({r["n_spheres"] for r in reps})[:3], "dups", sum(r["duplications"] for r in reps), flush=True)
print("OK")
