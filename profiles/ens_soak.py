"""Soak of the ensemble runner: argv = realizations threads steps_per_call calls"""
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import mcac_b200
from golden_lib import write_interpotential_file
from oracle.run_ref import merged_config

R, T, M, K = (int(x) for x in sys.argv[1:5])
tmp = tempfile.mkdtemp()
table = write_interpotential_file(Path(tmp) / "Interpotential_input.dat")
texts = [mcac_b200.ini_text(merged_config("classic", {"numerics": {"random_seed": 1000 + k}, "inter_potential": {"interpotential_file": table}}))
         for k in range(R)]
e = mcac_b200.Ensemble(texts)
for c in range(K):
    t0 = time.perf_counter()
    reps = e.run(M, threads=T)
    dt = time.perf_counter() - t0
    print(c, "steps/s %.0f" % (sum(r["steps"] for r in reps) / dt), "n_sph", sorted({r["n_spheres"] for r in reps})[:3], "dups", sum(r["duplications"] for r in reps), flush=True)
print("OK")
