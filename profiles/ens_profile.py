"""Short ensemble driver for ncu (run on the B200): N realizations of examples/classic.ini advanced to a late stage outside the
captured launch, then ONE k_ensemble_loop launch of a few steps each (ncu replays a kernel ~40 times: keep it short).
    ncu --profile-from-start off ... python profiles/ens_profile.py [realizations=148] [warm steps=2600] [captured steps=8]"""
import os
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import mcac_b200  # noqa: E402
from bench import classic_texts  # noqa: E402
from golden_lib import write_interpotential_file  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2600
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 8
os.environ.setdefault("MCAC_B200_RESERVE_SPHERES", "60000")
os.environ.setdefault("MCAC_B200_RESERVE_AGGREGATES", "4096")
tmp = tempfile.mkdtemp(prefix="mcac_ens_prof_")
table = write_interpotential_file(Path(tmp) / "Interpotential_input.dat")
e = mcac_b200.Ensemble(classic_texts(list(range(n)), table))
reps = e.run(warm, threads=16)
print("warm:", sum(r["steps"] for r in reps), "steps,", sum(r["n_spheres"] for r in reps) // n, "spheres per realization")
import torch  # noqa: E402  (only for cudaProfilerStart / Stop: ncu --profile-from-start off captures the launch below)

torch.cuda.profiler.start()
reps = e.run(cap, threads=16)
torch.cuda.profiler.stop()
print("captured:", sum(r["steps"] for r in reps), "steps,", sum(r["pair_tests_sphere"] for r in reps), "pair tests (reference count),",
      sum(r["pair_tests_executed"] for r in reps), "executed, kernel ms", reps[0]["device_ms"], "rounds", reps[0]["conflicts"])
