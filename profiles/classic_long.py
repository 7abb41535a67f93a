"""Debug / soak driver: one realization of examples/classic.ini for many steps (argv: seed, steps)."""
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import mcac_b200
from golden_lib import write_interpotential_file
from mcac_b200.configs import merged_config

seed, steps = int(sys.argv[1]), int(sys.argv[2])
tmp = tempfile.mkdtemp()
table = write_interpotential_file(Path(tmp) / "Interpotential_input.dat")
sim = mcac_b200.Simulation(mcac_b200.ini_text(merged_config("classic", {"numerics": {"random_seed": seed}, "inter_potential": {"interpotential_file": table}})))
done = 0
while done < steps:
    rep, _ = sim.run(min(250, steps - done))
    done += rep["steps"]
    print(done, {k: rep[k] for k in ("events", "n_aggregates", "n_spheres", "duplications", "device_ms", "kernel_launches")}, flush=True)
    if rep["finished"]:
        break
