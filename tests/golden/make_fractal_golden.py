"""Golden values of the reference's own morphology regression (AggregatList::get_instantaneous_fractal_law -> linreg,
src/aggregats/aggregat_list_fractal_law.cpp:23-33, src/tools/tools.cpp:126-157), evaluated by the UNMODIFIED reference
(oracle/_ref/MCAC_tap) on the snapshots the *.npz fixtures of this directory hold (same .ini + seed -> same states).
    python tests/golden/make_fractal_golden.py        -> tests/golden/fractal_law.json  {fixture: {state name: [ok, a, b, r]}}
"""
from __future__ import annotations

import json
import shutil
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from make_golden import FIXTURES  # noqa: E402
from oracle.run_ref import run_reference  # noqa: E402

NAMES = ["pytest_seed42", "monodisperse_seed42", "surface_growth_seed42", "classic_seed1000", "c2_small_seed42", "c3_small_seed42"]

if __name__ == "__main__":
    out = {}
    for name in NAMES:
        base, ov, total, keep, state_steps = FIXTURES[name]
        env = {"MCAC_TAP_MAX_STEPS": 0, "MCAC_TAP_STATE_STEPS": ",".join(map(str, state_steps))}
        if total:
            env["MCAC_TAP_EXIT_STEP"] = total
        wd, _ = run_reference(base, ov, env=env)
        out[name] = {}
        for f in sorted((wd / "tap").glob("*.fractal")):
            out[name][f.name[:-len(".bin.fractal")]] = [float(v) for v in np.fromfile(f, dtype=np.float64)]
        print(name, out[name])
        shutil.rmtree(wd, ignore_errors=True)
    (Path(__file__).parent / "fractal_law.json").write_text(json.dumps(out, indent=1) + "\n")
