"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference (oracle/_ref/MCAC_tap).

Run in the build container (needs oracle/_ref built from /root/reference; `make -C oracle ref`):
    python tests/golden/make_golden.py [name ...]
Each fixture <name>.npz holds the config overrides, the first `keep` per-step records of the reference run
(search + step taps), every merge, snapshots (initial, a few mid-run, final) and a sha256 digest over ALL
per-step records so long runs are pinned without committing them.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import ref_trace as rt  # noqa: E402
from oracle.run_ref import read_summary, run_reference  # noqa: E402

INTERPOT = "/root/reference/examples/Interpotential_input.dat"
# name -> (base config, overrides, total steps logged (None = whole run), records kept, state steps)
FIXTURES = {
    "pytest_seed42": ("pytest", {"numerics": {"random_seed": 42}}, None, 16000, [0, 432, 446, 5000]),
    "monodisperse_seed42": ("monodisperse", {"numerics": {"random_seed": 42}}, None, 3000, [0, 1000, 200000]),
    "polydisperse_seed42": ("polydisperse", {"numerics": {"random_seed": 42}}, 300000, 3000, [0, 50000]),
    "brownian_seed42": ("brownian", {"numerics": {"random_seed": 42}}, None, 3000, [0, 1000]),
    "surface_growth_seed42": ("surface_growth", {"numerics": {"random_seed": 42}}, 12000, 3000, [0, 500, 6000]),
    "caps_seed7": ("pytest", {"numerics": {"random_seed": 7}, "surface_growth": {"volsurf_method": "caps"},
                              "monomers": {"number": 60}}, 8000, 3000, [0, 4000]),
    "classic_seed1000": ("classic", {"numerics": {"random_seed": 1000},
                                     "inter_potential": {"interpotential_file": INTERPOT}}, 1500, 1500, [0, 700]),
    "c2_small_seed42": ("polydisperse", {"numerics": {"random_seed": 42, "n_verlet_divisions": 12,
                                                      "with_domain_duplication": "false"},
                                         "monomers": {"number": 4000}}, 60000, 3000, [0, 30000]),
    "c3_small_seed42": ("brownian", {"numerics": {"random_seed": 42, "n_verlet_divisions": 16, "with_collisions": "true",
                                                  "pick_method": "random", "with_domain_duplication": "false"},
                                     "environment": {"volume_fraction": "1000e-6"}, "monomers": {"number": 4000},
                                     "limits": {"physical_time": -1}}, 40000, 3000, [0, 20000]),
}
SEARCH_KEYS = ["rand_calls", "source", "dir", "full_distance", "distance", "moving_sphere", "other_sphere", "moving_label",
               "other_label", "n_agg", "time"]
STEP_KEYS = ["label", "dt", "proper_time", "pos", "lpm"]


def digest(searches: np.ndarray, steps: np.ndarray) -> str:
    h = hashlib.sha256()
    for k in SEARCH_KEYS:
        if len(searches):
            h.update(np.ascontiguousarray(searches[k]).tobytes())
    for k in STEP_KEYS:
        h.update(np.ascontiguousarray(steps[k]).tobytes())
    return h.hexdigest()


def flat_state(prefix: str, st: dict, out: dict) -> None:
    for k, v in st.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                out[f"{prefix}/{k}/{kk}"] = vv
        else:
            out[f"{prefix}/{k}"] = np.asarray(v)


def make(name: str) -> None:
    base, ov, total, keep, state_steps = FIXTURES[name]
    env = {"MCAC_TAP_MAX_STEPS": total if total else 10**9, "MCAC_TAP_STATE_STEPS": ",".join(map(str, state_steps)),
           "MCAC_TAP_SORT_CALLS": "0,1,2"}
    if total:
        env["MCAC_TAP_EXIT_STEP"] = total
    wd, _ = run_reference(base, ov, env=env)
    tap = wd / "tap"
    summ = read_summary(wd)
    searches = rt.read_searches(tap / "searches.bin") if (tap / "searches.bin").exists() else np.zeros(0, rt.SEARCH_DTYPE)
    if len(searches):
        # with potentials the orientation loop (calcul.cpp:119-141) may search several times per step: keep the last try
        last = np.r_[searches["step"][1:] != searches["step"][:-1], True]
        searches = searches[last]
    steps = rt.read_steps(tap / "steps.bin")
    merges = rt.read_merges(tap / "merges.bin") if (tap / "merges.bin").exists() else np.zeros(0, rt.MERGE_DTYPE)
    out: dict = {"meta": np.array(json.dumps(dict(base=base, overrides=ov, total_steps=int(summ["steps"]), bounded=bool(total),
                                                  keep=keep, state_steps=state_steps, summary=summ,
                                                  digest=digest(searches, steps))))}
    out["searches"] = searches[:keep]
    out["steps"] = steps[:keep]
    out["merges"] = merges
    flat_state("state_init", rt.read_state(tap / "state_init.bin"), out)
    flat_state("state_final", rt.read_state(tap / "state_final.bin"), out)
    for s in state_steps:
        f = tap / f"state_{s}.bin"
        if f.exists():
            flat_state(f"state_{s}", rt.read_state(f), out)
    for k in (0, 1, 2):
        f = tap / f"sort_{k}.bin"
        if f.exists():
            for kk, vv in rt.read_sort(f).items():
                out[f"sort_{k}/{kk}"] = np.asarray(vv)
    dst = Path(__file__).parent / f"{name}.npz"
    np.savez_compressed(dst, **out)
    print(f"{name}: {summ['steps']} steps, {len(merges)} merge calls, {dst.stat().st_size / 1e6:.2f} MB")
    shutil.rmtree(wd, ignore_errors=True)


def make_interpotential_fixture() -> None:
    """examples/Interpotential_input.dat (input table of examples/classic.ini) as a compact array; tests write it back as text."""
    vals = np.array(open(INTERPOT).read().split(), dtype=np.float64)
    np.savez_compressed(Path(__file__).parent / "interpotential_table.npz", values=vals)


if __name__ == "__main__":
    make_interpotential_fixture()
    for n in (sys.argv[1:] or list(FIXTURES)):
        make(n)
