"""TEST INFRASTRUCTURE ONLY: run the compiled reference (oracle/_ref/MCAC or MCAC_tap).

Writes an .ini (base file + overrides) into a fresh scratch directory, runs the binary there with a
fresh ``output_dir`` (an existing one makes the reference block on stdin,
src/physical_model/physical_model.cpp:196-210) and returns the scratch path.  Never reads
/root/reference at run time: base configs are the dict literals below, restated from
validation/*.ini and examples/classic.ini (SURVEY.md Appendix C lists the keys).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"

# validation/params_monodisperse.ini, params_polydisperse.ini, params_brownian.ini,
# params_surface_growth.ini, params_pytest.ini, examples/classic.ini — restated key by key.
_COMMON_DLCA = {
    "environment": dict(initial_time=0, fractal_dimension=1.78, fractal_prefactor=1.30, pressure=101300,
                        temperature=1700, volume_fraction="10e-6"),
}
CONFIGS = {
    "monodisperse": {
        **_COMMON_DLCA,
        "limits": dict(cpu=-1, mean_monomere_per_aggregate=100, n_iter_without_event=-1, number_of_aggregates=1,
                       physical_time=-1),
        "monomers": dict(density=1800, dispersion_diameter="1.00", initialisation_mode="lognormal", mean_diameter=10,
                         number=800),
        "numerics": dict(n_verlet_divisions=3, pick_method="random"),
        "output": dict(n_time_per_file=5000, write_between_event_frequency=100000, output_dir="out"),
    },
    "polydisperse": {
        **_COMMON_DLCA,
        "limits": dict(cpu=-1, mean_monomere_per_aggregate=100, n_iter_without_event=-1, number_of_aggregates=1,
                       physical_time=-1),
        "monomers": dict(density=1800, dispersion_diameter="1.25", initialisation_mode="lognormal", mean_diameter=20,
                         number=800),
        "numerics": dict(n_verlet_divisions=3, pick_method="random"),
        "output": dict(n_time_per_file=5000, write_between_event_frequency=100000, output_dir="out"),
    },
    "brownian": {
        "environment": dict(initial_time=0, volume_fraction="10e-30"),
        "limits": dict(physical_time="0.00005"),
        "monomers": dict(number=1000),
        "numerics": dict(with_collisions="false", pick_method="last"),
        "output": dict(output_dir="out", write_between_event_frequency=1000),
    },
    "surface_growth": {
        **_COMMON_DLCA,
        "limits": dict(mean_monomere_per_aggregate=40),
        "monomers": dict(density=1800, dispersion_diameter="1.20", initialisation_mode="lognormal", mean_diameter=10,
                         number=800),
        "output": dict(n_time_per_file=5000, output_dir="out"),
        "surface_growth": dict(with_surface_reactions="true", flux_surfgrowth="1e-04", volsurf_method="alphas",
                               full_aggregate_update_frequency=100),
    },
    "pytest": {
        **_COMMON_DLCA,
        "limits": dict(mean_monomere_per_aggregate=40),
        "monomers": dict(density=1800, dispersion_diameter="1.20", initialisation_mode="lognormal", mean_diameter=10,
                         number=20),
        "output": dict(n_time_per_file=5000, output_dir="out"),
        "surface_growth": dict(with_surface_reactions="true", flux_surfgrowth="1e-04", volsurf_method="alphas",
                               full_aggregate_update_frequency=100),
    },
    "classic": {
        "monomers": dict(number=100, density=1800, dispersion_diameter="1.25", mean_diameter=10,
                         initialisation_mode="normal"),
        "environment": dict(initial_time=0, volume_fraction="1e-3", temperature=1700, pressure=101300,
                            fractal_prefactor="1.4", fractal_dimension="1.8"),
        "surface_growth": dict(with_surface_reactions="true", flux_surfgrowth="1e-5", volsurf_method="none",
                               full_aggregate_update_frequency=100),
        "limits": dict(number_of_aggregates=1, n_iter_without_event=-1, cpu=-1, physical_time=-1,
                       mean_monomere_per_aggregate=-1),
        "numerics": dict(enforce_volume_fraction="true", with_collisions="true", n_verlet_divisions=10,
                         pick_method="random", individual_surf_reactions="true"),
        "nucleation": dict(with_nucleation="true", flux="5e23"),
        "flame_coupling": dict(with_flame_coupling="false"),
        "output": dict(output_dir="out", n_time_per_file=10, write_between_event_frequency=90),
        "inter_potential": dict(with_potentials="true", interpotential_file="Interpotential_input.dat",
                                with_external_potentials="true"),
    },
}


def merged_config(base: str, overrides: dict | None = None) -> dict:
    cfg = {sec: dict(kv) for sec, kv in CONFIGS[base].items()}
    for sec, kv in (overrides or {}).items():
        cfg.setdefault(sec, {}).update(kv)
    return cfg


def write_ini(path: Path, cfg: dict) -> None:
    with open(path, "w") as f:
        for sec, kv in cfg.items():
            f.write(f"[{sec}]\n")
            for k, v in kv.items():
                f.write(f"{k}={v}\n")
            f.write("\n")


def run_reference(base: str, overrides: dict | None = None, *, tap: bool = True, env: dict | None = None,
                  workdir: str | None = None, timeout: float | None = None, taskset_core: int | None = None):
    """Run the reference; returns (workdir Path, stdout str). Tap outputs land in workdir/tap."""
    exe = REF_DIR / ("MCAC_tap" if tap else "MCAC")
    if not exe.exists():
        raise FileNotFoundError(f"{exe} missing: run `make -C oracle/ref_build` where /root/reference exists")
    wd = Path(workdir or tempfile.mkdtemp(prefix="mcac_ref_"))
    wd.mkdir(parents=True, exist_ok=True)
    cfg = merged_config(base, overrides)
    out_dir = wd / cfg["output"]["output_dir"]
    if out_dir.exists():
        shutil.rmtree(out_dir)
    write_ini(wd / "params.ini", cfg)
    (wd / "tap").mkdir(exist_ok=True)
    e = dict(os.environ)
    e["MCAC_TAP_DIR"] = str(wd / "tap")
    e.update({k: str(v) for k, v in (env or {}).items()})
    cmd = [str(exe), "params.ini"]
    if taskset_core is not None and shutil.which("taskset"):
        cmd = ["taskset", "-c", str(taskset_core)] + cmd
    p = subprocess.run(cmd, cwd=wd, env=e, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"reference exited with {p.returncode}:\n{p.stdout[-2000:]}")
    return wd, p.stdout


def read_summary(wd: Path) -> dict:
    out = {}
    for line in (Path(wd) / "tap" / "summary.txt").read_text().splitlines():
        k, _, v = line.partition(" ")
        if k == "chunk_times":
            out[k] = [float(x) for x in v.split()]
            continue
        try:
            out[k] = int(v)
        except ValueError:
            try:
                out[k] = float(v)
            except ValueError:
                out[k] = v
    return out


if __name__ == "__main__":
    import argparse
    import json

    ap = argparse.ArgumentParser()
    ap.add_argument("base")
    ap.add_argument("--set", action="append", default=[], help="section.key=value")
    ap.add_argument("--env", action="append", default=[], help="NAME=value (tap controls)")
    ap.add_argument("--workdir")
    ap.add_argument("--no-tap", action="store_true")
    a = ap.parse_args()
    ov: dict = {}
    for s in a.set:
        k, v = s.split("=", 1)
        sec, key = k.split(".", 1)
        ov.setdefault(sec, {})[key] = v
    env = dict(s.split("=", 1) for s in a.env)
    wd, out = run_reference(a.base, ov, tap=not a.no_tap, env=env, workdir=a.workdir)
    print(out[-3000:])
    print("workdir:", wd)
    if not a.no_tap:
        print(json.dumps(read_summary(wd)))
