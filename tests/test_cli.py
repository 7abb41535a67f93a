"""`MCAC <params.ini>` command line of the product (host facade over the C ABI): same argument convention and exit codes
as the reference's src/main.cpp:26-56."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "mcac_b200" / "bin" / "MCAC"

INI = """[environment]
temperature=1700
volume_fraction=100e-6
fractal_dimension=1.78
[monomers]
number=2000
mean_diameter=10
dispersion_diameter=1.0
[limits]
mean_monomere_per_aggregate=20
[numerics]
random_seed=3
n_verlet_divisions=8
[output]
output_dir=cli_out
"""


def test_missing_argument_is_an_input_error():
    p = subprocess.run([str(EXE)], capture_output=True, text=True)
    assert p.returncode == 4 and "Missing argument" in p.stdout


def test_missing_file_is_an_input_error(tmp_path):
    p = subprocess.run([str(EXE), str(tmp_path / "nope.ini")], capture_output=True, text=True)
    assert p.returncode == 4


@pytest.mark.gpu
def test_cli_runs_to_the_npp_limit(tmp_path):
    (tmp_path / "params.ini").write_text(INI)
    p = subprocess.run([str(EXE), str(tmp_path / "params.ini")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "The End" in p.stdout
    adv = (tmp_path / "cli_out" / "advancement.dat").read_text().strip().splitlines()
    last = [float(x) for x in adv[-1].split()]
    assert len(last) == 9 and last[3] >= 20.0  # 9 columns (calcul.cpp:29-43); mean Npp reached the limit


ADV_CASES = {
    "pytest_seed42": ("pytest", {"numerics": {"random_seed": 42}}),
    "monodisperse_seed42": ("monodisperse", {"numerics": {"random_seed": 42}}),
    "c3_small_seed42": ("brownian", {"numerics": {"random_seed": 42, "n_verlet_divisions": 16, "with_collisions": "true", "pick_method": "random",
                                                  "with_domain_duplication": "false"},
                                     "environment": {"volume_fraction": "1000e-6"}, "monomers": {"number": 4000},
                                     "limits": {"physical_time": -1, "number_of_aggregates": 3800}}),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(ADV_CASES))
def test_cli_writes_the_reference_advancement_file(name, tmp_path):
    """`MCAC params.ini` end to end against the unmodified reference's own output file (tests/golden/advancement_*.dat, written by
    oracle/_ref/MCAC for the same .ini + seed): advancement.dat (src/calcul.cpp:29-43; rows at the loop tops PhysicalModel::
    time_to_write selects, physical_model.cpp:338-356) must have the same rows — same count, same 9 columns, values equal to the
    6 significant digits the stream prints (a last printed digit may differ where a 1e-12 relative difference crosses a rounding
    boundary: at most a handful of values per file).  pytest: growth + one duplication, 15 456 steps; monodisperse: 1 000 452
    steps, 1 688 events, two duplications (reference md5 1965853c... in SURVEY.md §8c)."""
    import numpy as np
    from mcac_b200.configs import merged_config
    from oracle.run_ref import write_ini

    base, ov = ADV_CASES[name]
    cfg = merged_config(base, ov)
    write_ini(tmp_path / "params.ini", cfg)
    p = subprocess.run([str(EXE), str(tmp_path / "params.ini")], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    got = (tmp_path / cfg["output"]["output_dir"] / "advancement.dat").read_text().splitlines()
    ref = (ROOT / "tests" / "golden" / f"advancement_{name}.dat").read_text().splitlines()
    assert len(got) == len(ref), (len(got), len(ref))
    differing = [i for i, (a, b) in enumerate(zip(got, ref)) if a != b]
    g = np.array([[float(x) for x in line.split()] for line in got])
    r = np.array([[float(x) for x in line.split()] for line in ref])
    assert g.shape == r.shape and g.shape[1] == 9
    np.testing.assert_allclose(g, r, rtol=2e-5, atol=0)      # printed with 6 significant digits
    assert len(differing) <= max(3, len(ref) // 200), f"{len(differing)} of {len(ref)} rows differ in a printed digit: {differing[:10]}"
    assert "The End" in p.stdout and "Final number of aggregates" in p.stdout
