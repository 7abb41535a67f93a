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
