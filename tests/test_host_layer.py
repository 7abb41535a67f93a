"""Host layer of the product (no GPU): .ini reader / derived constants / initial placement of libmcac_b200.so
against (i) the reference's golden metadata (pymcac/tests/test_read.py:31-48), (ii) the reference's own initial
state (state_init of every golden fixture, bit-exact: same libm on the host) and (iii) the exported C ABI."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import ref_trace as rt
from golden_lib import Golden
from oracle.run_ref import merged_config

import mcac_b200
from mcac_b200 import HostModel, ini_text

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "mcac_b200.h").read_text()
    declared = set(re.findall(r"\b(mcac_(?:gpu|sim|host|ensemble)_\w+)\s*\(", header))
    assert len(declared) >= 40
    L = mcac_b200.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/mcac_b200.h but not exported"
    from mcac_b200 import _capi
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)  # the ctypes binding covers the whole header


def test_golden_metadata_of_the_reference():
    """pymcac/tests/test_read.py:31-48 — the reference's only pinned physics constants (params_pytest.ini)."""
    m = HostModel(ini_text(merged_config("pytest", {"numerics": {"random_seed": 42}})), place=False)
    md = m.metadata()
    ref = {"flux_surfgrowth": 0.0001, "u_sg": 5.55556e-08, "dfe": 1.78, "kfe": 1.3, "lambda": 4.98113e-07,
           "rpeqmass": 5.25563e-09, "gamma_": 1.378, "P [Pa]": 101300.0, "T [K]": 1700.0, "Mu": 5.662e-05,
           "Rho [kg/m3]": 1800.0, "Dpm [nm]": 10.0, "sigmaDpm [nm]": 1.2, "FV [ppt]": 1e-05, "L": 1.06741e-06, "N []": 20.0}
    got = {k: float(v) for k, v in md.items()}
    assert got == ref


@pytest.mark.parametrize("name", ["pytest_seed42", "monodisperse_seed42", "polydisperse_seed42", "brownian_seed42",
                                  "surface_growth_seed42", "c2_small_seed42", "c3_small_seed42"])
def test_initial_placement_matches_reference_bit_for_bit(name):
    g = Golden(name)
    ref = g.state("state_init")
    st = HostModel(ini_text(merged_config(g.base, g.overrides))).state()
    for k in rt.SPHERE_FIELDS:
        np.testing.assert_array_equal(st["spheres"][k], ref["spheres"][k], err_msg=f"sphere {k}")
    for k in rt.AGG_FIELDS:
        if k == "electric_charge_field":
            continue
        np.testing.assert_array_equal(st["aggregates"][k], ref["aggregates"][k], err_msg=f"aggregate {k}")
    for k in ["members", "offsets", "agg_cell", "member_volumes", "member_surfaces", "member_distances_center"]:
        np.testing.assert_array_equal(st[k], ref[k], err_msg=k)
    assert st["maxradius"] == ref["maxradius"] and st["max_time_step"] == ref["max_time_step"]
    assert st["rand_consumed"] == ref["rand_calls"]


def test_input_errors_carry_the_reference_error_codes():
    with pytest.raises(mcac_b200.McacError) as e:
        HostModel("[numerics]\npick_method=nope\nrandom_seed=1\n", place=False)
    assert e.value.code == 4  # INPUT_ERROR
    with pytest.raises(mcac_b200.McacError) as e:  # too dense: 200 monomers at 60 % volume fraction
        HostModel("[monomers]\nnumber=200\n[environment]\nvolume_fraction=0.6\n[numerics]\nrandom_seed=1\n")
    assert e.value.code == 6  # TOO_DENSE_ERROR
