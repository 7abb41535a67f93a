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
    declared = set(re.findall(r"\b(mcac_(?:gpu|sim|host|ensemble|io)_\w+)\s*\(", header))
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


REFERENCE = Path("/root/reference")
REF_FLAGS = ["-std=c++17", "-include", str(ROOT / "oracle" / "ref_build" / "prelude.hpp"), f"-I{REFERENCE / 'include'}",
             f"-I{ROOT / 'oracle' / 'ref_build' / 'shim'}", f"-I{ROOT / 'include'}"]


@pytest.mark.skipif(not (REFERENCE / "include").is_dir(), reason="the reference's headers are not on this machine")
def test_reference_side_shim_compiles_against_the_reference_headers(tmp_path):
    """INTEGRATION.md §3 as code: tests/native/shim/aggregat_list_gpu.cpp defines the AggregatList members mcac::calcul calls
    (include/aggregats/aggregat_list.hpp:43-116) with bodies that forward to the C ABI.  Compiling it against the REAL reference
    headers proves that every signature, the weak_ptr contact info and the ErrorCodes / exception mapping fit the C ABI."""
    import subprocess
    obj = tmp_path / "aggregat_list_gpu.o"
    p = subprocess.run(["g++", *REF_FLAGS, "-Wall", "-c", "-o", str(obj), str(ROOT / "tests" / "native" / "shim" / "aggregat_list_gpu.cpp")],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-4000:]
    syms = subprocess.run(["nm", "-C", str(obj)], capture_output=True, text=True).stdout
    for method in ["distance_to_next_contact(unsigned long, std::array<double, 3ul> const&, double) const", "merge(mcac::AggregateContactInfo)",
                   "sort_time_steps(double)", "pick_random() const", "pick_last() const", "get_time_step(double) const", "refresh()",
                   "croissance_surface(double)", "croissance_surface(double, unsigned long)", "duplication()"]:
        assert f" T mcac::AggregatList::{method}" in syms, method
    for fn in ["mcac_gpu_contact_search", "mcac_gpu_merge", "mcac_gpu_sort_time_steps", "mcac_gpu_pick_random", "mcac_gpu_pick_last",
               "mcac_gpu_refresh", "mcac_gpu_grow", "mcac_gpu_duplicate", "mcac_gpu_translate", "mcac_gpu_update"]:
        assert f" U {fn}" in syms, fn


@pytest.mark.skipif(not (REFERENCE / "src").is_dir(), reason="the reference's sources are not on this machine")
def test_reference_binaries_in_tree_are_what_the_recipe_builds(tmp_path):
    """oracle/_ref/MCAC* are build artefacts (git-ignored).  Rebuild them from /root/reference with the committed recipe into a
    scratch directory and check that the fresh binary AND the one in the tree both write the advancement.dat SURVEY.md §8c records
    for params_pytest.ini + seed 42 (md5 68dddd7a...762b) — i.e. the binary the benchmark's reference arm times is the recipe's."""
    import hashlib
    import shutil
    import subprocess
    from oracle.run_ref import write_ini
    out = tmp_path / "ref"
    subprocess.check_call(["make", "-C", str(ROOT / "oracle" / "ref_build"), "-j8", f"OUT={out}", f"{out}/MCAC"], stdout=subprocess.DEVNULL)
    cfg = merged_config("pytest", {"numerics": {"random_seed": 42}})
    md5 = {}
    for tag, exe in [("fresh", out / "MCAC"), ("tree", ROOT / "oracle" / "_ref" / "MCAC"), ("tree_tap", ROOT / "oracle" / "_ref" / "MCAC_tap")]:
        wd = tmp_path / tag
        wd.mkdir()
        (wd / "tap").mkdir()
        write_ini(wd / "params.ini", cfg)
        subprocess.run([str(exe), "params.ini"], cwd=wd, stdin=subprocess.DEVNULL, stdout=subprocess.DEVNULL, check=True, timeout=300,
                       env={"MCAC_TAP_DIR": str(wd / "tap"), "PATH": "/usr/bin:/bin"})
        md5[tag] = hashlib.md5((wd / "out" / "advancement.dat").read_bytes()).hexdigest()
        shutil.rmtree(wd)
    assert md5["fresh"] == "68dddd7ad068b6939e7d8525991a762b", md5
    assert md5["tree"] == md5["fresh"] and md5["tree_tap"] == md5["fresh"], md5
    assert hashlib.md5((ROOT / "tests" / "golden" / "advancement_pytest_seed42.dat").read_bytes()).hexdigest() == md5["fresh"]


@pytest.mark.skipif(not (ROOT / "oracle" / "_ref" / "MCAC").exists(), reason="oracle/_ref/MCAC is not built")
@pytest.mark.parametrize("base", ["pytest", "classic"])
def test_params_ini_echo_is_the_reference_file(base, tmp_path):
    """PhysicalModel writes <output_dir>/params.ini = inipp::Ini::generate of the parsed file (physical_model.cpp:271-272): every
    key the constructor looked up, sorted, keys absent from the input with an empty value.  Compared with the file the unmodified
    reference writes for the same input (the reference's inipp is un-vendored: this pins the stand-in's and our reader's echo)."""
    import subprocess
    from oracle.run_ref import write_ini
    ov = {"numerics": {"random_seed": 42}, "limits": {"physical_time": "1e-9"}}
    if base == "classic":
        ov["inter_potential"] = {"with_potentials": "false", "with_external_potentials": "false"}
    cfg = merged_config(base, ov)
    write_ini(tmp_path / "params.ini", cfg)
    subprocess.run([str(ROOT / "oracle" / "_ref" / "MCAC"), "params.ini"], cwd=tmp_path, stdin=subprocess.DEVNULL, stdout=subprocess.DEVNULL,
                   check=True, timeout=300)
    ref = (tmp_path / "out" / "params.ini").read_text()
    got = HostModel(ini_text(cfg), place=False).ini_echo()
    assert got == ref


def test_options_that_are_not_built_are_refused_not_ignored():
    """options of the reference whose code is not on the built path must fail with InputError (exit code 4) instead of
    producing a trajectory that silently differs from the reference's."""
    for sec, key, val in [("inter_potential", "with_dynamic_random_charges", "true"), ("inter_potential", "with_electric_charges", "true"),
                          ("numerics", "with_domain_reduction", "true"), ("flame_coupling", "with_flame_coupling", "true"),
                          ("surface_growth", "volsurf_method", "sbl")]:
        with pytest.raises(mcac_b200.McacError) as e:
            HostModel(f"[{sec}]\n{key}={val}\n[numerics]\nrandom_seed=1\n", place=False)
        assert e.value.code == 4, (key, e.value)
    with pytest.raises(mcac_b200.McacError) as e:
        HostModel("[surface_growth]\nwith_surface_reactions=true\nflux_surfgrowth=-1e-4\n[numerics]\nrandom_seed=1\n", place=False)
    assert e.value.code == 4


def test_negative_random_seed_is_replaced_like_init_random():
    """src/tools/tools.cpp:41-50: random_seed < 0 (the default: none of the validation/*.ini sets one) seeds srand() with a hash of
    clock / time / pid.  The host layer must accept such files and report the seed it used (it is what a replay needs)."""
    text = "[monomers]\nnumber=50\n[environment]\nvolume_fraction=1e-5\n"
    for _ in range(20):  # the hash, read as an int, is negative half of the time: such a seed cannot be written back into an .ini
        m = HostModel(text)  # (the reference has the same limitation) — draw until it is replayable
        assert m.state()["n_agg"] == 50
        seed = m.params().random_seed
        if seed < 2 ** 31:
            break
    else:
        pytest.fail("20 clock/pid seeds in a row were negative as int")
    again = HostModel(text + f"[numerics]\nrandom_seed={seed}\n")
    assert again.params().random_seed == seed
    np.testing.assert_array_equal(m.state()["spheres"]["x"], again.state()["spheres"]["x"])
