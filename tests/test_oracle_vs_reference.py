"""Pins the CPU restatement (oracle/mcac_oracle.cpp) against the UNMODIFIED reference.

The golden fixtures are traces of oracle/_ref/MCAC_tap (the reference's own sources compiled from
/root/reference with link-time taps; tests/golden/make_golden.py).  Bar: BIT-EXACT on every field — same
machine, same glibc / libstdc++, no FMA contraction — for all five BASELINE configs plus caps / scaled variants.
"""
import numpy as np
import pytest

import ref_trace as rt
from golden_lib import SEARCH_KEYS, Golden, digest_from_oracle_records
from oracle_lib import Oracle

FIXTURES = ["pytest_seed42", "monodisperse_seed42", "polydisperse_seed42", "brownian_seed42", "surface_growth_seed42",
            "caps_seed7", "classic_seed1000", "c2_small_seed42", "c3_small_seed42"]


def assert_state_equal(got: dict, ref: dict, tag: str):
    for k in rt.SPHERE_FIELDS:
        np.testing.assert_array_equal(got["spheres"][k], ref["spheres"][k], err_msg=f"{tag}: sphere {k}")
    for k in rt.AGG_FIELDS:
        if k == "electric_charge_field":  # storage column the reference never writes (aggregat_storage.cpp:25-46)
            continue
        np.testing.assert_array_equal(got["aggregates"][k], ref["aggregates"][k], err_msg=f"{tag}: aggregate {k}")
    for k in ["sphere_label", "sphere_charge", "agg_n_spheres", "agg_charge", "members", "offsets", "agg_cell",
              "member_volumes", "member_surfaces", "member_distances_center"]:
        np.testing.assert_array_equal(got[k], ref[k], err_msg=f"{tag}: {k}")
    for k in rt.SCALARS:
        assert got[k] == ref[k], f"{tag}: scalar {k}: {got[k]!r} != {ref[k]!r}"


def run_like_reference(g: Golden, n_steps: int, partial_last: bool) -> tuple[Oracle, np.ndarray]:
    o = Oracle(g.base, g.overrides)
    if partial_last:
        recs = o.run(n_steps - 1)
        recs = np.concatenate([recs, o.run_partial_step()])
    else:
        recs = o.run(n_steps)
    return o, recs


@pytest.mark.parametrize("name", FIXTURES)
def test_full_trace(name):
    g = Golden(name)
    total = g.meta["total_steps"]
    o = Oracle(g.base, g.overrides)
    assert_state_equal(o.state(), g.state("state_init"), "initial state (a23 placement, enforce_volume_fraction)")
    if g.meta["bounded"]:
        recs = np.concatenate([o.run(total - 1), o.run_partial_step()])
    else:
        recs = o.run(total + 10)
        assert o.finished
    assert len(recs) == total
    with_coll = len(g.searches) > 0
    k = len(g.steps)
    if with_coll:
        for f in SEARCH_KEYS:
            a, b = recs[f][:k], g.searches[f]
            np.testing.assert_array_equal(a, b, err_msg=f"search tap field {f}")
    for f, src in [("label", "source"), ("dt", "dt"), ("proper_time", "proper_time"), ("pos", "pos"), ("lpm", "full_distance")]:
        np.testing.assert_array_equal(recs[src][:k], g.steps[f], err_msg=f"step tap field {f}")
    # every record of the whole run, through the digest
    assert digest_from_oracle_records(recs, with_coll) == g.meta["digest"]
    merged_steps = np.nonzero(recs["merged"])[0]
    np.testing.assert_array_equal(merged_steps, g.merges["step"][g.merges["ok"] == 1])
    assert_state_equal(o.state(), g.state("state_final"), "final state")
    c, s = o.counters(), g.meta["summary"]
    assert c["rand_calls"] == s["rand_calls"]
    if not g.meta["bounded"]:
        assert c["pair_sphere"] == s["pair_tests_sphere"] and c["pair_bounding"] == s["pair_tests_bounding"]


@pytest.mark.parametrize("name", ["pytest_seed42", "monodisperse_seed42", "surface_growth_seed42", "classic_seed1000",
                                  "c3_small_seed42"])
def test_mid_run_snapshots(name):
    g = Golden(name)
    for s in g.meta["state_steps"]:
        if s > 60000 or s >= g.meta["total_steps"]:
            continue
        o, _ = run_like_reference(g, s + 1, partial_last=True)  # tap dumps inside time_forward of step s
        assert_state_equal(o.state(), g.state(f"state_{s}"), f"snapshot after the move of step {s}")


@pytest.mark.parametrize("name", ["monodisperse_seed42", "polydisperse_seed42", "c3_small_seed42"])
def test_pick_table_matches_std_sort_of_reference(name):
    """index_sorted_time_steps / cumulative_time_steps after the first sort_time_steps call (H3: tie order)."""
    g = Golden(name)
    srt = g.sort(0)
    o = Oracle(g.base, g.overrides)
    o.run(1)
    idx, cum = o.pick_table()
    np.testing.assert_array_equal(idx, srt["idx"])
    np.testing.assert_array_equal(cum, srt["cum"])


def test_fractal_law_restatements_against_the_reference():
    """AggregatList::get_instantaneous_fractal_law -> linreg (aggregat_list_fractal_law.cpp:23-33, tools.cpp:126-157) evaluated by
    the unmodified reference on the fixture snapshots (tests/golden/fractal_law.json, made by make_fractal_golden.py): the oracle's
    C restatement must agree bit for bit, the numpy form used by mcac_b200/ensemble.py to summation-order rounding."""
    import json
    from pathlib import Path

    from golden_lib import Golden
    from mcac_b200 import ensemble as ens
    from oracle_lib import linreg
    gold = json.loads((Path(__file__).parent / "golden" / "fractal_law.json").read_text())
    checked = 0
    for name, states in gold.items():
        g = Golden(name)
        for key, (ok, a, b, r) in states.items():
            if key.startswith("state_") and key[6:].isdigit() and int(key[6:]) not in g.meta["state_steps"]:
                continue
            try:
                st = g.state(key)
            except KeyError:
                continue
            x, y = st["aggregates"]["dg_over_dp"], st["agg_n_spheres"].astype(float)
            got = linreg(x, y)
            assert got[0] == bool(ok), (name, key)
            for u, v in zip(got[1:], (a, b, r)):
                assert (np.isnan(u) and np.isnan(v)) or u == v, (name, key, got, (ok, a, b, r))
            if np.ptp(np.log(x)) < 1e-6:  # every aggregate still a monomer: the normal matrix is singular up to rounding noise, and
                checked += 1              # whether |denom| < 1e-9 then depends on the summation order (sequential in the reference)
                continue
            nok, na, nb_, nr = ens.linreg(x, y)
            assert nok == bool(ok)
            if ok and np.isfinite(r):
                np.testing.assert_allclose([na, nb_], [a, b], rtol=1e-9, atol=1e-12)
                np.testing.assert_allclose(nr, r, rtol=1e-6)
            checked += 1
    assert checked >= 12


def test_oracle_resynchronised_from_a_state_continues_the_same_trajectory():
    """orc_set_state (used by the full-size GPU tests to check late windows of a 1e6 run without replaying it from step 0 on the
    CPU): an oracle re-created from the state of another one, RNG repositioned, must produce the very same records."""
    from golden_lib import Golden
    g = Golden("c3_small_seed42")
    a = Oracle(g.base, g.overrides)
    a.run(5000, record=False)
    st = a.state()
    b = Oracle(g.base, g.overrides)
    b.set_state(st, a.counters()["rand_calls"])
    ra, rb = a.run(3000), b.run(3000)
    assert ra["merged"].sum() > 3
    for f in ra.dtype.names:
        if f == "step":
            continue
        np.testing.assert_array_equal(ra[f], rb[f], err_msg=f)
    sa, sb = a.state(), b.state()
    np.testing.assert_array_equal(sa["sphere_label"], sb["sphere_label"])
    np.testing.assert_array_equal(sa["aggregates"]["rg"], sb["aggregates"]["rg"])
    # Aggregate::update() of every aggregate reproduces the stored morphology from radii + relative positions alone
    # (of every aggregate that was updated since the initial radius rescale: aggregat_list_storage.cpp:75-87 re-runs only
    # compute_volume_surface on the monomers, so their stored Rg / time step still belong to the radii before the rescale)
    b.update_all()
    sc = b.state()
    multi = sb["agg_n_spheres"] > 1
    assert multi.sum() > 10
    for k in ("rg", "volume", "surface", "rmax", "lpm", "time_step", "f_agg", "d_m"):
        np.testing.assert_array_equal(sb["aggregates"][k][multi], sc["aggregates"][k][multi], err_msg=k)


def test_oracle_resynchronised_in_growth_mode_right_after_a_full_update():
    """Same, with surface growth (alphas volumes depend on when the contact graph was last refreshed): exact when the state is taken
    right after a step that fully updated every aggregate (n_iter_without_event % full_aggregate_update_frequency == 0 before it)."""
    from golden_lib import Golden
    g = Golden("pytest_seed42")
    a = Oracle(g.base, g.overrides)
    a.run(448, record=False)   # step 446 merged; step 447 (n_iter_without_event == 0) fully updated every aggregate
    st = a.state()
    b = Oracle(g.base, g.overrides, construct=False)
    b.set_state(st, a.counters()["rand_calls"])
    ra, rb = a.run(1500), b.run(1500)
    assert ra["merged"].sum() >= 3
    for f in ra.dtype.names:
        if f == "step":
            continue
        np.testing.assert_array_equal(ra[f], rb[f], err_msg=f)
    sa, sb = a.state(), b.state()
    for k in ("rg", "volume", "surface", "time_step", "overlapping", "coordination_number"):
        np.testing.assert_array_equal(sa["aggregates"][k], sb["aggregates"][k], err_msg=k)
