"""Parity at BASELINE.json's full sizes (run on the B200 with -m gpu): the head of each trajectory is replayed against the oracle
step by step (as long as the O(N)-per-event CPU restatement finishes in seconds), the rest is checked through size-independent
properties of the domain: labels partition the spheres, sphere volume is conserved by aggregation, every committed contact is a
contact, the speculative batch width is invisible."""
import numpy as np
import pytest

from mcac_b200.configs import merged_config
from oracle_lib import Oracle
from test_gpu_parity import FP_FIELDS, INT_FIELDS, assert_records_match, assert_states_match

from mcac_b200 import HostModel, Simulation, ini_text

pytestmark = pytest.mark.gpu


def check_structure(st):
    """pymcac/tests/test_data.py:166-205: per time step bincount(sphere Label) == aggregate Np, labels are 0..N_agg-1."""
    assert st["sphere_label"].min() == 0 and st["sphere_label"].max() == st["n_agg"] - 1
    np.testing.assert_array_equal(np.bincount(st["sphere_label"], minlength=st["n_agg"]), st["agg_n_spheres"])
    assert st["offsets"][0] == 0 and st["offsets"][-1] == st["n_sph"]
    np.testing.assert_array_equal(np.diff(st["offsets"]), st["agg_n_spheres"])
    assert np.array_equal(np.sort(st["members"]), np.arange(st["n_sph"]))
    for a in np.nonzero(st["agg_n_spheres"] > 1)[0][:200]:
        assert np.all(st["sphere_label"][st["members"][st["offsets"][a]:st["offsets"][a + 1]]] == a)


def test_c2_polydisperse_1e5_replays_the_collision_sequence():
    """BASELINE configs[1]: validation/params_polydisperse.ini at 1e5 spheres (lognormal Dpm, n_verlet_divisions=40),
    single-GPU replay against the reference algorithm's collision sequence: 30 000 steps, every record and the final state."""
    ov = {"monomers": {"number": 100000}, "numerics": {"n_verlet_divisions": 40, "with_domain_duplication": "false", "random_seed": 42}}
    steps = 30000
    sim = Simulation(ini_text(merged_config("polydisperse", ov)))
    rep, recs = sim.run(steps, batch=256, records=steps)
    o = Oracle("polydisperse", ov)
    ref = o.run(steps)
    box = o.scalars()["box_length"]
    assert rep["steps"] == len(ref) == steps
    assert_records_match(recs, ref, box, clock_rtol=1e-10)  # 1e5 aggregates: tree-summed pick total (see the 1e6 test)
    assert rep["events"] == int(ref["merged"].sum()) and rep["events"] > 10
    st = sim.state()
    assert_states_match(st, o.state(), box, clock_rtol=1e-10)
    check_structure(st)
    assert rep["pair_tests_sphere"] == o.counters()["pair_sphere"]


def test_c3_1e6_head_replay_and_properties():
    """BASELINE configs[2] (the bench workload): 1e6 monodisperse spheres at FV = 1000 ppm.  First 600 steps against the oracle
    (each of its merges costs O(N)), then 30 000 more steps checked through properties, and batch-width invisibility at full size."""
    ov = {"monomers": {"number": 1000000}, "environment": {"volume_fraction": "1000e-6"}, "limits": {"physical_time": -1},
          "numerics": {"with_collisions": "true", "pick_method": "random", "n_verlet_divisions": 100, "with_domain_duplication": "false",
                       "random_seed": 42}}
    text = ini_text(merged_config("brownian", ov))
    sim = Simulation(text)
    st0 = sim.state()
    head = 600
    rep, recs = sim.run(head, batch=256, records=head)
    o = Oracle("brownian", ov)
    ref = o.run(head)
    box = o.scalars()["box_length"]
    # every decision and every geometric quantity to the usual bar; the clocks to 1e-10: at 1e6 aggregates the reference's
    # sequential sum of the pick weights and the device's fixed tree sum differ by ~2e-11 relative (deviation 1)
    assert_records_match(recs, ref, box, clock_rtol=1e-10)
    assert rep["events"] == int(ref["merged"].sum()) and rep["events"] >= 3
    assert_states_match(sim.state(), o.state(), box, clock_rtol=1e-10)
    del o
    more = 30000
    rep2, recs2 = sim.run(more, batch=256, records=more)
    st = sim.state()
    check_structure(st)
    assert st["n_agg"] == st0["n_agg"] - rep["events"] - rep2["events"]
    # aggregation moves spheres, it never changes them
    np.testing.assert_array_equal(st["spheres"]["r"], st0["spheres"]["r"])
    np.testing.assert_array_equal(st["spheres"]["volume"], st0["spheres"]["volume"])
    np.testing.assert_allclose(st["aggregates"]["volume"].sum(), st0["spheres"]["volume"].sum(), rtol=1e-12)
    # every merge was a contact within the drawn mean free path, between two different aggregates
    m = recs2["merged"] == 1
    assert m.sum() == rep2["events"] > 100
    assert np.all(recs2["distance"][m] <= recs2["full_distance"][m]) and np.all(recs2["moving_label"][m] != recs2["other_label"][m])
    # merged spheres touch: |c_i - c_j| == r_i + r_j to rounding, for the dimers of the final state
    dim = np.nonzero(st["agg_n_spheres"] == 2)[0]
    a, b = st["members"][st["offsets"][dim]], st["members"][st["offsets"][dim] + 1]
    d = np.sqrt(sum((st["spheres"][k][a] - st["spheres"][k][b]) ** 2 for k in ("rx", "ry", "rz")))
    np.testing.assert_allclose(d, st["spheres"]["r"][a] + st["spheres"]["r"][b], rtol=1e-9)
    # time only moves forward, one RNG triple per step
    assert np.all(np.diff(recs2["time"]) > 0) and np.all(np.diff(recs2["rand_calls"]) == 3)
    # the batch width is invisible at full size too
    s1, s2 = Simulation(text), Simulation(text)
    _, r1 = s1.run(3000, batch=1, records=3000)
    _, r2 = s2.run(3000, batch=512, records=3000)
    for f in INT_FIELDS + FP_FIELDS:
        np.testing.assert_array_equal(r1[f], r2[f], err_msg=f)
    # ... and so is the pipelined submission of the batches (no records requested: batch i+1 is submitted before batch i is read back)
    s3 = Simulation(text)
    r3, _ = s3.run(3000, batch=512)
    a2, a3 = s2.state(), s3.state()
    assert r3["steps"] == 3000 and a2["n_agg"] == a3["n_agg"] and a2["time"] == a3["time"]
    np.testing.assert_array_equal(a2["sphere_label"], a3["sphere_label"])
    for k in ("x", "y", "z"):
        np.testing.assert_array_equal(a2["spheres"][k], a3["spheres"][k])
    np.testing.assert_array_equal(a2["aggregates"]["proper_time"], a3["aggregates"]["proper_time"])
    assert r3["tie_sorts"] > 0  # the pick table of this configuration is tie-dominated: sparse fast path of the sort


def as_upload(st):
    """Simulation.state() -> the dict Simulation.upload() takes (field-major stacks in the reference's field order)."""
    import ref_trace as rt
    out = dict(st)
    out["sphere_fields"] = np.stack([st["spheres"][k] for k in rt.SPHERE_FIELDS])
    out["agg_fields"] = np.stack([st["aggregates"][k] for k in rt.AGG_FIELDS])
    out["per_member"] = np.stack([st["member_volumes"], st["member_surfaces"], st["member_distances_center"]])
    return out


def relative_steps(recs):
    out = recs.copy()
    out["step"] -= out["step"][0]
    return out


C3_OV = {"monomers": {"number": 1000000}, "environment": {"volume_fraction": "1000e-6"}, "limits": {"physical_time": -1},
         "numerics": {"with_collisions": "true", "pick_method": "random", "n_verlet_divisions": 100, "with_domain_duplication": "false",
                      "random_seed": 42}}


def test_c3_1e6_parity_over_the_window_the_bench_times(monkeypatch):
    """VERDICT r1 item 1.  `bench.py --steps 20 --warmup 5` times MC steps 100 000 - 500 000 of this workload; the regimes of the
    pick-table sort change along the way (sparse elements of the tie-dominated table: ~100, ~2 000 — overlap on —, ~3 000 — handed-over
    segment > 4096, overlap lost —, ~7 000, and > 8 192 — general replay back).  The run is taken to 700 000 steps and at six
    checkpoints the device state is downloaded and checked:
      (a) the pick table the event kernel builds == libstdc++'s std::sort order of max_dt / time_step on that state (index order
          bit-exact; cumulative table == the sequential sum on the tie path, to the documented 1e-10 where the general replay sums by its
          tree), and at one checkpoint also with the sparse path
          switched off (forced general replay on the same state);
      (b) the structure invariants (labels partition the spheres, CSR membership);
      (c) every multi-sphere aggregate's V / S / Rg / rmax / f_agg / d_m / lpm / time_step == Aggregate::update() of the oracle on the
          same members (the monomers' stored fields predate the initial radius rescale, aggregat_list_storage.cpp:75-87);
      (d) the oracle is re-synchronised from the downloaded state and the next 200 steps are replayed step by step.
    aggregat_list.cpp:109-141 (sort), aggregat.cpp:247-483 (update)."""
    from oracle_lib import introsort_order
    text = ini_text(merged_config("brownian", C3_OV))
    consumed0 = HostModel(text).state()["rand_consumed"]
    sim = Simulation(text)
    st0 = sim.state()
    o = Oracle("brownian", C3_OV, construct=False)
    box = st0["box_length"]
    done, events, seen_paths = 0, 0, set()
    for cp in [7500, 150000, 225000, 520000, 650000, 700000]:
        rep, _ = sim.run(cp - done, batch=256)  # no records: the pipelined submission the bench uses
        assert rep["steps"] == cp - done and rep["sort_fallbacks"] == 0
        done, events = cp, events + rep["events"]
        st = sim.state()
        check_structure(st)
        assert st["n_agg"] == st0["n_agg"] - events
        # ---- (a) the pick table of this state
        ts = st["aggregates"]["time_step"]
        keys = st["max_time_step"] / ts
        assert st["max_time_step"] == ts.max()
        ref = introsort_order(keys)
        sim.sort_time_steps(st["max_time_step"])
        idx, cum = sim.pick_table()
        np.testing.assert_array_equal(idx, ref, err_msg=f"pick table order at step {cp}")
        seq = np.cumsum(keys[ref])
        n_sparse = int((keys != keys.max()).sum())
        r0, _ = sim.run(0)
        path = "tie" if r0["tie_sorts"] > 0 else "general"
        assert path == ("tie" if n_sparse <= 8192 else "general"), (cp, n_sparse, r0["tie_sorts"])
        # a draw u picks another entry than the reference's sequential table would iff u lies between cum_dev[i] / total_dev and
        # cum_seq[i] / total_seq for some i
        p_flip = float(np.abs(cum / cum[-1] - seq / seq[-1]).sum())
        print(f"step {cp}: {path} path, pick-flip probability per draw {p_flip:.3e} ({len(cum)} entries)")
        if path == "tie":  # the sequential sum itself (sparse head one by one + the W run in closed form, csrc/seq_cumsum.cuh)
            np.testing.assert_array_equal(cum, seq)
        else:  # general replay above 65 536 entries: fixed summation tree (deviation 1)
            np.testing.assert_allclose(cum, seq, rtol=1e-10, atol=0)
            assert p_flip < 1e-4
        seen_paths.add((path, n_sparse > 2500))
        if cp == 150000:  # the same state through the general replay only
            monkeypatch.setenv("MCAC_B200_TIE_MIN_N", "0")
            gen = Simulation(text)
            monkeypatch.delenv("MCAC_B200_TIE_MIN_N")
            gen.upload(as_upload(st))
            gen.sort_time_steps(st["max_time_step"])
            gidx, gcum = gen.pick_table()
            assert gen.run(0)[0]["tie_sorts"] == 0
            np.testing.assert_array_equal(gidx, ref, err_msg="general replay on the same state")
            np.testing.assert_allclose(gcum, cum, rtol=1e-10, atol=0)  # (the general replay sums this size by the fixed tree)
            del gen
        # ---- (c) morphology of every aggregate that was ever updated
        consumed = consumed0 + 3 * done
        o.set_state(st, consumed)
        o.update_all()
        so = o.state()
        multi = st["agg_n_spheres"] > 1
        assert multi.sum() == n_sparse
        for k in ("volume", "surface", "rg", "rmax", "f_agg", "d_m", "lpm", "time_step", "dp", "dg_over_dp"):
            np.testing.assert_allclose(st["aggregates"][k][multi], so["aggregates"][k][multi], rtol=1e-12, atol=0, err_msg=f"{k} at step {cp}")
        # ---- (d) 200 steps in lock-step from here
        o.set_state(st, consumed)
        ref_recs = o.run(200)
        rep2, recs = sim.run(200, batch=256, records=200)
        assert rep2["steps"] == 200 and recs["rand_calls"][0] == consumed + 3
        assert_records_match(relative_steps(recs), relative_steps(ref_recs), box, clock_rtol=1e-10)
        assert rep2["events"] == int(ref_recs["merged"].sum())
        assert_states_match(sim.state(), o.state(), box, clock_rtol=1e-10)
        done, events = done + 200, events + rep2["events"]
    assert seen_paths == {("tie", False), ("tie", True), ("general", True)}, seen_paths


def test_c4_surface_growth_1e6_through_its_first_merge():
    """BASELINE configs[3] at size, with the overlap volume / surface update stressed: the window around the FIRST MERGE of
    the 1e6-sphere surface-growth run — located by a scout run on the device — is replayed step by step against the oracle, which is
    re-synchronised from the device state right after a full-update step (every aggregate's contact graph is fresh there: calcul.cpp:
    184-206 with full_aggregate_update_frequency = 100), at least 50 steps including the merge and the alphas volume / surface of the
    merged aggregate (aggregat.cpp:321-430)."""
    ov = {"monomers": {"number": 1000000}, "numerics": {"n_verlet_divisions": 100, "random_seed": 42}}
    text = ini_text(merged_config("surface_growth", ov))
    scout = Simulation(text)
    scout.set_stop_at_event(True)
    rep, _ = scout.run(4000)
    assert rep["events"] == 1, "no merge within 4000 steps"
    merge_step = rep["steps"] - 1
    del scout
    # the last full update before the merge window: n_iter_without_event % 100 == 0 at steps 0, 100, 200, ... (no event yet)
    start = ((merge_step - 40) // 100 * 100 + 1) if merge_step >= 141 else 0
    n_replay = max(50, merge_step - start + 11)
    sim = Simulation(text)
    consumed0 = HostModel(text).state()["rand_consumed"]
    o = Oracle("surface_growth", ov, construct=start == 0)
    if start > 0:
        r, _ = sim.run(start)
        assert r["steps"] == start and r["events"] == 0
        st = sim.state()
        o.set_state(st, consumed0 + 3 * start)
    rep2, recs = sim.run(n_replay, records=n_replay)
    ref = o.run(n_replay)
    box = o.scalars()["box_length"]
    assert rep2["steps"] == len(ref) == n_replay
    assert_records_match(relative_steps(recs), relative_steps(ref), box, clock_rtol=1e-10)
    assert rep2["events"] == int(ref["merged"].sum()) >= 1
    got, want = sim.state(), o.state()
    assert_states_match(got, want, box, clock_rtol=1e-10)
    check_structure(got)
    merged = np.nonzero(got["agg_n_spheres"] > 1)[0]
    assert len(merged) >= 1
    for k in ("volume", "surface", "overlapping", "coordination_number"):
        np.testing.assert_allclose(got["aggregates"][k][merged], want["aggregates"][k][merged], rtol=1e-12, atol=1e-300, err_msg=k)


def test_c4_surface_growth_1e6_first_steps():
    """BASELINE configs[3]: validation/params_surface_growth.ini at 1e6 spheres (lognormal 10 nm, alphas method): every step grows
    every sphere, updates every aggregate and re-sorts the pick table.  Three steps against the oracle, full state."""
    ov = {"monomers": {"number": 1000000}, "numerics": {"n_verlet_divisions": 100, "random_seed": 42}}
    steps = 3
    sim = Simulation(ini_text(merged_config("surface_growth", ov)))
    rep, recs = sim.run(steps, records=steps)
    o = Oracle("surface_growth", ov)
    ref = o.run(steps)
    box = o.scalars()["box_length"]
    assert rep["steps"] == len(ref) == steps
    assert_records_match(recs, ref, box, clock_rtol=1e-10)
    st = sim.state()
    assert_states_match(st, o.state(), box, clock_rtol=1e-10)
    check_structure(st)
    # growth: V = 4 pi / 3 r^3 and S = 4 pi r^2 of the grown radii (sphere.cpp:113-120: plain multiplies)
    r = st["spheres"]["r"]
    np.testing.assert_allclose(st["spheres"]["volume"], 4 * np.pi / 3 * r * r * r, rtol=1e-15)
    np.testing.assert_allclose(st["spheres"]["surface"], 4 * np.pi * r * r, rtol=1e-15)
