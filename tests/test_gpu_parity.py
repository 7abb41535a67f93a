"""Parity of the CUDA path (through the C ABI of libmcac_b200.so) against the oracle — run on the B200 with -m gpu.

Bar (BASELINE.json north_star): bit-exact for integer / indexing work (picked aggregate, RNG position, chosen
collision partner, merge order, labels, membership, cells); <= 1e-12 relative for FP64 quantities.  The pair test
itself only uses +,-,*,/,sqrt,fmod,floor, so on identical inputs it is compared BIT-EXACTLY.
"""
import numpy as np
import pytest

import ref_trace as rt
from golden_lib import Golden
from oracle.run_ref import merged_config
from oracle_lib import Oracle, rand_stream

import mcac_b200
from mcac_b200 import HostModel, Simulation, ini_text

pytestmark = pytest.mark.gpu
RTOL = 1e-12

INT_FIELDS = ["step", "rand_calls", "source", "moving_sphere", "other_sphere", "moving_label", "other_label", "n_agg", "merged", "n_try"]
FP_FIELDS = ["dir", "full_distance", "distance", "time", "dt", "proper_time", "pos"]


def assert_records_match(got, ref, box, clock_rtol=RTOL):
    n = min(len(got), len(ref))
    assert n > 0
    for f in INT_FIELDS:
        bad = np.nonzero(got[f][:n] != ref[f][:n])[0]
        assert len(bad) == 0, f"{f}: first divergence at step {bad[0]}: gpu {got[f][bad[0]]} vs oracle {ref[f][bad[0]]}"
    fin = np.isfinite(ref["distance"][:n])
    assert np.array_equal(fin, np.isfinite(got["distance"][:n]))
    # the contact distance is a difference (proj - sqrt(R^2 - axis^2)): 1e-12 relative to the scale of its operands
    # (the displacement length); on bit-identical inputs it is checked bit-exactly in test_contact_search_*
    err = np.abs(got["distance"][:n][fin] - ref["distance"][:n][fin])
    assert np.all(err <= RTOL * ref["full_distance"][:n][fin]), f"contact distance off by {err.max()}"
    np.testing.assert_allclose(got["distance"][:n][fin], ref["distance"][:n][fin], rtol=1e-9, atol=0)
    np.testing.assert_allclose(got["dir"][:n], ref["dir"][:n], rtol=0, atol=1e-15)  # unit vector: a few ulp of CUDA sincos/acos vs glibc
    np.testing.assert_allclose(got["full_distance"][:n], ref["full_distance"][:n], rtol=RTOL, atol=0, err_msg="full_distance")
    # clocks: dt = max_dt / cumulative_time_steps.back().  Above 65 536 aggregates that total is a fixed tree sum here and a
    # sequential sum in the reference (whose own rounding error grows like n * 2^-53): callers at such sizes pass clock_rtol
    for f in ["time", "proper_time"]:
        np.testing.assert_allclose(got[f][:n], ref[f][:n], rtol=clock_rtol, atol=0, err_msg=f)
    # dt of a contact step = dt * (contact distance / lpm): inherits the conditioning of the contact distance
    np.testing.assert_allclose(got["dt"][:n], ref["dt"][:n], rtol=clock_rtol, atol=RTOL * ref["dt"][:n].max(), err_msg="dt")
    np.testing.assert_allclose(got["pos"][:n], ref["pos"][:n], rtol=0, atol=RTOL * box, err_msg="pos")


def assert_states_match(got, ref, box, clock_rtol=RTOL):
    assert got["n_sph"] == ref["n_sph"] and got["n_agg"] == ref["n_agg"]
    for k in ["sphere_label", "agg_n_spheres", "members", "offsets", "agg_cell"]:
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)
    for k in ["x", "y", "z", "rx", "ry", "rz"]:
        np.testing.assert_allclose(got["spheres"][k], ref["spheres"][k], rtol=0, atol=RTOL * box, err_msg=f"sphere {k}")
    for k in ["r", "volume", "surface"]:
        np.testing.assert_allclose(got["spheres"][k], ref["spheres"][k], rtol=RTOL, atol=0, err_msg=f"sphere {k}")
    for k in ["x", "y", "z", "rx", "ry", "rz"]:
        np.testing.assert_allclose(got["aggregates"][k], ref["aggregates"][k], rtol=0, atol=RTOL * box, err_msg=f"aggregate {k}")
    for k in ["rg", "f_agg", "lpm", "time_step", "rmax", "volume", "surface", "dp", "dg_over_dp", "coordination_number", "d_m"]:
        np.testing.assert_allclose(got["aggregates"][k], ref["aggregates"][k], rtol=RTOL, atol=1e-300, err_msg=f"aggregate {k}")
    np.testing.assert_allclose(got["aggregates"]["proper_time"], ref["aggregates"]["proper_time"], rtol=clock_rtol, atol=1e-300, err_msg="aggregate proper_time")
    # mean overlap coefficient c_ij = (r_i + r_j - d) / (r_i + r_j) is a ratio in [0, 1]; for spheres that just touch it is pure
    # rounding noise (1e-16), so it is compared on the scale of the ratio
    np.testing.assert_allclose(got["aggregates"]["overlapping"], ref["aggregates"]["overlapping"], rtol=RTOL, atol=RTOL, err_msg="overlapping")
    for k in ["member_volumes", "member_surfaces"]:
        np.testing.assert_allclose(got[k], ref[k], rtol=RTOL, atol=0, err_msg=k)
    np.testing.assert_allclose(got["member_distances_center"], ref["member_distances_center"], rtol=0, atol=RTOL * box)
    np.testing.assert_allclose(got["time"], ref["time"], rtol=clock_rtol, err_msg="time")
    for k in ["box_length", "maxradius", "max_time_step", "avg_npp"]:
        np.testing.assert_allclose(got[k], ref[k], rtol=RTOL, err_msg=k)


def test_device_rng_is_the_glibc_stream():
    text = ini_text(merged_config("monodisperse", {"numerics": {"random_seed": 42}, "monomers": {"number": 50}}))
    sim = Simulation(text)
    consumed = HostModel(text).state()["rand_consumed"]
    n = 3_000_000  # crosses two refills of the device buffer
    got = sim.rand(n)
    ref = rand_stream(42, consumed + n)[consumed:]
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("base,ov", [
    ("monodisperse", {"numerics": {"random_seed": 42}}),
    ("polydisperse", {"numerics": {"random_seed": 42, "n_verlet_divisions": 12}, "monomers": {"number": 4000}}),
    ("brownian", {"numerics": {"random_seed": 42, "n_verlet_divisions": 16, "with_collisions": "true", "pick_method": "random"},
                  "environment": {"volume_fraction": "1000e-6"}, "monomers": {"number": 4000}, "limits": {"physical_time": -1}}),
])
def test_contact_search_on_initial_state_bit_exact(base, ov):
    """K1 vs AggregatList::distance_to_next_contact on identical state and identical directions: bit-exact."""
    text = ini_text(merged_config(base, ov))
    sim = Simulation(text)
    o = Oracle(base, ov)
    st = o.state()
    rng = np.random.default_rng(1)
    n_agg = st["n_agg"]
    labels = rng.integers(0, n_agg, 600)
    v = rng.normal(size=(600, 3)); v /= np.linalg.norm(v, axis=1)[:, None]
    # long sweeps so that a good share of the queries really hit something
    dist = st["aggregates"]["lpm"][labels] * rng.choice([1.0, 30.0, 300.0], 600)
    got, pairs = sim.contact_search_batch(labels, v, dist)
    hits = 0
    for q in range(600):
        d, ids = o.search(int(labels[q]), v[q], float(dist[q]))
        if np.isinf(d):
            assert np.isinf(got["distance"][q])
            continue
        hits += 1
        assert got["distance"][q] == d, (q, got["distance"][q], d)
        assert (got["moving_sphere"][q], got["other_sphere"][q], got["moving_label"][q], got["other_label"][q]) == tuple(ids)
    assert hits > 20
    c = o.counters()
    assert pairs[0] == c["pair_sphere"] and pairs[1] == c["pair_bounding"]


@pytest.mark.parametrize("name,steps,batch", [("monodisperse_seed42", 60000, 128), ("c3_small_seed42", 40000, 256),
                                              ("c2_small_seed42", 30000, 256), ("polydisperse_seed42", 30000, 64)])
def test_run_replays_the_collision_sequence(name, steps, batch):
    """The device-resident loop (speculative batches + in-order commit) against the oracle, step by step."""
    g = Golden(name)
    text = ini_text(merged_config(g.base, g.overrides))
    sim = Simulation(text)
    rep, recs = sim.run(steps, batch=batch, records=steps)
    o = Oracle(g.base, g.overrides)
    ref = o.run(steps)
    box = o.scalars()["box_length"]
    assert rep["steps"] == len(ref)
    assert_records_match(recs, ref, box)
    assert rep["events"] == int(ref["merged"].sum())
    c = o.counters()
    # sphere-level tests of the examined suspects: exact.  Bounding prefilter tests: a speculative search sees the Verlet
    # cells as they were at the start of its batch, so the SIZE of its superset neighbourhood may differ by a few entries
    # from the sequential run although the eligible suspects (and hence every result) are identical.
    assert rep["pair_tests_sphere"] == c["pair_sphere"]
    assert abs(rep["pair_tests_bounding"] - c["pair_bounding"]) <= 1e-3 * c["pair_bounding"]
    assert_states_match(sim.state(), o.state(), box)
    # structural invariant pinned by the reference (pymcac/tests/test_data.py:166-205)
    st = sim.state()
    np.testing.assert_array_equal(np.bincount(st["sphere_label"], minlength=st["n_agg"]), st["agg_n_spheres"])


@pytest.mark.parametrize("name", ["monodisperse_seed42", "polydisperse_seed42", "c3_small_seed42", "c2_small_seed42"])
def test_device_sort_replays_std_sort_tie_order(name):
    """index_sorted_time_steps / cumulative_time_steps of the reference's first sort_time_steps call (taps: sort_0) —
    monodisperse: every weight ties, so this is purely libstdc++'s introsort order, replayed on the device."""
    g = Golden(name)
    srt = g.sort(0)
    sim = Simulation(ini_text(merged_config(g.base, g.overrides)))
    sim.sort_time_steps(float(srt["factor"]))
    idx, cum = sim.pick_table()
    np.testing.assert_array_equal(idx, srt["idx"])
    np.testing.assert_array_equal(cum, srt["cum"])  # n <= 65536: summed sequentially, the reference's rounding


@pytest.mark.parametrize("n,env", [(30000, {}), (30000, {"MCAC_B200_NO_SORT_WINDOWS": "1"}), (30000, {"MCAC_B200_SORT_LOCAL": "300"}),
                                   (250000, {})])
def test_device_sort_against_std_sort_on_random_weights(n, env, monkeypatch):
    """Same check on synthetic weight patterns (few classes, all equal, sorted, reversed, random) through the C ABI:
    the aggregates' time steps are overwritten by uploading a doctored state.  Variants: the block-local levels by one CTA per
    window of segments (default), by block 0 alone, with a small staging area, and on a table whose first levels stay grid-wide."""
    from oracle_lib import introsort_order
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    text = ini_text(merged_config("monodisperse", {"numerics": {"random_seed": 5, "n_verlet_divisions": 8}, "monomers": {"number": n},
                                                   "environment": {"volume_fraction": "10e-6"}}))
    hm = HostModel(text).state()
    rng = np.random.default_rng(3)
    assert n == hm["n_agg"]
    patterns = [rng.random(n) + 0.5, rng.integers(1, 4, n).astype(float), np.ones(n), np.sort(rng.integers(1, 60, n)).astype(float),
                np.sort(rng.random(n) + 0.5)[::-1].copy()]
    for ts in patterns:
        st = dict(hm)
        af = hm["agg_fields"].copy()
        af[3] = ts  # TIME_STEP column
        st["agg_fields"] = af
        sim = Simulation(text)
        sim.upload(st)
        sim.sort_time_steps(2.0)
        idx, cum = sim.pick_table()
        keys = 2.0 / ts
        ref = introsort_order(keys)
        np.testing.assert_array_equal(idx, ref)
        if n <= 65536:
            np.testing.assert_array_equal(cum, np.cumsum(keys[ref]))  # np.cumsum is the sequential sum


@pytest.mark.parametrize("depth", [1, 2, 3, 6])
def test_device_sort_replays_the_heap_sort_branch(depth, monkeypatch):
    """introsort's depth limit forced to a small value (MCAC_B200_SORT_DEPTH): the event kernel hands the sort back, the multi-launch
    device path replays the partition levels down to the limit and then libstdc++'s heap-sort branch (k_sort_heap,
    csrc/heap_sort.cuh) — the order must be that of libstdc++'s own __introsort_loop run with the same limit.  No host sort."""
    from oracle_lib import introsort_order_depth
    monkeypatch.setenv("MCAC_B200_SORT_DEPTH", str(depth))
    text = ini_text(merged_config("monodisperse", {"numerics": {"random_seed": 5, "n_verlet_divisions": 8}, "monomers": {"number": 3000}}))
    hm = HostModel(text).state()
    rng = np.random.default_rng(11 + depth)
    n = hm["n_agg"]
    for ts in [rng.random(n) + 0.5, rng.integers(1, 4, n).astype(float), np.ones(n)]:
        st = dict(hm)
        af = hm["agg_fields"].copy()
        af[3] = ts
        st["agg_fields"] = af
        sim = Simulation(text)
        sim.upload(st)
        sim.sort_time_steps(2.0)
        idx, cum = sim.pick_table()
        keys = 2.0 / ts
        ref = introsort_order_depth(keys, depth)
        np.testing.assert_array_equal(idx, ref)
        np.testing.assert_array_equal(cum, np.cumsum(keys[ref]))
        rep, _ = sim.run(0)
        assert rep["sort_heap_branches"] >= 1 and rep["sort_fallbacks"] >= 1


def test_suspect_list_overflow_is_an_error_not_a_wrong_contact(monkeypatch):
    """A contact search with more eligible suspects than its list holds does not know the first contact: the run must stop
    with an error instead of committing whichever suspects won the race.  The list is shrunk to 2 entries (MCAC_B200_CAND_CAP) in a
    dense box searched by the wide kernel only."""
    monkeypatch.setenv("MCAC_B200_CAND_CAP", "2")
    monkeypatch.setenv("MCAC_B200_SEARCH_GROUP", "0")
    text = ini_text(merged_config("monodisperse", {"numerics": {"random_seed": 3, "n_verlet_divisions": 3}, "monomers": {"number": 600},
                                                   "environment": {"volume_fraction": "0.05"}}))
    sim = Simulation(text)
    with pytest.raises(mcac_b200.McacError) as e:
        sim.run(20000, batch=64)
    assert "more eligible suspects" in str(e.value)
    monkeypatch.delenv("MCAC_B200_CAND_CAP")
    ok = Simulation(text)
    rep, _ = ok.run(1000, batch=64)
    assert rep["steps"] == 1000


@pytest.mark.parametrize("n,frac,classes,local", [(30000, 0.003, 3, 64), (30000, 0.05, 0, 4096), (200000, 0.0, 1, 4096),
                                                  (200000, 0.002, 0, 4096), (200000, 0.03, 4, 512), (65537, 0.01, 2, 4096)])
def test_tie_dominated_sort_fast_path_is_std_sort(n, frac, classes, local, monkeypatch):
    """Tables where one weight dominates and is the largest (monodisperse runs: every monomer ties) take the sparse fast path of
    the event kernel (csrc/tie_sort.cuh); its permutation must be libstdc++'s std::sort order, like the general replay's."""
    from oracle_lib import introsort_order
    monkeypatch.setenv("MCAC_B200_TIE_MIN_N", "1000")
    monkeypatch.setenv("MCAC_B200_SORT_LOCAL", str(local))
    text = ini_text(merged_config("monodisperse", {"numerics": {"random_seed": 5, "n_verlet_divisions": 8}, "monomers": {"number": n},
                                                   "environment": {"volume_fraction": "10e-6"}}))
    hm = HostModel(text).state()
    assert hm["n_agg"] == n
    rng = np.random.default_rng(n + local)
    ts = np.full(n, 0.5)
    sparse = rng.random(n) < frac
    ts[sparse] = (0.5 + rng.integers(1, classes + 1, sparse.sum())) if classes else (0.5 + rng.random(sparse.sum()))
    st = dict(hm)
    af = hm["agg_fields"].copy()
    af[3] = ts  # TIME_STEP column
    st["agg_fields"] = af
    sim = Simulation(text)
    sim.upload(st)
    sim.sort_time_steps(2.0)
    idx, cum = sim.pick_table()
    keys = 2.0 / ts
    ref = introsort_order(keys)
    np.testing.assert_array_equal(idx, ref)
    np.testing.assert_array_equal(keys[idx], np.sort(keys))
    # the reference's sequential sum, bit for bit, at any size: one addition after the other up to 65 536 entries, and above that the
    # sparse head added up one by one + the run of equal weights in closed form (csrc/seq_cumsum.cuh)
    np.testing.assert_array_equal(cum, np.cumsum(keys[ref]))  # np.cumsum is the sequential sum
    rep, _ = sim.run(0)
    assert rep["tie_sorts"] >= 1 and rep["tie_sparse"] >= int(sparse.sum())


@pytest.mark.parametrize("name,steps", [("surface_growth_seed42", 2500), ("caps_seed7", 4000), ("pytest_seed42", 20000),
                                        ("brownian_seed42", 4000)])
def test_general_step_loop_growth_picklast_nocollision(name, steps):
    """Configurations without a fixed pick sequence (surface growth: alphas / caps, all aggregates updated every step;
    pick_last without collisions) run one MC step per launch sequence in calcul()'s order — same records as the oracle.
    pytest_seed42 runs to its end (NPP_avg limit) through one domain duplication."""
    g = Golden(name)
    text = ini_text(merged_config(g.base, g.overrides))
    sim = Simulation(text)
    rep, recs = sim.run(steps, records=steps)
    o = Oracle(g.base, g.overrides)
    ref = o.run(steps)
    box = o.scalars()["box_length"]
    assert rep["steps"] == len(ref)
    assert rep["finished"] == int(o.finished)
    assert_records_match(recs, ref, box)
    assert rep["events"] == int(ref["merged"].sum())
    assert_states_match(sim.state(), o.state(), box)


@pytest.mark.parametrize("headroom", [None, "96"])
def test_classic_potentials_nucleation_individual_growth(tmp_path, monkeypatch, headroom):
    """examples/classic.ini (the ensemble workload): external interaction potentials with direction redraws, nucleation,
    individual surface reactions, normal-law diameters, three domain duplications within the first 1500 steps.
    headroom=96: the slot headroom for nucleated monomers is made tiny, so the state is re-laid out (slots renumbered) every few
    dozen nucleations — the pick table must be built after that, not before (regression: a stale table was read once)."""
    from golden_lib import write_interpotential_file
    if headroom:
        monkeypatch.setenv("MCAC_B200_NUCL_HEADROOM", headroom)
    g = Golden("classic_seed1000")
    ov = dict(g.overrides)
    ov["inter_potential"] = {"interpotential_file": write_interpotential_file(tmp_path / "Interpotential_input.dat")}
    steps = 1500
    sim = Simulation(ini_text(merged_config(g.base, ov)))
    rep, recs = sim.run(steps, records=steps)
    o = Oracle(g.base, ov)
    ref = o.run(steps)
    box = o.scalars()["box_length"]
    assert rep["steps"] == len(ref)
    assert_records_match(recs, ref, box)
    assert rep["events"] == o.counters()["events"]
    assert_states_match(sim.state(), o.state(), box)


def test_per_call_entry_points_follow_the_reference_methods():
    """translate / grow / update(partial, full) / merge / refresh called one by one through the C ABI (the per-call mode a
    reference shim would use) against the oracle driven the same way."""
    ov = {"numerics": {"random_seed": 11}, "monomers": {"number": 300}, "surface_growth": {"volsurf_method": "caps"},
          "environment": {"volume_fraction": "3000e-6"}}
    text = ini_text(merged_config("pytest", ov))
    sim = Simulation(text)
    o = Oracle("pytest", ov)
    st0 = o.state()
    rng = np.random.default_rng(0)
    box = st0["box_length"]
    # find a pair that collides: sweep label 0..n with long moves until the search reports a contact
    contact = None
    for lab in range(st0["n_agg"]):
        v = rng.normal(size=3); v /= np.linalg.norm(v)
        c = sim.contact_search(lab, v, box)
        d, ids = o.search(lab, v, box)
        assert c.distance == d
        if np.isfinite(d):
            contact = (lab, v, c)
            break
    assert contact is not None
    lab, v, c = contact
    sim.translate(lab, v * c.distance)
    assert sim.merge(c)
    st = sim.state()
    assert st["n_agg"] == st0["n_agg"] - 1
    kept = min(c.moving_label, c.other_label)
    assert st["agg_n_spheres"][kept] == 2
    np.testing.assert_array_equal(np.bincount(st["sphere_label"], minlength=st["n_agg"]), st["agg_n_spheres"])
    # growth of every sphere then a full update of every aggregate: volumes follow r^3, caps keep V below the sum
    r_before = st["spheres"]["r"].copy()
    sim.grow(1e-3)
    sim.update(-1, full=True)
    st2 = sim.state()
    u_sg = st2["u_sg"]
    np.testing.assert_allclose(st2["spheres"]["r"], r_before + u_sg * 1e-3, rtol=1e-15)
    np.testing.assert_allclose(st2["spheres"]["volume"], 4 * np.pi / 3 * st2["spheres"]["r"] ** 3, rtol=1e-14)
    members = st2["members"][st2["offsets"][kept]:st2["offsets"][kept + 1]]
    assert st2["aggregates"]["volume"][kept] < st2["spheres"]["volume"][members].sum()  # lens caps removed
    ref = sim.refresh()
    np.testing.assert_allclose(ref["total_volume"], st2["aggregates"]["volume"].sum(), rtol=1e-13)
    assert ref["max_time_step"] == st2["aggregates"]["time_step"].max()


def test_batch_width_does_not_change_the_trajectory():
    """Speculation must be invisible: batch = 1 (pure sequential) and batch = 512 give bit-identical device results."""
    text = ini_text(merged_config("monodisperse", {"numerics": {"random_seed": 3}}))
    a = Simulation(text); b = Simulation(text)
    ra, reca = a.run(20000, batch=1, records=20000)
    rb, recb = b.run(20000, batch=512, records=20000)
    assert ra["steps"] == rb["steps"]
    for f in INT_FIELDS + FP_FIELDS:
        np.testing.assert_array_equal(reca[f], recb[f], err_msg=f)
    assert rb["batches"] < ra["batches"]


@pytest.mark.parametrize("env", [{"MCAC_B200_FORCE_SORT_FAIL": "3"}, {"MCAC_B200_NO_OVERLAP": "1"}, {"MCAC_B200_SEARCH_GROUP": "0"},
                                 {"MCAC_B200_SEARCH_GROUP": "4", "MCAC_B200_SEARCH_MB": "4"}, {"MCAC_B200_NO_SORT_SMEM": "1"},
                                 {"MCAC_B200_TIE_MIN_N": "100", "MCAC_B200_SORT_LOCAL": "64"},
                                 {"MCAC_B200_TIE_MIN_N": "100", "MCAC_B200_SORT_LOCAL": "512", "MCAC_B200_TIE_MAX_SPARSE": "50"},
                                 {"MCAC_B200_TIE_MIN_N": "100", "MCAC_B200_SORT_LOCAL": "64", "MCAC_B200_TIE_NO_OVERLAP": "1"}])
def test_execution_variants_do_not_change_the_trajectory(env, monkeypatch):
    """Scheduling choices must be invisible in the results: a device sort that gives up (every 3rd one here) falls back to the
    multi-launch sort and the batch is redone; serialised vs overlapped cell rebuild; wide-only vs narrow-group contact search;
    the sparse fast path of the pick-table sort (tie_sort.cuh) forced onto this small table, with and without running out of room."""
    g = Golden("c3_small_seed42")
    text = ini_text(merged_config(g.base, g.overrides))
    base = Simulation(text)
    r0, rec0 = base.run(15000, batch=256, records=15000)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    var = Simulation(text)
    r1, rec1 = var.run(15000, batch=256, records=15000)
    assert r0["steps"] == r1["steps"] and r0["events"] == r1["events"] and r0["events"] > 10
    assert r0["tie_sorts"] == 0
    if "MCAC_B200_TIE_MIN_N" in env:
        assert r1["tie_sorts"] > 10 and r1["tie_levels"] >= 2 * r1["tie_sorts"]
        if "MCAC_B200_TIE_MAX_SPARSE" in env:
            assert r1["tie_sorts"] < r1["sorts"]  # more than 50 non-monomers later in the run: general sort again
    for f in INT_FIELDS + FP_FIELDS:
        np.testing.assert_array_equal(rec0[f], rec1[f], err_msg=f)
    s0, s1 = base.state(), var.state()
    np.testing.assert_array_equal(s0["sphere_label"], s1["sphere_label"])
    np.testing.assert_array_equal(s0["aggregates"]["rg"], s1["aggregates"]["rg"])


@pytest.mark.parametrize("env", [{}, {"MCAC_B200_FORCE_SORT_FAIL": "3"}, {"MCAC_B200_TIE_MIN_N": "100", "MCAC_B200_SORT_LOCAL": "64"}])
def test_pipelined_submission_is_invisible(env, monkeypatch):
    """Without step records mcac_gpu_run submits batch i+1 (event kernel, cell rebuild, queries, search, commit) before it has read
    batch i back; the kernels enforce the step limit / finished() / pool room themselves.  The state after any number of steps must be
    bit-identical to the one-batch-at-a-time loop (MCAC_B200_NO_PIPELINE=1), also across calls that stop mid-way, with device sorts
    that give up, and with the sparse fast path of the sort."""
    g = Golden("c3_small_seed42")
    text = ini_text(merged_config(g.base, g.overrides))
    monkeypatch.setenv("MCAC_B200_NO_PIPELINE", "1")
    base = Simulation(text)
    r0, _ = base.run(15000, batch=256)
    monkeypatch.delenv("MCAC_B200_NO_PIPELINE")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    one = Simulation(text)
    r1, _ = one.run(15000, batch=256)
    parts = Simulation(text)
    done = 0
    for chunk in (1, 700, 5000, 513, 8786):
        rp, _ = parts.run(chunk, batch=256)
        assert rp["steps"] == chunk
        done += chunk
    assert done == 15000 and r0["steps"] == r1["steps"] == 15000 and r0["events"] == r1["events"] > 10
    s0 = base.state()
    for other in (one.state(), parts.state()):
        assert s0["time"] == other["time"] and s0["n_agg"] == other["n_agg"]
        np.testing.assert_array_equal(s0["sphere_label"], other["sphere_label"])
        for k in ("x", "y", "z"):
            np.testing.assert_array_equal(s0["spheres"][k], other["spheres"][k])
        for k in ("rg", "proper_time", "time_step", "volume"):
            np.testing.assert_array_equal(s0["aggregates"][k], other["aggregates"][k])


def test_big_aggregate_search_form_is_the_same_search(monkeypatch):
    """The three-kernel search (phase 1 / sphere-pair tiles over the grid / phase 3) that single searches take once aggregates
    hold many spheres returns exactly what the one-CTA search returns: forced on for a whole growth run (pytest config, to its
    end) and compared with the default run, then both with the oracle."""
    g = Golden("pytest_seed42")
    text = ini_text(merged_config(g.base, g.overrides))
    base = Simulation(text)
    r0, rec0 = base.run(20000, records=20000)
    monkeypatch.setenv("MCAC_B200_BIG_NPP", "0")
    var = Simulation(text)
    r1, rec1 = var.run(20000, records=20000)
    assert r0["steps"] == r1["steps"] and r0["events"] == r1["events"] and r0["finished"] == 1
    for f in INT_FIELDS + FP_FIELDS:
        np.testing.assert_array_equal(rec0[f], rec1[f], err_msg=f)
    o = Oracle(g.base, g.overrides)
    ref = o.run(20000)
    assert_records_match(rec1, ref, o.scalars()["box_length"])


def test_monodisperse_full_run_through_two_duplications():
    """C1 literal: 800 -> 6400 -> 51200 spheres, 1 000 452 steps, 1 688 merges — decisions via the golden digest."""
    g = Golden("monodisperse_seed42")
    sim = Simulation(ini_text(merged_config(g.base, g.overrides)))
    rep, recs = sim.run(2_000_000, batch=256, records=1_100_000)
    assert rep["finished"] == 1
    assert rep["steps"] == g.meta["total_steps"]
    assert rep["events"] == 1688 and rep["duplications"] == 2
    np.testing.assert_array_equal(np.nonzero(recs["merged"])[0], g.merges["step"][g.merges["ok"] == 1])
    fin = g.state("state_final")
    st = sim.state()
    np.testing.assert_array_equal(st["sphere_label"], fin["sphere_label"])
    np.testing.assert_array_equal(st["members"], fin["members"])
    np.testing.assert_allclose(st["aggregates"]["rg"], fin["aggregates"]["rg"], rtol=1e-9)
    np.testing.assert_allclose(st["time"], fin["time"], rtol=1e-9)


LOOP_CASES = [("classic_seed1000", 2200), ("pytest_seed42", 4000), ("brownian_seed42", 1500), ("caps_seed7", 3000), ("surface_growth_seed42", 800)]


@pytest.mark.parametrize("name,steps", LOOP_CASES)
@pytest.mark.parametrize("env", [{"MCAC_B200_NO_LOOP": "1"}, {"MCAC_B200_LOOP_MAX_SLOTS": "150"}, {"MCAC_B200_NO_PRUNE": "1"},
                                 {"MCAC_B200_NO_LOOP_DUP": "1"}])
def test_step_loop_is_the_multi_launch_general_step(name, steps, env, monkeypatch, tmp_path):
    """The per-realization step loop (csrc/mcac_steploop.cuh: the whole general step of calcul() in one persistent CTA, in-kernel pool
    compaction, ordered CTA-wide sphere sweep) against the multi-launch sequence of the same device functions — bit-identical records
    and states, across duplications (in place on the device vs the host re-layout through the upload boundary, MCAC_B200_NO_LOOP_DUP) /
    nucleation regrows / the hand-over when the aggregate table outgrows the loop (MCAC_B200_LOOP_MAX_SLOTS).  Both are checked against the oracle elsewhere in this file; here: same trajectory, far fewer launches."""
    from golden_lib import write_interpotential_file
    g = Golden(name)
    ov = {k: dict(v) for k, v in g.overrides.items()}
    if g.base == "classic":
        ov.setdefault("inter_potential", {})["interpotential_file"] = write_interpotential_file(tmp_path / "Interpotential_input.dat")
    text = ini_text(merged_config(g.base, ov))
    loop = Simulation(text)
    r0, rec0 = loop.run(steps, records=steps)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    other = Simulation(text)
    r1, rec1 = other.run(steps, records=steps)
    assert r0["steps"] == r1["steps"] == len(rec0) and r0["events"] == r1["events"] and r0["duplications"] == r1["duplications"]
    assert r0["nucleated"] == r1["nucleated"] and r0["pair_tests_sphere"] == r1["pair_tests_sphere"]
    for f in INT_FIELDS + FP_FIELDS:
        np.testing.assert_array_equal(rec0[f], rec1[f], err_msg=f)
    s0, s1 = loop.state(), other.state()
    for k in ["sphere_label", "agg_n_spheres", "members", "offsets", "agg_cell"]:
        np.testing.assert_array_equal(s0[k], s1[k], err_msg=k)
    for k in s0["spheres"]:
        np.testing.assert_array_equal(s0["spheres"][k], s1["spheres"][k], err_msg=k)
    for k in s0["aggregates"]:
        np.testing.assert_array_equal(s0["aggregates"][k], s1["aggregates"][k], err_msg=k)
    assert s0["time"] == s1["time"] and s0["max_time_step"] == s1["max_time_step"]
    np.testing.assert_allclose(s0["volume_fraction"], s1["volume_fraction"], rtol=1e-13)  # totals: summed in a different fixed order
    if "MCAC_B200_NO_LOOP" in env:
        assert r0["kernel_launches"] * 20 < r1["kernel_launches"], (r0["kernel_launches"], r1["kernel_launches"])
    if "MCAC_B200_NO_PRUNE" in env:  # the pruned sweep executes fewer pair tests than the reference runs; unpruned, exactly as many
        assert r1["pair_tests_executed"] == r1["pair_tests_sphere"]
        assert r0["pair_tests_executed"] <= r0["pair_tests_sphere"] + 2 * r0["searches"] * max(1, r0["n_spheres"])
        if name == "classic_seed1000":
            assert r0["pair_tests_executed"] * 2 < r0["pair_tests_sphere"], (r0["pair_tests_executed"], r0["pair_tests_sphere"])


@pytest.mark.parametrize("name,steps", [("c3_small_seed42", 20000), ("monodisperse_seed42", 3000), ("polydisperse_seed42", 3000),
                                        ("classic_seed1000", 1500)])
def test_strict_direction_mode_meets_the_1e12_bar_on_contact_distances(name, steps, tmp_path):
    """BASELINE north_star: <= 1e-12 relative for the FP64 contact distance.  The only libm calls between the RNG stream and a contact
    distance are sin / cos / acos of random_direction() (tools.cpp:82-89); CUDA's versions differ from glibc's by <= 2 ulp, which a
    grazing contact amplifies (the default-mode tests allow 1e-9 there).  In strict replay mode (mcac_gpu_set_strict_direction) the
    directions come from the host's glibc: they must equal the oracle's BIT FOR BIT, and every contact distance to 1e-12 of itself."""
    from golden_lib import write_interpotential_file
    g = Golden(name)
    ov = {k: dict(v) for k, v in g.overrides.items()}
    if g.base == "classic":
        ov.setdefault("inter_potential", {})["interpotential_file"] = write_interpotential_file(tmp_path / "Interpotential_input.dat")
    sim = Simulation(ini_text(merged_config(g.base, ov)))
    sim.set_strict_direction(True)
    rep, recs = sim.run(steps, batch=256, records=steps)
    o = Oracle(g.base, ov)
    ref = o.run(steps)
    n = len(ref)
    assert rep["steps"] == n == len(recs)
    for f in INT_FIELDS:
        np.testing.assert_array_equal(recs[f], ref[f], err_msg=f)
    np.testing.assert_array_equal(recs["dir"], ref["dir"])
    fin = np.isfinite(ref["distance"])
    assert fin.sum() >= 3
    np.testing.assert_allclose(recs["distance"][fin], ref["distance"][fin], rtol=1e-12, atol=0)
    np.testing.assert_allclose(recs["full_distance"], ref["full_distance"], rtol=1e-12, atol=0)


def test_single_aggregate_getters_read_the_device_state():
    """Aggregate::get_lpm / get_time_step / size (include/aggregats/aggregat.hpp:83-133) through mcac_gpu_aggregate_fields: the 21
    AggregatesFields of one label equal the downloaded state's row."""
    g = Golden("c3_small_seed42")
    sim = Simulation(ini_text(merged_config(g.base, g.overrides)))
    sim.run(6000, batch=128)
    st = sim.state()
    for label in (0, 17, st["n_agg"] - 1, int(np.argmax(st["agg_n_spheres"]))):
        f, n = sim.aggregate_fields(label)
        assert n == st["agg_n_spheres"][label]
        for k, v in f.items():
            if k == "electric_charge_field":
                continue
            assert v == st["aggregates"][k][label], (label, k)
    with pytest.raises(mcac_b200.McacError):
        sim.aggregate_fields(st["n_agg"])
