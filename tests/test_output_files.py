"""The reference's output layout (src/io/*: Spheres_<k>.{h5,xmf}, Aggregats_<k>.{h5,xmf}) written by the product's libhdf5-free writer
(mcac_b200/host/xdmf_writer.cpp).  No libhdf5 / h5py exists in this image, so the files are read back by tests/h5_min_reader.py, which
walks every structure of the HDF5 format the way libhdf5 does and restates pymcac's XMF reader."""
import ctypes as C

import numpy as np
import pytest

import mcac_b200
from h5_min_reader import H5Min, read_xmf

PHYSICS = "flux_surfgrowth=0.0001\nu_sg=5.55556e-08\ndfe=1.78\nkfe=1.3\nN []=20\n"


def _writer(prefix, grid, per_file, n_width):
    L = mcac_b200.lib()
    w = C.c_void_p()
    assert L.mcac_io_writer_create(str(prefix).encode(), grid.encode(), per_file, n_width, PHYSICS.encode(), C.byref(w)) == 0
    return L, w


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def test_file_names_rotation_and_every_hdf5_structure(tmp_path):
    """7 steps, 3 per file -> _000000, _000001, _000002 (width = ceil(log10(N)) + 4, format.cpp:60-65; the last file is flushed by the
    destructor like ~ThreadedIO); datasets Data0, Data1, ... with f64 / i32 / i64 payloads read back bit for bit."""
    L, w = _writer(tmp_path / "Spheres", "Spheres", 3, 20)
    rng = np.random.default_rng(0)
    want = []
    for step in range(7):
        n = 5 + 3 * step
        t = 1e-9 * step
        xyz = rng.random(3 * n)
        q = rng.integers(-3, 3, n).astype(np.int32)
        r = rng.random(n)
        lab = rng.integers(0, 10**12, n).astype(np.int64)
        tt = np.array([t])
        assert L.mcac_io_begin_step(w, t) == 0
        assert L.mcac_io_attribute(w, b"Time", 0, _ptr(tt), 1, 0) == 0
        assert L.mcac_io_positions(w, _ptr(xyz), n) == 0
        assert L.mcac_io_attribute(w, b"electric_charge", 1, _ptr(q), n, 1) == 0
        assert L.mcac_io_attribute(w, b"Radius", 0, _ptr(r), n, 1) == 0
        assert L.mcac_io_attribute(w, b"Label", 2, _ptr(lab), n, 1) == 0
        assert L.mcac_io_end_step(w) == 0
        want.append((t, dict(Positions=xyz, electric_charge=q, Radius=r, Label=lab, Time=tt)))
    assert L.mcac_io_writer_destroy(w) == 0
    files = sorted(p.name for p in tmp_path.iterdir())
    assert files == [f"Spheres_{k:06d}.{ext}" for k in range(3) for ext in ("h5", "xmf")]
    got = []
    for k in range(3):
        meta, steps = read_xmf(tmp_path / f"Spheres_{k:06d}.xmf")
        assert meta == {"flux_surfgrowth": 0.0001, "u_sg": 5.55556e-08, "dfe": 1.78, "kfe": 1.3, "N []": 20.0}
        h5 = H5Min(tmp_path / f"Spheres_{k:06d}.h5")
        assert sorted(h5.datasets, key=lambda s: int(s[4:])) == [f"Data{i}" for i in range(5 * len(steps))]
        for t, items in steps:
            rec = {}
            for name, (fname, ds, dims) in items.items():
                assert fname == f"Spheres_{k:06d}.h5"
                rec[name] = h5.dataset(ds)
                assert len(rec[name]) == dims
            got.append((t, rec))
    assert len(got) == 7
    for (t0, a), (t1, b) in zip(want, got):
        assert t0 == t1 and set(a) == set(b)
        for name in a:
            assert a[name].dtype == b[name].dtype, name
            np.testing.assert_array_equal(a[name], b[name], err_msg=name)


def test_many_datasets_spill_into_several_symbol_table_nodes(tmp_path):
    """n_time_per_file = 5000 (the validation .ini files) puts tens of thousands of datasets in one group: names must stay sorted
    across symbol table nodes and B-tree keys (Data10 < Data100 < Data2 ... in strcmp order)."""
    L, w = _writer(tmp_path / "Aggregats", "Aggregats", 100000, 800)
    one = np.array([1.5])
    n_steps = 35000
    for step in range(n_steps):
        assert L.mcac_io_begin_step(w, float(step)) == 0
        assert L.mcac_io_positions(w, _ptr(np.array([0., 1., 2.])), 1) == 0
        assert L.mcac_io_attribute(w, b"Rg", 0, _ptr(one), 1, 1) == 0
        assert L.mcac_io_end_step(w) == 0
    assert L.mcac_io_writer_destroy(w) == 0
    h5 = H5Min(tmp_path / "Aggregats_0000000.h5")
    assert len(h5.datasets) == 2 * n_steps and 2 * h5.leaf_k < 2 * n_steps  # more than one node
    np.testing.assert_array_equal(h5.dataset("Data69998"), [0., 1., 2.])
    np.testing.assert_array_equal(h5.dataset("Data69999"), [1.5])


@pytest.mark.gpu
def test_saved_state_is_the_device_state(tmp_path):
    """SphereList::save + AggregatList::save through mcac_gpu_save: every dataset of the grids equals the downloaded state (sphere
    order = creation order, aggregate order = label order), attribute names and integer types as the reference writes them
    (io/sphere_list.cpp:36-57, io/aggregat_list.cpp:36-67)."""
    from golden_lib import Golden
    from mcac_b200.configs import merged_config
    g = Golden("pytest_seed42")
    sim = mcac_b200.Simulation(mcac_b200.ini_text(merged_config(g.base, g.overrides)))
    L, ws = _writer(tmp_path / "Spheres", "Spheres", 10, 20)
    _, wa = _writer(tmp_path / "Aggregats", "Aggregats", 10, 20)
    states = []
    for _ in range(3):
        sim.run(700)
        assert L.mcac_gpu_save(sim.h, ws, wa) == 0
        states.append(sim.state())
    assert L.mcac_io_writer_destroy(ws) == 0 and L.mcac_io_writer_destroy(wa) == 0
    _, s_steps = read_xmf(tmp_path / "Spheres_000000.xmf")
    _, a_steps = read_xmf(tmp_path / "Aggregats_000000.xmf")
    hs, ha = H5Min(tmp_path / "Spheres_000000.h5"), H5Min(tmp_path / "Aggregats_000000.h5")
    assert len(s_steps) == len(a_steps) == 3
    for st, (ts, si), (ta, ai) in zip(states, s_steps, a_steps):
        assert ts == ta == st["time"]
        assert set(si) == {"Positions", "Time", "electric_charge", "Radius", "Label"}
        assert set(ai) == {"Positions", "Time", "Rg", "Np", "f_agg", "lpm", "Deltat", "Rmax", "Volume", "Surface", "proper_time",
                           "coordination_number", "overlapping", "electric_charge", "d_m", "Label"}
        sp, ag = st["spheres"], st["aggregates"]
        np.testing.assert_array_equal(hs.dataset(si["Positions"][1]), np.stack([sp["x"], sp["y"], sp["z"]], axis=1).ravel())
        np.testing.assert_array_equal(hs.dataset(si["Radius"][1]), sp["r"])
        lab = hs.dataset(si["Label"][1])
        assert lab.dtype == np.int64
        np.testing.assert_array_equal(lab, st["sphere_label"])
        assert hs.dataset(si["electric_charge"][1]).dtype == np.int32
        np.testing.assert_array_equal(ha.dataset(ai["Positions"][1]), np.stack([ag["x"], ag["y"], ag["z"]], axis=1).ravel())
        for name, key in [("Rg", "rg"), ("f_agg", "f_agg"), ("lpm", "lpm"), ("Deltat", "time_step"), ("Rmax", "rmax"), ("Volume", "volume"),
                          ("Surface", "surface"), ("proper_time", "proper_time"), ("coordination_number", "coordination_number"),
                          ("overlapping", "overlapping"), ("d_m", "d_m")]:
            np.testing.assert_array_equal(ha.dataset(ai[name][1]), ag[key], err_msg=name)
        np.testing.assert_array_equal(ha.dataset(ai["Np"][1]), st["agg_n_spheres"])
        np.testing.assert_array_equal(ha.dataset(ai["Label"][1]), np.arange(st["n_agg"]))
        # pymcac/tests/test_data.py:166-205: bincount(sphere Label) == aggregate Np
        np.testing.assert_array_equal(np.bincount(lab, minlength=st["n_agg"]), ha.dataset(ai["Np"][1]))
