"""Readers for the raw little-endian records written by oracle/ref_build/tap.cpp (reference taps)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

SEARCH_DTYPE = np.dtype([
    ("step", "<i8"), ("rand_calls", "<i8"), ("source", "<i8"), ("dir", "<f8", 3), ("full_distance", "<f8"),
    ("distance", "<f8"), ("moving_sphere", "<i8"), ("other_sphere", "<i8"), ("moving_label", "<i8"),
    ("other_label", "<i8"), ("n_agg", "<i8"), ("time", "<f8"),
])
STEP_DTYPE = np.dtype([
    ("step", "<i8"), ("rand_calls", "<i8"), ("label", "<i8"), ("dt", "<f8"), ("proper_time", "<f8"),
    ("pos", "<f8", 3), ("lpm", "<f8"),
])
MERGE_DTYPE = np.dtype([("step", "<i8"), ("ok", "<i8"), ("n_agg", "<i8"), ("n_sph", "<i8")])

SPHERE_FIELDS = ["x", "y", "z", "r", "volume", "surface", "rx", "ry", "rz"]
AGG_FIELDS = ["rg", "f_agg", "lpm", "time_step", "rmax", "volume", "surface", "x", "y", "z", "rx", "ry", "rz",
              "proper_time", "dp", "dg_over_dp", "overlapping", "coordination_number", "electric_charge_field", "d_m",
              "CH_ratio"]
SCALARS = ["time", "box_length", "maxradius", "max_time_step", "avg_npp", "volume_fraction", "aggregate_concentration",
           "monomer_concentration", "total_volume_concent", "total_surface_concent"]


def read_searches(path) -> np.ndarray:
    return np.fromfile(path, dtype=SEARCH_DTYPE)


def read_steps(path) -> np.ndarray:
    return np.fromfile(path, dtype=STEP_DTYPE)


def read_merges(path) -> np.ndarray:
    return np.fromfile(path, dtype=MERGE_DTYPE)


def read_sort(path) -> dict:
    raw = Path(path).read_bytes()
    step, n = np.frombuffer(raw, "<i8", 2, 0)
    factor = np.frombuffer(raw, "<f8", 1, 16)[0]
    off = 24
    idx = np.frombuffer(raw, "<i8", n, off); off += 8 * n
    cum = np.frombuffer(raw, "<f8", n, off); off += 8 * n
    ts = np.frombuffer(raw, "<f8", n, off)
    return dict(step=int(step), n=int(n), factor=float(factor), idx=idx.copy(), cum=cum.copy(), time_step=ts.copy())


def read_state(path) -> dict:
    """Full SoA snapshot (layout = dump_state() in oracle/ref_build/tap.cpp)."""
    raw = Path(path).read_bytes()
    hdr_i = np.frombuffer(raw, "<i8", 7, 0)
    assert hdr_i[0] == 0x4D434143534E4150, "bad magic"
    off = 56
    sc = np.frombuffer(raw, "<f8", 10, off); off += 80
    n_sph, n_agg = int(hdr_i[3]), int(hdr_i[4])
    out = dict(step=int(hdr_i[1]), rand_calls=int(hdr_i[2]), n_sph=n_sph, n_agg=n_agg, n_monomeres=int(hdr_i[5]),
               n_iter_without_event=int(hdr_i[6]))
    out.update({k: float(v) for k, v in zip(SCALARS, sc)})

    def take(dt, count):
        nonlocal off
        a = np.frombuffer(raw, dt, count, off).copy()
        off += a.nbytes
        return a

    out["spheres"] = {k: take("<f8", n_sph) for k in SPHERE_FIELDS}
    out["sphere_label"] = take("<i8", n_sph)
    out["sphere_charge"] = take("<i8", n_sph)
    out["aggregates"] = {k: take("<f8", n_agg) for k in AGG_FIELDS}
    out["agg_n_spheres"] = take("<i8", n_agg)
    out["agg_label"] = take("<i8", n_agg)
    out["agg_charge"] = take("<i8", n_agg)
    out["agg_cell"] = np.stack([take("<i8", n_agg) for _ in range(3)])
    out["agg_bulk_density"] = take("<f8", n_agg)
    out["agg_alpha_vs_extreme"] = take("<f8", n_agg)
    out["offsets"] = take("<i8", n_agg + 1)
    out["members"] = take("<i8", n_sph)
    out["member_volumes"] = take("<f8", n_sph)
    out["member_surfaces"] = take("<f8", n_sph)
    out["member_distances_center"] = take("<f8", n_sph)
    assert off == len(raw), (off, len(raw))
    return out
