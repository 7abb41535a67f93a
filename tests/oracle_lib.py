"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (the CPU restatement of the reference)."""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.run_ref import merged_config  # noqa: E402

from ref_trace import AGG_FIELDS, SPHERE_FIELDS  # noqa: E402

STEP_DTYPE = np.dtype([
    ("step", "<i8"), ("rand_calls", "<i8"), ("source", "<i8"), ("dir", "<f8", 3), ("full_distance", "<f8"),
    ("distance", "<f8"), ("moving_sphere", "<i8"), ("other_sphere", "<i8"), ("moving_label", "<i8"),
    ("other_label", "<i8"), ("n_agg", "<i8"), ("time", "<f8"), ("dt", "<f8"), ("proper_time", "<f8"),
    ("pos", "<f8", 3), ("merged", "<i8"), ("n_try", "<i8"),
])
ORC_SCALARS = ["time", "box_length", "maxradius", "max_time_step", "avg_npp", "volume_fraction",
               "aggregate_concentration", "monomer_concentration", "total_volume_concent", "total_surface_concent", "u_sg",
               "gaz_mean_free_path", "mean_massic_radius", "friction_exponnant", "viscosity", "box_volume",
               "n_iter_without_event", "n_monomeres", "temperature", "nucleation_accum"]
COUNTERS = ["steps", "events", "searches", "pair_sphere", "pair_bounding", "sorts", "duplications", "rand_calls"]

_lib = None


def lib():
    global _lib
    if _lib is None:
        so = ROOT / "oracle" / "liboracle.so"
        src = ROOT / "oracle" / "mcac_oracle.cpp"
        if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
            subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(str(so))
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_run.restype = C.c_longlong
        L.orc_run.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong]
        for f in ("orc_n_spheres", "orc_n_aggregates", "orc_pick_table_size"):
            getattr(L, f).restype = C.c_longlong
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_finished.argtypes = [C.c_void_p]
        L.orc_set_stable_sort.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_stop_after_move.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_scalars.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_spheres.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.orc_get_aggregates.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.orc_get_pick_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_pair_distance_batch.argtypes = [C.c_longlong] + [C.c_void_p] * 6 + [C.c_double, C.c_void_p]
        L.orc_search.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_rand_stream.argtypes = [C.c_uint, C.c_longlong, C.c_void_p]
        L.orc_direction.argtypes = [C.c_double, C.c_double, C.c_void_p]
        L.orc_physics.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.orc_sort_introsort.argtypes = [C.c_longlong, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def ini_text(cfg: dict) -> str:
    out = []
    for sec, kv in cfg.items():
        out.append(f"[{sec}]")
        out += [f"{k}={v}" for k, v in kv.items()]
        out.append("")
    return "\n".join(out)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, base: str | None = None, overrides: dict | None = None, *, text: str | None = None,
                 base_dir: str = "", construct: bool = True):
        self.L = lib()
        if text is None:
            text = ini_text(merged_config(base, overrides))
        self.h = self.L.orc_create(text.encode(), base_dir.encode(), int(construct))
        if not self.h:
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def run(self, max_steps: int, record: bool = True) -> np.ndarray:
        recs = np.zeros(max_steps if record else 0, dtype=STEP_DTYPE)
        n = self.L.orc_run(self.h, max_steps, _p(recs) if record else None, len(recs))
        if n < 0:
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())
        self.last_steps = int(n)
        return recs[:n] if record else recs

    def run_partial_step(self) -> np.ndarray:
        """One step executed only up to time_forward (calcul.cpp:149) — where the tap's exit_step snapshot is taken."""
        self.L.orc_set_stop_after_move(self.h, 1)
        try:
            return self.run(1)
        finally:
            self.L.orc_set_stop_after_move(self.h, 0)

    @property
    def finished(self) -> bool:
        return bool(self.L.orc_finished(self.h))

    def counters(self) -> dict:
        a = np.zeros(8, dtype=np.int64)
        self.L.orc_get_counters(self.h, _p(a))
        return dict(zip(COUNTERS, (int(v) for v in a)))

    def scalars(self) -> dict:
        a = np.zeros(20, dtype=np.float64)
        self.L.orc_get_scalars(self.h, _p(a))
        return dict(zip(ORC_SCALARS, (float(v) for v in a)))

    def state(self) -> dict:
        ns, na = int(self.L.orc_n_spheres(self.h)), int(self.L.orc_n_aggregates(self.h))
        sf = np.zeros((9, ns)); lab = np.zeros(ns, np.int64); ch = np.zeros(ns, np.int64)
        self.L.orc_get_spheres(self.h, _p(sf), _p(lab), _p(ch))
        af = np.zeros((21, na)); nsp = np.zeros(na, np.int64); cells = np.zeros((3, na), np.int64)
        ach = np.zeros(na, np.int64); offs = np.zeros(na + 1, np.int64); mem = np.zeros(ns, np.int64)
        pm = np.zeros((3, ns))
        self.L.orc_get_aggregates(self.h, _p(af), _p(nsp), _p(cells), _p(ach), _p(offs), _p(mem), _p(pm))
        out = dict(n_sph=ns, n_agg=na, spheres=dict(zip(SPHERE_FIELDS, sf)), sphere_label=lab, sphere_charge=ch,
                   aggregates=dict(zip(AGG_FIELDS, af)), agg_n_spheres=nsp, agg_cell=cells, agg_charge=ach, offsets=offs,
                   members=mem, member_volumes=pm[0], member_surfaces=pm[1], member_distances_center=pm[2])
        out.update(self.scalars())
        return out

    def set_state(self, st: dict, rand_consumed: int) -> None:
        """Re-synchronise with a state in the layout Simulation.state() / Oracle.state() return (see orc_set_state)."""
        ns, na = int(st["n_sph"]), int(st["n_agg"])
        sf = np.ascontiguousarray(np.stack([st["spheres"][k] for k in SPHERE_FIELDS]), np.float64)
        af = np.ascontiguousarray(np.stack([st["aggregates"][k] for k in AGG_FIELDS]), np.float64)
        lab = np.ascontiguousarray(st["sphere_label"], np.int64)
        cells = np.ascontiguousarray(st["agg_cell"], np.int64)
        offs = np.ascontiguousarray(st["offsets"], np.int64); mem = np.ascontiguousarray(st["members"], np.int64)
        pm = np.ascontiguousarray(np.stack([st["member_volumes"], st["member_surfaces"], st["member_distances_center"]]), np.float64)
        sc = np.array([st["time"], st["maxradius"], st["max_time_step"], st["avg_npp"], st["n_iter_without_event"]], np.float64)
        self.L.orc_set_state.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong] + [C.c_void_p] * 8 + [C.c_longlong]
        if self.L.orc_set_state(self.h, ns, na, _p(sf), _p(lab), _p(af), _p(cells), _p(offs), _p(mem), _p(pm), _p(sc), int(rand_consumed)):
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())

    def update_all(self) -> None:
        self.L.orc_update_all.argtypes = [C.c_void_p]
        if self.L.orc_update_all(self.h):
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())

    def pick_table(self):
        n = int(self.L.orc_pick_table_size(self.h))
        idx = np.zeros(n, np.int64); cum = np.zeros(n)
        self.L.orc_get_pick_table(self.h, _p(idx), _p(cum))
        return idx, cum

    def search(self, source: int, direction, dist: float):
        d = np.ascontiguousarray(direction, dtype=np.float64)
        out = C.c_double(); ids = np.zeros(4, np.int64)
        rc = self.L.orc_search(self.h, source, _p(d), dist, C.byref(out), _p(ids))
        if rc:
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())
        return out.value, ids

    def physics(self, V: float, v: float, r: float) -> np.ndarray:
        out = np.zeros(6)
        self.L.orc_physics(self.h, V, v, r, _p(out))
        return out


def pair_distance(p1, r1, p2, r2, direction, dist, L) -> np.ndarray:
    p1 = np.ascontiguousarray(p1, np.float64).reshape(-1, 3); n = len(p1)
    p2 = np.ascontiguousarray(p2, np.float64).reshape(-1, 3)
    d = np.ascontiguousarray(direction, np.float64).reshape(-1, 3)
    r1 = np.ascontiguousarray(np.broadcast_to(r1, n), np.float64); r2 = np.ascontiguousarray(np.broadcast_to(r2, n), np.float64)
    dist = np.ascontiguousarray(np.broadcast_to(dist, n), np.float64)
    out = np.zeros(n)
    lib().orc_pair_distance_batch(n, _p(p1), _p(r1), _p(p2), _p(r2), _p(d), _p(dist), float(L), _p(out))
    return out


def rand_stream(seed: int, n: int) -> np.ndarray:
    out = np.zeros(n, np.int32)
    lib().orc_rand_stream(seed, n, _p(out))
    return out


def introsort_order(keys: np.ndarray) -> np.ndarray:
    keys = np.ascontiguousarray(keys, np.float64)
    idx = np.zeros(len(keys), np.int64)
    lib().orc_sort_introsort(len(keys), _p(keys), _p(idx))
    return idx


def introsort_order_depth(keys: np.ndarray, depth: int) -> np.ndarray:
    """std::sort's loop with the depth limit forced to `depth` (the heap-sort branch is taken below it)."""
    keys = np.ascontiguousarray(keys, np.float64)
    idx = np.zeros(len(keys), np.int64)
    lib().orc_sort_introsort_depth(len(keys), _p(keys), int(depth), _p(idx))
    return idx


def linreg(x: np.ndarray, y: np.ndarray) -> tuple:
    """mcac::linreg restated in C (oracle/mcac_oracle.cpp:orc_linreg): (ok, a, b, r)."""
    x = np.ascontiguousarray(x, np.float64); y = np.ascontiguousarray(y, np.float64)
    out = np.zeros(4)
    lib().orc_linreg(len(x), _p(x), _p(y), _p(out))
    return bool(out[0]), float(out[1]), float(out[2]), float(out[3])
