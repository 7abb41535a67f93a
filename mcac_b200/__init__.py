"""mcac_b200 — B200-native (sm_100a) Monte-Carlo aggregation hot path of MCAC behind a C ABI.

Python here is only the harness side of the boundary (tests, bench.py): `HostModel` mirrors
PhysicalModel(ini) + the initial placement, `Simulation` mirrors AggregatList + calcul() on the device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import (AGG_FIELDS, CONTACT_DTYPE, SCALARS, SPHERE_FIELDS, STEP_DTYPE, Contact, McacError, Params, RunReport, SweepReport,
                    lib, ptr)

__all__ = ["HostModel", "Simulation", "Ensemble", "McacError", "ini_text", "Params"]


def ini_text(cfg: dict) -> str:
    out = []
    for sec, kv in cfg.items():
        out.append(f"[{sec}]")
        out += [f"{k}={v}" for k, v in kv.items()]
        out.append("")
    return "\n".join(out)


class HostModel:
    """PhysicalModel(ini) (+ AggregatList placement when place=True) on the host — no GPU needed."""

    def __init__(self, text: str, place: bool = True):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.mcac_host_model_create(text.encode(), int(place), C.byref(h))
        if rc:
            raise McacError(rc, self.L.mcac_host_last_error().decode())
        self.h = h
        self.placed = place

    def __del__(self):
        if getattr(self, "h", None):
            self.L.mcac_host_model_destroy(self.h)
            self.h = None

    def params(self) -> Params:
        p = Params()
        self.L.mcac_host_model_params(self.h, C.byref(p))
        return p

    def metadata(self) -> dict:
        buf = C.create_string_buffer(4096)
        self.L.mcac_host_model_metadata(self.h, buf, 4096)
        return dict(line.split("=", 1) for line in buf.value.decode().splitlines())

    def ini_echo(self) -> str:
        buf = C.create_string_buffer(1 << 16)
        rc = self.L.mcac_host_model_ini_echo(self.h, buf, 1 << 16)
        if rc:
            raise McacError(rc, "ini echo does not fit")
        return buf.value.decode()

    def derived(self) -> dict:
        a = np.zeros(12)
        self.L.mcac_host_model_derived(self.h, ptr(a))
        names = ["box_length", "box_volume", "viscosity", "gaz_mean_free_path", "mean_massic_radius", "friction_exponnant", "u_sg",
                 "aggregate_concentration", "total_volume_concent", "total_surface_concent", "mass_nuclei", "volume_fraction"]
        return dict(zip(names, (float(v) for v in a)))

    def state(self) -> dict:
        ns, na = C.c_int64(), C.c_int64()
        self.L.mcac_host_model_sizes(self.h, C.byref(ns), C.byref(na))
        ns, na = ns.value, na.value
        sf = np.zeros((9, ns)); af = np.zeros((21, na)); cells = np.zeros((3, na), np.int64)
        offs = np.zeros(na + 1, np.int64); mem = np.zeros(ns, np.int64); pm = np.zeros((3, ns)); sc = np.zeros(3)
        rc = C.c_int64()
        self.L.mcac_host_model_state(self.h, ptr(sf), ptr(af), ptr(cells), ptr(offs), ptr(mem), ptr(pm), ptr(sc), C.byref(rc))
        return dict(n_sph=ns, n_agg=na, spheres=dict(zip(SPHERE_FIELDS, sf)), aggregates=dict(zip(AGG_FIELDS, af)), agg_cell=cells,
                    offsets=offs, members=mem, member_volumes=pm[0], member_surfaces=pm[1], member_distances_center=pm[2],
                    maxradius=float(sc[0]), max_time_step=float(sc[1]), avg_npp=float(sc[2]), rand_consumed=rc.value,
                    sphere_fields=sf, agg_fields=af, per_member=pm)


class Simulation:
    """One realization resident in HBM (AggregatList + the calcul() loop of the reference, on the device)."""

    def __init__(self, text: str | None = None, device: int = 0, *, params: Params | None = None):
        self.L = lib()
        h = C.c_void_p()
        if text is not None:
            rc = self.L.mcac_sim_create(text.encode(), device, C.byref(h))
            if rc:
                raise McacError(rc, self.L.mcac_host_last_error().decode())
        else:
            rc = self.L.mcac_gpu_create(C.byref(params), device, C.byref(h))
            if rc:
                msg = self.L.mcac_gpu_last_error(h).decode() if h else "create failed"
                if h:
                    self.L.mcac_gpu_destroy(h)
                raise McacError(rc, msg)
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.mcac_gpu_destroy(self.h)
            self.h = None

    def _ck(self, rc: int):
        if rc:
            raise McacError(rc, self.L.mcac_gpu_last_error(self.h).decode())

    # ---- state
    def set_rng(self, seed: int, consumed: int):
        self._ck(self.L.mcac_gpu_set_rng(self.h, seed, consumed))

    def upload(self, st: dict):
        sf = np.ascontiguousarray(st["sphere_fields"], np.float64); af = np.ascontiguousarray(st["agg_fields"], np.float64)
        ns, na = sf.shape[1], af.shape[1]
        sch = np.ascontiguousarray(st.get("sphere_charge", np.zeros(ns)), np.int64)
        ach = np.ascontiguousarray(st.get("agg_charge", np.zeros(na)), np.int64)
        cells = np.ascontiguousarray(st["agg_cell"], np.int64); offs = np.ascontiguousarray(st["offsets"], np.int64)
        mem = np.ascontiguousarray(st["members"], np.int64); pm = np.ascontiguousarray(st["per_member"], np.float64)
        self._ck(self.L.mcac_gpu_upload_state(self.h, ns, na, ptr(sf), ptr(sch), ptr(af), ptr(ach), ptr(cells), ptr(offs), ptr(mem),
                                              ptr(pm), float(st["maxradius"]), float(st["max_time_step"])))

    def sizes(self):
        ns, na = C.c_int64(), C.c_int64()
        self._ck(self.L.mcac_gpu_sizes(self.h, C.byref(ns), C.byref(na)))
        return ns.value, na.value

    def state(self) -> dict:
        ns, na = self.sizes()
        sf = np.zeros((9, ns)); lab = np.zeros(ns, np.int64); sch = np.zeros(ns, np.int64); af = np.zeros((21, na))
        nsp = np.zeros(na, np.int64); ach = np.zeros(na, np.int64); cells = np.zeros((3, na), np.int64)
        offs = np.zeros(na + 1, np.int64); mem = np.zeros(ns, np.int64); pm = np.zeros((3, ns)); sc = np.zeros(20)
        self._ck(self.L.mcac_gpu_download_state(self.h, ptr(sf), ptr(lab), ptr(sch), ptr(af), ptr(nsp), ptr(ach), ptr(cells), ptr(offs),
                                                ptr(mem), ptr(pm), ptr(sc)))
        out = dict(n_sph=ns, n_agg=na, spheres=dict(zip(SPHERE_FIELDS, sf)), sphere_label=lab, sphere_charge=sch,
                   aggregates=dict(zip(AGG_FIELDS, af)), agg_n_spheres=nsp, agg_charge=ach, agg_cell=cells, offsets=offs, members=mem,
                   member_volumes=pm[0], member_surfaces=pm[1], member_distances_center=pm[2])
        out.update(dict(zip(SCALARS, (float(v) for v in sc))))
        return out

    # ---- the same boundary on caller-owned page-locked arrays (no per-call allocation; full-bandwidth async copies)
    def pinned_state_buffers(self, n_sph: int | None = None, n_agg: int | None = None) -> dict:
        """Host arrays of the state boundary in page-locked memory, sized for (n_sph, n_agg) (default: the current state)."""
        if n_sph is None or n_agg is None:
            n_sph, n_agg = self.sizes()
        spec = dict(sphere_fields=((9, n_sph), np.float64), sphere_label=((n_sph,), np.int64), sphere_charge=((n_sph,), np.int64),
                    agg_fields=((21, n_agg), np.float64), agg_n_spheres=((n_agg,), np.int64), agg_charge=((n_agg,), np.int64),
                    agg_cell=((3, n_agg), np.int64), offsets=((n_agg + 1,), np.int64), members=((n_sph,), np.int64),
                    per_member=((3, n_sph), np.float64), scalars=((20,), np.float64))
        out = {"_pinned": []}
        for name, (shape, dt) in spec.items():
            nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize
            p = C.c_void_p()
            if self.L.mcac_host_alloc_pinned(nbytes, C.byref(p)):
                raise McacError(1, "cudaHostAlloc failed")
            out["_pinned"].append(p)
            buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
            out[name] = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
        return out

    def free_pinned(self, bufs: dict):
        for p in bufs.pop("_pinned", []):
            self.L.mcac_host_free_pinned(p)

    def download_into(self, b: dict, n_sph: int | None = None, n_agg: int | None = None):
        """mcac_gpu_download_state into arrays of matching size (see pinned_state_buffers)."""
        self._ck(self.L.mcac_gpu_download_state(self.h, ptr(b["sphere_fields"]), ptr(b["sphere_label"]), ptr(b["sphere_charge"]),
                                                ptr(b["agg_fields"]), ptr(b["agg_n_spheres"]), ptr(b["agg_charge"]), ptr(b["agg_cell"]),
                                                ptr(b["offsets"]), ptr(b["members"]), ptr(b["per_member"]), ptr(b["scalars"])))

    def upload_from(self, b: dict):
        ns, na = b["sphere_fields"].shape[1], b["agg_fields"].shape[1]
        self._ck(self.L.mcac_gpu_upload_state(self.h, ns, na, ptr(b["sphere_fields"]), ptr(b["sphere_charge"]), ptr(b["agg_fields"]),
                                              ptr(b["agg_charge"]), ptr(b["agg_cell"]), ptr(b["offsets"]), ptr(b["members"]),
                                              ptr(b["per_member"]), float(b["scalars"][2]), float(b["scalars"][3])))

    # ---- per-call mirror of the AggregatList / Aggregate methods
    def contact_search(self, label: int, direction, distance: float) -> Contact:
        d = np.ascontiguousarray(direction, np.float64)
        c = Contact()
        self._ck(self.L.mcac_gpu_contact_search(self.h, label, ptr(d), distance, C.byref(c)))
        return c

    def contact_search_batch(self, labels, directions, distances):
        labels = np.ascontiguousarray(labels, np.int64); n = len(labels)
        directions = np.ascontiguousarray(directions, np.float64).reshape(n, 3)
        distances = np.ascontiguousarray(distances, np.float64)
        out = np.zeros(n, CONTACT_DTYPE); pairs = np.zeros(2, np.int64)
        self._ck(self.L.mcac_gpu_contact_search_batch(self.h, n, ptr(labels), ptr(directions), ptr(distances), ptr(out), ptr(pairs)))
        return out, pairs

    def translate(self, label: int, vector):
        v = np.ascontiguousarray(vector, np.float64)
        self._ck(self.L.mcac_gpu_translate(self.h, label, ptr(v)))

    def merge(self, contact: Contact) -> bool:
        m = C.c_int()
        self._ck(self.L.mcac_gpu_merge(self.h, C.byref(contact), C.byref(m)))
        return bool(m.value)

    def grow(self, dt: float, label: int = -1):
        self._ck(self.L.mcac_gpu_grow(self.h, dt, label))

    def update(self, label: int = -1, full: bool = True):
        self._ck(self.L.mcac_gpu_update(self.h, label, int(full)))

    def aggregate_fields(self, label: int):
        """(the 21 AggregatesFields of one aggregate as a dict, n_spheres) — Aggregate::get_lpm / get_time_step / ..."""
        f = np.zeros(21); n = C.c_int64()
        self._ck(self.L.mcac_gpu_aggregate_fields(self.h, label, ptr(f), C.byref(n)))
        return dict(zip(AGG_FIELDS, (float(v) for v in f))), n.value

    def refresh(self):
        a, b, c, d = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        self._ck(self.L.mcac_gpu_refresh(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(max_time_step=a.value, avg_npp=b.value, total_volume=c.value, total_surface=d.value)

    def sort_time_steps(self, factor: float):
        self._ck(self.L.mcac_gpu_sort_time_steps(self.h, factor))

    def pick_table(self):
        _, na = self.sizes()
        idx = np.zeros(na, np.int64); cum = np.zeros(na); n = C.c_int64()
        self._ck(self.L.mcac_gpu_get_pick_table(self.h, ptr(idx), ptr(cum), C.byref(n)))
        return idx[:n.value], cum[:n.value]

    def pick_random(self, u: float):
        lab, dt = C.c_int64(), C.c_double()
        self._ck(self.L.mcac_gpu_pick_random(self.h, u, C.byref(lab), C.byref(dt)))
        return lab.value, dt.value

    def duplicate(self):
        self._ck(self.L.mcac_gpu_duplicate(self.h))

    def rand(self, n: int) -> np.ndarray:
        out = np.zeros(n, np.int32)
        self._ck(self.L.mcac_gpu_rand(self.h, n, ptr(out)))
        return out

    def search_sweep(self, n: int, repeats: int = 1) -> dict:
        rep = SweepReport()
        self._ck(self.L.mcac_gpu_search_sweep(self.h, n, repeats, C.byref(rep)))
        return rep.as_dict()

    KERNELS = {"cells": 0, "grow": 1, "update_partial": 2, "update_full": 3, "event_sort": 4, "event_nosort": 5, "grid_barriers_x100": 6,
               "rng_fill": 7, "morphology_stats": 8, "fp64_dfma": 9, "fp64_dmul_dadd": 10, "plan_probe": 11}

    def kernel_bench(self, which: str, reps: int = 5) -> dict:
        ms, units = C.c_double(), C.c_int64()
        self._ck(self.L.mcac_gpu_kernel_bench(self.h, self.KERNELS[which], reps, C.byref(ms), C.byref(units)))
        return {"kernel": which, "ms": ms.value, "units": units.value}

    def set_stop_at_event(self, on: bool = True):
        self._ck(self.L.mcac_gpu_set_stop_at_event(self.h, int(on)))

    def set_strict_direction(self, on: bool = True):
        """Directions from the host's glibc (bit-exact replay of the reference) instead of CUDA's sincos / acos."""
        self._ck(self.L.mcac_gpu_set_strict_direction(self.h, int(on)))

    def set_profile(self, on: bool = True):
        self._ck(self.L.mcac_gpu_set_profile(self.h, int(on)))

    def morphology_stats(self, n_bins: int = 24, rg_max: float = 1e-5) -> np.ndarray:
        out = np.zeros(2 * n_bins + 8)
        self._ck(self.L.mcac_gpu_morphology_stats(self.h, n_bins, rg_max, ptr(out)))
        return out

    def morphology_stats_device(self, device_ptr: int, n_bins: int = 24, rg_max: float = 1e-5):
        self._ck(self.L.mcac_gpu_morphology_stats_device(self.h, n_bins, rg_max, C.c_void_p(device_ptr)))

    # ---- the whole loop
    def run(self, max_steps: int, batch: int = 0, records: int = 0):
        recs = np.zeros(records, STEP_DTYPE) if records else None
        rep = RunReport()
        self._ck(self.L.mcac_gpu_run(self.h, max_steps, batch, ptr(recs), records, C.byref(rep)))
        if recs is not None:
            recs = recs[:min(records, rep.steps)]
        return rep.as_dict(), recs


class Ensemble:
    """Independent realizations of one configuration (distinct seeds) resident on one GPU and advanced concurrently."""

    def __init__(self, texts: list[str], device: int = 0, threads: int = 16):
        # creation = host placement + (classic.ini) parsing the interaction-potential table: the C calls release the GIL
        from concurrent.futures import ThreadPoolExecutor
        if len(texts) > 8 and threads > 1:
            with ThreadPoolExecutor(threads) as ex:
                self.sims = list(ex.map(lambda t: Simulation(t, device=device), texts))
        else:
            self.sims = [Simulation(t, device=device) for t in texts]
        self.L = lib()

    def __len__(self):
        return len(self.sims)

    def run(self, max_steps: int, batch: int = 0, threads: int = 8) -> list[dict]:
        n = len(self.sims)
        hs = (C.c_void_p * n)(*[s.h for s in self.sims])
        reps = (RunReport * n)()
        rc = self.L.mcac_ensemble_run(hs, n, max_steps, batch, threads, reps)
        if rc:
            msgs = [self.L.mcac_gpu_last_error(s.h).decode() for s in self.sims]
            distinct = list(dict.fromkeys(m for m in msgs if m))
            raise McacError(rc, " | ".join(distinct)[:1500])
        return [r.as_dict() for r in reps]

    def morphology_stats(self, n_bins: int = 24, rg_max: float = 1e-5) -> np.ndarray:
        """(n_realizations, 2*n_bins + 8) statistics rows (K11), ready for the all-gather across ranks"""
        return np.stack([s.morphology_stats(n_bins, rg_max) for s in self.sims])
