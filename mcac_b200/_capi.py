"""ctypes binding of libmcac_b200.so (the C ABI declared in include/mcac_b200.h).

Fails loudly when the shared library is missing: there is no Python / CPU fallback for the hot path.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libmcac_b200.so"

SPHERE_FIELDS = ["x", "y", "z", "r", "volume", "surface", "rx", "ry", "rz"]
AGG_FIELDS = ["rg", "f_agg", "lpm", "time_step", "rmax", "volume", "surface", "x", "y", "z", "rx", "ry", "rz", "proper_time", "dp",
              "dg_over_dp", "overlapping", "coordination_number", "electric_charge_field", "d_m", "CH_ratio"]
SCALARS = ["time", "box_length", "maxradius", "max_time_step", "avg_npp", "volume_fraction", "aggregate_concentration",
           "monomer_concentration", "total_volume_concent", "total_surface_concent", "u_sg", "gaz_mean_free_path",
           "mean_massic_radius", "friction_exponnant", "viscosity", "box_volume", "n_iter_without_event", "n_monomeres",
           "temperature", "nucleation_accum"]
ERROR_NAMES = ["NO_ERROR", "UNKNOWN_ERROR", "IO_ERROR", "VERLET_ERROR", "INPUT_ERROR", "ABANDON_ERROR", "TOO_DENSE_ERROR",
               "SBL_ERROR", "VOL_SURF_ERROR", "MERGE_ERROR", "ARVO_ERROR", "InterPotential_ERROR"]


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ["box_length", "time", "temperature", "pressure", "viscosity", "gaz_mean_free_path", "density", "fractal_dimension",
                 "u_sg", "rp_min_oxid", "flux_nucleation", "nucleation_accum", "box_volume", "mean_diameter_nucleation",
                 "dispersion_diameter_nucleation", "physical_time_limit"]] + \
               [(n, C.c_int64) for n in
                ["number_of_aggregates_limit", "n_iter_without_event_limit", "mean_monomere_per_aggregate_limit", "n_monomeres",
                 "full_aggregate_update_frequency"]] + \
               [(n, C.c_int32) for n in
                ["n_verlet_divisions", "pick_method", "volsurf_method", "with_collisions", "with_surface_reactions",
                 "individual_surf_reactions", "with_domain_duplication", "with_maturity", "with_potentials",
                 "with_external_potentials", "with_nucleation", "with_dynamic_random_charges", "normal_initialisation",
                 "sort_order"]] + \
               [("random_seed", C.c_uint32)]


class Contact(C.Structure):
    _fields_ = [("distance", C.c_double), ("moving_sphere", C.c_int64), ("other_sphere", C.c_int64), ("moving_label", C.c_int64),
                ("other_label", C.c_int64)]


CONTACT_DTYPE = np.dtype([("distance", "<f8"), ("moving_sphere", "<i8"), ("other_sphere", "<i8"), ("moving_label", "<i8"),
                          ("other_label", "<i8")])
STEP_DTYPE = np.dtype([
    ("step", "<i8"), ("rand_calls", "<i8"), ("source", "<i8"), ("dir", "<f8", 3), ("full_distance", "<f8"), ("distance", "<f8"),
    ("moving_sphere", "<i8"), ("other_sphere", "<i8"), ("moving_label", "<i8"), ("other_label", "<i8"), ("n_agg", "<i8"),
    ("time", "<f8"), ("dt", "<f8"), ("proper_time", "<f8"), ("pos", "<f8", 3), ("merged", "<i8"), ("n_try", "<i8"),
])


class RunReport(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ["steps", "events", "searches", "pair_tests_sphere", "pair_tests_bounding", "batches", "conflicts", "duplications",
                 "sorts", "kernel_launches", "n_aggregates", "n_spheres", "finished"]] + \
               [(n, C.c_double) for n in ["time", "box_length", "avg_npp", "max_time_step", "volume_fraction", "device_ms",
                                          "search_ms", "commit_ms"]] + \
               [(n, C.c_int64) for n in ["search_launches", "commit_launches"]] + \
               [(n, C.c_double) for n in ["event_ms", "cells_ms"]] + \
               [(n, C.c_int64) for n in ["event_launches", "cells_launches", "sort_span_elements", "sort_levels"]] + \
               [("event_phase_cycles", C.c_int64 * 8)] + \
               [(n, C.c_int64) for n in ["n_iter_without_event", "nucleated"]] + \
               [(n, C.c_double) for n in ["total_volume", "total_surface"]] + \
               [("tie_phase_cycles", C.c_int64 * 2)] + \
               [(n, C.c_int64) for n in ["tie_sorts", "tie_levels", "tie_sparse", "tie_handed"]] + \
               [("tie_sim_cycles", C.c_int64 * 3)] + \
               [(n, C.c_int64) for n in ["sort_fallbacks", "sort_heap_branches", "pair_tests_executed"]] + \
               [("loop_phase_cycles", C.c_int64 * 8)]

    def as_dict(self) -> dict:
        return {n: (list(getattr(self, n)) if n.endswith("_cycles") else getattr(self, n)) for n, _ in self._fields_}


class SweepReport(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ["n_queries", "contacts", "pair_tests_sphere", "pair_tests_bounding"]] + \
               [("distance_checksum", C.c_double), ("kernel_ms", C.c_double)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class McacError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERROR_NAMES[code] if 0 <= code < len(ERROR_NAMES) else code}: {msg}")
        self.code = code


_lib = None
EXPORTS = [
    "mcac_gpu_create", "mcac_gpu_destroy", "mcac_gpu_last_error", "mcac_gpu_set_rng", "mcac_gpu_upload_state", "mcac_gpu_sizes",
    "mcac_gpu_download_state", "mcac_gpu_contact_search", "mcac_gpu_contact_search_batch", "mcac_gpu_translate", "mcac_gpu_merge",
    "mcac_gpu_grow", "mcac_gpu_update", "mcac_gpu_refresh", "mcac_gpu_sort_time_steps", "mcac_gpu_get_pick_table",
    "mcac_gpu_pick_random", "mcac_gpu_pick_last", "mcac_gpu_duplicate", "mcac_gpu_rand", "mcac_gpu_run",
    "mcac_gpu_morphology_stats", "mcac_gpu_morphology_stats_device", "mcac_gpu_stream", "mcac_gpu_search_sweep",
    "mcac_gpu_set_profile", "mcac_gpu_set_interpotential", "mcac_host_alloc_pinned", "mcac_host_free_pinned", "mcac_gpu_kernel_bench", "mcac_ensemble_run", "mcac_gpu_set_stop_at_event",
    "mcac_host_last_error", "mcac_host_model_create", "mcac_host_model_destroy", "mcac_host_model_params", "mcac_host_model_sizes",
    "mcac_host_model_metadata", "mcac_host_model_derived", "mcac_host_model_ini_echo", "mcac_host_model_state", "mcac_sim_create",
    "mcac_io_writer_create", "mcac_io_begin_step", "mcac_io_positions", "mcac_io_attribute", "mcac_io_end_step", "mcac_io_writer_destroy",
    "mcac_gpu_save", "mcac_gpu_set_strict_direction", "mcac_gpu_reserve", "mcac_gpu_aggregate_fields",
]


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(mcac_b200 has no CPU fallback)")
        L = C.CDLL(str(LIB_PATH))
        vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
        L.mcac_gpu_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(vp)]
        L.mcac_gpu_destroy.argtypes = [vp]
        L.mcac_gpu_last_error.argtypes = [vp]
        L.mcac_gpu_last_error.restype = C.c_char_p
        L.mcac_gpu_set_rng.argtypes = [vp, C.c_uint32, i64]
        L.mcac_gpu_upload_state.argtypes = [vp, i64, i64] + [vp] * 8 + [dbl, dbl]
        L.mcac_gpu_sizes.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
        L.mcac_gpu_download_state.argtypes = [vp] + [vp] * 11
        L.mcac_gpu_contact_search.argtypes = [vp, i64, vp, dbl, C.POINTER(Contact)]
        L.mcac_gpu_contact_search_batch.argtypes = [vp, i64, vp, vp, vp, vp, vp]
        L.mcac_gpu_translate.argtypes = [vp, i64, vp]
        L.mcac_gpu_merge.argtypes = [vp, C.POINTER(Contact), C.POINTER(C.c_int)]
        L.mcac_gpu_grow.argtypes = [vp, dbl, i64]
        L.mcac_gpu_update.argtypes = [vp, i64, C.c_int]
        L.mcac_gpu_refresh.argtypes = [vp] + [C.POINTER(dbl)] * 4
        L.mcac_gpu_sort_time_steps.argtypes = [vp, dbl]
        L.mcac_gpu_get_pick_table.argtypes = [vp, vp, vp, C.POINTER(i64)]
        L.mcac_gpu_pick_random.argtypes = [vp, dbl, C.POINTER(i64), C.POINTER(dbl)]
        L.mcac_gpu_pick_last.argtypes = [vp, C.POINTER(i64)]
        L.mcac_gpu_duplicate.argtypes = [vp]
        L.mcac_gpu_rand.argtypes = [vp, i64, vp]
        L.mcac_gpu_run.argtypes = [vp, i64, C.c_int32, vp, i64, C.POINTER(RunReport)]
        L.mcac_gpu_morphology_stats.argtypes = [vp, C.c_int32, dbl, vp]
        L.mcac_gpu_morphology_stats_device.argtypes = [vp, C.c_int32, dbl, vp]
        L.mcac_gpu_search_sweep.argtypes = [vp, i64, C.c_int32, C.POINTER(SweepReport)]
        L.mcac_gpu_set_profile.argtypes = [vp, C.c_int32]
        L.mcac_gpu_set_stop_at_event.argtypes = [vp, C.c_int32]
        L.mcac_gpu_set_interpotential.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp, vp]
        L.mcac_gpu_stream.argtypes = [vp]
        L.mcac_gpu_kernel_bench.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(dbl), C.POINTER(i64)]
        L.mcac_ensemble_run.argtypes = [vp, C.c_int32, i64, C.c_int32, C.c_int32, vp]
        L.mcac_host_alloc_pinned.argtypes = [i64, C.POINTER(vp)]
        L.mcac_host_free_pinned.argtypes = [vp]
        L.mcac_gpu_stream.restype = vp
        L.mcac_host_last_error.restype = C.c_char_p
        L.mcac_host_model_create.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
        L.mcac_host_model_destroy.argtypes = [vp]
        L.mcac_host_model_params.argtypes = [vp, C.POINTER(Params)]
        L.mcac_host_model_sizes.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
        L.mcac_host_model_metadata.argtypes = [vp, C.c_char_p, i64]
        L.mcac_host_model_derived.argtypes = [vp, vp]
        L.mcac_host_model_ini_echo.argtypes = [vp, C.c_char_p, C.c_int64]
        L.mcac_host_model_state.argtypes = [vp] + [vp] * 8
        L.mcac_sim_create.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
        L.mcac_io_writer_create.argtypes = [C.c_char_p, C.c_char_p, i64, i64, C.c_char_p, C.POINTER(vp)]
        L.mcac_io_begin_step.argtypes = [vp, C.c_double]
        L.mcac_io_positions.argtypes = [vp, vp, i64]
        L.mcac_io_attribute.argtypes = [vp, C.c_char_p, C.c_int32, vp, i64, C.c_int32]
        L.mcac_io_end_step.argtypes = [vp]
        L.mcac_io_writer_destroy.argtypes = [vp]
        L.mcac_gpu_save.argtypes = [vp, vp, vp]
        L.mcac_gpu_set_strict_direction.argtypes = [vp, C.c_int32]
        L.mcac_gpu_reserve.argtypes = [vp, i64, i64]
        L.mcac_gpu_aggregate_fields.argtypes = [vp, i64, vp, C.POINTER(i64)]
        _lib = L
    return _lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
