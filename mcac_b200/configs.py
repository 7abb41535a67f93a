"""The named workloads of BASELINE.json as dict literals, restated key by key from the reference's validation/*.ini and
examples/classic.ini (SURVEY.md Appendix C lists the keys), so that tests, bench.py and the profiles build the same .ini texts
without reading /root/reference at run time.  Product-side helper: nothing here belongs to the oracle."""
from __future__ import annotations

# validation/params_monodisperse.ini, params_polydisperse.ini, params_brownian.ini,
# params_surface_growth.ini, params_pytest.ini, examples/classic.ini — restated key by key.
_COMMON_DLCA = {
    "environment": dict(initial_time=0, fractal_dimension=1.78, fractal_prefactor=1.30, pressure=101300,
                        temperature=1700, volume_fraction="10e-6"),
}
CONFIGS = {
    "monodisperse": {
        **_COMMON_DLCA,
        "limits": dict(cpu=-1, mean_monomere_per_aggregate=100, n_iter_without_event=-1, number_of_aggregates=1,
                       physical_time=-1),
        "monomers": dict(density=1800, dispersion_diameter="1.00", initialisation_mode="lognormal", mean_diameter=10,
                         number=800),
        "numerics": dict(n_verlet_divisions=3, pick_method="random"),
        "output": dict(n_time_per_file=5000, write_between_event_frequency=100000, output_dir="out"),
    },
    "polydisperse": {
        **_COMMON_DLCA,
        "limits": dict(cpu=-1, mean_monomere_per_aggregate=100, n_iter_without_event=-1, number_of_aggregates=1,
                       physical_time=-1),
        "monomers": dict(density=1800, dispersion_diameter="1.25", initialisation_mode="lognormal", mean_diameter=20,
                         number=800),
        "numerics": dict(n_verlet_divisions=3, pick_method="random"),
        "output": dict(n_time_per_file=5000, write_between_event_frequency=100000, output_dir="out"),
    },
    "brownian": {
        "environment": dict(initial_time=0, volume_fraction="10e-30"),
        "limits": dict(physical_time="0.00005"),
        "monomers": dict(number=1000),
        "numerics": dict(with_collisions="false", pick_method="last"),
        "output": dict(output_dir="out", write_between_event_frequency=1000),
    },
    "surface_growth": {
        **_COMMON_DLCA,
        "limits": dict(mean_monomere_per_aggregate=40),
        "monomers": dict(density=1800, dispersion_diameter="1.20", initialisation_mode="lognormal", mean_diameter=10,
                         number=800),
        "output": dict(n_time_per_file=5000, output_dir="out"),
        "surface_growth": dict(with_surface_reactions="true", flux_surfgrowth="1e-04", volsurf_method="alphas",
                               full_aggregate_update_frequency=100),
    },
    "pytest": {
        **_COMMON_DLCA,
        "limits": dict(mean_monomere_per_aggregate=40),
        "monomers": dict(density=1800, dispersion_diameter="1.20", initialisation_mode="lognormal", mean_diameter=10,
                         number=20),
        "output": dict(n_time_per_file=5000, output_dir="out"),
        "surface_growth": dict(with_surface_reactions="true", flux_surfgrowth="1e-04", volsurf_method="alphas",
                               full_aggregate_update_frequency=100),
    },
    "classic": {
        "monomers": dict(number=100, density=1800, dispersion_diameter="1.25", mean_diameter=10,
                         initialisation_mode="normal"),
        "environment": dict(initial_time=0, volume_fraction="1e-3", temperature=1700, pressure=101300,
                            fractal_prefactor="1.4", fractal_dimension="1.8"),
        "surface_growth": dict(with_surface_reactions="true", flux_surfgrowth="1e-5", volsurf_method="none",
                               full_aggregate_update_frequency=100),
        "limits": dict(number_of_aggregates=1, n_iter_without_event=-1, cpu=-1, physical_time=-1,
                       mean_monomere_per_aggregate=-1),
        "numerics": dict(enforce_volume_fraction="true", with_collisions="true", n_verlet_divisions=10,
                         pick_method="random", individual_surf_reactions="true"),
        "nucleation": dict(with_nucleation="true", flux="5e23"),
        "flame_coupling": dict(with_flame_coupling="false"),
        "output": dict(output_dir="out", n_time_per_file=10, write_between_event_frequency=90),
        "inter_potential": dict(with_potentials="true", interpotential_file="Interpotential_input.dat",
                                with_external_potentials="true"),
    },
}


def merged_config(base: str, overrides: dict | None = None) -> dict:
    cfg = {sec: dict(kv) for sec, kv in CONFIGS[base].items()}
    for sec, kv in (overrides or {}).items():
        cfg.setdefault(sec, {}).update(kv)
    return cfg
