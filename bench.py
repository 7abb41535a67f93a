#!/usr/bin/env python
"""bench.py — MC steps/s and contact-pair tests/s of the per-step aggregation hot path at N = 1e6 (BASELINE.json).

Workload (config.workload = "c3"): validation/params_brownian.ini scaled as SURVEY.md §8d prescribes —
number=1e6, volume_fraction=1000e-6, n_verlet_divisions=100, with_collisions=true, pick_method=random, duplication
off, random_seed=42(+rank): monodisperse 30 nm monomers, the reference's own placement procedure.  One bench
"step" = `--mc-steps` consecutive MC steps of one realization (pick -> direction -> contact search -> translate ->
merge/update).  N GPUs = N independent realizations (the path does not shard: replicas only, weak scaling),
followed by one NCCL all-gather of the morphology statistics.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n-monomers 1000000] [--mc-steps 20000]
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "mc_steps_per_sec_at_N1e6"
UNIT = "MC steps/s"


def workload_config(n_monomers: int, seed: int) -> tuple[str, dict]:
    ov = {"monomers": {"number": n_monomers}, "environment": {"volume_fraction": "1000e-6"},
          "limits": {"physical_time": -1},
          "numerics": {"with_collisions": "true", "pick_method": "random", "n_verlet_divisions": 100 if n_monomers >= 200000 else 16,
                       "with_domain_duplication": "false", "random_seed": seed}}
    return "brownian", ov


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def reference_run(n_monomers: int, seed: int, mc_steps: int, n_chunks: int, core: int | None = None) -> dict:
    """The reference's own CPU implementation of the path (oracle/_ref/MCAC_tap = unmodified sources + counters), one thread
    (the reference is single-threaded), bounded sample: n_chunks x mc_steps MC steps after its own initial placement."""
    from oracle.run_ref import read_summary, run_reference

    base, ov = workload_config(n_monomers, seed)
    ov["output"] = {"write_between_event_frequency": 1000000000, "write_events_frequency": 1000000000}
    kind = "reference"
    exe = ROOT / "oracle" / "_ref" / "MCAC_tap"
    if not exe.exists():
        raise FileNotFoundError("oracle/_ref/MCAC_tap is not built")
    t0 = time.perf_counter()
    wd, _ = run_reference(base, ov, env={"MCAC_TAP_EXIT_STEP": mc_steps * n_chunks, "MCAC_TAP_CHUNK": mc_steps}, taskset_core=core)
    total_wall = time.perf_counter() - t0
    s = read_summary(wd)
    shutil.rmtree(wd, ignore_errors=True)
    ct = [0.0] + s["chunk_times"]
    chunk_s = [b - a for a, b in zip(ct[:-1], ct[1:])]
    return dict(kind=kind, chunk_s=chunk_s, steps=s["steps"], pair_tests=s["pair_tests_sphere"] + s["pair_tests_bounding"],
                calcul_wall_s=s["calcul_wall_s"], init_s=total_wall - s["calcul_wall_s"])


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def classic_texts(indices: list[int], table_path: str) -> list[str]:
    import mcac_b200
    from mcac_b200 import ensemble as ens
    from mcac_b200.configs import merged_config

    return [mcac_b200.ini_text(merged_config("classic", {"numerics": {"random_seed": s}, "inter_potential": {"interpotential_file": table_path}}))
            for s in ens.seeds(1000, indices)]


def ensemble_reference(a, K: int, W: int, M: int) -> dict:
    """Reference arm of the ensemble: the unmodified reference binary, one process per realization on every host core in parallel,
    bounded sample: `cores` realizations x (W + K) x M MC steps of examples/classic.ini (seeds 1000+k)."""
    import tempfile
    from concurrent.futures import ThreadPoolExecutor

    sys.path.insert(0, str(ROOT / "tests"))
    from golden_lib import write_interpotential_file
    from oracle.run_ref import read_summary, run_reference

    cores = max(1, min(os.cpu_count() or 1, a.realizations, 32))
    tmp = tempfile.mkdtemp(prefix="mcac_ens_ref_")
    table = write_interpotential_file(Path(tmp) / "Interpotential_input.dat")

    def one(k: int):
        ov = {"numerics": {"random_seed": 1000 + k}, "inter_potential": {"interpotential_file": table},
              "output": {"write_between_event_frequency": 1000000000, "write_events_frequency": 1000000000}}
        wd, _ = run_reference("classic", ov, env={"MCAC_TAP_EXIT_STEP": M * (W + K), "MCAC_TAP_CHUNK": M}, taskset_core=k)
        s = read_summary(wd)
        shutil.rmtree(wd, ignore_errors=True)
        ct = [0.0] + s["chunk_times"]
        return [b - c for c, b in zip(ct[:-1], ct[1:])]

    with ThreadPoolExecutor(cores) as ex:
        chunks = list(ex.map(one, range(cores)))
    shutil.rmtree(tmp, ignore_errors=True)
    # every process advances M steps per chunk, all in parallel: ensemble rate = cores * M / (slowest process' mean chunk time)
    per_proc = [sum(c[W:W + K]) / max(1, len(c[W:W + K])) for c in chunks]
    return {"value": cores * M / max(per_proc), "cores": cores, "ms_per_step": 1e3 * max(per_proc)}


ENSEMBLE_TOTAL_STEPS = 3200  # per realization: what the unmodified reference reaches in classic.ini's own `cpu = 60` budget (3220 steps, this image)


def ensemble_main(a, rank: int, world: int, local: int):
    """--workload ensemble (the default for --gpus N > 1): R realizations of examples/classic.ini (seeds 1000+k), realization k on rank
    k mod N, no communication during the run.  Deterministic stop: an MC-step count — every realization advances exactly
    (W + K) * M steps with M = ceil(3200 / (W + K)) unless --mc-steps is given (classic.ini's only practical stop is the wall-clock
    `cpu = 60`, which the reference turns into ~3 220 steps on this image; physical time reached varies with the seed).  One bench
    step = every realization advances M steps (one launch of k_ensemble_loop per round of host services); the statistic rows (K11)
    are all-gathered over NCCL once at the end."""
    K, W = a.steps, max(a.warmup, 0)
    M = a.mc_steps or -(-ENSEMBLE_TOTAL_STEPS // max(1, W + K))
    R = a.realizations
    config = {"workload": "ensemble", "ini": "examples/classic.ini (100 monomers, growth + nucleation + external potentials), random_seed=1000+k",
              "realizations": R, "mc_steps_per_step": M, "stop": f"MC-step count: {(W + K) * M} steps per realization",
              "window_mc_steps": [W * M, (W + K) * M],
              "parallelism": f"realization k -> rank k mod {world}; replicas only, one all-gather of the K11 rows at the end",
              "l2": "per-realization state (<= 35 MB, 1024 realizations) exceeds L2 in total; the sphere sweep of a step re-reads one aggregate pair from L1/L2"}
    metric, unit = "ensemble_mc_steps_per_sec", "MC steps/s (sum over realizations)"
    if a.impl == "reference":
        if rank != 0:
            return
        try:
            r = ensemble_reference(a, K, W, M)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"impl": "reference", "unavailable": str(e)[:200]}))
            return
        print(json.dumps({"impl": "reference", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": a.gpus, "steps": K, "warmup": W,
                          "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "reference",
                                           "sample": f"{r['cores']} realizations in parallel (one pinned process per core), {W + K} x {M} MC steps each, cpu: {cpu_model()}"},
                          "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mcac_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = ensemble_measure(a, rank, world, local, dist, config, metric, unit, K, W, M, R)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def ensemble_measure(a, rank, world, local, dist, config, metric, unit, K, W, M, R):
    """The ensemble workload on the ranks of `dist` (None: this process alone); returns the JSON line on rank 0."""
    import tempfile

    import numpy as np  # noqa: F401
    import torch

    import mcac_b200
    from mcac_b200 import ensemble as ens

    sys.path.insert(0, str(ROOT / "tests"))
    from golden_lib import write_interpotential_file  # the committed copy of the table classic.ini points at (fixture data)

    tmp = tempfile.mkdtemp(prefix="mcac_ens_")
    table = write_interpotential_file(Path(tmp) / "Interpotential_input.dat")
    # room for the three domain duplications of the run (100 -> 51 200 spheres + nucleated monomers) from the start, so that the device
    # loop duplicates by itself (mcac_gpu_reserve; ~40 MB per realization)
    os.environ.setdefault("MCAC_B200_RESERVE_SPHERES", "60000")
    os.environ.setdefault("MCAC_B200_RESERVE_AGGREGATES", "4096")
    mine = ens.shard(R, rank, world)
    threads = a.threads or max(1, min(32, (os.cpu_count() or 8) // max(1, world)))
    t0 = time.perf_counter()
    e = mcac_b200.Ensemble(classic_texts(mine, table), device=local)
    init_s = time.perf_counter() - t0

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(W):
        e.run(M, threads=threads)
    sampler = ClockSampler(local)
    sync_all()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    steps_done = launches = events = pair_tests = pair_exec = rounds = 0
    kernel_ms = 0.0
    host_ms = [0.0] * 4
    host_services = [0] * 3
    phase_cycles = [0] * 8
    stats_s = 0.0
    d2h = 0
    rows = None
    for _ in range(K):
        reps = e.run(M, threads=threads)
        steps_done += sum(r["steps"] for r in reps)
        launches += sum(r["kernel_launches"] for r in reps)
        events += sum(r["events"] for r in reps)
        pair_tests += sum(r["pair_tests_sphere"] + r["pair_tests_bounding"] for r in reps)
        pair_exec += sum(r["pair_tests_executed"] for r in reps)
        kernel_ms += reps[0]["device_ms"] if reps else 0.0
        for k, name in enumerate(("search_ms", "commit_ms", "event_ms", "cells_ms")):
            host_ms[k] += reps[0][name] if reps else 0.0
        for k, name in enumerate(("search_launches", "commit_launches", "event_launches")):
            host_services[k] += reps[0][name] if reps else 0
        rounds += reps[0]["conflicts"] if reps else 0
        for k in range(8):
            phase_cycles[k] += sum(r["loop_phase_cycles"][k] for r in reps)
        # the step's result: the statistic rows of every realization cross to the host (timed apart: only e2e includes it)
        t_s = time.perf_counter()
        rows = e.morphology_stats(ens.N_BINS, 2e-6)
        stats_s += time.perf_counter() - t_s
        d2h = rows.nbytes
    sync_all()
    wall_s = time.perf_counter() - t_wall
    clocks = sampler.stop()
    del ev0, ev1
    t = torch.tensor([wall_s - stats_s, wall_s], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(steps_done), float(launches), float(events), float(pair_tests), float(pair_exec)] + [float(c) for c in phase_cycles],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    te = torch.tensor([float(t[1]), float(tot[0])], dtype=torch.float64, device="cuda")
    # the only collective of the path
    full = ens.gather_rows(rows, mine, R, dist=dist, device="cuda")
    line = None
    if rank == 0:
        value = float(tot[0]) / float(t[0])
        # the dominant kernel is k_ensemble_loop; its time is the sphere-pair sweep of the contact search (FP64-pipe-bound: both
        # aggregates of a pair are re-read from L1/L2, 104 FP64 ops per pair test).  Peaks measured here.
        sim0 = e.sims[0]
        dfma = sim0.kernel_bench("fp64_dfma", reps=3)
        dmad = sim0.kernel_bench("fp64_dmul_dadd", reps=3)
        peak_dfma = dfma["units"] / (dfma["ms"] * 1e-3) / 1e12
        peak_nofma = dmad["units"] / (dmad["ms"] * 1e-3) / 1e12
        pair_rate = float(tot[3]) / float(t[0])
        exec_rate = float(tot[4]) / float(t[0])  # sphere-pair tests the pruned sweeps actually executed
        fp64_tflops = exec_rate * 104.0 / 1e12
        cyc = [float(x) for x in tot[5:13]]
        phase_share = {n: (c / sum(cyc) if sum(cyc) else None) for n, c in
                       zip(["pick_table", "cells_and_search", "update_block", "nucleation_refresh", "loop_top", "move", "growth", "merge"], cyc)}
        peaks = {}
        try:
            peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_gbs = exec_rate * 32.0 / 1e9  # 32 B (x, y, z, r) of the candidate sphere per executed pair test
        cpu_baseline = None
        if world == 1 and not a.no_cpu_baseline:
            try:
                r = ensemble_reference(a, K, W, M)
                cpu_baseline = {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "reference",
                                "sample": f"{r['cores']} realizations in parallel (one pinned process per core), {W + K} x {M} MC steps each "
                                          f"(timed: the last {K} x {M}), cpu: {cpu_model()}"}
            except Exception as ex:  # noqa: BLE001
                cpu_baseline = {"value": None, "unit": unit, "cores": None, "kind": "reference", "sample": f"unavailable: {str(ex)[:160]}"}
        line = ({"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
                          "ms_per_step": 1e3 * float(t[0]) / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic", "config": dict(config, host_threads_per_rank=threads),
                          "timing": "wall clock between device synchronisations around the K steps (every step = rounds of one k_ensemble_loop launch + "
                                    "host services), max over ranks",
                          "mc_steps_timed": int(tot[0]), "events_timed": int(tot[2]), "pair_tests_per_sec": pair_rate,
                          "pair_tests_executed_per_sec": exec_rate, "loop_phase_share": phase_share, "init_s": init_s, "clocks": clocks,
                          "rank0_host_ms_per_step": dict(zip(["services", "launch_wait_readback", "bookkeeping", "whole_call"], [x / K for x in host_ms])),
                          "rank0_host_services_timed": dict(zip(["duplications", "table_regrows", "rng_refills"], host_services)),
                          "rank0_kernel_ms_per_step": kernel_ms / K, "rank0_rounds_per_step": rounds / K,
                          "rank0_kernel_share_of_wall": kernel_ms * 1e-3 / (wall_s - stats_s) if wall_s > stats_s else None,
                          "e2e": {"value": float(te[1]) / float(te[0]), "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(d2h),
                                  "what": "mcac_ensemble_run(M steps per realization) + mcac_gpu_morphology_stats of every realization to the host, per step; "
                                          "the realizations are created on the host (placement) and live in HBM from then on"},
                          "gpu_launches": int(tot[1]),
                          "roofline": {"kernel": "k_ensemble_loop (per-realization step loop; time = ordered sphere-pair sweep of the contact search)",
                                       "bound": "hbm", "achieved": alg_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": alg_gbs / hbm_peak, "traffic": None,
                                       "binding": "fp64", "fp64_achieved_tflops": fp64_tflops, "fp64_peak_dfma_tflops": peak_dfma,
                                       "fp64_peak_dmul_dadd_tflops": peak_nofma, "fp64_frac_of_dmul_dadd_peak": fp64_tflops / peak_nofma if peak_nofma else None,
                                       "note": "algorithmic bytes = 32 B per sphere-pair test; the sweep re-reads two aggregates from L1/L2, so the binding "
                                               "roofline is the FP64 pipe: 104 FP64 ops per pair test against the DMUL+DADD peak measured by "
                                               "k_fp64_peak<1> (the library is built --fmad=false)"},
                          "cpu_baseline": cpu_baseline, "ensemble_stats": ens.summarize(full)})
    shutil.rmtree(tmp, ignore_errors=True)
    del e
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mcac_b200", choices=["mcac_b200", "reference"])
    ap.add_argument("--n-monomers", type=int, default=1_000_000)
    ap.add_argument("--mc-steps", type=int, default=0, help="MC steps per bench step (default 20000; 400 for --impl reference)")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-table", action="store_true")
    ap.add_argument("--no-ensemble", action="store_true", help="c3 workload at one GPU: do not append the 1-GPU point of the ensemble workload")
    ap.add_argument("--workload", default=None, choices=["c3", "ensemble"],
                    help="c3: BASELINE's N=1e6 metric (default).  ensemble: --realizations independent runs of examples/classic.ini "
                         "(seeds 1000+k) sharded k -> rank k mod N, all-gather of the morphology statistics at the end")
    ap.add_argument("--realizations", type=int, default=1024)
    ap.add_argument("--threads", type=int, default=0, help="host threads driving the realizations of one rank (default: cores / ranks, <= 32)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    K, W = a.steps, max(a.warmup, 0)
    if a.workload is None:
        # one GPU: BASELINE's N = 1e6 metric (the path does not shard); several GPUs: the multi-GPU workload BASELINE names, the
        # 1024-realization ensemble of examples/classic.ini sharded over the ranks (replicas of the N = 1e6 run cannot fail to scale)
        a.workload = "c3" if a.gpus <= 1 and world <= 1 else "ensemble"
    if a.workload == "ensemble":
        return ensemble_main(a, rank, world, local)
    config = {"workload": "c3", "ini": "validation/params_brownian.ini + number, volume_fraction=1000e-6, n_verlet_divisions=100, "
              "with_collisions=true, pick_method=random, with_domain_duplication=false", "n_monomers": a.n_monomers,
              "volume_fraction_ppm": 1000, "random_seed": "42+rank", "parallelism": f"replicas x{a.gpus} (independent realizations)",
              "l2": "state (>300 MB of sphere + aggregate SoA at N=1e6) exceeds the 126 MB L2; no explicit flush"}

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if a.impl == "reference":
        if rank != 0:
            return
        M = a.mc_steps or 400
        # weak scaling like our arm: N GPUs = N independent realizations (seed 42 + k), here N single-threaded reference processes
        # side by side on the host cores (a single trajectory cannot use more than one thread)
        n_proc = max(1, min(a.gpus, os.cpu_count() or 1))
        try:
            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(n_proc) as ex:
                runs = list(ex.map(lambda k: reference_run(a.n_monomers, 42 + k, M, W + K, core=k), range(n_proc)))
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"impl": "reference", "unavailable": str(e)[:200]}))
            return
        r = runs[0]
        per_proc = [sum(x["chunk_s"][W:W + K]) for x in runs]
        timed = [max(per_proc) / K] * K  # the slowest replica bounds the job, as the max over ranks does on our arm
        value = n_proc * M * K / max(per_proc)
        r = dict(r, pair_tests=sum(x["pair_tests"] for x in runs))
        config["mc_steps_per_step"] = M
        # the CPU arm cannot reach the GPU arm's window within minutes (100 000 MC steps take ~3.5 min at ~480 steps/s): it times the
        # head of the same trajectory; its per-step cost is flat over the run (profiles/r2_tuning.md)
        config["window_mc_steps"] = [W * M, (W + K) * M]
        config["pinning"] = f"taskset -c k for process k (k < {n_proc})"
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * sum(timed) / len(timed), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_proc, "kind": r["kind"],
                                 "sample": f"{n_proc} realization(s) x {W + K} x {M} MC steps of the same N={a.n_monomers} workload, one thread "
                                           f"each (the reference is single-threaded), placement ({r['init_s']:.1f} s) excluded, cpu: {cpu_model()}"},
                "pair_tests_per_sec": r["pair_tests"] / r["calcul_wall_s"],
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    import numpy as np
    import torch

    import mcac_b200
    from mcac_b200.configs import merged_config

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mcac_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    M = a.mc_steps or 20000
    config["mc_steps_per_step"] = M
    config["window_mc_steps"] = [W * M, (W + K) * M]
    config["batch"] = a.batch
    base, ov = workload_config(a.n_monomers, 42 + rank)
    text = mcac_b200.ini_text(merged_config(base, ov))
    t0 = time.perf_counter()
    sim = mcac_b200.Simulation(text, device=local)
    init_s = time.perf_counter() - t0

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(W):
        sim.run(M, batch=a.batch)
    sim.set_profile(True)
    sampler = ClockSampler(local)
    sync_all()
    sampler.start()
    t_wall = time.perf_counter()
    reps = [sim.run(M, batch=a.batch)[0] for _ in range(K)]
    sync_all()
    wall_s = time.perf_counter() - t_wall
    clocks = sampler.stop()
    sim.set_profile(False)
    dev_ms = sum(r["device_ms"] for r in reps)
    steps_done = sum(r["steps"] for r in reps)
    pair_tests = sum(r["pair_tests_sphere"] + r["pair_tests_bounding"] for r in reps)
    t = torch.tensor([dev_ms, wall_s * 1e3], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(steps_done), float(pair_tests), float(sum(r["kernel_launches"] for r in reps))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    max_ms = float(t[0])
    value = float(tot[0]) / (max_ms * 1e-3)

    # ---- rooflines, measured live with CUDA events on the handle's stream (profile mode of mcac_gpu_run)
    searches = sum(r["searches"] for r in reps)
    ps = sum(r["pair_tests_sphere"] for r in reps); pb = sum(r["pair_tests_bounding"] for r in reps)
    n_launch = max(1, sum(r["search_launches"] for r in reps))
    search_ms = sum(r["search_ms"] for r in reps)
    commit_ms = sum(r["commit_ms"] for r in reps)
    event_ms = sum(r["event_ms"] for r in reps)
    n_event = max(1, sum(r["event_launches"] for r in reps))
    peaks = {}
    try:
        peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_source = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    k9_traffic = {}
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from an ncu --set full capture of THIS build and window
        k9_traffic = json.load(open(ROOT / "profiles" / "r2_k9_traffic.json"))
    except (OSError, ValueError):
        pass
    n_agg_now = reps[-1]["n_aggregates"]
    # K9 (dominant kernel of the step): the per-event pipeline in one cooperative launch.  Algorithmic bytes per aggregate
    # (one pass): liveness + label 12, refresh/totals 24, weight 8, sorted index 4, cumulative 8, pick slot 4 = 60 B.
    k9_bytes = 60.0 * n_agg_now
    k9_gbs = k9_bytes / (event_ms / n_event * 1e-3) / 1e9 if event_ms > 0 else 0.0
    phases = ["reduce", "labels", "sort_init", "grid_levels", "local_levels", "leaves", "cumulative", "pick_table"]
    sorts = max(1, sum(r["sorts"] for r in reps))
    roofline = {"kernel": "k_event (K9: labels + refresh + totals + 1/dt weights + replayed introsort [sparse simulation + routing for "
                          "tie-dominated tables] + cumulative table, one cooperative launch per merge)", "bound": "hbm", "achieved": k9_gbs, "peak": peak, "unit": "GB/s", "frac": k9_gbs / peak,
                "traffic": k9_traffic.get("dram_bytes_per_launch") if a.n_monomers == 1_000_000 else None,
                "traffic_source": k9_traffic.get("source", "no ncu capture of this build committed (profiles/r2_k9_traffic.json)"),
                "peak_source": peak_source, "launches": n_event, "avg_launch_us": 1e3 * event_ms / n_event,
                "algorithmic_bytes_per_launch": k9_bytes, "share_of_step": event_ms / dev_ms if dev_ms else None,
                "sort_levels_per_launch": sum(r["sort_levels"] for r in reps) / sorts,
                "sort_span_elements_per_launch": sum(r["sort_span_elements"] for r in reps) / sorts,
                "phase_us_per_launch": {**{p: sum(r["event_phase_cycles"][i] for r in reps) / 1965.0 / sorts for i, p in enumerate(phases)},
                                        **{p: sum(r["tie_phase_cycles"][i] for r in reps) / 1965.0 / sorts
                                           for i, p in enumerate(["tie_sparse_simulation", "tie_routing_pass"])},
                                        **{p: sum(r["tie_sim_cycles"][i] for r in reps) / 1965.0 / sorts
                                           for i, p in enumerate(["tie_sim_gather", "tie_sim_levels", "tie_sim_handover_barrier"])}},
                "tie_fast_path": {"sorts": sum(r["tie_sorts"] for r in reps), "of_sorts": sorts,
                                  "levels_per_sort": sum(r["tie_levels"] for r in reps) / max(1, sum(r["tie_sorts"] for r in reps)),
                                  "sparse_per_sort": sum(r["tie_sparse"] for r in reps) / max(1, sum(r["tie_sorts"] for r in reps)),
                                  "handed_per_sort": sum(r["tie_handed"] for r in reps) / max(1, sum(r["tie_sorts"] for r in reps))},
                "note": "latency-bound, not bandwidth-bound: dependent introsort levels x grid barriers; tie-dominated tables take the "
                        "sparse fast path (csrc/tie_sort.cuh) for the top levels",
                "commit_share_of_step": commit_ms / dev_ms if dev_ms else None, "search_share_of_step": search_ms / dev_ms if dev_ms else None,
                "avg_commit_us": 1e3 * commit_ms / max(1, sum(r["commit_launches"] for r in reps)),
                "avg_search_us": 1e3 * search_ms / n_launch}
    # K1 inside the loop (speculative batches of <= 256 searches: launch-latency-bound), algorithmic bytes per search
    # (SURVEY §8d restated for aggregate-level candidates): 36 B per bounding test (x,y,z,rmax + slot id), 32 B per sphere of the
    # sphere-level sweep (moving + other), 48 B of result
    alg_bytes = 36.0 * pb + 32.0 * (ps + searches) + 48.0 * searches
    achieved = alg_bytes / (search_ms * 1e-3) / 1e9 if search_ms > 0 else 0.0
    roofline_k1_inloop = {"kernel": "k_search_group<32> + k_search_wide (K1, speculative batch inside the MC loop)", "bound": "hbm",
                          "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "launches": n_launch,
                          "avg_launch_us": 1e3 * search_ms / n_launch, "algorithmic_bytes_per_launch": alg_bytes / n_launch}
    # the same kernel fed with one launch of ~3.5e5 independent searches (what an ensemble / wide speculation gives it)
    sw = sim.search_sweep(349000, repeats=3)
    sw_bytes = 36.0 * sw["pair_tests_bounding"] + 32.0 * (sw["pair_tests_sphere"] + sw["n_queries"]) + 48.0 * sw["n_queries"]
    sw_gbs = sw_bytes / (sw["kernel_ms"] * 1e-3) / 1e9
    roofline_sweep = {"kernel": "k_search_group<8> (K1), one launch of %d searches" % sw["n_queries"], "bound": "hbm", "achieved": sw_gbs,
                      "peak": peak, "unit": "GB/s", "frac": sw_gbs / peak, "traffic": None, "kernel_ms": sw["kernel_ms"],
                      "pair_tests_per_sec": (sw["pair_tests_bounding"] + sw["pair_tests_sphere"]) / (sw["kernel_ms"] * 1e-3),
                      "searches_per_sec": sw["n_queries"] / (sw["kernel_ms"] * 1e-3)}

    # ---- e2e: the same metric through the C ABI with HOST buffers: upload_state (H2D) + run + download_state (D2H) per step
    e2e = None
    if a.e2e_steps > 0:
        # The host owns the state (the reference's SoA arrays, here in page-locked memory); every step uploads it, runs M MC
        # steps and reads the new state back.  Array sizes change with the merges, so each step re-binds views of the same
        # pinned blocks (n_agg shrinks; n_sph is constant in this workload).
        ns0, na0 = sim.sizes()
        pin = sim.pinned_state_buffers(ns0, na0)

        def views(na):
            v = dict(pin)
            v["agg_fields"] = pin["agg_fields"].reshape(-1)[:21 * na].reshape(21, na)
            v["agg_cell"] = pin["agg_cell"].reshape(-1)[:3 * na].reshape(3, na)
            for k in ("agg_n_spheres", "agg_charge"):
                v[k] = pin[k][:na]
            v["offsets"] = pin["offsets"][:na + 1]
            return v

        cur = views(na0)
        sim.download_into(cur)
        h2d = d2h = 0
        torch.cuda.synchronize()
        t_e = time.perf_counter()
        steps_e = 0
        for _ in range(a.e2e_steps):
            sim.upload_from(cur)
            h2d = sum(cur[k].nbytes for k in ("sphere_fields", "sphere_charge", "agg_fields", "agg_charge", "agg_cell", "offsets",
                                              "members", "per_member"))
            r, _ = sim.run(M, batch=a.batch)
            cur = views(r["n_aggregates"])
            sim.download_into(cur)
            steps_e += r["steps"]
            d2h = sum(v.nbytes for k, v in cur.items() if k != "_pinned")
        torch.cuda.synchronize()
        e_s = time.perf_counter() - t_e
        sim.free_pinned(pin)
        e = torch.tensor([e_s], dtype=torch.float64, device="cuda")
        es = torch.tensor([float(steps_e)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e, op=dist.ReduceOp.MAX)
            dist.all_reduce(es, op=dist.ReduceOp.SUM)
        e2e = {"value": float(es[0]) / float(e[0]), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "what": "per step: mcac_gpu_upload_state(host SoA, pinned) + mcac_gpu_run(M MC steps) + mcac_gpu_download_state(host SoA, "
                       "pinned); wall clock over the steps, max over ranks"}

    # ---- the only collective of the path: all-gather of the per-realization morphology statistics (K11)
    n_bins = 24
    stats = torch.zeros(2 * n_bins + 8, dtype=torch.float64, device="cuda")
    sim.morphology_stats_device(stats.data_ptr(), n_bins, 2e-6)
    ensemble = None
    if world > 1:
        gathered = torch.zeros(world * (2 * n_bins + 8), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(gathered, stats)
        g = gathered.view(world, -1).cpu().numpy()
    else:
        g = stats.view(1, -1).cpu().numpy()
    if rank == 0:
        tail = g[:, 2 * n_bins:]
        ensemble = {"realizations": int(g.shape[0]), "n_agg": [int(x) for x in tail[:, 0]], "mean_npp": [float(x[1] / x[0]) for x in tail],
                    "mean_rg_nm": [float(1e9 * x[7] / x[0]) for x in tail]}

    # ---- per-kernel table on the resident state (last: growth / update rewrite derived fields)
    kernels = None
    if rank == 0 and not a.no_kernel_table:
        bytes_per_unit = {"cells": 88, "grow": 32, "update_partial": 208, "update_full": 208, "event_sort": 60, "event_nosort": 36,
                          "grid_barriers_x100": 0, "rng_fill": 4, "morphology_stats": 28}
        kernels = []
        for k, bpu in bytes_per_unit.items():
            r = sim.kernel_bench(k, reps=5)
            gbs = r["units"] * bpu / (r["ms"] * 1e-3) / 1e9 if r["ms"] > 0 else None
            kernels.append({"kernel": k, "us": 1e3 * r["ms"], "units": r["units"], "bytes_per_unit": bpu, "GBps": gbs,
                            "frac_of_hbm_peak": gbs / peak if gbs else None})

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            Mr = 400
            r = reference_run(a.n_monomers, 42, Mr, 8, core=0)
            timed = r["chunk_s"][2:]
            cpu_baseline = {"value": Mr * len(timed) / sum(timed), "unit": UNIT, "cores": 1, "kind": r["kind"],
                            "sample": f"{len(timed)} x {Mr} MC steps of the same N={a.n_monomers} workload (seed 42) after 2 warm-up chunks, one thread pinned with taskset "
                                      f"(the reference is single-threaded), placement ({r['init_s']:.1f} s) excluded, cpu: {cpu_model()}",
                            "pair_tests_per_sec": r["pair_tests"] / r["calcul_wall_s"]}
        except Exception as e:  # noqa: BLE001
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": f"unavailable: {str(e)[:160]}"}

    ensemble_point = None
    if world == 1 and not a.no_ensemble:
        # the multi-GPU workload of BASELINE.json (what `--gpus N` runs for N > 1) at ONE GPU, so that its scaling can be read against
        # this run: value(N) / (N * ensemble.value)
        del sim
        Me = -(-ENSEMBLE_TOTAL_STEPS // max(1, W + K))
        econf = {"workload": "ensemble", "realizations": a.realizations, "mc_steps_per_step": Me, "stop": f"MC-step count: {(W + K) * Me} steps per realization"}
        a_e = argparse.Namespace(**{**vars(a), "no_cpu_baseline": True})
        try:
            e_line = ensemble_measure(a_e, 0, 1, local, None, econf, "ensemble_mc_steps_per_sec", "MC steps/s (sum over realizations)", K, W, Me, a.realizations)
            ensemble_point = {k: e_line[k] for k in ("metric", "value", "unit", "ms_per_step", "config", "pair_tests_per_sec", "pair_tests_executed_per_sec",
                                                     "loop_phase_share", "e2e", "gpu_launches", "roofline")}
        except Exception as ex:  # noqa: BLE001
            ensemble_point = {"unavailable": str(ex)[:200]}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": max_ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "pair_tests_per_sec": float(tot[1]) / (max_ms * 1e-3), "mc_steps_timed": int(tot[0]),
                "events_timed": sum(r["events"] for r in reps), "wall_ms_per_step": float(t[1]) / K, "init_placement_s": init_s,
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(tot[2]), "roofline": roofline, "roofline_k1_inloop": roofline_k1_inloop,
                "roofline_k1_sweep": roofline_sweep, "kernels": kernels,
                "cpu_baseline": cpu_baseline, "ensemble_stats": ensemble, "ensemble": ensemble_point}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
