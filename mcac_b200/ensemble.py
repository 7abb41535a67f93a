"""Ensemble sharding + statistics: realization k runs on rank k mod world, no communication during the run, one
all-gather of the fixed-size per-realization morphology rows (K11) at the end.  Pure host logic + torch.distributed plumbing;
the realizations themselves are mcac_b200.Simulation handles (CUDA)."""
from __future__ import annotations

import numpy as np

N_BINS = 24


def shard(n_realizations: int, rank: int, world: int) -> list[int]:
    """Realization indices of this rank: round-robin, `k -> rank k mod world`."""
    return list(range(rank, n_realizations, world))


def seeds(first_seed: int, indices: list[int]) -> list[int]:
    """`random_seed = first_seed + k` (1000 + k)."""
    return [first_seed + k for k in indices]


def gather_rows(local_rows: np.ndarray, indices: list[int], n_realizations: int, dist=None, device=None) -> np.ndarray:
    """All-gather of the per-realization statistic rows.  Every rank contributes a fixed-size block (padded to the largest shard)
    and returns the full (n_realizations, row) table in realization order.  `dist` = torch.distributed (None: single process)."""
    row = local_rows.shape[1] if local_rows.size else 2 * N_BINS + 8
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        out = np.zeros((n_realizations, row))
        out[indices] = local_rows
        return out
    import torch

    world = dist.get_world_size()
    per = (n_realizations + world - 1) // world
    block = torch.zeros(per, row + 1, dtype=torch.float64, device=device)  # last column: realization index + 1 (0 = padding)
    if len(indices):
        block[:len(indices), :row] = torch.as_tensor(local_rows, dtype=torch.float64, device=device)
        block[:len(indices), row] = torch.as_tensor(np.asarray(indices, dtype=np.float64) + 1.0, device=device)
    gathered = torch.zeros(world * per, row + 1, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(gathered, block)
    g = gathered.cpu().numpy()
    out = np.zeros((n_realizations, row))
    for r in g:
        if r[row] > 0:
            out[int(r[row]) - 1] = r[:row]
    return out


def linreg(x: np.ndarray, y: np.ndarray):
    """mcac::linreg (src/tools/tools.cpp:126-157): least squares of log y on log x -> (ok, a, b, r).  Its `r` raises the product of
    the variances to the power 2 where a square root is meant (:153-155) — reproduced, not fixed."""
    lx, ly = np.log(np.asarray(x, dtype=np.float64)), np.log(np.asarray(y, dtype=np.float64))
    return linreg_from_sums(float(len(lx)), lx.sum(), (lx * lx).sum(), (lx * ly).sum(), ly.sum(), (ly * ly).sum())


def linreg_from_sums(n, sumx, sumx2, sumxy, sumy, sumy2):
    denom = n * sumx2 - sumx ** 2
    if int(n) == 0 or abs(denom) < 1e-9:  # "singular matrix. can't solve the problem."
        return False, 0.0, 0.0, 0.0
    a = (n * sumxy - sumx * sumy) / denom
    b = (sumy * sumx2 - sumx * sumxy) / denom
    with np.errstate(divide="ignore", invalid="ignore"):
        r = float(np.float64(sumxy - sumx * sumy / n) / np.float64(((sumx2 - sumx ** 2 / n) * (sumy2 - sumy ** 2 / n)) ** 2))
    return True, float(a), float(b), r


def fractal_law(rows: np.ndarray) -> list[tuple[float, float]]:
    """Per-realization (Df, kf) from the K11 sums: AggregatList::get_instantaneous_fractal_law
    (src/aggregats/aggregat_list_fractal_law.cpp:23-33): regression of log Np on log(Dg/Dp); Df = slope, kf = exp(intercept)."""
    out = []
    for r in rows:
        n, _, sx, sx2, sxy, sy, sy2, _ = r[2 * N_BINS:]
        ok, a, b, _ = linreg_from_sums(n, sx, sx2, sxy, sy, sy2)
        out.append((float(a), float(np.exp(b))) if ok else (float("nan"), float("nan")))
    return out


def summarize(rows: np.ndarray) -> dict:
    tail = rows[:, 2 * N_BINS:]
    n_agg = tail[:, 0]
    law = fractal_law(rows)
    df = np.array([d for d, _ in law]); kf = np.array([k for _, k in law])
    ok = np.isfinite(df)
    return {"realizations": int(rows.shape[0]), "n_agg_total": int(n_agg.sum()),
            "mean_npp": float(tail[:, 1].sum() / max(1.0, n_agg.sum())), "mean_rg_nm": float(1e9 * tail[:, 7].sum() / max(1.0, n_agg.sum())),
            "np_histogram_log2": [int(v) for v in rows[:, :N_BINS].sum(axis=0)],
            "rg_histogram": [int(v) for v in rows[:, N_BINS:2 * N_BINS].sum(axis=0)],
            "Df_mean": float(df[ok].mean()) if ok.any() else None, "kf_mean": float(kf[ok].mean()) if ok.any() else None}
