"""Loader for tests/golden/*.npz (made by tests/golden/make_golden.py from the unmodified reference)."""
from __future__ import annotations

import hashlib
import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
SEARCH_KEYS = ["rand_calls", "source", "dir", "full_distance", "distance", "moving_sphere", "other_sphere", "moving_label",
               "other_label", "n_agg", "time"]
STEP_KEYS = ["label", "dt", "proper_time", "pos", "lpm"]
ORC_TO_STEP = {"label": "source", "dt": "dt", "proper_time": "proper_time", "pos": "pos", "lpm": "full_distance"}


class Golden:
    def __init__(self, name: str):
        self.name = name
        z = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
        self.z = z
        self.meta = json.loads(str(z["meta"]))
        self.base = self.meta["base"]
        self.overrides = self.meta["overrides"]
        self.searches = z["searches"]
        self.steps = z["steps"]
        self.merges = z["merges"]

    def state(self, prefix: str) -> dict:
        out: dict = {}
        for k in self.z.files:
            if not k.startswith(prefix + "/"):
                continue
            parts = k.split("/")[1:]
            if len(parts) == 2:
                out.setdefault(parts[0], {})[parts[1]] = self.z[k]
            else:
                v = self.z[k]
                out[parts[0]] = v.item() if v.ndim == 0 else v
        return out

    def sort(self, k: int) -> dict | None:
        if f"sort_{k}/idx" not in self.z.files:
            return None
        return {kk: self.z[f"sort_{k}/{kk}"] for kk in ("step", "n", "factor", "idx", "cum", "time_step")}


def digest_from_oracle_records(recs: np.ndarray, with_collisions: bool) -> str:
    """Same digest as make_golden.digest(), computed from oracle/GPU step records."""
    h = hashlib.sha256()
    if with_collisions:
        for k in SEARCH_KEYS:
            h.update(np.ascontiguousarray(recs[k]).tobytes())
    for k in STEP_KEYS:
        h.update(np.ascontiguousarray(recs[ORC_TO_STEP[k]]).tobytes())
    return h.hexdigest()


def write_interpotential_file(path) -> str:
    """Re-create the text table examples/classic.ini points at from the committed fixture (one value per line is a valid layout:
    the reference reads it with operator>>)."""
    vals = np.load(GOLDEN / "interpotential_table.npz")["values"]
    with open(path, "w") as f:
        f.write("\n".join(repr(float(v)) for v in vals))
        f.write("\n")
    return str(path)
