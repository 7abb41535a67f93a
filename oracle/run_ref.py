"""TEST INFRASTRUCTURE ONLY: run the compiled reference (oracle/_ref/MCAC or MCAC_tap).

Writes an .ini (base file + overrides) into a fresh scratch directory, runs the binary there with a
fresh ``output_dir`` (an existing one makes the reference block on stdin,
src/physical_model/physical_model.cpp:196-210) and returns the scratch path.  Never reads
/root/reference at run time: base configs are the dict literals of mcac_b200/configs.py, restated from
validation/*.ini and examples/classic.ini (SURVEY.md Appendix C lists the keys).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import tempfile
from pathlib import Path

import sys

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
REF_DIR = HERE / "_ref"

from mcac_b200.configs import CONFIGS, merged_config  # noqa: E402,F401  (the workload dict literals live on the product side)


def write_ini(path: Path, cfg: dict) -> None:
    with open(path, "w") as f:
        for sec, kv in cfg.items():
            f.write(f"[{sec}]\n")
            for k, v in kv.items():
                f.write(f"{k}={v}\n")
            f.write("\n")


def run_reference(base: str, overrides: dict | None = None, *, tap: bool = True, env: dict | None = None,
                  workdir: str | None = None, timeout: float | None = None, taskset_core: int | None = None):
    """Run the reference; returns (workdir Path, stdout str). Tap outputs land in workdir/tap."""
    exe = REF_DIR / ("MCAC_tap" if tap else "MCAC")
    if not exe.exists():
        raise FileNotFoundError(f"{exe} missing: run `make -C oracle/ref_build` where /root/reference exists")
    wd = Path(workdir or tempfile.mkdtemp(prefix="mcac_ref_"))
    wd.mkdir(parents=True, exist_ok=True)
    cfg = merged_config(base, overrides)
    out_dir = wd / cfg["output"]["output_dir"]
    if out_dir.exists():
        shutil.rmtree(out_dir)
    write_ini(wd / "params.ini", cfg)
    (wd / "tap").mkdir(exist_ok=True)
    e = dict(os.environ)
    e["MCAC_TAP_DIR"] = str(wd / "tap")
    e.update({k: str(v) for k, v in (env or {}).items()})
    cmd = [str(exe), "params.ini"]
    if taskset_core is not None and shutil.which("taskset"):
        cmd = ["taskset", "-c", str(taskset_core)] + cmd
    p = subprocess.run(cmd, cwd=wd, env=e, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"reference exited with {p.returncode}:\n{p.stdout[-2000:]}")
    return wd, p.stdout


def read_summary(wd: Path) -> dict:
    out = {}
    for line in (Path(wd) / "tap" / "summary.txt").read_text().splitlines():
        k, _, v = line.partition(" ")
        if k == "chunk_times":
            out[k] = [float(x) for x in v.split()]
            continue
        try:
            out[k] = int(v)
        except ValueError:
            try:
                out[k] = float(v)
            except ValueError:
                out[k] = v
    return out


if __name__ == "__main__":
    import argparse
    import json

    ap = argparse.ArgumentParser()
    ap.add_argument("base")
    ap.add_argument("--set", action="append", default=[], help="section.key=value")
    ap.add_argument("--env", action="append", default=[], help="NAME=value (tap controls)")
    ap.add_argument("--workdir")
    ap.add_argument("--no-tap", action="store_true")
    a = ap.parse_args()
    ov: dict = {}
    for s in a.set:
        k, v = s.split("=", 1)
        sec, key = k.split(".", 1)
        ov.setdefault(sec, {})[key] = v
    env = dict(s.split("=", 1) for s in a.env)
    wd, out = run_reference(a.base, ov, tap=not a.no_tap, env=env, workdir=a.workdir)
    print(out[-3000:])
    print("workdir:", wd)
    if not a.no_tap:
        print(json.dumps(read_summary(wd)))
