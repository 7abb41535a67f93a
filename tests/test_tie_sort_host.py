"""Host check of the K9 fast path for tie-dominated pick tables (mcac_b200/csrc/tie_sort.cuh): the header is host/device neutral,
so its sparse simulation + routing are run here on the CPU against libstdc++'s own std::sort / __introsort_loop (the routine
AggregatList::sort_time_steps uses, aggregat_list.cpp:109-123) on thousands of random tie-dominated tables."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_tie_sort_header_reproduces_std_sort(tmp_path):
    exe = tmp_path / "tie_sort_host"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", str(exe), str(ROOT / "tests" / "native" / "tie_sort_host.cpp")])
    for seed in (1, 2):
        out = subprocess.run([str(exe), "400", str(seed)], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        assert out.stdout.startswith("ok 400 cases"), out.stdout


def test_heap_sort_header_reproduces_libstdcxx_heap_branch(tmp_path):
    """mcac_b200/csrc/heap_sort.cuh (the branch std::sort takes behind introsort's depth limit; k_sort_heap on the device) against
    std::partial_sort and against libstdc++'s own __introsort_loop run with a small depth limit."""
    exe = tmp_path / "heap_sort_host"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", str(exe), str(ROOT / "tests" / "native" / "heap_sort_host.cpp")])
    out = subprocess.run([str(exe), "300", "7"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok 300 cases"), out.stdout


def test_enclosing_ball_pruning_never_drops_a_finite_pair(tmp_path):
    """sweep_may_touch (csrc/mcac_math.cuh), the pruning of the step loop's sphere-pair sweeps, against the exact pair test on
    random aggregate pairs with un-wrapped positions, several periodic images, grazing contacts and overlapping starts."""
    exe = tmp_path / "prune_host"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-o", str(exe), str(ROOT / "tests" / "native" / "prune_host.cpp")])
    for seed in (3, 11):
        out = subprocess.run([str(exe), "20000", str(seed)], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        assert out.stdout.startswith("ok 20000 cases"), out.stdout


def test_seq_cumsum_header_is_the_sequential_sum(tmp_path):
    """mcac_b200/csrc/seq_cumsum.cuh (cumulative_time_steps over a run of equal weights in closed form, aggregat_list.cpp:131-140)
    against the plain loop cum[i] = cum[i-1] + w, every entry bit for bit, incl. round-to-even ties and binade crossings; the head of
    lighter weights as integer prefix sums between irregular steps (chunked the way a 512-thread CTA runs it, on the padded views, with
    perturbed approximate sums: accepted heads must be exact); and the building CTA's sort passes (bitonic_pass) in k_event's order."""
    exe = tmp_path / "seq_cumsum_host"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-o", str(exe), str(ROOT / "tests" / "native" / "seq_cumsum_host.cpp")])
    for seed in (1, 9):
        out = subprocess.run([str(exe), "600", str(seed)], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        assert out.stdout.startswith("ok 600 cases"), out.stdout
        assert "head: 600 cases" in out.stdout and "sort: 160 tables" in out.stdout, out.stdout
