"""The C-ABI shared library: it loads, exports every entry point include/e1b200.h declares, its
structs have the documented sizes, and its host-side restate equals the oracle's.  No kernel is
launched here (this file runs on boxes without a GPU)."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import e1util as U

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "galileo-sdr-sim_b200"))
import build as B  # noqa: E402
import e1b200 as E  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    B.build_lib()
    return E.load()


def declared_symbols():
    text = (ROOT / "include" / "e1b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(e1b200_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(E.SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", str(E.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (e1b200_\w+)", out))
    assert exported == set(names)


def test_library_contains_sm100a_kernels_and_no_oracle(lib):
    out = subprocess.run(["cuobjdump", "-lelf", str(E.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    syms = subprocess.run(["nm", "-D", str(E.LIB_PATH)], capture_output=True, text=True).stdout
    assert "e1o_" not in syms          # the CPU oracle is not linked into the product
    assert "libe1oracle" not in subprocess.run(["ldd", str(E.LIB_PATH)], capture_output=True, text=True).stdout


def test_struct_layouts_match_header():
    assert E.REC_DTYPE.itemsize == 176 and E.RANGE_DTYPE.itemsize == 168
    assert E.REC_DTYPE == U.REC_DTYPE and E.RANGE_DTYPE == U.RANGE_DTYPE
    assert C.sizeof(E.Config) == 32 and C.sizeof(E.Timing) == 20 and C.sizeof(E.Stats) == 64


def test_host_restate_equals_oracle(lib):
    rng = np.random.default_rng(5)
    for _ in range(200):
        rho0 = rng.uniform(2.0e7, 2.8e7)
        rho1 = rho0 + rng.uniform(-90, 90)
        grx = rng.uniform(0, 604800)
        assert E.restate(rho0, rho1, U.REF_DT, grx) == U.oracle_restate(rho0, rho1, U.REF_DT, grx)


def test_version_and_argument_checks(lib):
    assert b"sm_100a" in lib.e1b200_version()
    h = C.c_void_p()
    bad = E.Config(2.6e6, 260000, 0, 0, 0, 0.0)
    assert lib.e1b200_create(C.byref(bad), C.byref(h)) == -1 and not h.value      # E1B200_EINVAL
    bad = E.Config(2.6e6, 260000, 65, 0, 0, 0.0)
    assert lib.e1b200_create(C.byref(bad), C.byref(h)) == -1
    assert lib.e1b200_destroy(None) == -1


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(E.E1B200Error):
        E.Synth(2.6e6, 260000, 16)


def test_bench_reads_the_newest_ncu_summary():
    """bench.py takes the synthesis kernel's DRAM traffic and instruction count from the newest committed
    `ncu --set full` summary under profiles/: newest by round and build number (r1_v10 comes after r1_v9, r2_v1 after both), and the
    file it names exists and holds both figures."""
    import importlib.util
    root = Path(__file__).resolve().parent.parent
    spec = importlib.util.spec_from_file_location("bench_mod", root / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    traffic, name, inst = bench.ncu_summary_numbers()
    import re
    builds = sorted(tuple(map(int, re.match(r"r(\d+)_v(\d+)_", f.name).groups())) for f in (root / "profiles").glob("r*_v*_synth_ncu_summary.txt")
                    if "cfg3" not in f.name)
    assert name == "r%d_v%d_synth_ncu_summary.txt" % builds[-1] and (root / "profiles" / name).exists()
    assert 3.0e9 < traffic < 4.0e9 and 5.0e9 < inst < 2.0e10
    # the 25 MS/s family reads the captures of the event-driven kernel on the 30 s slice (750 M samples per launch)
    t3 = bench.ncu_summary_numbers("cfg3")
    assert t3 and "cfg3" in t3[1] and t3[1].startswith("r2_") and (root / "profiles" / t3[1]).exists()
    assert 3.0e9 < t3[0] < 4.0e9 and t3[2] * 32 / (750e6 * 36) < 6.0          # fewer than 6 instruction slots per channel-sample


def test_code_wraps_equals_the_literal_loop(lib):
    """e1b200_code_wraps against the reference's own statements (src/galileo-sdr.cpp:491-494, :528) run
    sample by sample, including starts at or above 4092 (wrap at sample 0) and blocks whose last addition
    reaches 4092 (not counted: the wrap test of that value belongs to the next block)."""
    rng = np.random.default_rng(11)
    for fs_nom, n in ((2.6e6, 260000), (25e6, 250000), (4e6, 1000)):
        fs = U.fs_as_reference(fs_nom)
        delt = 1.0 / fs
        for trial in range(6):
            f_code = 1.023e6 + rng.uniform(-3, 3)
            cp0 = float(rng.uniform(0, 4092)) if trial else 4092.0 + 1e-9
            sc = np.float64(f_code) * np.float64(delt)
            cp, wraps = np.float64(cp0), 0
            for _ in range(n):
                if cp >= 4092.0:
                    cp -= 4092.0
                    wraps += 1
                cp = cp + sc
            assert E.code_wraps(fs, n, cp0, f_code) == wraps, (fs_nom, trial)
    # a block that ends exactly on a wrap: the start is chosen so that the last addition lands on >= 4092
    fs, n, f_code = U.fs_as_reference(2.6e6), 10400, 1.023e6
    sc = np.float64(f_code) * np.float64(1.0 / fs)
    cp = np.float64(0.0)
    for _ in range(n):
        cp = cp + sc
    start = float(4092.0 - cp + 1e-9)          # after n additions the sum is just above 4092, never tested in this block
    lit, c = 0, np.float64(start)
    for _ in range(n):
        if c >= 4092.0:
            c -= 4092.0
            lit += 1
        c = c + sc
    assert c >= 4092.0 and lit == 0 and E.code_wraps(fs, n, start, f_code) == 0
    assert E.code_wraps(fs, n + 1, start, f_code) == 1


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the B200 arm) needs no GPU: one JSON line with the
    B200 arm's metric / unit / config keys, impl = reference, a cpu_baseline describing the run and an e2e of its own.
    Under torchrun only rank 0 works; the other ranks exit 0 without a line."""
    import json
    import os
    root = Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--workload", "rt1", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-400:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "E1B/C IQ Msamples/sec" and d["unit"] == "Msamples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("real-time call shape")
    other = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--workload", "rt1", "--steps", "1", "--warmup", "0", "--gpus", "2"],
                           capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert other.returncode == 0 and not other.stdout.strip()
