"""The drop-in, dropped in: the reference's OWN host code -- its main(), galileo_task(), RINEX reader, orbit /
pseudorange / I-NAV page generators and channel allocation, compiled from where they lie by `make -C oracle
refb200` -- with its sample loop (src/galileo-sdr.cpp:481-539) replaced by one e1b200_synth_epochs call per
0.1 s block (oracle/ref_patches/b200_dropin.sed + oracle/ref_hooks/b200_patch.cpp, n_epochs = 1 per call).
The files it writes must be the files the unmodified reference writes.  The binaries are built in the
build container (build()) and travel to the GPU box; nothing here reads /root/reference."""
import hashlib
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref"
GOLD = Path(__file__).parent / "golden"
NAV = GOLD / "week171_subset.rnx"


def run_dropin(binary, args, out):
    if not (REF / binary).exists():
        pytest.skip(f"oracle/_ref/{binary} not built (needs /root/reference at build time)")
    # the reference aborts at exit by design flaw (joinable std::thread destroyed) AFTER closing
    # its file: the exit status says nothing, the file is the result
    r = subprocess.run([str(REF / binary)] + args + ["-e", str(NAV), "-o", str(out), "-U", "1", "-b", "1"],
                       capture_output=True, text=True, timeout=900)
    assert "ERROR" not in r.stderr, r.stderr[-400:]
    assert out.exists(), r.stderr[-400:]
    return r


def golden_hashes(name):
    lines = (GOLD / f"{name}_sha256.txt").read_text().splitlines()
    return lines[0].split()[2], lines[1:]


@pytest.mark.gpu
@pytest.mark.parametrize("binary,name,n,args", [
    ("usrp_galileo_b200", "cfg1", 260000, ["-l", "-6,51,100", "-d", "10"]),
    ("usrp_galileo_b200", "paris45", 260000, ["-l", "48.85,2.35,35", "-t", "2021/06/20,11:59:40", "-d", "45"]),
    ("usrp_galileo_ch36_b200", "ch36", 260000, ["-l", "-6,51,100", "-d", "35"]),
    ("usrp_galileo_fs25ch36_b200", "fs25ch36", 2500000, ["-l", "-6,51,100", "-d", "3"]),
])
def test_reference_host_code_with_the_b200_sample_loop_writes_the_reference_file(tmp_path, binary, name, n, args):
    """configs[0] (md5 419622c8...), the 45 s Paris run (22 page turns and the 30 s re-allocation: the page
    hand-over decided by e1b200_code_wraps), 24 satellites in 36 slots for 35 s, and 24 satellites at 25 MS/s:
    every 0.1 s block's SHA-256 and the file's md5 equal the unmodified reference's."""
    out = tmp_path / "dropin.ishort"
    run_dropin(binary, args, out)
    md5, sha = golden_hashes(name)
    raw = out.read_bytes()
    assert len(raw) == len(sha) * n * 4
    iq = np.frombuffer(raw, np.int16).reshape(len(sha), n, 2)
    bad = [e for e in range(len(sha)) if hashlib.sha256(iq[e].tobytes()).hexdigest() != sha[e]]
    assert not bad, bad[:10]
    assert hashlib.md5(raw).hexdigest() == md5


def test_dropin_fails_loudly_without_a_device(tmp_path):
    """No CPU fallback behind the reference's command line either: without a usable CUDA device the patched
    executable stops like the reference does on its other fatal errors (message on stderr, exit(1))."""
    import os
    if not (REF / "usrp_galileo_b200").exists():
        pytest.skip("oracle/_ref/usrp_galileo_b200 not built")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([str(REF / "usrp_galileo_b200"), "-l", "-6,51,100", "-d", "1", "-e", str(NAV), "-o", str(tmp_path / "x.ishort"),
                        "-U", "1", "-b", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 1 and "e1b200_create failed" in r.stderr
