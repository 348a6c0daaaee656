#!/usr/bin/env python3
"""Generate tests/golden/* from the reference itself (build container only).

Runs oracle/_ref/usrp_galileo_trace (reference sources + state-dump hook, see oracle/Makefile)
on fixed scenarios and stores, per scenario:
  <name>_recs.npz     e1_epoch_rec[n_epochs][16] derived from the dumped channel state, plus the
                      reference's carrier phase at the top of every epoch (the exactness pin for
                      the carrier planner)
  <name>_sha256.txt   SHA-256 of every 0.1 s block of the reference's ishort output, md5 of the file
  <name>_samples.npz  raw int16 slices of a few blocks (first/page-turn/last) for quick diffs
It also checks the unhooked binary writes the same bytes, so the hook is arithmetic-neutral.
"""
import hashlib
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import e1util as U  # noqa: E402

REF = Path("/root/reference")
BIN = ROOT / "oracle" / "_ref"

NAV = str(REF / "rinex_files/week171.rnx")
# name: (binary suffix, samples per block, fs, MAX_CHAN of that build, args)
SCENARIOS = {
    # BASELINE.json configs[0]
    "cfg1": ("", 260000, 2.6e6, 16, ["-l", "-6,51,100", "-e", NAV, "-d", "10"]),
    # second pin: other site/time, grx ~ 43200 s, crosses the 30 s re-allocation check (:545-562)
    "paris45": ("", 260000, 2.6e6, 16, ["-l", "48.85,2.35,35", "-t", "2021/06/20,11:59:40", "-e", NAV, "-d", "45"]),
    # Patched builds of the reference (oracle/ref_patches/*.diff, `make -C oracle refp`): the rates and channel
    # counts of BASELINE configs[1]-[4], which the as-shipped build cannot run.
    #   fs25      SAMP_RATE 25e6, the 8 satellites of configs[0], 3 s
    #   ch36      MAX_CHAN 36 + elevation mask off: all 24 satellites of the RINEX file, 35 s (crosses the
    #             30 s re-allocation and every channel's page turns) -- the channel count configs[1] is quoted on
    #   fs25ch36  both: 24 satellites at 25 MS/s, 3 s -- the shape of configs[2]
    "fs25": ("_fs25", 2500000, 25e6, 16, ["-l", "-6,51,100", "-e", NAV, "-d", "3"]),
    "ch36": ("_ch36", 260000, 2.6e6, 36, ["-l", "-6,51,100", "-e", NAV, "-d", "35"]),
    "fs25ch36": ("_fs25ch36", 2500000, 25e6, 36, ["-l", "-6,51,100", "-e", NAV, "-d", "3"]),
}


def run(binary, args, out, trace=None):
    env = dict(os.environ)
    if trace:
        env["E1_TRACE_OUT"] = str(trace)
    # exit status is 134 by design flaw of the reference (joinable std::thread destroyed); file is complete
    subprocess.run([str(binary)] + args + ["-o", str(out), "-U", "1", "-b", "1"], env=env,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def main():
    subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle"), "ref", "refp", "libe1oracle.so"])
    gold = ROOT / "tests" / "golden"
    gold.mkdir(exist_ok=True)
    only = sys.argv[1:]
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for name, (suffix, N, fs_nom, max_chan, args) in SCENARIOS.items():
            if only and name not in only:
                continue
            out, tr, plain = td / "o.ishort", td / "t.bin", td / "p.ishort"
            run(BIN / f"usrp_galileo{suffix}_trace", args, out, tr)
            run(BIN / f"usrp_galileo{suffix}", args, plain)
            raw = out.read_bytes()
            assert raw == plain.read_bytes(), "trace hook changed the output"
            iq = np.frombuffer(raw, np.int16).reshape(-1, N, 2)
            trace = np.fromfile(tr, U.TRACE_DTYPE)
            recs, phase = U.trace_to_recs(trace, max_chan)
            assert recs.shape[0] == iq.shape[0]
            mine, _ = U.oracle_synth(U.fs_as_reference(fs_nom), N, recs, threads=8)
            assert np.array_equal(mine.reshape(iq.shape), iq), "oracle restatement differs from the reference"
            np.savez_compressed(gold / f"{name}_recs.npz", recs=recs, phase=phase,
                                grx=np.array(sorted(set(trace["grx"]))))
            with open(gold / f"{name}_sha256.txt", "w") as f:
                f.write("# md5 %s  bytes %d  args %s\n" % (hashlib.md5(raw).hexdigest(), len(raw), " ".join(args).replace(str(REF) + "/", "")))
                for e in range(iq.shape[0]):
                    f.write(hashlib.sha256(iq[e].tobytes()).hexdigest() + "\n")
            turn = [e for e in range(recs.shape[0]) if any(r["prn"] and r["ibit0"] + 25 >= 500 for r in recs[e])]
            keep = sorted(set([0, iq.shape[0] - 1] + turn[:2]))
            np.savez_compressed(gold / f"{name}_samples.npz", epochs=np.array(keep),
                                head=np.stack([iq[e, :4096] for e in keep]), tail=np.stack([iq[e, -4096:] for e in keep]))
            print(name, iq.shape, "satellites", sorted(set(int(x) for x in recs["prn"].ravel()) - {0}), "md5", hashlib.md5(raw).hexdigest(), "page-turn epochs", turn[:4])


if __name__ == "__main__":
    main()
