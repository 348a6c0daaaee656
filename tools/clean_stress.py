#!/usr/bin/env python3
"""Stress of the implication behind the tile-level ambiguity test (test infrastructure, CPU only): the host
build (tests/hostsim) runs the tracking sample loop beside the plain one on every run of every tile that
e1_par_clean marked and counts runs it flags or sums differently.  Prints one JSON line per workload.

    python tools/clean_stress.py [seeds]"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import e1util as U  # noqa: E402

hs = U.hostsim()
seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
CASES = [("2.6 MS/s, |f| < 4 kHz", 2.6e6, 260000, 40, 4000.0), ("2.6 MS/s, |f| < 100 Hz", 2.6e6, 260000, 40, 100.0),
         ("2.6 MS/s, |f| < 15 kHz", 2.6e6, 260000, 40, 15000.0), ("25 MS/s, |f| < 4 kHz", 25e6, 2500000, 4, 4000.0),
         ("4 MS/s, |f| < 6 kHz", 4.0e6, 400000, 24, 6000.0)]
for name, fs_nom, n_samp, n_ep, f_max in CASES:
    fs = U.fs_as_reference(fs_nom)
    c0, k0, t0, slow, exact = hs.hs_clean_tiles(), hs.hs_checked_tiles(), time.time(), 0, 0
    for seed in range(seeds):
        recs = U.synthetic_recs(n_ep, 36, fs, seed=500 + seed, max_chan=36, f_max=f_max)
        _, _, st = U.hostsim_synth(fs, n_samp, recs, planner=1)      # asserts hs_clean_violations() == 0
        slow += int(st[2])
        exact += int(st[0])
    print(json.dumps({"workload": name, "inputs": seeds, "blocks_each": n_ep, "channels": 36,
                      "tiles_marked_clean": hs.hs_clean_tiles() - c0, "tiles_tracked": hs.hs_checked_tiles() - k0,
                      "flagged_runs": slow, "samples_resolved_exactly": exact, "violations": hs.hs_clean_violations(),
                      "seconds": round(time.time() - t0, 1)}), flush=True)
