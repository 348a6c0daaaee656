#!/usr/bin/env python3
"""Soak test (test infrastructure): bench-size syntheses repeated and compared bit for bit with the oracle
(computed once per input).  A timing-dependent defect shows up as a repetition that differs -- this is how
the out-of-table lookup fixed in round 1 was found (one wrong sample per ~10 runs of 780 M samples).

    python tools/find_mismatch.py [reps]        -> one JSON line per (workload, input); exit status 1 on any difference"""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "galileo-sdr-sim_b200"))
sys.path.insert(0, str(ROOT / "tests"))
import e1b200 as E  # noqa: E402
import e1util as U  # noqa: E402

REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 6
CASES = [  # name, fs, samples per block, channels, blocks, seed, f_max, ranges?
    ("cfg2", 2.6e6, 260000, 36, 2999, 4242, 4000.0, False),
    ("cfg2-seed7", 2.6e6, 260000, 36, 2999, 7, 4000.0, False),
    ("cfg3s", 25e6, 2500000, 36, 300, 11, 4000.0, False),
    ("cfg2-lowdoppler", 2.6e6, 260000, 36, 1000, 5, 30.0, False),
    ("cfg1-16slots", 2.6e6, 260000, 16, 999, 3, 4000.0, False),
    ("64ch", 2.6e6, 260000, 64, 400, 9, 4500.0, False),
    ("cfg4s-ranges", 25e6, 2500000, 12, 120, 21, 0.0, True),
]
threads = os.cpu_count() or 1
failed = 0
for name, fs_nom, n_samp, n_chan, n_epochs, seed, f_max, ranges in CASES:
    fs = U.fs_as_reference(fs_nom)
    if ranges:
        rr, recs = U.synthetic_ranges(n_epochs, n_chan, seed=seed)
    else:
        recs = U.synthetic_recs_fast(n_epochs, n_chan, fs, seed=seed, f_max=f_max)
    ref, ph_ref = U.oracle_synth(fs, n_samp, recs, threads=min(threads, n_chan))
    out = np.empty((n_epochs * n_samp, 2), np.int16)
    counts, where = [], []
    for r in range(REPS):
        s = E.Synth(fs, n_samp, n_chan)
        if ranges:
            s.synth_ranges(rr, out)
        else:
            s.synth_epochs(recs, out)
        ph = s.carrier_phases()
        s.close()
        bad = np.nonzero((out != ref).any(axis=1))[0]
        counts.append(int(len(bad)) + (0 if np.array_equal(ph, ph_ref) else 1))
        where += [{"rep": r, "epoch": int(i // n_samp), "sample": int(i % n_samp), "cuda": out[i].tolist(), "oracle": ref[i].tolist()} for i in bad[:4]]
    failed += sum(1 for c in counts if c)
    print(json.dumps({"case": name, "samples": int(n_epochs * n_samp), "channels": n_chan, "reps": REPS, "differing_per_rep": counts, "where": where}), flush=True)
sys.exit(1 if failed else 0)
