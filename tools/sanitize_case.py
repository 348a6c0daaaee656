#!/usr/bin/env python3
"""A small synthesis (3 blocks x 40000 samples, 5 channels, one idle gap: carry-walked kernel, planner with an
irregular span; with `ev`: the same at 25 MS/s through the event-driven kernel -- columns in shared memory,
red.shared, team barriers) for compute-sanitizer:
    compute-sanitizer --tool memcheck|racecheck python tools/sanitize_case.py [ev]
Prints whether the output equals the oracle's."""
import sys
sys.path.insert(0, "galileo-sdr-sim_b200"); sys.path.insert(0, "tests")
import numpy as np, e1b200 as E, e1util as U
fs, n_samp, nch = U.fs_as_reference(25e6 if "ev" in sys.argv[1:] else 2.6e6), 40000, 5
recs = U.synthetic_recs(3, nch, fs, seed=2, max_chan=6)
recs[1, 3]["prn"] = 0
s = E.Synth(fs, n_samp, 6, device=0)
out = s.synth_epochs(recs)
ref, _ = U.oracle_synth(fs, n_samp, recs, threads=4)
print("equal", np.array_equal(out, ref), s.stats().kernel_name)
s.close()
