#!/usr/bin/env python3
"""Debug aid: repeat a bench-size synthesis under several library switches and count the samples that
differ from the oracle (computed once); saves the differing positions."""
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

fs, n_samp, n_chan, n_epochs = U.fs_as_reference(2.6e6), 260000, 36, 2999
recs = U.synthetic_recs_fast(n_epochs, n_chan, fs, seed=4242)
ref, ph_ref = U.oracle_synth(fs, n_samp, recs, threads=min(os.cpu_count() or 1, n_chan))
modes = [("default", {}, int(os.environ.get("REPS", "12"))), ("no_pair", {"E1B200_NO_PAIR": "1"}, 2), ("serial_planner", {"E1B200_SERIAL_PLANNER": "1"}, 1),
         ("no_tma", {"E1B200_NO_TMA": "1"}, 1)]
out = np.empty((n_epochs * n_samp, 2), np.int16)
for name, env, reps in modes:
    for k in list(os.environ):
        if k.startswith("E1B200_") and k != "E1B200_LIB":
            del os.environ[k]
    os.environ.update(env)
    for r in range(reps):
        s = E.Synth(fs, n_samp, n_chan)
        s.synth_epochs(recs, out)
        ph = s.carrier_phases()
        s.close()
        bad = np.nonzero((out != ref).any(axis=1))[0]
        info = [{"idx": int(i), "epoch": int(i // n_samp), "sample": int(i % n_samp), "tile": int((i % n_samp) // 8192), "in_tile": int((i % n_samp) % 8192),
                 "cuda": out[i].tolist(), "oracle": ref[i].tolist()} for i in bad[:8]]
        print(json.dumps({"mode": name, "rep": r, "differing": int(len(bad)), "phases_equal": bool(np.array_equal(ph, ph_ref)), "where": info}), flush=True)
