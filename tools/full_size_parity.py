#!/usr/bin/env python3
"""Bit-exact comparison of the CUDA path with the CPU oracle at BASELINE sizes (test infrastructure; the
pytest suite compares at sizes the oracle finishes in seconds and uses size-independent properties at
full size -- this is the brute-force version for a box with enough cores and memory).

    python tools/full_size_parity.py [cfg2|cfg3s|cfg1] ...

Prints one JSON line per workload: samples, differing int16 pairs, md5 of both streams, seconds."""
import hashlib
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "galileo-sdr-sim_b200"))
sys.path.insert(0, str(ROOT / "tests"))
import e1b200 as E  # noqa: E402
import e1util as U  # noqa: E402

WL = {"cfg1": (2.6e6, 260000, 8, 99), "cfg2": (2.6e6, 260000, 36, 2999), "cfg3s": (25e6, 2500000, 36, 300)}


def main():
    for name in (sys.argv[1:] or ["cfg2"]):
        fs_nom, n_samp, n_chan, n_epochs = WL[name]
        fs = U.fs_as_reference(fs_nom)
        recs = U.synthetic_recs_fast(n_epochs, n_chan, fs, seed=4242)
        t0 = time.perf_counter()
        s = E.Synth(fs, n_samp, n_chan)
        out = s.synth_epochs(recs)
        ph = s.carrier_phases()
        s.close()
        t1 = time.perf_counter()
        ref, ph_ref = U.oracle_synth(fs, n_samp, recs, threads=min(os.cpu_count() or 1, n_chan))
        t2 = time.perf_counter()
        diff = int(np.count_nonzero((out != ref).any(axis=1)))
        print(json.dumps({"workload": name, "samples": int(out.shape[0]), "channels": n_chan, "differing_samples": diff,
                          "carrier_phases_equal": bool(np.array_equal(ph, ph_ref)),
                          "md5_cuda": hashlib.md5(out.tobytes()).hexdigest(), "md5_oracle": hashlib.md5(ref.tobytes()).hexdigest(),
                          "cuda_s": round(t1 - t0, 2), "oracle_s": round(t2 - t1, 2)}), flush=True)


if __name__ == "__main__":
    main()
