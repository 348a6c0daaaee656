#!/usr/bin/env python3
"""Per-input timing of the default bench workload (test/diagnostic tool): planner and synthesis
milliseconds and the planner's serial walks for the record seeds bench.py gives ranks 0..7."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "galileo-sdr-sim_b200"))
sys.path.insert(0, str(ROOT / "tests"))
import e1b200 as E  # noqa: E402
import e1util as U  # noqa: E402

fs, n_samp, n_chan, n_ep = U.fs_as_reference(2.6e6), 260000, 36, 2999
for seed in range(1000, 1000 + int(sys.argv[1]) if len(sys.argv) > 1 else 1008):
    recs = U.synthetic_recs_fast(n_ep, n_chan, fs, seed=seed)
    d_recs = torch.from_numpy(recs.view(np.uint8).reshape(-1)).cuda()
    d_out = torch.empty(n_ep * n_samp * 2, dtype=torch.int16, device="cuda")
    s = E.Synth(fs, n_samp, n_chan)
    best = None
    for it in range(3):
        s.synth_epochs_device(n_ep, d_recs.data_ptr(), d_out.data_ptr())
        s.sync()
        t = s.timing()
        if best is None or t.plan_ms + t.synth_ms < best[0] + best[1]:
            best = (t.plan_ms, t.synth_ms)
    st = s.stats()
    f = recs[0]["f_carr"][:n_chan]
    print(json.dumps({"seed": seed, "plan_ms": round(best[0], 3), "synth_ms": round(best[1], 3), "serial": int(st.serial_epochs),
                      "exact_samples": int(st.exact_samples), "min_abs_f": round(float(np.abs(f).min()), 1),
                      "mean_abs_f": round(float(np.abs(f).mean()), 1)}), flush=True)
    s.close()
    del d_out, d_recs
