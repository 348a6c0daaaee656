#!/usr/bin/env python3
"""bench.py -- Galileo E1B/C IQ synthesis throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg1] [--impl reference]

A *step* is one pass of the hot path over one job of synthetic channel records (SURVEY.md
section 8d): default workload = BASELINE configs[1], static, 2.6 MS/s, 36 channels, 300 s
(2999 blocks of 0.1 s = 779.74 MS = 3.12 GB of int16 I/Q) on each GPU.
  value  samples of all ranks / device time, records and output resident in HBM (CUDA events on
         the library's stream, max over ranks)
  e2e    the same job through the C-ABI host entry point e1b200_synth_epochs(): records in
         pinned host memory, H2D + planner + synthesis + D2H into a pinned host buffer
  roofline  e1_synth_kernel: 4 B per output sample (algorithmic bytes) / its CUDA-event duration
         against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (oracle/e1_oracle.c, pinned to the reference's output) on a bounded
         slice of the same workload, all host threads
N > 1: one process per GPU (torchrun), time-axis shards with no data-path collective -- every
rank synthesises its own 300 s segment (weak scaling); NCCL only carries the barrier and the
max-over-ranks of the times.
--impl reference: the reference's CPU algorithm (oracle port; the reference binary itself is a
fixed 16-channel / 2.6 MS/s build that needs its RINEX tree) on rank 0's host cores.
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "galileo-sdr-sim_b200"))
sys.path.insert(0, str(ROOT / "tests"))

WORKLOADS = {
    # name: (fs, samples/epoch, channels, epochs, description)
    "cfg1": (2.6e6, 260000, 8, 99, "BASELINE configs[0]: static -l -6,51,100 -e week171.rnx, 2.6 MS/s, 8 satellites, 10 s (99 blocks)"),
    "cfg2": (2.6e6, 260000, 36, 2999, "BASELINE configs[1]: static, 2.6 MS/s, 36 ch, 300 s (2999 blocks)"),
    "cfg3": (25e6, 2500000, 36, 2999, "BASELINE configs[2]: static, 25 MS/s, 36 ch, 300 s (2999 blocks)"),
    "cfg3s": (25e6, 2500000, 36, 300, "BASELINE configs[2] slice: 25 MS/s, 36 ch, 30 s (300 blocks)"),
    "cfg4s": (25e6, 2500000, 36, 300, "BASELINE configs[3] slice: dynamic receiver, 25 MS/s, 36 ch, 30 s (300 blocks), per-block "
                                      "Doppler/code-phase restate on the device from pseudorange records"),
}
WORKLOADS["cfg5s"] = (25e6, 2500000, 36, 2400, "BASELINE configs[4] slice: 25 MS/s, 36 ch, 240 s (2400 blocks = 24 GB of int16 I/Q); meant for --gpus N: "
                                               "the `strong` object is ONE such scenario sharded on the time axis over the N GPUs and assembled in rank 0's HBM")
WORKLOADS["rt1"] = (2.6e6, 260000, 16, 1, "real-time call shape: ONE 0.1 s block per call (what the reference's galileo_task() loop hands over per iteration), "
                                          "2.6 MS/s, 16 slots, 8 satellites; a step = one call: ms_per_step is the latency of a call")
METRIC = "E1B/C IQ Msamples/sec"
UNIT = "Msamples/s"


def synthetic_range_records(n_epochs, n_chan, dtype, seed=0, dt=0.100000023142, grx0=43200.0):
    """Pseudorange-level inputs for the device-side restate (e1_range_rec): per channel a range of
    2.2-2.6e7 m, a range rate of +-760 m/s (+-4 kHz of Doppler) and an acceleration of +-0.5 m/s^2 (a
    moving receiver), random pages."""
    rng = np.random.default_rng(seed)
    rr = np.zeros((n_epochs, n_chan), dtype)
    t0 = (np.arange(n_epochs) * dt)[:, None]
    t1 = t0 + dt
    rho0, v, a = rng.uniform(2.2e7, 2.6e7, n_chan)[None, :], rng.uniform(-760, 760, n_chan)[None, :], rng.uniform(-0.5, 0.5, n_chan)[None, :]
    rr["prn"] = (np.arange(n_chan) % 50 + 1)[None, :]
    rr["rho_prev"] = rho0 + v * t0 + 0.5 * a * t0 * t0
    rr["rho_cur"] = rho0 + v * t1 + 0.5 * a * t1 * t1
    rr["grx_sec"] = grx0 + t1
    rr["flags"][0] = 1                                            # E1_REC_SET_PHASE
    rr["carr_phase_init"][0] = rng.uniform(0, 1, n_chan)
    pages = rng.integers(0, 256, (n_chan, 64), dtype=np.uint8)
    pages[:, 62] &= 0x0F                                           # 500 symbols: bits 500..511 stay zero
    pages[:, 63] = 0
    rr["page_cur"] = pages[None, :, :]
    rr["page_next"] = pages[None, :, :]
    return rr


def ncu_summary_numbers(family="cfg2"):
    """From the newest committed `ncu --set full` summary of the synthesis kernel (profiles/*synth_ncu_summary.txt, one
    launch): DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) and warp instructions executed
    (smsp__inst_executed.sum).  family "cfg2": the captures of the default workload (one launch = one step);
    "cfg3": the captures of the 25 MS/s slice (names with cfg3: one launch = 300 blocks of 2.5 M samples; the caller
    scales per sample)."""
    best = None

    def version(f):     # r<round>_v<build>_...: numeric order (v10 after v9)
        m = re.match(r"r(\d+)_v(\d+)_", f.name)
        return (int(m.group(1)), int(m.group(2))) if m else (0, 0)

    for f in sorted((ROOT / "profiles").glob("*synth*ncu_summary.txt"), key=version):
        if ("cfg3" in f.name) != (family == "cfg3"):
            continue
        tot, inst, mult = 0.0, None, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for ln in f.read_text().splitlines():
            p = ln.split()
            if len(p) >= 3 and p[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and p[2] in mult:
                tot += float(p[1]) * mult[p[2]]
            if len(p) >= 2 and p[0] == "smsp__inst_executed.sum":
                inst = float(p[1])
        if tot:
            best = (tot, f.name, inst)
    return best


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


try:
    ALL_CPUS = os.sched_getaffinity(0)
except Exception:
    ALL_CPUS = None


def unbind_cpus():
    """The CPU baseline uses every host core again."""
    if ALL_CPUS:
        try:
            os.sched_setaffinity(0, ALL_CPUS)
        except Exception:
            pass


def bind_to_gpu_numa_node(local_rank):
    """Run this rank (and first-touch its pinned buffers) on the CPUs of the NUMA node its GPU hangs
    off: D2H into memory of the other socket costs PCIe bandwidth (the e2e arm is a 3 GB copy per step)
    and with 8 ranks everybody would otherwise crowd node 0.  Returns the node number or None."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        if all(hasattr(pr, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bdf = f"{int(pr.pci_domain_id):04x}:{int(pr.pci_bus_id):02x}:{int(pr.pci_device_id):02x}.0"
        else:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = vis.split(",")[local_rank] if vis else str(local_rank)
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", idx],
                                 capture_output=True, text=True, timeout=20).stdout.strip()
            bdf = out.splitlines()[0].strip().lower()
            if len(bdf.split(":")[0]) == 8:      # nvidia-smi prints an 8-digit domain, sysfs uses 4
                bdf = bdf[4:]
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(fs, n_samp, n_chan, seconds_target=15.0):
    """CPU oracle on a bounded slice of the workload, all host threads.  Returns the dict for the JSON line."""
    import e1util as U
    cores = os.cpu_count() or 1
    threads = min(cores, n_chan)
    recs = U.synthetic_recs_fast(4, n_chan, fs, seed=123)
    t0 = time.perf_counter()
    U.oracle_synth(fs, n_samp, recs[:1], threads=threads)          # calibration + table build
    per_epoch = max(time.perf_counter() - t0, 1e-3)
    n_ep = int(min(max(seconds_target / per_epoch, 2), 400))
    recs = U.synthetic_recs_fast(n_ep, n_chan, fs, seed=124)
    t0 = time.perf_counter()
    U.oracle_synth(fs, n_samp, recs, threads=threads)
    dt = time.perf_counter() - t0
    return {"value": n_ep * n_samp / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_ep} blocks x {n_samp} samples x {n_chan} ch ({n_ep * n_samp / fs:.1f} s of signal) in {dt:.1f} s; "
                      f"oracle/e1_oracle.c (channels split over {threads} threads)"}


def parity_check(d_out, recs, fs, n_samp, n_chan, world, use_ranges, dt_epoch=0.100000023142):
    """The bytes the timed loop left in d_out against the CPU oracle (oracle/e1_oracle.c, checker only, after the
    timed regions).  All blocks when the host has the cores for it (~15 s budget), else a spread: first, last,
    blocks that turn a page, and evenly spaced ones -- each started from the phase the LITERAL carrier recurrence
    (e1o_carrier_phases, no planner involved) reaches at the top of that block.  Returns the dict for the JSON line."""
    import e1util as U
    t0 = time.perf_counter()
    n_epochs = recs.shape[0]
    if use_ranges:                       # pseudorange records: computeCodePhase by the oracle's restatement
        er = np.zeros(recs.shape, U.REC_DTYPE)
        for k in ("prn", "flags", "carr_phase_init", "page_cur", "page_next"):
            er[k] = recs[k]
        for e in range(n_epochs):
            for c in range(recs.shape[1]):
                r = recs[e, c]
                if r["prn"] > 0:
                    fc, fcode, cp, ib, _ = U.oracle_restate(float(r["rho_prev"]), float(r["rho_cur"]), dt_epoch, float(r["grx_sec"]))
                    er[e, c]["f_carr"], er[e, c]["f_code"], er[e, c]["code_phase0"], er[e, c]["ibit0"] = fc, fcode, cp, ib
        recs = er
    threads = max(1, min((os.cpu_count() or 1) // world, n_chan))
    budget = int(15.0 * 69e6 * threads / (n_samp * n_chan))          # blocks the oracle does in ~15 s (69 M channel-samples/s/thread)
    out_blocks = d_out.view(n_epochs, n_samp * 2)
    differing = 0
    if budget >= n_epochs:
        blocks, how = list(range(n_epochs)), "all blocks"
        ph, step = None, 256
        for a in range(0, n_epochs, step):
            ref, ph = U.oracle_synth(fs, n_samp, recs[a:a + step], ph, threads=threads)
            got = out_blocks[a:a + step].cpu().numpy().reshape(-1, 2)
            differing += int(np.count_nonzero((got != ref).any(axis=1)))
    else:
        act = recs["prn"] > 0
        turns = np.nonzero((act & (recs["ibit0"] + 26 >= 500)).any(axis=1))[0]
        n = max(min(budget * 2 // 3, n_epochs), 3)                     # a third of the budget goes to the carrier-only walk
        pick = set(np.linspace(0, n_epochs - 1, max(n - min(len(turns), n // 3), 2)).astype(int).tolist())
        pick |= set(turns[np.linspace(0, len(turns) - 1, min(len(turns), n // 3)).astype(int)].tolist()) if len(turns) else set()
        blocks, how = sorted(pick), "first, last, page-turn and evenly spaced blocks; start phases from the literal carrier recurrence"
        phases, _ = U.oracle_carrier_phases(fs, n_samp, recs, threads=threads)
        for b in blocks:
            ref, _ = U.oracle_synth(fs, n_samp, recs[b:b + 1], phases[b], threads=threads)
            got = out_blocks[b].cpu().numpy().reshape(-1, 2)
            differing += int(np.count_nonzero((got != ref).any(axis=1)))
    return {"blocks": len(blocks), "of": n_epochs, "samples_compared": len(blocks) * n_samp, "differing_samples": differing,
            "what": f"d_out of the last timed step vs oracle/e1_oracle.c ({how})", "threads": threads,
            "seconds": round(time.perf_counter() - t0, 1)}


REF_BIN = ROOT / "oracle" / "_ref" / "usrp_galileo"
NAV_SUBSET = ROOT / "tests" / "golden" / "week171_subset.rnx"


def reference_binary_cfg1(seconds=10):
    """The reference's OWN executable (oracle/_ref/usrp_galileo: its unmodified sources, built by
    oracle/Makefile where /root/reference exists; the binary travels with the repo) on BASELINE
    configs[0]: `-l -6,51,100 -e week171 -d 10`, 8 satellites, 2.6 MS/s, one thread (the reference's
    generator is single-threaded).  It aborts on exit by design flaw after closing its file (SURVEY
    fact 9), so only the file and its own "Process time" line count.  Returns None if unavailable."""
    if not (REF_BIN.exists() and NAV_SUBSET.exists()):
        return None
    import hashlib
    import re
    import tempfile
    out = Path(tempfile.gettempdir()) / f"e1_ref_{os.getpid()}.ishort"
    t0 = time.perf_counter()
    try:
        r = subprocess.run([str(REF_BIN), "-l", "-6,51,100", "-e", str(NAV_SUBSET), "-o", str(out), "-U", "1", "-b", "1", "-d", str(seconds)],
                           capture_output=True, text=True, timeout=600)
    except Exception:
        return None
    wall = time.perf_counter() - t0
    if not out.exists():
        return None
    n_samples = out.stat().st_size // 4
    md5 = hashlib.md5(out.read_bytes()).hexdigest() if n_samples * 4 < 2e8 else None
    out.unlink()
    m = re.search(r"Process time = ([0-9.]+)", r.stderr + r.stdout)
    t = float(m.group(1)) if m and float(m.group(1)) > 0 else wall
    return {"value": n_samples / t / 1e6, "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": f"oracle/_ref/usrp_galileo -l -6,51,100 -e week171_subset.rnx -d {seconds}: {n_samples} samples, 8 satellites, "
                      f"process time {t:.1f} s (wall {wall:.1f} s), md5 {md5}"}


def run_reference(args, wl):
    """--impl reference: the reference's CPU algorithm on the host cores (rank 0 only)."""
    import e1util as U
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fs_nom, n_samp, n_chan, n_epochs, desc = wl
    fs = U.fs_as_reference(fs_nom)
    threads = min(os.cpu_count() or 1, n_chan)
    if args.workload == "cfg1" and REF_BIN.exists() and NAV_SUBSET.exists():
        # the reference's own binary on its own runnable case: every step is one full run
        vals = [reference_binary_cfg1() for _ in range(args.warmup + args.steps)][args.warmup:]
        vals = [v for v in vals if v]
        if vals:
            val = sum(v["value"] for v in vals) / len(vals)
            ms = 1e3 * 99 * n_samp / (val * 1e6)
            print(json.dumps({
                "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64 phase + int32 accumulate", "data": "real: RINEX week171 subset, -l -6,51,100, 8 satellites",
                "config": {"workload": desc, "fs_hz": fs_nom, "channels": 8, "blocks_per_step": 99},
                "cpu_baseline": dict(vals[-1], value=val),
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
            return
    # bounded sample per step so the whole run ends within minutes
    recs1 = U.synthetic_recs_fast(2, n_chan, fs, seed=5)
    t0 = time.perf_counter()
    U.oracle_synth(fs, n_samp, recs1[:1], threads=threads)
    per_epoch = max(time.perf_counter() - t0, 1e-3)
    budget = 150.0 / max(args.steps + args.warmup, 1)
    n_ep = int(min(max(budget / per_epoch, 1), n_epochs))
    recs = U.synthetic_recs_fast(n_ep, n_chan, fs, seed=6)
    if args.workload == "rt1":
        recs[:, 8:]["prn"] = 0                 # the same 8 of 16 slots in use as in the B200 arm
    ph = None
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, ph = U.oracle_synth(fs, n_samp, recs, ph, threads=threads)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    val = n_ep * n_samp / (ms * 1e-3) / 1e6
    sample = f"{n_ep} of {n_epochs} blocks per step ({n_ep * n_samp / fs_nom:.1f} s of signal), oracle/e1_oracle.c, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 phase + int32 accumulate", "data": "synthetic",
        "config": {"workload": desc, "fs_hz": fs_nom, "channels": n_chan, "blocks_per_step": n_ep},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed output (profiling runs)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return

    # stdout carries exactly one JSON line: anything a library prints there while we run (NCCL's version
    # banner does, whatever NCCL_DEBUG_FILE says) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import e1b200 as E
    import e1util as U

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank)

    fs_nom, n_samp, n_chan, n_epochs, desc = wl
    fs = U.fs_as_reference(fs_nom)
    data_desc = "synthetic"
    recs = None
    if args.workload == "cfg1" and NAV_SUBSET.exists():
        try:                                                                 # the reference's own case from its own navigation data
            import build as B
            B.build_host()
            import e1host as H
            n_chan = 16                                                      # MAX_CHAN of the reference build
            recs, _ = H.Scenario(NAV_SUBSET, llh=(-6, 51, 100), duration_s=10, max_chan=n_chan).all()
            recs = np.ascontiguousarray(recs.astype(U.REC_DTYPE))
            data_desc = "real: RINEX week171 subset, -l -6,51,100, 8 satellites (records by galileo-sdr-sim_b200/host)"
        except Exception as ex:
            print(f"bench: host records unavailable ({ex}); synthetic records instead", file=sys.stderr)
            recs = None
    use_ranges = args.workload.startswith("cfg4")
    rec_dtype = E.RANGE_DTYPE if use_ranges else U.REC_DTYPE
    if use_ranges:
        recs = synthetic_range_records(n_epochs, n_chan, E.RANGE_DTYPE, seed=1000 + rank)
    if recs is None:
        recs = U.synthetic_recs_fast(n_epochs, n_chan, fs, seed=1000 + rank)     # each rank: its own time shard
        if args.workload == "rt1":
            recs[:, 8:]["prn"] = 0                                               # 8 of the 16 slots in use, like BASELINE configs[0]
    rec_bytes = recs.nbytes
    out_bytes = n_epochs * n_samp * 4
    samples_per_step = n_epochs * n_samp

    synth = E.Synth(fs, n_samp, n_chan, device=local_rank)                    # raises without the CUDA library
    st = synth.stats()
    ext = torch.cuda.ExternalStream(synth.stream(), device=dev)
    d_recs = torch.from_numpy(recs.view(np.uint8).reshape(-1)).to(dev)
    d_out = torch.empty(out_bytes // 2, dtype=torch.int16, device=dev)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm -------------------------------------------------------------
    dev_call = synth.synth_ranges_device if use_ranges else synth.synth_epochs_device
    host_call = synth.synth_ranges if use_ranges else synth.synth_epochs
    for _ in range(args.warmup):
        dev_call(n_epochs, d_recs.data_ptr(), d_out.data_ptr())
        synth.sync()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    synth_ms, plan_ms, launches, synth_launches = 0.0, 0.0, 0, 0
    ev0.record(ext)
    for _ in range(args.steps):
        dev_call(n_epochs, d_recs.data_ptr(), d_out.data_ptr())
        synth.sync()                      # also collects the per-kernel CUDA-event times of this step
        t = synth.timing()
        synth_ms += t.synth_ms
        plan_ms += t.plan_ms
        launches += t.kernel_launches
        synth_launches += t.synth_launches
    ev1.record(ext)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    value = world * samples_per_step / (dev_ms * 1e-3) / 1e6

    # ---- end-to-end arm: C-ABI host entry, pinned host buffers ---------------------------
    e2e = None
    if not args.no_e2e:
        h_recs = E.PinnedBuffer(rec_bytes)
        h_recs.u8[:] = recs.view(np.uint8).reshape(-1)
        h_out = E.PinnedBuffer(out_bytes)
        recs_view = h_recs.view(rec_dtype).reshape(n_epochs, n_chan)
        out_view = h_out.view(np.int16)
        for _ in range(max(1, min(args.warmup, 2))):
            host_call(recs_view, out_view)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            host_call(recs_view, out_view)                # returns after the last D2H completed
        torch.cuda.synchronize()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
        barrier()
        e2e = {"value": world * samples_per_step / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": rec_bytes, "d2h_bytes_per_step": out_bytes,
               "api": ("e1b200_synth_ranges" if use_ranges else "e1b200_synth_epochs") + " (host buffers, pinned), timed on the host clock around the call",
               "bound": f"PCIe D2H: {out_bytes / 1e9:.2f} GB per step leave the GPU at {out_bytes / (e2e_ms * 1e-3) / 1e9:.1f} GB/s (plain pinned copies on these boxes: "
                        "55 GB/s for one GPU, 118 GB/s for all eight together: profiles/r1_d2h_diag_8gpu.txt); the kernels take a quarter of the call"}
        h_recs.free(), h_out.free()

    # ---- strong-scaling arm (N > 1): ONE scenario of this size, time axis sharded over the N GPUs ----------------
    # rank 0's scenario (seed 1000) split into contiguous block ranges (shard.split_epochs).  Per step, timed:
    #   hand-off  every rank runs the carrier planner over all blocks BEFORE its range (shard.replan_start_phases_device: no
    #             communication, no serial chain; the alternative, a plan-and-send chain over NCCL, is timed as handoff_chain_ms)
    #   synthesis of the rank's range, its kernel storing straight into rank 0's stream buffer over NVLink (peer memory,
    #             e1b200_peer_*): compute and gather are one kernel, there is no gather step
    #   (for comparison) the same with a local buffer and an NCCL point-to-point gather of the device-resident segments
    # The stream that ends up in rank 0's HBM must equal rank 0's own single-GPU stream.
    strong = None
    if dist is not None and not use_ranges:
        import shard as S
        recs_all = recs if rank == 0 else U.synthetic_recs_fast(n_epochs, n_chan, fs, seed=1000)
        ranges = S.split_epochs(n_epochs, world)
        lo, hi = ranges[rank]
        d_seg_recs = torch.from_numpy(np.ascontiguousarray(recs_all[lo:hi]).view(np.uint8).reshape(-1)).to(dev)
        d_recs_all = d_recs if rank == 0 else torch.from_numpy(recs_all.view(np.uint8).reshape(-1)).to(dev)   # the whole scenario's records, resident
        handle = torch.zeros(64, dtype=torch.uint8, device=dev)
        full, peer_err = None, ""
        if rank == 0:
            try:
                full = E.PeerBuffer.alloc(local_rank, out_bytes)
                handle.copy_(torch.frombuffer(bytearray(full.handle), dtype=torch.uint8))
            except Exception as ex:
                peer_err = str(ex)
        dist.broadcast(handle, 0)
        if rank != 0:
            try:
                full = E.PeerBuffer.open(local_rank, bytes(handle.cpu().numpy().tobytes()), out_bytes)
            except Exception as ex:
                peer_err = str(ex)
        ok = torch.tensor([0.0 if peer_err else 1.0], dtype=torch.float64, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        peer_ok = bool(ok.item() > 0)
        if not peer_ok:          # e.g. ranks in separate containers (no CUDA IPC), GPUs that are not peers: the weak line still stands
            if full is not None:
                full.close()
            strong = {"scaling": "strong", "unavailable": "peer memory (CUDA IPC) between the ranks' GPUs: " + (peer_err or "another rank failed to open the handle")}
    if dist is not None and not use_ranges and strong is None:
        seg_engine = E.Synth(fs, n_samp, n_chan, device=local_rank)
        d_seg = torch.empty((hi - lo) * n_samp, dtype=torch.int32, device=dev)          # NCCL variant: one (I, Q) pair = one int32
        d_full = torch.empty(n_epochs * n_samp, dtype=torch.int32, device=dev) if rank == 0 else None

        h_seg_recs = E.PinnedBuffer(max(recs_all[lo:hi].nbytes, 1))
        h_seg_recs.u8[:recs_all[lo:hi].nbytes] = np.ascontiguousarray(recs_all[lo:hi]).view(np.uint8).reshape(-1)
        seg_recs_view = h_seg_recs.view(U.REC_DTYPE)[: (hi - lo) * n_chan].reshape(hi - lo, n_chan)

        def strong_step(fused, chain=False):
            t = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            t[0].record()
            if chain:
                S.handoff_start_phases(seg_engine, recs_all, rank, world, dist, dist_device=dev)
            else:
                S.replan_start_phases_device(seg_engine, d_recs_all.data_ptr(), n_epochs, rank, world)
                seg_engine.sync()                # only so that the hand-off can be timed apart from the synthesis
            t[1].record()
            if fused == "dma":       # slices by the copy engines over NVLink, behind the kernels (records from pinned host memory)
                seg_engine.synth_epochs_to(seg_recs_view, full.ptr + lo * n_samp * 4)
            elif fused:              # the kernel's own stores cross NVLink
                seg_engine.synth_epochs_device(hi - lo, d_seg_recs.data_ptr(), full.ptr + lo * n_samp * 4)
                seg_engine.sync()
            else:
                seg_engine.synth_epochs_device(hi - lo, d_seg_recs.data_ptr(), d_seg.data_ptr())
                seg_engine.sync()
            t[2].record()
            if fused:
                dist.barrier()                   # every rank's stores have landed in rank 0's buffer
            else:
                S.gather_segments_device(d_seg, rank, world, dist, n_epochs, n_samp, d_full)
            t[3].record()
            torch.cuda.synchronize()
            return [t[i].elapsed_time(t[i + 1]) for i in range(3)]

        def timed(fused, chain=False):
            for _ in range(max(args.warmup, 1)):
                strong_step(fused, chain)
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            parts = np.zeros(3)
            for _ in range(args.steps):
                parts += np.array(strong_step(fused, chain))
            s1.record()
            barrier()
            return max_over_ranks(s0.elapsed_time(s1)) / args.steps, [max_over_ranks(float(v)) / args.steps for v in parts]

        ms_d, parts_d = timed("dma")
        equal_d = None
        if rank == 0:
            import ctypes as C
            got = torch.empty(n_epochs * n_samp, dtype=torch.int32, device=dev)
            C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(got.data_ptr()), C.c_void_p(full.ptr), C.c_size_t(out_bytes), 3)   # device to device
            equal_d = bool(torch.equal(got, d_out.view(torch.int32)))
            C.CDLL("libcudart.so").cudaMemset(C.c_void_p(full.ptr), 0, C.c_size_t(out_bytes))
            del got
        ms_f, parts_f = timed(True)
        equal = None
        if rank == 0:
            got = torch.empty(n_epochs * n_samp, dtype=torch.int32, device=dev)
            import ctypes as C
            C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(got.data_ptr()), C.c_void_p(full.ptr), C.c_size_t(out_bytes), 3)   # device to device
            equal = bool(torch.equal(got, d_out.view(torch.int32)))
            del got
        ms_n, parts_n = timed(False)
        equal_n = bool(torch.equal(d_full, d_out.view(torch.int32))) if rank == 0 else None
        ms_c, parts_c = timed(True, chain=True)
        strong = {"scaling": "strong", "value": samples_per_step / (ms_d * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_d,
                  "handoff_ms": parts_d[0], "synth_and_gather_ms": parts_d[1], "barrier_ms": parts_d[2],
                  "what": f"ONE scenario ({n_epochs} blocks) split into {world} contiguous block ranges: every rank plans the carrier over the "
                          "blocks before its range (no communication), then synthesises its range through e1b200_synth_epochs with rank 0's "
                          "stream buffer (peer memory, e1b200_peer_*) as destination: every finished 96 MB slice travels over NVLink by the "
                          "copy engines while the SMs synthesise the next one -- no collective, no gather step after the kernels; "
                          "CUDA events, max over ranks",
                  "stream_in_rank0_hbm_equals_single_gpu_stream": equal_d,
                  "kernel_stores_variant": {"value": samples_per_step / (ms_f * 1e-3) / 1e6, "ms_per_step": ms_f, "handoff_ms": parts_f[0],
                                            "synth_ms": parts_f[1], "equals_single_gpu_stream": equal,
                                            "what": "the synthesis kernel's own 128-bit stores go to rank 0's buffer over NVLink (fused compute + gather)"},
                  "nccl_gather_variant": {"value": samples_per_step / (ms_n * 1e-3) / 1e6, "ms_per_step": ms_n, "handoff_ms": parts_n[0],
                                          "synth_ms": parts_n[1], "gather_ms": parts_n[2], "equals_single_gpu_stream": equal_n,
                                          "what": "local segment buffers + NCCL isend/irecv of the device-resident segments into rank 0's HBM"},
                  "handoff_chain_variant": {"ms_per_step": ms_c, "handoff_chain_ms": parts_c[0],
                                            "what": "plan-and-send chain (carrier planner on the own range, phases to the next rank over NCCL) instead of re-planning"}}
        seg_engine.close()
        barrier()
        full.close()
        h_seg_recs.free()
        del d_seg, d_full, d_seg_recs, d_recs_all

    # ---- the timed bytes against the oracle (after both timed regions) ---------------------
    parity = None
    if not args.no_parity:
        unbind_cpus()
        parity = parity_check(d_out, recs, fs, n_samp, n_chan, world, use_ranges)
        if dist is not None:                                            # every rank checked its own segment
            t = torch.tensor([parity["blocks"], parity["samples_compared"], parity["differing_samples"]], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            parity.update(blocks=int(t[0].item()), samples_compared=int(t[1].item()), differing_samples=int(t[2].item()),
                          of=n_epochs * world, what=parity["what"] + f", summed over {world} ranks")

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        per_launch_ms = synth_ms / max(synth_launches, 1)
        bytes_per_launch = out_bytes * args.steps / max(synth_launches, 1)
        achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
        tr = ncu_summary_numbers() if args.workload == "cfg2" else None
        if args.workload in ("cfg3", "cfg3s", "cfg4s", "cfg5s"):            # captured on the 30 s slice (750 M samples per launch): per-sample figures scale
            t3 = ncu_summary_numbers("cfg3")
            if t3:
                k = samples_per_step / (300 * 2500000)
                tr = (t3[0] * k, t3[1], t3[2] * k if t3[2] else None)
        traffic, traffic_src = (tr[0], f"profiles/{tr[1]} (ncu --set full, one launch" + (" of this workload)" if args.workload == "cfg2" else " of the 30 s slice, scaled by samples)")) if tr else (None, None)
        # what actually bounds the kernel: warp-instruction issue slots (4 schedulers per SM, one instruction
        # per clock each).  Instructions per launch from the committed ncu capture of this workload, duration live.
        issue = None
        if tr and tr[2] and clocks and clocks.get("sm_mhz"):
            peak_issue = st.sm_count * 4 * clocks["sm_mhz"] * 1e6
            ach_issue = tr[2] / (per_launch_ms * 1e-3)
            issue = {"bound": "issue slots (integer pipes; no tensor-core or HBM-bound formulation of this loop exists)",
                     "achieved": ach_issue / 1e9, "peak": peak_issue / 1e9, "unit": "G warp-inst/s", "frac": ach_issue / peak_issue,
                     "warp_inst_per_launch": tr[2], "inst_per_channel_sample": tr[2] * 32 / (samples_per_step * n_chan),
                     "source": f"smsp__inst_executed.sum in profiles/{tr[1]}; peak = {st.sm_count} SMs x 4 schedulers x SM clock under load",
                     "peak_source": "one warp instruction per clock and scheduler, measured on a B200 with tools/ubench/pipe_rates.cu "
                                    "(profiles/r1_pipe_rates_b200.txt: add.u64 = IADD3 + IADD3.X pairs 0.998, IMAD | SHF pairs 0.969 warp-inst/clk/SMSP; "
                                    "a single integer pipe -- IMAD, or IADD3 / LOP3 / SHF -- takes 0.50)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 phase + int32 accumulate", "data": data_desc,
            "config": {"workload": desc, "fs_hz": fs_nom, "channels": n_chan, "blocks_per_step": n_epochs,
                       "samples_per_step_per_gpu": samples_per_step, "tile": st.tile, "ctas_per_sm": st.ctas_per_sm,
                       "l2": f"each step writes {out_bytes / 1e9:.2f} GB per GPU (> 126 MB L2), no flush needed",
                       "shard": "time axis, one contiguous segment per rank, no data-path collective",
                       "host_numa_node": numa_node},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                         "kernel": st.kernel_name, "peak_source": peak_src,
                         "ms_per_launch": per_launch_ms, "launches_per_step": synth_launches / args.steps,
                         "planner_ms_per_step": plan_ms / args.steps, "synth_ms_per_step": synth_ms / args.steps,
                         "note": f"issue-bound, not HBM-bound: {n_chan} channel visits x ~{(issue or {}).get('inst_per_channel_sample', 12.3):.1f} integer instructions per 4-byte sample "
                                 "(see roofline_issue; ncu: profiles/*synth_ncu_summary.txt); HBM time of the same bytes "
                                 f"would be {bytes_per_launch / peak / 1e6:.2f} ms"},
            "roofline_issue": issue,
            "clocks": clocks, "gpu_launches": launches,
            "exact_fallback_samples": int(synth.stats().exact_samples),
            "planner": {"hat_epochs": int(synth.stats().hat_epochs), "serial_epochs": int(synth.stats().serial_epochs),
                        "note": "cumulative channel-epochs planned in parallel vs walked serially by the chain"},
        }
        if e2e:
            line["e2e"] = e2e
        line["parity_check"] = parity
        if strong:
            line["strong"] = strong
        if not args.no_cpu_baseline and world == 1:
            unbind_cpus()
            ref = reference_binary_cfg1() if args.workload == "cfg1" else None
            line["cpu_baseline"] = ref or cpu_baseline(fs, n_samp, n_chan)
            if args.workload != "cfg1":
                rb = reference_binary_cfg1()
                if rb:
                    line["reference_binary_configs0"] = rb      # the real executable on its own runnable case, for scale
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    synth.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
