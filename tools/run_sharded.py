#!/usr/bin/env python3
"""One scenario, time axis sharded over the GPUs of a node (BASELINE config 5 shape).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/run_sharded.py \
        [--fs 25e6] [--channels 36] [--seconds 60] [--sink pwrite|gather|peer|none] [--out /dev/shm/e1.ishort]

Every rank builds the same synthetic scenario (seeded), takes its contiguous block range, gets its
start phases through the hand-off chain (shard.py), synthesises through the C-ABI host entry point
(H2D + kernels + D2H into pinned memory) and either pwrite()s its byte range of the ishort file or
sends it to rank 0 over NCCL.  --sink peer is the single-writer path without a host round trip per segment: every
rank plans the carrier up to its range from its resident copy of the records (no chain), synthesises with rank 0's
stream buffer in HBM as destination (e1b200_peer_*: slices travel over NVLink behind the kernels), rank 0 then copies
the whole stream to pinned host memory once and writes the file with several pwrite threads.  Rank 0 prints one JSON line: whole-job Msamples/s (max over ranks of
the wall time between two barriers), the hand-off time and, with --check, whether the md5 of the file
equals a single-GPU run of the same scenario.
"""
import argparse
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fs", type=float, default=25e6)
    ap.add_argument("--channels", type=int, default=36)
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--sink", default="pwrite", choices=["pwrite", "gather", "peer", "none"])
    ap.add_argument("--writers", type=int, default=8, help="--sink peer: pwrite threads of the single writer")
    ap.add_argument("--out", default="/dev/shm/e1b200_sharded.ishort")
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import e1b200 as E
    import e1util as U
    import shard as S

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)

    fs = U.fs_as_reference(args.fs)
    n_samp = int(round(args.fs / 10))
    n_epochs = int(args.seconds * 10) - 1
    recs = U.synthetic_recs_fast(n_epochs, args.channels, fs, seed=42)
    lo, hi = S.split_epochs(n_epochs, world)[rank]
    eng = E.Synth(fs, n_samp, args.channels, device=local)
    h_out = E.PinnedBuffer(max(hi - lo, 1) * n_samp * 4)
    out_view = h_out.view(np.int16)[: (hi - lo) * n_samp * 2].reshape(-1, 2)
    eng.synth_epochs(recs[lo:min(hi, lo + 8)], out_view[: (min(hi, lo + 8) - lo) * n_samp])      # warm-up: allocations, tables
    if world > 1:                                                                                  # ... and NCCL's lazy P2P set-up
        S.handoff_start_phases(eng, recs[: 2 * world], rank, world, dist, dist_device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    full = None
    extra = {}
    if args.sink == "peer":
        import ctypes as C
        from concurrent.futures import ThreadPoolExecutor
        total = n_epochs * n_samp * 4
        handle = torch.zeros(64, dtype=torch.uint8, device=dev)
        if rank == 0:
            full = E.PeerBuffer.alloc(local, total)
            handle.copy_(torch.frombuffer(bytearray(full.handle), dtype=torch.uint8))
        if world > 1:
            dist.broadcast(handle, 0)
        if rank != 0:
            full = E.PeerBuffer.open(local, bytes(handle.cpu().numpy().tobytes()), total)
        d_recs = torch.from_numpy(recs.view(np.uint8).reshape(-1)).to(dev)
        h_all = E.PinnedBuffer(total) if rank == 0 else None
        if rank == 0:
            h_all.u8[::4096] = 0                                             # touch the pages before the clock starts
            with open(args.out, "wb") as f:
                f.truncate(total)
        barrier()
        t0 = time.perf_counter()
        S.replan_start_phases_device(eng, d_recs.data_ptr(), n_epochs, rank, world)
        eng.sync()
        t_hand = time.perf_counter() - t0
        if hi > lo:
            eng.synth_epochs_to(recs[lo:hi], full.ptr + lo * n_samp * 4)    # returns when the last slice has landed in rank 0's HBM
        barrier()
        t_synth = time.perf_counter() - t0
        if rank == 0:
            rt = C.CDLL("libcudart.so")
            assert rt.cudaMemcpy(C.c_void_p(h_all.ptr), C.c_void_p(full.ptr), C.c_size_t(total), 2) == 0    # one D2H of the whole stream
            extra["d2h_s"] = time.perf_counter() - t0 - t_synth
            buf = memoryview(h_all.u8)
            fd = os.open(args.out, os.O_WRONLY)
            step = -(-total // args.writers)
            with ThreadPoolExecutor(args.writers) as ex:
                list(ex.map(lambda o: os.pwrite(fd, buf[o:min(o + step, total)], o), range(0, total, step)))
            os.close(fd)
            del buf
            extra["write_s"] = time.perf_counter() - t0 - t_synth - extra["d2h_s"]
        barrier()
        dt = time.perf_counter() - t0
        if rank == 0:
            h_all.free()
        full.close()
        whole = None
    else:
        barrier()
        t0 = time.perf_counter()
        S.handoff_start_phases(eng, recs, rank, world, dist if world > 1 else None, dist_device=dev)
        t_hand = time.perf_counter() - t0
        seg = eng.synth_epochs(recs[lo:hi], out_view)
        t_synth = time.perf_counter() - t0
        whole = None
        if args.sink == "pwrite":
            S.write_segment(args.out, lo, n_samp, seg, total_epochs=n_epochs)
        elif args.sink == "gather" and world > 1:
            whole = S.gather_segments(seg, rank, world, dist, n_epochs, n_samp, dist_device=dev)
        barrier()
        dt = time.perf_counter() - t0
    times = torch.tensor([dt, t_hand, t_synth], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    if args.sink == "gather" and rank == 0:
        (whole if whole is not None else seg).tofile(args.out)
    if rank == 0:
        line = {"tool": "run_sharded", "n_gpus": world, "fs_hz": args.fs, "channels": args.channels, "blocks": n_epochs,
                "sink": args.sink, "Msamples_per_s": n_epochs * n_samp / float(times[0]) / 1e6, "wall_s": float(times[0]),
                "handoff_s_max": float(times[1]), "synth_s_max": float(times[2]), "bytes": n_epochs * n_samp * 4, **extra}
        if args.check and args.sink != "none":
            ref = E.Synth(fs, n_samp, args.channels, device=local)
            md5 = hashlib.md5()
            with open(args.out, "rb") as f:
                for e0 in range(0, n_epochs, 64):
                    exp = ref.synth_epochs(recs[e0:e0 + 64])
                    got = np.frombuffer(f.read(exp.nbytes), dtype=np.int16).reshape(-1, 2)
                    if not np.array_equal(exp, got):
                        line["check"] = f"MISMATCH in blocks {e0}..{e0 + 63}"
                        break
                    md5.update(exp.tobytes())
                else:
                    line["check"] = "file equals the single-GPU stream"
                    line["md5"] = md5.hexdigest()
            ref.close()
        print(json.dumps(line))
    eng.close()
    h_out.free()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
