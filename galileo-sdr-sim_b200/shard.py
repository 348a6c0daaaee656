"""Time-axis sharding of one scenario over several GPUs (BASELINE config 5).

One process per GPU.  A scenario is `recs[n_epochs][max_chan]` (e1_epoch_rec); rank r owns the
contiguous block range split_epochs(n_epochs, world)[r].  Blocks are independent except for the
carrier phase, which the reference integrates through the whole run
(src/galileo-sdr.cpp:531-532, state `chan[i].carr_phase`), so the only thing ranks exchange before
they synthesise is that phase:

    rank r:  start <- recv(rank r-1)          (rank 0: the scenario's initial phases)
             end    = plan_phases(own blocks)  (carrier planner only, milliseconds)
             send(end) -> rank r+1
             synthesise own blocks from `start`

Output segments are contiguous byte ranges of the reference's headerless ishort file
(src/galileo-sdr.cpp:536-542): sample k of block b sits at byte 4*(b*N + k).  Two sinks:
  write_segment()   every rank pwrite()s its own range -- no collective, each GPU uses its own
                    PCIe link (the default; the per-GPU stream is a few GB/s, far below NVLink)
  gather_segments_device()  the north star's single-writer variant: the device-resident segments travel to
                    rank 0's HBM over NCCL (NVLink) point-to-point, all transfers posted at once; rank 0 then
                    holds the whole stream on the device (one D2H, or none)
  gather_segments() the same through host arrays (gloo in the CPU tests; no GPU path uses it)

`engine` is anything with set_carrier_phases / plan_phases / synth_epochs / carrier_phases and
max_chan: the product passes e1b200.Synth (CUDA, no CPU fallback); the GPU-less tests pass a stub
so the split / hand-off / offset logic is covered without a device.
"""
import os

import numpy as np


def split_epochs(n_epochs, world):
    """Contiguous block ranges, sizes differing by at most one: [(lo, hi), ...] of length world."""
    if world < 1 or n_epochs < 0:
        raise ValueError("world >= 1 and n_epochs >= 0 required")
    base, extra = divmod(n_epochs, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def _phase_tensor(engine, dist_device):
    import torch
    return torch.zeros(engine.max_chan, dtype=torch.float64, device=dist_device)


def handoff_start_phases(engine, recs, rank, world, dist=None, phases0=None, dist_device="cpu"):
    """The phase hand-off chain.  Returns (lo, hi, start_phases) for this rank; afterwards the engine
    holds start_phases again, ready to synthesise recs[lo:hi]."""
    lo, hi = split_epochs(recs.shape[0], world)[rank]
    if rank == 0:
        start = np.zeros(engine.max_chan) if phases0 is None else np.asarray(phases0, dtype=np.float64).copy()
    else:
        t = _phase_tensor(engine, dist_device)
        dist.recv(t, src=rank - 1)
        start = t.cpu().numpy().copy()
    if rank + 1 < world:
        engine.set_carrier_phases(start)
        end = engine.plan_phases(recs[lo:hi])
        import torch
        t = torch.from_numpy(np.ascontiguousarray(end, dtype=np.float64)).to(dist_device)
        dist.send(t, dst=rank + 1)
    engine.set_carrier_phases(start)
    return lo, hi, start


def replan_start_phases(engine, recs, rank, world, phases0=None):
    """The hand-off without a chain: every rank runs the carrier planner over ALL blocks before its range, in
    parallel with the others -- redundant work (rank r plans lo_r blocks, the last rank nearly the whole scenario,
    ~1 us per channel-block) instead of world - 1 serial plan-and-send hops, each with the fixed cost of a planner
    pass.  No communication.  Returns (lo, hi, start_phases); the engine holds start_phases."""
    lo, hi = split_epochs(recs.shape[0], world)[rank]
    start = np.zeros(engine.max_chan) if phases0 is None else np.asarray(phases0, dtype=np.float64).copy()
    engine.set_carrier_phases(start)
    if lo > 0:
        start = engine.plan_phases(recs[:lo])
    return lo, hi, start


def replan_start_phases_device(engine, d_recs_ptr, n_epochs, rank, world, phases0=None):
    """replan_start_phases with the scenario's records already on the rank's GPU (d_recs_ptr: device address of
    e1_epoch_rec[n_epochs][max_chan]): the carrier-only planner runs over the blocks before the range straight from
    there (e1b200_plan_phases_device) -- no H2D of up to the whole record set per step, no read-back of the phases:
    they stay in the context, and the synthesis call that follows on the same stream starts from them.  Asynchronous;
    returns (lo, hi)."""
    lo, hi = split_epochs(n_epochs, world)[rank]
    engine.set_carrier_phases(np.zeros(engine.max_chan) if phases0 is None else np.asarray(phases0, dtype=np.float64))
    if lo > 0:
        engine.plan_phases_device(lo, d_recs_ptr)
    return lo, hi


def synth_shard(engine, recs, rank, world, dist=None, phases0=None, dist_device="cpu", out=None):
    """This rank's segment of the scenario: (lo, hi, int16 [ (hi-lo)*N, 2 ])."""
    lo, hi, _ = handoff_start_phases(engine, recs, rank, world, dist, phases0, dist_device)
    seg = engine.synth_epochs(recs[lo:hi], out) if hi > lo else np.zeros((0, 2), np.int16)
    return lo, hi, seg


def write_segment(path, lo, n_samp, samples, total_epochs=None):
    """pwrite this rank's samples at their place in the ishort file (4 bytes per sample)."""
    fd = os.open(path, os.O_WRONLY | os.O_CREAT, 0o644)
    try:
        if total_epochs is not None:
            os.ftruncate(fd, max(os.fstat(fd).st_size, total_epochs * n_samp * 4))
        buf = memoryview(np.ascontiguousarray(samples, dtype=np.int16).reshape(-1).view(np.uint8))
        off, done = lo * n_samp * 4, 0
        while done < len(buf):
            done += os.pwrite(fd, buf[done:done + (1 << 30)], off + done)
    finally:
        os.close(fd)


def gather_segments(seg, rank, world, dist, n_epochs, n_samp, dist_device="cpu"):
    """Single-writer variant: every rank's segment travels to rank 0 (point-to-point, one message per
    rank; the sizes are known from split_epochs).  Returns the whole stream on rank 0, None elsewhere."""
    import torch
    ranges = split_epochs(n_epochs, world)
    # NCCL has no int16: one (I, Q) pair travels as one int32
    if rank != 0:
        if seg.shape[0]:
            t = torch.from_numpy(np.ascontiguousarray(seg).view(np.int32).reshape(-1))
            dist.send(t.to(dist_device), dst=0)
        return None
    out = np.empty((n_epochs * n_samp, 2), np.int16)
    out[: seg.shape[0]] = seg
    for r in range(1, world):
        lo, hi = ranges[r]
        if hi == lo:
            continue
        t = torch.empty((hi - lo) * n_samp, dtype=torch.int32, device=dist_device)
        dist.recv(t, src=r)
        out[lo * n_samp:hi * n_samp] = t.cpu().numpy().view(np.int16).reshape(-1, 2)
    return out


def gather_segments_device(d_seg, rank, world, dist, n_epochs, n_samp, d_full=None):
    """Device-resident gather over NCCL: d_seg is this rank's segment as a 1-D int32 CUDA tensor (one (I, Q) pair
    per element: NCCL has no int16), d_full rank 0's receive buffer of n_epochs * n_samp int32 (allocated if
    None).  Every rank's isend / rank 0's irecvs are posted in one batch, so the N - 1 transfers share NVLink
    instead of queueing behind each other.  Returns d_full on rank 0, None elsewhere."""
    import torch
    ranges = split_epochs(n_epochs, world)
    ops = []
    if rank == 0:
        if d_full is None:
            d_full = torch.empty(n_epochs * n_samp, dtype=torch.int32, device=d_seg.device)
        lo, hi = ranges[0]
        d_full[lo * n_samp:hi * n_samp].copy_(d_seg[: (hi - lo) * n_samp])
        for r in range(1, world):
            lo, hi = ranges[r]
            if hi > lo:
                ops.append(dist.P2POp(dist.irecv, d_full[lo * n_samp:hi * n_samp], r))
    else:
        lo, hi = ranges[rank]
        if hi > lo:
            ops.append(dist.P2POp(dist.isend, d_seg[: (hi - lo) * n_samp], 0))
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
    return d_full if rank == 0 else None
