"""Time-axis sharding (galileo-sdr-sim_b200/shard.py): split, carrier-phase hand-off between ranks,
segment placement in the ishort file, gather to a single writer.  The CPU tests run two/three
`gloo` ranks with the oracle standing in for the engine (host logic only); the GPU test runs the
same hand-off with two real e1b200.Synth contexts."""
import hashlib
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

import e1util as U

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "galileo-sdr-sim_b200"))
import shard as S  # noqa: E402

FS = U.fs_as_reference(2.6e6)


def test_split_epochs_properties():
    for n in (0, 1, 2, 7, 99, 2999, 35999):
        for w in (1, 2, 3, 4, 8):
            r = S.split_epochs(n, w)
            assert len(r) == w and r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        S.split_epochs(5, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scenario(n_epochs, n_chan, max_chan, seed):
    recs = U.synthetic_recs(n_epochs, n_chan, FS, seed=seed, max_chan=max_chan)
    if n_epochs > 3:                       # a slot that is re-allocated inside the second shard, and an idle stretch
        recs[n_epochs - 2, 0]["flags"] = U.E1_REC_SET_PHASE
        recs[n_epochs - 2, 0]["carr_phase_init"] = 0.4321
        recs[1:3, 1]["prn"] = 0
    return recs


def _rank_main(rank, world, port, n_samp, n_epochs, n_chan, max_chan, seed, path, mode, use_gpu):
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        recs = _scenario(n_epochs, n_chan, max_chan, seed)
        if use_gpu:
            import e1b200 as E
            eng = E.Synth(FS, n_samp, max_chan, device=0)
        else:
            eng = U.OracleEngine(FS, n_samp, max_chan)
        phases0 = np.linspace(0.05, 0.6, max_chan)
        if mode == "peer":
            # the bench's strong-scaling arm in small: no hand-off chain (every rank re-plans the blocks before its
            # range), the finished segment copied by the library into the writer rank's device buffer (CUDA IPC)
            import e1b200 as E
            nbytes = n_epochs * n_samp * 4
            h = torch.zeros(64, dtype=torch.uint8)
            full = None
            if rank == 0:
                full = E.PeerBuffer.alloc(0, nbytes)
                h.copy_(torch.frombuffer(bytearray(full.handle), dtype=torch.uint8))
            dist.broadcast(h, 0)
            if rank != 0:
                full = E.PeerBuffer.open(0, h.numpy().tobytes(), nbytes)
            d_recs = torch.from_numpy(recs.view(np.uint8).reshape(-1)).cuda()       # the scenario's records, resident
            lo, hi = S.replan_start_phases_device(eng, d_recs.data_ptr(), n_epochs, rank, world, phases0=phases0)
            if hi > lo:
                eng.synth_epochs_to(recs[lo:hi], full.ptr + lo * n_samp * 4)
            dist.barrier()
            if rank == 0:
                import ctypes as C
                whole = np.empty((n_epochs * n_samp, 2), dtype=np.int16)
                assert C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(whole.ctypes.data), C.c_void_p(full.ptr), C.c_size_t(nbytes), 2) == 0
                whole.tofile(path)
            dist.barrier()
            full.close()
            eng.close()
            return
        if mode == "replan":
            lo, hi, _ = S.replan_start_phases(eng, recs, rank, world, phases0=phases0)
            seg = eng.synth_epochs(recs[lo:hi]) if hi > lo else np.zeros((0, 2), dtype=np.int16)
            mode = "pwrite"
        else:
            lo, hi, seg = S.synth_shard(eng, recs, rank, world, dist, phases0=phases0)
        assert seg.shape[0] == (hi - lo) * n_samp
        if mode == "pwrite":
            S.write_segment(path, lo, n_samp, seg, total_epochs=n_epochs)
        else:
            whole = S.gather_segments(seg, rank, world, dist, n_epochs, n_samp)
            if rank == 0:
                whole.tofile(path)
            else:
                assert whole is None
        dist.barrier()
        if use_gpu:
            eng.close()
    finally:
        dist.destroy_process_group()


def _run(world, n_samp, n_epochs, n_chan, max_chan, seed, tmp_path, mode, use_gpu=False):
    import torch.multiprocessing as mp
    path = str(tmp_path / f"out_{mode}_{world}.ishort")
    mp.spawn(_rank_main, args=(world, _free_port(), n_samp, n_epochs, n_chan, max_chan, seed, path, mode, use_gpu),
             nprocs=world, join=True)
    got = np.fromfile(path, dtype=np.int16).reshape(-1, 2)
    ref, _ = U.oracle_synth(FS, n_samp, _scenario(n_epochs, n_chan, max_chan, seed), np.linspace(0.05, 0.6, max_chan))
    assert got.shape == ref.shape
    assert hashlib.sha256(got.tobytes()).hexdigest() == hashlib.sha256(ref.tobytes()).hexdigest()


@pytest.mark.parametrize("mode", ["pwrite", "gather"])
def test_two_ranks_gloo_equal_single_run(tmp_path, mode):
    """world_size 2 over gloo: the two segments, each started from the phase handed over by the planner
    pass of the rank before it, form exactly the single-process stream (odd block count -> ragged split)."""
    _run(2, 13000, 7, 5, 6, 11, tmp_path, mode)


def test_three_ranks_with_an_empty_shard(tmp_path):
    _run(3, 5200, 2, 3, 4, 5, tmp_path, "pwrite")


def test_three_ranks_replan_instead_of_chain(tmp_path):
    """shard.replan_start_phases: every rank plans the carrier over all blocks before its range (no communication);
    same stream as the plan-and-send chain and as the single run."""
    _run(3, 5200, 8, 4, 5, 7, tmp_path, "replan")


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["pwrite", "gather"])
def test_two_ranks_cuda_engine(tmp_path, mode):
    """Same hand-off with two real synthesiser contexts (both on cuda:0; gloo carries the phases):
    e1b200_plan_phases + e1b200_set_carrier_phase reproduce the single-run stream bit for bit."""
    _run(2, 52000, 9, 7, 8, 3, tmp_path, mode, use_gpu=True)


@pytest.mark.gpu
def test_two_ranks_peer_buffer_sink(tmp_path):
    """e1b200_peer_alloc / _open: rank 0 owns the whole stream's device buffer, rank 1 (another process; on this
    one-GPU box the same device) opens the IPC handle and e1b200_synth_epochs delivers its segment there."""
    _run(2, 52000, 9, 7, 8, 3, tmp_path, "peer", use_gpu=True)
