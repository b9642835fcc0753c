"""GPU, two ranks (two devices when the box has them, else two processes on one device): the master mix over peer memory (nodey_peer_*, pipeline.PeerMaster).  Two processes render
16 tracks each through the plugin API; rank 0 mixes the group mixes of BOTH ranks with nodey_mix, reading rank 1's
through a CUDA-IPC mapped pointer.  The bus must be bit identical to the one-GPU render of all 32 tracks (the reduce of
partial buses, nodey_bus_reduce, is only within 1e-5).  One device: the export / open / read-through path is checked
inside one process (a block opened by its own exporter is not allowed by CUDA IPC, so only export + alloc are run)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_block_export_and_mix_on_raw_addresses(nd):
    import torch
    blk = nd.PeerBlock(1 << 20)
    assert blk.ptr and len(blk.handle) == nd.PEER_HANDLE_BYTES and blk.handle != b"\0" * 64
    # nodey_mix on raw addresses: one input inside the peer block, one ordinary tensor, volumes 0.5 / 0.25
    n = 4099
    a = torch.rand((2, n), device="cuda") - 0.5
    b = torch.rand((2, n), device="cuda") - 0.5
    plane = (n * 4 + 255) // 256 * 256
    import ctypes as C
    for ch in range(2):
        nd.check(nd.lib().nodey_memcpy_d2d(C.c_void_p(blk.ptr + ch * plane), C.c_void_p(a[ch].data_ptr()), n * 4, nd._stream()))
    out = torch.empty((2, n + 5), device="cuda")
    nd.mix_ptrs(out[0].data_ptr(), out[1].data_ptr(), [blk.ptr, b[0].data_ptr()], [blk.ptr + plane, b[1].data_ptr()], [n, n], [0.5, 0.25], n + 5)
    torch.cuda.synchronize()
    want = nd.mix([a, b], [0.5, 0.25], nframes=n + 5)
    assert torch.equal(out, want)
    blk.close()


_WORKER = r"""
import os, sys
import numpy as np
root, rank, world, port, outfile = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "nodey-audio-editor_b200", "bindings"))
import torch, torch.distributed as dist
import nodey, engine, pipeline
dev = rank % torch.cuda.device_count()          # one device: both ranks share it (CUDA IPC works between processes of one GPU)
torch.cuda.set_device(dev)
nodey.check(nodey.lib().nodey_set_device(dev))
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
n, total = 44100, 16 * world
first, cnt = pipeline.shard_tracks(total, world, rank)

def render(first, cnt):
    project, ids = engine.config5_project(cnt, [pipeline.track_gain(first + t) for t in range(cnt)], spectrum=False)
    e = engine.Engine(project.json())
    xs = [nodey.synth(n, 2, 44100, track=first + t) for t in range(cnt)]
    for t in range(cnt):
        e.bind_source(t, xs[t], nodey.FMT_FLT, 44100)
    e.run()
    return e, ids

e, ids = render(first, cnt)
groups = [e.product(g, "output") for g in ids["groups"]]
out = e.output()

def exchange(obj):
    box = [None] * world
    dist.all_gather_object(box, obj)
    return box

pm = pipeline.PeerMaster(rank, world, len(groups), groups[0].frames, exchange, dist.barrier)
pm.stage(groups)
bus = torch.zeros((2, out.frames), device="cuda")
pm.mix(bus[0].data_ptr(), bus[1].data_ptr(), out.frames, 1.0 / 16, torch.cuda.synchronize)
np.save(outfile + f".partial{rank}.npy", out.numpy())     # what nodey_bus_reduce would sum
if rank == 0:
    np.save(outfile + ".peer.npy", bus.cpu().numpy())
    ref, _ = render(0, total)                      # the one-GPU render of every track, same device
    np.save(outfile + ".ref.npy", ref.output().numpy())
    ref.close()
dist.barrier()
pm.close()
e.close()
dist.destroy_process_group()
"""


def test_two_rank_master_mix_over_peer_memory_is_bit_identical_to_one_gpu(nd, tmp_path):
    # two devices: one rank per GPU, rank 1's groups are read over NVLink; one device: both ranks share it, so the
    # cross-process path (export, open, ordering barriers, input order of the mix) is checked on every box
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    out = str(tmp_path / "bus")
    port = str(29600 + os.getpid() % 300)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", port, out]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    got, ref = np.load(out + ".peer.npy"), np.load(out + ".ref.npy")
    assert got.shape == ref.shape and np.abs(ref).max() > 1e-3
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), "peer-memory master mix differs from the one-GPU bus"
    # the other way to form the bus (nodey_bus_reduce): the sum of the ranks' partial master buses -- another summation
    # order, so the 1e-5 bar of BASELINE.json applies (relative to the bus peak)
    p0, p1 = np.load(out + ".partial0.npy"), np.load(out + ".partial1.npy")
    assert p0.shape == p1.shape == ref.shape
    assert np.abs((p0 + p1) - ref).max() <= 1e-5 * np.abs(ref).max()
