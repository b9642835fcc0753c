"""CPU, world_size 2 (gloo): the multi-GPU host logic -- contiguous sharding of the tracks in whole
amix groups, per-rank partial master bus, one reduce(sum) to rank 0 -- on the oracle graph.  The
reduced bus must equal the single-process bus within the float bar (summation order differs across
ranks: 1e-5 relative / -100 dBFS)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, total, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))
    import torch
    import torch.distributed as dist
    import pipeline
    from oracle import graph_oracle as G
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, cnt = pipeline.shard_tracks(total, world, rank)
    tracks = [O.synth_f32(n, 2, 44100, first + t) for t in range(cnt)]
    bus, _ = G.render(tracks, first_track=first, spectrum=False)
    t = torch.from_numpy(bus.copy())
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.destroy_process_group()


def test_sharding_helper():
    sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))
    import pipeline
    assert pipeline.shard_tracks(256, 8, 3) == (96, 32)
    assert pipeline.shard_tracks(256, 1, 0) == (0, 256)
    got = [pipeline.shard_tracks(256, 4, r) for r in range(4)]
    assert [g[0] for g in got] == [0, 64, 128, 192]
    with pytest.raises(ValueError):
        pipeline.shard_tracks(100, 8, 0)


def test_two_rank_partial_bus_reduce(tmp_path, orc):
    import torch.multiprocessing as mp
    from oracle import graph_oracle as G
    n, total = 22050, 32
    out = str(tmp_path / "bus.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, n, total, out), nprocs=2, join=True)
    got = np.load(out)
    tracks = [orc.synth_f32(n, 2, 44100, t) for t in range(total)]
    ref, _ = G.render(tracks, spectrum=False, threads=4)
    assert got.shape == ref.shape
    resid = np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()
    assert resid <= max(1e-5 * np.abs(ref).max(), 1e-5), resid


_EMIT = r"""
import ctypes, json, os, sys
sys.path.insert(0, sys.argv[1])
import bench
out = bench.JsonStdout()
print('python noise')
libc = ctypes.CDLL(None)
libc.puts(b'NCCL version 2.28.9+cuda12.9')
libc.fflush(None)
os.write(1, b'raw noise\n')
out.emit(json.dumps({'metric': 'm', 'value': 1.5}))
print('late noise')
"""


def test_bench_result_line_is_the_only_thing_on_stdout(tmp_path):
    """bench.py's JsonStdout: whatever a library prints to stdout while the bench runs (NCCL's version banner, C stdio
    included) lands on stderr; the real stdout carries exactly the one JSON line of the contract."""
    import json
    import subprocess
    script = tmp_path / "emit.py"
    script.write_text(_EMIT)
    r = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count("\n") == 1 and json.loads(r.stdout) == {"metric": "m", "value": 1.5}
    for noise in ("python noise", "NCCL version", "raw noise", "late noise"):
        assert noise in r.stderr
