"""GPU: the master-bus reduce through the C ABI (nodey_bus_reduce / nodey_bus_allreduce).  One rank: the sum over a
single partial bus is that bus, bit for bit, whichever planes / lengths / in-place form is used.  Two ranks (only on
a box with two devices): one process per GPU, id handed over through a file, the reduced bus against the float64 sum
and -- for config-5 style partial buses rendered by the engine -- against the one-GPU render within the 1e-5 bar."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bus1(nd):
    import torch
    torch.cuda.set_device(0)
    b = nd.Bus(nd.bus_unique_id(), 0, 1)
    yield b
    torch.cuda.synchronize()
    b.close()


def test_single_rank_info(bus1):
    assert bus1.info() == {"rank": 0, "nranks": 1, "device": 0}


@pytest.mark.parametrize("n", [1, 3, 1152, 100003, 6912000])
@pytest.mark.parametrize("nch", [1, 2])
def test_single_rank_reduce_is_the_identity(nd, bus1, n, nch):
    import torch
    g = torch.Generator(device="cuda").manual_seed(n + nch)
    send = torch.rand((nch, n), device="cuda", generator=g) - 0.5
    recv = torch.full((nch, n + 8), 7.0, device="cuda")          # guard band behind each plane
    bus1.reduce(send, recv[:, :n], root=0)
    torch.cuda.synchronize()
    assert torch.equal(recv[:, :n], send)
    assert bool((recv[:, n:] == 7.0).all()), "wrote behind the promised region"
    out = torch.zeros_like(send)
    bus1.allreduce(send, out)
    torch.cuda.synchronize()
    assert torch.equal(out, send)


def test_single_rank_in_place_and_empty(nd, bus1):
    import torch
    x = torch.rand((2, 4099), device="cuda")
    keep = x.clone()
    bus1.reduce(x, x, root=0)                                      # send == recv is allowed on the root
    bus1.reduce_ptrs(x[0].data_ptr(), x[1].data_ptr(), x[0].data_ptr(), x[1].data_ptr(), 0, root=0)   # nothing to do
    torch.cuda.synchronize()
    assert torch.equal(x, keep)


def test_argument_errors(nd, bus1):
    import torch
    x = torch.zeros((2, 16), device="cuda")
    with pytest.raises(nd.NodeyError, match="root"):
        bus1.reduce(x, x, root=1)
    with pytest.raises(nd.NodeyError, match="recv plane"):
        bus1.reduce(x, None, root=0)
    with pytest.raises(nd.NodeyError, match="negative"):
        bus1.reduce_ptrs(x[0].data_ptr(), 0, x[0].data_ptr(), 0, -1, root=0)


_WORKER = r"""
import os, sys, time
import numpy as np
root, rank, world, idfile, outfile = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "nodey-audio-editor_b200", "bindings"))
import torch, nodey
torch.cuda.set_device(rank)
nodey.check(nodey.lib().nodey_set_device(rank))
if rank == 0:
    with open(idfile + ".tmp", "wb") as f:
        f.write(nodey.bus_unique_id())
    os.replace(idfile + ".tmp", idfile)
t0 = time.time()
while not os.path.exists(idfile):
    assert time.time() - t0 < 60, "no id file"
    time.sleep(0.05)
bus = nodey.Bus(open(idfile, "rb").read(), rank, world)
n = 691200
g = torch.Generator(device="cuda").manual_seed(100 + rank)
send = torch.rand((2, n), device="cuda", generator=g) - 0.5
recv = torch.zeros((2, n), device="cuda")
bus.reduce(send, recv, root=0)
allr = torch.zeros((2, n), device="cuda")
bus.allreduce(send, allr)
torch.cuda.synchronize()
np.save(outfile + f".send{rank}.npy", send.cpu().numpy())
np.save(outfile + f".all{rank}.npy", allr.cpu().numpy())
if rank == 0:
    np.save(outfile + ".reduced.npy", recv.cpu().numpy())
bus.close()
"""


def test_two_rank_reduce_over_nvlink(nd, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    idfile, out = str(tmp_path / "bus.id"), str(tmp_path / "bus")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", idfile, out]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    s0, s1 = np.load(out + ".send0.npy"), np.load(out + ".send1.npy")
    want = s0.astype(np.float32) + s1.astype(np.float32)         # two addends: one rounding, order free
    assert np.array_equal(np.load(out + ".reduced.npy"), want)
    assert np.array_equal(np.load(out + ".all0.npy"), want)
    assert np.array_equal(np.load(out + ".all1.npy"), want)
