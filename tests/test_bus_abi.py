"""CPU: host-only behaviour of the master-bus entry points (nodey_bus_*, include/nodey_cuda.h): NCCL is bound at run
time, argument errors are reported before anything touches a device.  No communicator is created here (that needs
a GPU: tests/test_gpu_bus.py)."""
import ctypes as C

import pytest


def test_nccl_is_bound_at_run_time_and_reports_its_version():
    import nodey
    L = nodey.lib()
    v = C.c_int()
    nodey.check(L.nodey_bus_nccl_version(C.byref(v)))
    assert v.value >= 21800, v.value                   # NCCL 2.18+: the image carries 2.27 (system) and 2.28 (torch)


def test_unique_id_is_128_bytes_and_fresh():
    import nodey
    a, b = nodey.bus_unique_id(), nodey.bus_unique_id()
    assert len(a) == len(b) == nodey.BUS_ID_BYTES == 128
    assert a != b


def test_create_rejects_bad_arguments_without_touching_a_device():
    import nodey
    L = nodey.lib()
    h = C.c_void_p()
    assert L.nodey_bus_create(C.byref(h), None, 0, 1) == -1 and h.value is None
    assert b"id must point" in L.nodey_last_error()
    ident = C.c_char_p(b"\0" * 128)
    for rank, n in [(-1, 1), (1, 1), (3, 2), (0, 0)]:
        assert L.nodey_bus_create(C.byref(h), ident, rank, n) == -1, (rank, n)
        assert b"outside" in L.nodey_last_error()
    assert L.nodey_bus_create(None, ident, 0, 1) == -1


def test_collectives_reject_a_null_bus():
    import nodey
    L = nodey.lib()
    buf = (C.c_float * 4)()
    p = C.cast(buf, C.c_void_p)
    assert L.nodey_bus_reduce(None, p, None, p, None, 4, 0, None) == -1
    assert L.nodey_bus_allreduce(None, p, None, p, None, 4, None) == -1
    assert L.nodey_bus_info(None, None, None, None) == -1
    L.nodey_bus_destroy(None)                          # like free(NULL)
    with pytest.raises(nodey.NodeyError):
        nodey.check(L.nodey_bus_reduce(None, p, None, p, None, 4, -1, None))


def test_peer_master_orders_the_groups_like_the_graph(monkeypatch):
    """host logic of pipeline.PeerMaster without a device: the root hands nodey_mix the groups of rank 0, 1, 2 ... in
    that order (the master audio_amix's input order), each plane at its slot of the owner's block, after a barrier, and
    releases the blocks with a second barrier; other ranks launch nothing."""
    import nodey
    import pipeline

    class FakeBlock:
        def __init__(self, nbytes):
            self.ptr, self.nbytes, self.handle = 0x1000000, nbytes, b"me".ljust(64, b"\0")

        def close(self):
            self.ptr = None

    opened, closed, calls, events = [], [], [], []
    monkeypatch.setattr(nodey, "PeerBlock", FakeBlock)
    monkeypatch.setattr(nodey, "peer_open", lambda h: opened.append(h) or 0x10000000 * (1 + int(h[:1])))
    monkeypatch.setattr(nodey, "peer_close", lambda p: closed.append(p))
    monkeypatch.setattr(nodey, "mix_ptrs", lambda *a, **k: calls.append(a))
    world, groups_local, frames = 4, 4, 1000
    handles = [str(r).encode().ljust(64, b"\0") for r in range(world)]
    pm = pipeline.PeerMaster(0, world, groups_local, frames, lambda obj: handles, lambda: events.append("barrier"))
    assert [h[:1] for h in opened] == [b"1", b"2", b"3"]            # the root maps every other rank's block, not its own
    plane = pm.plane
    assert plane % 256 == 0 and plane >= frames * 4 and pm.block.nbytes == groups_local * 2 * plane
    pm.mix(111, 222, frames + 1152, 1.0 / 16, lambda: events.append("sync"))
    assert events == ["sync", "barrier", "sync", "barrier"]
    (out_l, out_r, in_l, in_r, lens, vols, total), = calls
    assert (out_l, out_r, total) == (111, 222, frames + 1152) and lens == [frames] * 16 and vols == [1.0 / 16] * 16
    bases = [0x1000000, 0x20000000, 0x30000000, 0x40000000]
    assert in_l == [bases[r] + (g * 2) * plane for r in range(world) for g in range(groups_local)]
    assert in_r == [bases[r] + (g * 2 + 1) * plane for r in range(world) for g in range(groups_local)]
    pm.close()
    assert sorted(closed) == bases[1:]
    # a non-root rank: no mapping, no launch, same two barriers
    opened.clear(); calls.clear(); events.clear()
    other = pipeline.PeerMaster(2, world, groups_local, frames, lambda obj: handles, lambda: events.append("barrier"))
    other.mix(0, 0, frames + 1152, 1.0 / 16, lambda: events.append("sync"))
    assert not opened and not calls and events == ["sync", "barrier", "barrier"]
