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
