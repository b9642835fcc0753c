#!/usr/bin/env python
"""bench.py -- offline-render realtime factor of the Nodey processor graph on B200 (BASELINE.json).

Workload (configs[4], the only configuration the metric is quoted on across 1/2/4/8 GPUs and one
that fits a single GPU): 256 independent 3 min stereo 44.1 kHz tracks through
resample -> pitch(+3 st) -> tempo(1.25, keep pitch) -> gain -> amix tree -> master bus -> spectrum.
Tracks are sharded over the ranks in contiguous groups of 16 (strong scaling: 256 tracks in total at
every N); the only collective is the reduce of the partial master buses.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [--steps K] [--warmup W]       # CPU reference arm (oracle port)

One JSON line on stdout (rank 0).  `value` = audio seconds rendered per second with the inputs
resident in HBM; `e2e` = the same through host buffers (pinned H2D of every track + D2H of bus and
spectrum inside the timed region).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
# the engine runs on up to ten CUDA streams: more hardware queues than the default 8, set before the CUDA context exists
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))

METRIC = "offline render realtime factor (audio-s/s) & HBM GB/s vs 8 TB/s per GPU"
UNIT = "audio-s/s"
TRACKS, SECONDS, IN_RATE = 256, 180, 44100
REF_TRACKS, REF_SECONDS = 16, 20          # bounded sample of the reference arm (see workload_config)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tracks", type=int, default=TRACKS, help="total tracks (development only; the contract is 256)")
    ap.add_argument("--seconds", type=int, default=SECONDS, help="track length (development only; the contract is 180)")
    ap.add_argument("--bus", default="nccl", choices=["nccl", "peer"],
                    help="N > 1: how the master bus is formed -- nccl: nodey_bus_reduce of the partial buses (1e-5 of the one-GPU bus); "
                         "peer: the root's master mix reads every rank's group mixes over NVLink peer memory (bit identical)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[0..3] block (N = 1) and the time-segment block")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-size parity checks after the timed region")
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {"workload": f"configs[4]: {args.tracks} x {args.seconds} s stereo 44.1 kHz float tracks, full graph "
                        "(audio_amix(1) resample -> pitch_modifier +3 -> velocity_modifier 1.25 keep_pitch -> "
                        "audio_volume_adjust -> audio_amix 16x tree -> master bus -> audio_spectrum 4096/1024)",
            "tracks": args.tracks, "track_seconds": args.seconds, "sharding": f"tracks/{n_gpus} per GPU, contiguous groups of 16",
            "collective": "none" if n_gpus == 1 else (
                "nodey_bus_reduce: ncclReduce(sum) of the partial master bus (C ABI, own communicator)" if getattr(args, "bus", "nccl") == "nccl"
                else "none: the root's master audio_amix reads every rank's group mixes over NVLink peer memory (nodey_peer_*, one kernel)"),
            "audio_seconds_per_step": args.tracks * args.seconds,
            "l2": "inputs larger than L2 (>= 2 GB per GPU per step), no flush needed",
            "reference_arm": f"--impl reference times the oracle port of the same graph on all host threads on a BOUNDED SAMPLE of this workload "
                             f"({REF_TRACKS} tracks x {REF_SECONDS} s per step, one track chain per thread); the realtime factor does not depend on "
                             "the number of tracks (independent chains) and the sample keeps a --steps 20 run within minutes"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference processors on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_sample(n_tracks, seconds, threads):
    """One pass of the same graph on a bounded sample; returns (audio seconds, wall seconds)."""
    from oracle import graph_oracle as G
    from oracle import oracle as O
    n = IN_RATE * seconds
    tracks = [O.synth_f32(n, 2, IN_RATE, t) for t in range(n_tracks)]
    t0 = time.perf_counter()
    G.render(tracks, threads=threads)
    return n_tracks * seconds, time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    n_tracks, seconds = REF_TRACKS, REF_SECONDS
    for _ in range(args.warmup):
        cpu_sample(n_tracks, seconds, cores)
    audio = wall = 0.0
    for _ in range(args.steps):
        a, w = cpu_sample(n_tracks, seconds, cores)
        audio += a; wall += w
    value = audio / wall
    sample = f"{n_tracks} tracks x {seconds} s of the same graph per step (oracle port of the reference processors)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall / max(args.steps, 1) * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def kernel_algo_bytes(name, r, batch):
    """ALGORITHMIC bytes one launch of `name` moves in this workload (kernel table):
    the bytes a perfect implementation must read plus the bytes it must write, per launch."""
    f8 = 8  # stereo float frame
    table = {
        "resample_tile_kernel": batch * (r.n_in * f8 + r.total1 * f8),           # one launch per batch of audio_amix(1): in + out planes
        "extract_planar_kernel": r.total1 * f8 * 2,
        "tds_offsets_kernel": None,                                               # filled per node below
        "st_post_kernel": None,
        "gain_f32_kernel": batch * r.m2 * f8 * 2,                                # one launch per batch of audio_volume_adjust nodes
        "to_fltp_kernel": r.m2 * f8 * 2,
        "mix_kernel": 16 * r.m2 * f8 + r.total2 * f8,
        "stft4096_kernel": r.total2 * f8 + 2 * r.spec_frames * 2049 * 8,
    }
    # SoundTouch kernels run once per node per sub-batch; both nodes read ~their input and write ~their output
    st_in = (r.total1 + r.m1) / 2.0
    st_out = (r.m1 + r.m2) / 2.0
    # the search reads, per sequence, the window (seek_length + overlap frames) and the mid buffer (overlap frames);
    # the frames between windows are never touched by this kernel.  Mean over the pitch and the tempo node.
    per_track = 0.0
    for st, n_in in ((r.st_pitch, r.total1), (r.st_tempo, r.m1)):
        info = st.info()
        _, nseq = st.out_frames(n_in, r.frame_size)
        per_track += 0.5 * max(nseq - 1, 0) * (info["seek_length"] + 2 * info["overlap"]) * f8
    table["tds_offsets_kernel"] = batch * per_track
    table["st_post_kernel"] = batch * (st_in + st_out) * f8                       # fused cross-fade + FIR + cubic: in once, out once
    return table.get(name)


def tds_fp32_ops(r, batch):
    """separately rounded FP32 operations one WSOLA offsets launch must issue (mean of the pitch and the tempo node):
    per sequence, seek_length candidates x (channels * overlap) samples x (multiply + add) for the correlation, plus
    half an add per (candidate, sample) for the norm sums shared between candidate classes."""
    ops = 0.0
    for st, n_in in ((r.st_pitch, r.total1), (r.st_tempo, r.m1)):
        info = st.info()
        _, nseq = st.out_frames(n_in, r.frame_size)
        ops += 0.5 * max(nseq - 1, 0) * info["seek_length"] * 2 * info["overlap"] * 2.5
    return batch * ops


class JsonStdout:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout when
    NCCL_DEBUG is WARN or VERSION), so the real stdout is put aside for the result line and file descriptor 1 points at
    stderr while the bench runs."""

    def __init__(self):
        sys.stdout.flush()
        self.fd = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        data = (line.rstrip("\n") + "\n").encode()
        while data:
            data = data[os.write(self.fd, data):]


def check_parity(args, world, rank, dev, eng, ids, first, t_local, bus_t, finish, bind):
    """Full-size parity of the render that was timed (not part of any timed region).
    N = 1: the per-track chains (audio_amix(1) -> pitch -> tempo -> gain) of the first and the last track of the 256 x 180 s
    render against the oracle, bit for bit (about 2 s of CPU each), and the master bus against the ordered sum of the group
    mixes the engine published.  N > 1: the reduced master bus against the one-GPU render of all tracks, made on rank 0 after
    the timed region -- another summation order, so the bar is BASELINE.json's 1e-5 (relative to the bus peak); the run
    fails above it."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import nodey
    import engine
    import pipeline
    L = nodey.lib()
    n_in = IN_RATE * args.seconds
    x = torch.empty((t_local, n_in, 2), dtype=torch.float32, device=dev)
    for t in range(t_local):
        nodey.check(L.nodey_synth(nodey._dp(x[t]), None, n_in, 2, IN_RATE, first + t, 0, None))
    torch.cuda.synchronize()
    bind(x)
    eng.run()
    finish()
    torch.cuda.synchronize()
    if world == 1:
        from oracle import graph_oracle as G
        from oracle import oracle as O
        O.build()
        res = {"what": "tracks 0 and %d of the %d x %d s render (per-track chain at the audio_volume_adjust output) vs oracle/graph_oracle.track_chain; "
                       "master bus vs the ordered, separately rounded sum of the published group mixes" % (t_local - 1, args.tracks, args.seconds)}
        ok = True
        for t in sorted({0, t_local - 1}):
            ref = G.track_chain(x[t].cpu().numpy(), G.track_gain(first + t))
            got = eng.product(ids["gains"][t], "output").numpy()
            same = got.shape == ref.shape and np.array_equal(got.view(np.uint32), ref.view(np.uint32))
            res[f"track_{t}_bit_exact"] = bool(same)
            ok = ok and same
        bus = eng.output().numpy()
        acc = np.zeros_like(bus)
        v = np.float32(1.0 / 16)
        for g in ids["groups"]:
            gm = eng.product(g, "output").numpy()
            acc[:, :gm.shape[1]] = acc[:, :gm.shape[1]] + gm * v
        same = np.array_equal(bus.view(np.uint32), acc.view(np.uint32))
        res["master_bus_bit_exact"] = bool(same)
        res["ok"] = bool(ok and same)
        if not res["ok"]:
            raise SystemExit(f"parity check failed: {res}")
        return res
    # N > 1: the reduced bus of this step against the one-GPU render of every track
    reduced = bus_t[0].clone() if rank == 0 else None
    del x
    eng.close()                    # the shard's products go back to the allocator: rank 0 needs the room for all tracks
    torch.cuda.empty_cache()
    nodey.check(L.nodey_trim_memory())
    res = None
    if rank == 0:
        gains = [pipeline.track_gain(t) for t in range(args.tracks)]
        project, ids1 = engine.config5_project(args.tracks, gains, spectrum=False)
        e1 = engine.Engine(project.json())
        xs = torch.empty((args.tracks, n_in, 2), dtype=torch.float32, device=dev)
        for t in range(args.tracks):
            nodey.check(L.nodey_synth(nodey._dp(xs[t]), None, n_in, 2, IN_RATE, t, 0, None))
        torch.cuda.synchronize()
        for t in range(args.tracks):
            e1.bind_source(t, xs[t], nodey.FMT_FLT, IN_RATE)
        e1.run()
        one = e1.output()
        ref = torch.empty((2, one.frames), dtype=torch.float32, device=dev)
        nodey.check(L.nodey_memcpy_d2d(nodey._dp(ref[0]), C.c_void_p(one.p0), one.frames * 4, None))
        nodey.check(L.nodey_memcpy_d2d(nodey._dp(ref[1]), C.c_void_p(one.p1), one.frames * 4, None))
        torch.cuda.synchronize()
        e1.close()
        del xs
        peak = float(ref.abs().max().item())
        err = float((reduced - ref).abs().max().item()) if tuple(reduced.shape) == tuple(ref.shape) else float("inf")
        rel = err / max(peak, 1e-30)
        res = {"what": f"master bus reduced over {world} ranks ({args.bus}) vs the one-GPU render of all {args.tracks} tracks made on rank 0",
               "bus_max_rel": rel, "bus_peak": peak, "frames": int(ref.shape[1]), "bar": 1e-5 if args.bus == "nccl" else 0.0,
               "ok": bool(rel <= (1e-5 if args.bus == "nccl" else 0.0))}
        torch.cuda.empty_cache()
        nodey.check(L.nodey_trim_memory())
    flag = torch.tensor([1 if (res is None or res["ok"]) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if not bool(flag.item()):
        raise SystemExit(f"parity check failed: {res}")
    return res


def measure_footprint(args, dev, eng, t_local):
    """High-water mark of device memory the library obtained from the driver for ONE render of the workload (sources
    resident, their 16.3 GB belong to the caller and are not counted), from an empty allocator cache, under both memory
    policies: keep (every link's product lives until the next run, like the reference's link_products) and release
    (infra::Runner::release_products: a link lets go once its consumer has enqueued; freed blocks are reused in stream
    order).  The timed steps use `keep`."""
    import torch
    import nodey
    import engine
    L = nodey.lib()
    n_in = IN_RATE * args.seconds
    eng._keep = []                         # sources bound earlier (parity check) may go
    torch.cuda.empty_cache()
    x = torch.empty((t_local, n_in, 2), dtype=torch.float32, device=dev)
    for t in range(t_local):
        nodey.check(L.nodey_synth(nodey._dp(x[t]), None, n_in, 2, IN_RATE, t, 0, None))
    for t in range(t_local):
        eng.bind_source(t, x[t], nodey.FMT_FLT, IN_RATE)
    out = {}
    for policy in ("keep", "release"):
        engine.set_release_products(policy == "release")
        try:
            eng.run()                      # drops the previous run's products
            torch.cuda.synchronize()
            if policy == "release":
                eng.run()
                torch.cuda.synchronize()
            nodey.check(L.nodey_trim_memory())
            held, _ = nodey.memory_reserved()      # what the last run still holds (sink, spectrum; every product under `keep`)
            nodey.memory_stats(reset_peak=True)
            t0 = time.perf_counter()
            eng.run()
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3
            _, peak = nodey.memory_reserved()
            out[policy] = {"peak_gb": round(peak / 1e9, 2), "held_before_gb": round(held / 1e9, 2), "cold_cache_run_ms": round(ms, 1)}
        finally:
            engine.set_release_products(False)
    out["note"] = "peak_gb: device memory obtained through the library's allocator during one render that starts from an empty cache (it " \
                  "includes what the previous run's surviving products still held: held_before_gb); the caller's sources (%.1f GB) are " \
                  "not counted; cold_cache_run_ms is that render's wall time with every block freshly cudaMalloc'ed" % (t_local * n_in * 8 / 1e9)
    eng.run()
    torch.cuda.synchronize()
    del x
    return out


def run_ours(args):
    result_out = JsonStdout()
    import ctypes as C
    import torch
    import torch.distributed as dist
    import nodey
    import engine
    import pipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = nodey.lib()        # fails loudly when the CUDA library is missing: there is no fallback path
    engine.lib()
    nodey.check(L.nodey_set_device(local))
    bus = None
    if world > 1:
        # torch.distributed is the plumbing (rendezvous, barriers, max over ranks); the bus reduce itself runs on the
        # library's own communicator, set up from an id that rank 0 makes and the process group hands round
        box = [nodey.bus_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        bus = nodey.Bus(box[0], rank, world)

    n_in = IN_RATE * args.seconds
    first, t_local = pipeline.shard_tracks(args.tracks, world, rank)
    plan = pipeline.Config5Renderer(n_in, sub_batch=16, device=dev)      # host-side lengths only (roofline bookkeeping)
    gains = [pipeline.track_gain(first + t) for t in range(t_local)]
    # the render is driven through the reference-facing plugin API: project JSON -> Graph -> Runner -> nodes
    project, ids = engine.config5_project(t_local, gains, spectrum=(world == 1))
    eng = engine.Engine(project.json())

    # synthetic sources, generated on the device by the same generator the oracle uses
    x_dev = torch.empty((t_local, n_in, 2), dtype=torch.float32, device=dev)
    for t in range(t_local):
        nodey.check(L.nodey_synth(nodey._dp(x_dev[t]), None, n_in, 2, IN_RATE, first + t, 0, None))
    torch.cuda.synchronize()

    bus_t = [None]
    peer = [None]

    def finish(want_host=None):
        """after eng.run(): reduce the partial buses (N > 1), spectrum on rank 0, optional D2H"""
        out = eng.output()
        if world > 1:
            if bus_t[0] is None:
                bus_t[0] = torch.empty((2, out.frames), dtype=torch.float32, device=dev)
            if args.bus == "peer":
                # compute + exchange in one kernel: the root's master mix over the group mixes of every rank, remote ones
                # read through CUDA-IPC mapped pointers over NVLink, in the graph's input order (bit identical to one GPU)
                groups = [eng.product(g, "output") for g in ids["groups"]]
                if peer[0] is None:
                    def exchange(obj):
                        box = [None] * world
                        dist.all_gather_object(box, obj)
                        return box
                    peer[0] = pipeline.PeerMaster(rank, world, len(groups), groups[0].frames, exchange, dist.barrier)
                peer[0].stage(groups)
                peer[0].mix(bus_t[0][0].data_ptr(), bus_t[0][1].data_ptr(), out.frames, 1.0 / 16, torch.cuda.synchronize)
            else:
                # the only collective of the path: sum of the partial master buses straight out of the engine's planes,
                # through the C ABI (nodey_bus_reduce: ncclReduce of both planes in one group, no staging copy)
                bus.reduce_ptrs(out.p0, out.p1, bus_t[0][0].data_ptr(), bus_t[0][1].data_ptr(), out.frames, root=0)
            spec_ptr, spec_elems = None, 0
            if rank == 0:
                if len(bus_t) < 2:     # the spectrum buffer is allocated once: a fresh 221 MB cudaMalloc per step stalled rank 0 for up to 50 ms
                    bus_t.append(torch.empty((2, nodey.stft_frames(out.frames), 2049), dtype=torch.complex64, device=dev))
                spec = nodey.stft(bus_t[0], False, out=bus_t[1])
                spec_ptr, spec_elems = spec.data_ptr(), spec.numel()
            p0, p1 = bus_t[0][0].data_ptr(), bus_t[0][1].data_ptr()
        else:
            sp = eng.product(ids["spectrum"], "output")
            spec_ptr, spec_elems = sp.p0, sp.ch * sp.frames * sp.bins
            p0, p1 = out.p0, out.p1
        if want_host is not None and rank == 0:
            hb, hs = want_host(out.frames, spec_elems)
            nodey.check(L.nodey_memcpy_d2h(nodey._dp(hb[0]), C.c_void_p(p0), out.frames * 4, None))
            nodey.check(L.nodey_memcpy_d2h(nodey._dp(hb[1]), C.c_void_p(p1), out.frames * 4, None))
            nodey.check(L.nodey_memcpy_d2h(nodey._dp(hs), C.c_void_p(spec_ptr), spec_elems * 8, None))
        return out.frames, spec_elems

    def bind(sources):
        for t in range(t_local):
            eng.bind_source(t, sources[t], nodey.FMT_FLT, IN_RATE)

    dev_split = [0.0, 0.0]

    def step_device():
        t0 = time.perf_counter()
        eng.run()
        t1 = time.perf_counter()
        finish()
        if world > 1:
            torch.cuda.synchronize()
        dev_split[0] += t1 - t0; dev_split[1] += time.perf_counter() - t1

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item())

    bind(x_dev)
    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = nodey.profile_launches()
    dev_split[0] = dev_split[1] = 0.0
    ms_total = timed(step_device, args.steps)
    print(f"[bench] rank {rank} device-resident split per step: engine run {dev_split[0] / args.steps * 1e3:.1f} ms, "
          f"bus reduce / spectrum {dev_split[1] / args.steps * 1e3:.1f} ms", file=sys.stderr, flush=True)
    launches = nodey.profile_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    audio = args.tracks * args.seconds * args.steps
    value = audio / (ms_total * 1e-3)

    # ---- per-kernel device time of one step (events around every launch; not part of the timed runs) ----
    roofline = None
    # In the timed steps the nodes of a track chain overlap chunk by chunk on the lane's three streams, so events around a
    # launch would time its share of the machine, not the kernel.  The profiled step therefore runs the same graph with
    # whole-track launches (NODEY_ST_CHUNKS=1) as ONE wave on one lane: every launch alone on the device, full batch.
    # Likewise the WSOLA chains run as ONE launch per node there (NODEY_ST_CHUNKS=1): in the timed steps the pitch and the
    # tempo node's chunk launches overlap on two streams.
    os.environ["NODEY_WAVE"] = "1000000"
    os.environ["NODEY_ST_CHUNKS"] = "1"
    step_device()
    nodey.profile_enable(True)
    step_device()
    rep = nodey.profile_report()
    nodey.profile_enable(False)
    del os.environ["NODEY_WAVE"]
    del os.environ["NODEY_ST_CHUNKS"]
    if rank == 0 and rep:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        name, st = max(rep.items(), key=lambda kv: kv[1]["ms"])
        per_launch_ms = st["ms"] / st["launches"]
        st_batch = min(256, t_local)                     # tracks per SoundTouch launch when the sources are resident (host/src/nodes.cpp)
        ab = kernel_algo_bytes(name, plan, st_batch)
        achieved = (ab / (per_launch_ms * 1e-3) / 1e9) if ab else None
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(name)
        except Exception:
            pass
        step_ms = sum(v["ms"] for v in rep.values())
        roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                    "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                    "launch_ms": per_launch_ms, "share_of_step": st["ms"] / step_ms if step_ms else None,
                    "note": "the dominant kernel is the WSOLA offset search: FP32-issue bound (about 610 separately rounded FP32 "
                            "operations per input frame, sequential per track), so its HBM fraction is small by construction; "
                            "`fp32` gives its fraction of the FP32 issue peak; see DESIGN.md 3.1",
                    "kernels_ms": {k: round(v["ms"], 3) for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])}}
        # every kernel of the step against the same HBM peak: algorithmic bytes of its launches / its device time
        per_kernel = {}
        batches = {"resample_tile_kernel": st_batch, "tds_offsets_kernel": st_batch, "st_post_kernel": st_batch, "gain_f32_kernel": st_batch}
        for k, v in rep.items():
            kb = kernel_algo_bytes(k, plan, batches.get(k, 1))
            if kb and v["ms"] > 0:
                gbs = kb * v["launches"] / (v["ms"] * 1e-3) / 1e9
                per_kernel[k] = {"launches": v["launches"], "gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}
        roofline["kernels_hbm"] = per_kernel
        if name == "tds_offsets_kernel":
            sm_mhz = float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0)
            fp_peak = nodey.device_info()["sm_count"] * 128 * sm_mhz * 1e6          # one FP32 operation per lane and clock (no FMA: -fmad=false)
            fp_ach = tds_fp32_ops(plan, st_batch) / (per_launch_ms * 1e-3)
            roofline["fp32"] = {"achieved": fp_ach / 1e12, "peak": fp_peak / 1e12, "unit": "Tops/s (separately rounded mul/add)",
                                "frac": fp_ach / fp_peak, "peak_source": f"SMs x 128 lanes x {sm_mhz:.0f} MHz"}

    # ---- end to end: pinned host inputs -> H2D (audio_input node) -> render -> D2H(bus, spectrum) ----
    e2e = None
    if not args.no_e2e:
        x_host = torch.empty((t_local, n_in, 2), dtype=torch.float32, pin_memory=True)
        x_host.copy_(x_dev)
        torch.cuda.synchronize()
        del x_dev
        torch.cuda.empty_cache()
        bind(x_host)
        host_out = {}

        def host_buffers(frames, spec_elems):
            if "bus" not in host_out:
                host_out["bus"] = torch.empty((2, frames), dtype=torch.float32, pin_memory=True)
                host_out["spec"] = torch.empty((spec_elems,), dtype=torch.complex64, pin_memory=True)
            return host_out["bus"], host_out["spec"]

        sizes = [0, 0]

        split = [0.0, 0.0]

        def step_e2e():
            t0 = time.perf_counter()
            eng.run()
            t1 = time.perf_counter()
            f, se = finish(want_host=host_buffers)
            torch.cuda.synchronize()
            split[0] += t1 - t0; split[1] += time.perf_counter() - t1
            sizes[0], sizes[1] = f, se

        for _ in range(2):
            step_e2e()
        split[0] = split[1] = 0.0
        ms_e2e = timed(step_e2e, args.steps)
        if rank == 0:
            print(f"[bench] e2e split per step: engine run {split[0] / args.steps * 1e3:.1f} ms, bus reduce / D2H {split[1] / args.steps * 1e3:.1f} ms",
                  file=sys.stderr, flush=True)
        e2e = {"value": audio / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": t_local * n_in * 8 * world,
               "d2h_bytes_per_step": sizes[0] * 8 + sizes[1] * 8, "ms_per_step": ms_e2e / args.steps,
               "api": "project JSON -> infra::Graph::deserialize -> infra::Runner::create_and_run (libnodey_host.so, "
                      "include/nodey_engine.h); audio_input uploads the pinned host PCM on its own stream, downstream nodes start per track"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        a, w = cpu_sample(16, 60, cores)
        a1, w1 = cpu_sample(16, 20, 1)
        cpu_baseline = {"value": a / w, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "16 tracks x 60 s of the same graph, oracle port of the reference processors, one track chain per host thread",
                        # the reference's Runner drives a whole graph from ONE thread: the same port on one core
                        "single_thread": {"value": a1 / w1, "unit": UNIT, "cores": 1, "sample": "16 tracks x 20 s of the same graph"}}

    # ---- parity of what was timed, at full size (after the timed regions) ----
    parity = None
    if not args.no_parity:
        parity = check_parity(args, world, rank, dev, eng, ids, first, t_local, bus_t, finish, bind)

    # ---- device-memory footprint of one render under both memory policies (not timed) ----
    memory = None
    if world == 1 and not args.no_parity and not args.no_e2e:
        memory = measure_footprint(args, dev, eng, t_local)

    configs = segments = None
    if not args.no_configs:
        eng.close()
        torch.cuda.empty_cache()
        nodey.check(L.nodey_trim_memory())
        import bench_configs
        peak_gbs = 6650.0
        try:
            peak_gbs = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak_gbs))
        except Exception:
            pass
        if world == 1:
            configs = bench_configs.run_configs(torch, nodey, engine, peak_gbs)
        segments = bench_configs.run_segments(torch, nodey, dist, world, rank, dev)

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": workload_config(args, world), "clocks": clocks,
               "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
               "parity": parity, "memory": memory, "configs": configs, "segments": segments}
        result_out.emit(json.dumps(out))
    eng.close()
    if world > 1:
        torch.cuda.synchronize()
        if peer[0] is not None:
            dist.barrier()
            peer[0].close()
        bus.close()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
