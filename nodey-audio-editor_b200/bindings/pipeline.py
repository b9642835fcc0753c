"""Config-5 render graph over the C ABI (BASELINE.json configs[4]), per GPU:

    audio_input(track t, 44.1 kHz FLT stereo)
      -> audio_amix(input_num=1, volume 1)            # the reference's only resampling node: 48 kHz FLTP
      -> pitch_modifier(+3 semitones)                 # SoundTouch, FLT interleaved out
      -> velocity_modifier(1.25, keep_pitch)          # SoundTouch
      -> audio_volume_adjust(gain_t)
      -> audio_amix(input_num=16) per group of 16     # level-1 buses
      -> audio_amix(input_num=n_groups)               # master bus (partial when tracks are sharded)
      -> audio_spectrum(4096, 1024, hann)

Every arrow is a reference node boundary: the value sequence on it equals what the reference's
frame stream would carry (zero padding of amix included).  Python only sequences C-ABI calls and
owns buffers (torch tensors); the C++ host layer (host/) does the same from project JSON.
"""
import numpy as np

import nodey as nd

IN_RATE = 44100
GROUP = 16


def track_gain(t):
    """per-track gain of the audio_volume_adjust node (deterministic, in the UI range [0, 10])."""
    return float(np.float32(0.5) + np.float32(0.0625) * np.float32(t % 9))


def shard_tracks(total_tracks, world, rank):
    """Tracks are sharded in contiguous blocks so that every rank owns whole amix groups of 16:
    returns (first_track, n_tracks) of `rank`.  The master bus is then sum over ranks of the partial
    level-2 mixes (the only collective of the render)."""
    if total_tracks % (GROUP * world) != 0:
        raise ValueError(f"{total_tracks} tracks do not split into groups of {GROUP} over {world} ranks")
    per = total_tracks // world
    return rank * per, per


class Config5Renderer:
    def __init__(self, n_in, sub_batch=32, frame_size=1152, pitch_semitones=3.0, velocity=1.25, device="cuda"):
        import torch
        self.torch = torch
        self.n_in, self.frame_size, self.sub_batch, self.device = n_in, frame_size, sub_batch, device
        assert sub_batch % GROUP == 0
        self.rs = nd.Resampler(IN_RATE, 48000)
        self.total1, segs, runs1 = nd.amix_plan([IN_RATE], [nd.uniform_runs(n_in, frame_size)])
        assert len(segs) == 1 and segs[0][1] == 0 and segs[0][2] == 0
        self.len1 = segs[0][3]
        self.st_pitch = nd.SoundTouch.pitch_node(48000, 2, pitch_semitones)
        self.st_tempo = nd.SoundTouch.velocity_node(48000, 2, velocity, True)
        self.m1, _ = self.st_pitch.out_frames(self.total1, frame_size)
        self.m2, _ = self.st_tempo.out_frames(self.m1, frame_size)
        # SoundTouch frames are 1152*k samples, the last one shorter (audio-velocity.cpp:416-424): the amix
        # bookkeeping only needs the total and a frame size; the reference pulls min(avail, 3*1152/velocity)
        # the SoundTouch nodes cut their output into frame_size frames (canonical chunking)
        self.total2, segs2, self.runs2 = nd.amix_plan([48000] * GROUP, [nd.uniform_runs(self.m2, frame_size)] * GROUP)
        assert len(segs2) == GROUP and all(s[1] == 0 and s[2] == 0 and s[3] == self.m2 for s in segs2)
        self.spec_frames = nd.stft_frames(self.total2)
        self._ws = None

    def out_seconds(self):
        return self.total2 / 48000.0

    def _workspace(self, b):
        t = self.torch
        if self._ws is None or self._ws["b"] < b:
            dev = self.device
            self._ws = dict(
                b=b,
                res=t.empty((2, self.total1), dtype=t.float32, device=dev),
                xi=t.empty((b, self.total1, 2), dtype=t.float32, device=dev),
                y1=t.empty((b, self.m1, 2), dtype=t.float32, device=dev),
                y2=t.empty((b, self.m2, 2), dtype=t.float32, device=dev),
                pl=t.empty((GROUP, 2, self.m2), dtype=t.float32, device=dev),
            )
        return self._ws

    def render_groups(self, x, first_track, group_vol=1.0 / 16, keep=None):
        """x: [B, n_in, 2] float32 device tensor, B multiple of 16 -> list of level-1 buses [2, total2].
        keep: optional dict that receives intermediate products of track 0 of the batch (tests)."""
        t = self.torch
        b = x.shape[0]
        assert b % GROUP == 0 and b <= self.sub_batch
        ws = self._workspace(self.sub_batch)
        res, xi, y1, y2, pl = ws["res"], ws["xi"][:b], ws["y1"][:b], ws["y2"][:b], ws["pl"]
        for k in range(b):
            # audio_amix(1): resample + (0 + x*1), zero padded to whole iterations; then A8 extraction
            self.rs.resample_mix([x[k]], [nd.FMT_FLT], [1.0], out_lens=[self.len1], out_frames=self.total1, out=res)
            nd.check(nd.lib().nodey_extract_interleaved(nd._dp(xi[k]), nd._dp(res[0]), nd._dp(res[1]), nd.FMT_FLTP,
                                                        self.total1, 2, nd._stream()))
            if keep is not None and k == 0:
                keep["amix1"] = res.clone()
        self.st_pitch.run(xi, self.frame_size, out=y1)
        self.st_tempo.run(y1, self.frame_size, out=y2)
        if keep is not None:
            keep["pitch"] = y1[0].clone(); keep["tempo"] = y2[0].clone()
        buses = []
        for g in range(b // GROUP):
            ins = []
            for k in range(GROUP):
                tr = g * GROUP + k
                nd.gain(y2[tr], nd.FMT_FLT, track_gain(first_track + tr), out=y2[tr])
                nd.check(nd.lib().nodey_to_fltp_stereo(nd._dp(pl[k][0]), nd._dp(pl[k][1]), nd._dp(y2[tr]), None,
                                                       nd.FMT_FLT, 2, self.m2, nd._stream()))
                ins.append(pl[k])
            buses.append(nd.mix(ins, [group_vol] * GROUP, nframes=self.total2))
        return buses

    def master(self, buses, vol=1.0 / 16):
        """level-2 audio_amix over the level-1 buses (<= 16) -> [2, total3] planar."""
        total3, segs, _ = nd.amix_plan([48000] * len(buses), [self.runs2] * len(buses))
        return nd.mix(buses, [vol] * len(buses), nframes=total3)

    def spectrum(self, bus, out=None):
        return nd.stft(bus, False, out=out)

    def render(self, x_all, first_track=0, host_input=False, copy_stream=None):
        """x_all: [T, n_in, 2] (device, or pinned host when host_input). Returns (bus, spectrum)."""
        t = self.torch
        T = x_all.shape[0]
        buses = []
        for s in range(0, T, self.sub_batch):
            xb = x_all[s:s + self.sub_batch]
            if host_input:
                xb = xb.to(self.device, non_blocking=True)
            buses += self.render_groups(xb, first_track + s)
        bus = self.master(buses)
        return bus


class PeerMaster:
    """Master mix of a track-sharded render as ONE kernel over NVLink peer memory (nodey_peer_*, include/nodey_cuda.h).

    Every rank stages the level-1 group mixes of its tracks (FLTP, 48 kHz) in a block the other processes can map; the
    root then runs the graph's master audio_amix -- nodey_mix, inputs in the graph's order, audio-amix.cpp:293-307 --
    over the groups of ALL ranks, reading the remote ones through the mapped pointers.  Compute and exchange are one
    kernel and the bus is bit identical to the one-GPU render (a reduce of partial buses sums in another order).

    exchange(obj) -> [obj of rank 0, obj of rank 1, ...] and barrier() are the caller's plumbing (torch.distributed,
    a launcher, files): the library only needs the 64-byte handles handed round once and two barriers per step.
    """

    def __init__(self, rank, world, groups_local, group_frames, exchange, barrier, root=0):
        self.rank, self.world, self.root = rank, world, root
        self.groups_local, self.frames = groups_local, int(group_frames)
        self.barrier = barrier
        self.plane = (self.frames * 4 + 255) // 256 * 256
        self.block = nd.PeerBlock(max(1, groups_local * 2 * self.plane))
        handles = exchange(self.block.handle)
        self.mapped = {}
        if rank == root:
            for r, h in enumerate(handles):
                self.mapped[r] = self.block.ptr if r == rank else nd.peer_open(h)

    def plane_ptr(self, base, group, ch):
        return base + (group * 2 + ch) * self.plane

    def stage(self, group_products):
        """copy this rank's group mixes (engine products: .p0 / .p1 planes, .frames) into the exported block"""
        import ctypes as C
        assert len(group_products) == self.groups_local
        for g, p in enumerate(group_products):
            assert p.frames == self.frames and p.fmt == nd.FMT_FLTP and p.rate == 48000, "group mixes are 48 kHz FLTP of one length"
            for ch, src in enumerate((p.p0, p.p1)):
                nd.check(nd.lib().nodey_memcpy_d2d(C.c_void_p(self.plane_ptr(self.block.ptr, g, ch)), C.c_void_p(src), self.frames * 4,
                                                   nd._stream()))

    def mix(self, out_l, out_r, total_frames, master_vol, sync):
        """sync(): wait for this rank's device work.  Every rank calls mix(); the root launches the kernel.
        out_l / out_r: device addresses of the bus planes on the root (ignored elsewhere)."""
        sync()
        self.barrier()                      # every rank's staged groups are complete and visible
        if self.rank == self.root:
            in_l, in_r = [], []
            for r in range(self.world):     # rank r holds groups [r * groups_local, (r + 1) * groups_local): the graph's input order
                for g in range(self.groups_local):
                    in_l.append(self.plane_ptr(self.mapped[r], g, 0))
                    in_r.append(self.plane_ptr(self.mapped[r], g, 1))
            n = len(in_l)
            nd.mix_ptrs(out_l, out_r, in_l, in_r, [self.frames] * n, [master_vol] * n, total_frames)
            sync()
        self.barrier()                      # the root has read everything: the blocks may be rewritten

    def close(self):
        for r, p in self.mapped.items():
            if r != self.rank:
                nd.peer_close(p)
        self.mapped = {}
        self.block.close()
