"""Guard-band checks of every kernel that writes through a caller-supplied pointer (compute-sanitizer is closed on
the GPU pool, so out-of-bounds WRITES are caught here): the output lives in the middle of a larger buffer filled with
a canary bit pattern, and after the launch every word outside the region the C ABI promises to write must still be
the canary.  Sizes are deliberately ragged (not multiples of a vector, a tile or a warp) because that is where the
vectorised tails, the TMA-staged tiles and the per-track strides can overrun.  Results inside the region are covered
by the parity tests; here only the boundaries matter."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CANARY = 0x7FC0DEAD          # a quiet NaN no kernel produces
PAD = 4096                   # guard words on either side


class Guarded:
    """a tensor view of `shape` / `dtype` with PAD canary words (32-bit) before and after it"""

    def __init__(self, shape, dtype=torch.float32, device="cuda"):
        n = int(np.prod(shape))
        words = n * torch.empty(0, dtype=dtype).element_size() // 4
        assert words * 4 == n * torch.empty(0, dtype=dtype).element_size()
        self.words = words
        self.raw = torch.full((words + 2 * PAD,), CANARY, dtype=torch.int32, device=device)
        self.view = self.raw[PAD:PAD + words].view(dtype).view(*shape)

    def check(self, what, written_words=None):
        torch.cuda.synchronize()
        lo = self.raw[:PAD]; hi = self.raw[PAD + self.words:]
        assert bool((lo == CANARY).all()), f"{what}: wrote BEFORE its output"
        assert bool((hi == CANARY).all()), f"{what}: wrote PAST its output"
        if written_words is not None:       # part of the view that must stay untouched as well
            rest = self.raw[PAD + written_words:PAD + self.words]
            assert bool((rest == CANARY).all()), f"{what}: wrote past the promised {written_words} words inside the view"


@pytest.mark.parametrize("n", [1, 3, 5, 127, 1001, 4099])
def test_gain_writes_exactly_its_elements(nd, n):
    for fmt, dt in ((nd.FMT_FLT, torch.float32), (nd.FMT_S32, torch.int32)):
        src = (torch.arange(n * 2, device="cuda") % 97).to(dt).view(n, 2)
        g = Guarded((n, 2), dt)
        nd.gain(src, fmt, 0.5, out=g.view)
        g.check(f"gain fmt {fmt} n {n}")
    # 16-bit stereo frames are whole 32-bit words
    src = (torch.arange(n * 2, device="cuda") % 97).to(torch.int16).view(n, 2)
    g = Guarded((n, 2), torch.int16)
    nd.gain(src, nd.FMT_S16, 0.5, out=g.view)
    g.check(f"gain s16 n {n}")


@pytest.mark.parametrize("rate", [44100, 22050, 96000, 47999])
@pytest.mark.parametrize("n", [40, 1000, 4705, 9001])
def test_resampler_writes_exactly_out_frames(nd, rate, n):
    x = nd.synth(n, 2, rate, track=1)
    r = nd.Resampler(rate, 48000)
    for flush in (True, False):
        m = r.out_count(n, flush)
        if m == 0:
            continue
        for mode in (0, 1, 2, 3, 4):
            g = Guarded((2, m))
            try:
                r.run(x, nd.FMT_FLT, flush=flush, mode=mode, out=g.view)
            except nd.NodeyError:
                continue        # the plan has no kernel for this mode
            g.check(f"resample {rate} n {n} flush {flush} mode {mode}")
    r.close()


@pytest.mark.parametrize("nin", [1, 3, 16])
def test_resample_mix_writes_exactly_out_frames(nd, nin):
    r = nd.Resampler(44100, 48000)
    xs = [nd.synth(5000 + 333 * k, 2, 44100, track=k) for k in range(nin)]
    m = max(r.out_count(int(x.shape[0]), True) for x in xs)
    g = Guarded((2, m))
    r.resample_mix(xs, [nd.FMT_FLT] * nin, [1.0 / nin] * nin, out=g.view)
    g.check(f"resample_mix {nin}")
    r.close()


@pytest.mark.parametrize("n", [4096, 4096 + 1023, 4096 * 3 + 1, 20000])
@pytest.mark.parametrize("interleaved", [True, False])
def test_stft_writes_exactly_its_frames(nd, n, interleaved):
    x = nd.synth(n, 2, 48000, track=2)
    if not interleaved:
        x = x.t().contiguous()
    m = nd.stft_frames(n)
    g = Guarded((2, m, 2049), torch.complex64)
    nd.stft(x, interleaved, out=g.view)
    g.check(f"stft n {n}")


@pytest.mark.parametrize("cluster", [1, 2, 4])
@pytest.mark.parametrize("node", ["pitch+3", "tempo1.25", "pitch-4"])
def test_soundtouch_writes_exactly_its_frames(nd, cluster, node):
    ntr, n = 3, 30011
    x = torch.stack([nd.synth(n, 2, 48000, track=k) for k in range(ntr)])
    st = {"pitch+3": lambda: nd.SoundTouch.pitch_node(48000, 2, 3.0),
          "tempo1.25": lambda: nd.SoundTouch.velocity_node(48000, 2, 1.25, True),
          "pitch-4": lambda: nd.SoundTouch.pitch_node(48000, 2, -4.0)}[node]()
    st.set_cluster(cluster)
    m, _ = st.out_frames(n)
    for unfused in (0, 1):
        st.set_unfused(unfused)
        g = Guarded((ntr, m, 2))
        st.run(x, out=g.view)
        g.check(f"soundtouch {node} cluster {cluster} unfused {unfused}")
    st.close()


def test_soundtouch_mono_and_short_inputs(nd):
    for n in (100, 2000, 5000, 44100):
        x = nd.synth(n, 1, 44100, track=3)
        st = nd.SoundTouch.pitch_node(44100, 1, 2.0)
        m, _ = st.out_frames(n)
        if m > 0:
            g = Guarded((1, m, 1))
            st.run(x.unsqueeze(0), out=g.view)
            g.check(f"soundtouch mono n {n}")
        st.close()
