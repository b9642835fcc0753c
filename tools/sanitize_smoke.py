"""Small invocation of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck).

    compute-sanitizer --tool memcheck  --error-exitcode 1 python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_smoke.py

No oracle, no comparison: the parity tests do that.  Sizes are the smallest that still reach every code path:
batched resampler (TMA-staged tile kernel, staging-row kernel, generic kernel incl. the interpolating plan), fused
resample + mix, SoundTouch offset search with clusters of 1 / 2 / 4 CTAs (stereo, interleaved and planar in place;
mono), fused and unfused tails, both pipeline orders, streaming kernels of every format, mixers, STFT, and the graph
through the plugin API (levels, waves, lanes).  FAMILY=name runs one family only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import nodey as nd  # noqa: E402


def fam_stream():
    x = nd.synth(5003, 2, 44100, track=1)
    s16 = nd.synth(5003, 2, 44100, track=2, want_s16=True)[1]
    s32 = (x.double() * 2147483647.0).round().clamp(-2147483648, 2147483647).to(torch.int32)
    for t, fmt in ((x, nd.FMT_FLT), (s16, nd.FMT_S16), (s32, nd.FMT_S32)):
        nd.gain(t, fmt, 0.8)
        nd.extract_interleaved(t, fmt)
        nd.split(t, fmt)
        nd.to_fltp_stereo(t, fmt)
        tp = t.t().contiguous()
        nd.gain(tp, fmt + 5, 0.8)
        nd.extract_interleaved(tp, fmt + 5)
        nd.split(tp, fmt + 5)
    planes = [nd.to_fltp_stereo(nd.synth(4099, 2, 48000, track=k), nd.FMT_FLT) for k in range(3)]
    nd.mix(planes, [0.5, 0.25, 1.0])
    nd.bimix(planes[0], planes[1], 0.3)
    nd.downmix_half(planes[0])


def fam_resample():
    for rate, n in ((44100, 9000), (22050, 5000), (96000, 9000), (47999, 6000), (8000, 2000)):
        x = nd.synth(n, 2, rate, track=3)
        r = nd.Resampler(rate, 48000)
        for mode in (0, 1, 2, 3, 4):
            try:
                r.run(x, nd.FMT_FLT, flush=True, mode=mode)
            except nd.NodeyError:
                pass            # a mode the plan has no kernel for
        r.close()
    r = nd.Resampler(44100, 48000)
    xs = [nd.synth(7000 + 300 * k, 2, 44100, track=k) for k in range(4)]
    r.resample_mix(xs, [nd.FMT_FLT] * 4, [0.5, 0.25, 1.0, 0.75])
    s16 = nd.synth(6000, 1, 44100, track=5, want_s16=True)[1]
    r.run(s16, nd.FMT_S16)
    r.run(nd.synth(40, 2, 44100, track=6), nd.FMT_FLT)
    r.close()


def fam_soundtouch():
    x = torch.stack([nd.synth(48000, 2, 48000, track=k) for k in range(3)])
    for cluster in (1, 2, 4, 8):
        st = nd.SoundTouch.pitch_node(48000, 2, 3.0)
        st.set_cluster(cluster)
        st.run(x, want_offsets=True)
        st.close()
    for kt, cluster in ((8, 1), (12, 2), (16, 1), (13, 4), (4, 4), (2, 8), (2, 1), (4, 8)):     # every code shape of the search: full / partial last group, rotating / two-group bodies
        st = nd.SoundTouch.pitch_node(48000, 2, 3.0)
        st.set_candidates_per_thread(kt)
        st.set_cluster(cluster)
        st.run(x, want_offsets=True)
        st.close()
    os.environ["NODEY_TDS_RUNTIME_SK"] = "1"                          # run-time sub-plane stride
    st = nd.SoundTouch.pitch_node(48000, 2, 3.0)
    st.run(x)
    st.close()
    os.environ.pop("NODEY_TDS_RUNTIME_SK")
    st = nd.SoundTouch.velocity_node(48000, 2, 1.25, True)
    st.run(x)
    st.set_unfused(1)
    st.run(x)
    st.close()
    st = nd.SoundTouch.pitch_node(48000, 2, -4.0)           # rate <= 1: cubic -> FIR -> TDStretch
    st.run(x)
    st.close()
    planes = [(x[k, :, 0].contiguous(), x[k, :, 1].contiguous()) for k in range(3)]
    st = nd.SoundTouch.pitch_node(48000, 2, 3.0)
    st.run_tracks(planes)
    st.close()
    st = nd.SoundTouch.pitch_node(44100, 1, 2.0)
    st.run(nd.synth(44100, 1, 44100, track=9))
    st.close()
    st = nd.SoundTouch.pitch_node(48000, 2, 3.0)            # shorter than one sequence
    st.run(nd.synth(2000, 2, 48000, track=10))
    st.close()


def fam_stft():
    nd.stft(nd.synth(4096 * 5 + 77, 2, 48000, track=11), True)
    nd.stft(nd.to_fltp_stereo(nd.synth(4096 * 3, 2, 48000, track=12), nd.FMT_FLT), False)


def fam_engine():
    import engine
    import pipeline as G                      # track_gain: host arithmetic
    n = 44100
    os.environ["NODEY_WAVE"] = "8"           # 16 tracks in two waves on rotating lanes
    project, ids = engine.config5_project(16, [G.track_gain(t) for t in range(16)])
    eng = engine.Engine(project.json())
    host = [nd.synth(n, 2, 44100, track=t).cpu().numpy() for t in range(16)]
    for t in range(16):
        eng.bind_source(t, host[t], nd.FMT_FLT, 44100)
    eng.run()
    eng.output().numpy()
    eng.close()
    del os.environ["NODEY_WAVE"]


FAMILIES = {"stream": fam_stream, "resample": fam_resample, "soundtouch": fam_soundtouch, "stft": fam_stft,
            "engine": fam_engine}

if __name__ == "__main__":
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    nd.lib()
    only = os.environ.get("FAMILY")
    for name, fn in FAMILIES.items():
        if only and name != only:
            continue
        before = nd.profile_launches()
        fn()
        torch.cuda.synchronize()
        print(f"{name}: {nd.profile_launches() - before} launches", flush=True)
    print("sanitize_smoke done")
