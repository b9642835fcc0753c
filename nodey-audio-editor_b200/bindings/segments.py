"""Time-segment sharding of ONE long stream across ranks: which part of the input a rank needs
and which outputs it owns, for every node class that can be cut in time.

  streaming nodes (gain, split, merge, mix, format conversion)   cut anywhere, no halo
  polyphase resampler (audio_amix / audio_bimix front end)        start >= 15 input frames (whole periods) early, drop the lead-in
  audio_spectrum (STFT 4096 / 1024)                               halo nfft - hop = 3072 frames on the right
  SoundTouch nodes                                                NOT cuttable: WSOLA / transposer state is sequential

Everything here is host arithmetic on lengths (no device, no oracle): a rank only needs its input slice, the
results are bit identical to the whole-stream render, and no collective is involved -- the segments are simply
concatenated (or written to disjoint ranges of the output file).  The C ABI states the same rule for the
resampler (nodey_resampler_segment); tests/test_gpu_segments.py checks that both agree.
"""


def split_even(total, world, rank, align=1):
    """contiguous near-equal split of range(total) into `world` parts whose inner boundaries are multiples of align"""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world / rank")
    units = -(-total // align)
    lo = (units * rank // world) * align
    hi = total if rank == world - 1 else (units * (rank + 1) // world) * align
    return min(lo, total), min(hi, total)


def stream_segment(n_frames, world, rank):
    """gain / split / merge / mix: frames [a, b) of every input give frames [a, b) of the output"""
    return split_even(n_frames, world, rank)


def stft_segment(n_frames, world, rank, nfft=4096, hop=1024):
    """spectrum frames [m0, m1) need input frames [in0, in1) = [m0*hop, (m1-1)*hop + nfft)"""
    frames = 0 if n_frames < nfft else (n_frames - nfft) // hop + 1
    m0, m1 = split_even(frames, world, rank)
    if m1 <= m0:
        return dict(m0=m0, m1=m0, in0=0, in1=0)
    return dict(m0=m0, m1=m1, in0=m0 * hop, in1=(m1 - 1) * hop + nfft)


def resample_segment(plan, n_in, total_out, k0, k1):
    """mirror of nodey_resampler_segment.  plan: the resampler's info dict (phase_count, filter_length,
    dst_incr_div, dst_incr_mod, index0).  Returns in0, in1, skip, flush: run the resampler on input frames
    [in0, in1) with `flush`, take skip + (k1 - k0) outputs, drop the first skip."""
    P, L, D = plan["phase_count"], plan["filter_length"], plan["dst_incr_div"]
    if plan["dst_incr_mod"] != 0:
        raise ValueError("only exact-rational plans can be cut into segments")
    if not (0 <= k0 <= k1 <= total_out) or k0 % P:
        raise ValueError("segment must start on a multiple of the phase count and lie inside the stream")
    center = (L - 1) // 2
    # lead-in: whole periods covering at least `center` input frames, so that no kept output sees the mirrored left edge
    lead = max(1, -(-center // D))
    periods = k0 // P
    in0 = 0 if periods <= lead else (periods - lead) * D
    skip = k0 if periods <= lead else lead * P
    in1, flush = n_in, True
    if k0 < k1 < total_out:
        s_last = (plan["index0"] + (skip + (k1 - k0) - 1) * D) // P
        need = max(s_last - center + L, L + 1)
        if in0 + need <= n_in:
            in1, flush = in0 + need, False
    # a conversion cannot start before filter_length + 1 frames are there (the initial mirror): a slice that runs to the
    # end of a short stream and is shorter than that starts more whole periods early (tiny streams cut into many segments)
    if in0 > 0 and in1 - in0 < L + 1:
        lead += -(-((L + 1) - (in1 - in0)) // D)
        in0 = 0 if periods <= lead else (periods - lead) * D
        skip = k0 if periods <= lead else lead * P
    return dict(in0=in0, in1=in1, skip=skip, flush=flush)


def resample_ranges(plan, n_in, total_out, world, rank):
    """outputs [k0, k1) of `rank` (period aligned) and the recipe to compute them"""
    k0, k1 = split_even(total_out, world, rank, align=plan["phase_count"])
    seg = resample_segment(plan, n_in, total_out, k0, k1)
    seg.update(k0=k0, k1=k1)
    return seg


# ---- execution through the C ABI (device tensors) -------------------------------------------------
def run_resample_segment(rs, x, fmt, n_in, total_out, world, rank):
    """x: the whole stream in HBM (only the slice is read) -> [2, k1-k0] of the rank's outputs"""
    seg = resample_ranges(rs.info(), n_in, total_out, world, rank)
    if seg["k1"] <= seg["k0"]:
        return seg, None
    xs = x[seg["in0"]:seg["in1"]] if x.dim() == 2 and fmt < 5 else x[..., seg["in0"]:seg["in1"]]
    out = rs.run(xs.contiguous(), fmt, flush=seg["flush"], out_frames=seg["skip"] + seg["k1"] - seg["k0"])
    return seg, out[:, seg["skip"]:]


def run_stft_segment(nd, x_planar, world, rank, nfft=4096, hop=1024):
    """x_planar: [nch, n] -> complex64 [nch, m1-m0, nfft/2+1] of the rank's frames"""
    seg = stft_segment(x_planar.shape[1], world, rank, nfft, hop)
    if seg["m1"] <= seg["m0"]:
        return seg, None
    return seg, nd.stft(x_planar[:, seg["in0"]:seg["in1"]].contiguous(), False, nfft, hop)
