"""CPU: what audio_input publishes for RIFF/WAVE files, through the C ABI's header probe.

The reference decodes files with libavformat / libavcodec and pushes every decoded frame as it is
(src/processor/audio-io.cpp:176-222).  For WAV that fixes two things the nodes behind the source can see:
  * the sample format: pcm_s16le -> S16, pcm_s24le and pcm_s32le -> S32, pcm_f32le -> FLT (libavcodec/pcm.c);
  * the frame size: the wav demuxer reads packets of max_size = 4096 bytes rounded down to whole blocks
    (libavformat/wavdec.c, wav_read_packet) and the PCM decoder returns one frame per packet.
The expected values below restate that arithmetic independently of the C++ code."""
import struct

import numpy as np
import pytest

from helpers import FMT_FLT, FMT_S16, FMT_S32


@pytest.fixture(scope="module")
def eng():
    import engine
    engine.lib()
    return engine


def _fmt_chunk(tag, ch, rate, bits, extensible=False):
    block = ch * bits // 8
    if not extensible:
        return b"fmt " + struct.pack("<IHHIIHH", 16, tag, ch, rate, rate * block, block, bits)
    guid_tail = bytes.fromhex("000000001000800000aa00389b71")
    return (b"fmt " + struct.pack("<IHHIIHH", 40, 0xFFFE, ch, rate, rate * block, block, bits)
            + struct.pack("<HHI", 22, bits, 3 if ch == 2 else 4) + struct.pack("<H", tag) + guid_tail)


def _wav(path, tag, ch, rate, bits, frames, extensible=False, extra=b"", claim=None):
    data = np.random.default_rng(frames).integers(0, 256, frames * ch * bits // 8, dtype=np.uint8).tobytes()
    body = b"WAVE" + _fmt_chunk(tag, ch, rate, bits, extensible) + extra
    body += b"data" + struct.pack("<I", len(data) if claim is None else claim) + data
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)
    return data


def ffmpeg_packet_frames(bits, ch):
    """wav_read_packet: size = 4096; if block_align > 1: size = max(size, block_align) / block_align * block_align"""
    block = ch * bits // 8
    size = 4096
    if block > 1:
        size = max(size, block) // block * block
    return size // block


@pytest.mark.parametrize("tag,bits,ch,fmt", [(1, 16, 2, FMT_S16), (1, 16, 1, FMT_S16), (1, 24, 2, FMT_S32), (1, 24, 1, FMT_S32),
                                             (1, 32, 2, FMT_S32), (1, 32, 1, FMT_S32), (3, 32, 2, FMT_FLT), (3, 32, 1, FMT_FLT)])
def test_format_and_frame_size_follow_the_demuxer(eng, tmp_path, tag, bits, ch, fmt):
    path = str(tmp_path / "x.wav")
    _wav(path, tag, ch, 44100, bits, 5000)
    got = eng.probe_wav(path)
    assert got == (fmt, 44100, ch, 5000, ffmpeg_packet_frames(bits, ch))


def test_known_frame_sizes():
    # the figures the header comment quotes
    assert ffmpeg_packet_frames(16, 2) == 1024 and ffmpeg_packet_frames(16, 1) == 2048
    assert ffmpeg_packet_frames(32, 2) == 512 and ffmpeg_packet_frames(24, 2) == 682 and ffmpeg_packet_frames(24, 1) == 1365


def test_extensible_header_list_chunk_and_odd_padding(eng, tmp_path):
    path = str(tmp_path / "ext.wav")
    # a LIST chunk of odd size (padded to even) between fmt and data, WAVE_FORMAT_EXTENSIBLE with the float sub-format
    _wav(path, 3, 2, 48000, 32, 777, extensible=True, extra=b"LIST" + struct.pack("<I", 5) + b"abcde\0")
    assert eng.probe_wav(path) == (FMT_FLT, 48000, 2, 777, 512)


def test_truncated_data_chunk_counts_what_is_there(eng, tmp_path):
    path = str(tmp_path / "cut.wav")
    # header claims 4000 frames, the file holds 1000 (a recording that was cut off): the demuxer reads to the end of the file
    _wav(path, 1, 2, 22050, 16, 1000, claim=4000 * 4)
    assert eng.probe_wav(path) == (FMT_S16, 22050, 2, 1000, 1024)


def test_streamed_file_with_an_unpatched_data_size(eng, tmp_path):
    # writers that stream to a pipe leave 0 or 0xFFFFFFFF in the data chunk's size: the demuxer reads to the end of the file
    for k, claim in enumerate((0, 0xFFFFFFFF)):
        path = str(tmp_path / f"stream{k}.wav")
        _wav(path, 3, 1, 16000, 32, 3001, claim=claim)
        assert eng.probe_wav(path) == (FMT_FLT, 16000, 1, 3001, 1024)


def test_oversized_fmt_chunk_is_refused_without_allocating_it(eng, tmp_path):
    path = str(tmp_path / "huge.wav")
    open(path, "wb").write(b"RIFF" + struct.pack("<I", 100) + b"WAVEfmt " + struct.pack("<I", 0xFFFFFFF0) + b"\0" * 64)
    with pytest.raises(eng.EngineError) as x:
        eng.probe_wav(path)
    assert x.value.code == eng.E_FILE


@pytest.mark.parametrize("tag,bits,ch", [(1, 8, 2), (3, 64, 2), (1, 16, 6), (85, 16, 2)])
def test_files_the_nodes_could_not_process_are_refused_at_the_source(eng, tmp_path, tag, bits, ch):
    path = str(tmp_path / "no.wav")
    _wav(path, tag, ch, 44100, bits, 100)
    with pytest.raises(eng.EngineError) as x:
        eng.probe_wav(path)
    assert x.value.code == eng.E_FILE and "Cannot open audio file" in x.value.message


def test_missing_file_and_missing_data_chunk(eng, tmp_path):
    with pytest.raises(eng.EngineError) as x:
        eng.probe_wav(str(tmp_path / "nope.wav"))
    assert x.value.code == eng.E_FILE
    path = str(tmp_path / "nodata.wav")
    body = b"WAVE" + _fmt_chunk(1, 2, 44100, 16)
    open(path, "wb").write(b"RIFF" + struct.pack("<I", len(body)) + body)
    with pytest.raises(eng.EngineError) as x:
        eng.probe_wav(path)
    assert "no data chunk" in x.value.message
