"""WAV ingest (whisperseg_b200/audio_io.py) against libsndfile/librosa semantics: PCM scaled by 1/2**(bits-1),
channels averaged, float32 -- the values `librosa.load(path, sr=None)` hands the reference at
scripts/segment.py:50.  Integer decode is bit-exact."""
import glob
import io
import os
import struct

import numpy as np
import pytest

from whisperseg_b200 import audio_io


def _wav_bytes(samples, sr, width, channels, tag=1, extensible=False):
    """Minimal RIFF writer (interleaved `samples` already encoded as bytes)."""
    block = width * channels
    if extensible:
        fmt = struct.pack("<HHIIHHHHIH14s", 0xFFFE, channels, sr, sr * block, block, width * 8, 22, width * 8, 0, tag,
                          b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71")
    else:
        fmt = struct.pack("<HHIIHH", tag, channels, sr, sr * block, block, width * 8)
    junk = b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\x00"          # odd-sized chunk + pad byte before the data
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + junk + b"data" + struct.pack("<I", len(samples)) + samples
    return b"RIFF" + struct.pack("<I", len(body)) + body


@pytest.mark.parametrize("channels", [1, 2, 3])
def test_pcm16_bit_exact(channels, tmp_path):
    rng = np.random.default_rng(channels)
    x = rng.integers(-32768, 32768, size=(5000, channels), dtype=np.int64).astype("<i2")
    x[0, 0], x[1, 0] = -32768, 32767
    data = _wav_bytes(x.tobytes(), 22050, 2, channels)
    path = tmp_path / "a.wav"
    path.write_bytes(data)
    for src in (str(path), data, io.BytesIO(data)):
        audio, sr = audio_io.load_audio(src)
        assert sr == 22050 and audio.dtype == np.float32 and audio.shape == (5000,)
        ref = (x.astype(np.float32) / 32768.0).T
        ref = ref[0] if channels == 1 else np.mean(ref, axis=0)
        assert np.array_equal(audio, ref)
    multi, _ = audio_io.load_audio(data, mono=False)
    assert multi.shape == ((5000,) if channels == 1 else (channels, 5000))
    if channels > 1:
        assert np.array_equal(multi[1], x[:, 1].astype(np.float32) / 32768.0)      # segment_service.py:79 channel_id


def test_other_sample_formats():
    rng = np.random.default_rng(7)
    n = 3000
    u8 = rng.integers(0, 256, size=n, dtype=np.uint8)
    a, _ = audio_io.load_audio(_wav_bytes(u8.tobytes(), 8000, 1, 1))
    assert np.array_equal(a, (u8.astype(np.float32) - 128.0) / 128.0)
    i24 = rng.integers(-(1 << 23), 1 << 23, size=n, dtype=np.int64)
    raw = b"".join(int(v & 0xFFFFFF).to_bytes(3, "little") for v in i24)
    a, _ = audio_io.load_audio(_wav_bytes(raw, 48000, 3, 1, extensible=True))
    assert np.array_equal(a, i24.astype(np.float32) / 8388608.0)
    i32 = rng.integers(-(1 << 31), 1 << 31, size=n, dtype=np.int64).astype("<i4")
    a, _ = audio_io.load_audio(_wav_bytes(i32.tobytes(), 48000, 4, 1))
    assert np.array_equal(a, (i32.astype(np.float64) / 2147483648.0).astype(np.float32))
    f32 = rng.standard_normal(n).astype("<f4")
    a, _ = audio_io.load_audio(_wav_bytes(f32.tobytes(), 16000, 4, 1, tag=3))
    assert np.array_equal(a, f32)
    with pytest.raises(ValueError):
        audio_io.load_audio(b"OggS" + bytes(100))
    with pytest.raises(ValueError):
        audio_io.load_audio(_wav_bytes(bytes(64), 8000, 2, 1, tag=0x55))


def test_resample_is_band_limited_and_length_correct():
    sr_in, sr_out = 48000, 16000
    t = np.arange(sr_in) / sr_in
    x = (0.5 * np.sin(2 * np.pi * 1000 * t)).astype(np.float32)
    data = _wav_bytes((x * 32767).astype("<i2").tobytes(), sr_in, 2, 1)
    y, sr = audio_io.load_audio(data, sr=sr_out)
    assert sr == sr_out and len(y) == sr_out
    ref = 0.5 * np.sin(2 * np.pi * 1000 * np.arange(sr_out) / sr_out)
    assert np.abs(y[200:-200] - ref[200:-200]).max() < 2e-3


def test_matches_scipy_on_reference_example_wavs():
    wavs = sorted(glob.glob("/root/reference/data/example_subset/*/*/*.wav"))[:6]
    if not wavs:
        pytest.skip("reference example recordings not present on this box")
    from scipy.io import wavfile
    for p in wavs:
        sr_ref, x = wavfile.read(p)
        audio, sr = audio_io.load_audio(p)
        assert sr == sr_ref
        if x.dtype == np.int16:
            ref = x.astype(np.float32) / 32768.0
        elif x.dtype == np.int32:
            ref = (x.astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            ref = x.astype(np.float32)
        if ref.ndim == 2:
            ref = np.mean(ref.T, axis=0)
        assert np.array_equal(audio, ref), p


class _FakeSegmenter:
    def __init__(self):
        self.calls = []

    def segment_many(self, audios, sr, **kw):
        self.calls.append((len(audios), sr, kw))
        return [{"onset": [0.0] * (len(a) // 8000), "offset": [0.1] * (len(a) // 8000), "cluster": ["x"] * (len(a) // 8000)}
                for a in audios]


def test_segment_files_groups_by_rate_and_builds_the_table(tmp_path):
    paths = []
    for i, (sr, n) in enumerate([(16000, 16000), (16000, 24000), (32000, 8000), (16000, 40000)]):
        p = tmp_path / ("f%d.wav" % i)
        p.write_bytes(_wav_bytes(np.zeros(n, "<i2").tobytes(), sr, 2, 1))
        paths.append(str(p))
    seg = _FakeSegmenter()
    per_file, table = audio_io.segment_files(seg, paths, workers=2, num_trials=1, num_beams=1)
    assert [c[:2] for c in seg.calls] == [(2, 16000), (1, 32000), (1, 16000)]
    assert seg.calls[0][2] == {"num_trials": 1, "num_beams": 1}
    assert list(per_file) and all(p in per_file for p in paths)
    assert table["filename"] == ["f0.wav"] * 2 + ["f1.wav"] * 3 + ["f2.wav"] * 1 + ["f3.wav"] * 5
    assert len(table["onset"]) == len(table["offset"]) == len(table["cluster"]) == 11
    seg2 = _FakeSegmenter()
    audio_io.segment_files(seg2, paths[:2], group_seconds=1.2)
    assert [c[0] for c in seg2.calls] == [1, 1]
    assert audio_io.segment_files(seg2, []) == ({}, {"filename": [], "onset": [], "offset": [], "cluster": []})


def test_wav_info_matches_decode(tmp_path):
    """Header-only length probe (what every rank reads of every file in sharded folder mode) against the decoder."""
    rng = np.random.default_rng(11)
    cases = [(2, 1, 16000, 12345, False), (2, 2, 22050, 777, False), (3, 1, 48000, 1001, True), (4, 3, 32000, 64, False),
             (1, 1, 8000, 0, False)]
    for i, (width, ch, sr, frames, ext) in enumerate(cases):
        raw = rng.integers(0, 256, size=frames * ch * width, dtype=np.uint8).tobytes()
        path = tmp_path / ("c%d.wav" % i)
        path.write_bytes(_wav_bytes(raw, sr, width, ch, extensible=ext))
        assert audio_io.wav_info(str(path)) == (frames, sr, ch)
        assert audio_io.wav_info(str(path), head_bytes=40) == (frames, sr, ch)      # data chunk beyond the first read
        audio, sr2 = audio_io.load_audio(str(path))
        assert (len(audio), sr2) == (frames, sr)
    trunc = tmp_path / "trunc.wav"                                                   # header announces more than is there
    trunc.write_bytes(_wav_bytes(b"\x00" * 2000, 16000, 2, 1)[:-500])
    assert audio_io.wav_info(str(trunc))[0] == len(audio_io.load_audio(str(trunc))[0]) == 750
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.wav"
        bad.write_bytes(b"OggS" + b"\x00" * 64)
        audio_io.wav_info(str(bad))
