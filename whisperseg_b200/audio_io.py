"""Audio ingest for the segmenter: WAV decode, channel handling, optional
resampling, and a folder runner that decodes files on host threads while the GPU segments.

Replaces, for WAV input, the `librosa.load` calls either side of the hot path:
  * reference scripts/segment.py:50,60,62  `librosa.load(path_or_buffer, sr=None)`  -> `load_audio(src)`:
    native sample rate, PCM scaled by 1 / 2**(bits-1) (what libsndfile hands librosa), channels averaged
    (`librosa.to_mono`), float32 -- bit-identical for PCM 16/24/32 and float WAV;
  * reference segment_service.py:76-80  `librosa.load(buf, sr=sr, mono=False)` + `audio[channel_id]`
    -> `load_audio(src, sr=sr, mono=False)[0][channel_id]`.  Resampling here is polyphase
    (`scipy.signal.resample_poly`), NOT librosa's soxr_hq: equal band-limited signal, not bit-identical
    samples -- resample upstream if exact agreement with the reference is required;
  * reference scripts/segment.py:41-56 (folder loop, one segment() call per file) -> `segment_files`:
    files are decoded by a small thread pool and handed to `segment_many` in groups, so decode, H2D and
    GPU work overlap and short clips share batches.
  * the same folder on several GPUs (BASELINE configs[4]) -> `segment_files_sharded`: every rank reads the WAV
    headers only (`wav_info`), decodes the clips of its own shard and joins one all-gather per sample-rate group.
Only RIFF/WAVE is handled (the reference globs *.wav / *.WAV); anything else raises ValueError.
"""
import io
import os
import struct
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_PCM, _FLOAT, _EXTENSIBLE = 0x0001, 0x0003, 0xFFFE


def _read_bytes(src):
    if isinstance(src, (bytes, bytearray, memoryview)):
        return bytes(src)
    if hasattr(src, "read"):
        return src.read()
    with open(os.fspath(src), "rb") as f:
        return f.read()


def _wav_layout(data, total_size=None):
    """Parse the RIFF chunk list: ((format tag, channels, sample rate, block align, bits), (data offset, data bytes)).
    `data` may be just the head of the file when `total_size` (the file size) is given."""
    if len(data) < 12 or data[:4] not in (b"RIFF", b"RF64") or data[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE stream")
    end = len(data) if total_size is None else total_size
    pos, fmt, payload = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
        body = pos + 8
        if cid == b"fmt ":
            if body + 16 > len(data):
                break
            tag, ch, sr, _, align, bits = struct.unpack_from("<HHIIHH", data, body)
            if tag == _EXTENSIBLE and size >= 26 and body + 26 <= len(data):
                tag = struct.unpack_from("<H", data, body + 24)[0]
            fmt = (tag, ch, sr, align, bits)
        elif cid == b"data":
            if size == 0xFFFFFFFF or body + size > end:            # streamed / truncated files: take what is there
                size = end - body
            payload = (body, size)
            break
        pos = body + size + (size & 1)
    if fmt is None or payload is None:
        raise ValueError("WAVE stream without fmt/data chunk")
    return fmt, payload


def wav_info(path, head_bytes=1 << 16):
    """(frames, sample_rate, channels) from the header of a WAV file, without decoding it (folder mode on several
    GPUs: every rank needs every clip's length for the window plan, but decodes only its own clips)."""
    path = os.fspath(path)
    total = os.path.getsize(path)
    with open(path, "rb") as f:
        head = f.read(head_bytes)
    try:
        fmt, (body, size) = _wav_layout(head, total)
    except ValueError:
        if total <= len(head):
            raise
        with open(path, "rb") as f:                                # a long chunk list before the data: parse the whole file
            fmt, (body, size) = _wav_layout(f.read())
    _, ch, sr, _, bits = fmt
    width = bits // 8
    if ch < 1 or width < 1:
        raise ValueError("bad WAVE header")
    return size // (width * ch), int(sr), int(ch)


def decode_wav(data):
    """RIFF/WAVE bytes -> (float32 array [channels, frames], sample_rate)."""
    fmt, payload = _wav_layout(data)
    tag, ch, sr, align, bits = fmt
    body, size = payload
    width = bits // 8
    if ch < 1 or width < 1:
        raise ValueError("bad WAVE header")
    n = size // (width * ch)
    raw = np.frombuffer(data, dtype=np.uint8, count=n * ch * width, offset=body)
    if tag == _PCM:
        if width == 1:
            x = (raw.astype(np.float32) - 128.0) / 128.0
        elif width == 2:
            x = raw.view("<i2").astype(np.float32) / 32768.0
        elif width == 3:
            b = raw.reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            v = np.where(v & 0x800000, v - 0x1000000, v)
            x = v.astype(np.float32) / 8388608.0
        elif width == 4:
            x = (raw.view("<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            raise ValueError("unsupported PCM width %d" % bits)
    elif tag == _FLOAT:
        if width == 4:
            x = raw.view("<f4").astype(np.float32)
        elif width == 8:
            x = raw.view("<f8").astype(np.float32)
        else:
            raise ValueError("unsupported float width %d" % bits)
    else:
        raise ValueError("unsupported WAVE format tag 0x%04x" % tag)
    return np.ascontiguousarray(x.reshape(n, ch).T), int(sr)


def resample(audio, sr_in, sr_out):
    """Polyphase resampling along the last axis (see the module docstring for the parity note)."""
    if sr_in == sr_out:
        return audio
    from math import gcd
    from scipy.signal import resample_poly
    g = gcd(int(sr_in), int(sr_out))
    return resample_poly(audio, int(sr_out) // g, int(sr_in) // g, axis=-1).astype(np.float32)


def load_audio(src, sr=None, mono=True):
    """`librosa.load(src, sr=sr, mono=mono)` for WAV input: (float32 audio, sample_rate).
    `src`: path, bytes or binary file object.  mono=True averages the channels ([frames]); mono=False keeps
    [channels, frames] for multi-channel files and [frames] for single-channel ones, like librosa."""
    x, native = decode_wav(_read_bytes(src))
    if mono:
        x = x[0] if x.shape[0] == 1 else np.mean(x, axis=0, dtype=np.float32)
    elif x.shape[0] == 1:
        x = x[0]
    if sr is not None and int(sr) != native:
        x = resample(x, native, int(sr))
        native = int(sr)
    return np.ascontiguousarray(x, dtype=np.float32), native


def segment_files(segmenter, paths, workers=4, group_seconds=1800.0, **segment_kwargs):
    """Folder mode (reference scripts/segment.py:41-56).  Returns (per_file, table): `per_file` maps every
    path to its prediction dict; `table` is the {"filename", "onset", "offset", "cluster"} column dict the
    reference builds its DataFrame from.  Files are decoded `workers` at a time ahead of the GPU; files of
    equal sample rate are segmented together (`segment_many`) in groups of at most `group_seconds` of audio."""
    paths = list(paths)
    per_file = {}
    table = {"filename": [], "onset": [], "offset": [], "cluster": []}
    if not paths:
        return per_file, table

    def flush(group, sr):
        if not group:
            return
        preds = segmenter.segment_many([a for _, a in group], sr, **segment_kwargs)
        for (path, _), pred in zip(group, preds):
            per_file[path] = pred

    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        # bounded decode-ahead: at most 2 x workers files are decoded (or being decoded) beyond the one being consumed,
        # so host memory holds one group plus that window -- not the whole folder (the reference streams file by file,
        # scripts/segment.py:48-55)
        ahead = max(2, 2 * max(1, workers))
        pending = {}
        next_submit = 0
        group, group_sr, seconds = [], None, 0.0
        for k, path in enumerate(paths):
            while next_submit < len(paths) and next_submit < k + ahead:
                pending[next_submit] = pool.submit(load_audio, paths[next_submit])
                next_submit += 1
            audio, sr = pending.pop(k).result()
            if group and (sr != group_sr or seconds + len(audio) / sr > group_seconds):
                flush(group, group_sr)
                group, seconds = [], 0.0
            group.append((path, audio))
            group_sr = sr
            seconds += len(audio) / sr
        flush(group, group_sr)
    for path in paths:
        pred = per_file[path]
        table["filename"] += [os.path.basename(path)] * len(pred["onset"])
        table["onset"] += pred["onset"]
        table["offset"] += pred["offset"]
        table["cluster"] += pred["cluster"]
    return per_file, table


def segment_files_sharded(segmenter, paths, workers=4, group=None, generate_fn=None, **segment_kwargs):
    """`segment_files` under torch.distributed (BASELINE configs[4]: a folder on all GPUs of a box): same
    (per_file, table) on every rank.  Every rank reads only the WAV *headers* of all files (clip lengths fix the
    window plan and the shards), decodes just the clips its shard of the flattened window list touches, and the
    token ids are exchanged with one all-gather per sample-rate group (`distributed.segment_many_sharded`).
    Files are segmented at their native sample rate, like the reference CLI (`librosa.load(path, sr=None)`)."""
    from .distributed import LazyClips, segment_many_sharded
    paths = list(paths)
    per_file = {}
    table = {"filename": [], "onset": [], "offset": [], "cluster": []}
    if not paths:
        return per_file, table
    infos = [wav_info(p) for p in paths]
    by_rate = {}
    for i, (_, sr, _) in enumerate(infos):
        by_rate.setdefault(sr, []).append(i)
    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        for sr, idx in by_rate.items():
            clips = LazyClips([infos[i][0] for i in idx], lambda k, _idx=idx: load_audio(paths[_idx[k]])[0], pool)
            preds = segment_many_sharded(segmenter, clips, sr, group=group, generate_fn=generate_fn, **segment_kwargs)
            for i, pred in zip(idx, preds):
                per_file[paths[i]] = pred
    for path in paths:
        pred = per_file[path]
        table["filename"] += [os.path.basename(path)] * len(pred["onset"])
        table["onset"] += pred["onset"]
        table["offset"] += pred["offset"]
        table["cluster"] += pred["cluster"]
    return per_file, table
