"""N>1 host logic on CPU: world_size-2 gloo run of the window sharding + single all-gather.  The
generation step is a scripted stand-in (tokens are a deterministic function of the window index), so
the test pins sharding, padding, gather order and rank-identical post-processing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from whisperseg_b200 import postprocess as pp
from whisperseg_b200.distributed import folder_window_table, segment_many_sharded, segment_sharded, shard_bounds
from whisperseg_b200.frontend import FrontendPlan, get_n_fft_given_sr

SR, STS, MAX_LEN = 16000, 0.01, 40
BOOK = {"vocal": 0, "b": 1}
TS0, EOT, PROMPT = 50364, 50257, [50258, 50259, 50363]


class _Tok:
    prompt_ids = PROMPT
    eos_token_id = pad_token_id = EOT

    def batch_decode(self, rows, skip_special_tokens=False):
        out = []
        for row in rows:
            s = ""
            for t in row:
                s += "<|%d|>" % (t - TS0) if t >= TS0 else ("<|endoftext|>" if t == EOT else str(t - 15))
            out.append(s)
        return out


class _Seg:
    default_segmentation_config = {}
    total_spec_columns = 1000
    cluster_codebook = BOOK
    precision_bits = 3
    tokenizer = _Tok()


def _tokens_for(global_index, n_new):
    rng = np.random.default_rng(1000 + global_index)
    row, t = [], 0
    while len(row) + 3 <= n_new - 1:
        t += int(rng.integers(5, 80))
        d = int(rng.integers(2, 40))
        if t + d > 500:
            break
        row += [TS0 + t, 15 + int(rng.integers(0, 2)), TS0 + t + d]
        t += d
    row.append(EOT)
    return row + [EOT] * (n_new - len(row))


def _expected(audio, num_trials):
    plan = FrontendPlan(SR, STS, 0)
    wins = plan.windows(len(audio), num_trials)
    texts = _Tok().batch_decode([_tokens_for(i, MAX_LEN - 3) for i in range(len(wins))])
    pred = pp.parse_generation(texts, [w.as_tuple() for w in wins], STS * 2, len(audio) / SR, STS, num_trials, STS * 8, STS,
                               "clustering", BOOK)
    return pp.correct_fft_blur_and_dedupe(pred, SR, get_n_fft_given_sr(SR))


def _worker(rank, world, port, n_samples, num_trials, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    audio = np.zeros(n_samples, dtype=np.float32)
    plan = FrontendPlan(SR, STS, 0)
    n_win = len(plan.windows(n_samples, num_trials))
    lo, hi, per = shard_bounds(n_win, world, rank)
    seen = {}

    def gen(windows, plan_, piece, slice_start):
        seen["n"] = len(windows)
        seen["slice"] = (slice_start, len(piece))
        return torch.tensor([_tokens_for(lo + i, MAX_LEN - 3) for i in range(len(windows))], dtype=torch.int32).reshape(-1, MAX_LEN - 3)
    res = segment_sharded(_Seg(), audio, SR, 0, STS, max_length=MAX_LEN, num_trials=num_trials, generate_fn=gen)
    q.put((rank, res, seen["n"], hi - lo, seen["slice"]))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("seconds,num_trials", [(73.0, 1), (31.0, 3), (4.0, 1)])
def test_sharded_segment_world2_gloo(seconds, num_trials):
    n = int(seconds * SR)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, num_trials, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    exp = _expected(np.zeros(n, dtype=np.float32), num_trials)
    assert len(exp["onset"]) > 0
    for rank, res, n_seen, n_mine, sl in got:
        assert res == exp, "rank %d result differs from the single-process result" % rank
        assert n_seen == n_mine
        assert sl[0] % 4 == 0 and sl[1] <= n


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 240, 361):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, world, r)[:2] for r in range(world)]
            flat = [i for a, b in spans for i in range(a, b)]
            assert flat == list(range(n))


# ------------------------------------------------------------------ folder mode (BASELINE configs[4]) sharded over ranks
def _folder_clips(seed, n_clips):
    """Ragged clips (0.3 .. 27 s at 16 kHz); clip i is filled with the constant i + 1 so that a worker can check
    that its window descriptors point into the right clip."""
    rng = np.random.default_rng(seed)
    secs = rng.uniform(0.3, 27.0, n_clips)
    if n_clips > 1:
        secs[1] = 0.05                                   # shorter than one FFT frame
    else:
        secs[0] = 3.0                                    # a single one-window clip: rank 1's shard is empty
    return [np.full(int(s_ * SR), i + 1, dtype=np.float32) for i, s_ in enumerate(secs)]


def _expected_folder(clips, num_trials):
    plan = FrontendPlan(SR, STS, 0)
    per_clip, owners, spans = folder_window_table(plan, [len(c) for c in clips], num_trials)
    texts = _Tok().batch_decode([_tokens_for(i, MAX_LEN - 3) for i in range(len(owners))])
    out = []
    for ci, (a, b) in enumerate(spans):
        pred = pp.parse_generation(texts[a:b], [w.as_tuple() for w in per_clip[ci]], STS * 2, len(clips[ci]) / SR, STS,
                                   num_trials, STS * 8, STS, "clustering", BOOK)
        out.append(pp.correct_fft_blur_and_dedupe(pred, SR, get_n_fft_given_sr(SR)))
    return out, owners, per_clip, spans


def _folder_worker(rank, world, port, seed, n_clips, num_trials, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    clips = _folder_clips(seed, n_clips)
    plan = FrontendPlan(SR, STS, 0)
    per_clip, owners, spans = folder_window_table(plan, [len(c) for c in clips], num_trials)
    checks = {"n": 0, "bad": 0}

    def gen(first, descs, plan_, piece):
        checks["n"] = len(descs)
        for i, (start, clo, chi) in enumerate(descs.tolist()):
            g = first + i
            ci = owners[g]
            w = per_clip[ci][g - spans[ci][0]]
            ok = (chi - clo == len(clips[ci]) and start - clo == w.start and clo % 4 == 0 and
                  (chi == clo or (piece[clo] == ci + 1 and piece[chi - 1] == ci + 1)))
            checks["bad"] += 0 if ok else 1
        return torch.tensor([_tokens_for(first + i, MAX_LEN - 3) for i in range(len(descs))], dtype=torch.int32).reshape(-1, MAX_LEN - 3)
    res = segment_many_sharded(_Seg(), clips, SR, 0, STS, max_length=MAX_LEN, num_trials=num_trials, generate_fn=gen)
    lo, hi, _ = shard_bounds(len(owners), world, rank)
    q.put((rank, res, checks["n"], hi - lo, checks["bad"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("seed,n_clips,num_trials", [(1, 9, 1), (2, 5, 3), (3, 2, 1), (4, 1, 1)])
def test_sharded_folder_world2_gloo(seed, n_clips, num_trials):
    """`segment_many_sharded`: the flattened window list of a ragged folder is cut into contiguous shards (a clip may
    straddle the cut), each rank's local buffer holds only its clips, and after ONE all-gather every rank returns
    the per-clip predictions of the single-process computation."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_folder_worker, args=(r, 2, port, seed, n_clips, num_trials, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    exp, owners, _, _ = _expected_folder(_folder_clips(seed, n_clips), num_trials)
    assert len(exp) == n_clips and sum(len(e["onset"]) for e in exp) > 0
    if n_clips == 1:
        assert len(owners) == 1 and sorted(g[3] for g in got) == [0, 1]      # one window: one rank idles
    for rank, res, n_seen, n_mine, bad in got:
        assert res == exp, "rank %d result differs from the single-process result" % rank
        assert n_seen == n_mine and bad == 0


# ------------------------------------------------------------------ folder of WAV files, sharded (audio_io.segment_files_sharded)
def _write_folder(root, seed, n_files):
    import struct
    rng = np.random.default_rng(seed)
    paths = []
    for i in range(n_files):
        sr = SR if i % 4 else 2 * SR                     # two sample-rate groups
        frames = int(rng.uniform(0.2, 18.0) * sr)
        pcm = np.full(frames, 100 * (i + 1), dtype="<i2").tobytes()
        fmt = struct.pack("<HHIIHH", 1, 1, sr, sr * 2, 2, 16)
        body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(pcm)) + pcm
        p = os.path.join(root, "clip%02d.wav" % i)
        with open(p, "wb") as f:
            f.write(b"RIFF" + struct.pack("<I", len(body)) + body)
        paths.append(p)
    return paths


def _files_worker(rank, world, port, paths, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from whisperseg_b200 import audio_io
    loaded = []
    real_load = audio_io.load_audio

    def counting_load(src, *a, **k):
        loaded.append(os.path.basename(src))
        return real_load(src, *a, **k)
    audio_io.load_audio = counting_load

    def gen(first, descs, plan_, piece):
        return torch.tensor([_tokens_for(first + i, MAX_LEN - 3) for i in range(len(descs))], dtype=torch.int32).reshape(-1, MAX_LEN - 3)
    per_file, table = audio_io.segment_files_sharded(_Seg(), paths, workers=2, generate_fn=gen, min_frequency=0,
                                                     spec_time_step=STS, max_length=MAX_LEN, num_trials=1)
    q.put((rank, per_file, table, sorted(loaded)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_files_world2_gloo(tmp_path):
    """Folder of WAV files on 2 ranks: headers fix the plan, each rank decodes only the clips of its shard, results
    are rank-identical and equal to the single-process per-clip post-processing of the same token rows."""
    from whisperseg_b200 import audio_io
    paths = _write_folder(str(tmp_path), 5, 9)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_files_worker, args=(r, 2, port, paths, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=180) for _ in procs], key=lambda g: g[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process expectation, one sample-rate group at a time (the window index restarts per group)
    exp = {}
    by_rate = {}
    for p in paths:
        by_rate.setdefault(audio_io.wav_info(p)[1], []).append(p)
    for sr, group in by_rate.items():
        plan = FrontendPlan(sr, STS, 0)
        lens = [audio_io.wav_info(p)[0] for p in group]
        per_clip, owners, spans = folder_window_table(plan, lens, 1)
        texts = _Tok().batch_decode([_tokens_for(i, MAX_LEN - 3) for i in range(len(owners))])
        for ci, (a, b) in enumerate(spans):
            pred = pp.parse_generation(texts[a:b], [w.as_tuple() for w in per_clip[ci]], STS * 2, lens[ci] / sr, STS, 1,
                                       STS * 8, STS, "clustering", BOOK)
            exp[group[ci]] = pp.correct_fft_blur_and_dedupe(pred, sr, get_n_fft_given_sr(sr))
    assert got[0][1] == exp and got[1][1] == exp and got[0][2] == got[1][2]
    assert got[0][2]["filename"] and len(got[0][2]["filename"]) == len(got[0][2]["onset"])
    names = sorted(os.path.basename(p) for p in paths)
    assert sorted(set(got[0][3]) | set(got[1][3])) == names            # every clip decoded somewhere ...
    assert len(got[0][3]) < len(names) and len(got[1][3]) < len(names)  # ... but no rank decoded the whole folder
