"""Multi-GPU: one process per GPU (torchrun), windows sharded by index, one small all-gather of segment lists.

The reference fans windows out to one Python *thread* per GPU inside a single process and
concatenates the per-thread text lists (reference model.py:169-189); there is no collective.
Here every rank computes the same window plan, takes a contiguous shard of the window list
(the reference's `ceil(n / n_devices)` split, model.py:172-173), uploads only the samples its
windows touch, runs front-end + encoder + decode locally, decodes ITS OWN token rows to text and
extracts their `<|on|>id<|off|>` triples (postprocess.extract_table), and the ranks exchange those
segment tables with ONE all-gather of a fixed-size float64 buffer (NCCL over NVLink on GPUs, gloo in
the CPU tests).  Only the cheap second half of the post-processing (stitching across window
boundaries, trial consolidation: postprocess.parse_table) runs on every rank.  (Round 1 gathered
the padded token ids and every rank decoded and parsed ALL windows: 8x redundant host work at N=8.)
`segment_many_sharded` does the same for folder mode: the windows of all clips, flattened in clip order, are
sharded, and the gathered token ids are regrouped per clip.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import postprocess as pp
from .frontend import FrontendPlan, get_n_fft_given_sr


def shard_bounds(n_items, world_size, rank):
    per = int(np.ceil(n_items / world_size)) if n_items else 0
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per), per


def all_gather_tokens(local_ids, per, max_new, pad_id, group=None):
    """local_ids: int32 [n_local, max_new] (torch, any device).  Returns int32 [world*per, max_new]
    on the same device; rows beyond a rank's shard are pad."""
    world = dist.get_world_size(group)
    buf = torch.full((per, max_new), pad_id, dtype=torch.int32, device=local_ids.device)
    if local_ids.numel():
        buf[:local_ids.shape[0]] = local_ids
    out = torch.empty((world * per, max_new), dtype=torch.int32, device=local_ids.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return out


def all_gather_tables(table, per, max_new, device, group=None):
    """table = (counts, onset, offset, cluster_id) of this rank's windows (postprocess.extract_table).  ONE all-gather
    of a float64 buffer [1 + per + 3 * cap] per rank -- n_segments | counts (padded to `per` windows) | onsets |
    offsets | cluster ids, cap = per * (max_new // 3) being the most triples `per` windows can hold.  Returns the
    table of all windows in rank order (the caller trims each rank's counts to its real shard length)."""
    world = dist.get_world_size(group)
    counts, on, off, cid = table
    cap = max(1, per * max(1, max_new // 3))
    n = len(on)
    assert n <= cap and len(counts) <= per
    buf = np.zeros(1 + per + 3 * cap, dtype=np.float64)
    buf[0] = n
    buf[1:1 + len(counts)] = counts
    base = 1 + per
    buf[base:base + n] = on
    buf[base + cap:base + cap + n] = off
    buf[base + 2 * cap:base + 2 * cap + n] = cid
    send = torch.from_numpy(buf).to(device)
    out = torch.empty((world, buf.shape[0]), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out.view(-1), send, group=group)
    rows = out.cpu().numpy()
    tables = []
    for r in range(world):
        k = int(rows[r, 0])
        tables.append((rows[r, 1:1 + per].astype(np.int32), rows[r, base:base + k].copy(), rows[r, base + cap:base + cap + k].copy(),
                       rows[r, base + 2 * cap:base + 2 * cap + k].astype(np.int32)))
    return tables


def _merge_tables(tables, n_items, world):
    """Per-rank tables -> one table over all `n_items` windows (each rank's counts trimmed to its shard length)."""
    counts = []
    for r, t in enumerate(tables):
        a, b, _ = shard_bounds(n_items, world, r)
        counts.append(t[0][:b - a])
    return (np.concatenate(counts) if counts else np.zeros(0, np.int32), np.concatenate([t[1] for t in tables]),
            np.concatenate([t[2] for t in tables]), np.concatenate([t[3] for t in tables]))


def local_slice(audio, windows, clip_len):
    """Samples the shard's windows touch: (slice_start, audio[slice_start:slice_end])."""
    if not windows:
        return 0, np.zeros(0, dtype=np.float32)
    lo = max(0, min(w.start for w in windows))
    hi = min(len(audio), max(w.start + clip_len for w in windows))
    hi = max(hi, lo)
    lo -= lo % 4                                        # keep float4 alignment of the device buffer
    return lo, np.ascontiguousarray(audio[lo:hi], dtype=np.float32)


def segment_sharded(segmenter, audio, sr, min_frequency=None, spec_time_step=None, min_segment_length=None, eps=None,
                    time_per_frame_for_voting=None, consolidation_method="clustering", max_length=448, num_trials=1,
                    group=None, generate_fn=None, num_beams=4, length_penalty=1.0):
    """Drop-in for `segmenter.segment(...)` under torch.distributed: same result on every rank.

    `generate_fn(windows, plan, audio_slice, slice_start) -> int32 tensor [n_local, max_new]` can be
    injected (the gloo CPU tests use a scripted generator); by default the rank's engine runs."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cfgd = segmenter.default_segmentation_config
    if min_frequency is None:
        min_frequency = cfgd.get("min_frequency", 0)
    if spec_time_step is None:
        spec_time_step = cfgd.get("spec_time_step", 0.0025)
    ratio = pp.RATIO_DECODING_TIME_STEP_TO_SPEC_TIME_STEP
    if min_segment_length is None:
        min_segment_length = spec_time_step * ratio
    if eps is None:
        eps = spec_time_step * ratio * 4
    if time_per_frame_for_voting is None:
        time_per_frame_for_voting = spec_time_step
    audio = np.asarray(audio)
    plan = FrontendPlan(sr, spec_time_step, min_frequency, total_spec_columns=segmenter.total_spec_columns)
    wins = plan.windows(len(audio), num_trials)
    lo, hi, per = shard_bounds(len(wins), world, rank)
    mine = wins[lo:hi]
    tok = segmenter.tokenizer
    max_new = max_length - len(tok.prompt_ids)
    s0, piece = local_slice(audio, mine, plan.clip_len)
    if generate_fn is None:
        local_ids = _engine_generate(segmenter, plan, mine, piece, s0, len(audio), max_length, num_beams, length_penalty)
    else:
        local_ids = generate_fn(mine, plan, piece, s0)
    texts = tok.batch_decode(local_ids.cpu().numpy()) if len(mine) else []
    table = pp.extract_table(texts, [w.as_tuple() for w in mine], spec_time_step, segmenter.cluster_codebook)
    counts, on, off, cid = _merge_tables(all_gather_tables(table, per, max_new, local_ids.device, group), len(wins), world)
    pred = pp.parse_table(counts, on, off, cid, [w.trial_id for w in wins], min_segment_length, len(audio) / sr, num_trials, eps,
                          time_per_frame_for_voting, consolidation_method, segmenter.cluster_codebook, segmenter.precision_bits)
    return pp.correct_fft_blur_and_dedupe(pred, sr, get_n_fft_given_sr(sr))


class LazyClips:
    """A folder of clips whose lengths are known up front (WAV headers) and whose samples are loaded on demand:
    `segment_many_sharded` asks for the length of every clip but indexes only the clips of the rank's own shard.
    `prefetch(range)` starts loading those on a thread pool."""

    def __init__(self, lengths, load, pool=None):
        self.lengths = [int(n) for n in lengths]
        self._load, self._pool, self._pending = load, pool, {}

    def __len__(self):
        return len(self.lengths)

    def prefetch(self, indices):
        if self._pool is not None:
            for k in indices:
                if k not in self._pending:
                    self._pending[k] = self._pool.submit(self._load, k)

    def __getitem__(self, k):
        fut = self._pending.pop(k, None)
        a = fut.result() if fut is not None else self._load(k)
        if len(a) != self.lengths[k]:
            raise ValueError("clip %d: %d samples decoded, %d announced by its header" % (k, len(a), self.lengths[k]))
        return a


def folder_window_table(plan, lengths, num_trials):
    """Folder mode: the windows of all clips flattened in clip order.  Returns (per_clip_windows, owner clip of
    every window, [first, last+1) window range of every clip)."""
    per_clip, owners, spans = [], [], []
    for ci, n in enumerate(lengths):
        wins = plan.windows(int(n), num_trials)
        spans.append((len(owners), len(owners) + len(wins)))
        per_clip.append(wins)
        owners += [ci] * len(wins)
    return per_clip, owners, spans


def local_folder_buffer(audios, per_clip, owners, lo, hi):
    """The samples rank-local windows [lo, hi) of the flattened list touch: the clips they belong to, concatenated
    (each padded to a multiple of 4 samples, so every clip stays 16-byte aligned on the device), and one
    [start, clip_lo, clip_hi) descriptor per window in local buffer coordinates -- the layout segment_many()
    uses for the whole folder (per-window bounds keep clips from leaking into each other)."""
    if hi <= lo:
        return np.zeros(0, dtype=np.float32), np.zeros((0, 3), dtype=np.int64)
    c_lo, c_hi = owners[lo], owners[hi - 1]
    if hasattr(audios, "prefetch"):
        audios.prefetch(range(c_lo, c_hi + 1))
    pieces, base_of, base = [], {}, 0
    for ci in range(c_lo, c_hi + 1):
        a = np.ascontiguousarray(np.asarray(audios[ci]), dtype=np.float32)
        base_of[ci] = (base, len(a))
        pieces.append(a)
        pad = (-len(a)) % 4
        if pad:
            pieces.append(np.zeros(pad, dtype=np.float32))
        base += len(a) + pad
    first_of = {}
    pos = 0
    for ci, wins in enumerate(per_clip):
        first_of[ci] = pos
        pos += len(wins)
    descs = []
    for g in range(lo, hi):
        ci = owners[g]
        w = per_clip[ci][g - first_of[ci]]
        b, n = base_of[ci]
        descs.append([b + w.start, b, b + n])
    return (np.concatenate(pieces) if pieces else np.zeros(0, dtype=np.float32)), np.asarray(descs, dtype=np.int64).reshape(-1, 3)


def segment_many_sharded(segmenter, audios, sr, min_frequency=None, spec_time_step=None, min_segment_length=None, eps=None,
                         time_per_frame_for_voting=None, consolidation_method="clustering", max_length=448, num_trials=1,
                         group=None, generate_fn=None, num_beams=4, length_penalty=1.0):
    """Folder mode under torch.distributed (BASELINE configs[4]: thousands of variable-length clips on 8 GPUs):
    drop-in for `segmenter.segment_many(...)`, same list of predictions on every rank.

    The windows of ALL clips are flattened in clip order (like segment_many) and that list is cut into contiguous
    shards, so a rank's batches are filled across clip boundaries and the shards are balanced in windows, not in
    clips; each rank uploads only the clips its windows belong to.  One all-gather of the token ids, then every
    rank regroups them per clip and post-processes exactly like segment().
    `generate_fn(first_global_window, local_descs, plan, local_audio) -> int32 [n_local, max_new]` can be injected
    (the gloo CPU tests use a scripted generator); by default the rank's engine runs."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cfgd = segmenter.default_segmentation_config
    if min_frequency is None:
        min_frequency = cfgd.get("min_frequency", 0)
    if spec_time_step is None:
        spec_time_step = cfgd.get("spec_time_step", 0.0025)
    ratio = pp.RATIO_DECODING_TIME_STEP_TO_SPEC_TIME_STEP
    if min_segment_length is None:
        min_segment_length = spec_time_step * ratio
    if eps is None:
        eps = spec_time_step * ratio * 4
    if time_per_frame_for_voting is None:
        time_per_frame_for_voting = spec_time_step
    plan = FrontendPlan(sr, spec_time_step, min_frequency, total_spec_columns=segmenter.total_spec_columns)
    lengths = list(audios.lengths) if hasattr(audios, "lengths") else [len(a) for a in audios]
    per_clip, owners, spans = folder_window_table(plan, lengths, num_trials)
    if not owners:
        return []
    lo, hi, per = shard_bounds(len(owners), world, rank)
    tok = segmenter.tokenizer
    max_new = max_length - len(tok.prompt_ids)
    piece, descs = local_folder_buffer(audios, per_clip, owners, lo, hi)
    if generate_fn is None:
        local_ids = _engine_generate_folder(segmenter, plan, piece, descs, max_length, num_beams, length_penalty)
    else:
        local_ids = generate_fn(lo, descs, plan, piece)
    flat = [w for wins in per_clip for w in wins]
    texts = tok.batch_decode(local_ids.cpu().numpy()) if hi > lo else []
    table = pp.extract_table(texts, [w.as_tuple() for w in flat[lo:hi]], spec_time_step, segmenter.cluster_codebook)
    counts, on, off, cid = _merge_tables(all_gather_tables(table, per, max_new, local_ids.device, group), len(owners), world)
    seg_start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    n_fft = get_n_fft_given_sr(sr)
    results = []
    for ci, (a, b) in enumerate(spans):
        s0, s1 = int(seg_start[a]), int(seg_start[b])
        pred = pp.parse_table(counts[a:b], on[s0:s1], off[s0:s1], cid[s0:s1], [w.trial_id for w in per_clip[ci]],
                              min_segment_length, lengths[ci] / sr, num_trials, eps, time_per_frame_for_voting,
                              consolidation_method, segmenter.cluster_codebook, segmenter.precision_bits)
        results.append(pp.correct_fft_blur_and_dedupe(pred, sr, n_fft))
    return results


def _generate_from_features(segmenter, feats, max_length, num_beams, length_penalty):
    eng = segmenter.engines[0]
    tok = segmenter.tokenizer
    outs = []
    if not 1 <= int(num_beams) <= 4:
        raise ValueError("whisperseg_b200 supports num_beams in [1, 4], got %r" % (num_beams,))
    per_call = eng.max_batch if num_beams == 1 else max(1, eng.max_batch // num_beams)
    for pos in range(0, feats.shape[0], per_call):
        chunk = feats[pos:pos + per_call].contiguous()
        eng.encode(chunk)
        if num_beams == 1:
            ids, _ = eng.generate(chunk.shape[0], tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, max_length)
        else:
            ids, _ = eng.generate_beam(chunk.shape[0], int(num_beams), tok.prompt_ids, tok.eos_token_id, tok.pad_token_id,
                                       max_length, length_penalty)
        outs.append(ids)
    return torch.cat(outs, 0)


def _engine_generate_folder(segmenter, plan, piece, descs, max_length, num_beams=1, length_penalty=1.0):
    eng = segmenter.engines[0]
    max_new = max_length - len(segmenter.tokenizer.prompt_ids)
    if len(descs) == 0:
        return torch.zeros((0, max_new), dtype=torch.int32, device=eng.device)
    audio_dev = eng.upload_audio(piece)
    desc_dev = torch.from_numpy(np.ascontiguousarray(descs, dtype=np.int64)).to(eng.device)
    feats = eng.features_device(plan, audio_dev, desc_dev, len(descs))
    return _generate_from_features(segmenter, feats, max_length, num_beams, length_penalty)


def _engine_generate(segmenter, plan, windows, piece, slice_start, n_total, max_length, num_beams=1, length_penalty=1.0):
    eng = segmenter.engines[0]
    tok = segmenter.tokenizer
    max_new = max_length - len(tok.prompt_ids)
    if not windows:
        return torch.zeros((0, max_new), dtype=torch.int32, device=eng.device)
    feats = eng.features_sliced(plan, piece, windows, slice_start, n_total)
    return _generate_from_features(segmenter, feats, max_length, num_beams, length_penalty)

