"""Token -> segment post-processing (host side, float64, bit-exact with the reference).

This is integer/index/float64 work that the reference does in Python after generation; it must
give *identical* results, so every float operation below is done in the reference's order
(`int * sts * 2`, then `+ offset_time`, comparisons on un-rounded values, `np.round(., 3)` last,
+-n_fft/2/sr after rounding).  Citations are to /root/reference/model.py.

Differences in *how* (not *what*):
  * parsing works on decoded text with the same regex (model.py:120) -- the text is produced by
    `tokens.TokenTable.decode`;
  * multi-trial consolidation by clustering does not build the reference's O(n^2) Python-callable
    distance matrix (model.py:305); neighbours are found with a sliding window over the onset-
    sorted segments (a pair within `eps` must have |onset difference| <= 2*eps) and DBSCAN's
    labels are derived from the connected components of the core-point graph, numbered and
    extended to border points the way sklearn's visiting order does -- the same clusters;
  * frame voting computes the per-frame mode with a counting pass instead of scipy.stats.mode.
"""
import re
from math import ceil

import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components

RATIO_DECODING_TIME_STEP_TO_SPEC_TIME_STEP = 2         # reference utils.py:5
_SEGMENT_RE = re.compile(r"<\|([0-9]+)\|>(\d+?)<\|([0-9]+)\|>")


def segments_from_text(text, spec_time_step, inverse_cluster_codebook):
    """model.py:191-207 -> list of [onset, offset, cluster_name] (window-relative seconds)."""
    found = []
    for on_s, cid_s, off_s in _SEGMENT_RE.findall(text):
        on = int(on_s) * spec_time_step * RATIO_DECODING_TIME_STEP_TO_SPEC_TIME_STEP
        off = int(off_s) * spec_time_step * RATIO_DECODING_TIME_STEP_TO_SPEC_TIME_STEP
        name = inverse_cluster_codebook.get(int(cid_s))
        if name is None or off - on <= 0:
            continue
        found.append([on, off, name])
    return found


def _stitch_trial(per_window):
    """Concatenate one trial's windows, fusing a segment cut by a window boundary (model.py:235-248)."""
    acc = []
    for segs in per_window:
        if acc and segs and acc[-1][1] == segs[0][0] and acc[-1][2] == segs[0][2]:
            acc[-1][1] = segs[0][1]
            segs = segs[1:]
        acc.extend(segs)
    return acc


def _neighbour_pairs(onsets, offsets, eps):
    """All ordered pairs (i, j) with (|d_onset| + |d_offset|)/2 <= eps (i == j included), sorted by (i, j).

    Candidates come from a window over the onset-sorted points (a pair within eps has |d_onset| <= 2 eps; the
    window is taken slightly wider, the exact test below decides), so the work is O(n * window), not O(n^2)."""
    n = len(onsets)
    order = np.argsort(onsets, kind="stable")
    so = onsets[order]
    reach = 2 * eps + 1e-6
    lo = np.searchsorted(so, so - reach, side="left")
    hi = np.searchsorted(so, so + reach, side="right")
    width = hi - lo
    src_r = np.repeat(np.arange(n), width)                       # rank of i in onset order
    starts = np.cumsum(width) - width
    dst_r = np.arange(int(width.sum())) - np.repeat(starts, width) + np.repeat(lo, width)
    i, j = order[src_r], order[dst_r]
    d = (np.abs(onsets[j] - onsets[i]) + np.abs(offsets[j] - offsets[i])) / 2
    keep = d <= eps
    i, j = i[keep], j[keep]
    srt = np.lexsort((j, i))
    return i[srt], j[srt]


def _dbscan_labels(onsets, offsets, eps, min_samples):
    """DBSCAN on the metric (|d_onset| + |d_offset|)/2 (model.py:285-288, 305-309).

    Same labels as sklearn's DBSCAN(metric="precomputed"), without replaying its stack: the core test counts the
    point itself; sklearn grows clusters from unlabeled core points in index order and finishes one cluster before
    it starts the next, so (a) the clusters of the core points are the connected components of the core-core
    neighbour graph, numbered by their smallest member index, and (b) a non-core point takes the smallest label
    among its core neighbours (the first cluster that reaches it), or -1 (noise) if it has none."""
    n = len(onsets)
    i, j = _neighbour_pairs(onsets, offsets, eps)
    core = np.bincount(i, minlength=n) >= min_samples
    labels = np.full(n, -1, dtype=np.int64)
    core_idx = np.nonzero(core)[0]
    if len(core_idx) == 0:
        return labels
    cc = core[i] & core[j]
    remap = np.full(n, -1, dtype=np.int64)
    remap[core_idx] = np.arange(len(core_idx))
    graph = coo_matrix((np.ones(int(cc.sum()), dtype=np.int8), (remap[i[cc]], remap[j[cc]])), shape=(len(core_idx),) * 2)
    _, comp = connected_components(graph, directed=False)
    # number the components by first appearance in index order (= order of their smallest core index)
    first = np.full(comp.max() + 1, n, dtype=np.int64)
    np.minimum.at(first, comp, core_idx)
    rank = np.empty_like(first)
    rank[np.argsort(first, kind="stable")] = np.arange(len(first))
    labels[core_idx] = rank[comp]
    # border points: smallest label among their core neighbours
    bc = ~core[i] & core[j]
    if bc.any():
        best = np.full(n, np.iinfo(np.int64).max, dtype=np.int64)
        np.minimum.at(best, i[bc], labels[j[bc]])
        reached = best != np.iinfo(np.int64).max
        labels[reached] = best[reached]
    return labels


def _group_means(values, starts, counts):
    """np.mean of every segment values[starts[g] : starts[g] + counts[g]], bit-identical to calling np.mean on the
    segment (what the reference does per cluster, model.py:322-326).  numpy adds fewer than 8 float64 values strictly
    left to right, so those segments are accumulated together, one member position per pass; longer segments (its
    blocked pairwise summation) are rare and go through np.mean itself.  (np.add.reduceat associates differently.)"""
    sums = np.zeros(len(starts), dtype=np.float64)
    small = counts < 8
    for k in range(int(counts[small].max()) if small.any() else 0):
        sel = small & (counts > k)
        sums[sel] += values[starts[sel] + k]
    means = sums / counts
    for g in np.nonzero(~small)[0]:
        means[g] = np.mean(values[starts[g]:starts[g] + counts[g]])
    return means


def consolidate_trials_by_clustering(trials, eps, min_samples):
    """model.py:291-337."""
    onsets = np.array([t for tr in trials for t in tr["onset"]], dtype=np.float64)
    offsets = np.array([t for tr in trials for t in tr["offset"]], dtype=np.float64)
    names = [c for tr in trials for c in tr["cluster"]]
    if len(onsets) == 0:
        return {"onset": [], "offset": [], "cluster": []}
    labels = _dbscan_labels(onsets, offsets, eps, min_samples)
    # members of every cluster in index order (stable sort by label); noise (-1) sorts first and is dropped
    order = np.argsort(labels, kind="stable")
    order = order[np.searchsorted(labels[order], 0, side="left"):]
    if len(order) == 0:
        return {"onset": [], "offset": [], "cluster": []}
    lab = labels[order]
    starts = np.concatenate([[0], np.nonzero(np.diff(lab))[0] + 1])
    counts = np.diff(np.concatenate([starts, [len(order)]]))
    mean_on = _group_means(onsets[order], starts, counts)
    mean_off = _group_means(offsets[order], starts, counts)
    # majority name, first-seen wins ties (dict insertion order in the reference loop)
    name_ids = {}
    ids = np.fromiter((name_ids.setdefault(c, len(name_ids)) for c in names), dtype=np.int64, count=len(names))[order]
    id_names = list(name_ids)
    lo_id, hi_id = np.minimum.reduceat(ids, starts), np.maximum.reduceat(ids, starts)
    chosen = []
    for g in range(len(starts)):
        if lo_id[g] == hi_id[g]:
            chosen.append(id_names[lo_id[g]])
            continue
        tally = {}
        for k in ids[starts[g]:starts[g] + counts[g]]:
            tally[k] = tally.get(k, 0) + 1
        best = max(tally.values())
        chosen.append(id_names[next(k for k, v in tally.items() if v == best)])
    srt = sorted(range(len(starts)), key=lambda g: mean_on[g])          # stable, like list.sort on the onset
    return {"onset": [mean_on[g] for g in srt], "offset": [mean_off[g] for g in srt], "cluster": [chosen[g] for g in srt]}


def consolidate_trials_by_voting(trials, time_per_frame, cluster_codebook):
    """model.py:339-394."""
    stamps = [t for tr in trials for t in list(tr["onset"]) + list(tr["offset"])]
    if len(stamps) == 0 or len(stamps) % 2 != 0:
        return {"onset": [], "offset": [], "cluster": []}
    t_min, t_max = np.min(stamps), np.max(stamps)
    n_frames = int(np.round((t_max - t_min) / time_per_frame))
    grid = np.full((len(trials), n_frames), -1, dtype=np.int64)
    for r, tr in enumerate(trials):
        # frame bounds of all segments at once (np.round is the same half-to-even rounding on arrays and scalars);
        # the fill stays sequential: where a trial's segments overlap, the later one wins, as in the reference
        a = np.round((np.asarray(tr["onset"], dtype=np.float64) - t_min) / time_per_frame).astype(np.int64)
        b = np.round((np.asarray(tr["offset"], dtype=np.float64) - t_min) / time_per_frame).astype(np.int64)
        ids = [cluster_codebook[name] for name in tr["cluster"]]
        row = grid[r]
        for lo, hi, cid in zip(a.tolist(), b.tolist(), ids):
            row[lo:hi] = cid
    # per-frame mode, smallest value on ties (scipy.stats.mode semantics)
    values = np.unique(grid)
    counts = np.stack([(grid == v).sum(axis=0) for v in values], axis=0) if n_frames else np.zeros((len(values), 0))
    voted = values[np.argmax(counts, axis=0)] if n_frames else np.zeros(0, dtype=np.int64)
    edges = np.nonzero(np.diff(np.concatenate([[-1], voted, [-1]])) != 0)[0]
    inverse = {v: k for k, v in cluster_codebook.items()}
    # a run between two change points is constant, so the reference's round(mean(run)) is the run's value itself
    run_ids = voted[edges[:-1]] if len(edges) > 1 else np.zeros(0, dtype=np.int64)
    keep = run_ids != -1
    starts, stops = edges[:-1][keep], edges[1:][keep]
    ons = list(starts * time_per_frame + t_min)
    offs = list(stops * time_per_frame + t_min)
    names = [inverse[int(c)] for c in run_ids[keep]]
    return {"onset": ons, "offset": offs, "cluster": names}


def extract_table(texts, windows, spec_time_step, cluster_codebook):
    """First half of model.py:210-232 for a run of windows: every window's text -> its `<|on|>id<|off|>` triples with
    the window's offset_time added (float64, `int * sts * 2` then `+ offset_time`, the reference's order).
    Returns (counts int32 [n_windows], onset f64 [n], offset f64 [n], cluster_id int32 [n]) -- plain arrays, so that a
    rank can compute the table of its own shard of windows and ranks exchange tables instead of token ids
    (distributed.py); `parse_table` finishes the job."""
    inverse = {v: k for k, v in cluster_codebook.items()}
    counts = np.zeros(len(texts), dtype=np.int32)
    ons, offs, cids = [], [], []
    for i, (text, win) in enumerate(zip(texts, windows)):
        offset_time = win[1]
        n = 0
        for on_s, cid_s, off_s in _SEGMENT_RE.findall(text):
            on = int(on_s) * spec_time_step * RATIO_DECODING_TIME_STEP_TO_SPEC_TIME_STEP
            off = int(off_s) * spec_time_step * RATIO_DECODING_TIME_STEP_TO_SPEC_TIME_STEP
            cid = int(cid_s)
            if cid not in inverse or off - on <= 0:
                continue
            ons.append(on + offset_time)
            offs.append(off + offset_time)
            cids.append(cid)
            n += 1
        counts[i] = n
    return counts, np.asarray(ons, dtype=np.float64), np.asarray(offs, dtype=np.float64), np.asarray(cids, dtype=np.int32)


def _stitch_clip_filter(on, off, cid, first, audio_duration, min_segment_length):
    """One trial, vectorised (model.py:235-258).  `first[i]`: segment i opens its window.  The reference fuses a
    window's first segment into the last accumulated one when `prev offset == onset` and the clusters agree; the
    accumulated segment's offset is then the fused segment's own offset, so the test only ever involves raw
    neighbours and a chain of fusions is a run of flags."""
    n = len(on)
    if n == 0:
        return on, off, cid
    fuse = np.zeros(n, dtype=bool)
    fuse[1:] = first[1:] & (off[:-1] == on[1:]) & (cid[:-1] == cid[1:])
    heads = np.flatnonzero(~fuse)
    last = np.concatenate([heads[1:], [n]]) - 1
    on, off, cid = on[heads], off[last], cid[heads]
    on = np.where(on <= 0, 0.0, on)                      # max(0, onset)
    off = np.minimum(off, audio_duration)                # min(offset, audio_duration)
    order = np.argsort(on, kind="stable")                # sorted(..., key=onset) is stable
    on, off, cid = on[order], off[order], cid[order]
    keep = off - on >= min_segment_length
    return on[keep], off[keep], cid[keep]


def parse_table(counts, on, off, cid, trial_ids, min_segment_length, audio_duration, num_trials, eps,
                time_per_frame_for_voting, consolidation_method, cluster_codebook, precision_bits=3):
    """Second half of model.py:210-281 on the table of ALL windows (generation order): per-trial stitching across
    window boundaries, clipping, sorting, minimum length, trial consolidation, rounding."""
    inverse = {v: k for k, v in cluster_codebook.items()}
    counts = np.asarray(counts, dtype=np.int64)
    trial_ids = np.asarray(trial_ids)
    seg_trial = np.repeat(trial_ids, counts)
    first = np.zeros(len(on), dtype=bool)
    starts = np.cumsum(counts) - counts
    first[starts[counts > 0]] = True
    seen = []
    for t in trial_ids.tolist():                         # trials in order of first appearance (dict insertion order)
        if t not in seen:
            seen.append(t)
    trials = []
    for t in seen:
        sel = seg_trial == t
        a, b, c = _stitch_clip_filter(on[sel], off[sel], cid[sel], first[sel], audio_duration, min_segment_length)
        trials.append({"onset": a.tolist(), "offset": b.tolist(), "cluster": [inverse[k] for k in c.tolist()]})
    if num_trials == 1:
        final = trials[0] if trials else {"onset": [], "offset": [], "cluster": []}
    elif consolidation_method == "clustering":
        final = consolidate_trials_by_clustering(trials, eps, max(2, int(ceil(num_trials * 0.5))))
    else:
        final = consolidate_trials_by_voting(trials, time_per_frame_for_voting, cluster_codebook)
    # np.round rounds arrays and scalars alike (x * 1000 -> rint -> / 1000); one call instead of one per segment
    final["onset"] = np.round(np.asarray(final["onset"], dtype=np.float64), precision_bits).tolist()
    final["offset"] = np.round(np.asarray(final["offset"], dtype=np.float64), precision_bits).tolist()
    return final


def parse_generation(texts, windows, min_segment_length, audio_duration, spec_time_step, num_trials, eps,
                     time_per_frame_for_voting, consolidation_method, cluster_codebook, precision_bits=3):
    """model.py:210-281.  `windows[i]` = (trial_id, offset_time, ...) in generation order."""
    counts, on, off, cid = extract_table(texts, windows, spec_time_step, cluster_codebook)
    return parse_table(counts, on, off, cid, [w[0] for w in windows][:len(texts)], min_segment_length, audio_duration,
                       num_trials, eps, time_per_frame_for_voting, consolidation_method, cluster_codebook, precision_bits)


def correct_fft_blur_and_dedupe(prediction, sr, n_fft):
    """model.py:439-468: shrink each segment by n_fft/2/sr per side, then drop exact duplicates."""
    delta = n_fft / 2 / sr
    rows = []
    for on, off, name in zip(prediction["onset"], prediction["offset"], prediction["cluster"]):
        a, b = on + delta, off - delta
        if a > b:
            a = b = (on + off) / 2
        rows.append((a, b, name))
    rows.sort(key=lambda r: r[0])
    out = {"onset": [], "offset": [], "cluster": []}
    for a, b, name in rows:
        if out["onset"] and a == out["onset"][-1] and b == out["offset"][-1] and name == out["cluster"][-1]:
            continue
        out["onset"].append(a)
        out["offset"].append(b)
        out["cluster"].append(name)
    return out


# ----------------------------------------------------------------------- scoring (model.py:474-569)
def compute_syllable_score(prediction_on_offset_list, label_on_offset_list, tolerance):
    n_pred, n_label, tp = len(prediction_on_offset_list), len(label_on_offset_list), 0
    for on, off, name in prediction_on_offset_list:
        for j, (lon, loff, lname) in enumerate(label_on_offset_list):
            if np.abs(on - lon) <= tolerance and np.abs(off - loff) <= tolerance and name == lname:
                tp += 1
                label_on_offset_list.pop(j)
                break
    return tp, n_pred, n_label


def segment_score(prediction, label, target_cluster=None, tolerance=None, default_spec_time_step=0.0025):
    if tolerance is None:
        tolerance = default_spec_time_step * 4
    def rows(d):
        return [[d["onset"][i], d["offset"][i], str(d["cluster"][i])] for i in range(len(d["onset"]))
                if target_cluster is None or str(target_cluster) == str(d["cluster"][i])]
    pred, lab = rows(prediction), rows(label)
    if target_cluster is not None and len(lab) == 0:
        print("Warning: the specified target cluster '%s' does not exist in the ground-truth labels." % str(target_cluster))
    tp, n_pred, n_label = compute_syllable_score(pred, lab, tolerance)
    precision = tp / max(n_pred, 1e-12)
    recall = tp / max(n_label, 1e-12)
    f1 = 2 / (1 / max(precision, 1e-12) + 1 / max(recall, 1e-12))
    return tp, n_pred, n_label, precision, recall, f1


def frame_score(prediction, label, target_cluster=None, time_per_frame_for_scoring=None, default_spec_time_step=0.0025):
    if time_per_frame_for_scoring is None:
        time_per_frame_for_scoring = min(0.001, default_spec_time_step)
    prediction["cluster"] = list(map(str, prediction["cluster"]))
    label["cluster"] = list(map(str, label["cluster"]))
    ids = {}
    for name in list(prediction["cluster"]) + list(label["cluster"]):
        ids.setdefault(name, len(ids))
    stamps = list(prediction["onset"]) + list(prediction["offset"]) + list(label["onset"]) + list(label["offset"])
    max_time = np.max(stamps) if len(stamps) else 1.0
    n_frames = int(np.round(max_time / time_per_frame_for_scoring)) + 1

    def rasterise(d):
        row = np.ones(n_frames) * -1
        for on, off, name in zip(d["onset"], d["offset"], d["cluster"]):
            row[int(np.round(on / time_per_frame_for_scoring)):int(np.round(off / time_per_frame_for_scoring))] = ids[name]
        return row
    fp, fl = rasterise(prediction), rasterise(label)
    if target_cluster is None:
        tp = np.logical_and(fl != -1, fp == fl).sum()
        p_pred, p_label = (fp != -1).sum(), (fl != -1).sum()
    else:
        tid = ids[target_cluster]
        tp = np.logical_and(fl == tid, fp == fl).sum()
        p_pred, p_label = (fp == tid).sum(), (fl == tid).sum()
    precision = tp / max(p_pred, 1e-12)
    recall = tp / max(p_label, 1e-12)
    f1 = 2 / (1 / max(precision, 1e-12) + 1 / max(recall, 1e-12))
    return tp, p_pred, p_label, precision, recall, f1
