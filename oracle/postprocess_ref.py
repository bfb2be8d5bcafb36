"""TEST INFRASTRUCTURE ONLY -- loop-level restatement of WhisperSeg's token->segment logic.

CPU oracle for the host-side post-processing (bit-exact float64 work).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it.

Restates, in the reference's own operation order (the float64 results depend on it):
  * extract_segments                    -- reference model.py:191-207, regex :120
  * parse_generation                    -- model.py:210-281
  * consolidate_trials_by_clustering    -- model.py:291-337 (+ custom_distance :285-288)
  * consolidate_trials_by_voting        -- model.py:339-394
  * the tail of segment()               -- model.py:439-468 (FFT-blur correction, de-dup)
  * segment_score / frame_score         -- model.py:474-569 (the parity metric)
Third-party pieces called as libraries, exactly like the reference: sklearn DBSCAN /
pairwise_distances (model.py:16-17), scipy.stats.mode (model.py:24).

Pinned by: tests/golden/postprocess_*.json (outputs of the unmodified reference, produced by
oracle/gen_golden.py) -- see tests/test_oracle_postprocess.py.
"""
import re

import numpy as np

RATIO = 2                      # reference utils.py:5
PRECISION_BITS = 3             # model.py:122
_MATCHER = re.compile(r"<\|([0-9]+)\|>(\d+?)<\|([0-9]+)\|>")    # model.py:120


def extract_segments(text, spec_time_step, cluster_codebook):
    inv = {v: k for k, v in cluster_codebook.items()}
    segs = []
    for on_t, cid_t, off_t in _MATCHER.findall(text):
        onset = int(on_t) * spec_time_step * RATIO
        offset = int(off_t) * spec_time_step * RATIO
        cid = int(cid_t)
        if cid not in inv:
            continue
        if offset - onset <= 0:
            continue
        segs.append([onset, offset, inv[cid]])
    return segs


def consolidate_by_clustering(trials, eps, min_samples):
    from sklearn.cluster import DBSCAN
    from sklearn.metrics import pairwise_distances
    segments = []
    for trial_id, trial in enumerate(trials):
        for on, off, cl in zip(trial["onset"], trial["offset"], trial["cluster"]):
            segments.append({"onset": on, "offset": off, "cluster": cl, "trial": trial_id})
    if len(segments) == 0:
        return {"onset": [], "offset": [], "cluster": []}

    def dist(a, b):
        return (abs(a[0] - b[0]) + abs(a[1] - b[1])) / 2

    dm = pairwise_distances([[s["onset"], s["offset"]] for s in segments], metric=dist)
    labels = DBSCAN(eps=eps, min_samples=min_samples, metric="precomputed").fit_predict(dm)
    merged = []
    for label in set(labels):
        if label == -1:
            continue
        group = [s for s, l in zip(segments, labels) if l == label]
        if not group:
            continue
        counts = {}
        for s in group:
            counts[s["cluster"]] = counts.get(s["cluster"], 0) + 1
        name = sorted(list(counts.items()), key=lambda x: -x[1])[0][0]
        merged.append({"onset": np.mean([s["onset"] for s in group]),
                       "offset": np.mean([s["offset"] for s in group]), "cluster": name})
    merged.sort(key=lambda x: x["onset"])
    return {"onset": [m["onset"] for m in merged], "offset": [m["offset"] for m in merged],
            "cluster": [m["cluster"] for m in merged]}


def consolidate_by_voting(trials, time_per_frame, cluster_codebook):
    from scipy.stats import mode
    stamps = []
    for t in trials:
        stamps += list(t["onset"])
        stamps += list(t["offset"])
    if len(stamps) == 0 or len(stamps) % 2 != 0:
        return {"onset": [], "offset": [], "cluster": []}
    t_min, t_max = np.min(stamps), np.max(stamps)
    n_frames = int(np.round((t_max - t_min) / time_per_frame))
    rows = []
    for t in trials:
        row = np.ones(n_frames) * -1
        for i in range(len(t["onset"])):
            a = t["onset"][i] - t_min
            b = t["offset"][i] - t_min
            row[int(np.round(a / time_per_frame)):int(np.round(b / time_per_frame))] = cluster_codebook[t["cluster"][i]]
        rows.append(row)
    voted, _ = mode(np.asarray(rows), axis=0)
    right = np.array(voted.tolist() + [-1])
    left = np.array([-1] + voted.tolist())
    events = np.argwhere(right - left != 0)[:, 0]
    inv = {v: k for k, v in cluster_codebook.items()}
    ons, offs, cls = [], [], []
    for i in range(0, len(events) - 1):
        a, b = events[i], events[i + 1]
        cid = int(np.round(np.mean(voted[a:b])))
        if cid == -1:
            continue
        ons.append(a * time_per_frame + t_min)
        offs.append(b * time_per_frame + t_min)
        cls.append(inv[cid])
    return {"onset": ons, "offset": offs, "cluster": cls}


def parse_generation(texts, windows, min_segment_length, audio_duration, spec_time_step, num_trials, eps,
                     time_per_frame_for_voting, consolidation_method, cluster_codebook):
    """`windows[i]` = (trial_id, offset_time, <anything>, clip_seconds) -- model.py:210-281."""
    per_trial = {}
    for text, win in zip(texts, windows):
        trial_id, offset_time = win[0], win[1]
        per_trial.setdefault(trial_id, [])
        segs = extract_segments(text, spec_time_step, cluster_codebook)
        for s in segs:
            s[0] += offset_time
            s[1] += offset_time
        per_trial[trial_id].append(segs)
    merged = {}
    for trial_id, clips in per_trial.items():
        acc = []
        for segs in clips:
            if acc and segs and acc[-1][1] == segs[0][0] and acc[-1][2] == segs[0][2]:
                acc[-1][1] = segs[0][1]
                segs = segs[1:]
            acc += segs
        merged[trial_id] = acc
    trials = []
    for trial_id, acc in merged.items():
        for s in acc:
            s[0] = max(0, s[0])
            s[1] = min(s[1], audio_duration)
        acc = sorted(acc, key=lambda x: x[0])
        acc = [s for s in acc if s[1] - s[0] >= min_segment_length]
        trials.append({"onset": [s[0] for s in acc], "offset": [s[1] for s in acc], "cluster": [s[2] for s in acc]})
    if num_trials == 1:
        final = trials[0]
    elif consolidation_method == "clustering":
        final = consolidate_by_clustering(trials, eps, max(2, int(np.ceil(num_trials * 0.5))))
    else:
        final = consolidate_by_voting(trials, time_per_frame_for_voting, cluster_codebook)
    final["onset"] = [float(np.round(t, PRECISION_BITS)) for t in final["onset"]]
    final["offset"] = [float(np.round(t, PRECISION_BITS)) for t in final["offset"]]
    return final


def finalize(pred, sr, n_fft):
    """FFT-blur correction and exact-duplicate removal -- model.py:439-468."""
    delta = n_fft / 2 / sr
    ons, offs = [], []
    for on, off in zip(pred["onset"], pred["offset"]):
        a, b = on + delta, off - delta
        if a > b:
            a = (on + off) / 2
            b = (on + off) / 2
        ons.append(a)
        offs.append(b)
    out = {"onset": ons, "offset": offs, "cluster": list(pred["cluster"])}
    if len(ons) > 0:
        co, cf, cc = [], [], []
        for on, off, cl in sorted(zip(out["onset"], out["offset"], out["cluster"]), key=lambda x: x[0]):
            if len(co) == 0 or on != co[-1] or off != cf[-1] or cl != cc[-1]:
                co.append(on)
                cf.append(off)
                cc.append(cl)
        out = {"onset": co, "offset": cf, "cluster": cc}
    return out


def segment_from_texts(texts, windows, n_samples, sr, spec_time_step, cluster_codebook, n_fft,
                       min_segment_length=None, eps=None, time_per_frame_for_voting=None,
                       consolidation_method="clustering", num_trials=1):
    """Everything in segment() after generation (model.py:420-470)."""
    if min_segment_length is None:
        min_segment_length = spec_time_step * RATIO
    if eps is None:
        eps = spec_time_step * RATIO * 4
    if time_per_frame_for_voting is None:
        time_per_frame_for_voting = spec_time_step
    pred = parse_generation(texts, windows, min_segment_length, n_samples / sr, spec_time_step, num_trials, eps,
                            time_per_frame_for_voting, consolidation_method, cluster_codebook)
    return finalize(pred, sr, n_fft)


# ------------------------------------------------------------------ scoring (model.py:474-569)
def segment_score(prediction, label, tolerance, target_cluster=None):
    pred = [[prediction["onset"][i], prediction["offset"][i], str(prediction["cluster"][i])]
            for i in range(len(prediction["onset"]))
            if target_cluster is None or str(target_cluster) == str(prediction["cluster"][i])]
    lab = [[label["onset"][i], label["offset"][i], str(label["cluster"][i])]
           for i in range(len(label["onset"]))
           if target_cluster is None or str(target_cluster) == str(label["cluster"][i])]
    n_pred, n_lab, tp = len(pred), len(lab), 0
    for on, off, cl in pred:
        hit = None
        for j, (lon, loff, lcl) in enumerate(lab):
            if np.abs(on - lon) <= tolerance and np.abs(off - loff) <= tolerance and cl == lcl:
                tp += 1
                hit = j
                break
        if hit is not None:
            lab.pop(hit)
    precision = tp / max(n_pred, 1e-12)
    recall = tp / max(n_lab, 1e-12)
    f1 = 2 / (1 / max(precision, 1e-12) + 1 / max(recall, 1e-12))
    return tp, n_pred, n_lab, precision, recall, f1
