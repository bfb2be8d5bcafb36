"""Token->segment post-processing: the product implementation and the oracle restatement must both
reproduce, BIT-EXACTLY, the outputs of the unmodified reference (tests/golden/postprocess.json) --
float64 equality, no tolerance -- plus the survey's hand-derived known answers and edge cases."""
import json

import numpy as np
import pytest

from oracle import postprocess_ref as PR
from oracle.ref_shim import reference_available
from whisperseg_b200 import postprocess as P
from whisperseg_b200.frontend import FrontendPlan


def _cases(golden_dir):
    return json.load(open(golden_dir + "/postprocess.json"))


def _run_product(c):
    cfg = c["config"]
    sts = cfg["sts"]
    plan = FrontendPlan(cfg["sr"], sts, 0)
    n = int(cfg["seconds"] * cfg["sr"])
    wins = plan.windows(n, cfg["num_trials"])
    assert [[w.trial_id, w.offset_time, w.clip_seconds] for w in wins] == c["windows"]
    pred = P.parse_generation(c["texts"], [w.as_tuple() for w in wins], cfg.get("min_segment_length", sts * 2), n / cfg["sr"],
                              sts, cfg["num_trials"], cfg.get("eps", sts * 8), cfg.get("time_per_frame_for_voting", sts),
                              cfg["method"], cfg["codebook"])
    return P.correct_fft_blur_and_dedupe(pred, cfg["sr"], plan.n_fft)


def _run_oracle(c):
    cfg = c["config"]
    plan = FrontendPlan(cfg["sr"], cfg["sts"], 0)
    return PR.segment_from_texts(c["texts"], c["windows"], int(cfg["seconds"] * cfg["sr"]), cfg["sr"], cfg["sts"],
                                 cfg["codebook"], plan.n_fft, cfg.get("min_segment_length"), cfg.get("eps"),
                                 cfg.get("time_per_frame_for_voting"), cfg["method"], cfg["num_trials"])


def test_product_bit_exact_vs_reference_golden(golden_dir):
    cases = _cases(golden_dir)
    assert len(cases) >= 8
    for c in cases:
        assert _run_product(c) == c["expected"]


def test_oracle_bit_exact_vs_reference_golden(golden_dir):
    for c in _cases(golden_dir):
        assert _run_oracle(c) == c["expected"]


def test_survey_known_answer(golden_dir):
    """zero-length dropped, unknown id dropped, boundary merge, clip to 25 s, and
    <|200|>1<|201|> dropped because 14.02-14.0 < 0.02 in float64."""
    c = _cases(golden_dir)[0]
    res = _run_product(c)
    assert res["cluster"] == ["vocal", "vocal", "vocal"]
    assert res["onset"] == [0.21600000000000003, 9.016, 22.016]
    assert res["offset"] == [1.184, 10.484, 24.984]


def test_clustering_and_voting_known_answer():
    """SURVEY.md section 4 consolidation example (one 12.00-12.50 s segment seen by 3 trials, trial 2
    onset one token late, a spurious segment in trial 1)."""
    trials = [{"onset": [12.0], "offset": [12.5], "cluster": ["vocal"]},
              {"onset": [12.0, 3.0], "offset": [12.5, 3.2], "cluster": ["vocal", "vocal"]},
              {"onset": [12.02], "offset": [12.5], "cluster": ["vocal"]}]
    a = P.consolidate_trials_by_clustering(trials, 0.08, 2)
    b = PR.consolidate_by_clustering(trials, 0.08, 2)
    assert a == b and len(a["onset"]) == 1 and a["cluster"] == ["vocal"]
    v = P.consolidate_trials_by_voting(trials, 0.01, {"vocal": 0})
    w = PR.consolidate_by_voting(trials, 0.01, {"vocal": 0})
    assert v == w and len(v["onset"]) == 1


def test_random_streams_product_equals_oracle():
    """Property test on seeded random token streams: product == oracle restatement, bit for bit."""
    rng = np.random.default_rng(123)
    book = {"a": 0, "b": 1, "c": 2}
    for trial in range(30):
        sr, sts = [(16000, 0.01), (32000, 0.0025), (48000, 0.0025)][trial % 3]
        nt = int(rng.integers(1, 5))
        secs = float(rng.uniform(0.3, 4.0)) * 1000 * sts
        plan = FrontendPlan(sr, sts, 0)
        n = int(secs * sr)
        wins = plan.windows(n, nt)
        texts = []
        for w in wins:
            toks, t = [], 0
            while t < 480:
                t += int(rng.integers(0, 90))
                d = int(rng.integers(0, 60))
                toks.append("<|%d|>%d<|%d|>" % (t, int(rng.integers(0, 4)), min(500, t + d)))
                t += d
            texts.append("".join(toks))
        method = "clustering" if trial % 2 == 0 else "voting"
        wt = [w.as_tuple() for w in wins]
        a = P.correct_fft_blur_and_dedupe(P.parse_generation(texts, wt, sts * 2, n / sr, sts, nt, sts * 8, sts, method, book),
                                          sr, plan.n_fft)
        b = PR.segment_from_texts(texts, [[w.trial_id, w.offset_time, w.clip_seconds] for w in wins], n, sr, sts, book,
                                  plan.n_fft, None, None, None, method, nt)
        assert a == b
        assert a["onset"] == sorted(a["onset"])


def test_dbscan_labels_equal_sklearn_precomputed():
    """`_dbscan_labels` (windowed neighbour search + connected components of the core graph, no O(n^2) matrix, no
    replay of sklearn's stack) against the call the reference makes (model.py:305-309): pairwise_distances with the
    Python-callable metric of model.py:285-288, then DBSCAN(metric="precomputed") -- identical labels, including
    the numbering of the clusters and which cluster a border point joins; and the merged segments bit for bit."""
    from sklearn.cluster import DBSCAN
    rng = np.random.default_rng(77)
    for case in range(60):
        n = int(rng.integers(1, 150))
        base = np.sort(rng.uniform(0, n * rng.choice([0.01, 0.05, 0.4]), n))
        on = base + rng.normal(0, rng.choice([1e-3, 1e-2]), n)
        off = on + rng.uniform(0.005, 0.2, n)
        if case % 2:                                   # a token grid makes exact ties and exact eps hits
            on, off = np.round(on / 0.005) * 0.005, np.round(off / 0.005) * 0.005
        eps, ms = float(rng.choice([0.005, 0.02, 0.04, 0.2])), int(rng.integers(1, 5))
        dist = (np.abs(on[:, None] - on[None, :]) + np.abs(off[:, None] - off[None, :])) / 2
        ref = DBSCAN(eps=eps, min_samples=ms, metric="precomputed").fit(dist).labels_
        assert np.array_equal(P._dbscan_labels(on, off, eps, ms), ref)
        names = rng.choice(["a", "b"], size=n).tolist()
        trials = [{"onset": on.tolist(), "offset": off.tolist(), "cluster": names}]
        assert P.consolidate_trials_by_clustering(trials, eps, ms) == PR.consolidate_by_clustering(trials, eps, ms)


def test_scoring_known_answer():
    pred = {"onset": [1, 2, 3], "offset": [1.5, 2.5, 3.5], "cluster": ["vocal", "vocal", "b"]}
    label = {"onset": [1.005, 2.2, 3], "offset": [1.5, 2.6, 3.5], "cluster": ["vocal", "vocal", "b"]}
    tp, n_pred, n_lab, p, r, f1 = P.segment_score(dict(pred), dict(label), tolerance=0.01)
    assert (tp, n_pred, n_lab) == (2, 3, 3) and abs(f1 - 2 / 3) < 1e-12
    assert PR.segment_score(pred, label, 0.01)[:3] == (2, 3, 3)
    fs = P.frame_score({k: list(v) for k, v in pred.items()}, {k: list(v) for k, v in label.items()},
                       default_spec_time_step=0.0025)
    assert (int(fs[0]), int(fs[1]), int(fs[2])) == (1295, 1500, 1395)


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted (GPU box)")
def test_product_vs_live_reference_random():
    from oracle.ref_shim import import_reference
    ref_model, _ = import_reference()
    rng = np.random.default_rng(5)
    for trial in range(6):
        sr, sts, nt = 16000, 0.01, int(rng.integers(1, 4))
        seg = ref_model.SegmenterBase()
        seg.total_spec_columns = 1000
        seg.cluster_codebook = {"vocal": 0, "x": 1}
        audio = np.zeros(int(rng.uniform(5, 35) * sr), dtype=np.float32)
        plan = FrontendPlan(sr, sts, 0)
        wins = plan.windows(len(audio), nt)
        texts = []
        for w in wins:
            t, toks = 0, []
            while t < 480:
                t += int(rng.integers(1, 120))
                d = int(rng.integers(1, 40))
                toks.append("<|%d|>%d<|%d|>" % (t, int(rng.integers(0, 2)), min(500, t + d)))
                t += d
            texts.append("".join(toks))
        seg.generate_segment_text = lambda *a, _t=texts, **k: list(_t)
        method = "voting" if trial % 2 else "clustering"
        ref = seg.segment(audio, sr, min_frequency=0, spec_time_step=sts, num_trials=nt, consolidation_method=method)
        mine = P.correct_fft_blur_and_dedupe(
            P.parse_generation(texts, [w.as_tuple() for w in wins], sts * 2, len(audio) / sr, sts, nt, sts * 8, sts, method,
                               seg.cluster_codebook), sr, plan.n_fft)
        assert mine == {k: list(v) for k, v in ref.items()}
