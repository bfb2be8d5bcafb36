"""The north-star correctness bar AS WRITTEN (BASELINE.json), asserted on the GPU:

  * decoder greedy tokens: RAW teacher-forced arg-max agreement with the fp32 oracle >= 99.9 % over >= 5000 positions
    (whisper-large on the whole bench workload: 240 windows of the 48 kHz clip; plus base on cfg1 and tiny);
  * final segments: segment-level F1 >= 0.99 with onsets/offsets within ONE spec_time_step, free-running
    `segment()` against the oracle's free-running output (reference scorer semantics, model.py:493-516).

Checkpoint: the "confident" synthetic recipe (tools/synth.py: script_vectors).  Why a recipe and not the round-1
"stress" recipe: a random network's logits are Gaussian, so a fixed fraction of positions are near-ties and ANY
bf16-operand implementation flips ~3x its relative error of them against fp32 (CPU model of the engine's rounding
points, oracle/bf16_emul.py + tools/noise_floor.py: 1.8-4 % flips for whisper-tiny..large with NO CUDA kernel in the
loop; half of that is the bf16 rounding of the WEIGHTS, which bf16-exact checkpoints now remove).  A trained segmenter's
logits are peaked; the confident recipe gives the random network that property (median top-1/top-2 margin ~3, against
0.2 for the stress recipe) while the positions that leave the script still depend on the audio through the whole
encoder/decoder stack.  The stress recipe stays in the suite (test_stress_recipe_reported) with its raw agreement and
the margin-conditional criterion of round 1.

Golden tokens/margins: tests/golden/tokens_*.npz from oracle/gen_golden_tokens.py (fp32 oracle, run in the container).
"""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(name, max_batch):
    from oracle.gen_golden_tokens import case_inputs
    from tools import synth
    from whisperseg_b200.segmenter import WhisperSegmenter
    arch, state, audio, sr, sts, n_win, max_length = case_inputs(name)
    tokdir = tempfile.mkdtemp()
    synth.token_table_files(tokdir)
    seg = WhisperSegmenter.from_state(state, tokdir, device="cuda", device_ids=[0], max_batch=max_batch)
    g = np.load(os.path.join(ROOT, "tests", "golden", "tokens_%s.npz" % name))
    assert int(g["n_windows"]) == n_win and int(g["max_length"]) == max_length
    return seg, state, audio, sr, sts, n_win, max_length, g["ids"].astype(np.int64), g["margins"].astype(np.float32)


def _valid_mask(ids, eos):
    valid = np.ones_like(ids, dtype=bool)
    for b in range(ids.shape[0]):
        e = np.nonzero(ids[b] == eos)[0]
        if len(e):
            valid[b, e[0] + 1:] = False
    return valid


def _teacher_forced(seg, audio, sr, sts, n_win, max_length, ids):
    import torch
    from whisperseg_b200.frontend import FrontendPlan
    eng, tok = seg.engines[0], seg.tokenizer
    plan = FrontendPlan(sr, sts, 0)
    wins = plan.windows(len(audio), 1)[:n_win]
    feats = eng.features(plan, audio, wins)
    got = np.zeros_like(ids)
    for c0 in range(0, n_win, eng.max_batch):
        c1 = min(n_win, c0 + eng.max_batch)
        forced = torch.full((c1 - c0, max_length), tok.eos_token_id, dtype=torch.int32)
        forced[:, :3] = torch.tensor(tok.prompt_ids, dtype=torch.int32)
        forced[:, 3:] = torch.from_numpy(ids[c0:c1]).to(torch.int32)
        eng.encode(feats[c0:c1].contiguous())
        out, _ = eng.generate(c1 - c0, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, max_length,
                              forced=forced.to(eng.device), use_graph=False)
        got[c0:c1] = out.cpu().numpy()
    return got


def _segments_from_ids(seg, ids, audio, sr, sts, n_win):
    """The oracle's tokens through the oracle's post-processing (the reference's own text -> segments path)."""
    from oracle import postprocess_ref as PR
    from whisperseg_b200.frontend import FrontendPlan, get_n_fft_given_sr
    texts = seg.tokenizer.batch_decode(ids)
    wins = FrontendPlan(sr, sts, 0).windows(len(audio), 1)[:n_win]
    return PR.segment_from_texts(texts, [w.as_tuple() for w in wins], len(audio), sr, sts, seg.cluster_codebook,
                                 get_n_fft_given_sr(sr))


@pytest.mark.parametrize("name,max_batch,min_positions", [("large_confident", 240, 5000), ("base_confident", 8, 100),
                                                          ("tiny_confident", 64, 1500)])
def test_raw_token_agreement_and_segment_f1(name, max_batch, min_positions):
    from oracle import postprocess_ref as PR
    seg, state, audio, sr, sts, n_win, max_length, ids, margins = _setup(name, max_batch)
    tok = seg.tokenizer
    valid = _valid_mask(ids, tok.eos_token_id)
    got = _teacher_forced(seg, audio, sr, sts, n_win, max_length, ids)
    n = int(valid.sum())
    flips = int(((got != ids) & valid).sum())
    raw = 1.0 - flips / n
    script = np.bincount(ids[valid].ravel()).max()           # not informative by itself; distinct rows is
    distinct = len({tuple(r) for r in ids.tolist()})
    print("%s: teacher-forced RAW agreement %.5f (%d flips / %d positions); oracle margins p1 %.3f p10 %.3f median %.3f; "
          "margins at the flips %s; distinct oracle rows %d / %d" %
          (name, raw, flips, n, np.percentile(margins[valid], 1), np.percentile(margins[valid], 10), np.median(margins[valid]),
           [round(float(v), 4) for v in margins[valid & (got != ids)][:8]], distinct, n_win))
    del script
    assert n >= min_positions
    assert raw >= 0.999
    # free-running segment() through the public API vs the oracle's free-running output
    n_samples = int(round(n_win * 1000 * sts * sr))
    clip = audio[:n_samples]
    res = seg.segment(clip, sr, min_frequency=0, spec_time_step=sts, num_trials=1, num_beams=1, max_length=max_length)
    gold = _segments_from_ids(seg, ids, clip, sr, sts, n_win)
    tp, n_pred, n_lab, p, r, f1 = PR.segment_score(res, gold, tolerance=sts)
    print("%s: free-running segments vs oracle within one spec_time_step: TP %d pred %d oracle %d  P %.4f R %.4f F1 %.4f" %
          (name, tp, n_pred, n_lab, p, r, f1))
    assert n_lab >= 8 * n_win // 2
    assert f1 >= 0.99


def test_stress_recipe_reported():
    """Round-1 "stress" recipe (Gaussian logits, the bench checkpoint), first 32 windows of the bench audio: raw
    agreement is bounded by the near-tie statistics, not by the kernels -- every flip must sit below the bf16 noise
    floor in oracle margin, and the confident positions must agree to 99.9 %."""
    seg, state, audio, sr, sts, n_win, max_length, ids, margins = _setup("large_stress32", 32)
    tok = seg.tokenizer
    valid = _valid_mask(ids, tok.eos_token_id)
    got = _teacher_forced(seg, audio, sr, sts, n_win, max_length, ids)
    mism = valid & (got != ids)
    raw = 1.0 - mism.sum() / valid.sum()
    conf = valid & (margins > 0.15)
    conf_agree = 1.0 - (mism & conf).sum() / max(1, conf.sum())
    print("large stress recipe: RAW agreement %.4f over %d positions; margin > 0.15: %.4f over %d; largest oracle margin at a flip %.4f"
          % (raw, valid.sum(), conf_agree, conf.sum(), margins[mism].max() if mism.any() else 0.0))
    assert valid.sum() >= 500
    assert conf_agree >= 0.999
    assert raw >= 0.93
