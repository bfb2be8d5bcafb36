"""Model-level parity on the GPU (through the C ABI / the segmenter API) against the CPU oracle.

Stated tolerances:
  * encoder hidden states (bf16 GEMM operands, fp32 accumulation/residual): max-abs error
    <= 0.08 and mean-abs error <= 0.01 on LayerNorm-ed outputs (unit scale);
  * decoder: teacher-forced per-position arg-max agreement >= 99.9 % on positions whose oracle
    top-1/top-2 logit margin exceeds BF16_MARGIN (the bf16 noise floor of random-init weights,
    measured and printed), raw agreement reported;
  * segments: bit-identical post-processing given identical tokens; F1 between free-running GPU
    output and the oracle's output is reported (random-init weights make free-running agreement
    chaotic after the first flipped token -- SURVEY.md 7.2-2).
"""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BF16_MARGIN = 0.15


@pytest.fixture(scope="module")
def setup(tiny_checkpoint):
    import torch
    from tools import synth
    from oracle import frontend_np as FO
    from oracle.whisper_torch import oracle_from_hf
    from whisperseg_b200.segmenter import WhisperSegmenter
    path, hf = tiny_checkpoint
    seg = WhisperSegmenter(path, device="cuda", device_ids=[0], max_batch=8)
    audio = synth.synth_audio(47.0, 16000, seed=11)
    ref_feats = FO.sliced_audio_features(audio, 16000, 0, 0.01, 1, dtype=np.float32)
    orc = oracle_from_hf(hf)
    x = torch.from_numpy(np.asarray([f[2] for f in ref_feats]))
    return dict(seg=seg, hf=hf, orc=orc, audio=audio, ref_feats=ref_feats, x=x)


def test_encoder_hidden_states(setup):
    import torch
    seg, orc, x = setup["seg"], setup["orc"], setup["x"]
    eng = seg.engines[0]
    hidden = eng.encode(x.to(eng.device).contiguous(), want_hidden=True)
    torch.cuda.synchronize()
    ref = orc.encode(x)
    err = (hidden.cpu() - ref).abs()
    print("encoder max-abs %.4f mean-abs %.5f (ref abs mean %.3f)" % (err.max(), err.mean(), ref.abs().mean()))
    assert err.max().item() <= 0.08 and err.mean().item() <= 0.01


def test_encoder_matches_golden_probe(setup, golden_dir):
    import torch
    g = np.load(golden_dir + "/model_tiny.npz")
    seg, x = setup["seg"], setup["x"]
    eng = seg.engines[0]
    hidden = eng.encode(x.to(eng.device).contiguous(), want_hidden=True).cpu().numpy()
    err = np.abs(hidden[:, ::50, ::16] - g["enc_probe"])
    assert err.max() <= 0.08


@pytest.mark.parametrize("path", ["persistent-kernel", "fused-folded-ln", "fused-exact-ln", "tcgen05-splitk"])
def test_decoder_teacher_forced(setup, monkeypatch, path):
    """All three decode implementations of the linear layers against the fp32 oracle: the <= 64-row fused kernels
    with the LayerNorm folded into the projection (default), the same kernels with the exact on-the-fly LayerNorm
    (WSB_NO_FOLD=1), and the tcgen05 split-K GEMM + reduce pair used above 64 rows (WSB_NO_GEMV=1)."""
    import torch
    from tools import synth
    if path == "persistent-kernel":             # opt-in: csrc/mega.cu, one launch per decoder position
        monkeypatch.setenv("WSB_MEGA", "1")
    if path == "fused-exact-ln":
        monkeypatch.setenv("WSB_NO_FOLD", "1")
    elif path == "tcgen05-splitk":
        monkeypatch.setenv("WSB_NO_GEMV", "1")
    seg, orc, hf, x = setup["seg"], setup["orc"], setup["hf"], setup["x"]
    eng = seg.engines[0]
    tok = seg.tokenizer
    max_length = 96
    enc = orc.encode(x)
    sup = hf.generation_config.suppress_tokens
    ids, margins = orc.greedy(enc, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, max_length,
                              suppress_tokens=sup, begin_suppress_tokens=hf.generation_config.begin_suppress_tokens,
                              return_margins=True)
    B, n_new = ids.shape
    forced = torch.full((B, max_length), tok.pad_token_id, dtype=torch.int32)
    forced[:, :3] = torch.tensor(tok.prompt_ids, dtype=torch.int32)
    forced[:, 3:3 + n_new] = ids.to(torch.int32)
    eng.encode(x.to(eng.device).contiguous())
    got, n_steps = eng.generate(B, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, max_length,
                                forced=forced.to(eng.device), use_graph=False)
    got = got.cpu()[:, :n_new].long()
    valid = torch.ones_like(ids, dtype=torch.bool)
    for b in range(B):                       # positions after the oracle's EOS are padding
        eos = (ids[b] == tok.eos_token_id).nonzero()
        if len(eos):
            valid[b, eos[0, 0] + 1:] = False
    agree = (got == ids) & valid
    raw = agree.sum().item() / valid.sum().item()
    confident = valid & (margins > BF16_MARGIN)
    conf = ((got == ids) & confident).sum().item() / max(1, confident.sum().item())
    mism = margins[valid & (got != ids)]
    print("teacher-forced agreement [%s]: raw %.4f (%d positions), margin>%.2f: %.4f (%d positions); "
          "oracle margins at mismatches: %s" % (path, raw, valid.sum().item(), BF16_MARGIN, conf, confident.sum().item(),
                                                 [round(float(v), 4) for v in mism[:12]]))
    assert conf >= 0.999
    assert raw >= 0.90


def test_generate_graph_equals_eager(setup):
    import torch
    seg, x = setup["seg"], setup["x"]
    eng, tok = seg.engines[0], seg.tokenizer
    eng.encode(x.to(eng.device).contiguous())
    a, _ = eng.generate(x.shape[0], tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, 64, use_graph=False)
    eng.encode(x.to(eng.device).contiguous())
    b, _ = eng.generate(x.shape[0], tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, 64, use_graph=True)
    assert torch.equal(a, b)


def test_segment_end_to_end(setup, golden_dir):
    """segment() through the public API; post-processing must be bit-identical to the oracle's
    when fed the GPU's own tokens, and the result is scored against the reference's output."""
    from oracle import postprocess_ref as PR
    from whisperseg_b200.frontend import FrontendPlan, get_n_fft_given_sr
    seg, audio = setup["seg"], setup["audio"]
    res = seg.segment(audio, 16000, num_trials=1, num_beams=1, max_length=96)
    assert set(res) == {"onset", "offset", "cluster"} and len(res["onset"]) == len(res["offset"]) == len(res["cluster"])
    assert all(isinstance(v, float) for v in res["onset"] + res["offset"])
    # replay the same tokens through the oracle's post-processing
    plan = FrontendPlan(16000, 0.01, 0)
    wins = plan.windows(len(audio), 1)
    sliced = seg.get_sliced_audios_features(audio, 16000, 0, 0.01, 1)
    texts = seg.generate_segment_text(sliced, 4, 96, 1)
    ref = PR.segment_from_texts(texts, [w.as_tuple() for w in wins], len(audio), 16000, 0.01, seg.cluster_codebook,
                                get_n_fft_given_sr(16000))
    assert res == ref
    g = np.load(golden_dir + "/model_tiny.npz")
    gold = json.loads(bytes(g["segments"]).decode())
    tp, n_pred, n_lab, p, r, f1 = PR.segment_score(res, gold, tolerance=0.01)
    print("free-running segments vs reference: TP %d pred %d ref %d F1 %.3f" % (tp, n_pred, n_lab, f1))


def test_segments_match_unmodified_reference_on_confident_checkpoint(tmp_path, golden_dir):
    """North-star bar against the REFERENCE ITSELF: tests/golden/model_tiny_confident.npz holds what the unmodified
    /root/reference WhisperSegmenterForEval.segment returned (oracle/gen_golden.py, HF fp32 on the CPU) for a seeded
    tiny checkpoint of the "confident" recipe; the same checkpoint directory through this repo's WhisperSegmenter on the
    GPU must give segment-level F1 >= 0.99 with onsets/offsets within one spec_time_step (reference scorer,
    model.py:493-516), for num_trials=1 and for the 3-trial consolidated output, and the same generated ids."""
    from oracle import postprocess_ref as PR
    from tools import synth
    from whisperseg_b200.segmenter import WhisperSegmenter
    hf = synth.make_hf_model("tiny", seed=0, confident=True, default_segmentation_config=dict(
        sr=16000, min_frequency=0, spec_time_step=0.01, species="human"))
    path = synth.save_checkpoint(hf, str(tmp_path / "ckpt"))
    seg = WhisperSegmenter(path, device="cuda", device_ids=[0], max_batch=8)
    g = np.load(golden_dir + "/model_tiny_confident.npz")
    audio = synth.synth_audio(47.0, 16000, seed=11)
    for key, trials in (("segments", 1), ("segments_trials3", 3)):
        gold = json.loads(bytes(g[key]).decode())
        res = seg.segment(audio, 16000, num_trials=trials, num_beams=1, batch_size=8, max_length=96)
        tp, n_pred, n_lab, p, r, f1 = PR.segment_score(res, gold, tolerance=0.01)
        print("confident tiny checkpoint, num_trials=%d: vs unmodified reference TP %d pred %d ref %d F1 %.4f" % (trials, tp, n_pred, n_lab, f1))
        if trials == 1:
            assert n_lab >= 20 and f1 >= 0.99
        else:                      # the scripted segments sit at window-relative times, so shifted trials never agree: both empty
            assert res == gold or f1 >= 0.99
    sliced = seg.get_sliced_audios_features(audio, 16000, 0, 0.01, 1)
    texts = seg.generate_segment_text(sliced, 8, 96, 1)
    ref_texts = json.loads(bytes(g["texts"]).decode())
    def strip(t):                  # the two sides pad finished rows differently (EOS runs / prompt tokens): compare the payload
        for special in ("<|endoftext|>", "<|startoftranscript|>", "<|en|>", "<|notimestamps|>"):
            t = t.replace(special, "")
        return t
    same = sum(strip(a) == strip(b) for a, b in zip(texts, ref_texts))
    print("generated texts identical to the reference's for %d / %d windows" % (same, len(ref_texts)))
    assert same >= len(ref_texts) - 1


def test_segment_multi_trial_and_empty(setup):
    seg, audio = setup["seg"], setup["audio"]
    res = seg.segment(audio[:16000 * 12], 16000, num_trials=3, num_beams=1, max_length=48)
    assert len(res["onset"]) == len(res["cluster"])
    assert res["onset"] == sorted(res["onset"])
    empty = seg.segment(np.zeros(0, np.float32), 16000, num_trials=1, num_beams=1, max_length=16)
    assert set(empty) == {"onset", "offset", "cluster"}


def test_batch_compaction_equals_uncompacted(tiny_checkpoint, monkeypatch):
    """Gathering the still-active rows into a smaller batch (decode.cu: batch compaction) must not change
    a single token: 96 windows, most of which stop early, decoded with and without compaction."""
    import torch
    from tools import synth
    from whisperseg_b200.frontend import FrontendPlan
    from whisperseg_b200.segmenter import WhisperSegmenter
    seg = WhisperSegmenter(tiny_checkpoint[0], device="cuda", device_ids=[0], max_batch=96)
    eng, tok = seg.engines[0], seg.tokenizer
    sr, sts = 16000, 0.001
    audio = synth.synth_audio(96.0, sr, seed=21)
    plan = FrontendPlan(sr, sts, 0)
    wins = plan.windows(len(audio), 1)
    assert len(wins) == 96
    feats = eng.features(plan, audio, wins)
    outs = {}
    monkeypatch.setenv("WSB_NO_GEMV", "1")       # same linear-layer kernels on both sides: this test is about the gather
    monkeypatch.setenv("WSB_ATTN_THREADS", "128")  # ... and the same attention variant whatever the row count
    for mode in ("compact", "plain"):
        if mode == "plain":
            monkeypatch.setenv("WSB_NO_COMPACT", "1")
        eng.encode(feats)
        ids, n_steps = eng.generate(96, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, 160)
        outs[mode] = (ids.cpu(), n_steps)
    lens = (outs["plain"][0] != tok.eos_token_id).sum(dim=1)
    print("row lengths: min %d median %d max %d; steps %d" % (lens.min(), lens.median(), lens.max(), outs["plain"][1]))
    assert torch.equal(outs["compact"][0], outs["plain"][0])
    assert outs["compact"][1] == outs["plain"][1]


@pytest.mark.parametrize("seconds,max_batch", [(13.0, 16), (29.0, 32), (45.0, 64), (85.0, 96)])
@pytest.mark.parametrize("recipe", ["stress", "confident"])
def test_small_batch_linear_path_matches_tensor_core_path(tiny_checkpoint, tiny_confident_checkpoint, monkeypatch, seconds, max_batch, recipe):
    """Batches of 65..256 rows can run one cluster split-K launch per linear layer (skinny.cu, opt-in with WSB_CLUSTER=1:
    the 85-window case) and
    batches of <= 64 rows (1, 2 or 4 m-tiles of 16) run the fused mma.sync linear kernels (gemv.cu) instead of the
    tcgen05 split-K GEMM + reduce pair, with the LayerNorm folded into the projection (the kernel reads bf16(x) and
    applies rstd (acc - mean c1) + c2), so the two paths round different quantities to bf16: independent rounding
    noise of the same size.  Teacher-forced on the tensor-core path's own tokens, the per-position arg-max must agree
    on >= 98.5 % of the positions (near-ties of a random-init model flip; 99.0-99.1 % measured, 99.2-99.7 % with the
    exact on-the-fly LayerNorm, WSB_NO_FOLD=1), and the free-running outputs of most rows must be identical.  Parity
    of this path against the fp32 oracle is asserted by the teacher-forced tests above (they run <= 64 rows).
    On the "confident" recipe (peaked logits) the two paths must agree on >= 99.9 % of the positions; on the "stress"
    recipe (Gaussian logits: 10 % of the positions are near-ties) the agreement is a noise measurement, >= 97 %."""
    import torch
    from tools import synth
    from whisperseg_b200.frontend import FrontendPlan
    from whisperseg_b200.segmenter import WhisperSegmenter
    ckpt = tiny_checkpoint if recipe == "stress" else tiny_confident_checkpoint
    seg = WhisperSegmenter(ckpt[0], device="cuda", device_ids=[0], max_batch=max_batch)
    eng, tok = seg.engines[0], seg.tokenizer
    sr, sts = 16000, 0.001
    audio = synth.synth_audio(seconds, sr, seed=23)
    plan = FrontendPlan(sr, sts, 0)
    wins = plan.windows(len(audio), 1)
    n = len(wins)
    assert max_batch // 2 < n <= max_batch
    feats = eng.features(plan, audio, wins)
    max_length = 64
    eng.encode(feats)
    def tensor_core_path(on):                  # reference = tcgen05 split-K pair; else gemv.cu (<= 64 rows) / skinny.cu (opt-in)
        monkeypatch.setenv("WSB_NO_GEMV", "1") if on else monkeypatch.delenv("WSB_NO_GEMV", raising=False)
        monkeypatch.delenv("WSB_CLUSTER", raising=False) if on else monkeypatch.setenv("WSB_CLUSTER", "1")

    tensor_core_path(True)
    ref, _ = eng.generate(n, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, max_length)
    tensor_core_path(False)
    got, _ = eng.generate(n, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, max_length)
    same_rows = (ref == got).all(dim=1).float().mean().item()
    forced = torch.full((n, max_length), tok.eos_token_id, dtype=torch.int32, device=eng.device)
    forced[:, :3] = torch.tensor(tok.prompt_ids, dtype=torch.int32, device=eng.device)
    forced[:, 3:] = ref
    tf_new, _ = eng.generate(n, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, max_length, forced=forced, use_graph=False)
    tensor_core_path(True)
    tf_old, _ = eng.generate(n, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, max_length, forced=forced, use_graph=False)
    agree = (tf_new == tf_old).float().mean().item()
    print("small-batch path [%s]: %d windows, free-running rows identical %.3f, teacher-forced arg-max agreement %.4f" % (recipe, n, same_rows, agree))
    if recipe == "confident":
        assert agree >= 0.999 and same_rows >= 0.95
    else:
        assert agree >= 0.97 and same_rows >= 0.6


def test_fused_prompt_prefill_matches_position_by_position(tiny_confident_checkpoint, monkeypatch):
    """Above 64 rows the three prompt positions run as ONE pass (3 x B virtual rows, causal 3x3 self-attention, 3-query
    cross-attention from one pass over a row's K/V block: engine.cu prefill_and_first_token) instead of three decoder
    positions (WSB_NO_PREFILL=1).  Same network, different reduction orders in the two attention kernels: on the confident
    checkpoint every generated token must be identical."""
    import torch
    from tools import synth
    from whisperseg_b200.frontend import FrontendPlan
    from whisperseg_b200.segmenter import WhisperSegmenter
    seg = WhisperSegmenter(tiny_confident_checkpoint[0], device="cuda", device_ids=[0], max_batch=96)
    eng, tok = seg.engines[0], seg.tokenizer
    sr, sts = 16000, 0.001
    audio = synth.synth_audio(90.0, sr, seed=37)
    plan = FrontendPlan(sr, sts, 0)
    wins = plan.windows(len(audio), 1)
    assert len(wins) == 90
    feats = eng.features(plan, audio, wins)
    outs = {}
    for mode in ("fused", "per-position"):
        if mode == "per-position":
            monkeypatch.setenv("WSB_NO_PREFILL", "1")
        else:
            monkeypatch.delenv("WSB_NO_PREFILL", raising=False)
        eng.encode(feats)
        ids, steps = eng.generate(len(wins), tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, 64)
        outs[mode] = (ids.cpu(), steps)
    same_rows = (outs["fused"][0] == outs["per-position"][0]).all(dim=1).float().mean().item()
    print("fused prefill vs position-by-position: %d windows, rows identical %.3f, positions %d / %d"
          % (len(wins), same_rows, outs["fused"][1], outs["per-position"][1]))
    assert same_rows >= 0.98
    assert outs["fused"][1] == outs["per-position"][1]
