"""Parity cases for the BASELINE.json configurations, through the public API.
cfg1 whisper-base 60 s @16 kHz | cfg2 whisper-large (2 windows of the bench workload) | cfg3 zebra-finch
parameters with num_trials=3 | cfg5 folder mode with ragged clips.  (cfg4 = sharding: tests/test_distributed_gloo.py
on CPU and bench.py --gpus N on GPUs.)"""
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BF16_MARGIN = 0.15


def _segmenter(arch, max_batch, **kw):
    from tools import synth
    from whisperseg_b200.segmenter import WhisperSegmenter
    state = synth.make_state(arch, seed=0, **kw)
    tokdir = tempfile.mkdtemp()
    synth.token_table_files(tokdir)
    return WhisperSegmenter.from_state(state, tokdir, device="cuda", device_ids=[0], max_batch=max_batch), state


def _teacher_forced(seg, state, x, max_length):
    """Returns (raw agreement, agreement on confident positions, #positions, encoder max/mean abs error)."""
    import torch
    from oracle.whisper_torch import WhisperOracle
    cfg, sd, gen = state
    orc = WhisperOracle(sd, cfg["encoder_attention_heads"], cfg["encoder_layers"])
    eng, tok = seg.engines[0], seg.tokenizer
    enc = orc.encode(x)
    hidden = eng.encode(x.to(eng.device).contiguous(), want_hidden=True).cpu()
    err = (hidden - enc).abs()
    ids, margins = orc.greedy(enc, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, max_length,
                              suppress_tokens=gen["suppress_tokens"], return_margins=True)
    B, n_new = ids.shape
    forced = torch.full((B, max_length), tok.pad_token_id, dtype=torch.int32)
    forced[:, :3] = torch.tensor(tok.prompt_ids, dtype=torch.int32)
    forced[:, 3:3 + n_new] = ids.to(torch.int32)
    got, _ = eng.generate(B, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, max_length, forced=forced.to(eng.device),
                          use_graph=False)
    got = got.cpu()[:, :n_new].long()
    valid = torch.ones_like(ids, dtype=torch.bool)
    for b in range(B):
        eos = (ids[b] == tok.eos_token_id).nonzero()
        if len(eos):
            valid[b, eos[0, 0] + 1:] = False
    raw = ((got == ids) & valid).sum().item() / max(1, valid.sum().item())
    conf_mask = valid & (margins > BF16_MARGIN)
    conf = ((got == ids) & conf_mask).sum().item() / max(1, conf_mask.sum().item())
    return raw, conf, int(valid.sum()), float(err.max()), float(err.mean())


def test_cfg1_whisper_base_60s_16k():
    import torch
    from oracle import frontend_np as FO
    from oracle import postprocess_ref as PR
    from tools import synth
    from whisperseg_b200.frontend import FrontendPlan, get_n_fft_given_sr
    seg, state = _segmenter("base", 8)
    audio = synth.synth_audio(60.0, 16000, seed=1)
    ref_feats = FO.sliced_audio_features(audio, 16000, 0, 0.01, 1, dtype=np.float32)
    assert len(ref_feats) == 6
    sliced = seg.get_sliced_audios_features(audio, 16000, 0, 0.01, 1)
    got = torch.stack([s[2] for s in sliced]).cpu().numpy()
    ref = np.asarray([f[2] for f in ref_feats])
    assert (np.abs(got - ref) / np.maximum(1.0, np.abs(ref))).max() <= 1e-4
    raw, conf, n, emax, emean = _teacher_forced(seg, state, torch.from_numpy(ref), 80)
    print("cfg1 base: encoder max-abs %.4f mean-abs %.5f; teacher-forced raw %.4f confident %.4f over %d positions"
          % (emax, emean, raw, conf, n))
    assert emax <= 0.12 and emean <= 0.012
    assert conf >= 0.999 and raw >= 0.9
    res = seg.segment(audio, 16000, min_frequency=0, spec_time_step=0.01, num_trials=1, num_beams=1, max_length=80)
    texts = seg.generate_segment_text(sliced, 4, 80, 1)
    wins = FrontendPlan(16000, 0.01, 0).windows(len(audio), 1)
    assert res == PR.segment_from_texts(texts, [w.as_tuple() for w in wins], len(audio), 16000, 0.01, seg.cluster_codebook,
                                        get_n_fft_given_sr(16000))


def test_cfg2_whisper_large_two_windows():
    import torch
    from oracle import frontend_np as FO
    from tools import synth
    seg, state = _segmenter("large", 4)
    audio = synth.synth_audio(5.0, 48000, seed=2)
    ref_feats = FO.sliced_audio_features(audio, 48000, 0, 0.0025, 1, dtype=np.float32)
    x = torch.from_numpy(np.asarray([f[2] for f in ref_feats]))
    raw, conf, n, emax, emean = _teacher_forced(seg, state, x, 40)
    print("cfg2 large: encoder max-abs %.4f mean-abs %.5f; teacher-forced raw %.4f confident %.4f over %d positions"
          % (emax, emean, raw, conf, n))
    assert emax <= 0.25 and emean <= 0.02          # 32 layers of bf16 GEMMs on unit-scale LayerNorm outputs
    assert conf >= 0.999 and raw >= 0.85


@pytest.mark.parametrize("arch,emax_bound,emean_bound", [("small", 0.08, 0.01), ("medium", 0.08, 0.01)])
def test_other_whisper_sizes(arch, emax_bound, emean_bound):
    """The published WhisperSeg checkpoints are base and large; the kernels take any Whisper size (head dim 64).  small
    (768 / 12 layers / 12 heads) and medium (1024 / 24 / 16) on two windows: encoder error, teacher-forced tokens."""
    import torch
    from oracle import frontend_np as FO
    from tools import synth
    seg, state = _segmenter(arch, 4)
    audio = synth.synth_audio(5.0, 32000, seed=4)
    ref_feats = FO.sliced_audio_features(audio, 32000, 0, 0.0025, 1, dtype=np.float32)
    x = torch.from_numpy(np.asarray([f[2] for f in ref_feats]))
    raw, conf, n, emax, emean = _teacher_forced(seg, state, x, 40)
    print("%s: encoder max-abs %.4f mean-abs %.5f; teacher-forced raw %.4f confident %.4f over %d positions"
          % (arch, emax, emean, raw, conf, n))
    assert emax <= emax_bound and emean <= emean_bound
    assert conf >= 0.999 and raw >= 0.85


def test_cfg3_zebra_finch_three_trials(tiny_checkpoint):
    from oracle import frontend_np as FO
    from oracle import postprocess_ref as PR
    from tools import synth
    from whisperseg_b200.frontend import FrontendPlan, get_n_fft_given_sr
    from whisperseg_b200.segmenter import WhisperSegmenter
    seg = WhisperSegmenter(tiny_checkpoint[0], device="cuda", device_ids=[0], max_batch=16)
    sr, sts = 32000, 0.0025
    audio = synth.synth_audio(7.3, sr, seed=3)
    sliced = seg.get_sliced_audios_features(audio, sr, 0, sts, 3)
    plan = FO.window_plan(len(audio), sr, sts, 3)
    assert [(s[0], s[1], s[3]) for s in sliced] == [(p[0], p[1], p[4]) for p in plan]
    assert len(sliced) == 11                                         # SURVEY.md section 4 golden
    for method in ("clustering", "voting"):
        res = seg.segment(audio, sr, min_frequency=0, spec_time_step=sts, num_trials=3, num_beams=1, max_length=64,
                          eps=0.02, min_segment_length=0.01, consolidation_method=method)
        texts = seg.generate_segment_text(sliced, 4, 64, 1)
        wins = FrontendPlan(sr, sts, 0).windows(len(audio), 3)
        ref = PR.segment_from_texts(texts, [w.as_tuple() for w in wins], len(audio), sr, sts, seg.cluster_codebook,
                                    get_n_fft_given_sr(sr), 0.01, 0.02, None, method, 3)
        assert res == ref


def test_cfg5_folder_mode_ragged_clips(tiny_checkpoint):
    from tools import synth
    from whisperseg_b200.segmenter import WhisperSegmenter
    seg = WhisperSegmenter(tiny_checkpoint[0], device="cuda", device_ids=[0], max_batch=8)
    sr, sts = 32000, 0.0025
    rng = np.random.default_rng(5)
    clips = [synth.synth_audio(float(rng.uniform(0.5, 6.0)), sr, seed=50 + i) for i in range(12)]
    clips.append(np.zeros(0, np.float32))
    clips.append(synth.synth_audio(0.013, sr, seed=99))
    many = seg.segment_many(clips, sr, min_frequency=0, spec_time_step=sts, max_length=48, num_beams=1)
    assert len(many) == len(clips)
    for clip, got in zip(clips, many):
        one = seg.segment(clip, sr, min_frequency=0, spec_time_step=sts, num_trials=1, num_beams=1, max_length=48)
        assert got == one
    many3 = seg.segment_many(clips[:5], sr, min_frequency=0, spec_time_step=sts, max_length=48, num_trials=3, num_beams=1)
    for clip, got in zip(clips[:5], many3):
        assert got == seg.segment(clip, sr, min_frequency=0, spec_time_step=sts, num_trials=3, num_beams=1, max_length=48)


def test_in_process_multi_device_fanout(tiny_checkpoint):
    """device_ids=[0, 1]: the reference's thread-per-device fan-out inside one process (model.py:169-189).
    Results must equal the single-device run (windows are independent)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tools import synth
    from whisperseg_b200.segmenter import WhisperSegmenter
    audio = synth.synth_audio(83.0, 16000, seed=31)
    one = WhisperSegmenter(tiny_checkpoint[0], device="cuda", device_ids=[0], max_batch=8)
    two = WhisperSegmenter(tiny_checkpoint[0], device="cuda", device_ids=[0, 1], max_batch=8)
    a = one.segment(audio, 16000, min_frequency=0, spec_time_step=0.01, num_trials=2, num_beams=1, max_length=64)
    b = two.segment(audio, 16000, min_frequency=0, spec_time_step=0.01, num_trials=2, num_beams=1, max_length=64)
    assert a == b and len(a["onset"]) == len(a["cluster"])
