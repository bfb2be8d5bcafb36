"""K5e (csrc/mega.cu, opt-in with WSB_MEGA=1): the persistent one-launch-per-position decode kernel for <= 64 rows against
the launch-per-layer fused path (gemv.cu + decode.cu, the default).  Both run the same arithmetic in the same order, so the
generated tokens must be BIT-IDENTICAL -- free-running, with and without CUDA graphs, across the 16 / 32 / 64-row
variants and through batch compaction; parity against the fp32 oracle is asserted for both paths in test_gpu_model.py
(test_decoder_teacher_forced) and test_gpu_parity_bar.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seconds,max_batch", [(5.0, 8), (13.0, 16), (29.0, 32), (61.0, 64)])
def test_mega_tokens_identical_to_launch_per_layer_path(tiny_checkpoint, monkeypatch, seconds, max_batch):
    import torch
    from tools import synth
    from whisperseg_b200.frontend import FrontendPlan
    from whisperseg_b200.segmenter import WhisperSegmenter
    seg = WhisperSegmenter(tiny_checkpoint[0], device="cuda", device_ids=[0], max_batch=max_batch)
    eng, tok = seg.engines[0], seg.tokenizer
    sr, sts = 16000, 0.001
    audio = synth.synth_audio(seconds, sr, seed=29)
    plan = FrontendPlan(sr, sts, 0)
    wins = plan.windows(len(audio), 1)
    n = len(wins)
    assert n <= max_batch
    feats = eng.features(plan, audio, wins)
    outs = {}
    monkeypatch.setenv("WSB_ATTN_THREADS", "128")     # the persistent kernel's attention units are decode.cu's 128-thread form
    for mode in ("mega", "launches"):
        if mode == "launches":
            monkeypatch.delenv("WSB_MEGA", raising=False)
        else:
            monkeypatch.setenv("WSB_MEGA", "1")
        for graph in (False, True):
            eng.encode(feats)
            ids, steps = eng.generate(n, tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, 120, use_graph=graph)
            outs[(mode, graph)] = (ids.cpu(), steps)
    ref_ids, ref_steps = outs[("launches", False)]
    lens = (ref_ids != tok.eos_token_id).sum(dim=1)
    print("%d windows: row lengths min %d median %d max %d, %d positions" % (n, lens.min(), lens.median(), lens.max(), ref_steps))
    for key, (ids, steps) in outs.items():
        assert steps == ref_steps, key
        assert torch.equal(ids, ref_ids), "tokens differ for %s" % (key,)


def test_mega_large_arch_small_batch(monkeypatch):
    """whisper-large widths (d 1280, ffn 5120: 8 x 10 KB weight rows per fc2 job) on 6 windows, confident recipe."""
    import tempfile
    import torch
    from tools import synth
    from whisperseg_b200.frontend import FrontendPlan
    from whisperseg_b200.segmenter import WhisperSegmenter
    state = synth.make_state("large", seed=0, calibrate="file")
    tokdir = tempfile.mkdtemp()
    synth.token_table_files(tokdir)
    seg = WhisperSegmenter.from_state(state, tokdir, device="cuda", device_ids=[0], max_batch=8)
    eng, tok = seg.engines[0], seg.tokenizer
    audio = synth.synth_audio(15.0, 48000, seed=2)
    plan = FrontendPlan(48000, 0.0025, 0)
    wins = plan.windows(len(audio), 1)
    feats = eng.features(plan, audio, wins)
    outs = {}
    monkeypatch.setenv("WSB_ATTN_THREADS", "128")
    for mode in ("mega", "launches"):
        if mode == "launches":
            monkeypatch.delenv("WSB_MEGA", raising=False)
        else:
            monkeypatch.setenv("WSB_MEGA", "1")
        eng.encode(feats)
        ids, steps = eng.generate(len(wins), tok.prompt_ids, tok.eos_token_id, tok.pad_token_id, 100)
        outs[mode] = (ids.cpu(), steps)
    assert outs["mega"][1] == outs["launches"][1]
    assert torch.equal(outs["mega"][0], outs["launches"][0])
