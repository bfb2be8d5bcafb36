"""The numpy front-end oracle is pinned against the reference's own outputs (golden fixtures made by
oracle/gen_golden.py from the UNMODIFIED reference) and, when /root/reference is present, live."""
import json
import zlib

import numpy as np
import pytest

from oracle import frontend_np as FO
from tools import synth
from oracle.ref_shim import reference_available


def _check(feat, g, key, tol=1e-4):
    assert np.abs(feat[::2, ::2] - g[key + "_grid"]).max() <= tol
    assert np.abs(feat[:, :6] - g[key + "_head"]).max() <= tol
    assert np.abs(feat[:, -6:] - g[key + "_tail"]).max() <= tol
    assert np.abs(feat.astype(np.float64).sum(0) - g[key + "_colsum"]).max() <= 80 * tol
    assert np.abs(feat.astype(np.float64).sum(1) - g[key + "_rowsum"]).max() <= 1000 * tol


def test_oracle_vs_golden_synth(golden_dir):
    g = np.load(golden_dir + "/frontend_synth.npz")
    meta = json.loads(bytes(g["meta"]).decode())
    assert len(meta) >= 8
    for k, m in enumerate(meta):
        audio = synth.synth_audio(m["seconds"], m["sr"], seed=100 + k)
        crc, n = g[m["name"] + "_audio_crc"]
        assert (zlib.crc32(audio.tobytes()), len(audio)) == (crc, n)
        plan = FO.window_plan(len(audio), m["sr"], m["spec_time_step"], m["num_trials"])
        assert [[p[0], p[1], p[4]] for p in plan] == g[m["name"] + "_plan"].tolist()
        feats = FO.sliced_audio_features(audio, m["sr"], m["min_frequency"], m["spec_time_step"], m["num_trials"])
        for w in m["keep"]:
            _check(feats[w][2], g, "%s_feat%d" % (m["name"], w))


def test_oracle_vs_golden_wav(golden_dir):
    g = np.load(golden_dir + "/frontend_wav.npz")
    for m in json.loads(bytes(g["meta"]).decode()):
        audio = g[m["name"] + "_pcm16"].astype(np.float32) / 32768.0
        feats = FO.sliced_audio_features(audio, m["sr"], m["min_frequency"], m["spec_time_step"], 1)
        assert len(feats) == m["n_windows"]
        for w in range(len(feats)):
            _check(feats[w][2], g, "%s_feat%d" % (m["name"], w))


def test_product_planning_matches_oracle():
    """Host-side planning of the product (window list, filterbank, n_fft) vs the oracle restatement."""
    from whisperseg_b200.frontend import FrontendPlan
    for sr, sts, mf, n, nt in [(16000, 0.01, 0, 400001, 1), (32000, 0.0025, 0, 233601, 3), (44100, 0.0025, 300, 90000, 2),
                               (300000, 0.0005, 35000, 210000, 1), (22050, 0.0029, 100, 88200, 2), (16000, 0.01, 0, 0, 1),
                               (48000, 0.0025, 0, 28800000, 1)]:
        plan = FrontendPlan(sr, sts, mf)
        ref = FO.FrontendConfig(sr, sts, mf)
        assert (plan.hop, plan.n_fft, plan.clip_len) == (ref.hop, ref.n_fft, ref.clip_len)
        assert np.array_equal(plan.mel_filters.astype(np.float32), ref.mel_filters.astype(np.float32))
        wins = plan.windows(n, nt)
        exp = FO.window_plan(n, sr, sts, nt)
        assert [(w.trial_id, w.offset_time, w.start, w.n_valid, w.clip_seconds) for w in wins] == exp
        assert (plan.mel_filters > 0).sum(axis=1).max() <= 2      # sparse: <= 2 triangles per FFT bin


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted (GPU box)")
def test_oracle_vs_live_reference():
    from oracle.ref_shim import import_reference
    ref_model, ref_au = import_reference()
    seg = ref_model.SegmenterBase()
    seg.total_spec_columns = 1000
    audio = synth.synth_audio(5.3, 32000, seed=77)
    ref = seg.get_sliced_audios_features(audio, 32000, 0, 0.0025, 2)
    mine = FO.sliced_audio_features(audio, 32000, 0, 0.0025, 2, dtype=np.float32)
    assert len(ref) == len(mine)
    for a, b in zip(ref, mine):
        assert (a[0], a[1], a[3]) == (b[0], b[1], b[3])
        assert np.abs(a[2] - b[2]).max() < 1e-4
    for sr in (8000, 32000, 32001, 80000, 150000, 300000, 300001):
        assert ref_au.get_n_fft_given_sr(sr) == FO.get_n_fft_given_sr(sr)
