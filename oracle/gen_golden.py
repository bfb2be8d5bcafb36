"""TEST INFRASTRUCTURE ONLY -- produce tests/golden/* by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden
The GPU box has no /root/reference, so the fixtures written here are what pins the oracle there.

Fixtures:
  frontend_synth.npz   reference get_sliced_audios_features (model.py:127-166) on seeded synthetic
                       audio for the species configurations of config/segment_config.json
  frontend_wav.npz     the same on short excerpts of the labelled example recordings
  postprocess.json     reference parse_generation + segment() tail on scripted token streams
  model_tiny.npz       reference WhisperSegmenterForEval.segment end to end (HF torch path) on a
                       seeded shaped whisper-tiny-architecture checkpoint: ids, texts, segments
"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from tools import synth  # noqa: E402
from oracle.ref_shim import GenerateAdapter, import_reference  # noqa: E402

FRONTEND_CASES = [
    # name, sr, spec_time_step, min_frequency, seconds, num_trials, keep windows
    ("human16k", 16000, 0.01, 0, 23.7, 1, [0, 2]),
    ("zebra32k", 32000, 0.0025, 0, 7.3, 3, [0, 2, 3, 6, 10]),
    ("marmoset48k", 48000, 0.0025, 0, 4.1, 1, [0, 1]),
    ("canary44k", 44100, 0.0025, 0, 3.0, 1, [0, 1]),
    ("meerkat16k", 16000, 0.001, 500, 2.2, 2, [1, 2]),
    ("bat250k", 250000, 0.0005, 20000, 0.8, 1, [0, 1]),
    ("mouse300k", 300000, 0.0005, 35000, 0.7, 1, [1]),
    ("empty16k", 16000, 0.01, 0, 0.0, 1, [0]),
    ("frac22k", 22050, 0.0029, 100, 4.0, 2, [0, 2]),
]


def _store_feat(out, key, feat):
    """Keep fixtures small: a 2x2-decimated grid plus the full first/last 6 columns (the
    reflect-padded edges) and the column/row sums (which see every element)."""
    out[key + "_grid"] = feat[::2, ::2]
    out[key + "_head"] = feat[:, :6]
    out[key + "_tail"] = feat[:, -6:]
    out[key + "_colsum"] = feat.astype(np.float64).sum(axis=0)
    out[key + "_rowsum"] = feat.astype(np.float64).sum(axis=1)


def gen_frontend(ref_model):
    seg = ref_model.SegmenterBase()
    seg.total_spec_columns = 1000
    out = {}
    meta = []
    for k, (name, sr, sts, mf, secs, nt, keep) in enumerate(FRONTEND_CASES):
        audio = synth.synth_audio(secs, sr, seed=100 + k)
        feats = seg.get_sliced_audios_features(audio, sr, mf, sts, nt)
        out[name + "_audio_crc"] = np.array([zlib.crc32(audio.tobytes()), len(audio)], dtype=np.int64)
        out[name + "_plan"] = np.array([[f[0], f[1], f[3]] for f in feats], dtype=np.float64)
        for w in keep:
            _store_feat(out, "%s_feat%d" % (name, w), feats[w][2])
        meta.append(dict(name=name, sr=sr, spec_time_step=sts, min_frequency=mf, seconds=secs, num_trials=nt,
                         keep=keep, n_windows=len(feats)))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLDEN, "frontend_synth.npz"), **out)
    print("frontend_synth:", len(meta), "cases")


def gen_frontend_wav(ref_model):
    from glob import glob
    from scipy.io import wavfile
    seg = ref_model.SegmenterBase()
    seg.total_spec_columns = 1000
    base = os.path.join(os.path.dirname(ref_model.__file__), "data", "example_subset")
    picks = []
    for sub in ("Zebra_finch/test_adults", "Bengalese_finch/test", "Canary/test", "Meerkat/test"):
        wavs = sorted(glob(os.path.join(base, sub, "*.wav")))
        if wavs:
            picks.append(wavs[0])
    out, meta = {}, []
    for k, path in enumerate(picks):
        sr, x = wavfile.read(path)
        label = json.load(open(path[:-4] + ".json"))
        sts, mf = label["spec_time_step"], label["min_frequency"]
        n = min(len(x), int(1.6 * 1000 * sts * sr))          # 1.6 windows: one full, one partial
        x16 = np.ascontiguousarray(x[:n])
        audio = (x16.astype(np.float32) / 32768.0)
        feats = seg.get_sliced_audios_features(audio, sr, mf, sts, 1)
        name = "wav%d" % k
        out[name + "_pcm16"] = x16
        for w in range(len(feats)):
            _store_feat(out, "%s_feat%d" % (name, w), feats[w][2])
        meta.append(dict(name=name, sr=int(sr), spec_time_step=sts, min_frequency=mf, n_windows=len(feats),
                         source=os.path.relpath(path, os.path.dirname(ref_model.__file__))))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLDEN, "frontend_wav.npz"), **out)
    print("frontend_wav:", [m["source"] for m in meta])


def _scripted_texts(rng, n_windows, n_clusters, boundary_merge_prob=0.3):
    """Random but grammar-valid generations with the edge cases the parser must survive."""
    texts = []
    carry = None
    for w in range(n_windows):
        toks, t = [], 0
        if carry is not None:
            toks.append("<|0|>%d<|%d|>" % (carry, int(rng.integers(1, 40))))
            t = 60
            carry = None
        while t < 470:
            t += int(rng.integers(1, 60))
            dur = int(rng.integers(0, 50))              # 0 => zero-length, dropped
            cid = int(rng.integers(0, n_clusters + 1))  # n_clusters => unknown id, dropped
            if t + dur > 500:
                break
            toks.append("<|%d|>%d<|%d|>" % (t, cid, t + dur))
            t += dur
        if rng.random() < boundary_merge_prob and w + 1 < n_windows:
            cid = int(rng.integers(0, n_clusters))
            a = int(rng.integers(max(t, 471), 500))
            toks.append("<|%d|>%d<|500|>" % (a, cid))
            carry = cid
        junk = ["", "<|endoftext|>", "<|12|>", "7", "<|3|><|4|>", " x "][int(rng.integers(0, 6))]
        texts.append("<|startoftranscript|><|en|><|notimestamps|>" + "".join(toks) + junk + "<|endoftext|>")
    return texts


def gen_postprocess(ref_model):
    cases = []
    rng = np.random.default_rng(7)
    configs = [
        dict(sr=16000, sts=0.01, seconds=25.0, num_trials=1, codebook={"vocal": 0, "b": 1}, method="clustering"),
        dict(sr=32000, sts=0.0025, seconds=11.3, num_trials=3, codebook={"vocal": 0}, method="clustering"),
        dict(sr=32000, sts=0.0025, seconds=11.3, num_trials=3, codebook={"a": 0, "b": 1, "c": 2}, method="voting"),
        dict(sr=48000, sts=0.0025, seconds=6.0, num_trials=2, codebook={"vocal": 0, "x": 1}, method="clustering"),
        dict(sr=16000, sts=0.001, seconds=3.3, num_trials=5, codebook={"vocal": 0}, method="clustering",
             eps=0.004, min_segment_length=0.003),
        dict(sr=44100, sts=0.0025, seconds=9.9, num_trials=3, codebook={"vocal": 0, "b": 1}, method="voting",
             time_per_frame_for_voting=0.001),
        dict(sr=16000, sts=0.01, seconds=0.0, num_trials=1, codebook={"vocal": 0}, method="clustering"),
        dict(sr=16000, sts=0.01, seconds=31.0, num_trials=3, codebook={"vocal": 0}, method="clustering", empty=True),
    ]
    # the survey's hand-written known-answer case
    survey_texts = ["<|startoftranscript|><|en|><|notimestamps|><|10|>0<|60|><|100|>1<|100|><|450|>0<|500|>",
                    "<|0|>0<|25|><|30|>7<|40|><|200|>1<|201|>", "<|100|>0<|400|>"]
    for ci, c in enumerate(configs):
        seg = ref_model.SegmenterBase()
        seg.total_spec_columns = 1000
        seg.cluster_codebook = c["codebook"]
        audio = np.zeros(int(c["seconds"] * c["sr"]), dtype=np.float32)
        feats = seg.get_sliced_audios_features(audio, c["sr"], 0, c["sts"], c["num_trials"])
        if ci == 0:
            texts = survey_texts
        elif c.get("empty"):
            texts = ["<|endoftext|>"] * len(feats)
        else:
            # the same underlying "events" seen by each trial would be the realistic case; random
            # per-window scripts plus a shared sub-stream exercise both consolidation branches
            texts = _scripted_texts(rng, len(feats), len(c["codebook"]))
            if c["num_trials"] > 1:
                texts = _shared_event_texts(rng, feats, c, texts)
        assert len(texts) == len(feats), (len(texts), len(feats))
        seg.generate_segment_text = lambda *a, _t=texts, **k: list(_t)
        kw = {k: c[k] for k in ("eps", "min_segment_length", "time_per_frame_for_voting") if k in c}
        res = seg.segment(audio, c["sr"], min_frequency=0, spec_time_step=c["sts"], num_trials=c["num_trials"],
                          consolidation_method=c["method"], **kw)
        cases.append(dict(config={k: v for k, v in c.items()}, texts=texts,
                          windows=[[int(f[0]), float(f[1]), float(f[3])] for f in feats],
                          expected={"onset": [float(x) for x in res["onset"]],
                                    "offset": [float(x) for x in res["offset"]],
                                    "cluster": list(res["cluster"])}))
        print("postprocess case", ci, "->", len(res["onset"]), "segments")
    json.dump(cases, open(os.path.join(GOLDEN, "postprocess.json"), "w"))


def _shared_event_texts(rng, feats, c, fallback):
    """Ground-truth events rendered into every trial's windows with +-1 token jitter."""
    sts, dur = c["sts"], c["seconds"]
    events, t = [], 0.05
    while t < dur - 0.1:
        d = float(rng.uniform(8 * sts, 120 * sts))
        events.append((t, min(t + d, dur), int(rng.integers(0, len(c["codebook"])))))
        t += d + float(rng.uniform(4 * sts, 200 * sts))
    texts = []
    for w, f in enumerate(feats):
        off = f[1]
        toks = []
        for a, b, cid in events:
            ta = int(round((a - off) / (2 * sts))) + int(rng.integers(-1, 2)) * (rng.random() < 0.3)
            tb = int(round((b - off) / (2 * sts))) + int(rng.integers(-1, 2)) * (rng.random() < 0.3)
            if tb <= 0 or ta >= 500:
                continue
            ta, tb = max(ta, 0), min(tb, 500)
            if rng.random() < 0.05:
                continue                                   # a trial misses the event
            toks.append("<|%d|>%d<|%d|>" % (ta, cid, tb))
        if rng.random() < 0.15:
            toks.append(fallback[w][len("<|startoftranscript|><|en|><|notimestamps|>"):])   # spurious extras
        texts.append("".join(toks) + "<|endoftext|>")
    return texts


def gen_model(ref_model):
    # the round-1 "stress" recipe (tests/conftest.py: tiny_checkpoint) and the "confident" recipe (peaked logits: the
    # checkpoint on which the north-star >= 0.99 F1 bar is asserted against the unmodified reference's own output)
    _gen_model(ref_model, "model_tiny.npz", {})
    _gen_model(ref_model, "model_tiny_confident.npz", dict(confident=True))


def _gen_model(ref_model, out_name, recipe_kw):
    import torch
    tok = synth.build_tokenizer()
    hf = synth.make_hf_model("tiny", seed=0, default_segmentation_config=dict(
        sr=16000, min_frequency=0, spec_time_step=0.01, species="human"), **recipe_kw)
    seg = ref_model.WhisperSegmenterForEval(model=GenerateAdapter(hf), tokenizer=tok)
    audio = synth.synth_audio(47.0, 16000, seed=11)
    captured = {}
    orig = hf.generate

    def spy(*a, **k):
        ids = orig(*a, **k)
        captured.setdefault("ids", []).append(ids.clone())
        return ids
    hf.generate = spy
    max_length = 96
    res = seg.segment(audio, 16000, num_trials=1, num_beams=1, batch_size=8, max_length=max_length)
    feats = seg.get_sliced_audios_features(audio, 16000, 0, 0.01, 1)
    texts = seg.generate_segment_text(feats, 8, max_length, 1)
    ids = captured["ids"][0]
    with torch.no_grad():
        enc = hf.model.encoder(torch.from_numpy(np.asarray([f[2] for f in feats]))).last_hidden_state
    res3 = seg.segment(audio, 16000, num_trials=3, num_beams=1, batch_size=8, max_length=max_length)
    np.savez_compressed(
        os.path.join(GOLDEN, out_name),
        ids=ids.numpy().astype(np.int32),
        enc_probe=enc[:, ::50, ::16].numpy(),
        texts=np.frombuffer(json.dumps(texts).encode(), dtype=np.uint8),
        segments=np.frombuffer(json.dumps(res).encode(), dtype=np.uint8),
        segments_trials3=np.frombuffer(json.dumps(res3).encode(), dtype=np.uint8),
        meta=np.frombuffer(json.dumps(dict(arch="tiny", seed=0, audio_seed=11, seconds=47.0, sr=16000,
                                           max_length=max_length)).encode(), dtype=np.uint8))
    print(out_name, ": ids", tuple(ids.shape), "segments", len(res["onset"]), "trials3", len(res3["onset"]))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref_model, _ = import_reference()
    which = sys.argv[1:] or ["frontend", "wav", "postprocess", "model"]
    if "frontend" in which:
        gen_frontend(ref_model)
    if "wav" in which:
        gen_frontend_wav(ref_model)
    if "postprocess" in which:
        gen_postprocess(ref_model)
    if "model" in which:
        gen_model(ref_model)


if __name__ == "__main__":
    main()
