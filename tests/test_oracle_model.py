"""The torch restatement of the Whisper network + greedy loop (oracle/whisper_torch.py) is pinned
against (a) HuggingFace transformers' own WhisperForConditionalGeneration -- the third-party code the
reference calls at model.py:609/655 -- and (b) the golden ids/segments produced by the unmodified
reference (tests/golden/model_tiny.npz)."""
import copy
import json
import os

import numpy as np
import pytest
import torch

from oracle import frontend_np as FO
from oracle import postprocess_ref as PR
from tools import synth
from oracle.whisper_torch import WhisperOracle, oracle_from_hf


@pytest.fixture(scope="module")
def tiny(tiny_checkpoint):
    path, hf = tiny_checkpoint
    audio = synth.synth_audio(47.0, 16000, seed=11)
    feats = FO.sliced_audio_features(audio, 16000, 0, 0.01, 1, dtype=np.float32)
    x = torch.from_numpy(np.asarray([f[2] for f in feats]))
    return dict(path=path, hf=hf, orc=oracle_from_hf(hf), audio=audio, feats=feats, x=x)


def test_restatement_exact_in_float64(tiny):
    """In float64 the restatement and HF agree to 1e-9: same function, only rounding differs."""
    hf64 = copy.deepcopy(tiny["hf"]).double()
    x = tiny["x"][:2].double()
    o64 = WhisperOracle(tiny["hf"].state_dict(), 6, 4)
    o64.w = {k: v.double() for k, v in tiny["hf"].state_dict().items()}
    with torch.no_grad():
        enc_hf = hf64.model.encoder(x).last_hidden_state
        h = o64.conv_stem(x)
        for i in range(4):
            h = o64.encoder_layer(h, i)
        import torch.nn.functional as F
        enc = F.layer_norm(h, (h.shape[-1],), o64.w["model.encoder.layer_norm.weight"], o64.w["model.encoder.layer_norm.bias"], 1e-5)
        assert (enc - enc_hf).abs().max().item() < 1e-9
        dec_in = torch.tensor([[synth.ID_SOT, synth.ID_EN, synth.ID_NOTIMESTAMPS, synth.ID_TS0 + 5, 15, synth.ID_TS0 + 9]] * 2)
        lg_hf = hf64(encoder_outputs=(enc_hf,), decoder_input_ids=dec_in).logits
        lg = o64.decode_logits(dec_in, enc=enc)
        assert (lg - lg_hf).abs().max().item() < 1e-8


def test_encoder_vs_hf_fp32(tiny):
    with torch.no_grad():
        enc_hf = tiny["hf"].model.encoder(tiny["x"]).last_hidden_state
    enc = tiny["orc"].encode(tiny["x"])
    assert (enc - enc_hf).abs().max().item() < 2e-3


def test_greedy_vs_golden_reference_ids(tiny, golden_dir):
    """ids produced by the reference's own generate call (HF, do_sample + top_k=1)."""
    g = np.load(golden_dir + "/model_tiny.npz")
    meta = json.loads(bytes(g["meta"]).decode())
    hf, orc = tiny["hf"], tiny["orc"]
    enc = orc.encode(tiny["x"])
    ids, margins = orc.greedy(enc, [synth.ID_SOT, synth.ID_EN, synth.ID_NOTIMESTAMPS], synth.ID_EOT, synth.ID_EOT,
                              meta["max_length"], suppress_tokens=hf.generation_config.suppress_tokens,
                              begin_suppress_tokens=hf.generation_config.begin_suppress_tokens, return_margins=True)
    gold = torch.from_numpy(g["ids"].astype(np.int64))
    n = min(ids.shape[1], gold.shape[1])
    same = (ids[:, :n] == gold[:, :n])
    # fp32 reduction-order differences can flip a near-tie; everything before the first flip must match
    for b in range(ids.shape[0]):
        bad = (~same[b]).nonzero()
        if len(bad):
            assert margins[b, bad[0, 0]].item() < 1e-2, "mismatch at a confident position"
    assert same.float().mean().item() > 0.9


def test_segments_from_golden_texts(tiny, golden_dir):
    g = np.load(golden_dir + "/model_tiny.npz")
    texts = json.loads(bytes(g["texts"]).decode())
    gold = json.loads(bytes(g["segments"]).decode())
    wins = FO.window_plan(len(tiny["audio"]), 16000, 0.01, 1)
    res = PR.segment_from_texts(texts, [[w[0], w[1], w[4]] for w in wins], len(tiny["audio"]), 16000, 0.01,
                                tiny["hf"].config.cluster_codebook, 512)
    assert res == gold


def test_token_table_matches_hf_tokenizer(tiny):
    from whisperseg_b200.tokens import TokenTable
    tok = synth.build_tokenizer()
    table = TokenTable.from_pretrained(tiny["path"])
    assert table.prompt_ids == [synth.ID_SOT, synth.ID_EN, synth.ID_NOTIMESTAMPS]
    assert (table.eos_token_id, table.pad_token_id) == (tok.eos_token_id, tok.pad_token_id)
    rng = np.random.default_rng(0)
    allowed = np.array(synth.allowed_token_ids())
    rows = [allowed[rng.integers(0, len(allowed), size=40)].tolist() for _ in range(20)]
    rows.append([synth.ID_TS0 + 10, 16, 17, synth.ID_TS0 + 60, synth.ID_EOT, synth.ID_EOT, 300, 220, 128, 200, 15])
    rows.append(rng.integers(0, 51372, size=64).tolist())
    assert table.batch_decode(rows) == tok.batch_decode(rows, skip_special_tokens=False)
    # rectangular generate() output: rows padded with EOS (or cut while still emitting) up to max_length
    rect = []
    for k, tail in [(0, synth.ID_EOT), (1, synth.ID_EOT), (7, synth.ID_EOT), (40, synth.ID_EOT), (12, 16), (5, synth.ID_TS0 + 3)]:
        rect.append((allowed[rng.integers(0, len(allowed), size=k)].tolist() + [tail] * 64)[:48])
    rect.append(rng.integers(0, 51372, size=48).tolist())
    arr = np.asarray(rect, dtype=np.int32)
    assert table.batch_decode(arr) == tok.batch_decode(rect, skip_special_tokens=False)
    assert table.batch_decode(arr[:, :1]) == tok.batch_decode([r[:1] for r in rect], skip_special_tokens=False)


def test_weight_preparation_layouts(tiny):
    """conv2's (k, ci) K-ordering over the strided im2col-free view, q pre-scaling, cross-K/V packing."""
    import torch.nn.functional as F
    from whisperseg_b200.weights import load_checkpoint, prepare_tensors
    cfg, sd, gen = load_checkpoint(tiny["path"])
    t = prepare_tensors(cfg, sd, gen, "cpu")
    d, T = cfg["d_model"], cfg["max_source_positions"]
    h1 = torch.randn(1, 1000, d)
    h1p = torch.cat([torch.zeros(1, 1, d), h1], 1).reshape(-1)
    rows = torch.stack([h1p[2 * tt * d:2 * tt * d + 3 * d] for tt in range(T)])          # the strided view
    got = rows @ sd["model.encoder.conv2.weight"].permute(0, 2, 1).reshape(d, 3 * d).t()
    ref = F.conv1d(h1.transpose(1, 2), sd["model.encoder.conv2.weight"], stride=2, padding=1)[0].t()
    assert (got - ref).abs().max().item() < 1e-3
    assert t["enc.conv2.w"].shape == (d, 3 * d) and t["enc.conv1.wt"].shape == (240, d)
    q = sd["model.encoder.layers.0.self_attn.q_proj.weight"] * 0.125
    assert torch.equal(t["enc.0.qkv.w"][:d].float(), q.to(torch.bfloat16).float())
    assert torch.all(t["enc.0.qkv.b"][d:2 * d] == 0)
    L = cfg["encoder_layers"]
    assert t["dec.crosskv.w"].shape == (2 * L * d, d)
    assert (t["dec.suppress"] == 0).sum().item() == len(synth.allowed_token_ids())


@pytest.mark.parametrize("safe", [True, False])
def test_sharded_checkpoint_layouts_load(tiny, tmp_path, safe):
    """HF writes checkpoints above max_shard_size as several files + an index (whisper-large in fp32 does by default);
    the loader must return the same state dict for the single-file and the sharded layout."""
    from whisperseg_b200.weights import load_checkpoint
    hf = tiny["hf"]
    out = str(tmp_path / ("sharded_%s" % safe))
    hf.save_pretrained(out, max_shard_size="20MB", safe_serialization=safe)
    names = os.listdir(out)
    assert any(n.endswith(".index.json") for n in names), names
    cfg, sd, gen = load_checkpoint(out)
    ref_cfg, ref_sd, _ = load_checkpoint(tiny["path"])
    assert cfg["d_model"] == ref_cfg["d_model"] and set(ref_sd) <= set(sd) | {"proj_out.weight"}
    for k, v in ref_sd.items():
        if k in sd:
            assert torch.equal(sd[k], v), k


def test_layernorm_fold_matches_hf_modules(tiny):
    """weights.py: fold_layernorm -- the three LayerNorm-consuming projections of a decoder layer with the LayerNorm
    affine folded in (what csrc/gemv.cu multiplies at <= 64 decode rows): rstd (Wf bf16(x) - mean c1) + c2 must
    reproduce HF's own LayerNorm -> Linear modules (modeling_whisper.py:417-506) within bf16 operand rounding."""
    from whisperseg_b200.weights import load_checkpoint, prepare_tensors
    cfg, sd, gen = load_checkpoint(tiny["path"])
    t = prepare_tensors(cfg, sd, gen, "cpu")
    d = cfg["d_model"]
    layer = tiny["hf"].model.decoder.layers[1]
    torch.manual_seed(5)
    x = torch.randn(24, d) * 2.5 + 0.3
    mean = x.mean(1, keepdim=True)
    rstd = torch.rsqrt(x.var(1, unbiased=False, keepdim=True) + 1e-5)
    xb = x.to(torch.bfloat16).float()

    def folded(name):
        wf, c1, c2 = t["dec.1.%s.wf" % name].float(), t["dec.1.%s.c1" % name], t["dec.1.%s.c2" % name]
        assert torch.equal(c1, wf.sum(1))                 # c1 = row sums of the weights exactly as stored
        return rstd * (xb @ wf.t() - mean * c1) + c2

    with torch.no_grad():
        ln1 = layer.self_attn_layer_norm(x)
        ref_qkv = torch.cat([layer.self_attn.q_proj(ln1) * 0.125, layer.self_attn.k_proj(ln1), layer.self_attn.v_proj(ln1)], 1)
        ref_cq = layer.encoder_attn.q_proj(layer.encoder_attn_layer_norm(x)) * 0.125
        ref_fc1 = layer.fc1(layer.final_layer_norm(x))
    for name, ref in (("sqkv", ref_qkv), ("cq", ref_cq), ("fc1", ref_fc1)):
        got = folded(name)
        err = (got - ref).abs().max().item() / max(1.0, ref.abs().max().item())
        assert got.shape == ref.shape and err < 6e-3, "%s: %g" % (name, err)


@pytest.mark.parametrize("eos_scale,length_penalty,max_length", [(1.0, 1.0, 48), (1.15, 1.0, 40), (1.15, 0.6, 40), (1.3, 2.0, 24)])
def test_beam_oracle_matches_hf_generate(eos_scale, length_penalty, max_length):
    """oracle/beam_np.py (HF beam search restated) against HF `generate(num_beams=4)` -- the reference's
    default decode mode (model.py:409, 614).  HF strips the all-EOS tail column; tokens must be identical."""
    from oracle.whisper_torch import beam_search
    hf = synth.make_hf_model("tiny", seed=3, eos_scale=eos_scale)
    orc = oracle_from_hf(hf)
    audio = synth.synth_audio(15.0, 32000, seed=5)
    feats = FO.sliced_audio_features(audio, 32000, 0, 0.0025, 1, dtype=np.float32)
    x = torch.from_numpy(np.asarray([f[2] for f in feats]))
    prompt = [synth.ID_SOT, synth.ID_EN, synth.ID_NOTIMESTAMPS]
    with torch.no_grad():
        out_hf = hf.generate(input_features=x, decoder_input_ids=torch.tensor([prompt] * x.shape[0]),
                             pad_token_id=synth.ID_EOT, eos_token_id=synth.ID_EOT, max_length=max_length,
                             num_beams=4, do_sample=False, length_penalty=length_penalty)
    out, st = beam_search(orc, orc.encode(x), prompt, synth.ID_EOT, synth.ID_EOT, max_length, 4, length_penalty,
                          hf.generation_config.suppress_tokens, hf.generation_config.begin_suppress_tokens,
                          return_state=True)
    n = min(out.shape[1], out_hf.shape[1])
    assert n >= out.shape[1] - 1
    assert torch.equal(out[:, :n], out_hf[:, :n])
    assert (out[:, n:] == synth.ID_EOT).all()


def test_token_table_reads_the_published_checkpoints_tokenizer_layout(tiny, tmp_path):
    """The published checkpoints (nccratliri/whisperseg-*) ship the slow-tokenizer trio -- vocab.json + added_tokens.json
    (+ merges.txt, tokenizer_config.json), no tokenizer.json -- with `<|0|>..<|1000|>` and the species tokens appended after
    the multilingual Whisper vocabulary (reference model.py:111-113).  Same ids and strings as the tokenizer.json form."""
    import json
    from whisperseg_b200.tokens import TokenTable
    full = json.load(open(os.path.join(tiny["path"], "tokenizer.json")))
    added = {a["content"]: a["id"] for a in full["added_tokens"]}
    pieces = {k: v for k, v in full["model"]["vocab"].items() if k not in added}
    legacy = tmp_path / "legacy"
    legacy.mkdir()
    json.dump(pieces, open(legacy / "vocab.json", "w"))
    json.dump(added, open(legacy / "added_tokens.json", "w"))
    json.dump({"pad_token": "<|endoftext|>", "eos_token": {"content": "<|endoftext|>"}}, open(legacy / "tokenizer_config.json", "w"))
    a = TokenTable.from_pretrained(tiny["path"])
    b = TokenTable.from_pretrained(str(legacy))
    assert (a.prompt_ids, a.eos_token_id, a.pad_token_id) == (b.prompt_ids, b.eos_token_id, b.pad_token_id)
    assert b.convert_tokens_to_ids(["<|0|>", "<|1000|>"]) == [synth.ID_TS0, synth.ID_TS0 + 1000]
    rng = np.random.default_rng(1)
    rows = rng.integers(0, 51372, size=(8, 40)).astype(np.int32)
    rows[:, 30:] = synth.ID_EOT
    assert a.batch_decode(rows) == b.batch_decode(rows)
