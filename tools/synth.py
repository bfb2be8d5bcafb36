"""BENCH / TEST DATA -- seeded synthetic checkpoints, tokenizer and audio (not part of the product,
not part of the oracle: nothing here restates the reference's algorithm).

There is no network, so neither the trained WhisperSeg checkpoints nor the Whisper tokenizer
files exist here.  This module builds, deterministically from a seed:

  * an offline `WhisperTokenizer` with the *real* multilingual Whisper id layout for every id
    the path touches (reference model.py:111-113: `<|i|>` tokens are appended after the 50364
    base vocabulary, so `<|0|>` == 50364; digits '0'..'9' are GPT-2 byte tokens 15..24);
  * a random-init `WhisperForConditionalGeneration` of a named architecture with
    `max_source_positions = total_spec_columns/2` (reference model.py:79-84) and the config
    contract the segmenter reads (model.py:590-595);
  * a weight-shaping recipe so greedy decoding is content dependent
    and terminates with EOS -- with HF's default init the decoder ignores the audio and every
    downstream parity check would be vacuous;
  * synthetic "vocal-like" audio.

Both the reference arm (HF torch) and the CUDA path load the SAME checkpoint directory written
by `save_checkpoint`, the way `WhisperSegmenter(model_path)` does (reference model.py:626-644).
"""
import json
import os

import numpy as np

ARCHS = {
    #         d_model layers heads ffn
    "tiny":  (384, 4, 6, 1536),
    "base":  (512, 6, 8, 2048),
    "small": (768, 12, 12, 3072),
    "medium": (1024, 24, 16, 4096),
    "large": (1280, 32, 20, 5120),
}
VOCAB_SIZE = 51865
TOTAL_SPEC_COLUMNS = 1000
ID_EOT = 50257
ID_SOT = 50258
ID_EN = 50259
ID_NOTIMESTAMPS = 50363
ID_TS0 = 50364            # "<|0|>"
ID_DIGIT0 = 15            # "0" in the GPT-2 byte alphabet
SPECIES = ["<|zebra_finch|>", "<|bengalese_finch|>", "<|mouse|>", "<|marmoset|>", "<|human|>",
           "<|unknown|>", "<|animal|>"]


def bytes_to_unicode():
    """GPT-2 byte<->unicode alphabet (published algorithm, openai/gpt-2 encoder.py)."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + \
        list(range(ord("\xae"), ord("\xff") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def build_tokenizer():
    from transformers import WhisperTokenizer
    alphabet = list(bytes_to_unicode().values())
    vocab = {ch: i for i, ch in enumerate(alphabet)}
    i = 256
    while len(vocab) < ID_EOT:
        vocab["~f%d" % i] = len(vocab)
        i += 1
    tok = WhisperTokenizer(vocab=vocab, merges=[], pad_token="<|endoftext|>")
    assert tok.convert_tokens_to_ids("<|endoftext|>") == ID_EOT
    specials = ["<|startoftranscript|>", "<|en|>"] + ["<|lang%d|>" % k for k in range(1, 99)] + \
        ["<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>", "<|nocaptions|>",
         "<|notimestamps|>"]
    tok.add_tokens(specials, special_tokens=True)
    # reference model.py:112-113
    tok.add_tokens(["<|%d|>" % i for i in range(TOTAL_SPEC_COLUMNS + 1)], special_tokens=True)
    tok.add_tokens(SPECIES, special_tokens=True)
    assert tok.convert_tokens_to_ids(["<|startoftranscript|>", "<|en|>", "<|notimestamps|>", "<|0|>"]) == \
        [ID_SOT, ID_EN, ID_NOTIMESTAMPS, ID_TS0]
    return tok


def allowed_token_ids(ts_step=25, n_digits=4):
    """Ids the shaped model may emit: every `ts_step`-th timestamp token <|0|>..<|1000|>, the first
    `n_digits` digit tokens and EOS.  The size of this set sets the statistics of the random model's
    output: P(EOS) = 1/len -> geometric decode lengths with mean ~len (the labelled marmoset example
    has 7.6 segments ~ 24 tokens per 2.5 s window), P(digit) = n_digits/len -> how often the
    `<|on|>cluster<|off|>` grammar is hit by chance."""
    return sorted(list(range(ID_TS0, ID_TS0 + TOTAL_SPEC_COLUMNS + 1, ts_step)) +
                  list(range(ID_DIGIT0, ID_DIGIT0 + n_digits)) + [ID_EOT])


def round_weights_bf16_(sd):
    """Make the checkpoint bf16-exact: every tensor the engine multiplies as bf16 (all matrices except conv1, which
    it keeps in fp32, and the fp32 position tables) is rounded to a bf16-representable fp32 value.  The engine's own
    conversion is then lossless, so engine-vs-fp32-oracle differences measure the KERNELS' arithmetic, not the
    weight quantisation (which accounts for half of the bf16-vs-fp32 noise of a random network:
    tools/noise_floor.py).  The engine's outputs are bit-identical with and without this rounding."""
    import torch
    for k, v in sd.items():
        if v.dim() >= 2 and "embed_positions" not in k and not k.endswith("encoder.conv1.weight"):
            sd[k] = v.to(torch.float32).to(torch.bfloat16).to(torch.float32)
    return sd


def script_tokens(seed, n_segments, ts_step=25, n_allowed_digits=4):
    """A seeded, grammar-valid token script `<|on|> cluster <|off|>` x n_segments + EOS with increasing times."""
    rng = np.random.default_rng(7000 + seed)
    n_ts = TOTAL_SPEC_COLUMNS // ts_step + 1
    times = np.sort(rng.choice(n_ts, size=2 * n_segments, replace=False))
    toks = []
    for k in range(n_segments):
        toks += [ID_TS0 + int(times[2 * k]) * ts_step, ID_DIGIT0 + int(rng.integers(0, n_allowed_digits)),
                 ID_TS0 + int(times[2 * k + 1]) * ts_step]
    return toks + [ID_EOT]


def script_vectors(emb, boost, seed, n_segments, prompt_len=3, ts_step=25, n_allowed_digits=4):
    """[448, d] term added to the decoder position table of the "confident" recipe: at decoder position p
    (p >= prompt_len - 1) `boost` along the unit embedding of the scripted NEXT token; after the script, EOS.
    Like the EOS ramp it rides the residual stream into the tied output projection, where it lifts the scripted
    token's logit by boost / rms(h) * |E| -- about `boost / rms(h)` standard deviations of the random logits.  The
    greedy decode then follows the script except where the audio-dependent random part of the logits beats the boost
    (a tail event whose probability the boost sets): most positions carry a wide top-1/top-2 margin, as a trained
    segmenter's do, and the remaining ones depend on the audio through the whole encoder/decoder stack."""
    import torch
    toks = script_tokens(seed, n_segments, ts_step, n_allowed_digits)
    out = torch.zeros(448, emb.shape[1], dtype=torch.float32)
    for p in range(prompt_len - 1, 448):
        j = min(p - (prompt_len - 1), len(toks) - 1)
        u = emb[toks[j]].float()
        out[p] = boost * u / u.norm()
    return out


# "confident" recipe constants: script boost = z * (rms of the final pre-LayerNorm hidden state, measured with the fp32
# oracle: tiny 4.68, base 5.53, large 13.1), z chosen so that ~1 % of the positions leave the script
ARCH_SCRIPT_BOOST = {"tiny": 14.0, "base": 16.5, "large": 35.0}

GELU_MEAN = 0.28209479177387814       # E[gelu(h)], h ~ N(0,1)
DEFAULT_CODEBOOK = {"vocal": 0, "b": 1, "c": 2, "d": 3}


def make_hf_model(arch="tiny", seed=0, shaped=True, cluster_codebook=None,
                  default_segmentation_config=None, **shape_kw):
    """Random-init HF Whisper of `arch` with the WhisperSeg config contract."""
    import torch
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    d, L, H, F = ARCHS[arch]
    cfg = WhisperConfig(vocab_size=VOCAB_SIZE, num_mel_bins=80, d_model=d,
                        encoder_layers=L, decoder_layers=L,
                        encoder_attention_heads=H, decoder_attention_heads=H,
                        encoder_ffn_dim=F, decoder_ffn_dim=F,
                        max_source_positions=TOTAL_SPEC_COLUMNS // 2, max_target_positions=448,
                        pad_token_id=ID_EOT, bos_token_id=ID_EOT, eos_token_id=ID_EOT,
                        decoder_start_token_id=ID_SOT, dropout=0.0, attention_dropout=0.0,
                        activation_dropout=0.0)
    cfg.total_spec_columns = TOTAL_SPEC_COLUMNS
    cfg.cluster_codebook = cluster_codebook if cluster_codebook is not None else dict(DEFAULT_CODEBOOK)
    cfg.species_codebook = {s[2:-2]: s for s in SPECIES}
    if default_segmentation_config is not None:
        cfg.default_segmentation_config = default_segmentation_config
    torch.manual_seed(seed)
    model = WhisperForConditionalGeneration(cfg)
    model.eval()
    if shaped:
        ts_step = shape_kw.pop("ts_step", 25)
        n_allowed_digits = shape_kw.pop("n_allowed_digits", 4)
        script_boost = shape_kw.pop("script_boost", None)
        script_segments = shape_kw.pop("script_segments", 8)
        confident = shape_kw.pop("confident", False)
        shape_weights_(model, seed, **shape_kw)
        allowed = set(allowed_token_ids(ts_step, n_allowed_digits))
        if getattr(model, "_wsb_calibrate", True):
            sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
            calibrate_output_bias_(sd, H, L, sorted(allowed), seed)
            with torch.no_grad():
                model.model.decoder.layer_norm.bias.copy_(sd["model.decoder.layer_norm.bias"])
        if confident and script_boost is None:
            script_boost = ARCH_SCRIPT_BOOST[arch]
        if script_boost:
            with torch.no_grad():
                model.model.decoder.embed_positions.weight.add_(script_vectors(
                    model.model.decoder.embed_tokens.weight.detach(), script_boost, seed, script_segments,
                    ts_step=ts_step, n_allowed_digits=n_allowed_digits))
        sup = [i for i in range(VOCAB_SIZE) if i not in allowed]
        model.generation_config.suppress_tokens = sup
        model.generation_config.begin_suppress_tokens = None
        model.config.suppress_tokens = sup
        model.config.begin_suppress_tokens = None
    return model


def shape_weights_(model, seed, n_digits=2, eos_scale=1.3, qk_cross=3.0, qk_self=2.0, emb_std=0.05,
                   digit_scale=1.0, conv_gain=2.0, bias_std=0.02, cancel_gelu_mean=True,
                   dec_pos_std=2.0, cross_out_gain=2.0, calibrate=True, bf16_exact=True):
    """SURVEY.md section 7.2-1 recipe (constants tuned so rows end with EOS at varied lengths)."""
    import torch
    g = torch.Generator().manual_seed(1000 + seed)
    sd = model.state_dict()
    with torch.no_grad():
        for name, p in sd.items():
            if "embed_positions" in name and "encoder" in name:
                continue                                    # keep the sinusoid table
            if "layer_norm" in name:
                continue                                    # identity LN
            if name.endswith("embed_tokens.weight") or name == "proj_out.weight":
                continue
            if name.endswith("decoder.embed_positions.weight"):
                p.copy_(torch.randn(p.shape, generator=g) * dec_pos_std)
                continue
            if p.dim() >= 2:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / fan_in ** 0.5)
                if name.endswith("q_proj.weight") or name.endswith("k_proj.weight"):
                    p.mul_(qk_cross if "encoder_attn" in name else qk_self)
                if name.endswith("encoder.conv1.weight"):
                    p.mul_(conv_gain)
                if name.endswith("encoder_attn.out_proj.weight"):
                    p.mul_(cross_out_gain)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * bias_std)
        if cancel_gelu_mean:
            for name, p in sd.items():
                if name.endswith("fc2.bias"):
                    p.sub_(GELU_MEAN * sd[name[:-4] + "weight"].sum(dim=1))
        emb = model.model.decoder.embed_tokens.weight
        emb.copy_(torch.randn(emb.shape, generator=g) * emb_std)
        emb[ID_DIGIT0:ID_DIGIT0 + n_digits] *= digit_scale
        emb[ID_EOT] *= eos_scale
        if bf16_exact:
            for name, p in sd.items():
                if p.dim() >= 2 and "embed_positions" not in name and not name.endswith("encoder.conv1.weight"):
                    p.copy_(p.to(torch.bfloat16).to(torch.float32))
    model.tie_weights()
    model._wsb_calibrate = calibrate


def save_checkpoint(model, path, tokenizer=None):
    os.makedirs(path, exist_ok=True)
    model.save_pretrained(path, safe_serialization=True)
    (tokenizer or build_tokenizer()).save_pretrained(path)
    return path


def synth_audio(seconds, sr, seed, band=(500.0, 8000.0), burst=(0.03, 0.25), gap=(0.02, 0.4)):
    """Mono float32 in [-1,1]: harmonic chirp bursts over pink-ish noise at -30 dB."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    # pink-ish noise by 1/sqrt(f) shaping of white noise, chunked to bound memory
    out = np.empty(n, dtype=np.float32)
    chunk = 1 << 20
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        w = rng.standard_normal(m)
        spec = np.fft.rfft(w)
        f = np.arange(spec.shape[0], dtype=np.float64)
        f[0] = 1.0
        spec /= np.sqrt(f)
        x = np.fft.irfft(spec, n=m)
        x *= 0.0316 / (np.abs(x).max() + 1e-12)
        out[s:s + m] = x.astype(np.float32)
    t = 0.0
    hi = min(band[1], 0.45 * sr)
    while t < seconds:
        t += rng.uniform(*gap)
        dur = rng.uniform(*burst)
        a, b = int(t * sr), min(n, int((t + dur) * sr))
        if b <= a:
            break
        tt = np.arange(b - a, dtype=np.float64) / sr
        f0 = rng.uniform(band[0], hi / 4)
        f1 = f0 * rng.uniform(0.7, 1.4)
        phase = 2 * np.pi * (f0 * tt + 0.5 * (f1 - f0) / max(dur, 1e-6) * tt * tt)
        sig = np.zeros_like(tt)
        for h in range(1, int(rng.integers(3, 6)) + 1):
            if f0 * h < hi:
                sig += np.sin(h * phase) / h
        env = np.hanning(b - a)
        out[a:b] += (0.5 * sig * env / (np.abs(sig).max() + 1e-12)).astype(np.float32)
        t += dur
    np.clip(out, -1.0, 1.0, out=out)
    return out


# ---------------------------------------------------------------------------------------------
# Direct construction of a shaped checkpoint state (no HF model object): used by bench.py, where
# instantiating a 1.5 B-parameter nn.Module just to overwrite its weights would waste minutes.
def hf_config_dict(arch, cluster_codebook=None, default_segmentation_config=None):
    d, L, H, F = ARCHS[arch]
    cfg = dict(model_type="whisper", vocab_size=VOCAB_SIZE, num_mel_bins=80, d_model=d, encoder_layers=L,
               decoder_layers=L, encoder_attention_heads=H, decoder_attention_heads=H, encoder_ffn_dim=F,
               decoder_ffn_dim=F, max_source_positions=TOTAL_SPEC_COLUMNS // 2, max_target_positions=448,
               pad_token_id=ID_EOT, bos_token_id=ID_EOT, eos_token_id=ID_EOT, decoder_start_token_id=ID_SOT,
               activation_function="gelu", scale_embedding=False, total_spec_columns=TOTAL_SPEC_COLUMNS,
               cluster_codebook=cluster_codebook if cluster_codebook is not None else dict(DEFAULT_CODEBOOK),
               suppress_tokens=None, begin_suppress_tokens=None)
    if default_segmentation_config is not None:
        cfg["default_segmentation_config"] = default_segmentation_config
    return cfg


def sinusoids(length, channels, max_timescale=10000):
    """HF modeling_whisper.py:55-64 (encoder positional table)."""
    import torch
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2, dtype=torch.float32))
    t = torch.arange(length, dtype=torch.float32).view(-1, 1) * inv.view(1, -1)
    return torch.cat([t.sin(), t.cos()], dim=1)


# per-architecture EOS boost: deeper networks need a larger one for rows to terminate (tuned on the GPU
# with tools/tune_large.py: large/3.0 -> median 16, mean ~110 generated tokens, ~5 % of rows never stop)
ARCH_EOS_RAMP = {"large": 0.1}
ARCH_EOS_SCALE = {"tiny": 1.3, "base": 1.3, "small": 1.3, "medium": 1.3, "large": 3.0}


def make_state(arch="large", seed=0, n_digits=2, eos_scale=None, qk_cross=3.0, qk_self=2.0, emb_std=0.05,
               digit_scale=1.0, conv_gain=2.0, ts_step=25, n_allowed_digits=4, bias_std=0.02,
               cancel_gelu_mean=True, dec_pos_std=2.0, cross_out_gain=2.0, calibrate="auto", dtype=None,
               eos_ramp=None, eos_ramp_start=10, bf16_exact=True, confident=False, script_boost=None, script_segments=8,
               dec_common_mode=0.0):
    """(config dict, state dict, generation dict) of a shaped random checkpoint -- same recipe as
    shape_weights_, HF parameter names, drawn tensor by tensor from one seeded CPU generator."""
    import torch
    d, L, H, F = ARCHS[arch]
    if eos_scale is None:
        eos_scale = ARCH_EOS_SCALE.get(arch, 1.3)
    g = torch.Generator().manual_seed(2000 + seed)
    sd = {}

    def mat(name, out_f, in_f, gain=1.0, shape=None):
        w = torch.randn(shape or (out_f, in_f), generator=g) * (gain / in_f ** 0.5)
        sd[name] = w if dtype is None else w.to(dtype)

    def vec(name, n, std=None):
        sd[name] = torch.randn(n, generator=g) * (bias_std if std is None else std)

    def ln(prefix):
        sd[prefix + ".weight"] = torch.ones(d)
        sd[prefix + ".bias"] = torch.zeros(d)

    def attn(prefix, qk, out_gain=1.0):
        mat(prefix + "q_proj.weight", d, d, qk)
        vec(prefix + "q_proj.bias", d)
        mat(prefix + "k_proj.weight", d, d, qk)
        mat(prefix + "v_proj.weight", d, d)
        vec(prefix + "v_proj.bias", d)
        mat(prefix + "out_proj.weight", d, d, out_gain)
        vec(prefix + "out_proj.bias", d)

    def mlp(prefix):
        mat(prefix + "fc1.weight", F, d)
        vec(prefix + "fc1.bias", F)
        mat(prefix + "fc2.weight", d, F)
        vec(prefix + "fc2.bias", d)
        if cancel_gelu_mean:
            # E[gelu(h)] for h ~ N(0,1) is 1/sqrt(4 pi): without this every MLP adds the same constant
            # vector at every position and window, and a handful of tokens win every arg-max
            sd[prefix + "fc2.bias"] = sd[prefix + "fc2.bias"] - GELU_MEAN * sd[prefix + "fc2.weight"].float().sum(dim=1)

    mat("model.encoder.conv1.weight", d, 80 * 3, gain=conv_gain, shape=(d, 80, 3))
    vec("model.encoder.conv1.bias", d)
    mat("model.encoder.conv2.weight", d, d * 3, shape=(d, d, 3))
    vec("model.encoder.conv2.bias", d)
    sd["model.encoder.embed_positions.weight"] = sinusoids(TOTAL_SPEC_COLUMNS // 2, d)
    for i in range(L):
        p = "model.encoder.layers.%d." % i
        ln(p + "self_attn_layer_norm")
        attn(p + "self_attn.", qk_self)
        ln(p + "final_layer_norm")
        mlp(p)
    ln("model.encoder.layer_norm")
    emb = torch.randn(VOCAB_SIZE, d, generator=g) * emb_std
    emb[ID_DIGIT0:ID_DIGIT0 + n_digits] *= digit_scale
    emb[ID_EOT] *= eos_scale
    sd["model.decoder.embed_tokens.weight"] = emb
    sd["model.decoder.embed_positions.weight"] = torch.randn(448, d, generator=g) * dec_pos_std
    for i in range(L):
        p = "model.decoder.layers.%d." % i
        ln(p + "self_attn_layer_norm")
        attn(p + "self_attn.", qk_self)
        ln(p + "encoder_attn_layer_norm")
        attn(p + "encoder_attn.", qk_cross, cross_out_gain)
        ln(p + "final_layer_norm")
        mlp(p)
    ln("model.decoder.layer_norm")
    if bf16_exact:
        round_weights_bf16_(sd)
    allowed = set(allowed_token_ids(ts_step, n_allowed_digits))
    gen = dict(suppress_tokens=[i for i in range(VOCAB_SIZE) if i not in allowed], begin_suppress_tokens=None)
    # calibrate: "file" = committed vector under tools/calibration/ only (bench.py's GPU arm: no oracle
    # code runs there); "auto" = that file when present, else one fp32 forward pass through the oracle
    # network (tests); False = skip
    if calibrate:
        path = calibration_path(arch, seed)
        if os.path.isfile(path):
            import torch
            mean = torch.from_numpy(np.load(path))
            sd["model.decoder.layer_norm.bias"] = sd["model.decoder.layer_norm.bias"].float() - mean
        elif calibrate == "file":
            raise FileNotFoundError("no committed calibration vector for (%s, seed %d): %s" % (arch, seed, path))
        else:
            calibrate_output_bias_(sd, H, L, sorted(allowed), seed)
    if confident:
        # "confident" recipe: the scripted grammar replaces the EOS ramp (see script_vectors)
        eos_ramp = 0.0
        if script_boost is None:
            script_boost = ARCH_SCRIPT_BOOST[arch]
    if script_boost:
        sd["model.decoder.embed_positions.weight"] = sd["model.decoder.embed_positions.weight"] + script_vectors(
            sd["model.decoder.embed_tokens.weight"], script_boost, seed, script_segments, ts_step=ts_step,
            n_allowed_digits=n_allowed_digits)
    if dec_common_mode:
        # a constant added to every channel of the decoder's residual stream: LayerNorm removes it exactly, but it makes
        # |mean| >> std for every row -- the regime in which a LayerNorm folded into the next projection (csrc/gemv.cu)
        # loses precision.  Parity case for the engine's fold guard (tests/test_gpu_fold_guard.py).
        sd["model.decoder.embed_positions.weight"] = sd["model.decoder.embed_positions.weight"] + float(dec_common_mode)
    if eos_ramp is None:
        eos_ramp = ARCH_EOS_RAMP.get(arch, 0.0)
    if eos_ramp:
        sd["model.decoder.embed_positions.weight"] = sd["model.decoder.embed_positions.weight"] + eos_ramp_vectors(
            sd["model.decoder.embed_tokens.weight"], eos_ramp, eos_ramp_start)
    return hf_config_dict(arch), sd, gen


def eos_ramp_vectors(emb, eos_ramp, start=10):
    """[448, d] term added to the decoder position table: eos_ramp * max(0, p - start) along the unit
    EOS embedding.  It rides the residual stream into the tied output projection, so the EOS logit grows
    linearly with the position and every row of a RANDOM checkpoint ends after a bounded number of
    tokens -- as a trained segmenter's do (2.5 s windows hold ~8 segments = ~24 tokens).
    Without it ~5 % of the rows never emit EOS and run to max_length."""
    import torch
    u = emb[ID_EOT].float()
    u = u / u.norm()
    ramp = torch.clamp(torch.arange(448, dtype=torch.float32) - float(start), min=0.0)
    return eos_ramp * ramp[:, None] * u[None, :]


def calibration_path(arch, seed):
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "calibration", "%s_seed%d.npy" % (arch, seed))


def calibrate_output_bias_(sd, n_heads, n_layers, allowed, seed, n_windows=2, n_positions=24, return_mean=False):
    """Remove the position- and window-independent component of the decoder's final hidden state by
    folding its negative mean into `decoder.layer_norm.bias`.  A random deep network otherwise carries
    a large constant vector to the output projection and the same 3-4 tokens win every arg-max; with
    the bias calibrated the emitted tokens follow the audio (cross-attention) and the position."""
    import torch
    from oracle import frontend_np as FO
    from oracle.whisper_torch import WhisperOracle
    audio = synth_audio(n_windows * 2.5, 32000, seed=9000 + seed)
    feats = FO.sliced_audio_features(audio, 32000, 0, 0.0025, 1, dtype=np.float32)
    x = torch.from_numpy(np.asarray([f[2] for f in feats]))
    orc = WhisperOracle(sd, n_heads, n_layers)
    enc = orc.encode(x)
    rng = np.random.default_rng(9000 + seed)
    ids = torch.from_numpy(rng.choice(np.array(allowed), size=(x.shape[0], n_positions)))
    ids[:, :3] = torch.tensor([ID_SOT, ID_EN, ID_NOTIMESTAMPS])
    hidden = orc.decode_logits(ids, enc=enc, return_hidden=True)
    mean = hidden[:, 3:].reshape(-1, hidden.shape[-1]).mean(dim=0)
    sd["model.decoder.layer_norm.bias"] = sd["model.decoder.layer_norm.bias"].float() - mean
    return mean


def token_table_files(path):
    """Write just the tokenizer files (tokenizer.json ...) so the product can build its TokenTable."""
    os.makedirs(path, exist_ok=True)
    build_tokenizer().save_pretrained(path)
    return path
