"""Checkpoint loading and weight preparation for libwsb.

Reads the same directory the reference hands to `WhisperForConditionalGeneration.from_pretrained`
(reference model.py:633-637): `config.json` (+ the WhisperSeg fields `total_spec_columns`,
`cluster_codebook`, `default_segmentation_config`, model.py:639-644), `model.safetensors` or
`pytorch_model.bin` (single file or HF's sharded `*.index.json` layout), optional `generation_config.json` (suppress_tokens / begin_suppress_tokens).

Prepared tensors handed to wsb_model_create (all on the target device):
  enc.conv1.wt  f32 [240][d]    (ci*3+k major, channel contiguous)     enc.conv1.b f32 [d]
  enc.conv1.wg  bf16 [d][512]   (k*160+{0,80}+ci along K, zero tail: conv1 as an im2col-free GEMM over time-major hi+lo bf16 features)
  enc.conv2.w   bf16 [d][3d]    (k*d+ci along K: matches the strided im2col-free view)
  enc.pos f32 [T][d];  enc.ln.{g,b};  per layer l: enc.l.{ln1,ln2}.{g,b}, qkv.{w bf16 [3d][d], b f32 [3d]}
  (q rows pre-scaled by head_dim^-0.5 -- HF multiplies q by `scaling`, modeling_whisper.py:279-310 --
  k has no bias), o / fc1 / fc2 {w bf16 [out][in], b f32}.
  dec.emb bf16 [V][d] (tied proj_out), dec.pos f32 [448][d], dec.crosskv.{w bf16 [L*2*d][d], b},
  per layer: ln1/ln2/ln3, sqkv, so, cq (pre-scaled), co, fc1, fc2;  dec.ln;  dec.suppress /
  dec.begin_suppress f32 [V] additive masks (0 or -inf).
"""
import json
import os

import torch


def load_checkpoint(model_path):
    cfg = json.load(open(os.path.join(model_path, "config.json")))
    st = os.path.join(model_path, "model.safetensors")
    sd = {}
    if os.path.isfile(st):
        from safetensors.torch import load_file
        sd = load_file(st)
    elif os.path.isfile(st + ".index.json"):            # sharded save (whisper-large in fp32 exceeds HF's 5 GB shard size)
        from safetensors.torch import load_file
        for shard in sorted(set(json.load(open(st + ".index.json"))["weight_map"].values())):
            sd.update(load_file(os.path.join(model_path, shard)))
    elif os.path.isfile(os.path.join(model_path, "pytorch_model.bin.index.json")):
        index = json.load(open(os.path.join(model_path, "pytorch_model.bin.index.json")))
        for shard in sorted(set(index["weight_map"].values())):
            sd.update(torch.load(os.path.join(model_path, shard), map_location="cpu", weights_only=True))
    else:
        sd = torch.load(os.path.join(model_path, "pytorch_model.bin"), map_location="cpu", weights_only=True)
    gen = {}
    gp = os.path.join(model_path, "generation_config.json")
    if os.path.isfile(gp):
        gen = json.load(open(gp))
    if "proj_out.weight" in sd and "model.decoder.embed_tokens.weight" not in sd:
        sd["model.decoder.embed_tokens.weight"] = sd["proj_out.weight"]
    return cfg, sd, gen


def logits_masks(cfg, gen, vocab):
    """Additive masks equivalent to HF's SuppressTokens / SuppressTokensAtBegin processors."""
    def pick(key):
        if key in gen:
            return gen[key]
        return cfg.get(key)
    sup = torch.zeros(vocab, dtype=torch.float32)
    beg = torch.zeros(vocab, dtype=torch.float32)
    s = pick("suppress_tokens")
    if s:
        sup[torch.tensor(s, dtype=torch.long)] = float("-inf")
    b = pick("begin_suppress_tokens")
    if b:
        beg[torch.tensor(b, dtype=torch.long)] = float("-inf")
    return sup, beg


def prepare_tensors(cfg, sd, gen, device):
    d = cfg["d_model"]
    L = cfg["encoder_layers"]
    H = cfg["encoder_attention_heads"]
    assert cfg["decoder_layers"] == L and cfg["decoder_attention_heads"] == H, "symmetric Whisper only"
    T = cfg["max_source_positions"]
    V = cfg["vocab_size"]
    scale = float((d // H) ** -0.5)
    out = {}

    def f32(t):
        return t.detach().to(torch.float32).contiguous().to(device)

    def bf16(t):
        return t.detach().to(torch.float32).to(torch.bfloat16).contiguous().to(device)

    g = lambda k: sd[k].to(torch.float32)   # noqa: E731
    w1 = g("model.encoder.conv1.weight")                              # [d][80][3]
    out["enc.conv1.wt"] = f32(w1.permute(1, 2, 0).reshape(-1, d))
    out["enc.conv1.b"] = f32(g("model.encoder.conv1.bias"))
    # conv1 as a tcgen05 GEMM over time-major bf16 features split into hi + lo parts (160 per frame):
    # Wg[co][k * 160 + ci] = Wg[co][k * 160 + 80 + ci] = w[co][ci][k], K padded from 480 to 512
    wg = torch.zeros(d, 512, dtype=torch.float32)
    wk = w1.permute(0, 2, 1)                                   # [d][3][80]
    wg[:, :480] = torch.cat([wk, wk], dim=2).reshape(d, 480)
    out["enc.conv1.wg"] = bf16(wg)
    w2 = g("model.encoder.conv2.weight")                              # [d][d][3]
    out["enc.conv2.w"] = bf16(w2.permute(0, 2, 1).reshape(d, 3 * d))
    out["enc.conv2.b"] = f32(g("model.encoder.conv2.bias"))
    out["enc.pos"] = f32(g("model.encoder.embed_positions.weight")[:T])
    out["enc.ln.g"] = f32(g("model.encoder.layer_norm.weight"))
    out["enc.ln.b"] = f32(g("model.encoder.layer_norm.bias"))

    def fold_layernorm(w, b, gamma, beta):
        """LayerNorm affine folded into the projection that consumes it (decode at <= 64 rows, csrc/gemv.cu):
        W LN(x) + b = rstd (Wf x - mean c1) + c2 with Wf = W o gamma, c1 = row sums of Wf *as rounded to bf16*
        (the kernel multiplies exactly those values), c2 = b + W beta."""
        wf = (w * gamma[None, :]).to(torch.bfloat16)
        c1 = wf.to(torch.float32).sum(1)
        c2 = (b if b is not None else 0.0) + w @ beta
        return wf.contiguous().to(device), c1.contiguous().to(device), c2.to(torch.float32).contiguous().to(device)

    def attn_qkv(prefix):
        qw, kw, vw = g(prefix + "q_proj.weight") * scale, g(prefix + "k_proj.weight"), g(prefix + "v_proj.weight")
        qb, vb = g(prefix + "q_proj.bias") * scale, g(prefix + "v_proj.bias")
        return torch.cat([qw, kw, vw], 0), torch.cat([qb, torch.zeros_like(qb), vb], 0)

    for l in range(L):
        p, o = "model.encoder.layers.%d." % l, "enc.%d." % l
        out[o + "ln1.g"], out[o + "ln1.b"] = f32(g(p + "self_attn_layer_norm.weight")), f32(g(p + "self_attn_layer_norm.bias"))
        w, b = attn_qkv(p + "self_attn.")
        out[o + "qkv.w"], out[o + "qkv.b"] = bf16(w), f32(b)
        out[o + "o.w"], out[o + "o.b"] = bf16(g(p + "self_attn.out_proj.weight")), f32(g(p + "self_attn.out_proj.bias"))
        out[o + "ln2.g"], out[o + "ln2.b"] = f32(g(p + "final_layer_norm.weight")), f32(g(p + "final_layer_norm.bias"))
        out[o + "fc1.w"], out[o + "fc1.b"] = bf16(g(p + "fc1.weight")), f32(g(p + "fc1.bias"))
        out[o + "fc2.w"], out[o + "fc2.b"] = bf16(g(p + "fc2.weight")), f32(g(p + "fc2.bias"))

    out["dec.emb"] = bf16(g("model.decoder.embed_tokens.weight"))
    out["dec.pos"] = f32(g("model.decoder.embed_positions.weight"))
    out["dec.ln.g"] = f32(g("model.decoder.layer_norm.weight"))
    out["dec.ln.b"] = f32(g("model.decoder.layer_norm.bias"))
    ckw, ckb = [], []
    for l in range(L):
        p, o = "model.decoder.layers.%d." % l, "dec.%d." % l
        out[o + "ln1.g"], out[o + "ln1.b"] = f32(g(p + "self_attn_layer_norm.weight")), f32(g(p + "self_attn_layer_norm.bias"))
        w, b = attn_qkv(p + "self_attn.")
        out[o + "sqkv.w"], out[o + "sqkv.b"] = bf16(w), f32(b)
        out[o + "sqkv.wf"], out[o + "sqkv.c1"], out[o + "sqkv.c2"] = fold_layernorm(
            w, b, g(p + "self_attn_layer_norm.weight"), g(p + "self_attn_layer_norm.bias"))
        out[o + "so.w"], out[o + "so.b"] = bf16(g(p + "self_attn.out_proj.weight")), f32(g(p + "self_attn.out_proj.bias"))
        out[o + "ln2.g"], out[o + "ln2.b"] = f32(g(p + "encoder_attn_layer_norm.weight")), f32(g(p + "encoder_attn_layer_norm.bias"))
        out[o + "cq.w"] = bf16(g(p + "encoder_attn.q_proj.weight") * scale)
        out[o + "cq.b"] = f32(g(p + "encoder_attn.q_proj.bias") * scale)
        out[o + "cq.wf"], out[o + "cq.c1"], out[o + "cq.c2"] = fold_layernorm(
            g(p + "encoder_attn.q_proj.weight") * scale, g(p + "encoder_attn.q_proj.bias") * scale,
            g(p + "encoder_attn_layer_norm.weight"), g(p + "encoder_attn_layer_norm.bias"))
        out[o + "co.w"], out[o + "co.b"] = bf16(g(p + "encoder_attn.out_proj.weight")), f32(g(p + "encoder_attn.out_proj.bias"))
        out[o + "ln3.g"], out[o + "ln3.b"] = f32(g(p + "final_layer_norm.weight")), f32(g(p + "final_layer_norm.bias"))
        out[o + "fc1.w"], out[o + "fc1.b"] = bf16(g(p + "fc1.weight")), f32(g(p + "fc1.bias"))
        out[o + "fc1.wf"], out[o + "fc1.c1"], out[o + "fc1.c2"] = fold_layernorm(
            g(p + "fc1.weight"), g(p + "fc1.bias"), g(p + "final_layer_norm.weight"), g(p + "final_layer_norm.bias"))
        out[o + "fc2.w"], out[o + "fc2.b"] = bf16(g(p + "fc2.weight")), f32(g(p + "fc2.bias"))
        vb = g(p + "encoder_attn.v_proj.bias")
        ckw += [g(p + "encoder_attn.k_proj.weight"), g(p + "encoder_attn.v_proj.weight")]
        ckb += [torch.zeros_like(vb), vb]
    out["dec.crosskv.w"] = bf16(torch.cat(ckw, 0))
    out["dec.crosskv.b"] = f32(torch.cat(ckb, 0))
    sup, beg = logits_masks(cfg, gen, V)
    out["dec.suppress"], out["dec.begin_suppress"] = sup.to(device), beg.to(device)
    return out
