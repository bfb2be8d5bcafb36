"""TEST INFRASTRUCTURE ONLY -- CPU model of WHERE the sm_100a engine rounds to bf16.

Not an oracle of the reference (that is whisper_torch.py, fp32) and never on the product path: it is the
fp32 oracle network with a round-to-bf16 inserted at every point where the CUDA engine stores or feeds a
bf16 value (GEMM operands, K/V caches, attention probabilities of the encoder kernel, GELU outputs), all
accumulation in fp32.  It exists to answer, on the CPU and before any GPU time is spent, "how often does
ANY bf16-operand implementation of this network flip a greedy arg-max of THIS checkpoint against fp32?" --
i.e. the noise floor a checkpoint recipe (tools/synth.py) imposes on the north-star token-agreement bar.
The same question asked of the reference's own GPU arm (CTranslate2 fp16, reference model.py:691) has the
same answer up to the 8x finer fp16 mantissa.

Rounding points follow whisperseg_b200/csrc:
  encoder: h1p = bf16(gelu(conv1)); xn = bf16(LN(x)); qkv, att, ff = bf16; P = bf16(exp(s - max)); enc_out = bf16
  decoder: cross K/V = bf16; dxn = bf16(LN(dx)); q, new k/v = bf16; attention output = bf16; dff = bf16(gelu);
           final LN -> bf16 -> tied projection with the bf16 embedding table.
"""
import torch
import torch.nn.functional as F

from .whisper_torch import WhisperOracle, _attn, _heads, _ln


def r16(x):
    return x.to(torch.bfloat16).to(torch.float32)


class Bf16EngineModel(WhisperOracle):
    """WhisperOracle whose matrices are the bf16-rounded ones the engine multiplies with."""

    def __init__(self, state_dict, n_heads, n_layers):
        super().__init__(state_dict, n_heads, n_layers)
        for k, v in self.w.items():
            if v.dim() >= 2 and "embed_positions" not in k and k != "model.encoder.conv1.weight":
                self.w[k] = r16(v)

    def _attn_enc(self, q, k, v):
        qh, kh, vh = _heads(q, self.H), _heads(k, self.H), _heads(v, self.H)
        s = qh @ kh.transpose(-1, -2)
        p = torch.exp(s - s.max(dim=-1, keepdim=True).values)
        o = (r16(p) @ vh) / p.sum(dim=-1, keepdim=True)
        B, H, T, hd = o.shape
        return o.transpose(1, 2).reshape(B, T, H * hd)

    def conv_stem(self, feats):
        w = self.w
        x = r16(F.gelu(F.conv1d(feats, w["model.encoder.conv1.weight"], w["model.encoder.conv1.bias"], padding=1)))
        x = F.gelu(F.conv1d(x, w["model.encoder.conv2.weight"], w["model.encoder.conv2.bias"], stride=2, padding=1))
        return x.permute(0, 2, 1) + w["model.encoder.embed_positions.weight"]

    def encoder_layer(self, x, i):
        w, p = self.w, "model.encoder.layers.%d." % i
        h = r16(_ln(x, w[p + "self_attn_layer_norm.weight"], w[p + "self_attn_layer_norm.bias"]))
        q = r16(F.linear(h, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"]) * self.scaling)
        k = r16(F.linear(h, w[p + "self_attn.k_proj.weight"]))
        v = r16(F.linear(h, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"]))
        a = r16(self._attn_enc(q, k, v))
        x = x + F.linear(a, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
        h = r16(_ln(x, w[p + "final_layer_norm.weight"], w[p + "final_layer_norm.bias"]))
        h = r16(F.gelu(F.linear(h, w[p + "fc1.weight"], w[p + "fc1.bias"])))
        return x + F.linear(h, w[p + "fc2.weight"], w[p + "fc2.bias"])

    @torch.no_grad()
    def encode(self, feats, return_all=False):
        out = super().encode(feats, return_all)
        return (r16(out[0]), out[1]) if return_all else r16(out)

    @torch.no_grad()
    def cross_kv(self, enc):
        return [(r16(k), r16(v)) for k, v in super().cross_kv(enc)]

    @torch.no_grad()
    def decode_logits(self, ids, enc=None, cross=None, cache=None, return_hidden=False):
        w = self.w
        cross = cross if cross is not None else self.cross_kv(enc)
        ids = torch.as_tensor(ids, dtype=torch.long)
        past = 0 if cache is None or cache[0] is None else cache[0][0].shape[1]
        T = ids.shape[1]
        x = w["model.decoder.embed_tokens.weight"][ids] + w["model.decoder.embed_positions.weight"][past:past + T]
        for i in range(self.L):
            p = "model.decoder.layers.%d." % i
            h = r16(_ln(x, w[p + "self_attn_layer_norm.weight"], w[p + "self_attn_layer_norm.bias"]))
            q = r16(F.linear(h, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"]) * self.scaling)
            k = r16(F.linear(h, w[p + "self_attn.k_proj.weight"]))
            v = r16(F.linear(h, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"]))
            if cache is not None:
                if cache[i] is not None:
                    k = torch.cat([cache[i][0], k], dim=1)
                    v = torch.cat([cache[i][1], v], dim=1)
                cache[i] = (k, v)
            a = r16(_attn(q, k, v, self.H, causal_offset=past))
            x = x + F.linear(a, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
            h = r16(_ln(x, w[p + "encoder_attn_layer_norm.weight"], w[p + "encoder_attn_layer_norm.bias"]))
            q = r16(F.linear(h, w[p + "encoder_attn.q_proj.weight"], w[p + "encoder_attn.q_proj.bias"]) * self.scaling)
            a = r16(_attn(q, cross[i][0], cross[i][1], self.H))
            x = x + F.linear(a, w[p + "encoder_attn.out_proj.weight"], w[p + "encoder_attn.out_proj.bias"])
            h = r16(_ln(x, w[p + "final_layer_norm.weight"], w[p + "final_layer_norm.bias"]))
            h = r16(F.gelu(F.linear(h, w[p + "fc1.weight"], w[p + "fc1.bias"])))
            x = x + F.linear(h, w[p + "fc2.weight"], w[p + "fc2.bias"])
        x = _ln(x, w["model.decoder.layer_norm.weight"], w["model.decoder.layer_norm.bias"])
        if return_hidden:
            return x
        return F.linear(r16(x), w["model.decoder.embed_tokens.weight"])
