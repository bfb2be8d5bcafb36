"""TEST INFRASTRUCTURE ONLY -- plain torch fp32 restatement of the Whisper network + greedy decode.

CPU oracle for the encoder (K2-K4) and decoder (K5) kernels.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it.

The arithmetic restated here lives in a third-party dependency that is NOT under /root/reference:
HuggingFace `transformers` (pinned 4.38.2, requirements.txt:1; 5.5.0 installed), reached from
the reference at model.py:609 / 655 (`model.generate(...)`):
  * WhisperEncoder.forward            -- HF models/whisper/modeling_whisper.py:593-647
  * WhisperEncoderLayer / Attention   -- :361-414, :241-357
  * WhisperDecoder.forward / layer    -- :691-796, :417-506; tied proj_out :964-975
  * generate(): greedy == (do_sample, top_k=1); SuppressTokens / SuppressTokensAtBegin logits
    processors (generation_whisper.py:1774-1812), stop on EOS or max_length (which counts the 3
    prompt tokens), finished rows padded with pad_token_id, prompt stripped from the output.

Weights are a flat {HF state_dict name: tensor} mapping -- the same files the product loads.
Pinned by: tests/test_oracle_model.py compares it against HF `WhisperForConditionalGeneration`
itself (encoder states, teacher-forced logits and `generate` ids) and tests/golden/model_*.npz.
"""
import math

import torch
import torch.nn.functional as F


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def _heads(x, n_heads):
    B, T, D = x.shape
    return x.view(B, T, n_heads, D // n_heads).transpose(1, 2)          # [B,H,T,hd]


def _attn(q, k, v, n_heads, causal_offset=None):
    """q:[B,Tq,D] already scaled; k,v:[B,Tk,D]; plain softmax(QK^T)V (fp32)."""
    qh, kh, vh = _heads(q, n_heads), _heads(k, n_heads), _heads(v, n_heads)
    s = qh @ kh.transpose(-1, -2)
    if causal_offset is not None:
        Tq, Tk = s.shape[-2:]
        i = torch.arange(Tq).view(-1, 1) + causal_offset
        j = torch.arange(Tk).view(1, -1)
        s = s.masked_fill(j > i, float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = p @ vh
    B, H, T, hd = o.shape
    return o.transpose(1, 2).reshape(B, T, H * hd)


class WhisperOracle:
    def __init__(self, state_dict, n_heads, n_layers):
        self.w = {k: v.detach().to(torch.float32) for k, v in state_dict.items()}
        self.H = n_heads
        self.L = n_layers
        self.d = self.w["model.encoder.conv1.weight"].shape[0]
        self.scaling = (self.d // n_heads) ** -0.5

    # ---------------------------------------------------------------- encoder
    def conv_stem(self, feats):
        w = self.w
        x = F.gelu(F.conv1d(feats, w["model.encoder.conv1.weight"], w["model.encoder.conv1.bias"], padding=1))
        x = F.gelu(F.conv1d(x, w["model.encoder.conv2.weight"], w["model.encoder.conv2.bias"], stride=2, padding=1))
        return x.permute(0, 2, 1) + w["model.encoder.embed_positions.weight"]

    def encoder_layer(self, x, i):
        w, p = self.w, "model.encoder.layers.%d." % i
        h = _ln(x, w[p + "self_attn_layer_norm.weight"], w[p + "self_attn_layer_norm.bias"])
        q = F.linear(h, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"]) * self.scaling
        k = F.linear(h, w[p + "self_attn.k_proj.weight"])
        v = F.linear(h, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"])
        a = _attn(q, k, v, self.H)
        x = x + F.linear(a, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
        h = _ln(x, w[p + "final_layer_norm.weight"], w[p + "final_layer_norm.bias"])
        h = F.gelu(F.linear(h, w[p + "fc1.weight"], w[p + "fc1.bias"]))
        return x + F.linear(h, w[p + "fc2.weight"], w[p + "fc2.bias"])

    @torch.no_grad()
    def encode(self, feats, return_all=False):
        """feats f32 [B,80,1000] -> hidden f32 [B,500,d]."""
        x = self.conv_stem(torch.as_tensor(feats, dtype=torch.float32))
        states = [x]
        for i in range(self.L):
            x = self.encoder_layer(x, i)
            states.append(x)
        x = _ln(x, self.w["model.encoder.layer_norm.weight"], self.w["model.encoder.layer_norm.bias"])
        return (x, states) if return_all else x

    # ---------------------------------------------------------------- decoder
    @torch.no_grad()
    def cross_kv(self, enc):
        kv = []
        for i in range(self.L):
            p = "model.decoder.layers.%d.encoder_attn." % i
            kv.append((F.linear(enc, self.w[p + "k_proj.weight"]),
                       F.linear(enc, self.w[p + "v_proj.weight"], self.w[p + "v_proj.bias"])))
        return kv

    @torch.no_grad()
    def decode_logits(self, ids, enc=None, cross=None, cache=None, return_hidden=False):
        """Teacher-forced logits f32 [B,T,V] for decoder input ids [B,T] (positions 0..T-1 unless
        `cache` holds past self-attention K/V, in which case ids are the new tokens)."""
        w = self.w
        cross = cross if cross is not None else self.cross_kv(enc)
        ids = torch.as_tensor(ids, dtype=torch.long)
        past = 0 if cache is None or cache[0] is None else cache[0][0].shape[1]
        T = ids.shape[1]
        x = w["model.decoder.embed_tokens.weight"][ids] + w["model.decoder.embed_positions.weight"][past:past + T]
        for i in range(self.L):
            p = "model.decoder.layers.%d." % i
            h = _ln(x, w[p + "self_attn_layer_norm.weight"], w[p + "self_attn_layer_norm.bias"])
            q = F.linear(h, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"]) * self.scaling
            k = F.linear(h, w[p + "self_attn.k_proj.weight"])
            v = F.linear(h, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"])
            if cache is not None:
                if cache[i] is not None:
                    k = torch.cat([cache[i][0], k], dim=1)
                    v = torch.cat([cache[i][1], v], dim=1)
                cache[i] = (k, v)
            a = _attn(q, k, v, self.H, causal_offset=past)
            x = x + F.linear(a, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
            h = _ln(x, w[p + "encoder_attn_layer_norm.weight"], w[p + "encoder_attn_layer_norm.bias"])
            q = F.linear(h, w[p + "encoder_attn.q_proj.weight"], w[p + "encoder_attn.q_proj.bias"]) * self.scaling
            a = _attn(q, cross[i][0], cross[i][1], self.H)
            x = x + F.linear(a, w[p + "encoder_attn.out_proj.weight"], w[p + "encoder_attn.out_proj.bias"])
            h = _ln(x, w[p + "final_layer_norm.weight"], w[p + "final_layer_norm.bias"])
            h = F.gelu(F.linear(h, w[p + "fc1.weight"], w[p + "fc1.bias"]))
            x = x + F.linear(h, w[p + "fc2.weight"], w[p + "fc2.bias"])
        if return_hidden == "pre":                       # residual stream before the final LayerNorm (recipe tuning)
            return x
        x = _ln(x, w["model.decoder.layer_norm.weight"], w["model.decoder.layer_norm.bias"])
        if return_hidden:
            return x
        return F.linear(x, w["model.decoder.embed_tokens.weight"])

    @torch.no_grad()
    def greedy(self, enc, prompt, eos_id, pad_id, max_length=448, suppress_tokens=None,
               begin_suppress_tokens=None, return_margins=False):
        """Greedy generation; returns int64 [B, n_generated] (prompt stripped, HF 5.x convention)."""
        B = enc.shape[0]
        cross = self.cross_kv(enc)
        cache = [None] * self.L
        ids = torch.tensor([list(prompt)] * B, dtype=torch.long)
        finished = torch.zeros(B, dtype=torch.bool)
        out, margins = [], []
        cur = ids
        length = ids.shape[1]
        while length < max_length:
            logits = self.decode_logits(cur, cross=cross, cache=cache)[:, -1, :]
            if suppress_tokens is not None and len(suppress_tokens):
                logits[:, list(suppress_tokens)] = float("-inf")
            if begin_suppress_tokens is not None and len(begin_suppress_tokens) and length == len(prompt):
                logits[:, list(begin_suppress_tokens)] = float("-inf")
            if return_margins:
                top2 = logits.topk(2, dim=-1).values
                margins.append(top2[:, 0] - top2[:, 1])
            nxt = logits.argmax(dim=-1)
            nxt = torch.where(finished, torch.full_like(nxt, pad_id), nxt)
            out.append(nxt)
            finished = finished | (nxt == eos_id)
            length += 1
            cur = nxt.view(B, 1)
            if bool(finished.all()):
                break
        res = torch.stack(out, dim=1) if out else torch.zeros(B, 0, dtype=torch.long)
        if return_margins:
            return res, (torch.stack(margins, dim=1) if margins else torch.zeros(B, 0))
        return res


def beam_search(orc, enc, prompt, eos_id, pad_id, max_length=448, num_beams=4, length_penalty=1.0,
                suppress_tokens=None, begin_suppress_tokens=None, return_state=False):
    """HF-equivalent beam search (oracle/beam_np.py) over the oracle network; int64 [B, n] (prompt stripped)."""
    import numpy as np
    from .beam_np import BeamState
    B = enc.shape[0]
    with torch.no_grad():
        cross = [(k.repeat_interleave(num_beams, dim=0), v.repeat_interleave(num_beams, dim=0))
                 for k, v in orc.cross_kv(enc)]
        cache = [None] * orc.L
        st = BeamState(B, num_beams, list(prompt), eos_id, pad_id, max_length, length_penalty)
        cur = torch.tensor([list(prompt)] * (B * num_beams), dtype=torch.long)
        while not st.finished:
            logits = orc.decode_logits(cur, cross=cross, cache=cache)[:, -1, :].numpy()
            parents = torch.from_numpy(st.step(logits, suppress_tokens, begin_suppress_tokens))
            cache = [(k.index_select(0, parents), v.index_select(0, parents)) for k, v in cache]
            cur = torch.from_numpy(st.rows_tokens()).view(-1, 1)
    out = torch.from_numpy(st.result())
    return (out, st) if return_state else out


def oracle_from_hf(hf_model):
    cfg = hf_model.config
    return WhisperOracle(hf_model.state_dict(), cfg.encoder_attention_heads, cfg.encoder_layers)
