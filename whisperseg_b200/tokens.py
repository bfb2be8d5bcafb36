"""Token table: ids <-> strings for the handful of ids the segmentation path needs.

The reference decodes generated ids with `WhisperTokenizer.batch_decode(..., skip_special_tokens=
False)` (model.py:620, 667) and regex-matches `<|on|>cluster<|off|>` triples (model.py:120).  This
module reads the checkpoint's `tokenizer.json` directly (plain JSON; no tokenizer library on the
product path) and reproduces that decoding: byte-level BPE pieces are mapped back to bytes and
decoded as UTF-8 with replacement, added/special tokens are emitted verbatim.
"""
import json
import os

import numpy as np


def _byte_decoder():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return {chr(c): b for b, c in zip(bs, cs)}


class TokenTable:
    PROMPT_TOKENS = ("<|startoftranscript|>", "<|en|>", "<|notimestamps|>")    # model.py:610, 656, 719

    def __init__(self, pieces, added, eos_token="<|endoftext|>", pad_token=None):
        """pieces: {str: id} byte-level vocabulary; added: {str: id} added/special tokens."""
        self.token_to_id = dict(pieces)
        self.token_to_id.update(added)
        size = max(self.token_to_id.values()) + 1
        self.id_to_token = [""] * size
        self.is_added = [False] * size
        for tok, i in pieces.items():
            self.id_to_token[i] = tok
        for tok, i in added.items():
            self.id_to_token[i] = tok
            self.is_added[i] = True
        self._bytes = _byte_decoder()
        self.eos_token_id = self.token_to_id[eos_token]
        self.pad_token_id = self.token_to_id[pad_token] if pad_token else self.eos_token_id
        self.prompt_ids = [self.token_to_id[t] for t in self.PROMPT_TOKENS]

    @classmethod
    def from_pretrained(cls, model_path):
        """Read `<model_path>/tokenizer.json` (HF `tokenizers` format), or the slow-tokenizer
        trio vocab.json + added_tokens.json (the format the published checkpoints ship)."""
        tj = os.path.join(model_path, "tokenizer.json")
        pad = None
        cfg_path = os.path.join(model_path, "tokenizer_config.json")
        if os.path.isfile(cfg_path):
            cfg = json.load(open(cfg_path))
            pad = cfg.get("pad_token")
            if isinstance(pad, dict):
                pad = pad.get("content")
        if os.path.isfile(tj):
            data = json.load(open(tj))
            pieces = dict(data["model"]["vocab"])
            added = {a["content"]: a["id"] for a in data.get("added_tokens", [])}
        else:
            pieces = json.load(open(os.path.join(model_path, "vocab.json")))
            added = {}
            ap = os.path.join(model_path, "added_tokens.json")
            if os.path.isfile(ap):
                added = json.load(open(ap))
        for tok in added:
            pieces.pop(tok, None)
        return cls(pieces, added, pad_token=pad if pad in added or pad in pieces else None)

    def convert_tokens_to_ids(self, tokens):
        if isinstance(tokens, str):
            return self.token_to_id[tokens]
        return [self.token_to_id[t] for t in tokens]

    def decode(self, ids):
        out, run = [], bytearray()
        n = len(self.id_to_token)
        # generated rows end in a long run of one special token (EOS / pad up to max_length): emit it in one piece
        ids = list(ids)
        tail_text = ""
        if len(ids) > 1:
            last = int(ids[-1])
            if 0 <= last < n and self.is_added[last]:
                k = len(ids)
                while k > 0 and ids[k - 1] == ids[-1]:
                    k -= 1
                tail_text = self.id_to_token[last] * (len(ids) - k)
                ids = ids[:k]
        for i in ids:
            i = int(i)
            if i < 0 or i >= n:
                continue
            if self.is_added[i]:
                if run:
                    out.append(run.decode("utf-8", errors="replace"))
                    run = bytearray()
                out.append(self.id_to_token[i])
            else:
                for ch in self.id_to_token[i]:
                    b = self._bytes.get(ch)
                    if b is None:
                        run.extend(ch.encode("utf-8"))
                    else:
                        run.append(b)
        if run:
            out.append(run.decode("utf-8", errors="replace"))
        out.append(tail_text)
        return "".join(out)

    def batch_decode(self, batch_ids, skip_special_tokens=False):
        """Same strings as HF `batch_decode(..., skip_special_tokens=False)` (reference model.py:668).  A rectangular
        int array (what generate() returns) has its trailing EOS / pad runs measured in one vectorised pass."""
        arr = batch_ids if isinstance(batch_ids, np.ndarray) else None
        if arr is None or arr.ndim != 2 or arr.shape[1] < 2 or arr.shape[0] == 0:
            return [self.decode(row) for row in batch_ids]
        n = len(self.id_to_token)
        differs = arr[:, ::-1] != arr[:, -1:]
        tail = np.where(differs.any(axis=1), differs.argmax(axis=1), arr.shape[1])
        out = []
        for row, t in zip(arr, tail.tolist()):
            last = int(row[-1])
            if 0 <= last < n and self.is_added[last]:
                out.append(self.decode(row[:arr.shape[1] - t].tolist()) + self.id_to_token[last] * t)
            else:
                out.append(self.decode(row.tolist()))
        return out
