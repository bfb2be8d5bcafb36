"""TEST INFRASTRUCTURE ONLY -- numpy restatement of HuggingFace beam search (the reference's default
decode mode, `num_beams=4`: reference model.py:409, 614, 662).

The algorithm lives in the third-party `transformers` package (pinned 4.38.2 in the reference's
requirements.txt:1; 5.5.0 installed here), reached from reference model.py:609/655 through
`model.generate(num_beams=..., length_penalty=...)`.  Restated from the published algorithm
(transformers/generation/utils.py `_beam_search` and its helpers `_get_top_k_continuations`,
`_get_running_beams_for_next_iteration`, `_update_finished_beams`, `_check_early_stop_heuristic`):

  per item: `nb` running beams with summed log-probs (start [0, -1e9, ...]) and a pool of `nb` finished
  hypotheses (scores start at -1e9).  Each step
    1. log_softmax over the FULL vocabulary in fp32, THEN the logits processors (suppressed ids -> -inf);
    2. add the running score, take the top K = 2*nb of the nb*V continuations (descending);
    3. a continuation "hits" when its token is EOS or the sequence reaches max_length;
    4. next running beams = first nb continuations after pushing the hits down by -1e9;
    5. hits among the TOP nb continuations enter the finished pool with score
       sum_logprob / (generated_len ** length_penalty) (generated_len counts the EOS); the pool keeps its
       best nb; nothing is added once the item's stop heuristic has fired;
    6. heuristic (early_stopping=False): stop the item when every pool slot is filled and
       best_running_sum / (generated_len ** length_penalty) <= worst pool score;
  the loop ends when every item has stopped or max_length is reached; pool slot 0 is returned, padded.

Pinned by tests/test_oracle_model.py::test_beam_oracle_matches_hf_generate against HF `generate` itself.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np

NEG = np.float32(-1.0e9)


def log_softmax_f32(logits):
    x = np.asarray(logits, dtype=np.float32)
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m, dtype=np.float32)
    return (x - m) - np.log(e.sum(axis=-1, keepdims=True, dtype=np.float32), dtype=np.float32)


def topk_desc(values, k):
    """Indices of the k largest values, descending, lowest index first among equals."""
    order = np.argsort(-values, kind="stable")
    return order[:k]


class BeamState:
    def __init__(self, batch, num_beams, prompt, eos_id, pad_id, max_length, length_penalty=1.0):
        self.B, self.nb, self.K = batch, num_beams, 2 * num_beams
        self.eos, self.pad, self.max_length, self.lp = eos_id, pad_id, max_length, float(length_penalty)
        self.prompt_len = len(prompt)
        self.cur_len = self.prompt_len
        self.running_seq = np.full((batch, num_beams, max_length), pad_id, dtype=np.int64)
        self.running_seq[:, :, :self.prompt_len] = np.asarray(prompt, dtype=np.int64)
        self.running_score = np.full((batch, num_beams), NEG, dtype=np.float32)
        self.running_score[:, 0] = 0.0
        self.fin_seq = self.running_seq.copy()
        self.fin_score = np.full((batch, num_beams), NEG, dtype=np.float32)
        self.fin_len = np.zeros((batch, num_beams), dtype=np.int64)      # generated tokens incl. EOS
        self.is_fin = np.zeros((batch, num_beams), dtype=bool)
        self.unsat = np.ones(batch, dtype=bool)                          # stop heuristic not yet satisfied
        self.last_hits = np.zeros((batch, self.K), dtype=bool)
        self.finished = False

    def rows_tokens(self):
        """Last token of every running beam, [B*nb] -- the decoder input of the next step."""
        return self.running_seq[:, :, self.cur_len - 1].reshape(-1)

    def step(self, logits, suppress_ids=None, begin_suppress_ids=None):
        """logits f32 [B*nb, V] for the running beams.  Returns parent rows [B*nb] (flat indices into the
        rows that produced `logits`) for reordering the self-attention cache."""
        B, nb, K = self.B, self.nb, self.K
        lp = log_softmax_f32(logits)
        V = lp.shape[-1]
        if suppress_ids is not None and len(suppress_ids):
            lp[:, list(suppress_ids)] = -np.inf
        if begin_suppress_ids is not None and len(begin_suppress_ids) and self.cur_len == self.prompt_len:
            lp[:, list(begin_suppress_ids)] = -np.inf
        acc = (lp.reshape(B, nb, V) + self.running_score[:, :, None]).astype(np.float32).reshape(B, nb * V)
        gen_len = self.cur_len + 1 - self.prompt_len
        parents = np.zeros((B, nb), dtype=np.int64)
        for b in range(B):
            idx = topk_desc(acc[b], K)
            top_lp = acc[b, idx]
            beam, tok = idx // V, idx % V
            cand_seq = self.running_seq[b, beam].copy()
            cand_seq[:, self.cur_len] = tok
            hits = (tok == self.eos) | (self.cur_len + 1 >= self.max_length)
            # next running beams
            run_lp = (top_lp + hits.astype(np.float32) * NEG).astype(np.float32)
            nxt = topk_desc(run_lp, nb)
            # finished pool
            just = hits & (np.arange(K) < nb)
            sc = (top_lp / np.float32(gen_len ** self.lp)).astype(np.float32)
            if not self.unsat[b]:
                sc = sc + NEG
            sc = (sc + (~just).astype(np.float32) * NEG).astype(np.float32)
            m_score = np.concatenate([self.fin_score[b], sc])
            m_seq = np.concatenate([self.fin_seq[b], cand_seq], axis=0)
            m_fin = np.concatenate([self.is_fin[b], just])
            m_len = np.concatenate([self.fin_len[b], np.full(K, gen_len)])
            keep = topk_desc(m_score, nb)
            self.fin_score[b], self.fin_seq[b], self.is_fin[b], self.fin_len[b] = m_score[keep], m_seq[keep], m_fin[keep], m_len[keep]
            self.running_seq[b] = cand_seq[nxt]
            self.running_score[b] = run_lp[nxt]
            parents[b] = b * nb + beam[nxt]
            self.last_hits[b] = hits
        self.cur_len += 1
        # early-stop heuristic (early_stopping=False): best possible running score vs worst finished
        best = self.running_score[:, 0] / np.float32((self.cur_len - self.prompt_len) ** self.lp)
        worst = np.where(self.is_fin, self.fin_score.min(axis=1, keepdims=True), NEG)
        self.unsat = self.unsat & (best[:, None] > worst).any(axis=1)
        self.finished = not (self.unsat.any() and not self.last_hits.all())
        return parents.reshape(-1)

    def active_items(self):
        return self.unsat.copy()

    def result(self):
        """Best hypothesis per item with the prompt stripped, padded to the longest: int64 [B, n]."""
        n = int(self.fin_len[:, 0].max()) if self.B else 0
        out = self.fin_seq[:, 0, self.prompt_len:self.prompt_len + n].copy()
        for b in range(self.B):
            out[b, self.fin_len[b, 0]:] = self.pad
        return out


def beam_search(logits_fn, reorder_fn, batch, num_beams, prompt, eos_id, pad_id, max_length, length_penalty=1.0,
                suppress_ids=None, begin_suppress_ids=None):
    """Drive BeamState with `logits_fn(tokens [B*nb, T]) -> f32 [B*nb, V]` (last position) and
    `reorder_fn(parent_rows)` (self-attention cache reorder)."""
    st = BeamState(batch, num_beams, prompt, eos_id, pad_id, max_length, length_penalty)
    tokens = np.tile(np.asarray(prompt, dtype=np.int64), (batch * num_beams, 1))
    while not st.finished:
        logits = logits_fn(tokens)
        parents = st.step(logits, suppress_ids, begin_suppress_ids)
        reorder_fn(parents)
        tokens = st.rows_tokens().reshape(-1, 1)
    return st.result(), st
