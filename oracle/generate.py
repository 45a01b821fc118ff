"""Greedy and beam-search decoding restated from the algorithm the reference's `generate()` runs.
TEST INFRASTRUCTURE ONLY (see oracle/model.py header).

The reference does not contain a decoder loop: `model.generate(...)` (INFER:798, 867; TRAIN:513-520)
dispatches into the third-party `transformers` package (pinned 4.18.0 in vacnic.yml:187, not
vendored, not executable here; installed and executable: 5.5.0).  What is restated below is the
published algorithm of transformers 5.5.0 `GenerationMixin._beam_search`
(generation/utils.py:3076-3400 with helpers :2876-3075) and `_sample` (greedy branch), for the
generation config a default `BartConfig` yields: decoder_start=2, eos=2, pad=1,
forced_eos_token_id=2 -> processors [ForcedEOSTokenLogitsProcessor], criteria [MaxLength, Eos],
early_stopping=False, do_sample=False.  Parity is pinned by tests/golden/make_golden.py, which runs
the unmodified reference classes through the real `generate()` and stores the token ids.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import model as M


# Test-only robustness probe: when set to (amplitude, torch.Generator), every next-token logit gets uniform
# noise in [-amplitude, amplitude].  tests/golden/make_golden.py uses it to keep only inputs whose decoded ids
# do not depend on perturbations of the size of the bf16 logit tolerance ("margin-vetted" goldens).
LOGIT_NOISE = None


def _encode(sd, cfg, batch_inputs, enc=None):
    return enc if enc is not None else M.encoder_forward(sd, cfg, **batch_inputs)


def _step_logits(sd, cfg, ids, enc_out, enc_mask, past):
    """Decoder on the last token with a KV cache (mathematically identical to the uncached run the
    reference does under transformers 5.5)."""
    if past is None:
        dec = M.decoder_forward(sd, cfg, ids, enc_out, enc_mask, past=None, use_cache=True)
    else:
        dec = M.decoder_forward(sd, cfg, ids[:, -1:], enc_out, enc_mask, past=past, use_cache=True)
    logits = M.lm_logits(sd, dec["last_hidden_state"][:, -1, :]).float()
    if LOGIT_NOISE is not None:
        amp, gen = LOGIT_NOISE
        logits = logits + (torch.rand(logits.shape, generator=gen) * 2 - 1).to(logits.device) * amp
    return logits, dec["past_key_values"]


def _forced_eos(log_probs, cur_len, max_length, eos):
    # ForcedEOSTokenLogitsProcessor: at cur_len == max_length - 1 everything but eos becomes -inf, eos 0
    if cur_len == max_length - 1:
        log_probs = torch.full_like(log_probs, -float("inf"))
        log_probs[:, eos] = 0
    return log_probs


@torch.no_grad()
def greedy(sd, cfg, enc_inputs, max_length=50, enc=None):
    eos, pad = cfg["eos_token_id"], cfg["pad_token_id"]
    enc = _encode(sd, cfg, enc_inputs, enc)
    enc_out, enc_mask = enc["last_hidden_state"], enc_inputs["attention_mask"]
    B = enc_out.shape[0]
    ids = torch.full((B, 1), cfg["decoder_start_token_id"], dtype=torch.long, device=enc_out.device)
    unfinished = torch.ones(B, dtype=torch.long, device=enc_out.device)
    past = None
    while True:
        logits, past = _step_logits(sd, cfg, ids, enc_out, enc_mask, past)
        scores = _forced_eos(logits, ids.shape[1], max_length, eos)
        nxt = scores.argmax(dim=-1)
        nxt = nxt * unfinished + pad * (1 - unfinished)
        ids = torch.cat([ids, nxt[:, None]], dim=-1)
        unfinished = unfinished & (nxt != eos).long() & int(ids.shape[1] < max_length)
        if unfinished.max() == 0:
            break
    return ids


def _gather_beams(t, idx):
    while idx.dim() < t.dim():
        idx = idx.unsqueeze(-1)
    return torch.take_along_dim(t, idx, dim=1)


@torch.no_grad()
def beam_search(sd, cfg, enc_inputs, num_beams=4, max_length=50, length_penalty=2.0, enc=None):
    """transformers 5.5.0 `_beam_search` (vectorised), early_stopping=False."""
    enc = _encode(sd, cfg, enc_inputs, enc)
    enc_out = enc["last_hidden_state"].repeat_interleave(num_beams, dim=0)
    enc_mask = enc_inputs["attention_mask"].repeat_interleave(num_beams, dim=0)
    B = enc_out.shape[0] // num_beams
    state = {"past": None}

    def step_fn(flat_ids, reorder):
        # g. reorder the self-attention cache by the source beam of each surviving running beam
        if reorder is not None:
            state["past"] = [(l[0].index_select(0, reorder), l[1].index_select(0, reorder), l[2], l[3]) for l in state["past"]]
        logits, state["past"] = _step_logits(sd, cfg, flat_ids, enc_out, enc_mask, state["past"])
        return logits

    return beam_search_core(step_fn, B, cfg["vocab"], enc_out.device, num_beams, max_length, length_penalty,
                            cfg["eos_token_id"], cfg["pad_token_id"], cfg["decoder_start_token_id"])


@torch.no_grad()
def beam_search_core(step_fn, B, V, dev, num_beams, max_length, length_penalty, eos, pad, start):
    """The search loop proper.  `step_fn(flat_ids [B*beams, cur_len], reorder or None)` returns fp32 next-token
    logits [B*beams, V]; `reorder` is the beam permutation chosen by the previous iteration."""
    K = 2 * num_beams  # beams_to_keep = max(2, 1 + n_eos) * num_beams
    cur_len = prompt = 1
    running = torch.full((B, num_beams, max_length), pad, dtype=torch.long, device=dev)
    running[:, :, 0] = start
    sequences = running.clone()
    running_scores = torch.zeros(B, num_beams, device=dev)
    running_scores[:, 1:] = -1e9
    beam_scores = torch.full((B, num_beams), -1e9, device=dev)
    finished = torch.zeros(B, num_beams, dtype=torch.bool, device=dev)
    unsat = torch.ones(B, 1, dtype=torch.bool, device=dev)
    run_idx = torch.full((B, num_beams, max_length - 1), -1, dtype=torch.int32, device=dev)
    beam_idx_out = run_idx.clone()
    top_mask = torch.cat((torch.ones(num_beams, dtype=torch.bool), torch.zeros(K - num_beams, dtype=torch.bool))).to(dev)
    bidx = None
    while True:
        flat = running[:, :, :cur_len].reshape(B * num_beams, cur_len)
        logits = step_fn(flat, bidx)
        lp = _forced_eos(F.log_softmax(logits, dim=-1), cur_len, max_length, eos)
        lp = (lp.view(B, num_beams, V) + running_scores[:, :, None]).reshape(B, num_beams * V)
        # c. top-K continuations
        topk_lp, topk_i = torch.topk(lp, k=K)
        src_beam = topk_i // V
        topk_run_idx = _gather_beams(run_idx, src_beam)
        topk_seq = _gather_beams(running, src_beam)
        topk_seq[:, :, cur_len] = topk_i % V
        topk_run_idx[:, :, cur_len - prompt] = (src_beam + torch.arange(B, device=dev).view(-1, 1) * num_beams).to(torch.int32)
        # d. stopping criteria on the K candidates: max length or eos
        hits = (topk_seq[:, :, cur_len] == eos) | (cur_len + 1 >= max_length)
        # e. running beams for the next iteration
        run_lp = topk_lp + hits.float() * -1.0e9
        nxt = torch.topk(run_lp, k=num_beams)[1]
        running = _gather_beams(topk_seq, nxt)
        running_scores = _gather_beams(run_lp, nxt)
        run_idx = _gather_beams(topk_run_idx, nxt)
        # f. finished beams
        just = hits & top_mask[None, :]
        fin_lp = topk_lp / ((cur_len + 1 - prompt) ** length_penalty)
        fin_lp = fin_lp + (~unsat).float() * -1.0e9
        fin_lp = fin_lp + (~just) * -1.0e9
        m_seq = torch.cat((sequences, topk_seq), dim=1)
        m_sc = torch.cat((beam_scores, fin_lp), dim=1)
        m_idx = torch.cat((beam_idx_out, topk_run_idx), dim=1)
        m_fin = torch.cat((finished, just), dim=1)
        sel = torch.topk(m_sc, k=num_beams)[1]
        sequences, beam_scores = _gather_beams(m_seq, sel), _gather_beams(m_sc, sel)
        beam_idx_out, finished = _gather_beams(m_idx, sel), _gather_beams(m_fin, sel)
        bidx = run_idx[..., cur_len - prompt].reshape(-1).long()
        cur_len += 1
        best_possible = running_scores[:, :1] / ((cur_len - prompt) ** length_penalty)
        worst_fin = torch.where(finished, beam_scores.min(dim=1, keepdim=True)[0], torch.tensor(-1.0e9, device=dev))
        unsat = unsat & torch.any(best_possible > worst_fin, dim=-1, keepdim=True)
        if not (bool(unsat.any()) and not bool(hits.all())):
            break
    seq = sequences[:, 0, :]
    gen_len = int(((beam_idx_out[:, 0, :] + 1).bool()).sum(dim=1).max())
    return seq[:, : prompt + gen_len], beam_scores[:, 0]
