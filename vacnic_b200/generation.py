"""Caption generation on the B200-native model: encoder once per caption, then a cached single-token
decoder step replayed as ONE CUDA graph per position, with greedy / beam search running on the device.

Replaces `model.generate(...)` as the VACNIC scripts call it (INFER:798, 867; TRAIN:513-520): greedy
(`num_beams=1`) and beam search with `length_penalty`, for the generation config a default `BartConfig`
yields (decoder_start=2, eos=2, pad=1, forced_eos_token_id=2, early_stopping=False).  The search
semantics are transformers 5.5 `_beam_search` / `_sample` (restated and pinned in oracle/generate.py).

Device design (see csrc/decode.cu): the encoder memory is projected once into per-layer cross K/V
buffers `[layer][captions*L][2d]` shared by the beams of a caption (the reference expands the encoder
output `num_beams` times and caches `num_beams` identical copies, MFULL:2066-2074); the self-attention
cache is never reordered (ancestry table instead of `_reorder_cache`); all kernels read the current
length from a device scalar so the captured step graph is position independent.  The host only replays
the graph and polls a stop flag every few steps.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import kernels as K


DECODE_LANES_DEFAULT = "1"


class _Lane:
    """Captions [c0, c1) of a Generator: row-slice views of its buffers, own search state, step graph and stream."""

    def __init__(self, g: "Generator", c0: int, c1: int, dev, multi: bool):
        nb = g.nb
        r0, r1 = c0 * nb, c1 * nb
        self.c0, self.c1, self.C, self.R = c0, c1, c1 - c0, r1 - r0
        for name in ("x", "x2", "attn", "proj", "qb", "x32", "x32b", "qkv", "h", "logits", "top_lp", "top_idx"):
            setattr(self, name, getattr(g, name)[r0:r1])
        self.kc, self.vc = g.kc[:, r0:r1], g.vc[:, r0:r1]            # [layer][rows][maxT][d]: kc[l] is contiguous
        self.cross_kv = g.cross_kv[:, c0:c1]                          # [layer][captions][k|v][head][L][64]
        self.key_mask, self.key_len = g.key_mask[c0:c1], g.key_len[c0:c1]
        bf, f32, i32, u8 = torch.bfloat16, torch.float32, torch.int32, torch.uint8
        C, maxT = self.C, g.maxT

        def buf(*shape, dtype=bf):
            return torch.empty(*shape, dtype=dtype, device=dev)

        st: Dict[str, torch.Tensor] = {"cur_len": buf(1, dtype=i32), "flags": buf(maxT + 1, 2, dtype=i32)}
        if nb > 1:
            st.update(run_seq=buf(2, C, nb, maxT, dtype=i32), run_anc=buf(2, C, nb, maxT, dtype=i32),
                      run_score=buf(2, C, nb, dtype=f32), fin_seq=buf(2, C, nb, maxT, dtype=i32),
                      fin_score=buf(2, C, nb, dtype=f32), fin_len=buf(2, C, nb, dtype=i32),
                      fin_flag=buf(2, C, nb, dtype=u8), unsat=buf(C, dtype=u8))
        else:
            st.update(seq=buf(C, maxT, dtype=i32), unfinished=buf(C, dtype=u8))
        self.st = st
        self.flags_host = torch.zeros(g.max_len + 1, 2, dtype=torch.int32).pin_memory()
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.stream = torch.cuda.Stream(device=dev) if multi else None
        self.t, self.stop_t, self.pending, self.done = 1, None, None, False


class Generator:
    """Decode engine specialised for (captions, beams, article length, max_length) on one model."""

    def __init__(self, model, captions: int, beams: int, L: int, max_length: int, length_penalty: float = 1.0,
                 use_graph: bool = True, poll_every: int = 4, varlen: bool = True, lanes: Optional[int] = None):
        cfg = model.cfg
        if max_length < 2 or max_length > 256:
            raise ValueError("max_length must be in 2..256")
        if beams < 1 or beams > 8:
            raise ValueError("num_beams must be in 1..8")
        if max_length - 1 + 2 > cfg.max_pos + 2:
            raise ValueError("max_length exceeds the decoder's position table")
        self.model, self.cfg = model, cfg
        self.C, self.nb, self.L, self.max_len = captions, beams, L, max_length
        self.lp = float(length_penalty)
        self.R = captions * beams
        self.maxT = max_length
        self.use_graph, self.poll_every = use_graph, max(1, poll_every)
        # packed (varlen) article rows through the encoder: the padding of the batch never enters a GEMM / LayerNorm /
        # attention tile (vacnic_b200.varlen); the memory is un-packed once for the per-caption cross K/V projection
        self.varlen = bool(varlen)
        dev = model.store.device
        d, f, V = cfg.d_model, cfg.ffn, cfg.vocab
        R, nl = self.R, cfg.dec_layers
        bf, f32, i32, u8 = torch.bfloat16, torch.float32, torch.int32, torch.uint8

        def buf(*shape, dtype=bf):
            return torch.empty(*shape, dtype=dtype, device=dev)

        self.x, self.x2, self.attn, self.proj, self.qb = (buf(R, d) for _ in range(5))
        self.x32, self.x32b = buf(R, d, dtype=f32), buf(R, d, dtype=f32)  # fp32 residual stream, like the training forward
        self.qkv, self.h = buf(R, 3 * d), buf(R, f)
        self.ldp = (V + 7) // 8 * 8
        self.logits = buf(R, self.ldp, dtype=f32)
        self.kc = buf(nl, R, self.maxT, d)
        self.vc = buf(nl, R, self.maxT, d)
        # cross K/V, head-major: [layer][caption][k|v][head][L][64] -> each (caption, head) streams one contiguous block
        self.cross_kv = buf(nl, captions, 2, cfg.heads, L, 64)
        self.key_mask = buf(captions, L, dtype=u8)
        self.key_len = buf(captions, dtype=i32)
        self.Kc = 2 * beams if beams > 1 else 1
        self.top_lp = buf(R, self.Kc, dtype=f32)
        self.top_idx = buf(R, self.Kc, dtype=i32)
        # Decode lanes.  A decode step is half HBM-bound (cross-attention streams every caption's K/V: 0.83 of the HBM
        # peak) and half latency-bound (73 small GEMMs, LayerNorms, top-k: a few dozen CTAs each).  Captions are independent,
        # so the batch is cut into `lanes` groups with their own search state, step graph and stream: one lane's small
        # kernels run under another lane's K/V streaming.  A lane is a set of row-slice VIEWS of the buffers above.
        if lanes is None:
            lanes = int(os.environ.get("VACNIC_DECODE_LANES", DECODE_LANES_DEFAULT))
        if not use_graph or captions % max(1, lanes) != 0 or captions // max(1, lanes) < 16:
            lanes = 1
        self.n_lanes = lanes
        self.lanes = []
        cl = captions // lanes
        for i in range(lanes):
            self.lanes.append(_Lane(self, i * cl, (i + 1) * cl, dev, multi=lanes > 1))
        self.launches_per_step = 0
        self.steps_run = 0
        self.sequences_scores = self.sequences_len = None   # beam search only; set by decode()

    # single-lane views kept for the kernel tests / profiling tools that drive the engine step by step
    @property
    def st(self):
        return self.lanes[0].st

    @property
    def graph(self):
        return self.lanes[0].graph

    # ------------------------------------------------------------------ encoder side
    def _reset_state(self):
        cfg = self.cfg
        for lane in self.lanes:
            st = lane.st
            st["cur_len"].fill_(1)
            st["flags"].zero_()
            if self.nb > 1:
                st["run_seq"].fill_(cfg.pad_token_id)
                st["run_seq"][:, :, :, 0] = cfg.decoder_start_token_id
                st["fin_seq"].copy_(st["run_seq"])
                st["run_anc"].zero_()
                st["run_score"].fill_(-1.0e9)
                st["run_score"][:, :, 0] = 0.0
                st["fin_score"].fill_(-1.0e9)
                st["fin_len"].zero_()
                st["fin_flag"].zero_()
                st["unsat"].fill_(1)
            else:
                st["seq"].fill_(cfg.pad_token_id)
                st["seq"][:, 0] = cfg.decoder_start_token_id
                st["unfinished"].fill_(1)

    @torch.no_grad()
    def encode(self, enc_inputs: dict):
        """Encoder forward (a7a) + the one-off cross-attention K/V projection of every decoder layer."""
        model, cfg = self.model, self.cfg
        if model.store.dirty_shadow:
            model.store.refresh_shadow()
        encoder = model.model.encoder
        was_training = encoder.training
        encoder.training = False  # dropout off for this call only; no train()/eval() round trip (that would re-cast the shadow)
        pack = self._pack(enc_inputs) if self.varlen else None
        try:
            enc = encoder(output_hidden_states=False, pack=pack, **enc_inputs)
        finally:
            encoder.training = was_training
        h = enc["last_hidden_state"]
        if pack is not None:  # packed [1, rows, d] -> [C, L, d] (pad rows zero: never attended, key_len below)
            h = K.unpack_rows(h[0], pack.start, pack.len, self.L)
            enc["last_hidden_state"] = h
        C, L, d = h.shape
        if (C, L) != (self.C, self.L):
            raise ValueError(f"generator built for {self.C} captions x {self.L} tokens, got {C} x {L}")
        mask = enc_inputs.get("attention_mask")
        if mask is None:
            self.key_mask.fill_(1)
        else:
            self.key_mask.copy_(mask.to(torch.uint8))
        self.key_len.copy_(K.mask_key_len(self.key_mask))
        lin = model.model.decoder.lin_cross_kv  # rows [l*2d, (l+1)*2d) = [k_l ; v_l]
        for l in range(cfg.dec_layers):  # batch dim = caption; 64-column groups (heads) scatter to their own [L][64] blocks
            K.gemm(h, lin.w16[l * 2 * d:(l + 1) * 2 * d].unsqueeze(0).expand(C, 2 * d, d), out=self.cross_kv[l],
                   bias=lin.b32[l * 2 * d:(l + 1) * 2 * d], head_major=(64, 2 * d * L, 0, L * 64))
        return enc

    def _pack(self, enc_inputs: dict):
        """Device-side batch assembly for the encoder (the collate's padding, DNYT:957-972, stays out of the compute):
        article tokens packed back to back, rows rounded up to varlen.ROW_BUCKET.  One small device -> host read (the
        lengths) sizes the buffers; everything else is index arithmetic on the device."""
        from .varlen import ROW_BUCKET, ArticlePack
        ids, mask = enc_inputs["input_ids"], enc_inputs.get("attention_mask")
        C, L = ids.shape
        if mask is None:
            return None
        mb = mask.bool()
        lens = mb.sum(1)
        ar = torch.arange(L, device=ids.device)
        if not bool((mb == (ar[None, :] < lens[:, None])).all()) or int(lens.min()) < 1:
            return None  # not the collate's right padding (or an empty article): keep the padded path
        M = int(lens.sum())
        M_pad = (M + ROW_BUCKET - 1) // ROW_BUCKET * ROW_BUCKET
        pid = torch.full((M_pad,), self.cfg.pad_token_id, dtype=torch.int64, device=ids.device)
        pos = torch.zeros(M_pad, dtype=torch.int32, device=ids.device)
        pid[:M] = ids[mb]
        pos[:M] = ar.to(torch.int32)[None, :].expand(C, L)[mb]
        ln = lens.to(torch.int32)
        start = (torch.cumsum(lens, 0) - lens).to(torch.int32)
        qlen = ln.clone()
        qlen[-1] += M_pad - M
        cfg = self.cfg
        prefix = 0 if cfg.stock else cfg.prompt_size + (0 if cfg.only_image else cfg.max_ner_type_len_gt)
        return ArticlePack({"ids": pid, "pos": pos, "start": start, "len": ln, "qlen": qlen}, C, L, prefix, 1)

    # ------------------------------------------------------------------ one decoding step (capturable)
    def _step(self, lane=None):
        """One decoding step of `lane` (default: every lane in turn) on the current stream."""
        if lane is None:
            for ln in self.lanes:
                self._step(ln)
            return
        model, cfg, st = self.model, self.cfg, lane.st
        dec = model.model.decoder
        store = model.store
        d, H, V = cfg.d_model, cfg.heads, cfg.vocab
        beam = self.nb > 1
        x, x2 = lane.x, lane.x2
        r32, r32b = (lane.x32, lane.x32b) if model.rt.res_fp32 else (None, None)
        K.decode_embed_ln(st["run_seq"] if beam else st["seq"], st["cur_len"], store.w16(dec.embed_tokens.weight),
                          store.w16(dec.embed_positions.weight), dec.ln_emb.g, dec.ln_emb.b, x, self.maxT, pos_offset=2,
                          pingpong=beam, y32=r32)
        anc = st["run_anc"] if beam else None
        for l, layer in enumerate(dec.layers):
            a = layer.self_attn
            K.gemm(x, a.lin_qkv.w16, out=lane.qkv, bias=a.lin_qkv.b32)
            K.decode_self_attn(lane.qkv, lane.kc[l], lane.vc[l], anc, st["cur_len"], lane.attn, H, self.maxT)
            K.gemm(lane.attn, a.lin_o.w16, out=lane.proj, bias=a.lin_o.b32)
            K.add_layernorm_fwd(lane.proj, x, a.ln.g, a.ln.b, out=x2, want_stats=False, res32=r32, y32_out=r32b)
            a = layer.encoder_attn
            K.gemm(x2, a.lin_q.w16, out=lane.qb, bias=a.lin_q.b32)
            K.decode_cross_attn(lane.qb, lane.cross_kv[l][:, 0], lane.cross_kv[l][:, 1], lane.key_mask, lane.key_len, lane.attn,
                                self.nb)
            K.gemm(lane.attn, a.lin_o.w16, out=lane.proj, bias=a.lin_o.b32)
            K.add_layernorm_fwd(lane.proj, x2, a.ln.g, a.ln.b, out=x, want_stats=False, res32=r32b, y32_out=r32)
            K.gemm(x, layer.lin_fc1.w16, out=lane.h, bias=layer.lin_fc1.b32, act=K.ACT_GELU)
            K.gemm(lane.h, layer.lin_fc2.w16, out=lane.proj, bias=layer.lin_fc2.b32)
            K.add_layernorm_fwd(lane.proj, x, layer.ln_final.g, layer.ln_final.b, out=x2, want_stats=False, res32=r32, y32_out=r32b)
            x, x2 = x2, x
            r32, r32b = r32b, r32
        K.gemm(x, model.lin_lm.w16, out=lane.logits[:, :V], bias=model.final_logits_bias.view(-1))
        K.decode_topk(lane.logits, V, self.Kc, lane.top_lp, lane.top_idx)
        if beam:
            K.beam_step(lane.top_lp, lane.top_idx, st, lane.C, self.nb, self.maxT, self.max_len, cfg.eos_token_id, V, self.lp)
        else:
            K.greedy_step(lane.top_idx, st["seq"], st["unfinished"], st["flags"], st["cur_len"], lane.R, self.maxT,
                          self.max_len, cfg.eos_token_id, cfg.pad_token_id)
        K.advance_len(st["cur_len"])

    def _capture(self):
        # one eager step first (lazy kernel attribute setup), on a side stream as CUDA graphs require;
        # the state is reset afterwards, so the warm-up leaves no trace
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for lane in self.lanes:  # one graph per lane: lanes are replayed on their own streams and drift freely
            lane.graph = torch.cuda.CUDAGraph()
            c0 = K._l.launch_count()
            with torch.cuda.graph(lane.graph):
                self._step(lane)
            self.launches_per_step = K._l.launch_count() - c0
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ search loop
    @torch.no_grad()
    def decode(self) -> torch.Tensor:
        """Run the search on the encoded captions; returns int64 ids like transformers' generate()."""
        if self.use_graph and self.lanes[0].graph is None:
            self._reset_state()
            self._capture()
        self._reset_state()
        # Search loop.  The host enqueues `poll_every` graph replays at a time and looks at the stop flags of the PREVIOUS
        # chunk (copied into pinned memory behind that chunk) while the GPU is already working on the current one, so the
        # device never idles on the stop test.  At most one extra chunk runs after every caption has finished; finished
        # hypotheses are frozen by the step kernels (`unsat` / `unfinished`), so the result does not depend on it.
        # With several lanes every lane has its own stream, flags and stop test; a lane that has stopped is not replayed
        # again, the loop ends when all have.
        cur = torch.cuda.current_stream()
        multi = self.n_lanes > 1
        for lane in self.lanes:
            lane.t, lane.stop_t, lane.pending, lane.done = 1, None, None, False
            if multi:
                lane.stream.wait_stream(cur)  # the encoder side (and the state reset above) ran on the caller's stream

        def check(lane, p):
            ev, t0, n = p
            ev.synchronize()  # device -> host read: the stop test of the search loop
            fl = lane.flags_host[t0:t0 + n]
            go = (fl[:, 0] != 0) & (fl[:, 1] != 0)
            return None if bool(go.all()) else t0 + int((~go).nonzero()[0])

        while not all(lane.done for lane in self.lanes):
            live = [lane for lane in self.lanes if not lane.done]
            n = min(self.poll_every, self.max_len - live[0].t)  # live lanes are in step: they all started at t = 1
            for _ in range(n):
                for lane in live:
                    if self.use_graph:
                        if multi:
                            with torch.cuda.stream(lane.stream):
                                lane.graph.replay()
                        else:
                            lane.graph.replay()
                    else:
                        c0 = K._l.launch_count()
                        self._step(lane)
                        self.launches_per_step = K._l.launch_count() - c0
            for lane in live:
                with torch.cuda.stream(lane.stream if multi else cur):
                    lane.flags_host[lane.t:lane.t + n].copy_(lane.st["flags"][lane.t:lane.t + n], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                if lane.pending is not None:
                    lane.stop_t = check(lane, lane.pending)
                lane.pending = (ev, lane.t, n)
                lane.t += n
                if lane.stop_t is not None or lane.t >= self.max_len:
                    lane.done = True
        for lane in self.lanes:
            if lane.stop_t is None and lane.pending is not None:
                lane.stop_t = check(lane, lane.pending)
            if lane.stop_t is None:
                lane.stop_t = self.max_len - 1
            if multi:
                cur.wait_stream(lane.stream)
        self.steps_run = max(lane.t for lane in self.lanes) - 1
        pad = self.cfg.pad_token_id
        outs = []
        if self.nb > 1:
            scores, lens = [], []
            for lane in self.lanes:
                st = lane.st
                ob = lane.t & 1  # buffer written by the last executed step
                gen_len = int(st["fin_len"][ob, :, 0].max().item())
                # transformers' `sequences_scores`: sum of token log-probs / generated_len ** length_penalty, best hypothesis
                scores.append(st["fin_score"][ob, :, 0].clone())
                lens.append(st["fin_len"][ob, :, 0].clone())
                outs.append(st["fin_seq"][ob, :, 0, :1 + gen_len].to(torch.int64))
            self.sequences_scores, self.sequences_len = torch.cat(scores), torch.cat(lens)
        else:
            outs = [lane.st["seq"][:, :lane.stop_t + 1].to(torch.int64) for lane in self.lanes]
        if len(outs) == 1:
            return outs[0]
        width = max(o.shape[1] for o in outs)  # hypotheses are padded after their EOS, exactly as within one lane
        return torch.cat([torch.nn.functional.pad(o, (0, width - o.shape[1]), value=pad) for o in outs], dim=0)

    def generate(self, enc_inputs: dict) -> torch.Tensor:
        self.encode(enc_inputs)
        return self.decode()


def _enc_inputs(model, input_ids, attention_mask, image_features, face_features, face_mask, name_ids, name_mask):
    kw = dict(input_ids=input_ids, attention_mask=attention_mask)
    if not model.cfg.stock:
        kw["image_features"] = image_features
        if not model.cfg.only_image:
            kw.update(face_features=face_features, face_mask=face_mask, name_ids=name_ids, name_mask=name_mask)
    return kw


@torch.no_grad()
def generate(model, input_ids=None, attention_mask=None, num_beams: int = 1, max_length: int = 20,
             length_penalty: float = 1.0, image_features=None, face_features=None, face_mask=None, name_ids=None,
             name_mask=None, use_graph: bool = True) -> torch.Tensor:
    """`model.generate(...)` of the reference scripts (INFER:798, 867).  Engines are cached on the model per
    (captions, beams, L, max_length, length_penalty) so repeated calls replay the captured step graph."""
    if input_ids is None:
        raise ValueError("generate() needs input_ids")
    if not input_ids.is_cuda:
        raise K._l.VacnicError("vacnic_b200 runs on CUDA tensors only (no CPU fallback)")
    C, L = input_ids.shape
    cache = model.__dict__.setdefault("_generators", {})
    key = (C, num_beams, L, max_length, float(length_penalty), use_graph)
    if key not in cache:
        cache[key] = Generator(model, C, num_beams, L, max_length, length_penalty, use_graph=use_graph)
    return cache[key].generate(_enc_inputs(model, input_ids, attention_mask, image_features, face_features, face_mask,
                                           name_ids, name_mask))
