"""One VACNIC training step (TRAIN:253-374) on the B200-native model: forward, token CE, CoLaM margin
loss against the frozen stock-BART guide, SECLA face-name loss, backward, (data-parallel gradient
all-reduce), fused AdamW with the linear warm-up/decay schedule — optionally captured once as a CUDA
graph and replayed, so that the ~2.5k kernel launches of a step cost no host time.

Host-side work per step is: copy the batch into static device buffers and replay.  The step counter, the
learning-rate schedule and the Adam bias corrections live ON THE DEVICE (`vacnic_optim_schedule`, first node of
the step): the host may run any number of replays ahead without a synchronisation.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import torch

from . import blocks as Bk
from . import kernels as K
from .dp import GradBuckets, PeerShards, vacnic_bucket_prefixes
from .modeling import VacnicBart, shift_tokens_right
from .varlen import ArticlePack, pack_articles


SIDE_STREAM_DEFAULT = "1"
GUIDE_STREAM_DEFAULT = "1"


def linear_schedule(step: int, warmup: int, total: int) -> float:
    """transformers.get_linear_schedule_with_warmup (TRAIN:102): multiplier applied to the base lr."""
    if step < warmup:
        return step / max(1, warmup)
    return max(0.0, (total - step) / max(1, total - warmup))


class TrainStep:
    def __init__(self, model: VacnicBart, guide: Optional[VacnicBart], lr: float = 3e-5, weight_decay: float = 0.01,
                 betas=(0.9, 0.999), eps: float = 1e-8, warmup_steps: int = 0, total_steps: int = 1_000_000,
                 margin: float = 1.0, alpha: float = 0.5, secla_weight: float = 1.0, use_graph: bool = True,
                 process_group=None, pipeline_optimizer: Optional[bool] = None, max_grad_norm: Optional[float] = None,
                 exchange: Optional[str] = None, varlen: bool = False, side_stream: Optional[bool] = None,
                 guide_stream: Optional[bool] = None):
        """`exchange` (world > 1): "p2p" = rank-sharded optimizer over NVLink peer memory (one fused reduce-scatter + AdamW +
        all-gather kernel per bucket, csrc/dp.cu; needs the model's store in symmetric memory), "p2p-mc" = the same through
        the NVSwitch multicast object (multimem.ld_reduce / multimem.st), "nccl" = in-place NCCL all-reduce of the fp32
        gradient buckets followed by the replicated fused AdamW.  None = "p2p" when available, else "nccl"."""
        self.model, self.guide = model, guide
        self.cfg = model.cfg
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.warmup, self.total = warmup_steps, total_steps
        self.margin, self.alpha, self.w_secla = margin, alpha, secla_weight
        self.use_graph = use_graph
        # torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm) of TRAIN:365-366 (--no_clip_norm True in the shipped
        # scripts): the global norm needs every gradient, so it switches the per-bucket optimizer pipelining off
        self.max_grad_norm = max_grad_norm
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        st = model.store
        dev = st.device
        self.m = torch.zeros_like(st.master)
        self.v = torch.zeros_like(st.master)
        self.hyper = torch.zeros(8, dtype=torch.float32, device=dev)
        self._clip_scratch = torch.zeros(2368, dtype=torch.float32, device=dev)  # VACNIC_CLIP_SCRATCH_FLOATS
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)  # optimizer updates applied so far (device truth)
        self.step_no = 0                                               # host mirror (bookkeeping only)
        self.graph = None                                  # graph / static inputs / losses of the bucket used last
        self.static: Dict[str, torch.Tensor] = {}
        self.losses: Dict[str, torch.Tensor] = {}
        self._graphs: Dict[int, tuple] = {}                # shape bucket -> (graph, static inputs, loss tensors)
        self._pool = None
        self.launches_per_step = 0
        # packed (varlen) article rows: the collate's padding never reaches the device (vacnic_b200.varlen)
        self.varlen = bool(varlen)
        self.buckets: Optional[GradBuckets] = None
        self.p2p: Optional[PeerShards] = None
        if exchange is None:
            exchange = os.environ.get("VACNIC_DP_EXCHANGE") or ("p2p" if (self.world > 1 and st.symmetric and max_grad_norm is None)
                                                                else "nccl")
        if exchange not in ("p2p", "p2p-mc", "nccl"):
            raise ValueError(f"exchange must be 'p2p', 'p2p-mc' or 'nccl', got {exchange!r}")
        if exchange != "nccl":
            if self.world == 1:
                exchange = "nccl"
            elif max_grad_norm is not None:
                raise ValueError("clip_grad_norm_ needs the full reduced gradient: use exchange='nccl' with max_grad_norm")
            else:
                self.p2p = PeerShards(st, process_group, use_multicast=(exchange == "p2p-mc"))
                if exchange == "p2p-mc" and not self.p2p.grad_mc:
                    raise RuntimeError("exchange='p2p-mc': this system exposes no multicast (NVLS) mapping")
                pipeline_optimizer = True
                st.gather_master = self.gather_master
        self.exchange = exchange
        if max_grad_norm is not None:
            pipeline_optimizer = False
        if pipeline_optimizer is None:
            # measured on B200: per-bucket updates behind the backward pass gain 0.6 % at 2 GPUs (the fused update no
            # longer waits for the last all-reduce) and lose 1 % on one GPU (HBM contention with the backward GEMMs)
            pipeline_optimizer = self.world > 1
        self.pipeline = bool(pipeline_optimizer)
        if pipeline_optimizer or self.world > 1:
            # Address-range buckets of the flat gradient buffer, completed from markers in the backward pass
            # (vacnic_b200.dp, blocks.GradMarkFn).  On a side stream each bucket is all-reduced in place (world > 1) and
            # then UPDATED by the fused AdamW kernel right away, while the backward pass of the earlier layers is still
            # running: a finished layer's weights are not read again in this step.
            spans = {n: (st.offsets[n], p.numel()) for n, p in st.params.items()}
            prefixes = vacnic_bucket_prefixes(self.cfg.enc_layers, self.cfg.dec_layers, group_size=3)
            self.buckets = GradBuckets(st.grad, spans, prefixes, group=process_group if (self.p2p is None and (
                self.world > 1 or os.environ.get("VACNIC_DP_FORCE"))) else None)
            self.comm_stream = torch.cuda.Stream(device=dev)
            self._tag_to_bucket = {("dec", 0): 0}
            hi, k = self.cfg.enc_layers, 1
            while hi > 0:
                lo = max(0, hi - 3)
                self._tag_to_bucket[("enc", lo)] = k
                hi, k = lo, k + 1
        # grid cap of the shard kernels that run BESIDE the backward pass (they only need enough loads in flight to fill
        # NVLink, not the whole GPU); the exposed tail bucket uses the full grid
        self.overlap_blocks = int(os.environ.get("VACNIC_DP_OVERLAP_BLOCKS", "64"))
        # Overlap of the exchange with the backward pass.  NCCL path: on (the fp32 all-reduce is ~10 ms of wire time).  Peer-
        # memory path: OFF by default -- measured on 2 x B200 (profiles/r2_dp_exchange_n2.md): the sharded update run entirely
        # after the backward pass costs no more than the single-GPU AdamW it replaces (533 samples/s = two independent
        # replicas), while shard kernels running beside the backward GEMMs slow those down by 5 ms / step (488 samples/s).
        # VACNIC_DP_SERIAL=1 / VACNIC_DP_OVERLAP=1 force either behaviour.
        env_serial, env_overlap = os.environ.get("VACNIC_DP_SERIAL"), os.environ.get("VACNIC_DP_OVERLAP")
        self.overlap = (self.p2p is None) if (env_serial is None and env_overlap is None) else (
            env_overlap == "1" if env_overlap is not None else env_serial != "1")
        # Prefix side of the encoder on a second (high-priority) stream, forward and backward (BartEncoder.forward).  Not with
        # exchange overlap: the backward-pass markers describe the main stream only.  VACNIC_SIDE_STREAM=0/1 overrides.
        self.side_stream = None
        want_side = (os.environ.get("VACNIC_SIDE_STREAM", SIDE_STREAM_DEFAULT) == "1") if side_stream is None else bool(side_stream)
        if want_side and not self.cfg.stock and not (self.buckets is not None and self.overlap):
            self.side_stream = torch.cuda.Stream(device=dev, priority=-1)
            K.register_side_stream(self.side_stream)
        # The frozen guide's forward (no gradient, independent of the model until the CoLaM loss) on a stream of its own,
        # enqueued BEFORE the model's forward: its latency-bound decoder kernels (1024 rows) then run beside the model's
        # encoder GEMMs and its encoder GEMMs beside the model's decoder.  VACNIC_GUIDE_STREAM=0/1 overrides.
        self.guide_stream = None
        want_g = (os.environ.get("VACNIC_GUIDE_STREAM", GUIDE_STREAM_DEFAULT) == "1") if guide_stream is None else bool(guide_stream)
        if want_g and guide is not None:
            self.guide_stream = torch.cuda.Stream(device=dev)
            K.register_side_stream(self.guide_stream)
        self._g_txt = torch.ones(1, device=dev)
        self._g_margin = torch.full((1,), alpha, device=dev)
        self._g_secla = torch.full((1,), secla_weight, device=dev)

    def _on_grad_ready(self, tag):
        """Backward-pass marker (runs inside the autograd engine): the gradients of bucket `tag` are final at this point
        of the compute stream.  Only an event is recorded here; the all-reduce that waits for it is enqueued on the
        communication stream by `_body` right after the (asynchronous) backward pass has been enqueued, so on the GPU
        timeline it still overlaps the remaining backward kernels — and no NCCL call is made from an engine thread."""
        i = self._tag_to_bucket.get(tag)
        if i is None or not self.overlap:
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._ready.append((i, ev))

    # ------------------------------------------------------------------ the step body (capturable)
    def _body(self, b: Dict[str, torch.Tensor], dry: bool = False):
        """`dry`: forward + backward only (graph warm-up): no schedule tick, no exchange, no optimizer, no collective."""
        model, guide, cfg = self.model, self.guide, self.cfg
        st = model.store
        if not dry:
            # t += 1; hyper = {lr_t, betas, eps, wd, bias corrections, 1/world} -- a device kernel inside the step
            K.optim_schedule(self.step_dev, self.hyper, self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.warmup,
                             self.total, 1.0 / self.world)
        st.begin_step()
        # the backward-pass markers call into THIS step object (another TrainStep may share the model)
        model.rt.grad_hook = self._on_grad_ready if self.buckets is not None else None
        if self.buckets is not None:
            self.buckets.begin_step()
            self._ready = []
        model.rt.rng.advance()
        model.rt.side_stream = self.side_stream
        try:
            return self._fwd_bwd_update(b, dry)
        finally:
            model.rt.side_stream = None
            model.rt.keepalive.clear()

    def _fwd_bwd_update(self, b: Dict[str, torch.Tensor], dry: bool):
        model, guide, cfg = self.model, self.guide, self.cfg
        st = model.store
        src, tgt = b["article_ids"], b["caption_ids"]
        dec_in = b["decoder_input_ids"]
        pack = None
        if "pk_ids" in b:  # packed article rows (varlen.pack_articles on the host batch)
            B_, L_ = src.shape
            prefix = 0 if cfg.stock else cfg.prompt_size + (0 if cfg.only_image else cfg.max_ner_type_len_gt)
            pack = ArticlePack({k[3:]: b[k] for k in ("pk_ids", "pk_pos", "pk_start", "pk_len", "pk_qlen")}, B_, L_, prefix,
                               dec_in.shape[1])
        kw = dict(input_ids=src, attention_mask=b["src_mask"], decoder_input_ids=dec_in,
                  image_features=b["image_features"], ce_targets=tgt, article_pack=pack)
        if not cfg.only_image:
            kw.update(face_features=b["face_emb"], face_mask=b["face_mask"], name_ids=b["names_art_ids"],
                      name_mask=b["name_mask"])
        gout, gs = None, self.guide_stream
        gkw = dict(input_ids=src, attention_mask=b["src_mask"], decoder_input_ids=dec_in, article_pack=pack,
                   need_logits=False)  # only decoder_hidden_states[-1] is read (TRAIN:293-296)
        if guide is not None and gs is not None:
            gs.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(gs), torch.no_grad():
                gout = guide(**gkw)
        out = model(**kw)
        heads, grads = [out["loss"]], [self._g_txt]
        losses = {"txt": out["loss"]}
        if guide is not None:
            if gout is None:
                with torch.no_grad():
                    gout = guide(**gkw)
            else:  # join; `gout` (allocated on the guide's stream, read on this one) stays referenced until the step returns
                torch.cuda.current_stream().wait_stream(gs)
                model.rt.keepalive.append(gout["decoder_hidden_states"][-1])
            margin = Bk.ColamFn.apply(out["decoder_hidden_states"][-1], gout["decoder_hidden_states"][-1], tgt, self.margin,
                                      cfg.pad_token_id)
            heads.append(margin); grads.append(self._g_margin)
            losses["margin"] = margin
        if not cfg.only_image:
            enc = model.model.encoder
            names = K.names_embed(b["names_ids"], st.w16(enc.embed_tokens_ner.weight), st.w16(enc.embed_positions_ner.weight),
                                  enc.ln_emb_ner.g, enc.ln_emb_ner.b)
            secla = Bk.SeclaFn.apply(out["hidden_states_face"], names)
            heads.append(secla); grads.append(self._g_secla)
            losses["secla"] = secla
        torch.autograd.backward(heads, grads)
        if self.side_stream is not None:  # join: the prefix side's backward nodes ran on the side stream
            torch.cuda.current_stream().wait_stream(self.side_stream)
        st.finish_backward()
        if dry:
            return losses
        if self.p2p is not None:
            # rank-sharded optimizer over NVLink peer memory.  Per bucket, on the communication stream: wait for the
            # backward-pass marker, barrier (every rank's gradients of the bucket are final and nobody reads its weights
            # again in this step), then ONE kernel per address range sums this rank's shard of all ranks' gradients,
            # updates its master / moments and stores the new bf16 weights into every rank's shadow.
            nb = len(self.buckets.buckets)
            for i, ev in self._ready:
                if not self.buckets.done[i]:
                    self.buckets.done[i] = True
                    self.comm_stream.wait_event(ev)
                    with torch.cuda.stream(self.comm_stream):
                        self._p2p_update(self.buckets.buckets[i], channel=i, max_blocks=self.overlap_blocks)
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                rest = [r for i, d in enumerate(self.buckets.done) if not d for r in self.buckets.buckets[i]] + list(self.buckets.rest)
                self.buckets.done = [True] * nb
                self._p2p_update(rest, channel=nb, max_blocks=0)       # exposed tail: the whole GPU
                self.p2p.barrier(nb + 1)   # all ranks finished reading my gradients and writing my shadow: the step is over
            torch.cuda.current_stream().wait_stream(self.comm_stream)
            st.master_sharded = True
        elif self.buckets is not None:  # sum over ranks; the 1/world mean is folded into hyper[7]
            for i, ev in self._ready:  # bucket i only depends on the point of the backward pass where it became final
                if not self.buckets.done[i]:
                    self.comm_stream.wait_event(ev)
                    with torch.cuda.stream(self.comm_stream):
                        r = self.buckets.reduce_bucket(i)
                        if self.pipeline:
                            self._adamw_ranges(r)
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                # embeddings / prefix modules + any bucket whose marker did not fire
                r = self.buckets.finish()
                if self.pipeline:
                    self._adamw_ranges(r)
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        if self.buckets is None or not self.pipeline:
            if self.max_grad_norm is not None:
                K.clip_grad_scale(st.grad, self.max_grad_norm, 1.0 / self.world, self._clip_scratch, self.hyper[7:8], self.grad_norm)
            K.adamw(st.master, st.grad, self.m, self.v, st.shadow, self.hyper)
        return losses

    def _p2p_update(self, ranges, channel: int, max_blocks: int):
        st, pz = self.model.store, self.p2p
        pz.barrier(channel)
        for a, b in ranges:
            begin, count = pz.shard(a, b)
            if count > 0:
                K.dp_adamw_shard(pz.grad_ptrs, pz.shadow_ptrs, pz.grad_mc, pz.shadow_mc, pz.world, pz.rank, st.master, self.m,
                                 self.v, self.hyper, begin, count, max_blocks)

    @torch.no_grad()
    def gather_master(self):
        """Rank-sharded optimizer: bring the fp32 master up to date on every rank (each shard is broadcast by its owner).
        COLLECTIVE; used before checkpoints / state_dict() / any re-cast of the bf16 shadow, never inside the step."""
        st = self.model.store
        if self.p2p is None or not st.master_sharded:
            return
        torch.cuda.synchronize()
        ranges = [r for bk in self.buckets.buckets for r in bk] + list(self.buckets.rest)
        from .dp import shard_of_range
        for a, b in ranges:
            for r in range(self.world):
                begin, count = shard_of_range(a, b, r, self.world)
                if count > 0:
                    torch.distributed.broadcast(st.master[begin:begin + count], src=torch.distributed.get_global_rank(
                        self.pg, r) if self.pg is not None else r, group=self.pg)
        st.master_sharded = False
        st._shadow_version = st.master._version  # the shadow is already the bf16 image of these weights

    def _adamw_ranges(self, ranges):
        st = self.model.store
        for a, b in ranges:
            K.adamw(st.master[a:b], st.grad[a:b], self.m[a:b], self.v[a:b], st.shadow[a:b], self.hyper)

    # ------------------------------------------------------------------ host side
    @staticmethod
    def prepare(batch: Dict[str, torch.Tensor], cfg, varlen: bool = False) -> Dict[str, torch.Tensor]:
        """Host-side derived inputs of the training loop (TRAIN:267-271): decoder inputs and pad masks; with `varlen` also
        the packed article rows (varlen.pack_articles: the device-side replacement of the collate's padding)."""
        b = dict(batch)
        if varlen:
            b.update({"pk_" + k: v for k, v in pack_articles(batch["article_ids"].cpu(), cfg.pad_token_id).items()})
        b["decoder_input_ids"] = shift_tokens_right(batch["caption_ids"], cfg.pad_token_id, cfg.eos_token_id)
        b["src_mask"] = (batch["article_ids"] != 1).to(torch.int64)
        if "face_emb" in batch:
            b["face_mask"] = (batch["face_emb"][:, :, -1] != 1).to(torch.int64)
            b["name_mask"] = (batch["names_art_ids"] != 1).to(torch.int64)
        return b

    def current_lr(self) -> float:
        """Learning rate the NEXT update will use (host-side restatement of the device schedule, for logging)."""
        return self.lr * linear_schedule(self.step_no, self.warmup, self.total)

    def set_step(self, t: int):
        """Resume at optimizer update count `t` (checkpoint restore)."""
        self.step_no = int(t)
        self.step_dev.fill_(int(t))

    def _load_static(self, static: Dict[str, torch.Tensor], b: Dict[str, torch.Tensor]):
        for k, v in b.items():
            if k not in static:
                static[k] = torch.empty(v.shape, dtype=v.dtype, device=self.model.store.device)
            if static[k].shape != v.shape:
                raise ValueError(f"batch field {k} changed shape {tuple(static[k].shape)} -> {tuple(v.shape)}: "
                                 "the captured step is shape-specialised (only the packed article rows may vary, by bucket)")
            static[k].copy_(v, non_blocking=True)

    @staticmethod
    def _bucket_key(b: Dict[str, torch.Tensor]) -> int:
        """Captured graphs are specialised on tensor shapes; with packed (varlen) articles the only shape that varies from
        batch to batch is the packed row count, a multiple of varlen.ROW_BUCKET."""
        return int(b["pk_ids"].numel()) if "pk_ids" in b else 0

    def _capture(self, key: int, b: Dict[str, torch.Tensor]):
        """Warm-up (allocator, lazy kernel attribute setup) + capture of the step for one shape bucket.  The warm-up runs
        forward + backward only -- no schedule tick, no gradient exchange, no optimizer -- so it changes no state and issues
        no collective: under data parallelism ranks meet different buckets at different steps, and a rank that has to
        capture must not run more collectives than the ranks that merely replay."""
        model = self.model
        static: Dict[str, torch.Tensor] = {}
        self._load_static(static, b)
        rng_saved = model.rt.rng.state.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._body(static, dry=True)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        model.rt.rng.state.copy_(rng_saved)
        if not self._graphs and self.world > 1:
            # FIRST capture (every rank is here at its first step): one eager collective of the largest bucket size, so
            # that NCCL / the symmetric-memory barrier have done their lazy allocations before any of them is captured
            with torch.cuda.stream(self.comm_stream):
                if self.p2p is not None:
                    self.p2p.barrier(0)
                elif self.buckets is not None and self.buckets.group is not None:
                    n = max(b_ - a_ for rs in self.buckets.buckets + [self.buckets.rest] for a_, b_ in rs)
                    torch.distributed.all_reduce(torch.zeros(n, dtype=torch.float32, device=model.store.device), group=self.pg)
            torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        c0 = K._l.launch_count()
        with torch.cuda.graph(graph, pool=self._pool):  # every bucket's graph shares one memory pool (they never overlap)
            losses = self._body(static)
        if self._pool is None:
            self._pool = graph.pool()
        self.launches_per_step = K._l.launch_count() - c0
        self._graphs[key] = (graph, static, losses)
        return self._graphs[key]

    def step(self, batch: Dict[str, torch.Tensor], prepared: bool = False) -> Dict[str, torch.Tensor]:
        """Run one optimisation step on a (host or device) batch; returns device scalars {txt, margin, secla}."""
        b = batch if prepared else self.prepare(batch, self.cfg, varlen=self.varlen)
        model, guide = self.model, self.guide
        if not model.training:
            model.train()
        if guide is not None and guide.training:
            guide.eval()
        # bring the bf16 compute shadows up to date HERE (outside any capture): the fused AdamW keeps the model's shadow
        # current from then on and the guide never changes, so the captured step holds no fp32 -> bf16 re-cast
        for mm in (model, guide):
            if mm is not None and mm.store.dirty_shadow:
                mm.store.refresh_shadow()
        st = model.store
        st.external_step = True  # this object drives begin_step / finish_backward (reset below: the plain
        try:                     # forward / backward / optimizer.step loop must keep working on the same model)
            self.step_no += 1
            if not self.use_graph:
                dev = st.device
                b = {k: v.to(dev, non_blocking=True) for k, v in b.items()}
                c0 = K._l.launch_count()
                self.losses = self._body(b)
                self.launches_per_step = K._l.launch_count() - c0
                return self.losses
            key = self._bucket_key(b)
            entry = self._graphs.get(key)
            if entry is None:
                entry = self._capture(key, b)
            else:
                self._load_static(entry[1], b)
            self.graph, self.static, self.losses = entry
            self.graph.replay()
            return self.losses
        finally:
            st.external_step = False
            if self.p2p is not None:
                st.master_sharded = True  # from here on only this rank's shard of the fp32 master is current

    def close(self):
        """Drop the captured graph and break the model <-> step reference cycle (the backward-pass markers hold a bound
        method of this object), so the step's device memory is returned as soon as the caller drops its reference."""
        if self.p2p is not None:
            self.gather_master()  # COLLECTIVE: leave every rank with the complete fp32 master
            self.model.store.gather_master = None
        self.graph = None
        self._graphs.clear()
        self._pool = None
        self.static = {}
        self.losses = {}
        if getattr(self.model.rt, "grad_hook", None) is not None:
            self.model.rt.grad_hook = None
