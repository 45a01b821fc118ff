"""Flat parameter storage: one fp32 master buffer, one bf16 compute shadow and one fp32 gradient buffer
for the whole model, with every nn.Parameter a view into them.

Why: (1) the tensor-core kernels consume bf16 weights while the reference trains fp32 parameters
(TRAIN:95, `bart-large-fp32`), (2) projections the reference runs as separate nn.Linear calls on the
same input (k/v/q of BartAttention MFULL:444-446; the twelve decoder cross-attention k/v projections
of the encoder memory, MFULL:856) become ONE GEMM when their weights are adjacent in memory, (3) the
optimizer and the gradient all-reduce become single passes over a flat buffer.  Parameter names and
shapes stay those of the reference state_dict.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import torch
from torch import nn

from . import kernels as K

ALIGN = 64  # elements; keeps every view 128-byte (bf16) / 256-byte (fp32) aligned


class Lin:
    """Views one (possibly fused) linear layer needs: bf16 weight, fp32 bias, fp32 gradient views."""
    __slots__ = ("w16", "b32", "gw", "gb", "key", "out_f", "in_f")

    def __init__(self, w16, b32, gw, gb, key):
        self.w16, self.b32, self.gw, self.gb, self.key = w16, b32, gw, gb, key
        self.out_f, self.in_f = w16.shape


class ParamStore:
    def __init__(self, module: nn.Module, device, first: Sequence[str] = (), frozen: bool = False, symmetric: bool = False):
        """`first`: parameter names to lay out first in their region, in the given order (used to make
        fused groups adjacent).  `frozen`: no gradient / optimizer buffers (the CoLaM guide).  `symmetric`: allocate
        the gradient buffer and the bf16 shadow as torch symmetric memory (cuMem allocations every rank of the job can
        map), so the data-parallel optimizer kernel reads / writes them across NVLink (vacnic_dp_adamw_shard)."""
        named = []
        seen = set()
        for n, p in module.named_parameters():
            if id(p) in seen:
                continue
            seen.add(id(p))
            named.append((n, p))
        by_name = dict(named)
        order = [n for n in first if n in by_name] + [n for n, _ in named if n not in set(first)]

        def is_matrix(n, p):
            return p.dim() == 2 and "embed_" not in n and "shared" not in n

        w_names = [n for n in order if is_matrix(n, by_name[n])]
        z_names = [n for n in order if not is_matrix(n, by_name[n])]
        self.offsets: Dict[str, int] = {}
        off = 0
        for n in w_names + z_names:
            if n == (z_names[0] if z_names else None):
                self.z_begin = off
            self.offsets[n] = off
            off += (by_name[n].numel() + ALIGN - 1) // ALIGN * ALIGN
        if not z_names:
            self.z_begin = off
        self.total = off
        self.device = torch.device(device)
        self.frozen = frozen
        self.master = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        self.symmetric = bool(symmetric) and not frozen
        if self.symmetric:
            import torch.distributed._symmetric_memory as symm_mem
            self.shadow = symm_mem.empty(self.total, dtype=torch.bfloat16, device=self.device).zero_()
            self.grad = symm_mem.empty(self.total, dtype=torch.float32, device=self.device).zero_()
        else:
            self.shadow = torch.zeros(self.total, dtype=torch.bfloat16, device=self.device)
            self.grad = None if frozen else torch.zeros(self.total, dtype=torch.float32, device=self.device)
        # rank-sharded optimizer (trainer.TrainStep over peer memory): each rank keeps only ITS shard of the fp32 master
        # current; `gather_master` (a collective every rank must reach) brings the rest up to date before anything reads it
        self.master_sharded = False
        self.gather_master = None
        self.params: Dict[str, nn.Parameter] = {}
        self._by_id: Dict[int, str] = {}
        remap: Dict[int, nn.Parameter] = {}
        for n, p in named:
            o, k = self.offsets[n], p.numel()
            view = self.master[o:o + k].view(p.shape)
            if p.device.type != "meta":
                view.copy_(p.data.to(self.device, torch.float32))
            q = nn.Parameter(view, requires_grad=not frozen)
            if not frozen:
                q.grad = self.grad[o:o + k].view(p.shape)
            remap[id(p)] = q
            self.params[n] = q
            self._by_id[id(q)] = n
        for m in module.modules():  # swap the (possibly meta) parameters for the flat views, ties preserved
            for k, p in list(m._parameters.items()):
                if p is not None and id(p) in remap:
                    m._parameters[k] = remap[id(p)]
        self._touched = set()
        self._untouched: list = []          # (offset, numel) of GEMM weights not written in the previous step
        self._untouched_key = None
        self._lin_span: Dict[str, tuple] = {}
        self._shadow_version = -1  # version counter of `master` when the bf16 shadow was last refreshed
        self.external_step = False  # True while a TrainStep drives begin_step / finish_backward itself
        # requires-grad anchor so that autograd records the first block of a forward pass
        self.anchor = torch.zeros(1, device=self.device, requires_grad=not frozen)

    # ------------------------------------------------------------------ views
    def name_of(self, p: nn.Parameter) -> str:
        return self._by_id[id(p)]

    def _span(self, ps: Sequence[nn.Parameter]):
        names = [self.name_of(p) for p in ps]
        o0 = self.offsets[names[0]]
        o = o0
        for n, p in zip(names, ps):
            if self.offsets[n] != o:
                raise RuntimeError(f"parameters {names} are not adjacent in the flat store")
            if p.numel() % ALIGN and n != names[-1]:
                raise RuntimeError(f"{n}: size {p.numel()} not a multiple of {ALIGN}; cannot fuse")
            o += p.numel()
        return o0, o - o0

    def w16(self, *ps: nn.Parameter) -> torch.Tensor:
        """bf16 shadow of one parameter, or of several adjacent ones stacked along dim 0."""
        o, k = self._span(ps)
        return self.shadow[o:o + k].view(-1, *ps[0].shape[1:])

    def f32(self, *ps: nn.Parameter) -> torch.Tensor:
        o, k = self._span(ps)
        return self.master[o:o + k].view(-1, *ps[0].shape[1:])

    def g32(self, *ps: nn.Parameter) -> Optional[torch.Tensor]:
        if self.frozen:
            return None
        o, k = self._span(ps)
        return self.grad[o:o + k].view(-1, *ps[0].shape[1:])

    def lin(self, weights: Sequence[nn.Parameter], biases: Optional[Sequence[nn.Parameter]]) -> Lin:
        b32 = self.f32(*biases).view(-1) if biases else None
        gb = self.g32(*biases).view(-1) if (biases and not self.frozen) else None
        key = f"{self.name_of(weights[0])}+{len(weights)}"
        self._lin_span[key] = self._span(weights)
        return Lin(self.w16(*weights), b32, self.g32(*weights), gb, key)

    # ------------------------------------------------------------------ step protocol
    def begin_step(self):
        """Call before the forward pass of a training step: zero the atomically-accumulated gradient
        region (biases, LayerNorm, embeddings); GEMM weight gradients are overwritten on first touch."""
        if self.frozen:
            return
        self.grad[self.z_begin:].zero_()
        # GEMM weights no kernel wrote in the previous step (parameters the configuration never uses): the set is a
        # property of the model, so zero them HERE, on the compute stream and before the forward pass -- not after the
        # backward pass, where a data-parallel bucket holding them may already be in flight on the communication stream
        for o, k in self._untouched:
            self.grad[o:o + k].zero_()
        self._touched.clear()
        for p in self.params.values():  # optimizer.zero_grad(set_to_none=True) drops the views
            if p.grad is None:
                o = self.offsets[self.name_of(p)]
                p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def finish_backward(self):
        """GEMM-weight gradients never touched in this step (unused parameters) must read as zero."""
        if self.frozen:
            return
        key = frozenset(self._touched)
        if key == self._untouched_key:
            return  # same set as last step: begin_step already zeroed them
        spans = sorted(self._lin_span[k] for k in self._touched)
        untouched = []
        for n, p in self.params.items():
            o = self.offsets[n]
            if o < self.z_begin and not any(a <= o < a + k for a, k in spans):
                untouched.append((o, p.numel()))
        for o, k in untouched:  # first step (or the used set changed): zero now; from the next step on begin_step does it
            if (o, k) not in self._untouched:
                self.grad[o:o + k].zero_()
        self._untouched, self._untouched_key = untouched, key

    def touch(self, key: str) -> bool:
        """True if `key`'s weight gradient already holds this step's partial sum (-> accumulate)."""
        seen = key in self._touched
        self._touched.add(key)
        return seen

    @property
    def dirty_shadow(self) -> bool:
        """True when the fp32 master changed through torch (optimizer.step(), load_state_dict, p.data edits made through
        the parameter views all bump the shared version counter) since the bf16 compute shadow was refreshed.  The fused
        AdamW kernel writes both copies itself and does not count."""
        return self.master._version != self._shadow_version

    def sync_master(self):
        """Make the fp32 master complete on this rank (no-op unless a rank-sharded optimizer owns it).  COLLECTIVE."""
        if self.master_sharded and self.gather_master is not None:
            self.gather_master()

    def refresh_shadow(self):
        self.sync_master()
        K.cast_bf16(self.master, self.shadow)
        self._shadow_version = self.master._version

    def load_state_dict_flat(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        missing = []
        for n, p in self.params.items():
            if n in sd:
                p.data.copy_(sd[n].to(self.device, torch.float32))
            else:
                missing.append(n)
        if strict and missing:
            raise KeyError(f"missing parameters: {missing[:5]} ...")
        self.refresh_shadow()
