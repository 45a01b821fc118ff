"""Data-parallel gradient exchange for the VACNIC training step (the reference wraps the model in
DistributedDataParallel, TRAIN:87; gradients are averaged over ranks before optimizer.step()).

One process per GPU.  All gradients live in ONE flat fp32 buffer (store.ParamStore), so the exchange is a
handful of in-place NCCL all-reduces (sum) over address ranges of that buffer, issued on a communication
stream as soon as the backward pass has finished the parameters of a bucket — the decoder + LM head first,
then the encoder layers in groups, last the embeddings / prefix modules whose gradients complete at the very
end.  The 1/world factor is folded into the fused AdamW kernel (hyper[7]), never applied to the buffer.

The bucket bookkeeping is plain tensor / process-group code with no CUDA dependency so that it is covered by
world-size-2 gloo tests on CPU (tests/test_dp_cpu.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def merge_ranges(ranges: Sequence[Tuple[int, int]]) -> List[Tuple[int, int]]:
    """Sort and coalesce half-open [a, b) ranges."""
    out: List[Tuple[int, int]] = []
    for a, b in sorted(r for r in ranges if r[1] > r[0]):
        if out and a <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], b))
        else:
            out.append((a, b))
    return out


def complement(ranges: Sequence[Tuple[int, int]], total: int) -> List[Tuple[int, int]]:
    out, pos = [], 0
    for a, b in merge_ranges(ranges):
        if a > pos:
            out.append((pos, a))
        pos = max(pos, b)
    if pos < total:
        out.append((pos, total))
    return out


class GradBuckets:
    """Address-range buckets over a flat gradient buffer.

    `spans`: parameter name -> (offset, numel) in the flat buffer; `bucket_prefixes`: for each bucket, the
    name prefixes of the parameters it owns, in the order the backward pass completes them.  Everything not
    claimed by a bucket forms the final bucket reduced by `finish()`."""

    def __init__(self, grad: torch.Tensor, spans: Dict[str, Tuple[int, int]], bucket_prefixes: Sequence[Sequence[str]],
                 group=None, max_gap: int = 63):
        self.grad, self.group = grad, group
        # only an EXPLICIT group communicates: a step object without one never issues a collective, even inside a
        # multi-rank job (e.g. a rank-local profiling pass)
        self.world = dist.get_world_size(group) if group is not None else 1
        self.buckets: List[List[Tuple[int, int]]] = []
        claimed: List[Tuple[int, int]] = []
        for prefixes in bucket_prefixes:
            rs = [(o, o + n) for name, (o, n) in spans.items() if any(name.startswith(p) for p in prefixes)]
            rs = merge_ranges(rs)
            # alignment padding (< 64 elements) between neighbouring parameters: bridge it so a layer is one range
            bridged: List[Tuple[int, int]] = []
            for a, b in rs:
                if bridged and a - bridged[-1][1] <= max_gap and not self._overlaps(claimed, bridged[-1][1], a):
                    bridged[-1] = (bridged[-1][0], b)
                else:
                    bridged.append((a, b))
            self.buckets.append(bridged)
            claimed += bridged
        srt = sorted(claimed)
        if any(srt[i][0] < srt[i - 1][1] for i in range(1, len(srt))):
            raise ValueError("gradient buckets overlap")
        self.rest = complement(claimed, grad.numel())
        self.done = [False] * len(self.buckets)
        self.bytes_reduced = 0

    @staticmethod
    def _overlaps(ranges, a, b) -> bool:
        return any(x < b and a < y for x, y in ranges)

    def _reduce(self, ranges):
        for a, b in ranges:
            if self.world > 1 or self.group is not None:  # an explicit group is always exercised (1-GPU debugging)
                dist.all_reduce(self.grad[a:b], op=dist.ReduceOp.SUM, group=self.group)
            self.bytes_reduced += (b - a) * self.grad.element_size()

    def begin_step(self):
        self.done = [False] * len(self.buckets)
        self.bytes_reduced = 0

    def reduce_bucket(self, i: int) -> List[Tuple[int, int]]:
        """All-reduce bucket i in place; returns its address ranges (the optimizer may update them right away)."""
        if self.done[i]:
            raise RuntimeError(f"gradient bucket {i} reduced twice in one step")
        self.done[i] = True
        self._reduce(self.buckets[i])
        return self.buckets[i]

    def finish(self) -> List[Tuple[int, int]]:
        """Reduce whatever has not been reduced yet (buckets whose hook never fired + the unclaimed remainder);
        returns those ranges."""
        out: List[Tuple[int, int]] = []
        for i, d in enumerate(self.done):
            if not d:
                out += self.reduce_bucket(i)
        self._reduce(self.rest)
        return out + list(self.rest)


def vacnic_bucket_prefixes(enc_layers: int, dec_layers: int, group_size: int = 3) -> List[List[str]]:
    """Bucket 0: LM head + every decoder layer (complete once the decoder backward is done); then the encoder
    layers from the last group to the first.  Embeddings, the ClipCap prefix MLP, visual_map and the face linear
    receive gradient until the very end of the backward pass and stay in the remainder."""
    buckets = [["lm_head."] + [f"model.decoder.layers.{i}." for i in range(dec_layers)]]
    hi = enc_layers
    while hi > 0:
        lo = max(0, hi - group_size)
        buckets.append([f"model.encoder.layers.{i}." for i in range(lo, hi)])
        hi = lo
    return buckets


def shard_of_range(a: int, b: int, rank: int, world: int, align: int = 64) -> Tuple[int, int]:
    """Sub-range of [a, b) owned by `rank` when the range is cut into `world` contiguous chunks whose length is a multiple
    of `align` elements (the last chunks may be short or empty).  Returns (begin, count)."""
    n = b - a
    chunk = (n + world - 1) // world
    chunk = (chunk + align - 1) // align * align
    lo = min(n, rank * chunk)
    hi = min(n, lo + chunk)
    return a + lo, hi - lo


class PeerShards:
    """Symmetric-memory view of the flat gradient buffer and bf16 shadow of every rank + the shard arithmetic of the
    fused reduce-scatter / AdamW / all-gather kernel (csrc/dp.cu).  Needs CUDA + NCCL-backed process group; the pure
    range logic (`shard_of_range`) is covered on CPU."""

    def __init__(self, store, group, use_multicast: bool = False):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm_mem
        if not getattr(store, "symmetric", False):
            raise ValueError("PeerShards needs a ParamStore allocated with symmetric=True")
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world not in (2, 4, 8):
            raise ValueError(f"peer-memory optimizer supports 2, 4 or 8 ranks, got {self.world}")
        self.h_grad = symm_mem.rendezvous(store.grad, group)
        self.h_shadow = symm_mem.rendezvous(store.shadow, group)
        self.grad_ptrs = (C.c_uint64 * self.world)(*[int(p) for p in self.h_grad.buffer_ptrs])
        self.shadow_ptrs = (C.c_uint64 * self.world)(*[int(p) for p in self.h_shadow.buffer_ptrs])
        if int(self.grad_ptrs[self.rank]) != store.grad.data_ptr() or int(self.shadow_ptrs[self.rank]) != store.shadow.data_ptr():
            raise RuntimeError("symmetric-memory rendezvous returned a different local address than the store's buffers")
        self.grad_mc = self.shadow_mc = 0
        if use_multicast:
            gm, sm = int(self.h_grad.multicast_ptr or 0), int(self.h_shadow.multicast_ptr or 0)
            if gm and sm:
                self.grad_mc, self.shadow_mc = gm, sm
        self.bytes_read_remote = 0
        self.bytes_written_remote = 0

    def barrier(self, channel: int):
        """All ranks reach this point of their communication stream (device-side, capturable in a CUDA graph)."""
        self.h_grad.barrier(channel=channel)

    def shard(self, a: int, b: int) -> Tuple[int, int]:
        return shard_of_range(a, b, self.rank, self.world)


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard of `n_items` independent items (captions at inference) owned by `rank`."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balance_shards(lengths: Sequence[int], world: int, per_rank: int) -> List[List[int]]:
    """Length-balanced assignment of the samples of a global batch to `world` ranks, `per_rank` samples each
    (len(lengths) == world * per_rank): longest first, each sample goes to the rank with the smallest token total that
    still has a free slot.  With packed (varlen) article rows a rank's step time is proportional to its token total, and a
    synchronous data-parallel step waits for the slowest rank; the reference's DistributedSampler (TRAIN:775) deals the
    samples out at random, which is harmless there only because every rank pads to the same [B, L] rectangle.  The
    gradient average is over the same global batch whichever rank processes which sample.  Deterministic."""
    if len(lengths) != world * per_rank:
        raise ValueError(f"balance_shards: {len(lengths)} samples for {world} ranks x {per_rank}")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    shards: List[List[int]] = [[] for _ in range(world)]
    totals = [0] * world
    for i in order:
        r = min((r for r in range(world) if len(shards[r]) < per_rank), key=lambda r: (totals[r], r))
        shards[r].append(i)
        totals[r] += int(lengths[i])
    return [sorted(sh) for sh in shards]
