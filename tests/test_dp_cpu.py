"""CPU (gloo, world size 2) tests of the data-parallel plumbing: bucketed in-place all-reduce of a flat gradient
buffer (vacnic_b200.dp.GradBuckets, the exchange step of the training path, TRAIN:87 DDP semantics) and the
contiguous caption sharding of inference.  No CUDA kernels are involved: the bucket bookkeeping is plain
torch.distributed code."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vacnic_b200 import dp, spec


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _layout(cfg):
    """Flat layout like store.ParamStore: matrices first (decoder cross k/v in front), then vectors / embeddings."""
    shapes = spec.param_shapes(cfg)
    names = [n for n in shapes if n != "final_logits_bias" and n not in spec.TIED_TO_SHARED]
    first = [f"model.decoder.layers.{i}.encoder_attn.{p}_proj.weight" for i in range(cfg.dec_layers) for p in ("k", "v")]
    is_mat = lambda n: len(shapes[n]) == 2 and "embed_" not in n and "shared" not in n  # noqa: E731
    order = first + [n for n in names if is_mat(n) and n not in first] + [n for n in names if not is_mat(n)]
    spans, off = {}, 0
    for n in order:
        k = 1
        for d in shapes[n]:
            k *= d
        spans[n] = (off, k)
        off += (k + 63) // 64 * 64
    return spans, off


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = spec.VacnicConfig(d_model=64, heads=1, ffn=128, enc_layers=5, dec_layers=3, vocab=300, max_pos=32, prompt_size=2)
        spans, total = _layout(cfg)
        g = torch.Generator().manual_seed(100 + rank)
        grad = torch.randn(total, generator=g)
        mine = grad.clone()
        prefixes = dp.vacnic_bucket_prefixes(cfg.enc_layers, cfg.dec_layers, group_size=2)
        b = dp.GradBuckets(grad, spans, prefixes, group=dist.group.WORLD)
        solo = dp.GradBuckets(grad.clone(), spans, prefixes)  # no group: never communicates
        assert solo.world == 1
        assert b.world == world and len(b.buckets) == 1 + 3
        # the decoder bucket owns the hoisted cross k/v block at the front of the buffer as well as the decoder layers
        assert b.buckets[0][0][0] == 0
        b.begin_step()
        # markers fire in backward order; one bucket (index 2) never fires and must be picked up by finish()
        b.reduce_bucket(0)
        b.reduce_bucket(1)
        b.reduce_bucket(3)
        with pytest.raises(RuntimeError):
            b.reduce_bucket(1)
        b.finish()
        assert all(b.done)
        # every element was reduced exactly once: grad == sum over ranks of the original buffers
        others = [torch.randn(total, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        want = sum(others)
        assert torch.allclose(grad, want, atol=1e-6), (grad - want).abs().max()
        assert b.bytes_reduced == total * 4
        # a second step on fresh gradients reuses the buckets
        grad.copy_(mine)
        b.begin_step()
        b.finish()
        assert torch.allclose(grad, want, atol=1e-6)
        # inference sharding: contiguous, disjoint, covering
        lo, hi = dp.shard_range(11, rank, world)
        t = torch.zeros(11)
        t[lo:hi] = 1
        dist.all_reduce(t)
        assert bool((t == 1).all())
        open(os.path.join(tmp, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_range_helpers():
    assert dp.merge_ranges([(5, 7), (0, 2), (2, 3), (6, 9)]) == [(0, 3), (5, 9)]
    assert dp.complement([(2, 5), (7, 9)], 12) == [(0, 2), (5, 7), (9, 12)]
    assert [dp.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert dp.vacnic_bucket_prefixes(4, 2, 3)[1:] == [["model.encoder.layers.1.", "model.encoder.layers.2.", "model.encoder.layers.3."],
                                                      ["model.encoder.layers.0."]]


def test_linear_schedule_matches_transformers():
    from vacnic_b200.trainer import linear_schedule
    try:
        from transformers import get_linear_schedule_with_warmup
    except Exception:  # pragma: no cover
        pytest.skip("transformers not importable")
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1.0)
    sch = get_linear_schedule_with_warmup(opt, 5, 40)
    for step in range(45):
        assert abs(opt.param_groups[0]["lr"] - linear_schedule(step, 5, 40)) < 1e-7, step
        opt.step()
        sch.step()


def test_shard_of_range_partitions_every_range_exactly_once():
    """Rank-sharded optimizer (csrc/dp.cu): for any world size the per-rank shards of an address range are disjoint,
    cover it exactly, start on 64-element boundaries relative to the range and have lengths that are multiples of 8
    whenever the range length is (the kernel processes 8 elements per thread)."""
    from vacnic_b200.dp import shard_of_range
    for world in (1, 2, 4, 8):
        for a, b in ((0, 64), (128, 128 + 1024), (64, 64 + 7 * 64), (0, 1048576 + 192), (4096, 4096 + 80 * 320), (0, 0)):
            covered = []
            for r in range(world):
                begin, count = shard_of_range(a, b, r, world)
                assert count >= 0 and (begin - a) % 64 == 0 and (count % 8 == 0 or (b - a) % 8 != 0)
                if count:
                    covered.append((begin, begin + count))
            covered.sort()
            pos = a
            for lo, hi in covered:
                assert lo == pos, (world, a, b, covered)
                pos = hi
            assert pos == b, (world, a, b, covered)


def test_balance_shards_is_a_partition_and_evens_out_token_totals():
    import random
    from vacnic_b200.dp import balance_shards
    rnd = random.Random(3)
    for world, per in ((1, 16), (2, 16), (8, 16), (4, 3)):
        lens = [rnd.randint(512, 1024) for _ in range(world * per)]
        sh = balance_shards(lens, world, per)
        assert sorted(i for s in sh for i in s) == list(range(world * per)) and all(len(s) == per for s in sh)
        totals = [sum(lens[i] for i in s) for s in sh]
        naive = [sum(lens[r * per:(r + 1) * per]) for r in range(world)]          # DistributedSampler-style contiguous deal
        assert max(totals) - min(totals) <= max(lens)                            # within one article of each other
        assert max(totals) <= max(naive)
        assert balance_shards(lens, world, per) == sh                            # deterministic
