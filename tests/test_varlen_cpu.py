"""CPU: host-side batch assembly for packed (varlen) article rows -- the replacement of the collate's right padding
(DNYT:957-972, TRAIN:255-271): index arithmetic only."""
import pytest
import torch

from vacnic_b200 import synthetic, varlen


@pytest.mark.parametrize("B,L,bucket", [(1, 17, 8), (4, 64, 32), (16, 1024, 512), (3, 130, 512)])
def test_pack_articles_geometry(B, L, bucket):
    ids = synthetic.make_batch(B=B, L=L, T=4, seed=B * 7 + L)["article_ids"]
    p = varlen.pack_articles(ids, pad_id=1, bucket=bucket)
    lens = (ids != 1).sum(1)
    M = int(lens.sum())
    assert p["ids"].numel() % bucket == 0 and 0 <= p["ids"].numel() - M < bucket
    assert p["len"].tolist() == lens.tolist() and p["start"].tolist() == [int(lens[:b].sum()) for b in range(B)]
    assert p["qlen"][:-1].tolist() == lens[:-1].tolist() and int(p["qlen"][-1]) == int(lens[-1]) + p["ids"].numel() - M
    for b in range(B):                                   # tokens and positions of every article, in order
        s, n = int(p["start"][b]), int(lens[b])
        assert torch.equal(p["ids"][s:s + n], ids[b, :n]) and p["pos"][s:s + n].tolist() == list(range(n))
    assert bool((p["ids"][M:] == 1).all()) and bool((p["pos"][M:] == 0).all())    # bucket tail = pad tokens
    # query ranges tile the packed rows exactly; key ranges leave only the tail uncovered
    assert int(p["start"][-1] + p["qlen"][-1]) == p["ids"].numel()
    # round trip
    x = torch.arange(p["ids"].numel(), dtype=torch.float32)[:, None].repeat(1, 2)
    u = varlen.unpack_rows(x, p["start"], p["len"], L, fill=-1.0)
    for b in range(B):
        n = int(lens[b])
        assert u[b, :n, 0].tolist() == list(range(int(p["start"][b]), int(p["start"][b]) + n)) and bool((u[b, n:] == -1).all())


def test_pack_articles_rejects_what_the_collate_never_produces():
    ids = torch.tensor([[0, 5, 1, 6, 2], [0, 7, 2, 1, 1]])      # pad in the middle of row 0
    with pytest.raises(ValueError):
        varlen.pack_articles(ids)
    with pytest.raises(ValueError):
        varlen.pack_articles(torch.tensor([[1, 1, 1], [0, 5, 2]]))   # empty article
