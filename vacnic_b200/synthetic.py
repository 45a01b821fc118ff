"""Synthetic NYTimes800k / GoodNews-shaped batches (host side), following the padding rules of the
reference collate functions (DNYT:804-913, 925-946, 957-972, 1028-1059, 1062-1113; SURVEY.md
Appendix C / §8d).  Data only: no model arithmetic happens here."""
from __future__ import annotations

from typing import Dict

import torch

PAD, BOS, EOS = 1, 0, 2
ENT, NONAME = 50265, 50266  # <ENT>, <NONAME> added at TRAIN:753-754


def _rand_row(g, length: int, total: int, vocab_hi: int) -> torch.Tensor:
    row = torch.full((total,), PAD, dtype=torch.int64)
    length = max(2, min(length, total))
    row[0] = BOS
    if length > 2:
        row[1:length - 1] = torch.randint(3, vocab_hi, (length - 2,), generator=g)
    row[length - 1] = EOS
    return row


def make_batch(B: int, L: int, T: int, seed: int = 42, F: int = 4, N: int = 8, E: int = 80, name_len: int = 8,
               vocab: int = 50267, full_length: bool = False) -> Dict[str, torch.Tensor]:
    """CPU tensors with the keys the reference training loop reads (TRAIN:255-281, 326-327).

    article_ids [B,L], caption_ids [B,T] int64 right-padded with 1; image_features [B,768] fp32 (stands
    in for CLIP ln_post(CLS), TRAIN:236); face_emb [B,F,512] L2-normalised rows, missing faces = ones
    (DNYT:831); names_art_ids [B,E] = [0] + spans joined by <ENT> + [2] padded to exactly E;
    names_ids [B,N,name_len] = [0]+tokens+[2], missing spans = [0,<NONAME>,2,1...] (DNYT:942)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    hi = min(vocab, 50265)
    art = torch.stack([_rand_row(g, L if full_length or i == 0 else int(torch.randint(L // 2, L + 1, (1,), generator=g)), L, hi)
                       for i in range(B)])
    cap = torch.stack([_rand_row(g, T if full_length or i == 0 else int(torch.randint(max(2, T // 4), T + 1, (1,), generator=g)), T, hi)
                       for i in range(B)])
    img = torch.randn(B, 768, generator=g)
    face = torch.ones(B, F, 512)
    for i in range(B):
        nf = int(torch.randint(0, F + 1, (1,), generator=g))
        if i == 0:
            nf = F
        v = torch.randn(nf, 512, generator=g)
        face[i, :nf] = v / v.norm(dim=-1, keepdim=True)
    names_art = torch.full((B, E), PAD, dtype=torch.int64)
    for i in range(B):
        toks = [BOS]
        nspan = int(torch.randint(0, 9, (1,), generator=g))
        if nspan == 0:
            toks.append(NONAME)
        for s in range(nspan):
            if s:
                toks.append(ENT)
            toks += torch.randint(3, hi, (int(torch.randint(1, 4, (1,), generator=g)),), generator=g).tolist()
        toks = toks[: E - 1] + [EOS]
        names_art[i, : len(toks)] = torch.tensor(toks)
    names = torch.full((B, N, name_len), PAD, dtype=torch.int64)
    for i in range(B):
        k = int(torch.randint(0, N + 1, (1,), generator=g))
        for n in range(N):
            if n < k:
                ln = int(torch.randint(1, name_len - 1, (1,), generator=g))
                row = [BOS] + torch.randint(3, hi, (ln,), generator=g).tolist() + [EOS]
            else:
                row = [BOS, NONAME, EOS]
            names[i, n, : len(row)] = torch.tensor(row)
    return dict(article_ids=art, caption_ids=cap, image_features=img, face_emb=face, names_art_ids=names_art,
                names_ids=names)


def to_device(batch: Dict[str, torch.Tensor], device, pinned: bool = False) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in batch.items():
        if pinned and device == "cpu":
            v = v.pin_memory()
        out[k] = v.to(device, non_blocking=True)
    return out
