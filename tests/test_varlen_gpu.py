"""Packed (varlen) article rows on the model path: the device never sees the collate's padding.
On the SAME model and weights the packed forward / backward must reproduce the padded one on everything the reference
defines: logits, losses, the encoder memory on valid tokens, and every parameter gradient (pad rows contribute exactly
zero in the reference: masked keys get probability 0, so their output gradient is 0).  Also against the fp32 oracle (which
pads), and through TrainStep with one captured graph per row bucket."""
import pytest
import torch

from oracle import model as OM
from vacnic_b200 import spec, synthetic, varlen

pytestmark = pytest.mark.gpu


def _models(dev, cfg, seed=5):
    from vacnic_b200.modeling import VacnicBart
    sd = spec.test_state_dict(cfg, seed, lm_scale=2.0)
    m = VacnicBart(cfg, device=dev, p_drop=0.0)
    m.load_reference_state_dict(sd)
    return m, sd


def _kw(cfg, batch):
    src = batch["article_ids"]
    kw = dict(input_ids=src, attention_mask=OM.src_mask(src), image_features=batch["image_features"])
    if not cfg.only_image:
        face = batch["face_emb"]
        kw.update(face_features=face, face_mask=OM.src_mask(face[:, :, -1]), name_ids=batch["names_art_ids"],
                  name_mask=OM.src_mask(batch["names_art_ids"]))
    return kw


def _pack(cfg, batch_cpu, dev, T):
    p = {k: v.to(dev) for k, v in varlen.pack_articles(batch_cpu["article_ids"], cfg.pad_token_id).items()}
    B, L = batch_cpu["article_ids"].shape
    prefix = cfg.prompt_size + (0 if cfg.only_image else cfg.max_ner_type_len_gt)
    return varlen.ArticlePack(p, B, L, prefix, T), p


@pytest.mark.parametrize("only_image", [False, True])
@pytest.mark.parametrize("B,L,T,seed", [(3, 300, 12, 1), (2, 1024, 9, 2), (5, 130, 7, 3)])
def test_packed_forward_backward_equals_padded(cuda_device, only_image, B, L, T, seed):
    dev = cuda_device
    cfg = spec.VacnicConfig(d_model=768, heads=12, ffn=1024, enc_layers=2, dec_layers=2, prompt_size=4, max_pos=1024,
                            only_image=only_image)
    m, sd = _models(dev, cfg)
    m.train()
    cpu = synthetic.make_batch(B=B, L=L, T=T, seed=seed)
    batch = synthetic.to_device(cpu, dev)
    tgt = batch["caption_ids"]
    dec_in = OM.shift_tokens_right(tgt, 1, 2)
    results = {}
    for mode in ("padded", "packed"):
        pack, p = _pack(cfg, cpu, dev, T) if mode == "packed" else (None, None)
        out = m(decoder_input_ids=dec_in, ce_targets=tgt, article_pack=pack, **_kw(cfg, batch))
        (out["loss"] + 1e-3 * out["decoder_hidden_states"][-1].float().pow(2).mean()).backward()
        torch.cuda.synchronize()
        enc = out["encoder_last_hidden_state"]
        if mode == "packed":
            assert enc.shape[0] == 1 and enc.shape[1] % varlen.ROW_BUCKET == 0
            enc = varlen.unpack_rows(enc[0], p["start"], p["len"], L)
        results[mode] = (out["logits"].float().clone(), float(out["loss"]), enc.float().clone(), m.store.grad.clone())
    (lg_a, loss_a, enc_a, g_a), (lg_b, loss_b, enc_b, g_b) = results["padded"], results["packed"]
    valid = OM.src_mask(batch["article_ids"]).bool()
    assert (lg_a - lg_b).abs().max().item() <= 2e-2 and abs(loss_a - loss_b) <= 1e-3 * abs(loss_a)
    assert (enc_a[valid] - enc_b[valid]).abs().max().item() <= 6e-2
    assert torch.nn.functional.cosine_similarity(enc_a[valid].flatten(), enc_b[valid].flatten(), dim=0).item() >= 0.9999
    assert torch.isfinite(g_b).all()
    worst = 1.0
    for n, q in m.store.params.items():
        o = m.store.offsets[n]
        a, b = g_a[o:o + q.numel()], g_b[o:o + q.numel()]
        na, nb = a.norm().item(), b.norm().item()
        if na < 1e-5 or n.endswith("k_proj.bias"):   # analytically zero (softmax is invariant to the key bias): rounding noise only
            assert nb < 2e-3, (n, nb)
            continue
        cos = (a @ b).item() / (na * nb + 1e-30)
        worst = min(worst, cos)
        assert cos >= 0.995 and abs(na - nb) <= 3e-2 * na, (n, cos, na, nb)
    # and against the fp32 oracle (padded): same tolerance as the padded path has
    with torch.no_grad():
        o = OM.model_forward({k: v.to(dev) for k, v in sd.items()}, cfg.as_dict(), decoder_input_ids=dec_in, **_kw(cfg, batch))
    assert (lg_b - o["logits"]).abs().max().item() <= 3e-2 * 2.0


def test_trainstep_varlen_buckets_and_matches_padded(cuda_device):
    """TrainStep(varlen=True): one captured graph per packed-row bucket (shared memory pool), batches of different
    buckets interleaved, losses equal to the padded TrainStep taking the same steps."""
    from vacnic_b200.modeling import VacnicBart
    from vacnic_b200.trainer import TrainStep
    dev = cuda_device
    cfg = spec.VacnicConfig(d_model=768, heads=12, ffn=1024, enc_layers=2, dec_layers=2, prompt_size=4, max_pos=1024)
    gcfg = spec.VacnicConfig(**{**cfg.as_dict(), "stock": True})
    batches = [synthetic.make_batch(B=4, L=800, T=10, seed=s) for s in (11, 12, 13, 11)]   # row buckets 3072, 3072, 2560, 3072
    losses = {}
    for vl in (False, True):
        m = VacnicBart(cfg, device=dev, p_drop=0.0, seed=3)
        g = VacnicBart(gcfg, device=dev, p_drop=0.0, seed=4, frozen=True)
        ts = TrainStep(m, g, lr=1e-5, use_graph=True, varlen=vl)
        out = []
        for b in batches:
            r = ts.step(b)
            out.append({k: float(v.detach()) for k, v in r.items()})
        torch.cuda.synchronize()
        losses[vl] = out
        if vl:
            keys = {varlen.pack_articles(b["article_ids"])["ids"].numel() for b in batches}
            assert set(ts._graphs) == keys and len(keys) >= 2      # really exercised more than one bucket
    for a, b in zip(losses[False], losses[True]):
        for k in a:
            assert abs(a[k] - b[k]) <= 5e-3 * max(1.0, abs(a[k])), (k, losses)


def test_generation_with_packed_encoder_is_identical(cuda_device):
    """Decode engine: the packed encoder gives the same memory on the valid tokens (same per-row arithmetic and key blocks;
    only the key block cut by the sequence end is evaluated by the mask-free kernel instantiation with one rounding less,
    i.e. last-bit differences), and identical beam-4 / greedy token ids; pad rows of the un-packed memory are zero and
    never attended."""
    from vacnic_b200 import generation
    from vacnic_b200.modeling import VacnicBart
    dev = cuda_device
    cfg = spec.VacnicConfig(d_model=768, heads=12, ffn=1024, enc_layers=2, dec_layers=2, prompt_size=4, max_pos=1024)
    m = VacnicBart(cfg, device=dev, p_drop=0.0)
    m.load_reference_state_dict(spec.test_state_dict(cfg, 9, lm_scale=8.0))
    m.eval()
    batch = synthetic.to_device(synthetic.make_batch(B=6, L=300, T=8, seed=4), dev)
    kw = _kw(cfg, batch)
    outs = {}
    for vl in (False, True):
        for nb in (1, 4):
            eng = generation.Generator(m, 6, nb, 300, 14, length_penalty=2.0, varlen=vl)
            enc = eng.encode(generation._enc_inputs(m, *(kw[k] for k in ("input_ids", "attention_mask", "image_features",
                                                                         "face_features", "face_mask", "name_ids", "name_mask"))))
            outs[(vl, nb)] = (eng.decode().clone(), enc["last_hidden_state"].clone(), eng.cross_kv.clone(), eng.key_len.clone())
    valid = kw["attention_mask"].bool()
    for nb in (1, 4):
        (ids_a, h_a, kv_a, kl_a), (ids_b, h_b, kv_b, kl_b) = outs[(False, nb)], outs[(True, nb)]
        assert torch.equal(ids_a, ids_b)
        d = (h_a[valid].float() - h_b[valid].float()).abs()
        assert d.max().item() <= 3.2e-2 and d.mean().item() <= 1e-4, (d.max().item(), d.mean().item())   # <= 2 bf16 ulp at |h| ~ 4, rare
        assert bool((h_b[~valid] == 0).all())
        assert torch.equal(kl_a, kl_b)
