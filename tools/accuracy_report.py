"""Logit accuracy of the B200 path next to the reference arithmetic in its own reduced-precision modes, all against
the fp32 oracle on the same device, weights and inputs (the comparator SURVEY.md §8d asks for):
  ours            vacnic_b200 (bf16 storage, fp32 accumulate / statistics)
  ref autocast    oracle under torch.autocast(bfloat16): bf16 matmuls, fp32 LayerNorm / softmax / residual stream
  ref bf16        oracle with every weight and activation in bf16 (model.bfloat16())
    python tools/accuracy_report.py [--large] [--cpu-ref-only]
`report()` is what tests/test_model_gpu.py::test_accuracy_next_to_reference_bf16_modes asserts on."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import model as OM  # noqa: E402
from vacnic_b200 import spec, synthetic  # noqa: E402


def _inputs(cfg, batch):
    src = batch["article_ids"]
    kw = dict(input_ids=src, attention_mask=OM.src_mask(src), image_features=batch["image_features"])
    if not cfg.only_image:
        face = batch["face_emb"]
        kw.update(face_features=face, face_mask=OM.src_mask(face[:, :, -1]), name_ids=batch["names_art_ids"],
                  name_mask=OM.src_mask(batch["names_art_ids"]))
    return kw


def _cast(obj, dt):
    if isinstance(obj, dict):
        return {k: _cast(v, dt) for k, v in obj.items()}
    return obj.to(dt) if torch.is_tensor(obj) and obj.is_floating_point() else obj


def reference_logits(sd, cfg, batch, mode):
    """fp32 | autocast | bf16 logits of the oracle restatement (float32 on return)."""
    kw = _inputs(cfg, batch)
    dec_in = OM.shift_tokens_right(batch["caption_ids"], 1, 2)
    with torch.no_grad():
        if mode == "fp32":
            return OM.model_forward(sd, cfg.as_dict(), decoder_input_ids=dec_in, **kw)["logits"].float()
        if mode == "autocast":
            with torch.autocast(device_type=batch["article_ids"].device.type, dtype=torch.bfloat16):
                return OM.model_forward(sd, cfg.as_dict(), decoder_input_ids=dec_in, **kw)["logits"].float()
        sd16, kw16 = _cast(sd, torch.bfloat16), _cast(kw, torch.bfloat16)
        return OM.model_forward(sd16, cfg.as_dict(), decoder_input_ids=dec_in, **kw16)["logits"].float()


def report(cfg, dev, B, L, T, weight_seed=7, batch_seed=3, ours=True):
    sd = {k: v.to(dev) for k, v in spec.test_state_dict(cfg, weight_seed).items()}
    batch = synthetic.to_device(synthetic.make_batch(B=B, L=L, T=T, seed=batch_seed), dev)
    ref = reference_logits(sd, cfg, batch, "fp32")
    valid = (batch["caption_ids"] != 1)
    out = {}
    cands = {"ref_autocast": lambda: reference_logits(sd, cfg, batch, "autocast"),
             "ref_bf16": lambda: reference_logits(sd, cfg, batch, "bf16")}
    if ours:
        from vacnic_b200.modeling import VacnicBart
        m = VacnicBart(cfg, device=dev, p_drop=0.0)
        m.load_reference_state_dict({k: v.cpu() for k, v in sd.items()})
        m.eval()

        def run_ours():
            with torch.no_grad():
                dec_in = OM.shift_tokens_right(batch["caption_ids"], 1, 2)
                return m(decoder_input_ids=dec_in, **_inputs(cfg, batch))["logits"].float()
        cands["ours"] = run_ours
    for name, fn in cands.items():
        err = (fn() - ref).abs()[valid]
        out[name] = {"max_abs": err.max().item(), "mean_abs": err.mean().item()}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--large", action="store_true")
    ap.add_argument("--cpu-ref-only", action="store_true", help="reference modes only, on the CPU (no CUDA extension needed)")
    a = ap.parse_args()
    dev = torch.device("cpu" if a.cpu_ref_only else "cuda:0")
    cfg = spec.bart_large() if a.large else spec.bart_base()
    B, L, T = (2, 1024, 64) if a.large else (2, 512, 40)
    if a.cpu_ref_only:
        B, L, T = 1, 64, 8
    r = report(cfg, dev, B, L, T, ours=not a.cpu_ref_only)
    print(f"logit error against the fp32 oracle ({'BART-large' if a.large else 'BART-base'}, B={B}, L={L}, T={T}; non-pad positions)")
    for k, v in r.items():
        print(f"  {k:14s} max-abs {v['max_abs']:.4f}  mean-abs {v['mean_abs']:.5f}")
