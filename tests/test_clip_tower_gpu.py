"""GPU parity of the CLIP ViT image tower on the sm_100a kernels (vacnic_b200.clip_tower) against the fp32 oracle
restatement (oracle/clip_vit.py, pinned on CPU to torch's own module tower), same weights and images, same device.
The reference runs this tower in fp16; tolerances are those of the BART path: LayerNorm outputs (|x| ~ 1..4) agree to
99.9 % <= htol = 2^-8 * 4 * sqrt(2 * layers + 2), every element <= 2 htol, mean-abs <= 1e-2; cosine >= 0.9995."""
import pytest
import torch

from oracle import clip_vit as CV

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,width,layers,image,B", [("vit_b16_224_full", 768, 12, 224, 4), ("small_ragged", 256, 2, 96, 3)])
def test_tower_matches_oracle(cuda_device, name, width, layers, image, B):
    from vacnic_b200.clip_tower import ClipVisionTower, extract_clip_img_feat
    cfg = CV.vit_cfg(width=width, layers=layers, patch=16, image=image)
    sd = CV.random_state_dict(cfg, 5)
    tower = ClipVisionTower({"visual." + k: v for k, v in sd.items()}, device=cuda_device)   # clip_model.state_dict() naming
    img = torch.randn(B, 3, image, image, generator=torch.Generator().manual_seed(2)).to(cuda_device)
    x, x_cls = extract_clip_img_feat(tower, img)
    dsd = {k: v.to(cuda_device) for k, v in sd.items()}
    with torch.no_grad():
        xr, cr = CV.extract_clip_img_feat(dsd, cfg, img)
    assert x.shape == xr.shape and x_cls.shape == cr.shape and x.dtype == torch.float32 and x_cls.dtype == torch.float32
    htol = 2 ** -8 * 4 * (2 * layers + 2) ** 0.5
    for got, want, what in ((x, xr, "patch tokens"), (x_cls, cr, "cls")):
        err = (got - want).abs()
        q999 = torch.quantile(err.flatten()[:1_000_000], 0.999).item()
        cos = torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
        assert err.mean().item() <= 1e-2 and q999 <= htol and err.max().item() <= 2 * htol and cos >= 0.9995, \
            (what, err.max().item(), q999, err.mean().item(), cos)


def test_tower_feeds_the_prefix_mlp(cuda_device):
    """x_cls [B, 768] is what the scripts hand to the model as `image_features` (TRAIN:236, 281)."""
    from vacnic_b200 import spec, synthetic
    from vacnic_b200.clip_tower import ClipVisionTower
    from vacnic_b200.modeling import VacnicBart
    cfgv = CV.vit_cfg(width=768, layers=2, patch=16, image=64)
    tower = ClipVisionTower(CV.random_state_dict(cfgv, 6), device=cuda_device)
    _, x_cls = tower(torch.randn(2, 3, 64, 64, device=cuda_device))
    cfg = spec.VacnicConfig(d_model=768, heads=12, ffn=1024, enc_layers=1, dec_layers=1, prompt_size=4, max_pos=128)
    m = VacnicBart(cfg, device=cuda_device, p_drop=0.0)
    m.eval()
    batch = synthetic.to_device(synthetic.make_batch(B=2, L=40, T=8, seed=1), cuda_device)
    face = batch["face_emb"]
    with torch.no_grad():
        out = m(input_ids=batch["article_ids"], attention_mask=(batch["article_ids"] != 1).long(), decoder_input_ids=batch["caption_ids"],
                image_features=x_cls, face_features=face, face_mask=(face[:, :, -1] != 1).long(), name_ids=batch["names_art_ids"],
                name_mask=(batch["names_art_ids"] != 1).long())
    assert out["logits"].shape[:2] == (2, 8) and torch.isfinite(out["logits"]).all()
