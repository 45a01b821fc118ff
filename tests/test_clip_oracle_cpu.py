"""CPU: the functional restatement of the CLIP ViT image tower (oracle/clip_vit.py) against the same tower assembled from
torch's own modules in OpenAI CLIP's published layout (clip/model.py: VisionTransformer, Transformer,
ResidualAttentionBlock with nn.MultiheadAttention, QuickGELU, LayerNorm) driven exactly like `extract_clip_img_feat`
(TRAIN:220-240).  The `clip` package itself is not installed (no network), so this is the strongest pin available: the
module tree below has the state_dict names of `clip_model.visual`, and the restatement must load them by name."""
from collections import OrderedDict

import torch
from torch import nn

from oracle import clip_vit as CV


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = nn.LayerNorm(d_model)

    def forward(self, x):
        x = x + self.attn(self.ln_1(x), self.ln_1(x), self.ln_1(x), need_weights=False)[0]
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads) for _ in range(layers)])

    def forward(self, x):
        return self.resblocks(x)


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution, patch_size, width, layers, heads):
        super().__init__()
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = nn.LayerNorm(width)


def script_extract_clip_img_feat(vit_backbone, x):
    """TRAIN:220-240 with `clip_model.eval().visual` already resolved."""
    with torch.no_grad():
        dtype = vit_backbone.conv1.weight.dtype
        x = vit_backbone.conv1(x.type(dtype))
        x = x.reshape(x.shape[0], x.shape[1], -1)
        x = x.permute(0, 2, 1)
        x = torch.cat([vit_backbone.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype), x], dim=1)
        x = x + vit_backbone.positional_embedding.to(x.dtype)
        x = vit_backbone.ln_pre(x)
        x = x.permute(1, 0, 2)
        x = vit_backbone.transformer(x)
        x = x.permute(1, 0, 2)
        x_cls = vit_backbone.ln_post(x[:, 0, :]).float()
        x = vit_backbone.ln_post(x[:, 1:, :]).float()
    return x, x_cls


def test_restatement_matches_the_torch_module_tower():
    cfg = CV.vit_cfg(width=256, layers=3, patch=16, image=96)
    sd = CV.random_state_dict(cfg, 3)
    vit = VisionTransformer(cfg["image"], cfg["patch"], cfg["width"], cfg["layers"], cfg["heads"]).eval()
    missing, unexpected = vit.load_state_dict(sd, strict=True)      # same names as clip_model.visual (minus `proj`)
    assert not missing and not unexpected
    assert set(sd) == set(CV.param_shapes(cfg))
    img = torch.randn(2, 3, 96, 96, generator=torch.Generator().manual_seed(1))
    x_ref, cls_ref = script_extract_clip_img_feat(vit, img)
    x, x_cls = CV.extract_clip_img_feat(sd, cfg, img)
    assert x.shape == (2, 36, 256) and x_cls.shape == (2, 256) and x.dtype == torch.float32
    assert (x - x_ref).abs().max().item() <= 1e-5 and (x_cls - cls_ref).abs().max().item() <= 1e-5


def test_vit_b16_shapes():
    cfg = CV.vit_cfg()  # ViT-B/16 at 224 (run_full_train.sh:6): 197 tokens, 12 layers, 12 heads
    assert cfg["tokens"] == 197 and cfg["heads"] == 12
    sh = CV.param_shapes(cfg)
    assert sh["conv1.weight"] == (768, 3, 16, 16) and sh["positional_embedding"] == (197, 768)
    assert sum(torch.Size(s).numel() for s in sh.values()) == 85_799_424   # + proj (768 x 512) = the 86.2 M of ViT-B/16's visual tower
