"""One eager (no CUDA graph) beam-4 generation at the bench workload, for ncu launch lists:
  python tools/profile_infer.py [--captions 64] [--max-length 8]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vacnic_b200 import generation, lib, spec, synthetic  # noqa: E402
from vacnic_b200.modeling import VacnicBart  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--captions", type=int, default=64)
ap.add_argument("--max-length", type=int, default=50)
ap.add_argument("--graph", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
cfg = spec.bart_large()
model = VacnicBart(cfg, device=dev, p_drop=0.0, seed=42)
model.eval()
b = synthetic.make_batch(B=args.captions, L=1024, T=8, seed=42)
face = b["face_emb"].to(dev)
kw = dict(input_ids=b["article_ids"].to(dev), attention_mask=(b["article_ids"] != 1).to(torch.int64).to(dev),
          image_features=b["image_features"].to(dev), face_features=face, face_mask=(face[:, :, -1] != 1).to(torch.int64),
          name_ids=b["names_art_ids"].to(dev), name_mask=(b["names_art_ids"] != 1).to(torch.int64).to(dev))
for _ in range(2):
    ids = generation.generate(model, num_beams=4, max_length=args.max_length, length_penalty=2.0, use_graph=args.graph, **kw)
torch.cuda.synchronize()
print("launches", lib.launch_count(), "ids", tuple(ids.shape), flush=True)
