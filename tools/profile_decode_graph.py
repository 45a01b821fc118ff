"""Per-kernel GPU time inside the replayed decode-step graph (CUPTI trace via torch.profiler), bench shape."""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from vacnic_b200 import generation, spec, synthetic  # noqa: E402
from vacnic_b200.modeling import VacnicBart  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--captions", type=int, default=256)
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
    generation.generate(model, num_beams=4, max_length=50, length_penalty=2.0, **kw)
torch.cuda.synchronize()
eng = next(iter(model._generators.values()))
eng._reset_state()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(8):
        eng.graph.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time_total > 0]
evs.sort(key=lambda e: e.time_range.start)
agg = collections.defaultdict(lambda: [0, 0.0])
busy = 0.0
for e in evs:
    name = e.name.split("(")[0].replace("void ", "").replace("vb::", "")[:70]
    agg[name][0] += 1
    agg[name][1] += e.device_time_total
    busy += e.device_time_total
span = evs[-1].time_range.end - evs[0].time_range.start
print(f"8 decode steps, {args.captions} captions x 4 beams: kernels {len(evs)}, GPU busy {busy / 8e3:.3f} ms/step, span {span / 8e3:.3f} ms/step, idle {100 * (1 - busy / span):.1f} %")
print("| kernel | launches/step | ms/step | share of span % | avg us |\n|---|---:|---:|---:|---:|")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"| `{name}` | {n / 8:.0f} | {t / 8e3:.3f} | {100 * t / span:.2f} | {t / n:.1f} |")
