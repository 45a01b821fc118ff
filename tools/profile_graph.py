"""Per-kernel GPU time INSIDE the captured training-step graph (CUPTI activity trace through torch.profiler):
unlike an ncu launch list (serialised, cold cache) this shows what each kernel costs back to back in the replayed
graph, and how much of the step the GPU idles between kernels.
  python tools/profile_graph.py [--small]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from vacnic_b200 import spec, synthetic  # noqa: E402
from vacnic_b200.modeling import VacnicBart  # noqa: E402
from vacnic_b200.trainer import TrainStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--small", action="store_true")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--no-varlen", action="store_true")
ap.add_argument("--pdl", default=None, help="set VACNIC_PDL (0 = kernel durations do not overlap)")
args = ap.parse_args()
if args.pdl is not None:
    os.environ["VACNIC_PDL"] = args.pdl
dev = torch.device("cuda:0")
cfg = spec.bart_base() if args.small else spec.bart_large()
gcfg = spec.bart_base(stock=True) if args.small else spec.VacnicConfig(stock=True)
model = VacnicBart(cfg, device=dev, p_drop=0.1, seed=1)
guide = VacnicBart(gcfg, device=dev, p_drop=0.0, seed=2, frozen=True)
ts = TrainStep(model, guide, use_graph=True, varlen=not args.no_varlen)
L, T = (512, 40) if args.small else (1024, 64)
b = {k: v.to(dev) for k, v in TrainStep.prepare(synthetic.make_batch(B=args.batch, L=L, T=T, seed=1), cfg, varlen=not args.no_varlen).items()}
for _ in range(4):
    ts.step(b, prepared=True)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ts.step(b, prepared=True)
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
print(f"kernels {len(evs)}, GPU busy {busy / 1e3:.2f} ms, first-to-last span {span / 1e3:.2f} ms, idle {100 * (1 - busy / span):.1f} %")
print("| kernel | launches | total ms | share of span % | avg us |\n|---|---:|---:|---:|---:|")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
    print(f"| `{name}` | {n} | {t / 1e3:.3f} | {100 * t / span:.2f} | {t / n:.1f} |")
