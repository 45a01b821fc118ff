"""Every vacnic_gemm launch of ONE eager training step at the bench workload, grouped by problem shape:
launch count, summed CUDA-event time and TFLOP/s per shape (same instrumentation as bench.py's roofline pass; eager, so
launches of a few microseconds include host gaps).  python tools/gemm_shapes.py > profiles/<name>.md"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vacnic_b200 import kernels as K, spec, synthetic  # noqa: E402
from vacnic_b200.modeling import VacnicBart  # noqa: E402
from vacnic_b200.trainer import TrainStep  # noqa: E402

dev = torch.device("cuda:0")
cfg, gcfg = spec.bart_large(), spec.VacnicConfig(stock=True)
model = VacnicBart(cfg, device=dev, p_drop=0.1, seed=1)
guide = VacnicBart(gcfg, device=dev, p_drop=0.0, seed=2, frozen=True)
ts = TrainStep(model, guide, use_graph=False)
b = {k: v.to(dev) for k, v in TrainStep.prepare(synthetic.make_batch(B=16, L=1024, T=64, seed=1), cfg).items()}
for _ in range(2):
    ts.step(b, prepared=True)
torch.cuda.synchronize()
K.PROFILE = []
ts.step(b, prepared=True)
torch.cuda.synchronize()
prof, K.PROFILE = K.PROFILE, None
agg = collections.OrderedDict()
for flops, e0, e1, shape, pair in prof:
    a = agg.setdefault((shape, pair), [0, 0.0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
    a[2] += flops
tot_ms = sum(v[1] for v in agg.values())
print(f"GEMM launches of one eager training step (B200, BART-large VACNIC, B=16, L=1024): {len(prof)} launches, "
      f"{tot_ms:.2f} ms, {sum(v[2] for v in agg.values()) / 1e12:.2f} TFLOP\n")
print("| M x N x K (x batch) | kernel | launches | total ms | share % | us / launch | TFLOP/s |")
print("|---|---|---:|---:|---:|---:|---:|")
for (shape, pair), (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    M, N, Kd, bt = shape
    print(f"| {M} x {N} x {Kd}{' x ' + str(bt) if bt > 1 else ''} | {'pair' if pair else 'single'} | {n} | {ms:.3f} | "
          f"{100 * ms / tot_ms:.1f} | {1e3 * ms / n:.1f} | {fl / 1e12 / (ms / 1e3):.0f} |")
