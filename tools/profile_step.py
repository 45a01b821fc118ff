"""One eager (no CUDA graph) training step at the bench workload, for ncu launch lists:
  python tools/profile_step.py [--small] [--warm N]
Prints the number of library launches of the warm-up steps so that `ncu -s` can skip them."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vacnic_b200 import lib, spec, synthetic  # noqa: E402
from vacnic_b200.modeling import VacnicBart  # noqa: E402
from vacnic_b200.trainer import TrainStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--small", action="store_true")
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--batch", type=int, default=16)
args = ap.parse_args()
dev = torch.device("cuda:0")
cfg = spec.bart_base() if args.small else spec.bart_large()
gcfg = spec.bart_base(stock=True) if args.small else spec.VacnicConfig(stock=True)
model = VacnicBart(cfg, device=dev, p_drop=0.1, seed=1)
guide = VacnicBart(gcfg, device=dev, p_drop=0.0, seed=2, frozen=True)
ts = TrainStep(model, guide, use_graph=False)
L, T = (512, 40) if args.small else (1024, 64)
b = TrainStep.prepare(synthetic.make_batch(B=args.batch, L=L, T=T, seed=1), cfg)
for _ in range(args.warm):
    ts.step(b, prepared=True)
torch.cuda.synchronize()
print("warm launches", lib.launch_count(), flush=True)
torch.cuda.nvtx.range_push("step")
ts.step(b, prepared=True)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("total launches", lib.launch_count(), "losses", {k: float(v) for k, v in ts.losses.items()}, flush=True)

if os.environ.get("GEMM_SHAPES"):
    import collections
    from vacnic_b200 import kernels as K
    K.PROFILE = []
    ts.step(b, prepared=True)
    torch.cuda.synchronize()
    prof, K.PROFILE = K.PROFILE, None
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for fl, e0, e1, shape, _pair in prof:
        a = agg[shape]
        a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += fl
    tot = sum(a[1] for a in agg.values())
    print(f"GEMM total {tot:.2f} ms, {sum(a[2] for a in agg.values())/1e12:.2f} TFLOP")
    for shape, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{a[1]:8.3f} ms {100*a[1]/tot:5.1f}% n={a[0]:4d} M,N,K,batch={shape} {a[2]/1e9/a[1]:8.1f} TF/s")
