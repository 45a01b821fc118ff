"""2+ GPU check (torchrun) of the rank-sharded optimizer over NVLink peer memory (csrc/dp.cu, trainer.TrainStep
exchange="p2p" / "p2p-mc") against the NCCL all-reduce + replicated AdamW path (exchange="nccl"):

  1. after K graph-captured steps on rank-local batches the bf16 shadows (what the forward pass computes with) are
     BIT-IDENTICAL across ranks, for every exchange;
  2. p2p == nccl: same losses step by step, and -- after gather_master() -- the same fp32 master up to the last-bit
     differences of two separately compiled AdamW kernels (checked to 1e-6 relative; at world 2 the gradient sum itself is
     order-free, at world > 2 NCCL's ring order differs from rank order by fp32 rounding);
  3. gather_master() leaves every rank with the complete master (equal across ranks, and shadow == bf16(master)).

Usage: torchrun --nproc-per-node N tools/dp_p2p_check.py [--large] [--steps K]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from vacnic_b200 import spec, synthetic  # noqa: E402
from vacnic_b200.modeling import VacnicBart  # noqa: E402
from vacnic_b200.trainer import TrainStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--modes", default="nccl,nccl,nccl-serial,p2p-serial,p2p,p2p-mc")
ap.add_argument("--lr", type=float, default=1e-5)
args = ap.parse_args()
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = spec.VacnicConfig(d_model=768, heads=12, ffn=1024, enc_layers=4, dec_layers=2, prompt_size=4, max_pos=128)
gcfg = spec.VacnicConfig(**{**cfg.as_dict(), "stock": True})
batches = [TrainStep.prepare(synthetic.make_batch(B=2, L=64, T=12, seed=10 + 7 * i + rank), cfg) for i in range(2)]


def all_equal(t, what):
    h = torch.stack([t.double().sum(), t.double().abs().sum(), (t.double() * torch.arange(t.numel(), device=dev) % 7).sum()])
    hs = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(hs, h)
    ok = all(torch.equal(hs[0], x) for x in hs)
    assert ok, (what, [x.tolist() for x in hs])


results = {}
FAILED = []
for mode_i, mode in enumerate(args.modes.split(",")):
    os.environ["VACNIC_DP_SERIAL"] = "1" if mode.endswith("-serial") else "0"
    xmode = mode.replace("-serial", "")
    model = VacnicBart(cfg, device=dev, p_drop=0.0, seed=5, symmetric=(xmode != "nccl"))   # same seed -> identical replicas
    guide = VacnicBart(gcfg, device=dev, p_drop=0.0, seed=6, frozen=True)
    try:
        ts = TrainStep(model, guide, use_graph=True, process_group=dist.group.WORLD, lr=args.lr, weight_decay=0.01, exchange=xmode)
    except RuntimeError as e:
        if mode == "p2p-mc":
            print(f"rank {rank}: {mode} unavailable: {e}", flush=True)
            continue
        raise
    losses = []
    for i in range(args.steps):
        out = ts.step(batches[i % 2], prepared=True)
        losses.append([float(out[k]) for k in ("txt", "margin", "secla")])
    torch.cuda.synchronize()
    dist.barrier()
    all_equal(model.store.shadow, f"{mode}: bf16 shadow differs across ranks")
    ts.gather_master()
    torch.cuda.synchronize()
    all_equal(model.store.master, f"{mode}: fp32 master differs across ranks after gather_master")
    sh = model.store.master.to(torch.bfloat16)
    assert torch.equal(sh, model.store.shadow), f"{mode}: shadow is not the bf16 image of the gathered master"
    results[mode if mode not in results else f"{mode}#{mode_i}"] = (losses, model.store.master.clone())
    print(f"rank {rank}: {mode:7s} ok  losses[-1]={losses[-1]}  launches/step={ts.launches_per_step}", flush=True)
    ts.close()
    del ts, model, guide
    torch.cuda.empty_cache()

ref_l, ref_p = results["nccl"]
init = VacnicBart(cfg, device=dev, p_drop=0.0, seed=5, symmetric=False).store.master
moved = (ref_p - init).abs()
floor = None
for mode, (l, p) in results.items():
    if mode == "nccl":
        continue
    d = (p - ref_p).abs()
    frac = (d > 0.5 * args.lr).float().mean().item()          # an Adam sign flip moves a weight by ~2 * lr
    dl = max(abs(x - y) / max(1.0, abs(y)) for a, b in zip(l, ref_l) for x, y in zip(a[:2], b[:2]))
    print(f"rank {rank}: {mode:9s} vs nccl: mean |dp| = {d.mean().item():.3e} (weights moved {moved.mean().item():.3e} on average), "
          f"max |dp| = {d.max().item():.3e}, fraction differing by > lr/2: {frac:.3e}, max rel loss diff (txt, margin) = {dl:.2e}", flush=True)
    if mode.startswith("nccl#"):
        floor = (d.mean().item(), frac, dl)      # run-to-run noise of the SAME path (fp32 atomics + sign-like Adam)
    else:
        # the peer-memory path must agree with the NCCL path as well as the NCCL path agrees with itself (x4 slack)
        f_mean, f_frac, f_dl = floor if floor is not None else (1e-7, 1e-3, 1e-4)
        # After ONE step the paths must agree to the run-to-run floor of the NCCL path itself (fp32 atomics): the fused
        # kernel IS all-reduce + AdamW.  After several steps parameters whose true gradient is zero (every k_proj.bias:
        # softmax is invariant to a constant added to all scores of a row) follow the sign of rounding noise under Adam
        # and decorrelate between ANY two runs that differ in the last bit; that is reported, not asserted.
        bad = args.steps == 1 and not (d.max().item() <= 1e-8 and dl <= 1e-6)
        if args.steps > 1 and rank == 0:
            print("   (multi-step run: cross-mode agreement is informative only; replica identity was asserted above)", flush=True)
        if bad:
            # name the parameters that differ
            st = VacnicBart(cfg, device=dev, p_drop=0.0, seed=5, symmetric=False).store
            rows = []
            for n, q in st.params.items():
                o = st.offsets[n]
                dd = d[o:o + q.numel()]
                fr = (dd > 0.5 * args.lr).float().mean().item()
                if fr > 1e-3:
                    rows.append((fr, n, o, q.numel()))
            rows.sort(reverse=True)
            if rank == 0:
                for fr, n, o, k in rows[:25]:
                    print(f"   differs: {n:70s} offset {o:10d} numel {k:9d} frac {fr:.3f}", flush=True)
            FAILED.append(mode)
dist.barrier()
assert not FAILED, FAILED
if rank == 0:
    print("dp_p2p_check: ALL OK", flush=True)
dist.destroy_process_group()
os._exit(0)
