#!/usr/bin/env python
"""Benchmark of the VACNIC hot path (BASELINE.json): BART-large VACNIC training step, bf16 compute,
batch 16 per GPU, 1024-token article segments + 20-token ClipCap prefix, CoLaM (margin 1.0, alpha 0.5)
against the frozen stock-BART guide, SECLA face-name loss, fused AdamW — synthetic NYTimes800k-shaped
batches, random-init weights (no network for data or checkpoints).

  python bench.py [--gpus N --steps K --warmup W]          our arm (sm_100a kernels through the C ABI)
  python bench.py --impl reference [...]                    the reference algorithm on host cores (oracle port)
  python bench.py --workload infer [...]                    beam-4 caption generation instead of training

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import sys
import threading
import time
import traceback

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

PEAKS_FALLBACK = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        d["_source"] = "measured"
        return d
    d = dict(PEAKS_FALLBACK)
    d["_source"] = "fallback"
    return d


# ----------------------------------------------------------------------------------------- FLOP model
def train_flops_per_sample(cfg, L, T, with_guide=True):
    """Algorithmic FLOPs (multiply-add = 2) per sample of one training step, SURVEY.md §8(d):
    backward = 2 x forward, attention recompute not counted, causal half not discounted."""
    d, f, P, G, F, E = cfg.d_model, cfg.ffn, cfg.prompt_size, cfg.max_ner_type_len_gt, 4, cfg.max_ner_type_len
    full = not cfg.only_image
    Gk = G if full else 0
    enc = 8 * L * d * d + 4 * L * L * d + 4 * L * d * d + 4 * (P + Gk) * d * d + 4 * L * (P + Gk) * d + 4 * L * d * f + 4 * P * d * f
    if full:
        enc += 4 * F * d * 3072 + 4 * E * d * d + 4 * (F + E) * d * d + 4 * E * (F + E) * d + 2 * d * E * E + 2 * d * E * G
    dec = 8 * T * d * d + 4 * T * T * d + 4 * T * d * d + 4 * L * d * d + 4 * T * L * d + 4 * T * d * f
    head = 2 * T * d * cfg.vocab
    prefix = 2 * 768 * 384 * P + 2 * 384 * P * 768 * P + (2 * P * 768 * d if d == 1024 else 0) + (2 * F * 512 * d if full else 0)
    fwd = cfg.enc_layers * enc + cfg.dec_layers * dec + head + prefix
    stock_enc = 8 * L * d * d + 4 * L * L * d + 4 * L * d * f
    # the frozen guide: encoder + decoder only.  (HF's forward also computes the guide's logits; CoLaM never reads them
    # (TRAIN:293-296), this implementation skips that GEMM and does not count it: 6.6 of SURVEY 8(d)'s 444.9 GF)
    guide = cfg.enc_layers * stock_enc + cfg.dec_layers * dec
    return 3 * fwd + (guide if with_guide else 0), fwd, guide


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.stop = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap"}
        while not self.stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, attr in names.items():
                    if r & getattr(nv, attr, 0):
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.05)

    def __enter__(self):
        if self.nv is not None:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.nv is not None:
            self.t.join(timeout=1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_train(threads: int, steps: int, warmup: int, batch: int = 16, micro: int = 4):
    """The reference algorithm (oracle/model.py, pinned to the reference classes) on host cores: one training step =
    forward, CE + CoLaM (frozen guide) + SECLA, backward, torch AdamW, at BART-large config-2 shapes.  Bounded sample of
    the batch-`batch` workload: ONE micro-batch of `micro` samples is run forward + backward and timed (t_fb), the
    optimizer update over all 912.7 M parameters is timed once per step (t_opt), and the batch-`batch` step time is
    (batch / micro) * t_fb + t_opt -- i.e. the optimizer is amortised over the real batch, not charged per micro-batch."""
    from oracle import model as OM
    from vacnic_b200 import spec, synthetic
    torch.set_num_threads(threads)
    cfg = spec.bart_large()
    gcfg = spec.VacnicConfig(stock=True)
    sd = spec.test_state_dict(cfg, 1)
    for k in sd:
        if sd[k].is_floating_point() and k != "final_logits_bias" and k not in spec.TIED_TO_SHARED:
            sd[k].requires_grad_(True)
    for k in spec.TIED_TO_SHARED:
        sd[k] = sd["model.shared.weight"]
    gsd = spec.test_state_dict(gcfg, 2)
    params = [v for k, v in sd.items() if v.requires_grad and k not in spec.TIED_TO_SHARED]
    opt = torch.optim.AdamW(params, lr=3e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    fb, op = [], []
    for i in range(warmup + steps):
        mb = synthetic.make_batch(B=micro, L=1024, T=64, seed=100 + i, full_length=True)
        t0 = time.perf_counter()
        res = OM.training_losses(sd, cfg.as_dict(), gsd, gcfg.as_dict(), mb, margin=1.0, alpha=0.5)
        res["loss"].backward()
        t1 = time.perf_counter()
        opt.step()
        opt.zero_grad(set_to_none=True)
        t2 = time.perf_counter()
        if i >= warmup:
            fb.append(t1 - t0)
            op.append(t2 - t1)
    t_fb, t_opt = sum(fb) / len(fb), sum(op) / len(op)
    ms = 1e3 * ((batch / micro) * t_fb + t_opt)
    return batch / (ms / 1e3), ms, (f"BART-large config-2 shapes (L=1024, T=64, P=20, F=4, N=8), fp32, {len(fb)} timed sample(s): forward+backward of a "
                                    f"{micro}-sample micro-batch ({t_fb:.2f} s) x {batch // micro} + one AdamW update ({t_opt:.2f} s) = one batch-{batch} step")


def cpu_reference_infer(threads: int, max_length: int = 50):
    """The reference algorithm for caption generation (oracle encoder + transformers-5.5 `_beam_search` restatement)
    on host cores: ONE BART-large caption, L=1024, beam 4, length penalty 2.0 (bounded sample of config 3)."""
    from oracle import generate as OG
    from oracle import model as OM
    from vacnic_b200 import spec, synthetic
    torch.set_num_threads(threads)
    cfg = spec.bart_large()
    sd = spec.test_state_dict(cfg, 1)
    batch = synthetic.make_batch(B=1, L=1024, T=16, seed=42, full_length=True)
    src = batch["article_ids"]
    face = batch["face_emb"]
    inp = dict(input_ids=src, attention_mask=OM.src_mask(src), image_features=batch["image_features"], face_features=face,
               face_mask=OM.src_mask(face[:, :, -1]), name_ids=batch["names_art_ids"], name_mask=OM.src_mask(batch["names_art_ids"]))
    t0 = time.perf_counter()
    OG.beam_search(sd, cfg.as_dict(), inp, num_beams=4, max_length=max_length, length_penalty=2.0)
    dt = time.perf_counter() - t0
    return 1.0 / dt, dt * 1e3, f"1 BART-large caption, L=1024, beam 4, length_penalty 2.0, max_length {max_length}, fp32, cached decoder"


def infer_flops_bytes(cfg, C, L, nb, steps, key_lens):
    """encoder + cross-K/V projection FLOPs (tensor bound) and decode bytes (HBM bound)."""
    d, f, P, G, F, E = cfg.d_model, cfg.ffn, cfg.prompt_size, cfg.max_ner_type_len_gt, 4, cfg.max_ner_type_len
    enc = 8 * L * d * d + 4 * L * L * d + 4 * L * d * d + 4 * (P + G) * d * d + 4 * L * (P + G) * d + 4 * L * d * f + 4 * P * d * f
    enc += 4 * F * d * 3072 + 4 * E * d * d + 4 * (F + E) * d * d + 4 * E * (F + E) * d + 2 * d * E * E + 2 * d * E * G
    enc_flops = C * (cfg.enc_layers * enc + cfg.dec_layers * 4 * L * d * d)
    w_bytes = 2 * (cfg.dec_layers * (4 * d * d + 4 * d * d + 2 * d * f) + d * cfg.vocab)
    cross_bytes = sum(int(k) for k in key_lens) * cfg.dec_layers * 2 * d * 2
    return enc_flops, steps * (w_bytes + cross_bytes), cross_bytes / cfg.dec_layers


def infer_roofline(args, model, engine, devb, cfg, C, L, nb, steps_run, pk):
    """Split of one generate() call: encoder (+ cross-K/V projection) vs decode loop, and the dominant decode kernel timed
    alone with CUDA events on the launching stream."""
    from vacnic_b200 import generation, kernels as K
    enc_in = generation._enc_inputs(model, *(devb[0][k] for k in (
        "input_ids", "attention_mask", "image_features", "face_features", "face_mask", "name_ids", "name_mask")))
    for _ in range(2):  # second pass is the measurement (the first re-warms the allocator after the timed runs)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record(); engine.encode(enc_in)
        ev[1].record(); engine.decode(); ev[2].record()
        torch.cuda.synchronize()
    enc_ms, dec_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    key_lens = engine.key_len.cpu().tolist()
    enc_flops, dec_bytes, cross_bytes_launch = infer_flops_bytes(cfg, C, L, nb, steps_run, key_lens)
    n_rep = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_rep):  # every layer's K/V in turn: 12 x cross_bytes >> L2, no reuse between launches
        l = i % cfg.dec_layers
        K.decode_cross_attn(engine.qb, engine.cross_kv[l, :, 0], engine.cross_kv[l, :, 1], engine.key_mask, engine.key_len, engine.attn, nb)
    e1.record()
    torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / n_rep
    ach = cross_bytes_launch / 1e9 / (k_ms / 1e3)
    big = L == 1024 and not args.small
    return {"bound": "hbm", "kernel": "decode_cross_attn_mma_kernel (beams of a caption over its shared encoder K/V; mma.sync + cp.async rings)",
            "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel from the committed ncu --set full capture
            # (64 captions, L=1024: 214.4 MB), scaled by captions/64 -- a profile figure, not measured in this run
            "traffic": 214.4e6 * C / 64 if big else None,
            "traffic_source": "profiles/r1_ncu_decode_cross_attn.md (ncu --set full, 64 captions x L=1024), scaled linearly to this run's captions" if big else None,
            "peak_source": pk["_source"], "us_per_launch": k_ms * 1e3, "algorithmic_bytes_per_launch": cross_bytes_launch,
            "encode_ms": enc_ms, "decode_ms": dec_ms, "decode_steps": steps_run,
            "encode_tensor_frac": enc_flops / 1e12 / (enc_ms / 1e3) / pk["bf16_tflops_sustained"],
            "decode_hbm_frac": dec_bytes / 1e9 / (dec_ms / 1e3) / pk["hbm_gbs"],
            "launches_per_decode_step": engine.launches_per_step}


def run_infer(args, rank, world, dev, pk, collectives=True):
    """Beam-4 caption generation, BASELINE.json configs[2]: encoder once per caption, cached decoder, device-side
    beam search (length penalty 2.0, max_length 50), `--captions` captions per GPU, no communication.
    `collectives=False`: time this rank alone (no barrier / max-reduce inside; the caller combines the ranks)."""
    from vacnic_b200 import generation, kernels as K, spec, synthetic
    from vacnic_b200.modeling import VacnicBart
    cfg = spec.bart_base() if args.small else spec.bart_large()
    C, L, nb, max_len = args.captions, args.article_len, 4, args.max_length
    model = VacnicBart(cfg, device=dev, p_drop=0.0, seed=42)  # seed 42: README.md:8
    model.eval()
    n_batches = 2
    host = []
    for i in range(n_batches):
        b = synthetic.make_batch(B=C, L=L, T=8, seed=42 + 1000 * rank + i)
        face = b["face_emb"]
        host.append({k: v.pin_memory() for k, v in dict(
            input_ids=b["article_ids"], attention_mask=(b["article_ids"] != 1).to(torch.int64), image_features=b["image_features"],
            face_features=face, face_mask=(face[:, :, -1] != 1).to(torch.int64), name_ids=b["names_art_ids"],
            name_mask=(b["names_art_ids"] != 1).to(torch.int64)).items()})
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    def barrier():
        if world > 1 and collectives:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def gen(b):
        return generation.generate(model, num_beams=nb, max_length=max_len, length_penalty=2.0, **b)

    def run(batches, steps, from_host):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        d2h = 0
        for i in range(steps):
            b = batches[i % len(batches)]
            if from_host:
                b = {k: v.to(dev, non_blocking=True) for k, v in b.items()}
            ids = gen(b)
            if from_host:
                ids = ids.cpu()
                d2h = ids.numel() * ids.element_size()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1 and collectives:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, d2h

    for i in range(max(3, args.warmup)):
        gen(devb[i % n_batches])
    torch.cuda.synchronize()
    engine = next(iter(model._generators.values()))
    with ClockSampler(dev.index or 0) as clk:
        ms_dev, _ = run(devb, args.steps, False)
    steps_run = engine.steps_run
    ms_e2e, d2h = run(host, args.steps, True)
    value = C * world * args.steps / (ms_dev / 1e3)
    e2e = C * world * args.steps / (ms_e2e / 1e3)
    roof = None
    if rank == 0:
        try:
            roof = infer_roofline(args, model, engine, devb, cfg, C, L, nb, steps_run, pk)
        except Exception as e:  # noqa: BLE001 -- an optional profiling pass must never cost the measurement
            traceback.print_exc(file=sys.stderr)
            roof = {"error": f"{type(e).__name__}: {e}"[:300]}
    return {"value": value, "e2e": e2e, "ms_dev": ms_dev, "ms_e2e": ms_e2e, "h2d": h2d, "d2h": d2h, "clocks": clk.summary(),
            "roofline": roof, "steps_run": steps_run, "launches_per_step": engine.launches_per_step, "C": C,
            "workload": f"BART-large VACNIC beam-search inference (BASELINE.json configs[2]): beam 4, length_penalty 2.0, "
                        f"max_length {max_len}, seed 42, {C} captions/GPU, L={L} article tokens + P=20 prefix, bf16"}


# ----------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "infer", "train-vis"],
                    help="train = BASELINE configs[1] (headline); infer = configs[2]; train-vis = configs[4] (only-visual-prompt model, CE only)")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--article-len", type=int, default=1024)
    ap.add_argument("--caption-len", type=int, default=64)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-pipeline-opt", action="store_true", help="one fused AdamW launch after the gradient exchange instead of per-bucket updates")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "p2p-mc", "nccl", "none"],
                    help="multi-GPU gradient exchange (auto = peer-memory sharded optimizer when available); none = debugging: "
                         "independent replicas, NOT data-parallel training")
    ap.add_argument("--seed-offset", type=int, default=0, help="debug: shift the synthetic batch seeds")
    ap.add_argument("--no-balance", action="store_true", help="multi-GPU + packed rows: deal samples to ranks contiguously instead of length-balanced")
    ap.add_argument("--no-varlen", action="store_true",
                    help="train workloads: keep the collate's padded [B, L] article rows on the device (the reference's layout) "
                         "instead of packed rows")
    ap.add_argument("--no-roofline", action="store_true", help="skip the rank-0 eager roofline pass (quick A/B runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the PyTorch-eager comparator of the reference arithmetic on the GPU")
    ap.add_argument("--small", action="store_true", help="BART-base config-1 shapes (debug)")
    ap.add_argument("--captions", type=int, default=256, help="captions per GPU (infer workload)")
    ap.add_argument("--max-length", type=int, default=50)
    ap.add_argument("--no-infer", action="store_true", help="train workload: skip the secondary beam-4 inference measurement")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl != "reference" and args.gpus != world:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}: launch one rank per GPU with "
                         f"`python -m torch.distributed.run --nnodes=1 --nproc-per-node {args.gpus} --master-addr 127.0.0.1 "
                         f"--master-port P bench.py --gpus {args.gpus} ...`")
    vis = args.workload == "train-vis"
    if vis:  # run_onlyvis_train.sh: GoodNews shapes, batch 32 per GPU, 512 article tokens, CE only, no guide
        args.batch = 32 if args.batch == 16 else args.batch
        args.article_len = 512 if args.article_len == 1024 else args.article_len
        args.no_infer = True
    metric = "beam4_captions_per_sec" if args.workload == "infer" else "train_samples_per_sec"
    unit = "captions/s" if args.workload == "infer" else "samples/s"
    config = {"workload": "BART-large VACNIC full-model training step (BASELINE.json configs[1]): bf16 compute, "
                          f"batch {args.batch}/GPU, L={args.article_len} article tokens + P=20 ClipCap prefix, T={args.caption_len}, "
                          "CE + 0.5*CoLaM(margin 1.0, frozen BART-large guide) + SECLA(F=4,N=8), dropout 0.1, AdamW",
              "global_batch": args.batch * world, "parallelism": f"dp{world}",
              "l2": "working set per step (>=1.8 GB bf16 weights + activations) exceeds the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        steps = max(1, min(args.steps, 2))
        if args.workload == "infer":
            val, ms, sample = cpu_reference_infer(threads, args.max_length)
            steps = 1
        else:
            val, ms, sample = cpu_reference_train(threads, steps, min(args.warmup, 1))
        line = {"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus, "steps": steps,
                "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the VACNIC hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: the JSON line must stand alone
    if world > 1:
        # NCCL / c10d print a version banner on stdout when the first communicator is created: route fd 1 to stderr until
        # then, so that the JSON result line is the only thing this script writes to stdout
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            torch.distributed.init_process_group("nccl", device_id=dev)
            pg = torch.distributed.group.WORLD
            warm = torch.zeros(1, device=dev)
            torch.distributed.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    from vacnic_b200 import kernels as K
    from vacnic_b200 import spec, synthetic
    from vacnic_b200.modeling import VacnicBart
    from vacnic_b200.trainer import TrainStep

    if args.workload == "infer":
        r = run_infer(args, rank, world, dev, peaks())
        if rank == 0:
            cpu = None
            if world == 1 and not args.no_cpu_baseline:
                threads = os.cpu_count() or 1
                v, ms, sample = cpu_reference_infer(threads, args.max_length)
                cpu = {"value": v, "unit": unit, "cores": threads, "kind": "port", "sample": sample}
            n_launch = int((r["steps_run"] * r["launches_per_step"]) * args.steps)
            print(json.dumps({
                "metric": metric, "value": r["value"], "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": r["ms_dev"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": r["workload"], "global_batch": r["C"] * world, "parallelism": f"dp{world} (captions sharded, no communication)",
                           "l2": "per-step decode traffic (weights 0.5 GB + cross K/V 2-3 GB) exceeds the 126 MB L2; no explicit flush"},
                "clocks": r["clocks"],
                "e2e": {"value": r["e2e"], "unit": unit, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                        "ms_per_step": r["ms_e2e"] / args.steps},
                "gpu_launches": n_launch, "roofline": r["roofline"], "cpu_baseline": cpu}))
        finish(world)
        return

    if args.small:
        cfg = spec.bart_base(only_image=vis)
        gcfg = spec.bart_base(stock=True)
    else:
        cfg = spec.bart_large(only_image=vis)
        gcfg = spec.VacnicConfig(stock=True)
    B, L, T = args.batch, args.article_len, args.caption_len
    if vis:
        config["workload"] = (f"BART-large only-visual-prompt VACNIC training step (BASELINE.json configs[4], MVIS / TRAINVIS): bf16, "
                              f"batch {B}/GPU, L={L} article tokens + P=20 ClipCap prefix, T={T}, token CE only, dropout 0.1, AdamW")
        config["global_batch"] = B * world
    model = VacnicBart(cfg, device=dev, p_drop=0.1, seed=684331)          # seed of run_full_train.sh:2
    guide = None if vis else VacnicBart(gcfg, device=dev, p_drop=0.0, seed=7, frozen=True)
    ts = TrainStep(model, guide, lr=3e-5, weight_decay=0.01, warmup_steps=100, total_steps=100000, margin=1.0, alpha=0.5,
                   use_graph=not args.no_graph, process_group=None if args.exchange == "none" else pg,
                   pipeline_optimizer=False if args.no_pipeline_opt else None,
                   exchange=None if args.exchange in ("auto", "none") else args.exchange, varlen=not args.no_varlen)
    config["prefix_side_stream"] = ts.side_stream is not None
    config["guide_stream"] = ts.guide_stream is not None
    config["exchange"] = ts.exchange if (world > 1 and args.exchange != "none") else ("none" if world > 1 else "single GPU")

    n_batches = 4
    # every rank draws the SAME global batch (seeded by the step only) and takes its shard of it: length-balanced shards
    # (dp.balance_shards) when the article rows are packed, the DistributedSampler-style contiguous deal otherwise
    from vacnic_b200.dp import balance_shards
    host = []
    for i in range(n_batches):
        gb = synthetic.make_batch(B=B * world, L=L, T=T, seed=1000 + i + args.seed_offset)
        if world > 1 and not args.no_varlen and not args.no_balance:
            mine = balance_shards((gb["article_ids"] != 1).sum(1).tolist(), world, B)[rank]
        else:
            mine = list(range(rank * B, (rank + 1) * B))
        host.append(TrainStep.prepare({k: v[mine].contiguous() for k, v in gb.items()}, cfg, varlen=not args.no_varlen))
    config["sampler"] = ("length-balanced shards of the global batch (dp.balance_shards)" if (world > 1 and not args.no_varlen
                         and not args.no_balance) else "contiguous shards of the global batch (DistributedSampler-style)")
    art_tokens = [int((b["article_ids"] != 1).sum()) for b in host]
    config["article_rows"] = ("packed (varlen): the collate's padding never reaches the device; mean "
                              f"{sum(art_tokens) / len(art_tokens) / (B * L):.3f} of B*L rows are real tokens, rows rounded up to a multiple of 512"
                              if not args.no_varlen else "padded [B, L] (reference layout)")
    host = [{k: v.pin_memory() for k, v in b.items()} for b in host]
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def run(batches, steps, read_loss):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        last = None
        for i in range(steps):
            losses = ts.step(batches[i % len(batches)], prepared=True)
            if read_loss:
                last = float(losses["txt"].item())  # device -> host read of the step's result
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last

    # warm-up: at least W (>= 3) steps AND every distinct batch once, so that every packed-row bucket the timed region will
    # meet has its step graph captured before the clock starts
    for i in range(max(3, args.warmup, n_batches)):
        ts.step(devb[i % n_batches], prepared=True)
    torch.cuda.synchronize()
    c0 = K._l.launch_count()
    with ClockSampler(local_rank) as clk:
        ms_dev, _ = run(devb, args.steps, read_loss=False)
    eager_launches = K._l.launch_count() - c0
    launches = ts.launches_per_step * args.steps if ts.use_graph else eager_launches
    ms_e2e, last_loss = run(host, args.steps, read_loss=True)

    value = B * world * args.steps / (ms_dev / 1e3)
    e2e = B * world * args.steps / (ms_e2e / 1e3)
    pk = peaks()
    flops_sample, fwd_f, guide_f = train_flops_per_sample(cfg, L, T, with_guide=not vis)
    step_tflops = flops_sample * B / 1e12
    # the same model evaluated at every article's TRUE length (what packed rows execute): sum over the samples of a batch
    real_lens = [(b["article_ids"] != 1).sum(1).tolist() for b in host]
    step_tflops_unpadded = sum(sum(train_flops_per_sample(cfg, n, T, with_guide=not vis)[0] for n in lens) for lens in real_lens) / len(real_lens) / 1e12

    # THE MEASUREMENT IS COMPLETE HERE.  Everything below (roofline pass, CPU / eager comparators, the secondary inference
    # figure) is optional: each part runs inside try/except and the result line is printed in a `finally`, so a failing
    # profiling pass can null its own key but never the measurement (round 1 lost its N>1 lines to exactly that).
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": config, "clocks": clk.summary(),
            "e2e": {"value": e2e, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                    "last_txt_loss": last_loss},
            "gpu_launches": int(launches), "roofline": None, "cpu_baseline": None, "infer": None}

    def guarded(name, fn):
        try:
            return fn()
        except Exception as e:  # noqa: BLE001
            traceback.print_exc(file=sys.stderr)
            sys.stderr.write(f"bench.py: optional pass '{name}' failed on rank {rank}; the measurement is unaffected\n")
            return {"error": f"{type(e).__name__}: {e}"[:300]}

    try:
        # ---- roofline of the dominant kernel (gemm2_sm100_kernel): one eager, event-instrumented step on rank 0
        if rank == 0 and not args.no_roofline:
            line["roofline"] = guarded("roofline", lambda: train_roofline(model, guide, ts, devb, pk, step_tflops, ms_dev / args.steps,
                                                                         args.small, step_tflops_unpadded))
        if world > 1:
            torch.distributed.barrier()
        # ---- free the trainer: drop the graph, break the model <-> step cycle, collect, return the blocks to the driver
        ts.close()
        del ts, model, guide, devb, host
        free_cuda()

        if not args.no_infer:
            # secondary headline (BASELINE.json metric names both): beam-4 captions/s on the same GPUs.  Each rank times
            # its own captions (no collective inside); ONE all-reduce afterwards combines them, and every rank reaches it
            iargs = argparse.Namespace(**{**vars(args), "steps": min(args.steps, 3), "warmup": 3})
            r = guarded("infer", lambda: run_infer(iargs, rank, world, dev, pk, collectives=False))
            ok = "error" not in r
            free_cuda()
            if world > 1:
                t = torch.tensor([r["ms_dev"] if ok else 0.0, r["ms_e2e"] if ok else 0.0, 0.0 if ok else 1.0], device=dev)
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                ms_i, ms_ie, bad = (float(x) for x in t.tolist())
            else:
                ms_i, ms_ie, bad = (r["ms_dev"], r["ms_e2e"], 0.0) if ok else (0.0, 0.0, 1.0)
            if bad or not ok:
                line["infer"] = r if not ok else {"error": "another rank failed"}
            else:
                Ci = r["C"]
                line["infer"] = {"metric": "beam4_captions_per_sec", "value": Ci * world * iargs.steps / (ms_i / 1e3), "unit": "captions/s",
                                 "steps": iargs.steps, "ms_per_step": ms_i / iargs.steps,
                                 "e2e": {"value": Ci * world * iargs.steps / (ms_ie / 1e3), "unit": "captions/s",
                                         "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
                                 "config": {"workload": r["workload"]}, "roofline": r["roofline"],
                                 "gpu_launches": int(r["steps_run"] * r["launches_per_step"] * iargs.steps)}

        if rank == 0 and world == 1 and not vis:
            if not args.no_gpu_eager:
                line["gpu_eager_baseline"] = guarded("gpu_eager_baseline", lambda: gpu_eager_reference(dev, B, L, T))
                free_cuda()
            if not args.no_cpu_baseline:
                def cpu_leg():
                    threads = os.cpu_count() or 1
                    v, ms, sample = cpu_reference_train(threads, 1, 0, batch=B)
                    return {"value": v, "unit": unit, "cores": threads, "kind": "port", "sample": sample}
                line["cpu_baseline"] = guarded("cpu_baseline", cpu_leg)
    finally:
        if rank == 0:
            print(json.dumps(line))
            sys.stdout.flush()
    finish(world)


def free_cuda():
    gc.collect()
    torch.cuda.empty_cache()


def train_roofline(model, guide, ts, devb, pk, step_tflops, ms_step, small, step_tflops_unpadded=None):
    """Dominant-kernel roofline of the training step: one eager pass with CUDA events around every GEMM launch (on the
    launching stream), Sum 2MNK over the launches routed to the CTA-pair kernel / Sum of their durations."""
    from vacnic_b200 import kernels as K
    from vacnic_b200.trainer import TrainStep
    # single stream: the per-launch CUDA events must bracket one kernel, not one kernel plus whatever a second stream ran
    eager = TrainStep(model, guide, use_graph=False, process_group=None, varlen=ts.varlen, side_stream=False, guide_stream=False)
    eager.m, eager.v = ts.m, ts.v
    eager.step_dev.copy_(ts.step_dev)
    try:
        eager.step(devb[0], prepared=True)  # warm the eager path
        torch.cuda.synchronize()
        K.PROFILE = []
        eager.step(devb[1], prepared=True)
        torch.cuda.synchronize()
        prof = K.PROFILE
    finally:
        K.PROFILE = None
        eager.close()
    big = [p for p in prof if p[4]]  # launches routed to the dominant kernel: the CTA-pair tcgen05 GEMM
    gf_all = sum(p[0] for p in prof)
    gms_all = sum(p[1].elapsed_time(p[2]) for p in prof)
    gf = sum(p[0] for p in big)
    gms = sum(p[1].elapsed_time(p[2]) for p in big)
    ach = gf / 1e12 / (gms / 1e3) if gms > 0 else 0.0
    ach_all = gf_all / 1e12 / (gms_all / 1e3) if gms_all > 0 else 0.0
    peak = pk["bf16_tflops_sustained"]
    rep = None
    if not small:
        # ONE representative launch timed alone (FFN fc1 + bias + GELU, 16384x4096x1024), cold L2 (three rotating operand
        # sets, 3 x 176 MB > 126 MB): the launch `traffic` below was captured on (ncu --set full)
        M_, N_, K_ = 16384, 4096, 1024
        dev = devb[0]["article_ids"].device
        sets = [(torch.randn(M_, K_, device=dev, dtype=torch.bfloat16), torch.randn(N_, K_, device=dev, dtype=torch.bfloat16) * 0.02,
                 torch.zeros(N_, device=dev), torch.empty(M_, N_, device=dev, dtype=torch.bfloat16),
                 torch.empty(M_, N_, device=dev, dtype=torch.bfloat16)) for _ in range(3)]
        for x, w, bb, o, aux in sets:
            K.gemm(x, w, out=o, bias=bb, act=K.ACT_GELU, aux_out=aux)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_rep = 12
        e0.record()
        for i in range(n_rep):
            x, w, bb, o, aux = sets[i % 3]
            K.gemm(x, w, out=o, bias=bb, act=K.ACT_GELU, aux_out=aux)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n_rep * 1e3
        tf = 2.0 * M_ * N_ * K_ / 1e12 / (us / 1e6)
        rep = {"shape": "fc1+bias+GELU M=16384 N=4096 K=1024 (bf16 out + bf16 pre-activation)", "us": us, "achieved": tf,
               "frac": tf / pk["bf16_tflops"], "peak": pk["bf16_tflops"], "peak_kind": "burst (kernel timed alone)",
               "algorithmic_bytes": 2.0 * (M_ * K_ + N_ * K_ + 2 * M_ * N_), "traffic": 319.6e6}
        del sets
    return {"bound": "tensor", "kernel": "gemm2_sm100_kernel (CTA-pair tcgen05.mma.cta_group::2 / TMEM / TMA batched bf16 GEMM)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (the representative launch below), from the committed
            # ncu --set full capture -- a profile figure per launch, not measured in this run
            "traffic": 319.6e6 if not small else None,
            "traffic_source": "profiles/r2_ncu_targets.md: fc1+bias+GELU 16384x4096x1024, 89.9 MB read + 229.7 MB written "
                              "(310 MB algorithmic: A + W read, bf16 output + bf16 pre-activation written)" if not small else None,
            "representative_launch": rep, "peak_source": pk["_source"] + " sustained",
            "launches_per_step": len(big), "kernel_ms_per_step": gms, "kernel_tflop_per_step": gf / 1e12,
            "kernel_share_of_step": gms / ms_step,
            # every vacnic_gemm launch of the step (skinny side-branch / decoder GEMMs included; eager pass, so the
            # small launches carry host gaps -- tools/profile_graph.py has their in-graph times)
            "all_gemm": {"achieved": ach_all, "frac": ach_all / peak, "launches_per_step": len(prof), "ms_per_step": gms_all},
            # SURVEY 8(d) counts the PADDED shapes (L tokens per article); with packed rows the device executes the unpadded
            # figure, so both are reported: step_mfu_unpadded is the honest utilisation, step_mfu the padded-equivalent rate
            "algorithmic_tflop_per_step": step_tflops, "step_mfu": step_tflops / (ms_step / 1e3) / peak,
            "algorithmic_tflop_per_step_unpadded": step_tflops_unpadded,
            "step_mfu_unpadded": None if step_tflops_unpadded is None else step_tflops_unpadded / (ms_step / 1e3) / peak}


def gpu_eager_reference(dev, B, L, T, steps=2):
    """SURVEY.md 8(d) "GPU-side comparator": the reference's arithmetic (the oracle port of its PyTorch modules and loss
    block, pinned to the reference classes) run by PyTorch eager ON THE SAME B200 -- fp32 and torch.autocast(bfloat16),
    batch B, torch.optim.AdamW (TRAIN:95).  A reported baseline: how far the sm_100a path is ahead of stock PyTorch."""
    from oracle import model as OM
    from vacnic_b200 import spec, synthetic
    cfg = spec.bart_large()
    gcfg = spec.VacnicConfig(stock=True)
    sd = {k: v.to(dev) for k, v in spec.test_state_dict(cfg, 1).items()}
    for k in sd:
        if sd[k].is_floating_point() and k != "final_logits_bias" and k not in spec.TIED_TO_SHARED:
            sd[k].requires_grad_(True)
    for k in spec.TIED_TO_SHARED:
        sd[k] = sd["model.shared.weight"]
    gsd = {k: v.to(dev) for k, v in spec.test_state_dict(gcfg, 2).items()}
    params = [v for k, v in sd.items() if v.requires_grad and k not in spec.TIED_TO_SHARED]
    opt = torch.optim.AdamW(params, lr=3e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    batches = [synthetic.to_device(synthetic.make_batch(B=B, L=L, T=T, seed=500 + i), dev) for i in range(2)]
    out = {"kind": "port (oracle/model.py) in PyTorch eager on the same GPU", "batch": B, "dropout": 0.0, "unit": "samples/s"}
    for mode in ("fp32", "autocast_bf16"):
        times = []
        for i in range(1 + steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode != "fp32")):
                res = OM.training_losses(sd, cfg.as_dict(), gsd, gcfg.as_dict(), batches[i % 2], margin=1.0, alpha=0.5)
            res["loss"].backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            e1.record()
            torch.cuda.synchronize()
            if i > 0:
                times.append(e0.elapsed_time(e1))
        ms = sum(times) / len(times)
        out[mode] = {"value": B / (ms / 1e3), "ms_per_step": ms, "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30}
    return out


def finish(world):
    """Multi-rank runs: leave together and skip interpreter teardown (destroying NCCL communicators while captured
    graphs that contain collectives are being garbage-collected was observed to hang a 4-GPU run AFTER its result line)."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        torch.cuda.synchronize()
        torch.distributed.barrier()
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    try:
        main()
    except SystemExit:
        raise
    except BaseException:  # noqa: BLE001 -- leave the root cause on stderr of the failing rank (torchrun only reports the code)
        sys.stderr.write(f"bench.py: rank {os.environ.get('RANK', '0')} failed:\n")
        traceback.print_exc(file=sys.stderr)
        sys.stderr.flush()
        os._exit(1)
