"""The UNCHANGED multi-GPU training loop of the reference: the script wraps the model in DistributedDataParallel
(TRAIN:86-87, and with find_unused_parameters=True at TRAINVIS:84), builds torch.optim.AdamW over `model.module.*`
parameters (TRAIN:91) and calls forward / loss.backward() / optimizer.step().  The kernels write parameter gradients
straight into the flat buffer, so `blocks.ParamTouchFn` is what makes DDP's reducer hooks fire; this test runs that loop
on two ranks and checks that (1) step 2 does not raise, (2) every rank ends with the SAME parameters, (3) the gradient
DDP leaves in `p.grad` is the average of the two ranks' local gradients.

Two processes share the one visible GPU through the gloo backend (NCCL refuses two ranks on one device); with two or
more GPUs each rank takes its own device and NCCL is used."""
import importlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MFULL = "src.models.modeling_mmbart_clip_inside_vis_clipcap_ent_type_final_fix_len_enc_self_face_name_ids_crossattn"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, find_unused, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    n_gpu = torch.cuda.device_count()
    dev_index = rank if n_gpu >= world else 0
    torch.cuda.set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    dist.init_process_group("nccl" if n_gpu >= world else "gloo", rank=rank, world_size=world)
    from test_dropin_gpu import _ctor_kwargs, _hf_config, _inputs
    from vacnic_b200 import spec, synthetic
    mod = importlib.import_module(MFULL)
    cfg = spec.VacnicConfig(d_model=768, heads=12, ffn=1024, enc_layers=2, dec_layers=2, prompt_size=4, max_pos=128)
    model = mod.BartForMultiModalGeneration(_hf_config(cfg), **_ctor_kwargs(cfg), seed=100 + rank)  # ranks start DIFFERENT
    if rank == 0:
        model.load_state_dict(spec.test_state_dict(cfg, 31), strict=False)
    ddp = torch.nn.parallel.DistributedDataParallel(model.cuda(), device_ids=[dev_index], output_device=dev_index,
                                                    find_unused_parameters=find_unused)   # TRAIN:87 / TRAINVIS:84
    ddp.to(dev)                                                                           # TRAIN:88
    params = list(ddp.module.model.parameters()) + list(ddp.module.lm_head.parameters())   # TRAIN:91
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0.01)
    ce = torch.nn.CrossEntropyLoss(ignore_index=1)
    ddp.train()
    result = {"losses": []}
    for it in range(3):
        batch = synthetic.to_device(synthetic.make_batch(B=2, L=40, T=12, seed=50 + 10 * it + rank), dev)  # DistributedSampler
        tgt = batch["caption_ids"]
        dec_in = mod.shift_tokens_right(tgt, 1, 2)
        if it == 2:   # local gradient of THIS rank's batch through the bare module (no DDP hooks involved)
            with ddp.no_sync():
                out = ddp(decoder_input_ids=dec_in, **_inputs(cfg, batch))
                ce(out["logits"].reshape(-1, out["logits"].shape[-1]), tgt.reshape(-1)).backward()
            result["local_grad"] = model.store.grad.clone().cpu()
            opt.zero_grad()
        out = ddp(decoder_input_ids=dec_in, **_inputs(cfg, batch))
        logits = out["logits"]
        loss = ce(logits.reshape(-1, logits.shape[-1]), tgt.reshape(-1))
        loss.backward()
        if it == 2:
            result["ddp_grad"] = model.store.grad.clone().cpu()
        opt.step()
        opt.zero_grad()
        result["losses"].append(float(loss.detach()))
    torch.cuda.synchronize()
    result["master"] = model.store.master.clone().cpu()
    torch.save(result, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("find_unused", [False, True])
def test_unchanged_ddp_loop_two_ranks(cuda_device, tmp_path, find_unused):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), find_unused, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"rank{r}.pt") for r in range(world))
    # (2) replicas identical after three optimizer steps (DDP broadcast rank 0's weights, then averaged every gradient)
    assert torch.equal(r0["master"], r1["master"])
    # (3) the gradient DDP left in p.grad is the mean of the ranks' local gradients
    want = 0.5 * (r0["local_grad"] + r1["local_grad"])
    for r in (r0, r1):
        err = (r["ddp_grad"] - want).abs().max().item()
        assert err <= 1e-5 * max(1.0, want.abs().max().item()) + 1e-6, err
    assert (r0["local_grad"] - r1["local_grad"]).abs().max().item() > 1e-4   # the ranks really saw different data
    assert r0["losses"][2] < r0["losses"][0]
