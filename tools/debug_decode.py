import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_generation_gpu import _build, _gen_kwargs
from vacnic_b200 import generation, kernels as K
from oracle import model as OM
dev = torch.device("cuda:0")
fx, cfg, m, batch = _build("tests/golden/base_full_mini.pt", dev)
kw = _gen_kwargs(cfg, batch)
gold = fx["greedy_ids"].to(dev)
print("gold", gold.tolist())
g = generation.generate(m, num_beams=1, max_length=fx["max_length"], use_graph=False, **kw)
print("got ", g.tolist())
# teacher-forced full forward of our model on the golden ids: logits argmax per position
with torch.no_grad():
    out = m(decoder_input_ids=gold[:, :-1].contiguous(), **kw)
lg = out["logits"].float()
print("tf argmax", lg.argmax(-1).tolist())
top2 = lg.topk(2, -1).values
print("tf margins", (top2[..., 0] - top2[..., 1]).min().item())
# step-by-step: cached decode logits vs teacher-forced logits
gen = generation.Generator(m, gold.shape[0], 1, kw["input_ids"].shape[1], fx["max_length"], use_graph=False)
gen.encode(generation._enc_inputs(m, kw["input_ids"], kw["attention_mask"], kw["image_features"], kw.get("face_features"), kw.get("face_mask"), kw.get("name_ids"), kw.get("name_mask")))
gen._reset_state()
for t in range(1, 6):
    gen.st["seq"][:, :t] = gold[:, :t].int()
    gen._step()
    torch.cuda.synchronize()
    V = cfg.vocab
    err = (gen.logits[:, :V] - lg[:, t - 1]).abs().max().item()
    print("step", t, "cur_len", gen.st["cur_len"].item(), "max|dlogit|", err, "argmax", gen.logits[:, :V].argmax(-1).tolist(), "top_idx", gen.top_idx[:, 0].tolist())
