"""Where the decode loop's time goes per position (bench shape: 256 captions x 4 beams, BART-large, L = 1024): CUDA events
between the replays of the step graph over all 49 positions, the loop of Generator.decode() itself, and the same loop with
the stop-flag polling removed."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vacnic_b200 import generation, spec, synthetic  # noqa: E402
from vacnic_b200.modeling import VacnicBart  # noqa: E402

dev = torch.device("cuda:0")
cfg = spec.bart_large()
model = VacnicBart(cfg, device=dev, p_drop=0.0, seed=42)
model.eval()
C = 256
b = synthetic.make_batch(B=C, L=1024, T=8, seed=42)
face = b["face_emb"].to(dev)
kw = dict(input_ids=b["article_ids"].to(dev), attention_mask=(b["article_ids"] != 1).to(torch.int64).to(dev),
          image_features=b["image_features"].to(dev), face_features=face, face_mask=(face[:, :, -1] != 1).to(torch.int64),
          name_ids=b["names_art_ids"].to(dev), name_mask=(b["names_art_ids"] != 1).to(torch.int64).to(dev))
for _ in range(2):
    generation.generate(model, num_beams=4, max_length=50, length_penalty=2.0, **kw)
torch.cuda.synchronize()
eng = next(iter(model._generators.values()))
for rep in range(2):
    eng._reset_state()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(50)]
    evs[0].record()
    for t in range(49):
        eng.graph.replay()
        evs[t + 1].record()
    torch.cuda.synchronize()
ms = [evs[t].elapsed_time(evs[t + 1]) for t in range(49)]
print("per-position ms (events between replays):", " ".join(f"{x:.2f}" for x in ms))
print(f"sum {sum(ms):.1f} ms, mean {sum(ms) / 49:.3f} ms/step, first 8 mean {sum(ms[:8]) / 8:.3f}, last 8 mean {sum(ms[-8:]) / 8:.3f}")
for rep in range(2):
    eng._reset_state()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for t in range(49):
        eng.graph.replay()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
print(f"49 replays back to back, no events / polling: {e0.elapsed_time(e1):.1f} ms GPU, host enqueue {1e3 * (t1 - t0):.1f} ms")
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    eng.decode()
    e1.record()
    torch.cuda.synchronize()
print(f"Generator.decode(): {e0.elapsed_time(e1):.1f} ms for {eng.steps_run} steps")
