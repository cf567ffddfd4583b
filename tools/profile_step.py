"""One-step kernel-time breakdown of the bench workload with torch.profiler (CUPTI).  Not a benchmark: numbers
under a profiler are never reported as bench values; this only shows where the step time goes."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from open_flamingo_b200.testing import build_flamingo, synthetic_batch
from open_flamingo_b200.train import FlatTrainer

model_name = sys.argv[1] if len(sys.argv) > 1 else "of3b"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
vit_cfg, mpt_kw, every = bench.model_dims(model_name)
dev = torch.device("cuda", 0)
model, _, tok = build_flamingo(vit_cfg, mpt_kw, cross_attn_every_n_layers=every, device=dev, gate_init=1.0)
model.train()
media_id, eoc_id = tok.encode("<image>")[-1], tok.encode("<|endofchunk|>")[-1]
trainer = FlatTrainer(model)
batch = {k: v.to(dev) for k, v in synthetic_batch(B, 2, 256, media_id, eoc_id, mpt_kw["vocab_size"],
                                                  image_size=vit_cfg["image_size"]).items()}


def step():
    trainer.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(vision_x=batch["vision_x"], lang_x=batch["lang_x"], attention_mask=batch["attention_mask"],
                    labels=batch["labels"])
    out.loss.backward()
    trainer.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name
        short = name.split("<")[0].split("(")[0][:70]
        tot[short][0] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        tot[short][1] += 1
total = sum(v[0] for v in tot.values())
print(f"total device kernel time {total/1e3:.2f} ms over {sum(v[1] for v in tot.values())} kernels")
mine = sum(v[0] for k, v in tot.items() if "ofk::" in k)
print(f"libofk kernels: {mine/1e3:.2f} ms ({100*mine/total:.1f}%)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{v[0]/1e3:9.3f} ms {100*v[0]/total:5.1f}% x{v[1]:5d}  {k}")
