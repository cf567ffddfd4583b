"""Four eager OF-3B training steps (target for `ncu --metrics gpu__time_duration.sum --profile-from-start off`: the launch list of the last one)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from open_flamingo_b200.testing import build_flamingo, synthetic_batch
from open_flamingo_b200.train import FlatTrainer

vit_cfg, mpt_kw, every = bench.model_dims("of3b")
model, _, tok = build_flamingo(vit_cfg, mpt_kw, cross_attn_every_n_layers=every, device="cuda", gate_init=1.0)
model.train()
media_id, eoc_id = tok.encode("<image>")[-1], tok.encode("<|endofchunk|>")[-1]
trainer = FlatTrainer(model)
batch = {k: v.cuda() for k, v in synthetic_batch(32, 2, 256, media_id, eoc_id, mpt_kw["vocab_size"], image_size=224).items()}
for it in range(4):
    if it == 3:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()      # ncu --profile-from-start off: only this step is profiled
    trainer.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(vision_x=batch["vision_x"], lang_x=batch["lang_x"], attention_mask=batch["attention_mask"], labels=batch["labels"])
    out.loss.backward()
    trainer.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
