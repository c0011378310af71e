"""Per-kernel time breakdown of one fine-tuning step (BASELINE configs[3]) with torch.profiler (CUPTI): cheap enough to run on
the whole step (ncu serialises ~0.1 s per launch).  Prints the kernels by total device time.

    python scripts/profile_finetune.py [batch]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_sequencing_b200 import OrderingEngine  # noqa: E402
from oracle import berson_oracle as O  # noqa: E402  (synthetic inputs only)
from oracle import synth  # noqa: E402

torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = dict(synth.BERT_BASE)
vit = dict(synth.VIT_B32)
cfg.update(vit=vit, rn=None, para_ff=3072)
eng = OrderingEngine(synth.full_state_dict(cfg, vit, seed=0), cfg, precise=False)
ids, labels, images = O.synthetic_manuals(B, 6, 64, image_px=224, seed=1)
pb = eng.prepare(ids, labels, 6, images).to(eng.device)
grads = eng.new_grad_buffer()
if os.environ.get("MSQ_PROFILE_DROPOUT", "1") == "1":
    eng.set_dropout(0.1, 0.1, 0.1, seed=1234)   # the reference's fine-tuning workload (bench.py --config 3)


def step():
    grads.zero_()
    eng.train_step(pb, grads)
    eng.adamw_step(grads, 5e-6)


for _ in range(2):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print("total device time %.2f ms over %d launches" % (tot / 1e3, sum(e.count for e in rows)))
for e in rows[:28]:
    print("%9.2f ms %5.1f%%  n=%5d  avg=%8.1f us  %s" % (e.device_time_total / 1e3, 100 * e.device_time_total / tot, e.count,
                                                        e.device_time_total / e.count, e.key[:100]))
