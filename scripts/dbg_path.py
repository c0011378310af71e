import sys, torch
sys.path.insert(0, ".")
from multimodal_sequencing_b200 import OrderingEngine
from oracle import synth, berson_oracle as O
torch.set_grad_enabled(False)
cfg = dict(synth.BERT_BASE); cfg.update(vit=dict(synth.VIT_B32), rn=None, para_ff=3072)
sd = synth.full_state_dict(cfg, cfg["vit"], seed=0)
eng = OrderingEngine(sd, cfg, precise="bf16x3")
B = int(sys.argv[1])
ids, labels, images = O.synthetic_manuals(B, 5, 64, image_px=224, seed=1)
pb = eng.prepare(ids, labels, 5, images).to(eng.device)
for i in range(3):
    perm = eng.order_device(pb, 4)
    torch.cuda.synchronize()
    print("iter", i, perm[:2].tolist())
