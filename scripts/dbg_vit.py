import sys, torch
sys.path.insert(0, ".")
from multimodal_sequencing_b200 import OrderingEngine
from oracle import synth
torch.set_grad_enabled(False)
cfg = dict(synth.BERT_BASE); cfg.update(vit=dict(synth.VIT_B32), rn=None, para_ff=3072, num_hidden_layers=1)
vit = dict(synth.VIT_B32); vit["vision_layers"] = int(sys.argv[2]); cfg["vit"] = vit
sd = synth.full_state_dict(cfg, vit, seed=0)
eng = OrderingEngine(sd, cfg, precise="bf16x3")
R = int(sys.argv[1])
images = torch.randn(8, 3, 224, 224)
idx = torch.randint(0, 8, (R, 2), dtype=torch.int32)
for i in range(3):
    out = eng.vit_forward(images, idx, R)
    torch.cuda.synchronize()
    print("iter", i, float(out.abs().mean()))
