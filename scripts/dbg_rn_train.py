"""Per-parameter gradient errors of the ModifiedResNet fine-tuning step vs oracle autograd (development aid)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_sequencing_b200 import OrderingEngine
from oracle import berson_oracle as O, train_oracle as TO
torch.set_grad_enabled(False)
precise = sys.argv[1] == "fp32"
g = torch.load(os.path.join(ROOT, "tests/golden/mm_rn_tiny.pt"), weights_only=False)
r = torch.load(os.path.join(ROOT, "tests/golden/grads_tiny.pt"), weights_only=False)["mm_rn_bntrain"]
c = g["cfg"]
cfg = dict(hidden_size=c["hidden_size"], num_hidden_layers=c["num_hidden_layers"], num_attention_heads=c["num_attention_heads"],
           intermediate_size=c["intermediate_size"], vocab_size=c["vocab_size_or_config_json_file"],
           max_position_embeddings=c["max_position_embeddings"], vit=None, rn=g["rn"], para_ff=g["ff_size"])
eng = OrderingEngine(g["sd"], cfg, precise=precise)
ids, labels, images = O.synthetic_manuals(r["B"], r["N"], r["L"], vocab=1000, image_px=224, seed=r["seed"])
pb = eng.prepare(ids, labels, r["N"], images)
grads = eng.new_grad_buffer()
loss = float(eng.train_step(pb, grads))
ocfg = dict(num_hidden_layers=c["num_hidden_layers"], num_attention_heads=c["num_attention_heads"], vit=None, rn=dict(g["rn"], bn_train=True))
oloss, ref = TO.loss_grads(g["sd"], ocfg, O.prepare_inputs(ids, labels, r["N"], images))
got = eng.grads_by_name(grads)
print("loss", loss, oloss)
for n in ref:
    if "visual" not in n or n not in got:
        continue
    a, b = got[n].float().cpu().reshape(-1), ref[n].float().reshape(-1)
    print("%-75s %.3e  |ref| %.3e" % (n[len("bert.encoder.visual_model.visual."):], float((a - b).norm() / (b.norm() + 1e-12)), float(b.norm())))
