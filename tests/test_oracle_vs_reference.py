"""Live pin of the oracle against the REAL reference when /root/reference is present (build container).
Skipped on the GPU box, where only the committed fixtures exist.  Runs in a subprocess: the reference's
top-level `models` package must not collide with the drop-in mirror other tests import."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_harness as rh  # noqa: E402

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference checkout not present")

SCRIPT = r'''
import json, sys, torch
sys.path.insert(0, %(golden)r); sys.path.insert(0, %(root)r)
import ref_harness as rh
from oracle import berson_oracle as O
torch.set_grad_enabled(False)
TINY = dict(vocab_size_or_config_json_file=600, hidden_size=128, num_hidden_layers=1, num_attention_heads=2,
            intermediate_size=256, max_position_embeddings=128)
VIT = dict(embed_dim=64, image_resolution=224, vision_layers=1, vision_width=128, vision_patch_size=32)
mm, N, W = %(mm)r, %(N)d, %(W)d
ns = rh.load()
args = rh.make_args(N, W, multimodal=mm)
args.ff_size = 128
model = rh.build_multimodal_model(ns, TINY, args, VIT, seed=5) if mm else rh.build_text_model(ns, TINY, args, seed=5)
sd = {k: v.clone() for k, v in model.state_dict().items()}
cfg = dict(num_hidden_layers=1, num_attention_heads=2, vit=VIT if mm else None)
ids, labels, images = O.synthetic_manuals(2, N, 12, vocab=600, image_px=224 if mm else None, seed=77)
ref = []
for b in range(2):
    inputs = {"input_ids": ids[b:b + 1], "attention_mask": torch.ones_like(ids[b:b + 1]), "labels": labels[b:b + 1]}
    if mm:
        inputs["images"] = images[b:b + 1]
    ref.append(ns.berson.berson_pointer_network(args, model, rh.StubTokenizer(), inputs))
print("RESULT " + json.dumps({"ref": ref, "oracle": O.order_manuals(sd, cfg, ids, labels, N, W, images)}))
'''


@pytest.mark.parametrize("mm,N,W", [(False, 5, 4), (False, 7, 16), (True, 5, 4)])
def test_live(mm, N, W):
    code = SCRIPT % dict(golden=os.path.join(HERE, "golden"), root=os.path.dirname(HERE), mm=mm, N=N, W=W)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert res["oracle"] == res["ref"]


TC_SCRIPT = r'''
import json, sys, numpy as np, torch
sys.path.insert(0, %(golden)r); sys.path.insert(0, %(root)r)
import ref_harness as rh
from oracle import berson_oracle as O
torch.set_grad_enabled(False)
TINY = dict(vocab_size_or_config_json_file=600, hidden_size=128, num_hidden_layers=1, num_attention_heads=2,
            intermediate_size=256, max_position_embeddings=128)
N = 6
ns = rh.load()
args = rh.make_args(N, 4)
args.ff_size = 128
args.additional_wrapper_level_objectives = ["time_contrastive"]
model = rh.build_text_model(ns, TINY, args, seed=9)
model.tokenizer = rh.StubTokenizer()
assert model.time_contrastive
sd = {k: v.clone() for k, v in model.state_dict().items()}
cfg = dict(num_hidden_layers=1, num_attention_heads=2, vit=None)
ids, labels, _ = O.synthetic_manuals(3, N, 12, vocab=600, seed=21)
inputs = {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": labels}
np.random.seed(1234)
ref = float(model(inputs)[0])
inp = O.prepare_inputs(ids, labels, N, None)
np.random.seed(1234)
trip = O.time_contrastive_triplets(inp["ground_truth"])
print("RESULT " + json.dumps({"ref": ref, "oracle": float(O.training_loss(sd, cfg, inp, triplets=trip)),
                              "plain": float(O.training_loss(sd, cfg, inp))}))
'''


def test_time_contrastive_objective_live():
    """modeling_bert.py:1176-1216 on the real reference (numpy seeded) against the oracle's restatement of the index draw
    and of the triplet term."""
    code = TC_SCRIPT % dict(golden=os.path.join(HERE, "golden"), root=os.path.dirname(HERE))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert abs(res["oracle"] - res["ref"]) < 2e-5, res
    assert abs(res["oracle"] - res["plain"]) > 1e-3, res      # the term was active


ML_SCRIPT = r'''
import json, sys, torch
sys.path.insert(0, %(golden)r); sys.path.insert(0, %(root)r)
import ref_harness as rh
from oracle import berson_oracle as O
torch.set_grad_enabled(False)
TINY = dict(vocab_size_or_config_json_file=600, hidden_size=128, num_hidden_layers=1, num_attention_heads=2,
            intermediate_size=256, max_position_embeddings=128)
VIT = dict(embed_dim=64, image_resolution=224, vision_layers=1, vision_width=128, vision_patch_size=32)
N = 5
ns = rh.load()
args = rh.make_args(N, 4, multimodal=True)
args.ff_size = 128
args.multimodal_loss = True
model = rh.build_multimodal_model(ns, TINY, args, VIT, seed=11, v_feature_size=128)
assert hasattr(model, "img_projection")
sd = {k: v.clone() for k, v in model.state_dict().items()}
cfg = dict(num_hidden_layers=1, num_attention_heads=2, vit=VIT)
ids, labels, images = O.synthetic_manuals(2, N, 12, vocab=600, image_px=224, seed=31)
inputs = {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": labels, "images": images}
ref = float(model(inputs)[0])
perm = ns.berson.berson_pointer_network(args, model, rh.StubTokenizer(),
                                        {k: v[:1] for k, v in inputs.items()})
inp = O.prepare_inputs(ids, labels, N, images)
print("RESULT " + json.dumps({"ref": ref, "oracle": float(O.training_loss(sd, cfg, inp, multimodal_loss=True)),
                              "plain": float(O.training_loss(sd, cfg, inp)), "perm": perm,
                              "oracle_perm": O.order_manuals(sd, cfg, ids[:1], labels[:1], N, 4, images[:1])[0]}))
'''


def test_multimodal_loss_objective_live():
    """args.multimodal_loss (modeling_bert.py:897-898, 1359-1364, 1218-1225) on the real reference against the oracle's
    restatement of the image pairwise term; decoding with the flag set (1432-1433) is unchanged."""
    code = ML_SCRIPT % dict(golden=os.path.join(HERE, "golden"), root=os.path.dirname(HERE))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert abs(res["oracle"] - res["ref"]) < 2e-5, res
    assert abs(res["oracle"] - res["plain"]) > 1e-2, res      # the term was active
    assert res["perm"] == res["oracle_perm"], res
