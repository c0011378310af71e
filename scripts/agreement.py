"""bf16 tensor-core path vs fp32 parity path on the full-size model: permutation agreement and metric deltas
over a larger sample than the oracle-checked tests can afford (the fp32 path itself is checked against the
oracle / reference fixtures in tests/)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_sequencing_b200 import OrderingEngine  # noqa: E402
from oracle import berson_oracle as O  # noqa: E402
from oracle import synth  # noqa: E402

torch.set_grad_enabled(False)
out = []
for N, W, B in ((5, 4, 96), (6, 8, 32), (10, 16, 16)):
    cfg = dict(synth.BERT_BASE)
    cfg.update(vit=dict(synth.VIT_B32), para_ff=3072)
    sd = synth.full_state_dict(cfg, cfg["vit"], seed=0)
    ids, labels, images = O.synthetic_manuals(B, N, 64, image_px=224, seed=100 + N)
    res = {}
    for precise in (True, False):
        eng = OrderingEngine(sd, cfg, precise=precise)
        res[precise] = eng.order(ids, labels, N, W, images)
        del eng
        torch.cuda.empty_cache()
    agree = sum(a == b for a, b in zip(res[True], res[False]))
    m32, m16 = O.cal_result(labels.tolist(), res[True]), O.cal_result(labels.tolist(), res[False])
    out.append(dict(N=N, beam=W, manuals=B, identical_permutations=agree, rate=agree / B,
                    fp32_acc_pmr_tau=m32, bf16_acc_pmr_tau=m16))
    print(json.dumps(out[-1]), flush=True)
