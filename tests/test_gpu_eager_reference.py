""""also run the reference on the same B200 through PyTorch eager as the honest GPU baseline".  The oracle port
of the reference path (plain torch ops, one manual per call as berson_pointer_network does) is moved to the GPU with
torch's default-device context and timed next to this library on the same workload (BASELINE configs[1] shape).  A report,
not a parity gate: it only asserts that both sides return permutations."""
import os
import time

import pytest
import torch

from oracle import berson_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


def test_report_eager_gpu_reference_speed():
    from multimodal_sequencing_b200 import OrderingEngine
    cfg = dict(synth.BERT_BASE)
    vit = dict(synth.VIT_B32)
    cfg.update(vit=vit, rn=None, para_ff=3072)
    sd = synth.full_state_dict(cfg, vit, seed=0)
    N, W, B = 5, 4, 6
    ids, labels, images = O.synthetic_manuals(B, N, 64, image_px=224, seed=1)
    ocfg = dict(num_hidden_layers=12, num_attention_heads=12, vit=vit)
    eager = None
    try:
        with torch.device("cuda"):
            sd_gpu = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in sd.items()}
            times, perms = [], []
            for b in range(B):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                perms += O.order_manuals(sd_gpu, ocfg, ids[b:b + 1].cuda(), labels[b:b + 1].cuda(), N, W, images[b:b + 1].cuda())
                torch.cuda.synchronize()
                times.append(time.perf_counter() - t0)
        assert all(sorted(p) == list(range(N)) for p in perms)
        eager = (len(times) - 1) / sum(times[1:])
        del sd_gpu
    except Exception as e:   # the oracle is CPU test infrastructure; if a device mix-up stops it, say so and go on
        print("eager-GPU run of the oracle port failed: %r" % (e,))
    eng = OrderingEngine(sd, cfg, precise=False)
    pb = eng.prepare(ids, labels, N, images).to(eng.device)
    for _ in range(2):
        ours = eng.order_device(pb, W)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        ours = eng.order_device(pb, W)
    torch.cuda.synchronize()
    batched = 5 * B / (time.perf_counter() - t0)
    one = eng.prepare(ids[:1], labels[:1], N, images[:1]).to(eng.device)
    eng.order_device(one, W)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        eng.order_device(one, W)
    torch.cuda.synchronize()
    single = 10 / (time.perf_counter() - t0)
    assert all(sorted(p) == list(range(N)) for p in ours.cpu().tolist())
    print("torch-eager oracle port on the B200: %s manuals/s (one manual per call); this library: %.1f manuals/s one manual per call, "
          "%.1f manuals/s at batch %d" % ("%.2f" % eager if eager else "n/a", single, batched, B))
