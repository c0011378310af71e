"""Gradient fixtures for the fine-tuning path, produced by the REAL reference (/root/reference) on CPU:

    python tests/golden/make_golden_grads.py          # writes tests/golden/grads_tiny.pt

BertForOrdering._forward (modeling_bert.py:943-1174) -> loss.backward() through the unmodified reference modules in
eval() mode (dropout off == the p = 0 semantics of the parity runs), for
  text  : the tiny text-only model / batch of text_tiny.pt's loss_case (3 five-step manuals)
  mm    : the tiny LXRT + CLIP-ViT model of mm_tiny.pt, one five-step manual (seed 45, 16 tokens per step)
  mm_rn / mm_rn_bntrain : the tiny LXRT + ModifiedResNet ("RN50" wiring) model of mm_rn_tiny.pt, one five-step manual (seed 65);
          first with the tower's BatchNorm in eval mode (running statistics), then with the WHOLE model in train() mode and every
          dropout probability set to 0 -- the mode the reference actually fine-tunes in (BatchNorm uses the statistics of the
          batch of materialised pair images).  Pins the oracle for the next scope row (backward through the RN50 tower).
Per parameter the fixture keeps a compact summary (numel, float64 sum, float64 L2 norm, 32 strided samples): enough
to pin the oracle's autograd (tests/test_oracle_golden.py) without committing two more copies of the weights.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_harness as rh  # noqa: E402
import make_golden as mg  # noqa: E402
from oracle import berson_oracle as O  # noqa: E402

torch.set_num_threads(1)


def summarize(g):
    g = g.detach().double().reshape(-1)
    idx = torch.linspace(0, g.numel() - 1, min(32, g.numel())).long()
    return dict(numel=g.numel(), sum=float(g.sum()), norm=float(g.norm()), idx=idx, val=g[idx].float())


def grads_of(ns, model, args, ids, labels, images):
    inputs = {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": labels}
    if images is not None:
        inputs["images"] = images
    bi = ns.prep.prepare_berson_inputs(inputs, rh.StubTokenizer(), args=args)
    model.zero_grad()
    with torch.enable_grad():
        loss = model._forward(**bi)[0]
        loss.backward()
    out = {}
    for n, p in model.named_parameters():
        if p.grad is not None and not any(d in n for d in mg.DEAD):
            out[n] = summarize(p.grad)
    return float(loss), out


def main():
    ns = rh.load()
    res = {}
    args = rh.make_args(5, 4)
    args.ff_size = 256
    model = rh.build_text_model(ns, mg.TINY, args, seed=0)
    ids, labels, _ = O.synthetic_manuals(3, 5, 24, vocab=1000, seed=33)
    loss, g = grads_of(ns, model, args, ids, labels, None)
    res["text"] = dict(loss=loss, grads=g, seed=33, B=3, N=5, L=24)
    args = rh.make_args(5, 4, multimodal=True)
    args.ff_size = 256
    model = rh.build_multimodal_model(ns, mg.TINY, args, mg.TINY_VIT, seed=0)
    ids, labels, images = O.synthetic_manuals(1, 5, 16, vocab=1000, image_px=224, seed=45)
    loss, g = grads_of(ns, model, args, ids, labels, images)
    res["mm"] = dict(loss=loss, grads=g, seed=45, B=1, N=5, L=16, image_checksum=float(images.double().sum()))
    # RN50 wiring: eval-mode BatchNorm, then train() with all dropouts at 0
    args = rh.make_args(5, 4, multimodal=True)
    args.ff_size = 256
    args.para_dropout = 0.0
    tiny0 = dict(mg.TINY, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    ids, labels, images = O.synthetic_manuals(1, 5, 16, vocab=1000, image_px=224, seed=65)
    for key, train in (("mm_rn", False), ("mm_rn_bntrain", True)):
        model = rh.build_multimodal_model(ns, tiny0, args, seed=0, rn_cfg=mg.TINY_RN)
        if train:
            model.train()
            assert all(m.p == 0.0 for m in model.modules() if isinstance(m, torch.nn.Dropout)), "a dropout is still active"
        loss, g = grads_of(ns, model, args, ids, labels, images)
        res[key] = dict(loss=loss, grads=g, seed=65, B=1, N=5, L=16, image_checksum=float(images.double().sum()))
    torch.save(res, os.path.join(HERE, "grads_tiny.pt"))
    for k, v in res.items():
        print(k, "loss", v["loss"], "params with grad", len(v["grads"]))


if __name__ == "__main__":
    main()
