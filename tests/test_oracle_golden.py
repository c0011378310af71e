"""The oracle restatement (oracle/berson_oracle.py) against the committed golden fixtures that the
REAL reference produced (tests/golden/make_golden.py).  CPU only."""
import os

import pytest
import torch

from oracle import berson_oracle as O
from oracle import synth

torch.set_grad_enabled(False)
ENC = ["sents", "para", "h0", "key", "cls", "cls_mat", "cls_score", "score_mat", "his1", "his2"]
TOL = 2e-5  # fp32, summation-order differences only (reference = torch/MKL kernels, oracle = explicit formulas)


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _cfg(g):
    c = g["cfg"]
    return dict(num_hidden_layers=c["num_hidden_layers"], num_attention_heads=c["num_attention_heads"],
                vit=g.get("vit"), rn=g.get("rn"))


def _check_case(sd, cfg, c, images=None):
    inp = O.prepare_inputs(c["ids"], c["labels"], c["N"], images)
    for k, v in c["prep"].items():
        assert torch.equal(inp[k], v), k          # integer host logic: bit-exact
    enc = O.encode(sd, cfg, inp)
    for k in ENC:
        assert (enc[k] - c["enc"][k]).abs().max() < TOL, k
    tr = []
    perm = O.beam_search(sd, enc, c["N"], c["W"], 0, tr)
    assert perm == c["perm"]
    assert len(tr) == len(c["steps"])
    for a, b in zip(tr, c["steps"]):
        assert torch.equal(a["beam_ix"], b["beam_ix"]) and torch.equal(a["tok_ix"], b["tok_ix"])
        assert (a["logp"] - b["logp"]).abs().max() < TOL
    return enc


def test_text_tiny(golden_dir):
    g = _load(golden_dir, "text_tiny.pt")
    for c in g["cases"]:
        _check_case(g["sd"], _cfg(g), c)


def test_training_loss(golden_dir):
    g = _load(golden_dir, "text_tiny.pt")
    lc = g["loss_case"]
    loss = O.training_loss(g["sd"], _cfg(g), O.prepare_inputs(lc["ids"], lc["labels"], 5))
    assert abs(loss.item() - lc["loss"].item()) < 1e-5


def test_mm_tiny(golden_dir):
    g = _load(golden_dir, "mm_tiny.pt")
    for c in g["cases"]:
        ids, labels, images = O.synthetic_manuals(1, c["N"], c["L"], vocab=1000, image_px=224, seed=c["seed"])
        assert abs(float(images.double().sum()) - c["image_checksum"]) < 1e-6, "torch RNG drift"
        assert torch.equal(ids, c["ids"])
        _check_case(g["sd"], _cfg(g), c, images)
        inp = O.prepare_inputs(ids, labels, c["N"], images)
        B, P, Lt = inp["input_ids"].shape
        im = inp["images"].reshape(B * P * 2, 3, 224, 224)
        tower = O.vit_pair_tower(g["sd"], "bert.encoder.visual_model.visual.", im[:6], g["vit"])
        assert (tower - c["tower"]).abs().max() < TOL
        lang, visn, pooled = O.lxrt_forward(g["sd"], _cfg(g), inp["input_ids"].reshape(B * P, Lt)[:3],
                                            inp["token_type_ids"].reshape(B * P, Lt)[:3],
                                            inp["attention_mask"].reshape(B * P, Lt)[:3], im[:6])
        assert (lang - c["lang"]).abs().max() < TOL and (visn - c["visn"]).abs().max() < TOL
        assert (pooled - c["pooled"][:3]).abs().max() < TOL


def test_mm_rn_tiny(golden_dir):
    """The "RN50" branch (ModifiedResNet + AttentionPool2d + visual_pos / visual_token_type) on a narrow tower."""
    g = _load(golden_dir, "mm_rn_tiny.pt")
    for c in g["cases"]:
        ids, labels, images = O.synthetic_manuals(1, c["N"], c["L"], vocab=1000, image_px=224, seed=c["seed"])
        assert abs(float(images.double().sum()) - c["image_checksum"]) < 1e-6, "torch RNG drift"
        _check_case(g["sd"], _cfg(g), c, images)
        inp = O.prepare_inputs(ids, labels, c["N"], images)
        B, P, Lt = inp["input_ids"].shape
        im = inp["images"].reshape(B * P * 2, 3, 224, 224)
        tower = O.rn_pair_tower(g["sd"], "bert.encoder.visual_model.visual.", im[:6], g["rn"])
        assert (tower - c["tower"]).abs().max() < TOL
        lang, visn, pooled = O.lxrt_forward_rn(g["sd"], _cfg(g), inp["input_ids"].reshape(B * P, Lt)[:3],
                                               inp["token_type_ids"].reshape(B * P, Lt)[:3],
                                               inp["attention_mask"].reshape(B * P, Lt)[:3], im[:6])
        assert (lang - c["lang"]).abs().max() < TOL and (visn - c["visn"]).abs().max() < TOL
        assert (pooled - c["pooled"][:3]).abs().max() < TOL


def test_decode_full_width(golden_dir):
    g = _load(golden_dir, "decode_full.pt")
    for c in g["cases"]:
        sd = synth.decode_head_weights(g["H"], c["head_seed"])
        enc = synth.synthetic_encode(c["N"], g["H"], c["enc_seed"])
        tr = []
        perm = O.beam_search(sd, enc, c["N"], c["W"], 0, tr)
        assert perm == c["perm"], (c["N"], c["W"])
        for a, b in zip(tr, c["steps"]):
            assert torch.equal(a["beam_ix"], b["beam_ix"]) and torch.equal(a["tok_ix"], b["tok_ix"])
            assert (a["logp"] - b["logp"]).abs().max() < TOL


def test_metrics_known_answers():
    """tau per models/berson/eval.py:237-247; hand-checkable cases."""
    acc, pmr, tau = O.cal_result([[0, 1, 2, 3, 4]], [[0, 1, 2, 3, 4]])
    assert (acc, pmr, tau) == (1.0, 1.0, 1.0)
    acc, pmr, tau = O.cal_result([[0, 1, 2, 3, 4]], [[4, 3, 2, 1, 0]])
    assert pmr == 0.0 and tau == -1.0 and abs(acc - 0.2) < 1e-12
    acc, pmr, tau = O.cal_result([[0, 1, 2]], [[1, 0, 2]])
    assert abs(tau - (1 - 2 * 1 / 3)) < 1e-12


def test_pairs_generator():
    p, n = O.pairs_generator(4)
    assert n == 12 and p[:6] == [[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]] and p[6] == [1, 0] and p[-1] == [3, 2]
