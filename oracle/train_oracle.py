"""ORACLE (test infrastructure only) for the fine-tuning path, SURVEY.md §8(f).2 / BASELINE config 4.

Gradients: torch autograd through the fp32 restatement in berson_oracle.py -- the same thing the reference gets from
`loss.backward()` (trainers/train.py:340-351) through models/berson/modeling_bert.py:943-1174 and
models/CLIP/src/lxrt/modeling.py:1513-1598.  PIN: tests/golden/grads_tiny.pt holds summaries of the gradients the REAL
reference produced (tests/golden/make_golden_grads.py); tests/test_oracle_golden.py checks `loss_grads` against them.

Optimizer: the reference uses transformers.AdamW (trainers/train.py:36,185; transformers==3.4.0 per requirements.txt),
which the installed transformers 5.5 no longer ships -- but the reference tree vendors the same class in
models/berson/optimization.py:107-189.  `hf_adamw_step` restates it (correct_bias=True):
    exp_avg    = b1 exp_avg + (1 - b1) g
    exp_avg_sq = b2 exp_avg_sq + (1 - b2) g^2
    p -= lr sqrt(1 - b2^t) / (1 - b1^t) * exp_avg / (sqrt(exp_avg_sq) + eps)
    p -= lr * weight_decay * p                      (decoupled, after the Adam update)
PIN: tests/test_train_oracle.py runs the vendored class itself (loaded from /root/reference) against this function; the
clip is torch.nn.utils.clip_grad_norm_ (train.py:358) and is checked against torch itself.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import torch

from . import berson_oracle as O


def _leaf_sd(sd):
    return {k: (v.detach().clone().requires_grad_(True) if torch.is_tensor(v) and v.is_floating_point() else v)
            for k, v in sd.items()}


def _padding_idx_rows(grads, pre, lxrt):
    """nn.Embedding(padding_idx=0) accumulates no gradient into row 0.  The LXRT embeddings set it on all three tables
    (lxrt/modeling.py:347-349), the text-only BertModel on the word table only (modeling_bert.py:153-155); the oracle's
    forward indexes plain tensors, so the rows are cleared here."""
    tables = ["word_embeddings"] + (["position_embeddings", "token_type_embeddings"] if lxrt else [])
    for t in tables:
        k = pre + "embeddings." + t + ".weight"
        if k in grads:
            grads[k][0].zero_()
    return grads


def inner_grads(sd, cfg, ids, tt, mask, images, g_lang, g_visn=None, pre="bert."):
    """d/dparam of  sum(lang * g_lang) + sum(visn * g_visn)  through the inner encoder
    (LXRTModel.forward BERSON mode, or the text-only BertModel when images is None)."""
    leaf = _leaf_sd(sd)
    with torch.enable_grad():
        if images is not None:
            lang, visn, _ = O.lxrt_forward(leaf, cfg, ids, tt, mask, images, pre)
            obj = (lang * g_lang).sum() + ((visn * g_visn).sum() if g_visn is not None else 0.0)
        else:
            lang, _ = O.text_bert(leaf, cfg, ids, mask, tt, pre)
            visn = None
            obj = (lang * g_lang).sum()
        obj.backward()
    grads = {k: v.grad for k, v in leaf.items() if torch.is_tensor(v) and v.is_floating_point() and v.grad is not None}
    return lang.detach(), (None if visn is None else visn.detach()), _padding_idx_rows(grads, pre, images is not None)


def loss_grads(sd, cfg, inp, lam=0.6, dropout=None, triplets=None, multimodal_loss=False):
    """(loss, {name: grad}) of BertForOrdering._forward's default objective (modeling_bert.py:943-1174).
    dropout: None (p = 0) or an oracle.dropout.DropSpec -- the training-mode forward with those masks."""
    leaf = _leaf_sd(sd)
    O.DROPOUT = dropout
    try:
        with torch.enable_grad():
            loss = O.training_loss(leaf, cfg, inp, lam, triplets=triplets, multimodal_loss=multimodal_loss)
            loss.backward()
    finally:
        O.DROPOUT = None
    grads = {k: v.grad for k, v in leaf.items() if torch.is_tensor(v) and v.is_floating_point() and v.grad is not None}
    lxrt = (cfg.get("vit") is not None or cfg.get("rn") is not None) and inp.get("images") is not None
    grads = {k: v for k, v in grads.items() if "running_" not in k}   # BatchNorm buffers are not parameters
    return float(loss.detach()), _padding_idx_rows(grads, "bert.", lxrt)


def decays(name):
    """trainers/train.py:172-181: no weight decay for names containing "bias" or "LayerNorm.weight"."""
    return not any(nd in name for nd in ("bias", "LayerNorm.weight"))


def clip_coef(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (train.py:358): returns (total_norm, coefficient applied to every gradient)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = max_norm / (total + 1e-6)
    return float(total), float(torch.clamp(coef, max=1.0))


def hf_adamw_step(p, g, m, v, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """One transformers.AdamW step on fp32 tensors (in place on p, m, v); `step` counts from 1."""
    b1, b2 = betas
    m.mul_(b1).add_(g, alpha=1.0 - b1)
    v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
    denom = v.sqrt().add_(eps)
    step_size = lr * (1.0 - b2 ** step) ** 0.5 / (1.0 - b1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)
    return p
