"""GPU parity tests of the fine-tuning path: training forward / backward / AdamW of the inner encoder
through the C ABI, against torch autograd through the oracle (oracle/train_oracle.py, itself pinned to gradients the
real reference produced, tests/test_train_oracle.py).

Tolerances: fp32 ("precise") mode -- every parameter gradient within 2e-4 of the oracle's in relative L2 norm (fp32
summation order over a few thousand rows); bf16 tensor-core mode -- 6e-2 relative L2 per parameter (bf16 activations and
bf16 gradient operands, fp32 accumulation), the bound is the test's own statement, not a reference figure."""
import os

import pytest
import torch

from oracle import berson_oracle as O
from oracle import train_oracle as TO

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


def _engine(sd, cfg, precise):
    from multimodal_sequencing_b200 import OrderingEngine
    return OrderingEngine(sd, cfg, precise=precise)


def _cfg_from_golden(g):
    c = g["cfg"]
    return dict(hidden_size=c["hidden_size"], num_hidden_layers=c["num_hidden_layers"],
                num_attention_heads=c["num_attention_heads"], intermediate_size=c["intermediate_size"],
                vocab_size=c["vocab_size_or_config_json_file"], max_position_embeddings=c["max_position_embeddings"],
                vit=g.get("vit"), rn=g.get("rn"), para_ff=g["ff_size"])


def _ocfg(g, bn_train=None):
    c = g["cfg"]
    rn = g.get("rn")
    if rn is not None and bn_train is not None:
        rn = dict(rn, bn_train=bn_train)
    return dict(num_hidden_layers=c["num_hidden_layers"], num_attention_heads=c["num_attention_heads"], vit=g.get("vit"), rn=rn)


def _ragged(n_steps, vocab, seed):
    gen = torch.Generator().manual_seed(seed)
    rows = []
    for _ in range(n_steps):
        ln = int(torch.randint(5, 22, (1,), generator=gen))
        rows.append(torch.cat([torch.tensor([101]), torch.randint(200, vocab, (ln,), generator=gen), torch.tensor([102])]))
    return torch.cat(rows)[None], torch.randperm(n_steps, generator=gen)[None]


def _case(g, multimodal, R, seed):
    """R pair rows of one ragged 4-step manual (padding -> the additive attention mask matters)."""
    ids, labels = _ragged(4, 1000, seed)
    images = torch.randn(1, 4, 3, 224, 224, generator=torch.Generator().manual_seed(seed + 1)) if multimodal else None
    return ids, labels, images


# biases that only shift all logits of a softmax: attention key biases, the token-score bias, the pointer-score bias
_ZERO_GRAD = ("attention.self.key.bias", "self_attn.linear_keys.bias", "tanh_linear.bias", "sentence_tran_2.bias")


def _compare(got, ref, tol, skip=()):
    worst = ("", 0.0)
    gmax = max(float(r.float().norm()) for r in ref.values())
    for n, r in ref.items():
        if n not in got or any(s in n for s in skip):
            continue
        a, b = got[n].detach().float().cpu().reshape(-1), r.float().reshape(-1)
        if n.endswith(_ZERO_GRAD):
            # softmax is invariant to a constant added to every score, so these gradients are exactly zero and both sides
            # hold rounding noise only: bound the noise against the largest gradient of the model.
            assert float(a.norm()) <= tol * gmax + 1e-6, "%s: %.3e vs largest gradient norm %.3e" % (n, float(a.norm()), gmax)
            continue
        err = float((a - b).norm() / (b.norm() + 1e-12)) if float(b.norm()) > 1e-9 else float((a - b).norm())
        if err > worst[1]:
            worst = (n, err)
        assert err <= tol, "%s: relative L2 error %.3e > %.1e (|ref| = %.3e)" % (n, err, tol, float(b.norm()))
    return worst


def _run(g, multimodal, precise, R=5, seed=7):
    eng = _engine(g["sd"], _cfg_from_golden(g), precise)
    ids, labels, images = _case(g, multimodal, R, seed)
    pb = eng.prepare(ids, labels, 4, images)
    iid, tt, am = pb.input_ids[0][:R], pb.token_type_ids[0][:R], pb.attention_mask[0][:R]
    assert int((am == 0).sum()) > 0, "the case must contain padding"
    idx = pb.img_index[0][:R] if multimodal else None
    lang, visn = eng.inner_forward_train(iid, tt, am, pb.images if multimodal else None, idx)
    gen = torch.Generator().manual_seed(seed + 2)
    g_lang = torch.randn(lang.shape, generator=gen)
    g_visn = torch.randn(visn.shape, generator=gen) if multimodal else None
    grads = eng.new_grad_buffer()
    eng.inner_backward(g_lang, g_visn, grads)
    torch.cuda.synchronize()
    om = pb.images[idx.reshape(-1).long()].cpu() if multimodal else None
    olang, ovisn, ref = TO.inner_grads(g["sd"], _ocfg(g), iid.cpu(), tt.cpu(), am.cpu(), om, g_lang, g_visn)
    return eng, grads, lang, visn, olang, ovisn, ref, (iid, tt, am, idx, pb)


@pytest.mark.parametrize("precise", [True, False])
def test_text_encoder_backward_vs_oracle_autograd(golden_dir, precise):
    g = torch.load(os.path.join(golden_dir, "text_tiny.pt"), weights_only=False)
    eng, grads, lang, _, olang, _, ref, _ = _run(g, False, precise)
    tol_f, tol_g = (2e-5, 2e-4) if precise else (3e-2, 6e-2)
    assert (lang.cpu() - olang).abs().max() <= tol_f * max(1.0, float(olang.abs().max()))
    got = {n: v for n, v in eng.grads_by_name(grads).items() if n.startswith("bert.")}   # the heads take no part here
    assert set(got) <= set(ref) | {"bert.pooler.dense.weight", "bert.pooler.dense.bias"}
    worst = _compare(got, ref, tol_g)
    print("text backward (%s): worst relative L2 %.2e at %s" % ("fp32" if precise else "bf16", worst[1], worst[0]))
    # nn.Embedding(padding_idx=0): the text-only BertModel keeps row 0 of the word table gradient-free only
    assert float(got["bert.embeddings.word_embeddings.weight"].reshape(1000, -1)[0].abs().max()) == 0.0
    assert float(got["bert.embeddings.position_embeddings.weight"].reshape(256, -1)[0].abs().max()) > 0.0


@pytest.mark.parametrize("precise", [True, False])
def test_multimodal_encoder_backward_vs_oracle_autograd(golden_dir, precise):
    g = torch.load(os.path.join(golden_dir, "mm_tiny.pt"), weights_only=False)
    eng, grads, lang, visn, olang, ovisn, ref, _ = _run(g, True, precise)
    tol_f, tol_g = (4e-5, 3e-4) if precise else (3e-2, 8e-2)
    assert (lang.cpu() - olang).abs().max() <= tol_f * max(1.0, float(olang.abs().max()))
    assert (visn.cpu() - ovisn).abs().max() <= tol_f * max(1.0, float(ovisn.abs().max()))
    got = {n: v for n, v in eng.grads_by_name(grads).items() if n.startswith("bert.")}   # the heads take no part here
    names = [n for n in got if n in ref]
    assert len(names) == len(got), sorted(set(got) - set(ref))
    worst = _compare(got, ref, tol_g)
    print("multimodal backward (%s): worst relative L2 %.2e at %s" % ("fp32" if precise else "bf16", worst[1], worst[0]))
    # LXRT embeddings: padding_idx=0 on all three tables (lxrt/modeling.py:347-349)
    H = g["cfg"]["hidden_size"]
    for t in ("word_embeddings", "position_embeddings", "token_type_embeddings"):
        assert float(got["bert.embeddings.%s.weight" % t].reshape(-1, H)[0].abs().max()) == 0.0, t


def test_backward_accumulates_and_is_repeatable(golden_dir):
    g = torch.load(os.path.join(golden_dir, "text_tiny.pt"), weights_only=False)
    eng, grads, lang, _, _, _, _, (iid, tt, am, _, _) = _run(g, False, True)
    once = grads.clone()
    g_lang = torch.randn(lang.shape, generator=torch.Generator().manual_seed(9))   # seed + 2 of _run
    eng.inner_forward_train(iid, tt, am)
    eng.inner_backward(g_lang, None, grads)
    torch.cuda.synchronize()
    # embedding tables are scattered with fp32 atomics (order-dependent rounding); everything else is deterministic
    lay = {n: (o, k) for n, o, k, _ in eng.train_layout()}
    for n, (o, k) in lay.items():
        a, b = grads[o:o + k], 2 * once[o:o + k]
        if "embeddings.word" in n or "embeddings.position" in n or "embeddings.token_type" in n:
            assert (a - b).abs().max() <= 1e-5 * max(1.0, float(b.abs().max())), n
        else:
            assert torch.equal(a, b), n


@pytest.mark.parametrize("multimodal", [False, True])
def test_adamw_step_and_repack(golden_dir, multimodal):
    """clip_grad_norm_ + transformers.AdamW on the fp32 masters, twice; afterwards the eval path (which reads the packed
    copies) must see the updated weights."""
    g = torch.load(os.path.join(golden_dir, "mm_tiny.pt" if multimodal else "text_tiny.pt"), weights_only=False)
    eng, grads, _, _, _, _, ref, (iid, tt, am, idx, pb) = _run(g, multimodal, True)
    lay = eng.train_layout()
    sd = {n: g["sd"][n].clone().float() for n, _, _, _ in lay}
    m = {n: torch.zeros_like(v) for n, v in sd.items()}
    v = {n: torch.zeros_like(v_) for n, v_ in sd.items()}
    lr, wd, max_norm = 1e-3, 0.01, 0.5
    for step in (1, 2):
        dev = eng.grads_by_name(grads)
        host = {n: dev[n].detach().cpu().reshape(sd[n].shape).clone() for n in sd}
        total, coef = TO.clip_coef(list(host.values()), max_norm)
        norm = eng.adamw_step(grads, lr, weight_decay=wd, max_grad_norm=max_norm)
        torch.cuda.synchronize()
        assert abs(float(norm[0]) - total) <= 1e-4 * total and abs(float(norm[1]) - coef) <= 1e-4 * coef
        for n, _, k, dec in lay:
            assert dec == TO.decays(n), n
            TO.hf_adamw_step(sd[n], host[n] * coef, m[n], v[n], step, lr, weight_decay=wd if dec else 0.0)
            got = eng.read_param(n, sd[n].shape).cpu()
            assert (got - sd[n]).abs().max() <= 2e-6 + 1e-5 * float(sd[n].abs().max()), (step, n)
        # a second step with the same gradients exercises the moment buffers
    full = dict(g["sd"])
    full.update(sd)
    om = pb.images[idx.reshape(-1).long()].cpu() if multimodal else None
    if multimodal:
        olang, _, _ = O.lxrt_forward(full, _ocfg(g), iid.cpu(), tt.cpu(), am.cpu(), om)
    else:
        olang, _ = O.text_bert(full, _ocfg(g), iid.cpu(), am.cpu(), tt.cpu())
    lang, _, _ = eng.inner_forward(iid, tt, am, pb.images if multimodal else None, idx)
    assert (lang.cpu() - olang).abs().max() <= 4e-5 * max(1.0, float(olang.abs().max()))
    lang2, _ = eng.inner_forward_train(iid, tt, am, pb.images if multimodal else None, idx)
    assert (lang2.cpu() - olang).abs().max() <= 4e-5 * max(1.0, float(olang.abs().max()))


def test_backward_without_forward_fails(golden_dir):
    g = torch.load(os.path.join(golden_dir, "text_tiny.pt"), weights_only=False)
    eng = _engine(g["sd"], _cfg_from_golden(g), True)
    grads = eng.new_grad_buffer()
    with pytest.raises(RuntimeError):
        eng.inner_backward(torch.zeros(1, 4, 128), None, grads)


# ---- full fine-tuning step: loss + gradients of EVERY parameter (encoder and BERSON heads) -----------------------------
def _full_step(g, multimodal, precise, B, N, L, seed):
    eng = _engine(g["sd"], _cfg_from_golden(g), precise)
    ids, labels, images = O.synthetic_manuals(B, N, L, vocab=1000, image_px=224 if multimodal else None, seed=seed)
    pb = eng.prepare(ids, labels, N, images)
    grads = eng.new_grad_buffer()
    loss = eng.train_step(pb, grads)
    torch.cuda.synchronize()
    oloss, ref = TO.loss_grads(g["sd"], _ocfg(g), O.prepare_inputs(ids, labels, N, images))
    return eng, pb, grads, float(loss), oloss, ref


@pytest.mark.parametrize("precise", [True, False])
def test_text_train_step_vs_reference_pinned_oracle(golden_dir, precise):
    """Same batch as the reference-generated gradient fixture (grads_tiny.pt 'text': 3 five-step manuals)."""
    g = torch.load(os.path.join(golden_dir, "text_tiny.pt"), weights_only=False)
    r = torch.load(os.path.join(golden_dir, "grads_tiny.pt"), weights_only=False)["text"]
    eng, pb, grads, loss, oloss, ref = _full_step(g, False, precise, r["B"], r["N"], r["L"], r["seed"])
    assert abs(oloss - r["loss"]) < 2e-5                      # the oracle reproduces the reference's loss ...
    assert abs(loss - r["loss"]) < (5e-5 if precise else 2e-2)     # ... and so does the CUDA path
    got = eng.grads_by_name(grads)
    assert set(got) == set(r["grads"]) - {"bert.pooler.dense.weight", "bert.pooler.dense.bias"}, sorted(set(got) ^ set(r["grads"]))
    worst = _compare(got, ref, 5e-4 if precise else 1e-1)
    print("text train step (%s): loss %.6f (reference %.6f), worst relative L2 %.2e at %s" %
          ("fp32" if precise else "bf16", loss, r["loss"], worst[1], worst[0]))
    if precise:   # directly against the reference's own numbers (norms + strided samples)
        for n, s in r["grads"].items():
            if n not in got:
                continue
            a = got[n].detach().double().cpu().reshape(-1)
            assert abs(float(a.norm()) - s["norm"]) <= 1e-3 * s["norm"] + 1e-7, n
            assert (a[s["idx"]].float() - s["val"]).abs().max() <= 1e-3 * s["norm"] / max(s["numel"], 1) ** 0.5 + 1e-5 * float(s["val"].abs().max()) + 1e-7, n


@pytest.mark.parametrize("precise", [True, False])
def test_multimodal_train_step_vs_reference_pinned_oracle(golden_dir, precise):
    g = torch.load(os.path.join(golden_dir, "mm_tiny.pt"), weights_only=False)
    r = torch.load(os.path.join(golden_dir, "grads_tiny.pt"), weights_only=False)["mm"]
    eng, pb, grads, loss, oloss, ref = _full_step(g, True, precise, r["B"], r["N"], r["L"], r["seed"])
    assert abs(oloss - r["loss"]) < 2e-5
    assert abs(loss - r["loss"]) < (5e-5 if precise else 2e-2)
    got = eng.grads_by_name(grads)
    assert set(got) == set(r["grads"]) - {"bert.pooler.dense.weight", "bert.pooler.dense.bias"}, sorted(set(got) ^ set(r["grads"]))
    worst = _compare(got, ref, 5e-4 if precise else 1e-1)
    print("multimodal train step (%s): loss %.6f (reference %.6f), worst relative L2 %.2e at %s" %
          ("fp32" if precise else "bf16", loss, r["loss"], worst[1], worst[0]))


@pytest.mark.parametrize("bn_train", [True, False])
@pytest.mark.parametrize("precise", [True, False])
def test_resnet_train_step_vs_reference_pinned_oracle(golden_dir, precise, bn_train):
    """Fine-tuning step through the ModifiedResNet tower (the reference's wired backbone): loss and EVERY parameter gradient
    against torch autograd through the oracle, on the batch of the reference-generated fixtures (grads_tiny.pt 'mm_rn_bntrain':
    nn.BatchNorm2d in train() mode, the mode the reference fine-tunes in; 'mm_rn': frozen running statistics)."""
    g = torch.load(os.path.join(golden_dir, "mm_rn_tiny.pt"), weights_only=False)
    r = torch.load(os.path.join(golden_dir, "grads_tiny.pt"), weights_only=False)["mm_rn_bntrain" if bn_train else "mm_rn"]
    eng = _engine(g["sd"], _cfg_from_golden(g), precise)
    eng.set_bn_mode(use_running_stats=not bn_train)
    ids, labels, images = O.synthetic_manuals(r["B"], r["N"], r["L"], vocab=1000, image_px=224, seed=r["seed"])
    pb = eng.prepare(ids, labels, r["N"], images)
    grads = eng.new_grad_buffer()
    loss = float(eng.train_step(pb, grads))
    torch.cuda.synchronize()
    oloss, ref = TO.loss_grads(g["sd"], _ocfg(g, bn_train), O.prepare_inputs(ids, labels, r["N"], images))
    assert abs(oloss - r["loss"]) < 2e-5                            # the oracle reproduces the reference's loss ...
    assert abs(loss - r["loss"]) < (5e-5 if precise else 2e-2)     # ... and so does the CUDA path
    got = eng.grads_by_name(grads)
    assert set(got) == set(r["grads"]) - {"bert.pooler.dense.weight", "bert.pooler.dense.bias"}, sorted(set(got) ^ set(r["grads"]))
    # fp32, batch statistics: the BatchNorm backward subtracts the mean and the xhat-projection of the incoming gradient, so the
    # convolution gradients in front of it are small differences of large fp32 sums -- against a float64 run BOTH the fp32
    # oracle and the fp32 reference are 2-4e-3 away on the tower's early layers (tests/test_train_oracle.py), hence 1e-2 here;
    # frozen statistics have no such cancellation: 5e-4.  bf16, frozen statistics: 1e-1 (13 convolutions deep).
    # bf16 + batch statistics ON THIS FIXTURE (one manual, 5 images, 8-128 channels, random init): the fp32 run above shows a
    # condition number of ~5e4 (6e-8 -> 7e-3) for everything in front of the last BatchNorm, i.e. bf16's 4e-3 per operand
    # leaves those gradients noise-dominated in ANY bf16 implementation; their concatenation is bounded loosely (0.6: no
    # blow-up, no sign flip of the bulk), while the loss and every gradient behind the tower stay within 1.5e-1.
    tower = "visual_model.visual."
    if precise or not bn_train:
        worst = _compare(got, ref, (1e-2 if bn_train else 5e-4) if precise else 1e-1)
    else:
        worst = _compare({n: v for n, v in got.items() if tower not in n or "attnpool" in n}, ref, 1.5e-1)
        names = [n for n in got if tower in n and "attnpool" not in n]
        a = torch.cat([got[n].detach().float().cpu().reshape(-1) for n in names])
        b = torch.cat([ref[n].float().reshape(-1) for n in names])
        assert bool(torch.isfinite(a).all())
        rel = float((a - b).norm() / b.norm())
        print("  tower convolutions / BatchNorms (noise-dominated in bf16, see above): relative L2 of the concatenated gradient %.2e" % rel)
        assert rel <= 0.6, rel
    print("ResNet train step (%s, BatchNorm %s): loss %.6f (reference %.6f), worst relative L2 %.2e at %s" %
          ("fp32" if precise else "bf16", "batch statistics" if bn_train else "running statistics", loss, r["loss"], worst[1], worst[0]))
    if precise:   # directly against the reference's own numbers (norms + strided samples)
        for n, s_ in r["grads"].items():
            if n not in got:
                continue
            a = got[n].detach().double().cpu().reshape(-1)
            assert abs(float(a.norm()) - s_["norm"]) <= (1e-2 if bn_train else 1e-3) * s_["norm"] + 1e-7, n


def test_resnet_running_statistics_follow_the_batch(golden_dir):
    """nn.BatchNorm2d.train() bookkeeping: after one training forward the registered running_mean / running_var of the stem's
    first two BatchNorms equal torch's own update (momentum 0.1, unbiased batch variance) over the MATERIALISED pair images."""
    import torch.nn.functional as F
    g = torch.load(os.path.join(golden_dir, "mm_rn_tiny.pt"), weights_only=False)
    sd = g["sd"]
    eng = _engine(sd, _cfg_from_golden(g), True)
    ids, labels, images = O.synthetic_manuals(1, 5, 16, vocab=1000, image_px=224, seed=65)
    pb = eng.prepare(ids, labels, 5, images)
    eng.train_step(pb, eng.new_grad_buffer())
    torch.cuda.synchronize()
    v = "bert.encoder.visual_model.visual."
    x = pb.images[pb.img_index.reshape(-1).long()].cpu()                      # every image once per pair slot
    for i, stride in ((1, 2), (2, 1)):
        y = F.conv2d(x, sd[v + "conv%d.weight" % i], stride=stride, padding=1)
        rm, rv = sd[v + "bn%d.running_mean" % i].clone(), sd[v + "bn%d.running_var" % i].clone()
        x = torch.relu(F.batch_norm(y, rm, rv, sd[v + "bn%d.weight" % i], sd[v + "bn%d.bias" % i], training=True, momentum=0.1, eps=1e-5))
        got_m = eng.read_param(v + "bn%d.running_mean" % i, tuple(rm.shape)).cpu()
        got_v = eng.read_param(v + "bn%d.running_var" % i, tuple(rv.shape)).cpu()
        assert (got_m - rm).abs().max() <= 1e-5 * max(1.0, float(rm.abs().max())), (i, got_m, rm)
        assert (got_v - rv).abs().max() <= 1e-5 * max(1.0, float(rv.abs().max())), (i, got_v, rv)
        assert (got_m - sd[v + "bn%d.running_mean" % i]).abs().max() > 1e-4      # they did move
    # frozen statistics: nothing moves
    eng2 = _engine(sd, _cfg_from_golden(g), True)
    eng2.set_bn_mode(use_running_stats=True)
    eng2.train_step(pb, eng2.new_grad_buffer())
    torch.cuda.synchronize()
    assert torch.equal(eng2.read_param(v + "bn1.running_mean", tuple(sd[v + "bn1.running_mean"].shape)).cpu(), sd[v + "bn1.running_mean"])


@pytest.mark.parametrize("precise", [True, False])
def test_train_step_with_time_contrastive_objective(golden_dir, precise):
    """The reference's optional time-contrastive term (modeling_bert.py:1176-1216, 0.1 * TripletMarginLoss over sentence
    vectors): loss and every gradient against torch autograd through the oracle with the same triplets."""
    g = torch.load(os.path.join(golden_dir, "text_tiny.pt"), weights_only=False)
    eng = _engine(g["sd"], _cfg_from_golden(g), precise)
    B, N, L = 3, 5, 24
    ids, labels, _ = O.synthetic_manuals(B, N, L, vocab=1000, seed=33)
    pb = eng.prepare(ids, labels, N, None)
    inp = O.prepare_inputs(ids, labels, N, None)
    import numpy as np
    trip = O.time_contrastive_triplets(inp["ground_truth"], np.random.RandomState(5))
    assert all(len(set(t)) == 3 for t in trip.tolist())
    grads = eng.new_grad_buffer()
    loss = float(eng.train_step(pb, grads, triplets=trip))
    torch.cuda.synchronize()
    oloss, ref = TO.loss_grads(g["sd"], _ocfg(g), inp, triplets=trip)
    o0, _ = TO.loss_grads(g["sd"], _ocfg(g), inp)
    assert oloss - o0 > 1e-3, "the hinge must be active for the test to mean anything"
    assert abs(loss - oloss) < (5e-5 if precise else 2e-2), (loss, oloss)
    worst = _compare(eng.grads_by_name(grads), ref, 5e-4 if precise else 1e-1)
    print("time-contrastive train step (%s): loss %.6f (oracle %.6f, without the term %.6f), worst relative L2 %.2e at %s" %
          ("fp32" if precise else "bf16", loss, oloss, o0, worst[1], worst[0]))
    # one-shot: the next step runs the default objective again
    grads.zero_()
    assert abs(float(eng.train_step(pb, grads)) - o0) < (5e-5 if precise else 2e-2)


@pytest.mark.parametrize("precise", [True, False])
def test_train_step_with_image_pairwise_objective(golden_dir, precise):
    """args.multimodal_loss of the reference (modeling_bert.py:897-898, 1359-1364, 1218-1225): img_projection of the first visual
    token through the shared pairwise_relationship head.  Loss and every gradient (img_projection.*, the shared head, the whole
    encoder below the visual token) against torch autograd through the oracle, whose term is pinned live to the reference
    (tests/test_oracle_vs_reference.py::test_multimodal_loss_objective_live)."""
    g = torch.load(os.path.join(golden_dir, "mm_tiny.pt"), weights_only=False)
    H = g["cfg"]["hidden_size"]
    gen = torch.Generator().manual_seed(123)
    sd = dict(g["sd"])
    sd["img_projection.weight"] = torch.randn(H, H, generator=gen) * 0.05
    sd["img_projection.bias"] = torch.randn(H, generator=gen) * 0.05
    eng = _engine(sd, _cfg_from_golden(g), precise)
    eng.set_multimodal_loss(True)
    B, N, L = 2, 4, 12
    ids, labels, images = O.synthetic_manuals(B, N, L, vocab=1000, image_px=224, seed=57)
    pb = eng.prepare(ids, labels, N, images)
    inp = O.prepare_inputs(ids, labels, N, images)
    grads = eng.new_grad_buffer()
    loss = float(eng.train_step(pb, grads))
    torch.cuda.synchronize()
    oloss, ref = TO.loss_grads(sd, _ocfg(g), inp, multimodal_loss=True)
    o0, ref0 = TO.loss_grads(sd, _ocfg(g), inp)
    assert oloss - o0 > 1e-2, "the image term must contribute for the test to mean anything"
    assert "img_projection.weight" in ref and float(ref["img_projection.weight"].norm()) > 0
    assert abs(loss - oloss) < (5e-5 if precise else 2e-2), (loss, oloss)
    got = eng.grads_by_name(grads)
    assert "img_projection.weight" in got and "img_projection.bias" in got
    worst = _compare(got, ref, 5e-4 if precise else 1e-1)
    print("image pairwise objective (%s): loss %.6f (oracle %.6f, without the term %.6f), worst relative L2 %.2e at %s" %
          ("fp32" if precise else "bf16", loss, oloss, o0, worst[1], worst[0]))
    # switching it off restores the default objective
    eng.set_multimodal_loss(False)
    grads.zero_()
    assert abs(float(eng.train_step(pb, grads)) - o0) < (5e-5 if precise else 2e-2)
    assert float(eng.grads_by_name(grads)["img_projection.weight"].abs().max()) == 0.0


def test_fine_tuning_lowers_the_loss(golden_dir):
    """Five optimizer steps on one batch: the loss the step reports must fall (end-to-end sign / wiring check), and the
    eval-mode loss entry point must agree with the training one before and after."""
    g = torch.load(os.path.join(golden_dir, "text_tiny.pt"), weights_only=False)
    eng = _engine(g["sd"], _cfg_from_golden(g), True)
    ids, labels, _ = O.synthetic_manuals(4, 5, 16, vocab=1000, seed=3)
    pb = eng.prepare(ids, labels, 5, None)
    losses = []
    for step in range(5):
        grads = eng.new_grad_buffer()
        losses.append(float(eng.train_step(pb, grads)))
        if step == 0:
            assert abs(losses[0] - float(eng.training_loss(pb))) < 1e-5
        eng.adamw_step(grads, 1e-3, max_grad_norm=1.0)
    final = float(eng.training_loss(pb))
    print("losses", losses, "final", final)
    assert final < losses[0] - 0.05 and all(b < a + 1e-3 for a, b in zip(losses, losses[1:]))


# ---- full size (BERT-base + ViT-B/32): BASELINE configs[3] shape, one 6-step manual ------------------------------------
def test_full_size_train_step_vs_oracle_autograd():
    """RecipeQA-shaped manual (6 steps x 64 tokens, 30 pairs of 227 joint tokens) through the full-size model: loss and every
    gradient against torch autograd through the oracle on the box's host cores.  fp32 mode: 2e-3 relative L2 per parameter
    (fp32 sums over 6 810 joint rows and 24 layers); bf16 tensor-core mode: 1.5e-1, reported."""
    from oracle import synth
    cfg = dict(synth.BERT_BASE)
    vit = dict(synth.VIT_B32)
    cfg.update(vit=vit, rn=None, para_ff=3072)
    sd = synth.full_state_dict(cfg, vit, seed=0)
    N = 6
    ids, labels, images = O.synthetic_manuals(1, N, 64, image_px=224, seed=4)
    torch.set_num_threads(os.cpu_count() or 1)
    ocfg = dict(num_hidden_layers=12, num_attention_heads=12, vit=vit)
    oloss, ref = TO.loss_grads(sd, ocfg, O.prepare_inputs(ids, labels, N, images))
    for precise in (True, False):
        eng = _engine(sd, cfg, precise)
        pb = eng.prepare(ids, labels, N, images)
        grads = eng.new_grad_buffer()
        loss = float(eng.train_step(pb, grads))
        torch.cuda.synchronize()
        assert abs(loss - oloss) < (2e-4 if precise else 5e-2), (loss, oloss)
        got = eng.grads_by_name(grads)
        worst = _compare(got, ref, 2e-3 if precise else 1.5e-1)
        print("full-size train step (%s): loss %.6f (oracle %.6f), worst relative L2 %.2e at %s" %
              ("fp32" if precise else "bf16", loss, oloss, worst[1], worst[0]))
        del eng, grads, got
        torch.cuda.empty_cache()


# ---- kernel level: tcgen05 GEMM with MN-major ("TN") operands, the weight-gradient shape -------------------------------
@pytest.mark.parametrize("shape", [(768, 768, 6810), (3072, 768, 1000), (128, 512, 77), (2304, 768, 54480)])
def test_gemm_tn_operands_vs_float64(shape):
    """C[M,N] = A^T W + resid with A [K,M], W [K,N] bf16 row-major read as MN-major UMMA operands (no transposes);
    K is deliberately not a multiple of 64 (TMA zero-fills the tail rows)."""
    import ctypes as C
    from multimodal_sequencing_b200 import _lib
    lib = _lib.load()
    if not lib.msq_tc_available():
        pytest.skip("no tcgen05 device")
    M, N, K = shape
    gen = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(K, M, generator=gen).to(torch.bfloat16).cuda()
    W = torch.randn(K, N, generator=gen).to(torch.bfloat16).cuda()
    R = torch.randn(M, N, generator=gen).cuda()
    out = R.clone()
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(lib.msq_gemm(5, p(A), p(W), None, p(out), p(out), M, N, K, 0, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    ref = A.double().t() @ W.double() + R.double()
    err = (out.double() - ref).abs().max().item()
    bound = 2e-5 * (K ** 0.5) * 4 + 1e-3      # fp32 accumulation of K bf16 x bf16 products of O(1) magnitude
    assert err <= bound * max(1.0, ref.abs().max().item() / (K ** 0.5)), (err, bound)


@pytest.mark.parametrize("shape,act", [((6810, 3072, 768), 1), ((23760, 3072, 768), 2), ((100, 512, 128), 1), ((300, 264, 64), 2)])
def test_gemm_dual_output_vs_float64(shape, act):
    """EPI_DUALACT of the tcgen05 GEMM (the up-projections of the fine-tuning forward): C = bf16(A W^T + b) and
    C2 = bf16(act(A W^T + b)) from one accumulator read; CTA-pair path with a partial last row tile, and the single-CTA path
    with partial column tiles.  act 1 = erf-GELU (lxrt/modeling.py:116-122), 2 = QuickGELU (clip/model.py:199-201)."""
    import ctypes as C
    from multimodal_sequencing_b200 import _lib
    lib = _lib.load()
    if not lib.msq_tc_available():
        pytest.skip("no tcgen05 device")
    M, N, K = shape
    gen = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=gen) * 0.5).to(torch.bfloat16).cuda()
    W = (torch.randn(N, K, generator=gen) * (2.0 / K ** 0.5)).to(torch.bfloat16).cuda()
    b = torch.randn(N, generator=gen).cuda()
    u = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    h = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(lib.msq_gemm_deferred_ln(3, 1, p(A), p(W), p(b), None, None, None, None, 0, 0, 0.0, p(u), p(h), None, M, N, K, act,
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t() + b.double()
    ract = ref * 0.5 * (1 + torch.erf(ref / 2 ** 0.5)) if act == 1 else ref * torch.sigmoid(1.702 * ref)
    assert not torch.isnan(u.float()).any() and not torch.isnan(h.float()).any(), "every element of both outputs must be written"
    scale = max(1.0, ref.abs().max().item())
    assert (u.double() - ref).abs().max().item() <= 2 ** -8 * scale + 1e-3          # bf16 rounding of the stored value
    assert (h.double() - ract).abs().max().item() <= 2 ** -8 * scale + 2e-3         # + the tanh-fit of erf (|err| < 3e-5)


@pytest.mark.parametrize("precise", [True, False])
def test_multimodal_train_step_two_manuals(golden_dir, precise):
    """B = 2 multimodal manuals of 4 steps (unique-image table shared across the batch, pair rows of both manuals in one
    encoder pass) against oracle autograd."""
    g = torch.load(os.path.join(golden_dir, "mm_tiny.pt"), weights_only=False)
    eng, pb, grads, loss, oloss, ref = _full_step(g, True, precise, 2, 4, 12, 91)
    assert abs(loss - oloss) < (5e-5 if precise else 2e-2)
    worst = _compare(eng.grads_by_name(grads), ref, 5e-4 if precise else 1e-1)
    print("multimodal B=2 train step (%s): loss %.6f (oracle %.6f), worst relative L2 %.2e at %s" %
          ("fp32" if precise else "bf16", loss, oloss, worst[1], worst[0]))


def test_full_size_gradient_additivity_over_the_batch():
    """Size-independent property at BASELINE configs[3] size (BERT-base + ViT-B/32, six-step manuals, bf16 tensor-core path):
    the loss is a batch mean, so one step on {a, b} must equal the two single-manual steps accumulated into one buffer and
    halved -- loss exactly so up to fp32 rounding, gradients up to the summation order of the weight-gradient GEMMs."""
    from oracle import synth
    cfg = dict(synth.BERT_BASE)
    vit = dict(synth.VIT_B32)
    cfg.update(vit=vit, rn=None, para_ff=3072)
    eng = _engine(synth.full_state_dict(cfg, vit, seed=0), cfg, False)
    ids, labels, images = O.synthetic_manuals(2, 6, 64, image_px=224, seed=12)
    both = eng.new_grad_buffer()
    loss_ab = float(eng.train_step(eng.prepare(ids, labels, 6, images), both))
    acc = eng.new_grad_buffer()
    la = float(eng.train_step(eng.prepare(ids[:1], labels[:1], 6, images[:1]), acc))
    lb = float(eng.train_step(eng.prepare(ids[1:], labels[1:], 6, images[1:]), acc))   # accumulates
    torch.cuda.synchronize()
    assert abs(loss_ab - 0.5 * (la + lb)) < 2e-5
    worst = ("", 0.0)
    for n, o, k, _ in eng.train_layout():
        a, b = both[o:o + k], 0.5 * acc[o:o + k]
        ref = float(b.norm())
        if ref < 1e-7:
            continue
        err = float((a - b).norm()) / ref
        if err > worst[1]:
            worst = (n, err)
        assert err < 2e-2 or n.endswith(_ZERO_GRAD), "%s: %.3e" % (n, err)
    print("batch additivity at full size: loss %.6f vs %.6f, worst relative L2 %.2e at %s" % (loss_ab, 0.5 * (la + lb), worst[1], worst[0]))


# ---------------------------------------------------------------------------------------------------
# dropout (the reference fine-tunes with p = 0.1): counter-based masks, regenerated in the backward pass
# ---------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("multimodal", [False, True])
@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("probs", [(0.1, 0.1, 0.1), (0.3, 0.0, 0.0), (0.0, 0.25, 0.0), (0.0, 0.0, 0.2)])
def test_train_step_with_dropout_vs_oracle_same_masks(golden_dir, multimodal, precise, probs):
    """msq_train_step with dropout at the embedding / visn_fc / attention-prob / dense-output / token-attention / paragraph-encoder
    sites against torch autograd through the oracle with the SAME masks (oracle/dropout.py regenerates them from seed, forward
    counter, site and element index).  Loss and every parameter gradient; then a second step (a new counter -> new masks)."""
    from oracle.dropout import DropSpec
    g = torch.load(os.path.join(golden_dir, "mm_tiny.pt" if multimodal else "text_tiny.pt"), weights_only=False)
    eng = _engine(g["sd"], _cfg_from_golden(g), precise)
    B, N, L = 2, 5, 12
    ids, labels, images = O.synthetic_manuals(B, N, L, vocab=1000, image_px=224 if multimodal else None, seed=41)
    pb = eng.prepare(ids, labels, N, images)
    inp = O.prepare_inputs(ids, labels, N, images)
    ph, pa, pp = probs
    eng.set_dropout(ph, pa, pp, seed=1234)
    losses = []
    for it in range(2):
        grads = eng.new_grad_buffer()
        loss = float(eng.train_step(pb, grads))
        step = eng.dropout_step()
        assert step == it
        oloss, ref = TO.loss_grads(g["sd"], _ocfg(g), inp, dropout=DropSpec(1234, step, ph, pa, pp))
        assert abs(loss - oloss) < (1e-4 if precise else 3e-2), (it, loss, oloss)
        worst = _compare(eng.grads_by_name(grads), ref, 1e-3 if precise else 1.5e-1)
        losses.append(loss)
        print("dropout %s %s %s step %d: loss %.6f (oracle %.6f), worst relative L2 %.2e at %s" %
              (probs, "mm" if multimodal else "text", "fp32" if precise else "bf16", it, loss, oloss, worst[1], worst[0]))
    assert abs(losses[0] - losses[1]) > 1e-6, "the two steps must see different masks"
    # p = 0 restores the deterministic path exactly
    eng.set_dropout(0.0, 0.0, 0.0, seed=1234)
    g0 = eng.new_grad_buffer()
    l0 = float(eng.train_step(pb, g0))
    o0, _ = TO.loss_grads(g["sd"], _ocfg(g), inp)
    assert abs(l0 - o0) < (1e-4 if precise else 3e-2)
    # evaluation never drops anything
    eng.set_dropout(0.5, 0.5, 0.5, seed=1)
    assert abs(float(eng.training_loss(pb)) - o0) < (1e-4 if precise else 3e-2)


def test_dropout_keep_rate_and_scaling():
    """statistics of the device masks: keep rate ~ 1 - p and E[dropout(x)] ~ x at a site large enough to measure."""
    import ctypes as C
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "text_tiny.pt"), weights_only=False)
    eng = _engine(g["sd"], _cfg_from_golden(g), True)
    ids, labels, _ = O.synthetic_manuals(4, 5, 20, vocab=1000, seed=3)
    pb = eng.prepare(ids, labels, 5)
    R = pb.input_ids.shape[0] * pb.input_ids.shape[1]
    iid, tt, am = [t.reshape(R, -1) for t in (pb.input_ids, pb.token_type_ids, pb.attention_mask)]
    eng.set_dropout(0.0, 0.0, 0.0, 7)
    base, _ = eng.inner_forward_train(iid, tt, am)
    # dropout on the LAST layer's dense outputs only is not separable; measure the embedding site through a 0-layer model instead
    cfg0 = dict(_cfg_from_golden(g), num_hidden_layers=0)
    e0 = _engine(g["sd"], cfg0, True)
    e0.set_dropout(0.0, 0.0, 0.0, 7)
    x0, _ = e0.inner_forward_train(iid, tt, am)
    e0.set_dropout(0.2, 0.0, 0.0, 7)
    x1, _ = e0.inner_forward_train(iid, tt, am)
    kept = (x1 != 0) | (x0 == 0)
    rate = float(kept.float().mean())
    assert abs(rate - 0.8) < 0.01, rate
    assert torch.allclose(x1[kept], x0[kept] / 0.8, rtol=1e-5, atol=1e-6)
