"""The drop-in mirror of the reference's module interface (multimodal_sequencing_b200/dropin): same import
paths, constructors, state_dict keys (CPU checks) and — on the GPU — the reference's own call sequence
(berson_pointer_network, and beam_search_pointer's step-by-step loop through model.step + Beam) reproducing
the reference-generated fixtures."""
import os
import types

import pytest
import torch

from multimodal_sequencing_b200 import dropin

dropin.install()
from models.beam import Beam  # noqa: E402
from models.berson import BertConfig, BertForOrdering, beam_search_pointer  # noqa: E402
from models.berson.modeling_bert import berson_pointer_network  # noqa: E402
from models.CLIP.src.lxrt.modeling import BertConfig as LxrtBertConfig  # noqa: E402
from models.CLIP.src.lxrt.modeling import LXRTModel  # noqa: E402
from oracle import berson_oracle as O  # noqa: E402

torch.set_grad_enabled(False)


class Tok:
    cls_token, sep_token, pad_token = "[CLS]", "[SEP]", "[PAD]"

    def convert_tokens_to_ids(self, t):
        return {"[CLS]": 101, "[SEP]": 102, "[PAD]": 0}[t]


def _args(N, W, ff, device="cpu", mm=False):
    return types.SimpleNamespace(ff_size=ff, heads=8, para_dropout=0.1, inter_layers=2, beam_size=W, pairwise_loss_lam=0.6,
                                 multimodal_loss=False, additional_wrapper_level_objectives=None, device=device,
                                 multimodal=mm, use_multimodal_model=False, multimodal_model_type="clip" if mm else None,
                                 multimodal_img_part=False, per_seq_max_length=64, max_story_length=N)


def _build(g, N, W, device="cpu", multimodal_loss=False):
    c = g["cfg"]
    cfg = BertConfig(c["vocab_size_or_config_json_file"], hidden_size=c["hidden_size"], num_hidden_layers=c["num_hidden_layers"],
                     num_attention_heads=c["num_attention_heads"], intermediate_size=c["intermediate_size"],
                     max_position_embeddings=c["max_position_embeddings"])
    mm = g.get("vit") is not None or g.get("rn") is not None
    args = _args(N, W, g["ff_size"], device, mm)
    if multimodal_loss:
        args.multimodal_loss = True
        cfg.v_feature_size = c["hidden_size"]   # train.py:2022 hard-wires 1024 (= RoBERTa-large's H); the tiny model's H here
    if mm:
        inner = LXRTModel(LxrtBertConfig(c["vocab_size_or_config_json_file"], hidden_size=c["hidden_size"],
                                         num_hidden_layers=c["num_hidden_layers"], num_attention_heads=c["num_attention_heads"],
                                         intermediate_size=c["intermediate_size"],
                                         max_position_embeddings=c["max_position_embeddings"]),
                          clip_model_name="RN50" if g.get("rn") else "ViT-B/32", clip_config=g.get("rn") or g["vit"],
                          cls_id=101, sep_id=102, max_story_length=N)
        model = BertForOrdering(cfg, args, tokenizer=Tok())
        model.bert = inner
    else:
        model = BertForOrdering(cfg, args, tokenizer=None)
    return model, args


@pytest.mark.gpu
def test_reference_training_loop_with_the_resnet_tower(golden_dir):
    """trainers/train.py:340-363 on the drop-in module with the reference's wired backbone (ModifiedResNet, train() mode =
    BatchNorm over the batch): loss and gradient norms of the first step equal the reference-generated fixture
    (grads_tiny.pt 'mm_rn_bntrain'), the BatchNorm buffers of the module move like nn.BatchNorm2d's, the loop lowers the loss."""
    g = torch.load(os.path.join(golden_dir, "mm_rn_tiny.pt"), weights_only=False)
    r = torch.load(os.path.join(golden_dir, "grads_tiny.pt"), weights_only=False)["mm_rn_bntrain"]
    ids, labels, images = O.synthetic_manuals(r["B"], r["N"], r["L"], vocab=1000, image_px=224, seed=r["seed"])
    model, args = _build(g, r["N"], 4, "cuda")
    model.config.hidden_dropout_prob = model.config.attention_probs_dropout_prob = 0.0
    model.bert.config.hidden_dropout_prob = model.bert.config.attention_probs_dropout_prob = 0.0
    args.para_dropout = 0.0
    model.load_state_dict(g["sd"], strict=False)
    model = model.cuda().train()
    for mod in model.modules():
        mod.precise = True
    inp = O.prepare_inputs(ids, labels, r["N"], images)
    inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inp.items()}
    v = "bert.encoder.visual_model.visual."
    rm0 = model.state_dict()[v + "bn1.running_mean"].clone()
    nbt0 = int(model.state_dict()[v + "bn1.num_batches_tracked"]) if (v + "bn1.num_batches_tracked") in model.state_dict() else None
    opt = torch.optim.AdamW([p for p in model.parameters()], lr=1e-3, weight_decay=0.0)
    losses = []
    with torch.enable_grad():
        for step in range(3):
            loss = model._forward(**inp)[0]
            loss.backward()
            if step == 0:
                assert abs(loss.item() - r["loss"]) < 5e-5
                named = dict(model.named_parameters())
                for n, s_ in r["grads"].items():
                    if n.startswith("bert.pooler"):
                        continue
                    a = named[n].grad.detach().double().cpu().reshape(-1)
                    assert abs(float(a.norm()) - s_["norm"]) <= 1e-2 * s_["norm"] + 1e-7, n
                assert (model.state_dict()[v + "bn1.running_mean"] - rm0).abs().max() > 1e-5   # nn.BatchNorm2d.train() bookkeeping
                if nbt0 is not None:
                    assert int(model.state_dict()[v + "bn1.num_batches_tracked"]) == nbt0 + 1
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            model.zero_grad()
            losses.append(loss.item())
    assert losses[-1] < losses[0] - 0.01, losses


@pytest.mark.parametrize("name", ["text_tiny.pt", "mm_tiny.pt", "mm_rn_tiny.pt"])
def test_state_dict_keys_match_reference(golden_dir, name):
    g = torch.load(os.path.join(golden_dir, name), weights_only=False)
    model, _ = _build(g, 5, 4)
    missing, unexpected = model.load_state_dict(g["sd"], strict=False)
    assert not unexpected, unexpected
    assert all("proj" in k or "box_" in k for k in missing), missing  # tensors make_golden did not need to save
    for k, v in g["sd"].items():
        assert model.state_dict()[k].shape == v.shape, k


def test_beam_matches_generator_semantics():
    prev = Beam(4)
    prev.candidates, prev.scores = [[]], [0]
    done, remain = Beam(4).step(torch.tensor([[0.3, 0.1, 0.2]]), prev, lambda c: len(c) == 2)
    assert done == [] and remain == [0, 0, 0]


@pytest.mark.gpu
def test_reference_training_loop_runs_unchanged(golden_dir):
    """trainers/train.py:340-363 against the drop-in module: loss = model(inputs)[0]; loss.backward();
    clip_grad_norm_(model.parameters()); optimizer.step(); model.zero_grad().  p.grad must equal the gradients the real
    reference produced (grads_tiny.pt), and the loop must lower the loss."""
    g = torch.load(os.path.join(golden_dir, "text_tiny.pt"), weights_only=False)
    r = torch.load(os.path.join(golden_dir, "grads_tiny.pt"), weights_only=False)["text"]
    from oracle import berson_oracle as O
    ids, labels, _ = O.synthetic_manuals(r["B"], r["N"], r["L"], vocab=1000, seed=r["seed"])
    model, args = _build(g, 5, 4, "cuda")
    # the reference's gradient fixture was produced without dropout: switch it off the way a user would (config / args)
    model.config.hidden_dropout_prob = model.config.attention_probs_dropout_prob = 0.0
    args.para_dropout = 0.0
    model.load_state_dict(g["sd"], strict=False)
    model = model.cuda().train()
    for mod in model.modules():
        mod.precise = True
    model.tokenizer = Tok()
    inputs = {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": labels}
    opt = torch.optim.AdamW([p for p in model.parameters()], lr=1e-3, weight_decay=0.0)
    losses = []
    with torch.enable_grad():
        for step in range(3):
            loss = model(inputs)[0]
            loss.backward()
            if step == 0:
                assert abs(loss.item() - r["loss"]) < 5e-5
                named = dict(model.named_parameters())
                for n, s_ in r["grads"].items():
                    if n.startswith("bert.pooler"):
                        continue
                    a = named[n].grad.detach().double().cpu().reshape(-1)
                    assert abs(float(a.norm()) - s_["norm"]) <= 1e-3 * s_["norm"] + 1e-7, n
                for n in ("two_level_encoder.h1_relationship.weight", "classifier.weight", "encoder.transformer_inter.0.layer_norm.weight"):
                    assert named[n].grad is None, n      # no gradient in the reference either
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            model.zero_grad()
            losses.append(loss.item())
    assert losses[-1] < losses[0] - 0.02, losses
    # fused device-side step: same semantics without the weight round trip
    l0 = model.finetune_step(inputs, 1e-3).item()
    l1 = model.finetune_step(inputs, 1e-3).item()
    assert l1 < l0
    model.pull_weights()
    model.eval()
    assert abs(model(inputs)[0].item() - model.engine().training_loss(model.engine().prepare(ids, labels, 5, None)).item()) < 1e-6


@pytest.mark.gpu
def test_training_loop_with_data_updating_optimizer(golden_dir):
    """The reference's optimizers (transformers.AdamW 3.4, trainers/train.py:185; models/berson/optimization.py:176,187) update
    weights through `p.data.addcdiv_` / `p.data.add_`, which does NOT bump `p._version`.  The drop-in must still see every
    update (it refreshes its packed model in place on each train-mode forward) -- and must not rebuild the model per step."""
    import time
    g = torch.load(os.path.join(golden_dir, "text_tiny.pt"), weights_only=False)
    r = torch.load(os.path.join(golden_dir, "grads_tiny.pt"), weights_only=False)["text"]
    from oracle import berson_oracle as O
    ids, labels, _ = O.synthetic_manuals(r["B"], r["N"], r["L"], vocab=1000, seed=r["seed"])
    model, args = _build(g, 5, 4, "cuda")
    model.load_state_dict(g["sd"], strict=False)
    model = model.cuda().train()
    for mod in model.modules():
        mod.precise = True
    model.tokenizer = Tok()
    inputs = {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": labels}
    params = [p for p in model.parameters()]
    state = {id(p): (torch.zeros_like(p), torch.zeros_like(p)) for p in params}

    def data_adamw_step(step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-6):
        # the update rule of the vendored AdamW, written as that file writes it: in place on p.data
        for p in params:
            if p.grad is None:
                continue
            m, v = state[id(p)]
            m.mul_(b1).add_(p.grad.data, alpha=1 - b1)
            v.mul_(b2).addcmul_(p.grad.data, p.grad.data, value=1 - b2)
            step_size = lr * (1 - b2 ** step) ** 0.5 / (1 - b1 ** step)
            p.data.addcdiv_(m, v.sqrt().add_(eps), value=-step_size)

    versions = [p._version for p in params]
    losses = []
    with torch.enable_grad():
        for step in range(1, 5):
            loss = model(inputs)[0]
            loss.backward()
            data_adamw_step(step)
            model.zero_grad()
            losses.append(loss.item())
    assert [p._version for p in params] == versions, "the test optimizer must not bump version counters"
    assert losses[-1] < losses[0] - 0.02, "stale packed weights: the loss does not move (%s)" % losses
    assert all(abs(a - b) > 1e-7 for a, b in zip(losses, losses[1:])), losses
    assert model.__dict__["_eng_builds"] == 1, "the packed model was rebuilt %d times" % model.__dict__["_eng_builds"]
    assert model.__dict__["_eng_uploads"] >= 3
    # eval after the last optimizer step: the first eval-mode call must see that step too
    model.eval()
    ev = model(inputs)[0].item()
    eng = model.engine()
    named = dict(model.named_parameters())
    for n, _, _, _ in eng.train_layout()[:40:7]:
        assert torch.equal(eng.read_param(n, tuple(named[n].shape)), named[n].data), n
    assert abs(ev - eng.training_loss(eng.prepare(ids, labels, 5, None)).item()) < 1e-6
    ups = model.__dict__["_eng_uploads"]
    model(inputs)
    model(inputs)
    assert model.__dict__["_eng_uploads"] == ups, "an eval-only model must not re-upload its weights"
    # cost of the in-place refresh on this (tiny) model: it must be a refresh, not a rebuild
    model.train()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        model.engine()
    torch.cuda.synchronize()
    per = (time.perf_counter() - t0) / 5
    print("in-place weight refresh (tiny model): %.2f ms per train-mode forward" % (per * 1e3))
    assert model.__dict__["_eng_builds"] == 1


@pytest.mark.gpu
def test_validation_loss_matches_reference(golden_dir):
    """BertForOrdering._forward loss value (modeling_bert.py:943-1174) against the reference's own number."""
    g = torch.load(os.path.join(golden_dir, "text_tiny.pt"), weights_only=False)
    lc = g["loss_case"]
    model, args = _build(g, 5, 4, "cuda")
    model.load_state_dict(g["sd"], strict=False)
    model = model.cuda().eval()
    for mod in model.modules():
        mod.precise = True
    model.tokenizer = Tok()
    (loss,) = model({"input_ids": lc["ids"], "attention_mask": torch.ones_like(lc["ids"]), "labels": lc["labels"]})
    assert abs(loss.item() - lc["loss"].item()) < 1e-5, (loss.item(), lc["loss"].item())
    model.precise = False   # bf16 encoder: same loss within the bf16 bound
    (loss16,) = model({"input_ids": lc["ids"], "attention_mask": torch.ones_like(lc["ids"]), "labels": lc["labels"]})
    assert abs(loss16.item() - lc["loss"].item()) < 2e-2


def test_cal_result_matches_reference_formula():
    from models.berson.eval import cal_result
    truth = [[0, 1, 2, 3, 4], [2, 0, 1, 4, 3], [4, 3, 2, 1, 0]]
    pred = [[0, 1, 2, 3, 4], [0, 2, 1, 4, 3], [0, 1, 2, 3, 4]]
    got = cal_result(truth, pred)
    want = O.cal_result(truth, pred)
    assert all(abs(a - b) < 1e-12 for a, b in zip(got, want))
    assert got[1] == 1 / 3 and abs(got[2] - (1 + 0.8 - 1) / 3) < 1e-12


@pytest.mark.gpu
def test_berson_evaluate_batched(golden_dir, tmp_path):
    """berson_evaluate (models/berson/eval.py:39) with a DataLoader batch of 3 manuals: artefacts + metrics."""
    from models.berson.eval import berson_evaluate
    g = torch.load(os.path.join(golden_dir, "text_tiny.pt"), weights_only=False)
    cases = [c for c in g["cases"] if c["N"] == 5 and c["W"] == 4 and c["kind"] == "full"]
    model, args = _build(g, 5, 4, "cuda")
    model.load_state_dict(g["sd"], strict=False)
    model = model.cuda().eval()
    for mod in model.modules():
        mod.precise = True
    ids = torch.cat([c["ids"] for c in cases])
    labels = torch.cat([c["labels"] for c in cases])
    ds = torch.utils.data.TensorDataset(ids, torch.ones_like(ids), torch.zeros_like(ids), labels, torch.arange(len(cases)))
    args.task_names, args.output_dir, args.per_gpu_eval_batch_size, args.n_gpu = ["wikihow"], str(tmp_path), 3, 1
    args.local_rank, args.max_eval_steps = -1, 0
    res = berson_evaluate(args, model, lambda *a, **k: ds, Tok())
    want = O.cal_result(labels.tolist(), [c["perm"] for c in cases])
    assert abs(res["acc_dev"] - want[0]) < 1e-12 and res["pmr_dev"] == want[1] and abs(res["taus_dev"] - want[2]) < 1e-12
    lines = open(os.path.join(str(tmp_path), "output_order.txt")).read().strip().split("\n")
    assert lines == ["%s|||%s" % (" ".join(map(str, c["perm"])), " ".join(map(str, c["labels"][0].tolist()))) for c in cases]


def _reference_style_search(args, model, berson_inputs):
    """beam_search_pointer's loop exactly as the reference writes it (modeling_bert.py:1429-1552), driving
    OUR model.encode / model.step / Beam one step at a time."""
    bi = {k: v for k, v in berson_inputs.items() if k not in ("cuda", "_pair_batch")}
    (sentences, _, dec_init, original_keys, _, cls_mat, _, score_mat, _, _) = model.encode(**bi)
    num_sen = int(bi["passage_length"][0])
    sentences, original_keys = sentences[:, :num_sen, :], original_keys[:, :num_sen, :]
    document = sentences.squeeze(0)
    T, H = document.size()
    dev = document.device
    rela_vec = model.rela_encode(cls_mat, score_mat).contiguous()
    hist_left1, hist_left2 = model.history_encode(cls_mat, score_mat, score_mat)
    eye_zeros = (1 - torch.eye(T, device=dev)).byte()
    W = args.beam_size
    prev_beam = Beam(W)
    prev_beam.candidates, prev_beam.scores = [[]], [0]
    target_t, valid_size, hyp_list, logps = T - 1, W, [], []
    f_done = lambda x: len(x) == target_t
    for t in range(target_t):
        candidates = prev_beam.candidates
        if t == 0:
            dec_input = sentences.new_zeros(1, 1, H)
            pointed_mask = sentences.new_zeros(1, T).byte()
            rela_mask = eye_zeros.unsqueeze(0).clone()
            l1_mask, l2_mask = torch.zeros_like(rela_mask), torch.zeros_like(rela_mask)
        else:
            index = torch.tensor([c[-1] for c in candidates], device=dev)
            dec_input = document[index].unsqueeze(1)
            ar = torch.arange(index.size(0), device=dev)
            pointed_mask[ar, index] = 1
            rela_mask[ar, :, index] = 0
            rela_mask[ar, index] = 0
            l1_mask, l2_mask = torch.zeros_like(rela_mask), torch.zeros_like(rela_mask)
            l1_mask[ar, index, :] = 1
            if t > 1:
                l2_mask[ar, torch.tensor([c[-2] for c in candidates], device=dev), :] = 1
        dec_h, dec_c, log_prob = model.step(dec_input, dec_init, original_keys, pointed_mask, rela_vec, rela_mask,
                                            hist_left1, hist_left2, l1_mask, l2_mask)
        logps.append(log_prob.cpu())
        next_beam = Beam(valid_size)
        done_list, remain_list = next_beam.step(-log_prob, prev_beam, f_done)
        hyp_list.extend(done_list)
        valid_size -= len(done_list)
        if valid_size == 0:
            break
        ix = torch.tensor(remain_list, device=dev)
        dec_init = (dec_h.index_select(1, ix), dec_c.index_select(1, ix))
        pointed_mask, rela_mask = pointed_mask.index_select(0, ix), rela_mask.index_select(0, ix)
        rela_vec = rela_vec.index_select(0, ix).contiguous()
        hist_left1, hist_left2 = hist_left1.index_select(0, ix), hist_left2.index_select(0, ix)
        prev_beam = next_beam
    best = sorted(hyp_list, key=lambda h: h[1])[0][0]
    best = best + [sorted(set(range(T)) - set(best))[0]]
    return best, logps


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["text_tiny.pt", "mm_tiny.pt", "mm_rn_tiny.pt"])
def test_dropin_reproduces_reference_fixtures(golden_dir, name):
    g = torch.load(os.path.join(golden_dir, name), weights_only=False)
    mm = g.get("vit") is not None or g.get("rn") is not None
    for c in g["cases"]:
        model, args = _build(g, c["N"], c["W"], "cuda")
        model.load_state_dict(g["sd"], strict=False)
        model = model.cuda().eval()
        for mod in model.modules():   # fp32 parity mode for every engine owner (wrapper, inner model, tower)
            mod.precise = True
        images = None
        if mm:
            _, _, images = O.synthetic_manuals(1, c["N"], c["L"], vocab=1000, image_px=224, seed=c["seed"])
        inputs = {"input_ids": c["ids"], "attention_mask": torch.ones_like(c["ids"]), "labels": c["labels"]}
        if mm:
            inputs["images"] = images
        # (1) the reference's top-level call
        assert berson_pointer_network(args, model, Tok(), dict(inputs)) == c["perm"]
        # (2) the reference's own step-by-step loop through model.step + Beam
        from models.berson.process_inputs_for_berson import prepare_berson_inputs
        best, logps = _reference_style_search(args, model, prepare_berson_inputs(dict(inputs), Tok(), args=args))
        assert best == c["perm"]
        for lp, s in zip(logps, c["steps"]):
            live = s["logp"] > -1e8
            assert (lp[live] - s["logp"][live]).abs().max() < 4e-5
        # (3) inner model stand-alone, same signature as LXRTModel.forward / BertModel.forward
        pb = prepare_berson_inputs(dict(inputs), Tok(), args=args)
        B, P, Lt = pb["input_ids"].shape
        ids2, tt2, am2 = (pb[k].reshape(B * P, Lt) for k in ("input_ids", "token_type_ids", "attention_mask"))
        if mm:
            im = pb["images"].reshape(B * P * 2, 3, 224, 224)
            (lang, visn), pooled = model.bert(ids2, token_type_ids=tt2, attention_mask=am2, visual_feats=im)
            assert (lang[:3].cpu() - c["lang"]).abs().max() < 4e-5 and (pooled.cpu() - c["pooled"]).abs().max() < 4e-5
            if g.get("rn"):
                tower = model.bert.encoder.visual_model.visual(im[:6], img_len=2)
            else:
                tower = model.bert.encoder.visual_model.visual(im[:6], skip_last_layer=True, img_len=2)
            assert (tower.cpu() - c["tower"]).abs().max() < 4e-5
        else:
            seq, pooled = model.bert(ids2, attention_mask=am2, token_type_ids=tt2)
            assert (pooled.cpu() - c["enc"]["cls"]).abs().max() < 4e-5


def test_topological_sort_graph_matches_reference(golden_dir):
    """models/topological.py::Graph against trainers/topological_sort.py::Graph outputs (random tournaments, cycles,
    assert_head) recorded by tests/golden/make_golden.py --only-graph."""
    import json
    from models.topological import Graph, debatch_stories
    cases = json.load(open(os.path.join(golden_dir, "graph_cases.json")))
    assert len(cases) == 60
    for c in cases:
        g = Graph(c["n"])
        for u, v in c["edges"]:
            g.addEdge(u, v)
        try:
            order = g.topologicalSort(assert_head=c["head"]) if c["head"] is not None else g.topologicalSort()
        except AssertionError:
            order = "assert"
        assert order == c["order"], c
    assert debatch_stories([["a0", "b0"], ["a1", "b1"], ["a2", "b2"]]) == [["a0", "a1", "a2"], ["b0", "b1", "b2"]]


class TopoStubTokenizer:
    """Same stand-in as tests/golden/make_golden.py: a text is a string of token ids; <s>=0, </s>=2, <pad>=1."""

    def __call__(self, texts, max_length=None, padding=None, truncation=None, **kw):
        rows = []
        for t in texts:
            ids = ([0] + [int(x) for x in t.split()])[:max_length - 1] + [2]
            rows.append(ids + [1] * (max_length - len(ids)))
        return {"input_ids": rows}


@pytest.mark.gpu
def test_topological_inference_matches_reference(golden_dir):
    """trainers/eval.py::topological_inference: predicted orders and every pairwise logit against what the reference's
    own function produced with the reference LXRTModel (tests/golden/topo_inference_tiny.pt); here all C(N,2) pairs of a
    story go through the device in one call."""
    from models.topological import topological_inference
    mm = torch.load(os.path.join(golden_dir, "mm_tiny.pt"), weights_only=False)
    t = torch.load(os.path.join(golden_dir, "topo_inference_tiny.pt"), weights_only=False)
    c = mm["cfg"]
    model = LXRTModel(LxrtBertConfig(c["vocab_size_or_config_json_file"], hidden_size=c["hidden_size"],
                                     num_hidden_layers=c["num_hidden_layers"], num_attention_heads=c["num_attention_heads"],
                                     intermediate_size=c["intermediate_size"], max_position_embeddings=c["max_position_embeddings"]),
                      clip_model_name="ViT-B/32", clip_config=mm["vit"], cls_id=101, sep_id=102, max_story_length=5, num_labels=2)
    sd = {k[len("bert."):]: v for k, v in mm["sd"].items() if k.startswith("bert.")}
    sd.update(t["classifier"])
    assert not model.load_state_dict(sd, strict=False).unexpected_keys
    model = model.cuda().eval()
    model.precise = True
    # the generator drew the texts from the same stream first; re-draw them to land on the same image bits
    g = torch.Generator().manual_seed(t["image_seed"])
    for _ in range(t["N"]):
        for _ in range(t["B"]):
            torch.randint(10, 1000, (int(torch.randint(4, 12, (1,), generator=g)),), generator=g)
    images = torch.randn(t["B"], t["N"], 3, 224, 224, generator=g)
    assert abs(float(images.double().sum()) - t["image_checksum"]) < 1e-6, "torch RNG drift"
    args = types.SimpleNamespace(**t["args"])
    args.device = "cuda"
    logits = []
    fwd = model.forward

    def spy(*a, **k):
        out = fwd(*a, **k)
        logits.append(out[0].detach().float().cpu())
        return out

    model.forward = spy
    preds, loss = topological_inference(args, model, t["seqs"], TopoStubTokenizer(), images=images)
    assert len(logits) == t["B"], "one device call per story, not one per pair"
    assert (torch.cat(logits) - t["logits"]).abs().max() < 8e-6
    assert preds == t["preds"] and loss == t["loss"]


@pytest.mark.gpu
def test_lxrt_topo_sort_classifier_mode(golden_dir):
    """LXRTModel(..., num_labels=2) as the pairwise classifier of trainers/eval.py:topological_inference
    (lxrt/modeling.py:1502-1511, 1586-1594): logits against the reference's."""
    mm = torch.load(os.path.join(golden_dir, "mm_tiny.pt"), weights_only=False)
    t = torch.load(os.path.join(golden_dir, "topo_tiny.pt"), weights_only=False)
    c = mm["cfg"]
    model = LXRTModel(LxrtBertConfig(c["vocab_size_or_config_json_file"], hidden_size=c["hidden_size"],
                                     num_hidden_layers=c["num_hidden_layers"], num_attention_heads=c["num_attention_heads"],
                                     intermediate_size=c["intermediate_size"], max_position_embeddings=c["max_position_embeddings"]),
                      clip_model_name="ViT-B/32", clip_config=mm["vit"], cls_id=101, sep_id=102, max_story_length=5, num_labels=2)
    sd = {k[len("bert."):]: v for k, v in mm["sd"].items() if k.startswith("bert.")}
    sd.update(t["classifier"])
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("proj" in k or "box_" in k for k in missing), (missing, unexpected)
    model = model.cuda().eval()
    model.precise = True
    ids, labels, images = O.synthetic_manuals(1, 5, 16, vocab=1000, image_px=224, seed=t["seed"])
    assert abs(float(images.double().sum()) - t["image_checksum"]) < 1e-6
    inp = O.prepare_inputs(ids, labels, 5, images)
    R = t["R"]
    (logits,) = model(inp["input_ids"][0, :R].cuda(), attention_mask=inp["attention_mask"][0, :R].cuda(),
                      token_type_ids=inp["token_type_ids"][0, :R].cuda(), visual_feats=inp["images"][0, :R].cuda())
    assert (logits.cpu() - t["logits"]).abs().max() < 2e-5
    loss, _ = model(inp["input_ids"][0, :R].cuda(), attention_mask=inp["attention_mask"][0, :R].cuda(),
                    token_type_ids=inp["token_type_ids"][0, :R].cuda(), visual_feats=inp["images"][0, :R].cuda(),
                    labels=torch.zeros(R, dtype=torch.long))
    assert abs(loss.item() - torch.nn.functional.cross_entropy(t["logits"], torch.zeros(R, dtype=torch.long)).item()) < 2e-5


def test_pointer_module_state_dict_keys(golden_dir):
    from models.pointer_module import PointerOutput
    g = torch.load(os.path.join(golden_dir, "pointer_p1.pt"), weights_only=False)
    cfg = types.SimpleNamespace(hierarchical_version="p1", hidden_size=g["H"], max_story_length=g["N"], hl_include_objectives=None,
                                cls_id=101)
    m = PointerOutput(cfg)
    keys = set(m.state_dict().keys())
    assert set(g["sd"].keys()) <= keys and {k.replace("lstm_decoder.", "lstm_pointer.decoder.") for k in g["sd"]} <= keys
    with pytest.raises(NotImplementedError):
        PointerOutput(types.SimpleNamespace(hierarchical_version="p0", hidden_size=768, max_story_length=5))


@pytest.mark.gpu
def test_pointer_module_p1_matches_reference(golden_dir):
    """PointerOutput.forward / LSTMPointerModule (models/pointer_module.py:153-576, 690-749) vs the reference's outputs."""
    from models.pointer_module import PointerOutput
    g = torch.load(os.path.join(golden_dir, "pointer_p1.pt"), weights_only=False)
    cfg = types.SimpleNamespace(hierarchical_version="p1", hidden_size=g["H"], max_story_length=g["N"], hl_include_objectives=None,
                                cls_id=101)
    m = PointerOutput(cfg)
    m.load_state_dict(g["sd"], strict=False)
    m = m.cuda().eval()
    loss, out = m({"input_ids": g["ids"].cuda(), "labels": g["labels"].cuda()}, g["seq"].cuda())
    assert torch.equal(out.cpu(), g["outputs"])                      # greedy picks: bit-exact index work
    assert abs(loss.item() - g["loss"].item()) < 1e-5 * max(1.0, abs(g["loss"].item()))
    (out2,) = m({"input_ids": g["ids"].cuda()}, g["seq"].cuda())
    assert torch.equal(out2.cpu(), g["outputs"])


def test_train_mode_autograd_bridge_host_logic(golden_dir):
    """BertForOrdering.forward in train() mode, host side only (a stand-in engine supplies a known flat gradient buffer):
    loss.backward() must hand every parameter its slice, scaled by the upstream gradient, leave parameters without a slot
    at grad None, and accumulate over two backward passes like autograd does."""
    g = torch.load(os.path.join(golden_dir, "text_tiny.pt"), weights_only=False)
    model, _ = _build(g, 5, 4)
    model.train()
    named = [(n, p) for n, p in model.named_parameters()]
    skip = {"classifier.weight", "classifier.bias", "two_level_encoder.h1_relationship.weight"}
    layout, off = [], 0
    for n, p in named:
        if n in skip:
            continue
        layout.append((n, off, p.numel(), True))
        off += (p.numel() + 63) // 64 * 64

    class FakeEngine:
        device = torch.device("cpu")

        def train_layout(self):
            return layout

        def new_grad_buffer(self):
            return torch.zeros(off)

        def train_step(self, pb, flat, lam, triplets=None):
            for i, (n, o, k, _) in enumerate(layout):
                flat[o:o + k] += float(i + 1)
            return torch.tensor(2.5)

        def set_dropout(self, p_hidden, p_attn, p_para, seed=0):
            dropped.append((p_hidden, p_attn, p_para))

    dropped = []
    fake = FakeEngine()
    model.engine = lambda: fake
    from oracle import berson_oracle as O
    ids, labels, _ = O.synthetic_manuals(1, 5, 8, vocab=1000, seed=2)
    bi = O.prepare_inputs(ids, labels, 5)
    with torch.enable_grad():
        loss = model._forward(**bi)[0]
        assert loss.requires_grad and abs(loss.item() - 2.5) < 1e-6
        (loss * 0.5).backward()
        loss2 = model._forward(**bi)[0]
        loss2.backward()
    # train(): the reference's dropout probabilities reach the engine, once (BertConfig 0.1 / 0.1, args.para_dropout 0.1)
    assert dropped == [(model.config.hidden_dropout_prob, model.config.attention_probs_dropout_prob, 0.1)], dropped
    idx = {n: i for i, (n, _, _, _) in enumerate(layout)}
    for n, p in named:
        if n in skip:
            assert p.grad is None, n
        else:
            assert p.grad.shape == p.shape and torch.all(p.grad == 1.5 * (idx[n] + 1)), n


def test_save_and_from_pretrained_round_trip(golden_dir, tmp_path):
    """trainers/train.py:404 save_pretrained(dir) and 2030-2035 from_pretrained(dir, inner_model=, tokenizer=, config=,
    load_inner_model=True, args=): weights and configuration survive the round trip, a bare-BertModel checkpoint loads into
    the model with heads (base_model_prefix rule), a shape mismatch raises like torch's loader does."""
    from models.berson.modeling_bert import BertModel
    g = torch.load(os.path.join(golden_dir, "text_tiny.pt"), weights_only=False)
    model, args = _build(g, 5, 4)
    model.load_state_dict(g["sd"], strict=False)
    d = tmp_path / "ckpt"
    d.mkdir()
    model.save_pretrained(str(d))
    assert (d / "pytorch_model.bin").exists() and (d / "config.json").exists()
    again = BertForOrdering.from_pretrained(str(d), args=args)              # config read back from config.json
    assert not again.training and again.config.hidden_size == model.config.hidden_size
    for k, v in model.state_dict().items():
        assert torch.equal(again.state_dict()[k], v), k
    again2, info = BertForOrdering.from_pretrained(str(d), config=model.config, args=args, output_loading_info=True)
    assert info["missing_keys"] == [] and info["unexpected_keys"] == []
    # bare inner model checkpoint -> model with heads, and the other way round
    inner_dir = tmp_path / "inner"
    inner_dir.mkdir()
    model.bert.save_pretrained(str(inner_dir))
    fresh = BertForOrdering.from_pretrained(str(inner_dir), config=model.config, args=args)
    assert torch.equal(fresh.bert.state_dict()["embeddings.word_embeddings.weight"], model.bert.state_dict()["embeddings.word_embeddings.weight"])
    bare = BertModel.from_pretrained(str(d), config=model.config)
    assert torch.equal(bare.state_dict()["encoder.layer.1.output.dense.weight"], model.state_dict()["bert.encoder.layer.1.output.dense.weight"])
    bad = dict(model.state_dict())
    bad["key_linear.weight"] = torch.zeros(3, 3)
    with pytest.raises(RuntimeError, match="size mismatch"):
        BertForOrdering.from_pretrained(None, config=model.config, state_dict=bad, args=args)
    with pytest.raises(EnvironmentError):
        BertForOrdering.from_pretrained(str(tmp_path / "nowhere"), config=model.config, args=args)


def test_lxrt_from_pretrained_prefix_rules(golden_dir, tmp_path):
    """LXRTModel.from_pretrained(path, **kwargs) (train.py:1869-1880; lxrt/modeling.py:1257-1430): config.json +
    pytorch_model.bin from a directory; `bert.`-prefixed checkpoints load into the prefix-less body."""
    g = torch.load(os.path.join(golden_dir, "mm_tiny.pt"), weights_only=False)
    model, _ = _build(g, 5, 4)
    model.load_state_dict(g["sd"], strict=False)
    kw = dict(clip_model_name="ViT-B/32", clip_config=g["vit"], cls_id=101, sep_id=102, max_story_length=5)
    d = tmp_path / "lxrt"
    d.mkdir()
    model.bert.save_pretrained(str(d))
    a = LXRTModel.from_pretrained(str(d), **kw)
    for k, v in model.bert.state_dict().items():
        assert torch.equal(a.state_dict()[k], v), k
    # a checkpoint written by the wrapper (keys "bert.*" + heads) loads into the bare body as well
    d2 = tmp_path / "wrapper"
    d2.mkdir()
    model.bert.save_pretrained(str(d2))
    torch.save(model.state_dict(), str(d2 / "pytorch_model.bin"))
    b = LXRTModel.from_pretrained(str(d2), **kw)
    assert torch.equal(b.state_dict()["encoder.layer.0.output.dense.weight"], model.state_dict()["bert.encoder.layer.0.output.dense.weight"])
    with pytest.raises(EnvironmentError):
        LXRTModel.from_pretrained(str(tmp_path / "missing"), **kw)


def _hf_tiny():
    import transformers
    cfg = transformers.BertConfig(vocab_size=1000, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512,
                                  max_position_embeddings=256, type_vocab_size=2, hidden_dropout_prob=0.1,
                                  attention_probs_dropout_prob=0.1)
    torch.manual_seed(11)
    return transformers.BertModel(cfg).eval(), cfg


def test_oracle_matches_huggingface_automodel_outputs():
    """trainers/train.py:1928-1933 builds the text-only inner encoder with transformers' AutoModel; BertForOrdering.encode takes
    outputs[0] / outputs[1] of it (modeling_bert.py:1308-1315), and outputs[1] of an HF BertModel is the TANH POOLER, not
    seq[:,0].  Pin the oracle's restatement of that (text_bert with cfg["cls_pooler"]) to the real HF module."""
    from oracle import berson_oracle as O
    hf, cfg = _hf_tiny()
    ids = torch.randint(1, 1000, (6, 23), generator=torch.Generator().manual_seed(0))
    am = torch.ones_like(ids)
    am[2, 15:] = 0
    tt = (torch.arange(23)[None] > 9).long().expand(6, -1).contiguous()
    out = hf(input_ids=ids, attention_mask=am, token_type_ids=tt)
    sd = {"bert." + k: v for k, v in hf.state_dict().items()}
    ocfg = dict(num_hidden_layers=2, num_attention_heads=2, cls_pooler=True)
    seq, pooled = O.text_bert(sd, ocfg, ids, am, tt)
    assert (seq - out.last_hidden_state).abs().max() < 2e-5
    assert (pooled - out.pooler_output).abs().max() < 2e-5
    assert (pooled - seq[:, 0]).abs().max() > 1e-2       # and it really differs from the CLS row


@pytest.mark.gpu
def test_bert_for_ordering_with_huggingface_inner_model():
    """BertForOrdering(config, args, inner_model=<HF AutoModel>, tokenizer=..., load_inner_model=True) as train.py:1928-1933 +
    2012-2022 assemble the text-only task: encode() and berson_pointer_network through the drop-in equal the oracle with the
    tanh-pooled CLS (and differ from the seq[:,0] variant)."""
    from oracle import berson_oracle as O
    hf, hcfg = _hf_tiny()
    cfg = BertConfig(1000, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512, max_position_embeddings=256)
    N, W = 5, 4
    args = _args(N, W, 256, "cuda")
    torch.manual_seed(5)
    model = BertForOrdering(cfg, args, inner_model=hf, tokenizer=Tok(), load_inner_model=True)
    assert model.bert is hf
    model = model.cuda().eval()
    for mod in model.modules():
        mod.precise = True
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ids, labels, _ = O.synthetic_manuals(3, N, 14, vocab=1000, seed=8)
    inp = O.prepare_inputs(ids, labels, N)
    ocfg = dict(num_hidden_layers=2, num_attention_heads=2, vit=None, cls_pooler=True)
    want = O.encode(sd, ocfg, inp)
    plain = O.encode(sd, dict(ocfg, cls_pooler=False), inp)
    eng = model.engine()
    got = eng.encode(eng.prepare(ids, labels, N), want_top_vec=True)
    for k in ("sents", "para", "h0", "key", "cls", "cls_mat", "cls_score", "score_mat", "his1", "his2", "top_vec"):
        a, b = got[k].reshape(want[k].shape).cpu(), want[k]
        assert (a - b).abs().max() <= 4e-5 * max(1.0, float(b.abs().max())), k
    assert (got["cls"].cpu().reshape(plain["cls"].shape) - plain["cls"]).abs().max() > 1e-2
    perms = eng.order(ids, labels, N, W)
    assert perms == [O.beam_search(sd, want, N, W, b) for b in range(3)]
    # and the module-level call sites of the reference: encode() / berson_pointer_network
    one = {"input_ids": ids[:1], "attention_mask": torch.ones_like(ids[:1]), "labels": labels[:1]}
    assert berson_pointer_network(args, model, Tok(), dict(one)) == perms[0]


@pytest.mark.gpu
def test_multimodal_loss_objective_through_the_module(golden_dir):
    """args.multimodal_loss (modeling_bert.py:897-898, 1218-1225, 1359-1364) on the drop-in module: img_projection exists under the
    reference's name, the training forward returns the loss with the image pairwise term and hands its gradients to p.grad
    (checked against autograd through the oracle, itself pinned live to the reference), the eval-mode loss agrees, encode()
    returns the reference's (sentences, visual tokens) / (text, image) score tuples and decoding is unchanged."""
    from oracle import train_oracle as TO
    g = torch.load(os.path.join(golden_dir, "mm_tiny.pt"), weights_only=False)
    N = 4
    ids, labels, images = O.synthetic_manuals(2, N, 12, vocab=1000, image_px=224, seed=57)
    model, args = _build(g, N, 4, "cuda", multimodal_loss=True)
    assert isinstance(model.img_projection, torch.nn.Linear) and "img_projection.weight" in model.state_dict()
    model.config.hidden_dropout_prob = model.config.attention_probs_dropout_prob = 0.0
    model.bert.config.hidden_dropout_prob = model.bert.config.attention_probs_dropout_prob = 0.0
    args.para_dropout = 0.0
    model.load_state_dict(g["sd"], strict=False)
    model = model.cuda().train()
    for mod in model.modules():
        mod.precise = True
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    ocfg = dict(num_hidden_layers=g["cfg"]["num_hidden_layers"], num_attention_heads=g["cfg"]["num_attention_heads"], vit=g["vit"])
    inp = O.prepare_inputs(ids, labels, N, images)
    oloss, ref = TO.loss_grads(sd, ocfg, inp, multimodal_loss=True)
    o0, _ = TO.loss_grads(sd, ocfg, inp)
    inputs = {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": labels, "images": images}
    with torch.enable_grad():
        loss = model(inputs)[0]
        loss.backward()
    assert abs(loss.item() - oloss) < 5e-5 and oloss - o0 > 1e-2, (loss.item(), oloss, o0)
    for n in ("img_projection.weight", "img_projection.bias", "two_level_encoder.pairwise_relationship.weight",
              "bert.encoder.visn_fc.visn_fc.weight"):
        got, want = dict(model.named_parameters())[n].grad.detach().cpu(), ref[n]
        assert float((got - want).norm() / want.norm()) < 5e-4, n
    model.eval()
    assert abs(model(inputs)[0].item() - oloss) < 5e-5
    cuda_inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inp.items()}
    enc = model.encode(**cuda_inp)
    (sents, visn), (score, score_img) = enc[0], enc[6]
    oe = O.encode(sd, ocfg, inp)
    assert visn.shape == oe["visn"].shape and (visn.cpu() - oe["visn"]).abs().max() < 1e-4
    want_img = O._lin(sd, "two_level_encoder.pairwise_relationship", O._lin(sd, "img_projection", oe["visn"][:, 0]))
    assert (score_img.cpu() - want_img).abs().max() < 1e-4 and (sents.cpu() - oe["sents"]).abs().max() < 1e-4
    got = [berson_pointer_network(args, model, Tok(), {k: v[b:b + 1] for k, v in inputs.items()}) for b in range(2)]
    assert got == O.order_manuals(sd, ocfg, ids, labels, N, 4, images)


def test_mixed_step_counts_in_one_batch_are_refused(golden_dir):
    """The B200 path batches manuals of ONE length (as the reference's loaders do); a mixed batch must fail loudly."""
    g = torch.load(os.path.join(golden_dir, "text_tiny.pt"), weights_only=False)
    model, _ = _build(g, 5, 4)
    ids = torch.zeros(2, 20, 16, dtype=torch.long)
    with pytest.raises(ValueError, match="different step counts"):
        model._pair_batch(ids, ids, ids, None, torch.tensor([5, 4]), None, None, None, None, None)
