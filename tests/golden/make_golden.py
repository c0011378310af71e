"""Generate the committed golden fixtures by running the REAL reference (/root/reference) on CPU.

    python tests/golden/make_golden.py          # writes tests/golden/*.pt

Fixtures (all fp32, produced by the unmodified reference code through the shims in ref_harness.py):
  text_tiny.pt    tiny text-only BERSON (vendored BertModel): state_dict, ragged + full inputs,
                  prepare_berson_inputs output, encode 10-tuple, per-step log-probs and beam
                  indices, final permutations, training loss.
  mm_tiny.pt      tiny LXRT + CLIP-ViT multimodal BERSON (ViT adapter of SURVEY §8(c)); images are
                  regenerated from the stored seed (a checksum guards RNG drift).
  decode_full.pt  full-width (H=768) decode stage: reference beam_search_pointer fed with seeded
                  synthetic `encode` outputs and seeded head weights, for
                  N in {5,6,10} x W in {1,4,8,16}: per-step beam indices / log-probs / permutation.
                  Weights and inputs are regenerated from seeds by oracle.synth (too big to commit).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_harness as rh  # noqa: E402
from oracle import berson_oracle as O  # noqa: E402
from oracle import synth  # noqa: E402

torch.set_grad_enabled(False)
torch.set_num_threads(1)  # deterministic summation order for the committed numbers

TINY = dict(vocab_size_or_config_json_file=1000, hidden_size=128, num_hidden_layers=2,
            num_attention_heads=2, intermediate_size=512, max_position_embeddings=256)
TINY_VIT = dict(embed_dim=64, image_resolution=224, vision_layers=2, vision_width=128, vision_patch_size=32)
DEAD = ("visual_model.transformer.", "visual_model.token_embedding", "visual_model.positional_embedding",
        "visual_model.ln_final", "visual_model.text_projection", "visual_model.logit_scale")


def traced_beam(ns, args, model, berson_inputs):
    """Run reference beam_search_pointer while recording model.step log-probs and Beam.step picks."""
    rec = []
    orig_step = model.step
    orig_bstep = ns.gen.Beam.step

    def step(*a, **k):
        h, c, logp = orig_step(*a, **k)
        rec.append(dict(logp=logp.clone()))
        return h, c, logp

    def bstep(self, prob, prev_beam, f_done):
        pre = prob.new_tensor(prev_beam.scores)
        score = prob + pre.unsqueeze(-1).expand_as(prob)
        k = min(self.beam_size, score.numel())
        s, ix = score.view(-1).topk(k, largest=False)
        rec[-1].update(score=s.clone(), beam_ix=(ix // prob.size(1)).clone(),
                       tok_ix=(ix - (ix // prob.size(1)) * prob.size(1)).clone())
        return orig_bstep(self, prob, prev_beam, f_done)

    model.step = step
    ns.gen.Beam.step = bstep
    try:
        perm = ns.berson.beam_search_pointer(args, model, **berson_inputs)
    finally:
        model.step = orig_step
        ns.gen.Beam.step = orig_bstep
    return perm, rec


ENC_NAMES = ["sents", "para", "hc", "key", "cls", "cls_mat", "cls_score", "score_mat", "his1", "his2"]


def run_case(ns, model, args, ids, labels, images):
    tok = rh.StubTokenizer()
    inputs = {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": labels}
    if images is not None:
        inputs["images"] = images
    bi = ns.prep.prepare_berson_inputs(inputs, tok, args=args)
    enc = model.encode(**bi)
    encd = {}
    for n, t in zip(ENC_NAMES, enc):
        if n == "hc":
            encd["h0"], encd["c0"] = t[0].clone(), t[1].clone()
        else:
            encd[n] = t.clone()
    bi = ns.prep.prepare_berson_inputs(inputs, tok, args=args)
    perm, rec = traced_beam(ns, args, model, bi)
    prep = {k: v.clone() for k, v in bi.items() if torch.is_tensor(v) and k != "images"}
    return dict(ids=ids.clone(), labels=labels.clone(), prep=prep, enc=encd, steps=rec, perm=perm)


def ragged_manual(n_steps, vocab, seed):
    g = torch.Generator().manual_seed(seed)
    rows = []
    for _ in range(n_steps):
        ln = int(torch.randint(5, 30, (1,), generator=g))
        rows.append(torch.cat([torch.tensor([101]), torch.randint(200, vocab, (ln,), generator=g), torch.tensor([102])]))
    ids = torch.cat(rows)[None]
    return ids, torch.randperm(n_steps, generator=g)[None]


def gen_text(ns):
    out = dict(cfg=dict(TINY), ff_size=256, cases=[])
    for N, W in ((5, 4), (5, 1), (6, 8), (10, 16)):
        args = rh.make_args(N, W)
        args.ff_size = 256
        model = rh.build_text_model(ns, TINY, args, seed=0)
        if "sd" not in out:
            out["sd"] = {k: v.clone() for k, v in model.state_dict().items()}
        ids, labels, _ = O.synthetic_manuals(2, N, 24, vocab=1000, seed=10 + N)
        for b in range(2):
            c = run_case(ns, model, args, ids[b:b + 1], labels[b:b + 1], None)
            c.update(N=N, W=W, kind="full")
            out["cases"].append(c)
        rid, rlab = ragged_manual(N, 1000, 20 + N)
        c = run_case(ns, model, args, rid, rlab, None)
        c.update(N=N, W=W, kind="ragged")
        out["cases"].append(c)
    # training loss (modeling_bert.py:943-1174), batch of 3 five-step manuals
    args = rh.make_args(5, 4)
    args.ff_size = 256
    model = rh.build_text_model(ns, TINY, args, seed=0)
    ids, labels, _ = O.synthetic_manuals(3, 5, 24, vocab=1000, seed=33)
    bi = ns.prep.prepare_berson_inputs({"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": labels},
                                       rh.StubTokenizer(), args=args)
    out["loss_case"] = dict(ids=ids, labels=labels, loss=model._forward(**bi)[0].clone())
    return out


def gen_mm(ns):
    out = dict(cfg=dict(TINY), vit=dict(TINY_VIT), ff_size=256, cases=[])
    for N, W, L in ((5, 4, 64), (6, 8, 16)):
        args = rh.make_args(N, W, multimodal=True)
        args.ff_size = 256
        model = rh.build_multimodal_model(ns, TINY, args, TINY_VIT, seed=0)
        if "sd" not in out:
            out["sd"] = {k: v.clone() for k, v in model.state_dict().items() if not any(d in k for d in DEAD)}
        seed = 40 + N
        ids, labels, images = O.synthetic_manuals(1, N, L, vocab=1000, image_px=224, seed=seed)
        c = run_case(ns, model, args, ids, labels, images)
        # intermediate tower / joint outputs of the inner model for stage-wise parity
        bi = ns.prep.prepare_berson_inputs({"input_ids": ids, "attention_mask": torch.ones_like(ids),
                                            "labels": labels, "images": images}, rh.StubTokenizer(), args=args)
        B, P, Lt = bi["input_ids"].shape
        im = bi["images"].reshape(B * P * 2, 3, 224, 224)
        tower = model.bert.encoder.visual_model.visual(im, skip_last_layer=True, img_len=2)
        (lang, visn), pooled = model.bert(input_ids=bi["input_ids"].reshape(B * P, Lt),
                                          token_type_ids=bi["token_type_ids"].reshape(B * P, Lt),
                                          attention_mask=bi["attention_mask"].reshape(B * P, Lt), visual_feats=im)
        c.update(N=N, W=W, L=L, seed=seed, image_checksum=float(images.double().sum()),
                 tower=tower[:3].clone(), lang=lang[:3].clone(), visn=visn[:3].clone(), pooled=pooled.clone())
        out["cases"].append(c)
    return out


TINY_RN = dict(embed_dim=256, image_resolution=224, vision_layers=(1, 2, 1, 1), vision_width=16)


def gen_mm_rn(ns):
    """The reference's wired default backbone ("RN50": ModifiedResNet + AttentionPool2d, visual_pos / visual_token_type
    embeddings) on a narrow tower -- same code path as RN50, width 16 instead of 64."""
    out = dict(cfg=dict(TINY), rn=dict(TINY_RN), ff_size=256, cases=[])
    for N, W, L in ((5, 4, 32),):
        args = rh.make_args(N, W, multimodal=True)
        args.ff_size = 256
        model = rh.build_multimodal_model(ns, TINY, args, seed=0, rn_cfg=TINY_RN)
        out["sd"] = {k: v.clone() for k, v in model.state_dict().items()
                     if not any(d in k for d in DEAD) and v.is_floating_point()}
        seed = 60 + N
        ids, labels, images = O.synthetic_manuals(1, N, L, vocab=1000, image_px=224, seed=seed)
        c = run_case(ns, model, args, ids, labels, images)
        bi = ns.prep.prepare_berson_inputs({"input_ids": ids, "attention_mask": torch.ones_like(ids),
                                            "labels": labels, "images": images}, rh.StubTokenizer(), args=args)
        B, P, Lt = bi["input_ids"].shape
        im = bi["images"].reshape(B * P * 2, 3, 224, 224)
        tower = model.bert.encoder.visual_model.visual(im[:6], img_len=2)
        (lang, visn), pooled = model.bert(input_ids=bi["input_ids"].reshape(B * P, Lt),
                                          token_type_ids=bi["token_type_ids"].reshape(B * P, Lt),
                                          attention_mask=bi["attention_mask"].reshape(B * P, Lt), visual_feats=im)
        c.update(N=N, W=W, L=L, seed=seed, image_checksum=float(images.double().sum()),
                 tower=tower[:3].clone(), lang=lang[:3].clone(), visn=visn[:3].clone(), pooled=pooled.clone())
        out["cases"].append(c)
    return out


def gen_decode_full(ns):
    H = 768
    out = dict(H=H, cases=[])
    full = dict(vocab_size_or_config_json_file=64, hidden_size=H, num_hidden_layers=1, num_attention_heads=12,
                intermediate_size=64, max_position_embeddings=8)
    for N in (5, 6, 10):
        for W in (1, 4, 8, 16):
            args = rh.make_args(N, W)
            args.ff_size = 64
            model = rh.build_text_model(ns, full, args, seed=0)
            heads = synth.decode_head_weights(H, seed=7)
            missing = model.load_state_dict(heads, strict=False)
            assert not missing.unexpected_keys
            enc = synth.synthetic_encode(N, H, seed=100 + N)
            tup = (enc["sents"], enc["para"], (enc["h0"], enc["c0"]), enc["key"], enc["cls"], enc["cls_mat"],
                   enc["cls_score"], enc["score_mat"], enc["his1"], enc["his2"])
            model.encode = lambda *a, _t=tup, **k: tuple(x.clone() if torch.is_tensor(x) else tuple(y.clone() for y in x) for x in _t)
            dummy = dict(input_ids=torch.zeros(1, N * (N - 1), 4, dtype=torch.long),
                         passage_length=torch.tensor([N]))
            perm, rec = traced_beam(ns, args, model, dummy)
            out["cases"].append(dict(N=N, W=W, enc_seed=100 + N, head_seed=7, perm=perm, steps=rec))
    return out


def gen_topo(ns, mm):
    """LXRTModel in topo-sort classifier mode (lxrt/modeling.py:1502-1511, 1586-1594; used by
    trainers/eval.py:topological_inference 509-520).  Inner weights = mm_tiny.pt's `bert.*`; only the
    classifier head and the reference logits are stored."""
    import contextlib
    import io
    torch.manual_seed(123)
    ns.fake_clip._vit_cfg = dict(TINY_VIT)
    ns.param.VISUAL_CONFIG.set_visual_dims(TINY_VIT["vision_width"], 4)
    ns.param.VISUAL_CONFIG.clip_model_name = "ViT-B/32"
    cfg = ns.lxrt.BertConfig(**TINY)
    cfg.classifier_dropout = None  # attribute newer transformers' RobertaClassificationHead reads
    with contextlib.redirect_stdout(io.StringIO()):
        m = ns.lxrt.LXRTModel(cfg, multimodal_text_part=False, multimodal_img_part=False, cls_id=101, sep_id=102,
                              max_story_length=5, clip_model_name="ViT-B/32", num_labels=2)
    m.encoder.skip_last_layer = True
    inner = {k[len("bert."):]: v for k, v in mm["sd"].items() if k.startswith("bert.")}
    res = m.load_state_dict(inner, strict=False)
    assert not res.unexpected_keys
    with torch.no_grad():
        m.classifier.out_proj.weight.normal_(0, 0.5)  # spread the logits (the 0.02 init gives ~1e-2 values)
    m.eval()
    seed, R = 61, 6
    ids, labels, images = O.synthetic_manuals(1, 5, 16, vocab=1000, image_px=224, seed=seed)
    inp = O.prepare_inputs(ids, labels, 5, images)
    out = m(inp["input_ids"][0, :R], attention_mask=inp["attention_mask"][0, :R], token_type_ids=inp["token_type_ids"][0, :R],
            visual_feats=inp["images"][0, :R])
    return dict(seed=seed, R=R, logits=out[0].clone(), image_checksum=float(images.double().sum()),
                classifier={k: v.clone() for k, v in m.state_dict().items() if k.startswith("classifier.")})


class TopoStubTokenizer:
    """Stands in for the HF tokenizer inside topological_inference: a "text" is a string of space-separated token ids;
    <s>=0 ... </s>=2, <pad>=1 (the function hard-codes the RoBERTa pad id, trainers/eval.py:466)."""

    def __call__(self, texts, max_length=None, padding=None, truncation=None, **kw):
        rows = []
        for t in texts:
            ids = ([0] + [int(x) for x in t.split()])[:max_length - 1] + [2]
            rows.append(ids + [1] * (max_length - len(ids)))
        return {"input_ids": rows}


def gen_topo_inference(ns, mm, topo):
    """trainers/eval.py::topological_inference (425-529) run on the REAL reference code: the function (and
    debatch_stories) are compiled from the reference file's own source, without importing the module (its imports need
    packages this image lacks), against the reference LXRTModel in classifier mode and the reference Graph."""
    import ast
    import contextlib
    import io
    import types
    import numpy as np
    from trainers.topological_sort import Graph
    src = open(os.path.join(rh.REF if hasattr(rh, "REF") else "/root/reference", "trainers", "eval.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("topological_inference", "debatch_stories")]
    glb = dict(np=np, torch=torch, Graph=Graph)
    exec(compile(ast.Module(body=keep, type_ignores=[]), "reference:trainers/eval.py", "exec"), glb)
    torch.manual_seed(123)
    ns.fake_clip._vit_cfg = dict(TINY_VIT)
    ns.param.VISUAL_CONFIG.set_visual_dims(TINY_VIT["vision_width"], 4)
    ns.param.VISUAL_CONFIG.clip_model_name = "ViT-B/32"
    cfg = ns.lxrt.BertConfig(**TINY)
    cfg.classifier_dropout = None
    with contextlib.redirect_stdout(io.StringIO()):
        m = ns.lxrt.LXRTModel(cfg, multimodal_text_part=False, multimodal_img_part=False, cls_id=101, sep_id=102,
                              max_story_length=5, clip_model_name="ViT-B/32", num_labels=2)
    m.encoder.skip_last_layer = True
    inner = {k[len("bert."):]: v for k, v in mm["sd"].items() if k.startswith("bert.")}
    inner.update(topo["classifier"])
    assert not m.load_state_dict(inner, strict=False).unexpected_keys
    m.eval()
    N, B = 5, 2
    g = torch.Generator().manual_seed(77)
    # seqs[j][b] = text of step j of story b (the DataLoader's collated layout, debatch_stories 1088-1097)
    seqs = [[" ".join(str(int(t)) for t in torch.randint(10, 1000, (int(torch.randint(4, 12, (1,), generator=g)),), generator=g))
             for _ in range(B)] for _ in range(N)]
    images = torch.randn(B, N, 3, 224, 224, generator=g)
    args = types.SimpleNamespace(multimodal=True, use_multimodal_model=False, use_cached=False, per_seq_max_length=16,
                                 max_seq_length=40, device="cpu", multimodal_text_part=False,
                                 include_num_img_regional_features=None, replace_token_type_embeddings=True,
                                 multimodal_model_type="clip")
    logits = []
    fwd = m.forward

    def spy(*a, **k):
        out = fwd(*a, **k)
        logits.append(out[0].detach().clone())
        return out

    m.forward = spy
    glb["topological_inference"](args, m, seqs, TopoStubTokenizer(), images=images)
    # a random-init classifier votes "ordered" for every pair; centre its decision boundary on the median margin so that
    # the fixture exercises both edge directions (and a cyclic tournament or two)
    d = torch.cat(logits)
    with torch.no_grad():
        ds = torch.sort(d[:, 1] - d[:, 0]).values
        m.classifier.out_proj.bias[0] += 0.5 * (ds[len(ds) // 2 - 1] + ds[len(ds) // 2])   # midpoint: no zero margins
    del logits[:]
    preds, loss = glb["topological_inference"](args, m, seqs, TopoStubTokenizer(), images=images)
    return dict(seqs=seqs, image_seed=77, image_checksum=float(images.double().sum()), preds=preds, loss=float(loss),
                logits=torch.cat(logits), args=vars(args), N=N, B=B,
                classifier={k: v.clone() for k, v in m.state_dict().items() if k.startswith("classifier.")})


def gen_graph_cases():
    """trainers/topological_sort.py::Graph on random tournaments (cycles included), with and without assert_head."""
    import json
    import random
    from trainers.topological_sort import Graph
    rnd = random.Random(5)
    cases = []
    for t in range(60):
        n = rnd.randint(2, 10)
        edges = []
        for i in range(n):
            for j in range(n):
                if i < j:
                    edges.append((i, j) if rnd.random() < 0.5 else (j, i))
        head = rnd.randrange(n) if t % 3 == 0 else None
        g = Graph(n)
        for u, v in edges:
            g.addEdge(u, v)
        try:
            order = g.topologicalSort(assert_head=head) if head is not None else g.topologicalSort()
        except AssertionError:
            order = "assert"
        cases.append(dict(n=n, edges=edges, head=head, order=order))
    with open(os.path.join(HERE, "graph_cases.json"), "w") as f:
        json.dump(cases, f)
    return cases


def gen_pointer_p1(ns):
    """models/pointer_module.py::PointerOutput, p1 variant (dead code in the reference):
    LSTMPointerModule greedy decoding + CE loss on random hidden states."""
    import types
    import transformers
    import transformers.file_utils as fu
    transformers.__dict__["AdamW"] = torch.optim.AdamW               # symbols of transformers 3.4 the module imports
    if not hasattr(fu, "requires_sklearn"):
        fu.requires_sklearn = lambda *a, **k: None
    if not hasattr(fu, "is_sklearn_available"):
        fu.is_sklearn_available = lambda: True
    from models import pointer_module as pm
    torch.manual_seed(11)
    H, N, B, L = 128, 5, 3, 40
    cfg = types.SimpleNamespace(hierarchical_version="p1", hidden_size=H, max_story_length=N, hl_include_objectives=None, cls_id=101)
    m = pm.PointerOutput(cfg).eval()
    for p_ in m.parameters():
        p_.data.mul_(3.0)   # spread the attention logits
    ids = torch.randint(200, 900, (B, L))
    for b in range(B):
        posn = torch.randperm(L - 1)[:N - 1].add(1).sort().values
        ids[b, 0] = 101
        ids[b, posn] = 101
    seq = torch.randn(B, L + 7, H)   # text + a few visual positions
    labels = torch.stack([torch.randperm(N) for _ in range(B)])
    loss, out = m({"input_ids": ids, "labels": labels}, seq)
    sd = {k: v.clone() for k, v in m.state_dict().items() if not k.startswith("lstm_pointer.")}
    return dict(H=H, N=N, sd=sd, ids=ids, seq=seq, labels=labels, loss=loss.clone(), outputs=out.clone())


def main():
    ns = rh.load()
    if "--only-graph" in sys.argv:
        print(len(gen_graph_cases()), "graph cases")
        return
    if "--only-topo-inference" in sys.argv:
        mm = torch.load(os.path.join(HERE, "mm_tiny.pt"), weights_only=False)
        topo = torch.load(os.path.join(HERE, "topo_tiny.pt"), weights_only=False)
        torch.save(gen_topo_inference(ns, mm, topo), os.path.join(HERE, "topo_inference_tiny.pt"))
        print("topo_inference_tiny.pt", os.path.getsize(os.path.join(HERE, "topo_inference_tiny.pt")) // 1024, "KiB")
        return
    if "--only-rn" in sys.argv:   # regenerate just the RN fixture (each generator seeds itself)
        torch.save(gen_mm_rn(ns), os.path.join(HERE, "mm_rn_tiny.pt"))
        print("mm_rn_tiny.pt", os.path.getsize(os.path.join(HERE, "mm_rn_tiny.pt")) // 1024, "KiB")
        return
    torch.save(gen_text(ns), os.path.join(HERE, "text_tiny.pt"))
    mm = gen_mm(ns)
    torch.save(mm, os.path.join(HERE, "mm_tiny.pt"))
    topo = gen_topo(ns, mm)
    torch.save(topo, os.path.join(HERE, "topo_tiny.pt"))
    torch.save(gen_topo_inference(ns, mm, topo), os.path.join(HERE, "topo_inference_tiny.pt"))
    gen_graph_cases()
    gen_graph_cases()
    torch.save(gen_mm_rn(ns), os.path.join(HERE, "mm_rn_tiny.pt"))
    torch.save(gen_decode_full(ns), os.path.join(HERE, "decode_full.pt"))
    torch.save(gen_pointer_p1(ns), os.path.join(HERE, "pointer_p1.pt"))
    for f in ("text_tiny.pt", "mm_tiny.pt", "topo_tiny.pt", "mm_rn_tiny.pt", "decode_full.pt", "pointer_p1.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
