"""ORACLE — test infrastructure only (never imported by the product path).

CPU restatement (plain torch fp32 on the host) of the reference's multimodal step-ordering hot
path, SURVEY.md §8(a) rows a1-a16/a19.  Every function cites the reference file:line it follows
(paths relative to telin0411/multimodal_sequencing).  Weights are handed over as ONE state_dict
with the reference's own key names, so the same tensors can be loaded into
the reference, this oracle and the CUDA path.

PARITY PIN: the reference's own tests do not cover this path, so this oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF, executed in the build container from
/root/reference (tests/golden/make_golden.py -> tests/golden/*.pt, and
tests/test_oracle_vs_reference.py which re-runs the reference live whenever it is present).
Third-party arithmetic (torch nn.Linear / nn.LSTM / nn.MultiheadAttention / LayerNorm / topk,
pinned torch==1.8.0 in the reference, 2.11 here) is restated with explicit formulas.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.
"""
import itertools
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------------


def _lin(sd, name, x, bias=True):
    w = sd[name + ".weight"]
    y = x @ w.t()
    if bias and (name + ".bias") in sd:
        y = y + sd[name + ".bias"]
    return y


def _ln(sd, name, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def gelu_erf(x):
    """models/berson/modeling_bert.py:125-131, models/CLIP/src/lxrt/modeling.py:116-122."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def gelu_tanh(x):
    """models/berson/neural.py:7-8."""
    return 0.5 * x * (1 + torch.tanh(math.sqrt(2 / math.pi) * (x + 0.044715 * torch.pow(x, 3))))


def quick_gelu(x):
    """models/CLIP/clip/model.py:199-201."""
    return x * torch.sigmoid(1.702 * x)


# --------------------------------------------------------------------------------------------
# a1 — host-side pair expansion (models/berson/process_inputs_for_berson.py)
# --------------------------------------------------------------------------------------------


def pairs_generator(n):
    """process_inputs_for_berson.py:246-261: all (i<j) lexicographic, then the same list as (j,i)."""
    one = [[a, b] for a, b in itertools.combinations(range(n), 2)]
    return one + [[b, a] for a, b in one], 2 * len(one)


def parse_steps(ids, cls_id, sep_id):
    """process_inputs_for_berson.py:100-110: split a flat id row into [CLS]..[SEP] steps."""
    ids = ids.tolist() if torch.is_tensor(ids) else list(ids)
    starts = [i for i, t in enumerate(ids) if t == cls_id]
    ends = [i for i, t in enumerate(ids) if t == sep_id]
    assert len(starts) == len(ends)
    return [ids[s:e + 1] for s, e in zip(starts, ends)]


def prepare_inputs(input_ids, labels, n_steps, images=None, cls_id=101, sep_id=102, pad_id=0):
    """process_inputs_for_berson.py:13-79 (+113-243, 264-368, 82-97).

    input_ids [B, L] int64 (steps concatenated), labels [B, N] int64, images [B, N, 3, S, S] or None.
    Returns the reference's dict (without the unused "cuda" string)."""
    B = len(input_ids)
    pairs, P = pairs_generator(n_steps)
    per = []
    for b in range(B):
        steps = parse_steps(input_ids[b], cls_id, sep_id)
        assert len(steps) == n_steps
        gt = [int(v) for v in labels[b]]
        rows, tts, seps, plab = [], [], [], []
        for i, j in pairs:
            # pairwise label: 1 iff ground-truth position of i precedes that of j (162-172)
            plab.append(1 if gt.index(i) < gt.index(j) else 0)
            s1, s2 = steps[i], steps[j]
            rows.append(s1 + s2)
            # token types 0/1; 0/0 when cls_id == 0 (RoBERTa) (203-208)
            tts.append([0] * len(s1) + [0 if cls_id == 0 else 1] * len(s2))
            seps.append([len(s1) - 1, len(s1) + len(s2) - 1])
        per.append((rows, tts, seps, plab, gt))
    Lt = max(len(r) for rows, *_ in per for r in rows)
    ids_o = torch.full((B, P, Lt), pad_id, dtype=torch.long)
    am_o = torch.full((B, P, Lt), pad_id, dtype=torch.long)  # padded with pad_id, as the reference (309)
    tt_o = torch.zeros((B, P, Lt), dtype=torch.long)
    for b, (rows, tts, seps, plab, gt) in enumerate(per):
        for p, (r, t) in enumerate(zip(rows, tts)):
            ids_o[b, p, :len(r)] = torch.tensor(r)
            am_o[b, p, :len(r)] = 1
            tt_o[b, p, :len(t)] = torch.tensor(t)
    out = {
        "input_ids": ids_o, "attention_mask": am_o, "token_type_ids": tt_o,
        "pairs_list": torch.tensor([pairs] * B, dtype=torch.long),
        "passage_length": torch.full((B,), n_steps, dtype=torch.long),
        "pairs_num": torch.full((B,), P, dtype=torch.long),
        "sep_positions": torch.tensor([x[2] for x in per], dtype=torch.long),
        "ground_truth": torch.tensor([x[4] for x in per], dtype=torch.long),
        "mask_cls": torch.ones((B, n_steps), dtype=torch.long),
        "pairwise_labels": torch.tensor([x[3] for x in per], dtype=torch.long),
    }
    if images is not None:
        # process_images (82-97): [img_i, img_j] stacked per pair -> [B, P, 2, 3, S, S]
        idx = torch.tensor(pairs)
        out["images"] = images[:, idx]  # [B, P, 2, 3, S, S]
    return out


# --------------------------------------------------------------------------------------------
# a3 / a7 / a8 — BERT embeddings, layer, text-only model
# --------------------------------------------------------------------------------------------


# training-mode dropout: None (eval) or an oracle.dropout.DropSpec -- a callable (x, kind, layer) -> x with the elements of
# that site dropped and the rest scaled by 1/(1-p).  Set by oracle.train_oracle around a training forward.
DROPOUT = None


def _drop(x, kind, layer=0):
    return x if DROPOUT is None else DROPOUT(x, kind, layer)


def _layer_index(pre):
    """'...encoder.layer.7.' / '...transformer_inter.1.' -> 7 / 1"""
    return int(pre.rstrip(".").rsplit(".", 1)[1])


def bert_embeddings(sd, pre, ids, tt, eps=1e-12):
    """lxrt/modeling.py:342-370 == models/berson/modeling_bert.py:148-180 (eval: dropout off)."""
    L = ids.shape[1]
    e = sd[pre + "word_embeddings.weight"][ids] + sd[pre + "position_embeddings.weight"][:L][None] \
        + sd[pre + "token_type_embeddings.weight"][tt]
    return _drop(_ln(sd, pre + "LayerNorm", e, eps), "E")


def _heads(x, h):
    B, L, H = x.shape
    return x.view(B, L, h, H // h).permute(0, 2, 1, 3)


def bert_layer(sd, pre, x, add_mask, heads, eps=1e-12, lxrt=True):
    """Post-LN BERT layer.  lxrt/modeling.py:373-507 (keys attention.self.* / attention.output.*),
    same math as models/berson/modeling_bert.py:183-337."""
    a = pre + "attention."
    q = _heads(_lin(sd, a + "self.query", x), heads)
    k = _heads(_lin(sd, a + "self.key", x), heads)
    v = _heads(_lin(sd, a + "self.value", x), heads)
    s = q @ k.transpose(-1, -2) / math.sqrt(q.shape[-1])
    if add_mask is not None:
        s = s + add_mask
    li = _layer_index(pre) if DROPOUT is not None else 0
    p = _drop(torch.softmax(s, dim=-1), "A", li)
    c = (p @ v).permute(0, 2, 1, 3).reshape(x.shape)
    x1 = _ln(sd, a + "output.LayerNorm", _drop(_lin(sd, a + "output.dense", c), "O", li) + x, eps)
    inter = gelu_erf(_lin(sd, pre + "intermediate.dense", x1))
    return _ln(sd, pre + "output.LayerNorm", _drop(_lin(sd, pre + "output.dense", inter), "F", li) + x1, eps)


def ext_mask(attention_mask):
    """(1 - mask) * -10000 broadcast to [R,1,1,L] (lxrt/modeling.py:1537-1545; modeling_bert.py:625-635)."""
    return (1.0 - attention_mask[:, None, None, :].float()) * -10000.0


def text_bert(sd, cfg, ids, attention_mask, tt, pre="bert."):
    """models/berson/modeling_bert.py:563-663: returns (sequence_output, seq[:,0]).
    cfg["cls_pooler"]: the inner model is a HuggingFace AutoModel, as trainers/train.py:1928-1933 builds it for the text-only
    task: its outputs[1] is pooler_output = tanh(pooler.dense(seq[:,0])) (transformers BertPooler), and that is what
    BertForOrdering.encode takes as cls_pooled_output (modeling_bert.py:1315)."""
    x = bert_embeddings(sd, pre + "embeddings.", ids, tt, cfg.get("layer_norm_eps", 1e-12))
    m = ext_mask(attention_mask)
    for i in range(cfg["num_hidden_layers"]):
        x = bert_layer(sd, pre + "encoder.layer.%d." % i, x, m, cfg["num_attention_heads"],
                       cfg.get("layer_norm_eps", 1e-12))
    if cfg.get("cls_pooler"):
        return x, torch.tanh(_lin(sd, pre + "pooler.dense", x[:, 0]))
    return x, x[:, 0]


# --------------------------------------------------------------------------------------------
# a4 — CLIP ViT tower with the reference's pair-joint token sequence
# --------------------------------------------------------------------------------------------


def vit_pair_tower(sd, pre, images, vit, img_len=2):
    """models/CLIP/clip/model.py:262-305 (+204-226, 190-201), skip_last_layer=True branch (oracle
    decision).  images [R*img_len, 3, S, S] -> [R, 1 + img_len*g*g, width].

    Quirks reproduced: one class token per PAIR; positional embedding cat(pos[0:g²+1], pos[0:g²])
    (271-275): the second image's patches reuse rows 0..g²-1."""
    W, patch, heads = vit["vision_width"], vit["vision_patch_size"], vit["vision_width"] // 64
    x = F.conv2d(images, sd[pre + "conv1.weight"], stride=patch)  # [R*il, W, g, g]
    g2 = x.shape[2] * x.shape[3]
    x = x.reshape(x.shape[0], W, g2).permute(0, 2, 1)
    R = x.shape[0] // img_len
    x = x.reshape(R, img_len * g2, W)
    cls = sd[pre + "class_embedding"][None, None, :].expand(R, 1, W)
    x = torch.cat([cls, x], dim=1)
    pos = sd[pre + "positional_embedding"]
    pos_all = torch.cat([pos] + [pos[:g2]] * (img_len - 1), dim=0)
    x = x + pos_all[None]
    x = _ln(sd, pre + "ln_pre", x, 1e-5)
    d = W // heads
    for l in range(vit["vision_layers"]):
        b = pre + "transformer.resblocks.%d." % l
        y = _ln(sd, b + "ln_1", x, 1e-5)
        qkv = y @ sd[b + "attn.in_proj_weight"].t() + sd[b + "attn.in_proj_bias"]
        q, k, v = (_heads(t, heads) for t in qkv.split(W, dim=-1))
        # nn.MultiheadAttention: q scaled by d**-0.5 before q@k^T, no mask
        s = (q * (d ** -0.5)) @ k.transpose(-1, -2)
        c = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(x.shape)
        x = x + _lin(sd, b + "attn.out_proj", c)
        y = _ln(sd, b + "ln_2", x, 1e-5)
        x = x + _lin(sd, b + "mlp.c_proj", quick_gelu(_lin(sd, b + "mlp.c_fc", y)))
    return _ln(sd, pre + "ln_post", x, 1e-5)


# --------------------------------------------------------------------------------------------
# a2 / a6 — LXRTModel.forward, CLIP / visualbert branch
# --------------------------------------------------------------------------------------------


def lxrt_forward(sd, cfg, ids, tt, attention_mask, images, pre="bert."):
    """lxrt/modeling.py:1513-1598 -> LXRTEncoder.forward 838-1107 (use_clip, visualbert_style).
    Returns (lang [R,Lt,H], visn [R,Lv,H], pooled [R,H])."""
    emb = bert_embeddings(sd, pre + "embeddings.", ids, tt, 1e-12)
    tower = vit_pair_tower(sd, pre + "encoder.visual_model.visual.", images, cfg["vit"])
    v = _drop(_ln(sd, pre + "encoder.visn_fc.visn_layer_norm", _lin(sd, pre + "encoder.visn_fc.visn_fc", tower), 1e-12), "V")
    R, Lt = ids.shape
    joint = torch.cat([emb, v], dim=1)
    m = torch.cat([ext_mask(attention_mask), torch.zeros(R, 1, 1, v.shape[1])], dim=-1)
    for i in range(cfg["num_hidden_layers"]):
        joint = bert_layer(sd, pre + "encoder.layer.%d." % i, joint, m, cfg["num_attention_heads"], 1e-12)
    lang, visn = joint[:, :Lt], joint[:, Lt:]
    pooled = _lin(sd, pre + "pooler.dense", lang[:, 0])
    return lang, visn, pooled


# --------------------------------------------------------------------------------------------
# a10 — HierarchicalAttention
# --------------------------------------------------------------------------------------------


def hierarchical_attention(sd, top_vec, cls, pairs_list, n_steps, sep_positions, pre="two_level_encoder."):
    """models/berson/modeling_bert.py:686-817.  top_vec [R,L,H], cls [R,H], pairs_list [B,P,2],
    sep_positions [B,P,2].  All manuals share n_steps (reference eval semantics)."""
    R, L, H = top_vec.shape
    B, P, _ = sep_positions.shape
    N = n_steps
    sep = sep_positions.reshape(R, 2)
    score = _lin(sd, pre + "sentence_tran_2", torch.tanh(_lin(sd, pre + "sentence_tran", top_vec))).squeeze(-1)
    pos = torch.arange(L)[None, :]
    m0 = ((pos >= 1) & (pos <= sep[:, 0:1])).float()            # span0 = 1..sep0      (711)
    m1 = ((pos > sep[:, 0:1]) & (pos <= sep[:, 1:2])).float()   # span1 = sep0+1..sep1 (712)
    sel = torch.stack([m0, m1], dim=1)                          # [R,2,L]
    att = sel * score[:, None, :] + (1.0 - sel) * -10000.0      # (722-731)
    mix = _drop(torch.softmax(att, -1), "H") @ top_vec          # [R,2,H]  (attention-prob dropout, 735)
    mix = mix.reshape(B, P, 2, H)
    cls_score = _lin(sd, pre + "pairwise_relationship", cls)    # [R,2]
    his1 = _lin(sd, pre + "h1_relationship", cls)
    his2 = _lin(sd, pre + "h2_relationship", cls)
    final = torch.zeros(B, N, H)
    cls_mat = torch.zeros(B, N, N, H)
    score_mat = torch.zeros(B, N, N, 2)
    his1_mat = torch.zeros(B, N, N, 2)
    his2_mat = torch.zeros(B, N, N, 2)
    clsb, sb, h1b, h2b = cls.reshape(B, P, H), cls_score.reshape(B, P, 2), his1.reshape(B, P, 2), his2.reshape(B, P, 2)
    edge = 2 * (N - 1)  # int(P / N * 2) (771)
    for b in range(B):
        count = [0] * N
        sent = torch.zeros(N, edge, H)
        for p, (i, j) in enumerate(pairs_list[b].tolist()):
            sent[i, count[i]] = mix[b, p, 0]; count[i] += 1
            sent[j, count[j]] = mix[b, p, 1]; count[j] += 1
            cls_mat[b, i, j] = clsb[b, p]
            score_mat[b, i, j] = sb[b, p]
            his1_mat[b, i, j] = h1b[b, p]
            his2_mat[b, i, j] = h2b[b, p]
        w = torch.softmax(_lin(sd, pre + "linear_in_2", sent, bias=False).squeeze(-1), -1)  # [N, edge]
        final[b] = (w[:, None, :] @ sent).squeeze(1)
    return final, cls_mat, cls_score, score_mat, his1_mat, his2_mat


# --------------------------------------------------------------------------------------------
# a11 — paragraph encoder
# --------------------------------------------------------------------------------------------


def paragraph_encoder(sd, x, mask, heads=8, layers=2, pre="encoder."):
    """models/berson/encoder.py:46-59 + 9-29; models/berson/neural.py:11-33, 98-232.  x [B,N,H], mask [B,N]."""
    x = x * mask[:, :, None].float()
    B, N, H = x.shape
    d = H // heads
    addm = ((1 - mask.float()) * -10000.0)[:, None, None, :]
    for i in range(layers):
        p = pre + "transformer_inter.%d." % i
        y = _ln(sd, p + "layer_norm", x, 1e-6) if i != 0 else x
        k = _heads(_lin(sd, p + "self_attn.linear_keys", y), heads)
        v = _heads(_lin(sd, p + "self_attn.linear_values", y), heads)
        q = _heads(_lin(sd, p + "self_attn.linear_query", y), heads) / math.sqrt(d)
        a = _drop(torch.softmax(q @ k.transpose(2, 3) + addm, -1), "PA", i)
        c = (a @ v).transpose(1, 2).reshape(B, N, H)
        out = _drop(_lin(sd, p + "self_attn.final_linear", c), "PC", i) + x
        f = p + "feed_forward."
        inter = _drop(gelu_tanh(_lin(sd, f + "w_1", _ln(sd, f + "layer_norm", out, 1e-6))), "PF1", i)
        x = _drop(_lin(sd, f + "w_2", inter), "PF2", i) + out
    return _ln(sd, pre + "layer_norm", x, 1e-6)


# --------------------------------------------------------------------------------------------
# a9 / a12 — BertForOrdering.encode, rela_encode
# --------------------------------------------------------------------------------------------


def encode(sd, cfg, inp):
    """models/berson/modeling_bert.py:1239-1366.  Returns a dict with the 10-tuple's tensors."""
    ids = inp["input_ids"]
    B, P, Lt = ids.shape
    N = int(inp["passage_length"][0])
    ids2, am2, tt2 = ids.reshape(B * P, Lt), inp["attention_mask"].reshape(B * P, Lt), inp["token_type_ids"].reshape(B * P, Lt)
    if (cfg.get("vit") is not None or cfg.get("rn") is not None) and inp.get("images") is not None:
        im = inp["images"]
        im = im.reshape(B * P * 2, *im.shape[3:])
        top_vec, visn, _ = (lxrt_forward_rn if cfg.get("rn") is not None else lxrt_forward)(sd, cfg, ids2, tt2, am2, im)
        cls = top_vec[:, 0]
    else:
        top_vec, cls = text_bert(sd, cfg, ids2, am2, tt2)
        visn = None
    final, cls_mat, cls_score, score_mat, his1, his2 = hierarchical_attention(
        sd, top_vec, cls, inp["pairs_list"], N, inp["sep_positions"])
    mask_cls = inp["mask_cls"]
    sents = final * mask_cls[:, :, None].float()
    para = paragraph_encoder(sd, sents, mask_cls.float(), cfg.get("para_heads", 8), cfg.get("para_layers", 2))
    para = para * mask_cls[:, :, None]
    h0 = (para.sum(1) / (inp["passage_length"].float() + 1e-20)[:, None])[None]
    key = _lin(sd, "key_linear", torch.cat([sents, para], -1))
    return dict(top_vec=top_vec, visn=visn, sents=sents, para=para, h0=h0, c0=torch.zeros_like(h0), key=key,
                cls=cls, cls_mat=cls_mat, cls_score=cls_score, score_mat=score_mat, his1=his1, his2=his2)


def rela_encode(cls_mat, score_mat):
    """modeling_bert.py:919-925 (history_encode 927-935 is called with the same score matrix, 1449)."""
    return torch.cat([cls_mat, torch.softmax(score_mat, -1)], -1)


# --------------------------------------------------------------------------------------------
# a13 / a14 / a15 — decode step, beam search (materialised formulation, as the reference)
# --------------------------------------------------------------------------------------------


def lstm_cell(sd, x, h, c, pre="decoder."):
    """nn.LSTM single layer/step (modeling_bert.py:886,1375): gate order i,f,g,o; two bias vectors."""
    g = x @ sd[pre + "weight_ih_l0"].t() + sd[pre + "bias_ih_l0"] + h @ sd[pre + "weight_hh_l0"].t() + sd[pre + "bias_hh_l0"]
    i, f, gg, o = g.chunk(4, -1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    return torch.sigmoid(o) * torch.tanh(c2), c2


def decode_step(sd, prev_y, h, c, key0, pointed, rela_vec, rela_mask, hist1, hist2, l1_mask, l2_mask):
    """modeling_bert.py:1368-1402.  prev_y [W,H]; h,c [W,H]; key0 [1,N,H]; pointed [W,N] bool;
    rela_vec/hist [W,N,N,H+2] (rela_vec is zeroed IN PLACE, 1385); masks [W,N,N]."""
    h, c = lstm_cell(sd, prev_y, h, c)
    q = _lin(sd, "query_linear", h)[:, None, :]
    left1 = (hist1 * l1_mask[..., None]).sum(1)
    left2 = (hist2 * l2_mask[..., None]).sum(1)
    rela_vec.mul_(rela_mask[..., None].to(rela_vec.dtype))
    forw, back = rela_vec.mean(2), rela_vec.mean(1)  # divide by N including zeros (1386-1387)
    keys = torch.cat([left1, left2, forw, back], -1) @ sd["pw_k.weight"].t()
    e = _lin(sd, "tanh_linear", torch.tanh(q + keys + key0)).squeeze(2)
    e = e.masked_fill(pointed, -1e9)
    return h, c, torch.log_softmax(e, -1)


def beam_step(cost, prev_scores, prev_cands, beam_size, target_len):
    """models/berson/generator.py:15-38.  cost [W',N] = -logp.  Returns (done, remain_ix, cands, scores, nbest)."""
    score = cost + torch.tensor(prev_scores, dtype=cost.dtype)[:, None]
    k = min(beam_size, score.numel())
    nb_score, nb_ix = score.reshape(-1).topk(k, largest=False)
    n = cost.shape[1]
    beam_ix = nb_ix // n
    tok_ix = nb_ix - beam_ix * n
    done, remain, cands, scores = [], [], [], []
    for s, b, t in zip(nb_score.tolist(), beam_ix.tolist(), tok_ix.tolist()):
        cand = prev_cands[b] + [t]
        if len(cand) == target_len:
            done.append([cand, s])
        else:
            remain.append(b); cands.append(cand); scores.append(s)
    return done, remain, cands, scores, (nb_score, beam_ix, tok_ix)


def beam_search(sd, enc, n_steps, beam_size, manual=0, trace=None):
    """modeling_bert.py:1411-1552 for ONE manual of an encode() result.  Returns the permutation
    (position t holds the index of the step predicted t-th).  `trace` (list) receives per-step
    dicts {logp, score, beam_ix, tok_ix}."""
    N = n_steps
    b = manual
    sents, key0 = enc["sents"][b, :N], enc["key"][b:b + 1, :N]
    h, c = enc["h0"][0, b:b + 1].clone(), enc["c0"][0, b:b + 1].clone()
    rela = rela_encode(enc["cls_mat"][b:b + 1], enc["score_mat"][b:b + 1]).clone()
    hist1, hist2 = rela.clone(), rela.clone()
    H = sents.shape[-1]
    eye_zeros = (1 - torch.eye(N)).to(rela.dtype)
    cands, scores = [[]], [0]
    target = N - 1
    valid = beam_size
    hyp = []
    pointed = rela_mask = None
    for t in range(target):
        if t == 0:
            x = sents.new_zeros(1, H)
            pointed = torch.zeros(1, N, dtype=torch.bool)
            rela_mask = eye_zeros[None].clone()
            l1 = torch.zeros_like(rela_mask); l2 = torch.zeros_like(rela_mask)
        else:
            idx = torch.tensor([cd[-1] for cd in cands])
            x = sents[idx]
            ar = torch.arange(len(idx))
            pointed[ar, idx] = True
            rela_mask[ar, :, idx] = 0
            rela_mask[ar, idx] = 0
            l1 = torch.zeros_like(rela_mask); l2 = torch.zeros_like(rela_mask)
            l1[ar, idx, :] = 1
            if t > 1:
                idx2 = torch.tensor([cd[-2] for cd in cands])
                l2[ar, idx2, :] = 1
        h, c, logp = decode_step(sd, x, h, c, key0, pointed, rela, rela_mask, hist1, hist2, l1, l2)
        done, remain, cands, scores, nbest = beam_step(-logp, scores, cands, valid, target)
        if trace is not None:
            trace.append(dict(logp=logp.clone(), score=nbest[0].clone(), beam_ix=nbest[1].clone(), tok_ix=nbest[2].clone()))
        hyp.extend(done)
        valid -= len(done)
        if valid == 0:
            break
        r = torch.tensor(remain, dtype=torch.long)
        h, c = h[r], c[r]
        pointed, rela_mask, rela = pointed[r], rela_mask[r], rela[r]
        hist1, hist2 = hist1[r], hist2[r]
    score = torch.tensor([hp[1] for hp in hyp])
    _, order = torch.sort(score)
    best = list(hyp[order[0].item()][0])
    best.append(sorted(set(range(N)) - set(best))[0])
    return best


def order_manuals(sd, cfg, input_ids, labels, n_steps, beam_size, images=None, traces=None):
    """berson_pointer_network (modeling_bert.py:1405-1408) applied per manual, as berson_evaluate
    does (models/berson/eval.py:85-129).  Returns list of permutations."""
    out = []
    for b in range(len(input_ids)):
        inp = prepare_inputs(input_ids[b:b + 1], labels[b:b + 1], n_steps,
                             None if images is None else images[b:b + 1])
        enc = encode(sd, cfg, inp)
        tr = [] if traces is not None else None
        out.append(beam_search(sd, enc, n_steps, beam_size, 0, tr))
        if traces is not None:
            traces.append(tr)
    return out


# --------------------------------------------------------------------------------------------
# a16 — teacher-forced training loss
# --------------------------------------------------------------------------------------------


def time_contrastive_triplets(target, rng=None):
    """modeling_bert.py:1176-1209: per manual an anchor POSITION is drawn, a neighbouring position as the positive, a position at
    least two away as the negative (np.random.choice, in this call order), and each is mapped through the ground-truth order
    `target[b]` to a sentence index.  -> long [B, 3] (anchor, positive, negative) rows of the sentence matrix."""
    import numpy as np
    rng = np.random if rng is None else rng
    B, N = target.shape
    out = []
    for b in range(B):
        anchor = rng.choice(list(range(N)), 1, replace=False)[0]
        pos = [i for i in (anchor - 1, anchor + 1) if 0 <= i < N]
        positive = rng.choice(pos, 1, replace=False)[0]
        negative = rng.choice([j for j in range(N) if abs(j - anchor) >= 2], 1, replace=False)[0]
        out.append([int(target[b][anchor]), int(target[b][positive]), int(target[b][negative])])
    return torch.tensor(out, dtype=torch.long)


def training_loss(sd, cfg, inp, lam=0.6, triplets=None, multimodal_loss=False):
    """modeling_bert.py:943-1174 (pointer NLL + lam * pairwise NLL), plus the optional time-contrastive term (1176-1216:
    0.1 * TripletMarginLoss(margin 1, p 2) over the sentence vectors picked by `triplets` [B, 3], see
    time_contrastive_triplets) and the optional image pairwise term (args.multimodal_loss: 1359-1364 img_projection of the first
    visual token -> the SAME pairwise_relationship head, 1218-1225 its NLL / P averaged over the batch, times lam).
    All manuals in the batch have N steps (tgt_len == num)."""
    enc = encode(sd, cfg, inp)
    target = inp["ground_truth"]
    B, N = target.shape
    sents, key0 = enc["sents"], enc["key"]
    ar = torch.arange(B)
    dec_in = torch.cat([sents.new_zeros(B, 1, sents.shape[-1]), sents[ar[:, None], target[:, :-1]]], 1)
    rela = rela_encode(enc["cls_mat"], enc["score_mat"]).clone()
    hist = rela.clone()
    rela_mask = (1 - torch.eye(N))[None].repeat(B, 1, 1)
    h, c = enc["h0"][0], enc["c0"][0]
    pointed = [torch.zeros(B, 1, N)]
    outs, pw_keys = [], []
    for t in range(N):
        l1 = torch.zeros(B, N, N); l2 = torch.zeros(B, N, N)
        if t > 0:
            tar = target[:, t - 1]
            rela_mask[ar, tar] = 0
            rela_mask[ar, :, tar] = 0
            l1[ar, tar, :] = 1
            if t > 1:
                l2[ar, target[:, t - 2], :] = 1
            pm = pointed[-1].clone(); pm[ar, :, tar] = 1; pointed.append(pm)
        left1 = (hist * l1[..., None]).sum(1)
        left2 = (hist * l2[..., None]).sum(1)
        rela = rela * rela_mask[..., None].clone()   # (the reference zeroes in place, 1385; out of place here so autograd can replay it)
        pw = torch.cat([left1, left2, rela.mean(2), rela.mean(1)], -1)
        pw_keys.append((pw @ sd["pw_k.weight"].t())[:, None])
        h, c = lstm_cell(sd, dec_in[:, t], h, c)
        outs.append(h[:, None])
    query = _lin(sd, "query_linear", torch.cat(outs, 1))[:, :, None]
    e = _lin(sd, "tanh_linear", torch.tanh(query + torch.cat(pw_keys, 1) + key0[:, None])).squeeze(-1)
    e = e.masked_fill(torch.cat(pointed, 1) == 1, -1e9)
    logp = torch.log_softmax(e, -1).reshape(B * N, N)
    nll = -logp[torch.arange(B * N), target.reshape(-1)].reshape(B, N)
    ptr = (nll.sum(-1) / (inp["passage_length"].float() + 1e-20 - 1)).sum() / B
    lp = torch.log_softmax(enc["cls_score"], -1)
    pl = inp["pairwise_labels"].reshape(-1)
    pair = (-lp[torch.arange(lp.shape[0]), pl]).reshape(B, -1)
    pair = (pair.sum(-1) / (inp["pairs_num"].float() + 1e-20)).sum() / B
    loss = ptr + lam * pair
    if triplets is not None:
        a, p, n = (sents[ar, triplets[:, i]] for i in range(3))
        loss = loss + 0.1 * F.triplet_margin_loss(a, p, n, margin=1.0, p=2)
    if multimodal_loss:
        score_img = _lin(sd, "two_level_encoder.pairwise_relationship", _lin(sd, "img_projection", enc["visn"][:, 0]))
        lpi = torch.log_softmax(score_img, -1)
        pair_img = (-lpi[torch.arange(lpi.shape[0]), pl]).reshape(B, -1)
        loss = loss + lam * (pair_img.sum(-1) / (inp["pairs_num"].float() + 1e-20)).sum() / B
    return loss


# --------------------------------------------------------------------------------------------
# a19 — metrics
# --------------------------------------------------------------------------------------------


def cal_result(truth, predicted):
    """models/berson/eval.py:190-260: (acc, pmr, tau) — the triple berson_evaluate returns.
    tau = 1 - 2 * (#pairs ordered differently) / C(N,2) via sets of ordered pairs (237-247)."""
    accs, taus, pmr_right = [], [], 0
    for t, p in zip(truth, predicted):
        eq = [int(a == b) for a, b in zip(t, p)]
        accs.append(sum(eq) / len(t))
        pmr_right += int(all(eq))
        s_t = set(itertools.combinations(t, 2))
        s_p = set(itertools.combinations(p, 2))
        cn2 = len(p) * (len(p) - 1) / 2
        taus.append(1 - 2 * (len(s_p) - len(s_p & s_t)) / cn2)
    n = len(truth)
    return sum(accs) / n, pmr_right / n, sum(taus) / n


# --------------------------------------------------------------------------------------------
# synthetic inputs shared by tests / bench
# --------------------------------------------------------------------------------------------


def synthetic_manuals(B, n_steps, tokens_per_step=64, vocab=30522, image_px=None, seed=1, lo=1000):
    """`[101] + (l-2) ids ~U{lo..vocab-1} + [102]` per step; labels = randperm; images ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    lo = min(lo, max(103, vocab // 4))
    body = torch.randint(lo, vocab, (B, n_steps, tokens_per_step - 2), generator=g)
    ids = torch.cat([torch.full((B, n_steps, 1), 101), body, torch.full((B, n_steps, 1), 102)], -1)
    ids = ids.reshape(B, n_steps * tokens_per_step)
    labels = torch.stack([torch.randperm(n_steps, generator=g) for _ in range(B)])
    images = None
    if image_px:
        images = torch.randn(B, n_steps, 3, image_px, image_px, generator=g)
    return ids, labels, images


# --------------------------------------------------------------------------------------------
# a5 — CLIP ModifiedResNet tower (the reference's wired default backbone) + RN-only LXRT embeddings
# --------------------------------------------------------------------------------------------


def _bn(sd, name, x, train=False):
    """BatchNorm2d, eps 1e-5.  eval: running statistics.  train=True (the mode the reference fine-tunes in): statistics
    of the batch itself -- mean and BIASED variance over (N, H, W), torch.nn.functional.batch_norm(training=True)."""
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    if train:
        m = x.mean(dim=(0, 2, 3))
        v = ((x - m[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))
    else:
        m, v = sd[name + ".running_mean"], sd[name + ".running_var"]
    return (x - m[None, :, None, None]) / torch.sqrt(v[None, :, None, None] + 1e-5) * w[None, :, None, None] + b[None, :, None, None]


def _bottleneck(sd, pre, x, stride, bn_train=False):
    """models/CLIP/clip/model.py:10-53: 1x1 -> 3x3 -> (avgpool) -> 1x1, anti-aliased down-sampling branch."""
    out = torch.relu(_bn(sd, pre + "bn1", F.conv2d(x, sd[pre + "conv1.weight"]), bn_train))
    out = torch.relu(_bn(sd, pre + "bn2", F.conv2d(out, sd[pre + "conv2.weight"], padding=1), bn_train))
    if stride > 1:
        out = F.avg_pool2d(out, stride)
    out = _bn(sd, pre + "bn3", F.conv2d(out, sd[pre + "conv3.weight"]), bn_train)
    identity = x
    if (pre + "downsample.0.weight") in sd:
        identity = F.avg_pool2d(x, stride) if stride > 1 else x
        identity = _bn(sd, pre + "downsample.1", F.conv2d(identity, sd[pre + "downsample.0.weight"]), bn_train)
    return torch.relu(out + identity)


def rn_pair_tower(sd, pre, images, rn, img_len=2):
    """ModifiedResNet.forward + AttentionPool2d.forward (clip/model.py:171-187, 71-125), skip_last_layer=False
    (as hard-wired, lxrt/modeling.py:783).  images [R*img_len,3,S,S] -> [R, 1 + img_len*g*g, 2*output_dim].

    Quirk reproduced (model.py:76): the NCHW feature maps of a pair are reshaped [R, C, g*g*img_len] WITHOUT moving the
    image axis, so token t / channel c' reads flat element c'*(g*g*img_len) + t of the pair's [img, C, g*g] block."""
    x = images
    bn_train = bool(rn.get("bn_train", False))   # fine-tuning: BatchNorm over the batch of MATERIALISED pair images
    for i in (1, 2, 3):
        x = torch.relu(_bn(sd, pre + "bn%d" % i, F.conv2d(x, sd[pre + "conv%d.weight" % i], stride=2 if i == 1 else 1, padding=1),
                           bn_train))
    x = F.avg_pool2d(x, 2)
    for li, nblk in enumerate(rn["vision_layers"], start=1):
        for bi in range(nblk):
            x = _bottleneck(sd, pre + "layer%d.%d." % (li, bi), x, 2 if (li > 1 and bi == 0) else 1, bn_train)
    a = pre + "attnpool."
    R = x.shape[0] // img_len
    C, g2 = x.shape[1], x.shape[2] * x.shape[3]
    t = x.reshape(R, C, g2 * img_len).permute(2, 0, 1)                      # (HW*il) R C
    t = torch.cat([t.mean(dim=0, keepdim=True), t], dim=0)
    pos = sd[a + "positional_embedding"]
    t = t + torch.cat([pos] + [pos[:g2]] * (img_len - 1), dim=0)[:, None, :]
    heads = rn["vision_width"] * 32 // 64
    d = C // heads
    L = t.shape[0]
    q = (t @ sd[a + "q_proj.weight"].t() + sd[a + "q_proj.bias"]).reshape(L, R, heads, d).permute(1, 2, 0, 3) * (d ** -0.5)
    k = (t @ sd[a + "k_proj.weight"].t() + sd[a + "k_proj.bias"]).reshape(L, R, heads, d).permute(1, 2, 0, 3)
    v = (t @ sd[a + "v_proj.weight"].t() + sd[a + "v_proj.bias"]).reshape(L, R, heads, d).permute(1, 2, 0, 3)
    o = (torch.softmax(q @ k.transpose(-1, -2), -1) @ v).permute(0, 2, 1, 3).reshape(R, L, C)
    o = o @ sd[a + "c_proj.weight"].t() + sd[a + "c_proj.bias"]
    return torch.cat([o, o], dim=-1)                                        # model.py:106


def lxrt_forward_rn(sd, cfg, ids, tt, attention_mask, images, pre="bert."):
    """LXRTModel.forward with the RN backbone as wired (lxrt/modeling.py:874-882, 1014-1030, 1063-1107):
    tower -> + LinearPositionEmbedding (x/y, 621-660) -> + VisualTokenTypeEmbedding (663-705) -> visn_fc -> joint BERT."""
    rn = cfg["rn"]
    tower = rn_pair_tower(sd, pre + "encoder.visual_model.visual.", images, rn)
    g = rn["image_resolution"] // 32
    xe, ye = sd[pre + "encoder.visual_pos.x_position_embedding.weight"][:g], sd[pre + "encoder.visual_pos.y_position_embedding.weight"][:g]
    pe = (xe[:, None, :] + ye[None, :, :]).reshape(g * g, -1)
    pe = torch.cat([pe[0:1], pe, pe], dim=0)                                # skip_last_layer=False, img_len=2 (649-651)
    te = sd[pre + "encoder.visual_token_type.token_type_embedding.weight"]
    type_ids = torch.zeros(1 + 2 * g * g, dtype=torch.long)
    type_ids[1 + g * g:1 + 2 * g * g] = 1                                    # 691-696
    v = tower + pe[None] + te[type_ids][None]
    v = _ln(sd, pre + "encoder.visn_fc.visn_layer_norm", _lin(sd, pre + "encoder.visn_fc.visn_fc", v), 1e-12)
    emb = bert_embeddings(sd, pre + "embeddings.", ids, tt, 1e-12)
    R, Lt = ids.shape
    joint = torch.cat([emb, v], dim=1)
    m = torch.cat([ext_mask(attention_mask), torch.zeros(R, 1, 1, v.shape[1])], dim=-1)
    for i in range(cfg["num_hidden_layers"]):
        joint = bert_layer(sd, pre + "encoder.layer.%d." % i, joint, m, cfg["num_attention_heads"], 1e-12)
    return joint[:, :Lt], joint[:, Lt:], _lin(sd, pre + "pooler.dense", joint[:, 0])
