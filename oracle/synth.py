"""ORACLE-side synthetic weights and inputs (test infrastructure; see berson_oracle.py header).

Seeded generators for state_dicts and `encode` outputs that are too large to commit.  The
distributions follow the reference's initialisers (SURVEY.md §8(c) "Weights hand-over", Appendix
A.19): nn.Linear / nn.Embedding ~ N(0, 0.02), LayerNorm = (1, 0), biases 0
(models/berson/modeling_bert.py:464-474, models/CLIP/src/lxrt/modeling.py:1243-1254), nn.LSTM
~ U(-1/sqrt(H), 1/sqrt(H)) (torch default; not touched by init_weights), CLIP tower parameters
not matched by init_bert_weights keep CLIP's own init (models/CLIP/clip/model.py:248-258).
Matching the reference's RNG *stream* is never required: one state_dict is loaded on both sides.
"""
import math

import torch


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _normal(g, *shape, std=0.02):
    return torch.randn(*shape, generator=g) * std


def decode_head_weights(H, seed=7):
    """Weights used by BertForOrdering.step (modeling_bert.py:1368-1402) with the reference's keys."""
    g = _g(seed)
    k = 1.0 / math.sqrt(H)
    u = lambda *s: (torch.rand(*s, generator=g) * 2 - 1) * k
    return {
        "decoder.weight_ih_l0": u(4 * H, H), "decoder.weight_hh_l0": u(4 * H, H),
        "decoder.bias_ih_l0": u(4 * H), "decoder.bias_hh_l0": u(4 * H),
        "query_linear.weight": _normal(g, H, H), "query_linear.bias": torch.zeros(H),
        "tanh_linear.weight": _normal(g, 1, H), "tanh_linear.bias": torch.zeros(1),
        "pw_k.weight": _normal(g, H, 4 * (H + 2)),
    }


def synthetic_encode(N, H, seed, B=1):
    """decode-only inputs, `sents,key,cls_mat,score_mat ~ N(0,1)`; the
    diagonal of the pair tables is zero as HierarchicalAttention leaves it (modeling_bert.py:746,761)."""
    g = _g(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    off = (1 - torch.eye(N))[None, :, :, None]
    sents, para, key = r(B, N, H), r(B, N, H), r(B, N, H)
    cls_mat = r(B, N, N, H) * off
    score_mat = r(B, N, N, 2) * off
    h0 = para.mean(1)[None]
    P = N * (N - 1)
    return dict(sents=sents, para=para, h0=h0, c0=torch.zeros_like(h0), key=key,
                cls=r(B * P, H), cls_mat=cls_mat, cls_score=r(B * P, 2), score_mat=score_mat,
                his1=r(B, N, N, 2) * off, his2=r(B, N, N, 2) * off)


def berson_head_weights(H, ff=3072, para_layers=2, seed=3):
    """All BertForOrdering head parameters (Appendix B) — everything except `bert.*`."""
    g = _g(seed)
    sd = decode_head_weights(H, seed + 1000)
    lin = lambda name, o, i, bias=True: sd.update(
        {name + ".weight": _normal(g, o, i), **({name + ".bias": torch.zeros(o)} if bias else {})})
    ln = lambda name: sd.update({name + ".weight": torch.ones(H), name + ".bias": torch.zeros(H)})
    lin("classifier", 2, H)
    for i in range(para_layers):
        p = "encoder.transformer_inter.%d." % i
        for n in ("linear_keys", "linear_values", "linear_query", "final_linear"):
            lin(p + "self_attn." + n, H, H)
        lin(p + "feed_forward.w_1", ff, H)
        lin(p + "feed_forward.w_2", H, ff)
        ln(p + "feed_forward.layer_norm")
        ln(p + "layer_norm")
    ln("encoder.layer_norm")
    lin("key_linear", H, 2 * H)
    t = "two_level_encoder."
    lin(t + "linear_in_2", 1, H, bias=False)
    lin(t + "sentence_tran", H, H)
    lin(t + "sentence_tran_2", 1, H)
    for n in ("pairwise_relationship", "h1_relationship", "h2_relationship"):
        lin(t + n, 2, H)
    return sd


def bert_weights(pre, H, layers, inter, vocab, max_pos, seed, lxrt=False, type_vocab=2):
    """BERT embeddings + layer stack under prefix `pre` (keys of Appendix B)."""
    g = _g(seed)
    sd = {}
    lin = lambda name, o, i: sd.update({name + ".weight": _normal(g, o, i), name + ".bias": torch.zeros(o)})
    ln = lambda name: sd.update({name + ".weight": torch.ones(H), name + ".bias": torch.zeros(H)})
    e = pre + "embeddings."
    sd[e + "word_embeddings.weight"] = _normal(g, vocab, H)
    sd[e + "position_embeddings.weight"] = _normal(g, max_pos, H)
    sd[e + "token_type_embeddings.weight"] = _normal(g, type_vocab, H)
    ln(e + "LayerNorm")
    for i in range(layers):
        l = pre + "encoder.layer.%d." % i
        for n in ("query", "key", "value"):
            lin(l + "attention.self." + n, H, H)
        lin(l + "attention.output.dense", H, H)
        ln(l + "attention.output.LayerNorm")
        lin(l + "intermediate.dense", inter, H)
        lin(l + "output.dense", H, inter)
        ln(l + "output.LayerNorm")
    if lxrt:
        lin(pre + "pooler.dense", H, H)
    return sd


def vit_weights(pre, width, layers, patch, res, seed):
    """CLIP VisualTransformer parameters (clip/model.py:242-258) after LXRTModel's
    `apply(init_bert_weights)` (lxrt/modeling.py:1464): nn.Linear -> N(0,.02)/0, LayerNorm -> 1/0;
    conv1, class_embedding, positional_embedding, in_proj_* keep CLIP / torch defaults."""
    g = _g(seed)
    sd = {}
    scale = width ** -0.5
    n_tok = (res // patch) ** 2 + 1
    fan_in = 3 * patch * patch
    bound = 1.0 / math.sqrt(fan_in)  # kaiming_uniform(a=sqrt(5)) bound of nn.Conv2d
    sd[pre + "conv1.weight"] = (torch.rand(width, 3, patch, patch, generator=g) * 2 - 1) * bound
    sd[pre + "class_embedding"] = scale * torch.randn(width, generator=g)
    sd[pre + "positional_embedding"] = scale * torch.randn(n_tok, width, generator=g)
    ln = lambda name: sd.update({name + ".weight": torch.ones(width), name + ".bias": torch.zeros(width)})
    ln(pre + "ln_pre")
    xb = math.sqrt(6.0 / (width + 3 * width))  # xavier_uniform of in_proj_weight
    for i in range(layers):
        b = pre + "transformer.resblocks.%d." % i
        sd[b + "attn.in_proj_weight"] = (torch.rand(3 * width, width, generator=g) * 2 - 1) * xb
        sd[b + "attn.in_proj_bias"] = torch.zeros(3 * width)
        sd[b + "attn.out_proj.weight"] = _normal(g, width, width)
        sd[b + "attn.out_proj.bias"] = torch.zeros(width)
        ln(b + "ln_1")
        sd[b + "mlp.c_fc.weight"] = _normal(g, 4 * width, width)
        sd[b + "mlp.c_fc.bias"] = torch.zeros(4 * width)
        sd[b + "mlp.c_proj.weight"] = _normal(g, width, 4 * width)
        sd[b + "mlp.c_proj.bias"] = torch.zeros(width)
        ln(b + "ln_2")
    ln(pre + "ln_post")
    return sd


def rn_weights(pre, rn, seed):
    """CLIP ModifiedResNet parameters (clip/model.py:128-187): kaiming-uniform convolutions, BatchNorm affine and
    running statistics drawn away from the identity (a trained tower's are), attnpool as CLIP initialises it."""
    g = _g(seed)
    sd = {}
    w, E = rn["vision_width"], rn["embed_dim"]
    C = 32 * w

    def conv(name, cout, cin, k):
        bound = math.sqrt(6.0 / (cin * k * k))  # He-uniform keeps activations O(1) through ~50 ReLU convolutions
        sd[name + ".weight"] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound

    def bn(name, c, gain=1.0):
        sd[name + ".weight"] = (torch.rand(c, generator=g) * 0.5 + 0.75) * gain
        sd[name + ".bias"] = torch.randn(c, generator=g) * 0.1
        sd[name + ".running_mean"] = torch.randn(c, generator=g) * 0.1
        sd[name + ".running_var"] = torch.rand(c, generator=g) * 0.5 + 0.75

    conv(pre + "conv1", w // 2, 3, 3); bn(pre + "bn1", w // 2)
    conv(pre + "conv2", w // 2, w // 2, 3); bn(pre + "bn2", w // 2)
    conv(pre + "conv3", w, w // 2, 3); bn(pre + "bn3", w)
    cin = w
    for s, nblk in enumerate(rn["vision_layers"]):
        p = w << s
        for b in range(nblk):
            k = pre + "layer%d.%d." % (s + 1, b)
            conv(k + "conv1", p, cin, 1); bn(k + "bn1", p)
            conv(k + "conv2", p, p, 3); bn(k + "bn2", p)
            conv(k + "conv3", 4 * p, p, 1); bn(k + "bn3", 4 * p, gain=0.5)
            if b == 0:
                conv(k + "downsample.0", 4 * p, cin, 1); bn(k + "downsample.1", 4 * p, gain=0.5)
            cin = 4 * p
    n_tok = (rn["image_resolution"] // 32) ** 2 + 1
    sd[pre + "attnpool.positional_embedding"] = torch.randn(n_tok, C, generator=g) / C ** 0.5
    std = C ** -0.5
    for n, o in (("q_proj", C), ("k_proj", C), ("v_proj", C), ("c_proj", E)):
        sd[pre + "attnpool.%s.weight" % n] = torch.randn(o, C, generator=g) * std
        sd[pre + "attnpool.%s.bias" % n] = torch.randn(o, generator=g) * 0.02
    return sd


def rn_lxrt_extras(pre, rn, H, seed):
    """visual_pos / visual_token_type / visn_fc of the RN branch (lxrt/modeling.py:621-705, 874-882) after init_bert_weights."""
    g = _g(seed)
    F = 2 * rn["embed_dim"]
    return {pre + "encoder.visual_pos.x_position_embedding.weight": _normal(g, 25, F),
            pre + "encoder.visual_pos.y_position_embedding.weight": _normal(g, 25, F),
            pre + "encoder.visual_token_type.token_type_embedding.weight": _normal(g, 5, F),
            pre + "encoder.visn_fc.visn_fc.weight": _normal(g, H, F),
            pre + "encoder.visn_fc.visn_fc.bias": torch.zeros(H),
            pre + "encoder.visn_fc.visn_layer_norm.weight": torch.ones(H),
            pre + "encoder.visn_fc.visn_layer_norm.bias": torch.zeros(H)}


RN50 = dict(embed_dim=1024, image_resolution=224, vision_layers=(3, 4, 6, 3), vision_width=64)
BERT_BASE = dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                 vocab_size=30522, max_position_embeddings=512)
# scripts/wikihow_finetune.sh: --config_name roberta-large (the LXRT embeddings / layers are built from this config)
ROBERTA_LARGE = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                     vocab_size=50265, max_position_embeddings=514, type_vocab_size=1)
VIT_B32 = dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=32)


def full_state_dict(cfg=None, vit=None, seed=0, ff=3072, rn=None):
    """Seeded state_dict of a whole BertForOrdering (+ LXRT and the ViT / ResNet tower when `vit` / `rn` is given)."""
    cfg = dict(BERT_BASE if cfg is None else cfg)
    H = cfg["hidden_size"]
    sd = berson_head_weights(H, ff=ff, seed=seed + 3)
    sd.update(bert_weights("bert.", H, cfg["num_hidden_layers"], cfg["intermediate_size"], cfg["vocab_size"],
                           cfg["max_position_embeddings"], seed + 11, lxrt=vit is not None or rn is not None,
                           type_vocab=cfg.get("type_vocab_size", 2)))
    if rn is not None:
        sd.update(rn_weights("bert.encoder.visual_model.visual.", rn, seed + 23))
        sd.update(rn_lxrt_extras("bert.", rn, H, seed + 17))
    if vit is not None:
        g = _g(seed + 17)
        W = vit["vision_width"]
        sd["bert.encoder.visn_fc.visn_fc.weight"] = _normal(g, H, W)
        sd["bert.encoder.visn_fc.visn_fc.bias"] = torch.zeros(H)
        sd["bert.encoder.visn_fc.visn_layer_norm.weight"] = torch.ones(H)
        sd["bert.encoder.visn_fc.visn_layer_norm.bias"] = torch.zeros(H)
        sd.update(vit_weights("bert.encoder.visual_model.visual.", W, vit["vision_layers"],
                              vit["vision_patch_size"], vit["image_resolution"], seed + 23))
    return sd
