"""Import the REAL reference (telin0411/multimodal_sequencing) from /root/reference on CPU.

Only usable in the build container (the GPU box has no /root/reference).  Used by
`make_golden.py` to generate the committed fixtures and by `tests/test_oracle_vs_reference.py`
(skipped when the reference is absent) to pin the oracle restatement to the reference itself.

The shims below touch no arithmetic:
  * boto3 / botocore stubs                  (models/berson/file_utils.py:20-22)
  * transformers.modeling_roberta alias      (models/CLIP/src/lxrt/modeling.py:36)
  * uint8 -> bool cast for masked_fill[_]    (models/berson/modeling_bert.py:1399 with .byte() masks;
                                              identical semantics under the pinned torch 1.8)
  * fake `clip` module: random-init CLIP instead of a network download
                                             (models/CLIP/clip/clip.py:63-83)
  * ViT-B/32 adapter (oracle DECISION): VisualTransformer.forward accepts and
    ignores `img_len=`; encoder.skip_last_layer=True -> ln_post tokens (width-d);
    VISUAL_CONFIG.set_visual_dims(width, 4) before LXRTModel construction.
"""
import os
import sys
import types

import torch

REF = os.environ.get("MSQ_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "models", "berson"))


_state = {}


class StubTokenizer:
    """prepare_berson_inputs only needs cls/sep/pad ids (process_inputs_for_berson.py:44,121-122)."""
    cls_token, sep_token, pad_token = "[CLS]", "[SEP]", "[PAD]"
    _ids = {"[CLS]": 101, "[SEP]": 102, "[PAD]": 0}

    def convert_tokens_to_ids(self, tok):
        return self._ids[tok]


def load():
    """Install shims and import the reference modules.  Returns a namespace of modules."""
    if _state:
        return _state["ns"]
    assert available(), "reference not present at %s" % REF
    for p in (REF, os.path.join(REF, "models/CLIP/src"), os.path.join(REF, "models/CLIP/clip")):
        if p not in sys.path:
            sys.path.insert(0, p)

    # --- boto3 / botocore stubs
    for name in ("boto3", "botocore", "botocore.config", "botocore.exceptions"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["botocore.config"].Config = object
    sys.modules["botocore.exceptions"].ClientError = Exception
    sys.modules["botocore"].config = sys.modules["botocore.config"]
    sys.modules["botocore"].exceptions = sys.modules["botocore.exceptions"]

    # --- transformers.modeling_roberta alias
    import transformers
    if "transformers.modeling_roberta" not in sys.modules:
        from transformers.models.roberta import modeling_roberta as _mr
        alias = types.ModuleType("transformers.modeling_roberta")
        alias.RobertaClassificationHead = _mr.RobertaClassificationHead
        sys.modules["transformers.modeling_roberta"] = alias

    # --- uint8 masks for masked_fill
    if not getattr(torch.Tensor, "_msq_mf_patched", False):
        _mf_, _mf = torch.Tensor.masked_fill_, torch.Tensor.masked_fill

        def masked_fill_(self, mask, value):
            return _mf_(self, mask.bool() if mask.dtype == torch.uint8 else mask, value)

        def masked_fill(self, mask, value):
            return _mf(self, mask.bool() if mask.dtype == torch.uint8 else mask, value)

        torch.Tensor.masked_fill_ = masked_fill_
        torch.Tensor.masked_fill = masked_fill
        torch.Tensor._msq_mf_patched = True

    # --- fake clip module (random-init, no download)
    import model as clip_model_mod  # models/CLIP/clip/model.py

    fake_clip = types.ModuleType("clip")
    fake_clip._vit_cfg = dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768,
                              vision_patch_size=32)

    fake_clip._rn_cfg = dict(embed_dim=1024, image_resolution=224, vision_layers=(3, 4, 6, 3), vision_width=64)

    def clip_load(name, device="cpu", jit=False, img_len=None, img_only=False):
        if name == "ViT-B/32":
            c = fake_clip._vit_cfg
            m = clip_model_mod.CLIP(c["embed_dim"], c["image_resolution"], c["vision_layers"], c["vision_width"],
                                    c["vision_patch_size"], 77, 49408, 512, 8, 12, img_len=img_len,
                                    img_only=img_only)
        elif name == "RN50":
            c = fake_clip._rn_cfg
            m = clip_model_mod.CLIP(c["embed_dim"], c["image_resolution"], tuple(c["vision_layers"]), c["vision_width"],
                                    None, 77, 49408, 512, 8, 12, img_len=img_len, img_only=img_only)
        else:
            raise RuntimeError(name)
        return m.eval().float(), None

    fake_clip.load = clip_load
    sys.modules["clip"] = fake_clip

    # --- ViT adapter: accept-and-ignore img_len
    VT = clip_model_mod.VisualTransformer
    if not getattr(VT, "_msq_patched", False):
        _fwd = VT.forward

        def forward(self, x, skip_last_layer=False, text_embedding=None, text_mask=None, img_len=None):
            return _fwd(self, x, skip_last_layer=skip_last_layer, text_embedding=text_embedding,
                        text_mask=text_mask)

        VT.forward = forward
        VT._msq_patched = True

    import param
    from models.berson import modeling_bert as berson_mb
    from models.berson import generator as berson_gen
    from models.berson import process_inputs_for_berson as berson_prep
    from models.berson.configuration_bert import BertConfig as BersonBertConfig
    from lxrt import modeling as lxrt_modeling
    from trainers import metrics as ref_metrics

    ns = types.SimpleNamespace(param=param, berson=berson_mb, gen=berson_gen, prep=berson_prep,
                               BersonBertConfig=BersonBertConfig, lxrt=lxrt_modeling,
                               clip_model=clip_model_mod, fake_clip=fake_clip, metrics=ref_metrics)
    _state["ns"] = ns
    return ns


def make_args(N, beam, multimodal=False, device="cpu"):
    """argparse namespace with the fields the path reads."""
    return types.SimpleNamespace(
        ff_size=3072, heads=8, para_dropout=0.1, inter_layers=2, beam_size=beam, pairwise_loss_lam=0.6,
        multimodal_loss=False, additional_wrapper_level_objectives=None, device=device,
        multimodal=multimodal, use_multimodal_model=False,
        multimodal_model_type="clip" if multimodal else None, multimodal_img_part=False,
        multimodal_text_part=False, per_seq_max_length=64, max_story_length=N)


def build_text_model(ns, cfg_kwargs, args, seed=0):
    """BertForOrdering with the vendored text-only BertModel (modeling_bert.py:860-861)."""
    torch.manual_seed(seed)
    cfg = ns.BersonBertConfig(**cfg_kwargs)
    cfg.wrapper_model_with_heatmap = False
    cfg.v_feature_size = 1024
    m = ns.berson.BertForOrdering(cfg, args, tokenizer=None)
    return m.eval()


def build_multimodal_model(ns, cfg_kwargs, args, vit_cfg=None, seed=0, rn_cfg=None, v_feature_size=1024):
    """BertForOrdering around LXRTModel + CLIP tower (train.py:1869-1880, 2024-2028).  vit_cfg -> ViT-B/32-shaped tower
    with skip_last_layer=True; rn_cfg -> the "RN50" ModifiedResNet branch exactly as wired (skip_last_layer=False,
    visual_feat_dim = 2*embed_dim), BatchNorm statistics randomised so eval-mode BN is not the identity."""
    torch.manual_seed(seed)
    name = "RN50" if rn_cfg is not None else "ViT-B/32"
    if rn_cfg is not None:
        ns.fake_clip._rn_cfg = dict(rn_cfg)
        width = 2 * rn_cfg["embed_dim"]
    else:
        if vit_cfg is not None:
            ns.fake_clip._vit_cfg = dict(vit_cfg)
        width = ns.fake_clip._vit_cfg["vision_width"]
    ns.param.VISUAL_CONFIG.set_visual_dims(width, 4)
    ns.param.VISUAL_CONFIG.clip_model_name = name
    cfg = ns.BersonBertConfig(**cfg_kwargs)
    cfg.wrapper_model_with_heatmap = False
    cfg.v_feature_size = v_feature_size   # train.py:2022 hard-wires 1024 (= RoBERTa-large's hidden size)
    lx = dict(cfg_kwargs)
    lx.pop("layer_norm_eps", None)
    lxcfg = ns.lxrt.BertConfig(**lx)  # lxrt has its own BertConfig class (lxrt/modeling.py:147)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        inner = ns.lxrt.LXRTModel(lxcfg, multimodal_text_part=False, multimodal_img_part=False,
                                  cls_id=101, sep_id=102, max_story_length=args.max_story_length,
                                  clip_model_name=name)
    if rn_cfg is None:
        inner.encoder.skip_last_layer = True
    else:
        g = torch.Generator().manual_seed(seed + 77)
        for mod in inner.encoder.visual_model.visual.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 0.5 + 0.75)
                mod.weight.data.copy_(torch.rand(mod.weight.shape, generator=g) * 0.5 + 0.75)
                mod.bias.data.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
    m = ns.berson.BertForOrdering(cfg, args, tokenizer=StubTokenizer())
    m.bert = inner
    return m.eval()
