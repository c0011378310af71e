"""models/CLIP/src/lxrt/modeling.py of the reference, BERSON surface: BertConfig (147-232) and LXRTModel
(1456-1598) in CLIP / visualbert-style mode.  Same constructor kwargs, forward signature and state_dict
keys; the forward is msq_inner_forward (ViT or ResNet pair tower -> [+ visual_pos +
visual_token_type for the ResNet] -> visn_fc -> joint BERT)."""
import json
import os

import torch
import torch.nn as nn

from multimodal_sequencing_b200.engine import OrderingEngine
from multimodal_sequencing_b200.dropin._owner import owned_engine
from models.berson.modeling_bert import _bert_layer, _holder, _init_bert_weights
from models.CLIP.clip.model import CLIP

CLIP_CONFIGS = {
    # clip.load("ViT-B/32") geometry (models/CLIP/clip/model.py:471-508 rebuilds it from the checkpoint)
    "ViT-B/32": dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=32),
    # the reference's wired default (param.py VISUAL_CONFIG.clip_model_name): ModifiedResNet + AttentionPool2d
    "RN50": dict(embed_dim=1024, image_resolution=224, vision_layers=(3, 4, 6, 3), vision_width=64, vision_patch_size=None),
}


class BertConfig(object):
    """lxrt/modeling.py:147-210."""

    def __init__(self, vocab_size_or_config_json_file, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                 max_position_embeddings=512, type_vocab_size=2, initializer_range=0.02):
        self.vocab_size = vocab_size_or_config_json_file
        self.hidden_size, self.num_hidden_layers, self.num_attention_heads = hidden_size, num_hidden_layers, num_attention_heads
        self.hidden_act, self.intermediate_size = hidden_act, intermediate_size
        self.hidden_dropout_prob, self.attention_probs_dropout_prob = hidden_dropout_prob, attention_probs_dropout_prob
        self.max_position_embeddings, self.type_vocab_size, self.initializer_range = max_position_embeddings, type_vocab_size, initializer_range


class LXRTModel(nn.Module):
    """lxrt/modeling.py:1456-1598.  kwargs as the reference: multimodal_text_part, multimodal_img_part, cls_id,
    sep_id, max_story_length, hl_include_objectives, mlm_ignore_index, clip_model_name, num_labels
    (+ clip_config=dict(...) to override the tower geometry, e.g. for small test models)."""

    def __init__(self, config, **kwargs):
        super().__init__()
        self.config = config
        if kwargs.get("multimodal_text_part") or kwargs.get("multimodal_img_part"):
            raise NotImplementedError("text-only / image-only LXRT parts are outside the scoped path")
        # topo-sort pairwise classifier mode (lxrt/modeling.py:1502-1511): RobertaClassificationHead on the pooled CLS
        self.topo_sort = kwargs.get("num_labels") is not None
        if self.topo_sort:
            self.num_labels = kwargs["num_labels"]
            config.num_labels = self.num_labels
        name = kwargs.get("clip_model_name", "ViT-B/32")
        vit = kwargs.get("clip_config") or CLIP_CONFIGS.get(name)
        if vit is None:
            raise NotImplementedError("visual backbone %r: the towers built are %s" % (name, sorted(CLIP_CONFIGS)))
        self.vit_config = dict(vit)
        self.is_resnet = isinstance(vit["vision_layers"], (tuple, list))
        # features per visual token handed to visn_fc (VISUAL_CONFIG.visual_feat_dim): ViT width, or 2*embed_dim for the
        # ResNet whose attention pool returns cat([x, x]) (clip/model.py:106)
        F = 2 * vit["embed_dim"] if self.is_resnet else vit["vision_width"]
        self.cls_id, self.sep_id = kwargs.get("cls_id"), kwargs.get("sep_id")
        H = config.hidden_size
        self.embeddings = _holder(word_embeddings=nn.Embedding(config.vocab_size, H, padding_idx=0),
                                  position_embeddings=nn.Embedding(config.max_position_embeddings, H, padding_idx=0),
                                  token_type_embeddings=nn.Embedding(config.type_vocab_size, H, padding_idx=0),
                                  LayerNorm=nn.LayerNorm(H, eps=1e-12))
        enc = _holder(layer=nn.ModuleList([_bert_layer(H, config.intermediate_size, 1e-12) for _ in range(config.num_hidden_layers)]),
                      visn_fc=_holder(visn_fc=nn.Linear(F, H), visn_layer_norm=nn.LayerNorm(H, eps=1e-12),
                                      box_fc=nn.Linear(4, H), box_layer_norm=nn.LayerNorm(H, eps=1e-12)),
                      visual_model=CLIP(vit["embed_dim"], vit["image_resolution"], vit["vision_layers"], vit["vision_width"],
                                        vit.get("vision_patch_size"), 77, 49408, 512, 8, 12,
                                        img_len=2, img_only=True))
        # embeddings the reference allocates regardless of the backbone (lxrt/modeling.py:828-831, 621-705); added to the
        # tower output on the ResNet path only ("RN" guard, 1014) but present in every checkpoint
        enc.visual_pos = _holder(x_position_embedding=nn.Embedding(25, F), y_position_embedding=nn.Embedding(25, F))
        enc.visual_token_type = _holder(token_type_embedding=nn.Embedding(5, F))
        # ViT: oracle decision; ResNet: hard-wired False (lxrt/modeling.py:783)
        enc.skip_last_layer = not self.is_resnet
        self.encoder = enc
        self.pooler = _holder(dense=nn.Linear(H, H))
        if self.topo_sort:
            self.classifier = _holder(dense=nn.Linear(H, H), out_proj=nn.Linear(H, self.num_labels))
        self.apply(lambda m: _init_bert_weights(m, config.initializer_range))  # 1464: re-initialises the tower's Linears too

    # ---- checkpoints (lxrt/modeling.py:1257-1450): local directories only -- no downloads, no tar archives, no TF
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, state_dict=None, cache_dir=None, from_tf=False, *inputs, **kwargs):
        """Directory with config.json + pytorch_model.bin (what save_pretrained writes).  As in the reference, a checkpoint whose
        keys carry a `bert.` or `roberta.` prefix is loaded into this (prefix-less) body, old gamma/beta LayerNorm names are
        renamed, missing / unexpected keys are tolerated and a shape mismatch raises RuntimeError."""
        if from_tf:
            raise NotImplementedError("TensorFlow checkpoints are not supported by the B200 drop-in")
        path = pretrained_model_name_or_path
        if not os.path.isdir(path):
            raise EnvironmentError("Model name '{}' was not found: only local checkpoint directories are supported".format(path))
        with open(os.path.join(path, "config.json"), "r", encoding="utf-8") as fh:
            d = json.load(fh)
        vocab = d.pop("vocab_size", d.pop("vocab_size_or_config_json_file", 30522))
        known = ("hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size", "hidden_act", "hidden_dropout_prob",
                 "attention_probs_dropout_prob", "max_position_embeddings", "type_vocab_size", "initializer_range")
        config = BertConfig(vocab, **{k: d[k] for k in known if k in d})
        model = cls(config, *inputs, **kwargs)
        if state_dict is None:
            state_dict = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu")
        sd = {}
        for k, v in state_dict.items():
            k = k.replace("gamma", "weight") if "gamma" in k else k
            k = k.replace("beta", "bias") if "beta" in k else k
            sd[k] = v
        for pre in ("bert.", "roberta."):
            if any(k.startswith(pre) for k in sd):
                sd = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
                break
        own = model.state_dict()
        errors = ["size mismatch for {}: copying a param with shape {} from checkpoint, the shape in current model is {}.".format(
            k, tuple(v.shape), tuple(own[k].shape)) for k, v in sd.items() if k in own and tuple(own[k].shape) != tuple(v.shape)]
        if errors:
            raise RuntimeError("Error(s) in loading state_dict for {}:\n\t{}".format(cls.__name__, "\n\t".join(errors)))
        model.load_state_dict(sd, strict=False)
        return model

    def save_pretrained(self, save_directory):
        assert os.path.isdir(save_directory), "Saving path should be a directory where the model and configuration can be saved"
        model_to_save = self.module if hasattr(self, "module") else self
        cfg = {k: v for k, v in vars(model_to_save.config).items() if isinstance(v, (int, float, str, bool, list, dict, type(None)))}
        with open(os.path.join(save_directory, "config.json"), "w", encoding="utf-8") as fh:
            json.dump(cfg, fh, indent=2, sort_keys=True)
        torch.save(model_to_save.state_dict(), os.path.join(save_directory, "pytorch_model.bin"))

    def _engine(self):
        def make_config():
            c = self.config
            cfg = dict(hidden_size=c.hidden_size, num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
                       intermediate_size=c.intermediate_size, vocab_size=c.vocab_size,
                       max_position_embeddings=c.max_position_embeddings, type_vocab_size=c.type_vocab_size)
            cfg["rn" if self.is_resnet else "vit"] = self.vit_config
            return cfg
        # built once; parameter values are pushed in place (dropin/_owner.py)
        return owned_engine(self, make_config, lambda: {"bert." + k: v for k, v in self.state_dict().items()},
                            self.pooler.dense.weight.device)

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, visual_feats=None, visual_attention_mask=None,
                pretraining_objective=None, labels=None):
        if pretraining_objective is not None or visual_attention_mask is not None:
            raise NotImplementedError("pre-training objectives are outside the scoped path (SURVEY §2 row 18)")
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        R = input_ids.shape[0]
        idx = None
        if self.topo_sort and visual_feats is not None and visual_feats.dim() == 5:   # [R, img_len, 3, S, S] (1516-1518)
            visual_feats = visual_feats.reshape(-1, *visual_feats.shape[2:])
        if visual_feats is not None:
            assert visual_feats.shape[0] == 2 * R, "BERSON mode feeds two images per pair row"
            idx = torch.arange(2 * R, dtype=torch.int32, device=visual_feats.device)
        with torch.no_grad():
            lang, visn, pooled = self._engine().inner_forward(input_ids, token_type_ids, attention_mask, visual_feats, idx,
                                                              want_pooled=True)
        if self.topo_sort:
            # RobertaClassificationHead on pooled_output.unsqueeze(1) (1586-1594): dense -> tanh -> out_proj (eval: no dropout)
            with torch.no_grad():
                eng = self._engine()
                hid = eng.linear(pooled, self.classifier.dense.weight, self.classifier.dense.bias, act=3)
                logits = eng.linear(hid, self.classifier.out_proj.weight, self.classifier.out_proj.bias)
            if labels is not None:
                return torch.nn.functional.cross_entropy(logits, labels.to(logits.device)), logits
            return (logits,)
        return (lang, visn), pooled
