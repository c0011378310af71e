"""models/berson/modeling_bert.py of the reference, hot-path surface only: BertConfig, BertModel,
HierarchicalAttention (parameter holder), BertForOrdering, berson_pointer_network, beam_search_pointer.

Same constructors, forward signatures and state_dict keys; the arithmetic runs in
libmsq_b200.so through multimodal_sequencing_b200.OrderingEngine.  Reference lines are cited per method.
Training: `forward(inputs)` in train() mode returns a loss whose .backward() fills p.grad of every nn.Parameter with the
gradients the CUDA backward pass produced (msq_train_step), so trainers/train.py:340-363 (loss.backward();
clip_grad_norm_(model.parameters()); optimizer.step(); model.zero_grad()) runs unchanged; `finetune_step` is the fused
device-side alternative (clip + transformers.AdamW inside the library, no weight round trip).  Dropout is not applied."""
import copy
import json
import os
import types
import warnings

import torch
import torch.nn as nn

from multimodal_sequencing_b200.engine import OrderingEngine, PairBatch
from multimodal_sequencing_b200.dropin._owner import mark_dirty, owned_engine
from .process_inputs_for_berson import prepare_berson_inputs
from .generator import Beam


class BertConfig(object):
    """models/berson/configuration_bert.py:46-112 (the attributes the path reads)."""

    def __init__(self, vocab_size_or_config_json_file=30522, hidden_size=768, num_hidden_layers=12,
                 num_attention_heads=12, intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
                 attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
                 initializer_range=0.02, layer_norm_eps=1e-12, **kwargs):
        self.vocab_size = vocab_size_or_config_json_file
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.hidden_act = hidden_act
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.max_position_embeddings = max_position_embeddings
        self.type_vocab_size = type_vocab_size
        self.initializer_range = initializer_range
        self.layer_norm_eps = layer_norm_eps
        self.num_labels = kwargs.pop("num_labels", 2)
        self.wrapper_model_with_heatmap = kwargs.pop("wrapper_model_with_heatmap", False)
        self.v_feature_size = kwargs.pop("v_feature_size", 1024)
        self.output_attentions = False
        self.output_hidden_states = False
        for k, v in kwargs.items():
            setattr(self, k, v)


def _holder(**mods):
    m = nn.Module()
    for k, v in mods.items():
        m.add_module(k, v)
    return m


def _bert_layer(H, inter, eps):
    """parameter names of BertLayer (modeling_bert.py:324-337)."""
    return _holder(
        attention=_holder(self=_holder(query=nn.Linear(H, H), key=nn.Linear(H, H), value=nn.Linear(H, H)),
                          output=_holder(dense=nn.Linear(H, H), LayerNorm=nn.LayerNorm(H, eps=eps))),
        intermediate=_holder(dense=nn.Linear(H, inter)),
        output=_holder(dense=nn.Linear(inter, H), LayerNorm=nn.LayerNorm(H, eps=eps)))


def _init_bert_weights(module, std):
    """BertPreTrainedModel._init_weights (modeling_bert.py:464-474)."""
    if isinstance(module, (nn.Linear, nn.Embedding)):
        module.weight.data.normal_(mean=0.0, std=std)
    elif isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)
    if isinstance(module, nn.Linear) and module.bias is not None:
        module.bias.data.zero_()


class _EngineOwner:
    """Owns the packed device model: built once, its weights refreshed in place (never rebuilt because values changed)."""

    def _engine_config(self):
        raise NotImplementedError

    def _engine_state(self):
        return self.state_dict()

    def engine(self):
        # built once; parameter VALUES are pushed in place (see dropin/_owner.py for when)
        return owned_engine(self, self._engine_config, self._engine_state, next(self.parameters()).device)

    def mark_dirty(self):
        """parameter values were changed behind the module's back in eval mode (`p.data` edits): re-upload before the next call"""
        mark_dirty(self)


class _DeviceBackward(torch.autograd.Function):
    """Bridge between the device-side backward pass and torch.autograd: forward passes the loss through, backward hands
    every parameter its slice of the flat gradient buffer (scaled by the incoming gradient, e.g. 1 / accumulation steps)."""

    @staticmethod
    def forward(ctx, loss, flat, spans, *params):
        ctx.flat, ctx.spans, ctx.shapes = flat, spans, [p.shape for p in params]
        return loss.detach().clone()

    @staticmethod
    def backward(ctx, g):
        out = [None, None, None]
        for span, shp in zip(ctx.spans, ctx.shapes):
            out.append(None if span is None else ctx.flat[span[0]:span[0] + span[1]].reshape(shp) * g)
        return tuple(out)


WEIGHTS_NAME, CONFIG_NAME = "pytorch_model.bin", "config.json"   # models/berson/file_utils / modeling_utils constants


class _Pretrained:
    """save_pretrained / from_pretrained of the reference's PreTrainedModel (models/berson/modeling_utils.py:190-400) for local
    directories and files -- the forms trainers/train.py uses (404, 2030-2035, 2194-2199).  No hub downloads, no TF checkpoints."""

    base_model_prefix = "bert"

    def save_pretrained(self, save_directory):
        assert os.path.isdir(save_directory), "Saving path should be a directory where the model and configuration can be saved"
        model_to_save = self.module if hasattr(self, "module") else self
        if model_to_save.__dict__.get("_flat") is not None:     # fused finetune_step ran: the fp32 masters live in the library
            model_to_save.pull_weights()
        cfg = {k: v for k, v in vars(model_to_save.config).items() if isinstance(v, (int, float, str, bool, list, dict, type(None)))}
        with open(os.path.join(save_directory, CONFIG_NAME), "w", encoding="utf-8") as fh:
            json.dump(cfg, fh, indent=2, sort_keys=True)
        torch.save(model_to_save.state_dict(), os.path.join(save_directory, WEIGHTS_NAME))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, **kwargs):
        config = kwargs.pop("config", None)
        state_dict = kwargs.pop("state_dict", None)
        output_loading_info = kwargs.pop("output_loading_info", False)
        for k in ("cache_dir", "force_download", "proxies"):
            kwargs.pop(k, None)
        if kwargs.pop("from_tf", False):
            raise NotImplementedError("TensorFlow checkpoints are not supported by the B200 drop-in")
        path = pretrained_model_name_or_path
        if config is None:
            cfg_file = os.path.join(path, CONFIG_NAME) if os.path.isdir(path) else path
            with open(cfg_file, "r", encoding="utf-8") as fh:
                d = json.load(fh)
            vocab = d.pop("vocab_size", d.pop("vocab_size_or_config_json_file", 30522))
            config = BertConfig(vocab, **d)
        if state_dict is None and path is not None:
            if os.path.isdir(path):
                archive = os.path.join(path, WEIGHTS_NAME)
                if not os.path.isfile(archive):
                    raise EnvironmentError("Error no file named {} found in directory {}".format(WEIGHTS_NAME, path))
            elif os.path.isfile(path):
                archive = path
            else:
                raise EnvironmentError("Model name '{}' was not found: only local directories / files are supported".format(path))
            state_dict = torch.load(archive, map_location="cpu")
        model = cls(config, *model_args, **kwargs)
        missing, unexpected, errors = [], [], []
        if state_dict is not None:
            # old TF-style LayerNorm names (modeling_utils.py: gamma -> weight, beta -> bias)
            state_dict = {k.replace("gamma", "weight").replace("beta", "bias") if ("gamma" in k or "beta" in k) else k: v
                          for k, v in state_dict.items()}
            pre = cls.base_model_prefix
            has_base, ckpt_has_base = hasattr(model, pre), any(k.startswith(pre + ".") for k in state_dict)
            target = model
            if not has_base and ckpt_has_base:          # bare inner model <- checkpoint of a model with heads
                state_dict = {k[len(pre) + 1:]: v for k, v in state_dict.items() if k.startswith(pre + ".")}
            elif has_base and not ckpt_has_base:        # model with heads <- checkpoint of the bare inner model
                target = getattr(model, pre)
            own = target.state_dict()
            for k, v in state_dict.items():
                if k in own and tuple(own[k].shape) != tuple(v.shape):
                    errors.append("size mismatch for {}: copying a param with shape {} from checkpoint, the shape in current model "
                                  "is {}.".format(k, tuple(v.shape), tuple(own[k].shape)))
            if errors:
                raise RuntimeError("Error(s) in loading state_dict for {}:\n\t{}".format(model.__class__.__name__, "\n\t".join(errors)))
            res = target.load_state_dict(state_dict, strict=False)
            missing, unexpected = list(res.missing_keys), list(res.unexpected_keys)
        model.eval()
        if output_loading_info:
            return model, {"missing_keys": missing, "unexpected_keys": unexpected, "error_msgs": errors}
        return model


class BertModel(nn.Module, _EngineOwner, _Pretrained):
    """Text-only inner encoder (modeling_bert.py:563-663): forward -> (sequence_output, sequence_output[:, 0])."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        H = config.hidden_size
        self.embeddings = _holder(word_embeddings=nn.Embedding(config.vocab_size, H, padding_idx=0),
                                  position_embeddings=nn.Embedding(config.max_position_embeddings, H),
                                  token_type_embeddings=nn.Embedding(config.type_vocab_size, H),
                                  LayerNorm=nn.LayerNorm(H, eps=config.layer_norm_eps))
        self.encoder = _holder(layer=nn.ModuleList(
            [_bert_layer(H, config.intermediate_size, config.layer_norm_eps) for _ in range(config.num_hidden_layers)]))
        self.apply(lambda m: _init_bert_weights(m, config.initializer_range))

    def _engine_config(self):
        c = self.config
        return dict(hidden_size=c.hidden_size, num_hidden_layers=c.num_hidden_layers,
                    num_attention_heads=c.num_attention_heads, intermediate_size=c.intermediate_size,
                    vocab_size=c.vocab_size, max_position_embeddings=c.max_position_embeddings,
                    type_vocab_size=c.type_vocab_size, vit=None)

    def _engine_state(self):
        return {"bert." + k: v for k, v in self.state_dict().items()}

    def forward(self, input_ids, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None):
        if position_ids is not None or head_mask is not None:
            raise NotImplementedError("position_ids / head_mask are not used on the ordering path")
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        with torch.no_grad():
            seq, _, _ = self.engine().inner_forward(input_ids, token_type_ids, attention_mask)
        return seq, seq[:, 0]


class HierarchicalAttention(nn.Module):
    """Parameter holder with the reference's names (modeling_bert.py:666-684); the pooling itself is
    csrc/pooling.cu, reached through BertForOrdering.encode."""

    def __init__(self, config, args=None):
        super().__init__()
        H = config.hidden_size
        self.linear_in_2 = nn.Linear(H, 1, bias=False)
        self.sentence_tran = nn.Linear(H, H)
        self.sentence_tran_2 = nn.Linear(H, 1)
        self.pairwise_relationship = nn.Linear(H, 2)
        self.h1_relationship = nn.Linear(H, 2)
        self.h2_relationship = nn.Linear(H, 2)
        self.args = args


class _ParaLayer(nn.Module):
    def __init__(self, H, ff):
        super().__init__()
        self.self_attn = _holder(linear_keys=nn.Linear(H, H), linear_values=nn.Linear(H, H), linear_query=nn.Linear(H, H),
                                 final_linear=nn.Linear(H, H))
        self.feed_forward = _holder(w_1=nn.Linear(H, ff), w_2=nn.Linear(ff, H), layer_norm=nn.LayerNorm(H, eps=1e-6))
        self.layer_norm = nn.LayerNorm(H, eps=1e-6)


class TransformerInterEncoder(nn.Module):
    """Parameter holder of models/berson/encoder.py:32-44."""

    def __init__(self, d_model, d_ff, heads, dropout, num_inter_layers=0):
        super().__init__()
        self.d_model, self.heads, self.d_ff, self.num_inter_layers = d_model, heads, d_ff, num_inter_layers
        self.transformer_inter = nn.ModuleList([_ParaLayer(d_model, d_ff) for _ in range(num_inter_layers)])
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)


class BertForOrdering(nn.Module, _EngineOwner, _Pretrained):
    """models/berson/modeling_bert.py:825-1402."""

    def __init__(self, config, args, inner_model=None, tokenizer=None, load_inner_model=False, **kwargs):
        super().__init__()
        if tokenizer is None:
            self.bert = BertModel(config)
        else:
            self.bert = inner_model if load_inner_model else BertModel(config)
            self.tokenizer = tokenizer
        self.config, self.args = config, args
        if getattr(config, "wrapper_model_with_heatmap", False):
            raise NotImplementedError("heat-map head (models/heatmap_module.py) is missing from the reference too")
        H = config.hidden_size
        self.num_labels, self.hidden_size = config.num_labels, H
        self.classifier = nn.Linear(H, config.num_labels)
        self.encoder = TransformerInterEncoder(H, args.ff_size, args.heads, args.para_dropout, args.inter_layers)
        self.key_linear = nn.Linear(H * 2, H)
        self.query_linear = nn.Linear(H, H)
        self.tanh_linear = nn.Linear(H, 1)
        self.decoder = nn.LSTM(H, H, batch_first=True)
        self.two_level_encoder = HierarchicalAttention(config, args=args)
        self.pairwise_loss_lam = args.pairwise_loss_lam
        # optional image pairwise objective (modeling_bert.py:897-898): Linear(v_feature_size, H) applied to the H-d first visual
        # token, i.e. the reference only runs it when v_feature_size == H (train.py:2022 hard-wires 1024 = RoBERTa-large)
        self.multimodal_loss = bool(getattr(args, "multimodal_loss", False))
        if self.multimodal_loss:
            self.img_projection = nn.Linear(config.v_feature_size, H)
        self.pw_k = nn.Linear((H + 2) * 4, H, False)
        # "other losses" (modeling_bert.py:904-911): the optional time-contrastive objective
        self.time_contrastive = "time_contrastive" in (getattr(args, "additional_wrapper_level_objectives", None) or ())
        for name, mod in self.named_modules():      # init_weights() (913): heads only, the inner model keeps its own
            if not name.startswith("bert"):
                _init_bert_weights(mod, config.initializer_range)

    # ---- engine plumbing ---------------------------------------------------------------------
    def _engine_config(self):
        c = self.config
        inner = self.bert
        vit = getattr(inner, "vit_config", None)
        rn = None
        if vit is not None and getattr(inner, "is_resnet", False):
            vit, rn = None, vit
        ic = getattr(inner, "config", c)
        # a HuggingFace AutoModel as the inner encoder (trainers/train.py:1928-1933, load_inner_model=True): same parameter names
        # as the vendored BertModel, but outputs[1] is the tanh pooler -- the engine then pools the pair CLS the same way
        hf_inner = not isinstance(inner, BertModel) and vit is None and rn is None and getattr(inner, "pooler", None) is not None
        return dict(rn=rn, cls_pooler=hf_inner, hidden_size=c.hidden_size, num_hidden_layers=ic.num_hidden_layers,
                    num_attention_heads=ic.num_attention_heads, intermediate_size=ic.intermediate_size,
                    vocab_size=ic.vocab_size, max_position_embeddings=ic.max_position_embeddings,
                    type_vocab_size=getattr(ic, "type_vocab_size", 2), vit=vit, para_heads=self.args.heads,
                    para_ff=self.args.ff_size, para_layers=self.args.inter_layers)

    def equip(self, critic):
        self.critic = critic

    def rela_encode(self, cls_output_matrix_nn, cls_score_matrix_nn):
        """modeling_bert.py:919-925 (tiny glue, kept as the reference writes it)."""
        return torch.cat((cls_output_matrix_nn, torch.softmax(cls_score_matrix_nn, -1)), -1)

    def history_encode(self, cls_output_matrix_nn, cls_score_matrix_nn_his1, cls_score_matrix_nn_his2):
        """modeling_bert.py:927-935."""
        return (torch.cat((cls_output_matrix_nn, torch.softmax(cls_score_matrix_nn_his1, -1)), -1),
                torch.cat((cls_output_matrix_nn, torch.softmax(cls_score_matrix_nn_his2, -1)), -1))

    def forward(self, inputs):
        """modeling_bert.py:937-941 -> (loss,): pointer NLL / (N-1) + lam * pairwise NLL / P (943-1174).  eval(): the loss
        value only.  train(): the value AND the whole backward pass run on the device (msq_train_step); the returned tensor
        carries an autograd node that hands those gradients to p.grad when the caller does loss.backward()."""
        berson_inputs = inputs
        if getattr(self, "tokenizer", None) is not None:
            berson_inputs = prepare_berson_inputs(berson_inputs, self.tokenizer, args=self.args)
        return self._forward(**berson_inputs)

    def _pair_batch(self, input_ids, attention_mask, token_type_ids, pairs_list, passage_length, sep_positions, ground_truth,
                    pairwise_labels, images, _pair_batch):
        if _pair_batch is not None:
            return _pair_batch
        B, P, Lt = input_ids.shape
        if passage_length.numel() > 1 and bool((passage_length != passage_length.reshape(-1)[0]).any()):
            raise ValueError("manuals with different step counts in one batch are not supported by the B200 path "
                             "(the reference's data loaders batch manuals of one length): %s" % passage_length.tolist())
        img = idx = None
        if images is not None:
            img = images.reshape(B * P * 2, *images.shape[3:])
            idx = torch.arange(B * P * 2, dtype=torch.int32).reshape(B, P, 2)
        return PairBatch(input_ids, attention_mask, token_type_ids, sep_positions, pairs_list, pairwise_labels, ground_truth,
                         int(passage_length[0]), img, idx)

    def _apply_dropout_config(self, eng):
        """train(): the reference's dropout probabilities (BertConfig.hidden_dropout_prob / attention_probs_dropout_prob of the
        INNER model's config, args.para_dropout for the paragraph encoder), seeded from torch's RNG seed once per engine."""
        inner_cfg = getattr(self.bert, "config", self.config)
        probs = (float(getattr(inner_cfg, "hidden_dropout_prob", 0.0) or 0.0),
                 float(getattr(inner_cfg, "attention_probs_dropout_prob", 0.0) or 0.0),
                 float(getattr(self.args, "para_dropout", 0.0) or 0.0))
        key = (id(eng), probs)
        if self.__dict__.get("_drop_key") != key:
            eng.set_dropout(*probs, seed=torch.initial_seed() & 0xFFFFFFFF)
            self.__dict__["_drop_key"] = key
        if self.multimodal_loss and self.__dict__.get("_ml_key") != id(eng):
            if self.img_projection.in_features != self.hidden_size:
                raise RuntimeError("multimodal_loss: img_projection expects %d features but the visual token has %d (the reference "
                                   "fails on the same shapes)" % (self.img_projection.in_features, self.hidden_size))
            eng.set_multimodal_loss(True)
            self.__dict__["_ml_key"] = id(eng)

    def _pull_bn_buffers(self, eng, count=True):
        """ModifiedResNet tower: a training forward moved the library's running_mean / running_var (nn.BatchNorm2d.train()
        bookkeeping, clip/model.py:10-53); mirror them in this module's buffers so that state_dict() / the next weight refresh
        carry them, and advance num_batches_tracked as torch does."""
        with torch.no_grad():
            for n, b in self.named_buffers():
                if n.endswith("running_mean") or n.endswith("running_var"):
                    b.copy_(eng.read_param(n, tuple(b.shape)))
                elif count and n.endswith("num_batches_tracked"):
                    b += 1

    @staticmethod
    def _time_contrastive_triplets(target):
        """modeling_bert.py:1176-1209: anchor position, a neighbouring positive, a negative at least two positions away -- drawn
        with numpy's global generator in the reference's call order -- each mapped through the ground-truth order of its manual."""
        import numpy as np
        target = target.tolist()
        out = []
        for tgt in target:
            n = len(tgt)
            anchor = np.random.choice(list(range(n)), 1, replace=False)[0]
            positive = np.random.choice([i for i in (anchor - 1, anchor + 1) if 0 <= i < n], 1, replace=False)[0]
            negative = np.random.choice([j for j in range(n) if abs(j - anchor) >= 2], 1, replace=False)[0]
            out.append([tgt[anchor], tgt[positive], tgt[negative]])
        return torch.tensor(out, dtype=torch.int32)

    def _train_forward(self, pb):
        eng = self.engine()
        self._apply_dropout_config(eng)
        flat = eng.new_grad_buffer()
        trip = self._time_contrastive_triplets(pb.ground_truth) if self.time_contrastive else None
        with torch.no_grad():
            loss = eng.train_step(pb, flat, self.pairwise_loss_lam, triplets=trip)
        self._pull_bn_buffers(eng)
        lay = {n: (o, k) for n, o, k, _ in eng.train_layout()}
        named = list(self.named_parameters())
        spans = [lay.get(n) for n, _ in named]    # None: the reference gives this parameter no gradient either
        with torch.enable_grad():
            return _DeviceBackward.apply(loss, flat, spans, *[p for _, p in named])

    def finetune_step(self, inputs, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0, grad_scale=None,
                      allreduce=True):
        """Fused alternative to loss.backward() + clip_grad_norm_ + AdamW.step() (trainers/train.py:340-363): one
        msq_train_step, one all-reduce of the flat gradient buffer when torch.distributed is initialised, one
        msq_adamw_step on the fp32 masters inside the library.  The nn.Parameters of this module go stale until
        pull_weights() (call it before state_dict() / save_pretrained()).  Returns the loss (0-d device tensor)."""
        from multimodal_sequencing_b200.sharding import allreduce_gradients
        berson_inputs = inputs
        if getattr(self, "tokenizer", None) is not None and "pairs_list" not in inputs:
            berson_inputs = prepare_berson_inputs(inputs, self.tokenizer, args=self.args)
        b = berson_inputs
        pb = self._pair_batch(b["input_ids"], b.get("attention_mask"), b.get("token_type_ids"), b.get("pairs_list"),
                              b["passage_length"], b.get("sep_positions"), b.get("ground_truth"), b.get("pairwise_labels"),
                              b.get("images"), b.get("_pair_batch"))
        eng = self.engine()
        self._apply_dropout_config(eng)
        flat = self.__dict__.get("_flat")
        if flat is None or flat.device != eng.device or flat.numel() != eng.new_grad_buffer().numel():
            flat = self.__dict__["_flat"] = eng.new_grad_buffer()
        flat.zero_()
        self.__dict__["_lib_masters"] = True   # from here on the library's fp32 masters are the truth (until pull_weights)
        trip = self._time_contrastive_triplets(pb.ground_truth) if self.time_contrastive else None
        with torch.no_grad():
            loss = eng.train_step(pb, flat, self.pairwise_loss_lam, triplets=trip)
            scale = allreduce_gradients(flat) if allreduce else 1.0
            eng.adamw_step(flat, lr, betas, eps, weight_decay, max_grad_norm, scale if grad_scale is None else grad_scale)
        for n, b in self.named_buffers():
            if n.endswith("num_batches_tracked"):
                b += 1
        return loss

    def pull_weights(self):
        """Copy the library's fp32 masters back into this module's nn.Parameters (after finetune_step)."""
        eng = self.engine()
        with torch.no_grad():
            for n, p in self.named_parameters():
                if any(n == k for k, _, _, _ in eng.train_layout()):
                    p.data.copy_(eng.read_param(n, tuple(p.shape)))
        self._pull_bn_buffers(eng, count=False)
        # both sides hold the same values again: the nn.Parameters are the masters from here on
        tensors = list(self.parameters()) + list(self.buffers())
        self.__dict__["_eng_vers"] = tuple((t._version, t.data_ptr()) for t in tensors)
        self.__dict__["_lib_masters"], self.__dict__["_eng_dirty"] = False, False

    def _forward(self, input_ids, attention_mask=None, token_type_ids=None, pairs_list=None, passage_length=None,
                 pairs_num=None, sep_positions=None, ground_truth=None, mask_cls=None, pairwise_labels=None, cuda=None,
                 head_mask=None, images=None, _pair_batch=None):
        pb = self._pair_batch(input_ids, attention_mask, token_type_ids, pairs_list, passage_length, sep_positions, ground_truth,
                              pairwise_labels, images, _pair_batch)
        if self.training:
            return (self._train_forward(pb),)
        with torch.no_grad():
            loss = self.engine().training_loss(pb, self.pairwise_loss_lam)
            if self.time_contrastive:   # evaluation: the same extra term on the sentence vectors (a [B, H]-sized host-side op)
                trip = self._time_contrastive_triplets(pb.ground_truth).long().to(loss.device)
                sents = self.engine().encode(pb)["sents"]
                ar = torch.arange(sents.shape[0], device=sents.device)
                a, p, n = (sents[ar, trip[:, i]] for i in range(3))
                loss = loss + 0.1 * torch.nn.functional.triplet_margin_loss(a, p, n, margin=1.0, p=2)
            if self.multimodal_loss:    # evaluation: the image pairwise term (1218-1225) on the [R, 2] logits
                _, score_img = self._image_pair_scores(pb)
                nll = -torch.log_softmax(score_img, -1).gather(1, pb.pairwise_labels.reshape(-1, 1).to(score_img.device))
                B, P = pb.input_ids.shape[:2]
                loss = loss + self.pairwise_loss_lam * (nll.reshape(B, P).sum(-1) / (P + 1e-20)).sum() / B
            return (loss,)

    def _image_pair_scores(self, pb):
        """modeling_bert.py:1295-1297, 1359-1362: the visual half of the inner model's output and
        pairwise_relationship(img_projection(visn[:, 0])).  Evaluation-only helper (a second pass through the inner model)."""
        eng = self.engine()
        B, P, Lt = pb.input_ids.shape
        _, visn, _ = eng.inner_forward(pb.input_ids.reshape(B * P, Lt), pb.token_type_ids.reshape(B * P, Lt),
                                       pb.attention_mask.reshape(B * P, Lt), pb.images, pb.img_index.reshape(-1))
        u = eng.linear(visn[:, 0].contiguous(), self.img_projection.weight, self.img_projection.bias)
        rel = self.two_level_encoder.pairwise_relationship
        return visn, eng.linear(u, rel.weight, rel.bias)

    # ---- encode ------------------------------------------------------------------------------
    def encode(self, input_ids, attention_mask=None, token_type_ids=None, pairs_list=None, passage_length=None,
               pairs_num=None, sep_positions=None, ground_truth=None, mask_cls=None, pairwise_labels=None, cuda=None,
               head_mask=None, images=None, _pair_batch=None):
        """modeling_bert.py:1239-1366 -> the same 10-tuple."""
        B, P, Lt = input_ids.shape
        N = int(passage_length[0])
        if _pair_batch is not None:
            pb = _pair_batch
        else:
            img = idx = None
            if images is not None:   # materialised [B,P,2,3,S,S] as the reference passes it
                img = images.reshape(B * P * 2, *images.shape[3:])
                idx = torch.arange(B * P * 2, dtype=torch.int32).reshape(B, P, 2)
            pb = PairBatch(input_ids, attention_mask, token_type_ids, sep_positions, pairs_list, pairwise_labels,
                           ground_truth, N, img, idx)
        with torch.no_grad():
            e = self.engine().encode(pb)
            sents, cls_score = e["sents"], e["cls_score"]
            if self.multimodal_loss:   # 1359-1364: (sentence vectors, visual tokens) and (text, image) pair logits
                visn, score_img = self._image_pair_scores(pb)
                sents, cls_score = (sents, visn), (cls_score, score_img)
        hcn = (e["h0"].unsqueeze(0), e["c0"].unsqueeze(0))
        return (sents, e["para"], hcn, e["key"], e["cls"], e["cls_mat"], cls_score, e["score_mat"], e["his1"], e["his2"])

    # ---- one decode step ---------------------------------------------------------------------
    def step(self, prev_y, prev_handc, original_keys, mask, rela_vec, rela_mask, hist_left1, hist_left2, l1_mask, l2_mask):
        """modeling_bert.py:1368-1402 (rela_vec is zeroed in place, as there)."""
        with torch.no_grad():
            return self.engine().decode_step(prev_y, prev_handc, original_keys, mask, rela_vec, rela_mask, hist_left1,
                                             hist_left2, l1_mask, l2_mask)


def berson_pointer_network(args, model, tokenizer, inputs):
    """modeling_bert.py:1405-1408."""
    berson_inputs = prepare_berson_inputs(inputs, tokenizer, args=args)
    return beam_search_pointer(args, model, **berson_inputs)


def beam_search_pointer(args, model, input_ids, attention_mask=None, token_type_ids=None, pairs_list=None,
                        passage_length=None, pairs_num=None, sep_positions=None, ground_truth=None, mask_cls=None,
                        pairwise_labels=None, cuda=None, head_mask=None, images=None, _pair_batch=None):
    """modeling_bert.py:1411-1552: encode + beam search, fused on the device.  One manual -> list[int] like the
    reference; a batch of manuals -> list of lists (batched beam search is a new capability)."""
    (sents, _, hcn, key, _, cls_mat, _, score_mat, _, _) = model.encode(
        input_ids, attention_mask, token_type_ids, pairs_list, passage_length, pairs_num, sep_positions, ground_truth,
        mask_cls, pairwise_labels, cuda, head_mask, images=images, _pair_batch=_pair_batch)
    if isinstance(sents, tuple):   # args.multimodal_loss (1432-1433)
        sents, _ = sents
    N = int(passage_length[0])
    enc = dict(sents=sents, key=key, h0=hcn[0], cls_mat=cls_mat, score_mat=score_mat)
    perm = model.engine().beam_search(enc, N, args.beam_size).cpu().tolist()
    return perm[0] if len(perm) == 1 else perm
