"""Host side of the B200 step-ordering path: weight hand-over, pair expansion, and thin wrappers over
the C ABI (include/msq_b200.h).  PyTorch is used for device memory and streams only.

Reference call sites mirrored here (telin0411/multimodal_sequencing):
  prepare_berson_inputs      models/berson/process_inputs_for_berson.py:13-79  -> OrderingEngine.prepare
  BertForOrdering.encode     models/berson/modeling_bert.py:1239-1366         -> OrderingEngine.encode
  beam_search_pointer        models/berson/modeling_bert.py:1411-1552         -> OrderingEngine.beam_search
  berson_pointer_network     models/berson/modeling_bert.py:1405-1408         -> OrderingEngine.order / order_host
"""
import ctypes as C
import itertools
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib

_DEAD = ("visual_model.transformer.", "visual_model.token_embedding", "visual_model.positional_embedding",
         "visual_model.ln_final", "visual_model.text_projection", "visual_model.logit_scale")


def pairs_generator(n):
    """All (i<j) pairs in lexicographic order, then the mirrored list
    (models/berson/process_inputs_for_berson.py:246-261)."""
    one = [[a, b] for a, b in itertools.combinations(range(n), 2)]
    return one + [[b, a] for a, b in one], 2 * len(one)


@dataclass
class PairBatch:
    """One batch of manuals expanded to ordered pairs.  Images stay UNIQUE ([B*N,3,S,S]); the pair ->
    image relation is the int32 table `img_index` (the reference materialises [B,P,2,3,S,S] instead,
    process_inputs_for_berson.py:82-97)."""
    input_ids: torch.Tensor        # [B,P,Lt] int64
    attention_mask: torch.Tensor   # [B,P,Lt] int64
    token_type_ids: torch.Tensor   # [B,P,Lt] int64
    sep_positions: torch.Tensor    # [B,P,2]  int64
    pairs_list: torch.Tensor       # [B,P,2]  int64
    pairwise_labels: torch.Tensor  # [B,P]    int64
    ground_truth: torch.Tensor     # [B,N]    int64
    n_steps: int
    images: Optional[torch.Tensor] = None     # [B*N,3,S,S] fp32
    img_index: Optional[torch.Tensor] = None  # [B,P,2] int32

    @property
    def B(self):
        return self.input_ids.shape[0]

    @property
    def Lt(self):
        return self.input_ids.shape[2]

    _DTYPES = dict(images=torch.float32, img_index=torch.int32)

    def to(self, device=None, non_blocking=False):
        """Move to `device` (None: stay) AND canonicalise what the kernels read through raw pointers: int64 ids / masks /
        token types / sep positions / labels, int32 img_index, fp32 images, all contiguous (a sliced or transposed view or an
        int32 id tensor would otherwise be misread silently).  Tensors already in that form are passed through untouched."""
        kw = {}
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                v = v.to(device=device if device is not None else v.device, dtype=self._DTYPES.get(k, torch.long),
                         non_blocking=non_blocking).contiguous()
            kw[k] = v
        return PairBatch(**kw)

    def reference_dict(self, materialize_images=True):
        """The dict prepare_berson_inputs returns (process_inputs_for_berson.py:46-79)."""
        B, P = self.input_ids.shape[:2]
        d = {
            "input_ids": self.input_ids, "attention_mask": self.attention_mask, "token_type_ids": self.token_type_ids,
            "pairs_list": self.pairs_list,
            "passage_length": torch.full((B,), self.n_steps, dtype=torch.long, device=self.input_ids.device),
            "pairs_num": torch.full((B,), P, dtype=torch.long, device=self.input_ids.device),
            "sep_positions": self.sep_positions, "ground_truth": self.ground_truth,
            "mask_cls": torch.ones((B, self.n_steps), dtype=torch.long, device=self.input_ids.device),
            "pairwise_labels": self.pairwise_labels, "cuda": "cuda:0",
        }
        if self.images is not None and materialize_images:
            d["images"] = self.images[self.img_index.long()]  # [B,P,2,3,S,S]
        return d


def prepare_pairs(input_ids, labels, n_steps, images=None, cls_id=101, sep_id=102, pad_id=0):
    """Manual -> P = N(N-1) ordered pairs (host logic of process_inputs_for_berson.py:113-243,264-368).

    input_ids [B,L] int64 rows of concatenated `[CLS] ... [SEP]` steps; labels [B,N]; images [B,N,3,S,S]."""
    input_ids = torch.as_tensor(input_ids).cpu()
    labels = torch.as_tensor(labels).cpu()
    B = input_ids.shape[0]
    N = n_steps
    pairs, P = pairs_generator(N)
    pi = torch.tensor(pairs)  # [P,2]
    is_cls, is_sep = input_ids == cls_id, input_ids == sep_id
    if not (is_cls.sum(1) == N).all() or not (is_sep.sum(1) == N).all():
        raise AssertionError("every manual must hold exactly max_story_length [CLS]..[SEP] steps")
    starts = is_cls.nonzero()[:, 1].view(B, N)
    ends = is_sep.nonzero()[:, 1].view(B, N)
    lens = ends - starts + 1                               # [B,N]
    l1, l2 = lens[:, pi[:, 0]], lens[:, pi[:, 1]]          # [B,P]
    Lt = int((l1 + l2).max())
    pos = torch.arange(Lt)[None, None, :]
    in1 = pos < l1[..., None]
    in2 = (pos >= l1[..., None]) & (pos < (l1 + l2)[..., None])
    src = torch.where(in1, starts[:, pi[:, 0]][..., None] + pos,
                      starts[:, pi[:, 1]][..., None] + pos - l1[..., None]).clamp_(0, input_ids.shape[1] - 1)
    ids = torch.gather(input_ids[:, None, :].expand(B, P, -1), 2, src)
    valid = in1 | in2
    ids = torch.where(valid, ids, torch.full_like(ids, pad_id))
    am = torch.where(valid, torch.ones_like(ids), torch.full_like(ids, pad_id))
    tt = (in2 & (cls_id != 0)).long()
    sep = torch.stack([l1 - 1, l1 + l2 - 1], -1)
    # pairwise label = 1 iff the ground-truth position of i precedes that of j (162-172)
    posn = torch.argsort(labels, dim=1)  # posn[b, step] = index of `step` inside labels[b]
    plab = (posn[:, pi[:, 0]] < posn[:, pi[:, 1]]).long()
    img = idx = None
    if images is not None:
        images = torch.as_tensor(images)
        img = images.reshape(B * N, *images.shape[2:]).float().contiguous()
        idx = (torch.arange(B)[:, None, None] * N + pi[None]).int().contiguous()
    return PairBatch(ids.contiguous(), am.contiguous(), tt.contiguous(), sep.contiguous(),
                     pi[None].expand(B, P, 2).contiguous(), plab, labels.clone(), N, img, idx)


# msq_config.precise: 0 = bf16 operands (fastest), 1 = fp32 FFMA (CUDA cores), 2 = bf16x3 (hi + lo bf16 operands,
# three tcgen05 MMAs per product: fp32-grade results at tensor-core speed)
PRECISION = {False: 0, True: 1, 2: 2, "bf16": 0, "fp32": 1, "bf16x3": 2}


def precision_code(precise):
    try:
        return PRECISION[precise]
    except (KeyError, TypeError):
        raise ValueError("precise=%r: expected False/'bf16', True/'fp32' or 2/'bf16x3'" % (precise,))


class OrderingEngine:
    """Owns one packed device model (msq_model) and runs the path through the C ABI.

    precise: False / "bf16" (bf16 tensor-core operands), True / "fp32" (CUDA-core fp32 parity mode) or
    2 / "bf16x3" (split-bf16 operands on the tensor cores; evaluation only; ViT, ModifiedResNet or text-only models)."""

    def __init__(self, state_dict, config, precise=False, device="cuda:0", inner_prefix="bert."):
        if not torch.cuda.is_available():
            raise RuntimeError("multimodal_sequencing_b200 needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device(device)
        vit, rn = config.get("vit"), config.get("rn")
        if vit and rn:
            raise ValueError("config names both a ViT and a ResNet visual tower")
        self.cfg = dict(config)
        self.H = config["hidden_size"]
        c = _lib.MsqConfig(
            hidden=config["hidden_size"], layers=config["num_hidden_layers"], heads=config["num_attention_heads"],
            inter=config["intermediate_size"], vocab=config["vocab_size"], max_pos=config["max_position_embeddings"],
            type_vocab=config.get("type_vocab_size", 2), vit_width=vit["vision_width"] if vit else 0,
            vit_layers=vit["vision_layers"] if vit else 0, vit_patch=vit["vision_patch_size"] if vit else 0,
            vit_res=vit["image_resolution"] if vit else 0, para_heads=config.get("para_heads", 8),
            para_ff=config.get("para_ff", 3072), para_layers=config.get("para_layers", 2), precise=precision_code(precise),
            reserved=1 if config.get("cls_pooler") else 0)
        if rn:
            # CLIP ModifiedResNet (clip/model.py:128-187): the tower hands 2*embed_dim features per token to visn_fc
            blocks = tuple(rn["vision_layers"])
            c.vit_width, c.vit_layers, c.vit_patch, c.vit_res = 2 * rn["embed_dim"], 0, 32, rn["image_resolution"]
            c.rn_width, c.rn_embed = rn["vision_width"], rn["embed_dim"]
            c.rn_blocks0, c.rn_blocks1, c.rn_blocks2, c.rn_blocks3 = blocks
        self.multimodal = bool(vit or rn)
        self.vit, self.rn = vit, rn
        # visual tokens per pair row and their feature width as the tower returns them
        self._grid = (vit["image_resolution"] // vit["vision_patch_size"]) if vit else (rn["image_resolution"] // 32 if rn else 0)
        self._vis_width = vit["vision_width"] if vit else (2 * rn["embed_dim"] if rn else 0)
        self.precise = precision_code(precise) == 1
        self.precision = ("bf16", "fp32", "bf16x3")[precision_code(precise)]
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.msq_model_create(C.byref(c), C.byref(self._h)))
            st = self._stream()
            for k, v in state_dict.items():
                if not torch.is_tensor(v) or not v.is_floating_point() or any(d in k for d in _DEAD):
                    continue
                if inner_prefix != "bert." and k.startswith(inner_prefix):
                    k = "bert." + k[len(inner_prefix):]
                t = v.detach().to(self.device, torch.float32).contiguous()
                _lib.check(self.lib.msq_model_set_weight(self._h, k.encode(), t.data_ptr(), t.numel(), st))
            torch.cuda.synchronize(self.device)
            _lib.check(self.lib.msq_model_pack(self._h, st))

    def refresh_weights(self, state_dict, inner_prefix="bert."):
        """Overwrite the registered fp32 masters IN PLACE with the tensors of `state_dict` (same keys / shapes as at
        construction) and re-derive every packed copy -- no model rebuild, no allocation.  For trainers that update
        nn.Parameter.data themselves (transformers.AdamW 3.4 / models/berson/optimization.py do): call it before the next
        forward.  Returns the number of tensors uploaded."""
        n = 0
        with torch.cuda.device(self.device):
            st = self._stream()
            for k, v in state_dict.items():
                if not torch.is_tensor(v) or not v.is_floating_point() or any(d in k for d in _DEAD):
                    continue
                if inner_prefix != "bert." and k.startswith(inner_prefix):
                    k = "bert." + k[len(inner_prefix):]
                t = v.detach()
                if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.to(self.device, torch.float32).contiguous()
                _lib.check(self.lib.msq_model_update_weight(self._h, k.encode(), t.data_ptr(), t.numel(), st))
                n += 1
            _lib.check(self.lib.msq_model_refresh(self._h, st))
        return n

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self.lib.msq_model_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()

    def launch_count(self):
        return int(self.lib.msq_launch_count())

    prepare = staticmethod(prepare_pairs)

    # ------------------------------------------------------------------------------------------
    def vit_forward(self, images, img_index, R):
        """CLIP pair tower: VisualTransformer (models/CLIP/clip/model.py:262-305, skip_last_layer=True) or
        ModifiedResNet + AttentionPool2d (model.py:171-187, 71-125, skip_last_layer=False)."""
        g = self._grid
        out = torch.empty(R, 1 + 2 * g * g, self._vis_width, device=self.device)
        images = images.to(self.device, torch.float32).contiguous()
        img_index = img_index.to(self.device, torch.int32).contiguous()
        _lib.check(self.lib.msq_vit_forward(self._h, self._p(images), images.shape[0], self._p(img_index), R,
                                            self._p(out), self._stream()))
        return out

    def inner_forward(self, ids, tt, mask, images=None, img_index=None, want_pooled=False):
        """LXRTModel.forward / BertModel.forward on R pair rows -> (lang, visn|None, pooled|None)."""
        ids = ids.to(self.device, torch.long).contiguous()
        tt = tt.to(self.device, torch.long).contiguous()
        mask = mask.to(self.device, torch.long).contiguous()
        R, Lt = ids.shape
        lang = torch.empty(R, Lt, self.H, device=self.device)
        visn = pooled = None
        n_img = 0
        if self.multimodal and images is not None:
            images = images.to(self.device, torch.float32).contiguous()
            img_index = img_index.to(self.device, torch.int32).contiguous()
            g = self._grid
            visn = torch.empty(R, 1 + 2 * g * g, self.H, device=self.device)
            n_img = images.shape[0]
        if want_pooled:
            pooled = torch.empty(R, self.H, device=self.device)
        _lib.check(self.lib.msq_inner_forward(self._h, self._p(ids), self._p(tt), self._p(mask), R, Lt,
                                              self._p(images if visn is not None else None), n_img,
                                              self._p(img_index if visn is not None else None), self._p(lang),
                                              self._p(visn), self._p(pooled), self._stream()))
        return lang, visn, pooled

    def encode(self, batch: PairBatch, want_top_vec=False):
        """BertForOrdering.encode -> dict with the reference's 10-tuple tensors (fp32, on device)."""
        b = batch.to(self.device)
        B, P, Lt = b.input_ids.shape
        N, H, dev = b.n_steps, self.H, self.device
        o = dict(sents=torch.empty(B, N, H, device=dev), para=torch.empty(B, N, H, device=dev),
                 h0=torch.empty(B, H, device=dev), key=torch.empty(B, N, H, device=dev),
                 cls=torch.empty(B * P, H, device=dev), cls_mat=torch.empty(B, N, N, H, device=dev),
                 cls_score=torch.empty(B * P, 2, device=dev), score_mat=torch.empty(B, N, N, 2, device=dev),
                 his1=torch.empty(B, N, N, 2, device=dev), his2=torch.empty(B, N, N, 2, device=dev))
        if want_top_vec:
            o["top_vec"] = torch.empty(B * P, Lt, H, device=dev)
        eo = _lib.MsqEncodeOut(**{k: (o[k].data_ptr() if k in o else None) for k, _ in _lib.MsqEncodeOut._fields_})
        n_img = 0 if b.images is None else b.images.shape[0]
        _lib.check(self.lib.msq_encode(self._h, self._p(b.input_ids), self._p(b.token_type_ids),
                                       self._p(b.attention_mask), self._p(b.sep_positions), B, N, Lt,
                                       self._p(b.images), n_img, self._p(b.img_index), C.byref(eo), self._stream()))
        o["c0"] = torch.zeros_like(o["h0"])
        return o

    def linear(self, x, weight, bias=None, act=0):
        """y = act(x W^T + b) in fp32 on the FFMA GEMM (small heads, e.g. the topo-sort classifier).
        act: 0 none, 1 erf-GELU, 2 QuickGELU, 3 tanh, 4 tanh-GELU."""
        x = x.to(self.device, torch.float32).contiguous()
        M, K = x.shape
        N = weight.shape[0]
        Np, Kp = (N + 3) // 4 * 4, (K + 15) // 16 * 16
        w = torch.zeros(Np, Kp, device=self.device)
        w[:N, :K] = weight.detach().to(self.device, torch.float32)
        b = torch.zeros(Np, device=self.device)
        if bias is not None:
            b[:N] = bias.detach().to(self.device, torch.float32)
        if Kp != K:
            x = torch.nn.functional.pad(x, (0, Kp - K)).contiguous()
        out = torch.empty(M, Np, device=self.device)
        _lib.check(self.lib.msq_gemm(0, self._p(x), self._p(w), self._p(b), None, self._p(out), M, Np, Kp, act, self._stream()))
        return out[:, :N]

    # ---- fine-tuning of the inner encoder: training forward / backward / AdamW -----------
    def train_layout(self):
        """[(state_dict key, offset, numel, decays)] of the flat gradient buffer (msq_train_param_info)."""
        if getattr(self, "_layout", None) is None:
            n = int(self.lib.msq_train_param_count(self._h, self._stream()))
            if n < 0:
                _lib.check(1)
            out = []
            name, off, numel, dec = C.c_char_p(), C.c_int64(), C.c_int64(), C.c_int32()
            for i in range(n):
                _lib.check(self.lib.msq_train_param_info(self._h, i, C.byref(name), C.byref(off), C.byref(numel), C.byref(dec)))
                out.append((name.value.decode(), off.value, numel.value, bool(dec.value)))
            self._layout = out
            self._grad_numel = int(self.lib.msq_train_grad_numel(self._h, self._stream()))
        return self._layout

    def new_grad_buffer(self):
        """Zeroed flat fp32 gradient buffer (one torch tensor: all-reduce it as a whole in data-parallel runs)."""
        self.train_layout()
        return torch.zeros(self._grad_numel, device=self.device, dtype=torch.float32)

    def grads_by_name(self, flat):
        """Views of the flat buffer keyed by the reference's state_dict names."""
        return {n: flat[o:o + k] for n, o, k, _ in self.train_layout()}

    def inner_forward_train(self, ids, tt, mask, images=None, img_index=None):
        """LXRTModel.forward / BertModel.forward in training mode (activations recorded for inner_backward)."""
        ids = ids.to(self.device, torch.long).contiguous()
        tt = tt.to(self.device, torch.long).contiguous()
        mask = mask.to(self.device, torch.long).contiguous()
        R, Lt = ids.shape
        lang = torch.empty(R, Lt, self.H, device=self.device)
        visn, n_img = None, 0
        if self.multimodal and images is not None:
            images = images.to(self.device, torch.float32).contiguous()
            img_index = img_index.to(self.device, torch.int32).contiguous()
            g = self._grid
            visn = torch.empty(R, 1 + 2 * g * g, self.H, device=self.device)
            n_img = images.shape[0]
        self._train_keep = (images, img_index)   # the backward pass re-reads the images (patch-embedding weight gradient)
        _lib.check(self.lib.msq_inner_forward_train(self._h, self._p(ids), self._p(tt), self._p(mask), R, Lt,
                                                    self._p(images if visn is not None else None), n_img,
                                                    self._p(img_index if visn is not None else None), self._p(lang),
                                                    self._p(visn), self._stream()))
        return lang, visn

    def inner_backward(self, d_lang, d_visn, grads):
        """grads += dL/dparam for the recorded forward, given dL/d(lang), dL/d(visn)."""
        d_lang = None if d_lang is None else d_lang.to(self.device, torch.float32).contiguous()
        d_visn = None if d_visn is None else d_visn.to(self.device, torch.float32).contiguous()
        _lib.check(self.lib.msq_inner_backward(self._h, self._p(d_lang), self._p(d_visn), self._p(grads), self._stream()))
        return grads

    def adamw_step(self, grads, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0, grad_scale=1.0):
        """clip_grad_norm_ + transformers.AdamW step (trainers/train.py:185-190, 353-363) on the fp32 masters, then re-pack.
        Returns a 2-element device tensor (gradient norm, factor applied to the gradients)."""
        norm = torch.empty(2, device=self.device)
        _lib.check(self.lib.msq_adamw_step(self._h, self._p(grads), lr, betas[0], betas[1], eps, weight_decay, max_grad_norm,
                                           grad_scale, self._p(norm), self._stream()))
        return norm

    def read_param(self, name, shape):
        """Current fp32 master of a registered weight (e.g. for save_pretrained)."""
        out = torch.empty(*shape, device=self.device)
        _lib.check(self.lib.msq_train_read_param(self._h, name.encode(), self._p(out), out.numel(), self._stream()))
        return out

    def set_dropout(self, p_hidden=0.1, p_attn=0.1, p_para=0.1, seed=0):
        """Training-mode dropout for every later train_step / inner_forward_train (the reference's defaults are 0.1 each:
        BertConfig.hidden_dropout_prob / attention_probs_dropout_prob, args.para_dropout).  0 switches a family off."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.msq_train_set_dropout(self._h, float(p_hidden), float(p_attn), float(p_para), int(seed) & 0xFFFFFFFF,
                                                      self._stream()))

    def set_multimodal_loss(self, on=True):
        """args.multimodal_loss of the reference (modeling_bert.py:1218-1225): every later train_step adds the image pairwise term
        lam * NLL(pairwise_relationship(img_projection(visn[:, 0]))).  The state dict must carry img_projection.weight [H, H]
        and .bias [H]."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.msq_train_set_multimodal_loss(self._h, 1 if on else 0, self._stream()))

    def set_bn_mode(self, use_running_stats=False):
        """BatchNorm of the ModifiedResNet tower inside training steps: batch statistics (False, nn.BatchNorm2d.train(), what the
        reference fine-tunes with) or the frozen running statistics (True, eval() semantics)."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.msq_train_set_bn_mode(self._h, 1 if use_running_stats else 0, self._stream()))

    def dropout_step(self):
        """counter that keyed the masks of the last training forward (-1: none yet); what oracle.dropout.DropSpec needs"""
        return int(self.lib.msq_train_dropout_step(self._h))

    def train_step(self, batch: PairBatch, grads, lam=0.6, triplets=None, triplet_weight=0.1):
        """One fine-tuning forward + backward of BertForOrdering._forward's default objective (modeling_bert.py:943-1174:
        pointer NLL / (N-1) + lam * pairwise NLL / P, batch mean): grads += dL/dparam for every parameter of the path
        (encoder and heads).  Returns the loss as a 0-d device tensor.  Follow with adamw_step(grads, ...).
        triplets [B,3] (anchor, positive, negative sentence index per manual): adds the reference's optional time-contrastive
        term triplet_weight * TripletMarginLoss(margin 1, p 2) (modeling_bert.py:1176-1216)."""
        b = batch.to(self.device)
        B, P, Lt = b.input_ids.shape
        if triplets is not None:
            tr = torch.as_tensor(triplets).to(device=self.device, dtype=torch.int32).contiguous()
            if tuple(tr.shape) != (B, 3):
                raise ValueError("triplets must be [B, 3], got %s" % (tuple(tr.shape),))
            _lib.check(self.lib.msq_train_set_triplets(self._h, self._p(tr), B, float(triplet_weight), self._stream()))
        gt = b.ground_truth.to(torch.int32).contiguous()
        loss = torch.zeros(1, device=self.device)
        n_img = 0 if b.images is None else b.images.shape[0]
        self._train_keep = (b, gt)
        _lib.check(self.lib.msq_train_step(self._h, self._p(b.input_ids), self._p(b.token_type_ids), self._p(b.attention_mask),
                                           self._p(b.sep_positions), B, b.n_steps, Lt, self._p(b.images), n_img, self._p(b.img_index),
                                           self._p(gt), self._p(b.pairwise_labels.contiguous()), float(lam), self._p(grads),
                                           self._p(loss), self._stream()))
        return loss[0]

    def training_loss(self, batch: PairBatch, lam=0.6):
        """BertForOrdering._forward loss value (modeling_bert.py:943-1174), forward only -> 0-d device tensor."""
        b = batch.to(self.device)
        B, P, Lt = b.input_ids.shape
        gt = b.ground_truth.to(torch.int32).contiguous()
        perm = torch.empty(B, b.n_steps, dtype=torch.int32, device=self.device)
        loss = torch.zeros(1, device=self.device)
        n_img = 0 if b.images is None else b.images.shape[0]
        _lib.check(self.lib.msq_training_loss(self._h, self._p(b.input_ids), self._p(b.token_type_ids), self._p(b.attention_mask),
                                              self._p(b.sep_positions), B, b.n_steps, Lt, self._p(b.images), n_img,
                                              self._p(b.img_index), self._p(gt), self._p(b.pairwise_labels.contiguous()),
                                              float(lam), self._p(perm), self._p(loss), self._stream()))
        return loss[0]

    def beam_search(self, enc, n_steps, beam, trace=False):
        """beam_search_pointer for B manuals at once.  Returns perm [B,N] int32 (+ trace dict)."""
        f = lambda t: t.to(self.device, torch.float32).contiguous()
        sents, key, h0 = f(enc["sents"]), f(enc["key"]), f(enc["h0"]).reshape(-1, self.H)
        cls_mat, score_mat = f(enc["cls_mat"]), f(enc["score_mat"])
        B, N = sents.shape[0], n_steps
        perm = torch.empty(B, N, dtype=torch.int32, device=self.device)
        tr = None
        if trace:
            tr = dict(ix=torch.empty(B, N - 1, beam, dtype=torch.int32, device=self.device),
                      cost=torch.zeros(B, N - 1, beam, device=self.device),
                      logp=torch.zeros(B, N - 1, beam, N, device=self.device))
        _lib.check(self.lib.msq_beam_search(self._h, self._p(sents), self._p(key), self._p(h0), self._p(cls_mat),
                                            self._p(score_mat), B, N, beam, self._p(perm),
                                            self._p(tr["ix"] if tr else None), self._p(tr["cost"] if tr else None),
                                            self._p(tr["logp"] if tr else None), self._stream()))
        return (perm, tr) if trace else perm

    def decode_step(self, prev_y, prev_handc, original_keys, mask, rela_vec, rela_mask, hist_left1, hist_left2,
                    l1_mask, l2_mask):
        """BertForOrdering.step (modeling_bert.py:1368-1402) with the reference's materialised tensors.
        rela_vec is zeroed IN PLACE where rela_mask == 0, exactly as the reference does."""
        dev, H = self.device, self.H
        f = lambda t: t.to(dev, torch.float32).contiguous()
        u8 = lambda t: (t != 0).to(dev, torch.uint8).contiguous()
        Wb, N = mask.shape
        x = f(prev_y).reshape(Wb, H)
        h, c = f(prev_handc[0]).reshape(Wb, H), f(prev_handc[1]).reshape(Wb, H)
        key0 = f(original_keys).reshape(N, H)
        if not (rela_vec.is_cuda and rela_vec.dtype == torch.float32 and rela_vec.is_contiguous()):
            raise ValueError("rela_vec must be a contiguous fp32 CUDA tensor (it is updated in place)")
        h1, h2 = f(hist_left1), f(hist_left2)
        pm, rm, m1, m2 = u8(mask), u8(rela_mask), u8(l1_mask), u8(l2_mask)
        ho, co = torch.empty(1, Wb, H, device=dev), torch.empty(1, Wb, H, device=dev)
        logp = torch.empty(Wb, N, device=dev)
        _lib.check(self.lib.msq_decode_step(self._h, self._p(x), self._p(h), self._p(c), self._p(key0), self._p(pm),
                                            self._p(rela_vec), self._p(rm), self._p(h1), self._p(h2), self._p(m1),
                                            self._p(m2), Wb, N, self._p(ho), self._p(co), self._p(logp), self._stream()))
        return ho, co, logp

    def order_device(self, batch: PairBatch, beam):
        """encode + beam search on device-resident inputs; returns perm [B,N] int32 (device, async)."""
        b = batch.to(self.device)
        B, P, Lt = b.input_ids.shape
        perm = torch.empty(B, b.n_steps, dtype=torch.int32, device=self.device)
        n_img = 0 if b.images is None else b.images.shape[0]
        _lib.check(self.lib.msq_order_manuals_dev(self._h, self._p(b.input_ids), self._p(b.token_type_ids),
                                                  self._p(b.attention_mask), self._p(b.sep_positions), B, b.n_steps, Lt,
                                                  self._p(b.images), n_img, self._p(b.img_index), beam, self._p(perm),
                                                  self._stream()))
        return perm

    def order_host(self, batch: PairBatch, beam, perm_out=None):
        """Whole path from HOST (ideally pinned) buffers, H2D and D2H included; synchronous."""
        b = batch.to(None)       # canonical dtypes / contiguity, stays in host memory
        if b.input_ids.is_cuda:
            raise ValueError("order_host takes HOST buffers (use order_device for a device-resident batch)")
        B, P, Lt = b.input_ids.shape
        if perm_out is None:
            perm_out = torch.empty(B, b.n_steps, dtype=torch.int32).pin_memory()
        n_img = 0 if b.images is None else b.images.shape[0]
        with torch.cuda.device(self.device):
            _lib.check(self.lib.msq_order_manuals_host(self._h, self._p(b.input_ids), self._p(b.token_type_ids),
                                                       self._p(b.attention_mask), self._p(b.sep_positions), B, b.n_steps,
                                                       Lt, self._p(b.images), n_img, self._p(b.img_index), beam,
                                                       self._p(perm_out), self._stream()))
        return perm_out

    def expand_pairs_device(self, input_ids, n_steps, cls_id=101, sep_id=102, pad_id=0, with_images=True):
        """prepare_berson_inputs' pair expansion on the DEVICE (msq_scan_steps + msq_expand_pairs): token rows [B,L] ->
        (input_ids, attention_mask, token_type_ids [B,P,Lt], sep_positions [B,P,2], img_index [B,P,2] int32), all on device.
        Bit-identical to prepare_pairs (process_inputs_for_berson.py:113-368)."""
        ids = torch.as_tensor(input_ids).to(self.device, torch.long).contiguous()
        B, L = ids.shape
        N, P = n_steps, n_steps * (n_steps - 1)
        starts = torch.empty(B, N, dtype=torch.int32, device=self.device)
        lens = torch.empty(B, N, dtype=torch.int32, device=self.device)
        meta = torch.empty(2, dtype=torch.int32, device=self.device)
        lt = C.c_int32(0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.msq_scan_steps(self._p(ids), B, L, N, cls_id, sep_id, self._p(starts), self._p(lens), self._p(meta),
                                               C.byref(lt), self._stream()))
            Lt = int(lt.value)
            out = [torch.empty(B, P, Lt, dtype=torch.long, device=self.device) for _ in range(3)]
            sep = torch.empty(B, P, 2, dtype=torch.long, device=self.device)
            idx = torch.empty(B, P, 2, dtype=torch.int32, device=self.device) if with_images else None
            _lib.check(self.lib.msq_expand_pairs(self._p(ids), B, L, N, Lt, cls_id, pad_id, self._p(starts), self._p(lens),
                                                 self._p(out[0]), self._p(out[1]), self._p(out[2]), self._p(sep), self._p(idx),
                                                 self._stream()))
        return out[0], out[1], out[2], sep, idx

    def order_raw_host(self, input_ids, images, n_steps, beam, perm_out=None, cls_id=101, sep_id=102, pad_id=0):
        """berson_pointer_network from the DataLoader tuple: token rows [B,L] int64 and step images [B,N,3,S,S] fp32 in HOST
        memory (pinned for full H2D bandwidth) -> permutations [B,N] int32 (host).  Pair expansion, H2D and D2H all happen
        inside this call (msq_order_manuals_raw_host); synchronous."""
        ids = torch.as_tensor(input_ids)
        if ids.dtype != torch.long or not ids.is_contiguous() or ids.is_cuda:
            ids = ids.cpu().to(torch.long).contiguous()
        B, L = ids.shape
        if images is not None:
            if images.dtype != torch.float32 or not images.is_contiguous() or images.is_cuda:
                images = images.cpu().float().contiguous()
            if images.shape[0] * (images.shape[1] if images.dim() == 5 else 1) != B * n_steps:
                raise ValueError("images must hold one image per step: [B,N,3,S,S] or [B*N,3,S,S]")
        if perm_out is None:
            perm_out = torch.empty(B, n_steps, dtype=torch.int32).pin_memory()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.msq_order_manuals_raw_host(self._h, self._p(ids), B, L, n_steps, cls_id, sep_id, pad_id,
                                                           self._p(images), beam, self._p(perm_out), self._stream()))
        return perm_out

    def order(self, input_ids, labels, n_steps, beam, images=None, cls_id=101, sep_id=102, pad_id=0):
        """berson_pointer_network for a batch of manuals: list of permutations (python ints)."""
        batch = prepare_pairs(input_ids, labels, n_steps, images, cls_id, sep_id, pad_id).to(self.device)
        return self.order_device(batch, beam).cpu().tolist()
