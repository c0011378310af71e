"""models/CLIP/clip/model.py of the reference, visual-tower surface: CLIP(...).visual(x, skip_last_layer=True)
(model.py:242-323).  Parameter names match the reference so CLIP checkpoints load unchanged; the forward is
the pair-joint ViT tower of csrc (patch-embed GEMM, fused token assembly + ln_pre, 12 residual blocks on
tcgen05 GEMMs / fused attention, ln_post)."""
from collections import OrderedDict

import torch
import torch.nn as nn

from multimodal_sequencing_b200.engine import OrderingEngine
from multimodal_sequencing_b200.dropin._owner import owned_engine


class LayerNorm(nn.LayerNorm):
    """model.py:190-196 (fp32 LayerNorm, eps 1e-5): parameter holder."""


class QuickGELU(nn.Module):
    """model.py:199-201; evaluated in the GEMM epilogue on the device path."""

    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    """model.py:204-226 (parameter holder: attn.in_proj_*, attn.out_proj, ln_1, mlp.c_fc, mlp.c_proj, ln_2)."""

    def __init__(self, d_model, n_head, attn_mask=None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)


class Transformer(nn.Module):
    def __init__(self, width, layers, heads, attn_mask=None):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])


class VisualTransformer(nn.Module):
    """model.py:242-305.  forward(x [R*img_len,3,S,S]) -> [R, 1 + img_len*g*g, width] (skip_last_layer=True)."""

    def __init__(self, input_resolution, patch_size, width, layers, heads, output_dim, img_len=None):
        super().__init__()
        self.input_resolution, self.output_dim, self.heads, self.img_len = input_resolution, output_dim, heads, img_len
        self.patch_size, self.width, self.layers = patch_size, width, layers
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def vit_config(self):
        return dict(embed_dim=self.output_dim, image_resolution=self.input_resolution, vision_layers=self.layers,
                    vision_width=self.width, vision_patch_size=self.patch_size)

    def _engine(self):
        return owned_engine(self, lambda: dict(hidden_size=self.width, num_hidden_layers=0, num_attention_heads=self.width // 64,
                                               intermediate_size=4 * self.width, vocab_size=1, max_position_embeddings=1,
                                               vit=self.vit_config()),
                            lambda: {"bert.encoder.visual_model.visual." + k: v for k, v in self.state_dict().items()},
                            self.conv1.weight.device)

    def forward(self, x, skip_last_layer=False, text_embedding=None, text_mask=None, img_len=None):
        if not skip_last_layer or text_embedding is not None:
            # the `x @ proj` tail feeds a visn_fc of the wrong width in the reference and the
            # vilt-style joint branch is off by default (param.py:243-279): neither is on the scoped path
            raise NotImplementedError("only the skip_last_layer=True tower (ln_post tokens) is on the ordering path")
        il = self.img_len or img_len or 2
        if il != 2:
            raise NotImplementedError("BERSON feeds image PAIRS (max_subsample_image_length = 2)")
        R = x.shape[0] // il
        idx = torch.arange(R * il, dtype=torch.int32, device=x.device)
        with torch.no_grad():
            return self._engine().vit_forward(x, idx, R)


class Bottleneck(nn.Module):
    """model.py:10-53 (parameter holder).  conv1/bn1 (1x1) -> conv2/bn2 (3x3) -> AvgPool2d(stride) -> conv3/bn3 (1x1),
    shortcut `downsample` = AvgPool2d + 1x1 conv + BN when the shape changes; all of it runs as im2col + GEMM with the
    BatchNorm folded in (csrc/rn.cu, api.cu run_rn_trunk)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1, self.bn1 = nn.Conv2d(inplanes, planes, 1, bias=False), nn.BatchNorm2d(planes)
        self.conv2, self.bn2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False), nn.BatchNorm2d(planes)
        self.conv3, self.bn3 = nn.Conv2d(planes, planes * 4, 1, bias=False), nn.BatchNorm2d(planes * 4)
        self.stride, self.downsample = stride, None
        if stride > 1 or inplanes != planes * 4:
            self.downsample = nn.Sequential(OrderedDict([("-1", nn.AvgPool2d(stride)),
                                                         ("0", nn.Conv2d(inplanes, planes * 4, 1, stride=1, bias=False)),
                                                         ("1", nn.BatchNorm2d(planes * 4))]))


class AttentionPool2d(nn.Module):
    """model.py:55-125 (parameter holder): positional_embedding, q/k/v/c_proj, and the token_type_embedding table the
    reference allocates for img_len > 1 but only reads behind a disabled flag."""

    def __init__(self, spacial_dim, embed_dim, num_heads, output_dim=None, img_len=None):
        super().__init__()
        self.img_len, self.num_heads, self.embed_dim = img_len, num_heads, embed_dim
        self.positional_embedding = nn.Parameter(torch.randn(spacial_dim ** 2 + 1, embed_dim) / embed_dim ** 0.5)
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.c_proj = nn.Linear(embed_dim, output_dim or embed_dim)
        if img_len is not None and img_len > 1:
            self.token_type_embedding = nn.Embedding(5, embed_dim)


class ModifiedResNet(nn.Module):
    """model.py:128-187.  forward(x [R*img_len,3,S,S]) -> [R, 1 + img_len*(S/32)^2, 2*output_dim]: 3-conv stem, AvgPool,
    four bottleneck stages, AttentionPool2d over the PAIR's tokens, concatenated with itself (model.py:106)."""

    def __init__(self, layers, output_dim, heads, input_resolution=224, width=64, img_len=None):
        super().__init__()
        self.output_dim, self.input_resolution, self.width, self.layers, self.img_len = output_dim, input_resolution, width, tuple(layers), img_len
        self.conv1, self.bn1 = nn.Conv2d(3, width // 2, kernel_size=3, stride=2, padding=1, bias=False), nn.BatchNorm2d(width // 2)
        self.conv2, self.bn2 = nn.Conv2d(width // 2, width // 2, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(width // 2)
        self.conv3, self.bn3 = nn.Conv2d(width // 2, width, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(width)
        self._inplanes = width
        self.layer1 = self._make_layer(width, layers[0])
        self.layer2 = self._make_layer(width * 2, layers[1], stride=2)
        self.layer3 = self._make_layer(width * 4, layers[2], stride=2)
        self.layer4 = self._make_layer(width * 8, layers[3], stride=2)
        self.attnpool = AttentionPool2d(input_resolution // 32, width * 32, heads, output_dim, img_len=img_len)

    def _make_layer(self, planes, blocks, stride=1):
        mods = [Bottleneck(self._inplanes, planes, stride)]
        self._inplanes = planes * Bottleneck.expansion
        mods += [Bottleneck(self._inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*mods)

    def rn_config(self):
        return dict(embed_dim=self.output_dim, image_resolution=self.input_resolution, vision_layers=self.layers,
                    vision_width=self.width)

    def _engine(self):
        return owned_engine(self, lambda: dict(hidden_size=128, num_hidden_layers=0, num_attention_heads=2, intermediate_size=512,
                                               vocab_size=1, max_position_embeddings=1, rn=self.rn_config()),
                            lambda: {"bert.encoder.visual_model.visual." + k: v for k, v in self.state_dict().items()},
                            self.conv1.weight.device)

    def forward(self, x, skip_last_layer=False, text_embedding=None, text_mask=None, img_len=None):
        if skip_last_layer or text_embedding is not None:
            # LXRTModel hard-wires skip_last_layer=False for this tower (lxrt/modeling.py:783)
            raise NotImplementedError("the ordering path runs the ResNet tower with its attention pool (skip_last_layer=False)")
        if self.training:
            raise NotImplementedError("BatchNorm is folded for inference; training the tower is outside the drop-in's path")
        il = self.img_len or img_len or 2
        if il != 2:
            raise NotImplementedError("BERSON feeds image PAIRS (max_subsample_image_length = 2)")
        R = x.shape[0] // il
        idx = torch.arange(R * il, dtype=torch.int32, device=x.device)
        with torch.no_grad():
            return self._engine().vit_forward(x, idx, R)


class CLIP(nn.Module):
    """model.py:307-367.  Only the visual tower is built: the text tower is dead weight on this path
    (63.4 M parameters the reference allocates and never runs)."""

    def __init__(self, embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size, context_length,
                 vocab_size, transformer_width, transformer_heads, transformer_layers, img_len=None, img_only=False):
        super().__init__()
        self.context_length, self.img_only = context_length, img_only
        if isinstance(vision_layers, (tuple, list)):
            self.visual = ModifiedResNet(vision_layers, embed_dim, vision_width * 32 // 64, image_resolution, vision_width,
                                         img_len=img_len)
        else:
            self.visual = VisualTransformer(image_resolution, vision_patch_size, vision_width, vision_layers,
                                            vision_width // 64, embed_dim, img_len=img_len)

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype
