"""ORACLE side of the fine-tuning path's counter-based dropout (test infrastructure only): the same integer hash as
multimodal_sequencing_b200/csrc/dropout.cuh in numpy, so that the CUDA path and the torch-autograd oracle drop EXACTLY the
same elements.  The reference itself uses torch's stateful nn.Dropout (models/berson/modeling_bert.py:160-319, 677-735,
models/CLIP/src/lxrt/modeling.py:354-601, models/berson/encoder.py:17-28, neural.py:27-32, 93, 228): which elements are
dropped is not part of its contract, the sites, probabilities and the 1/(1-p) scaling are -- those are what this restates."""
import numpy as np
import torch

KIND = dict(E=1, V=2, A=3, O=4, F=5, H=6, PA=7, PC=8, PF1=9, PF2=10)
_M = np.uint64(0xFFFFFFFF)


def _fmix(h):
    h = h.astype(np.uint64) & _M
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & _M
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & _M
    h ^= h >> np.uint64(16)
    return h


def keep_mask(seed, step, kind, layer, p, shape):
    """bool array of `shape`: element with flat index i is kept iff hash(seed, step, site, i) >= p * 2^32."""
    site = np.uint64(KIND[kind] + 16 * layer)
    key = _fmix(np.uint64(seed) ^ _fmix(np.array((step + 0x9E3779B9) & 0xFFFFFFFF, dtype=np.uint64)) ^ ((site * np.uint64(0x85EBCA6B)) & _M))
    idx = np.arange(int(np.prod(shape)), dtype=np.uint64)
    h = _fmix(key ^ (idx & _M) ^ (((idx >> np.uint64(32)) * np.uint64(0x9E3779B1)) & _M))
    thresh = np.uint64(int(np.float64(np.float32(p)) * 4294967296.0))   # (uint32)((double)(float)p * 2^32), as the device computes it
    return (h >= thresh).reshape(shape)


class DropSpec:
    """what msq_train_set_dropout configures + the step counter of the forward being checked"""

    def __init__(self, seed, step, p_hidden=0.0, p_attn=0.0, p_para=0.0):
        self.seed, self.step, self.p = seed, step, dict(hidden=p_hidden, attn=p_attn, para=p_para)

    def __call__(self, x, kind, layer=0):
        p = self.p[{"A": "attn", "PA": "para", "PC": "para", "PF1": "para", "PF2": "para"}.get(kind, "hidden")]
        if not p > 0:
            return x
        m = torch.from_numpy(keep_mask(self.seed, self.step, kind, layer, p, tuple(x.shape)))
        scale = float(np.float32(1.0) / (np.float32(1.0) - np.float32(p)))   # 1.0f / (1.0f - p), as the device computes it
        return x * m.to(x.dtype) * scale
