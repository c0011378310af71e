"""One shape of the split-operand attention, a few launches (for ncu captures): python scripts/attn_only.py [L] [R]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_sequencing_b200 import _lib  # noqa: E402

lib = _lib.load()
L = int(sys.argv[1]) if len(sys.argv) > 1 else 227
R = int(sys.argv[2]) if len(sys.argv) > 2 else 640
heads = 12
qkv = torch.randn(R * L, 2 * 3 * heads * 64, device="cuda").bfloat16()
qkv[:, 3 * heads * 64:] *= 2.0 ** -9
mask = torch.zeros(R, 128, device="cuda")
ctx = torch.empty(R * L, 2 * heads * 64, device="cuda", dtype=torch.bfloat16)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(4):
    _lib.check(lib.msq_attention(2, qkv.data_ptr(), R, L, heads, 0.125, mask.data_ptr() if L > 128 else None, 128 if L > 128 else 0,
                                 ctx.data_ptr(), st))
torch.cuda.synchronize()
print("ok")
