import ctypes as C, sys, torch
sys.path.insert(0, ".")
from multimodal_sequencing_b200 import _lib
lib = _lib.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
R, L, heads = int(sys.argv[1]), int(sys.argv[2]), 12
H = heads * 64
qkv = torch.randn(R * L, 3 * H, device="cuda")
qs = torch.empty(R * L, 6 * H, device="cuda", dtype=torch.bfloat16)
_lib.check(lib.msq_f32_to_bf16_split(qkv.data_ptr(), qs.data_ptr(), R * L, 3 * H, st))
ctx = torch.empty(R * L, 2 * H, device="cuda", dtype=torch.bfloat16)
for i in range(3):
    _lib.check(lib.msq_attention(2, qs.data_ptr(), R, L, heads, 0.125, None, 0, ctx.data_ptr(), st))
    torch.cuda.synchronize()
    print("iter", i, "ok")
q, k, v = [t.reshape(R, L, heads, 64).permute(0, 2, 1, 3) for t in qkv.split(H, dim=1)]
ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).permute(0, 2, 1, 3).reshape(R * L, H)
got = ctx[:, :H].float() + ctx[:, H:].float()
print("max err", (got - ref).abs().max().item())
