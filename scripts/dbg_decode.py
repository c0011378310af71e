import os, sys, torch
sys.path.insert(0, ".")
from multimodal_sequencing_b200 import OrderingEngine
from oracle import synth
torch.set_grad_enabled(False)
H = 768
cfg = dict(hidden_size=H, num_hidden_layers=1, num_attention_heads=12, intermediate_size=64, vocab_size=64,
           max_position_embeddings=8, vit=None, para_ff=64)
eng = OrderingEngine(synth.full_state_dict(cfg, None, seed=0, ff=64), cfg, precise=(os.environ.get("MSQ_BENCH_PRECISE", "0") == "1"))
N, W, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
enc = {k: v.cuda() for k, v in synth.synthetic_encode(N, H, seed=3, B=B).items()}
for i in range(3):
    p = eng.beam_search(enc, N, W)
    torch.cuda.synchronize()
print(p[0].tolist())
