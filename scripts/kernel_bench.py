"""Per-kernel roofline lines (CUDA events, warm, inputs > L2 or L2 flushed between iterations):
  * LayerNorm (HBM bound): algorithmic bytes = rows*H*(4 in + 4 out_f32 + 2 out_bf16)
  * tcgen05 GEMM shapes of the encoder (tensor bound)
  * tcgen05 attention (seq 227 / 99)
  * pointer decoder + beam search sweep (BASELINE configs[4]: N=10, W=16, B swept) against the
    base-tensor algorithmic bytes of SURVEY.md §8(d)
Writes one JSON object per line to stdout."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_sequencing_b200 import OrderingEngine, _lib  # noqa: E402
from oracle import synth  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else \
    dict(hbm_gbs=6650.0, bf16_tflops=1590.0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10, warm=3, do_flush=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if do_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def bench_ln():
    rows, H = 145280, 768
    x = torch.randn(rows, H, device=dev)
    g, b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    of, ob = torch.empty_like(x), torch.empty(rows, H, device=dev, dtype=torch.bfloat16)
    ms = timed(lambda: lib.msq_layernorm(1, x.data_ptr(), rows, H, g.data_ptr(), b.data_ptr(), 1e-12, ob.data_ptr(), st()))
    byt = rows * H * (4 + 2)
    print(json.dumps(dict(kernel="layernorm_kernel<bf16> (fp32 in -> bf16 out)", rows=rows, H=H, ms=ms, gbs=byt / ms / 1e6,
                          frac_hbm=byt / ms / 1e6 / PK["hbm_gbs"], bound="hbm")))


def bench_gemm():
    for name, M, N, K, act, obf in (("bert_qkv", 145280, 2304, 768, 0, True), ("bert_ffn_up_gelu", 145280, 3072, 768, 1, True),
                                    ("bert_ffn_down_resid", 145280, 768, 3072, 0, False), ("bert_out_resid", 145280, 768, 768, 0, False),
                                    ("vit_fc_quickgelu", 63360, 3072, 768, 2, True)):
        A = torch.randn(M, K, device=dev).bfloat16()
        W = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
        bias = torch.zeros(N, device=dev)
        res = None if obf else torch.randn(M, N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if obf else torch.float32)
        ms = timed(lambda: lib.msq_gemm(2 if obf else 1, A.data_ptr(), W.data_ptr(), bias.data_ptr(), res.data_ptr() if res is not None else None,
                                        out.data_ptr(), M, N, K, act, st()), iters=8)
        tf = 2.0 * M * N * K / ms / 1e9
        print(json.dumps(dict(kernel="gemm_tc_kernel " + name, M=M, N=N, K=K, ms=ms, tflops=tf, frac_burst=tf / PK["bf16_tflops"],
                              frac_sustained=tf / PK.get("bf16_tflops_sustained", PK["bf16_tflops"]), bound="tensor")))
        del A, W, out, res


def bench_deferred():
    """The deferred-LayerNorm epilogues next to the plain GEMM + separate LayerNorm they replace (BERT layer shapes)."""
    M, H, I = 145280, 768, 3072
    sp = 6
    gam, bet = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    stats = torch.zeros(M, sp, 2, device=dev)
    stats[:, :, 1] = 128.0
    stats_o = torch.empty(M, sp, 2, device=dev)
    y = torch.randn(M, H, device=dev)
    yt = y.bfloat16()
    for name, N, K, act in (("qkv", 3 * H, H, 0), ("ffn_up_gelu", I, H, 1)):
        W = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
        bias, sv = torch.zeros(N, device=dev), torch.zeros(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ms0 = timed(lambda: lib.msq_gemm(2, yt.data_ptr(), W.data_ptr(), bias.data_ptr(), None, out.data_ptr(), M, N, K, act, st()), iters=8)
        ms1 = timed(lambda: lib.msq_gemm_deferred_ln(1, 1, yt.data_ptr(), W.data_ptr(), bias.data_ptr(), None, sv.data_ptr(), None,
                                                     stats.data_ptr(), sp, H, 1e-12, out.data_ptr(), None, None, M, N, K, act, st()), iters=8)
        print(json.dumps(dict(kernel="gemm_tc " + name, plain_ms=ms0, folded_ms=ms1, tflops_folded=2.0 * M * N * K / ms1 / 1e9)))
        del W, out
    ln_ms = timed(lambda: lib.msq_layernorm(1, y.data_ptr(), M, H, gam.data_ptr(), bet.data_ptr(), 1e-12, yt.data_ptr(), st()))
    for name, K in (("out_proj", H), ("ffn_down", I)):
        A = torch.randn(M, K, device=dev).bfloat16()
        W = (torch.randn(H, K, device=dev) * 0.02).bfloat16()
        bias = torch.zeros(H, device=dev)
        tmp = torch.empty(M, H, device=dev)
        ms0 = timed(lambda: lib.msq_gemm(1, A.data_ptr(), W.data_ptr(), bias.data_ptr(), y.data_ptr(), tmp.data_ptr(), M, H, K, 0, st()), iters=8)
        ms1 = timed(lambda: lib.msq_gemm_deferred_ln(2, 0, A.data_ptr(), W.data_ptr(), bias.data_ptr(), y.data_ptr(), gam.data_ptr(),
                                                     bet.data_ptr(), stats.data_ptr(), sp, H, 1e-12, y.data_ptr(), yt.data_ptr(),
                                                     stats_o.data_ptr(), M, H, K, 0, st()), iters=8)
        byt = M * (K * 2 + H * (4 + 4 + 2))
        print(json.dumps(dict(kernel="gemm_tc " + name, plain_ms=ms0, layernorm_ms=ln_ms, plain_plus_ln_ms=ms0 + ln_ms, deferred_ms=ms1,
                              deferred_gbs=byt / ms1 / 1e6, deferred_tflops=2.0 * M * H * K / ms1 / 1e9)))
        del A, W, tmp


def bench_attn():
    for split in (False, True):
        for L, masked in ((227, True), (99, False)):
            R, heads = 640, 12
            pl = 2 if split else 1
            qkv = torch.randn(R * L, pl * 3 * heads * 64, device=dev).bfloat16()
            if split:
                qkv[:, 3 * heads * 64:] *= 2.0 ** -9
            mask = torch.zeros(R, 128, device=dev)
            ctx = torch.empty(R * L, pl * heads * 64, device=dev, dtype=torch.bfloat16)
            ms = timed(lambda: lib.msq_attention(2 if split else 1, qkv.data_ptr(), R, L, heads, 0.125, mask.data_ptr() if masked else None,
                                                 128 if masked else 0, ctx.data_ptr(), st()))
            fl = 4.0 * L * L * 64 * heads * R
            byt = R * L * heads * 64 * 2 * 4 * pl
            print(json.dumps(dict(kernel="attention (bf16x3)" if split else "attention (bf16)", L=L, R=R, ms=ms, tflops=fl / ms / 1e9,
                                  executed_tflops=fl * (3 if split else 1) / ms / 1e9, gbs=byt / ms / 1e6,
                                  frac_hbm=byt / ms / 1e6 / PK["hbm_gbs"], bound="tensor+sfu")))


def bench_decode():
    H = 768
    cfg = dict(hidden_size=H, num_hidden_layers=1, num_attention_heads=12, intermediate_size=64, vocab_size=64,
               max_position_embeddings=8, vit=None, para_ff=64)
    eng = OrderingEngine(synth.full_state_dict(cfg, None, seed=0, ff=64), cfg, precise=(os.environ.get("MSQ_BENCH_PRECISE", "0") == "1"))
    for N, W in ((5, 4), (10, 16)):
        for B in (1, 8, 64, 256):
            enc = {k: v.to(dev) for k, v in synth.synthetic_encode(N, H, seed=3, B=B).items()}
            ms = timed(lambda: eng.beam_search(enc, N, W), iters=5, do_flush=False)
            per_step = N * N * 770 * 4 + N * 768 * 4 + W * (4 * 768 * 4 + 2 * N * 4)
            byt = B * per_step * (N - 1)
            live, fma = 1, 0
            for t in range(N - 1):       # recurrent GEMMs: gates (4H columns) + query (H columns) for every live row
                fma += B * live * 5 * H * H
                live = min(W, live * N)
            peak32 = 148 * 128 * 2 * 1.965e9 / 1e12   # fp32 FFMA peak at the maximum SM clock (TFLOP/s)
            print(json.dumps(dict(kernel="decode (2 pre-projection GEMMs + %s)" % ("beam_search_kernel" if os.environ.get("MSQ_DECODE_FUSED") == "1"
                                                                                  else "per step: dec_gemm x2 + dec_select"),
                                  N=N, W=W, B=B, ms=ms, manuals_per_s=B / ms * 1e3, algorithmic_gbs=byt / ms / 1e6,
                                  frac_hbm=byt / ms / 1e6 / PK["hbm_gbs"], recurrent_fp32_tflops=2 * fma / ms / 1e9,
                                  frac_fp32_peak=2 * fma / ms / 1e9 / peak32, bound="fp32 FMA (recurrent GEMMs) + HBM (T4 rows)")))


if __name__ == "__main__":
    which = sys.argv[1:] or ["ln", "gemm", "deferred", "attn", "decode"]
    for w in which:
        {"ln": bench_ln, "gemm": bench_gemm, "deferred": bench_deferred, "attn": bench_attn, "decode": bench_decode}[w]()
