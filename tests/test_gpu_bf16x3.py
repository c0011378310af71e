"""GPU parity tests of the bf16x3 precision mode (msq_config.precise == 2): every tensor-core operand is carried
as hi + lo bf16 (16 significand bits) and every product is three tcgen05 MMAs (hi*hi + lo*hi + hi*lo, fp32
accumulation in TMEM).  north_star's gate for the tensor-core path: encoder outputs within 1e-3 relative, predicted
permutations and beam indices bit-exact.  Everything goes through the C ABI; the oracle is only the checker."""
import ctypes as C
import os

import pytest
import torch

from oracle import berson_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)

ENC = ["sents", "para", "h0", "key", "cls", "cls_mat", "cls_score", "score_mat", "his1", "his2"]
REL_X3 = 1e-3      # north_star: "within 1e-3 relative in bf16" -- the asserted bound; measured values are ~100x smaller
GAP_EPS = 2e-4     # a permutation may differ from the fp32 path only where the fp32 path's own decision margin is below this


def _lib_and_stream():
    from multimodal_sequencing_b200 import _lib
    lib = _lib.load()
    assert lib.msq_tc_available() == 1, "tcgen05 path unavailable on this device"
    return _lib, lib, C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _split(lib, _lib, x, st):
    """fp32 [rows, K] (device) -> split rows [hi(K) | lo(K)] as a bf16 tensor [rows, 2K]."""
    rows, K = x.shape
    out = torch.empty(rows, 2 * K, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.msq_f32_to_bf16_split(x.data_ptr(), out.data_ptr(), rows, K, st))
    return out


def _join(t, K):
    return t[:, :K].float() + t[:, K:].float()


def _rel_l2(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_split_roundtrip_is_16_bit():
    _lib, lib, st = _lib_and_stream()
    x = torch.randn(513, 768, device="cuda") * 3
    s = _split(lib, _lib, x, st)
    torch.cuda.synchronize()
    assert torch.equal(s[:, :768], x.bfloat16())
    err = (_join(s, 768) - x).abs().max().item()
    assert err <= 2.0 ** -16 * x.abs().max().item(), err


@pytest.mark.parametrize("M,N,K,act,split_out", [(300, 768, 768, 0, False), (4540, 2304, 768, 0, True), (1980, 3072, 768, 2, True),
                                                 (777, 768, 3072, 1, False), (1000, 3072, 768, 1, True), (50, 128, 128, 3, False),
                                                 (37000, 768, 768, 0, False), (129, 64, 64, 5, True)])
def test_gemm_bf16x3(M, N, K, act, split_out):
    """split-bf16 tcgen05 GEMM against float64 on the ORIGINAL fp32 operands (not on rounded copies): the error left is
    the 2^-17 operand representation + fp32 accumulation."""
    _lib, lib, st = _lib_and_stream()
    g = torch.Generator().manual_seed(M + N + K)
    A, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.05
    bias = torch.randn(N, generator=g)
    resid = None if split_out else torch.randn(M, N, generator=g)
    ref = A.double() @ W.double().t() + bias.double()
    ref = {0: lambda x: x, 1: O.gelu_erf, 2: O.quick_gelu, 3: torch.tanh, 5: torch.relu}[act](ref)
    if resid is not None:
        ref = ref + resid.double()
    As, Ws = _split(lib, _lib, A.cuda(), st), _split(lib, _lib, W.cuda(), st)
    bd = bias.cuda()
    rd = resid.cuda() if resid is not None else None
    out = torch.empty(M, 2 * N, device="cuda", dtype=torch.bfloat16) if split_out else torch.empty(M, N, device="cuda")
    _lib.check(lib.msq_gemm(7 if split_out else 6, As.data_ptr(), Ws.data_ptr(), bd.data_ptr(), rd.data_ptr() if rd is not None else None,
                            out.data_ptr(), M, N, K, act, st))
    torch.cuda.synchronize()
    got = _join(out, N) if split_out else out
    err = (got.double().cpu() - ref).abs().max().item()
    scale = max(1.0, ref.abs().max().item())
    print("bf16x3 gemm M=%d N=%d K=%d act=%d split_out=%s: max err %.2e (max |ref| %.2f), rel-L2 %.2e" %
          (M, N, K, act, split_out, err, ref.abs().max().item(), _rel_l2(got, ref)))
    assert err <= 4e-5 * scale, err
    assert _rel_l2(got, ref) < 2e-5


@pytest.mark.parametrize("R,L,heads,mask_len", [(3, 227, 12, 128), (5, 99, 12, 0), (2, 128, 2, 128), (4, 60, 4, 60), (1, 256, 12, 100),
                                                (300, 227, 12, 128), (1, 129, 1, 0), (2, 1, 1, 0), (200, 99, 12, 0), (37, 200, 3, 128),
                                                (149, 130, 1, 64)])
def test_attention_bf16x3(R, L, heads, mask_len):
    _lib, lib, st = _lib_and_stream()
    g = torch.Generator().manual_seed(R * 1000 + L)
    H = heads * 64
    qkv = torch.randn(R * L, 3 * H, generator=g) * 1.5
    mask = None
    if mask_len:
        keep = torch.ones(R, mask_len)
        for r in range(R):
            keep[r, mask_len - 1 - (r * 7) % (mask_len // 2):] = 0
        keep[0] = 1                                             # one row with nothing masked (fast path)
        mask = (1.0 - keep) * -10000.0
    q, k, v = [t.reshape(R, L, heads, 64).permute(0, 2, 1, 3).double() for t in qkv.split(H, dim=1)]
    s = q @ k.transpose(-1, -2) * 0.125
    if mask is not None:
        full = torch.zeros(R, L, dtype=torch.double)
        full[:, :mask_len] = mask.double()
        s = s + full[:, None, None, :]
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(R * L, H)
    qs = _split(lib, _lib, qkv.cuda(), st)
    ctx = torch.empty(R * L, 2 * H, device="cuda", dtype=torch.bfloat16)
    md = mask.cuda().contiguous() if mask is not None else None
    _lib.check(lib.msq_attention(2, qs.data_ptr(), R, L, heads, 0.125, md.data_ptr() if md is not None else None, mask_len,
                                 ctx.data_ptr(), st))
    torch.cuda.synchronize()
    got = _join(ctx, H)
    err = (got.double().cpu() - ref).abs().max().item()
    print("bf16x3 attention R=%d L=%d heads=%d: max err %.2e rel-L2 %.2e" % (R, L, heads, err, _rel_l2(got, ref)))
    assert err <= 5e-5 * max(1.0, ref.abs().max().item()), err
    assert _rel_l2(got, ref) < 3e-5


def test_layernorm_split_output():
    _lib, lib, st = _lib_and_stream()
    x = torch.randn(1001, 768, device="cuda") * 2 + 0.3
    gm, bt = torch.randn(768, device="cuda"), torch.randn(768, device="cuda")
    out = torch.empty(1001, 2 * 768, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.msq_layernorm(2, x.data_ptr(), 1001, 768, gm.data_ptr(), bt.data_ptr(), 1e-12, out.data_ptr(), st))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.double(), (768,), gm.double(), bt.double(), 1e-12)
    assert (_join(out, 768).double() - ref).abs().max().item() < 5e-5 * ref.abs().max().item()


# ---------------------------------------------------------------------------------------------------------------
# whole path
# ---------------------------------------------------------------------------------------------------------------

def _engine(sd, cfg, precise):
    from multimodal_sequencing_b200 import OrderingEngine
    return OrderingEngine(sd, cfg, precise=precise)


def _cfg_from_golden(g):
    c = g["cfg"]
    return dict(hidden_size=c["hidden_size"], num_hidden_layers=c["num_hidden_layers"],
                num_attention_heads=c["num_attention_heads"], intermediate_size=c["intermediate_size"],
                vocab_size=c["vocab_size_or_config_json_file"], max_position_embeddings=c["max_position_embeddings"],
                vit=g.get("vit"), rn=g.get("rn"), para_ff=g["ff_size"])


def _assert_enc(enc, ref, what):
    """every tensor of the encode 10-tuple: max-abs within REL_X3 * max(1, max|ref|) AND relative L2 within REL_X3."""
    worst = 0.0
    for k in ENC:
        a, b = enc[k].reshape(ref[k].shape).float().cpu(), ref[k].float()
        err = (a - b).abs().max().item()
        assert err <= REL_X3 * max(1.0, b.abs().max().item()), "%s %s: max err %.3e" % (what, k, err)
        rl = _rel_l2(a, b)
        assert rl <= REL_X3, "%s %s: relative L2 %.3e" % (what, k, rl)
        worst = max(worst, rl)
    return worst


def _fixture_margin(steps, N):
    """Smallest decision gap of the REFERENCE's own beam search (fixture trace): last kept vs best dropped candidate at every
    step, best vs second-best hypothesis at the last one, and the gaps between neighbouring kept candidates (their order
    decides ties further down).  candidate = parent score + log-prob."""
    margin = float("inf")
    parent = torch.zeros(1)
    for t, s in enumerate(steps):
        cand = (parent[:, None] + s["logp"][:parent.numel()]).reshape(-1)
        cand = cand[cand > -1e8].sort(descending=True).values
        k = s["beam_ix"].numel()
        upto = min(k + 1, cand.numel())
        if upto > 1:
            margin = min(margin, float((cand[:upto - 1] - cand[1:upto]).min()))
        parent = s["score"]
    return margin


@pytest.mark.parametrize("name", ["text_tiny.pt", "mm_tiny.pt"])
def test_tiny_goldens_bf16x3(golden_dir, name):
    """fixtures the REAL reference produced: encode tensors within the bf16 gate; permutations and beam indices bit-exact
    wherever the reference's own decision gaps exceed GAP_EPS (the N=10 / W=16 fixtures hold exact ties)."""
    g = torch.load(os.path.join(golden_dir, name), weights_only=False)
    eng = _engine(g["sd"], _cfg_from_golden(g), "bf16x3")
    n_exact = 0
    for c in g["cases"]:
        if name.startswith("mm"):
            ids, labels, images = O.synthetic_manuals(1, c["N"], c["L"], vocab=1000, image_px=224, seed=c["seed"])
        else:
            ids, labels, images = c["ids"], c["labels"], None
        enc = eng.encode(eng.prepare(ids, labels, c["N"], images))
        _assert_enc(enc, c["enc"], name)
        # the decode kernel itself: the reference's encoder outputs in -> the reference's indices out, bit-exact
        perm, tr = eng.beam_search(c["enc"], c["N"], c["W"], trace=True)
        assert perm[0].tolist() == c["perm"]
        for t, s in enumerate(c["steps"]):
            k = s["beam_ix"].numel()
            assert torch.equal(tr["ix"][0, t, :k].cpu(), (s["beam_ix"] * c["N"] + s["tok_ix"]).int()), "step %d beam indices" % t
        # end to end through the bf16x3 encoder
        margin = _fixture_margin(c["steps"], c["N"])
        perm, tr = eng.beam_search(enc, c["N"], c["W"], trace=True)
        same = perm[0].tolist() == c["perm"] and all(
            torch.equal(tr["ix"][0, t, :s["beam_ix"].numel()].cpu(), (s["beam_ix"] * c["N"] + s["tok_ix"]).int())
            for t, s in enumerate(c["steps"]))
        n_exact += same
        assert same or margin < GAP_EPS, "N=%d W=%d: trace differs although the reference's smallest gap is %.3e" % (c["N"], c["W"], margin)
        assert sorted(perm[0].tolist()) == list(range(c["N"]))
        if margin >= GAP_EPS:
            assert eng.order(ids, labels, c["N"], c["W"], images) == [c["perm"]]
    print("%s: %d / %d cases bit-exact end to end (the others have a reference decision gap < %.0e)" % (name, n_exact, len(g["cases"]), GAP_EPS))
    assert n_exact >= len(g["cases"]) // 2


def _full_cfg(mm):
    cfg = dict(synth.BERT_BASE)
    cfg.update(vit=dict(synth.VIT_B32) if mm is True else None, rn=dict(synth.RN50) if mm == "rn50" else None, para_ff=3072)
    return cfg


@pytest.mark.parametrize("mm", [False, True, "rn50"])
def test_full_size_vs_oracle_bf16x3(mm):
    """BERT-base (+ ViT-B/32, or the ModifiedResNet "RN50" the reference wires by default) against the CPU oracle: 10-tuple
    within 1e-3 relative (asserted), permutations identical."""
    cfg = _full_cfg(mm)
    sd = synth.full_state_dict(cfg, cfg["vit"], seed=0, rn=cfg["rn"])
    N, W, B = 5, 4, 2
    ids, labels, images = O.synthetic_manuals(B, N, 64, image_px=224 if mm else None, seed=1)
    torch.set_num_threads(os.cpu_count() or 1)
    ocfg = dict(num_hidden_layers=12, num_attention_heads=12, vit=cfg["vit"], rn=cfg["rn"])
    oenc = O.encode(sd, ocfg, O.prepare_inputs(ids, labels, N, images))
    operm = [O.beam_search(sd, oenc, N, W, b) for b in range(B)]
    eng = _engine(sd, cfg, "bf16x3")
    enc = eng.encode(eng.prepare(ids, labels, N, images), want_top_vec=True)
    worst = _assert_enc(enc, oenc, "full-size")
    tv = _rel_l2(enc["top_vec"].reshape(oenc["top_vec"].shape), oenc["top_vec"])
    assert tv <= REL_X3
    print("full-size %s bf16x3: worst relative L2 over the 10-tuple %.2e, top_vec %.2e; max-abs %s" %
          ({False: "text", True: "mm"}.get(mm, mm), worst, tv,
           {k: "%.1e" % (enc[k].reshape(oenc[k].shape).cpu() - oenc[k]).abs().max().item() for k in ENC}))
    assert eng.order(ids, labels, N, W, images) == operm
    assert eng.order_host(eng.prepare(ids, labels, N, images), W).tolist() == operm


def _decision_margin(tr, b, N, W):
    """Smallest gap the fp32 beam search saw for manual b: at every step, between the last kept and the best dropped
    candidate; at the last step, between the best and the second-best hypothesis (candidate = parent cost + log-prob)."""
    margin = float("inf")
    cost, logp = tr["cost"][b].cpu(), tr["logp"][b].cpu()
    ix = tr["ix"][b].cpu()
    for t in range(N - 1):
        live = int((ix[t - 1] >= 0).sum()) if t else 1
        parent = cost[t - 1, :live] if t else torch.zeros(1)
        cand = (parent[:, None] + logp[t, :live]).reshape(-1)
        cand = cand[cand > -1e8].sort(descending=True).values
        k = int((ix[t] >= 0).sum())
        if t == N - 2:
            if cand.numel() > 1:
                margin = min(margin, float(cand[0] - cand[1]))
        elif cand.numel() > k:
            margin = min(margin, float(cand[k - 1] - cand[k]))
    return margin


@pytest.mark.parametrize("N,W,B,min_equal", [(5, 4, 96, 0.99), (6, 8, 32, 0.0), (10, 16, 16, 0.0)])
def test_permutation_agreement_bf16x3_vs_fp32(N, W, B, min_equal):
    """Full-size multimodal model, random-init weights (microscopic beam margins): the bf16x3 path must reproduce the
    fp32 path's permutations; a manual may differ only where the fp32 path's OWN decision margin is below GAP_EPS,
    and at N=5 / W=4 (BASELINE configs[1]) at least 99 % must be identical regardless."""
    cfg = _full_cfg(True)
    sd = synth.full_state_dict(cfg, cfg["vit"], seed=0)
    ids, labels, images = O.synthetic_manuals(B, N, 64, image_px=224, seed=100 + N)
    e32 = _engine(sd, cfg, True)
    pb = e32.prepare(ids, labels, N, images)
    enc32 = e32.encode(pb)
    p32, tr = e32.beam_search(enc32, N, W, trace=True)
    p32 = p32.cpu().tolist()
    margins = [_decision_margin(tr, b, N, W) for b in range(B)]
    enc32 = {k: v.cpu() for k, v in enc32.items()}
    del e32
    torch.cuda.empty_cache()
    ex3 = _engine(sd, cfg, "bf16x3")
    encx = ex3.encode(pb)
    worst = max(_rel_l2(encx[k], enc32[k]) for k in ENC)
    px = ex3.order(ids, labels, N, W, images)
    same = [px[b] == p32[b] for b in range(B)]
    rate = sum(same) / B
    print("bf16x3 vs fp32 path, N=%d W=%d: %d / %d identical permutations (%.1f %%), worst relative L2 of the encode tensors %.2e, "
          "median fp32 decision margin %.2e, smallest %.2e" %
          (N, W, sum(same), B, 100 * rate, worst, sorted(margins)[B // 2], min(margins)))
    for b in range(B):
        if not same[b]:
            print("  manual %d differs: fp32 %s bf16x3 %s, fp32 decision margin %.3e" % (b, p32[b], px[b], margins[b]))
            assert margins[b] < GAP_EPS, "manual %d differs although the fp32 margin is %.3e" % (b, margins[b])
    assert worst <= REL_X3
    assert rate >= min_equal
    assert O.cal_result(labels.tolist(), px)[0] == pytest.approx(O.cal_result(labels.tolist(), p32)[0], abs=0.02)
