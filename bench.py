#!/usr/bin/env python
"""bench.py — the BASELINE.json configs of the step-ordering hot path on N B200s, one JSON line per run.

    python bench.py [--config 1|2|3|4] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision ...]
    torchrun ... bench.py --gpus N --steps K --warmup W          (one rank per GPU, NCCL)

  --config 1 (default, the headline)  multimodal BERSON + CLIP ViT-B/32, 5-step manuals (224-px images, 64 tokens/step, 20
        ordered pairs of 227 joint tokens), beam 4, 64 manuals per step per GPU (weak scaling, manuals sharded by rank, no
        collective).  Arithmetic: bf16x3 (hi+lo bf16 tensor-core operands, three tcgen05 MMAs per product) -- the mode that
        meets north_star's gate (permutations identical to the fp32 path, encoder outputs ~2e-5 relative); the plain-bf16
        fast mode is timed in the same run and reported beside it with ITS agreement rate (`fast_mode`).
  --config 2   same model, 256 manuals in total, beam 8, strong-sharded r::G over the ranks.
  --config 3   fine-tuning step (bf16, AdamW) on 6-step manuals, 8 per GPU, NCCL all-reduce of the gradients (bucketed,
        overlapped with the rest of the backward pass).
  --config 4   long-manual decode sweep: 10-step manuals, beam 16, 256 manuals per GPU through pre-projections + beam search
        only; HBM roofline of the T4-row streaming kernel and tensor roofline of the recurrent GEMMs.
  The default run also measures configs 2-4 briefly and embeds their lines under "configs" (the driver only runs the
  default command line); --no-extra-configs skips that.

  value : whole-job throughput with the batch already resident in HBM (CUDA events, max over ranks)
  e2e   : the same metric from the DataLoader tuple in pinned HOST memory through the C ABI
          (msq_order_manuals_raw_host: H2D of token rows + images, device-side pair expansion, D2H of the orders)
  roofline : the dominant kernel, algorithmic FLOPs (bytes) per launch / CUDA-event duration of every launch in the
          timed region, against MEASURED_PEAKS.json
  cpu_baseline / --impl reference : the oracle port of the reference path on this box's host cores (all threads, ONE host
          process also when N > 1; /root/reference does not exist on the GPU box), bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOKENS, IMG = 64, 224
CFG = {1: dict(n_steps=5, beam=4, batch=64), 2: dict(n_steps=5, beam=8, total=256), 3: dict(n_steps=6, batch=8),
       4: dict(n_steps=10, beam=16, batch=256)}
FLOP_PER_PAIR = 58.37e9        # one ordered pair through the ViT pair tower + 12 joint BERT layers
METRIC = {1: "5-step manuals ordered/sec (beam=4)", 2: "5-step manuals ordered/sec (beam=8, batch 256)",
          3: "6-step manuals fine-tuned/sec", 4: "10-step manuals decoded/sec (beam=16)"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="manuals per step per GPU (configs 1, 3, 4)")
    ap.add_argument("--precision", default=None, choices=["bf16", "bf16x3", "fp32"],
                    help="encoder arithmetic (default: bf16x3 for configs 1/2, bf16 for 3/4)")
    ap.add_argument("--precise", action="store_true", help="same as --precision fp32")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="config 1 only: do not embed short runs of configs 2-4")
    ap.add_argument("--no-parity", action="store_true", help="config 1/2: skip the fp32-path agreement check")
    ap.add_argument("--lr", type=float, default=5e-6)
    ap.add_argument("--dropout", type=float, default=0.1, help="config 3: hidden / attention / paragraph-encoder dropout (the reference trains with 0.1)")
    ap.add_argument("--text", default="bert-base", choices=["bert-base", "roberta-large"])
    ap.add_argument("--backbone", default="vit", choices=["vit", "rn50"])
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.rows.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for l in self.rows:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def _towers(backbone):
    from oracle import synth
    return (dict(synth.VIT_B32), None) if backbone == "vit" else (None, dict(synth.RN50))


def config_dict(cfg_id, per_gpu, n_gpus, backbone="vit", text="bert-base"):
    """identical keys for both arms; only `manuals_per_step_per_gpu` differs (the reference orders one manual per call)."""
    c = CFG[cfg_id]
    N = c["n_steps"]
    work = {1: "configs[1]: multimodal BERSON + CLIP ViT-B/32, 5 steps x 64 tokens + 224px images, beam=4, eval",
            2: "configs[2]: same model, 256 manuals in total, beam=8, data-parallel eval strong-sharded r::G",
            3: "configs[3]: multimodal fine-tuning step (AdamW lr 5e-6 eps 1e-8 wd 0, clip 1.0, dropout 0.1 as the reference trains) on "
               "6-step manuals, NCCL gradient all-reduce",
            4: "configs[4]: long-manual stress, 10-step manuals, beam=16, pointer/beam decode (pre-projections + search) from encoder outputs"}[cfg_id]
    if backbone != "vit":
        work += " [RN50 tower instead of ViT-B/32: secondary line, not the headline]"
    d = {"workload": work, "manuals_per_step_per_gpu": per_gpu, "n_steps": N, "beam": c.get("beam"), "tokens_per_step": TOKENS,
         "pairs_per_manual": N * (N - 1), "joint_tokens_per_pair": 227, "text_encoder": text,
         "parallelism": {1: "manuals sharded by rank, no data-path collective (dp%d)", 2: "256 manuals strong-sharded r::G, no data-path collective (dp%d)",
                         3: "replicas, NCCL all-reduce of the flat fp32 gradient buffer per optimizer step (dp%d)",
                         4: "manuals sharded by rank, no collective (dp%d)"}[cfg_id] % n_gpus,
         "l2": "per-step inputs and activations (GBs) exceed the 126 MB L2; no flush needed"}
    return d


# =====================================================================================================
# reference arm / cpu_baseline: the oracle port on the host cores
# =====================================================================================================

def cpu_order_run(cfg_id, max_manuals, budget_s, backbone="vit"):
    import torch
    from oracle import berson_oracle as O
    from oracle import synth
    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c = CFG[cfg_id]
    N, W = c["n_steps"], c["beam"]
    vit, rn = _towers(backbone)
    cfg = dict(num_hidden_layers=12, num_attention_heads=12, vit=vit, rn=rn)
    sd = synth.full_state_dict(None, vit, seed=0, rn=rn)
    times, t_start = [], time.time()
    if cfg_id == 4:   # decode only: seeded encoder outputs -> beam search
        for b in range(max_manuals):
            enc = synth.synthetic_encode(N, 768, seed=3 + b)
            t0 = time.time()
            O.beam_search(sd, enc, N, W, 0)
            times.append(time.time() - t0)
            if time.time() - t_start > budget_s:
                break
        return times, cores
    ids, labels, images = O.synthetic_manuals(max_manuals, N, TOKENS, image_px=IMG, seed=1)
    for b in range(max_manuals):
        t0 = time.time()
        O.order_manuals(sd, cfg, ids[b:b + 1], labels[b:b + 1], N, W, images[b:b + 1])
        times.append(time.time() - t0)
        if time.time() - t_start > budget_s:
            break
    return times, cores


def cpu_finetune_run(n, budget_s):
    import torch
    from oracle import berson_oracle as O
    from oracle import synth
    from oracle import train_oracle as TO
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vit = dict(synth.VIT_B32)
    cfg = dict(num_hidden_layers=12, num_attention_heads=12, vit=vit)
    sd = synth.full_state_dict(None, vit, seed=0)
    N = CFG[3]["n_steps"]
    ids, labels, images = O.synthetic_manuals(1, N, TOKENS, image_px=IMG, seed=1)
    inp = O.prepare_inputs(ids, labels, N, images)
    times = []
    with torch.enable_grad():
        for _ in range(n):
            t0 = time.time()
            loss, grads = TO.loss_grads(sd, cfg, inp)
            TO.clip_coef(list(grads.values()), 1.0)
            times.append(time.time() - t0)
            if sum(times) > budget_s:
                break
    return times, cores


def cpu_sample(cfg_id, n, budget_s, backbone="vit", warm=1):
    """-> (manuals/s, cores, description) of a bounded CPU sample of config cfg_id."""
    if cfg_id == 3:
        times, cores = cpu_finetune_run(max(2, min(n, 3)), budget_s)
        what = "optimizer step(s) of ONE manual (forward + autograd backward + gradient norm, no AdamW update)"
    else:
        times, cores = cpu_order_run(cfg_id, n, budget_s, backbone)
        what = "manual(s) of the same workload, one per call (the reference's batch size)"
    warm = max(0, min(warm, len(times) - 1))   # a run the time budget cut short keeps at least one timed call
    timed = times[warm:]
    return len(timed) / sum(timed), cores, "%d %s after %d warm-up, oracle port (torch fp32) on %d host threads, one host process" % (
        len(timed), what, len(times) - len(timed), cores), len(timed), len(times) - len(timed)


def run_reference(args, rank):
    """--impl reference: the reference's own CPU path (oracle port), all host threads, rank 0 only."""
    if rank != 0:
        return
    n = args.steps + args.warmup
    val, cores, sample, nt, nw = cpu_sample(args.config, n, 170.0, args.backbone, warm=args.warmup)
    if args.gpus > 1:
        sample += "; at N > 1 this arm is still ONE host process on the box's cores (the ratio is N GPUs against one host)"
    line = {"metric": METRIC[args.config], "value": val, "unit": "manuals/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": nt, "warmup": nw, "ms_per_step": 1e3 / val,
            "higher_is_better": True, "scaling": "strong" if args.config == 2 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.config, 1, args.gpus, args.backbone, args.text),
            "cpu_baseline": {"value": val, "unit": "manuals/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "manuals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# =====================================================================================================
# our arm
# =====================================================================================================

class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback in the product path)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # NCCL prints its version banner to stdout when NCCL_DEBUG is set; stdout carries exactly one JSON line here
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed(self, fn, steps, warmup, profile_lib=None, sample_clocks=False):
        """W warm-up calls, then K calls bracketed by barrier + synchronize; CUDA events; max over ranks -> total ms (+ extras)."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        sampler = ClockSampler(self.local) if (sample_clocks and self.rank == 0) else None
        if sampler:
            sampler.start()
        if profile_lib is not None:
            profile_lib.msq_profile_enable(1)
        l0 = profile_lib.msq_launch_count() if profile_lib is not None else 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        self.barrier()
        ms = self.max_over_ranks(e0.elapsed_time(e1))
        extra = {"out": out, "clocks": sampler.stop() if sampler else None}
        if profile_lib is not None:
            pm, pf, pl = C.c_double(), C.c_double(), C.c_int64()
            profile_lib.msq_profile_read(C.byref(pm), C.byref(pf), C.byref(pl))
            profile_lib.msq_profile_enable(0)
            extra.update(kernel_ms=pm.value, kernel_flops=pf.value, kernel_launches=pl.value,
                         launches=int(profile_lib.msq_launch_count() - l0))
        return ms, extra


def _model_cfg(args):
    from oracle import synth
    cfg = dict(synth.BERT_BASE)
    if args.text == "roberta-large":   # sizing only: ids stay BERT-style ([CLS]=101 ...), so keep two token types
        cfg = dict(synth.ROBERTA_LARGE, type_vocab_size=2)
    vit, rn = _towers(args.backbone)
    cfg.update(vit=vit, rn=rn, para_ff=3072)
    return cfg, vit, rn


def gemm_roofline(extra, steps, ms_total, pk, kernel, note=None):
    ach = (extra["kernel_flops"] / 1e12) / (extra["kernel_ms"] / 1e3) if extra.get("kernel_ms") else None
    r = {"bound": "tensor", "kernel": kernel, "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s",
         "frac": (ach / pk["tf_sust"]) if ach else None, "traffic": None,
         "peak_source": pk["src"] + ", bf16_tflops_sustained (kernel timed inside a long step)",
         "launches_timed": extra.get("kernel_launches"), "kernel_ms_per_step": extra["kernel_ms"] / steps,
         "kernel_share_of_step": (extra["kernel_ms"] / steps) / (ms_total / steps)}
    if note:
        r["note"] = note
    return r


def run_ordering(ctx, args, cfg_id, steps, warmup, full=True):
    """configs 1 and 2: encode + beam search.  full=False -> the compact record embedded in the default line."""
    torch = ctx.torch
    from multimodal_sequencing_b200 import OrderingEngine, _lib
    from oracle import berson_oracle as O   # synthetic inputs, the parity checker and the cpu_baseline leg only
    from oracle import synth
    torch.set_grad_enabled(False)
    lib = _lib.load()
    c = CFG[cfg_id]
    N, W = c["n_steps"], c["beam"]
    precision = args.precision or ("fp32" if args.precise else "bf16x3")
    mcfg, vit, rn = _model_cfg(args)
    sd = synth.full_state_dict(mcfg, vit, seed=0, rn=rn)
    eng = OrderingEngine(sd, mcfg, precise=precision, device=ctx.dev)

    if cfg_id == 2:   # strong scaling: a fixed total, rank r orders manuals r::G
        total = c["total"]
        ids, labels, images = O.synthetic_manuals(total, N, TOKENS, image_px=IMG, seed=1)
        mine = list(range(ctx.rank, total, ctx.world))
        ids, labels, images = ids[mine], labels[mine], images[mine]
        B = len(mine)
    else:
        B = args.batch or c["batch"]
        total = B * ctx.world
        ids, labels, images = O.synthetic_manuals(B, N, TOKENS, image_px=IMG, seed=1 + ctx.rank)
    devb = eng.prepare(ids, labels, N, images).to(ctx.dev)
    ids_pin, img_pin = ids.contiguous().pin_memory(), images.contiguous().pin_memory()
    perm_host = torch.empty(B, N, dtype=torch.int32).pin_memory()
    h2d = ids_pin.numel() * 8 + img_pin.numel() * 4
    d2h = perm_host.numel() * 4

    ms_total, ex = ctx.timed(lambda: eng.order_device(devb, W), steps, warmup, profile_lib=lib, sample_clocks=full)
    perm_dev = ex["out"].cpu()
    e2e_ms, _ = ctx.timed(lambda: eng.order_raw_host(ids_pin, img_pin, N, W, perm_host), steps, max(1, warmup // 2))
    assert perm_host.tolist() == perm_dev.tolist(), "host and device paths disagree"
    assert all(sorted(p) == list(range(N)) for p in perm_host.tolist())

    rec = {"value": total * steps / (ms_total / 1e3), "ms_per_step": ms_total / steps,
           "e2e_value": total * steps / (e2e_ms / 1e3), "h2d": h2d, "d2h": d2h, "launches": ex["launches"], "clocks": ex["clocks"],
           "extra": ex, "precision": precision, "B": B, "total": total, "ms_total": ms_total}

    # ---- parity of the timed batch against the fp32 CUDA-core path (and the fast mode beside it)
    rec["parity"] = None
    if not args.no_parity and precision != "fp32":
        n_chk = min(B, 64)
        sub = eng.prepare(ids[:n_chk], labels[:n_chk], N, images[:n_chk]).to(ctx.dev)
        enc_x = {k: v.float().cpu() for k, v in eng.encode(sub).items()} if full else None
        del eng
        torch.cuda.empty_cache()
        e32 = OrderingEngine(sd, mcfg, precise="fp32", device=ctx.dev)
        p32 = e32.order_device(sub, W).cpu()
        worst = None
        if full:
            enc32 = e32.encode(sub)
            worst = max(float((enc_x[k] - enc32[k].float().cpu()).norm() / (enc32[k].float().cpu().norm() + 1e-30))
                        for k in ("sents", "para", "h0", "key", "cls", "cls_mat", "cls_score", "score_mat", "his1", "his2"))
        del e32
        torch.cuda.empty_cache()
        same = int((perm_dev[:n_chk] == p32).all(1).sum())
        same = int(ctx.sum_over_ranks(same))
        rec["parity"] = {"mode": precision, "manuals_checked": n_chk * ctx.world, "identical_permutations_vs_fp32_path": same,
                         "agreement": same / (n_chk * ctx.world), "worst_rel_l2_encode_vs_fp32_path": worst,
                         "checker": "this library's fp32 FFMA mode, itself pinned to the oracle / reference goldens in tests/"}
        if full and precision == "bf16x3" and cfg_id == 1:
            fast = OrderingEngine(sd, mcfg, precise="bf16", device=ctx.dev)
            fms, fex = ctx.timed(lambda: fast.order_device(devb, W), max(2, steps // 2), 2, profile_lib=lib)
            pf_ = fast.order_device(sub, W).cpu()
            fsame = int(ctx.sum_over_ranks(int((pf_ == p32).all(1).sum())))
            fach = (fex["kernel_flops"] / 1e12) / (fex["kernel_ms"] / 1e3) if fex.get("kernel_ms") else None
            rec["fast_mode"] = {"dtype": "bf16", "value": total * max(2, steps // 2) / (fms / 1e3), "unit": "manuals/s",
                                "agreement_vs_fp32_path": fsame / (n_chk * ctx.world),
                                # the same tcgen05 GEMM kernel with one MMA per product: algorithmic = executed FLOPs
                                "gemm_tflops": fach, "gemm_frac_of_bf16_peak": (fach / peaks()["tf_sust"]) if fach else None,
                                "note": "plain bf16 operands: ~2.5x the headline's speed, outside north_star's numeric gate "
                                        "(encoder outputs ~1e-2, some permutations flip under random-init margins)"}
            del fast
            torch.cuda.empty_cache()
    return rec


def line_ordering(ctx, args, cfg_id, rec, steps, warmup):
    pk = peaks()
    N = CFG[cfg_id]["n_steps"]
    flop_per_manual = N * (N - 1) * FLOP_PER_PAIR
    x3 = rec["precision"] == "bf16x3"
    line = {"metric": METRIC[cfg_id], "value": rec["value"], "unit": "manuals/s", "n_gpus": ctx.world, "steps": steps, "warmup": warmup,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong" if cfg_id == 2 else "weak",
            "vs_baseline": None, "dtype": {"fp32": "f32"}.get(rec["precision"], rec["precision"]), "data": "synthetic",
            "config": config_dict(cfg_id, rec["B"], ctx.world, args.backbone, args.text), "impl": "ours",
            "e2e": {"value": rec["e2e_value"], "unit": "manuals/s", "h2d_bytes_per_step": rec["h2d"], "d2h_bytes_per_step": rec["d2h"],
                    "api": "OrderingEngine.order_raw_host -> msq_order_manuals_raw_host: DataLoader tuple (token rows + step images, pinned "
                           "host memory) -> orders; pair expansion on the device, inside the timed region"},
            "gpu_launches": rec["launches"], "clocks": rec["clocks"], "parity": rec["parity"],
            "roofline": gemm_roofline(rec["extra"], steps, rec["ms_total"], pk,
                                      "msq::gemm_tc_kernel (tcgen05 GEMM, all encoder linears" + ("; bf16x3: 3 MMAs per algorithmic product)" if x3 else ")"),
                                      note=("`achieved` counts ALGORITHMIC FLOPs (2MNK); the bf16x3 kernel executes 3x that on the tensor pipe, i.e. "
                                            "its MMA rate is 3 x achieved (compare THAT with the bf16 peak: executed_frac)") if x3 else None)}
    if x3 and line["roofline"]["achieved"]:
        line["roofline"]["executed_tflops"] = 3 * line["roofline"]["achieved"]
        line["roofline"]["executed_frac"] = 3 * line["roofline"]["achieved"] / pk["tf_sust"]
    if (args.text, args.backbone) == ("bert-base", "vit"):
        line["roofline"]["whole_step_tflops_per_gpu"] = flop_per_manual * rec["B"] * steps / (rec["ms_total"] / 1e3) / 1e12
    tp = os.path.join(ROOT, "profiles", "r2_gemm_tc_traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp))
        line["roofline"]["traffic"] = t.get("traffic_bytes_per_launch")
        line["roofline"]["traffic_note"] = t.get("note")
    if "fast_mode" in rec:
        line["fast_mode"] = rec["fast_mode"]
    return line


def run_finetune(ctx, args, steps, warmup, full=True):
    """config 3: msq_train_step + gradient all-reduce + msq_adamw_step."""
    torch = ctx.torch
    from multimodal_sequencing_b200 import OrderingEngine, _lib, sharding
    from oracle import berson_oracle as O
    from oracle import synth
    torch.set_grad_enabled(False)
    lib = _lib.load()
    N = CFG[3]["n_steps"]
    B = args.batch or CFG[3]["batch"]
    precision = args.precision or ("fp32" if args.precise else "bf16")
    mcfg, vit, rn = _model_cfg(args)
    sd = synth.full_state_dict(mcfg, vit, seed=0, rn=rn)
    eng = OrderingEngine(sd, mcfg, precise=precision, device=ctx.dev)
    del sd
    ids, labels, images = O.synthetic_manuals(B, N, TOKENS, image_px=IMG, seed=1 + ctx.rank)
    host = eng.prepare(ids, labels, N, images)
    pinned = type(host)(**{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host.__dict__.items()})
    devb = host.to(ctx.dev)
    grads = eng.new_grad_buffer()
    eng.set_dropout(args.dropout, args.dropout, args.dropout, seed=1234 + ctx.rank)   # every rank drops its own elements, as DDP replicas do
    h2d = sum(v.numel() * v.element_size() for v in (pinned.input_ids, pinned.token_type_ids, pinned.attention_mask, pinned.sep_positions,
                                                     pinned.images, pinned.img_index, pinned.ground_truth, pinned.pairwise_labels))
    reducer = sharding.GradientReducer(eng, grads) if ctx.world > 1 else None

    def step(batch):
        grads.zero_()
        loss = eng.train_step(batch, grads)
        scale = reducer.allreduce() if reducer else 1.0
        eng.adamw_step(grads, args.lr, eps=1e-8, weight_decay=0.0, max_grad_norm=1.0, grad_scale=scale)
        return loss

    losses = []
    ms_total, ex = ctx.timed(lambda: losses.append(step(devb)) or losses[-1], steps, warmup, profile_lib=lib, sample_clocks=full)
    loss_host = torch.empty(1).pin_memory()

    def e2e_step():
        loss_host.copy_(step(pinned.to(ctx.dev, non_blocking=True)).reshape(1), non_blocking=True)
        return loss_host
    e2e_ms, _ = ctx.timed(e2e_step, steps, max(2, warmup // 2))
    total = B * ctx.world
    traj = [float(x) for x in losses[::max(1, len(losses) // 6)]] + [float(loss_host)]
    del eng, grads
    torch.cuda.empty_cache()
    return {"value": total * steps / (ms_total / 1e3), "ms_per_step": ms_total / steps, "e2e_value": total * steps / (e2e_ms / 1e3),
            "h2d": h2d, "d2h": 4, "launches": ex["launches"], "clocks": ex["clocks"], "extra": ex, "precision": precision, "B": B,
            "ms_total": ms_total, "loss_trajectory": traj, "dropout": args.dropout,
            "collective": ("NCCL all-reduce (SUM) of the flat fp32 gradient buffer in %d buckets" % reducer.n_buckets) if reducer else None}


def line_finetune(ctx, args, rec, steps, warmup):
    pk = peaks()
    N = CFG[3]["n_steps"]
    line = {"metric": METRIC[3], "value": rec["value"], "unit": "manuals/s", "n_gpus": ctx.world, "steps": steps, "warmup": warmup,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32"}.get(rec["precision"], rec["precision"]), "data": "synthetic",
            "config": config_dict(3, rec["B"], ctx.world, args.backbone, args.text), "impl": "ours",
            "e2e": {"value": rec["e2e_value"], "unit": "manuals/s", "h2d_bytes_per_step": rec["h2d"], "d2h_bytes_per_step": rec["d2h"],
                    "api": "OrderingEngine.train_step + gradient all-reduce + adamw_step from a pinned host PairBatch; the loss comes back"},
            "gpu_launches": rec["launches"], "clocks": rec["clocks"], "loss_trajectory": rec["loss_trajectory"], "collective": rec["collective"],
            "dropout": rec["dropout"],
            "roofline": gemm_roofline(rec["extra"], steps, rec["ms_total"], pk, "msq::gemm_tc_kernel (forward, dgrad and wgrad GEMMs)",
                                      note="split-K slices of a weight gradient run concurrently on auxiliary streams: their event durations "
                                           "overlap, so `achieved` is a lower bound")}
    line["roofline"]["whole_step_tflops_per_gpu"] = 3 * N * (N - 1) * FLOP_PER_PAIR * rec["B"] * steps / (rec["ms_total"] / 1e3) / 1e12
    return line


def run_decode(ctx, args, steps, warmup, full=True):
    """config 4: pre-projections (XG, T4) + beam search from seeded encoder outputs (H = 768)."""
    torch = ctx.torch
    from multimodal_sequencing_b200 import OrderingEngine, _lib
    from oracle import synth
    torch.set_grad_enabled(False)
    lib = _lib.load()
    c = CFG[4]
    N, W, H = c["n_steps"], c["beam"], 768
    B = args.batch or c["batch"]
    precision = args.precision or ("fp32" if args.precise else "bf16")
    cfg = dict(hidden_size=H, num_hidden_layers=1, num_attention_heads=12, intermediate_size=64, vocab_size=64,
               max_position_embeddings=8, vit=None, para_ff=64)
    eng = OrderingEngine(synth.full_state_dict(cfg, None, seed=0, ff=64), cfg, precise=precision, device=ctx.dev)
    enc = {k: v.to(ctx.dev) for k, v in synth.synthetic_encode(N, H, seed=3 + ctx.rank, B=B).items()}
    enc_pin = {k: v.cpu().contiguous().pin_memory() for k, v in enc.items()}
    ms_total, ex = ctx.timed(lambda: eng.beam_search(enc, N, W), steps, warmup, profile_lib=lib, sample_clocks=full)
    perm = ex["out"].cpu()
    assert all(sorted(p) == list(range(N)) for p in perm.tolist())
    keys = ("sents", "key", "h0", "cls_mat", "score_mat")
    h2d = sum(enc_pin[k].numel() * 4 for k in keys)

    def e2e_step():
        d = {k: enc_pin[k].to(ctx.dev, non_blocking=True) for k in keys}
        return eng.beam_search(d, N, W).cpu()
    e2e_ms, _ = ctx.timed(e2e_step, steps, max(2, warmup // 2))
    total = B * ctx.world
    # algorithmic bytes (base-tensor formulation) per manual and decode step, and the recurrent GEMM work
    per_step = N * N * 770 * 4 + N * 768 * 4 + W * (4 * 768 * 4 + 2 * N * 4)
    live, rows = 1, 0
    for _ in range(N - 1):
        rows += live
        live = min(W, live * N)
    del eng
    torch.cuda.empty_cache()
    return {"value": total * steps / (ms_total / 1e3), "ms_per_step": ms_total / steps, "e2e_value": total * steps / (e2e_ms / 1e3),
            "h2d": h2d, "d2h": perm.numel() * 4, "launches": ex["launches"], "clocks": ex["clocks"], "extra": ex, "precision": precision,
            "B": B, "ms_total": ms_total, "alg_bytes_per_step": B * per_step * (N - 1),
            "recurrent_flops_per_step": 2.0 * B * rows * 5 * H * H + 2.0 * B * (N * N * 4 * H * (H + 2) + (N + 1) * 4 * H * H)}


def line_decode(ctx, args, rec, steps, warmup):
    pk = peaks()
    gbs = rec["alg_bytes_per_step"] / (rec["ms_per_step"] / 1e3) / 1e9
    tc = rec["precision"] != "fp32"
    line = {"metric": METRIC[4], "value": rec["value"], "unit": "manuals/s", "n_gpus": ctx.world, "steps": steps, "warmup": warmup,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("f32 state / scoring; recurrent + pre-projection GEMMs on tcgen05 as " +
                      ("bf16x6 (three bf16 planes, six MMAs per product, fp32-grade)" if os.environ.get("MSQ_DEC_X3", "1") == "0"
                       else "bf16x3 (hi + lo bf16 planes, three MMAs per product; MSQ_DEC_X3=0: bf16x6)")) if tc else "f32",
            "data": "synthetic (seeded encoder outputs)", "config": config_dict(4, rec["B"], ctx.world, args.backbone, args.text), "impl": "ours",
            "e2e": {"value": rec["e2e_value"], "unit": "manuals/s", "h2d_bytes_per_step": rec["h2d"], "d2h_bytes_per_step": rec["d2h"],
                    "api": "OrderingEngine.beam_search -> msq_beam_search from pinned host encoder outputs; the orders come back"},
            "gpu_launches": rec["launches"], "clocks": rec["clocks"],
            "roofline": {"bound": "hbm", "kernel": "whole decode (cell + tcgen05 GEMM + dec_select per step, XG / T4 pre-projections)",
                         "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"], "traffic": None,
                         "peak_source": pk["src"] + ", hbm_gbs",
                         "note": "algorithmic bytes of SURVEY 8(d) over the WHOLE decode time.  The decode is not HBM-bound at these sizes: "
                                 "per manual and step it evaluates beam x N x H tanh terms (dec_select is issue-bound, profiles/) and "
                                 "2 x 5 x H^2 MACs per live beam; the GEMM work is given below against the tensor peak",
                         "gemm_algorithmic_tflops": rec["recurrent_flops_per_step"] / (rec["ms_per_step"] / 1e3) / 1e12,
                         "gemm_kernel_tflops": ((rec["extra"]["kernel_flops"] / 1e12) / (rec["extra"]["kernel_ms"] / 1e3)
                                                if rec["extra"].get("kernel_ms") else None),
                         "gemm_kernel_share_of_step": ((rec["extra"]["kernel_ms"] / steps) / rec["ms_per_step"]
                                                       if rec["extra"].get("kernel_ms") else None)}}
    return line


def compact(line):
    keep = ("metric", "value", "unit", "ms_per_step", "scaling", "dtype", "gpu_launches", "parity", "collective", "steps", "warmup")
    d = {k: line[k] for k in keep if k in line}
    d["e2e"] = line["e2e"]["value"]
    d["manuals_per_step_per_gpu"] = line["config"]["manuals_per_step_per_gpu"]
    r = line["roofline"]
    d["roofline"] = {k: r[k] for k in ("bound", "achieved", "peak", "unit", "frac", "gemm_kernel_tflops", "executed_frac") if k in r}
    return d


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    ctx = Ctx()
    steps, warmup = args.steps, max(args.warmup, 1)
    if args.config in (1, 2):
        rec = run_ordering(ctx, args, args.config, steps, warmup)
        line = line_ordering(ctx, args, args.config, rec, steps, warmup) if ctx.rank == 0 else None
    elif args.config == 3:
        rec = run_finetune(ctx, args, steps, warmup)
        line = line_finetune(ctx, args, rec, steps, warmup) if ctx.rank == 0 else None
    else:
        rec = run_decode(ctx, args, steps, warmup)
        line = line_decode(ctx, args, rec, steps, warmup) if ctx.rank == 0 else None

    if args.config == 1 and not args.no_extra_configs and (args.text, args.backbone) == ("bert-base", "vit"):
        # short runs of the other BASELINE configs, embedded so that the driver's single default command records them
        saved = (args.batch, args.precision, args.no_parity)
        args.batch, args.precision = None, None
        extra = {}
        r2 = run_ordering(ctx, args, 2, 2, 2, full=False)
        if ctx.rank == 0:
            extra["2"] = compact(line_ordering(ctx, args, 2, r2, 2, 2))
        r3 = run_finetune(ctx, args, 3, 3, full=False)
        if ctx.rank == 0:
            extra["3"] = compact(line_finetune(ctx, args, r3, 3, 3))
        r4 = run_decode(ctx, args, 5, 3, full=False)
        if ctx.rank == 0:
            extra["4"] = compact(line_decode(ctx, args, r4, 5, 3))
            line["configs"] = extra
        args.batch, args.precision, args.no_parity = saved

    if ctx.rank == 0:
        if not args.no_cpu_baseline and ctx.world == 1:
            val, cores, sample, _, _ = cpu_sample(args.config, 3, 25.0, args.backbone)
            line["cpu_baseline"] = {"value": val, "unit": "manuals/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
