#!/usr/bin/env python
"""bench.py — 5-step manuals ordered / second (beam = 4) on N B200s, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    torchrun ... bench.py --gpus N --steps K --warmup W          (one rank per GPU, NCCL)

Workload (BASELINE.json configs[1]): multimodal BERSON + CLIP ViT-B/32, WikiHow-shaped synthetic manuals
(5 steps, 224-px images, 64 tokens/step, 20 ordered pairs, 227 joint tokens/pair), beam = 4, random-init
weights (oracle/synth.py seed 0), eval only (no collective on the data path; manuals shard by rank).
A "step" = one batch of `--batch` manuals per GPU through encode + beam search.

  value : manuals/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through OrderingEngine.order_host (C ABI msq_order_manuals_host) from pinned HOST
          buffers: H2D of ids/masks/images and D2H of the permutations are inside the timed region
  roofline : the tcgen05 GEMM kernel (dominant): algorithmic 2*M*N*K per launch / CUDA-event duration of
          every launch inside the timed region, against MEASURED_PEAKS.json bf16_tflops_sustained
  cpu_baseline : the oracle port of the reference path on this box's host cores, bounded sample
  --impl reference : the reference's CPU implementation (oracle port; /root/reference does not exist on
          the GPU box) timed alone with all host threads, same metric / config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_STEPS, BEAM, TOKENS, IMG = 5, 4, 64, 224
FLOP_PER_MANUAL = 1.167e12  # 20 pairs x 58.37 GFLOP


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="manuals per step per GPU")
    ap.add_argument("--precise", action="store_true", help="fp32 FFMA parity mode instead of bf16 tcgen05")
    ap.add_argument("--precision", default=None, choices=["bf16", "bf16x3", "fp32"],
                    help="encoder arithmetic: bf16 tcgen05 operands, bf16x3 (hi+lo bf16 operands, 3 MMAs per product) or fp32 FFMA")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--text", default="bert-base", choices=["bert-base", "roberta-large"],
                    help="joint encoder sizing: BERT-base (BASELINE configs, the headline) or the roberta-large config the "
                         "reference's scripts pass (H=1024, 24 layers); secondary line")
    ap.add_argument("--backbone", default="vit", choices=["vit", "rn50"],
                    help="visual tower: ViT-B/32 (BASELINE configs[1], the headline) or the reference's wired default RN50")
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


def config_dict(batch, n_gpus, backbone="vit"):
    return {"workload": ("configs[1]: multimodal BERSON + CLIP ViT-B/32, 5 steps x 64 tokens + 224px images, beam=4, eval"
                         if backbone == "vit" else
                         "configs[1] with the reference's wired RN50 tower instead of ViT-B/32 (secondary line, not the headline)"),
            "manuals_per_step_per_gpu": batch, "n_steps": N_STEPS, "beam": BEAM, "tokens_per_step": TOKENS,
            "pairs_per_manual": N_STEPS * (N_STEPS - 1), "joint_tokens_per_pair": 227,
            "parallelism": "manuals sharded by rank, no data-path collective (dp%d)" % n_gpus,
            "l2": "per-step inputs (%.0f MB images) and activations exceed the 126 MB L2" %
                  (batch * N_STEPS * 3 * IMG * IMG * 4 / 1e6)}


def cpu_reference_run(max_manuals, budget_s, seed=1, backbone="vit"):
    """Oracle port of berson_pointer_network on the host cores (reference semantics: one manual per call)."""
    import torch
    from oracle import berson_oracle as O
    from oracle import synth
    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vit, rn = _towers(backbone)
    cfg = dict(num_hidden_layers=12, num_attention_heads=12, vit=vit, rn=rn)
    sd = synth.full_state_dict(None, vit, seed=0, rn=rn)
    ids, labels, images = O.synthetic_manuals(max_manuals, N_STEPS, TOKENS, image_px=IMG, seed=seed)
    times = []
    t_start = time.time()
    for b in range(max_manuals):
        t0 = time.time()
        O.order_manuals(sd, cfg, ids[b:b + 1], labels[b:b + 1], N_STEPS, BEAM, images[b:b + 1])
        times.append(time.time() - t0)
        if time.time() - t_start > budget_s:
            break
    return times, cores


def run_reference(args, rank):
    """--impl reference: the reference's own CPU path (oracle port), all host threads, rank 0 only."""
    if rank != 0:
        return
    n = args.steps + args.warmup
    times, cores = cpu_reference_run(n, budget_s=170.0, backbone=args.backbone)
    timed = times[min(args.warmup, max(0, len(times) - 1)):]
    ms = 1e3 * sum(timed) / len(timed)
    val = 1e3 / ms
    sample = "%d manual(s) timed one per call after %d warm-up (torch fp32, %d threads)%s" % (
        len(timed), len(times) - len(timed), cores, "" if len(times) == n else "; stopped early at the 170 s budget")
    line = {"metric": "5-step manuals ordered/sec (beam=4)", "value": val, "unit": "manuals/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": len(timed), "warmup": len(times) - len(timed), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(1, args.gpus, args.backbone),
            "cpu_baseline": {"value": val, "unit": "manuals/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "manuals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from multimodal_sequencing_b200 import OrderingEngine, _lib
    from oracle import berson_oracle as O   # synthetic inputs + cpu_baseline leg only
    from oracle import synth
    import ctypes as C

    torch.set_grad_enabled(False)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback in the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner to stdout when NCCL_DEBUG is set; stdout carries exactly one JSON line here
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    cfg = dict(synth.BERT_BASE)
    if args.text == "roberta-large":   # sizing only: ids stay BERT-style ([CLS]=101 ...), so keep two token types
        cfg = dict(synth.ROBERTA_LARGE, type_vocab_size=2)
        args.no_cpu_baseline = True
    vit, rn = _towers(args.backbone)
    cfg.update(vit=vit, rn=rn, para_ff=3072)
    sd = synth.full_state_dict(cfg, vit, seed=0, rn=rn)
    precision = args.precision or ("fp32" if args.precise else "bf16")
    eng = OrderingEngine(sd, cfg, precise=precision, device=dev)
    del sd
    lib = _lib.load()

    B = args.batch
    ids, labels, images = O.synthetic_manuals(B, N_STEPS, TOKENS, image_px=IMG, seed=1 + rank)
    host = eng.prepare(ids, labels, N_STEPS, images)
    pinned = type(host)(**{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host.__dict__.items()})
    devb = host.to(dev)
    perm_host = torch.empty(B, N_STEPS, dtype=torch.int32).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in (pinned.input_ids, pinned.token_type_ids, pinned.attention_mask,
                                                     pinned.sep_positions, pinned.images, pinned.img_index))
    d2h = perm_host.numel() * perm_host.element_size()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident throughput (value) + per-launch roofline of the GEMM kernel
    for _ in range(args.warmup):
        perm = eng.order_device(devb, BEAM)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.msq_profile_enable(1)
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        perm = eng.order_device(devb, BEAM)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launch_count() - l0
    pm, pf, pl = C.c_double(), C.c_double(), C.c_int64()
    lib.msq_profile_read(C.byref(pm), C.byref(pf), C.byref(pl))
    lib.msq_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    perm_dev = perm.cpu()

    # ---------------- end to end from pinned host buffers through the C ABI
    for _ in range(max(1, args.warmup // 2)):
        eng.order_host(pinned, BEAM, perm_host)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        eng.order_host(pinned, BEAM, perm_host)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    assert perm_host.tolist() == perm_dev.tolist(), "host and device paths disagree"
    assert all(sorted(p) == list(range(N_STEPS)) for p in perm_host.tolist())

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    pk = peaks()
    traffic = None  # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
    tp = os.path.join(ROOT, "profiles", "r1h_gemm_tc_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("traffic_bytes_per_launch")
    total = world * B * args.steps
    value = total / (ms_total / 1e3)
    e2e = total / (e2e_ms / 1e3)
    ach = (pf.value / 1e12) / (pm.value / 1e3) if pm.value > 0 else None
    line = {"metric": "5-step manuals ordered/sec (beam=4)", "value": value, "unit": "manuals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32"}.get(precision, precision), "data": "synthetic",
            "config": dict(config_dict(B, world, args.backbone), text_encoder=args.text), "impl": "ours",
            "e2e": {"value": e2e, "unit": "manuals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "OrderingEngine.order_host -> msq_order_manuals_host (pinned host buffers)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "msq::gemm_tc_kernel (tcgen05 bf16 GEMM, all encoder linears; LayerNorms ride in its epilogues)",
                         "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                         "frac": (ach / pk["tf_sust"]) if ach else None, "traffic": traffic,
                         "traffic_note": "mean dram read+write bytes per launch over 8 BERT-layer GEMM launches at 640 pair rows, "
                                         "ncu --set full (profiles/r1h_gemm_tc_traffic.json); algorithmic bytes beside it there",
                         "peak_source": pk["src"] + ", bf16_tflops_sustained (kernel timed inside a long step)",
                         "launches_timed": pl.value, "kernel_ms_per_step": pm.value / args.steps,
                         "kernel_share_of_step": (pm.value / args.steps) / (ms_total / args.steps),
                         "whole_step_tflops_per_gpu": (FLOP_PER_MANUAL * B * args.steps / (ms_total / 1e3) / 1e12
                                                       if (args.text, args.backbone) == ("bert-base", "vit") else None)},
            }
    if not args.no_cpu_baseline and world == 1:
        times, cores = cpu_reference_run(3, budget_s=25.0, backbone=args.backbone)
        timed = times[1:] if len(times) > 1 else times
        cv = len(timed) / sum(timed)
        line["cpu_baseline"] = {"value": cv, "unit": "manuals/s", "cores": cores, "kind": "port",
                                "sample": "%d manual(s) of the same workload, one per call after %d warm-up, oracle port "
                                          "(torch fp32) on %d host threads" % (len(timed), len(times) - len(timed), cores)}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
