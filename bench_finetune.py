#!/usr/bin/env python
"""bench_finetune.py — BASELINE.json configs[3]: multimodal fine-tuning step (bf16, AdamW) on RecipeQA-shaped synthetic
manuals (6 steps, 64 tokens/step, 224-px images, 30 ordered pairs of 227 joint tokens) at 1/2/4/8 B200.  Secondary line
(bench.py carries the headline eval metric); same JSON shape.

    python bench_finetune.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--precise] [--impl ours|reference]
    torchrun ... bench_finetune.py --gpus N ...          (one rank per GPU; ONE NCCL all-reduce of the flat gradient buffer)

A "step" = what trainers/train.py:340-363 does per optimizer step with gradient_accumulation_steps = 1:
loss = model(inputs)[0]; loss.backward(); clip_grad_norm_(1.0); AdamW.step()  -- here msq_train_step (forward with recorded
activations, loss, full backward), all-reduce of the gradients when N > 1, msq_adamw_step (clip + transformers.AdamW +
re-pack of every derived weight copy).  Dropout is off (p = 0, the parity configuration of SURVEY §8(d) cfg4).

  value : manuals/s, batch resident in HBM, CUDA events, max over ranks
  e2e   : the same step fed from pinned HOST buffers (H2D of ids / masks / images inside the timed region, D2H of the loss)
  roofline : tcgen05 GEMM launches of the step (forward, dgrad and wgrad all run on it): algorithmic FLOP / event time
  --impl reference : the oracle port (torch autograd, fp32, all host threads) doing the same step on the host cores
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from bench import ClockSampler, peaks  # noqa: E402

N_STEPS, TOKENS, IMG = 6, 64, 224
FWD_FLOP_PER_MANUAL = 1.751e12   # 30 pairs x 58.37 GFLOP; backward = 2x


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="manuals per optimizer step per GPU (scripts/recipeqa_finetune.sh: 1-8)")
    ap.add_argument("--precise", action="store_true")
    ap.add_argument("--lr", type=float, default=5e-6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def config_dict(batch, world):
    return {"workload": "configs[3]: multimodal BERSON + CLIP ViT-B/32 fine-tuning step, 6 steps x 64 tokens + 224px images, AdamW "
                        "lr 5e-6 eps 1e-8 wd 0, clip 1.0, dropout 0",
            "manuals_per_step_per_gpu": batch, "n_steps": N_STEPS, "tokens_per_step": TOKENS, "pairs_per_manual": N_STEPS * (N_STEPS - 1),
            "joint_tokens_per_pair": 227,
            "parallelism": "replicas, one all-reduce of the flat fp32 gradient buffer per optimizer step (dp%d)" % world,
            "l2": "activations recorded for the backward pass (GBs) exceed the 126 MB L2"}


def cpu_reference_steps(n, budget_s):
    """Oracle port of one fine-tuning step (forward + autograd backward + gradient norm) of ONE manual on the host cores."""
    import torch
    from oracle import berson_oracle as O
    from oracle import synth
    from oracle import train_oracle as TO
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vit = dict(synth.VIT_B32)
    cfg = dict(num_hidden_layers=12, num_attention_heads=12, vit=vit)
    sd = synth.full_state_dict(None, vit, seed=0)
    ids, labels, images = O.synthetic_manuals(1, N_STEPS, TOKENS, image_px=IMG, seed=1)
    inp = O.prepare_inputs(ids, labels, N_STEPS, images)
    times = []
    for _ in range(n):
        t0 = time.time()
        loss, grads = TO.loss_grads(sd, cfg, inp)
        total, coef = TO.clip_coef(list(grads.values()), 1.0)
        times.append(time.time() - t0)
        if sum(times) > budget_s:
            break
    return times, cores


def run_reference(args, rank):
    if rank != 0:
        return
    times, cores = cpu_reference_steps(min(args.steps + args.warmup, 3), 150)
    timed = times[1:] if len(times) > 1 else times
    ms = 1e3 * sum(timed) / len(timed)
    val = 1e3 / ms
    sample = "%d optimizer step(s) of ONE manual (forward + autograd backward + grad norm; the AdamW update itself is left out) " \
             "after %d warm-up, oracle port on %d host threads" % (len(timed), len(times) - len(timed), cores)
    print(json.dumps({"metric": "6-step manuals fine-tuned/sec", "value": val, "unit": "manuals/s", "impl": "reference", "n_gpus": args.gpus,
                      "steps": len(timed), "warmup": len(times) - len(timed), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(1, args.gpus),
                      "cpu_baseline": {"value": val, "unit": "manuals/s", "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": "manuals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


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
    from multimodal_sequencing_b200 import OrderingEngine, _lib, sharding
    from oracle import berson_oracle as O   # synthetic inputs only
    from oracle import synth

    torch.set_grad_enabled(False)
    if not torch.cuda.is_available():
        raise SystemExit("bench_finetune.py needs a CUDA device (no CPU fallback in the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(synth.BERT_BASE)
    vit = dict(synth.VIT_B32)
    cfg.update(vit=vit, rn=None, para_ff=3072)
    sd = synth.full_state_dict(cfg, vit, seed=0)
    eng = OrderingEngine(sd, cfg, precise=args.precise, device=dev)
    del sd
    lib = _lib.load()
    B = args.batch
    ids, labels, images = O.synthetic_manuals(B, N_STEPS, TOKENS, image_px=IMG, seed=1 + rank)
    host = eng.prepare(ids, labels, N_STEPS, images)
    pinned = type(host)(**{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host.__dict__.items()})
    devb = host.to(dev)
    grads = eng.new_grad_buffer()
    h2d = sum(v.numel() * v.element_size() for v in (pinned.input_ids, pinned.token_type_ids, pinned.attention_mask, pinned.sep_positions,
                                                     pinned.images, pinned.img_index, pinned.ground_truth, pinned.pairwise_labels))

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

    def step(batch):
        grads.zero_()
        loss = eng.train_step(batch, grads)
        scale = sharding.allreduce_gradients(grads)          # SUM over ranks (no-op at world size 1) -> 1 / world
        eng.adamw_step(grads, args.lr, eps=1e-8, weight_decay=0.0, max_grad_norm=1.0, grad_scale=scale)
        return loss

    losses = []
    for _ in range(args.warmup):
        losses.append(float(step(devb)))
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.msq_profile_enable(1)
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step(devb)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launch_count() - l0
    pm, pf, pl = C.c_double(), C.c_double(), C.c_int64()
    lib.msq_profile_read(C.byref(pm), C.byref(pf), C.byref(pl))
    lib.msq_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    losses.append(float(loss))

    # end to end: the batch travels from pinned host memory every step, the loss comes back
    loss_host = torch.empty(1).pin_memory()
    for _ in range(1):
        loss_host.copy_(step(pinned.to(dev, non_blocking=True)).reshape(1))
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss_host.copy_(step(pinned.to(dev, non_blocking=True)).reshape(1), non_blocking=True)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    losses.append(float(loss_host))
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    pk = peaks()
    total = world * B * args.steps
    value = total / (ms_total / 1e3)
    ach = (pf.value / 1e12) / (pm.value / 1e3) if pm.value > 0 else None
    line = {"metric": "6-step manuals fine-tuned/sec", "value": value, "unit": "manuals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precise else "bf16", "data": "synthetic", "config": config_dict(B, world), "impl": "ours",
            "e2e": {"value": total / (e2e_ms / 1e3), "unit": "manuals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "api": "OrderingEngine.train_step + sharding.allreduce_gradients + adamw_step from a pinned host PairBatch"},
            "gpu_launches": launches, "clocks": clocks, "loss_trajectory": losses,
            "roofline": {"bound": "tensor", "kernel": "msq::gemm_tc_kernel (forward, dgrad and wgrad GEMMs)", "achieved": ach,
                         "note": "the split-K slices of a weight gradient run concurrently on auxiliary streams: their event durations overlap, so "
                                 "the per-launch sum over-counts kernel time and `achieved` is a lower bound (MSQ_WGRAD_SPLITK=0 gives the clean figure)",
                         "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": (ach / pk["tf_sust"]) if ach else None, "traffic": None,
                         "peak_source": pk["src"] + ", bf16_tflops_sustained", "launches_timed": pl.value,
                         "kernel_ms_per_step": pm.value / args.steps,
                         "kernel_share_of_step": (pm.value / args.steps) / (ms_total / args.steps),
                         "whole_step_tflops_per_gpu": 3 * FWD_FLOP_PER_MANUAL * B * args.steps / (ms_total / 1e3) / 1e12}}
    if not args.no_cpu_baseline and world == 1:
        with torch.enable_grad():
            times, cores = cpu_reference_steps(2, 25.0)
        timed = times[1:] if len(times) > 1 else times
        line["cpu_baseline"] = {"value": len(timed) / sum(timed), "unit": "manuals/s", "cores": cores, "kind": "port",
                                "sample": "%d optimizer step(s) of ONE manual of the same workload (forward + autograd backward + gradient "
                                          "norm, no AdamW update) after %d warm-up, oracle port (torch fp32) on %d host threads" %
                                          (len(timed), len(times) - len(timed), cores)}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
