#!/usr/bin/env python
"""bench_finetune.py — BASELINE.json configs[3] (multimodal fine-tuning step, AdamW, NCCL gradient all-reduce at 1/2/4/8 B200).
Kept as an entry point for existing command lines; the implementation lives in bench.py (`python bench.py --config 3 ...`,
same flags, same one-line JSON contract)."""
import sys

import bench

if __name__ == "__main__":
    sys.argv = [sys.argv[0], "--config", "3"] + sys.argv[1:]
    bench.main()
