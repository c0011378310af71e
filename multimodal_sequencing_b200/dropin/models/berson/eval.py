"""models/berson/eval.py of the reference: berson_evaluate (39-187) and cal_result (190-260).

Same signature, same artefacts (`output_order.txt` lines `pred ||| truth`, `eval_results_split_*.txt`) and
the same result dict {"acc_dev","pmr_dev","taus_dev"}.  Difference: a DataLoader batch of ANY size goes to
the device in one call (batched beam search); the reference supports batch size 1 only and
produces the same per-manual permutations."""
import itertools
import logging
import os

import numpy as np
import torch
from torch.utils.data import DataLoader, SequentialSampler

from .modeling_bert import berson_pointer_network

logger = logging.getLogger(__name__)


def cal_result(truth, predicted, best_acc=None, f=None, args=None):
    """eval.py:190-260 -> (mean per-manual accuracy, perfect-match ratio, mean Kendall tau)."""
    right = total = pmr_right = 0
    taus, accs = [], []
    for t, p in zip(truth, predicted):
        if np.asarray(t).ndim > 1:
            t = t[0]
        if len(p) == 1:
            right += 1; total += 1; pmr_right += 1
            accs.append(1); taus.append(1)
            continue
        eq = np.equal(t, p)
        right += eq.sum()
        accs.append(eq.sum() / len(t))
        total += len(t)
        pmr_right += eq.all()
        s_t = set(itertools.combinations(t, 2))
        s_p = set(itertools.combinations(p, 2))
        cn_2 = len(p) * (len(p) - 1) / 2
        taus.append(1 - 2 * (len(s_p) - len(s_p.intersection(s_t))) / cn_2)
    if best_acc is not None:
        best_acc.append(right / max(total, 1))
    if f is not None:
        f.close()
    return float(np.mean(accs)), pmr_right / len(truth), float(np.mean(taus))


def berson_evaluate(args, model, load_and_cache_examples, tokenizer, prefix="", data_split="test", human_evaluate=False):
    results = {}
    for eval_task, eval_output_dir in zip(args.task_names, [args.output_dir] * len(args.task_names)):
        eval_dataset = load_and_cache_examples(args, [eval_task], tokenizer, evaluate=True, data_split=data_split)
        if not os.path.exists(eval_output_dir) and getattr(args, "local_rank", -1) in [-1, 0]:
            os.makedirs(eval_output_dir)
        args.eval_batch_size = args.per_gpu_eval_batch_size * max(1, getattr(args, "n_gpu", 1))
        loader = DataLoader(eval_dataset, sampler=SequentialSampler(eval_dataset), batch_size=args.eval_batch_size)
        logger.info("***** Running evaluation on split: %s %s *****", data_split, prefix)
        truth, predicted, best_acc = [], [], []
        f = open(os.path.join(args.output_dir, "output_order.txt"), "w")
        steps = 0
        model.eval()
        for batch in loader:
            labels = batch[3]
            multiref = labels.ndim > 2
            tru = [(l[0] if multiref else l).tolist() for l in labels]
            with torch.no_grad():
                inputs = {"input_ids": batch[0], "attention_mask": batch[1], "labels": labels[:, 0, :] if multiref else labels}
                if getattr(args, "multimodal", False):
                    inputs["images"] = batch[-1]
                if len(tru[0]) == 1 and not multiref:
                    preds = tru
                else:
                    preds = berson_pointer_network(args, model, tokenizer, inputs)
                    if len(tru) == 1:
                        preds = [preds]
            for p, t in zip(preds, tru):
                truth.append(t)
                predicted.append(p)
                print("{}|||{}".format(" ".join(map(str, p)), " ".join(map(str, t))), file=f)
            steps += 1
            if getattr(args, "max_eval_steps", 0) > 0 and steps >= args.max_eval_steps:
                break
        accs, pmr, taus = cal_result(truth, predicted, best_acc, f, args=args)
        results.update(acc_dev=accs, pmr_dev=pmr, taus_dev=taus)
        out = os.path.join(eval_output_dir, "eval_results_split_{}.txt".format(data_split))
        with open(out, "w") as writer:
            for key in sorted(results.keys()):
                writer.write("%s = %s\n" % (key, str(results[key])))
    return results
