"""Pairwise-classifier + topological-sort ordering (the reference's `--sort_method topological` eval branch):
trainers/eval.py::topological_inference (425-529), debatch_stories (1088-1097) and trainers/topological_sort.py::Graph.

The reference calls the pairwise classifier once per (i < j) pair of every story — C(N,2) forward passes of batch 1,
each with its own host tokenisation and `.cpu()` read-back.  Here the C(N,2) pair rows of a story are built on the
host exactly as the reference builds them and go through LXRTModel (topo-sort classifier mode, msq_inner_forward +
the classifier head) in ONE call; the tournament is then sorted with the same DFS.

Not shadowing `trainers/`: a maintainer patches the single call site
(`from models.topological import topological_inference`)."""
import numpy as np
import torch


class Graph:
    """trainers/topological_sort.py: DFS post-order, vertices visited in index order, adjacency in insertion order.
    (Iterative; the order of the result is the recursion's.)"""

    def __init__(self, vertices):
        self.V = vertices
        self.graph = {}

    def addEdge(self, u, v):
        self.graph.setdefault(u, []).append(v)

    def _dfs(self, root, visited, order):
        visited[root] = True
        stack = [(root, iter(self.graph.get(root, ())))]
        while stack:
            v, it = stack[-1]
            for w in it:
                if not visited[w]:
                    visited[w] = True
                    stack.append((w, iter(self.graph.get(w, ()))))
                    break
            else:
                stack.pop()
                order.insert(0, v)

    def topologicalSort(self, assert_head=None):
        if assert_head is not None:
            for v in list(self.graph.keys()):
                if v != assert_head and v not in self.graph.setdefault(assert_head, []):
                    self.graph[assert_head].insert(0, v)
        visited = [False] * self.V
        order = []
        for i in range(self.V):
            if not visited[i] and i != assert_head:
                self._dfs(i, visited, order)
        if assert_head is not None:
            if assert_head in order:
                order.remove(assert_head)
            self._dfs(assert_head, visited, order)
            assert order[0] == assert_head, "Asserting head failed"
        return order


def debatch_stories(seqs):
    """seqs[j][b] (collated by the DataLoader) -> stories[b][j] (trainers/eval.py:1088-1097)."""
    return [[seqs[j][b] for j in range(len(seqs))] for b in range(len(seqs[0]))]


def _pair_row(tokenizer, text_a, text_b, args):
    """One (text_a, text_b) classifier input as the reference assembles it (trainers/eval.py:447-478): both steps
    tokenised to per_seq_max_length, pads (id 1) stripped, concatenated with segment ids 0 / 1, re-padded with 1 to
    max_seq_length; attention mask = ids != 1."""
    enc = tokenizer([text_a, text_b], max_length=args.per_seq_max_length, padding="max_length", truncation=True)
    ids = np.asarray(enc["input_ids"])
    cat_ids, cat_tt = [], []
    for k in range(len(ids)):
        unpad = ids[k][ids[k] != 1]
        cat_ids.append(unpad)
        cat_tt.append(np.full(len(unpad), k, dtype=np.int64))
    cat_ids, cat_tt = np.concatenate(cat_ids), np.concatenate(cat_tt)
    n = min(args.max_seq_length, len(cat_ids))
    row = np.ones(args.max_seq_length, dtype=np.int64)
    tt = np.zeros(args.max_seq_length, dtype=np.int64)
    row[:n], tt[:n] = cat_ids[:n], cat_tt[:n]
    return row, tt


def topological_inference(args, model, seqs, tokenizer, images=None, batch=None):
    """Same arguments and return value as the reference: (list of predicted orders, loss).  `model` is the LXRTModel
    mirror built with num_labels (classifier mode)."""
    clip_mm = getattr(args, "multimodal", False) and getattr(args, "multimodal_model_type", None) == "clip"
    if not clip_mm or getattr(args, "multimodal_text_part", False):
        raise NotImplementedError("only the CLIP multimodal pairwise classifier (LXRTModel, num_labels=...) is on this path")
    if getattr(args, "include_num_img_regional_features", None) is not None:
        raise NotImplementedError("regional image features are outside the scoped path")
    stories = debatch_stories(seqs)
    n = len(seqs)
    pairs = [(i, j) for i in range(n) for j in range(n) if i < j]
    device = next(model.parameters()).device
    preds = []
    for b, story in enumerate(stories):
        rows = [_pair_row(tokenizer, story[i], story[j], args) for i, j in pairs]
        ids = torch.from_numpy(np.stack([r[0] for r in rows])).to(device)
        tt = torch.from_numpy(np.stack([r[1] for r in rows])).to(device)
        enc = {"input_ids": ids, "attention_mask": (ids != 1).long(),
               "token_type_ids": tt if getattr(args, "replace_token_type_embeddings", False) else None}
        img = torch.as_tensor(images)[b]
        enc["visual_feats"] = torch.stack([torch.stack([img[i], img[j]]) for i, j in pairs]).to(device)   # [C(n,2), 2, 3, S, S]
        with torch.no_grad():
            logits = model(**enc)[0]
        labels = logits.detach().float().cpu().numpy().argmax(-1)       # ties -> 0 ("unordered"), as np.argmax does
        graph = Graph(n)
        for (i, j), lab in zip(pairs, labels):
            if lab == 1:
                graph.addEdge(i, j)
            else:
                graph.addEdge(j, i)
        preds.append(graph.topologicalSort())
    return preds, 0.0   # the reference accumulates no loss here (loss = 0 / cnt)
