"""N>1 host path on CPU: two gloo ranks shard a dataset of manuals, order their shards (with the oracle
standing in for the device path — this test is about the sharding / merge plumbing) and the merged
predictions must equal the single-process result set."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_manuals, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multimodal_sequencing_b200.sharding import merge_predictions, shard_indices
    from oracle import berson_oracle as O
    from oracle import synth
    torch.set_grad_enabled(False)
    torch.set_num_threads(1)
    cfg = dict(hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=256, vocab_size=300,
               max_position_embeddings=64)
    sd = synth.full_state_dict(cfg, None, seed=0, ff=128)
    ids, labels, _ = O.synthetic_manuals(n_manuals, 5, 8, vocab=300, seed=9)
    idx = shard_indices(n_manuals, rank, world)
    ocfg = dict(num_hidden_layers=1, num_attention_heads=2, vit=None)
    local = O.order_manuals(sd, ocfg, ids[idx], labels[idx], 5, 4)
    merged = merge_predictions(torch.tensor(local, dtype=torch.int32).reshape(len(idx), 5), idx, n_manuals)
    torch.save(merged, os.path.join(out_dir, "merged_%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_manuals", [7, 4])
def test_two_rank_sharding_equals_single_process(tmp_path, n_manuals):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_manuals, str(tmp_path)), nprocs=2, join=True)
    from oracle import berson_oracle as O
    from oracle import synth
    torch.set_grad_enabled(False)
    cfg = dict(hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=256, vocab_size=300,
               max_position_embeddings=64)
    sd = synth.full_state_dict(cfg, None, seed=0, ff=128)
    ids, labels, _ = O.synthetic_manuals(n_manuals, 5, 8, vocab=300, seed=9)
    single = O.order_manuals(sd, dict(num_hidden_layers=1, num_attention_heads=2, vit=None), ids, labels, 5, 4)
    for r in range(2):
        merged = torch.load(os.path.join(str(tmp_path), "merged_%d.pt" % r))
        assert merged.tolist() == single


def test_shard_indices_partition():
    from multimodal_sequencing_b200.sharding import shard_indices
    for n in (0, 1, 5, 256):
        for w in (1, 2, 4, 8):
            parts = [shard_indices(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _grad_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multimodal_sequencing_b200.sharding import allreduce_gradients
    from oracle import berson_oracle as O
    from oracle import synth
    from oracle import train_oracle as TO
    torch.set_num_threads(1)
    cfg = dict(hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=256, vocab_size=300,
               max_position_embeddings=64)
    sd = synth.full_state_dict(cfg, None, seed=0, ff=128)
    ids, labels, _ = O.synthetic_manuals(4, 4, 6, vocab=300, seed=5)
    mine = list(range(rank, 4, world))
    ocfg = dict(num_hidden_layers=1, num_attention_heads=2, vit=None)
    _, grads = TO.loss_grads(sd, ocfg, O.prepare_inputs(ids[mine], labels[mine], 4))   # the oracle stands in for the device path
    names = sorted(grads)
    flat = torch.cat([grads[n].reshape(-1) for n in names])
    scale = allreduce_gradients(flat)
    torch.save((names, flat * scale), os.path.join(out_dir, "g_%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch(tmp_path):
    """Replicas + one all-reduce of the flat gradient buffer == the gradient of the whole batch (the loss is a batch mean)."""
    port = _free_port()
    mp.spawn(_grad_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from oracle import berson_oracle as O
    from oracle import synth
    from oracle import train_oracle as TO
    cfg = dict(hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=256, vocab_size=300,
               max_position_embeddings=64)
    sd = synth.full_state_dict(cfg, None, seed=0, ff=128)
    ids, labels, _ = O.synthetic_manuals(4, 4, 6, vocab=300, seed=5)
    _, full = TO.loss_grads(sd, dict(num_hidden_layers=1, num_attention_heads=2, vit=None), O.prepare_inputs(ids, labels, 4))
    for r in range(2):
        names, flat = torch.load(os.path.join(str(tmp_path), "g_%d.pt" % r))
        ref = torch.cat([full[n].reshape(-1) for n in names])
        assert (flat - ref).abs().max() <= 1e-5 * max(1.0, float(ref.abs().max()))
