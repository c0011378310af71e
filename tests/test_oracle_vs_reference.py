"""Live pin of the oracle against the REAL reference when /root/reference is present (build
container).  Skipped on the GPU box, where only the committed fixtures exist."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import ref_harness as rh  # noqa: E402
from oracle import berson_oracle as O  # noqa: E402

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference checkout not present")
torch.set_grad_enabled(False)

TINY = dict(vocab_size_or_config_json_file=600, hidden_size=128, num_hidden_layers=1, num_attention_heads=2,
            intermediate_size=256, max_position_embeddings=128)
VIT = dict(embed_dim=64, image_resolution=224, vision_layers=1, vision_width=128, vision_patch_size=32)


@pytest.mark.parametrize("mm,N,W", [(False, 5, 4), (False, 7, 16), (True, 5, 4)])
def test_live(mm, N, W):
    ns = rh.load()
    args = rh.make_args(N, W, multimodal=mm)
    args.ff_size = 128
    model = rh.build_multimodal_model(ns, TINY, args, VIT, seed=5) if mm else rh.build_text_model(ns, TINY, args, seed=5)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    cfg = dict(num_hidden_layers=1, num_attention_heads=2, vit=VIT if mm else None)
    ids, labels, images = O.synthetic_manuals(2, N, 12, vocab=600, image_px=224 if mm else None, seed=77)
    ref = []
    for b in range(2):
        inputs = {"input_ids": ids[b:b + 1], "attention_mask": torch.ones_like(ids[b:b + 1]), "labels": labels[b:b + 1]}
        if mm:
            inputs["images"] = images[b:b + 1]
        ref.append(ns.berson.berson_pointer_network(args, model, rh.StubTokenizer(), inputs))
    assert O.order_manuals(sd, cfg, ids, labels, N, W, images) == ref
