"""CPU-side checks: host pair expansion against the reference-generated fixtures, and that the C-ABI
library loads and exports every symbol include/msq_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

from multimodal_sequencing_b200 import _lib, prepare_pairs
from oracle import berson_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_prepare_pairs_matches_reference_fixture(golden_dir):
    g = torch.load(os.path.join(golden_dir, "text_tiny.pt"), weights_only=False)
    for c in g["cases"]:
        pb = prepare_pairs(c["ids"], c["labels"], c["N"])
        ref = c["prep"]
        d = pb.reference_dict()
        for k in ("input_ids", "attention_mask", "token_type_ids", "pairs_list", "passage_length", "pairs_num",
                  "sep_positions", "ground_truth", "mask_cls", "pairwise_labels"):
            assert torch.equal(d[k], ref[k]), (k, c["kind"])


def test_prepare_pairs_batch_ragged_and_images():
    ids1, lab1, img = O.synthetic_manuals(3, 5, 12, vocab=500, image_px=32, seed=3)
    pb = prepare_pairs(ids1, lab1, 5, img)
    ref = O.prepare_inputs(ids1, lab1, 5, img)
    d = pb.reference_dict()
    for k in ("input_ids", "attention_mask", "token_type_ids", "sep_positions", "pairwise_labels", "images"):
        assert torch.equal(d[k], ref[k]), k
    assert pb.images.shape == (15, 3, 32, 32) and pb.img_index.dtype == torch.int32


def test_prepare_pairs_rejects_wrong_step_count():
    ids, lab, _ = O.synthetic_manuals(1, 5, 12, vocab=500)
    with pytest.raises(AssertionError):
        prepare_pairs(ids, lab[:, :4], 4)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "msq_b200.h")).read()
    declared = set(re.findall(r"\b(msq_[a-z0-9_]+)\s*\(", header))
    declared -= {"msq_model", "msq_config", "msq_encode_out"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert isinstance(getattr(lib, name), ctypes._CFuncPtr)
    assert lib.msq_version() == 100


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from multimodal_sequencing_b200 import OrderingEngine
    with pytest.raises(RuntimeError):
        OrderingEngine({}, dict(hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=256,
                                vocab_size=10, max_position_embeddings=16))
    lib = _lib.load()
    cfg = _lib.MsqConfig(hidden=128, layers=1, heads=2, inter=256, vocab=10, max_pos=16, type_vocab=2, para_heads=8,
                         para_ff=256, para_layers=2)
    h = ctypes.c_void_p()
    assert lib.msq_model_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"no CUDA device" in lib.msq_last_error()
