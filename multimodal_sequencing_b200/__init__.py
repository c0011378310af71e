"""B200-native (sm_100a) implementation of the multimodal step-ordering hot path of
telin0411/multimodal_sequencing: CLIP ViT pair tower -> joint BERT encoder -> BERSON pooling ->
pointer decoder with batched beam search.  CUDA kernels + C ABI in csrc/, host mirror of the
reference's module interface in engine.py and dropin/."""
from .engine import OrderingEngine, PairBatch, pairs_generator, prepare_pairs  # noqa: F401

__all__ = ["OrderingEngine", "PairBatch", "pairs_generator", "prepare_pairs"]
