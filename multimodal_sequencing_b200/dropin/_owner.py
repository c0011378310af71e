"""How the drop-in nn.Modules keep their packed device model (OrderingEngine) in step with their nn.Parameters.

Version counters are NOT a dirty signal: transformers.AdamW 3.4 (pinned by the reference, trainers/train.py:185) and the
vendored models/berson/optimization.py:176,187 update weights through `p.data.addcdiv_` / `p.data.add_`, which leaves
`p._version` untouched.  So:
  * the engine is BUILT once per (parameter identities, precision) and never rebuilt because values changed;
  * values are pushed with OrderingEngine.refresh_weights (in-place upload + re-pack, a few ms) whenever they may have changed:
    on every forward in train() mode, on the first forward after a train-mode forward (an optimizer step normally follows
    it), whenever a version counter or a storage pointer did move, and after `mark_dirty()`.
An eval-only model therefore uploads its weights exactly once."""
import torch

from ..engine import OrderingEngine


def mark_dirty(module):
    """Tell the drop-in that parameter VALUES changed outside its sight (e.g. `p.data` edits in eval mode)."""
    module.__dict__["_eng_dirty"] = True


def owned_engine(module, make_config, make_state, device, tensors=None):
    """-> the module's OrderingEngine, built on first use and refreshed in place when its weights may have changed.
    module.__dict__["_eng_builds"] / ["_eng_uploads"] count constructions and in-place refreshes (tests read them)."""
    tensors = tensors if tensors is not None else list(module.parameters()) + list(module.buffers())
    precise = getattr(module, "precise", False)
    struct = (tuple(id(t) for t in tensors), tuple(tuple(t.shape) for t in tensors), precise)
    vers = tuple((t._version, t.data_ptr()) for t in tensors)
    d = module.__dict__
    eng = d.get("_eng")
    if eng is None or d.get("_eng_struct") != struct:
        if device.type != "cuda":
            raise RuntimeError("the B200 path has no CPU fallback: move the model to a CUDA device first")
        eng = OrderingEngine(make_state(), make_config(), device=device, precise=precise)
        d["_eng"], d["_eng_struct"], d["_eng_vers"], d["_eng_dirty"] = eng, struct, vers, False
        d["_eng_builds"] = d.get("_eng_builds", 0) + 1
        d["_lib_masters"] = False
    elif not d.get("_lib_masters") and (module.training or d.get("_eng_dirty") or d.get("_eng_vers") != vers):
        # (_lib_masters: the fused finetune_step keeps the fp32 masters inside the library; the nn.Parameters are the stale side)
        eng.refresh_weights(make_state())
        d["_eng_vers"], d["_eng_dirty"] = vers, False
        d["_eng_uploads"] = d.get("_eng_uploads", 0) + 1
    if module.training and not d.get("_lib_masters"):
        d["_eng_dirty"] = True     # an optimizer step normally follows a train-mode forward
    return eng
