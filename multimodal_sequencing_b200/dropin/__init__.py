"""Drop-in mirror of the reference's module interface for the step-ordering path.

Put this directory FIRST on sys.path (see `install()`), and the reference's own call sites
(trainers/train.py:2006-2037, 2193-2220; models/berson/eval.py:111) resolve to the B200 path:

    from models.berson import BertForOrdering, beam_search_pointer, BertConfig
    from models.berson.modeling_bert import berson_pointer_network
    from models.CLIP.src.lxrt.modeling import LXRTModel
    from models.beam import Beam

Constructors, forward signatures and state_dict keys follow the reference; every forward runs through
the C ABI of libmsq_b200.so (no eager PyTorch math on the hot path)."""
import os
import sys


def install():
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    return here
