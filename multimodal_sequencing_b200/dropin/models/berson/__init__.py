"""`from models.berson import BertForOrdering, beam_search_pointer, BertConfig` (trainers/train.py:2008-2009)."""
from .modeling_bert import (BertConfig, BertForOrdering, BertModel, HierarchicalAttention, TransformerInterEncoder,  # noqa: F401
                            beam_search_pointer, berson_pointer_network)
from .generator import Beam  # noqa: F401
