"""models/pointer_module.py of the reference, p1 surface: PointerOutput (32-576) with LSTMPointerModule (681-749),
LSTMDecoder (651-678), LSTMAttention (616-648), SimpleClassifier (579-590), GeLU (593-613).

The p0 variant (HF Roberta decoder + `pretrained_models/roberta/*/decoder_config.json`, 46-66) and the auxiliary
objectives need files / packages the reference itself does not ship and raise.  p1 keeps the
reference's parameter names (`lstm_decoder.*`, aliased as `lstm_pointer.decoder.*`) and behaviour — greedy
pointing WITHOUT a permutation mask — and runs as one CUDA kernel (msq_pointer_p1)."""
import ctypes as C
import math

import torch
import torch.nn as nn

from multimodal_sequencing_b200 import _lib
from .beam import Beam  # noqa: F401  (imported by the reference module too)


def gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


class GeLU(nn.Module):
    def forward(self, x):
        return gelu(x)


class SimpleClassifier(nn.Module):
    """pointer_module.py:579-590 (parameter holder for the aux heads; not on the p1 path)."""

    def __init__(self, in_dim, hid_dim, out_dim, dropout):
        super().__init__()
        self.logit_fc = nn.Sequential(nn.Linear(in_dim, hid_dim), GeLU(), nn.LayerNorm(hid_dim, eps=1e-12),
                                      nn.Linear(hid_dim, out_dim))


class LSTMAttention(nn.Module):
    def __init__(self, hidden_size, units):
        super().__init__()
        self.W1 = nn.Linear(hidden_size, units, bias=False)
        self.W2 = nn.Linear(hidden_size, units, bias=False)
        self.V = nn.Linear(units, 1, bias=False)


class LSTMDecoder(nn.Module):
    def __init__(self, hidden_size, attention_units=10):
        super().__init__()
        self.lstm = nn.LSTM(hidden_size * 2, hidden_size, batch_first=True)
        self.attention = LSTMAttention(hidden_size, attention_units)


class LSTMPointerModule(nn.Module):
    """forward(encoder_out [B,N,H], encoder_cls [B,H], y [B,N]) -> (outputs [B,N], batch_loss)  (690-749)."""

    def __init__(self, decoder, beam_size=None):
        super().__init__()
        self.decoder = decoder
        self.beam_size = beam_size  # hard-wired to None by PointerOutput (38): the beam branch never ran

    def forward(self, encoder_out, encoder_cls, y, teacher_force_ratio=.5):
        if not encoder_out.is_cuda:
            raise RuntimeError("the B200 path has no CPU fallback: move the tensors to a CUDA device first")
        lib = _lib.load()
        B, N, H = encoder_out.shape
        d = self.decoder
        U = d.attention.W1.weight.shape[0]
        f = lambda t: t.detach().to(encoder_out.device, torch.float32).contiguous()
        enc, cls, yy = f(encoder_out), f(encoder_cls), y.to(encoder_out.device, torch.long).contiguous()
        w = [f(d.attention.W1.weight), f(d.attention.W2.weight), f(d.attention.V.weight).reshape(-1), f(d.lstm.weight_ih_l0),
             f(d.lstm.weight_hh_l0), f(d.lstm.bias_ih_l0), f(d.lstm.bias_hh_l0)]
        preds = torch.empty(B, N, device=enc.device)
        ce = torch.empty(B, device=enc.device)
        loss = torch.zeros(1, device=enc.device)
        p = lambda t: C.c_void_p(t.data_ptr())
        st = C.c_void_p(torch.cuda.current_stream(enc.device).cuda_stream)
        _lib.check(lib.msq_pointer_p1(p(enc), p(cls), p(yy), *[p(t) for t in w], B, N, H, U, p(preds), p(ce), p(loss), st))
        return preds.type_as(encoder_out), loss[0]


class PointerOutput(nn.Module):
    """pointer_module.py:32-576, `config.hierarchical_version == "p1"`."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        config.beam_size = None
        if config.hierarchical_version != "p1":
            raise NotImplementedError("p0 needs pretrained_models/roberta/*/decoder_config.json, absent from the reference")
        self.lstm_decoder = LSTMDecoder(config.hidden_size, config.max_story_length)
        self.lstm_pointer = LSTMPointerModule(self.lstm_decoder, config.beam_size)
        self.hl_include_objectives = getattr(config, "hl_include_objectives", None)
        if self.hl_include_objectives:
            raise NotImplementedError("auxiliary objectives are outside the scoped path")

    def forward(self, batch, sequence_output, itm_repr=None):
        input_ids = batch["input_ids"]
        bz, text_len = input_ids.size()
        if isinstance(sequence_output, tuple):
            raise NotImplementedError("Not done yet!")  # as the reference (177-179)
        # hidden states at the [CLS] positions of every manual (160-164, 197-200, 300-303)
        pos = (input_ids == self.config.cls_id).nonzero(as_tuple=False)
        n = pos.shape[0] // bz
        cls_pointer = sequence_output[pos[:, 0], pos[:, 1]].reshape(bz, n, -1)
        self.cls_pointer = cls_pointer
        encoder_cls = sequence_output[:, :text_len][:, 0, :]
        labels = batch.get("labels")
        y = labels if labels is not None else torch.zeros(bz, n, dtype=torch.long, device=input_ids.device)
        outputs, loss = self.lstm_pointer(cls_pointer, encoder_cls, y)
        if labels is not None:
            return loss, outputs
        return (outputs,)
