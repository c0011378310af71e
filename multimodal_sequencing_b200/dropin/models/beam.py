"""Host-side stand-in for the reference's beam container (models/beam.py:8-38; live twin models/berson/generator.py:15-38):
same attributes (`beam_size`, `candidates`, `scores`) and the same `step(prob, prev_beam, f_done) -> (done, remain)`
contract, so code written against the reference's step-by-step loop keeps working.

The product path never calls it: top-k, expansion and the permutation-mask update run inside the decode kernels
(csrc/decode.cu).  The selection rule here is the device's: ascending total cost, ties to the lowest flat index
`beam * n_tokens + token` (a stable sort), integer index arithmetic throughout (the reference's `/` breaks on torch >= 1.5,
SURVEY §0.4)."""
import torch


class Beam(object):
    def __init__(self, beam_size):
        self.beam_size = beam_size
        self.candidates = []   # token sequences of the hypotheses kept by step()
        self.scores = []       # their accumulated costs

    def step(self, prob, prev_beam, f_done):
        """prob [live beams, tokens]: cost of appending each token; prev_beam: the Beam of the previous step.
        Returns (finished [[sequence, cost], ...], parents of the hypotheses that stay alive)."""
        n_tokens = prob.size(1)
        totals = (prob + prob.new_tensor(prev_beam.scores)[:, None]).reshape(-1)
        keep = min(self.beam_size, totals.numel())
        order = torch.sort(totals, stable=True).indices[:keep].tolist()
        finished, parents = [], []
        for flat in order:
            parent, token = divmod(flat, n_tokens)
            sequence = prev_beam.candidates[parent] + [token]
            cost = float(totals[flat])
            if f_done(sequence):
                finished.append([sequence, cost])
                continue
            parents.append(parent)
            self.candidates.append(sequence)
            self.scores.append(cost)
        return finished, parents
