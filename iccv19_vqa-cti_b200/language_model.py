"""Drop-in for ``QuestionEmbedding`` of the reference's ``src/language_model.py`` (:50-98), the step right before the
hot path: the one-layer unidirectional GRU that turns word embeddings into the per-token
question / answer features ``q`` and ``a``.

Same constructor and methods (``forward`` -> last hidden state, ``forward_all`` -> every hidden state, ``init_hidden``);
the parameters live in a submodule called ``rnn`` under nn.GRU's own names (``rnn.weight_ih_l0``, ``rnn.weight_hh_l0``,
``rnn.bias_ih_l0``, ``rnn.bias_hh_l0``), so reference checkpoints load.  Only what the shipped builders construct runs
on the accelerated path -- ``QuestionEmbedding(in_dim, num_hid, 1, False, .0)`` with ``rnn_type='GRU'``
(src/MC/base_model.py:188-191, src/FFOE/base_model.py:177-180); LSTM, stacked or bidirectional variants raise.
``WordEmbedding`` (an embedding lookup) stays the reference's.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import functions as F_


class GRUParams(nn.Module):
    """The parameters of ``nn.GRU(in_dim, num_hid, 1, batch_first=True)`` in its registration order and with its
    initialisation (uniform(-1/sqrt(H), 1/sqrt(H)) for every tensor, drawn in this order)."""

    def __init__(self, in_dim: int, num_hid: int):
        super().__init__()
        self.input_size, self.hidden_size = in_dim, num_hid
        k = 1.0 / math.sqrt(num_hid)
        self.weight_ih_l0 = nn.Parameter(torch.empty(3 * num_hid, in_dim).uniform_(-k, k))
        self.weight_hh_l0 = nn.Parameter(torch.empty(3 * num_hid, num_hid).uniform_(-k, k))
        self.bias_ih_l0 = nn.Parameter(torch.empty(3 * num_hid).uniform_(-k, k))
        self.bias_hh_l0 = nn.Parameter(torch.empty(3 * num_hid).uniform_(-k, k))
        self._pack = None

    def extra_repr(self) -> str:
        return f"{self.input_size}, {self.hidden_size}, batch_first=True"

    def packed(self):
        key = (self.weight_ih_l0._version, self.weight_hh_l0._version, self.weight_ih_l0.data_ptr())
        if self._pack is None or self._pack[0] != key:
            self._pack = (key, F_.gru_pack(self.weight_ih_l0, self.weight_hh_l0))
        return self._pack[1]

    def forward(self, x, hidden=None):
        """(output, h_n) like nn.GRU; ``hidden`` must be None or zeros (all the reference ever passes)."""
        if not x.is_cuda:
            raise RuntimeError("cti_b200 modules run on CUDA tensors only (no CPU fallback)")
        if hidden is not None and (hidden.requires_grad or bool(torch.count_nonzero(hidden).item())):
            raise RuntimeError("the accelerated GRU starts from a zero state (reference src/language_model.py:83-84,"
                               "95-96 always passes init_hidden()); a non-zero or differentiable `hidden` is not supported")
        out = F_.GRUFn.apply(x if x.dtype == torch.float32 else x.float(), self.weight_ih_l0, self.weight_hh_l0,
                             self.bias_ih_l0, self.bias_hh_l0, self.packed())
        return out, out[:, -1].unsqueeze(0)


class QuestionEmbedding(nn.Module):
    def __init__(self, in_dim, num_hid, nlayers, bidirect, dropout, rnn_type='GRU'):
        """Module for question embedding (same signature as reference src/language_model.py:51)."""
        super().__init__()
        assert rnn_type == 'LSTM' or rnn_type == 'GRU'
        if rnn_type != 'GRU' or nlayers != 1 or bidirect:
            raise NotImplementedError("the accelerated QuestionEmbedding covers what the builders construct: a one-layer "
                                      "unidirectional GRU (reference src/MC/base_model.py:188-191)")
        self.rnn = GRUParams(in_dim, num_hid)
        self.in_dim = in_dim
        self.num_hid = num_hid
        self.nlayers = nlayers
        self.rnn_type = rnn_type
        self.ndirections = 1 + int(bidirect)

    def init_hidden(self, batch):
        weight = next(self.parameters()).data
        return weight.new_zeros((self.nlayers * self.ndirections, batch, self.num_hid))

    def forward(self, x):
        """x: [batch, sequence, in_dim] -> last hidden state [batch, num_hid]."""
        output, _ = self.rnn(x)
        return output[:, -1]

    def forward_all(self, x):
        """x: [batch, sequence, in_dim] -> every hidden state [batch, sequence, num_hid]."""
        output, _ = self.rnn(x)
        return output
