"""Parameter holder for the bidirectional GRU (reference: src/models/RNN.py:7-16, nn.GRU names
`rnn.weight_ih_l0`, `rnn.weight_hh_l0_reverse`, ...).  The recurrence runs in csrc/gru.cu."""
from .CNN import _Holder


class BidirectionalGRU(_Holder):
    def __init__(self, n_in, n_hidden, dropout=0, num_layers=1):
        super().__init__()
        self.n_in, self.n_hidden, self.num_layers = n_in, n_hidden, num_layers
        self.rnn = _Holder()
