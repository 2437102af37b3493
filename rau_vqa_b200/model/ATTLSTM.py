"""Drop-in for model/ATTLSTM.lua: ``LSTM.create(input_size, rnn_size, num_layers, dropout)`` returns a module
taking ``{x, prev_c, prev_h}`` and returning ``{next_c, next_h}`` (A:30, A:70-71).  Gate chunks are
(in, in_transform, forget, out) via Reshape(4,H)+SplitTable (A:12-19); Dropout sits on every layer's input,
layer 1 included (A:52); prev_c/prev_h are [B, num_layers*H] narrowed per layer (A:43-44)."""
from ..core import GATES_IGFO
from ._lstm_stack import LSTMStack


class LSTM:
    @staticmethod
    def create(input_size, rnn_size, num_layers, dropout=0.0):
        return LSTMStack(input_size, rnn_size, num_layers, dropout, GATES_IGFO, packed_state=False, dropout_on_first=True)


create = LSTM.create
