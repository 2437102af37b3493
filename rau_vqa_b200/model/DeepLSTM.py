"""Drop-in for model/DeepLSTM.lua: ``LSTM.create(input_size, rnn_size, n, dropout)`` returns a module taking
``{x, h_old}`` and returning ``h_new`` (D:14, D:70), the state packed as [c1|h1|c2|h2|...] (D:23-24, D:68).
Gate chunks are (in, forget, out, in_transform) (D:47-54); Dropout only between layers (D:38-39)."""
from ..core import GATES_IFOG
from ._lstm_stack import LSTMStack


class LSTM:
    @staticmethod
    def create(input_size, rnn_size, n, dropout=0.0):
        return LSTMStack(input_size, rnn_size, n, dropout, GATES_IFOG, packed_state=True, dropout_on_first=False)


create = LSTM.create
