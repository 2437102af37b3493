-- Drop-in for model/ATTLSTM.lua of HyeonwooNoh/RAU_VQA (same module path: experiments/*/model is a symlink,
-- gen_simulinks.sh:2-4).  LSTM.create(input_size, rnn_size, num_layers, dropout) returns an nn.Module taking
-- {x, prev_c, prev_h} and returning {next_c, next_h} (A:30, A:70-71); gate chunks are
-- (in, in_transform, forget, out) (A:12-19); Dropout sits on every layer's input (A:52); prev_c / prev_h are
-- [B, num_layers*rnn_size] narrowed per layer (A:43-44).  The cell runs on librau.so (rau_lstm_cell_fwd/bwd).
local S = require 'rau.LSTMStack'
local LSTM = {}

function LSTM.create(input_size, rnn_size, num_layers, dropout)
  return nn.RauLSTMStack(input_size, rnn_size, num_layers, dropout or 0, S.GATES_IGFO, false, true)
end

return LSTM
