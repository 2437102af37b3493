-- Drop-in for model/DeepLSTM.lua.  Like the reference it defines the GLOBAL table LSTM (D:12) and returns it.
-- LSTM.create(input_size, rnn_size, n, dropout) returns an nn.Module taking {x, h_old} and returning h_new,
-- the state packed [c1|h1|...|cn|hn] (D:23-24, D:68); gate chunks (in, forget, out, transform) (D:47-54);
-- Dropout only on the input of layers 2..n (D:39).  Runs on librau.so.
local S = require 'rau.LSTMStack'
LSTM = {}

function LSTM.create(input_size, rnn_size, n, dropout)
  return nn.RauLSTMStack(input_size, rnn_size, n, dropout or 0, S.GATES_IFOG, true, false)
end

return LSTM
