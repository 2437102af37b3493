"""rau_vqa_b200 -- B200-native implementation of the RAU_VQA recurrent-answering-unit hot path.

Host-side mirror of the reference's module surface (the Lua shims under lua/ bind the same C ABI through LuaJIT
FFI): model.ATTLSTM, model.DeepLSTM, model.RAU, utils.model_utils, utils.optim_updates, plus `core` for the fused
training step.  The product path is librau.so (CUDA, sm_100a); nothing here falls back to the CPU."""
from .core import (Context, RauConfig, StepBuffers, draw_masks, feval, noise_clip, optim_step, predict, predict_answers,  # noqa: F401
                   train_step)
from ._ffi import RauError  # noqa: F401

__version__ = "0.1.0"
