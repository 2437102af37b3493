"""Drop-in for utils/optim_updates.lua: sgd, sgdm, sgdmom, adagrad, rmsprop, adam -- in place on flat
float32 CUDA tensors, state kept in a dict exactly as the Lua `state` table (OU:7-87).  Each call is one fused
kernel launch (rau_optim_step) instead of ~9 tensor ops."""
from __future__ import annotations

import torch

from .. import core, nn


def _ctx(x):
    return nn.context(x.device)


def _state(state, key, like):
    if key not in state:
        state[key] = torch.zeros_like(like)
    return state[key]


def sgd(x, dx, lr):
    core.optim_step(_ctx(x), core.OPT_SGD, x, dx, lr)                                         # OU:7-9


def sgdm(x, dx, lr, alpha, state):
    core.optim_step(_ctx(x), core.OPT_SGDM, x, dx, lr, h0=alpha, state0=_state(state, "v", x))   # OU:11-19


def sgdmom(x, dx, lr, alpha, state):
    core.optim_step(_ctx(x), core.OPT_SGDMOM, x, dx, lr, h0=alpha, state0=_state(state, "m", x))  # OU:21-31


def adagrad(x, dx, lr, epsilon, state):
    core.optim_step(_ctx(x), core.OPT_ADAGRAD, x, dx, lr, h0=epsilon, state0=_state(state, "m", x))   # OU:33-43


def rmsprop(x, dx, lr, alpha, epsilon, state):
    core.optim_step(_ctx(x), core.OPT_RMSPROP, x, dx, lr, h0=alpha, h1=epsilon, state0=_state(state, "m", x))   # OU:46-57


def adam(x, dx, lr, beta1=None, beta2=None, epsilon=None, state=None):
    beta1 = 0.9 if beta1 is None else beta1          # OU:60-62
    beta2 = 0.999 if beta2 is None else beta2
    epsilon = 1e-8 if epsilon is None else epsilon
    state = {} if state is None else state
    m, v = _state(state, "m", dx), _state(state, "v", dx)
    state["t"] = state.get("t", 0) + 1                # OU:79
    core.optim_step(_ctx(x), core.OPT_ADAM, x, dx, lr, h0=beta1, h1=beta2, h2=epsilon, state0=m, state1=v, t=state["t"])
    return state
