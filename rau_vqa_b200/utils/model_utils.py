"""Drop-in for utils/model_utils.lua (char-rnn lineage): clone_list, clone_many_times, combine_all_parameters."""
from __future__ import annotations

import copy

import torch


def clone_list(tensor_list, zero_too=False):
    """MU:4-13: deep copy of a list of tensors, optionally zeroed."""
    out = []
    for t in tensor_list:
        c = t.clone()
        if zero_too:
            c.zero_()
        out.append(c)
    return out


def clone_many_times(net, T):
    """MU:15-36: T copies of `net` whose parameters and gradients alias the prototype's (serialise, then :set)."""
    clones = []
    params, grads = net.parameters() if hasattr(net, "parameters") else ([], [])
    for _ in range(T):
        c = copy.deepcopy(net)
        if params:
            slots, pslots = c._param_slots(), net._param_slots()
            for (o, w, g), (po, pw, pg) in zip(slots, pslots):
                setattr(o, w, getattr(po, pw))      # cloneParams[i]:set(params[i])          MU:29
                setattr(o, g, getattr(po, pg))      # cloneGradParams[i]:set(gradParams[i])  MU:30
        clones.append(c)
    return clones


def combine_all_parameters(*networks):
    """MU:38-137: one flat parameter vector and one flat gradient vector for several networks; tensors that
    already share storage (weight tying) are laid out once."""
    slots = []
    for net in networks:
        slots += net._param_slots()
    seen = {}
    order = []
    for o, w, g in slots:
        t = getattr(o, w)
        key = (t.data_ptr(), t.numel())
        if key not in seen:
            seen[key] = None
            order.append((o, w, g))
    n = sum(getattr(o, w).numel() for o, w, _ in order)
    dev = getattr(order[0][0], order[0][1]).device if order else "cpu"
    flat = torch.empty(n, dtype=torch.float32, device=dev)
    gflat = torch.zeros(n, dtype=torch.float32, device=dev)
    off = 0
    for o, w, g in order:
        t = getattr(o, w)
        k = t.numel()
        flat[off:off + k].copy_(t.reshape(-1))
        gflat[off:off + k].copy_(getattr(o, g).reshape(-1))
        seen[(t.data_ptr(), k)] = (flat[off:off + k].view(t.shape), gflat[off:off + k].view(t.shape))
        off += k
    for o, w, g in slots:
        t = getattr(o, w)
        key = (t.data_ptr(), t.numel())
        if key in seen and seen[key] is not None:
            nw, ng = seen[key]
            setattr(o, w, nw)
            setattr(o, g, ng)
    return flat, gflat
