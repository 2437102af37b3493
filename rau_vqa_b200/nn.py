"""The slice of the Torch7 ``nn.Module`` protocol the experiment scripts use on the hot-path modules
(SURVEY.md 8b): forward/backward, updateOutput/updateGradInput/accGradParameters, parameters/getParameters,
clone(...)/share(...), training/evaluate, zeroGradParameters.

Host-side mirror of the Lua shims in lua/ (LuaJIT is not available in this image).  The
rules the Lua boundary forces are kept here too: parameters live in fields literally named weight / bias /
gradWeight / gradBias on child modules listed in ``self.modules``; native code receives raw pointers taken from
those tensors at every call (getParameters re-points them); no native handle is stored in a module (the context
is looked up per device), so modules survive ``copy.deepcopy`` the way they survive torch.MemoryFile.
"""
from __future__ import annotations

import copy

import torch

from . import core

_ctx_by_device = {}


def context(device=0) -> core.Context:
    """The per-device native context (file-local upvalue in the Lua shims)."""
    idx = device.index if isinstance(device, torch.device) else int(device)
    idx = 0 if idx is None else idx
    if idx not in _ctx_by_device:
        _ctx_by_device[idx] = core.Context(idx)
    return _ctx_by_device[idx]


class Module:
    def __init__(self):
        self.modules = []
        self.train = True
        self.output = None
        self.gradInput = None

    # --- protocol
    def updateOutput(self, input):
        raise NotImplementedError

    def updateGradInput(self, input, gradOutput):
        raise NotImplementedError

    def accGradParameters(self, input, gradOutput, scale=1.0):
        pass

    def forward(self, input):
        return self.updateOutput(input)

    def backward(self, input, gradOutput, scale=1.0):
        self.updateGradInput(input, gradOutput)
        self.accGradParameters(input, gradOutput, scale)
        return self.gradInput

    def training(self):
        self.train = True
        for m in self.modules:
            m.training()
        return self

    def evaluate(self):
        self.train = False
        for m in self.modules:
            m.evaluate()
        return self

    # --- parameters
    def parameters(self):
        ws, gs = [], []
        for name, gname in (("weight", "gradWeight"), ("bias", "gradBias")):
            if getattr(self, name, None) is not None:
                ws.append(getattr(self, name))
                gs.append(getattr(self, gname))
        for m in self.modules:
            w, g = m.parameters()
            ws += w
            gs += g
        return ws, gs

    def _param_slots(self):
        slots = []
        for name, gname in (("weight", "gradWeight"), ("bias", "gradBias")):
            if getattr(self, name, None) is not None:
                slots.append((self, name, gname))
        for m in self.modules:
            slots += m._param_slots()
        return slots

    def getParameters(self):
        """Flatten every parameter into one storage and re-point the module fields at views of it
        (Module:getParameters, used at F:322-324)."""
        slots = self._param_slots()
        n = sum(getattr(o, w).numel() for o, w, _ in slots)
        dev = getattr(slots[0][0], slots[0][1]).device if slots else "cpu"
        flat = torch.empty(n, dtype=torch.float32, device=dev)
        gflat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for o, w, g in slots:
            t = getattr(o, w)
            k = t.numel()
            flat[off:off + k].copy_(t.reshape(-1))
            gflat[off:off + k].copy_(getattr(o, g).reshape(-1))
            setattr(o, w, flat[off:off + k].view(t.shape))
            setattr(o, g, gflat[off:off + k].view(t.shape))
            off += k
        return flat, gflat

    def zeroGradParameters(self):
        for g in self.parameters()[1]:
            g.zero_()

    def share(self, other, *names):
        """nn.Module:share: self[name]:set(other[name]) for the named fields OF THIS MODULE ONLY.  Stock Torch7 does not
        recurse here -- nn.Container:share does -- so a module that keeps its weights in child modules must be a
        Container, or clone('weight', ...) silently shares nothing (the Lua shims follow the same rule)."""
        for n in names:
            if getattr(other, n, None) is not None and getattr(self, n, None) is not None:
                setattr(self, n, getattr(other, n))
        return self

    def clone(self, *names):
        """Serialise + deserialise, then share the named fields with the prototype (F:339-347)."""
        c = copy.deepcopy(self)
        if names:
            c.share(self, *names)
        return c

    def cuda(self, device=0):
        return self.type_(torch.device("cuda", device))

    def type_(self, device):
        for k, v in list(self.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device))
        for m in self.modules:
            m.type_(device)
        return self


class Container(Module):
    """nn.Container: a module whose parameters live in the child modules of self.modules; share() recurses."""

    def add(self, m):
        self.modules.append(m)
        return self

    def share(self, other, *names):
        for a, b in zip(self.modules, other.modules):
            a.share(b, *names)
        return self


class Linear(Module):
    """Parameter holder with nn.Linear's field names; the arithmetic happens in the owning module's native call."""

    def __init__(self, in_size, out_size, device="cpu", bias=True):
        super().__init__()
        self.weight = torch.zeros(out_size, in_size, dtype=torch.float32, device=device)
        self.gradWeight = torch.zeros_like(self.weight)
        if bias:                       # nn.Linear(i, o, false) has no bias field
            self.bias = torch.zeros(out_size, dtype=torch.float32, device=device)
            self.gradBias = torch.zeros_like(self.bias)


class Add(Module):
    """Parameter holder with nn.Add's field names (a bias without a weight)."""

    def __init__(self, size, device="cpu"):
        super().__init__()
        self.bias = torch.zeros(size, dtype=torch.float32, device=device)
        self.gradBias = torch.zeros_like(self.bias)
