"""CPU oracle for the RAU_VQA recurrent-answering-unit hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (rau_vqa_b200/, librau.so) may
import or call this file; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs use it, and only as the checker / CPU arm.

PARITY UNPINNED: the reference (HyeonwooNoh/RAU_VQA) ships no tests, golden vectors or
fixtures for this path, and its arithmetic lives in un-vendored, un-pinned Torch7 rocks
(torch7/nn/nngraph, README.md:34-40) that cannot run in this image (no LuaJIT).  This
file restates the algorithm from the reference's own Lua sources, module by module with
explicit backward passes (the updateGradInput/accGradParameters of each nn module), in
float64 numpy (the reference's CPU mode is float64).  It is cross-checked against an
independent derivation (torch.autograd over torch.nn modules mirroring the nngraph,
oracle/torch_graph.py) and finite differences in tests/test_oracle.py.

Citations are into /root/reference:
  F:  experiments/Ours_Full/LstmAttCtrlGradNoiseDontSelect.lua
  A:  model/ATTLSTM.lua      D:  model/DeepLSTM.lua      OU: utils/optim_updates.lua

Parameter naming (role -> reference constructor line):
  embed group : E[V,200]                                   F:204
  rnn group   : l{1,2}.Wi[4H,in] l.bi[4H] l.Wh[4H,H] l.bh   D:43-44   gate order (i,f,o,g)
  mult group  : Wq,bq F:233 | Wh,bh F:234 | Wi,bi F:240 | Wqa,bqa F:246 | Wa,ba F:247 |
                ws F:251 | Wm,bm F:287 | Wp,bp F:271 | Wx,bx A:6 | Whh,bhh A:7 |
                Wo,bo F:279 | Ws,bso F:280 | wd F:281 | bs F:251 | bd F:281   ATTLSTM gate order (i,g,f,o)
The flat order inside each group is the order listed above (our own layout; nngraph's
traversal order is not recoverable from the reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# ----------------------------------------------------------------------------- config


@dataclass
class RauConfig:
    """Constants hard-coded in the experiment scripts (F:202-229) plus run-time sizes."""
    V: int = 16384          # vocab size (vqa_data.vocab_size, F:204)
    embed: int = 200        # F:202
    Hq: int = 512           # rnn_size F:208
    nlayer: int = 2         # F:209
    C: int = 512            # cnnout_dim F:216 (2048 for Ours_ResNet, RN:217)
    S: int = 196            # 14x14 grid (F:217-219)
    M: int = 512            # multfeat_dim F:220
    A: int = 256            # attfeat_dim F:221
    H: int = 512            # att_rnn_size F:225
    N: int = 2000           # netout_dim = answer_size F:222
    nHop: int = 8           # F:53
    T: int = 26             # seq_len (LD:1418)
    p_embed: float = 0.5    # F:205
    p_rnn: float = 0.5      # F:210
    p_q: float = 0.5        # F:233
    p_x: float = 0.5        # F:239
    p_m: float = 0.5        # F:277
    grad_clip: float = 0.1  # F:49
    noisy_eta: float = 0.01     # F:54
    noisy_gamma: float = 0.55   # F:55
    lr: float = 3e-3        # F:43
    mult_lr: float = 3e-4   # F:45

    @property
    def Q(self) -> int:
        return 2 * self.Hq * self.nlayer  # rnnout_dim F:211


# --------------------------------------------------------------------- parameter layout

def rnn_param_shapes(cfg: RauConfig):
    out = []
    for L in range(1, cfg.nlayer + 1):
        in_sz = cfg.embed if L == 1 else cfg.Hq
        out += [(f"l{L}.Wi", (4 * cfg.Hq, in_sz)), (f"l{L}.bi", (4 * cfg.Hq,)),
                (f"l{L}.Wh", (4 * cfg.Hq, cfg.Hq)), (f"l{L}.bh", (4 * cfg.Hq,))]
    return out


def mult_param_shapes(cfg: RauConfig):
    return [
        ("Wq", (cfg.M, cfg.Q)), ("bq", (cfg.M,)),
        ("Wh", (cfg.M, cfg.H)), ("bh", (cfg.M,)),
        ("Wi", (cfg.M, cfg.C)), ("bi", (cfg.M,)),
        ("Wqa", (cfg.A, cfg.M)), ("bqa", (cfg.A,)),
        ("Wa", (cfg.A, cfg.M)), ("ba", (cfg.A,)),
        ("ws", (1, cfg.A)),
        ("Wm", (cfg.S, cfg.H)), ("bm", (cfg.S,)),
        ("Wp", (cfg.M, cfg.S)), ("bp", (cfg.M,)),
        ("Wx", (4 * cfg.H, cfg.M)), ("bx", (4 * cfg.H,)),
        ("Whh", (4 * cfg.H, cfg.H)), ("bhh", (4 * cfg.H,)),
        ("Wo", (cfg.M, cfg.H)), ("bo", (cfg.M,)),
        ("Ws", (cfg.N, cfg.M)), ("bso", (cfg.N,)),
        ("wd", (1, cfg.M)),
        ("bs", (1,)), ("bd", (1,)),      # the two scalars last: everything above stays 16-byte aligned in the flat buffer
    ]


def embed_param_shapes(cfg: RauConfig):
    return [("E", (cfg.V, cfg.embed))]


GROUPS = ("embed", "rnn", "mult")


def group_shapes(cfg: RauConfig, group: str):
    return {"embed": embed_param_shapes, "rnn": rnn_param_shapes, "mult": mult_param_shapes}[group](cfg)


def group_size(cfg: RauConfig, group: str) -> int:
    return sum(int(np.prod(s)) for _, s in group_shapes(cfg, group))


def views(cfg: RauConfig, group: str, flat: np.ndarray) -> dict:
    """Named views into a flat group buffer (what getParameters() gives the script, F:322-324)."""
    out, off = {}, 0
    for name, shp in group_shapes(cfg, group):
        n = int(np.prod(shp))
        out[name] = flat[off:off + n].reshape(shp)
        off += n
    assert off == flat.size, (off, flat.size)
    return out


def init_params(cfg: RauConfig, seed: int = 123, dtype=np.float64) -> dict:
    """params:uniform(-0.08, 0.08) on the three flat buffers (F:352-354)."""
    rng = np.random.default_rng(seed)
    return {g: rng.uniform(-0.08, 0.08, group_size(cfg, g)).astype(dtype) for g in GROUPS}


# --------------------------------------------------------------------------- primitives

def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def dropout_fwd(x, mask, p):
    """nn.Dropout v2 (train): x * Bernoulli(1-p) / (1-p); mask is the 0/1 keep tensor; None = eval."""
    if mask is None or p <= 0.0:
        return x
    return x * mask * (1.0 / (1.0 - p))


dropout_bwd = dropout_fwd  # gradient flows through the same scaled mask


def lstm_gates_fwd(G, c_prev, order):
    """Pointwise LSTM cell on pre-activations G[B,4H].
    order 'ifog' -> DeepLSTM chunks [i|f|o|g] (D:47-54); 'igfo' -> ATTLSTM chunks [i|g|f|o] (A:12-19)."""
    H = G.shape[1] // 4
    ch = [G[:, k * H:(k + 1) * H] for k in range(4)]
    if order == "ifog":
        i, f, o, g = sigmoid(ch[0]), sigmoid(ch[1]), sigmoid(ch[2]), np.tanh(ch[3])
    elif order == "igfo":
        i, g, f, o = sigmoid(ch[0]), np.tanh(ch[1]), sigmoid(ch[2]), sigmoid(ch[3])
    else:
        raise ValueError(order)
    c = f * c_prev + i * g          # D:56-59 / A:21-24
    tc = np.tanh(c)
    h = o * tc                      # D:61 / A:25
    return c, h, (i, f, o, g, tc)


def lstm_gates_bwd(dc_out, dh_out, c_prev, saved, order):
    """Backward of lstm_gates_fwd; returns dG[B,4H] in the module's chunk order and dc_prev."""
    i, f, o, g, tc = saved
    do = dh_out * tc
    dc = dc_out + dh_out * o * (1.0 - tc * tc)
    df = dc * c_prev
    dc_prev = dc * f
    di = dc * g
    dg = dc * i
    dGi = di * i * (1.0 - i)
    dGf = df * f * (1.0 - f)
    dGo = do * o * (1.0 - o)
    dGg = dg * (1.0 - g * g)
    if order == "ifog":
        dG = np.concatenate([dGi, dGf, dGo, dGg], axis=1)
    else:
        dG = np.concatenate([dGi, dGg, dGf, dGo], axis=1)
    return dG, dc_prev


# ------------------------------------------------------------------ word embedding (a3)

def word_embed_fwd(E, x_t, mask, p):
    """protos.word_embed = LookupTable -> Dropout(0.5) -> Tanh (F:203-206). x_t is 1-based."""
    idx = np.asarray(x_t).astype(np.int64) - 1
    raw = E[idx]
    e = np.tanh(dropout_fwd(raw, mask, p))
    return e, (idx, e)


def word_embed_bwd(gE, cache, mask, p, de):
    idx, e = cache
    d = dropout_bwd(de * (1.0 - e * e), mask, p)
    np.add.at(gE, idx, d)          # LookupTable accGradParameters scatter-add; no gradInput


# ---------------------------------------------------------------- DeepLSTM one step (a1)

def deeplstm_fwd(P, cfg, x, s_prev, masks):
    """model/DeepLSTM.lua:14-71. s = [c1|h1|c2|h2]; dropout only on the input of layers >= 2 (D:39).
    masks: list (len nlayer-1) of keep masks [B,Hq] for layers 2.. or None (eval)."""
    H = cfg.Hq
    outs, cache, u = [], [], x
    for L in range(1, cfg.nlayer + 1):
        c_prev = s_prev[:, 2 * (L - 1) * H: 2 * (L - 1) * H + H]        # D:23
        h_prev = s_prev[:, 2 * (L - 1) * H + H: 2 * L * H]              # D:24
        if L > 1:
            mk = None if masks is None else masks[L - 2]
            u = dropout_fwd(outs[-1], mk, cfg.p_rnn)                    # D:38-39
        G = u @ P[f"l{L}.Wi"].T + P[f"l{L}.bi"] + h_prev @ P[f"l{L}.Wh"].T + P[f"l{L}.bh"]  # D:43-45
        c, h, saved = lstm_gates_fwd(G, c_prev, "ifog")
        cache.append((u, c_prev, h_prev, saved))
        outs += [c, h]
    return np.concatenate(outs, axis=1), cache                         # D:68


def deeplstm_bwd(P, gP, cfg, cache, masks, ds_new):
    """Returns (dx, ds_prev); accumulates into gP (accGradParameters adds, shared across clones)."""
    H = cfg.Hq
    B = ds_new.shape[0]
    ds_prev = np.zeros((B, 2 * H * cfg.nlayer), dtype=ds_new.dtype)
    dh_from_above = None
    dx = None
    for L in range(cfg.nlayer, 0, -1):
        u, c_prev, h_prev, saved = cache[L - 1]
        dc_out = ds_new[:, 2 * (L - 1) * H: 2 * (L - 1) * H + H]
        dh_out = ds_new[:, 2 * (L - 1) * H + H: 2 * L * H].copy()
        if dh_from_above is not None:
            dh_out = dh_out + dh_from_above          # next_h feeds both the output and layer L+1
        dG, dc_prev = lstm_gates_bwd(dc_out, dh_out, c_prev, saved, "ifog")
        gP[f"l{L}.Wi"] += dG.T @ u
        gP[f"l{L}.bi"] += dG.sum(0)
        gP[f"l{L}.Wh"] += dG.T @ h_prev
        gP[f"l{L}.bh"] += dG.sum(0)
        du = dG @ P[f"l{L}.Wi"]
        dh_prev = dG @ P[f"l{L}.Wh"]
        ds_prev[:, 2 * (L - 1) * H: 2 * (L - 1) * H + H] = dc_prev
        ds_prev[:, 2 * (L - 1) * H + H: 2 * L * H] = dh_prev
        if L > 1:
            mk = None if masks is None else masks[L - 2]
            dh_from_above = dropout_bwd(du, mk, cfg.p_rnn)
        else:
            dx = du
    return dx, ds_prev


# ------------------------------------------------------------------ ATTLSTM one step (a2)

def attlstm_fwd(Wx, bx, Whh, bhh, x, c_prev, h_prev):
    """model/ATTLSTM.lua:4-28 with num_layers=1, dropout=0 (F:225-229): Dropout(0) is identity (A:52)."""
    G = x @ Wx.T + bx + h_prev @ Whh.T + bhh                           # A:6-8
    c, h, saved = lstm_gates_fwd(G, c_prev, "igfo")
    return c, h, (x, c_prev, h_prev, saved)


def attlstm_bwd(Wx, Whh, g, cache, dc_out, dh_out):
    """g: dict with gWx,gbx,gWhh,gbhh arrays to accumulate into. Returns (dx, dc_prev, dh_prev)."""
    x, c_prev, h_prev, saved = cache
    dG, dc_prev = lstm_gates_bwd(dc_out, dh_out, c_prev, saved, "igfo")
    g["Wx"] += dG.T @ x
    g["bx"] += dG.sum(0)
    g["Whh"] += dG.T @ h_prev
    g["bhh"] += dG.sum(0)
    return dG @ Wx, dc_prev, dG @ Whh


# ------------------------------------------------------------------------ RAU hop (a5-a11)

def softmax_rows(z):
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True)


def hop_fwd(P, cfg, q, X, c, h, masks):
    """protos.multimodal forward (F:292-307): {q[B,Q], X[B,C,S], c[B,H], h[B,H]} ->
    {score[B,N], do_pred[B], p[B,S], c'[B,H], h'[B,H]}.
    masks: dict(q=[B,Q], X=[B,C,S], m=[B,M]) of 0/1 keep masks, or None for evaluate()."""
    mq = mX = mm = None
    if masks is not None:
        mq, mX, mm = masks.get("q"), masks.get("X"), masks.get("m")
    B = q.shape[0]
    qd = dropout_fwd(q, mq, cfg.p_q)                                   # F:233
    qf = np.tanh(qd @ P["Wq"].T + P["bq"] + h @ P["Wh"].T + P["bh"])   # F:233-235
    Xd = dropout_fwd(X.reshape(B, cfg.C, cfg.S), mX, cfg.p_x)          # F:239
    I = np.tanh(np.matmul(P["Wi"], Xd) + P["bi"][None, :, None])   # F:240-242
    qatt = qf @ P["Wqa"].T + P["bqa"]                                  # F:246
    Z = np.matmul(P["Wa"], I) + P["ba"][None, :, None]  # F:247-249
    E = np.tanh(Z + qatt[:, :, None])                                  # F:250
    s = np.matmul(P["ws"][0], E) + P["bs"][0]             # F:251
    mem = h @ P["Wm"].T + P["bm"]                                      # F:287
    p = softmax_rows(s + mem)                                          # F:288-289
    a = np.matmul(I, p[:, :, None])[:, :, 0]                                  # F:254-263
    fp = p @ P["Wp"].T + P["bp"]                                       # F:271
    j = qf + a + fp                                                    # F:270,272
    c2, h2, lcache = attlstm_fwd(P["Wx"], P["bx"], P["Whh"], P["bhh"], j, c, h)   # F:273
    pre_m = j + h2 @ P["Wo"].T + P["bo"]                               # F:277-279 (Dropout(0.0) on h' = id)
    m = dropout_fwd(pre_m, mm, cfg.p_m)
    score = m @ P["Ws"].T + P["bso"]                                   # F:280
    do_pred = sigmoid(m @ P["wd"][0] + P["bd"][0])                     # F:281 (Sum(2) squeezes)
    cache = dict(q=q, qd=qd, qf=qf, Xd=Xd, I=I, E=E, p=p, h=h, j=j, h2=h2, m=m,
                 do_pred=do_pred, lcache=lcache, masks=(mq, mX, mm))
    return score, do_pred, p, c2, h2, cache


def hop_bwd(P, gP, cfg, cache, dscore, ddo_pred, dp_att, dc_next, dh_next, want_dX=False):
    """multimodals[h]:backward (F:590-593). Upstream grads {dscore, ddo_pred, dp_att, dc', dh'}.
    Returns (dq, dX or None, dc_prev, dh_prev); accumulates parameter grads into gP."""
    mq, mX, mm = cache["masks"]
    qd, qf, Xd, I, E, p, h, j, h2, m = (cache[k] for k in ("qd", "qf", "Xd", "I", "E", "p", "h", "j", "h2", "m"))
    # heads
    dm = dscore @ P["Ws"]
    gP["Ws"] += dscore.T @ m
    gP["bso"] += dscore.sum(0)
    dlogit_d = ddo_pred * cache["do_pred"] * (1.0 - cache["do_pred"])
    dm = dm + dlogit_d[:, None] * P["wd"][0][None, :]
    gP["wd"] += (dlogit_d[:, None] * m).sum(0)[None, :]
    gP["bd"] += dlogit_d.sum()
    du = dropout_bwd(dm, mm, cfg.p_m)
    dj = du.copy()
    dh2 = dh_next + du @ P["Wo"]
    gP["Wo"] += du.T @ h2
    gP["bo"] += du.sum(0)
    # attlstm
    g_l = dict(Wx=gP["Wx"], bx=gP["bx"], Whh=gP["Whh"], bhh=gP["bhh"])
    dx_l, dc_prev, dh_prev = attlstm_bwd(P["Wx"], P["Whh"], g_l, cache["lcache"], dc_next, dh2)
    dj = dj + dx_l
    # join: j = qf + a + Wp p + bp
    dqf = dj.copy()
    da = dj
    dp = dp_att + dj @ P["Wp"]
    gP["Wp"] += dj.T @ p
    gP["bp"] += dj.sum(0)
    # attselect: a = I p
    dI = da[:, :, None] * p[:, None, :]
    dp = dp + np.matmul(da[:, None, :], I)[:, 0, :]
    # softmax
    ds = p * (dp - (p * dp).sum(axis=1, keepdims=True))
    dh_prev = dh_prev + ds @ P["Wm"]
    gP["Wm"] += ds.T @ h
    gP["bm"] += ds.sum(0)
    # score conv: s = ws.E + bs
    dE = P["ws"][0][None, :, None] * ds[:, None, :]
    gP["ws"] += np.matmul(E, ds[:, :, None])[:, :, 0].sum(0)[None, :]
    gP["bs"] += ds.sum()
    dZ = dE * (1.0 - E * E)
    dI = dI + np.matmul(P["Wa"].T, dZ)
    gP["Wa"] += np.tensordot(dZ, I, axes=([0, 2], [0, 2]))
    gP["ba"] += dZ.sum(axis=(0, 2))
    dqa = dZ.sum(axis=2)
    dqf = dqf + dqa @ P["Wqa"]
    gP["Wqa"] += dqa.T @ qf
    gP["bqa"] += dqa.sum(0)
    # i_embed
    dY = dI * (1.0 - I * I)
    gP["Wi"] += np.tensordot(dY, Xd, axes=([0, 2], [0, 2]))
    gP["bi"] += dY.sum(axis=(0, 2))
    dX = None
    if want_dX:   # computed by the reference then discarded (F:598)
        dX = dropout_bwd(np.matmul(P["Wi"].T, dY), mX, cfg.p_x)
    # q_embed
    dpre = dqf * (1.0 - qf * qf)
    dq = dropout_bwd(dpre @ P["Wq"], mq, cfg.p_q)
    dh_prev = dh_prev + dpre @ P["Wh"]
    gP["Wq"] += dpre.T @ qd
    gP["bq"] += dpre.sum(0)
    gP["Wh"] += dpre.T @ h
    gP["bh"] += dpre.sum(0)
    return dq, dX, dc_prev, dh_prev


# ------------------------------------------------------------------------- criteria (a12)

def cross_entropy_fwd(score, y):
    """nn.CrossEntropyCriterion = LogSoftMax + ClassNLLCriterion(sizeAverage) (F:310); y 1-based."""
    B = score.shape[0]
    z = score - score.max(axis=1, keepdims=True)
    lse = np.log(np.exp(z).sum(axis=1))
    idx = np.asarray(y).astype(np.int64) - 1
    return float((lse - z[np.arange(B), idx]).mean())


def cross_entropy_bwd(score, y):
    B = score.shape[0]
    d = softmax_rows(score)
    d[np.arange(B), np.asarray(y).astype(np.int64) - 1] -= 1.0
    return d / B


def bce_fwd(x, t):
    """nn.BCECriterion: mean(-(t log(x+eps) + (1-t) log(1-x+eps))), eps = 1e-12 (F:311)."""
    eps = 1e-12
    return float(-(t * np.log(x + eps) + (1.0 - t) * np.log(1.0 - x + eps)).mean())


def argmax1(score):
    """torch.max(score, 2) index, 1-based; ties -> lowest index (Torch's tie order is unspecified)."""
    return score.argmax(axis=1) + 1


# ---------------------------------------------------------------- encoder unroll (a4)

def encoder_fwd(Pe, Pr, cfg, x, x_len, masks):
    """F:460-479. x[T,B] 1-based tokens (pad = 1), x_len[B]. masks: dict(embed=[T,B,200], rnn=[T,B,Hq])
    or None. Runs t = 1..max_len like the reference; rnn_out[k] = state_t[k] at t == x_len[k]."""
    T, B = x.shape
    max_len = int(np.max(x_len))
    state = np.zeros((B, cfg.Q))
    rnn_out = np.zeros((B, cfg.Q))
    caches = []
    for t in range(max_len):
        me = None if masks is None else masks["embed"][t]
        mr = None if masks is None else [masks["rnn"][t]]
        e, ecache = word_embed_fwd(Pe["E"], x[t], me, cfg.p_embed)     # F:468
        state, lcache = deeplstm_fwd(Pr, cfg, e, state, mr)            # F:469
        sel = (np.asarray(x_len) == t + 1)                             # F:472-478
        rnn_out[sel] = state[sel]
        caches.append((ecache, lcache, me, mr))
    return rnn_out, caches


def encoder_bwd(Pr, gPe, gPr, cfg, x_len, caches, dq):
    """F:600-615. dq[B,Q] = summed gradient of rnn_out over hops (branch:backward, F:598)."""
    B = dq.shape[0]
    dstate = np.zeros((B, cfg.Q))
    for t in range(len(caches) - 1, -1, -1):
        ecache, lcache, me, mr = caches[t]
        drnn_out = dstate.copy()                                       # F:603
        sel = (np.asarray(x_len) == t + 1)
        drnn_out[sel] = dq[sel]                                        # F:604-610 (replacement)
        dwe, dstate = deeplstm_bwd(Pr, gPr, cfg, lcache, mr, drnn_out)  # F:611
        word_embed_bwd(gPe["E"], ecache, me, cfg.p_embed, dwe)         # F:612


# ----------------------------------------------------------- optimizers (a14) and a13

def adam(x, dx, lr, state, beta1=0.9, beta2=0.999, eps=1e-8):
    """utils/optim_updates.lua:59-87 (epsilon added after sqrt, OU:78). In place on x."""
    if "m" not in state:
        state["t"] = 0
        state["m"] = np.zeros_like(dx)
        state["v"] = np.zeros_like(dx)
    state["m"] *= beta1
    state["m"] += (1 - beta1) * dx
    state["v"] *= beta2
    state["v"] += (1 - beta2) * dx * dx
    tmp = np.sqrt(state["v"]) + eps
    state["t"] += 1
    bc1 = 1 - beta1 ** state["t"]
    bc2 = 1 - beta2 ** state["t"]
    step = lr * math.sqrt(bc2) / bc1
    x -= step * state["m"] / tmp


def rmsprop(x, dx, lr, alpha, eps, state):
    """utils/optim_updates.lua:46-57."""
    if "m" not in state:
        state["m"] = np.zeros_like(x)
    state["m"] *= alpha
    state["m"] += (1.0 - alpha) * dx * dx
    x -= lr * dx / (np.sqrt(state["m"]) + eps)


def sgd(x, dx, lr):
    """OU:7-9."""
    x -= lr * dx


def sgdm(x, dx, lr, alpha, state):
    """OU:11-19."""
    if "v" not in state:
        state["v"] = np.zeros_like(x)
    state["v"] *= alpha
    state["v"] += lr * dx
    x -= state["v"]


def sgdmom(x, dx, lr, alpha, state):
    """OU:21-31 (nesterov form)."""
    if "m" not in state:
        state["m"] = np.zeros_like(x)
    tmp = state["m"].copy()
    state["m"] *= alpha
    state["m"] -= lr * dx
    x -= alpha * tmp
    x += (1 + alpha) * state["m"]


def adagrad(x, dx, lr, eps, state):
    """OU:33-43."""
    if "m" not in state:
        state["m"] = np.zeros_like(x)
    state["m"] += dx * dx
    x -= lr * dx / (np.sqrt(state["m"]) + eps)


def noise_std(cfg: RauConfig, step_t: int) -> float:
    """F:617-618: var = eta / ((step_t+1) * gamma)  (a product, not a power)."""
    return math.sqrt(cfg.noisy_eta / ((step_t + 1) * cfg.noisy_gamma))


def noise_and_clip(cfg, grad, noise):
    """F:619-648 for one group: g += noise; if ||g|| > clip: g *= clip/||g||. Returns pre-clip norm."""
    if noise is not None:
        grad += noise
    n = float(np.linalg.norm(grad))
    if n > cfg.grad_clip:
        grad *= cfg.grad_clip / n
    return n


# --------------------------------------------------------------------- feval (a6, a11-13)

@dataclass
class StepResult:
    loss: np.ndarray            # tab_loss[1..nHop+2]   (F:535, F:548, F:557)
    loss_do_pred: np.ndarray    # tab_loss_do_pred[1..nHop] (F:572)
    grads: dict                 # flat grads per group (after noise + clip if requested)
    scores: list                # per-hop score[B,N]
    attprob: list               # per-hop p[B,S]
    do_pred: list               # per-hop do_pred[B]
    answers: np.ndarray         # argmax per hop (+uni, +select) [nHop+2, B], 1-based
    norms: dict = field(default_factory=dict)
    rnn_out: np.ndarray | None = None


def feval(cfg: RauConfig, params: dict, X, x, x_len, y, masks=None, hop_mask=None,
          noise=None, step_t=None, clip=True):
    """One training step's forward + backward (F:445-650) on flat float64 params.
    masks: None (all dropout off, i.e. eval-mode forward with training-mode backward semantics) or
      dict(embed=[T,B,200], rnn=[T,B,Hq], hops=[dict(q,X,m)]*nHop).
    hop_mask[h]: tab_multhop_compute_loss (F:587-589); noise: dict group->array or None.
    Returns StepResult with grads *after* noise and per-group clip (clip=False stops before F:617)."""
    Pe = views(cfg, "embed", params["embed"])
    Pr = views(cfg, "rnn", params["rnn"])
    Pm = views(cfg, "mult", params["mult"])
    grads = {g: np.zeros_like(params[g]) for g in GROUPS}             # F:446-448
    gPe, gPr, gPm = (views(cfg, g, grads[g]) for g in GROUPS)
    B = X.shape[0]
    nHop = cfg.nHop
    hop_mask = [True] * nHop if hop_mask is None else hop_mask
    enc_masks = None if masks is None else dict(embed=masks["embed"], rnn=masks["rnn"])
    rnn_out, ecaches = encoder_fwd(Pe, Pr, cfg, x, x_len, enc_masks)
    c = np.zeros((B, cfg.H))
    h = np.zeros((B, cfg.H))
    caches, scores, attp, dps, cs, hs = [], [], [], [], [c], [h]
    uni = np.zeros((B, cfg.N))
    sel = np.zeros((B, cfg.N))
    did_pred = np.zeros(B)
    loss = np.zeros(nHop + 2)
    loss_dp = np.zeros(nHop)
    answers = np.zeros((nHop + 2, B), dtype=np.int64)
    for hp in range(nHop):
        hm = None if masks is None else masks["hops"][hp]
        score, do_pred, p, c, h, cache = hop_fwd(Pm, cfg, rnn_out, X, c, h, hm)   # F:497
        uni += score                                                   # F:499
        ans = argmax1(score)                                           # F:505
        answers[hp] = ans
        is_correct = (ans == np.asarray(y).astype(np.int64)).astype(np.float64)
        dp_bin = (do_pred > 0.5).astype(np.float64)                    # F:518
        cur = np.clip(dp_bin - did_pred, 0, 1)                         # F:522
        sel += score * cur[:, None]                                    # F:524
        did_pred = np.clip(did_pred + dp_bin, 0, 1)                    # F:532
        loss[hp] = cross_entropy_fwd(score, y)                         # F:535
        loss_dp[hp] = bce_fwd(do_pred, is_correct)                     # F:572
        caches.append(cache); scores.append(score); attp.append(p); dps.append(do_pred)
        cs.append(c); hs.append(h)
    uni /= nHop                                                        # F:539
    answers[nHop] = argmax1(uni)
    loss[nHop] = cross_entropy_fwd(uni, y)                             # F:547-548
    answers[nHop + 1] = argmax1(sel)
    loss[nHop + 1] = cross_entropy_fwd(sel, y)                         # F:556-557
    # backward through hops (F:578-597)
    dc = np.zeros((B, cfg.H))
    dh = np.zeros((B, cfg.H))
    dq_sum = np.zeros((B, cfg.Q))
    zeros_p = np.zeros((B, cfg.S))
    zeros_d = np.zeros(B)
    for hp in range(nHop - 1, -1, -1):
        dscore = cross_entropy_bwd(scores[hp], y)                      # F:585
        if not hop_mask[hp]:
            dscore = dscore * 0.0                                      # F:587-589
        dq, _, dc, dh = hop_bwd(Pm, gPm, cfg, caches[hp], dscore, zeros_d, zeros_p, dc, dh)  # F:582-583,590
        dq_sum += dq                                                   # branch:backward sums (F:598)
    encoder_bwd(Pr, gPe, gPr, cfg, x_len, ecaches, dq_sum)
    norms = {}
    if clip:
        for g in GROUPS:                                               # F:617-648
            nz = None if noise is None else noise[g]
            norms[g] = noise_and_clip(cfg, grads[g], nz)
    return StepResult(loss=loss, loss_do_pred=loss_dp, grads=grads, scores=scores, attprob=attp,
                      do_pred=dps, answers=answers, norms=norms, rnn_out=rnn_out)


def train_step(cfg, params, opt_state, X, x, x_len, y, masks=None, hop_mask=None, noise=None,
               optim="adam", lrs=None):
    """feval + the three optimizer calls (F:787-791). Mutates params/opt_state in place."""
    res = feval(cfg, params, X, x, x_len, y, masks, hop_mask, noise)
    lrs = lrs or dict(embed=cfg.lr, rnn=cfg.lr, mult=cfg.mult_lr)
    for g in GROUPS:
        st = opt_state.setdefault(g, {})
        if optim == "adam":
            adam(params[g], res.grads[g], lrs[g], st)
        elif optim == "rmsprop":
            rmsprop(params[g], res.grads[g], lrs[g], 0.99, 1e-8, st)
        elif optim == "sgd":
            sgd(params[g], res.grads[g], lrs[g])
        else:
            raise ValueError(optim)
    return res


# ------------------------------------------------------------------- predict_result (a8)

def predict(cfg, params, X, x, x_len):
    """predict_result (F:652-724): evaluate()-mode forward; returns (tab_pred[nHop+2], tab_att[nHop+2])."""
    Pe = views(cfg, "embed", params["embed"])
    Pr = views(cfg, "rnn", params["rnn"])
    Pm = views(cfg, "mult", params["mult"])
    B = X.shape[0]
    rnn_out, _ = encoder_fwd(Pe, Pr, cfg, x, x_len, None)
    c = np.zeros((B, cfg.H)); h = np.zeros((B, cfg.H))
    uni = np.zeros((B, cfg.N)); uni_att = np.zeros((B, cfg.S))
    sel = np.zeros((B, cfg.N)); sel_att = np.zeros((B, cfg.S))
    did = np.zeros(B)
    preds, atts = [], []
    for hp in range(cfg.nHop):
        score, do_pred, p, c, h, _ = hop_fwd(Pm, cfg, rnn_out, X, c, h, None)
        uni += score; uni_att += p                                     # F:699-700
        dp = (do_pred > 0.5).astype(np.float64)
        if hp == cfg.nHop - 1:
            dp[:] = 1.0                                                # F:704
        cur = np.clip(dp - did, 0, 1)                                  # F:705
        sel += score * cur[:, None]; sel_att += p * cur[:, None]       # F:706-707
        did = np.clip(did + dp, 0, 1)                                  # F:716
        preds.append(score); atts.append(p)
    preds += [uni / cfg.nHop, sel]                                     # F:718-721
    atts += [uni_att / cfg.nHop, sel_att]
    return preds, atts


# -------------------------------------------------------------------- synthetic batches

def synth_batch(cfg: RauConfig, B: int, seed: int = 123, min_len: int = 8):
    """X = max(0, N(0,1)); tokens U[2,V], pad=1 past the length; len U[min_len,T]; y U[1,N]."""
    rng = np.random.default_rng(seed)
    X = np.maximum(rng.standard_normal((B, cfg.C, cfg.S)), 0.0)
    x_len = rng.integers(min(min_len, cfg.T), cfg.T + 1, B)
    x = rng.integers(2, cfg.V + 1, (cfg.T, B))
    for b in range(B):
        x[x_len[b]:, b] = 1
    y = rng.integers(1, cfg.N + 1, B)
    return X, x, x_len, y


def synth_masks(cfg: RauConfig, B: int, seed: int = 7):
    rng = np.random.default_rng(seed)
    bern = lambda shape, p: (rng.random(shape) >= p).astype(np.float64)
    return dict(
        embed=bern((cfg.T, B, cfg.embed), cfg.p_embed),
        rnn=bern((cfg.T, B, cfg.Hq), cfg.p_rnn),
        hops=[dict(q=bern((B, cfg.Q), cfg.p_q), X=bern((B, cfg.C, cfg.S), cfg.p_x),
                   m=bern((B, cfg.M), cfg.p_m)) for _ in range(cfg.nHop)],
    )
