"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the LSTM head arithmetic.

Independent of torch's nn.LSTM: spells out the published cell equations that
the reference reaches through torch (call site mobileposer/models/rnn.py:15,27;
torch/nn/modules/rnn.py "LSTM" docstring):

    gates = W_ih x_t + b_ih + W_hh h_{t-1} + b_hh          (row blocks i, f, g, o)
    c_t = sigmoid(f) * c_{t-1} + sigmoid(i) * tanh(g)
    h_t = sigmoid(o) * tanh(c_t)

Layer 1 consumes concat(fwd, rev) of layer 0; with packed sequences the
reverse direction of sequence b starts at its own last valid frame, and
pad_packed_sequence zero-fills frames >= len before linear2 (rnn.py:31-32).

Runs in float64 by default so tests can use it as the arbiter between two
float32 implementations (reference-on-CPU vs CUDA); pinned against the live
reference through tests/golden/ (tests/test_oracle.py).  Never imported by the
product package.
"""
from __future__ import annotations

import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_direction(x, w_ih, w_hh, b_ih, b_hh, reverse=False, h0=None, c0=None):
    """x [T, In] -> h sequence [T, H], (h_T, c_T) for one sequence, one direction."""
    T = x.shape[0]
    H = w_hh.shape[1]
    h = np.zeros(H, x.dtype) if h0 is None else h0.astype(x.dtype)
    c = np.zeros(H, x.dtype) if c0 is None else c0.astype(x.dtype)
    pre = x @ w_ih.T + (b_ih + b_hh)
    out = np.zeros((T, H), x.dtype)
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        g = pre[t] + w_hh @ h
        i, f, gg, o = g[:H], g[H:2 * H], g[2 * H:3 * H], g[3 * H:]
        c = _sigmoid(f) * c + _sigmoid(i) * np.tanh(gg)
        h = _sigmoid(o) * np.tanh(c)
        out[t] = h
    return out, h, c


def rnn_head(sd, prefix, x, lengths, bidirectional, state=None, dtype=np.float64):
    """x [B, T, n_in] -> y [B, T, n_out], (h_n, c_n) each [layers*dirs, B, H]."""
    g = lambda k: np.asarray(sd[prefix + k], dtype=dtype)
    x = np.asarray(x, dtype=dtype)
    B, T, _ = x.shape
    dirs = 2 if bidirectional else 1
    H = g('linear1.weight').shape[0]
    a = np.maximum(x @ g('linear1.weight').T + g('linear1.bias'), 0.0)
    hn = np.zeros((2 * dirs, B, H), dtype)
    cn = np.zeros((2 * dirs, B, H), dtype)
    top = np.zeros((B, T, dirs * H), dtype)
    for b in range(B):
        L = int(lengths[b])
        inp = a[b, :L]
        for layer in range(2):
            outs = []
            for d in range(dirs):
                sfx = f'_l{layer}' + ('_reverse' if d else '')
                idx = layer * dirs + d
                h0 = c0 = None
                if state is not None:
                    h0, c0 = np.asarray(state[0])[idx, b], np.asarray(state[1])[idx, b]
                o, h, c = lstm_direction(inp, g('rnn.weight_ih' + sfx), g('rnn.weight_hh' + sfx),
                                         g('rnn.bias_ih' + sfx), g('rnn.bias_hh' + sfx),
                                         reverse=bool(d), h0=h0, c0=c0)
                outs.append(o)
                hn[idx, b], cn[idx, b] = h, c
            inp = np.concatenate(outs, axis=1)
        top[b, :L] = inp
    y = top @ g('linear2.weight').T + g('linear2.bias')
    return y, (hn, cn)
