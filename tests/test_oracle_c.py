"""Pins the plain-C restatement (oracle/osq_oracle.c) against the reference's golden vectors and
against the torch oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import build_c
from oracle import osq_oracle as O


@pytest.fixture(scope="module")
def lib():
    l = C.CDLL(build_c.build())
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    l.osqo_fq_per_tensor.argtypes = [fp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, fp, fp]
    l.osqo_fq_per_channel.argtypes = [fp, C.c_int64, C.c_int64, fp, ip, C.c_float, C.c_float, fp, fp]
    l.osqo_qparams.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, fp, fp]
    l.osqo_lsqplus_effective.argtypes = [C.c_float, C.c_float, C.c_float, fp, fp]
    l.osqo_token_minmax.argtypes = [fp] + [C.c_int64] * 8 + [C.POINTER(C.c_int64), C.c_int64, fp, fp]
    l.osqo_token_minmax.restype = C.c_int64
    return l


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def P(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def test_c_fq_per_tensor_golden(lib, golden):
    g = golden("fq_per_tensor")
    for i in range(int(g["n"])):
        scale, zp, qmin, qmax = g["p%d" % i]
        x = f32(g["x%d" % i]); y = np.empty_like(x); q = np.empty_like(x)
        lib.osqo_fq_per_tensor(P(x), x.size, scale, zp, qmin, qmax, P(y), P(q))
        np.testing.assert_array_equal(y, g["y%d" % i])
        np.testing.assert_array_equal(q, g["q%d" % i])


def test_c_fq_per_channel_and_qparams_golden(lib, golden):
    g = golden("fq_per_channel")
    for i in range(int(g["n"])):
        bit, sym, qmin, qmax = (int(v) for v in g["p%d" % i])
        w = f32(g["w2_%d" % i]); y = np.empty_like(w)
        scale = f32(g["scale%d" % i]); zp = np.ascontiguousarray(g["zp%d" % i], dtype=np.int32)
        lib.osqo_fq_per_channel(P(w), w.shape[0], w.shape[1], P(scale), P(zp, C.c_int32), qmin, qmax, P(y), None)
        np.testing.assert_array_equal(y, g["y2_%d" % i])
        for r in range(w.shape[0]):
            s, z = C.c_float(), C.c_float()
            lib.osqo_qparams(float(g["min%d" % i][r]), float(g["max%d" % i][r]), qmin, qmax, sym, C.byref(s), C.byref(z))
            assert np.float32(s.value) == scale[r] and int(z.value) == zp[r]


def test_c_lsqplus_effective_matches_torch_oracle(lib):
    rng = np.random.default_rng(0)
    for _ in range(2000):
        s, z = np.float32(rng.uniform(1e-3, 2)), np.float32(rng.uniform(0, 63))
        n = int(rng.integers(10, 10 ** 7))
        se, ze, g = O.lsqplus_effective_qparams(torch.tensor([s]), torch.tensor([z]), n, 63)
        a, b = C.c_float(), C.c_float()
        lib.osqo_lsqplus_effective(float(s), float(z), float(np.float32(g)), C.byref(a), C.byref(b))
        assert np.float32(a.value) == se.numpy()[0] and np.float32(b.value) == ze.numpy()[0]


def test_c_token_minmax_golden(lib, golden):
    g = golden("observers")
    x = torch.from_numpy(g["q4d_x0"])  # [B,h,S,d], seq_pos=2
    lens = np.array([12, 7, 1, 9], dtype=np.int64)
    xs = f32(x.numpy())
    B, h, S, d = x.shape
    tmin = np.empty(B * S, np.float32); tmax = np.empty(B * S, np.float32)
    T = lib.osqo_token_minmax(P(xs), B, S, h, d, h * S * d, d, S * d, 1, P(lens, C.c_int64), 4, P(tmin), P(tmax))
    tok = O.token_matrix(x, lens.tolist(), 2)
    assert T == tok.shape[0]
    a, b = O.token_minmax(tok)
    np.testing.assert_array_equal(tmin[:T], a.numpy()); np.testing.assert_array_equal(tmax[:T], b.numpy())
