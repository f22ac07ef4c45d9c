"""ctypes front end of oracle/pc_codec.c (the CPU oracle of the PC1 entropy coder).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__ (build + smoke) and bench.py's CPU legs, never by dsin_b200."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "pc_codec.c")
_OUT = os.path.join(_HERE, "_build", "libpc_codec.so")
PC = "imgcomp/probclass3d/logits/"
ENC_CENTERS = "encoder/encoder_body/encoder_body/autoencoder/encoder/centers"

# live taps (kd, kh, kw) of the masked (2,3,3) kernels in coding order (pc_codec.c TAPS)
TAPS_OTHER = [(0, kh, kw) for kh in range(3) for kw in range(3)] + [(1, 0, 0), (1, 0, 1), (1, 0, 2), (1, 1, 0), (1, 1, 1)]
TAPS_FIRST = TAPS_OTHER[:-1]


class _Model(C.Structure):
    _fields_ = [("C", C.c_int), ("H", C.c_int), ("W", C.c_int), ("L", C.c_int), ("K", C.c_int),
                ("centers", C.c_void_p), ("pad", C.c_float)] + [(n, C.c_void_p) for n in
                                                                  ("w0", "b0", "w1", "b1", "w2", "b2", "w3", "b3")]


def build(force=False):
    """gcc the C oracle (no-op when the .so is newer than the source)."""
    if not force and os.path.isfile(_OUT) and os.path.getmtime(_OUT) >= os.path.getmtime(_SRC):
        return _OUT
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", _OUT, _SRC, "-lm"])
    return _OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.pc1_encode.restype = C.c_int
        _lib.pc1_decode.restype = C.c_int
        _lib.pc1_exp.restype = C.c_float
        _lib.pc1_exp.argtypes = [C.c_float]
    return _lib


def pack_layer(w_dhwio, first):
    """TF conv3d weights (2,3,3,cin,cout) -> live-tap-major (taps, cin, cout) float32."""
    taps = TAPS_FIRST if first else TAPS_OTHER
    w = np.asarray(w_dhwio, np.float32)
    return np.ascontiguousarray(np.stack([w[kd, kh, kw] for kd, kh, kw in taps]))


def pack_weights(W):
    """dict keyed by TF variable names -> the eight arrays of pc1_model + centres."""
    sc = [PC + "conv3d_conv0_mask", PC + "res1/conv3d_conv1_mask", PC + "res1/conv3d_conv2_mask", PC + "conv3d_conv2_mask"]
    arrs = []
    for i, s in enumerate(sc):
        arrs.append(pack_layer(W[s + "/weights"], first=(i == 0)))
        arrs.append(np.ascontiguousarray(W[s + "/biases"], np.float32))
    return arrs, np.ascontiguousarray(W[ENC_CENTERS], np.float32)


def _model(shape, arrs, centers):
    m = _Model()
    m.C, m.H, m.W = (int(v) for v in shape)
    m.L, m.K = int(centers.size), int(arrs[1].size)
    m.centers = centers.ctypes.data
    m.pad = float(centers[0])
    for name, a in zip(("w0", "b0", "w1", "b1", "w2", "b2", "w3", "b3"), arrs):
        setattr(m, name, a.ctypes.data)
    return m


def encode(symbols_chw, W, nstreams=8, cap=None):
    """-> (list of per-stream bytes, ideal code length in bits of the quantised frequencies)."""
    sym = np.ascontiguousarray(symbols_chw, np.int32)
    arrs, centers = pack_weights(W)
    m = _model(sym.shape, arrs, centers)
    cap = int(cap or (2 * sym.size + 64))  # a symbol costs at most 16 bits (every frequency is >= 1 of 65536)
    out = np.zeros((nstreams, cap), np.uint8)
    sizes = np.zeros(nstreams, np.int64)
    ideal = C.c_double(0.0)
    rc = lib().pc1_encode(C.byref(m), C.c_void_p(sym.ctypes.data), C.c_int(nstreams), C.c_void_p(out.ctypes.data),
                          C.c_int64(cap), C.c_void_p(sizes.ctypes.data), C.byref(ideal))
    if rc != 0:
        raise RuntimeError("pc1_encode failed: %d" % rc)
    return [out[k, :sizes[k]].tobytes() for k in range(nstreams)], ideal.value


def decode(streams, shape_chw, W):
    arrs, centers = pack_weights(W)
    m = _model(shape_chw, arrs, centers)
    nstreams = len(streams)
    cap = max(len(s) for s in streams) + 8
    buf = np.zeros((nstreams, cap), np.uint8)
    for k, s in enumerate(streams):
        buf[k, :len(s)] = np.frombuffer(s, np.uint8)
    sizes = np.array([len(s) for s in streams], np.int64)
    sym = np.zeros(shape_chw, np.int32)
    ideal = C.c_double(0.0)
    rc = lib().pc1_decode(C.byref(m), C.c_void_p(sym.ctypes.data), C.c_int(nstreams), C.c_void_p(buf.ctypes.data),
                          C.c_int64(cap), C.c_void_p(sizes.ctypes.data), C.byref(ideal))
    if rc != 0:
        raise RuntimeError("pc1_decode failed: %d" % rc)
    return sym


def freqs(logits):
    l = np.ascontiguousarray(logits, np.float32)
    f = np.zeros(l.size, np.uint32)
    lib().pc1_freqs(C.c_void_p(l.ctypes.data), C.c_int(l.size), C.c_void_p(f.ctypes.data))
    return f


def exp_det(x):
    return float(lib().pc1_exp(C.c_float(x)))


def rc_selftest(freq_tables, symbols):
    """Range coder alone: (n, L) uint32 tables summing to 65536 + n symbols -> (mismatches after decode, bytes)."""
    f = np.ascontiguousarray(freq_tables, np.uint32)
    s = np.ascontiguousarray(symbols, np.int32)
    buf = np.zeros(s.size * 3 + 64, np.uint8)
    size = C.c_int64(0)
    lib().pc1_rc_selftest.restype = C.c_int
    bad = lib().pc1_rc_selftest(C.c_void_p(f.ctypes.data), C.c_void_p(s.ctypes.data), C.c_int(s.size), C.c_int(f.shape[1]),
                                C.c_void_p(buf.ctypes.data), C.c_int64(buf.size), C.byref(size))
    return bad, buf[:size.value].tobytes()
