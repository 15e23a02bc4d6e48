"""ctypes binding of oracle/rans_oracle.c + an independent pure-Python rANS for tiny cases.

TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/__init__.py).
"""
import ctypes
import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    """Compile oracle/rans_oracle.c with gcc (oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "rans_oracle.c")
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i32p = ctypes.POINTER(ctypes.c_int32)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        f32p = ctypes.POINTER(ctypes.c_float)
        L.oracle_pmf_to_quantized_cdf.argtypes = [f32p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]
        L.oracle_pmf_to_quantized_cdf.restype = ctypes.c_int
        L.oracle_rans_encode.argtypes = [i32p, i32p, ctypes.c_int, i32p, ctypes.c_int, i32p, i32p, u8p, ctypes.c_int64]
        L.oracle_rans_encode.restype = ctypes.c_int64
        L.oracle_rans_decode.argtypes = [u8p, ctypes.c_int64, i32p, ctypes.c_int, i32p, ctypes.c_int, i32p, i32p, i32p]
        L.oracle_rans_decode.restype = ctypes.c_int
        L.oracle_quantize_symbols.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p, f32p, f32p, i32p]
        L.oracle_quantize_symbols.restype = None
        L.oracle_rans_encode_batch.argtypes = [i32p, ctypes.c_int64, ctypes.c_int, i32p, ctypes.c_int, i32p, i32p, u8p, ctypes.c_int64, i32p]
        L.oracle_rans_encode_batch.restype = ctypes.c_int
        L.oracle_rans_decode_batch.argtypes = [u8p, ctypes.c_int64, i32p, ctypes.c_int64, ctypes.c_int, i32p, ctypes.c_int, i32p, i32p, i32p]
        L.oracle_rans_decode_batch.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


#: worst case bytes for one stream of c symbols: 16 bits/symbol + 9 bypass nibbles, +2 flush words
def stream_cap_bytes(c):
    return 4 * ((c * (16 + 36) + 31) // 32 + 4)


def pmf_to_quantized_cdf(pmf, precision=16):
    """compressai._CXX.pmf_to_quantized_cdf (ops.cpp).  pmf: float32[n] -> int32[n+1]."""
    pmf = np.ascontiguousarray(pmf, dtype=np.float32)
    out = np.zeros(pmf.size + 1, dtype=np.uint32)
    rc = lib().oracle_pmf_to_quantized_cdf(_p(pmf, ctypes.c_float), pmf.size, precision, _p(out, ctypes.c_uint32))
    if rc != 0:
        raise ValueError(f"oracle_pmf_to_quantized_cdf failed rc={rc}")
    return out.astype(np.int32)


def encode_with_indexes(symbols, indexes, cdfs, cdf_sizes, offsets):
    """compressai.ans.RansEncoder.encode_with_indexes -> bytes."""
    symbols, indexes = _i32(symbols), _i32(indexes)
    cdfs, cdf_sizes, offsets = _i32(cdfs), _i32(cdf_sizes), _i32(offsets)
    cap = stream_cap_bytes(symbols.size)
    out = np.zeros(cap, dtype=np.uint8)
    nb = lib().oracle_rans_encode(
        _p(symbols, ctypes.c_int32), _p(indexes, ctypes.c_int32), symbols.size,
        _p(cdfs, ctypes.c_int32), cdfs.shape[1], _p(cdf_sizes, ctypes.c_int32),
        _p(offsets, ctypes.c_int32), _p(out, ctypes.c_uint8), cap)
    if nb < 0:
        raise ValueError(f"oracle_rans_encode failed rc={nb}")
    return out[:nb].tobytes()


def decode_with_indexes(encoded, indexes, cdfs, cdf_sizes, offsets):
    """compressai.ans.RansDecoder.decode_with_indexes -> int32[n]."""
    indexes = _i32(indexes)
    cdfs, cdf_sizes, offsets = _i32(cdfs), _i32(cdf_sizes), _i32(offsets)
    enc = np.frombuffer(encoded, dtype=np.uint8).copy()
    out = np.zeros(indexes.size, dtype=np.int32)
    rc = lib().oracle_rans_decode(
        _p(enc, ctypes.c_uint8), enc.size, _p(indexes, ctypes.c_int32), indexes.size,
        _p(cdfs, ctypes.c_int32), cdfs.shape[1], _p(cdf_sizes, ctypes.c_int32),
        _p(offsets, ctypes.c_int32), _p(out, ctypes.c_int32))
    if rc != 0:
        raise ValueError(f"oracle_rans_decode failed rc={rc}")
    return out


def quantize_symbols(z, bias, exp_s, medians):
    """int(rint((z.float() + b) * exp(s) - median)), one float32 rounding per op."""
    z = np.ascontiguousarray(z, dtype=np.float32)
    rows, c = z.shape
    bias = np.ascontiguousarray(bias, dtype=np.float32)
    exp_s = np.ascontiguousarray(exp_s, dtype=np.float32)
    medians = np.ascontiguousarray(medians, dtype=np.float32)
    sym = np.zeros((rows, c), dtype=np.int32)
    lib().oracle_quantize_symbols(_p(z, ctypes.c_float), rows, c, _p(bias, ctypes.c_float),
                                  _p(exp_s, ctypes.c_float), _p(medians, ctypes.c_float),
                                  _p(sym, ctypes.c_int32))
    return sym


def encode_batch(symbols, cdfs, cdf_sizes, offsets):
    """Encode [rows, c] symbols (indexes = arange(c)) -> list[bytes] (single thread)."""
    symbols = _i32(symbols)
    rows, c = symbols.shape
    cdfs, cdf_sizes, offsets = _i32(cdfs), _i32(cdf_sizes), _i32(offsets)
    cap = stream_cap_bytes(c)
    out = np.zeros((rows, cap), dtype=np.uint8)
    lens = np.zeros(rows, dtype=np.int32)
    rc = lib().oracle_rans_encode_batch(
        _p(symbols, ctypes.c_int32), rows, c, _p(cdfs, ctypes.c_int32), cdfs.shape[1],
        _p(cdf_sizes, ctypes.c_int32), _p(offsets, ctypes.c_int32), _p(out, ctypes.c_uint8), cap,
        _p(lens, ctypes.c_int32))
    if rc != 0:
        raise ValueError(f"oracle_rans_encode_batch failed rc={rc}")
    return [out[r, : lens[r]].tobytes() for r in range(rows)]


def decode_batch(strings, c, cdfs, cdf_sizes, offsets):
    """Decode list[bytes] -> int32[rows, c] (indexes = arange(c), single thread)."""
    rows = len(strings)
    cdfs, cdf_sizes, offsets = _i32(cdfs), _i32(cdf_sizes), _i32(offsets)
    cap = max(8, max(len(s) for s in strings))
    enc = np.zeros((rows, cap), dtype=np.uint8)
    lens = np.zeros(rows, dtype=np.int32)
    for r, s in enumerate(strings):
        enc[r, : len(s)] = np.frombuffer(s, dtype=np.uint8)
        lens[r] = len(s)
    out = np.zeros((rows, c), dtype=np.int32)
    rc = lib().oracle_rans_decode_batch(
        _p(enc, ctypes.c_uint8), cap, _p(lens, ctypes.c_int32), rows, c,
        _p(cdfs, ctypes.c_int32), cdfs.shape[1], _p(cdf_sizes, ctypes.c_int32),
        _p(offsets, ctypes.c_int32), _p(out, ctypes.c_int32))
    if rc != 0:
        raise ValueError(f"oracle_rans_decode_batch failed rc={rc}")
    return out


# ---------------------------------------------------------------------------------------
# Independent pure-Python restatement (big ints, no C) -- small cases only; used to check
# that the C oracle and the written description of the format agree.
# ---------------------------------------------------------------------------------------
_M32 = (1 << 32) - 1
_L = 1 << 31


def py_encode(symbols, indexes, cdfs, cdf_sizes, offsets):
    pushed = []
    for s, ci in zip(symbols, indexes):
        cdf = cdfs[ci]
        maxv = int(cdf_sizes[ci]) - 2
        v = int(s) - int(offsets[ci])
        raw = 0
        if v < 0:
            raw, v = -2 * v - 1, maxv
        elif v >= maxv:
            raw, v = 2 * (v - maxv), maxv
        pushed.append((int(cdf[v]), int(cdf[v + 1]) - int(cdf[v]), False))
        if v == maxv:
            nb = 0
            while (raw >> (4 * nb)) != 0:
                nb += 1
            val = nb
            while val >= 15:
                pushed.append((15, 16, True))
                val -= 15
            pushed.append((val, val + 1, True))
            for j in range(nb):
                pushed.append(((raw >> (4 * j)) & 15, 0, True))
    x, emitted = _L, []
    for start, rng, byp in reversed(pushed):
        if not byp:
            if x >= ((_L >> 16) << 32) * rng:
                emitted.append(x & _M32)
                x >>= 32
            x = ((x // rng) << 16) + (x % rng) + start
        else:
            if x >= ((_L >> 16) << 32) * (1 << 12):
                emitted.append(x & _M32)
                x >>= 32
            x = (x << 4) | start
    emitted.append((x >> 32) & _M32)
    emitted.append(x & _M32)
    words = emitted[::-1]
    return struct.pack("<%dI" % len(words), *words)


def py_decode(encoded, indexes, cdfs, cdf_sizes, offsets):
    w = struct.unpack("<%dI" % (len(encoded) // 4), encoded)
    x, p, out = w[0] | (w[1] << 32), 2, []

    def bits():
        nonlocal x, p
        val = x & 15
        x >>= 4
        if x < _L:
            x = (x << 32) | w[p]
            p += 1
        return val

    for ci in indexes:
        cdf, size = cdfs[ci], int(cdf_sizes[ci])
        maxv = size - 2
        cf = x & 0xFFFF
        s = next(k for k in range(size) if int(cdf[k]) > cf) - 1
        start, fr = int(cdf[s]), int(cdf[s + 1]) - int(cdf[s])
        x = fr * (x >> 16) + (x & 0xFFFF) - start
        if x < _L:
            x = (x << 32) | w[p]
            p += 1
        v = s
        if v == maxv:
            val = bits()
            nb = val
            while val == 15:
                val = bits()
                nb += val
            raw = 0
            for j in range(nb):
                raw |= bits() << (4 * j)
            v = raw >> 1
            v = -v - 1 if raw & 1 else v + maxv
        out.append(v + int(offsets[ci]))
    return out
