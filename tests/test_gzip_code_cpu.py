"""CPU: the host half of the GPU gzip writer (pbsim_b200/csrc/gz_host.hpp via pbsim_host_deflate_code).  A plain
Python encoder driven by the product's code table and block header must yield a DEFLATE stream that zlib inflates
back to the input — for record-like text, for skewed histograms that force the length limit, and for bytes the
histogram never saw."""
import zlib

import numpy as np
import pytest

from tests import hostsim_util as H


def deflate_code(hist):
    L = H.lib()
    from pbsim_b200 import capi
    capi.declare_host(L)
    import ctypes as C
    h = np.asarray(hist, dtype=np.int64)
    lit = np.zeros(257, dtype=np.uint32)
    hdr = np.zeros(128, dtype=np.uint32)
    nbits = C.c_uint32()
    rc = L.pbsim_host_deflate_code(h.ctypes.data, lit.ctypes.data, C.byref(nbits), hdr.ctypes.data, 128)
    assert rc == 0
    return lit, hdr, nbits.value


def encode(data, lit, hdr, hdr_bits):
    """one dynamic-Huffman block of literals, LSB-first bit packing (what k_gz_encode does per member)"""
    bits = []
    for i in range(hdr_bits):
        bits.append((int(hdr[i >> 5]) >> (i & 31)) & 1)
    for b in data + [256]:
        e = int(lit[b])
        code, n = e & 0xFFFF, e >> 16
        assert 1 <= n <= 12
        bits.extend((code >> k) & 1 for k in range(n))
    while len(bits) % 8:
        bits.append(0)
    a = np.array(bits, dtype=np.uint8).reshape(-1, 8)
    return bytes((a * (1 << np.arange(8))).sum(axis=1).astype(np.uint8))


def roundtrip(data, hist):
    lit, hdr, nbits = deflate_code(hist)
    raw = encode(list(data), lit, hdr, nbits)
    out = zlib.decompressobj(-15).decompress(raw)
    assert out == bytes(data)
    return len(raw), lit


def test_fastq_like_text_roundtrips_and_compresses():
    rng = np.random.default_rng(1)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 3000)]
    qual = (33 + rng.integers(2, 16, 3000)).astype(np.uint8)
    text = b"@S1_1\n" + seq.tobytes() + b"\n+S1_1\n" + qual.tobytes() + b"\n"
    hist = np.bincount(np.frombuffer(text, dtype=np.uint8), minlength=256)
    n, lit = roundtrip(text, hist)
    assert n < 0.6 * len(text)
    assert (int(lit[ord("A")]) >> 16) <= 4


def test_skewed_histogram_hits_the_length_limit():
    hist = np.zeros(256, dtype=np.int64)
    hist[ord("A")] = 10 ** 12
    for i, c in enumerate(range(100, 150)):
        hist[c] = 1 + i
    data = bytes([ord("A")] * 50 + list(range(100, 150)) + [0, 255, 7])  # three bytes the histogram never saw
    n, lit = roundtrip(data, hist)
    lens = (lit >> 16).astype(int)
    assert 10 <= lens.max() <= 12 and lens.min() == 1 and (lens[:256] > 0).all()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_histograms_give_valid_streams(seed):
    rng = np.random.default_rng(seed)
    hist = (rng.pareto(0.7, 256) * 100).astype(np.int64) * (rng.random(256) < 0.6)
    data = bytes(rng.integers(0, 256, 2000, dtype=np.uint8))
    roundtrip(data, hist)


def test_kraft_equality_and_uniform_input():
    lit, _, _ = deflate_code(np.ones(256, dtype=np.int64))
    lens = (lit >> 16).astype(int)
    assert sum(2.0 ** -int(x) for x in lens) == 1.0  # complete code: 257 symbols
    roundtrip(bytes(range(256)), np.ones(256, dtype=np.int64))


@pytest.mark.parametrize("seed", range(36))
def test_extreme_histograms_give_valid_streams(seed):
    """empty, single-symbol, power-of-two, Fibonacci-like (deepest Huffman trees) and heavy-tailed histograms: the code
    stays within 12 bits, is a prefix code, and zlib inflates what it encodes — for any input bytes"""
    rng = np.random.default_rng(50000 + seed)
    kind = seed % 6
    if kind == 0:
        hist = np.zeros(256, np.int64)
        hist[int(rng.integers(0, 256))] = int(rng.integers(1, 10 ** 15))
    elif kind == 1:
        hist = np.zeros(256, np.int64)
    elif kind == 2:
        hist = (2 ** rng.integers(0, 50, 256)).astype(np.int64) * (rng.random(256) < rng.random())
    elif kind == 3:
        hist = np.array([int(1.6180339 ** min(i, 80)) for i in range(256)], dtype=np.int64)
        rng.shuffle(hist)
    elif kind == 4:
        hist = rng.integers(0, 3, 256).astype(np.int64)
    else:
        hist = (rng.pareto(0.4, 256) * 1000).astype(np.int64)
    data = bytes(rng.integers(0, 256, int(rng.integers(0, 600)), dtype=np.uint8))
    _, lit = roundtrip(data, hist)
    lens = (lit >> 16).astype(int)
    assert 1 <= lens.min() and lens.max() <= 12
    assert sum(2.0 ** -int(x) for x in lens) <= 1.0
