"""Pins oracle/glibc_rand.c (restatement of glibc srand/rand, the RNG behind every draw of the
reference, pbsim.cpp:543) and oracle/philox.h."""
import ctypes
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.golden_util import GOLDEN, Case, case_names


def test_glibc_rand_matches_committed_known_answers():
    with open(os.path.join(GOLDEN, "glibc_rand_kat.json")) as f:
        kat = json.load(f)["values"]
    for seed, vals in kat.items():
        got = O.glibc_rand(int(seed), len(vals))
        assert got.tolist() == vals, "seed %s" % seed


def test_glibc_rand_matches_system_libc_when_it_is_glibc():
    libc = ctypes.CDLL(None)
    if not hasattr(libc, "gnu_get_libc_version"):
        pytest.skip("not glibc")
    for seed in (3, 77, 123456789):
        libc.srand(ctypes.c_uint(seed))
        want = [int(libc.rand()) for _ in range(2000)]
        assert O.glibc_rand(seed, 2000).tolist() == want


@pytest.mark.parametrize("name", case_names())
def test_rand_head_of_every_golden_case(name):
    c = Case(name)
    head = np.load(os.path.join(c.dir, "rand_head.npy"))
    assert np.array_equal(O.glibc_rand(c.seed, len(head)), head)


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    vec = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in vec:
        assert tuple(int(x) for x in O.philox_block(ctr, key)) == want
