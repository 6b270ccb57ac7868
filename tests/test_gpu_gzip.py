"""-m gpu: the GPU gzip writer (option "deflate", pbsim_b200/csrc/gz_kernels.cuh).  The members handed out by
host delivery must gunzip (CRC-32 and ISIZE checked by zlib) to exactly the text the engine delivers without the
option — which the other parity tests pin to the reference's bytes."""
import gzip
import zlib

import pytest

from pbsim_b200 import capi, simulator
from tests.golden_util import Case, SetCase, model_path
from tests.gpu_util import engine_model, run_case_on_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = simulator.Engine(0)
    yield e
    e.set_option("deflate", 0)
    e.close()


def _members(blob):
    """split a concatenation of gzip members, checking every CRC / ISIZE; returns (n_members, text)"""
    out, n = [], 0
    while blob:
        d = zlib.decompressobj(16 + zlib.MAX_WBITS)
        out.append(d.decompress(blob))
        assert d.eof
        blob = d.unused_data
        n += 1
    return n, b"".join(out)


@pytest.mark.parametrize("name", ["qs_rsii_basic", "err_sequel_multipass", "qs_delheavy_uniform"])
def test_members_gunzip_to_the_text(eng, name):
    c = Case(name)
    eng.set_option("deflate", 0)
    text = run_case_on_gpu(c, eng, "philox")
    eng.set_option("deflate", 1)
    try:
        comp = run_case_on_gpu(c, eng, "philox")
    finally:
        eng.set_option("deflate", 0)
    for (r0, m0, st0, t0), (r1, m1, st1, t1) in zip(text, comp):
        assert gzip.decompress(r1) == r0
        assert gzip.decompress(m1) == m0
        assert t0 == t1
        n, _ = _members(r1)
        # one member per 32 KiB of text, per batch (the quota's tail reads are batches of their own)
        assert n >= (len(r0) + 32767) // 32768
        assert len(r1) % 4 == 0 and len(r1) < 0.62 * len(r0)
        assert len(m1) < 0.5 * len(m0)


def test_many_batches_and_small_pieces(eng):
    """members must survive batch boundaries (a partial last unit per batch and stream) and 4 KiB delivery pieces"""
    L = capi.load()
    hm = capi.HostModel(L, capi.host_params("qshmm"), model_path("QSHMM-RSII.model"))
    eng.set_model(hm)
    n = 1000000
    eng.set_synthetic_sequence(n, 1, 5)
    ref = eng.simulate(2 * n, rng_mode=capi.RNG_PHILOX, seed=9)
    eng.set_option("deflate", 1)
    eng.set_option("stage_bytes", 4096)
    try:
        got = eng.simulate(2 * n, rng_mode=capi.RNG_PHILOX, seed=9, batch_reads=37)
    finally:
        eng.set_option("deflate", 0)
        eng.set_option("stage_bytes", 128 << 20)
    assert got[3] > 50
    assert gzip.decompress(got[0]) == ref[0]
    assert gzip.decompress(got[1]) == ref[1]
    assert got[2].res_len_total == ref[2].res_len_total
    assert got[2].deflate_seconds > 0


def test_transcript_run_compressed(eng):
    c = SetCase("tr_qs_rsii_basic")
    hm = capi.HostModel(capi.load(), capi.host_params(c.method, **c.okw), c.model)
    run = simulator.SetRun(eng, hm, c.strategy)
    want, _ = c.run_oracle("philox")
    eng.set_option("deflate", 1)
    try:
        reads, maf, st, text = run.simulate(c.seqset, rng_mode=capi.RNG_PHILOX, seed=c.seed)
    finally:
        eng.set_option("deflate", 0)
    assert gzip.decompress(reads) == want["reads"]
    assert gzip.decompress(maf) == want["maf"]
