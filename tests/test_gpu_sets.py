"""-m gpu: the transcript (--strategy trans) and template (--strategy templ) strategies on the CUDA engine,
through the C ABI (pbsim_cuda_set_seqset): byte parity with the reference in replay mode and with the oracle in
PHILOX mode, over the nine reference runs captured in tests/golden/sets/."""
import numpy as np
import pytest

from oracle import oracle as O
from pbsim_b200 import capi, simulator
from tests.golden_util import SetCase, set_case_names

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = simulator.Engine(0)
    yield e
    e.close()


def _host_model(c):
    return capi.HostModel(capi.load(), capi.host_params(c.method, **c.okw), c.model)


@pytest.mark.parametrize("name", set_case_names())
def test_replay_reproduces_reference_set_run(eng, name):
    c = SetCase(name)
    draws = O.glibc_rand(c.seed, c.ndraws)
    starts = np.concatenate([[0], c.marks[:-1]]).astype(np.int64)
    run = simulator.SetRun(eng, _host_model(c), c.strategy, hp_del_bias=c.okw.get("hp_del_bias", 1.0))
    reads, maf, st, text = run.simulate(c.seqset, rng_mode=capi.RNG_REPLAY, replay_draws=draws, replay_starts=starts)
    assert reads == c.reads(), "reads differ from the reference"
    assert maf == c.maf(), "maf differs from the reference"
    assert text == c.stats_text


@pytest.mark.parametrize("name", set_case_names())
@pytest.mark.parametrize("seg_min_len", [2048, 1024])
def test_philox_equals_oracle_set_run(eng, name, seg_min_len):
    c = SetCase(name)
    want, _ = c.run_oracle("philox")
    run = simulator.SetRun(eng, _host_model(c), c.strategy, hp_del_bias=c.okw.get("hp_del_bias", 1.0))
    eng.set_option("seg_min_len", seg_min_len)
    try:
        reads, maf, st, text = run.simulate(c.seqset, rng_mode=capi.RNG_PHILOX, seed=c.seed)
    finally:
        eng.set_option("seg_min_len", 2048)
    assert reads == want["reads"], "reads differ from the oracle"
    assert maf == want["maf"], "maf differs from the oracle"
    assert text == want["stats_text"]


def test_set_run_is_independent_of_batching_and_read_ranges(eng):
    """reads depend only on (seed, read number): small batches and a split into two read ranges give the same bytes"""
    c = SetCase("tr_qs_rsii_basic")
    run = simulator.SetRun(eng, _host_model(c), c.strategy)
    whole = run.simulate(c.seqset, rng_mode=capi.RNG_PHILOX, seed=5)
    small = run.simulate(c.seqset, rng_mode=capi.RNG_PHILOX, seed=5, batch_reads=3)
    assert whole[0] == small[0] and whole[1] == small[1] and whole[3] == small[3]
    n = whole[2].res_num
    a = eng.simulate(0, rng_mode=capi.RNG_PHILOX, seed=5, first_read=0, max_reads=n // 2)
    b = eng.simulate(0, rng_mode=capi.RNG_PHILOX, seed=5, first_read=n // 2)
    assert a[0] + b[0] == whole[0] and a[1] + b[1] == whole[1]
    assert a[2].res_num + b[2].res_num == n


def test_set_rejects_what_the_reference_cannot_handle(eng):
    c = SetCase("tr_qs_rsii_basic")
    eng.set_model(_host_model(c))
    with pytest.raises(simulator.EngineError, match="longer than 20 bases"):
        eng.set_seqset("trans", [("t1", 1, 0, b"ACGTACGTAC")], [0.0] + [1.0] * 10 + [0.0])
    with pytest.raises(simulator.EngineError, match="no reads"):
        eng.set_seqset("trans", [("t1", 0, 0, b"ACGT" * 30)], [0.0] + [1.0] * 10 + [0.0])
    # a WGS sequence afterwards resets the strategy
    eng.set_sequence(b"ACGT" * 1000, 1, [0.0] + [1.0] * 10 + [0.0])
    reads, maf, st, _ = eng.simulate(8000, rng_mode=capi.RNG_PHILOX, seed=1)
    assert b"s ref " in maf
