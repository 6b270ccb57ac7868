"""-m gpu: the C ABI fails loudly and recoverably on misuse (call order, missing inputs, exhausted replay log)."""
import numpy as np
import pytest

from oracle import oracle as O
from pbsim_b200 import capi, simulator
from tests.golden_util import Case
from tests.gpu_util import engine_model

pytestmark = pytest.mark.gpu
BIAS = [0.0] + [1.0] * 10 + [0.0]


def test_call_order_is_enforced_and_the_engine_stays_usable():
    c = Case("qs_rsii_basic")
    eng = simulator.Engine(0)
    try:
        with pytest.raises(simulator.EngineError, match="set_model and set_sequence must precede"):
            eng.begin(1000)
        eng.set_model(engine_model(c))
        with pytest.raises(simulator.EngineError, match="set_model and set_sequence must precede"):
            eng.begin(1000)
        with pytest.raises(simulator.EngineError, match="simulate_begin was not called"):
            eng.next_chunk()
        eng.set_sequence(c.contigs[0][1], 1, BIAS)
        with pytest.raises(simulator.EngineError, match="replay mode needs the draw log"):
            eng.begin(1000, rng_mode=capi.RNG_REPLAY)
        with pytest.raises(simulator.EngineError, match="unknown option"):
            eng.set_option("no_such_option", 1)
        with pytest.raises(simulator.EngineError, match="seg_min_len must be at least"):
            eng.set_option("seg_min_len", 10)
        # host and device delivery cannot be mixed inside a run
        eng.begin(5 * len(c.contigs[0][1]), rng_mode=capi.RNG_PHILOX, seed=1, batch_reads=16)
        assert eng.next_chunk() is not None
        with pytest.raises(simulator.EngineError, match="cannot be mixed"):
            eng.next_chunk(device=True)
        eng.end()
        # ... and after all that a normal run gives the oracle's bytes
        out, _ = c.run_oracle("philox")
        reads, maf, st, _ = eng.simulate(int(c.depth * len(c.contigs[0][1])), rng_mode=capi.RNG_PHILOX, seed=c.seed)
        assert reads == out[0]["reads"] and maf == out[0]["maf"]
    finally:
        eng.close()


def test_replay_log_that_ends_too_early_is_reported():
    c = Case("qs_rsii_basic")
    out, _ = c.run_oracle("glibc")
    draws = O.glibc_rand(c.seed, c.ndraws)
    starts = np.concatenate([[0], c.marks[:-1]]).astype(np.int64)
    n = len(out[0]["info"])
    eng = simulator.Engine(0)
    try:
        eng.set_model(engine_model(c))
        eng.set_sequence(c.contigs[0][1], 1, BIAS)
        with pytest.raises(simulator.EngineError, match="replay"):
            # only the first half of the sub-reads is in the log, the quota needs all of them
            eng.simulate(int(c.depth * len(c.contigs[0][1])), rng_mode=capi.RNG_REPLAY,
                         replay_draws=draws[:int(starts[n // 2])], replay_starts=starts[:n // 2])
    finally:
        eng.close()


def test_create_rejects_a_device_that_does_not_exist():
    with pytest.raises(simulator.EngineError, match="out of range"):
        simulator.Engine(97)
