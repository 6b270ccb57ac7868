"""GPU: the seeded random configurations of tests/test_fuzz_core_cpu.py through the CUDA engine and the C ABI —
method, model, options, --hp-del-bias, --pass-num, genome features, batch size, segments on / off — against the oracle
in PHILOX mode (FASTQ / SAM, MAF and the statistics block)."""
import pytest

from oracle import oracle as O
from pbsim_b200 import capi, simulator
from tests.golden_util import model_path
from tests.test_fuzz_core_cpu import _config

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", range(24))
def test_random_configuration_engine_equals_oracle(k):
    import numpy as np
    cfg = _config(k)
    okw, genome = cfg["okw"], cfg["genome"]
    path = model_path(cfg["model"]) if cfg["model"] else None
    try:
        o = O.Oracle(cfg["method"], path, **okw)
    except RuntimeError as e:
        pytest.skip(str(e))
    o.rng_philox(cfg["seed"])
    if okw["hp_del_bias"] != 1.0:
        o.hp_bias_prepass([genome])
    o.set_sequence(genome, 1)
    pool = None
    try:
        if cfg["method"] == "sample":
            pool = [bytes(cfg["rng"].integers(33 + 3, 33 + 25, int(n)).astype(np.uint8))
                    for n in cfg["rng"].integers(100, 3000, int(cfg["rng"].integers(4, 40)))]
            want = o.simulate_sample(cfg["depth"], pool)
        else:
            want = o.simulate_wgs(cfg["depth"])
    except RuntimeError as e:
        pytest.skip(str(e))
    batch_reads = int(cfg["rng"].integers(1, 20))
    hm = capi.HostModel(capi.load(), capi.host_params(cfg["method"], **okw), path)
    eng = simulator.Engine(0)
    try:
        eng.set_option("pipeline", k % 3)
        eng.set_option("segments", 1 if cfg["segments"] else 0)
        eng.set_option("seg_min_len", 1024)
        eng.set_option("chain_chunk", cfg["chain_chunk"] or 32)
        eng.set_option("sample_spec", k % 2)
        run = simulator.WgsRun(eng, hm, cfg["depth"], hp_del_bias=okw["hp_del_bias"])
        if pool is not None:
            eng.set_pool(pool)
        if okw["hp_del_bias"] != 1.0:
            run.prepass([("g", genome)])
        reads, maf, st, text = run.simulate_sequence(genome, 1, rng_mode=capi.RNG_PHILOX, seed=cfg["seed"],
                                                     batch_reads=batch_reads)
    finally:
        eng.close()
    assert reads == want[0], "reads differ"
    assert maf == want[1], "maf differs"
    assert text == O.format_stats(want[2], 1)
