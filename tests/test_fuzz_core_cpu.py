"""CPU: seeded random configurations of the hot path — method, model, length / accuracy / ratio options, --hp-del-bias,
--pass-num, genomes with N runs, IUPAC codes, lower case and planted homopolymers — run through the engine's per-read
core (sim_core.cuh compiled by tests/hostsim, the code the CUDA kernels execute) in PHILOX mode, sequential and
segment-parallel, and compared byte for byte with the oracle.  The golden cases pin the oracle to the reference; this
widens the engine == oracle side beyond the goldens' parameters."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import refrun as R
from pbsim_b200 import capi
from tests import hostsim_util as H
from tests.golden_util import model_path

QS = ["QSHMM-RSII.model", "QSHMM-ONT.model", "QSHMM-ONT-HQ.model"]
ER = ["ERRHMM-RSII.model", "ERRHMM-ONT.model", "ERRHMM-ONT-HQ.model", "ERRHMM-SEQUEL.model"]


def _config(k):
    rng = np.random.default_rng(4200 + k)
    method = ["qshmm", "errhmm", "sample"][k % 3]
    okw = dict(len_min=100, len_max=int(rng.integers(3000, 30000)),
               ratio=tuple(int(x) for x in rng.integers(1, 60, 3)),
               hp_del_bias=float(rng.choice([1.0, 1.0, 2.5, 4.0])))
    model = None
    if method != "sample":
        model = str(rng.choice(QS if method == "qshmm" else ER))
        mean = float(rng.integers(1200, 5000))
        okw.update(len_mean=mean, len_sd=float(rng.uniform(0.3, 0.9)) * mean, pass_num=int(rng.choice([1, 1, 2])),
                   accuracy_mean=float(rng.integers(82, 97)) / 100.0, accuracy_mean_set=True)
    genome = R.synth_genome(900 + k, [("g", int(rng.integers(15000, 60000)))], n_runs=int(rng.integers(0, 4)),
                            hp_plants=int(rng.integers(0, 40)), lowercase_frac=float(rng.choice([0.0, 0.1])),
                            iupac=int(rng.integers(0, 5)), long_runs=(13, 25) if k % 4 == 0 else ())[0][1]
    return dict(method=method, model=model, okw=okw, genome=genome, depth=float(rng.uniform(1.5, 4.0)), seed=int(rng.integers(1, 1 << 30)),
                segments=bool(k % 2), chain_chunk=int(rng.choice([0, 1, 4, 32])), rng=rng)


@pytest.mark.parametrize("k", range(48))
def test_random_configuration_core_equals_oracle(k):
    cfg = _config(k)
    okw, genome = cfg["okw"], cfg["genome"]
    try:
        o = O.Oracle(cfg["method"], model_path(cfg["model"]) if cfg["model"] else None, **okw)
    except RuntimeError as e:  # a parameter set the reference itself rejects ("... are not appropriate")
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
    except RuntimeError as e:  # e.g. a drawn accuracy the model cannot serve (the engine reports the same)
        pytest.skip(str(e))
    hm = capi.HostModel(H.lib(), capi.host_params(cfg["method"], **okw), model_path(cfg["model"]) if cfg["model"] else None)
    L = H.lib()
    L.hostsim_use_segments(1 if cfg["segments"] else 0, 1025)
    L.hostsim_set_chain_chunk(cfg["chain_chunk"])
    try:
        sub = H.run(hm, o.seq_upper(), o.hp(), 1, o.bias(), capi.RNG_PHILOX, cfg["seed"], None,
                    int(cfg["depth"] * len(genome)), pool=pool, batch_reads=int(cfg["rng"].integers(1, 20)))
    finally:
        L.hostsim_use_segments(0, 2048)
        L.hostsim_set_chain_chunk(0)
    reads, maf = H.records_from_events(hm, sub, o.seq_upper(), 1)
    assert reads == want[0], "reads differ"
    assert maf == want[1], "maf differs"
    info = o.readinfo()
    assert [s["nsub"] for s in sub] == info["nsub"].tolist()
    assert [s["ndel"] for s in sub] == info["ndel"].tolist()
    assert np.array_equal(np.array([s["accuracy"] for s in sub]), info["accuracy"])
