"""CPU-only: the C-ABI library loads without a GPU and exports every symbol include/pbsim_cuda.h declares;
engine creation fails loudly (no CPU fallback) when there is no device."""
import ctypes as C
import os
import re

import pytest

import __graft_entry__ as G
from pbsim_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    G.build_engine()
    return capi.load()


def test_header_symbols_are_exported(lib):
    with open(os.path.join(ROOT, "include", "pbsim_cuda.h")) as f:
        text = f.read()
    names = sorted(set(re.findall(r"\b(pbsim_(?:cuda|host)_[a-z0-9_]+)\s*\(", text)))
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(capi.HOST_EXPORTS + capi.ENGINE_EXPORTS)
    assert lib.pbsim_cuda_abi_version() == 4


def test_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: the header must compile as C99 on its own (no C++, no torch types)"""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "pbsim_cuda.h"\nint main(void) { pbsim_run r; pbsim_chunk c; pbsim_stats s; pbsim_seqset q; '
                   'pbsim_sample_stats t; (void)r; (void)c; (void)s; (void)q; (void)t; return PBSIM_ABI_VERSION == 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                           "-I", os.path.join(ROOT, "include"), str(src)])


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.pbsim_cuda_create(C.byref(h), 0)
    assert rc < 0
    assert b"no CPU fallback" in lib.pbsim_cuda_last_error(None)


def test_host_front_end_reports_reference_errors(lib):
    from tests.golden_util import model_path
    # kappa = mean^2/sd^2 > ~52 overflows the reference's pow() -> "length parameters are not appropriate" (SURVEY B-12)
    p = capi.host_params("qshmm", len_mean=50000.0, len_sd=5000.0)
    with pytest.raises(RuntimeError, match="length parameters are not appropriate"):
        capi.HostModel(lib, p, model_path("QSHMM-RSII.model"))
    with pytest.raises(RuntimeError, match="Cannot open"):
        capi.HostModel(lib, capi.host_params("qshmm"), "/nonexistent.model")


def test_length_parameters_are_checked_for_every_pass_number(lib):
    """the reference tests len_rand_value only with --pass-num 1 (pbsim.cpp:2022, :3664) and crashes otherwise; the
    product's table builder (and the oracle) refuse the parameters whatever the pass number"""
    from oracle import oracle as O
    from tests.golden_util import model_path
    for method, model in (("qshmm", "QSHMM-RSII.model"), ("errhmm", "ERRHMM-ONT.model")):
        for pass_num in (1, 2, 10):
            with pytest.raises(RuntimeError, match="length parameters are not appropriate"):
                capi.HostModel(lib, capi.host_params(method, len_mean=50000.0, len_sd=5000.0, pass_num=pass_num), model_path(model))
            with pytest.raises(RuntimeError, match="length parameters are not appropriate"):
                O.Oracle(method, model_path(model), len_mean=50000.0, len_sd=5000.0, pass_num=pass_num)
