"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/nanollama_cuda.h declares, and
refuses (loudly) to compute without a GPU — there is no CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nanollama_b200 import build as B
from nanollama_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    B.build()
    return capi.lib()


def test_header_and_binding_agree(lib):
    hdr = open(os.path.join(ROOT, "include", "nanollama_cuda.h")).read()
    declared = set(re.findall(r"^(?:int|void|int64_t|const char \*)\s*(nl_[a-z0-9_]+)\(", hdr, flags=re.M))
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name  # dlsym succeeds
    assert lib.nl_abi_version() == 1


def test_config_struct_layout():
    # 8 int32 + 2 float + 6 int32, no padding: must match struct nl_config
    assert C.sizeof(capi.NlConfig) == 16 * 4


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.nl_device_count() == 0
    src = np.zeros(18, np.uint8)
    dst = np.zeros(32, np.float32)
    rc = lib.nl_dequant(2, capi.ptr(src), 32, capi.ptr(dst))
    assert rc == capi.NL_ERR_CUDA
    assert b"no CPU fallback" in lib.nl_last_error()
    cfg = capi.NlConfig(2, 128, 2, 1, 64, 256, 64, 512, 1e-5, 10000.0, 0, 0, 0, 0, 1, 1)
    h = C.c_void_p()
    assert lib.nl_create(C.byref(cfg), C.byref(h)) == capi.NL_ERR_CUDA and not h.value
    with pytest.raises(capi.NlError):
        from nanollama_b200 import model as M
        M.matmul_dispatch(np.zeros(18 * 4, np.uint8), 2, np.zeros(128, np.float32), 1, 128)


def test_invalid_arguments_rejected_before_touching_cuda(lib):
    h = C.c_void_p()
    assert lib.nl_create(None, C.byref(h)) == capi.NL_ERR_INVALID
    cfg = capi.NlConfig(2, 128, 3, 2, 0, 256, 64, 512, 1e-5, 10000.0, 0, 0, 0, 0, 1, 1)  # 3 heads / 2 kv heads
    assert lib.nl_create(C.byref(cfg), C.byref(h)) == capi.NL_ERR_INVALID
    assert b"n_kv_heads" in lib.nl_last_error()
    assert lib.nl_forward(None, 0, 0, None) == capi.NL_ERR_INVALID
    assert lib.nl_dequant(3, capi.ptr(np.zeros(20, np.uint8)), 32, capi.ptr(np.zeros(32, np.float32))) == capi.NL_ERR_UNSUPPORTED  # Q4_1
