"""Self-test of the CUDA execution-model emulation used by the CPU tests of kernel sources (tests/native/cuda_emu.h)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emuself") / "libemuself.so")
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-shared", "-fPIC", "-I", os.path.join(ROOT, "tests", "native"),
                    os.path.join(ROOT, "tests", "native", "emu_selftest.cpp"), "-o", so], check=True)
    return C.CDLL(so)


def test_block_reduction_over_grid_and_block_sizes(lib):
    x = np.arange(1.0, 1001.0)
    for grid, block in ((1, 32), (3, 64), (7, 256), (2, 1024)):
        out = C.c_double()
        lib.run_block_reduce(x.ctypes.data_as(C.POINTER(C.c_double)), len(x), grid, block, C.byref(out))
        assert out.value == x.sum()


def test_ballot_compaction_with_dynamic_shared_memory(lib):
    rng = np.random.default_rng(0)
    x = rng.integers(-5, 6, 1000).astype(np.int32)
    out = np.zeros(1000, dtype=np.int32)
    cnt = C.c_int()
    lib.run_compact(x.ctypes.data_as(C.POINTER(C.c_int)), len(x), 128, out.ctypes.data_as(C.POINTER(C.c_int)), C.byref(cnt))
    assert cnt.value == int((x > 0).sum())
    assert np.array_equal(out[: cnt.value], x[x > 0])  # blocks run in order, so the global order is kept


def test_shuffle_variants(lib):
    out = np.zeros(64, dtype=np.int32)
    lib.run_shuffle(out.ctypes.data_as(C.POINTER(C.c_int)))
    lane = np.arange(64) % 32
    down = np.where(lane + 1 < 32, lane + 1, lane)
    assert np.array_equal(out, 30 + down + (lane ^ 1))
