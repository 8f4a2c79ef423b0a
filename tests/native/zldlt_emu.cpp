// Compiles quantax_b200/csrc/zldlt.cu for the CPU (tests/native/cuda_emu.h) so that the factorisation and solve
// kernels run as written, with small blocks (kNB = 8), against NumPy in tests/test_zldlt_emu_cpu.py.
// TEST INFRASTRUCTURE ONLY.
#define QTX_HOST_EMULATION 1
#include "zldlt.cu"

extern "C" {
int emu_zldlt_block_size() { return qtx::zldlt_block_size(); }
size_t emu_zldlt_scratch_bytes(int64_t n) { return qtx::zldlt_scratch_bytes(n); }
int emu_zldlt_factor(double* M_c128, int64_t n, void* scratch, int32_t* info) {
  return qtx::zldlt_factor((cuDoubleComplex*)M_c128, n, scratch, info, nullptr);
}
int emu_zldlt_solve(const double* M_c128, int64_t n, double* x_c128, void* scratch) {
  return qtx::zldlt_solve((const cuDoubleComplex*)M_c128, n, (cuDoubleComplex*)x_c128, scratch, nullptr);
}
}
