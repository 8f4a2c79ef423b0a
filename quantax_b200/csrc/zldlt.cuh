// Complex-symmetric LDL^T (no pivoting) and its triangular solves: host entry points of zldlt.cu.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace qtx {
size_t zldlt_scratch_bytes(int64_t n);
// in place on the lower triangle of the row-major M [n, n]: L below the diagonal, D on it; *info (device) receives the
// 1-based index of the first zero pivot (left untouched otherwise)
int zldlt_factor(cuDoubleComplex* M, int64_t n, void* scratch, int32_t* info, cudaStream_t st);
// x <- (L D L^T)^-1 x with the factors (and the scratch) left by zldlt_factor
int zldlt_solve(const cuDoubleComplex* M, int64_t n, cuDoubleComplex* x, void* scratch, cudaStream_t st);
int zldlt_block_size();
}  // namespace qtx
