// riccati_block_launch.h -- host interface of the CTA-per-instance sweep for run-time
// dimensions (block_kernel.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace ab2 {
struct SweepParams;
// threads per CTA for this shape (0 = a row count exceeds one CTA)
int block_threads(int nx, int nu, int nc, int nc0, int nth = 0);
size_t block_smem_bytes(int nx, int nu, int nc, int nc0, int nth = 0);
bool block_supported(int nx, int nu, int nc, int nc0, int nth = 0);
// info != nullptr: fill {threads, smem, threads, grid, regs, CTAs/SM} instead of launching
cudaError_t launch_block(const SweepParams &p, int nx, int nu, int nc, cudaStream_t st, int *info);
// leg mode (gar::ParallelRiccatiSolver): condensed block-tridiagonal solve per instance, collapseFeedback
bool condensed_supported(int nx, int nc0, int legs);
cudaError_t launch_condensed(const SweepParams &p, int nx, cudaStream_t st);
// the stage-dense solver (gar::RiccatiSolverDense, riccati_dense.cuh)
bool dense_supported(int nx, int nu, int nc, int nct, int nc0);
cudaError_t launch_dense(const SweepParams &p, int nx, int nu, int nc, cudaStream_t st);
cudaError_t launch_collapse(const SweepParams &p, int nx, int nu, int nc, cudaStream_t st);
} // namespace ab2
