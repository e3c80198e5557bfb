// lq_assemble.cu -- the step BEFORE the sweep, batched on the device:
// SolverProxDDP::updateLQSubproblem (solvers/proxddp/solver-proxddp.hxx:734-805) fused with
// computeProjectedJacobians (:25-69).  Turns the derivative buffers of every (instance,
// knot) into the packed stage / terminal records the sweep reads -- one pass over HBM,
// so a device-resident caller never moves the knots over PCIe.
//
// Pure streaming work (HBM-bound): a warp per record, grid-stride over 148 x 8 CTAs of 256
// threads; every source block and the destination record are contiguous and walked with
// unit stride, four elements per lane in flight.
#include <cuda_runtime.h>

#include "../../include/aligator_b200/gar.h"
#include "lq_assemble.h"

namespace ab2 {

// active row of the product constraint set: the normal-cone projection Jacobian keeps it
// (core/constraint-set.hxx:25-37 with computeActiveSet of equality-constraint.hpp:52-55,
// negative-orthant.hpp:30-33, box-constraint.hpp:39-43 expressed as one test on [lo, hi])
__device__ __forceinline__ bool row_active(double z, double lo, double hi) { return z > hi || z < lo; }

// One element of the stage record of (instance, knot) `rec`: e indexes
// [A | B | f | Q | S | R | q | r | C | D | d | pad].
__device__ __forceinline__ double stage_element(const ab2_lq_inputs &in, long rec, int t, long inst, int e,
                                                int nx, int nu, int nc) {
  const int nxx = nx * nx, nxu = nx * nu, nuu = nu * nu;
  if (e < nxx) // knot.A = dd.Jx()  (:755)
    return in.Jx[rec * nxx + e];
  e -= nxx;
  if (e < nxu) // knot.B = dd.Ju()
    return in.Ju[rec * nxu + e];
  e -= nxu;
  if (e < nx) // knot.f = dyn_slacks[t+1]
    return in.slack[rec * nx + e];
  e -= nx;
  if (e < nxx) { // knot.Q = Lxx; diag += preg; (+= Hxx, EXACT :770-774) (; += id.Hxx_ on stage 0, :803-804)
    double v = in.Lxx[rec * nxx + e];
    if (e % (nx + 1) == 0)
      v += in.preg;
    if (in.Hxx)
      v += in.Hxx[rec * nxx + e];
    if (t == 0 && in.Hxx0)
      v += in.Hxx0[inst * nxx + e];
    return v;
  }
  e -= nxx;
  if (e < nxu) {
    double v = in.Lxu[rec * nxu + e];
    if (in.Hxu)
      v += in.Hxu[rec * nxu + e];
    return v;
  }
  e -= nxu;
  if (e < nuu) {
    double v = in.Luu[rec * nuu + e];
    if (e % (nu + 1) == 0)
      v += in.preg;
    if (in.Huu)
      v += in.Huu[rec * nuu + e];
    return v;
  }
  e -= nuu;
  if (e < nx + nu) {
    // q = Lxs[t] + cstr_lx_corr, r = Lus[t] + cstr_lu_corr (:764-765, 782-783) with
    // corr = P^T lv - Ptilde^T lv, lv = Lvs * mu_inv: both products over ALL rows, then subtracted (:46-52)
    const bool isx = e < nx;
    const int jj = isx ? e : e - nx;
    double full = 0.0, proj = 0.0;
    if (nc > 0) {
      const double *P = isx ? in.cJx + rec * nc * nx : in.cJu + rec * nc * nu;
      for (int i = 0; i < nc; ++i) {
        const double lv = in.Lv[rec * nc + i] * in.mu_inv;
        const double pij = P[i + (long)jj * nc];
        const double a = row_active(in.shifted[rec * nc + i], in.lo[i], in.hi[i]) ? 1.0 : 0.0;
        full += pij * lv;
        proj += (pij * a) * lv;
      }
    }
    return (isx ? in.Lx[rec * nx + jj] : in.Lu[rec * nu + jj]) + (full - proj);
  }
  e -= nx + nu;
  if (e < nc * nx) { // knot.C = projected Jx: rows of inactive constraints zeroed (:49-50, 778)
    const int i = e % nc;
    return row_active(in.shifted[rec * nc + i], in.lo[i], in.hi[i]) ? in.cJx[rec * nc * nx + e] : 0.0;
  }
  e -= nc * nx;
  if (e < nc * nu) {
    const int i = e % nc;
    return row_active(in.shifted[rec * nc + i], in.lo[i], in.hi[i]) ? in.cJu[rec * nc * nu + e] : 0.0;
  }
  e -= nc * nu;
  if (e < nc) // knot.d = Lvs[t]
    return in.Lv[rec * nc + e];
  return 0.0; // pad to even
}

// A warp per record (grid-stride), lanes stride over the record's elements, UNR elements per
// lane in flight: unit-stride reads of every source block and unit-stride writes of the record.
__global__ void __launch_bounds__(256)
    lq_assemble_stage_kernel(const ab2_lq_inputs in, double *__restrict__ stage, long nrec, int N, int nx,
                             int nu, int nc, int srec) {
  constexpr int UNR = 4;
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long rec = warp; rec < nrec; rec += nwarps) {
    const int t = (int)(rec % N);
    const long inst = rec / N;
    double *dst = stage + rec * srec;
    for (int e0 = 0; e0 < srec; e0 += 32 * UNR) {
      double v[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int e = e0 + 32 * u + lane;
        v[u] = (e < srec) ? stage_element(in, rec, t, inst, e, nx, nu, nc) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int e = e0 + 32 * u + lane;
        if (e < srec)
          dst[e] = v[u];
      }
    }
  }
}

// Terminal knot (:785-795), initial condition (:797-800): a warp per instance.
__global__ void __launch_bounds__(256)
    lq_assemble_term_kernel(const ab2_lq_inputs in, double *__restrict__ term, double *__restrict__ G0,
                            double *__restrict__ g0, int batch, int nx, int nct, int nc0, int trec) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const int nxx = nx * nx;
  for (long b = warp; b < batch; b += nwarps) {
    double *dst = term + b * trec;
    for (int e = lane; e < nxx; e += 32) // knot.Q = tcd.Lxx_; diag += preg
      dst[e] = in.Lxx_N[b * nxx + e] + ((e % (nx + 1) == 0) ? in.preg : 0.0);
    for (int j = lane; j < nx; j += 32) { // knot.q = Lxs[N] + cstr_lx_corr[N]
      double full = 0.0, proj = 0.0;
      for (int i = 0; i < nct; ++i) {
        const double lv = in.Lv_N[b * nct + i] * in.mu_inv;
        const double pij = in.cJx_N[b * nct * nx + i + (long)j * nct];
        const double a = row_active(in.shifted_N[b * nct + i], in.loN[i], in.hiN[i]) ? 1.0 : 0.0;
        full += pij * lv;
        proj += (pij * a) * lv;
      }
      dst[nxx + j] = in.Lx_N[b * nx + j] + (full - proj);
    }
    for (int e = lane; e < nct * nx; e += 32) {
      const int i = e % nct;
      dst[nxx + nx + e] =
          row_active(in.shifted_N[b * nct + i], in.loN[i], in.hiN[i]) ? in.cJx_N[b * nct * nx + e] : 0.0;
    }
    for (int e = lane; e < nct; e += 32)
      dst[nxx + nx + nct * nx + e] = in.Lv_N[b * nct + e];
    for (int e = lane; e < nc0 * nx; e += 32) // prob.G0 = id.Jx(), prob.g0 = id.value_
      G0[b * nc0 * nx + e] = in.G0[b * nc0 * nx + e];
    for (int e = lane; e < nc0; e += 32)
      g0[b * nc0 + e] = in.g0[b * nc0 + e];
  }
}

cudaError_t launch_lq_assemble(const ab2_lq_inputs &in, double *stage, double *term, double *G0, double *g0,
                               int batch, int N, int nx, int nu, int nc, int nct, int nc0, int srec, int trec,
                               cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long full = (long)sms * 8; // 8 CTAs of 256 threads per SM: 2048 threads, the SM's limit
  if (N > 0) {
    const long nrec = (long)batch * N;
    long grid = (nrec + 7) / 8; // 8 warps per CTA
    if (grid > full)
      grid = full;
    lq_assemble_stage_kernel<<<(int)grid, 256, 0, st>>>(in, stage, nrec, N, nx, nu, nc, srec);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
      return e;
  }
  long grid = ((long)batch + 7) / 8;
  if (grid > full)
    grid = full;
  lq_assemble_term_kernel<<<(int)grid, 256, 0, st>>>(in, term, G0, g0, batch, nx, nct, nc0, trec);
  return cudaGetLastError();
}

} // namespace ab2
