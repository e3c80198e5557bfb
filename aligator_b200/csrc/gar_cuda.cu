// gar_cuda.cu -- sm_100a kernels + C-ABI implementation (include/aligator_b200/gar.h).
//
// The arithmetic lives in riccati_group.cuh (one group of G lanes per problem
// instance).  This file supplies the device execution context -- TMA bulk copies
// (cp.async.bulk, SASS UBLKCP) completing on per-group mbarriers, or cp.async
// (LDGSTS) staging -- the persistent-sweep kernel, the shape dispatch and the
// host-side handle.  No CPU fallback: every entry point fails loudly without a
// CUDA device.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/aligator_b200/gar.h"
#include "kkt_error.h"
#include "linesearch.h"
#include "lq_assemble.h"
#include "riccati_block_launch.h"
#include "riccati_configs.h"
#include "riccati_launch.cuh"

namespace ab2 {

// [K_0 | k_0] of every instance -> dst [batch][nu][nx+1]
// (head: physical slot of stage knot 0 in the factor arrays, non-zero only between a cycleAppend and the next backward)
__global__ void first_step_policy_kernel(const double *__restrict__ fb, const double *__restrict__ ff,
                                         double *__restrict__ dst, int batch, int N, int nr, int nu, int nx, int head) {
  const int per = nu * (nx + 1);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < (long)batch * per; i += (long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per), e = (int)(i % per), r = e / (nx + 1), c = e % (nx + 1);
    const size_t k0 = ((size_t)b * N + head) * nr + r;
    dst[i] = (c < nx) ? fb[k0 * nx + c] : ff[k0];
  }
}

// ---------------------------------------------------------------------------
// Multi-GPU: the all-gather of the first-step policy FUSED into the kernel that computes it, over
// NVLink peer memory.  Every rank owns a receive buffer [3][world][batch][nu][nx+1] (three slots,
// step s lives in slot s mod 3) + flag words, mapped into every peer by CUDA IPC.
//  * warp-per-instance sweeps (SweepParams::peer_*): the sweep kernel itself stores an instance's
//    [K0 | k0] into slot [s mod 3][r] of EVERY rank's buffer the moment its backward pass reaches
//    knot 0 (posted peer stores over NVLink / NVSwitch that overlap the rest of the sweep; no pack
//    kernel, no NCCL kernel, no staging copy); ab2_gar_policy_allgather then only publishes the
//    step number in every peer's data flag (policy_publish_kernel);
//  * every other kernel (CTA per instance, legs, dense): policy_allgather_kernel packs from the
//    factor arrays and stores each element into every rank's slot, the last CTA publishes.
// Flow control: the WAIT kernel of step s (stream-ordered behind the consumers of step s-1 on the
// consumer's stream) first tells every peer "everything up to s-1 is consumed here" (ack flag),
// then waits for the data flags of step s; the publish kernel of step s holds the producer's
// stream until every peer has acknowledged step s-2 -- what the slot of step s+1 held.  A rank may
// thus run two steps ahead of the slowest consumer, and NO flag is ever awaited inside the
// persistent sweep (it would starve the kernel that writes the flag of an SM slot).
// ---------------------------------------------------------------------------
constexpr int kMaxPeers = 8;
struct PeerPtrs {
  double *buf[kMaxPeers];               // receive buffer of rank w
  unsigned long long *data_flag[kMaxPeers]; // [world] of rank w: step of the last block received from each sender
  unsigned long long *ack_flag[kMaxPeers];  // [world] of rank w: last step each peer has finished consuming
};
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(256)
    policy_allgather_kernel(const double *__restrict__ fb, const double *__restrict__ ff, const PeerPtrs peers,
                            const int world, const int rank, const int batch, const int N, const int nr, const int nu,
                            const int nx, const unsigned long long step, unsigned int *done_counter) {
  const int per = nu * (nx + 1);
  const long total = (long)batch * per;
  if (threadIdx.x == 0) {
    if (step >= 3) // the slot written now held step-3: every peer has consumed it once it acknowledges step-2
      for (int w = 0; w < world; ++w)
        while (ld_acquire_sys(peers.ack_flag[rank] + w) + 2 < step)
          __nanosleep(64);
  }
  __syncthreads();
  const size_t half = (size_t)(step % 3) * world * total; // three slots
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per), e = (int)(i % per), r = e / (nx + 1), c = e % (nx + 1);
    const double v = (c < nx) ? fb[((size_t)b * N * nr + r) * nx + c] : ff[(size_t)b * N * nr + r];
    for (int w = 0; w < world; ++w) // own buffer first-class: w == rank is a local store
      peers.buf[w][half + (size_t)rank * total + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done_counter, 1u);
    if (prev == gridDim.x - 1) { // every CTA's stores are fenced: publish
      *done_counter = 0;
      __threadfence_system();
      for (int w = 0; w < world; ++w)
        st_release_sys(peers.data_flag[w] + rank, step);
    }
  }
}
// the sweep kernel stored this rank's blocks into every peer itself (SweepParams::peer_*): order them
// before the flags (the kernel boundary orders the sweep's stores before this kernel; the fence and the
// releases carry that to system scope)
// It also holds the stream until every peer has consumed step - 2: the NEXT sweep stores into the slot that held
// step - 2 (three slots).  That wait sits here, in a one-warp kernel behind the sweep, never inside the
// persistent sweep: a kernel that fills every SM and spins on a flag which another kernel of the same GPU must
// write (this rank's own wait kernel, on the side stream) can starve that kernel of an SM slot for ever.
__global__ void policy_publish_kernel(const PeerPtrs peers, const int world, const int rank, const unsigned long long step) {
  const int w = threadIdx.x;
  __threadfence_system();
  if (w < world) {
    st_release_sys(peers.data_flag[w] + rank, step);
    if (step >= 3)
      while (ld_acquire_sys(peers.ack_flag[rank] + w) + 2 < step)
        __nanosleep(64);
  }
}
// the stream waits until the blocks of every sender have arrived for `step`; before that it
// acknowledges to every peer that this rank is done with step-1 (everything enqueued earlier on
// this stream -- the consumers of step-1 -- has completed)
__global__ void policy_wait_kernel(const PeerPtrs peers, const int world, const int rank, const unsigned long long step) {
  const int w = threadIdx.x;
  if (w < world) {
    st_release_sys(peers.ack_flag[w] + rank, step - 1);
    while (ld_acquire_sys(peers.data_flag[rank] + w) < step)
      __nanosleep(64);
  }
}

// row-major fb [nr][nx] + ff [nr]  ->  column-major [nr][nx+1] with column 0 = ff, per (instance, knot)
__global__ void gains_kernel(const double *__restrict__ fb, const double *__restrict__ ff, double *__restrict__ dst,
                             long nrec, int nr, int nx, int N, int head) {
  const int per = nr * (nx + 1);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nrec * per; i += (long)gridDim.x * blockDim.x) {
    long rec = i / per; // logical (instance, knot) -> physical slot
    if (head) {
      const long b = rec / N;
      const int t = (int)(rec % N) + head;
      rec = b * N + (t >= N ? t - N : t);
    }
    const int e = (int)(i % per), c = e / nr, r = e % nr; // destination is column-major
    dst[i] = (c == 0) ? ff[rec * nr + r] : fb[(rec * nr + r) * nx + (c - 1)];
  }
}

// ---- FDDP backwardPass bookkeeping around the sweep (solver-fddp.hxx:204-277) ----
// before: slack_t = fs[t+1] (the knot's affine term), G0 = -I, g0 = fs[0]
__global__ void fddp_prep_kernel(const double *__restrict__ fs, double *__restrict__ slack, double *__restrict__ G0,
                                 double *__restrict__ g0, int batch, int N, int nx) {
  const long nS = (long)batch * N * nx, nG = (long)batch * nx * nx, ng = (long)batch * nx;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nS + nG + ng; i += (long)gridDim.x * blockDim.x) {
    if (i < nS) {
      const long b = i / ((long)N * nx), r = i % ((long)N * nx);
      slack[i] = fs[b * (N + 1) * nx + nx + r];
    } else if (i < nS + nG) {
      const long e = (i - nS) % ((long)nx * nx);
      G0[i - nS] = (e % nx == e / nx) ? -1.0 : 0.0;
    } else {
      const long j = i - nS - nG, b = j / nx, c = j % nx;
      g0[j] = fs[b * (N + 1) * nx + c];
    }
  }
}
// after: Vx_i = vx_i + sym(Vxx_i) fs[i] (:219-220, 272-276), then Quuks_i = -(Lu_i + Ju_i^T Vx_{i+1}) = Quu_i k_i (:264)
__global__ void fddp_vx_kernel(const double *__restrict__ Vxx, const double *__restrict__ vx, const double *__restrict__ fs,
                               double *__restrict__ Vx_out, int batch, int N, int nx) {
  const long total = (long)batch * (N + 1) * nx;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long kn = i / nx;
    const int r = (int)(i % nx);
    const double *V = Vxx + kn * nx * nx, *f = fs + kn * nx;
    double acc = 0.0;
    for (int c = 0; c < nx; ++c) // the lower triangle mirrored (selfadjointView<Lower>, :272)
      acc += (r >= c ? V[r + (size_t)c * nx] : V[c + (size_t)r * nx]) * f[c];
    Vx_out[i] = vx[i] + acc;
  }
}
__global__ void fddp_quuks_kernel(const double *__restrict__ Ju, const double *__restrict__ Lu, const double *__restrict__ Vx,
                                  double *__restrict__ out, int batch, int N, int nx, int nu) {
  const long total = (long)batch * N * nu;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long kn = i / nu, b = kn / N, t = kn % N;
    const int c = (int)(i % nu);
    const double *B = Ju + kn * nx * nu + (size_t)c * nx, *v = Vx + (b * (N + 1) + t + 1) * nx;
    double acc = Lu[i];
    for (int r = 0; r < nx; ++r)
      acc += B[r] * v[r];
    out[i] = -acc;
  }
}

// one KernelEntry per compile-time shape, each defined in its own object file
// (kernel_inst.cu compiled with -DAB2_NX=.. -DAB2_NU=.. -DAB2_NC=.. -DAB2_G=..)
#define X(NX, NU, NC, G) extern const KernelEntry kEntry_##NX##_##NU##_##NC;
AB2_FOR_EACH_CONFIG(X)
#undef X
static const KernelEntry *const kTable[] = {
#define X(NX, NU, NC, G) &kEntry_##NX##_##NU##_##NC,
    AB2_FOR_EACH_CONFIG(X)
#undef X
};

static const KernelEntry *find_kernel(int nx, int nu, int nc) {
  for (const KernelEntry *e : kTable)
    if (e->nx == nx && e->nu == nu && e->nc == nc)
      return e;
  return nullptr;
}

} // namespace ab2

// ===========================================================================
// C ABI
// ===========================================================================
namespace {
thread_local std::string g_err;
int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return fail(AB2_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
  } while (0)
} // namespace

struct ab2_gar_solver {
  ab2_gar_dims d;
  const ab2::KernelEntry *k;
  int srec, trec, nr;
  ab2::SweepParams p;
  // owned device storage
  double *own_stage = nullptr, *own_term = nullptr, *own_G0 = nullptr, *own_g0 = nullptr;
  double *own_stage_sym = nullptr; // triangle-packed stage records as uploaded by ab2_gar_sweep_host_sym
  double *gains_tmp = nullptr, *kkt_tmp = nullptr, *theta_dev = nullptr, *ls_tmp = nullptr;
  double *fddp_slack = nullptr, *fddp_G0 = nullptr, *fddp_g0 = nullptr, *fddp_vx = nullptr;
  int nth = 0;  // parameter dimension of the value function outputs (= nx in leg mode)
  int rec_nth = 0; // parameter blocks carried by the knot records (0 in leg mode)
  int legs = 0;    // >= 2: gar::ParallelRiccatiSolver (leg mode)
  bool dense = false; // gar::RiccatiSolverDense (one CTA per instance, stage-dense KKT): FF/FB have nu+nc+2nx rows
  // O(1) cycleAppend: ring heads.  fac_head: physical slot of stage knot 0 in the per-knot FACTOR arrays
  // (FF, FB, VXX, VX) -- non-zero only between a cycle_append and the next backward, which rewrites every slot
  // in place; seen by the getters only.  p.stage_head: the same for the solver-owned copy of the stage records,
  // read by the kernels through stage_slot().
  int fac_head = 0;
  double *cond = nullptr;
  // fused pack + all-gather over peer memory (multi-GPU)
  int pg_world = 0, pg_rank = 0;
  unsigned long long pg_step = 0;
  void *pg_local = nullptr;          // cudaMalloc: [3][world][batch][per] doubles, then flags
  void *pg_peer_base[8] = {};        // IPC-opened bases (own entry = pg_local)
  ab2::PeerPtrs pg_ptrs{};
  unsigned int *pg_done = nullptr;
  size_t pg_buf_doubles = 0;
  unsigned long long pg_pushed_step = 0; // the step whose blocks the last sweep stored into the peers itself
  bool pg_in_sweep = true;               // env AB2_PEER_IN_SWEEP=0: always use the separate pack + store kernel
  double *out[AB2_OUT_COUNT] = {};
  size_t out_doubles[AB2_OUT_COUNT] = {};
  size_t out_rec[AB2_OUT_COUNT] = {};  // doubles per knot (or per instance)
  int out_knots[AB2_OUT_COUNT] = {};   // knots per instance (1 for per-instance arrays)
  int *status = nullptr, *pivstat = nullptr;
  bool have_problem = false, have_backward = false, have_forward = false;
  long launches = 0;
  int variant = -1;
  int group_doubles[4] = {0, 0, 0, 0};
  // ab2_gar_sweep_host: internal streams, one event per stream + a fork event
  static constexpr int kPipeStreams = 4;
  cudaStream_t pipe_stream[kPipeStreams] = {};
  cudaEvent_t pipe_done[kPipeStreams] = {};
  cudaEvent_t pipe_fork = nullptr;
};

static size_t stage_total(const ab2_gar_solver *s) {
  return (size_t)s->d.batch * s->d.horizon * s->srec;
}

extern "C" {

const char *ab2_gar_last_error(void) { return g_err.c_str(); }
const char *ab2_gar_version(void) { return "aligator_b200 gar 0.1 (sm_100a)"; }

size_t ab2_gar_stage_record_doubles(int nx, int nu, int nc) {
  size_t n = 2 * (size_t)nx * nx + 2 * (size_t)nx * nu + (size_t)nu * nu + 2 * (size_t)nx + nu +
             (size_t)nc * (nx + nu + 1);
  return (n + 1) & ~(size_t)1;
}
size_t ab2_gar_term_record_doubles(int nx, int nct) {
  return (size_t)nx * nx + nx + (size_t)nct * nx + nct;
}
size_t ab2_gar_stage_record_doubles_th(int nx, int nu, int nc, int nth) {
  size_t n = 2 * (size_t)nx * nx + 2 * (size_t)nx * nu + (size_t)nu * nu + 2 * (size_t)nx + nu +
             (size_t)nc * (nx + nu + 1) + (size_t)nth * (nx + nu + nc + nth + 1);
  return (n + 1) & ~(size_t)1;
}
size_t ab2_gar_term_record_doubles_th(int nx, int nct, int nth) {
  return ab2_gar_term_record_doubles(nx, nct) + (size_t)nth * (nx + nct + nth + 1);
}
int ab2_gar_supported(int nx, int nu, int nc, int nc0) {
  if (nx < 1 || nu < 1 || nc < 0 || nc0 < 0)
    return 0;
  const ab2::KernelEntry *k = ab2::find_kernel(nx, nu, nc);
  if (k && nx + nc0 <= k->G)
    return 1; // compile-time shape, one warp (or part of one) per instance
  return ab2::block_supported(nx, nu, nc, nc0) ? 2 : 0; // run-time shape, one CTA per instance
}

static int create_impl(const ab2_gar_dims *dims, int nth, int legs, ab2_gar_solver **out);
int ab2_gar_create(const ab2_gar_dims *dims, ab2_gar_solver **out) { return create_impl(dims, 0, 0, out); }
int ab2_gar_create_parametric(const ab2_gar_dims *dims, int nth, ab2_gar_solver **out) {
  return create_impl(dims, nth, 0, out);
}
int ab2_gar_create_dense(const ab2_gar_dims *dims, ab2_gar_solver **out) { return create_impl(dims, 0, -1, out); }
int ab2_gar_create_parallel(const ab2_gar_dims *dims, int num_legs, ab2_gar_solver **out) {
  if (num_legs < 2) // parallel-solver.hxx:42-46 throws "numThreads should be greater than or equal to 2"
    return fail(AB2_ERR_INVALID, "num_legs (" + std::to_string(num_legs) + ") should be greater than or equal to 2");
  if (dims && dims->horizon + 1 < num_legs)
    return fail(AB2_ERR_INVALID, "every leg needs at least one knot: horizon + 1 >= num_legs");
  return create_impl(dims, dims ? dims->nx : 0, num_legs, out);
}

static int create_impl(const ab2_gar_dims *dims, int nth, int legs, ab2_gar_solver **out) {
  if (!dims || !out)
    return fail(AB2_ERR_INVALID, "null argument");
  const ab2_gar_dims &d = *dims;
  if (d.nx < 1 || d.nu < 1 || d.nc < 0 || d.nct < 0 || d.nc0 < 0 || d.horizon < 0 || d.batch < 1 || nth < 0)
    return fail(AB2_ERR_INVALID, "bad dimensions");
  const bool dense = legs < 0; // (legs = -1 selects the stage-dense solver)
  if (dense)
    legs = 0;
  const int rec_nth = legs > 1 ? 0 : nth; // leg mode: plain records, the parameterisation is implicit
  if (dense && !ab2::dense_supported(d.nx, d.nu, d.nc, d.nct, d.nc0))
    return fail(AB2_ERR_UNSUPPORTED, "the stage-dense KKT system (nu + nc + 2 nx rows) does not fit one CTA");
  // compile-time shapes run one warp (or part of one) per instance; every other shape runs
  // the CTA-per-instance kernel with run-time dimensions (block_kernel.cu)
  const ab2::KernelEntry *k = ab2::find_kernel(d.nx, d.nu, d.nc);
  if (k && (d.nx + d.nc0 > k->G || nth > 0))
    k = nullptr; // parametric problems run the CTA-per-instance kernel
  if (legs > 1 && !ab2::condensed_supported(d.nx, d.nc0, legs))
    return fail(AB2_ERR_UNSUPPORTED, "the condensed system of " + std::to_string(legs) + " legs does not fit one CTA's shared memory");
  if (!k && !ab2::block_supported(d.nx, d.nu, d.nc, d.nc0, nth))
    return fail(AB2_ERR_UNSUPPORTED,
                "(nx,nu,nc,nc0) = (" + std::to_string(d.nx) + "," + std::to_string(d.nu) + "," +
                    std::to_string(d.nc) + "," + std::to_string(d.nc0) +
                    ") does not fit one CTA (227 KB of shared memory, 256 rows)");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (d.device < 0 || d.device >= ndev)
    return fail(AB2_ERR_CUDA, "no such CUDA device");
  CUDA_TRY(cudaSetDevice(d.device));
  auto *s = new ab2_gar_solver();
  s->d = d;
  s->k = k;
  s->nth = nth;
  s->rec_nth = rec_nth;
  s->legs = legs;
  s->srec = (int)ab2_gar_stage_record_doubles_th(d.nx, d.nu, d.nc, rec_nth);
  if (k && k->srec_pad != s->srec) {
    delete s;
    return fail(AB2_ERR_INVALID, "internal: record size mismatch");
  }
  s->trec = (int)ab2_gar_term_record_doubles_th(d.nx, d.nct, rec_nth);
  s->nr = d.nu + d.nc + d.nx;
  s->dense = dense;
  if (dense) {
    s->k = nullptr;
    s->nr = d.nu + d.nc + 2 * d.nx; // rows of ff / fb: [k; z; l; y] (dense-kernel.hpp:28-31)
  }
  if (s->k)
    s->k->group_doubles(d.nc0, s->group_doubles);
  const int N = d.horizon, B = d.batch, nx = d.nx;
  auto setup = [&](int what, size_t rec, int knots) {
    s->out_rec[what] = rec;
    s->out_knots[what] = knots;
    s->out_doubles[what] = (size_t)B * knots * rec;
  };
  setup(AB2_OUT_FF, s->nr, N);
  setup(AB2_OUT_FB, (size_t)s->nr * nx, N);
  setup(AB2_OUT_VXX, (size_t)nx * nx, N + 1);
  setup(AB2_OUT_VX, nx, N + 1);
  setup(AB2_OUT_FFT, d.nct, 1);
  setup(AB2_OUT_FBT, (size_t)d.nct * nx, 1);
  setup(AB2_OUT_KKT0, nx + d.nc0, 1);
  setup(AB2_OUT_XS, nx, N + 1);
  setup(AB2_OUT_US, d.nu, N);
  setup(AB2_OUT_VS, d.nc, N);
  setup(AB2_OUT_VST, d.nct, 1);
  setup(AB2_OUT_LBD0, d.nc0, 1);
  setup(AB2_OUT_LBDAS, nx, N);
  setup(AB2_OUT_FTH, (size_t)s->nr * nth, N);
  setup(AB2_OUT_VXT, (size_t)nx * nth, N + 1);
  setup(AB2_OUT_VTT, (size_t)nth * nth, N + 1);
  setup(AB2_OUT_VT, nth, N + 1);
  setup(AB2_OUT_KKT0FTH, (size_t)(nx + d.nc0) * nth, 1);
  setup(AB2_OUT_THGRAD, nth, 1);
  setup(AB2_OUT_THHESS, (size_t)nth * nth, 1);
  for (int w = 0; w < AB2_OUT_COUNT; ++w) {
    // (+2: the forward pass of the CTA-per-instance kernel fetches odd-sized gain records
    // with 16-byte granularity, up to one double past the end of the array)
    const size_t bytes = (s->out_doubles[w] > 0 ? s->out_doubles[w] + 2 : 2) * sizeof(double);
    cudaError_t e = cudaMalloc(&s->out[w], bytes);
    if (e == cudaSuccess)
      e = cudaMemset(s->out[w], 0, bytes);
    if (e != cudaSuccess) {
      ab2_gar_destroy(s);
      return fail(AB2_ERR_CUDA, std::string("cudaMalloc outputs: ") + cudaGetErrorString(e));
    }
  }
  {
    cudaError_t e = cudaMalloc(&s->status, sizeof(int) * B);
    if (e == cudaSuccess)
      e = cudaMemset(s->status, 0, sizeof(int) * B);
    if (e != cudaSuccess) {
      ab2_gar_destroy(s);
      return fail(AB2_ERR_CUDA, std::string("cudaMalloc status: ") + cudaGetErrorString(e));
    }
  }
  {
    cudaError_t e = cudaMalloc(&s->pivstat, sizeof(int) * B);
    if (e == cudaSuccess)
      e = cudaMemset(s->pivstat, 0, sizeof(int) * B);
    if (e != cudaSuccess) {
      ab2_gar_destroy(s);
      return fail(AB2_ERR_CUDA, std::string("cudaMalloc pivstat: ") + cudaGetErrorString(e));
    }
  }
  if (legs > 1) {
    const size_t td = (size_t)d.nc0 + (size_t)d.nx * (2 * legs - 1);
    cudaError_t e = cudaMalloc(&s->cond, sizeof(double) * B * td);
    if (e == cudaSuccess)
      e = cudaMemset(s->cond, 0, sizeof(double) * B * td);
    if (e != cudaSuccess) {
      ab2_gar_destroy(s);
      return fail(AB2_ERR_CUDA, std::string("cudaMalloc cond: ") + cudaGetErrorString(e));
    }
  }
  ab2::SweepParams &p = s->p;
  std::memset(&p, 0, sizeof(p));
  p.legs = legs;
  p.cond = s->cond;
  p.N = N;
  p.nct = d.nct;
  p.nc0 = d.nc0;
  p.batch = B;
  p.ff = s->out[AB2_OUT_FF];
  p.fb = s->out[AB2_OUT_FB];
  p.Vxx = s->out[AB2_OUT_VXX];
  p.vx = s->out[AB2_OUT_VX];
  p.ffT = s->out[AB2_OUT_FFT];
  p.fbT = s->out[AB2_OUT_FBT];
  p.kkt0 = s->out[AB2_OUT_KKT0];
  p.xs = s->out[AB2_OUT_XS];
  p.us = s->out[AB2_OUT_US];
  p.vs = s->out[AB2_OUT_VS];
  p.vsT = s->out[AB2_OUT_VST];
  p.lbd0 = s->out[AB2_OUT_LBD0];
  p.lbdas = s->out[AB2_OUT_LBDAS];
  p.status = s->status;
  p.pivstat = s->pivstat;
  p.nth = nth;
  p.theta = nullptr;
  p.fth = s->out[AB2_OUT_FTH];
  p.Vxt = s->out[AB2_OUT_VXT];
  p.Vtt = s->out[AB2_OUT_VTT];
  p.vt = s->out[AB2_OUT_VT];
  p.kkt0fth = s->out[AB2_OUT_KKT0FTH];
  p.thGrad = s->out[AB2_OUT_THGRAD];
  p.thHess = s->out[AB2_OUT_THHESS];
  {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d.device);
    p.num_sms = sms > 0 ? sms : 148;
  }
  if (const char *f = std::getenv("AB2_PEER_IN_SWEEP")) // 0: the exchange always runs as its own pack + store kernel
    s->pg_in_sweep = std::atoi(f) != 0;
  if (const char *f = std::getenv("AB2_DEBUG_FLAGS")) // experiment switches, see SweepParams::dbg
    p.dbg = std::atoi(f);
  if (std::getenv("AB2_PHASE_CLOCKS")) { // profiling aid of the CTA-per-instance kernel: 16 phase counters
    cudaMalloc(&p.clk, 16 * sizeof(long long));
    cudaMemset(p.clk, 0, 16 * sizeof(long long));
  }
  *out = s;
  return AB2_OK;
}

int ab2_gar_destroy(ab2_gar_solver *s) {
  if (!s)
    return AB2_OK;
  cudaSetDevice(s->d.device);
  for (int w = 0; w < AB2_OUT_COUNT; ++w)
    if (s->out[w])
      cudaFree(s->out[w]);
  if (s->status)
    cudaFree(s->status);
  if (s->pivstat)
    cudaFree(s->pivstat);
  for (int w = 0; w < s->pg_world; ++w)
    if (w != s->pg_rank && s->pg_peer_base[w])
      cudaIpcCloseMemHandle(s->pg_peer_base[w]);
  if (s->pg_local)
    cudaFree(s->pg_local);
  if (s->pg_done)
    cudaFree(s->pg_done);
  for (double *q : {s->own_stage_sym, s->own_stage, s->own_term, s->own_G0, s->own_g0, s->gains_tmp, s->kkt_tmp, s->theta_dev, s->cond, s->ls_tmp, s->fddp_slack, s->fddp_G0,
                    s->fddp_g0, s->fddp_vx})
    if (q)
      cudaFree(q);
  for (int i = 0; i < ab2_gar_solver::kPipeStreams; ++i) {
    if (s->pipe_done[i])
      cudaEventDestroy(s->pipe_done[i]);
    if (s->pipe_stream[i])
      cudaStreamDestroy(s->pipe_stream[i]);
  }
  if (s->pipe_fork)
    cudaEventDestroy(s->pipe_fork);
  delete s;
  return AB2_OK;
}

int ab2_gar_set_tuning(ab2_gar_solver *s, const ab2_gar_tuning *t) {
  if (!s || !t)
    return fail(AB2_ERR_INVALID, "null argument");
  if (t->variant < -1 || t->variant > 10)
    return fail(AB2_ERR_INVALID, "variant must be -1 (default) or 0..10");
  if (t->variant == 9 && !ab2::block_supported(s->d.nx, s->d.nu, s->d.nc, s->d.nc0))
    return fail(AB2_ERR_UNSUPPORTED, "variant 9 (CTA per instance) does not fit this shape");
  if (t->stagger_ns < 0 || t->stagger_ns > 100000 || t->ctas_per_sm < 0 || t->ctas_per_sm > 32)
    return fail(AB2_ERR_INVALID, "stagger_ns must be in [0, 100000], ctas_per_sm in [0, 32]");
  s->variant = t->variant;
  s->p.stagger_ns = t->stagger_ns;
  s->p.ctas_per_sm = t->ctas_per_sm;
  return AB2_OK;
}

int ab2_gar_set_problem(ab2_gar_solver *s, const double *stage, const double *term, const double *G0,
                        const double *g0, int memspace, void *stream) {
  if (!s)
    return fail(AB2_ERR_INVALID, "null solver");
  CUDA_TRY(cudaSetDevice(s->d.device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n_stage = stage_total(s), n_term = (size_t)s->d.batch * s->trec,
               n_G0 = (size_t)s->d.batch * s->d.nc0 * s->d.nx, n_g0 = (size_t)s->d.batch * s->d.nc0;
  if (memspace == AB2_DEVICE) {
    if (stage) {
      s->p.stage = stage;
      s->p.stage_head = 0; // a caller-owned array is in knot order (the caller rotates it itself)
    }
    if (term)
      s->p.term = term;
    if (G0)
      s->p.G0 = G0;
    if (g0)
      s->p.g0 = g0;
  } else if (memspace == AB2_HOST) {
    auto up = [&](double *&own, const double *src, size_t n, const double *&dst) -> int {
      if (!src)
        return AB2_OK;
      if (!own)
        CUDA_TRY(cudaMalloc(&own, (n > 0 ? n : 1) * sizeof(double)));
      if (n)
        CUDA_TRY(cudaMemcpyAsync(own, src, n * sizeof(double), cudaMemcpyHostToDevice, st));
      dst = own;
      return AB2_OK;
    };
    int rc;
    if ((rc = up(s->own_stage, stage, n_stage, s->p.stage)) != AB2_OK)
      return rc;
    if (stage)
      s->p.stage_head = 0;
    if ((rc = up(s->own_term, term, n_term, s->p.term)) != AB2_OK)
      return rc;
    if ((rc = up(s->own_G0, G0, n_G0, s->p.G0)) != AB2_OK)
      return rc;
    if ((rc = up(s->own_g0, g0, n_g0, s->p.g0)) != AB2_OK)
      return rc;
  } else {
    return fail(AB2_ERR_INVALID, "memspace must be AB2_HOST or AB2_DEVICE");
  }
  s->have_problem = s->p.stage && s->p.term && (s->p.G0 || s->d.nc0 == 0) && (s->p.g0 || s->d.nc0 == 0);
  if (s->d.horizon == 0 && s->p.term)
    s->have_problem = true;
  return AB2_OK;
}

// SweepParams of the instances [b0, b0 + nb): every array leads with the batch index.
static ab2::SweepParams slice_params(const ab2_gar_solver *s, int b0, int nb) {
  ab2::SweepParams q = s->p;
  const size_t b = (size_t)b0;
  const int N = s->d.horizon, nx = s->d.nx;
  q.batch = nb;
  q.stage += b * N * s->srec;
  q.term += b * s->trec;
  if (q.G0)
    q.G0 += b * s->d.nc0 * nx;
  if (q.g0)
    q.g0 += b * s->d.nc0;
  q.ff += b * s->out_knots[AB2_OUT_FF] * s->out_rec[AB2_OUT_FF];
  q.fb += b * s->out_knots[AB2_OUT_FB] * s->out_rec[AB2_OUT_FB];
  q.Vxx += b * s->out_knots[AB2_OUT_VXX] * s->out_rec[AB2_OUT_VXX];
  q.vx += b * s->out_knots[AB2_OUT_VX] * s->out_rec[AB2_OUT_VX];
  q.ffT += b * s->out_rec[AB2_OUT_FFT];
  q.fbT += b * s->out_rec[AB2_OUT_FBT];
  q.kkt0 += b * s->out_rec[AB2_OUT_KKT0];
  q.xs += b * s->out_knots[AB2_OUT_XS] * s->out_rec[AB2_OUT_XS];
  q.us += b * s->out_knots[AB2_OUT_US] * s->out_rec[AB2_OUT_US];
  q.vs += b * s->out_knots[AB2_OUT_VS] * s->out_rec[AB2_OUT_VS];
  q.vsT += b * s->out_rec[AB2_OUT_VST];
  q.lbd0 += b * s->out_rec[AB2_OUT_LBD0];
  q.lbdas += b * s->out_knots[AB2_OUT_LBDAS] * s->out_rec[AB2_OUT_LBDAS];
  q.status += b;
  q.pivstat += b;
  if (s->nth > 0) {
    q.fth += b * s->out_knots[AB2_OUT_FTH] * s->out_rec[AB2_OUT_FTH];
    q.Vxt += b * s->out_knots[AB2_OUT_VXT] * s->out_rec[AB2_OUT_VXT];
    q.Vtt += b * s->out_knots[AB2_OUT_VTT] * s->out_rec[AB2_OUT_VTT];
    q.vt += b * s->out_knots[AB2_OUT_VT] * s->out_rec[AB2_OUT_VT];
    q.kkt0fth += b * s->out_rec[AB2_OUT_KKT0FTH];
    q.thGrad += b * s->out_rec[AB2_OUT_THGRAD];
    q.thHess += b * s->out_rec[AB2_OUT_THHESS];
    if (q.theta)
      q.theta += b * s->nth;
  }
  if (q.cond)
    q.cond += b * ((size_t)s->d.nc0 + (size_t)nx * (2 * s->legs - 1));
  return q;
}

// The kernels of one (sub-)batch.  Serial solver: ONE launch (backward and/or forward).  Leg mode
// (gar::ParallelRiccatiSolver): the legs' backward recursions of all instances in one launch, the
// condensed block-tridiagonal systems in a second, the legs' rollouts in a third
// (parallel-solver.hxx:150-164, 166-203, 221-241).
static int run_kernels(ab2_gar_solver *s, ab2::SweepParams q, int bwd, int fwd, cudaStream_t st) {
  if (s->legs > 1) {
    if (bwd) {
      CUDA_TRY(cudaMemsetAsync(q.status, 0, sizeof(int) * q.batch, st)); // the legs OR / add into these
      CUDA_TRY(cudaMemsetAsync(q.pivstat, 0, sizeof(int) * q.batch, st));
      q.do_bwd = 1;
      q.do_fwd = 0;
      CUDA_TRY(ab2::launch_block(q, s->d.nx, s->d.nu, s->d.nc, st, nullptr));
      CUDA_TRY(ab2::launch_condensed(q, s->d.nx, st));
      s->launches += 2;
    }
    if (fwd) {
      q.do_bwd = 0;
      q.do_fwd = 1;
      CUDA_TRY(ab2::launch_block(q, s->d.nx, s->d.nu, s->d.nc, st, nullptr));
      s->launches += 1;
    }
    return AB2_OK;
  }
  q.do_bwd = bwd;
  q.do_fwd = fwd;
  if (s->dense)
    CUDA_TRY(ab2::launch_dense(q, s->d.nx, s->d.nu, s->d.nc, st));
  else if (s->k && s->variant != 9)
    CUDA_TRY(s->k->launch(q, s->variant, s->group_doubles, st, nullptr));
  else
    CUDA_TRY(ab2::launch_block(q, s->d.nx, s->d.nu, s->d.nc, st, nullptr));
  s->launches += 1;
  return AB2_OK;
}

static int launch(ab2_gar_solver *s, double mueq, int bwd, int fwd, void *stream) {
  if (!s)
    return fail(AB2_ERR_INVALID, "null solver");
  if (!s->have_problem)
    return fail(AB2_ERR_STATE, "set_problem has not been called with all four buffers");
  if (fwd && !bwd && !s->have_backward)
    return fail(AB2_ERR_STATE, "forward() before backward()");
  if (bwd && !(mueq > 0.0) && (s->d.nc > 0 || s->d.nct > 0))
    return fail(AB2_ERR_INVALID, "mueq must be > 0 when constraints are present");
  CUDA_TRY(cudaSetDevice(s->d.device));
  s->p.mueq = mueq;
  s->p.do_bwd = bwd;
  s->p.do_fwd = fwd;
  ab2::SweepParams q = s->p;
  q.peer_world = 0;
  if (bwd && s->pg_in_sweep && s->pg_peer_base[0] && s->k && s->variant != 9 && s->legs <= 1 && !s->dense &&
      s->d.horizon > 0) {
    // sharded batch: the warp-per-instance sweep stores each instance's [K0 | k0] into every rank's receive
    // buffer as soon as its backward pass is done (the exchange of step pg_step + 1)
    const unsigned long long step = s->pg_step + 1;
    const size_t total = (size_t)s->d.batch * s->d.nu * (s->d.nx + 1);
    q.peer_world = s->pg_world;
    for (int w = 0; w < s->pg_world; ++w)
      q.peer_dst[w] = s->pg_ptrs.buf[w];
    q.peer_off = (long long)((step % 3) * s->pg_world * total + (size_t)s->pg_rank * total);
    s->pg_pushed_step = step;
  }
  if (int rc = run_kernels(s, q, bwd, fwd, (cudaStream_t)stream))
    return rc;
  if (bwd) {
    s->have_backward = true;
    s->fac_head = 0; // every factor slot was rewritten in knot order
  }
  s->have_forward = fwd != 0; // a backward-only launch invalidates the previous trajectory
  return AB2_OK;
}

int ab2_gar_backward(ab2_gar_solver *s, double mueq, void *stream) { return launch(s, mueq, 1, 0, stream); }
int ab2_gar_forward(ab2_gar_solver *s, void *stream) { return launch(s, s ? s->p.mueq : 0.0, 0, 1, stream); }
int ab2_gar_sweep(ab2_gar_solver *s, double mueq, void *stream) { return launch(s, mueq, 1, 1, stream); }
int ab2_gar_forward_theta(ab2_gar_solver *s, const double *theta, int memspace, void *stream) {
  if (!s)
    return fail(AB2_ERR_INVALID, "null solver");
  if (theta && (s->nth == 0 || s->legs > 1))
    return fail(AB2_ERR_INVALID, "theta given to a solver without parameters (nth = 0; the parallel solver ignores theta, parallel-solver.hxx:211)");
  CUDA_TRY(cudaSetDevice(s->d.device));
  s->p.theta = nullptr;
  if (theta) {
    if (memspace == AB2_DEVICE) {
      s->p.theta = theta;
    } else {
      if (!s->theta_dev)
        CUDA_TRY(cudaMalloc(&s->theta_dev, (size_t)s->d.batch * s->nth * sizeof(double)));
      CUDA_TRY(cudaMemcpyAsync(s->theta_dev, theta, (size_t)s->d.batch * s->nth * sizeof(double),
                               cudaMemcpyHostToDevice, (cudaStream_t)stream));
      s->p.theta = s->theta_dev;
    }
  }
  const int rc = launch(s, s->p.mueq, 0, 1, stream);
  s->p.theta = nullptr;
  return rc;
}

int ab2_gar_assemble(ab2_gar_solver *s, const ab2_lq_inputs *in, void *stream) {
  if (!s || !in)
    return fail(AB2_ERR_INVALID, "null argument");
  if (s->rec_nth > 0)
    return fail(AB2_ERR_UNSUPPORTED, "assemble: parametric problems (nth > 0) are not supported");
  const ab2_gar_dims &d = s->d;
  const bool stage_ok = d.horizon == 0 || (in->Jx && in->Ju && in->slack && in->Lxx && in->Lxu && in->Luu &&
                                           in->Lx && in->Lu);
  const bool cstr_ok = d.nc == 0 || d.horizon == 0 || (in->cJx && in->cJu && in->Lv && in->shifted && in->lo && in->hi);
  const bool hess_ok = (!in->Hxx && !in->Hxu && !in->Huu) || (in->Hxx && in->Hxu && in->Huu);
  const bool term_ok = in->Lxx_N && in->Lx_N &&
                       (d.nct == 0 || (in->cJx_N && in->Lv_N && in->shifted_N && in->loN && in->hiN));
  const bool init_ok = d.nc0 == 0 || (in->G0 && in->g0);
  if (!stage_ok || !cstr_ok || !hess_ok || !term_ok || !init_ok)
    return fail(AB2_ERR_INVALID, "ab2_lq_inputs: a required array is NULL for these dimensions");
  CUDA_TRY(cudaSetDevice(d.device));
  auto own = [&](double *&buf, size_t n) -> int {
    if (!buf)
      CUDA_TRY(cudaMalloc(&buf, (n > 0 ? n : 1) * sizeof(double)));
    return AB2_OK;
  };
  int rc;
  if ((rc = own(s->own_stage, stage_total(s))) != AB2_OK || (rc = own(s->own_term, (size_t)d.batch * s->trec)) != AB2_OK ||
      (rc = own(s->own_G0, (size_t)d.batch * d.nc0 * d.nx)) != AB2_OK || (rc = own(s->own_g0, (size_t)d.batch * d.nc0)) != AB2_OK)
    return rc;
  CUDA_TRY(ab2::launch_lq_assemble(*in, s->own_stage, s->own_term, s->own_G0, s->own_g0, d.batch, d.horizon, d.nx,
                                   d.nu, d.nc, d.nct, d.nc0, s->srec, s->trec, (cudaStream_t)stream));
  s->launches += d.horizon > 0 ? 2 : 1;
  s->p.stage = s->own_stage;
  s->p.stage_head = 0;
  s->p.term = s->own_term;
  s->p.G0 = s->own_G0;
  s->p.g0 = s->own_g0;
  s->have_problem = true;
  return AB2_OK;
}

static int copy_ring(const double *base, size_t rec, int knots, int nring, int head, int b0, int nb, int t0, int nt,
                     double *dst, cudaMemcpyKind kind, cudaStream_t st);

int ab2_gar_problem_ptr(ab2_gar_solver *s, int what, const double **out) {
  if (!s || !out || what < 0 || what > 3)
    return fail(AB2_ERR_INVALID, "bad argument");
  const double *ptrs[4] = {s->p.stage, s->p.term, s->p.G0, s->p.g0};
  *out = ptrs[what];
  return AB2_OK;
}

int ab2_gar_get_problem(ab2_gar_solver *s, int what, double *dst, int memspace, void *stream) {
  if (!s || !dst || what < 0 || what > 3)
    return fail(AB2_ERR_INVALID, "bad argument");
  if (!s->have_problem)
    return fail(AB2_ERR_STATE, "no problem set");
  const double *ptrs[4] = {s->p.stage, s->p.term, s->p.G0, s->p.g0};
  const size_t n[4] = {stage_total(s), (size_t)s->d.batch * s->trec, (size_t)s->d.batch * s->d.nc0 * s->d.nx,
                       (size_t)s->d.batch * s->d.nc0};
  CUDA_TRY(cudaSetDevice(s->d.device));
  if (what == 0 && s->p.stage_head != 0 && n[0]) // the solver-owned copy after cycle_append: knot order through the head
    return copy_ring(s->p.stage, (size_t)s->srec, s->d.horizon, s->d.horizon, s->p.stage_head, 0, s->d.batch, 0,
                     s->d.horizon, dst, memspace == AB2_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                     (cudaStream_t)stream);
  if (n[what])
    CUDA_TRY(cudaMemcpyAsync(dst, ptrs[what], n[what] * sizeof(double),
                             memspace == AB2_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
  return AB2_OK;
}

// ---- symmetric blocks travel as lower triangles (ab2_gar_sweep_host_sym) ----
// source offset, in the triangle-packed record, of element e of the full record [A|B|f|Q|S|R|q|r|C|D|d]
static inline __host__ __device__ int sym_source(int e, int nx, int nu, int nc) {
  const int off_q = nx * nx + nx * nu + nx, tq = nx * (nx + 1) / 2, tr = nu * (nu + 1) / 2;
  if (e < off_q)
    return e;
  e -= off_q;
  if (e < nx * nx) { // Q(i, j), column-major; the lower triangle column by column: column j holds rows j..nx-1
    int i = e % nx, j = e / nx;
    if (i < j) {
      const int t = i;
      i = j;
      j = t;
    }
    return off_q + j * nx - j * (j - 1) / 2 + (i - j);
  }
  e -= nx * nx;
  if (e < nx * nu)
    return off_q + tq + e;
  e -= nx * nu;
  if (e < nu * nu) {
    int i = e % nu, j = e / nu;
    if (i < j) {
      const int t = i;
      i = j;
      j = t;
    }
    return off_q + tq + nx * nu + j * nu - j * (j - 1) / 2 + (i - j);
  }
  e -= nu * nu;
  return off_q + tq + nx * nu + tr + e; // q, r, C, D, d
}
namespace ab2 {
// full records from triangle-packed ones: HBM -> HBM, a table look-up per element (built once per CTA)
__global__ void __launch_bounds__(256) expand_sym_kernel(const double *__restrict__ src, double *__restrict__ dst, const long nrec,
                                                         const int nx, const int nu, const int nc, const int srec_full,
                                                         const int srec_pad, const int srec_sym) {
  extern __shared__ int lut[];
  for (int e = threadIdx.x; e < srec_pad; e += blockDim.x)
    lut[e] = e < srec_full ? sym_source(e, nx, nu, nc) : -1;
  __syncthreads();
  const long total = nrec * srec_pad;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / srec_pad;
    const int e = (int)(i - r * srec_pad), o = lut[e];
    dst[i] = o >= 0 ? src[r * srec_sym + o] : 0.0;
  }
}
} // namespace ab2

size_t ab2_gar_stage_record_doubles_sym(int nx, int nu, int nc) {
  if (nx < 1 || nu < 0 || nc < 0)
    return 0;
  return (size_t)nx * nx + (size_t)nx * nu + nx + (size_t)nx * (nx + 1) / 2 + (size_t)nx * nu + (size_t)nu * (nu + 1) / 2 + nx + nu +
         (size_t)nc * nx + (size_t)nc * nu + nc;
}

int ab2_gar_pack_stage_sym(int nx, int nu, int nc, const double *stage, double *stage_sym, long nrec) {
  if (!stage || !stage_sym || nrec < 0 || nx < 1 || nu < 0 || nc < 0)
    return fail(AB2_ERR_INVALID, "bad argument");
  const int full = nx * nx + nx * nu + nx + nx * nx + nx * nu + nu * nu + nx + nu + nc * nx + nc * nu + nc;
  const size_t pad = ab2_gar_stage_record_doubles(nx, nu, nc), sym = ab2_gar_stage_record_doubles_sym(nx, nu, nc);
  std::vector<int> dst_of(sym, 0); // one full-record element per packed slot (the lower-triangle one)
  for (int e = full - 1; e >= 0; --e) {
    const int off_q = nx * nx + nx * nu + nx;
    int keep = 1;
    int r = e - off_q;
    if (r >= 0 && r < nx * nx)
      keep = (r % nx) >= (r / nx);
    r -= nx * nx + nx * nu;
    if (r >= 0 && r < nu * nu)
      keep = (r % nu) >= (r / nu);
    if (keep)
      dst_of[sym_source(e, nx, nu, nc)] = e;
  }
  for (long k = 0; k < nrec; ++k)
    for (size_t o = 0; o < sym; ++o)
      stage_sym[(size_t)k * sym + o] = stage[(size_t)k * pad + dst_of[o]];
  return AB2_OK;
}

static int sweep_host_impl(ab2_gar_solver *s, const double *stage, const double *term, const double *G0,
                           const double *g0, double mueq, int nchunks, const int *whats,
                           double *const *dsts, int nwhat, void *stream, bool sym);
int ab2_gar_sweep_host(ab2_gar_solver *s, const double *stage, const double *term, const double *G0,
                       const double *g0, double mueq, int nchunks, const int *whats,
                       double *const *dsts, int nwhat, void *stream) {
  return sweep_host_impl(s, stage, term, G0, g0, mueq, nchunks, whats, dsts, nwhat, stream, false);
}
int ab2_gar_sweep_host_sym(ab2_gar_solver *s, const double *stage_sym, const double *term, const double *G0,
                           const double *g0, double mueq, int nchunks, const int *whats,
                           double *const *dsts, int nwhat, void *stream) {
  if (s && (s->rec_nth > 0 || s->legs > 1 || s->dense))
    return fail(AB2_ERR_UNSUPPORTED, "sweep_host_sym: plain (nth = 0) serial handles only");
  return sweep_host_impl(s, stage_sym, term, G0, g0, mueq, nchunks, whats, dsts, nwhat, stream, true);
}

static int sweep_host_impl(ab2_gar_solver *s, const double *stage, const double *term, const double *G0,
                           const double *g0, double mueq, int nchunks, const int *whats,
                           double *const *dsts, int nwhat, void *stream, const bool sym) {
  if (!s || !stage || !term || (s->d.nc0 > 0 && (!G0 || !g0)) || nwhat < 0 || (nwhat > 0 && (!whats || !dsts)))
    return fail(AB2_ERR_INVALID, "bad argument");
  for (int i = 0; i < nwhat; ++i)
    if (whats[i] < 0 || whats[i] >= AB2_OUT_COUNT || !dsts[i])
      return fail(AB2_ERR_INVALID, "bad output selector");
  if (!(mueq > 0.0) && (s->d.nc > 0 || s->d.nct > 0))
    return fail(AB2_ERR_INVALID, "mueq must be > 0 when constraints are present");
  CUDA_TRY(cudaSetDevice(s->d.device));
  const int B = s->d.batch, N = s->d.horizon, nx = s->d.nx, nc0 = s->d.nc0;
  constexpr int NS = ab2_gar_solver::kPipeStreams;
  if (!s->pipe_fork) {
    CUDA_TRY(cudaEventCreateWithFlags(&s->pipe_fork, cudaEventDisableTiming));
    for (int i = 0; i < NS; ++i) {
      CUDA_TRY(cudaStreamCreateWithFlags(&s->pipe_stream[i], cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&s->pipe_done[i], cudaEventDisableTiming));
    }
  }
  auto own = [&](double *&buf, size_t n) -> int {
    if (!buf)
      CUDA_TRY(cudaMalloc(&buf, (n > 0 ? n : 1) * sizeof(double)));
    return AB2_OK;
  };
  int rc;
  if ((rc = own(s->own_stage, stage_total(s))) != AB2_OK || (rc = own(s->own_term, (size_t)B * s->trec)) != AB2_OK ||
      (rc = own(s->own_G0, (size_t)B * nc0 * nx)) != AB2_OK || (rc = own(s->own_g0, (size_t)B * nc0)) != AB2_OK)
    return rc;
  const size_t srec_sym = ab2_gar_stage_record_doubles_sym(nx, s->d.nu, s->d.nc);
  if (sym && (rc = own(s->own_stage_sym, (size_t)B * N * srec_sym)) != AB2_OK)
    return rc;
  s->p.stage = s->own_stage;
  s->p.stage_head = 0;
  s->p.term = s->own_term;
  s->p.G0 = s->own_G0;
  s->p.g0 = s->own_g0;
  s->have_problem = true;
  s->p.mueq = mueq;
  s->p.do_bwd = 1;
  s->p.do_fwd = 1;
  if (nchunks <= 0) { // enough slices to hide the first upload / last download, each still one full wave
    nchunks = B / 512;
    if (nchunks > 16)
      nchunks = 16;
  }
  if (nchunks < 1)
    nchunks = 1;
  if (nchunks > B)
    nchunks = B;
  cudaStream_t user = (cudaStream_t)stream;
  CUDA_TRY(cudaEventRecord(s->pipe_fork, user));
  for (int i = 0; i < NS && i < nchunks; ++i)
    CUDA_TRY(cudaStreamWaitEvent(s->pipe_stream[i], s->pipe_fork, 0));
  for (int c = 0; c < nchunks; ++c) {
    const int b0 = (int)((long long)B * c / nchunks), b1 = (int)((long long)B * (c + 1) / nchunks);
    const int nb = b1 - b0;
    if (nb <= 0)
      continue;
    cudaStream_t st = s->pipe_stream[c % NS];
    auto up = [&](double *dev, const double *host, size_t per_inst) -> int {
      if (per_inst)
        CUDA_TRY(cudaMemcpyAsync(dev + (size_t)b0 * per_inst, host + (size_t)b0 * per_inst,
                                 (size_t)nb * per_inst * sizeof(double), cudaMemcpyHostToDevice, st));
      return AB2_OK;
    };
    // the small uploads first: enqueued behind the expansion kernel they would sit in the copy engine's queue
    // behind the NEXT slice's records and hold this slice's sweep back by a whole upload
    if ((rc = up(s->own_term, term, s->trec)) != AB2_OK ||
        (rc = up(s->own_G0, G0, (size_t)nc0 * nx)) != AB2_OK || (rc = up(s->own_g0, g0, nc0)) != AB2_OK)
      return rc;
    if (sym) { // lower triangles over PCIe, full records rebuilt in HBM
      if ((rc = up(s->own_stage_sym, stage, (size_t)N * srec_sym)) != AB2_OK)
        return rc;
      const long nrec = (long)nb * N;
      if (nrec > 0) {
        const int full = nx * nx + nx * s->d.nu + nx + nx * nx + nx * s->d.nu + s->d.nu * s->d.nu + nx + s->d.nu +
                         s->d.nc * nx + s->d.nc * s->d.nu + s->d.nc;
        long blocks = (nrec * (long)s->srec + 255) / 256;
        if (blocks > 148 * 8)
          blocks = 148 * 8;
        // same shared-memory carve-out as the sweeps it runs beside (an SM is configured for one carve-out at a time)
        static bool carve = false;
        if (!carve) {
          CUDA_TRY(cudaFuncSetAttribute(ab2::expand_sym_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        (int)cudaSharedmemCarveoutMaxShared));
          carve = true;
        }
        if (s->srec * sizeof(int) > 48 * 1024) // (records beyond 12 288 doubles: opt in to the larger table)
          CUDA_TRY(cudaFuncSetAttribute(ab2::expand_sym_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(s->srec * sizeof(int))));
        ab2::expand_sym_kernel<<<(int)blocks, 256, s->srec * sizeof(int), st>>>(
            s->own_stage_sym + (size_t)b0 * N * srec_sym, s->own_stage + (size_t)b0 * N * s->srec, nrec, nx, s->d.nu, s->d.nc,
            full, (int)s->srec, (int)srec_sym);
        CUDA_TRY(cudaGetLastError());
        s->launches += 1;
      }
    } else if ((rc = up(s->own_stage, stage, (size_t)N * s->srec)) != AB2_OK)
      return rc;
    const ab2::SweepParams q = slice_params(s, b0, nb);
    if (int rc2 = run_kernels(s, q, 1, 1, st))
      return rc2;
    for (int i = 0; i < nwhat; ++i) {
      const int w = whats[i];
      const size_t per_inst = (size_t)s->out_knots[w] * s->out_rec[w];
      if (per_inst)
        CUDA_TRY(cudaMemcpyAsync(dsts[i] + (size_t)b0 * per_inst, s->out[w] + (size_t)b0 * per_inst,
                                 (size_t)nb * per_inst * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
  }
  for (int i = 0; i < NS && i < nchunks; ++i) {
    CUDA_TRY(cudaEventRecord(s->pipe_done[i], s->pipe_stream[i]));
    CUDA_TRY(cudaStreamWaitEvent(user, s->pipe_done[i], 0));
  }
  s->have_backward = true;
  s->have_forward = true;
  s->fac_head = 0;
  return AB2_OK;
}

size_t ab2_gar_output_doubles(const ab2_gar_solver *s, int what) {
  if (!s || what < 0 || what >= AB2_OUT_COUNT)
    return 0;
  return s->out_doubles[what];
}

// Dense copy of knots [t0, t0 + nt) of instances [b0, b0 + nb) of a per-knot array whose first `nring` knots
// are ring-indexed with head `head` (knot t in slot (t + head) mod nring; knots >= nring -- the terminal
// entry of VXX / VX -- stay in place): at most three strided copies.
static int copy_ring(const double *base, size_t rec, int knots, int nring, int head, int b0, int nb, int t0, int nt,
                     double *dst, cudaMemcpyKind kind, cudaStream_t st) {
  auto piece = [&](int lt, int pt, int len) -> int { // logical start, physical start, length
    if (len <= 0)
      return AB2_OK;
    CUDA_TRY(cudaMemcpy2DAsync(dst + (size_t)(lt - t0) * rec, (size_t)nt * rec * sizeof(double),
                               base + ((size_t)b0 * knots + pt) * rec, (size_t)knots * rec * sizeof(double),
                               (size_t)len * rec * sizeof(double), (size_t)nb, kind, st));
    return AB2_OK;
  };
  const int t1 = t0 + nt;
  auto clip = [&](int lo, int hi, int shift) { // logical [lo, hi) of the ring, physical = logical + shift
    const int a = t0 > lo ? t0 : lo, b = t1 < hi ? t1 : hi;
    return piece(a, a + shift, b - a);
  };
  int rc;
  if ((rc = clip(0, nring - head, head)) != AB2_OK || (rc = clip(nring - head, nring, head - nring)) != AB2_OK ||
      (rc = clip(nring, knots, 0)) != AB2_OK)
    return rc;
  return AB2_OK;
}
static bool ring_indexed(const ab2_gar_solver *s, int what) {
  return s->fac_head != 0 && (what == AB2_OUT_FF || what == AB2_OUT_FB || what == AB2_OUT_VXX || what == AB2_OUT_VX);
}

int ab2_gar_get(ab2_gar_solver *s, int what, double *dst, int memspace, void *stream) {
  if (!s || !dst || what < 0 || what >= AB2_OUT_COUNT)
    return fail(AB2_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(s->d.device));
  if (s->out_doubles[what] == 0)
    return AB2_OK;
  if (ring_indexed(s, what)) // between a cycle_append and the next backward: knot order through the ring head
    return copy_ring(s->out[what], s->out_rec[what], s->out_knots[what], s->d.horizon, s->fac_head, 0, s->d.batch, 0,
                     s->out_knots[what], dst, memspace == AB2_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                     (cudaStream_t)stream);
  CUDA_TRY(cudaMemcpyAsync(dst, s->out[what], s->out_doubles[what] * sizeof(double),
                           memspace == AB2_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                           (cudaStream_t)stream));
  return AB2_OK;
}

int ab2_gar_get_range(ab2_gar_solver *s, int what, int b0, int nb, int t0, int nt, double *dst,
                      int memspace, void *stream) {
  if (!s || !dst || what < 0 || what >= AB2_OUT_COUNT)
    return fail(AB2_ERR_INVALID, "bad argument");
  const int knots = s->out_knots[what];
  if (b0 < 0 || nb < 0 || b0 + nb > s->d.batch || t0 < 0 || nt < 0 || t0 + nt > knots)
    return fail(AB2_ERR_INVALID, "range out of bounds");
  CUDA_TRY(cudaSetDevice(s->d.device));
  const size_t rec = s->out_rec[what];
  if (rec == 0 || nb == 0 || nt == 0)
    return AB2_OK;
  if (ring_indexed(s, what))
    return copy_ring(s->out[what], rec, knots, s->d.horizon, s->fac_head, b0, nb, t0, nt, dst,
                     memspace == AB2_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  const double *src = s->out[what] + ((size_t)b0 * knots + t0) * rec;
  CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)nt * rec * sizeof(double), src, (size_t)knots * rec * sizeof(double),
                             (size_t)nt * rec * sizeof(double), (size_t)nb,
                             memspace == AB2_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
  return AB2_OK;
}

int ab2_gar_first_step_policy(ab2_gar_solver *s, double *dst, void *stream) {
  if (!s || !dst)
    return fail(AB2_ERR_INVALID, "bad argument");
  if (s->d.horizon < 1)
    return fail(AB2_ERR_INVALID, "first_step_policy needs horizon >= 1");
  if (s->dense)
    return fail(AB2_ERR_UNSUPPORTED, "first_step_policy: not offered by the stage-dense solver");
  if (!s->have_backward)
    return fail(AB2_ERR_STATE, "first_step_policy before backward()");
  CUDA_TRY(cudaSetDevice(s->d.device));
  const long total = (long)s->d.batch * s->d.nu * (s->d.nx + 1);
  const int threads = 256;
  long blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 8)
    blocks = 148 * 8;
  ab2::first_step_policy_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(
      s->out[AB2_OUT_FB], s->out[AB2_OUT_FF], dst, s->d.batch, s->d.horizon, s->nr, s->d.nu, s->d.nx, s->fac_head);
  CUDA_TRY(cudaGetLastError());
  s->launches += 1;
  return AB2_OK;
}

int ab2_gar_get_gains(ab2_gar_solver *s, double *dst, int memspace, void *stream) {
  if (!s || !dst)
    return fail(AB2_ERR_INVALID, "bad argument");
  if (!s->have_backward)
    return fail(AB2_ERR_STATE, "get_gains before backward()");
  if (s->dense)
    return fail(AB2_ERR_UNSUPPORTED, "get_gains: the results_.gains_ layout is the proximal solver's");
  CUDA_TRY(cudaSetDevice(s->d.device));
  const long nrec = (long)s->d.batch * s->d.horizon;
  const size_t total = (size_t)nrec * s->nr * (s->d.nx + 1);
  if (total == 0)
    return AB2_OK;
  double *out = dst;
  if (memspace != AB2_DEVICE) { // stage on the device, then one copy
    if (!s->gains_tmp)
      CUDA_TRY(cudaMalloc(&s->gains_tmp, total * sizeof(double)));
    out = s->gains_tmp;
  }
  long blocks = ((long)total + 255) / 256;
  if (blocks > 148 * 8)
    blocks = 148 * 8;
  ab2::gains_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(s->out[AB2_OUT_FB], s->out[AB2_OUT_FF], out, nrec,
                                                                    s->nr, s->d.nx, s->d.horizon, s->fac_head);
  CUDA_TRY(cudaGetLastError());
  s->launches += 1;
  if (memspace != AB2_DEVICE)
    CUDA_TRY(cudaMemcpyAsync(dst, out, total * sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return AB2_OK;
}

int ab2_gar_kkt_error(ab2_gar_solver *s, double mueq, double *dst, int memspace, void *stream) {
  if (!s || !dst)
    return fail(AB2_ERR_INVALID, "bad argument");
  if (!s->have_problem || !s->have_backward || !s->have_forward)
    return fail(AB2_ERR_STATE, "kkt_error needs a problem and a completed sweep (no forward pass since the last backward)");
  if (s->rec_nth > 0)
    return fail(AB2_ERR_UNSUPPORTED, "kkt_error: parametric problems (nth > 0) are not supported");
  CUDA_TRY(cudaSetDevice(s->d.device));
  if (!s->kkt_tmp)
    CUDA_TRY(cudaMalloc(&s->kkt_tmp, (size_t)s->d.batch * 3 * sizeof(double)));
  ab2::KktErrorArgs a;
  a.batch = s->d.batch;
  a.N = s->d.horizon;
  a.nx = s->d.nx;
  a.nu = s->d.nu;
  a.nc = s->d.nc;
  a.nct = s->d.nct;
  a.nc0 = s->d.nc0;
  a.srec = s->srec;
  a.trec = s->trec;
  a.stage_head = s->p.stage_head;
  a.mueq = mueq;
  a.stage = s->p.stage;
  a.term = s->p.term;
  a.G0 = s->p.G0;
  a.g0 = s->p.g0;
  a.xs = s->out[AB2_OUT_XS];
  a.us = s->out[AB2_OUT_US];
  a.vs = s->out[AB2_OUT_VS];
  a.vsT = s->out[AB2_OUT_VST];
  a.lbd0 = s->out[AB2_OUT_LBD0];
  a.lbdas = s->out[AB2_OUT_LBDAS];
  a.out = (memspace == AB2_DEVICE) ? dst : s->kkt_tmp;
  CUDA_TRY(ab2::launch_kkt_error(a, (cudaStream_t)stream));
  s->launches += 1;
  if (memspace != AB2_DEVICE)
    CUDA_TRY(cudaMemcpyAsync(dst, s->kkt_tmp, (size_t)s->d.batch * 3 * sizeof(double), cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
  return AB2_OK;
}

int ab2_gar_device_ptr(ab2_gar_solver *s, int what, double **out) {
  if (!s || !out || what < 0 || what >= AB2_OUT_COUNT)
    return fail(AB2_ERR_INVALID, "bad argument");
  *out = s->out[what];
  return AB2_OK;
}

int ab2_gar_status(ab2_gar_solver *s, int *dst, int memspace, void *stream) {
  if (!s || !dst)
    return fail(AB2_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(s->d.device));
  CUDA_TRY(cudaMemcpyAsync(dst, s->status, sizeof(int) * s->d.batch,
                           memspace == AB2_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                           (cudaStream_t)stream));
  return AB2_OK;
}

// ---- line-search consumers (linesearch.cu) ----
static ab2::LineSearchArgs ls_args(const ab2_gar_solver *s) {
  ab2::LineSearchArgs a;
  a.batch = s->d.batch;
  a.N = s->d.horizon;
  a.nx = s->d.nx;
  a.nu = s->d.nu;
  a.nc = s->d.nc;
  a.nct = s->d.nct;
  a.nc0 = s->d.nc0;
  a.dxs = s->out[AB2_OUT_XS];
  a.dus = s->out[AB2_OUT_US];
  a.dvs = s->out[AB2_OUT_VS];
  a.dvsT = s->out[AB2_OUT_VST];
  a.dlam0 = s->out[AB2_OUT_LBD0];
  a.dlams = s->out[AB2_OUT_LBDAS];
  return a;
}
static int ls_result(ab2_gar_solver *s, double *dst, int memspace, cudaStream_t st, double **dev) {
  if (memspace == AB2_DEVICE) {
    *dev = dst;
    return AB2_OK;
  }
  if (!s->ls_tmp)
    CUDA_TRY(cudaMalloc(&s->ls_tmp, (size_t)s->d.batch * sizeof(double)));
  *dev = s->ls_tmp;
  (void)st;
  return AB2_OK;
}
int ab2_gar_linear_step(ab2_gar_solver *s, double alpha, const ab2_ls_iterate *cur, const ab2_ls_trial *trial,
                        void *stream) {
  if (!s || !cur || !trial)
    return fail(AB2_ERR_INVALID, "null argument");
  if (!s->have_forward)
    return fail(AB2_ERR_STATE, "linear_step needs the step of a forward pass");
  const ab2_gar_dims &d = s->d;
  const bool ok = cur->xs && trial->xs && (d.horizon == 0 || (cur->us && trial->us && cur->lams && trial->lams)) &&
                  (d.nc == 0 || d.horizon == 0 || (cur->vs && trial->vs)) && (d.nct == 0 || (cur->vsT && trial->vsT)) &&
                  (d.nc0 == 0 || (cur->lam0 && trial->lam0));
  if (!ok)
    return fail(AB2_ERR_INVALID, "linear_step: a required array is NULL for these dimensions");
  CUDA_TRY(cudaSetDevice(d.device));
  ab2::LinearStepIO io{cur->xs, cur->us, cur->vs, cur->vsT, cur->lam0, cur->lams,
                       trial->xs, trial->us, trial->vs, trial->vsT, trial->lam0, trial->lams};
  CUDA_TRY(ab2::launch_linear_step(ls_args(s), io, alpha, (cudaStream_t)stream));
  s->launches += 1;
  return AB2_OK;
}
int ab2_gar_directional_derivative(ab2_gar_solver *s, const double *Lxs, const double *Lus, double *dst, int memspace,
                                   void *stream) {
  if (!s || !Lxs || !dst || (s->d.horizon > 0 && !Lus))
    return fail(AB2_ERR_INVALID, "null argument");
  if (!s->have_forward)
    return fail(AB2_ERR_STATE, "directional_derivative needs the step of a forward pass");
  CUDA_TRY(cudaSetDevice(s->d.device));
  double *dev = nullptr;
  if (int rc = ls_result(s, dst, memspace, (cudaStream_t)stream, &dev))
    return rc;
  CUDA_TRY(ab2::launch_directional_derivative(ls_args(s), Lxs, Lus, dev, (cudaStream_t)stream));
  s->launches += 1;
  if (memspace != AB2_DEVICE)
    CUDA_TRY(cudaMemcpyAsync(dst, dev, (size_t)s->d.batch * sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return AB2_OK;
}
int ab2_gar_al_value(ab2_gar_solver *s, const ab2_ls_iterate *plus, const double *cost, double mudyn, double mucstr,
                     double *dst, int memspace, void *stream) {
  if (!s || !plus || !dst)
    return fail(AB2_ERR_INVALID, "null argument");
  const ab2_gar_dims &d = s->d;
  if ((d.nc0 > 0 && !plus->lam0) || (d.horizon > 0 && !plus->lams) || (d.nc > 0 && d.horizon > 0 && !plus->vs) ||
      (d.nct > 0 && !plus->vsT))
    return fail(AB2_ERR_INVALID, "al_value: a required multiplier array is NULL for these dimensions");
  CUDA_TRY(cudaSetDevice(d.device));
  double *dev = nullptr;
  if (int rc = ls_result(s, dst, memspace, (cudaStream_t)stream, &dev))
    return rc;
  CUDA_TRY(ab2::launch_al_value(d.batch, d.horizon, d.nx, d.nc, d.nct, d.nc0, plus->lam0, plus->lams, plus->vs, plus->vsT,
                                cost, mudyn, mucstr, dev, (cudaStream_t)stream));
  s->launches += 1;
  if (memspace != AB2_DEVICE)
    CUDA_TRY(cudaMemcpyAsync(dst, dev, (size_t)d.batch * sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return AB2_OK;
}

int ab2_fddp_backward_pass(ab2_gar_solver *s, const ab2_fddp_inputs *in, double *Vx_out, double *Quuks_out, void *stream) {
  if (!s || !in)
    return fail(AB2_ERR_INVALID, "null argument");
  const ab2_gar_dims &d = s->d;
  if (d.nc != 0 || d.nct != 0 || d.nc0 != d.nx || s->rec_nth != 0 || s->legs > 1 || d.horizon < 1)
    return fail(AB2_ERR_INVALID, "fddp_backward_pass needs a serial solver with nc = nct = 0, nc0 = nx, horizon >= 1");
  if (!in->Jx || !in->Ju || !in->fs || !in->Lxx || !in->Lxu || !in->Luu || !in->Lx || !in->Lu || !in->Lxx_N || !in->Lx_N)
    return fail(AB2_ERR_INVALID, "ab2_fddp_inputs: a required array is NULL");
  CUDA_TRY(cudaSetDevice(d.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d.batch, N = d.horizon, nx = d.nx, nu = d.nu;
  auto own = [&](double *&buf, size_t n) -> int {
    if (!buf)
      CUDA_TRY(cudaMalloc(&buf, (n > 0 ? n : 1) * sizeof(double)));
    return AB2_OK;
  };
  int rc;
  if ((rc = own(s->fddp_slack, (size_t)B * N * nx)) != AB2_OK || (rc = own(s->fddp_G0, (size_t)B * nx * nx)) != AB2_OK ||
      (rc = own(s->fddp_g0, (size_t)B * nx)) != AB2_OK || (rc = own(s->fddp_vx, (size_t)B * (N + 1) * nx)) != AB2_OK)
    return rc;
  ab2::fddp_prep_kernel<<<148 * 4, 256, 0, st>>>(in->fs, s->fddp_slack, s->fddp_G0, s->fddp_g0, B, N, nx);
  CUDA_TRY(cudaGetLastError());
  s->launches += 1;
  ab2_lq_inputs lq;
  std::memset(&lq, 0, sizeof(lq));
  lq.Jx = in->Jx;
  lq.Ju = in->Ju;
  lq.slack = s->fddp_slack;
  lq.Lxx = in->Lxx;
  lq.Lxu = in->Lxu;
  lq.Luu = in->Luu;
  lq.Lx = in->Lx;
  lq.Lu = in->Lu;
  lq.Lxx_N = in->Lxx_N;
  lq.Lx_N = in->Lx_N;
  lq.G0 = s->fddp_G0;
  lq.g0 = s->fddp_g0;
  lq.preg = in->preg; // Q, R and the terminal Q carry + preg I (:217, :246, :273)
  lq.mu_inv = 1.0;
  if ((rc = ab2_gar_assemble(s, &lq, stream)) != AB2_OK)
    return rc;
  if ((rc = ab2_gar_backward(s, 1.0, stream)) != AB2_OK) // (mueq is unused without constraints)
    return rc;
  double *vxo = Vx_out ? Vx_out : s->fddp_vx;
  ab2::fddp_vx_kernel<<<148 * 4, 256, 0, st>>>(s->out[AB2_OUT_VXX], s->out[AB2_OUT_VX], in->fs, vxo, B, N, nx);
  CUDA_TRY(cudaGetLastError());
  s->launches += 1;
  if (Quuks_out) {
    ab2::fddp_quuks_kernel<<<148 * 4, 256, 0, st>>>(in->Ju, in->Lu, vxo, Quuks_out, B, N, nx, nu);
    CUDA_TRY(cudaGetLastError());
    s->launches += 1;
  }
  return AB2_OK;
}

int ab2_gar_collapse_feedback(ab2_gar_solver *s, void *stream) {
  if (!s)
    return fail(AB2_ERR_INVALID, "null solver");
  if (s->legs <= 1)
    return AB2_OK; // the serial solver's collapseFeedback is a no-op (riccati-base.hpp:32)
  if (!s->have_backward)
    return fail(AB2_ERR_STATE, "collapse_feedback before backward()");
  if (s->d.horizon < 1)
    return AB2_OK;
  CUDA_TRY(cudaSetDevice(s->d.device));
  CUDA_TRY(ab2::launch_collapse(s->p, s->d.nx, s->d.nu, s->d.nc, (cudaStream_t)stream));
  s->launches += 1;
  return AB2_OK;
}

int ab2_gar_pivot_stats(ab2_gar_solver *s, int *dst, int memspace, void *stream) {
  if (!s || !dst)
    return fail(AB2_ERR_INVALID, "bad argument");
  if (!s->have_backward)
    return fail(AB2_ERR_STATE, "pivot_stats before backward()");
  CUDA_TRY(cudaSetDevice(s->d.device));
  CUDA_TRY(cudaMemcpyAsync(dst, s->pivstat, sizeof(int) * s->d.batch,
                           memspace == AB2_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                           (cudaStream_t)stream));
  return AB2_OK;
}

int ab2_gar_cycle_append(ab2_gar_solver *s, const double *new_last, int memspace, void *stream) {
  if (!s || !new_last)
    return fail(AB2_ERR_INVALID, "bad argument");
  const int N = s->d.horizon, B = s->d.batch;
  if (N < 1)
    return fail(AB2_ERR_INVALID, "cycle_append needs horizon >= 1");
  if (s->rec_nth > 0)
    return fail(AB2_ERR_UNSUPPORTED, "cycle_append: parametric problems (nth > 0) are not supported");
  CUDA_TRY(cudaSetDevice(s->d.device));
  cudaStream_t st = (cudaStream_t)stream;
  // factors: datas[0..N-1] rotate left, datas[N-1] re-created (zeros), terminal kept
  // (proximal-riccati.hxx:79-83).  Vxx/vx have N+1 entries; the last is the terminal's.
  if (s->legs > 1) { // the parallel solver drops every factor and starts over (parallel-solver.hxx:246-258)
    for (int w : {AB2_OUT_FF, AB2_OUT_FB, AB2_OUT_VXX, AB2_OUT_VX, AB2_OUT_FTH, AB2_OUT_VXT, AB2_OUT_VTT, AB2_OUT_VT})
      if (s->out_doubles[w])
        CUDA_TRY(cudaMemsetAsync(s->out[w], 0, s->out_doubles[w] * sizeof(double), st));
  }
  // O(1) in the horizon: nothing moves.  The per-knot factor arrays and the solver-owned stage records are rings;
  // rotating left = advancing the head by one.  What is touched is ONE knot slot per instance and array:
  // the factor slot of the new last knot is zeroed (datas[N-1] re-created, :82-83), its record is written.
  if (s->legs <= 1) {
    s->fac_head = (s->fac_head + 1) % N;
    const int slot = (N - 1 + s->fac_head) % N; // physical slot of the new stage knot N-1 (= the old knot 0's)
    for (int w : {AB2_OUT_FF, AB2_OUT_FB, AB2_OUT_VXX, AB2_OUT_VX}) {
      const size_t rec = s->out_rec[w];
      if (rec == 0)
        continue;
      CUDA_TRY(cudaMemset2DAsync(s->out[w] + (size_t)slot * rec, (size_t)s->out_knots[w] * rec * sizeof(double), 0,
                                 rec * sizeof(double), (size_t)B, st));
    }
  }
  // kkt0 zeroed (:84-86)
  if (s->out_doubles[AB2_OUT_KKT0])
    CUDA_TRY(cudaMemsetAsync(s->out[AB2_OUT_KKT0], 0, s->out_doubles[AB2_OUT_KKT0] * sizeof(double), st));
  // the problem itself: our own device copy (host-fed problems) rotates the same way; a device-resident
  // caller rotates its own buffers, like cycleProblem does for the reference's problem (solver-proxddp.hxx:202-209).
  if (s->own_stage && s->p.stage == s->own_stage) {
    s->p.stage_head = (s->p.stage_head + 1) % N;
    const int slot = (N - 1 + s->p.stage_head) % N;
    CUDA_TRY(cudaMemcpy2DAsync(s->own_stage + (size_t)slot * s->srec, (size_t)N * s->srec * sizeof(double),
                               new_last, (size_t)s->srec * sizeof(double), (size_t)s->srec * sizeof(double),
                               (size_t)B,
                               memspace == AB2_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  }
  CUDA_TRY(cudaGetLastError());
  s->have_backward = false;
  s->have_forward = false;
  return AB2_OK;
}

int ab2_gar_phase_clocks(ab2_gar_solver *s, long long *dst16) { // AB2_PHASE_CLOCKS=1: cycles per phase, instance 0
  if (!s || !dst16 || !s->p.clk)
    return fail(AB2_ERR_STATE, "phase clocks are off (set AB2_PHASE_CLOCKS=1 before create)");
  CUDA_TRY(cudaMemcpy(dst16, s->p.clk, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemset(s->p.clk, 0, 16 * sizeof(long long)));
  return AB2_OK;
}

int ab2_gar_ring_heads(const ab2_gar_solver *s, int *factor_head, int *stage_head) {
  if (!s)
    return fail(AB2_ERR_INVALID, "null solver");
  if (factor_head)
    *factor_head = s->fac_head;
  if (stage_head)
    *stage_head = s->p.stage_head;
  return AB2_OK;
}

int ab2_gar_synchronize(ab2_gar_solver *s, void *stream) {
  if (!s)
    return fail(AB2_ERR_INVALID, "null solver");
  CUDA_TRY(cudaSetDevice(s->d.device));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return AB2_OK;
}

// ---- multi-GPU: fused pack + all-gather of the first-step policy over NVLink peer memory ----
int ab2_gar_peer_gather_init(ab2_gar_solver *s, int world, int rank, void *ipc_handle_out) {
  if (!s || !ipc_handle_out || world < 1 || world > ab2::kMaxPeers || rank < 0 || rank >= world)
    return fail(AB2_ERR_INVALID, "peer_gather_init: bad argument (world <= 8)");
  if (s->d.horizon < 1)
    return fail(AB2_ERR_INVALID, "peer_gather needs horizon >= 1");
  if (s->pg_local)
    return fail(AB2_ERR_STATE, "peer_gather_init called twice");
  CUDA_TRY(cudaSetDevice(s->d.device));
  const size_t per = (size_t)s->d.nu * (s->d.nx + 1);
  s->pg_buf_doubles = 3 * (size_t)world * s->d.batch * per; // three slots: step s lives in slot s mod 3
  const size_t bytes = s->pg_buf_doubles * sizeof(double) + 2 * ab2::kMaxPeers * sizeof(unsigned long long);
  CUDA_TRY(cudaMalloc(&s->pg_local, bytes));
  CUDA_TRY(cudaMemset(s->pg_local, 0, bytes));
  CUDA_TRY(cudaMalloc(&s->pg_done, sizeof(unsigned int)));
  CUDA_TRY(cudaMemset(s->pg_done, 0, sizeof(unsigned int)));
  s->pg_world = world;
  s->pg_rank = rank;
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, s->pg_local));
  static_assert(sizeof(h) == 64, "IPC handle size");
  std::memcpy(ipc_handle_out, &h, sizeof(h));
  return AB2_OK;
}

int ab2_gar_peer_gather_connect(ab2_gar_solver *s, const void *all_handles) {
  if (!s || !all_handles || !s->pg_local)
    return fail(AB2_ERR_STATE, "peer_gather_connect before peer_gather_init");
  CUDA_TRY(cudaSetDevice(s->d.device));
  for (int w = 0; w < s->pg_world; ++w) {
    void *base = s->pg_local;
    if (w != s->pg_rank) {
      cudaIpcMemHandle_t h;
      std::memcpy(&h, (const char *)all_handles + 64 * (size_t)w, sizeof(h));
      const cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) { // no peer access / IPC not permitted: leave the handle unconnected (callers fall back)
        cudaGetLastError();
        for (int v = 0; v < w; ++v) {
          if (v != s->pg_rank && s->pg_peer_base[v])
            cudaIpcCloseMemHandle(s->pg_peer_base[v]);
          s->pg_peer_base[v] = nullptr;
        }
        return fail(AB2_ERR_CUDA, std::string("peer_gather_connect: cudaIpcOpenMemHandle(rank ") + std::to_string(w) +
                                      "): " + cudaGetErrorString(e));
      }
    }
    s->pg_peer_base[w] = base;
    s->pg_ptrs.buf[w] = (double *)base;
    unsigned long long *fl = (unsigned long long *)((double *)base + s->pg_buf_doubles);
    s->pg_ptrs.data_flag[w] = fl;
    s->pg_ptrs.ack_flag[w] = fl + ab2::kMaxPeers;
  }
  return AB2_OK;
}

int ab2_gar_policy_allgather(ab2_gar_solver *s, void *stream) {
  if (!s || !s->pg_peer_base[0])
    return fail(AB2_ERR_STATE, "policy_allgather before peer_gather_connect");
  if (!s->have_backward)
    return fail(AB2_ERR_STATE, "policy_allgather before backward()");
  CUDA_TRY(cudaSetDevice(s->d.device));
  const long total = (long)s->d.batch * s->d.nu * (s->d.nx + 1);
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 2)
    blocks = 148 * 2; // all resident at once: the last-CTA publication never waits on an unscheduled CTA
  s->pg_step += 1;
  if (s->pg_pushed_step == s->pg_step) // the sweep stored the blocks itself: publish the flags
    ab2::policy_publish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(s->pg_ptrs, s->pg_world, s->pg_rank, s->pg_step);
  else
    ab2::policy_allgather_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
        s->out[AB2_OUT_FB], s->out[AB2_OUT_FF], s->pg_ptrs, s->pg_world, s->pg_rank, s->d.batch, s->d.horizon, s->nr,
        s->d.nu, s->d.nx, s->pg_step, s->pg_done);
  CUDA_TRY(cudaGetLastError());
  s->launches += 1;
  return AB2_OK;
}

int ab2_gar_policy_allgather_wait(ab2_gar_solver *s, void *stream) {
  if (!s || !s->pg_peer_base[0] || s->pg_step == 0)
    return fail(AB2_ERR_STATE, "policy_allgather_wait before policy_allgather");
  CUDA_TRY(cudaSetDevice(s->d.device));
  static bool carve = false; // (an SM holds one shared-memory carve-out at a time: ask for the sweeps' so that this
  if (!carve) {              //  kernel can start while a sweep is resident)
    CUDA_TRY(cudaFuncSetAttribute(ab2::policy_wait_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxShared));
    carve = true;
  }
  ab2::policy_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(s->pg_ptrs, s->pg_world, s->pg_rank, s->pg_step);
  CUDA_TRY(cudaGetLastError());
  s->launches += 1;
  return AB2_OK;
}

int ab2_gar_peer_gather_buffer(ab2_gar_solver *s, double **out, long *step) {
  if (!s || !out || !s->pg_local)
    return fail(AB2_ERR_STATE, "peer_gather_buffer before peer_gather_init");
  const size_t half = (size_t)(s->pg_step % 3) * s->pg_world * s->d.batch * s->d.nu * (s->d.nx + 1);
  *out = (double *)s->pg_local + half;
  if (step)
    *step = (long)s->pg_step;
  return AB2_OK;
}

int ab2_gar_pinned_alloc(size_t bytes, void **out) {
  if (!out)
    return fail(AB2_ERR_INVALID, "null argument");
  *out = nullptr;
  CUDA_TRY(cudaHostAlloc(out, bytes > 0 ? bytes : 1, cudaHostAllocPortable));
  return AB2_OK;
}
void ab2_gar_pinned_free(void *p) {
  if (p)
    cudaFreeHost(p);
}

long ab2_gar_launch_count(const ab2_gar_solver *s) { return s ? s->launches : 0; }

int ab2_gar_kernel_info(const ab2_gar_solver *s, int *group_lanes, int *smem_bytes_per_cta,
                        int *threads_per_cta, int *grid, int *regs_per_thread) {
  if (!s)
    return fail(AB2_ERR_INVALID, "null solver");
  int info[6] = {0, 0, 0, 0, 0, 0};
  CUDA_TRY(cudaSetDevice(s->d.device));
  if (s->k && s->variant != 9)
    CUDA_TRY(s->k->launch(s->p, s->variant, s->group_doubles, 0, info));
  else
    CUDA_TRY(ab2::launch_block(s->p, s->d.nx, s->d.nu, s->d.nc, 0, info));
  if (group_lanes)
    *group_lanes = info[0];
  if (smem_bytes_per_cta)
    *smem_bytes_per_cta = info[1];
  if (threads_per_cta)
    *threads_per_cta = info[2];
  if (grid)
    *grid = info[3];
  if (regs_per_thread)
    *regs_per_thread = info[4] | (info[5] << 16); /* high half: resident CTAs per SM */
  return AB2_OK;
}

} // extern "C"
