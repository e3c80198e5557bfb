// Host emulation of the CTA-per-instance sweep (aligator_b200/csrc/riccati_block.cuh):
// a CTA is 32*NW std::threads, sync() is a CTA-wide std::barrier, the warp-wide mma is
// a fragment exchange behind a per-warp barrier, the TMA bulk copy is a memcpy by
// thread 0.  TEST INFRASTRUCTURE (see group_emu.cpp).
#include <barrier>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

#include "../../aligator_b200/csrc/riccati_block.cuh"

namespace {
struct HostBlockCtx {
  int tid, nthreads, warp, lane, nwarps;
  std::barrier<> *cta;
  std::barrier<> **sub; // sub[w-1]: barrier over the first w warps
  std::barrier<> *wbar;
  double *xa, *xb; // this warp's exchange slots
  void sync() { cta->arrive_and_wait(); }
  void sync_sub(int nth) { sub[nth / 32 - 1]->arrive_and_wait(); }
  void mma(double (&d)[2], double a, double b) {
    xa[lane] = a;
    xb[lane] = b;
    wbar->arrive_and_wait();
    const int g = lane >> 2, q = lane & 3;
    for (int e = 0; e < 2; ++e) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k)
        s += xa[4 * g + k] * xb[4 * (2 * q + e) + k];
      d[e] += s;
    }
    wbar->arrive_and_wait();
  }
  void wsync() { wbar->arrive_and_wait(); }
  double shfl(double v, int src) {
    xa[lane] = v;
    wbar->arrive_and_wait();
    const double r = xa[src & 31];
    wbar->arrive_and_wait();
    return r;
  }
  bool all(bool p) {
    xb[lane] = p ? 1.0 : 0.0;
    wbar->arrive_and_wait();
    bool r = true;
    for (int i = 0; i < 32; ++i)
      r = r && (xb[i] != 0.0);
    wbar->arrive_and_wait();
    return r;
  }
  void issue_copy(int, double *dst, const double *src, int nd) {
    if (tid == 0)
      std::memcpy(dst, src, sizeof(double) * (size_t)nd);
  }
  void wait_copy(int) { sync(); }
  void proxy_fence() {}
};
} // namespace

extern "C" int emu_block_stage_record(int nx, int nu, int nc) {
  return ab2::make_block_dims(nx, nu, nc, 0).srec_pad;
}

template <class D> static int run_block(const D &d, int nwarps, const ab2::SweepParams &p);

// the compile-time specialisation (StaticBlockDims) of two small shapes, to execute that
// instantiation of the code on the CPU
extern "C" int emu_block_sweep_static(int nx, int nu, int nc, int nwarps, const ab2::SweepParams *pp) {
  if (nx == 7 && nu == 3 && nc == 0 && pp->nc0 == 7)
    return run_block(ab2::StaticBlockDims<7, 3, 0, 7>{}, nwarps, *pp);
  if (nx == 9 && nu == 5 && nc == 3 && pp->nc0 == 9)
    return run_block(ab2::StaticBlockDims<9, 5, 3, 9>{}, nwarps, *pp);
  return 1;
}

// parametric problems (nth > 0)
extern "C" int emu_block_stage_record_th(int nx, int nu, int nc, int nth) {
  return ab2::make_block_dims(nx, nu, nc, 0, nth).srec_pad;
}
extern "C" int emu_block_sweep_th(int nx, int nu, int nc, int nth, int nwarps, const ab2::SweepParams *pp) {
  return run_block(ab2::make_block_dims(nx, nu, nc, pp->nc0, nth), nwarps, *pp);
}

extern "C" int emu_block_sweep(int nx, int nu, int nc, int nwarps, const ab2::SweepParams *pp) {
  return run_block(ab2::make_block_dims(nx, nu, nc, pp->nc0), nwarps, *pp);
}

template <class D> static int run_block(const D &d, int nwarps, const ab2::SweepParams &p) {
  const int nx = d.nx;
  const int T = 32 * nwarps;
  if (nx + 1 > T || d.nk > T || nx + p.nc0 > T || d.nr > T || d.nth > T)
    return 2;
  for (int inst = 0; inst < p.batch; ++inst) {
    std::vector<double> sm((size_t)d.s_end, std::numeric_limits<double>::quiet_NaN());
    std::barrier<> cta(T);
    std::vector<std::unique_ptr<std::barrier<>>> wb;
    for (int w = 0; w < nwarps; ++w)
      wb.emplace_back(new std::barrier<>(32));
    std::vector<std::unique_ptr<std::barrier<>>> subs;
    std::vector<std::barrier<> *> subp;
    for (int w = 1; w <= nwarps; ++w) {
      subs.emplace_back(new std::barrier<>(32 * w));
      subp.push_back(subs.back().get());
    }
    std::vector<double> xa(T), xb(T);
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, t] {
        HostBlockCtx ctx{t, T, t / 32, t % 32, nwarps, &cta, subp.data(), wb[t / 32].get(),
                         xa.data() + 32 * (t / 32), xb.data() + 32 * (t / 32)};
        ab2::riccati_block_sweep(ctx, p, d, inst, sm.data());
      });
    for (auto &t : th)
      t.join();
  }
  return 0;
}
