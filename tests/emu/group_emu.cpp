// Host emulation of the CUDA group program (aligator_b200/csrc/riccati_group.cuh):
// a group is G std::threads, sync() is a std::barrier, the TMA bulk copy is a
// memcpy by lane 0.  TEST INFRASTRUCTURE: lets the CPU suite execute the exact
// index arithmetic of the kernel and compare it with the oracle.
#include <barrier>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../../aligator_b200/csrc/riccati_configs.h"
#include "../../aligator_b200/csrc/riccati_group.cuh"

namespace {
struct HostCtx {
  int lane;
  std::barrier<> *bar;
  void sync() { bar->arrive_and_wait(); }
  void issue_copy(int, double *dst, const double *src, int nd) {
    if (lane == 0)
      std::memcpy(dst, src, sizeof(double) * (size_t)nd);
  }
  void wait_copy(int) { sync(); }
};

template <int NX, int NU, int NC, int G, bool DB> int run(const ab2::SweepParams &p) {
  using C = ab2::Cfg<NX, NU, NC, G, DB>;
  if (NX + p.nc0 > G)
    return 2;
  for (int inst = 0; inst < p.batch; ++inst) {
    std::vector<double> sm((size_t)C::group_doubles(p.nc0),
                           std::numeric_limits<double>::quiet_NaN());
    std::barrier<> bar(G);
    std::vector<std::thread> th;
    for (int l = 0; l < G; ++l)
      th.emplace_back([&, l] {
        HostCtx ctx{l, &bar};
        ab2::riccati_group_sweep<C>(ctx, p, inst, sm.data());
      });
    for (auto &t : th)
      t.join();
  }
  return 0;
}
} // namespace

extern "C" int emu_stage_record(int nx, int nu, int nc) {
#define X(NX, NU, NC, G)                                                        \
  if (nx == NX && nu == NU && nc == NC)                                         \
    return ab2::Cfg<NX, NU, NC, G>::SREC_PAD;
  AB2_FOR_EACH_CONFIG(X)
#undef X
  return -1;
}

extern "C" int emu_sweep(int nx, int nu, int nc, int db, const ab2::SweepParams *p) {
#define X(NX, NU, NC, G)                                                        \
  if (nx == NX && nu == NU && nc == NC)                                         \
    return db ? run<NX, NU, NC, G, true>(*p) : run<NX, NU, NC, G, false>(*p);
  AB2_FOR_EACH_CONFIG(X)
#undef X
  return 1;
}
