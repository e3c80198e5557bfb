// Compile-time shapes instantiated by the CUDA library (and by the host emulation
// used in tests).  X(NX, NU, NC, G): G lanes per instance (sub-warp or warp);
// requires nx+nu+1 <= G, nu+nc <= G and nx+nc0 <= G at run time.
#pragma once
#ifdef AB2_ONLY_C2
#define AB2_FOR_EACH_CONFIG(X) X(12, 6, 0, 32)
#else
#define AB2_FOR_EACH_CONFIG(X)                                                  \
  X(2, 2, 0, 8)   /* tests/gar/riccati.cpp short-horizon shape            */    \
  X(2, 2, 2, 8)   /* ... with the control-constrained knot                */    \
  X(3, 2, 0, 8)   /* SE2 car, LQ dims (ndx=3, nu=2): bench/se2-car.cpp    */    \
  X(4, 2, 2, 8)   /* BASELINE config 3: nx4 nu2 nc2 (examples/clqr.cpp)   */    \
  X(4, 2, 0, 8)                                                                 \
  X(5, 2, 2, 16)                                                                \
  X(6, 3, 0, 16)  /* BASELINE config 1                                    */    \
  X(8, 3, 0, 16)                                                                \
  X(10, 4, 0, 32) /* tests/gar/riccati.cpp parametric shape (nth=0 part)  */    \
  X(12, 6, 0, 32) /* BASELINE config 2 (headline)                         */    \
  X(12, 6, 6, 32)                                                               \
  X(14, 7, 0, 32) /* BASELINE config 4: Talos arm                         */

#endif
