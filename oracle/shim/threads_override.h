// TEST INFRASTRUCTURE ONLY -- build scaffolding for oracle/_ref/avx2_nproc.
//
// Force-included AFTER the reference's own paramSetting.hpp (which is `#pragma once`, so the
// sources' later #include of it is a no-op): re-defines the compile-time OpenMP team size
// THREADS_NUM_USED (paramSetting.hpp:11, shipped as 8) as a run-time value -- MSN_REF_THREADS if
// set, else the number of processors OpenMP sees.  Every use in matchers.cpp / featextract.cpp is
// a num_threads(...) clause, a loop bound or omp_set_num_threads(...), all of which accept it.
#pragma once
#include <omp.h>
#include <stdlib.h>
static inline int msn_ref_threads(void) {
  static int n = 0;
  if (n == 0) {
    const char* e = getenv("MSN_REF_THREADS");
    n = e ? atoi(e) : omp_get_num_procs();
    if (n < 1) n = 1;
  }
  return n;
}
#undef THREADS_NUM_USED
#define THREADS_NUM_USED msn_ref_threads()
