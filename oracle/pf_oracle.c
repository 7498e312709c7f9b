/*
 * pf_oracle.c -- CPU oracle (test infrastructure, see pf_oracle.h).
 * Instantiates pf_oracle_impl.h for dim = 2 and dim = 3, like the reference's
 * template instantiation at cracks.cc:4648-4657.
 */
#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "pf_oracle.h"

#define DIM 2
#define SUF(name) name##_2d
#include "pf_oracle_impl.h"
#undef DIM
#undef SUF

#define DIM 3
#define SUF(name) name##_3d
#include "pf_oracle_impl.h"
#undef DIM
#undef SUF

int
pfo_num_threads (void)
{
#ifdef _OPENMP
  return omp_get_max_threads ();
#else
  return 1;
#endif
}

/* bench.py's CPU arm: torchrun exports OMP_NUM_THREADS=1 to every rank, which would make the
 * baseline a single-threaded one; the arm asks for all host cores explicitly. */
void
pfo_set_num_threads (int n)
{
#ifdef _OPENMP
  if (n > 0)
    omp_set_num_threads (n);
#else
  (void) n;
#endif
}
