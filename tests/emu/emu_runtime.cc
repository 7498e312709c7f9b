// emu_runtime.cc -- definitions behind tests/emu/cuda_shim_full/cuda_runtime.h (test infrastructure)
#include <cuda_runtime.h>

#include <cstdio>
#include <map>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

namespace pf {
// `extern __shared__ unsigned char smem_raw[]` of the tiled kernels; blocks run one after the other
alignas (16) unsigned char smem_raw[256 * 1024];
} // namespace pf

namespace pf_emu {
std::barrier<> *block_barrier = nullptr;
double shfl_slots[1024];
long long launches_sequential = 0, launches_cooperative = 0;

Pool::Pool (unsigned n_) : n (n_), start ((std::ptrdiff_t) n_ + 1), done ((std::ptrdiff_t) n_ + 1), inner ((std::ptrdiff_t) n_)
{
  for (unsigned t = 0; t < n; ++t)
    workers.emplace_back ([this, t]() {
      for (;;)
        {
          start.arrive_and_wait ();
          if (stop)
            return;
          fn (job, t);
          done.arrive_and_wait ();
        }
    });
}

Pool::~Pool ()
{
  stop = true;
  start.arrive_and_wait ();
  for (auto &w : workers)
    w.join ();
}

void
Pool::run_block (void (*f) (void *, unsigned), void *j)
{
  fn = f;
  job = j;
  start.arrive_and_wait ();
  done.arrive_and_wait ();
}

Pool &
pool_for (unsigned block)
{
  static std::map<unsigned, Pool *> pools; // leaked on purpose: workers must outlive static destruction order
  auto it = pools.find (block);
  if (it == pools.end ())
    it = pools.emplace (block, new Pool (block)).first;
  return *it->second;
}

void
fail (const char *what)
{
  std::fprintf (stderr, "pf_emu: %s\n", what);
  std::abort ();
}
} // namespace pf_emu

extern "C" void
pf_emu_launch_counts (long long *sequential, long long *cooperative)
{
  *sequential = pf_emu::launches_sequential;
  *cooperative = pf_emu::launches_cooperative;
}
