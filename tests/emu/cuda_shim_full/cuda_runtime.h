// cuda_runtime.h -- TEST SHIM (full): enough of the CUDA runtime API and of the device-side vocabulary
// for g++ to compile a pre-processed copy of cracks_b200/csrc/pf_api.cu (tests/emu/build_emulated_library.py
// rewrites the <<<...>>> launches into pf_emu::launch calls) and run the WHOLE library on the CPU:
// "device" memory is host memory, streams and events are no-ops, a kernel launch executes its blocks one
// after the other -- sequentially thread by thread for thread-per-item kernels, with one OS thread per CUDA
// thread, a std::barrier for __syncthreads / __syncwarp / warp shuffles and std::atomic_ref for atomicAdd for
// the cooperative ones.  Test infrastructure only (tests/test_emulated_library_cpu.py): it is a checker of
// the library's host logic and kernel sources, never a fallback -- nothing under cracks_b200/ refers to it.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <tuple>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n)
#define __grid_constant__

struct uint3
{
  unsigned x, y, z;
};
struct dim3
{
  unsigned x = 1, y = 1, z = 1;
};
struct double2
{
  double x, y;
};
struct double4
{
  double x, y, z, w;
};
struct alignas (16) float4
{
  float x, y, z, w;
};
inline double4
make_double4 (double x, double y, double z, double w)
{
  return {x, y, z, w};
}

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

namespace pf_emu {
extern std::barrier<> *block_barrier; // null while a kernel runs in sequential mode
extern double shfl_slots[1024];
[[noreturn]] void fail (const char *what);
inline void
sync ()
{
  if (!block_barrier)
    fail ("a kernel launched in sequential mode reached a barrier / shuffle: add it to COOPERATIVE in "
          "tests/emu/build_emulated_library.py");
  block_barrier->arrive_and_wait ();
}
} // namespace pf_emu

inline double
atomicAdd (double *p, double v)
{
  std::atomic_ref<double> a (*p);
  double old = a.load (std::memory_order_relaxed);
  while (!a.compare_exchange_weak (old, old + v, std::memory_order_relaxed))
    {
    }
  return old;
}
inline float
atomicAdd (float *p, float v)
{
  std::atomic_ref<float> a (*p);
  float old = a.load (std::memory_order_relaxed);
  while (!a.compare_exchange_weak (old, old + v, std::memory_order_relaxed))
    {
    }
  return old;
}
inline unsigned long long
atomicAdd (unsigned long long *p, unsigned long long v)
{
  return std::atomic_ref<unsigned long long> (*p).fetch_add (v);
}
inline void
__syncthreads ()
{
  pf_emu::sync ();
}
inline void
__syncwarp ()
{
  pf_emu::sync (); // called from uniform control flow in these kernels: a block barrier is a valid stand-in
}
// every thread of the block calls the shuffle (true for all reductions in these kernels)
inline double
__shfl_down_sync (unsigned, double v, int offset)
{
  const unsigned t = threadIdx.x;
  pf_emu::shfl_slots[t] = v;
  pf_emu::sync ();
  const unsigned lane = t & 31u;
  const double r = (lane + (unsigned) offset < 32u && t + (unsigned) offset < blockDim.x) ? pf_emu::shfl_slots[t + offset] : v;
  pf_emu::sync ();
  return r;
}
inline double2
make_double2 (double x, double y)
{
  return {x, y};
}
inline double
__longlong_as_double (long long v)
{
  double d;
  std::memcpy (&d, &v, sizeof d);
  return d;
}
inline double
rsqrt (double x)
{
  return 1.0 / std::sqrt (x);
}
using std::fabs;
using std::fma;
using std::fmax;
using std::fmin;
using std::sqrt;

// ---- runtime API ---------------------------------------------------------------------------------------
typedef int cudaError_t;
enum
{
  cudaSuccess = 0,
  cudaErrorNotSupported = 801
};
typedef struct pf_emu_stream *cudaStream_t;
typedef struct pf_emu_event *cudaEvent_t;
typedef struct pf_emu_graph *cudaGraph_t;
typedef struct pf_emu_graph_exec *cudaGraphExec_t;
struct cudaGraphExecUpdateResultInfo
{
  int result;
};
enum cudaMemcpyKind
{
  cudaMemcpyHostToDevice,
  cudaMemcpyDeviceToHost,
  cudaMemcpyDeviceToDevice,
  cudaMemcpyHostToHost
};
enum
{
  cudaStreamNonBlocking = 1,
  cudaEventDisableTiming = 2,
  cudaStreamCaptureModeThreadLocal = 1,
  cudaFuncAttributeMaxDynamicSharedMemorySize = 8,
  cudaFuncAttributePreferredSharedMemoryCarveout = 9,
  cudaDevAttrMultiProcessorCount = 16
};
inline const char *
cudaGetErrorString (cudaError_t e)
{
  return e == cudaSuccess ? "no error" : "emulation: unsupported CUDA call";
}
inline cudaError_t cudaGetLastError () { return cudaSuccess; }
inline cudaError_t cudaSetDevice (int) { return cudaSuccess; }
template <class T>
inline cudaError_t
cudaMalloc (T **p, size_t bytes)
{
  *p = static_cast<T *> (std::calloc (bytes ? bytes : 1, 1));
  return *p ? cudaSuccess : 2;
}
template <class T>
inline cudaError_t
cudaMallocHost (T **p, size_t bytes)
{
  return cudaMalloc (p, bytes);
}
inline cudaError_t
cudaFree (void *p)
{
  std::free (p);
  return cudaSuccess;
}
inline cudaError_t cudaFreeHost (void *p) { return cudaFree (p); }
inline cudaError_t
cudaMemcpy (void *d, const void *s, size_t n, cudaMemcpyKind)
{
  std::memmove (d, s, n);
  return cudaSuccess;
}
inline cudaError_t
cudaMemcpyAsync (void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr)
{
  std::memmove (d, s, n);
  return cudaSuccess;
}
inline cudaError_t
cudaMemsetAsync (void *d, int v, size_t n, cudaStream_t = nullptr)
{
  std::memset (d, v, n);
  return cudaSuccess;
}
inline cudaError_t
cudaStreamCreateWithFlags (cudaStream_t *s, unsigned)
{
  *s = reinterpret_cast<cudaStream_t> (std::malloc (1));
  return cudaSuccess;
}
inline cudaError_t
cudaStreamDestroy (cudaStream_t s)
{
  std::free (s);
  return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize (cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent (cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t
cudaEventCreateWithFlags (cudaEvent_t *e, unsigned)
{
  *e = reinterpret_cast<cudaEvent_t> (std::malloc (1));
  return cudaSuccess;
}
inline cudaError_t cudaEventCreate (cudaEvent_t *e) { return cudaEventCreateWithFlags (e, 0); }
inline cudaError_t
cudaEventDestroy (cudaEvent_t e)
{
  std::free (e);
  return cudaSuccess;
}
inline cudaError_t cudaEventRecord (cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t
cudaEventElapsedTime (float *ms, cudaEvent_t, cudaEvent_t)
{
  *ms = 0.f;
  return cudaSuccess;
}
template <class F>
inline cudaError_t
cudaFuncSetAttribute (F, int, int)
{
  return cudaSuccess;
}
inline cudaError_t
cudaDeviceGetAttribute (int *v, int, int)
{
  *v = 148;
  return cudaSuccess;
}
// graphs: not emulated (PF_MG_GRAPH is opt-in)
inline cudaError_t cudaStreamBeginCapture (cudaStream_t, int) { return cudaErrorNotSupported; }
inline cudaError_t cudaStreamEndCapture (cudaStream_t, cudaGraph_t *) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphInstantiate (cudaGraphExec_t *, cudaGraph_t, int) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphExecUpdate (cudaGraphExec_t, cudaGraph_t, cudaGraphExecUpdateResultInfo *) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphLaunch (cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphDestroy (cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy (cudaGraphExec_t) { return cudaSuccess; }

// ---- kernel launch ---------------------------------------------------------------------------------------
namespace pf_emu {
extern long long launches_sequential, launches_cooperative;

// persistent workers, one set per block size: a cooperative launch hands every block to the same OS
// threads instead of spawning blockDim threads per block
struct Pool
{
  explicit Pool (unsigned n);
  ~Pool ();
  void run_block (void (*trampoline) (void *, unsigned), void *job);
  unsigned n;
  std::barrier<> start, done, inner;
  std::vector<std::thread> workers;
  void (*fn) (void *, unsigned) = nullptr;
  void *job = nullptr;
  bool stop = false;
};
Pool &pool_for (unsigned block);

template <class K, class... A>
void
launch (bool cooperative, K kernel, unsigned grid, unsigned block, size_t /*smem*/, cudaStream_t, A... args)
{
  if (!cooperative)
    {
      ++launches_sequential;
      block_barrier = nullptr;
      gridDim.x = grid;
      blockDim.x = block;
      for (unsigned b = 0; b < grid; ++b)
        for (unsigned t = 0; t < block; ++t)
          {
            blockIdx.x = b;
            threadIdx.x = t;
            kernel (args...);
          }
      return;
    }
  ++launches_cooperative;
  if (block > 1024)
    fail ("block too large for the emulation");
  Pool &pool = pool_for (block);
  block_barrier = &pool.inner;
  struct Job
  {
    unsigned grid, block, b;
    decltype (std::make_tuple (args...)) a;
    K k;
  } job{grid, block, 0, std::make_tuple (args...), kernel};
  auto trampoline = [](void *p, unsigned t) {
    Job &j = *static_cast<Job *> (p);
    gridDim.x = j.grid;
    blockDim.x = j.block;
    blockIdx.x = j.b;
    threadIdx.x = t;
    std::apply (j.k, j.a);
  };
  for (unsigned b = 0; b < grid; ++b)
    {
      job.b = b;
      pool.run_block (trampoline, &job);
    }
  block_barrier = nullptr;
}
} // namespace pf_emu
