// cuda_runtime.h -- TEST SHIM (full): enough of the CUDA runtime API and of the device-side vocabulary
// for g++ to compile a pre-processed copy of cracks_b200/csrc/pf_api.cu (tests/emu/build_emulated_library.py
// rewrites the <<<...>>> launches into pf_emu::launch calls) and run the WHOLE library on the CPU:
// "device" memory is host memory, streams and events are no-ops, a kernel launch hands its blocks to a small
// pool of OS threads.  Inside a block the CUDA threads run one after the other -- as plain calls for
// thread-per-item kernels, as fibers (ucontext) that yield at __syncthreads / __syncwarp / warp shuffles for the
// cooperative ones; shared memory is thread_local storage of the OS thread that runs the block, atomicAdd is
// std::atomic_ref (blocks run concurrently).  Test infrastructure only (tests/test_emulated_library_cpu.py): it is a checker of
// the library's host logic and kernel sources, never a fallback -- nothing under cracks_b200/ refers to it.
#pragma once
#define PF_EMULATION 1 // kernel sources take their plain-C++ path where the device path is inline PTX
#include <atomic>
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
#define __maxnreg__(...)
#define __shared__ static thread_local
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
struct float2
{
  float x, y;
};
struct double4
{
  double x, y, z, w;
};
struct alignas (16) float4
{
  float x, y, z, w;
};
// packed FP32 intrinsics of sm_100 (pf_apply3d_v6.cuh: two cells per thread), lane by lane
inline float2 make_float2 (float x, float y) { float2 r; r.x = x; r.y = y; return r; }
inline float2 __ffma2_rn (float2 a, float2 b, float2 c) { return make_float2 (std::fma (a.x, b.x, c.x), std::fma (a.y, b.y, c.y)); }
inline float2 __fadd2_rn (float2 a, float2 b) { return make_float2 (a.x + b.x, a.y + b.y); }
inline float2 __fmul2_rn (float2 a, float2 b) { return make_float2 (a.x * b.x, a.y * b.y); }

inline double4
make_double4 (double x, double y, double z, double w)
{
  return {x, y, z, w};
}

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

namespace pf_emu {
extern thread_local double shfl_slots[1024];
[[noreturn]] void fail (const char *what);
void sync (); // block barrier: the running fiber yields until every thread of the block has arrived
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

// one block: run `body (job, t)` for t = 0 .. block-1 on the calling OS thread; cooperative blocks as fibers
void run_block (bool cooperative, unsigned block, void (*body) (void *, unsigned), void *job);
// all blocks of a launch on the worker pool; returns when the last one has finished
void run_grid (unsigned grid, unsigned block, void (*per_block) (void *, unsigned), void *job);

template <class K, class... A>
void
launch (bool cooperative, K kernel, unsigned grid, unsigned block, size_t /*smem*/, cudaStream_t, A... args)
{
  if (block > 1024)
    fail ("block too large for the emulation");
  ++(cooperative ? launches_cooperative : launches_sequential);
  struct Job
  {
    unsigned grid, block;
    bool cooperative;
    decltype (std::make_tuple (args...)) a;
    K k;
  } job{grid, block, cooperative, std::make_tuple (args...), kernel};
  struct BlockJob
  {
    Job *j;
    unsigned b;
  };
  auto per_block = [](void *p, unsigned b) {
    Job &j = *static_cast<Job *> (p);
    BlockJob bj{&j, b};
    run_block (j.cooperative, j.block,
               [](void *q, unsigned t) {
                 BlockJob &x = *static_cast<BlockJob *> (q);
                 gridDim.x = x.j->grid;
                 blockDim.x = x.j->block;
                 blockIdx.x = x.b;
                 threadIdx.x = t;
                 std::apply (x.j->k, x.j->a);
               },
               &bj);
  };
  run_grid (grid, block, per_block, &job);
}
} // namespace pf_emu
