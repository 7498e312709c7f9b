// cuda_runtime.h -- TEST SHIM for block-cooperative kernels (shared memory, __syncthreads, __syncwarp,
// atomics from concurrent threads).  Every CUDA thread of a block is an OS thread, __syncthreads() is a
// std::barrier over the block, the dynamic shared-memory array is one buffer per block, atomicAdd is a
// real atomic.  Blocks run one after the other.  See tests/emu/emu_tiled.cc.
#pragma once
#define PF_EMULATION 1 // kernel sources take their plain-C++ path where the device path is inline PTX
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __shared__
#define __align__(n)

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
struct float4
{
  float x, y, z, w;
};
// packed FP32 intrinsics of sm_100 (pf_apply3d_v6.cuh: two cells per thread), lane by lane
inline float2 make_float2 (float x, float y) { float2 r; r.x = x; r.y = y; return r; }
inline float2 __ffma2_rn (float2 a, float2 b, float2 c) { return make_float2 (std::fma (a.x, b.x, c.x), std::fma (a.y, b.y, c.y)); }
inline float2 __fadd2_rn (float2 a, float2 b) { return make_float2 (a.x + b.x, a.y + b.y); }
inline float2 __fmul2_rn (float2 a, float2 b) { return make_float2 (a.x * b.x, a.y * b.y); }

struct double4
{
  double x, y, z, w;
};
inline double4
make_double4 (double x, double y, double z, double w)
{
  return {x, y, z, w};
}

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
extern std::barrier<> *emu_block_barrier;

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
  emu_block_barrier->arrive_and_wait ();
}
// the kernels call __syncwarp() from uniform control flow: a block barrier is a valid (stronger) stand-in
inline void
__syncwarp ()
{
  emu_block_barrier->arrive_and_wait ();
}
template <class T>
inline T
__shfl_down_sync (unsigned, T v, int)
{
  return v; // compile-only: kernels with warp shuffles are not executed by the emulation
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
