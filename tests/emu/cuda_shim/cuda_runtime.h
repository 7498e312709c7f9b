// cuda_runtime.h -- TEST SHIM, not the CUDA header.  Lets g++ compile the kernel sources under
// cracks_b200/csrc/*.cuh for the host so that tests/test_kernel_emulation_cpu.py can execute
// thread-per-item kernels sequentially on the CPU (one "thread" at a time) and hold them against
// the oracle without a GPU.  Only kernels that do not use shared memory, barriers or warp shuffles
// may be *run* this way; the others merely have to compile.  Nothing under cracks_b200/ includes this.
#pragma once
#define PF_EMULATION 1 // kernel sources take their plain-C++ path where the device path is inline PTX
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
#define __shared__ static
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

inline double
atomicAdd (double *p, double v)
{
  const double old = *p;
  *p = old + v;
  return old;
}
inline unsigned long long
atomicAdd (unsigned long long *p, unsigned long long v)
{
  const unsigned long long old = *p;
  *p = old + v;
  return old;
}
// compile-only stand-ins (kernels using them are never executed by the emulation)
inline void __syncthreads () {}
inline void __syncwarp () {}
template <class T>
inline T
__shfl_down_sync (unsigned, T v, int)
{
  return v;
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
