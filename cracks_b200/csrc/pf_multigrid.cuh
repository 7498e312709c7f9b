// pf_multigrid.cuh -- transfer and smoother kernels of the matrix-free geometric
// multigrid preconditioner that stands in for the reference's two Trilinos ML
// AMG hierarchies (cracks.cc:2477-2497, applied at 2717-2740).  The hierarchy
// is the uniform 10 * 2^l family the Sneddon meshes come from (cracks.cc:1248,
// 1534); coarse operators are re-discretisations with the injected state, the
// smoother is Chebyshev-Jacobi on the matrix-free diagonal.  The arithmetic of
// the preconditioner is unpinned by the reference's goldens (only #LinIts
// depends on it).
#pragma once
#include "pf_common.cuh"

namespace pf {

struct Dims3
{
  int n[3]; // global nodes per direction
  int kb;   // first node plane of the slowest coordinate held locally (z-slab decomposition)
};

// local index of global node (i,j,k); the caller guarantees that plane k is local
__device__ __forceinline__ long long
node_id (const Dims3 &d, int i, int j, int k)
{
  return i + (long long) d.n[0] * (j + (long long) d.n[1] * (k - d.kb));
}

// coarse node (I,J,K) <- fine node (2I,2J,2K)
template <int NCOMP, typename T>
__global__ void
k_inject (Dims3 dc, Dims3 df, int Ka, int Ke, const T *__restrict__ src, T *__restrict__ dst)
{
  // coarse planes [Ka, Ke): the caller picks those whose fine plane 2K is local
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long) dc.n[0] * dc.n[1] * (Ke - Ka);
  if (t >= total)
    return;
  const int i = (int) (t % dc.n[0]), j = (int) ((t / dc.n[0]) % dc.n[1]), k = Ka + (int) (t / ((long long) dc.n[0] * dc.n[1]));
  const long long n = node_id (dc, i, j, k);
  const long long f = node_id (df, 2 * i, 2 * j, 2 * k);
  for (int c = 0; c < NCOMP; ++c)
    dst[n * NCOMP + c] = src[f * NCOMP + c];
}

// xf += P xc (trilinear), constrained fine dofs receive nothing
__global__ void
k_prolong_add (Dims3 dc, Dims3 df, int ka, int ke, const double *__restrict__ xc,
               const uint8_t *__restrict__ fmask, double *__restrict__ xf)
{
  // fine planes [ka, ke); the coarse planes k/2 and (k+1)/2 must be local
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long) df.n[0] * df.n[1] * (ke - ka);
  if (t >= total)
    return;
  const int i = (int) (t % df.n[0]), j = (int) ((t / df.n[0]) % df.n[1]), k = ka + (int) (t / ((long long) df.n[0] * df.n[1]));
  const long long n = node_id (df, i, j, k);
  const int i0 = i >> 1, j0 = j >> 1, k0 = k >> 1;
  const int ni = i & 1, nj = j & 1, nk = k & 1; // odd -> average of two coarse neighbours
  double acc[4] = {0, 0, 0, 0};
  for (int c3 = 0; c3 <= nk; ++c3)
    for (int c2 = 0; c2 <= nj; ++c2)
      for (int c1 = 0; c1 <= ni; ++c1)
        {
          const double w = (ni ? 0.5 : 1.0) * (nj ? 0.5 : 1.0) * (nk ? 0.5 : 1.0);
          const long long cn = node_id (dc, i0 + c1, j0 + c2, k0 + c3);
          const double4 v = *reinterpret_cast<const double4 *> (xc + 4 * cn);
          acc[0] += w * v.x;
          acc[1] += w * v.y;
          acc[2] += w * v.z;
          acc[3] += w * v.w;
        }
  const uint8_t m = fmask[n];
  for (int c = 0; c < 4; ++c)
    if (!((m >> c) & 1))
      xf[4 * n + c] += acc[c];
}

// rc = P^T rf with constrained fine rows treated as zero; constrained coarse rows get zero
__global__ void
k_restrict (Dims3 dc, Dims3 df, int Ka, int Ke, const double *__restrict__ rf, const uint8_t *__restrict__ fmask,
            const uint8_t *__restrict__ cmask, double *__restrict__ rc)
{
  // coarse planes [Ka, Ke); the fine planes 2K-1 .. 2K+1 (where they exist) must be local
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long) dc.n[0] * dc.n[1] * (Ke - Ka);
  if (t >= total)
    return;
  const int I = (int) (t % dc.n[0]), J = (int) ((t / dc.n[0]) % dc.n[1]), K = Ka + (int) (t / ((long long) dc.n[0] * dc.n[1]));
  const long long n = node_id (dc, I, J, K);
  double acc[4] = {0, 0, 0, 0};
  for (int dk = -1; dk <= 1; ++dk)
    for (int dj = -1; dj <= 1; ++dj)
      for (int di = -1; di <= 1; ++di)
        {
          const int i = 2 * I + di, j = 2 * J + dj, k = 2 * K + dk;
          if (i < 0 || j < 0 || k < 0 || i >= df.n[0] || j >= df.n[1] || k >= df.n[2])
            continue;
          const double w = (di ? 0.5 : 1.0) * (dj ? 0.5 : 1.0) * (dk ? 0.5 : 1.0);
          const long long fn = node_id (df, i, j, k);
          const uint8_t m = fmask[fn];
          const double4 v = *reinterpret_cast<const double4 *> (rf + 4 * fn);
          acc[0] += ((m >> 0) & 1) ? 0.0 : w * v.x;
          acc[1] += ((m >> 1) & 1) ? 0.0 : w * v.y;
          acc[2] += ((m >> 2) & 1) ? 0.0 : w * v.z;
          acc[3] += ((m >> 3) & 1) ? 0.0 : w * v.w;
        }
  const uint8_t m = cmask[n];
  for (int c = 0; c < 4; ++c)
    rc[4 * n + c] = ((m >> c) & 1) ? 0.0 : acc[c];
}

// one Chebyshev step, fused: r = b - y (y = A x, or r = b when first), d = c1 d + c2 r / diag, x += d
__global__ void
k_cheb_step (long long n, int first, double c1, double c2, const double *__restrict__ b,
             const double *__restrict__ y, const double *__restrict__ diag, double *__restrict__ d,
             double *__restrict__ x)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long) gridDim.x * blockDim.x)
    {
      const double r = first ? b[i] : b[i] - y[i];
      const double dn = first ? c2 * r / diag[i] : fma (c1, d[i], c2 * r / diag[i]);
      d[i] = dn;
      x[i] = first ? dn : x[i] + dn;
    }
}

__global__ void
k_sub (long long n, const double *__restrict__ b, const double *__restrict__ y, double *__restrict__ r)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long) gridDim.x * blockDim.x)
    r[i] = b[i] - y[i];
}

// deterministic pseudo-random start vector for the power iteration
__global__ void
k_fill_hash (long long n, double *__restrict__ v)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long) gridDim.x * blockDim.x)
    {
      unsigned long long h = (unsigned long long) i * 0x9E3779B97F4A7C15ull;
      h ^= h >> 29;
      h *= 0xBF58476D1CE4E5B9ull;
      h ^= h >> 32;
      v[i] = (double) (h & 0xFFFFF) / 1048576.0 - 0.5;
    }
}

} // namespace pf
