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

// ---- 2-D hierarchy (box meshes and the unit square with the slit of the Miehe tests) -------------------
// Nodes: the regular (n[0] x n[1]) grid, x fastest, followed by the upper copies of the slit nodes
// (slit_i0 + k, slit_j) -- the layout of Grid::slit_* (pf_common.cuh).  slit_j < 0: no slit.
struct Dims2
{
  int n[2];
  int slit_j, slit_i0;
  long long slit_base;
};

__device__ __forceinline__ long long
n_nodes2 (const Dims2 &d)
{
  return (long long) d.n[0] * d.n[1] + (d.slit_j >= 0 ? d.n[0] - d.slit_i0 : 0);
}

// node (i, j) as seen from the upper (upper = true) or the lower side of the slit
__device__ __forceinline__ long long
node_id2 (const Dims2 &d, int i, int j, bool upper)
{
  if (upper && j == d.slit_j && i >= d.slit_i0)
    return d.slit_base + (i - d.slit_i0);
  return i + (long long) d.n[0] * j;
}

// t-th node -> (i, j, side)
__device__ __forceinline__ void
node_ij2 (const Dims2 &d, long long t, int &i, int &j, bool &upper)
{
  const long long regular = (long long) d.n[0] * d.n[1];
  if (t < regular)
    {
      i = (int) (t % d.n[0]);
      j = (int) (t / d.n[0]);
      upper = d.slit_j >= 0 && j > d.slit_j;
    }
  else
    {
      i = d.slit_i0 + (int) (t - regular);
      j = d.slit_j;
      upper = true;
    }
}

// coarse node <- the fine node at the same place (same side of the slit)
template <int NCOMP, typename T>
__global__ void
k_inject2d (Dims2 dc, Dims2 df, const T *__restrict__ src, T *__restrict__ dst)
{
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_nodes2 (dc))
    return;
  int I, J;
  bool upper;
  node_ij2 (dc, t, I, J, upper);
  const long long f = node_id2 (df, 2 * I, 2 * J, upper);
  for (int c = 0; c < NCOMP; ++c)
    dst[t * NCOMP + c] = src[f * NCOMP + c];
}

// xf += P xc (bilinear; a fine node above the slit takes the upper copies of its coarse parents on the slit
// line, one below the lower ones), constrained fine dofs receive nothing.  3 components per node.
__global__ void
k_prolong_add2d (Dims2 dc, Dims2 df, const double *__restrict__ xc, const uint8_t *__restrict__ fmask,
                 double *__restrict__ xf)
{
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_nodes2 (df))
    return;
  int i, j;
  bool upper;
  node_ij2 (df, t, i, j, upper);
  const int i0 = i >> 1, j0 = j >> 1, ni = i & 1, nj = j & 1;
  const double w = (ni ? 0.5 : 1.0) * (nj ? 0.5 : 1.0);
  double acc[3] = {0, 0, 0};
  for (int c2 = 0; c2 <= nj; ++c2)
    for (int c1 = 0; c1 <= ni; ++c1)
      {
        const long long cn = node_id2 (dc, i0 + c1, j0 + c2, upper);
        for (int c = 0; c < 3; ++c)
          acc[c] += w * xc[3 * cn + c];
      }
  const uint8_t m = fmask[t];
  for (int c = 0; c < 3; ++c)
    if (!((m >> c) & 1))
      xf[3 * t + c] += acc[c];
}

// rc += P^T rf (the transpose of k_prolong_add2d as a scatter; rc zeroed by the caller, constrained fine rows
// count as zero; constrained coarse rows are zeroed afterwards with k_zero_constrained)
__global__ void
k_restrict_add2d (Dims2 dc, Dims2 df, const double *__restrict__ rf, const uint8_t *__restrict__ fmask,
                  double *__restrict__ rc)
{
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_nodes2 (df))
    return;
  int i, j;
  bool upper;
  node_ij2 (df, t, i, j, upper);
  const uint8_t m = fmask[t];
  double r[3];
  for (int c = 0; c < 3; ++c)
    r[c] = ((m >> c) & 1) ? 0.0 : rf[3 * t + c];
  const int i0 = i >> 1, j0 = j >> 1, ni = i & 1, nj = j & 1;
  const double w = (ni ? 0.5 : 1.0) * (nj ? 0.5 : 1.0);
  for (int c2 = 0; c2 <= nj; ++c2)
    for (int c1 = 0; c1 <= ni; ++c1)
      {
        const long long cn = node_id2 (dc, i0 + c1, j0 + c2, upper);
        for (int c = 0; c < 3; ++c)
          atomicAdd (&rc[3 * cn + c], w * r[c]);
      }
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
