// pf_vector.cuh -- vector, constraint and active-set kernels around the apply:
// the glue of newton_active_set() (cracks.cc:2780-2994) and of SolverGMRES
// (cracks.cc:2764-2771) that the reference runs through Trilinos vectors.
#pragma once
#include "pf_common.cuh"

namespace pf {

constexpr int RED_BLOCKS = 592; // 4 x 148 SMs
constexpr int RED_THREADS = 256;

__device__ __forceinline__ double
block_reduce_sum (double v, double *sh /* [RED_THREADS/32] */)
{
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_down_sync (0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads ();
  if (l == 0)
    sh[w] = v;
  __syncthreads ();
  double s = 0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int) (blockDim.x >> 5); ++i)
      s += sh[i];
  return s; // valid on thread 0
}

__device__ __forceinline__ double
block_reduce_max (double v, double *sh)
{
  for (int o = 16; o > 0; o >>= 1)
    v = fmax (v, __shfl_down_sync (0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads ();
  if (l == 0)
    sh[w] = v;
  __syncthreads ();
  double s = 0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int) (blockDim.x >> 5); ++i)
      s = fmax (s, sh[i]);
  return s;
}

// ---- layout conversion: block layout [u | phi] <-> node-major interleaved
template <int DIM>
__global__ void
k_block_to_nodal (long long n_nodes, const double *__restrict__ ub, const double *__restrict__ pb,
                  double *__restrict__ v)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes)
    return;
  for (int c = 0; c < DIM; ++c)
    v[n * (DIM + 1) + c] = ub[n * DIM + c];
  v[n * (DIM + 1) + DIM] = pb[n];
}

template <int DIM>
__global__ void
k_nodal_to_block (long long n_nodes, const double *__restrict__ v, double *__restrict__ ub,
                  double *__restrict__ pb)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes)
    return;
  for (int c = 0; c < DIM; ++c)
    ub[n * DIM + c] = v[n * (DIM + 1) + c];
  pb[n] = v[n * (DIM + 1) + DIM];
}

// masks: block-layout bytes -> one byte per node with a bit per component
template <int DIM>
__global__ void
k_mask_from_block (long long n_nodes, const uint8_t *__restrict__ ub, const uint8_t *__restrict__ pb,
                   int set_u, int set_p, uint8_t *__restrict__ mask)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes)
    return;
  uint8_t m = mask[n];
  if (set_u)
    {
      m &= (uint8_t) (1u << DIM);
      for (int c = 0; c < DIM; ++c)
        if (ub[n * DIM + c])
          m |= (uint8_t) (1u << c);
    }
  if (set_p)
    {
      m &= (uint8_t) ~(1u << DIM);
      if (pb[n])
        m |= (uint8_t) (1u << DIM);
    }
  mask[n] = m;
}

// Dirichlet rows: every displacement component on every face of the box
template <int DIM>
__global__ void
k_mask_dirichlet_faces (Grid g, uint8_t *__restrict__ mask)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= g.n_local_nodes)
    return;
  long long rem = n;
  bool on = false;
  for (int d = 0; d < DIM; ++d)
    {
      int i;
      if (d < DIM - 1)
        {
          i = (int) (rem % g.nn[d]);
          rem /= g.nn[d];
        }
      else
        i = (int) rem + g.plane_begin;
      on = on || i == 0 || i == g.nn[d] - 1;
    }
  uint8_t m = mask[n] & (uint8_t) (1u << DIM);
  if (on)
    m |= (uint8_t) ((1u << DIM) - 1u);
  mask[n] = m;
}

// nodal extrapolation phi~ = phi_oo + ct (phi_o - phi_oo), cracks.cc:2268-2269
// (or phi_o when use_old_timestep_pf, 2276-2277); the clamp happens at the
// q-points after interpolation, like the reference.
template <int DIM>
__global__ void
k_extrapolate (long long n_nodes, double ct, int use_old, const double *__restrict__ old,
               const double *__restrict__ oldold, double *__restrict__ pt)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes)
    return;
  const double po = old[n * (DIM + 1) + DIM], poo = oldold[n * (DIM + 1) + DIM];
  pt[n] = use_old ? po : poo + ct * (po - poo);
}

// InitialValuesSneddon, cracks.cc:381-406
template <int DIM>
__global__ void
k_interpolate_sneddon (Grid g, double h_diam, double *__restrict__ sol)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= g.n_local_nodes)
    return;
  long long rem = n;
  double xc[3] = {0, 0, 0};
  for (int d = 0; d < DIM; ++d)
    {
      int i;
      if (d < DIM - 1)
        {
          i = (int) (rem % g.nn[d]);
          rem /= g.nn[d];
        }
      else
        i = (int) rem + g.plane_begin;
      xc[d] = g.origin[d] + g.h[d] * i;
    }
  const double r2 = (DIM == 2) ? xc[0] * xc[0] : xc[0] * xc[0] + xc[2] * xc[2];
  const bool broken = (r2 <= 1.0) && (fabs (2.0 * xc[1]) <= 2.0 * h_diam);
  for (int c = 0; c < DIM; ++c)
    sol[n * (DIM + 1) + c] = 0.0;
  sol[n * (DIM + 1) + DIM] = broken ? 0.0 : 1.0;
}

// y = constrained ? diag * x : 0   (positive diagonal that
// distribute_local_to_global keeps on constrained rows)
template <int DIM>
__global__ void
k_apply_init (long long n_nodes, const double *__restrict__ x, const double *__restrict__ diag,
              const uint8_t *__restrict__ mask, double *__restrict__ y)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes)
    return;
  const uint8_t m = mask[n];
  for (int c = 0; c <= DIM; ++c)
    {
      const long long i = n * (DIM + 1) + c;
      y[i] = ((m >> c) & 1) ? diag[i] * x[i] : 0.0;
    }
}

// r_pde = set_zero(r_total) and partial sums of ||r_pde||^2 over owned nodes
template <int DIM>
__global__ void __launch_bounds__ (RED_THREADS)
k_residual_finish (long long n_nodes, long long owned_lo, long long owned_hi,
                   const double *__restrict__ r_total, const uint8_t *__restrict__ mask,
                   double *__restrict__ r_pde, double *__restrict__ partial)
{
  __shared__ double sh[RED_THREADS / 32];
  double acc = 0;
  for (long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x; n < n_nodes;
       n += (long long) gridDim.x * blockDim.x)
    {
      const uint8_t m = mask[n];
      const bool owned = n >= owned_lo && n < owned_hi;
      for (int c = 0; c <= DIM; ++c)
        {
          const long long i = n * (DIM + 1) + c;
          const double v = ((m >> c) & 1) ? 0.0 : r_total[i];
          r_pde[i] = v;
          if (owned)
            acc += v * v;
        }
    }
  const double s = block_reduce_sum (acc, sh);
  if (threadIdx.x == 0)
    partial[blockIdx.x] = s;
}

// final reduction of per-block partials: out[j] = sum_b partial[j * nb + b]
__global__ void
k_reduce_partials (int nb, int nvals, const double *__restrict__ partial, double *__restrict__ out)
{
  __shared__ double sh[RED_THREADS / 32];
  for (int j = blockIdx.x; j < nvals; j += gridDim.x)
    {
      double acc = 0;
      for (int b = threadIdx.x; b < nb; b += blockDim.x)
        acc += partial[(long long) j * nb + b];
      const double s = block_reduce_sum (acc, sh);
      if (threadIdx.x == 0)
        out[j] = s;
    }
}

__global__ void
k_reduce_partials_max (int nb, const double *__restrict__ partial, double *__restrict__ out)
{
  __shared__ double sh[RED_THREADS / 32];
  double acc = 0;
  for (int b = threadIdx.x; b < nb; b += blockDim.x)
    acc = fmax (acc, partial[b]);
  const double s = block_reduce_max (acc, sh);
  if (threadIdx.x == 0)
    out[0] = s;
}

// active-set update, cracks.cc:2849-2885 (uniform mesh: no hanging nodes)
// counts[0] = owned active, counts[1] = owned cycling, counts[2] = changed
template <int DIM>
__global__ void
k_active_set (long long n_nodes, long long owned_lo, long long owned_hi, double c_scale,
              const double *__restrict__ r_total, const double *__restrict__ mass,
              const double *__restrict__ old, double *__restrict__ sol, int *__restrict__ cycle,
              uint8_t *__restrict__ mask, unsigned long long *__restrict__ counts)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes || n < owned_lo || n >= owned_hi)
    return;
  const long long i = n * (DIM + 1) + DIM;
  const double old_value = old[i], new_value = sol[i];
  const double gap = new_value - old_value;
  const uint8_t m = mask[n];
  if (m & PF_HANGING_BIT) // hanging nodes are skipped (cracks.cc:2855-2857); forest meshes only
    return;
  const bool was = (m >> DIM) & 1;
  int cyc = cycle[n];
  const bool inactive = (r_total[i] / mass[n] + c_scale * gap <= 0.0) && (cyc < 5);
  if (!inactive)
    {
      sol[i] = old_value;
      atomicAdd (&counts[0], 1ull);
      if (cyc >= 5)
        atomicAdd (&counts[1], 1ull);
    }
  // cycle detection: ++ for dofs that left the set (cracks.cc:2901-2907)
  if (was && inactive)
    cycle[n] = cyc + 1;
  if (was != !inactive)
    atomicAdd (&counts[2], 1ull);
  mask[n] = inactive ? (uint8_t) (m & ~(1u << DIM)) : (uint8_t) (m | (1u << DIM));
}

template <int DIM>
__global__ void
k_clear_active (long long n_nodes, uint8_t *__restrict__ mask, int *__restrict__ cycle)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes)
    return;
  mask[n] &= (uint8_t) ~(1u << DIM);
  cycle[n] = 0;
}

template <int DIM>
__global__ void
k_get_active (long long n_nodes, const uint8_t *__restrict__ mask, uint8_t *__restrict__ act)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n < n_nodes)
    act[n] = (mask[n] >> DIM) & 1;
}

template <int DIM>
__global__ void
k_project_phi (long long n_nodes, double *__restrict__ sol)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n < n_nodes)
    {
      const long long i = n * (DIM + 1) + DIM;
      sol[i] = fmax (0.0, fmin (sol[i], 1.0));
    }
}

// lumped mass of the phi block on a uniform mesh: vol/2^dim per adjacent cell
template <int DIM>
__global__ void
k_lumped_mass (Grid g, double *__restrict__ mass)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= g.n_local_nodes)
    return;
  long long rem = n;
  double cnt = 1, vol = 1;
  for (int d = 0; d < DIM; ++d)
    {
      int i;
      if (d < DIM - 1)
        {
          i = (int) (rem % g.nn[d]);
          rem /= g.nn[d];
        }
      else
        i = (int) rem + g.plane_begin;
      cnt *= (i == 0 || i == g.nn[d] - 1) ? 1.0 : 2.0;
      vol *= g.h[d];
    }
  mass[n] = cnt * (vol / (1 << DIM));
}

// ---- BLAS-1 style kernels for the Krylov solver --------------------------
__global__ void
k_axpy (long long n, double a, const double *__restrict__ x, double *__restrict__ y)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long) gridDim.x * blockDim.x)
    y[i] = fma (a, x[i], y[i]);
}

__global__ void
k_scale_copy (long long n, const double *__restrict__ a_dev, double a_mul, int invert,
              const double *__restrict__ x, double *__restrict__ y)
{
  // y = s * x with s = a_mul * (invert ? 1/sqrt(a_dev[0]) : 1)
  const double s = a_mul * (invert ? rsqrt (a_dev[0]) : 1.0);
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long) gridDim.x * blockDim.x)
    y[i] = s * x[i];
}

// z = x / diag on unconstrained rows, z = x / diag on constrained rows too
// (their operator row is diag): plain Jacobi
__global__ void
k_jacobi (long long n, const double *__restrict__ diag, const double *__restrict__ x,
          double *__restrict__ z)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long) gridDim.x * blockDim.x)
    z[i] = x[i] / diag[i];
}

// partial[j*nb + b] = sum over this block's chunk of V_j . w   (j < k)
// and partial[k*nb + b] = w . w.  V_j = V + j * stride.  Range [lo, hi).
template <int KB>
__global__ void __launch_bounds__ (RED_THREADS)
k_multi_dot (long long lo, long long hi, int j0, int k, const double *__restrict__ V, long long stride,
             const double *__restrict__ w, int with_norm, double *__restrict__ partial)
{
  __shared__ double sh[RED_THREADS / 32];
  double acc[KB], nrm = 0;
#pragma unroll
  for (int j = 0; j < KB; ++j)
    acc[j] = 0;
  for (long long i = lo + (long long) blockIdx.x * blockDim.x + threadIdx.x; i < hi;
       i += (long long) gridDim.x * blockDim.x)
    {
      const double wi = w[i];
#pragma unroll
      for (int j = 0; j < KB; ++j)
        if (j0 + j < k)
          acc[j] = fma (V[(j0 + j) * stride + i], wi, acc[j]);
      nrm = fma (wi, wi, nrm);
    }
#pragma unroll
  for (int j = 0; j < KB; ++j)
    if (j0 + j < k)
      {
        const double s = block_reduce_sum (acc[j], sh);
        if (threadIdx.x == 0)
          partial[(long long) (j0 + j) * gridDim.x + blockIdx.x] = s;
      }
  if (with_norm)
    {
      const double s = block_reduce_sum (nrm, sh);
      if (threadIdx.x == 0)
        partial[(long long) k * gridDim.x + blockIdx.x] = s;
    }
}

// w -= sum_j h[j] V_j over the whole local vector; hacc[j] += h[j] (thread 0)
__global__ void __launch_bounds__ (RED_THREADS)
k_multi_axpy (long long n, int k, const double *__restrict__ V, long long stride,
              const double *__restrict__ h, double *__restrict__ w)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long) gridDim.x * blockDim.x)
    {
      double wi = w[i];
      for (int j = 0; j < k; ++j)
        wi = fma (-h[j], V[j * stride + i], wi);
      w[i] = wi;
    }
}

// first Gram-Schmidt update and the dot products of the second pass in one sweep over the basis:
//   w <- w - sum_j h[j] V_j,   partial[j*nb + b] = V_j . w (new w),   partial[k*nb + b] = w . w (new w)
// The update is element-wise, so the second pass's dots can be taken of the updated element at once: the k basis
// vectors are read from DRAM once instead of twice (the second read of V_j[i] hits L1).  k <= KMAX.
// Dots over [lo, hi) only (owned entries), the update over the whole local vector [0, n).
template <int KMAX>
__global__ void __launch_bounds__ (RED_THREADS)
k_multi_axpy_dot (long long n, long long lo, long long hi, int k, const double *__restrict__ V, long long stride,
                  const double *__restrict__ h, double *__restrict__ w, double *__restrict__ partial)
{
  __shared__ double sh[RED_THREADS / 32];
  double acc[KMAX], nrm = 0, hj[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j)
    {
      acc[j] = 0;
      hj[j] = j < k ? h[j] : 0.0;
    }
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
    {
      double wi = w[i];
#pragma unroll
      for (int j = 0; j < KMAX; ++j)
        if (j < k)
          wi = fma (-hj[j], V[j * stride + i], wi);
      w[i] = wi;
      if (i >= lo && i < hi)
        {
#pragma unroll
          for (int j = 0; j < KMAX; ++j)
            if (j < k)
              acc[j] = fma (V[j * stride + i], wi, acc[j]);
          nrm = fma (wi, wi, nrm);
        }
    }
#pragma unroll
  for (int j = 0; j < KMAX; ++j)
    if (j < k)
      {
        const double s = block_reduce_sum (acc[j], sh);
        if (threadIdx.x == 0)
          partial[(long long) j * gridDim.x + blockIdx.x] = s;
      }
  const double s = block_reduce_sum (nrm, sh);
  if (threadIdx.x == 0)
    partial[(long long) k * gridDim.x + blockIdx.x] = s;
}

// second Gram-Schmidt pass and normalisation in one sweep: v_next = (w - sum_j h[j] V_j) / hk1 with
// hk1^2 = |w|^2 - sum_j h[j]^2 (h[k] = |w|^2 of the vector BEFORE this pass; the h[j] of a second pass are at
// round-off level, so the Pythagorean form loses nothing).  hk1 == 0: v_next = the unscaled vector (breakdown,
// detected by the host from the same numbers).
__global__ void __launch_bounds__ (RED_THREADS)
k_multi_axpy_scale (long long n, int k, const double *__restrict__ V, long long stride, const double *__restrict__ h,
                    const double *__restrict__ w, double *__restrict__ v_next)
{
  double s2 = h[k];
  for (int j = 0; j < k; ++j)
    s2 = fma (-h[j], h[j], s2);
  const double inv = s2 > 0 ? rsqrt (s2) : 1.0;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long) gridDim.x * blockDim.x)
    {
      double wi = w[i];
      for (int j = 0; j < k; ++j)
        wi = fma (-h[j], V[j * stride + i], wi);
      v_next[i] = wi * inv;
    }
}

// x += sum_j yv[j] V_j  (solution update; yv on device)
__global__ void __launch_bounds__ (RED_THREADS)
k_combine (long long n, int k, const double *__restrict__ V, long long stride,
           const double *__restrict__ yv, double *__restrict__ x)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long) gridDim.x * blockDim.x)
    {
      double xi = x[i];
      for (int j = 0; j < k; ++j)
        xi = fma (yv[j], V[j * stride + i], xi);
      x[i] = xi;
    }
}

// one component of a node-major vector <-> a vector with one value per node (compact Krylov basis of the phi stage)
__global__ void __launch_bounds__ (RED_THREADS)
k_gather_component (long long n_nodes, int nc, int comp, const double *__restrict__ src, double *__restrict__ dst)
{
  for (long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x; n < n_nodes; n += (long long) gridDim.x * blockDim.x)
    dst[n] = src[n * nc + comp];
}
__global__ void __launch_bounds__ (RED_THREADS)
k_scatter_component (long long n_nodes, int nc, int comp, const double *__restrict__ src, double *__restrict__ dst)
{
  for (long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x; n < n_nodes; n += (long long) gridDim.x * blockDim.x)
    dst[n * nc + comp] = src[n];
}

// constraint masks of the two stages of the block-triangular solve: every phi dof constrained (u stage), every u dof
// constrained (phi stage); bit c of a node's byte = component c constrained, phi is component dim
__global__ void
k_block_masks (long long n_nodes, int dim, const uint8_t *__restrict__ mask, uint8_t *__restrict__ mask_u,
               uint8_t *__restrict__ mask_phi)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes)
    return;
  const uint8_t m = mask[n], phi = (uint8_t) (1u << dim);
  mask_u[n] = m | phi;
  mask_phi[n] = m | (uint8_t) (phi - 1u);
}

template <int DIM>
__global__ void
k_zero_constrained (long long n_nodes, const uint8_t *__restrict__ mask, double *__restrict__ v)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes)
    return;
  const uint8_t m = mask[n];
  for (int c = 0; c <= DIM; ++c)
    if ((m >> c) & 1)
      v[n * (DIM + 1) + c] = 0.0;
}

__global__ void __launch_bounds__ (RED_THREADS)
k_absdiff_max (long long lo, long long hi, const double *__restrict__ a, const double *__restrict__ b,
               double *__restrict__ partial)
{
  __shared__ double sh[RED_THREADS / 32];
  double acc = 0;
  for (long long i = lo + (long long) blockIdx.x * blockDim.x + threadIdx.x; i < hi;
       i += (long long) gridDim.x * blockDim.x)
    acc = fmax (acc, fabs (a[i] - b[i]));
  const double s = block_reduce_max (acc, sh);
  if (threadIdx.x == 0)
    partial[blockIdx.x] = s;
}

template <int DIM>
__global__ void
k_set_component (long long n_nodes, int c, double value, double *__restrict__ v)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n < n_nodes)
    v[n * (DIM + 1) + c] = value;
}

// partial[b] = max over the block's nodes in [lo, hi) of (1 - phi): the refinement
// indicator of refine_mesh() looks for phi below a threshold (cracks.cc:3971-3995)
__global__ void __launch_bounds__ (RED_THREADS)
k_one_minus_phi_max (long long lo, long long hi, int ncomp, const double *__restrict__ sol,
                     double *__restrict__ partial)
{
  __shared__ double sh[RED_THREADS / 32];
  double acc = -1e300;
  for (long long n = lo + (long long) blockIdx.x * blockDim.x + threadIdx.x; n < hi;
       n += (long long) gridDim.x * blockDim.x)
    acc = fmax (acc, 1.0 - sol[n * ncomp + ncomp - 1]);
  const double s = block_reduce_max (acc, sh);
  if (threadIdx.x == 0)
    partial[blockIdx.x] = s;
}

} // namespace pf
