// pf_apply3d_v6.cuh -- fourth tuning step of the hot kernel: y += J(U) x on cubic cells with the exact
// 27-point rule (cracks.cc:2235-2389), same mapping as pf_apply3d_v4.cuh (staged z-collapsed node columns,
// one thread per cell, plane -> row -> point walk, transposed collapse into a shared y tile).
//
// The kernel is bound by the FP64 pipe (profiles/: 3290 FP64 instructions per cell, DRAM at 6 % of its peak),
// and 40 % of those instructions evaluate quantities that depend on the linearisation state U only -- they are
// the same in every Jacobian application of a Newton step (about 12 Krylov iterations).  v6 trades idle HBM
// bandwidth for FP64 instructions:
//   * k_point_coeffs (once per pf_setup_jacobian) stores two scalars per quadrature point,
//       wg = JxW g(phi~) 2 mu gamma^2     the degraded elastic weight of the (u,u) block (cracks.cc:2359-2364,
//                                         clamp of phi~ at the point included, 2270-2273),
//       c2 = JxW/8 [(1-kappa) sigma(u):E(u) + G_c/eps - 2 (alpha-1) p div u]   the (phi,phi) reaction
//                                         coefficient (cracks.cc:2377-2382),
//     27 x 16 bytes per cell, laid out [tile][point][cell of the tile] so that a warp reads 512 contiguous bytes;
//   * the point loop keeps only what depends on x: with the symmetric strain of x in units of 1/gamma,
//       sigma'(x) = G + lam2 tr(G) I                      (stress in units of 2 mu gamma^2, one fma per diagonal entry)
//       a = pf K1 w [ (lam2 tr(U) - beta) tr(G) + U:G ] + dphi c2   (the phi row, cracks.cc:2375-2382)
//     and every accumulation of the transposed x-collapse is one fma with the point weight folded in.
// Instruction count per cell (SASS, see profiles/kernels.json): 3290 -> about 2300.
//
// R = double: the exact operator (Krylov operator, parity <= 1e-12 against the oracle).
// R = float : the same kernel in FP32 for the inexact-Newton Jacobian (pf_set_jacobian_precision); global vectors
//             stay FP64 (converted on load, z-differences formed in FP64 before the conversion), the residual that
//             defines the Newton fixed point is never evaluated in reduced precision.
// R = f32x2 : two x-adjacent cells per thread in packed FP32 (FFMA2 / FADD2 / FMUL2, sm_100 only): the FP32 kernels
//             are bound by instruction issue, not by the FP32 pipe, and a packed instruction does the work of two.
#pragma once
#include "pf_apply3d_v2.cuh"

namespace pf {

template <typename R> struct Pair;
template <> struct Pair<double> { using type = double2; };
template <> struct Pair<float> { using type = float2; };
template <typename V> struct Quad;
template <> struct Quad<double> { using type = double4; };
template <> struct Quad<float> { using type = float4; };

__device__ __forceinline__ double fma_r (double a, double b, double c) { return fma (a, b, c); }
__device__ __forceinline__ float fma_r (float a, float b, float c) { return fmaf (a, b, c); }

// ---- packed pair of floats: lane x = left cell, lane y = right cell of a thread's pair ---------------------------
struct f32x2
{
  float2 v;
  __device__ __forceinline__ f32x2 () {}
  __device__ __forceinline__ f32x2 (float a) { v = make_float2 (a, a); }
  __device__ __forceinline__ f32x2 (double a) { v = make_float2 ((float) a, (float) a); }
  __device__ __forceinline__ f32x2 (int a) { v = make_float2 ((float) a, (float) a); }
  __device__ __forceinline__ f32x2 (float a, float b) { v = make_float2 (a, b); }
};
__device__ __forceinline__ f32x2 operator- (f32x2 a) { return f32x2 (-a.v.x, -a.v.y); }
__device__ __forceinline__ f32x2 operator+ (f32x2 a, f32x2 b) { f32x2 r; r.v = __fadd2_rn (a.v, b.v); return r; }
__device__ __forceinline__ f32x2 operator- (f32x2 a, f32x2 b) { f32x2 r; r.v = __fadd2_rn (a.v, (-b).v); return r; }
__device__ __forceinline__ f32x2 operator* (f32x2 a, f32x2 b) { f32x2 r; r.v = __fmul2_rn (a.v, b.v); return r; }
__device__ __forceinline__ f32x2 &operator+= (f32x2 &a, f32x2 b) { a = a + b; return a; }
__device__ __forceinline__ f32x2 &operator-= (f32x2 &a, f32x2 b) { a = a - b; return a; }
__device__ __forceinline__ f32x2 fma_r (f32x2 a, f32x2 b, f32x2 c) { f32x2 r; r.v = __ffma2_rn (a.v, b.v, c.v); return r; }

template <typename T>
__device__ __forceinline__ T
v6_ldg (const T *p)
{
#ifndef PF_EMULATION
  return __ldg (p);
#else
  return *p;
#endif
}

// ---- how a thread sees the staging arrays: one cell (scalar R) or a pair of x-adjacent cells (f32x2) ---------------
template <typename R> struct Lane
{
  using S = R;                       // element type of the staging arrays in shared memory
  static constexpr int W = 1;        // cells per thread
  static __device__ __forceinline__ R ld (const S *p) { return p[0]; }        // node column c of the thread
  static __device__ __forceinline__ R ld1 (const S *p, R) { return p[1]; }    // node column c + 1
  static __device__ __forceinline__ void rec (const S *r, R &wg, R &c2) { wg = r[0], c2 = r[1]; }
  static __device__ __forceinline__ void rec_g (const S *r, R &wg, R &c2) // the same record straight from global memory
  {
    const typename Pair<S>::type q = v6_ldg (reinterpret_cast<const typename Pair<S>::type *> (r));
    wg = q.x, c2 = q.y;
  }
  // y tile: phase A adds the contributions to the thread's first node column(s), phase B to its last one
  static __device__ __forceinline__ void add_a (S *p, R v0, R) { p[0] += v0; }
  static __device__ __forceinline__ void add_b (S *p, R v1) { p[1] += v1; }
  static __device__ __forceinline__ R drop_right (R v) { return v; }
};
template <> struct Lane<f32x2>
{
  using S = float;
  static constexpr int W = 2;
  static __device__ __forceinline__ f32x2 ld (const float *p)
  {
    f32x2 r;
    r.v = *reinterpret_cast<const float2 *> (p); // columns (c, c + 1): c is even, the row pitch too
    return r;
  }
  static __device__ __forceinline__ f32x2 ld1 (const float *p, f32x2 a) { return f32x2 (a.v.y, p[2]); } // columns (c + 1, c + 2)
  static __device__ __forceinline__ void rec (const float *r, f32x2 &wg, f32x2 &c2)
  {
    const float4 q = *reinterpret_cast<const float4 *> (r); // (wg left, wg right, c2 left, c2 right)
    wg = f32x2 (q.x, q.y), c2 = f32x2 (q.z, q.w);
  }
  static __device__ __forceinline__ void rec_g (const float *r, f32x2 &wg, f32x2 &c2)
  {
    const float4 q = v6_ldg (reinterpret_cast<const float4 *> (r));
    wg = f32x2 (q.x, q.y), c2 = f32x2 (q.z, q.w);
  }
  // the middle column c + 1 belongs to both cells of the pair: summed in the thread
  static __device__ __forceinline__ void add_a (float *p, f32x2 v0, f32x2 v1)
  {
    float2 t = *reinterpret_cast<float2 *> (p);
    t.x += v0.v.x;
    t.y += v1.v.x + v0.v.y;
    *reinterpret_cast<float2 *> (p) = t;
  }
  static __device__ __forceinline__ void add_b (float *p, f32x2 v1) { p[2] += v1.v.y; }
  static __device__ __forceinline__ f32x2 drop_right (f32x2 v) { return f32x2 (v.v.x, 0.0f); }
};

// ---- the coefficient stream: TMA bulk copies (cp.async.bulk, completion on an mbarrier) ---------------------------
// One contiguous block per (tile, Gauss plane), 9 (or 4) records per cell: thread 0 of the CTA issues the copy, the
// copy engine fills shared memory while the CTA stages x and U, no register or issue slot of the FP64-bound cell walk
// is spent on it.  PF_EMULATION (tests/emu shims): the same data movement as a cooperative copy.
#ifndef PF_EMULATION
__device__ __forceinline__ unsigned
v6_smem_u32 (const void *p)
{
  return (unsigned) __cvta_generic_to_shared (p);
}
__device__ __forceinline__ void
v6_mbar_init (void *mbar)
{
  asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(v6_smem_u32 (mbar)) : "memory");
}
__device__ __forceinline__ void
v6_mbar_init_fence ()
{
  asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void
v6_bulk_load (void *dst, const void *src, unsigned bytes, void *mbar)
{
  asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(v6_smem_u32 (mbar)), "r"(bytes) : "memory");
  asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                  v6_smem_u32 (dst)),
                "l"(reinterpret_cast<unsigned long long> (src)), "r"(bytes), "r"(v6_smem_u32 (mbar))
                : "memory");
}
__device__ __forceinline__ void
v6_mbar_wait (void *mbar, unsigned parity)
{
  asm volatile ("{\n"
                ".reg .pred p;\n"
                "WAIT_%=:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra DONE_%=;\n"
                "bra WAIT_%=;\n"
                "DONE_%=:\n"
                "}" ::"r"(v6_smem_u32 (mbar)),
                "r"(parity)
                : "memory");
}
__device__ __forceinline__ void
v6_async_proxy_fence ()
{
  asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
}
#endif

// TX x TY cells per tile, W cells per thread; the row pitch of the staging arrays is even when W == 2
// COUPLED = false: the operator without its (phi,u) block, i.e. block diagonal like the reference's own preconditioner
// (BlockDiagonalPreconditioner, cracks.cc:2717-2740) -- multigrid smoother only; the state U is then not staged at all
template <int TX, int TY, int NQ = 3, int W = 1, bool COUPLED = true> struct Tile3v6
{
  static_assert (TX % W == 0, "a thread's cells lie in one tile row");
  static constexpr int NX = TX + 1, NY = TY + 1;
  static constexpr int PX = (W == 2 && NX % 2) ? NX + 1 : NX; // row pitch of the node arrays
  static constexpr int NN = PX * NY * 2; // nodes of the tile (one cell layer), padded rows
  static constexpr int NC2 = PX * NY;    // node columns
  static constexpr int NXC = PX * TY;    // (x-node, cell row)
  static constexpr int NT = TX / W * TY; // threads
  static constexpr int SY = PX, SZ = PX * NY;
  static constexpr int NF = COUPLED ? 8 : 4;  // staged nodal fields: x_u (3), x_phi / 8, u (3), phi / 8
  static constexpr int NFZ = COUPLED ? 7 : 4; // fields with a z-difference chain: all but phi
  static constexpr int NQP = NQ * NQ * NQ;
  static constexpr size_t scratch = (NFZ * NC2 > 4 * NN) ? (size_t) NFZ * NC2 : (size_t) 4 * NN; // DZ, then the y tile
  static constexpr size_t smem_elems = (size_t) NQ * NF * NC2 + (size_t) NFZ * NQ * NXC + NXC + scratch;
  // coefficient scalars (wg and c2 of every cell) of one tile, and of one Gauss plane of it (one bulk copy)
  static constexpr size_t coef_per_tile = (size_t) NQP * TX * TY * 2;
  static constexpr size_t coef_per_plane = (size_t) NQ * NQ * TX * TY * 2;
  // shared memory of the kernel by coefficient feed (CFM, see k_apply3d_v6): staging arrays, record ring, mbarriers
  static constexpr size_t ring_planes (int cfm) { return cfm == 0 ? 2 : (cfm == 2 ? 1 : 0); }
  template <typename S> static constexpr size_t smem_bytes (int cfm)
  {
    return (smem_elems * sizeof (S) + 15) / 16 * 16 + ring_planes (cfm) * coef_per_plane * sizeof (S) + (cfm == 1 ? 0 : 32);
  }
};

// constants of one launch, derived on the host from Phys / K3 (cubic cells: h = hx = hy = hz)
struct K6
{
  double s;      // sqrt(3/5)
  double gam;    // 1 / (4 h): a difference of nodal sums times gam is a gradient
  double lam2;   // lambda / (2 mu)
  double beta;   // (alpha - 1) p gam / ((1 - kappa) 2 mu gam^2): pressure term of the (phi,u) block in units of the stress term
  double k1;     // (1 - kappa) mu gam^2 / 2: scale of the (phi,u) stress term (JxW and the 1/8 of the test function not included)
  double w[3];   // JxW of the three point classes of a Gauss plane, per unit plane weight: (5/9)^2, (5/9)(8/9), (8/9)^2
  double wz[3];  // h^3/8 times the Gauss weight of the plane
  double kl[3];  // closed-form G_c eps Laplacian: G_c eps h {1/2, 1/3, 1/6}
  double s2;     // sqrt(1/3): abscissa of the 2-point rule (multigrid smoother operator, preconditioner only)
  double wvol;   // h^3 / 8: JxW of every point of the 2-point rule
  double cge;    // G_c eps 8 gam^2: phi-gradient flux of the 2-point rule (the 3-point rule uses the closed form)
};

#ifndef PF_POINT_COEFFS_THREADS
#define PF_POINT_COEFFS_THREADS 256 // resident threads per SM the set-up kernel is built for: no register cap (0.45 ms;
                                    // 384, i.e. 168 registers and 200 B of spills: 0.54 ms at 16.7 M DoF)
#endif
// ---- set-up: the two state coefficients per quadrature point ---------------------------------------------
// One thread per cell; record order exactly as the apply kernel reads it: per (tile, point, thread of the apply
// kernel) the W values of wg, then the W values of c2.  Everything is computed in FP64 and stored as CS.
template <typename CS, int TX, int TY, int NQ, int W>
__global__ void __launch_bounds__ (TX * TY, PF_POINT_COEFFS_THREADS / (TX * TY))
k_point_coeffs (Grid g, Phys p, K3 k, int tiles_x, int tiles_y, int layer0, const double *__restrict__ sol,
                const double *__restrict__ pt, CS *__restrict__ coef)
{
  using T = Tile3v6<TX, TY, NQ, W>;
  const int tid = threadIdx.x, lx = tid % TX, ly = tid / TX;
  int b = blockIdx.x;
  const int bx = b % tiles_x;
  b /= tiles_x;
  const int by = b % tiles_y, bz = b / tiles_y;
  const int cx = bx * TX + lx, cy = by * TY + ly, cz = g.cell_begin + bz;
  const int thread = lx / W + (TX / W) * ly, lane = lx % W; // thread of the apply kernel that owns this cell
  CS *out = coef + ((size_t) ((cz - layer0) * tiles_y + by) * tiles_x + bx) * T::coef_per_tile + (size_t) thread * 2 * W + lane;
  const bool valid = cx < g.n[0] && cy < g.n[1] && cz < g.cell_end;
  double u[8][3], pe[8];
#pragma unroll
  for (int v = 0; v < 8; ++v)
    {
      u[v][0] = u[v][1] = u[v][2] = pe[v] = 0;
      if (valid)
        {
          const long long n = (cx + (v & 1)) + (long long) g.nn[0] * (cy + ((v >> 1) & 1))
                              + g.nodes_per_plane * (cz + (v >> 2) - g.plane_begin);
          const double4 s = *reinterpret_cast<const double4 *> (sol + 4 * n);
          u[v][0] = s.x, u[v][1] = s.y, u[v][2] = s.z, pe[v] = pt[n];
        }
    }
  const double gam = k.gu[0];
  const double two_mu_g2 = 2.0 * p.mu * gam * gam, omk = 1.0 - p.kappa, g2 = gam * gam;
  const double es[3] = {NQ == 3 ? -k.s : -k.s2, NQ == 3 ? 0.0 : k.s2, k.s};
  // The interpolation collapses z -> y -> x like the apply (vertex v = vx + 2 vy + 4 vz, shape function
  // prod (1 +- xi) / 8): per node column the z-sum s and z-difference r of a field; a gradient is gam times a
  // difference of such sums.  Fields 0..2 = u, 3 = phi~ (value only).
  double sz[4][4], rz[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        sz[c][j] = u[j + 4][c] + u[j][c], rz[c][j] = u[j + 4][c] - u[j][c];
      sz[3][j] = pe[j + 4] + pe[j], rz[3][j] = pe[j + 4] - pe[j];
    }
#pragma unroll 1
  for (int qz = 0; qz < NQ; ++qz)
    {
      const double ez = qz == 0 ? es[0] : (qz == 1 ? es[1] : es[2]);
      const double wzq = k.wvol * (qz == 1 ? k.wq[1] : k.wq[0]); // the rule is symmetric: wq[2] = wq[0]
      // z-derivative chain: y-collapse at the NQ abscissae, then sum / difference in x (the same in every Gauss plane;
      // recomputed per plane, 18 fewer live values than carried through the loop)
      double PB[3][NQ], RB[3][NQ];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        {
          const double p0 = rz[c][0] + rz[c][2], d0 = rz[c][2] - rz[c][0], p1 = rz[c][1] + rz[c][3], d1 = rz[c][3] - rz[c][1];
#pragma unroll
          for (int qy = 0; qy < NQ; ++qy)
            {
              const double b0 = fma (es[qy], d0, p0), b1 = fma (es[qy], d1, p1);
              PB[c][qy] = b0 + b1, RB[c][qy] = b1 - b0;
            }
        }
      // y-sums and y-differences of the plane values, per x-node; x-sum / x-difference of the y-differences
      double Py[4][2], Ry[4][2], PDy[3], RDy[3];
#pragma unroll
      for (int f = 0; f < 4; ++f)
        {
          const double a0 = fma (ez, rz[f][0], sz[f][0]), a1 = fma (ez, rz[f][1], sz[f][1]);
          const double a2 = fma (ez, rz[f][2], sz[f][2]), a3 = fma (ez, rz[f][3], sz[f][3]);
          Py[f][0] = a0 + a2, Ry[f][0] = a2 - a0, Py[f][1] = a1 + a3, Ry[f][1] = a3 - a1;
          if (f < 3)
            PDy[f] = Ry[f][0] + Ry[f][1], RDy[f] = Ry[f][1] - Ry[f][0];
        }
#pragma unroll
      for (int qy = 0; qy < NQ; ++qy)
        {
          const double ey = es[qy];
          double dxv[3], Pp, Rp;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            dxv[c] = fma (ey, Ry[c][1], Py[c][1]) - fma (ey, Ry[c][0], Py[c][0]);
          {
            const double v0 = fma (ey, Ry[3][0], Py[3][0]), v1 = fma (ey, Ry[3][1], Py[3][1]);
            Pp = v0 + v1, Rp = v1 - v0;
          }
#pragma unroll
          for (int qx = 0; qx < NQ; ++qx)
            {
              const double ex = es[qx];
              const int q = (qz * NQ + qy) * NQ + qx;
              const double w = NQ == 3 ? wzq * k.wq[qx] * k.wq[qy] : k.wvol;
              double G[3][3];
#pragma unroll
              for (int c = 0; c < 3; ++c)
                {
                  G[c][0] = dxv[c];
                  G[c][1] = fma (ex, RDy[c], PDy[c]);
                  G[c][2] = fma (ex, RB[c][qy], PB[c][qy]);
                }
              double pte = 0.125 * fma (ex, Rp, Pp);
              if (p.clamp_extra)
                pte = fmin (fmax (pte, 0.0), 1.0);
              const double gdeg = fma (omk * pte, pte, p.kappa);
              // E = gam sym(G): tr E = gam tr G, E:E = gam^2 (sum of squared diagonal + half the squared off-diagonal sums)
              const double trG = G[0][0] + G[1][1] + G[2][2];
              const double o01 = G[0][1] + G[1][0], o02 = G[0][2] + G[2][0], o12 = G[1][2] + G[2][1];
              const double dd = fma (G[0][0], G[0][0], fma (G[1][1], G[1][1], G[2][2] * G[2][2]));
              const double od = fma (o01, o01, fma (o02, o02, o12 * o12));
              const double tr = gam * trG, ee = g2 * fma (0.5, od, dd);
              const double sE = p.lambda * tr * tr + 2.0 * p.mu * ee; // sigma(u) : E(u)
              CS *rec = out + (size_t) q * T::NT * 2 * W;
              rec[0] = valid ? (CS) (w * gdeg * two_mu_g2) : (CS) 0;
              rec[W] = valid ? (CS) (0.125 * w * (omk * sE + p.G_c / p.eps - 2.0 * p.P1 * tr)) : (CS) 0;
            }
        }
    }
}

// ---- the diagonal of the Jacobian from the coefficient records (Jacobi / Chebyshev preconditioner) ---------------
// NQ = 3: records of the exact rule (fine level: the diagonal pf_jacobian_diagonal returns); NQ = 2: records of the
// smoother's 2-point operator (coarse multigrid levels, where only those exist).
// diag(J)[v, c] of a cell, in the units of the records: u rows  sum_q wg [(lam2 + 1/2) d_c^2 + |d|^2 / 2], phi row
// sum_q c2 a^2 / 8 + G_c eps h / 3 (the Q1 Laplacian in closed form), with a = prod (1 +- xi_k) and d_k = a without its
// k-th factor: separable in the three directions, so the 27 points collapse x -> y -> z like the apply itself.  Per cell
// the absolute value, and the cell's mean where an entry is zero, as k_diag_generic (pf_generic.cuh) does to mimic
// AffineConstraints::distribute_local_to_global.  One thread per cell, thread / record mapping of k_point_coeffs;
// cells that share a node within the tile are ordered by eight barrier-separated vertex phases (no shared atomics).
template <typename CS, int TX, int TY, int W, int NQ>
__global__ void __launch_bounds__ (TX * TY)
k_diag_v6 (Grid g, K6 k, int tiles_x, int tiles_y, int layer0, const CS *__restrict__ coef, double *__restrict__ diag)
{
  using T = Tile3v6<TX, TY, NQ, W>;
  constexpr int NN = T::NN, PX = T::PX, NX = T::NX, NY = T::NY, NTH = TX * TY;
  __shared__ double dt[4 * NN];
  const int tid = threadIdx.x, lx = tid % TX, ly = tid / TX;
  int b = blockIdx.x;
  const int bx = b % tiles_x;
  b /= tiles_x;
  const int by = b % tiles_y, bz = b / tiles_y;
  const int cx0 = bx * TX, cy0 = by * TY, cz = g.cell_begin + bz;
  const int cx = cx0 + lx, cy = cy0 + ly;
  const int thread = lx / W + (TX / W) * ly, lane = lx % W;
  const CS *rec = coef + ((size_t) ((cz - layer0) * tiles_y + by) * tiles_x + bx) * T::coef_per_tile + (size_t) thread * 2 * W + lane;
  const bool valid = cx < g.n[0] && cy < g.n[1] && cz < g.cell_end;
  for (int i = tid; i < 4 * NN; i += NTH)
    dt[i] = 0;
  // (1 -+ xi)^2 at the three abscissae, for the lower (0) and the upper (1) node of a direction
  const double sa = NQ == 3 ? k.s : k.s2, sm = (1.0 - sa) * (1.0 - sa), sp = (1.0 + sa) * (1.0 + sa);
  const double sq[3][2] = {{sp, sm}, {NQ == 3 ? 1.0 : sm, NQ == 3 ? 1.0 : sp}, {sm, sp}};
  double T0[2][2] = {{0, 0}, {0, 0}}, T1[2][2] = {{0, 0}, {0, 0}}, T2[2][2] = {{0, 0}, {0, 0}};
  double T3[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};
  if (valid)
    {
#pragma unroll
      for (int qz = 0; qz < NQ; ++qz)
        {
          double Y0[2] = {0, 0}, Y1[2] = {0, 0}, Y2[2][2] = {{0, 0}, {0, 0}}, Yc[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
          for (int qy = 0; qy < NQ; ++qy)
            {
              double X0 = 0, X1[2] = {0, 0}, Xc[2] = {0, 0};
#pragma unroll
              for (int qx = 0; qx < NQ; ++qx)
                {
                  const CS *r = rec + (size_t) ((qz * NQ + qy) * NQ + qx) * T::NT * 2 * W;
                  const double wg = (double) v6_ldg (r), c2 = (double) v6_ldg (r + W);
                  X0 += wg;
#pragma unroll
                  for (int b0 = 0; b0 < 2; ++b0)
                    {
                      X1[b0] = fma (wg, sq[qx][b0], X1[b0]);
                      Xc[b0] = fma (c2, sq[qx][b0], Xc[b0]);
                    }
                }
#pragma unroll
              for (int b1 = 0; b1 < 2; ++b1)
                {
                  Y0[b1] = fma (X0, sq[qy][b1], Y0[b1]);
                  Y1[b1] += X1[b1];
#pragma unroll
                  for (int b0 = 0; b0 < 2; ++b0)
                    {
                      Y2[b0][b1] = fma (X1[b0], sq[qy][b1], Y2[b0][b1]);
                      Yc[b0][b1] = fma (Xc[b0], sq[qy][b1], Yc[b0][b1]);
                    }
                }
            }
#pragma unroll
          for (int b2 = 0; b2 < 2; ++b2)
#pragma unroll
            for (int a = 0; a < 2; ++a)
              {
                T0[a][b2] = fma (Y0[a], sq[qz][b2], T0[a][b2]); // d_0^2 = a_1^2 a_2^2: [v1][v2]
                T1[a][b2] = fma (Y1[a], sq[qz][b2], T1[a][b2]); // d_1^2 = a_0^2 a_2^2: [v0][v2]
                T3[0][a][b2] = fma (Yc[0][a], sq[qz][b2], T3[0][a][b2]);
                T3[1][a][b2] = fma (Yc[1][a], sq[qz][b2], T3[1][a][b2]);
              }
#pragma unroll
          for (int b0 = 0; b0 < 2; ++b0)
#pragma unroll
            for (int b1 = 0; b1 < 2; ++b1)
              T2[b0][b1] += Y2[b0][b1]; // d_2^2 = a_0^2 a_1^2: [v0][v1]
        }
    }
  double out[8][4], avg = 0;
#pragma unroll
  for (int v = 0; v < 8; ++v)
    {
      const int b0 = v & 1, b1 = (v >> 1) & 1, b2 = v >> 2;
      const double d0 = T0[b1][b2], d1 = T1[b0][b2], d2 = T2[b0][b1], hs = 0.5 * (d0 + d1 + d2);
      out[v][0] = fabs (fma (k.lam2 + 0.5, d0, hs));
      out[v][1] = fabs (fma (k.lam2 + 0.5, d1, hs));
      out[v][2] = fabs (fma (k.lam2 + 0.5, d2, hs));
      out[v][3] = fabs (fma (0.125, T3[b0][b1][b2], k.kl[1]));
      avg += out[v][0] + out[v][1] + out[v][2] + out[v][3];
    }
  avg *= 1.0 / 32.0;
  __syncthreads ();
#pragma unroll
  for (int v = 0; v < 8; ++v)
    {
      if (valid)
        {
          double *d = dt + (lx + (v & 1)) + PX * (ly + ((v >> 1) & 1)) + PX * NY * (v >> 2);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            d[c * NN] += out[v][c] != 0.0 ? out[v][c] : avg;
        }
      __syncthreads ();
    }
  const int nnx = g.nn[0], nny = g.nn[1];
  for (int i = tid; i < NN; i += NTH)
    {
      const int ix = i % PX, iy = (i / PX) % NY, iz = i / (PX * NY);
      const int gx = cx0 + ix, gy = cy0 + iy, gz = cz + iz;
      if (ix < NX && gx < nnx && gy < nny && cz < g.cell_end)
        {
          const long long n = gx + (long long) nnx * gy + g.nodes_per_plane * (gz - g.plane_begin);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            atomicAdd (&diag[4 * n + c], dt[c * NN + i]);
        }
    }
}

// ---- stages 3 and 4 of the apply for one tile (contains block barriers: all threads of the CTA call it) ---
// NQ = 3: the exact rule, G_c eps grad(dphi).grad(psi) in closed form.  NQ = 2: the under-integrated operator of the
// multigrid smoother (preconditioner only), phi-gradient flux inside the quadrature.
template <typename R, int TX, int TY, int NQ, bool COUPLED, int CFM>
__device__ __forceinline__ void
tile_cells_v6 (const Grid &g, const K6 &k, const int tid, const int cx0, const int cy0, const int cz0,
               const typename Lane<R>::S *__restrict__ AZ, const typename Lane<R>::S *__restrict__ BZ,
               const typename Lane<R>::S *__restrict__ BR, const typename Lane<R>::S *__restrict__ coef_tile,
               typename Lane<R>::S *CF, unsigned long long *mbar, typename Lane<R>::S *__restrict__ ys)
{
  using L = Lane<R>;
  using S = typename L::S;
  constexpr int W = L::W;
  using T = Tile3v6<TX, TY, NQ, W, COUPLED>;
  constexpr int NN = T::NN, PX = T::PX, NC2 = T::NC2, NXC = T::NXC, NF = T::NF, NT = T::NT;
  constexpr unsigned plane_bytes = (unsigned) (T::coef_per_plane * sizeof (S));
  constexpr bool CEN = NQ == 3; // the 3-point rule has a centre point and uses the closed-form Laplacian
  constexpr int NCQ = CEN ? 3 : 4; // components whose fluxes go through the quadrature
  const R Sq = (R) (CEN ? k.s : k.s2);
  const int tx = tid % (TX / W), ty = tid / (TX / W);
  const bool valid = (cx0 + W * tx < g.n[0]) && (cy0 + ty < g.n[1]) && (cz0 < g.cell_end);
  const bool right_valid = W == 1 || (cx0 + W * tx + 1 < g.n[0]); // W == 2: the second cell of the pair exists
  const int c00 = W * tx + PX * ty;  // first node column of the thread in AZ; (x-node, cell row) in BZ alike
  const int nbase = W * tx + T::SY * ty;
  const R lam2 = (R) k.lam2, nbeta = (R) -k.beta;
  const R es[3] = {-Sq, CEN ? (R) 0 : Sq, Sq};
  const R kl1 = (R) k.kl[0], kl2 = (R) k.kl[1], kl3 = (R) k.kl[2];
  const R wb = (R) (k.wvol * k.cge); // NQ == 2: weight of the phi-gradient flux (every point has JxW = h^3/8)
  const R half = (R) 0.5, two = (R) 2;
  // CFM == 1: the records of the next row of points (qz, qy) are in flight in registers while this one is evaluated
  const S *cg = coef_tile + (size_t) tid * 2 * W;
  R wgN[NQ], c2N[NQ];
  if (CFM == 1)
    {
#pragma unroll
      for (int qx = 0; qx < NQ; ++qx)
        L::rec_g (cg + (size_t) qx * NT * 2 * W, wgN[qx], c2N[qx]);
    }

#pragma unroll 1
  for (int qz = 0; qz < NQ; ++qz)
    {
      const R ez = (qz == 0) ? -Sq : ((CEN && qz == 1) ? (R) 0 : Sq);
      const S *Aq = AZ + qz * NF * NC2 + c00;
      // the coefficient records of this plane: buffer qz % 2, filled by the bulk copy that signals mbar[qz]
      const S *cf = CF + (size_t) (CFM == 0 ? (qz & 1) : 0) * T::coef_per_plane + (size_t) tid * 2 * W;
#ifndef PF_EMULATION
      if (CFM != 1)
        v6_mbar_wait (mbar + qz, 0);
#endif
      // (phi,u) weight of the point classes of this plane: JxW k1
      const double wzk = (CEN ? k.wz[qz] : k.wvol) * k.k1;
      const R wk[3] = {(R) (CEN ? wzk * k.w[0] : wzk), (R) (wzk * k.w[1]), (R) (wzk * k.w[2])};
      R VP[4][2], VR[4][2], DP[4][2], DR[4][2], YP[4], YR[4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        {
          YP[c] = YR[c] = (R) 0;
#pragma unroll
          for (int vx = 0; vx < 2; ++vx)
            VP[c][vx] = VR[c][vx] = DP[c][vx] = DR[c][vx] = (R) 0;
        }
      if (valid)
        {
          // plane level: in-plane sums and differences of the four node columns of a cell
          R s0[NF], s1[NF], r0[NF], r1[NF];
#pragma unroll
          for (int f = 0; f < NF; ++f)
            {
              const R a00 = L::ld (Aq + f * NC2), a10 = L::ld1 (Aq + f * NC2, a00);
              const R a01 = L::ld (Aq + f * NC2 + PX), a11 = L::ld1 (Aq + f * NC2 + PX, a01);
              s0[f] = a00 + a01, s1[f] = a10 + a11, r0[f] = a01 - a00, r1[f] = a11 - a10;
            }
#pragma unroll
          for (int qy = 0; qy < NQ; ++qy)
            {
              const R ey = es[qy];
              const bool ceny = CEN && qy == 1;
              R wg3[3], c23[3];
              if (CFM == 1)
                {
                  const int row = qz * NQ + qy + 1;
#pragma unroll
                  for (int qx = 0; qx < NQ; ++qx)
                    {
                      wg3[qx] = wgN[qx], c23[qx] = c2N[qx];
                      if (row < NQ * NQ)
                        L::rec_g (cg + (size_t) (row * NQ + qx) * NT * 2 * W, wgN[qx], c2N[qx]);
                    }
                }
              else
                {
#pragma unroll
                  for (int qx = 0; qx < NQ; ++qx)
                    L::rec (cf + (size_t) (qy * NQ + qx) * NT * 2 * W, wg3[qx], c23[qx]);
                }
              // x-derivative (constant along x), y-derivative and z-derivative (linear in xi_x: P + ex R) of the
              // displacement-like fields f = 0..2 (x), 4..6 (U) and -- 2-point rule only -- of x's phi (f = 3)
              R dx[7], PxDy[7], RxDy[7], PxBz[7], RxBz[7];
#pragma unroll
              for (int f = 0; f < 7; ++f)
                {
                  if ((CEN && f == 3) || (!COUPLED && f > 3))
                    continue;
                  const R ds = s1[f] - s0[f], dr = r1[f] - r0[f];
                  dx[f] = ceny ? ds : fma_r (ey, dr, ds);
                  PxDy[f] = r0[f] + r1[f];
                  RxDy[f] = dr;
                  const S *bz = BZ + (f * NQ + qy) * NXC + c00;
                  const R z0 = L::ld (bz), z1 = L::ld1 (bz, z0);
                  PxBz[f] = z0 + z1;
                  RxBz[f] = z1 - z0;
                }
              const R b0p = ceny ? s0[3] : fma_r (ey, r0[3], s0[3]), b1p = ceny ? s1[3] : fma_r (ey, r1[3], s1[3]);
              R Ppf = (R) 0, Rpf = (R) 0;
              if (COUPLED)
                {
                  const R b0f = ceny ? s0[NF - 1] : fma_r (ey, r0[NF - 1], s0[NF - 1]);
                  const R b1f = ceny ? s1[NF - 1] : fma_r (ey, r1[NF - 1], s1[NF - 1]);
                  Ppf = b0f + b1f, Rpf = b1f - b0f;
                }
              const R Pdphi = b0p + b1p, Rdphi = b1p - b0p;
              // symmetric off-diagonal strain sums, linear in xi_x
              const R oP01 = PxDy[0] + dx[1], oP02 = PxBz[0] + dx[2], oP12 = PxBz[1] + PxDy[2], oR12 = RxBz[1] + RxDy[2];
              R uP01 = (R) 0, uP02 = (R) 0, uP12 = (R) 0, uR12 = (R) 0;
              if (COUPLED)
                uP01 = PxDy[4] + dx[5], uP02 = PxBz[4] + dx[6], uP12 = PxBz[5] + PxDy[6], uR12 = RxBz[5] + RxDy[6];
              // accumulators of the transposed x-collapse: stresses S00 S01 S02 S11 S12 S22 (sum and xi_x-weighted sum)
              R P00 = (R) 0, P01 = (R) 0, P02 = (R) 0, P11 = (R) 0, P12 = (R) 0, P22 = (R) 0;
              R R01 = (R) 0, R02 = (R) 0, R11 = (R) 0, R12 = (R) 0, R22 = (R) 0;
              R AP = (R) 0, AR = (R) 0;
              R F1R = (R) 0, F2R = (R) 0; // 2-point rule: xi_x-weighted sums of the phi-gradient (its plain sums are closed forms)
#pragma unroll
              for (int qx = 0; qx < NQ; ++qx)
                {
                  const R ex = es[qx];
                  const bool cen = CEN && qx == 1;
                  const R G00 = dx[0];
                  const R G11 = cen ? PxDy[1] : fma_r (ex, RxDy[1], PxDy[1]);
                  const R G22 = cen ? PxBz[2] : fma_r (ex, RxBz[2], PxBz[2]);
                  const R o01 = cen ? oP01 : fma_r (ex, RxDy[0], oP01);
                  const R o02 = cen ? oP02 : fma_r (ex, RxBz[0], oP02);
                  const R o12 = cen ? oP12 : fma_r (ex, oR12, oP12);
                  const R dphi = cen ? Pdphi : fma_r (ex, Rdphi, Pdphi);
                  const R trG = G00 + G11 + G22;
                  R wa = dphi * c23[qx]; // (phi,phi): dphi c2
                  if (COUPLED)
                    {
                      // (phi,u): pf w k1 [ (lam2 trU - beta) trG + U:G ]
                      const R U00 = dx[4];
                      const R U11 = cen ? PxDy[5] : fma_r (ex, RxDy[5], PxDy[5]);
                      const R U22 = cen ? PxBz[6] : fma_r (ex, RxBz[6], PxBz[6]);
                      const R u01 = cen ? uP01 : fma_r (ex, RxDy[4], uP01);
                      const R u02 = cen ? uP02 : fma_r (ex, RxBz[4], uP02);
                      const R u12 = cen ? uP12 : fma_r (ex, uR12, uP12);
                      const R pf = cen ? Ppf : fma_r (ex, Rpf, Ppf);
                      const R trU = U00 + U11 + U22;
                      const R tU = fma_r (lam2, trU, nbeta);
                      const R dd = fma_r (U00, G00, fma_r (U11, G11, U22 * G22));
                      const R od = fma_r (u01, o01, fma_r (u02, o02, u12 * o12));
                      const R spg = fma_r (tU, trG, fma_r (half, od, dd));
                      const int wc = CEN ? (qx == 1) + (qy == 1) : 0; // point class: corner-, edge-, centre-like in the plane
                      wa = fma_r (pf * wk[wc], spg, wa);
                    }
                  AP += wa;
                  if (!cen)
                    AR = fma_r (ex, wa, AR);
                  // (u,u): stress in units of 2 mu gam^2, weight wg = JxW g(phi~) 2 mu gam^2 from the coefficient record
                  const R wg = wg3[qx], wgh = half * wg;
                  const R t00 = fma_r (lam2, trG, G00), t11 = fma_r (lam2, trG, G11), t22 = fma_r (lam2, trG, G22);
                  P00 = fma_r (wg, t00, P00);
                  P11 = fma_r (wg, t11, P11);
                  P22 = fma_r (wg, t22, P22);
                  P01 = fma_r (wgh, o01, P01);
                  P02 = fma_r (wgh, o02, P02);
                  P12 = fma_r (wgh, o12, P12);
                  if (!cen)
                    {
                      const R wgx = wg * ex, wghx = wgh * ex;
                      R11 = fma_r (wgx, t11, R11);
                      R22 = fma_r (wgx, t22, R22);
                      R01 = fma_r (wghx, o01, R01);
                      R02 = fma_r (wghx, o02, R02);
                      R12 = fma_r (wghx, o12, R12);
                    }
                  if (!CEN)
                    {
                      // grad(dphi) = (dx, P + ex R, P + ex R): the xi_x-weighted sums over the two points +-S are 2 S^2 R
                      F1R = fma_r (ex, fma_r (ex, RxDy[3], PxDy[3]), F1R);
                      F2R = fma_r (ex, fma_r (ex, RxBz[3], PxBz[3]), F2R);
                    }
                }
              // fx = (S00, S01, S02), fy = (S01, S11, S12), fz = (S02, S12, S22); NQ == 2: + G_c eps grad(dphi)
              const R XS[4] = {P00, P01, P02, CEN ? (R) 0 : two * wb * dx[3]};
              const R yP[4] = {P01, P11, P12, CEN ? (R) 0 : two * wb * PxDy[3]}, yR[4] = {R01, R11, R12, wb * F1R};
              const R ZP[4] = {P02, P12, P22, CEN ? (R) 0 : two * wb * PxBz[3]}, ZR[4] = {R02, R12, R22, wb * F2R};
              if (CEN && qy == 1 && qz == 1)
                {
                  // closed-form G_c eps grad(dphi).grad(psi) (cracks.cc:2378): the Q1 Laplacian is diagonal in the
                  // sum / difference basis of the 8 cell nodes; x-inverse here, y and z by stage 4 (xi_z = 0 here)
                  const R r0y = L::ld (BR + c00), r1y = L::ld1 (BR + c00, r0y);
                  const R dxp = s1[3] - s0[3];                       // x-difference of the y-sum (xi_y = 0)
                  const R pdy = r0[3] + r1[3], rdy = r1[3] - r0[3];  // y-difference: x-sum and x-difference
                  const S *bz = BZ + (3 * NQ + 1) * NXC + c00;
                  const R z0 = L::ld (bz), z1 = L::ld1 (bz, z0);
                  const R pbz = z0 + z1, rbz = z1 - z0;
                  const R o_rpp = kl1 * dxp, o_prp = kl1 * pdy, o_ppr = kl1 * pbz;
                  const R o_rrp = kl2 * rdy, o_rpr = kl2 * rbz, o_prr = kl2 * (r0y + r1y);
                  const R o_rrr = kl3 * (r1y - r0y);
                  VP[3][0] -= o_rpp;
                  VP[3][1] += o_rpp;
                  VR[3][0] += o_prp - o_rrp;
                  VR[3][1] += o_prp + o_rrp;
                  DP[3][0] += o_ppr - o_rpr;
                  DP[3][1] += o_ppr + o_rpr;
                  DR[3][0] += o_prr - o_rrr;
                  DR[3][1] += o_prr + o_rrr;
                }
              {
                // phi row: the value terms
                const R v0 = AP - AR, v1 = AP + AR;
                VP[3][0] += v0;
                VP[3][1] += v1;
                if (!ceny)
                  {
                    VR[3][0] = fma_r (ey, v0, VR[3][0]);
                    VR[3][1] = fma_r (ey, v1, VR[3][1]);
                  }
              }
#pragma unroll
              for (int c = 0; c < NCQ; ++c)
                {
                  VP[c][0] -= XS[c];
                  VP[c][1] += XS[c];
                  const R z0 = ZP[c] - ZR[c], z1 = ZP[c] + ZR[c];
                  DP[c][0] += z0;
                  DP[c][1] += z1;
                  if (!ceny)
                    {
                      VR[c][0] = fma_r (-ey, XS[c], VR[c][0]);
                      VR[c][1] = fma_r (ey, XS[c], VR[c][1]);
                      DR[c][0] = fma_r (ey, z0, DR[c][0]);
                      DR[c][1] = fma_r (ey, z1, DR[c][1]);
                    }
                  YP[c] += yP[c];
                  YR[c] += yR[c];
                }
            }
        }
      // ---- stage 4: plane -> shared y tile.  x-neighbours are lanes of one warp: the writes to a thread's first
      // node column(s) (phase A) and to its last one (phase B, the first column of the next thread) are ordered with
      // __syncwarp; y-neighbours may sit in other warps: block barriers between the vy phases
      const R omez = (R) 1 - ez, opez = (R) 1 + ez;
#pragma unroll
      for (int vy = 0; vy < 2; ++vy)
        {
#pragma unroll
          for (int vz = 0; vz < 2; ++vz)
            {
              R val[2][4];
#pragma unroll
              for (int vx = 0; vx < 2; ++vx)
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  {
                    R a;
                    if (c >= NCQ)
                      a = (vy == 0) ? VP[c][vx] - VR[c][vx] : VP[c][vx] + VR[c][vx];
                    else
                      {
                        const R yv = vx == 0 ? YP[c] - YR[c] : YP[c] + YR[c];
                        a = (vy == 0) ? VP[c][vx] - VR[c][vx] - yv : VP[c][vx] + VR[c][vx] + yv;
                      }
                    const R d = (vy == 0) ? DP[c][vx] - DR[c][vx] : DP[c][vx] + DR[c][vx];
                    const R v = (vz == 0) ? fma_r (a, omez, -d) : fma_r (a, opez, d);
                    val[vx][c] = right_valid ? v : L::drop_right (v);
                  }
              S *yn = ys + nbase + T::SY * vy + T::SZ * vz;
              if (valid)
                {
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                    L::add_a (yn + c * NN, val[0][c], val[1][c]);
                }
              __syncwarp ();
              if (valid)
                {
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                    L::add_b (yn + c * NN, val[1][c]);
                }
              __syncwarp ();
            }
          __syncthreads ();
          if (CFM == 2 && vy == 0 && qz + 1 < NQ)
            {
              // single-slot ring: every thread has read this plane's records, the next plane can land on them
#ifndef PF_EMULATION
              if (tid == 0)
                {
                  v6_async_proxy_fence ();
                  v6_bulk_load (CF, coef_tile + (size_t) (qz + 1) * T::coef_per_plane, plane_bytes, mbar + qz + 1);
                }
#else
              for (int i = tid; i < (int) T::coef_per_plane; i += NT)
                CF[i] = coef_tile[(size_t) (qz + 1) * T::coef_per_plane + i];
#endif
            }
        }
      if (CFM == 0 && NQ == 3 && qz == 0)
        {
          // every thread has passed the barrier above: buffer 0 is free for the records of plane 2
#ifndef PF_EMULATION
          if (tid == 0)
            {
              v6_async_proxy_fence ();
              v6_bulk_load (CF, coef_tile + 2 * T::coef_per_plane, plane_bytes, mbar + 2);
            }
#else
          for (int i = tid; i < (int) T::coef_per_plane; i += NT)
            CF[i] = coef_tile[2 * T::coef_per_plane + i];
          __syncthreads ();
#endif
        }
    }
}

// the whole kernel: staging (stages 1 and 2), the cell walk, the flush of the y tile.
// V = type of the global vectors x, sol, y (FP64 for the Krylov operator, FP32 inside the FP32 V-cycle);
// R = arithmetic type of the cell walk (double, float, or f32x2 = two cells per thread in packed FP32).
// The z-collapse of stage 1 is done in V.
// CFM = how the coefficient records reach the cell walk: 0 = two-plane ring in shared memory filled by TMA bulk copies one
// Gauss plane ahead; 1 = no ring, each thread loads its own records one row of points ahead into registers (least
// shared memory: more CTAs per SM); 2 = one-plane ring, refilled by TMA as soon as every thread has read the plane.
template <typename R, typename V, int NQ, int TX, int TY, int MINB, bool COUPLED = true, int CFM = 0>
__global__ void __launch_bounds__ (TX / Lane<R>::W * TY, MINB)
k_apply3d_v6 (Grid g, K6 k, int tiles_x, int tiles_y, int layer0, const V *__restrict__ x,
              const V *__restrict__ sol, const uint8_t *__restrict__ mask,
              const typename Lane<R>::S *__restrict__ coef, V *__restrict__ y)
{
  using S = typename Lane<R>::S;
  constexpr int W = Lane<R>::W;
  using T = Tile3v6<TX, TY, NQ, W, COUPLED>;
  using V4 = typename Quad<V>::type;
  constexpr int NN = T::NN, NT = T::NT, NX = T::NX, PX = T::PX, NY = T::NY, NC2 = T::NC2, NXC = T::NXC, NF = T::NF,
                NFZ = T::NFZ;
  extern __shared__ __align__ (16) unsigned char smem_raw[];
  S *AZ = reinterpret_cast<S *> (smem_raw); // [NQ][NF][NC2]
  S *BZ = AZ + NQ * NF * NC2;               // [NFZ][NQ][NXC]
  S *BR = BZ + NFZ * NQ * NXC;              // [NXC]: y-difference of the z-difference of x's phi
  S *DZ = BR + NXC;                         // [NFZ][NC2], stage 1 -> 2 only
  S *ys = DZ;                               // [4][NN], aliases DZ
  constexpr size_t off_cf = (T::smem_elems * sizeof (S) + 15) / 16 * 16,
                   ring = (CFM == 0 ? 2 : (CFM == 2 ? 1 : 0)) * T::coef_per_plane, off_mbar = off_cf + ring * sizeof (S);
  S *CF = reinterpret_cast<S *> (smem_raw + off_cf);                                       // [2][NQ*NQ][NT][2 W]
  unsigned long long *mbar = reinterpret_cast<unsigned long long *> (smem_raw + off_mbar); // one per Gauss plane

  const int tid = threadIdx.x;
  int bx, by, bz;
  decode_tile (g, (int) blockIdx.x, tiles_x, tiles_y, bx, by, bz);
  const int cx0 = bx * TX, cy0 = by * TY, cz0 = g.cell_begin + bz * g.layer_stride;
  const S *coef_tile = coef + ((size_t) ((cz0 - layer0) * tiles_y + by) * tiles_x + bx) * T::coef_per_tile;
  // the records of the first two Gauss planes start to arrive while x and U are staged
#ifndef PF_EMULATION
  if (CFM != 1 && tid == 0)
    {
      for (int q = 0; q < NQ; ++q)
        v6_mbar_init (mbar + q);
      v6_mbar_init_fence ();
      constexpr unsigned plane_bytes = (unsigned) (T::coef_per_plane * sizeof (S));
      v6_bulk_load (CF, coef_tile, plane_bytes, mbar);
      if (CFM == 0)
        v6_bulk_load (CF + T::coef_per_plane, coef_tile + T::coef_per_plane, plane_bytes, mbar + 1);
    }
#else
  for (int i = tid; i < (int) ring; i += NT)
    CF[i] = coef_tile[i];
#endif
  const int nnx = g.nn[0], nny = g.nn[1];
  const long long pstride = g.nodes_per_plane;
  const V Sq = (V) (NQ == 3 ? k.s : k.s2);
  const V eighth = (V) 0.125;

  // ---- stage 1: z-collapse per node column, in the precision of the global vectors ----------------
  for (int i = tid; i < NC2; i += NT)
    {
      const int ix = i % PX, iy = i / PX;
      const int gx = cx0 + ix, gy = cy0 + iy;
      V f0[NF], f1[NF];
#pragma unroll
      for (int f = 0; f < NF; ++f)
        f0[f] = f1[f] = 0;
      if (ix < NX && gx < nnx && gy < nny && cz0 < g.cell_end)
        {
          const long long n0 = gx + (long long) nnx * gy + pstride * (cz0 - g.plane_begin);
          const long long n1 = n0 + pstride;
          const V4 xa = *reinterpret_cast<const V4 *> (x + 4 * n0);
          const V4 xb = *reinterpret_cast<const V4 *> (x + 4 * n1);
          const uint8_t m0 = mask[n0], m1 = mask[n1];
          f0[0] = (m0 & 1) ? (V) 0 : xa.x;
          f0[1] = (m0 & 2) ? (V) 0 : xa.y;
          f0[2] = (m0 & 4) ? (V) 0 : xa.z;
          f0[3] = (m0 & 8) ? (V) 0 : eighth * xa.w;
          f1[0] = (m1 & 1) ? (V) 0 : xb.x;
          f1[1] = (m1 & 2) ? (V) 0 : xb.y;
          f1[2] = (m1 & 4) ? (V) 0 : xb.z;
          f1[3] = (m1 & 8) ? (V) 0 : eighth * xb.w;
          if (COUPLED)
            {
              const V4 sa = *reinterpret_cast<const V4 *> (sol + 4 * n0);
              const V4 sb = *reinterpret_cast<const V4 *> (sol + 4 * n1);
              f0[NF - 4] = sa.x, f0[NF - 3] = sa.y, f0[NF - 2] = sa.z, f0[NF - 1] = eighth * sa.w;
              f1[NF - 4] = sb.x, f1[NF - 3] = sb.y, f1[NF - 2] = sb.z, f1[NF - 1] = eighth * sb.w;
            }
        }
#pragma unroll
      for (int f = 0; f < NF; ++f)
        {
          const V s = f0[f] + f1[f], r = f1[f] - f0[f];
          AZ[(0 * NF + f) * NC2 + i] = (S) fma_r (-Sq, r, s);
          AZ[(1 * NF + f) * NC2 + i] = (NQ == 3) ? (S) s : (S) fma_r (Sq, r, s);
          if (NQ == 3)
            AZ[(2 * NF + f) * NC2 + i] = (S) fma_r (Sq, r, s);
          if (f < NFZ)
            DZ[f * NC2 + i] = (S) r;
        }
    }
  __syncthreads ();
  // ---- stage 2: y-collapse of the z-difference chain ------------------------------
  for (int i = tid; i < NXC; i += NT)
    {
      const int ix = i % PX, cy = i / PX;
      const int c0 = ix + PX * cy;
#pragma unroll
      for (int f = 0; f < NFZ; ++f)
        {
          const S d0 = DZ[f * NC2 + c0], d1 = DZ[f * NC2 + c0 + PX];
          const S P = d0 + d1, Rd = d1 - d0;
          if (f == 3)
            BR[i] = Rd;
          BZ[(f * NQ + 0) * NXC + i] = fma_r ((S) -Sq, Rd, P);
          BZ[(f * NQ + 1) * NXC + i] = (NQ == 3) ? P : fma_r ((S) Sq, Rd, P);
          if (NQ == 3)
            BZ[(f * NQ + 2) * NXC + i] = fma_r ((S) Sq, Rd, P);
        }
    }
  __syncthreads ();
  for (int i = tid; i < 4 * NN; i += NT)
    ys[i] = 0;
  __syncthreads ();

  tile_cells_v6<R, TX, TY, NQ, COUPLED, CFM> (g, k, tid, cx0, cy0, cz0, AZ, BZ, BR, coef_tile, CF, mbar, ys);

  // ---- flush the y tile -----------------------------------------------------------
  for (int i = tid; i < NN; i += NT)
    {
      const int ix = i % PX, iy = (i / PX) % NY, iz = i / (PX * NY);
      const int gx = cx0 + ix, gy = cy0 + iy, gz = cz0 + iz;
      if (ix < NX && gx < nnx && gy < nny && cz0 < g.cell_end)
        {
          const long long n = gx + (long long) nnx * gy + pstride * (gz - g.plane_begin);
          const uint8_t m = mask[n];
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (!((m >> c) & 1))
              atomicAdd (&y[4 * n + c], (V) ys[c * NN + i]);
        }
    }
}

} // namespace pf
