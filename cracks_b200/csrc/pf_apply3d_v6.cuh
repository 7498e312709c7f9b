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

template <int TX, int TY, int NQ = 3> struct Tile3v6
{
  static constexpr int NX = TX + 1, NY = TY + 1;
  static constexpr int NN = NX * NY * 2; // nodes of the tile (one cell layer)
  static constexpr int NC2 = NX * NY;    // node columns
  static constexpr int NXC = NX * TY;    // (x-node, cell row)
  static constexpr int NT = TX * TY;
  static constexpr int SY = NX, SZ = NX * NY;
  static constexpr int NF = 8;  // staged nodal fields: x_u (3), x_phi / 8, u (3), phi / 8
  static constexpr int NFZ = 7; // fields with a z-difference chain: all but phi
  static constexpr int NQP = NQ * NQ * NQ;
  static constexpr size_t scratch = (NFZ * NC2 > 4 * NN) ? (size_t) NFZ * NC2 : (size_t) 4 * NN; // DZ, then the y tile
  static constexpr size_t smem_elems = (size_t) NQ * NF * NC2 + (size_t) NFZ * NQ * NXC + NXC + scratch;
  // coefficient records of one tile, and of one Gauss plane of it (one bulk copy)
  static constexpr size_t coef_per_tile = (size_t) NQP * NT;
  static constexpr size_t coef_per_plane = (size_t) NQ * NQ * NT;
  // staging arrays (rounded up to 16 bytes) + two plane buffers of coefficient records + 4 mbarriers
  template <typename R> static constexpr size_t off_cf () { return (smem_elems * sizeof (R) + 15) / 16 * 16; }
  template <typename R> static constexpr size_t off_mbar () { return off_cf<R> () + 2 * coef_per_plane * 2 * sizeof (R); }
  template <typename R> static constexpr size_t smem_bytes () { return off_mbar<R> () + 4 * sizeof (unsigned long long); }
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

// ---- set-up: the two state coefficients per quadrature point ---------------------------------------------
// One thread per cell, tiles and record order exactly as the apply kernel reads them.  Everything in FP64.
template <typename R, int TX, int TY, int NQ = 3>
__global__ void __launch_bounds__ (TX * TY)
k_point_coeffs (Grid g, Phys p, K3 k, int tiles_x, int tiles_y, int layer0, const double *__restrict__ sol,
                const double *__restrict__ pt, typename Pair<R>::type *__restrict__ coef)
{
  using T = Tile3v6<TX, TY, NQ>;
  const int tid = threadIdx.x;
  int b = blockIdx.x;
  const int bx = b % tiles_x;
  b /= tiles_x;
  const int by = b % tiles_y, bz = b / tiles_y;
  const int cx = bx * TX + tid % TX, cy = by * TY + tid / TX, cz = g.cell_begin + bz;
  typename Pair<R>::type *out = coef + ((size_t) ((cz - layer0) * tiles_y + by) * tiles_x + bx) * T::coef_per_tile + tid;
  const bool valid = cx < g.n[0] && cy < g.n[1] && cz < g.cell_end;
  double u[8][3], ph[8], pe[8];
#pragma unroll
  for (int v = 0; v < 8; ++v)
    {
      u[v][0] = u[v][1] = u[v][2] = ph[v] = pe[v] = 0;
      if (valid)
        {
          const long long n = (cx + (v & 1)) + (long long) g.nn[0] * (cy + ((v >> 1) & 1))
                              + g.nodes_per_plane * (cz + (v >> 2) - g.plane_begin);
          const double4 s = *reinterpret_cast<const double4 *> (sol + 4 * n);
          u[v][0] = s.x, u[v][1] = s.y, u[v][2] = s.z, ph[v] = s.w, pe[v] = pt[n];
        }
    }
  const double gam = k.gu[0];
  const double two_mu_g2 = 2.0 * p.mu * gam * gam, omk = 1.0 - p.kappa;
  const double es[3] = {NQ == 3 ? -k.s : -k.s2, NQ == 3 ? 0.0 : k.s2, k.s};
#pragma unroll 1
  for (int q = 0; q < NQ * NQ * NQ; ++q)
    {
      const double e[3] = {es[q % NQ], es[(q / NQ) % NQ], es[q / (NQ * NQ)]};
      const double w = NQ == 3 ? k.wvol * k.wq[q % 3] * k.wq[(q / 3) % 3] * k.wq[q / 9] : k.wvol;
      double G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, pte = 0;
#pragma unroll
      for (int v = 0; v < 8; ++v)
        {
          // shape function (1 +- e0)(1 +- e1)(1 +- e2) / 8, gradient (+-1/h)(1 +- e)(1 +- e) / 4
          const double a0 = (v & 1) ? 1.0 + e[0] : 1.0 - e[0], a1 = (v & 2) ? 1.0 + e[1] : 1.0 - e[1],
                       a2 = (v & 4) ? 1.0 + e[2] : 1.0 - e[2];
          const double d0 = ((v & 1) ? 1.0 : -1.0) * a1 * a2, d1 = ((v & 2) ? 1.0 : -1.0) * a0 * a2,
                       d2 = ((v & 4) ? 1.0 : -1.0) * a0 * a1;
          pte = fma (0.125 * a0 * a1 * a2, pe[v], pte);
#pragma unroll
          for (int c = 0; c < 3; ++c)
            {
              G[c][0] = fma (d0, u[v][c], G[c][0]);
              G[c][1] = fma (d1, u[v][c], G[c][1]);
              G[c][2] = fma (d2, u[v][c], G[c][2]);
            }
        }
      // true gradient = k.ih * G / 4 = (4 gam) G / 4 = gam G
      if (p.clamp_extra)
        pte = fmin (fmax (pte, 0.0), 1.0);
      const double gdeg = fma (omk * pte, pte, p.kappa);
      double E[3][3];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int d = 0; d < 3; ++d)
          E[c][d] = 0.5 * gam * (G[c][d] + G[d][c]);
      const double tr = E[0][0] + E[1][1] + E[2][2];
      double ee = 0;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int d = 0; d < 3; ++d)
          ee = fma (E[c][d], E[c][d], ee);
      const double sE = p.lambda * tr * tr + 2.0 * p.mu * ee; // sigma(u) : E(u)
      typename Pair<R>::type rec;
      rec.x = (R) (w * gdeg * two_mu_g2);
      rec.y = (R) (0.125 * w * (omk * sE + p.G_c / p.eps - 2.0 * p.P1 * tr));
      if (valid)
        out[(size_t) q * T::NT] = rec;
      else
        {
          rec.x = rec.y = 0;
          out[(size_t) q * T::NT] = rec;
        }
    }
}

// ---- stages 3 and 4 of the apply for one tile (contains block barriers: all threads of the CTA call it) ---
// NQ = 3: the exact rule, G_c eps grad(dphi).grad(psi) in closed form.  NQ = 2: the under-integrated operator of the
// multigrid smoother (preconditioner only), phi-gradient flux inside the quadrature.
template <typename R, int TX, int TY, int NQ>
__device__ __forceinline__ void
tile_cells_v6 (const Grid &g, const K6 &k, const int tid, const int cx0, const int cy0, const int cz0,
               const R *__restrict__ AZ, const R *__restrict__ BZ, const R *__restrict__ BR,
               const typename Pair<R>::type *__restrict__ coef_tile, typename Pair<R>::type *CF,
               unsigned long long *mbar, R *__restrict__ ys)
{
  using T = Tile3v6<TX, TY, NQ>;
  using R2 = typename Pair<R>::type;
  constexpr unsigned plane_bytes = (unsigned) (T::coef_per_plane * sizeof (R2));
  constexpr int NN = T::NN, NX = T::NX, NC2 = T::NC2, NXC = T::NXC, NF = T::NF, NT = T::NT;
  constexpr bool CEN = NQ == 3; // the 3-point rule has a centre point and uses the closed-form Laplacian
  constexpr int NCQ = CEN ? 3 : 4; // components whose fluxes go through the quadrature
  const R S = (R) (CEN ? k.s : k.s2);
  const int tx = tid % TX, ty = tid / TX;
  const bool valid = (cx0 + tx < g.n[0]) && (cy0 + ty < g.n[1]) && (cz0 < g.cell_end);
  const int c00 = tx + NX * ty;  // node column (tx, ty) of AZ
  const int it0 = tx + NX * ty;  // (x-node tx, cell row ty) of BZ (NXC = NX * TY: same linear index)
  const int nbase = tx + T::SY * ty;
  const R lam2 = (R) k.lam2, nbeta = (R) -k.beta;
  const R es[3] = {-S, CEN ? (R) 0 : S, S};
  const R kl1 = (R) k.kl[0], kl2 = (R) k.kl[1], kl3 = (R) k.kl[2];
  const R wb = (R) (k.wvol * k.cge); // NQ == 2: weight of the phi-gradient flux (every point has JxW = h^3/8)

#pragma unroll 1
  for (int qz = 0; qz < NQ; ++qz)
    {
      const R ez = (qz == 0) ? -S : ((CEN && qz == 1) ? (R) 0 : S);
      const R *Aq = AZ + qz * NF * NC2 + c00;
      // the coefficient records of this plane: buffer qz % 2, filled by the bulk copy that signals mbar[qz]
      const R2 *cf = CF + (size_t) (qz & 1) * T::coef_per_plane + tid;
#ifndef PF_EMULATION
      v6_mbar_wait (mbar + qz, 0);
#endif
      // (phi,u) weight of the point classes of this plane: JxW k1
      const R wzk = (R) ((CEN ? k.wz[qz] : k.wvol) * k.k1);
      const R wk[3] = {CEN ? wzk * (R) k.w[0] : wzk, wzk * (R) k.w[1], wzk * (R) k.w[2]};
      R VP[4][2], VR[4][2], DP[4][2], DR[4][2], YP[4], YR[4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        {
          YP[c] = YR[c] = 0;
#pragma unroll
          for (int vx = 0; vx < 2; ++vx)
            VP[c][vx] = VR[c][vx] = DP[c][vx] = DR[c][vx] = 0;
        }
      if (valid)
        {
          // plane level: in-plane sums and differences of the four node columns of the cell
          R s0[NF], s1[NF], r0[NF], r1[NF];
#pragma unroll
          for (int f = 0; f < NF; ++f)
            {
              const R a00 = Aq[f * NC2], a10 = Aq[f * NC2 + 1], a01 = Aq[f * NC2 + NX], a11 = Aq[f * NC2 + NX + 1];
              s0[f] = a00 + a01, s1[f] = a10 + a11, r0[f] = a01 - a00, r1[f] = a11 - a10;
            }
#pragma unroll
          for (int qy = 0; qy < NQ; ++qy)
            {
              const R ey = es[qy];
              const bool ceny = CEN && qy == 1;
              const R2 c0 = cf[(qy * NQ + 0) * NT], c1 = cf[(qy * NQ + 1) * NT], c2 = CEN ? cf[(qy * NQ + 2) * NT] : c1;
              // x-derivative (constant along x), y-derivative and z-derivative (linear in xi_x: P + ex R) of the
              // displacement-like fields f = 0..2 (x), 4..6 (U) and -- 2-point rule only -- of x's phi (f = 3)
              R dx[7], PxDy[7], RxDy[7], PxBz[7], RxBz[7];
#pragma unroll
              for (int f = 0; f < 7; ++f)
                {
                  if (CEN && f == 3)
                    continue;
                  const R ds = s1[f] - s0[f], dr = r1[f] - r0[f];
                  dx[f] = ceny ? ds : fma_r (ey, dr, ds);
                  PxDy[f] = r0[f] + r1[f];
                  RxDy[f] = dr;
                  const R z0 = BZ[(f * NQ + qy) * NXC + it0], z1 = BZ[(f * NQ + qy) * NXC + it0 + 1];
                  PxBz[f] = z0 + z1;
                  RxBz[f] = z1 - z0;
                }
              const R b0p = ceny ? s0[3] : fma_r (ey, r0[3], s0[3]), b1p = ceny ? s1[3] : fma_r (ey, r1[3], s1[3]);
              const R b0f = ceny ? s0[7] : fma_r (ey, r0[7], s0[7]), b1f = ceny ? s1[7] : fma_r (ey, r1[7], s1[7]);
              const R Pdphi = b0p + b1p, Rdphi = b1p - b0p, Ppf = b0f + b1f, Rpf = b1f - b0f;
              // symmetric off-diagonal strain sums, linear in xi_x
              const R oP01 = PxDy[0] + dx[1], oP02 = PxBz[0] + dx[2], oP12 = PxBz[1] + PxDy[2], oR12 = RxBz[1] + RxDy[2];
              const R uP01 = PxDy[4] + dx[5], uP02 = PxBz[4] + dx[6], uP12 = PxBz[5] + PxDy[6], uR12 = RxBz[5] + RxDy[6];
              // accumulators of the transposed x-collapse: stresses S00 S01 S02 S11 S12 S22 (sum and xi_x-weighted sum)
              R P00 = 0, P01 = 0, P02 = 0, P11 = 0, P12 = 0, P22 = 0, R01 = 0, R02 = 0, R11 = 0, R12 = 0, R22 = 0;
              R AP = 0, AR = 0;
              R F1R = 0, F2R = 0; // 2-point rule: xi_x-weighted sums of the phi-gradient (its plain sums are closed forms)
#pragma unroll
              for (int qx = 0; qx < NQ; ++qx)
                {
                  const R ex = es[qx];
                  const R2 c = qx == 0 ? c0 : (qx == 1 ? c1 : c2);
                  const bool cen = CEN && qx == 1;
                  const R G00 = dx[0], U00 = dx[4];
                  const R G11 = cen ? PxDy[1] : fma_r (ex, RxDy[1], PxDy[1]);
                  const R G22 = cen ? PxBz[2] : fma_r (ex, RxBz[2], PxBz[2]);
                  const R U11 = cen ? PxDy[5] : fma_r (ex, RxDy[5], PxDy[5]);
                  const R U22 = cen ? PxBz[6] : fma_r (ex, RxBz[6], PxBz[6]);
                  const R o01 = cen ? oP01 : fma_r (ex, RxDy[0], oP01);
                  const R o02 = cen ? oP02 : fma_r (ex, RxBz[0], oP02);
                  const R o12 = cen ? oP12 : fma_r (ex, oR12, oP12);
                  const R u01 = cen ? uP01 : fma_r (ex, RxDy[4], uP01);
                  const R u02 = cen ? uP02 : fma_r (ex, RxBz[4], uP02);
                  const R u12 = cen ? uP12 : fma_r (ex, uR12, uP12);
                  const R dphi = cen ? Pdphi : fma_r (ex, Rdphi, Pdphi);
                  const R pf = cen ? Ppf : fma_r (ex, Rpf, Ppf);
                  const R trG = G00 + G11 + G22, trU = U00 + U11 + U22;
                  // (phi,u) and (phi,phi): a = pf w k1 [ (lam2 trU - beta) trG + U:G ] + dphi c2
                  const R tU = fma_r (lam2, trU, nbeta);
                  const R dd = fma_r (U00, G00, fma_r (U11, G11, U22 * G22));
                  const R od = fma_r (u01, o01, fma_r (u02, o02, u12 * o12));
                  const R spg = fma_r (tU, trG, fma_r ((R) 0.5, od, dd));
                  const int wc = CEN ? (qx == 1) + (qy == 1) : 0; // point class: corner-, edge-, centre-like in the plane
                  const R wa = fma_r (pf * wk[wc], spg, dphi * c.y);
                  AP += wa;
                  if (!cen)
                    AR = fma_r (ex, wa, AR);
                  // (u,u): stress in units of 2 mu gam^2, weight wg = JxW g(phi~) 2 mu gam^2 from the coefficient record
                  const R wg = c.x, wgh = (R) 0.5 * c.x;
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
              const R XS[4] = {P00, P01, P02, CEN ? (R) 0 : (R) 2 * wb * dx[3]};
              const R yP[4] = {P01, P11, P12, CEN ? (R) 0 : (R) 2 * wb * PxDy[3]}, yR[4] = {R01, R11, R12, wb * F1R};
              const R ZP[4] = {P02, P12, P22, CEN ? (R) 0 : (R) 2 * wb * PxBz[3]}, ZR[4] = {R02, R12, R22, wb * F2R};
              if (CEN && qy == 1 && qz == 1)
                {
                  // closed-form G_c eps grad(dphi).grad(psi) (cracks.cc:2378): the Q1 Laplacian is diagonal in the
                  // sum / difference basis of the 8 cell nodes; x-inverse here, y and z by stage 4 (xi_z = 0 here)
                  const R r0y = BR[it0], r1y = BR[it0 + 1];
                  const R dxp = s1[3] - s0[3];                       // x-difference of the y-sum (xi_y = 0)
                  const R pdy = r0[3] + r1[3], rdy = r1[3] - r0[3];  // y-difference: x-sum and x-difference
                  const R z0 = BZ[(3 * NQ + 1) * NXC + it0], z1 = BZ[(3 * NQ + 1) * NXC + it0 + 1];
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
      // ---- stage 4: plane -> shared y tile.  x-neighbours are lanes of one warp (TX == 16 or 32): the two vx
      // phases are ordered with __syncwarp; y-neighbours may sit in other warps: block barriers between vy phases
      const R omez = (R) 1 - ez, opez = (R) 1 + ez;
#pragma unroll
      for (int vy = 0; vy < 2; ++vy)
        {
#pragma unroll
          for (int vz = 0; vz < 2; ++vz)
            {
#pragma unroll
              for (int vx = 0; vx < 2; ++vx)
                {
                  R val[4];
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
                      val[c] = (vz == 0) ? fma_r (a, omez, -d) : fma_r (a, opez, d);
                    }
                  const int n0 = nbase + vx + T::SY * vy + T::SZ * vz;
                  if (valid)
                    {
#pragma unroll
                      for (int c = 0; c < 4; ++c)
                        ys[c * NN + n0] += val[c];
                    }
                  __syncwarp ();
                }
            }
          __syncthreads ();
        }
      if (NQ == 3 && qz == 0)
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
// R = arithmetic type of the cell walk.  The z-collapse of stage 1 is done in V.
template <typename R, typename V, int NQ, int TX, int TY, int MINB>
__global__ void __launch_bounds__ (TX * TY, MINB)
k_apply3d_v6 (Grid g, K6 k, int tiles_x, int tiles_y, int layer0, const V *__restrict__ x,
              const V *__restrict__ sol, const uint8_t *__restrict__ mask,
              const typename Pair<R>::type *__restrict__ coef, V *__restrict__ y)
{
  using T = Tile3v6<TX, TY, NQ>;
  using V4 = typename Quad<V>::type;
  constexpr int NN = T::NN, NT = T::NT, NX = T::NX, NY = T::NY, NC2 = T::NC2, NXC = T::NXC, NF = T::NF, NFZ = T::NFZ;
  extern __shared__ __align__ (16) unsigned char smem_raw[];
  R *AZ = reinterpret_cast<R *> (smem_raw); // [NQ][NF][NC2]
  R *BZ = AZ + NQ * NF * NC2;               // [NFZ][NQ][NXC]
  R *BR = BZ + NFZ * NQ * NXC;              // [NXC]: y-difference of the z-difference of x's phi
  R *DZ = BR + NXC;                         // [NFZ][NC2], stage 1 -> 2 only
  R *ys = DZ;                               // [4][NN], aliases DZ
  using R2 = typename Pair<R>::type;
  constexpr size_t off_cf = (T::smem_elems * sizeof (R) + 15) / 16 * 16, off_mbar = off_cf + 2 * T::coef_per_plane * sizeof (R2);
  R2 *CF = reinterpret_cast<R2 *> (smem_raw + off_cf);                                // [2][NQ*NQ][NT]
  unsigned long long *mbar = reinterpret_cast<unsigned long long *> (smem_raw + off_mbar); // one per Gauss plane

  const int tid = threadIdx.x;
  int b = blockIdx.x;
  const int bx = b % tiles_x;
  b /= tiles_x;
  const int by = b % tiles_y;
  const int bz = b / tiles_y;
  const int cx0 = bx * TX, cy0 = by * TY, cz0 = g.cell_begin + bz * g.layer_stride;
  const R2 *coef_tile = coef + ((size_t) ((cz0 - layer0) * tiles_y + by) * tiles_x + bx) * T::coef_per_tile;
  // the records of the first two Gauss planes start to arrive while x and U are staged
#ifndef PF_EMULATION
  if (tid == 0)
    {
      for (int q = 0; q < NQ; ++q)
        v6_mbar_init (mbar + q);
      v6_mbar_init_fence ();
      constexpr unsigned plane_bytes = (unsigned) (T::coef_per_plane * sizeof (R2));
      v6_bulk_load (CF, coef_tile, plane_bytes, mbar);
      v6_bulk_load (CF + T::coef_per_plane, coef_tile + T::coef_per_plane, plane_bytes, mbar + 1);
    }
#else
  for (int i = tid; i < 2 * (int) T::coef_per_plane; i += NT)
    CF[i] = coef_tile[i];
#endif
  const int nnx = g.nn[0], nny = g.nn[1];
  const long long pstride = g.nodes_per_plane;
  const V S = (V) (NQ == 3 ? k.s : k.s2);
  const V eighth = (V) 0.125;

  // ---- stage 1: z-collapse per node column, in the precision of the global vectors ----------------
  for (int i = tid; i < NC2; i += NT)
    {
      const int ix = i % NX, iy = i / NX;
      const int gx = cx0 + ix, gy = cy0 + iy;
      V f0[NF], f1[NF];
#pragma unroll
      for (int f = 0; f < NF; ++f)
        f0[f] = f1[f] = 0;
      if (gx < nnx && gy < nny && cz0 < g.cell_end)
        {
          const long long n0 = gx + (long long) nnx * gy + pstride * (cz0 - g.plane_begin);
          const long long n1 = n0 + pstride;
          const V4 xa = *reinterpret_cast<const V4 *> (x + 4 * n0);
          const V4 xb = *reinterpret_cast<const V4 *> (x + 4 * n1);
          const V4 sa = *reinterpret_cast<const V4 *> (sol + 4 * n0);
          const V4 sb = *reinterpret_cast<const V4 *> (sol + 4 * n1);
          const uint8_t m0 = mask[n0], m1 = mask[n1];
          f0[0] = (m0 & 1) ? (V) 0 : xa.x;
          f0[1] = (m0 & 2) ? (V) 0 : xa.y;
          f0[2] = (m0 & 4) ? (V) 0 : xa.z;
          f0[3] = (m0 & 8) ? (V) 0 : eighth * xa.w;
          f1[0] = (m1 & 1) ? (V) 0 : xb.x;
          f1[1] = (m1 & 2) ? (V) 0 : xb.y;
          f1[2] = (m1 & 4) ? (V) 0 : xb.z;
          f1[3] = (m1 & 8) ? (V) 0 : eighth * xb.w;
          f0[4] = sa.x, f0[5] = sa.y, f0[6] = sa.z, f0[7] = eighth * sa.w;
          f1[4] = sb.x, f1[5] = sb.y, f1[6] = sb.z, f1[7] = eighth * sb.w;
        }
#pragma unroll
      for (int f = 0; f < NF; ++f)
        {
          const V s = f0[f] + f1[f], r = f1[f] - f0[f];
          AZ[(0 * NF + f) * NC2 + i] = (R) fma_r (-S, r, s);
          AZ[(1 * NF + f) * NC2 + i] = (NQ == 3) ? (R) s : (R) fma_r (S, r, s);
          if (NQ == 3)
            AZ[(2 * NF + f) * NC2 + i] = (R) fma_r (S, r, s);
          if (f < NFZ)
            DZ[f * NC2 + i] = (R) r;
        }
    }
  __syncthreads ();
  // ---- stage 2: y-collapse of the z-difference chain ------------------------------
  for (int i = tid; i < NXC; i += NT)
    {
      const int ix = i % NX, cy = i / NX;
      const int c0 = ix + NX * cy;
#pragma unroll
      for (int f = 0; f < NFZ; ++f)
        {
          const R d0 = DZ[f * NC2 + c0], d1 = DZ[f * NC2 + c0 + NX];
          const R P = d0 + d1, Rd = d1 - d0;
          if (f == 3)
            BR[i] = Rd;
          BZ[(f * NQ + 0) * NXC + i] = fma_r ((R) -S, Rd, P);
          BZ[(f * NQ + 1) * NXC + i] = (NQ == 3) ? P : fma_r ((R) S, Rd, P);
          if (NQ == 3)
            BZ[(f * NQ + 2) * NXC + i] = fma_r ((R) S, Rd, P);
        }
    }
  __syncthreads ();
  for (int i = tid; i < 4 * NN; i += NT)
    ys[i] = 0;
  __syncthreads ();

  tile_cells_v6<R, TX, TY, NQ> (g, k, tid, cx0, cy0, cz0, AZ, BZ, BR, coef_tile, CF, mbar, ys);

  // ---- flush the y tile -----------------------------------------------------------
  for (int i = tid; i < NN; i += NT)
    {
      const int ix = i % NX, iy = (i / NX) % NY, iz = i / (NX * NY);
      const int gx = cx0 + ix, gy = cy0 + iy, gz = cz0 + iz;
      if (gx < nnx && gy < nny && cz0 < g.cell_end)
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
