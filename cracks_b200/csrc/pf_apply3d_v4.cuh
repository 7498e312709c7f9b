// pf_apply3d_v4.cuh -- third tuning step of the hot kernel (same mapping as
// pf_apply3d_v2.cuh: staged z-collapsed node columns, one thread per cell,
// plane -> row -> point walk, transposed collapse into a shared y tile).
// What changes is the FP64 instruction count per cell (the binding resource,
// DESIGN.md 5.1):
//   * the strain only enters through its symmetric part, so the off-diagonal
//     sums G01+G10, G02+G20, G12+G21 (and the same for the state U) are formed
//     from row-level combinations with ONE fma per point instead of two fmas
//     and an add;
//   * cubic cells, 3-point rule: the constant-coefficient term
//     G_c eps grad(dphi).grad(psi) (cracks.cc:2378) is a polynomial that the
//     Gauss rule integrates exactly, so it is applied in closed form: in the
//     sum/difference (Walsh-Hadamard) basis of the 8 cell nodes the Q1
//     Laplacian is diagonal (1-D mass: h/4, h/12; 1-D stiffness: 0, 1/h), the
//     8 coefficients are by-products of the centre row of the centre plane and
//     the result joins the plane accumulators there.  The phi-gradient leaves
//     the 27-point loop altogether.  Differences to the quadrature are
//     round-off (parity tests: 1e-12).
#pragma once
#include "pf_apply3d_v2.cuh"

namespace pf {

template <int TX, int TY, int TZ> struct Tile3v4 : Tile3v2<TX, TY, TZ>
{
  using B = Tile3v2<TX, TY, TZ>;
  // + BR[NXC]: y-difference of the z-difference of the phi column of x (closed-form Laplacian)
  static constexpr size_t smem_doubles = B::smem_doubles + B::NXC;
  static constexpr size_t smem_bytes = smem_doubles * sizeof (double);
};

// Stages 3 and 4 for one tile: every thread walks its cell through the 27 Gauss
// points (plane -> row -> point) reading the staged arrays AZ / BZ, and adds
// its nodal contributions to the shared y tile `ys`.  Contains block barriers:
// must be called by all threads of the CTA.
template <int TX, int TY, int TZ, int NQ = 3, bool ISO = false>
__device__ __forceinline__ void
tile_cells_v4 (const Grid &g, const Phys &p, const K3 &k, const int tid, const int cx0, const int cy0,
               const int cz0, const double *__restrict__ AZ, const double *__restrict__ BZ,
               const double *__restrict__ BR, double *__restrict__ ys)
{
  using T = Tile3v2<TX, TY, TZ>;
  constexpr int NN = T::NN, NX = T::NX, NY = T::NY, NC2 = T::NC2, NXC = T::NXC;
  constexpr bool CEN = (NQ == 3); // the 3-point rule has a centre point (xi = 0), the 2-point rule has not
  constexpr bool CLOSED = ISO && CEN; // closed-form G_c eps Laplacian (needs cubic cells and a plane with xi_z = 0)
  const double S = CEN ? k.s : k.s2;

  // ---- stage 3: one thread per cell -----------------------------------------
  const int tx = tid % TX, ty = (tid / TX) % TY, tz = tid / (TX * TY);
  const bool valid = (cx0 + tx < g.n[0]) && (cy0 + ty < g.n[1]) && (cz0 + tz < g.cell_end);
  const int c00 = tx + NX * (ty + NY * tz);
  const int it0 = tx + NX * (ty + TY * tz);
  const int nbase = tx + T::SY * ty + T::SZ * tz;

  const double omk = 1.0 - p.kappa;
  const double c_gce = p.G_c / p.eps;
  // ISO (hx == hy == hz): the 1/(4h) gradient scales of input and output sides are
  // folded into the material constants instead of being applied per row / plane
  const double gam = k.gu[0];
  const double lamq = ISO ? p.lambda * gam * gam : p.lambda;
  const double muq = ISO ? p.mu * gam * gam : p.mu;
  const double p1g = ISO ? p.P1 * gam : p.P1;
  const double vs = ISO ? 0.125 : 1.0; // 1/8 of the phi value weights, folded into a_val when ISO
  const double c_gceps = ISO ? p.G_c * p.eps * 8.0 * gam * gam : p.G_c * p.eps;
  const double two_mu = 2.0 * muq;
  const double ca1 = 2.0 * omk * vs, ca2 = 2.0 * p1g * vs, ca3 = omk * vs, ca4 = c_gce * vs;
  const double es[3] = {-S, CEN ? 0.0 : S, S};
  // closed-form Laplacian: G_c eps * 8 (input pre-scaled by 1/8) * {h/16, h/24, h/48}
  const double kl1 = p.G_c * p.eps * g.h[0] * 0.5, kl2 = p.G_c * p.eps * g.h[0] * (1.0 / 3.0),
               kl3 = p.G_c * p.eps * g.h[0] * (1.0 / 6.0);

#pragma unroll 1
  for (int qz = 0; qz < NQ; ++qz)
    {
      const double ez = (qz == 0) ? -S : ((CEN && qz == 1) ? 0.0 : S);
      const double *Aq = AZ + qz * 9 * NC2 + c00;
      double VP[4][2], VR[4][2], DP[4][2], DR[4][2], YP[4], YR[4];
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
#pragma unroll
          for (int qy = 0; qy < NQ; ++qy)
            {
              const double ey = es[qy];
              double PxB[9], RxB[9], PxBz[7], RxBz[7], dx[7], PxDy[7], RxDy[7];
#pragma unroll
              for (int f = 0; f < 9; ++f)
                {
                  const double a00 = Aq[f * NC2], a10 = Aq[f * NC2 + 1];
                  const double a01 = Aq[f * NC2 + NX], a11 = Aq[f * NC2 + NX + 1];
                  const double r0 = a01 - a00, r1 = a11 - a10;
                  const double b0 = (CEN && qy == 1) ? a00 + a01 : fma (ey, r0, a00 + a01);
                  const double b1 = (CEN && qy == 1) ? a10 + a11 : fma (ey, r1, a10 + a11);
                  PxB[f] = b0 + b1;
                  RxB[f] = b1 - b0;
                  if (f < 7)
                    {
                      const double gys = (f == 3) ? k.gp[1] : k.gu[1];
                      dx[f] = ISO ? RxB[f] : RxB[f] * ((f == 3) ? k.gp[0] : k.gu[0]);
                      PxDy[f] = ISO ? r0 + r1 : (r0 + r1) * gys;
                      RxDy[f] = ISO ? r1 - r0 : (r1 - r0) * gys;
                      const double z0 = BZ[(f * NQ + qy) * NXC + it0], z1 = BZ[(f * NQ + qy) * NXC + it0 + 1];
                      PxBz[f] = z0 + z1;
                      RxBz[f] = z1 - z0;
                    }
                }
              // symmetric off-diagonal strain sums, linear in xi_x: P + ex R
              const double oP01 = PxDy[0] + dx[1], oP02 = PxBz[0] + dx[2], oP12 = PxBz[1] + PxDy[2],
                           oR12 = RxBz[1] + RxDy[2];
              const double uP01 = PxDy[4] + dx[5], uP02 = PxBz[4] + dx[6], uP12 = PxBz[5] + PxDy[6],
                           uR12 = RxBz[5] + RxDy[6];
              double XS[4], ZP[4], ZR[4], yP[4], yR[4];
              double AP = 0, AR = 0;
#pragma unroll
              for (int c = 0; c < 4; ++c)
                XS[c] = ZP[c] = ZR[c] = yP[c] = yR[c] = 0;

#pragma unroll
              for (int qx = 0; qx < NQ; ++qx)
                {
                  const double ex = es[qx];
                  // diagonal strains and the symmetric off-diagonal sums
                  const double G00 = dx[0], U00 = dx[4];
                  const double G11 = (CEN && qx == 1) ? PxDy[1] : fma (ex, RxDy[1], PxDy[1]);
                  const double G22 = (CEN && qx == 1) ? PxBz[2] : fma (ex, RxBz[2], PxBz[2]);
                  const double U11 = (CEN && qx == 1) ? PxDy[5] : fma (ex, RxDy[5], PxDy[5]);
                  const double U22 = (CEN && qx == 1) ? PxBz[6] : fma (ex, RxBz[6], PxBz[6]);
                  const double o01 = (CEN && qx == 1) ? oP01 : fma (ex, RxDy[0], oP01);
                  const double o02 = (CEN && qx == 1) ? oP02 : fma (ex, RxBz[0], oP02);
                  const double o12 = (CEN && qx == 1) ? oP12 : fma (ex, oR12, oP12);
                  const double u01 = (CEN && qx == 1) ? uP01 : fma (ex, RxDy[4], uP01);
                  const double u02 = (CEN && qx == 1) ? uP02 : fma (ex, RxBz[4], uP02);
                  const double u12 = (CEN && qx == 1) ? uP12 : fma (ex, uR12, uP12);
                  double gph[3] = {0, 0, 0};
                  if (!CLOSED)
                    {
                      gph[0] = dx[3];
                      gph[1] = (CEN && qx == 1) ? PxDy[3] : fma (ex, RxDy[3], PxDy[3]);
                      gph[2] = (CEN && qx == 1) ? PxBz[3] : fma (ex, RxBz[3], PxBz[3]);
                    }
                  const double dphi = (CEN && qx == 1) ? PxB[3] : fma (ex, RxB[3], PxB[3]);
                  const double pf = (CEN && qx == 1) ? PxB[7] : fma (ex, RxB[7], PxB[7]);
                  double pte = (CEN && qx == 1) ? PxB[8] : fma (ex, RxB[8], PxB[8]);
                  if (p.clamp_extra)
                    pte = fmin (fmax (pte, 0.0), 1.0);

                  const double gdeg = fma (omk * pte, pte, p.kappa);
                  const double trU = U00 + U11 + U22;
                  const double trG = G00 + G11 + G22;
                  const double ddot = fma (U00, G00, fma (U11, G11, U22 * G22));
                  const double odot = fma (u01, o01, fma (u02, o02, u12 * o12));
                  const double spG = fma (lamq * trU, trG, two_mu * fma (0.5, odot, ddot));
                  const double dd2 = fma (U00, U00, fma (U11, U11, U22 * U22));
                  const double od2 = fma (u01, u01, fma (u02, u02, u12 * u12));
                  const double spE = fma (lamq * trU, trU, two_mu * fma (0.5, od2, dd2));
                  const double a_val = pf * (ca1 * spG - ca2 * trG) + dphi * (fma (ca3, spE, ca4) - ca2 * trU);
                  const double w = CEN ? k.wvol * k.wq[qx] * k.wq[qy] * k.wq[qz] : k.wvol;
                  const double wg = w * gdeg;
                  const double wgl = wg * lamq * trG, wgm = wg * muq, wg2m = wg * two_mu;
                  const double S00 = fma (wg2m, G00, wgl), S11 = fma (wg2m, G11, wgl), S22 = fma (wg2m, G22, wgl);
                  const double S01 = wgm * o01, S02 = wgm * o02, S12 = wgm * o12;
                  const double wa = w * a_val, wb = CLOSED ? 0.0 : w * c_gceps;
                  const double fx[4] = {S00, S01, S02, wb * gph[0]};
                  const double fy[4] = {S01, S11, S12, wb * gph[1]};
                  const double fz[4] = {S02, S12, S22, wb * gph[2]};
#pragma unroll
                  for (int c = 0; c < (CLOSED ? 3 : 4); ++c)
                    {
                      XS[c] += fx[c];
                      yP[c] += fy[c];
                      ZP[c] += fz[c];
                      if (!(CEN && qx == 1))
                        {
                          yR[c] = fma (ex, fy[c], yR[c]);
                          ZR[c] = fma (ex, fz[c], ZR[c]);
                        }
                    }
                  AP += wa;
                  if (!(CEN && qx == 1))
                    AR = fma (ex, wa, AR);
                }
              if (CLOSED && qy == 1 && qz == 1)
                {
                  // Walsh-Hadamard coefficients c_{xyz} of the (1/8-scaled) phi column of x on this
                  // cell: p = sum, r = difference per direction.  The Q1 Laplacian is diagonal in
                  // this basis; the x-inverse is done here, y and z by stage 4 (xi_z = 0 on this plane).
                  const double r0y = BR[it0], r1y = BR[it0 + 1];
                  const double o_rpp = kl1 * dx[3], o_prp = kl1 * PxDy[3], o_ppr = kl1 * PxBz[3];
                  const double o_rrp = kl2 * RxDy[3], o_rpr = kl2 * RxBz[3], o_prr = kl2 * (r0y + r1y);
                  const double o_rrr = kl3 * (r1y - r0y);
                  VP[3][0] -= o_rpp;
                  VP[3][1] += o_rpp;
                  VR[3][0] += o_prp - o_rrp;
                  VR[3][1] += o_prp + o_rrp;
                  DP[3][0] += o_ppr - o_rpr;
                  DP[3][1] += o_ppr + o_rpr;
                  DR[3][0] += o_prr - o_rrr;
                  DR[3][1] += o_prr + o_rrr;
                }
#pragma unroll
              for (int c = 0; c < 4; ++c)
                {
                  if (CLOSED && c == 3)
                    {
                      // only the value terms are left in the quadrature
                      const double v0 = AP - AR, v1 = AP + AR;
                      VP[3][0] += v0;
                      VP[3][1] += v1;
                      if (!(CEN && qy == 1))
                        {
                          VR[3][0] = fma (ey, v0, VR[3][0]);
                          VR[3][1] = fma (ey, v1, VR[3][1]);
                        }
                      continue;
                    }
                  const double gxs = (c == 3) ? k.gp[0] : k.gu[0];
                  const double xv = ISO ? XS[c] : XS[c] * gxs;
                  double v0 = -xv, v1 = xv;
                  if (c == 3)
                    {
                      v0 += AP - AR;
                      v1 += AP + AR;
                    }
                  VP[c][0] += v0;
                  VP[c][1] += v1;
                  const double z0 = ZP[c] - ZR[c], z1 = ZP[c] + ZR[c];
                  DP[c][0] += z0;
                  DP[c][1] += z1;
                  if (!(CEN && qy == 1))
                    {
                      VR[c][0] = fma (ey, v0, VR[c][0]);
                      VR[c][1] = fma (ey, v1, VR[c][1]);
                      DR[c][0] = fma (ey, z0, DR[c][0]);
                      DR[c][1] = fma (ey, z1, DR[c][1]);
                    }
                  YP[c] += yP[c];
                  YR[c] += yR[c];
                }
            }
        }
      // ---- stage 4: plane -> shared y tile.  Two cells hit the same node only
      // if they are neighbours: x-neighbours are lanes of one warp (TX == 16 or
      // 32), so the two vx phases are ordered with __syncwarp; y- and z-
      // neighbours may sit in other warps, so vy (and vz when TZ > 1) phases
      // are separated by block barriers.
#pragma unroll
      for (int vy = 0; vy < 2; ++vy)
        {
#pragma unroll
          for (int vz = 0; vz < 2; ++vz)
            {
#pragma unroll
              for (int vx = 0; vx < 2; ++vx)
                {
                  double val[4];
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                    {
                      const double gys = (c == 3) ? k.gp[1] : k.gu[1];
                      const double gzs = (c == 3) ? k.gp[2] : k.gu[2];
                      const double yv0 = vx == 0 ? YP[c] - YR[c] : YP[c] + YR[c];
                      const double yv = ISO ? yv0 : yv0 * gys;
                      const double a = (CLOSED && c == 3)
                                         ? ((vy == 0) ? VP[c][vx] - VR[c][vx] : VP[c][vx] + VR[c][vx])
                                         : ((vy == 0) ? VP[c][vx] - VR[c][vx] - yv : VP[c][vx] + VR[c][vx] + yv);
                      const double d0 = (vy == 0) ? DP[c][vx] - DR[c][vx] : DP[c][vx] + DR[c][vx];
                      const double d = ISO ? d0 : d0 * gzs;
                      const double v = (vz == 0) ? fma (-ez, a, a) - d : fma (ez, a, a) + d;
                      val[c] = (c == 3 && !ISO) ? 0.125 * v : v;
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
              if (TZ > 1)
                __syncthreads ();
            }
          if (TZ == 1)
            __syncthreads ();
        }
    }

}

// the whole kernel: staging (stages 1 and 2), the cell walk, the flush of the y tile
template <int TX, int TY, int TZ, int NQ, bool ISO>
__device__ __forceinline__ void
apply3d_v4_body (const Grid &g, const Phys &p, const K3 &k, int tiles_x, int tiles_y,
                 const double *__restrict__ x, const double *__restrict__ sol,
                 const double *__restrict__ pt, const uint8_t *__restrict__ mask,
                 double *__restrict__ y)
{
  using T = Tile3v2<TX, TY, TZ>;
  constexpr int NN = T::NN, NT = T::NT, NX = T::NX, NY = T::NY, NC2 = T::NC2, NXC = T::NXC;
  extern __shared__ __align__ (16) unsigned char smem_raw[];
  double *AZ = reinterpret_cast<double *> (smem_raw); // [3][9][NC2]
  double *BZ = AZ + 27 * NC2;                         // [7][3][NXC]
  double *BR = BZ + 21 * NXC;                         // [NXC]: y-difference of the z-difference of x's phi
  double *DZ = BR + NXC;                              // [7][NC2], stage 1 -> 2 only
  double *ys = DZ;                                    // [4][NN], aliases DZ

  const int tid = threadIdx.x;
  int bx, by, bz;
  decode_tile (g, (int) blockIdx.x, tiles_x, tiles_y, bx, by, bz);
  const int cx0 = bx * TX, cy0 = by * TY, cz0 = g.cell_begin + bz * TZ * g.layer_stride;
  const int nnx = g.nn[0], nny = g.nn[1];
  const int lz_off = g.plane_begin;
  const long long pstride = g.nodes_per_plane;
  const double S = (NQ == 3) ? k.s : k.s2;

  // ---- stage 1: z-collapse per node column --------------------------------
  for (int i = tid; i < NC2; i += NT)
    {
      const int ix = i % NX, iy = (i / NX) % NY, tz = i / (NX * NY);
      const int gx = cx0 + ix, gy = cy0 + iy, gz = cz0 + tz;
      double f0[9], f1[9];
#pragma unroll
      for (int f = 0; f < 9; ++f)
        f0[f] = f1[f] = 0;
      if (gx < nnx && gy < nny && gz < g.cell_end)
        {
          const long long n0 = gx + (long long) nnx * gy + pstride * (gz - lz_off);
          const long long n1 = n0 + pstride;
          const double4 xa = *reinterpret_cast<const double4 *> (x + 4 * n0);
          const double4 xb = *reinterpret_cast<const double4 *> (x + 4 * n1);
          const double4 sa = *reinterpret_cast<const double4 *> (sol + 4 * n0);
          const double4 sb = *reinterpret_cast<const double4 *> (sol + 4 * n1);
          const uint8_t m0 = mask[n0], m1 = mask[n1];
          f0[0] = (m0 & 1) ? 0.0 : xa.x;
          f0[1] = (m0 & 2) ? 0.0 : xa.y;
          f0[2] = (m0 & 4) ? 0.0 : xa.z;
          f0[3] = (m0 & 8) ? 0.0 : 0.125 * xa.w;
          f1[0] = (m1 & 1) ? 0.0 : xb.x;
          f1[1] = (m1 & 2) ? 0.0 : xb.y;
          f1[2] = (m1 & 4) ? 0.0 : xb.z;
          f1[3] = (m1 & 8) ? 0.0 : 0.125 * xb.w;
          f0[4] = sa.x, f0[5] = sa.y, f0[6] = sa.z, f0[7] = 0.125 * sa.w, f0[8] = 0.125 * pt[n0];
          f1[4] = sb.x, f1[5] = sb.y, f1[6] = sb.z, f1[7] = 0.125 * sb.w, f1[8] = 0.125 * pt[n1];
        }
#pragma unroll
      for (int f = 0; f < 9; ++f)
        {
          const double s = f0[f] + f1[f], r = f1[f] - f0[f];
          AZ[(0 * 9 + f) * NC2 + i] = fma (-S, r, s);
          AZ[(1 * 9 + f) * NC2 + i] = (NQ == 3) ? s : fma (S, r, s);
          if (NQ == 3)
            AZ[(2 * 9 + f) * NC2 + i] = fma (S, r, s);
          if (f < 7)
            DZ[f * NC2 + i] = ISO ? r : r * ((f == 3) ? k.gp[2] : k.gu[2]);
        }
    }
  __syncthreads ();
  // ---- stage 2: y-collapse of the z-derivative chain ------------------------
  for (int i = tid; i < NXC; i += NT)
    {
      const int ix = i % NX, cy = (i / NX) % TY, tz = i / (NX * TY);
      const int c0 = ix + NX * (cy + NY * tz);
#pragma unroll
      for (int f = 0; f < 7; ++f)
        {
          const double d0 = DZ[f * NC2 + c0], d1 = DZ[f * NC2 + c0 + NX];
          const double P = d0 + d1, R = d1 - d0;
          if (f == 3)
            BR[i] = R;
          BZ[(f * NQ + 0) * NXC + i] = fma (-S, R, P);
          BZ[(f * NQ + 1) * NXC + i] = (NQ == 3) ? P : fma (S, R, P);
          if (NQ == 3)
            BZ[(f * NQ + 2) * NXC + i] = fma (S, R, P);
        }
    }
  __syncthreads ();
  for (int i = tid; i < 4 * NN; i += NT)
    ys[i] = 0;
  __syncthreads ();

  tile_cells_v4<TX, TY, TZ, NQ, ISO> (g, p, k, tid, cx0, cy0, cz0, AZ, BZ, BR, ys);

  // ---- flush the y tile ---------------------------------------------------------
  for (int i = tid; i < NN; i += NT)
    {
      const int ix = i % NX, iy = (i / NX) % NY, iz = i / (NX * NY);
      const int gx = cx0 + ix, gy = cy0 + iy, gz = cz0 + iz;
      if (gx < nnx && gy < nny && gz <= g.cell_end)
        {
          const long long n = gx + (long long) nnx * gy + pstride * (gz - lz_off);
          const uint8_t m = mask[n];
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (!((m >> c) & 1))
              atomicAdd (&y[4 * n + c], ys[c * NN + i]);
        }
    }
}

template <int TX, int TY, int TZ, int MINB, int NQ = 3, bool ISO = false>
__global__ void __launch_bounds__ (TX * TY * TZ, MINB)
k_apply3d_v4 (Grid g, Phys p, K3 k, int tiles_x, int tiles_y,
              const double *__restrict__ x, const double *__restrict__ sol,
              const double *__restrict__ pt, const uint8_t *__restrict__ mask,
              double *__restrict__ y)
{
  apply3d_v4_body<TX, TY, TZ, NQ, ISO> (g, p, k, tiles_x, tiles_y, x, sol, pt, mask, y);
}

// the same kernel under an explicit register cap (tuning variants: 200 registers = 5 CTAs of 64 threads
// per SM, 10 warps, where __launch_bounds__ (64, 5) makes ptxas go down to 168 and spill)
template <int TX, int TY, int TZ, int MAXR, int NQ = 3, bool ISO = false>
__global__ void __launch_bounds__ (TX * TY * TZ) __maxnreg__ (MAXR)
k_apply3d_v4_maxr (Grid g, Phys p, K3 k, int tiles_x, int tiles_y,
                   const double *__restrict__ x, const double *__restrict__ sol,
                   const double *__restrict__ pt, const uint8_t *__restrict__ mask,
                   double *__restrict__ y)
{
  apply3d_v4_body<TX, TY, TZ, NQ, ISO> (g, p, k, tiles_x, tiles_y, x, sol, pt, mask, y);
}

} // namespace pf
