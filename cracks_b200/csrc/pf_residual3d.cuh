// pf_residual3d.cuh -- tiled 3-D residual kernel: r_total += local_rhs of
// cracks.cc:2393-2432 (sign as in the reference, rhs = -F), same mapping as
// the apply kernel (pf_apply3d_v2.cuh): z-collapsed node columns staged in
// shared memory, one thread per cell, plane -> row -> point loops, transposed
// collapse into a shared tile, red.global.add flush.  Only the 5 state fields
// are interpolated (u_x, u_y, u_z, phi, phi~); phi needs value and gradient.
//
// Called 1 + (line search trials) times per Newton step (cracks.cc:2790, 2946);
// algorithmic bytes 16 N_dof + 17 N_node.
#pragma once
#include "pf_apply3d_v2.cuh"

namespace pf {

template <int TX, int TY, int TZ> struct TileR3
{
  static constexpr int NX = TX + 1, NY = TY + 1, NZ = TZ + 1;
  static constexpr int NN = NX * NY * NZ;
  static constexpr int NC2 = NX * NY * TZ;
  static constexpr int NXC = NX * TY * TZ;
  static constexpr int NT = TX * TY * TZ;
  static constexpr int SY = NX, SZ = NX * NY;
  static constexpr size_t dz_or_y = (4 * NC2 > 4 * NN) ? 4 * NC2 : 4 * NN;
  static constexpr size_t smem_doubles = (size_t) 15 * NC2 + 12 * NXC + dz_or_y;
  static constexpr size_t smem_bytes = smem_doubles * sizeof (double);
};

template <int TX, int TY, int TZ, int MINB>
__global__ void __launch_bounds__ (TX * TY * TZ, MINB)
k_residual3d (Grid g, Phys p, K3 k, int tiles_x, int tiles_y, const double *__restrict__ sol,
              const double *__restrict__ pt, double *__restrict__ r)
{
  using T = TileR3<TX, TY, TZ>;
  constexpr int NN = T::NN, NT = T::NT, NX = T::NX, NY = T::NY, NC2 = T::NC2, NXC = T::NXC;
  constexpr int NF = 5, NG = 4; // fields: u_x u_y u_z phi/8 pt/8 ; the first NG need gradients
  extern __shared__ __align__ (16) unsigned char smem_raw[];
  double *AZ = reinterpret_cast<double *> (smem_raw); // [3][NF][NC2]
  double *BZ = AZ + 3 * NF * NC2;                     // [NG][3][NXC]
  double *DZ = BZ + 3 * NG * NXC;                     // [NG][NC2]
  double *ys = DZ;                                    // [4][NN], aliases DZ

  const int tid = threadIdx.x;
  int bx, by, bz;
  decode_tile (g, (int) blockIdx.x, tiles_x, tiles_y, bx, by, bz);
  const int cx0 = bx * TX, cy0 = by * TY, cz0 = g.cell_begin + bz * TZ;
  const int nnx = g.nn[0], nny = g.nn[1];
  const int lz_off = g.plane_begin;
  const long long pstride = g.nodes_per_plane;
  const double S = k.s;

  // ---- stage 1: z-collapse per node column
  for (int i = tid; i < NC2; i += NT)
    {
      const int ix = i % NX, iy = (i / NX) % NY, tz = i / (NX * NY);
      const int gx = cx0 + ix, gy = cy0 + iy, gz = cz0 + tz;
      double f0[NF], f1[NF];
#pragma unroll
      for (int f = 0; f < NF; ++f)
        f0[f] = f1[f] = 0;
      if (gx < nnx && gy < nny && gz < g.cell_end)
        {
          const long long n0 = gx + (long long) nnx * gy + pstride * (gz - lz_off);
          const long long n1 = n0 + pstride;
          const double4 sa = *reinterpret_cast<const double4 *> (sol + 4 * n0);
          const double4 sb = *reinterpret_cast<const double4 *> (sol + 4 * n1);
          f0[0] = sa.x, f0[1] = sa.y, f0[2] = sa.z, f0[3] = 0.125 * sa.w, f0[4] = 0.125 * pt[n0];
          f1[0] = sb.x, f1[1] = sb.y, f1[2] = sb.z, f1[3] = 0.125 * sb.w, f1[4] = 0.125 * pt[n1];
        }
#pragma unroll
      for (int f = 0; f < NF; ++f)
        {
          const double s = f0[f] + f1[f], rr = f1[f] - f0[f];
          AZ[(0 * NF + f) * NC2 + i] = fma (-S, rr, s);
          AZ[(1 * NF + f) * NC2 + i] = s;
          AZ[(2 * NF + f) * NC2 + i] = fma (S, rr, s);
          if (f < NG)
            DZ[f * NC2 + i] = rr * ((f == 3) ? k.gp[2] : k.gu[2]);
        }
    }
  __syncthreads ();
  // ---- stage 2: y-collapse of the z-derivative chain
  for (int i = tid; i < NXC; i += NT)
    {
      const int ix = i % NX, cy = (i / NX) % TY, tz = i / (NX * TY);
      const int c0 = ix + NX * (cy + NY * tz);
#pragma unroll
      for (int f = 0; f < NG; ++f)
        {
          const double d0 = DZ[f * NC2 + c0], d1 = DZ[f * NC2 + c0 + NX];
          const double P = d0 + d1, R = d1 - d0;
          BZ[(f * 3 + 0) * NXC + i] = fma (-S, R, P);
          BZ[(f * 3 + 1) * NXC + i] = P;
          BZ[(f * 3 + 2) * NXC + i] = fma (S, R, P);
        }
    }
  __syncthreads ();
  for (int i = tid; i < 4 * NN; i += NT)
    ys[i] = 0;
  __syncthreads ();

  // ---- stage 3: one thread per cell
  const int tx = tid % TX, ty = (tid / TX) % TY, tz = tid / (TX * TY);
  const bool valid = (cx0 + tx < g.n[0]) && (cy0 + ty < g.n[1]) && (cz0 + tz < g.cell_end);
  const int c00 = tx + NX * (ty + NY * tz);
  const int it0 = tx + NX * (ty + TY * tz);
  const int nbase = tx + T::SY * ty + T::SZ * tz;

  const double omk = 1.0 - p.kappa;
  const double c_gce = p.G_c / p.eps;
  const double c_gceps = p.G_c * p.eps;
  const double two_mu = 2.0 * p.mu;
  const double es[3] = {-S, 0.0, S};

#pragma unroll 1
  for (int qz = 0; qz < 3; ++qz)
    {
      const double ez = (qz == 0) ? -S : (qz == 1 ? 0.0 : S);
      const double *Aq = AZ + qz * NF * NC2 + c00;
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
          for (int qy = 0; qy < 3; ++qy)
            {
              const double ey = es[qy];
              double PxB[NF], RxB[NF], PxBz[NG], RxBz[NG], dx[NG], PxDy[NG], RxDy[NG];
#pragma unroll
              for (int f = 0; f < NF; ++f)
                {
                  const double a00 = Aq[f * NC2], a10 = Aq[f * NC2 + 1];
                  const double a01 = Aq[f * NC2 + NX], a11 = Aq[f * NC2 + NX + 1];
                  const double r0 = a01 - a00, r1 = a11 - a10;
                  const double b0 = (qy == 1) ? a00 + a01 : fma (ey, r0, a00 + a01);
                  const double b1 = (qy == 1) ? a10 + a11 : fma (ey, r1, a10 + a11);
                  PxB[f] = b0 + b1;
                  RxB[f] = b1 - b0;
                  if (f < NG)
                    {
                      const double gys = (f == 3) ? k.gp[1] : k.gu[1];
                      dx[f] = RxB[f] * ((f == 3) ? k.gp[0] : k.gu[0]);
                      PxDy[f] = (r0 + r1) * gys;
                      RxDy[f] = (r1 - r0) * gys;
                      const double z0 = BZ[(f * 3 + qy) * NXC + it0], z1 = BZ[(f * 3 + qy) * NXC + it0 + 1];
                      PxBz[f] = z0 + z1;
                      RxBz[f] = z1 - z0;
                    }
                }
              double XS[4], ZP[4], ZR[4], yP[4], yR[4];
              double AP = 0, AR = 0;
#pragma unroll
              for (int c = 0; c < 4; ++c)
                XS[c] = ZP[c] = ZR[c] = yP[c] = yR[c] = 0;

#pragma unroll
              for (int qx = 0; qx < 3; ++qx)
                {
                  const double ex = es[qx];
                  double U[3][3], gpf[3];
#pragma unroll
                  for (int c = 0; c < 3; ++c)
                    {
                      U[c][0] = dx[c];
                      U[c][1] = (qx == 1) ? PxDy[c] : fma (ex, RxDy[c], PxDy[c]);
                      U[c][2] = (qx == 1) ? PxBz[c] : fma (ex, RxBz[c], PxBz[c]);
                    }
                  gpf[0] = dx[3];
                  gpf[1] = (qx == 1) ? PxDy[3] : fma (ex, RxDy[3], PxDy[3]);
                  gpf[2] = (qx == 1) ? PxBz[3] : fma (ex, RxBz[3], PxBz[3]);
                  const double pf = (qx == 1) ? PxB[3] : fma (ex, RxB[3], PxB[3]);
                  double pte = (qx == 1) ? PxB[4] : fma (ex, RxB[4], PxB[4]);
                  if (p.clamp_extra)
                    pte = fmin (fmax (pte, 0.0), 1.0);

                  // weak form at the q-point (cracks.cc:2404-2429)
                  const double pe2 = pte * pte;
                  const double gdeg = fma (omk, pe2, p.kappa);
                  const double trU = U[0][0] + U[1][1] + U[2][2];
                  const double u01 = U[0][1] + U[1][0], u02 = U[0][2] + U[2][0], u12 = U[1][2] + U[2][1];
                  const double dd2 = fma (U[0][0], U[0][0], fma (U[1][1], U[1][1], U[2][2] * U[2][2]));
                  const double od2 = fma (u01, u01, fma (u02, u02, u12 * u12));
                  const double ltr = p.lambda * trU;
                  const double spE = fma (ltr, trU, two_mu * fma (0.5, od2, dd2));
                  const double w = k.wvol * k.wq[qx] * k.wq[qy] * k.wq[qz];
                  const double mw = -w;
                  // -(g sigma - P1 pte^2 I) w
                  const double mwg = mw * gdeg;
                  const double iso = fma (mwg, ltr, w * p.P1 * pe2);
                  const double mwg2m = mwg * two_mu, mwgm = mwg * p.mu;
                  const double S00 = fma (mwg2m, U[0][0], iso), S11 = fma (mwg2m, U[1][1], iso),
                               S22 = fma (mwg2m, U[2][2], iso);
                  const double S01 = mwgm * u01, S02 = mwgm * u02, S12 = mwgm * u12;
                  const double cphi = fma (omk * spE, pf, -c_gce * (1.0 - pf)) - 2.0 * p.P1 * pf * trU;
                  const double wa = mw * cphi, wb = mw * c_gceps;
                  const double fx[4] = {S00, S01, S02, wb * gpf[0]};
                  const double fy[4] = {S01, S11, S12, wb * gpf[1]};
                  const double fz[4] = {S02, S12, S22, wb * gpf[2]};
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                    {
                      XS[c] += fx[c];
                      yP[c] += fy[c];
                      ZP[c] += fz[c];
                      if (qx != 1)
                        {
                          yR[c] = fma (ex, fy[c], yR[c]);
                          ZR[c] = fma (ex, fz[c], ZR[c]);
                        }
                    }
                  AP += wa;
                  if (qx != 1)
                    AR = fma (ex, wa, AR);
                }
#pragma unroll
              for (int c = 0; c < 4; ++c)
                {
                  const double gxs = (c == 3) ? k.gp[0] : k.gu[0];
                  const double xv = XS[c] * gxs;
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
                  if (qy != 1)
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
      // ---- stage 4: plane -> shared tile (see pf_apply3d_v2.cuh for the phase argument)
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
                      const double yv = (vx == 0 ? YP[c] - YR[c] : YP[c] + YR[c]) * gys;
                      const double a = (vy == 0) ? VP[c][vx] - VR[c][vx] - yv : VP[c][vx] + VR[c][vx] + yv;
                      const double d = ((vy == 0) ? DP[c][vx] - DR[c][vx] : DP[c][vx] + DR[c][vx]) * gzs;
                      const double sc = (c == 3) ? 0.125 : 1.0;
                      val[c] = (vz == 0) ? sc * (fma (-ez, a, a) - d) : sc * (fma (ez, a, a) + d);
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

  // ---- flush
  for (int i = tid; i < NN; i += NT)
    {
      const int ix = i % NX, iy = (i / NX) % NY, iz = i / (NX * NY);
      const int gx = cx0 + ix, gy = cy0 + iy, gz = cz0 + iz;
      if (gx < nnx && gy < nny && gz <= g.cell_end)
        {
          const long long n = gx + (long long) nnx * gy + pstride * (gz - lz_off);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            atomicAdd (&r[4 * n + c], ys[c * NN + i]);
        }
    }
}

} // namespace pf
