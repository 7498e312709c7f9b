// pf_apply3d.cuh -- the hot kernel: y += J(U) x for dim = 3, Q1, uniform brick.
//
// Replaces "assemble 32x32 cell matrices into Trilinos, then SpMV"
// (cracks.cc:2308-2389 + 2459-2463 + 2770) by a matrix-free evaluation of the
// same bilinear form with the same 27-point Gauss rule.
//
// Mapping: one CTA owns a TX x TY x TZ tile of cells, one thread per cell.
//   1. the (TX+1)(TY+1)(TZ+1) nodal halo of x (constrained columns zeroed),
//      of the state displacement, phi and the extrapolated phi~ is staged in
//      shared memory as SoA (conflict-free for x-contiguous threads);
//   2. each thread sum-factorises its cell: z-collapse per q-plane, y-collapse
//      per q-row, x per q-point, exploiting that d/dx of a Q1 function does
//      not depend on xi_x; the weak form is evaluated at the q-point and the
//      transposed collapse accumulates the 32 local outputs in registers;
//   3. the outputs are summed into a shared-memory y tile in 8 conflict-free
//      phases (phase k = local vertex k, distinct cells hit distinct nodes),
//      and the tile is flushed with one red.global.add.f64 per (node, comp).
//
// FP64 throughout; no tensor cores (tcgen05 has no f64 kind).
#pragma once
#include "pf_common.cuh"

namespace pf {

struct K3
{
  double s;          // sqrt(3/5): Gauss abscissa on [-1,1]
  double s2;         // sqrt(1/3): abscissa of the 2-point rule (preconditioner-only operator)
  double gu[3];      // 1/(4 h_d): gradient scale of an unscaled nodal field
  double gp[3];      // 2/h_d:     gradient scale of a field pre-scaled by 1/8
  double wq[3];      // per-direction Gauss weights (5/9, 8/9, 5/9)
  double wvol;       // hx hy hz / 8
  double ih[3];      // 1/h_d
};

template <int TX, int TY, int TZ> struct Tile3
{
  static constexpr int NX = TX + 1, NY = TY + 1, NZ = TZ + 1;
  static constexpr int NN = NX * NY * NZ;
  static constexpr int NT = TX * TY * TZ;
  static constexpr int SY = NX, SZ = NX * NY;
  // doubles: x[4], u[3], phi, pt, y[4] = 13 per node, + 1 byte mask
  static constexpr size_t smem_bytes = (size_t) NN * 13 * sizeof (double) + ((NN + 15) / 16) * 16;
};

template <int TX, int TY, int TZ>
__global__ void __launch_bounds__ (TX * TY * TZ)
k_apply3d (Grid g, Phys p, K3 k, int tiles_x, int tiles_y,
           const double *__restrict__ x, const double *__restrict__ sol,
           const double *__restrict__ pt, const uint8_t *__restrict__ mask,
           double *__restrict__ y)
{
  using T = Tile3<TX, TY, TZ>;
  constexpr int NN = T::NN, NT = T::NT, SY = T::SY, SZ = T::SZ;
  extern __shared__ __align__ (16) unsigned char smem_raw[];
  double *xs = reinterpret_cast<double *> (smem_raw); // [4][NN]
  double *us = xs + 4 * NN;                           // [3][NN]
  double *ps = us + 3 * NN;                           // [NN] phi / 8
  double *ts = ps + NN;                               // [NN] pt / 8
  double *ys = ts + NN;                               // [4][NN]
  uint8_t *ms = reinterpret_cast<uint8_t *> (ys + 4 * NN);

  const int tid = threadIdx.x;
  int b = blockIdx.x;
  const int bx = b % tiles_x;
  b /= tiles_x;
  const int by = b % tiles_y;
  const int bz = b / tiles_y;
  // tile origin in cells; z in local cell layers
  const int cx0 = bx * TX, cy0 = by * TY, cz0 = g.cell_begin + bz * TZ;
  const int nnx = g.nn[0], nny = g.nn[1];
  const int lz_off = g.plane_begin;        // local plane = global plane - plane_begin
  const long long pstride = g.nodes_per_plane;

  // ---- 1. stage the nodal halo ------------------------------------------
  for (int i = tid; i < NN; i += NT)
    {
      const int ix = i % T::NX, iy = (i / T::NX) % T::NY, iz = i / (T::NX * T::NY);
      const int gx = cx0 + ix, gy = cy0 + iy, gz = cz0 + iz;
      double4 xv = make_double4 (0, 0, 0, 0), sv = make_double4 (0, 0, 0, 0);
      double tv = 0;
      uint8_t m = 0xff;
      if (gx < nnx && gy < nny && gz <= g.cell_end)
        {
          const long long n = gx + (long long) nnx * gy + pstride * (gz - lz_off);
          xv = *reinterpret_cast<const double4 *> (x + 4 * n);
          sv = *reinterpret_cast<const double4 *> (sol + 4 * n);
          tv = pt[n];
          m = mask[n];
        }
      xs[i] = (m & 1) ? 0.0 : xv.x;
      xs[NN + i] = (m & 2) ? 0.0 : xv.y;
      xs[2 * NN + i] = (m & 4) ? 0.0 : xv.z;
      xs[3 * NN + i] = (m & 8) ? 0.0 : 0.125 * xv.w;
      us[i] = sv.x;
      us[NN + i] = sv.y;
      us[2 * NN + i] = sv.z;
      ps[i] = 0.125 * sv.w;
      ts[i] = 0.125 * tv;
      ys[i] = 0;
      ys[NN + i] = 0;
      ys[2 * NN + i] = 0;
      ys[3 * NN + i] = 0;
      ms[i] = m;
    }
  __syncthreads ();

  // ---- 2. per-cell evaluation --------------------------------------------
  const int tx = tid % TX, ty = (tid / TX) % TY, tz = tid / (TX * TY);
  const bool valid = (cx0 + tx < g.n[0]) && (cy0 + ty < g.n[1]) && (cz0 + tz < g.cell_end);
  const int base = tx + SY * ty + SZ * tz;

  double out[4][8];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int v = 0; v < 8; ++v)
      out[c][v] = 0;

  if (valid)
    {
      // physics constants
      const double omk = 1.0 - p.kappa;
      const double c_gce = p.G_c / p.eps;
      const double c_gceps = p.G_c * p.eps;
      const double two_mu = 2.0 * p.mu;
      const double es[3] = {-k.s, 0.0, k.s};

#pragma unroll 1
      for (int qz = 0; qz < 3; ++qz)
        {
          const double ez = es[qz];
          // plane level: for field f and x-column vx, the y-pair (P,R) of the
          // z-collapsed value A and of the z-difference D.
          // fields 0..2: trial u, 3: trial phi, 4..6: state u, 7: state phi, 8: pt
          double PyA[9][2], RyA[9][2], PyD[7][2], RyD[7][2];
#pragma unroll
          for (int f = 0; f < 9; ++f)
            {
              const double *F = (f < 4) ? xs + f * NN : (f < 7) ? us + (f - 4) * NN : (f == 7) ? ps : ts;
#pragma unroll
              for (int vx = 0; vx < 2; ++vx)
                {
                  const double f00 = F[base + vx], f01 = F[base + vx + SZ];
                  const double f10 = F[base + vx + SY], f11 = F[base + vx + SY + SZ];
                  const double r0 = f01 - f00, r1 = f11 - f10;
                  const double a0 = fma (ez, r0, f00 + f01), a1 = fma (ez, r1, f10 + f11);
                  PyA[f][vx] = a0 + a1;
                  RyA[f][vx] = a1 - a0;
                  if (f < 7)
                    {
                      const double gzs = (f == 3) ? k.gp[2] : k.gu[2];
                      PyD[f][vx] = (r0 + r1) * gzs;
                      RyD[f][vx] = (r1 - r0) * gzs;
                    }
                }
            }
          // d/dy depends on (qx, qz) only: x-pair of RyA, scaled
          double PxDy[7], RxDy[7];
#pragma unroll
          for (int f = 0; f < 7; ++f)
            {
              const double gys = (f == 3) ? k.gp[1] : k.gu[1];
              PxDy[f] = (RyA[f][0] + RyA[f][1]) * gys;
              RxDy[f] = (RyA[f][1] - RyA[f][0]) * gys;
            }
          // plane-level output accumulators, per x-column vx and y-pair form:
          // VP/VR: value-like part (P and R coefficient in y), DP/DR: z-derivative part
          double VP[4][2], VR[4][2], DP[4][2], DR[4][2];
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int vx = 0; vx < 2; ++vx)
              VP[c][vx] = VR[c][vx] = DP[c][vx] = DR[c][vx] = 0;
          // y-derivative part of the output depends on (qx,qz) only: accumulate
          // over the rows in x-pair form
          double YP[4], YR[4];
#pragma unroll
          for (int c = 0; c < 4; ++c)
            YP[c] = YR[c] = 0;

#pragma unroll
          for (int qy = 0; qy < 3; ++qy)
            {
              const double ey = es[qy];
              // row level: x-pairs of value chain B and z-derivative chain Bz
              double PxB[9], RxB[9], PxBz[7], RxBz[7], dx[7];
#pragma unroll
              for (int f = 0; f < 9; ++f)
                {
                  const double b0 = (qy == 1) ? PyA[f][0] : fma (ey, RyA[f][0], PyA[f][0]);
                  const double b1 = (qy == 1) ? PyA[f][1] : fma (ey, RyA[f][1], PyA[f][1]);
                  PxB[f] = b0 + b1;
                  RxB[f] = b1 - b0;
                  if (f < 7)
                    {
                      const double z0 = (qy == 1) ? PyD[f][0] : fma (ey, RyD[f][0], PyD[f][0]);
                      const double z1 = (qy == 1) ? PyD[f][1] : fma (ey, RyD[f][1], PyD[f][1]);
                      PxBz[f] = z0 + z1;
                      RxBz[f] = z1 - z0;
                      dx[f] = RxB[f] * ((f == 3) ? k.gp[0] : k.gu[0]);
                    }
                }
              // row-level output accumulators (x-pair form)
              double XS[4];          // x-derivative part (independent of qx)
              double ZP[4], ZR[4];   // z-derivative part
              double AP = 0, AR = 0; // phi value part
#pragma unroll
              for (int c = 0; c < 4; ++c)
                XS[c] = ZP[c] = ZR[c] = 0;
              double yP[4], yR[4];
#pragma unroll
              for (int c = 0; c < 4; ++c)
                yP[c] = yR[c] = 0;

#pragma unroll
              for (int qx = 0; qx < 3; ++qx)
                {
                  const double ex = es[qx];
                  // ---- interpolate to the q-point
                  double G[3][3], U[3][3], gph[3];
#pragma unroll
                  for (int c = 0; c < 3; ++c)
                    {
                      G[c][0] = dx[c];
                      G[c][1] = (qx == 1) ? PxDy[c] : fma (ex, RxDy[c], PxDy[c]);
                      G[c][2] = (qx == 1) ? PxBz[c] : fma (ex, RxBz[c], PxBz[c]);
                      U[c][0] = dx[4 + c];
                      U[c][1] = (qx == 1) ? PxDy[4 + c] : fma (ex, RxDy[4 + c], PxDy[4 + c]);
                      U[c][2] = (qx == 1) ? PxBz[4 + c] : fma (ex, RxBz[4 + c], PxBz[4 + c]);
                    }
                  gph[0] = dx[3];
                  gph[1] = (qx == 1) ? PxDy[3] : fma (ex, RxDy[3], PxDy[3]);
                  gph[2] = (qx == 1) ? PxBz[3] : fma (ex, RxBz[3], PxBz[3]);
                  const double dphi = (qx == 1) ? PxB[3] : fma (ex, RxB[3], PxB[3]);
                  const double pf = (qx == 1) ? PxB[7] : fma (ex, RxB[7], PxB[7]);
                  double pte = (qx == 1) ? PxB[8] : fma (ex, RxB[8], PxB[8]);
                  if (p.clamp_extra)
                    pte = fmin (fmax (pte, 0.0), 1.0);

                  // ---- weak form at the q-point (cracks.cc:2359-2382)
                  const double gdeg = fma (omk * pte, pte, p.kappa);
                  const double trU = U[0][0] + U[1][1] + U[2][2];
                  const double trG = G[0][0] + G[1][1] + G[2][2];
                  const double o01 = G[0][1] + G[1][0], o02 = G[0][2] + G[2][0], o12 = G[1][2] + G[2][1];
                  const double u01 = U[0][1] + U[1][0], u02 = U[0][2] + U[2][0], u12 = U[1][2] + U[2][1];
                  // sigma(u):G  and  sigma(u):E(u)
                  const double ddot = fma (U[0][0], G[0][0], fma (U[1][1], G[1][1], U[2][2] * G[2][2]));
                  const double odot = fma (u01, o01, fma (u02, o02, u12 * o12));
                  const double spG = fma (p.lambda * trU, trG, two_mu * fma (0.5, odot, ddot));
                  const double dd2 = fma (U[0][0], U[0][0], fma (U[1][1], U[1][1], U[2][2] * U[2][2]));
                  const double od2 = fma (u01, u01, fma (u02, u02, u12 * u12));
                  const double spE = fma (p.lambda * trU, trU, two_mu * fma (0.5, od2, dd2));
                  const double a_val = pf * (2.0 * omk * spG - 2.0 * p.P1 * trG)
                                       + dphi * (fma (omk, spE, c_gce) - 2.0 * p.P1 * trU);
                  const double w = k.wvol * k.wq[qx] * k.wq[qy] * k.wq[qz];
                  const double wg = w * gdeg;
                  const double wgl = wg * p.lambda * trG, wgm = wg * p.mu, wg2m = wg * two_mu;
                  // w * Sigma (symmetric)
                  const double S00 = fma (wg2m, G[0][0], wgl), S11 = fma (wg2m, G[1][1], wgl),
                               S22 = fma (wg2m, G[2][2], wgl);
                  const double S01 = wgm * o01, S02 = wgm * o02, S12 = wgm * o12;
                  const double wa = w * a_val, wb = w * c_gceps;
                  const double fx[4] = {S00, S01, S02, wb * gph[0]};
                  const double fy[4] = {S01, S11, S12, wb * gph[1]};
                  const double fz[4] = {S02, S12, S22, wb * gph[2]};
                  // ---- transposed x-collapse
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
              // ---- transposed y-collapse: row -> plane accumulators
#pragma unroll
              for (int c = 0; c < 4; ++c)
                {
                  const double gxs = (c == 3) ? k.gp[0] : k.gu[0];
                  // x-derivative part acts as -/+ on vx = 0/1, value-like in y
                  const double xv = XS[c] * gxs;
                  double v0 = -xv, v1 = xv;
                  if (c == 3)
                    {
                      // phi value part: x-pair (AP, AR) -> vx0 = AP - AR, vx1 = AP + AR
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
          // ---- transposed z-collapse: plane -> nodal outputs
#pragma unroll
          for (int c = 0; c < 4; ++c)
            {
              const double gys = (c == 3) ? k.gp[1] : k.gu[1];
              const double gzs = (c == 3) ? k.gp[2] : k.gu[2];
#pragma unroll
              for (int vx = 0; vx < 2; ++vx)
                {
                  // y-derivative part: x-pair (YP, YR) -> this vx, then -/+ on vy
                  const double yv = (vx == 0 ? YP[c] - YR[c] : YP[c] + YR[c]) * gys;
                  // value-like at (vx, vy): VP -/+ VR ; then z-pair with ez
                  const double a0 = VP[c][vx] - VR[c][vx] - yv; // vy = 0
                  const double a1 = VP[c][vx] + VR[c][vx] + yv; // vy = 1
                  const double d0 = (DP[c][vx] - DR[c][vx]) * gzs;
                  const double d1 = (DP[c][vx] + DR[c][vx]) * gzs;
                  // vertex v = vx + 2 vy + 4 vz
                  out[c][vx + 0] += fma (-ez, a0, a0) - d0;
                  out[c][vx + 4] += fma (ez, a0, a0) + d0;
                  out[c][vx + 2] += fma (-ez, a1, a1) - d1;
                  out[c][vx + 6] += fma (ez, a1, a1) + d1;
                }
            }
        }
    }

  // ---- 3. conflict-free accumulation into the y tile ----------------------
#pragma unroll
  for (int v = 0; v < 8; ++v)
    {
      const int n = base + (v & 1) + SY * ((v >> 1) & 1) + SZ * (v >> 2);
      if (valid)
        {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            ys[c * NN + n] += out[c][v];
          // phi row: the 1/8 of the trilinear value weights (gp[] = 2/h already
          // carries the matching factor for the gradient parts)
          ys[3 * NN + n] += 0.125 * out[3][v];
        }
      __syncthreads ();
    }
  for (int i = tid; i < NN; i += NT)
    {
      const int ix = i % T::NX, iy = (i / T::NX) % T::NY, iz = i / (T::NX * T::NY);
      const int gx = cx0 + ix, gy = cy0 + iy, gz = cz0 + iz;
      if (gx < nnx && gy < nny && gz <= g.cell_end)
        {
          const long long n = gx + (long long) nnx * gy + pstride * (gz - lz_off);
          const uint8_t m = ms[i];
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (!((m >> c) & 1))
              atomicAdd (&y[4 * n + c], ys[c * NN + i]);
        }
    }
}

} // namespace pf
