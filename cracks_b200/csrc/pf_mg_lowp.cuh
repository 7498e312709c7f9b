// pf_mg_lowp.cuh -- the multigrid V-cycle in a second precision R (float).
//
// The preconditioner that stands in for the reference's ML AMG hierarchies
// (cracks.cc:2477-2497, 2717-2740) is only ever applied to Krylov vectors; its
// arithmetic is unpinned by the goldens and the outer GMRES,
// the Jacobian apply and every residual stay FP64.  Running the V-cycle in
// FP32 halves the bytes of its vector kernels and moves the smoother operator
// from the FP64 pipe to the FP32 pipe.  This header holds the R-typed twins
// of the kernels a V-cycle launches: the smoother operator (the 2-point-rule
// variant of pf_apply3d_v4.cuh, same staging and mapping), its row
// initialisation, the Chebyshev step, and the grid transfers.  Opt-in
// (pf_set_multigrid_precision / PF_MG_FP32=1) until it has been timed.
#pragma once
#include "pf_apply3d_v2.cuh"
#include "pf_multigrid.cuh"

namespace pf {

template <typename R> struct Real4;
template <> struct Real4<double> { using type = double4; };
template <> struct Real4<float> { using type = float4; };

__device__ __forceinline__ double rfma (double a, double b, double c) { return fma (a, b, c); }
__device__ __forceinline__ float rfma (float a, float b, float c) { return fmaf (a, b, c); }
__device__ __forceinline__ double rclamp01 (double a) { return fmin (fmax (a, 0.0), 1.0); }
__device__ __forceinline__ float rclamp01 (float a) { return fminf (fmaxf (a, 0.0f), 1.0f); }

template <typename R, int TX, int TY, int TZ> struct Tile3mg : Tile3v2<TX, TY, TZ>
{
  using B = Tile3v2<TX, TY, TZ>;
  // AZ [2][9][NC2], BZ [7][2][NXC], DZ [7][NC2] aliased with the y tile [4][NN]
  static constexpr size_t smem_reals = (size_t) 18 * B::NC2 + 14 * B::NXC + B::dz_or_y;
  static constexpr size_t smem_bytes = smem_reals * sizeof (R);
};

// y += J_2pt(U) x on one tile of TX x TY x TZ cells; y pre-initialised by k_apply_init_r.
// Same weak form as k_apply3d_v4 (cracks.cc:2359-2382) with the 2-point Gauss rule.
template <typename R, int TX, int TY, int TZ, int MINB, bool ISO>
__global__ void __launch_bounds__ (TX * TY * TZ, MINB)
k_apply3d_mg (Grid g, Phys p, K3 k, int tiles_x, int tiles_y, const R *__restrict__ x, const R *__restrict__ sol,
              const R *__restrict__ pt, const uint8_t *__restrict__ mask, R *__restrict__ y)
{
  using T = Tile3v2<TX, TY, TZ>;
  using R4 = typename Real4<R>::type;
  constexpr int NN = T::NN, NT = T::NT, NX = T::NX, NY = T::NY, NC2 = T::NC2, NXC = T::NXC;
  extern __shared__ __align__ (16) unsigned char smem_raw[];
  R *AZ = reinterpret_cast<R *> (smem_raw); // [2][9][NC2]
  R *BZ = AZ + 18 * NC2;                    // [7][2][NXC]
  R *DZ = BZ + 14 * NXC;                    // [7][NC2], stage 1 -> 2 only
  R *ys = DZ;                               // [4][NN], aliases DZ

  const int tid = threadIdx.x;
  int b = blockIdx.x;
  const int bx = b % tiles_x;
  b /= tiles_x;
  const int by = b % tiles_y;
  const int bz = b / tiles_y;
  const int cx0 = bx * TX, cy0 = by * TY, cz0 = g.cell_begin + bz * TZ * g.layer_stride;
  const int nnx = g.nn[0], nny = g.nn[1];
  const int lz_off = g.plane_begin;
  const long long pstride = g.nodes_per_plane;
  const R S = (R) k.s2;
  const R eighth = (R) 0.125;
  const R gu0 = (R) k.gu[0], gu1 = (R) k.gu[1], gu2 = (R) k.gu[2];
  const R gp0 = (R) k.gp[0], gp1 = (R) k.gp[1], gp2 = (R) k.gp[2];

  // ---- stage 1: z-collapse per node column --------------------------------
  for (int i = tid; i < NC2; i += NT)
    {
      const int ix = i % NX, iy = (i / NX) % NY, tz = i / (NX * NY);
      const int gx = cx0 + ix, gy = cy0 + iy, gz = cz0 + tz;
      R f0[9], f1[9];
#pragma unroll
      for (int f = 0; f < 9; ++f)
        f0[f] = f1[f] = 0;
      if (gx < nnx && gy < nny && gz < g.cell_end)
        {
          const long long n0 = gx + (long long) nnx * gy + pstride * (gz - lz_off);
          const long long n1 = n0 + pstride;
          const R4 xa = *reinterpret_cast<const R4 *> (x + 4 * n0);
          const R4 xb = *reinterpret_cast<const R4 *> (x + 4 * n1);
          const R4 sa = *reinterpret_cast<const R4 *> (sol + 4 * n0);
          const R4 sb = *reinterpret_cast<const R4 *> (sol + 4 * n1);
          const uint8_t m0 = mask[n0], m1 = mask[n1];
          f0[0] = (m0 & 1) ? (R) 0 : xa.x;
          f0[1] = (m0 & 2) ? (R) 0 : xa.y;
          f0[2] = (m0 & 4) ? (R) 0 : xa.z;
          f0[3] = (m0 & 8) ? (R) 0 : eighth * xa.w;
          f1[0] = (m1 & 1) ? (R) 0 : xb.x;
          f1[1] = (m1 & 2) ? (R) 0 : xb.y;
          f1[2] = (m1 & 4) ? (R) 0 : xb.z;
          f1[3] = (m1 & 8) ? (R) 0 : eighth * xb.w;
          f0[4] = sa.x, f0[5] = sa.y, f0[6] = sa.z, f0[7] = eighth * sa.w, f0[8] = eighth * pt[n0];
          f1[4] = sb.x, f1[5] = sb.y, f1[6] = sb.z, f1[7] = eighth * sb.w, f1[8] = eighth * pt[n1];
        }
#pragma unroll
      for (int f = 0; f < 9; ++f)
        {
          const R s = f0[f] + f1[f], r = f1[f] - f0[f];
          AZ[(0 * 9 + f) * NC2 + i] = rfma (-S, r, s);
          AZ[(1 * 9 + f) * NC2 + i] = rfma (S, r, s);
          if (f < 7)
            DZ[f * NC2 + i] = ISO ? r : r * ((f == 3) ? gp2 : gu2);
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
          const R d0 = DZ[f * NC2 + c0], d1 = DZ[f * NC2 + c0 + NX];
          const R P = d0 + d1, Rd = d1 - d0;
          BZ[(f * 2 + 0) * NXC + i] = rfma (-S, Rd, P);
          BZ[(f * 2 + 1) * NXC + i] = rfma (S, Rd, P);
        }
    }
  __syncthreads ();
  for (int i = tid; i < 4 * NN; i += NT)
    ys[i] = 0;
  __syncthreads ();

  // ---- stage 3: one thread per cell, plane -> row -> point --------------------
  {
    const int tx = tid % TX, ty = (tid / TX) % TY, tz = tid / (TX * TY);
    const bool valid = (cx0 + tx < g.n[0]) && (cy0 + ty < g.n[1]) && (cz0 + tz < g.cell_end);
    const int c00 = tx + NX * (ty + NY * tz);
    const int it0 = tx + NX * (ty + TY * tz);
    const int nbase = tx + T::SY * ty + T::SZ * tz;

    // the constants are formed in double and rounded once
    const double gam = k.gu[0];
    const double omk_d = 1.0 - p.kappa;
    const double vs_d = ISO ? 0.125 : 1.0;
    const R omk = (R) omk_d, kappa = (R) p.kappa;
    const R lamq = (R) (ISO ? p.lambda * gam * gam : p.lambda);
    const R muq = (R) (ISO ? p.mu * gam * gam : p.mu);
    const R two_mu = (R) 2 * muq;
    const R c_gceps = (R) (ISO ? p.G_c * p.eps * 8.0 * gam * gam : p.G_c * p.eps);
    const R ca1 = (R) (2.0 * omk_d * vs_d), ca2 = (R) (2.0 * (ISO ? p.P1 * gam : p.P1) * vs_d), ca3 = (R) (omk_d * vs_d),
            ca4 = (R) (p.G_c / p.eps * vs_d);
    const R w = (R) k.wvol; // 2-point rule: unit weights
    const R wb = w * c_gceps;
    const R half = (R) 0.5;
    const bool clamp = p.clamp_extra != 0;

#pragma unroll 1
    for (int qz = 0; qz < 2; ++qz)
      {
        const R ez = qz == 0 ? -S : S;
        const R *Aq = AZ + qz * 9 * NC2 + c00;
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
#pragma unroll
            for (int qy = 0; qy < 2; ++qy)
              {
                const R ey = qy == 0 ? -S : S;
                R PxB[9], RxB[9], PxBz[7], RxBz[7], dx[7], PxDy[7], RxDy[7];
#pragma unroll
                for (int f = 0; f < 9; ++f)
                  {
                    const R a00 = Aq[f * NC2], a10 = Aq[f * NC2 + 1];
                    const R a01 = Aq[f * NC2 + NX], a11 = Aq[f * NC2 + NX + 1];
                    const R r0 = a01 - a00, r1 = a11 - a10;
                    const R b0 = rfma (ey, r0, a00 + a01);
                    const R b1 = rfma (ey, r1, a10 + a11);
                    PxB[f] = b0 + b1;
                    RxB[f] = b1 - b0;
                    if (f < 7)
                      {
                        const R gys = (f == 3) ? gp1 : gu1;
                        dx[f] = ISO ? RxB[f] : RxB[f] * ((f == 3) ? gp0 : gu0);
                        PxDy[f] = ISO ? r0 + r1 : (r0 + r1) * gys;
                        RxDy[f] = ISO ? r1 - r0 : (r1 - r0) * gys;
                        const R z0 = BZ[(f * 2 + qy) * NXC + it0], z1 = BZ[(f * 2 + qy) * NXC + it0 + 1];
                        PxBz[f] = z0 + z1;
                        RxBz[f] = z1 - z0;
                      }
                  }
                // symmetric off-diagonal strain sums, linear in xi_x: P + ex R
                const R oP01 = PxDy[0] + dx[1], oP02 = PxBz[0] + dx[2], oP12 = PxBz[1] + PxDy[2], oR12 = RxBz[1] + RxDy[2];
                const R uP01 = PxDy[4] + dx[5], uP02 = PxBz[4] + dx[6], uP12 = PxBz[5] + PxDy[6], uR12 = RxBz[5] + RxDy[6];
                R XS[4], ZP[4], ZR[4], yP[4], yR[4];
                R AP = 0, AR = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  XS[c] = ZP[c] = ZR[c] = yP[c] = yR[c] = 0;

#pragma unroll
                for (int qx = 0; qx < 2; ++qx)
                  {
                    const R ex = qx == 0 ? -S : S;
                    const R G00 = dx[0], U00 = dx[4];
                    const R G11 = rfma (ex, RxDy[1], PxDy[1]), G22 = rfma (ex, RxBz[2], PxBz[2]);
                    const R U11 = rfma (ex, RxDy[5], PxDy[5]), U22 = rfma (ex, RxBz[6], PxBz[6]);
                    const R o01 = rfma (ex, RxDy[0], oP01), o02 = rfma (ex, RxBz[0], oP02), o12 = rfma (ex, oR12, oP12);
                    const R u01 = rfma (ex, RxDy[4], uP01), u02 = rfma (ex, RxBz[4], uP02), u12 = rfma (ex, uR12, uP12);
                    const R gph0 = dx[3], gph1 = rfma (ex, RxDy[3], PxDy[3]), gph2 = rfma (ex, RxBz[3], PxBz[3]);
                    const R dphi = rfma (ex, RxB[3], PxB[3]);
                    const R pf = rfma (ex, RxB[7], PxB[7]);
                    R pte = rfma (ex, RxB[8], PxB[8]);
                    if (clamp)
                      pte = rclamp01 (pte);

                    const R gdeg = rfma (omk * pte, pte, kappa);
                    const R trU = U00 + U11 + U22;
                    const R trG = G00 + G11 + G22;
                    const R ddot = rfma (U00, G00, rfma (U11, G11, U22 * G22));
                    const R odot = rfma (u01, o01, rfma (u02, o02, u12 * o12));
                    const R spG = rfma (lamq * trU, trG, two_mu * rfma (half, odot, ddot));
                    const R dd2 = rfma (U00, U00, rfma (U11, U11, U22 * U22));
                    const R od2 = rfma (u01, u01, rfma (u02, u02, u12 * u12));
                    const R spE = rfma (lamq * trU, trU, two_mu * rfma (half, od2, dd2));
                    const R a_val = pf * (ca1 * spG - ca2 * trG) + dphi * (rfma (ca3, spE, ca4) - ca2 * trU);
                    const R wg = w * gdeg;
                    const R wgl = wg * lamq * trG, wgm = wg * muq, wg2m = wg * two_mu;
                    const R S00 = rfma (wg2m, G00, wgl), S11 = rfma (wg2m, G11, wgl), S22 = rfma (wg2m, G22, wgl);
                    const R S01 = wgm * o01, S02 = wgm * o02, S12 = wgm * o12;
                    const R wa = w * a_val;
                    const R fx[4] = {S00, S01, S02, wb * gph0};
                    const R fy[4] = {S01, S11, S12, wb * gph1};
                    const R fz[4] = {S02, S12, S22, wb * gph2};
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                      {
                        XS[c] += fx[c];
                        yP[c] += fy[c];
                        ZP[c] += fz[c];
                        yR[c] = rfma (ex, fy[c], yR[c]);
                        ZR[c] = rfma (ex, fz[c], ZR[c]);
                      }
                    AP += wa;
                    AR = rfma (ex, wa, AR);
                  }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  {
                    const R gxs = (c == 3) ? gp0 : gu0;
                    const R xv = ISO ? XS[c] : XS[c] * gxs;
                    R v0 = -xv, v1 = xv;
                    if (c == 3)
                      {
                        v0 += AP - AR;
                        v1 += AP + AR;
                      }
                    VP[c][0] += v0;
                    VP[c][1] += v1;
                    const R z0 = ZP[c] - ZR[c], z1 = ZP[c] + ZR[c];
                    DP[c][0] += z0;
                    DP[c][1] += z1;
                    VR[c][0] = rfma (ey, v0, VR[c][0]);
                    VR[c][1] = rfma (ey, v1, VR[c][1]);
                    DR[c][0] = rfma (ey, z0, DR[c][0]);
                    DR[c][1] = rfma (ey, z1, DR[c][1]);
                    YP[c] += yP[c];
                    YR[c] += yR[c];
                  }
              }
          }
        // ---- stage 4: plane -> shared y tile, 8 conflict-free phases (see pf_apply3d_v4.cuh) ----
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
                        const R gys = (c == 3) ? gp1 : gu1;
                        const R gzs = (c == 3) ? gp2 : gu2;
                        const R yv0 = vx == 0 ? YP[c] - YR[c] : YP[c] + YR[c];
                        const R yv = ISO ? yv0 : yv0 * gys;
                        const R a = (vy == 0) ? VP[c][vx] - VR[c][vx] - yv : VP[c][vx] + VR[c][vx] + yv;
                        const R d0 = (vy == 0) ? DP[c][vx] - DR[c][vx] : DP[c][vx] + DR[c][vx];
                        const R d = ISO ? d0 : d0 * gzs;
                        const R v = (vz == 0) ? rfma (-ez, a, a) - d : rfma (ez, a, a) + d;
                        val[c] = (c == 3 && !ISO) ? eighth * v : v;
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

// y = constrained ? diag * x : 0 (3-D, 4 components per node); the level keeps 1 / diag
template <typename R>
__global__ void
k_apply_init_r (long long n_nodes, const R *__restrict__ x, const R *__restrict__ idiag, const uint8_t *__restrict__ mask,
                R *__restrict__ y)
{
  using R4 = typename Real4<R>::type;
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes)
    return;
  const uint8_t m = mask[n];
  R4 out;
  out.x = out.y = out.z = out.w = 0;
  if (m & 15)
    {
      const R4 xv = *reinterpret_cast<const R4 *> (x + 4 * n), dv = *reinterpret_cast<const R4 *> (idiag + 4 * n);
      out.x = (m & 1) ? xv.x / dv.x : (R) 0;
      out.y = (m & 2) ? xv.y / dv.y : (R) 0;
      out.z = (m & 4) ? xv.z / dv.z : (R) 0;
      out.w = (m & 8) ? xv.w / dv.w : (R) 0;
    }
  *reinterpret_cast<R4 *> (y + 4 * n) = out;
}

// dst = (B) src
template <typename A, typename B>
__global__ void
k_convert (long long n, const A *__restrict__ src, B *__restrict__ dst)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
    dst[i] = (B) src[i];
}

// dst = 1 / src: the V-cycle multiplies by the inverse diagonal
template <typename A, typename B>
__global__ void
k_convert_inverse (long long n, const A *__restrict__ src, B *__restrict__ dst)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
    dst[i] = (B) (1.0 / src[i]);
}

// one Chebyshev step (see k_cheb_step) with the inverse diagonal; `mode` 0: r = b - y, x += d;
// 1: first step from a zero guess (r = b, x = d); 2: first step from a given x (r = b - y, d = c2 r / diag, x += d)
template <typename R>
__global__ void
k_cheb_step_r (long long n, int mode, R c1, R c2, const R *__restrict__ b, const R *__restrict__ y,
               const R *__restrict__ idiag, R *__restrict__ d, R *__restrict__ x)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
    {
      const R r = mode == 1 ? b[i] : b[i] - y[i];
      const R dn = mode == 0 ? rfma (c1, d[i], c2 * r * idiag[i]) : c2 * r * idiag[i];
      d[i] = dn;
      x[i] = mode == 1 ? dn : x[i] + dn;
    }
}

template <typename R>
__global__ void
k_sub_r (long long n, const R *__restrict__ b, const R *__restrict__ y, R *__restrict__ r)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
    r[i] = b[i] - y[i];
}

// xf += P xc (trilinear), constrained fine dofs receive nothing
template <typename R>
__global__ void
k_prolong_add_r (Dims3 dc, Dims3 df, int ka, int ke, const R *__restrict__ xc, const uint8_t *__restrict__ fmask,
                 R *__restrict__ xf)
{
  using R4 = typename Real4<R>::type;
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long) df.n[0] * df.n[1] * (ke - ka);
  if (t >= total)
    return;
  const int i = (int) (t % df.n[0]), j = (int) ((t / df.n[0]) % df.n[1]), k = ka + (int) (t / ((long long) df.n[0] * df.n[1]));
  const long long n = node_id (df, i, j, k);
  const int i0 = i >> 1, j0 = j >> 1, k0 = k >> 1;
  const int ni = i & 1, nj = j & 1, nk = k & 1;
  const R w = (R) ((ni ? 0.5 : 1.0) * (nj ? 0.5 : 1.0) * (nk ? 0.5 : 1.0));
  R acc[4] = {0, 0, 0, 0};
  for (int c3 = 0; c3 <= nk; ++c3)
    for (int c2 = 0; c2 <= nj; ++c2)
      for (int c1 = 0; c1 <= ni; ++c1)
        {
          const long long cn = node_id (dc, i0 + c1, j0 + c2, k0 + c3);
          const R4 v = *reinterpret_cast<const R4 *> (xc + 4 * cn);
          acc[0] += w * v.x;
          acc[1] += w * v.y;
          acc[2] += w * v.z;
          acc[3] += w * v.w;
        }
  const uint8_t m = fmask[n];
  R4 o = *reinterpret_cast<R4 *> (xf + 4 * n);
  o.x += (m & 1) ? (R) 0 : acc[0];
  o.y += (m & 2) ? (R) 0 : acc[1];
  o.z += (m & 4) ? (R) 0 : acc[2];
  o.w += (m & 8) ? (R) 0 : acc[3];
  *reinterpret_cast<R4 *> (xf + 4 * n) = o;
}

// rc = P^T rf with constrained fine rows treated as zero; constrained coarse rows get zero
template <typename R>
__global__ void
k_restrict_r (Dims3 dc, Dims3 df, int Ka, int Ke, const R *__restrict__ rf, const uint8_t *__restrict__ fmask,
              const uint8_t *__restrict__ cmask, R *__restrict__ rc)
{
  using R4 = typename Real4<R>::type;
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long) dc.n[0] * dc.n[1] * (Ke - Ka);
  if (t >= total)
    return;
  const int I = (int) (t % dc.n[0]), J = (int) ((t / dc.n[0]) % dc.n[1]), K = Ka + (int) (t / ((long long) dc.n[0] * dc.n[1]));
  const long long n = node_id (dc, I, J, K);
  R acc[4] = {0, 0, 0, 0};
  for (int dk = -1; dk <= 1; ++dk)
    for (int dj = -1; dj <= 1; ++dj)
      for (int di = -1; di <= 1; ++di)
        {
          const int i = 2 * I + di, j = 2 * J + dj, k = 2 * K + dk;
          if (i < 0 || j < 0 || k < 0 || i >= df.n[0] || j >= df.n[1] || k >= df.n[2])
            continue;
          const R w = (R) ((di ? 0.5 : 1.0) * (dj ? 0.5 : 1.0) * (dk ? 0.5 : 1.0));
          const long long fn = node_id (df, i, j, k);
          const uint8_t m = fmask[fn];
          const R4 v = *reinterpret_cast<const R4 *> (rf + 4 * fn);
          acc[0] += (m & 1) ? (R) 0 : w * v.x;
          acc[1] += (m & 2) ? (R) 0 : w * v.y;
          acc[2] += (m & 4) ? (R) 0 : w * v.z;
          acc[3] += (m & 8) ? (R) 0 : w * v.w;
        }
  const uint8_t m = cmask[n];
  R4 o;
  o.x = (m & 1) ? (R) 0 : acc[0];
  o.y = (m & 2) ? (R) 0 : acc[1];
  o.z = (m & 4) ? (R) 0 : acc[2];
  o.w = (m & 8) ? (R) 0 : acc[3];
  *reinterpret_cast<R4 *> (rc + 4 * n) = o;
}

} // namespace pf
