// pf_apply3d_phi.cuh -- the (phi,phi) block of the Jacobian alone, y_phi += B(U) x_phi, on cubic cells.
//
// The reference's Jacobian is block lower triangular: the linearised stresses are zeroed for phi trial functions
// (cracks.cc:2333-2337), so block (u,phi) is identically zero and J dx = b splits into  A du = b_u  followed by
// B dphi = b_phi - C du  (SURVEY.md 8f-1).  After the first Newton step of a time step the u equation -- linear in u
// for the extrapolated phi~ (cracks.cc:2262-2277, 2404-2410) -- is solved, and the remaining Newton / active-set
// steps only move phi: their linear solves need B alone.  B = sum_q c2_q psi_i psi_j + G_c eps grad psi_i . grad psi_j
// (cracks.cc:2377-2382): the reaction coefficient c2 is the second scalar of the cached coefficient records of
// pf_apply3d_v6.cuh (k_point_coeffs), the Laplacian is the closed form used there (exact for the 2-point rule of the
// multigrid smoother as well).  About 350 flops per cell instead of 2400: the kernel is bound by the record stream.
//
// One thread per cell on the record tiles of v6 (TX x TY cells, W cells per record group); the vectors keep the
// 4-component node-major layout, only component 3 is read and written.  CS = type of the records, V = type of the
// vectors and of the arithmetic.
#pragma once
#include "pf_apply3d_v6.cuh"

namespace pf {

template <typename CS, typename V, int TX, int TY, int W, int NQ>
__global__ void __launch_bounds__ (TX * TY)
k_apply3d_phi (Grid g, K6 k, int tiles_x, int tiles_y, int layer0, const V *__restrict__ x,
               const uint8_t *__restrict__ mask, const CS *__restrict__ coef, V *__restrict__ y)
{
  using T = Tile3v6<TX, TY, NQ, W>;
  constexpr int NX = T::NX, NY = T::NY, PX = T::PX, NN = T::NN, NTH = TX * TY;
  __shared__ V xs[NN], ys[NN]; // [2][NY][PX]: x_phi / 8 of the tile's nodes (0 where constrained), the y tile
  const int tid = threadIdx.x, lx = tid % TX, ly = tid / TX;
  int bx, by, bz;
  decode_tile (g, (int) blockIdx.x, tiles_x, tiles_y, bx, by, bz);
  const int cx0 = bx * TX, cy0 = by * TY, cz0 = g.cell_begin + bz * g.layer_stride;
  const int nnx = g.nn[0], nny = g.nn[1];
  const long long pstride = g.nodes_per_plane;
  for (int i = tid; i < NN; i += NTH)
    {
      const int ix = i % PX, iy = (i / PX) % NY, iz = i / (PX * NY);
      const int gx = cx0 + ix, gy = cy0 + iy;
      V v = 0;
      if (ix < NX && gx < nnx && gy < nny && cz0 < g.cell_end)
        {
          const long long n = gx + (long long) nnx * gy + pstride * (cz0 + iz - g.plane_begin);
          if (!(mask[n] & 8))
            v = (V) 0.125 * x[4 * n + 3];
        }
      xs[i] = v;
      ys[i] = 0;
    }
  __syncthreads ();
  const bool valid = (cx0 + lx < g.n[0]) && (cy0 + ly < g.n[1]) && (cz0 < g.cell_end);
  V out[8];
#pragma unroll
  for (int v = 0; v < 8; ++v)
    out[v] = 0;
  if (valid)
    {
      // record of point q of this cell: thread / lane mapping of k_point_coeffs, c2 is the second half of a record
      const int thread = lx / W + (TX / W) * ly, lane = lx % W;
      const CS *rec = coef + ((size_t) ((cz0 - layer0) * tiles_y + by) * tiles_x + bx) * T::coef_per_tile
                      + (size_t) thread * 2 * W + W + lane;
      V f[8];
#pragma unroll
      for (int v = 0; v < 8; ++v)
        f[v] = xs[(lx + (v & 1)) + PX * (ly + ((v >> 1) & 1)) + PX * NY * (v >> 2)];
      const V Sq = (V) (NQ == 3 ? k.s : k.s2);
      const V es[3] = {-Sq, NQ == 3 ? (V) 0 : Sq, Sq};
      // ---- value term: interpolate z -> y -> x, weight with c2, transposed collapse x -> y -> z
      V sz[4], rz[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sz[j] = f[j + 4] + f[j], rz[j] = f[j + 4] - f[j];
#pragma unroll
      for (int qz = 0; qz < NQ; ++qz)
        {
          const V ez = es[qz];
          const V a0 = fma_r (ez, rz[0], sz[0]), a1 = fma_r (ez, rz[1], sz[1]);
          const V a2 = fma_r (ez, rz[2], sz[2]), a3 = fma_r (ez, rz[3], sz[3]);
          const V Py[2] = {a0 + a2, a1 + a3}, Ry[2] = {a2 - a0, a3 - a1};
          V Pl[2] = {0, 0}, Rl[2] = {0, 0};
#pragma unroll
          for (int qy = 0; qy < NQ; ++qy)
            {
              const V ey = es[qy];
              const V v0 = fma_r (ey, Ry[0], Py[0]), v1 = fma_r (ey, Ry[1], Py[1]);
              const V Pp = v0 + v1, Rp = v1 - v0;
              V AP = 0, AR = 0;
#pragma unroll
              for (int qx = 0; qx < NQ; ++qx)
                {
                  const V c2 = (V) v6_ldg (rec + (size_t) ((qz * NQ + qy) * NQ + qx) * T::NT * 2 * W);
                  const V wa = fma_r (es[qx], Rp, Pp) * c2;
                  AP += wa;
                  AR = fma_r (es[qx], wa, AR);
                }
              const V w0 = AP - AR, w1 = AP + AR; // x-nodes 0 and 1: weights (1 - xi_x), (1 + xi_x)
              Pl[0] += w0, Pl[1] += w1;
              Rl[0] = fma_r (ey, w0, Rl[0]), Rl[1] = fma_r (ey, w1, Rl[1]);
            }
          const V omez = (V) 1 - ez, opez = (V) 1 + ez;
#pragma unroll
          for (int vx = 0; vx < 2; ++vx)
            {
              const V t0 = Pl[vx] - Rl[vx], t1 = Pl[vx] + Rl[vx]; // y-nodes 0 and 1
              out[vx] = fma_r (omez, t0, out[vx]);
              out[vx + 2] = fma_r (omez, t1, out[vx + 2]);
              out[vx + 4] = fma_r (opez, t0, out[vx + 4]);
              out[vx + 6] = fma_r (opez, t1, out[vx + 6]);
            }
        }
      // ---- G_c eps grad psi_i . grad psi_j in closed form: the Q1 stiffness of a cube is diagonal in the basis of
      // sums and differences per direction, eigenvalue G_c eps h {0, 1/2, 1/3, 1/6} for {0, 1, 2, 3} differences
      V c[8];
#pragma unroll
      for (int v = 0; v < 8; ++v)
        c[v] = f[v];
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int v = 0; v < 8; ++v)
          if (!(v & (1 << d)))
            {
              const V a = c[v], b = c[v | (1 << d)];
              c[v] = a + b, c[v | (1 << d)] = b - a;
            }
      const V kl[4] = {(V) 0, (V) k.kl[0], (V) k.kl[1], (V) k.kl[2]};
#pragma unroll
      for (int v = 0; v < 8; ++v)
        c[v] *= kl[(v & 1) + ((v >> 1) & 1) + (v >> 2)];
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int v = 0; v < 8; ++v)
          if (!(v & (1 << d)))
            {
              const V s = c[v], r = c[v | (1 << d)];
              c[v] = s - r, c[v | (1 << d)] = s + r;
            }
#pragma unroll
      for (int v = 0; v < 8; ++v)
        out[v] += c[v];
    }
  // cells that share a node are ordered by eight barrier-separated vertex phases
#pragma unroll
  for (int v = 0; v < 8; ++v)
    {
      if (valid)
        ys[(lx + (v & 1)) + PX * (ly + ((v >> 1) & 1)) + PX * NY * (v >> 2)] += out[v];
      __syncthreads ();
    }
  for (int i = tid; i < NN; i += NTH)
    {
      const int ix = i % PX, iy = (i / PX) % NY, iz = i / (PX * NY);
      const int gx = cx0 + ix, gy = cy0 + iy;
      if (ix < NX && gx < nnx && gy < nny && cz0 < g.cell_end)
        {
          const long long n = gx + (long long) nnx * gy + pstride * (cz0 + iz - g.plane_begin);
          if (!(mask[n] & 8))
            atomicAdd (&y[4 * n + 3], ys[i]);
        }
    }
}

} // namespace pf
