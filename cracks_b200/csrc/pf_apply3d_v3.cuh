// pf_apply3d_v3.cuh -- third generation of the hot kernel: v2's arithmetic
// (tile_cells_v2) inside a persistent CTA that receives the nodal halo of its
// next tile by TMA while the current tile is being computed.
//
// ncu on v2 (profiles/r1_v2_apply3d_ncu_summary.txt): the plane loop already
// keeps the FP64 pipe busy, but 40 % of the samples sit outside it -- stage-1
// global-load latency once per tile, block barriers, CTA prologue.  v3:
//   * grid = #SMs x MINB persistent CTAs, tiles assigned round-robin;
//   * one elected thread issues three cp.async.bulk.tensor.4d loads per tile
//     (x, U as [z][y][x][4] doubles; aux = {phi~, mask} as [z][y][x][2]) into
//     a raw shared buffer, completion tracked by an mbarrier (complete_tx);
//     out-of-range parts of ragged tiles are zero-filled by the TMA unit;
//   * the raw buffer is consumed by stage 1 (z-collapse) and is free again
//     after one barrier, so the loads of tile i+1 are issued right there and
//     land during stage 3 of tile i: no global-load latency on the critical
//     path, single raw buffer.
#pragma once
#include <cuda.h>

#include "pf_apply3d_v2.cuh"

namespace pf {

template <int TX, int TY, int TZ> struct Tile3v3
{
  using V2 = Tile3v2<TX, TY, TZ>;
  static constexpr int NN = V2::NN;
  static constexpr size_t raw_x_bytes = (size_t) NN * 4 * sizeof (double);
  static constexpr size_t raw_a_bytes = (size_t) NN * 2 * sizeof (double);
  static constexpr size_t align128 (size_t v) { return (v + 127) / 128 * 128; }
  static constexpr size_t off_mbar = 0;
  static constexpr size_t off_raw_x = 128;
  static constexpr size_t off_raw_s = off_raw_x + align128 (raw_x_bytes);
  static constexpr size_t off_raw_a = off_raw_s + align128 (raw_x_bytes);
  static constexpr size_t off_work = off_raw_a + align128 (raw_a_bytes);
  static constexpr size_t off_ms = off_work + V2::smem_bytes;
  static constexpr size_t smem_bytes = off_ms + align128 (NN);
  static constexpr unsigned tx_bytes = (unsigned) (2 * raw_x_bytes + raw_a_bytes);
};

__device__ __forceinline__ unsigned
smem_u32 (const void *p)
{
  return (unsigned) __cvta_generic_to_shared (p);
}

__device__ __forceinline__ void
tma_load_4d (void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, void *mbar)
{
  asm volatile ("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
                " [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32 (dst)),
                "l"(reinterpret_cast<unsigned long long> (tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
                "r"(smem_u32 (mbar))
                : "memory");
}

__device__ __forceinline__ void
mbar_wait (void *mbar, unsigned parity)
{
  asm volatile ("{\n"
                ".reg .pred p;\n"
                "WAIT_%=:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra DONE_%=;\n"
                "bra WAIT_%=;\n"
                "DONE_%=:\n"
                "}" ::"r"(smem_u32 (mbar)),
                "r"(parity)
                : "memory");
}

template <int TX, int TY, int TZ, int MINB>
__global__ void __launch_bounds__ (TX * TY * TZ, MINB)
k_apply3d_v3 (Grid g, Phys p, K3 k, int tiles_x, int tiles_y, int n_tiles,
              unsigned long long *__restrict__ tile_counter, unsigned long long epoch_base,
              const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_s,
              const __grid_constant__ CUtensorMap tm_a, double *__restrict__ y)
{
  using T = Tile3v2<TX, TY, TZ>;
  using L = Tile3v3<TX, TY, TZ>;
  constexpr int NN = T::NN, NT = T::NT, NX = T::NX, NY = T::NY, NC2 = T::NC2, NXC = T::NXC;
  extern __shared__ __align__ (128) unsigned char smem_raw[];
  unsigned long long *mbar = reinterpret_cast<unsigned long long *> (smem_raw + L::off_mbar);
  const double4 *raw_x = reinterpret_cast<const double4 *> (smem_raw + L::off_raw_x);
  const double4 *raw_s = reinterpret_cast<const double4 *> (smem_raw + L::off_raw_s);
  const double2 *raw_a = reinterpret_cast<const double2 *> (smem_raw + L::off_raw_a);
  double *AZ = reinterpret_cast<double *> (smem_raw + L::off_work); // [3][9][NC2]
  double *BZ = AZ + 27 * NC2;                                       // [7][3][NXC]
  double *DZ = BZ + 21 * NXC;                                       // [7][NC2]
  double *ys = DZ;                                                  // [4][NN], aliases DZ
  uint8_t *ms = smem_raw + L::off_ms;                               // [NN]

  const int tid = threadIdx.x;
  const int nnx = g.nn[0], nny = g.nn[1];
  const int lz_off = g.plane_begin;
  const long long pstride = g.nodes_per_plane;
  const double S = k.s;

  auto issue = [&](int t) {
    int b = t;
    const int bx = b % tiles_x;
    b /= tiles_x;
    const int by = b % tiles_y;
    const int bz = b / tiles_y;
    const int cx = bx * TX, cy = by * TY, lz = g.cell_begin + bz * TZ - lz_off;
    asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32 (mbar)),
                  "r"(L::tx_bytes)
                  : "memory");
    tma_load_4d (smem_raw + L::off_raw_x, &tm_x, 0, cx, cy, lz, mbar);
    tma_load_4d (smem_raw + L::off_raw_s, &tm_s, 0, cx, cy, lz, mbar);
    tma_load_4d (smem_raw + L::off_raw_a, &tm_a, 0, cx, cy, lz, mbar);
  };

  // dynamic tile scheduler: tiles are handed out in order by a global counter
  // that is never reset (the host advances epoch_base by n_tiles + gridDim.x
  // per launch: every CTA overshoots exactly once).  next_tile[] is the
  // two-entry mailbox thread 0 uses to publish the tile it prefetched.
  int *next_tile = reinterpret_cast<int *> (smem_raw + L::off_mbar + 16);
  auto grab = [&]() -> int {
    const unsigned long long v = atomicAdd (tile_counter, 1ull) - epoch_base;
    return v < (unsigned long long) n_tiles ? (int) v : n_tiles;
  };
  if (tid == 0)
    {
      asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32 (mbar)) : "memory");
      asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
      const int t0 = grab ();
      next_tile[0] = t0;
      if (t0 < n_tiles)
        issue (t0);
    }
  __syncthreads ();

  unsigned parity = 0;
  int slot = 0;
  for (int t = next_tile[0]; t < n_tiles; t = next_tile[slot])
    {
      int b = t;
      const int bx = b % tiles_x;
      b /= tiles_x;
      const int by = b % tiles_y;
      const int bz = b / tiles_y;
      const int cx0 = bx * TX, cy0 = by * TY, cz0 = g.cell_begin + bz * TZ;

      mbar_wait (mbar, parity);
      parity ^= 1u;

      // ---- stage 1: z-collapse per node column, from the TMA-filled raw buffer
      for (int i = tid; i < NC2; i += NT)
        {
          const int n0 = i % (NX * NY) + (NX * NY) * (i / (NX * NY)); // node (ix, iy, tz)
          const int n1 = n0 + NX * NY;
          const double4 xa = raw_x[n0], xb = raw_x[n1], sa = raw_s[n0], sb = raw_s[n1];
          const double2 aa = raw_a[n0], ab = raw_a[n1];
          const unsigned m0 = (unsigned) __double_as_longlong (aa.y), m1 = (unsigned) __double_as_longlong (ab.y);
          double f0[9], f1[9];
          f0[0] = (m0 & 1) ? 0.0 : xa.x;
          f0[1] = (m0 & 2) ? 0.0 : xa.y;
          f0[2] = (m0 & 4) ? 0.0 : xa.z;
          f0[3] = (m0 & 8) ? 0.0 : 0.125 * xa.w;
          f1[0] = (m1 & 1) ? 0.0 : xb.x;
          f1[1] = (m1 & 2) ? 0.0 : xb.y;
          f1[2] = (m1 & 4) ? 0.0 : xb.z;
          f1[3] = (m1 & 8) ? 0.0 : 0.125 * xb.w;
          f0[4] = sa.x, f0[5] = sa.y, f0[6] = sa.z, f0[7] = 0.125 * sa.w, f0[8] = 0.125 * aa.x;
          f1[4] = sb.x, f1[5] = sb.y, f1[6] = sb.z, f1[7] = 0.125 * sb.w, f1[8] = 0.125 * ab.x;
#pragma unroll
          for (int f = 0; f < 9; ++f)
            {
              const double s = f0[f] + f1[f], r = f1[f] - f0[f];
              AZ[(0 * 9 + f) * NC2 + i] = fma (-S, r, s);
              AZ[(1 * 9 + f) * NC2 + i] = s;
              AZ[(2 * 9 + f) * NC2 + i] = fma (S, r, s);
              if (f < 7)
                DZ[f * NC2 + i] = r * ((f == 3) ? k.gp[2] : k.gu[2]);
            }
        }
      for (int i = tid; i < NN; i += NT)
        ms[i] = (uint8_t) __double_as_longlong (raw_a[i].y);
      __syncthreads ();
      // the raw buffer is free: prefetch the next tile of this CTA
      slot ^= 1;
      if (tid == 0)
        {
          const int tn = grab ();
          next_tile[slot] = tn; // read by everybody after the barriers below
          if (tn < n_tiles)
            {
              asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
              issue (tn);
            }
        }
      // ---- stage 2: y-collapse of the z-derivative chain
      for (int i = tid; i < NXC; i += NT)
        {
          const int ix = i % NX, cy = (i / NX) % TY, tz = i / (NX * TY);
          const int c0 = ix + NX * (cy + NY * tz);
#pragma unroll
          for (int f = 0; f < 7; ++f)
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

      // ---- stages 3 + 4
      tile_cells_v2<TX, TY, TZ> (g, p, k, tid, cx0, cy0, cz0, AZ, BZ, ys);

      // ---- flush the y tile
      for (int i = tid; i < NN; i += NT)
        {
          const int ix = i % NX, iy = (i / NX) % NY, iz = i / (NX * NY);
          const int gx = cx0 + ix, gy = cy0 + iy, gz = cz0 + iz;
          if (gx < nnx && gy < nny && gz <= g.cell_end)
            {
              const long long n = gx + (long long) nnx * gy + pstride * (gz - lz_off);
              const unsigned m = ms[i];
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (!((m >> c) & 1))
                  atomicAdd (&y[4 * n + c], ys[c * NN + i]);
            }
        }
      __syncthreads (); // ys / ms / AZ are rewritten by the next tile
    }
}

// {phi~, mask} packed as 16-byte records so that the pair is a legal TMA tensor
// (global strides must be multiples of 16 bytes)
__global__ void
k_pack_aux (long long n_nodes, const double *__restrict__ pt, const uint8_t *__restrict__ mask,
            double2 *__restrict__ aux)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n < n_nodes)
    aux[n] = make_double2 (pt[n], __longlong_as_double ((long long) mask[n]));
}

} // namespace pf
