// emu_tiled.cc -- TEST INFRASTRUCTURE: the tiled 3-D kernels of the hot path (k_apply3d_v4 / v2 and
// k_residual3d: shared-memory staging, block barriers, atomics) executed on the CPU with one OS thread
// per CUDA thread (tests/emu/cuda_shim_block), so that the hot kernel's SOURCE is held against the
// oracle in the CPU suite as well (tests/test_kernel_emulation_cpu.py).  Not a CPU fallback of the
// product; nothing under cracks_b200/ links it.
#include <cuda_runtime.h>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
std::barrier<> *emu_block_barrier = nullptr;
namespace pf {
// `extern __shared__ unsigned char smem_raw[]` inside the kernels (namespace pf) resolves to this buffer;
// blocks run one after the other, so one buffer serves them all
alignas (16) unsigned char smem_raw[256 * 1024];
} // namespace pf

#include <thread>
#include <vector>

#include "../../cracks_b200/csrc/pf_apply3d_v4.cuh"
#include "../../cracks_b200/csrc/pf_apply3d_v5.cuh"
#include "../../cracks_b200/csrc/pf_apply3d_v6.cuh"
#include "../../cracks_b200/csrc/pf_residual3d.cuh"
#include "../../cracks_b200/csrc/pf_vector.cuh"

using namespace pf;

template <class K, class... A>
static void
launch_blocks (K kernel, unsigned grid, unsigned block, A... args)
{
  for (unsigned b = 0; b < grid; ++b)
    {
      std::barrier<> bar ((std::ptrdiff_t) block);
      emu_block_barrier = &bar;
      std::vector<std::thread> threads;
      for (unsigned t = 0; t < block; ++t)
        threads.emplace_back ([=]() {
          gridDim.x = grid;
          blockDim.x = block;
          blockIdx.x = b;
          threadIdx.x = t;
          kernel (args...);
        });
      for (auto &th : threads)
        th.join ();
    }
}

static Grid
box_grid (const int *n, const double *h)
{
  Grid g{};
  g.dim = 3;
  g.nodes_per_plane = 1;
  g.n_global_nodes = 1;
  for (int d = 0; d < 3; ++d)
    {
      g.n[d] = n[d];
      g.nn[d] = n[d] + 1;
      g.h[d] = h[d];
      g.n_global_nodes *= g.nn[d];
      if (d < 2)
        g.nodes_per_plane *= g.nn[d];
    }
  g.plane_begin = g.owned_begin = 0;
  g.plane_end = g.owned_end = g.nn[2];
  g.cell_begin = 0;
  g.cell_end = n[2];
  g.n_local_nodes = g.n_global_nodes;
  g.n_local_cells = (long long) n[0] * n[1] * n[2];
  g.slit_row = -1;
  return g;
}

static K3
make_k3 (const double *h)
{
  K3 k;
  k.s = std::sqrt (3.0 / 5.0);
  k.s2 = std::sqrt (1.0 / 3.0);
  k.wvol = 1.0;
  for (int d = 0; d < 3; ++d)
    {
      k.gu[d] = 1.0 / (4.0 * h[d]);
      k.gp[d] = 2.0 / h[d];
      k.ih[d] = 1.0 / h[d];
      k.wvol *= h[d] / 2.0;
    }
  k.wq[0] = k.wq[2] = 5.0 / 9.0;
  k.wq[1] = 8.0 / 9.0;
  return k;
}

extern "C" {

// y = J x on a box mesh with the tiled kernel, the way apply_dev launches it (k_apply_init, then the
// tiles).  variant 16 = v4, 3 = v2; nq = 3 (exact) or 2 (multigrid smoother operator); vectors are
// node-major interleaved with 4 doubles per node.
void
emu_apply3d (int variant, int nq, const int *n, const double *h, const double *phys, const double *x, const double *sol,
             const double *pt, const unsigned char *mask, const double *diag, double *y)
{
  const Grid g = box_grid (n, h);
  const K3 k = make_k3 (h);
  Phys p;
  std::memset (&p, 0, sizeof p);
  p.lambda = phys[0], p.mu = phys[1], p.G_c = phys[2], p.kappa = phys[3], p.eps = phys[4], p.P1 = phys[5];
  p.clamp_extra = (int) phys[6];
  const long long nn = g.n_local_nodes;
  for (long long i = 0; i < nn; ++i) // k_apply_init<3>
    for (int c = 0; c < 4; ++c)
      y[4 * i + c] = ((mask[i] >> c) & 1) ? diag[4 * i + c] * x[4 * i + c] : 0.0;
  constexpr int TX = 16, TY = 4, TZ = 1;
  const int tiles_x = (n[0] + TX - 1) / TX, tiles_y = (n[1] + TY - 1) / TY, tiles_z = (n[2] + TZ - 1) / TZ;
  const unsigned grid = (unsigned) (tiles_x * tiles_y * tiles_z);
  const bool iso = h[0] == h[1] && h[1] == h[2];
  if (variant == 26 || variant == 27 || variant == 28)
    {
      // v6: cubic cells only; the coefficient records first (k_point_coeffs), then the tiles -- as pf_setup_jacobian /
      // launch_apply3d_v6 do.  27 = the FP32 Jacobian of the inexact-Newton path.
      using T6 = Tile3v6<16, 4>;
      K6 k6;
      const double gam = k.gu[0], omk = 1.0 - p.kappa;
      k6.s = k.s, k6.gam = gam, k6.lam2 = p.lambda / (2.0 * p.mu);
      k6.beta = p.P1 * gam / (omk * 2.0 * p.mu * gam * gam);
      k6.k1 = 0.5 * omk * p.mu * gam * gam;
      k6.w[0] = 25.0 / 81.0, k6.w[1] = 40.0 / 81.0, k6.w[2] = 64.0 / 81.0;
      for (int q = 0; q < 3; ++q)
        k6.wz[q] = k.wvol * k.wq[q];
      k6.kl[0] = p.G_c * p.eps * h[0] * 0.5, k6.kl[1] = p.G_c * p.eps * h[0] / 3.0, k6.kl[2] = p.G_c * p.eps * h[0] / 6.0;
      k6.s2 = k.s2, k6.wvol = k.wvol, k6.cge = p.G_c * p.eps * 8.0 * gam * gam;
      if (variant == 26 && nq == 3)
        {
          std::vector<double> coef (T6::coef_per_tile * grid);
          launch_blocks (k_point_coeffs<double, 16, 4, 3, 1>, grid, 64u, g, p, k, tiles_x, tiles_y, 0, sol, pt, coef.data ());
          launch_blocks (k_apply3d_v6<double, double, 3, 16, 4, 4>, grid, 64u, g, k6, tiles_x, tiles_y, 0, x, sol, mask,
                         (const double *) coef.data (), y);
        }
      else if (variant == 28)
        {
          // the 2-point-rule smoother operator without its (phi,u) block (pf_set_multigrid_coupling 0)
          std::vector<double> coef (Tile3v6<16, 4, 2>::coef_per_tile * grid);
          launch_blocks (k_point_coeffs<double, 16, 4, 2, 1>, grid, 64u, g, p, k, tiles_x, tiles_y, 0, sol, pt, coef.data ());
          launch_blocks (k_apply3d_v6<double, double, 2, 16, 4, 4, false>, grid, 64u, g, k6, tiles_x, tiles_y, 0, x, sol, mask,
                         (const double *) coef.data (), y);
        }
      else if (variant == 26)
        {
          std::vector<double> coef (Tile3v6<16, 4, 2>::coef_per_tile * grid);
          launch_blocks (k_point_coeffs<double, 16, 4, 2, 1>, grid, 64u, g, p, k, tiles_x, tiles_y, 0, sol, pt, coef.data ());
          launch_blocks (k_apply3d_v6<double, double, 2, 16, 4, 4>, grid, 64u, g, k6, tiles_x, tiles_y, 0, x, sol, mask,
                         (const double *) coef.data (), y);
        }
      else
        {
          // the FP32 Jacobian: two cells per thread in packed arithmetic on 32 x 4 tiles (pf_api.cu: V6Shape<f32x2>)
          const int fx = (n[0] + 31) / 32, fy = (n[1] + 3) / 4;
          const unsigned fgrid = (unsigned) (fx * fy * n[2]);
          if (nq == 3)
            {
              std::vector<float> coef (Tile3v6<32, 4, 3, 2>::coef_per_tile * fgrid);
              launch_blocks (k_point_coeffs<float, 32, 4, 3, 2>, fgrid, 128u, g, p, k, fx, fy, 0, sol, pt, coef.data ());
              launch_blocks (k_apply3d_v6<f32x2, double, 3, 32, 4, 4>, fgrid, 64u, g, k6, fx, fy, 0, x, sol, mask,
                             (const float *) coef.data (), y);
            }
          else
            {
              std::vector<float> coef (Tile3v6<32, 4, 2, 2>::coef_per_tile * fgrid);
              launch_blocks (k_point_coeffs<float, 32, 4, 2, 2>, fgrid, 128u, g, p, k, fx, fy, 0, sol, pt, coef.data ());
              launch_blocks (k_apply3d_v6<f32x2, double, 2, 32, 4, 4>, fgrid, 64u, g, k6, fx, fy, 0, x, sol, mask,
                             (const float *) coef.data (), y);
            }
        }
      return;
    }
#define EMU_LAUNCH(KERNEL) launch_blocks (KERNEL, grid, (unsigned) (TX * TY * TZ), g, p, k, tiles_x, tiles_y, x, sol, pt, mask, y)
  if (variant == 19)
    iso ? EMU_LAUNCH ((k_apply3d_v5<TX, TY, TZ, 2, 3, true>) ) : EMU_LAUNCH ((k_apply3d_v5<TX, TY, TZ, 2, 3, false>) );
  else if (variant == 16 && nq == 3)
    iso ? EMU_LAUNCH ((k_apply3d_v4<TX, TY, TZ, 2, 3, true>) ) : EMU_LAUNCH ((k_apply3d_v4<TX, TY, TZ, 2, 3, false>) );
  else if (variant == 16)
    iso ? EMU_LAUNCH ((k_apply3d_v4<TX, TY, TZ, 2, 2, true>) ) : EMU_LAUNCH ((k_apply3d_v4<TX, TY, TZ, 2, 2, false>) );
  else if (nq == 3)
    iso ? EMU_LAUNCH ((k_apply3d_v2<TX, TY, TZ, 2, 3, true>) ) : EMU_LAUNCH ((k_apply3d_v2<TX, TY, TZ, 2, 3, false>) );
  else
    iso ? EMU_LAUNCH ((k_apply3d_v2<TX, TY, TZ, 2, 2, true>) ) : EMU_LAUNCH ((k_apply3d_v2<TX, TY, TZ, 2, 2, false>) );
#undef EMU_LAUNCH
}

// r_total of residual_dev (memset + k_residual3d)
void
emu_residual3d (const int *n, const double *h, const double *phys, const double *sol, const double *pt, double *r)
{
  const Grid g = box_grid (n, h);
  const K3 k = make_k3 (h);
  Phys p;
  std::memset (&p, 0, sizeof p);
  p.lambda = phys[0], p.mu = phys[1], p.G_c = phys[2], p.kappa = phys[3], p.eps = phys[4], p.P1 = phys[5];
  p.clamp_extra = (int) phys[6];
  std::memset (r, 0, sizeof (double) * 4 * (size_t) g.n_local_nodes);
  const int tiles_x = (n[0] + 15) / 16, tiles_y = (n[1] + 3) / 4, tiles_z = n[2];
  launch_blocks (k_residual3d<16, 4, 1, 2>, (unsigned) (tiles_x * tiles_y * tiles_z), 64u, g, p, k, tiles_x, tiles_y, sol, pt, r);
}
}
