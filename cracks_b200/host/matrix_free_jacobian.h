// matrix_free_jacobian.h -- C++ face of the C ABI.
//
// The reference hands `system_pde_matrix` to SolverGMRES, which only needs a
// duck-typed `vmult(dst, src)` (cracks.cc:2764-2771; its own
// BlockDiagonalPreconditioner, 2717-2740, is such a class).  MatrixFreeJacobian
// is the drop-in for that object: same call, GPU underneath.  Error codes of
// the C ABI are mapped back to exceptions here, so that
// SolverControl::NoConvergence stays control flow for the time-step cut
// (cracks.cc:4333-4355).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cracks_b200.h"

namespace cracks {

struct NoConvergence : std::runtime_error // SolverControl::NoConvergence
{
  using std::runtime_error::runtime_error;
};

struct DeviceError : std::runtime_error
{
  using std::runtime_error::runtime_error;
};

inline void
pf_check (pf_ctx *ctx, int rc)
{
  if (rc == PF_OK)
    return;
  const std::string msg = ctx ? pf_last_error (ctx) : "cracks_b200 call failed";
  if (rc == PF_NO_CONVERGENCE)
    throw NoConvergence (msg);
  throw DeviceError ("cracks_b200 error " + std::to_string (rc) + ": " + msg);
}

// Block vector in the reference's layout: block(0) = displacements
// (dim per node), block(1) = phase field.
struct BlockVector
{
  std::vector<double> data;
  size_t n_u = 0;
  BlockVector () = default;
  BlockVector (size_t n_nodes, int dim) : data (n_nodes * (dim + 1), 0.0), n_u (n_nodes * dim) {}
  double *block (int b) { return data.data () + (b == 0 ? 0 : n_u); }
  const double *block (int b) const { return data.data () + (b == 0 ? 0 : n_u); }
  size_t size () const { return data.size (); }
  double l2_norm () const;
};

class MatrixFreeJacobian
{
public:
  explicit MatrixFreeJacobian (pf_ctx *ctx) : ctx_ (ctx) {}
  // linearise at the state / constraints currently set on the context
  void reinit () { pf_check (ctx_, pf_setup_jacobian (ctx_)); }
  // dst = J(U) src, exactly the call SolverGMRES makes on system_pde_matrix
  void vmult (BlockVector &dst, const BlockVector &src) const
  {
    pf_check (ctx_, pf_apply_jacobian (ctx_, src.data.data (), dst.data.data ()));
  }
  size_t m () const { return (size_t) pf_n_dofs (ctx_); }

private:
  pf_ctx *ctx_;
};

} // namespace cracks
