// fracture_problem.h -- host driver mirroring FracturePhaseFieldProblem<dim>
// (cracks.cc:1024-1170) for the part of the program that sits on the hot path:
// parameter handling, the time-step loop of run() (4166-4581) and the
// primal-dual active-set Newton iteration (2780-2994).  All vectors live on
// the GPU behind the C ABI; this class only sequences calls and keeps the
// reference's screen / statistics output format.
//
// Supported here, `outer solver = active set`:
//   `test case = sneddon` on the uniform (globally refined) box, dim 2 or 3;
//   `test case = miehe tension / miehe shear` (dim 2) on the globally refined
//   meshes/unit_slit.inp topology, with the stress split and the load
//   functional, for as long as refine_mesh() would not change the mesh;
//   on the host forest (forest.h; device side verified on the CPU emulation,
//   first GPU run pending): `sneddon`, dim 2, with local pre-refinement and
//   refinement cycles (`fixed preref sneddon`), and `multiple het`, dim 3,
//   with the phase-field pre-refinement and the bitmap E-modulus field.
// Everything else the .prm surface can express is parsed and rejected with
// ExcNotImplemented-style errors (SURVEY.md section 8f lists it as "next").
#pragma once
#include <iosfwd>
#include <string>
#include <vector>

#include <memory>

#include "forest.h"
#include "function_parser.h"
#include "matrix_free_jacobian.h"
#include "parameter_handler.h"

namespace cracks {

struct NotImplemented : std::runtime_error
{
  using std::runtime_error::runtime_error;
};

struct StatisticsRow
{
  unsigned timestep_no;
  double time;
  long long dofs;
  double h_min, bulk_energy, crack_energy;
  double load = 0; // "Load y" (miehe tension) / "Load x" (miehe shear), cracks.cc:3791-3804
};

class FracturePhaseFieldProblem
{
public:
  FracturePhaseFieldProblem (ParameterHandler &prm, int dim, std::ostream &out);
  ~FracturePhaseFieldProblem ();
  static void declare_parameters (ParameterHandler &prm);
  void run ();

  const std::vector<StatisticsRow> &statistics () const { return statistics_; }
  double tcv () const { return tcv_; }
  const std::vector<std::pair<double, double>> &cod () const { return cod_; }
  unsigned total_newton_iterations () const { return total_newton_its_; }
  unsigned total_linear_iterations () const { return total_linear_its_; }
  // knobs that are not part of the reference's .prm surface
  int device = 0;
  // run the Miehe tests with `Adaptive refinement cycles` > 0 on the host forest and follow
  // refine_mesh() (phase-field flags, level cap, solution transfer, redo of the step) instead of stopping
  // where the mesh would change
  bool adaptive_forest = false;
  // write the .vtu / .pvtu / .visit / .pvd files of output_results() (cracks.cc:3142-3258); the command line
  // switches it off with --no-output (16.7 M DoF are 0.8 GB per time step)
  bool write_output = true;
  std::string source_dir = ".";   // where test.pgm lives ($SRC of cracks.cc:1541, a compile-time path in the reference)
  int gmres_max_iterations = 200; // SolverControl(200, ...) at cracks.cc:2762
  double gmres_tolerance = 1e-8;

private:
  void set_runtime_parameters ();
  void setup_system ();
  void determine_mesh_dependent_parameters ();
  double newton_active_set ();
  void write_statistics () const;
  void output_results (); // cracks.cc:3142-3258
  bool miehe () const { return test_case == "miehe tension" || test_case == "miehe shear"; }
  // forest path (GPU suite: tests/test_gpu_forest.py): Sneddon 2-D with local pre-refinement
  // / refinement cycles on the host forest, strategy `fixed preref sneddon`
  bool use_forest () const { return forest_ != nullptr; }
  bool hetero () const { return test_case == "multiple het"; }
  void forest_refine_fixed_preref_sneddon ();
  void forest_refine_phase_field_on_initial_values ();
  void forest_prerefine ();
  void forest_create_context ();
  std::vector<double> forest_initial_sneddon () const;
  std::vector<double> forest_initial_values () const;
  bool forest_refine_phase_field_and_transfer ();
  std::vector<double> forest_miehe_boundary_values (double t) const;
  long long n_nodes () const;
  long long n_cells () const;
  int miehe_kind () const { return test_case == "miehe tension" ? 1 : 2; }

  ParameterHandler &prm_;
  int dim_;
  std::ostream &pcout_;
  pf_ctx *ctx_ = nullptr;
  std::unique_ptr<Forest> forest_;
  std::vector<long long> forest_top_nodes_;  // nodes on boundary id 3 (y = 1) of the Miehe forest
  std::vector<int64_t> forest_top_cells_;    // cells whose top edge lies on it (compute_load)
  pf_mesh mesh_{};
  pf_params params_{};

  // run-time parameters, names as in the reference (cracks.cc:1111-1166)
  unsigned n_global_pre_refine = 0, n_local_pre_refine = 0, n_refinement_cycles = 0;
  unsigned max_no_timesteps = 0, switch_timestep = 0;
  double timestep = 1, timestep_size_2 = 1, time = 0, old_timestep = 1, old_old_timestep = 1;
  unsigned timestep_number = 0;
  std::string outer_solver, test_case, output_folder;
  double G_c = 0, poisson_ratio_nu = 0, E_modulus = 0, lame_coefficient_mu = 0, lame_coefficient_lambda = 0;
  double constant_k = 0, alpha_eps = 0, min_cell_diameter = 0;
  bool direct_solver = false, use_old_timestep_pf = false;
  double lower_bound_newton_residual = 1e-10, line_search_damping = 0.5;
  unsigned max_no_newton_steps = 10, max_no_line_search_steps = 5;
  double decompose_stress_rhs = 0, decompose_stress_matrix = 0;
  double value_phase_field_for_refinement = 0;
  std::string refinement_strategy;
  FunctionParser func_pressure;

  std::vector<StatisticsRow> statistics_;
  int output_cycle_ = -1; // `refinement_cycle` of output_results()
  std::string filename_base;
  std::vector<std::vector<std::string>> output_file_names_by_timestep_;
  std::vector<std::pair<double, std::string>> times_and_names_;
  double tcv_ = 0;
  double active_set_E_ = 1.0; // `multiple het`: E_modulus as the last assembled cell leaves it (cracks.cc:2209-2210, 2859)
  std::vector<std::pair<double, double>> cod_; // (x, COD(x)) lines of compute_functional_values
  unsigned total_newton_its_ = 0, total_linear_its_ = 0;
};

} // namespace cracks
