#include "fracture_problem.h"
#include "vtu_writer.h"

#include "bitmap_function.h"

#include <cmath>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <sys/stat.h>

namespace cracks {

double
BlockVector::l2_norm () const
{
  double s = 0;
  for (double v : data)
    s += v * v;
  return std::sqrt (s);
}

using P = ParameterHandler::Pattern;

// the 32 entries of cracks.cc:1307-1405, same names, defaults and patterns
void
FracturePhaseFieldProblem::declare_parameters (ParameterHandler &prm)
{
  prm.enter_subsection ("Global parameters");
  prm.declare_entry ("Dimension", "2", P::Integer, "", 0);
  prm.declare_entry ("FE degree", "1", P::Integer, "", 1);
  prm.declare_entry ("Global pre-refinement steps", "1", P::Integer, "", 0);
  prm.declare_entry ("Local pre-refinement steps", "0", P::Integer, "", 0);
  prm.declare_entry ("Adaptive refinement cycles", "0", P::Integer, "", 0);
  prm.declare_entry ("Max No of timesteps", "1", P::Integer, "", 0);
  prm.declare_entry ("Timestep size", "1.0", P::Double, "", 0);
  prm.declare_entry ("Timestep size to switch to", "1.0", P::Double, "", 0);
  prm.declare_entry ("Switch timestep after steps", "0", P::Integer, "", 0);
  prm.declare_entry ("outer solver", "active set", P::Selection, "active set|simple monolithic");
  prm.declare_entry ("test case", "sneddon", P::Selection,
                     "sneddon|miehe tension|miehe shear|multiple homo|multiple het|three point bending");
  prm.declare_entry ("ref strategy", "phase field", P::Selection,
                     "phase field|fixed preref sneddon|fixed preref miehe tension|fixed preref miehe shear|"
                     "fixed preref multiple homo|fixed preref multiple het|global|mix|phase field three point top");
  prm.declare_entry ("value phase field for refinement", "0.0", P::Double, "", 0);
  prm.declare_entry ("Output directory", "output", P::Anything);
  prm.declare_entry ("Output filename", "solution_", P::Anything);
  prm.leave_subsection ();

  prm.enter_subsection ("Problem dependent parameters");
  prm.declare_entry ("K reg", "1.0 * h", P::Anything);
  prm.declare_entry ("Eps reg", "1.0 * h", P::Anything);
  prm.declare_entry ("Gamma penalization", "0.0", P::Double, "", 0);
  prm.declare_entry ("Pressure", "0.0", P::Anything);
  prm.declare_entry ("Fracture toughness G_c", "0.0", P::Double, "", 0);
  prm.declare_entry ("Poisson ratio nu", "0.0", P::Double, "", 0);
  prm.declare_entry ("E modulus", "0.0", P::Double, "", 0);
  prm.declare_entry ("Lame mu", "0.0", P::Double, "", 0);
  prm.declare_entry ("Lame lambda", "0.0", P::Double, "", 0);
  prm.leave_subsection ();

  prm.enter_subsection ("Solver parameters");
  prm.declare_entry ("Use Direct Inner Solver", "false", P::Bool);
  prm.declare_entry ("Newton lower bound", "1.0e-10", P::Double, "", 0);
  prm.declare_entry ("Newton maximum steps", "10", P::Integer, "", 0);
  prm.declare_entry ("Upper Newton rho", "0.999", P::Double, "", 0);
  prm.declare_entry ("Line search maximum steps", "5", P::Integer, "", 0);
  prm.declare_entry ("Line search damping", "0.5", P::Double, "", 0);
  prm.declare_entry ("Decompose stress in rhs", "0.0", P::Double, "", 0);
  prm.declare_entry ("Decompose stress in matrix", "0.0", P::Double, "", 0);
  prm.leave_subsection ();
}

FracturePhaseFieldProblem::FracturePhaseFieldProblem (ParameterHandler &prm, int dim, std::ostream &out)
  : prm_ (prm), dim_ (dim), pcout_ (out)
{
  if (dim != 2 && dim != 3)
    throw NotImplemented ("Dimension must be 2 or 3");
}

FracturePhaseFieldProblem::~FracturePhaseFieldProblem ()
{
  if (ctx_)
    pf_destroy (ctx_);
}

// cracks.cc:1411-1575
void
FracturePhaseFieldProblem::set_runtime_parameters ()
{
  prm_.enter_subsection ("Global parameters");
  if (prm_.get_integer ("FE degree") != 1)
    throw NotImplemented ("only FE degree = 1 is available on the GPU path");
  n_global_pre_refine = (unsigned) prm_.get_integer ("Global pre-refinement steps");
  n_local_pre_refine = (unsigned) prm_.get_integer ("Local pre-refinement steps");
  n_refinement_cycles = (unsigned) prm_.get_integer ("Adaptive refinement cycles");
  max_no_timesteps = (unsigned) prm_.get_integer ("Max No of timesteps");
  timestep = prm_.get_double ("Timestep size");
  timestep_size_2 = prm_.get_double ("Timestep size to switch to");
  switch_timestep = (unsigned) prm_.get_integer ("Switch timestep after steps");
  outer_solver = prm_.get ("outer solver");
  test_case = prm_.get ("test case");
  output_folder = prm_.get ("Output directory");
  filename_base = prm_.get ("Output filename");
  refinement_strategy = prm_.get ("ref strategy");
  value_phase_field_for_refinement = prm_.get_double ("value phase field for refinement");
  prm_.leave_subsection ();

  if (outer_solver != "active set")
    throw NotImplemented ("outer solver = simple monolithic is not part of the GPU hot path");
  if (test_case != "sneddon" && !miehe () && !hetero ())
    throw NotImplemented ("test case <" + test_case + "> needs meshes / boundary data outside this round's scope");
  if (hetero () && (dim_ != 3 || refinement_strategy != "phase field"))
    throw NotImplemented ("test case = multiple het is available for dim 3 with ref strategy = phase field");
  if (miehe () && dim_ != 2)
    throw NotImplemented ("the Miehe tests are 2-D (meshes/unit_slit.inp)");
  const bool forest_case = (test_case == "sneddon" && dim_ == 2 && refinement_strategy == "fixed preref sneddon"
                            && (n_local_pre_refine != 0 || n_refinement_cycles != 0))
                           || hetero () || (miehe () && n_refinement_cycles != 0 && adaptive_forest);
  if (hetero () && n_refinement_cycles != 0)
    throw NotImplemented ("multiple het: adaptive refinement cycles during the run are not available");
  if (n_local_pre_refine != 0 && !forest_case)
    throw NotImplemented ("local pre-refinement (hanging nodes) is only available for sneddon, dim 2, "
                          "ref strategy = fixed preref sneddon");
  if (n_refinement_cycles != 0 && !forest_case && !(miehe () && refinement_strategy == "phase field"))
    throw NotImplemented ("adaptive refinement (hanging nodes) is not available: use global refinement");

  prm_.enter_subsection ("Problem dependent parameters");
  func_pressure.initialize ("time", prm_.get ("Pressure"));
  G_c = prm_.get_double ("Fracture toughness G_c");
  poisson_ratio_nu = prm_.get_double ("Poisson ratio nu");
  E_modulus = prm_.get_double ("E modulus");
  lame_coefficient_mu = E_modulus / (2.0 * (1 + poisson_ratio_nu));
  lame_coefficient_lambda = (2 * poisson_ratio_nu * lame_coefficient_mu) / (1.0 - 2 * poisson_ratio_nu);
  if (miehe ())
    {
      // Miehe 2010: the Lame coefficients are given directly (cracks.cc:1512-1521)
      lame_coefficient_mu = prm_.get_double ("Lame mu");
      lame_coefficient_lambda = prm_.get_double ("Lame lambda");
    }
  prm_.leave_subsection ();

  timestep_number = 0;
  time = 0;

  // setup_mesh(): "rect -10 -10 [-10] 10 10 [10]", 10 cells per direction,
  // then refine_global (cracks.cc:1207-1253, 1534)
  // Miehe tests: meshes/unit_slit.inp, 2 x 2 cells on the unit square with the slit (1202-1205)
  mesh_.dim = dim_;
  mesh_.slit = miehe () ? 1 : 0;
  const int n = (miehe () ? 2 : 10) << n_global_pre_refine;
  const double length = miehe () ? 1.0 : 20.0, lower = miehe () ? 0.0 : -10.0;
  for (int d = 0; d < 3; ++d)
    {
      mesh_.n[d] = d < dim_ ? n : 1;
      mesh_.h[d] = d < dim_ ? length / n : 1.0;
      mesh_.origin[d] = d < dim_ ? lower : 0.0;
    }
  long long cells = 1;
  for (int d = 0; d < dim_; ++d)
    cells *= n;
  if (hetero ())
    cells = 1ll << (3 * n_global_pre_refine); // one coarse hexahedron
  pcout_ << "Cells:\t" << cells << std::endl;
  if (forest_case && hetero ())
    {
      // meshes/unit_cube_10.inp: one hexahedron [0,10]^3 (cracks.cc:1216-1222), then refine_global
      const int nc[3] = {1, 1, 1};
      const double lo[3] = {0.0, 0.0, 0.0}, hi[3] = {10.0, 10.0, 10.0};
      forest_.reset (new Forest (3, nc, lo, hi));
      forest_->refine_global ((int) n_global_pre_refine);
    }
  else if (forest_case && miehe ())
    {
      // meshes/unit_slit.inp: 2 x 2 cells with the slit (cracks.cc:1202-1205), then refine_global
      const int nc[3] = {2, 2, 1};
      const double lo[3] = {0.0, 0.0, 0.0}, hi[3] = {1.0, 1.0, 0.0};
      forest_.reset (new Forest (2, nc, lo, hi, true));
      forest_->refine_global ((int) n_global_pre_refine);
    }
  else if (forest_case)
    {
      const int nc[3] = {n, n, 1};
      const double lo[3] = {-10.0, -10.0, 0.0}, hi[3] = {10.0, 10.0, 0.0};
      forest_.reset (new Forest (2, nc, lo, hi));
    }

  prm_.enter_subsection ("Solver parameters");
  direct_solver = prm_.get_bool ("Use Direct Inner Solver");
  lower_bound_newton_residual = prm_.get_double ("Newton lower bound");
  max_no_newton_steps = (unsigned) prm_.get_integer ("Newton maximum steps");
  max_no_line_search_steps = (unsigned) prm_.get_integer ("Line search maximum steps");
  line_search_damping = prm_.get_double ("Line search damping");
  decompose_stress_rhs = prm_.get_double ("Decompose stress in rhs");
  decompose_stress_matrix = prm_.get_double ("Decompose stress in matrix");
  prm_.leave_subsection ();
  if ((decompose_stress_rhs > 0 || decompose_stress_matrix > 0) && dim_ != 2)
    throw NotImplemented ("the Miehe stress split is 2-D only (cracks.cc:1923-2120)");
  if (direct_solver)
    pcout_ << "note: 'Use Direct Inner Solver' is ignored, the GPU path always solves matrix-free" << std::endl;
  use_old_timestep_pf = false;
}

// cracks.cc:3820-3892 (uniform mesh: every cell has the same diameter)
void
FracturePhaseFieldProblem::determine_mesh_dependent_parameters ()
{
  double d2 = 0;
  for (int d = 0; d < dim_; ++d)
    d2 += mesh_.h[d] * mesh_.h[d];
  min_cell_diameter = std::sqrt (d2);
  if (use_forest ())
    min_cell_diameter = forest_->min_cell_diameter ();
  // the Miehe tests use the h the mesh will have on its final level (cracks.cc:3839-3854);
  // the coarse cells of unit_slit.inp have diameter sqrt(2)/2
  if (miehe ())
    min_cell_diameter = 0.5 * std::sqrt (2.0)
                        * std::pow (2.0, -1.0 * (n_global_pre_refine + n_refinement_cycles + n_local_pre_refine));
  FunctionParser func;
  prm_.enter_subsection ("Problem dependent parameters");
  func.initialize ("h", prm_.get ("K reg"));
  constant_k = func.value (min_cell_diameter);
  func.initialize ("h", prm_.get ("Eps reg"));
  alpha_eps = func.value (min_cell_diameter);
  prm_.leave_subsection ();
}

// cracks.cc:1579-1680 without the sparse matrix
void
FracturePhaseFieldProblem::setup_system ()
{
  determine_mesh_dependent_parameters ();
  params_.lambda = lame_coefficient_lambda;
  params_.mu = lame_coefficient_mu;
  params_.G_c = G_c;
  params_.kappa = constant_k;
  params_.eps = alpha_eps;
  params_.alpha_biot = 0.0; // cracks.cc:1497
  if (use_forest ())
    {
      forest_create_context ();
      pcout_ << std::endl;
      pcout_ << "DoFs: " << n_nodes () * dim_ << " solid + " << n_nodes () << " phase"
             << " = " << n_nodes () * (dim_ + 1) << std::endl;
      return;
    }
  const int rc = pf_create (&mesh_, &params_, device, 0, 1, nullptr, &ctx_);
  pf_check (ctx_, rc);
  if (miehe ())
    {
      pf_check (ctx_, pf_dirichlet_miehe (ctx_, miehe_kind (), 0.0, 0)); // set_newton_bc, cracks.cc:2584-2625
      // the reference solves these small, ill-conditioned systems with a sparse direct solver or
      // AMG (2750-2771); Jacobi-GMRES needs a basis as long as its iteration count
      pf_check (ctx_, pf_set_krylov_dim (ctx_, 300));
      gmres_max_iterations = std::max (gmres_max_iterations, 3000);
      // from 64 x 64 cells on Jacobi-GMRES needs thousands of iterations per solve: geometric multigrid on the
      // slit square (the golden-sized meshes stay on the Jacobi path their GPU runs were verified with)
      if (mesh_.n[0] >= 64)
        pf_check (ctx_, pf_set_preconditioner (ctx_, 3, 2, 8.0));
    }
  else
    pf_check (ctx_, pf_set_dirichlet_all_faces (ctx_)); // set_newton_bc, cracks.cc:2575-2583 / 2686-2694
  long long nodes = pf_n_dofs (ctx_) / (dim_ + 1);
  pcout_ << std::endl;
  pcout_ << "DoFs: " << nodes * dim_ << " solid + " << nodes << " phase"
         << " = " << nodes * (dim_ + 1) << std::endl;
}

// cracks.cc:2780-2994
double
FracturePhaseFieldProblem::newton_active_set ()
{
  pcout_ << "It.\t#A.Set\t#CycDoF\tResidual\tReduction\tLSrch\t#LinIts" << std::endl;
  double newton_residual = 0;
  pf_check (ctx_, pf_residual (ctx_, nullptr, nullptr, &newton_residual));
  double old_newton_residual = newton_residual;
  unsigned newton_step = 0;
  char buf[256];
  std::snprintf (buf, sizeof buf, "0\t\t\t%e", newton_residual);
  pcout_ << buf << std::endl;
  pf_check (ctx_, pf_active_set_reset (ctx_)); // active_set.clear(), fresh cycle_counter
  unsigned sum_lin_it = 0;
  double new_newton_residual = 0.0;
  while (true)
    {
      int64_t n_active = 0, n_cycling = 0;
      int changed = 0;
      pf_check (ctx_, pf_active_set_update (ctx_, 1e+1 * (hetero () ? active_set_E_ : E_modulus), nullptr, &n_active, &n_cycling, &changed));
      pf_check (ctx_, pf_setup_jacobian (ctx_));                       // assemble_system()
      pf_check (ctx_, pf_residual (ctx_, nullptr, nullptr, nullptr));  // its rhs + set_zero
      int no_linear_iterations = 0;
      pf_check (ctx_, pf_solve (ctx_, gmres_tolerance, gmres_max_iterations, nullptr, &no_linear_iterations));
      sum_lin_it += (unsigned) no_linear_iterations;
      pf_check (ctx_, pf_save_solution (ctx_));
      unsigned line_search_step = 0;
      for (; line_search_step < max_no_line_search_steps; ++line_search_step)
        {
          pf_check (ctx_, pf_update_solution (ctx_, 1.0));
          pf_check (ctx_, pf_residual (ctx_, nullptr, nullptr, &new_newton_residual));
          if (new_newton_residual < newton_residual)
            break;
          pf_check (ctx_, pf_restore_saved_solution (ctx_));
          pf_check (ctx_, pf_scale_update (ctx_, line_search_damping));
        }
      std::snprintf (buf, sizeof buf, "%u\t%lld\t%lld\t%e\t%e\t%u\t%d", newton_step + 1, (long long) n_active,
                     (long long) n_cycling, new_newton_residual, new_newton_residual / newton_residual,
                     line_search_step, no_linear_iterations);
      pcout_ << buf << std::endl;
      old_newton_residual = newton_residual;
      newton_residual = new_newton_residual;
      ++newton_step;
      if (newton_residual < lower_bound_newton_residual && changed == 0)
        {
          pcout_ << "\tNewton iterations: " << newton_step << " total linear iterations: " << sum_lin_it
                 << std::endl;
          break;
        }
      if (newton_step >= max_no_newton_steps)
        {
          pcout_ << "Newton iteration did not converge in " << newton_step << " steps." << std::endl;
          total_newton_its_ += newton_step;
          total_linear_its_ += sum_lin_it;
          throw NoConvergence ("Newton iteration did not converge");
        }
    }
  total_newton_its_ += newton_step;
  total_linear_its_ += sum_lin_it;
  return new_newton_residual / old_newton_residual;
}

// TableHandler::write_text(simple_table_with_separate_column_description), cracks.cc:4469-4475
void
FracturePhaseFieldProblem::write_statistics () const
{
  ::mkdir (output_folder.c_str (), 0755);
  std::ofstream f ((output_folder + "/statistics").c_str ());
  f << "# 1: Timestep No\n# 2: Time\n# 3: DoFs\n# 4: minimum cell diameter\n# 5: Bulk Energy\n# 6: Crack Energy\n";
  if (miehe ())
    f << (miehe_kind () == 1 ? "# 7: Load y\n" : "# 7: Load x\n");
  char buf[256];
  for (const auto &r : statistics_)
    {
      std::snprintf (buf, sizeof buf, "%u %.4f %lld %.8e %.8e %.8e ", r.timestep_no, r.time, r.dofs, r.h_min,
                     r.bulk_energy, r.crack_energy);
      f << buf;
      if (miehe ())
        {
          std::snprintf (buf, sizeof buf, "%.8e ", r.load);
          f << buf;
        }
      f << "\n";
    }
}

// output_results(), cracks.cc:3142-3258: displacement, phasefield and active_set per node, subdomain (and
// emodulus for `multiple het`) per cell; one .vtu piece (this driver runs one rank), the .pvtu / .visit / .pvd
// records, and the two screen lines of the reference.  Not written: the SneddonExactPostProc fields (3164-3167).
void
FracturePhaseFieldProblem::output_results ()
{
  ++output_cycle_;
  if (!write_output)
    return;
  const long long nn = n_nodes ();
  std::vector<double> blk ((size_t) nn * (dim_ + 1));
  pf_check (ctx_, pf_get_state (ctx_, 0, blk.data ()));
  std::vector<uint8_t> active ((size_t) nn, 0);
  pf_check (ctx_, pf_get_active_set (ctx_, active.data ()));
  std::vector<VtuPointField> pfields (3);
  pfields[0].name = "displacement";
  pfields[0].n_components = 3;
  pfields[0].values.assign ((size_t) nn * 3, 0.0);
  pfields[1].name = "phasefield";
  pfields[1].n_components = 1;
  pfields[1].values.resize ((size_t) nn);
  pfields[2].name = "active_set";
  pfields[2].n_components = 1;
  pfields[2].values.resize ((size_t) nn);
  for (long long i = 0; i < nn; ++i)
    {
      for (int d = 0; d < dim_; ++d)
        pfields[0].values[(size_t) (3 * i + d)] = blk[(size_t) (dim_ * i + d)];
      pfields[1].values[(size_t) i] = blk[(size_t) (dim_ * nn + i)];
      pfields[2].values[(size_t) i] = active[(size_t) i];
    }
  // mesh tables: the forest's, or the box (with the doubled nodes of the slit, include/cracks_b200.h pf_mesh.slit)
  std::vector<double> coords;
  std::vector<long long> conn;
  long long nc = 0;
  if (use_forest ())
    {
      coords = forest_->coordinates ();
      conn = forest_->connectivity ();
      nc = forest_->n_cells ();
    }
  else
    {
      const int *n = mesh_.n;
      const int nx1 = n[0] + 1, ny1 = n[1] + 1, nz1 = dim_ == 3 ? n[2] + 1 : 1;
      const long long regular = (long long) nx1 * ny1 * nz1;
      coords.assign ((size_t) nn * dim_, 0.0);
      for (long long p = 0; p < regular; ++p)
        {
          const long long ijk[3] = {p % nx1, (p / nx1) % ny1, p / ((long long) nx1 * ny1)};
          for (int d = 0; d < dim_; ++d)
            coords[(size_t) (dim_ * p + d)] = mesh_.origin[d] + mesh_.h[d] * ijk[d];
        }
      const int slit_row = mesh_.slit ? n[1] / 2 : -1, slit_i0 = n[0] / 2 + 1;
      for (long long p = regular; p < nn; ++p) // copies of the nodes (slit_i0 + k, n[1]/2)
        {
          coords[(size_t) (2 * p)] = mesh_.origin[0] + mesh_.h[0] * (slit_i0 + (p - regular));
          coords[(size_t) (2 * p + 1)] = mesh_.origin[1] + mesh_.h[1] * (n[1] / 2);
        }
      const int nv = 1 << dim_;
      nc = (long long) n[0] * n[1] * (dim_ == 3 ? n[2] : 1);
      conn.resize ((size_t) nc * nv);
      for (long long c = 0; c < nc; ++c)
        {
          const long long ci = c % n[0], cj = (c / n[0]) % n[1], ck = c / ((long long) n[0] * n[1]);
          for (int v = 0; v < nv; ++v)
            {
              const long long i = ci + (v & 1), j = cj + ((v >> 1) & 1), k = ck + ((v >> 2) & 1);
              long long node = i + nx1 * (j + (long long) ny1 * k);
              if (cj == slit_row && j == slit_row && i >= slit_i0)
                node = regular + (i - slit_i0); // the cell row above the slit sees the upper copies
              conn[(size_t) (nv * c + v)] = node;
            }
        }
    }
  std::vector<VtuCellField> cfields;
  if (hetero ())
    {
      // e_mod(cell) = 1.0 + func_emodulus->value(cell->center()), cracks.cc:3170-3183
      const BitmapFunction func_emodulus (source_dir + "/test.pgm", 0, 10, 0, 10, E_modulus, 10.0 * E_modulus);
      VtuCellField e;
      e.name = "emodulus";
      e.values.resize ((size_t) nc);
      for (long long c = 0; c < nc; ++c)
        {
          double x[3];
          forest_->cell_centre (c, x);
          e.values[(size_t) c] = (float) (1.0 + func_emodulus.value (x, 3));
        }
      cfields.push_back (e);
    }
  VtuCellField sub;
  sub.name = "subdomain";
  sub.values.assign ((size_t) nc, 0.0f);
  cfields.push_back (sub);

  pcout_ << "Write solution " << output_cycle_ << std::endl;
  ::mkdir (output_folder.c_str (), 0755);
  char num[16];
  std::snprintf (num, sizeof num, "%05d", output_cycle_);
  const std::string piece = filename_base + num + ".0000.vtu", master = filename_base + num + ".pvtu";
  write_vtu (output_folder + "/" + piece, dim_, nn, coords.data (), nc, conn.data (), pfields, cfields);
  write_pvtu_record (output_folder + "/" + master, {piece}, pfields, cfields);
  const std::string visit = output_folder + "/" + filename_base + num + ".visit";
  write_visit_record (visit, {{piece}});
  output_file_names_by_timestep_.push_back ({piece});
  write_visit_record (output_folder + "/solution.visit", output_file_names_by_timestep_);
  pcout_ << "\tas " << visit << std::endl;
  times_and_names_.emplace_back (time, master);
  write_pvd_record (output_folder + "/solution.pvd", times_and_names_);
}

// cracks.cc:4166-4581, Sneddon branch
void
FracturePhaseFieldProblem::run ()
{
  pcout_ << "Running on 1 GPU" << std::endl;
  set_runtime_parameters ();
  setup_system ();
  // local pre-refinement (cracks.cc:4177-4211): interpolate the initial condition, refine, set up again
  for (unsigned i = 0; i < n_local_pre_refine; ++i)
    {
      pcout_ << "Prerefinement step with h= " << min_cell_diameter << std::endl;
      forest_prerefine ();
      setup_system ();
    }
  if (!(alpha_eps >= min_cell_diameter))
    throw ParameterError ("You need to pick eps >= h");
  if (!(constant_k < 1.0))
    throw ParameterError ("You need to pick K < 1");

  pcout_ << "\n==============================" << "=====================================" << std::endl;
  pcout_ << "Parameters\n" << "==========\n" << "h (min):           " << min_cell_diameter << "\n"
         << "k:                 " << constant_k << "\n" << "eps:               " << alpha_eps << "\n"
         << "G_c:               " << G_c << "\n" << "gamma penal:       " << 0 << "\n"
         << "Poisson nu:        " << poisson_ratio_nu << "\n" << "E modulus:         " << E_modulus << "\n"
         << "Lame mu:           " << lame_coefficient_mu << "\n" << "Lame lambda:       " << lame_coefficient_lambda
         << "\n" << std::endl;

  // initial condition, project_back_phase_field, old = old_old = solution (4233-4277)
  if (miehe ())
    pf_check (ctx_, pf_interpolate_unbroken (ctx_)); // InitialValuesTensionOrShear, cracks.cc:679-691
  else if (use_forest ())
    {
      const std::vector<double> ic = forest_initial_values ();
      pf_check (ctx_, pf_set_state (ctx_, ic.data (), ic.data (), ic.data (), timestep, timestep, 0, func_pressure.value (0.0)));
    }
  else
    pf_check (ctx_, pf_interpolate_sneddon (ctx_, min_cell_diameter));
  output_results (); // cracks.cc:4263
  pf_check (ctx_, pf_project_phase_field (ctx_));
  old_timestep = timestep;
  old_old_timestep = timestep;
  long long nodes = pf_n_dofs (ctx_) / (dim_ + 1), cells = 1;
  for (int d = 0; d < dim_; ++d)
    cells *= mesh_.n[d];
  if (use_forest ())
    cells = n_cells ();
  unsigned refinement_cycle = 0;
  double finishing_timestep_loop = 0;

  do
    {
      if (timestep_number > switch_timestep && switch_timestep > 0)
        timestep = timestep_size_2;
      const double tmp_timestep = timestep;
      old_old_timestep = old_timestep;
      old_timestep = timestep;
      pf_check (ctx_, pf_advance_timestep (ctx_)); // old_old = old; old = solution

      pcout_ << std::endl;
      pcout_ << "\n==============================" << "=========================================" << std::endl;
      pcout_ << "Timestep " << timestep_number << ": " << time << " (" << timestep << ")"
             << "   " << "Cells: " << cells << "   " << "DoFs: " << nodes * (dim_ + 1);
      pcout_ << "\n--------------------------------" << "---------------------------------------" << std::endl;
      pcout_ << std::endl;

    redo_step: // cracks.cc:4307
      time += timestep;
      do
        {
          // catch NoConvergence, retry with a tenth of the step and old_timestep_pf (4320-4358)
          use_old_timestep_pf = false;
          pf_check (ctx_, pf_set_time_parameters (ctx_, old_timestep, old_old_timestep, 0, func_pressure.value (time)));
          // the split is active from the second time step on (cracks.cc:2294, 2338)
          pf_check (ctx_, pf_set_stress_split (ctx_, decompose_stress_matrix > 0 && timestep_number > 0,
                                               decompose_stress_rhs, decompose_stress_matrix));
          if (miehe () && use_forest ())
            {
              const std::vector<double> bc = forest_miehe_boundary_values (time);
              pf_check (ctx_, pf_set_dirichlet_values (ctx_, bc.data ()));     // set_initial_bc(time), 2787
            }
          else if (miehe ())
            pf_check (ctx_, pf_dirichlet_miehe (ctx_, miehe_kind (), time, 1)); // set_initial_bc(time), 2787
          try
            {
              newton_active_set ();
              break;
            }
          catch (NoConvergence &)
            {
              pcout_ << "Solver did not converge! Adjusting time step to " << timestep / 10 << std::endl;
            }
          pcout_ << "Taking old_timestep_pf" << std::endl;
          use_old_timestep_pf = true;
          pf_check (ctx_, pf_restore_old_solution (ctx_));
          time -= timestep;
          timestep = timestep / 10.0;
          time += timestep;
          if (timestep < 1e-12)
            throw NoConvergence ("time step underflow");
          // NB: the reference resets use_old_timestep_pf to false at the top of the retry
          // loop (cracks.cc:4325), so the flag never reaches assemble_system; kept as is.
        }
      while (true);

      pf_check (ctx_, pf_project_phase_field (ctx_));
      if ((miehe () || hetero ()) && use_forest ())
        {
          // predictor-corrector refinement (cracks.cc:4419-4431; every non-sneddon case, so also `multiple het`,
          // whose shipped .prm grows its cracks through cells below the level cap): if refine_mesh() changed the
          // mesh, redo the step
          if (forest_refine_phase_field_and_transfer ())
            {
              pcout_ << "MESH CHANGED!" << std::endl;
              time -= timestep;
              pf_check (ctx_, pf_restore_old_solution (ctx_));
              nodes = n_nodes ();
              cells = n_cells ();
              goto redo_step;
            }
        }
      else if (miehe () && n_refinement_cycles > 0)
        {
          // refine_mesh(), strategy "phase field" (cracks.cc:3971-3995): cells with a phase-field dof
          // below the threshold would be refined and the step redone.  Hanging nodes are out of scope.
          double phi_min = 1.0;
          pf_check (ctx_, pf_phase_field_min (ctx_, &phi_min));
          if (phi_min < value_phase_field_for_refinement)
            throw NotImplemented ("refine_mesh() would refine the mesh in time step "
                                  + std::to_string (timestep_number)
                                  + " (predictor-corrector refinement with hanging nodes is not available)");
        }
      timestep = tmp_timestep;

      double bulk = 0, crack = 0, load = 0;
      pf_check (ctx_, pf_energy (ctx_, &bulk, &crack));
      pcout_ << std::endl;
      pcout_ << "No " << timestep_number << " time " << time << " bulk energy: " << bulk
             << " crack energy: " << crack;
      if (miehe ())
        {
          double lx = 0, ly = 0;
          if (use_forest ())
            pf_check (ctx_, pf_load_cells (ctx_, forest_top_cells_.data (), (int64_t) forest_top_cells_.size (), &lx, &ly));
          else
            pf_check (ctx_, pf_load (ctx_, &lx, &ly)); // compute_load(), cracks.cc:3728-3816
          load = miehe_kind () == 1 ? ly : lx;
          pcout_ << (miehe_kind () == 1 ? "  Load y: " : "  Load x: ") << load;
        }
      pcout_ << std::endl;
      statistics_.push_back ({timestep_number, time, nodes * (dim_ + 1), min_cell_diameter, bulk, crack, load});
      output_results (); // cracks.cc:4466-4467
      write_statistics ();

      pf_check (ctx_, pf_timestep_difference (ctx_, &finishing_timestep_loop));
      if (test_case == "sneddon")
        pcout_ << "Timestep difference linfty: " << finishing_timestep_loop << std::endl;
      ++timestep_number;

      if (test_case == "sneddon" && finishing_timestep_loop < 1.0e-5)
        {
          pf_check (ctx_, pf_tcv (ctx_, &tcv_));
          const double p = func_pressure.value (time), nu = poisson_ratio_nu, E = 1.0, l_0 = 1.0;
          const double ref = dim_ == 2 ? 2.0 * p * l_0 * l_0 * (1.0 - nu * nu) * M_PI / E
                                       : 16.0 * p * l_0 * l_0 * l_0 * (1.0 - nu * nu) / E / 3.0;
          pcout_ << "TCV: value= " << tcv_ << " exact= " << ref << " error= " << std::abs (tcv_ - ref) << std::endl;
          // compute_functional_values(), cracks.cc:3704-3725: COD on the lines x = -1.5 + i/256;
          // lines that carry no mesh face print nothing (compute_cod returns -1e300 there)
          const unsigned N = 16 * 16;
          for (unsigned i = 0; i <= 3 * N && !use_forest (); ++i) // pf_cod: box meshes only so far
            {
              const double x = -1.5 + i * (1.0 / N);
              const double t = (x - mesh_.origin[0]) / mesh_.h[0];
              if (std::abs (t - std::round (t)) * mesh_.h[0] > 1e-8)
                continue;
              double cod = 0;
              int64_t n_faces = 0;
              pf_check (ctx_, pf_cod (ctx_, x, &cod, &n_faces));
              if (n_faces > 0)
                {
                  pcout_ << x << "  " << cod << std::endl;
                  cod_.push_back ({x, cod});
                }
            }
          if (n_refinement_cycles == 0)
            break;
          // refinement cycle (cracks.cc:4525-4560): refine, carry old / old_old over, restart from the
          // re-interpolated initial condition
          --n_refinement_cycles;
          pcout_ << std::endl;
          pcout_ << "\n================== " << std::endl;
          pcout_ << "Refinement cycle " << refinement_cycle << "\n------------------ " << std::endl;
          if (!use_forest ())
            throw NotImplemented ("refinement cycles need the forest path (sneddon, dim 2, fixed preref sneddon)");
          {
            const Forest coarse = *forest_;
            const long long nd_old = pf_n_dofs (ctx_), nn_old = nd_old / 3;
            std::vector<double> old_b ((size_t) nd_old), oldold_b ((size_t) nd_old);
            pf_check (ctx_, pf_get_state (ctx_, 1, old_b.data ()));
            pf_check (ctx_, pf_get_state (ctx_, 2, oldold_b.data ()));
            forest_refine_fixed_preref_sneddon ();
            setup_system ();
            const long long nn_new = n_nodes ();
            // block layout [u | phi] <-> nodal interleaved for the transfer
            auto to_nodal = [](const std::vector<double> &b, long long nn) {
              std::vector<double> v ((size_t) nn * 3);
              for (long long i = 0; i < nn; ++i)
                {
                  v[(size_t) (3 * i)] = b[(size_t) (2 * i)];
                  v[(size_t) (3 * i + 1)] = b[(size_t) (2 * i + 1)];
                  v[(size_t) (3 * i + 2)] = b[(size_t) (2 * nn + i)];
                }
              return v;
            };
            auto to_block = [](const std::vector<double> &v, long long nn) {
              std::vector<double> b ((size_t) nn * 3);
              for (long long i = 0; i < nn; ++i)
                {
                  b[(size_t) (2 * i)] = v[(size_t) (3 * i)];
                  b[(size_t) (2 * i + 1)] = v[(size_t) (3 * i + 1)];
                  b[(size_t) (2 * nn + i)] = v[(size_t) (3 * i + 2)];
                }
              return b;
            };
            std::vector<double> tmp ((size_t) nn_new * 3);
            forest_->transfer (coarse, to_nodal (old_b, nn_old).data (), tmp.data (), 3);
            const std::vector<double> old_new = to_block (tmp, nn_new);
            forest_->transfer (coarse, to_nodal (oldold_b, nn_old).data (), tmp.data (), 3);
            const std::vector<double> oldold_new = to_block (tmp, nn_new);
            const std::vector<double> ic = forest_initial_sneddon ();
            pf_check (ctx_, pf_set_state (ctx_, ic.data (), old_new.data (), oldold_new.data (), old_timestep,
                                          old_old_timestep, 0, func_pressure.value (time)));
            nodes = nn_new;
            cells = n_cells ();
            ++refinement_cycle;
          }
        }
    }
  while (timestep_number <= max_no_timesteps);

  pcout_ << std::endl;
  pcout_ << "Finishing time step loop: " << finishing_timestep_loop << std::endl;
}

// ---- forest path (locally refined meshes) -----------------------------------------

long long
FracturePhaseFieldProblem::n_nodes () const
{
  return use_forest () ? forest_->n_nodes () : pf_n_dofs (ctx_) / (dim_ + 1);
}

long long
FracturePhaseFieldProblem::n_cells () const
{
  return forest_->n_cells ();
}

// refine_mesh(), strategy `fixed preref sneddon` (cracks.cc:3901-3923): cells with a vertex in
// [-2.5, 2.5] x [-1.25, 1.25]
void
FracturePhaseFieldProblem::forest_refine_fixed_preref_sneddon ()
{
  const Forest &f = *forest_;
  std::vector<char> flags ((size_t) f.n_cells (), 0);
  for (long long c = 0; c < f.n_cells (); ++c)
    for (int v = 0; v < 4; ++v)
      {
        const long long n = f.connectivity ()[(size_t) (c * 4 + v)];
        const double x = f.coordinates ()[(size_t) (2 * n)], y = f.coordinates ()[(size_t) (2 * n + 1)];
        if (x <= 2.5 && x >= -2.5 && y <= 1.25 && y >= -1.25)
          flags[(size_t) c] = 1;
      }
  forest_->refine (flags);
}

// setup_system() on the forest: flat tables -> pf_create_forest, Dirichlet rows u = 0 on the boundary
void
FracturePhaseFieldProblem::forest_create_context ()
{
  if (ctx_)
    {
      pf_destroy (ctx_);
      ctx_ = nullptr;
    }
  const Forest &f = *forest_;
  const int dim = f.dim (), nv = 1 << dim, ncomp = dim + 1;
  const long long nn = f.n_nodes (), nc = f.n_cells ();
  std::vector<int64_t> conn (f.connectivity ().begin (), f.connectivity ().end ());
  std::vector<uint8_t> level ((size_t) nc);
  for (long long c = 0; c < nc; ++c)
    level[(size_t) c] = (uint8_t) f.cells ()[(size_t) c].level;
  std::vector<double> level_h ((size_t) (f.max_level () + 1) * dim);
  for (int l = 0; l <= f.max_level (); ++l)
    f.cell_size (l, &level_h[(size_t) (dim * l)]);
  std::vector<int64_t> hang (f.hanging_nodes ().size () * 5);
  for (size_t h = 0; h < f.hanging_nodes ().size (); ++h)
    {
      const HangingNode &hn = f.hanging_nodes ()[h];
      hang[5 * h] = hn.node;
      for (int q = 0; q < 4; ++q)
        hang[5 * h + 1 + (size_t) q] = q < hn.n_parents ? hn.parents[q] : -1;
    }
  pf_forest_mesh fm{};
  fm.dim = dim;
  fm.n_cells = nc;
  fm.n_nodes = nn;
  fm.conn = conn.data ();
  fm.cell_level = level.data ();
  fm.n_levels = f.max_level () + 1;
  fm.level_h = level_h.data ();
  fm.n_hanging = (int64_t) f.hanging_nodes ().size ();
  fm.hanging = hang.empty () ? nullptr : hang.data ();
  // heterogeneous material: Lame coefficients per cell from E(cell centre) (cracks.cc:2207-2216); the assembly
  // uses E + 1, compute_energy uses E (3646-3656)
  std::vector<double> lame_assembly, lame_energy;
  if (hetero ())
    {
      const BitmapFunction func_emodulus (source_dir + "/test.pgm", 0, 10, 0, 10, E_modulus, 10.0 * E_modulus); // 1543
      lame_assembly.resize ((size_t) nc * 2);
      lame_energy.resize ((size_t) nc * 2);
      const double nu = poisson_ratio_nu;
      for (long long c = 0; c < nc; ++c)
        {
          double x[3];
          f.cell_centre (c, x);
          const double E = func_emodulus.value (x, 3);
          // the reference overwrites its member E_modulus cell by cell (cracks.cc:2209-2210), so the active-set
          // constant c = 10 E_modulus (2859) is taken with the value the LAST assembled cell left there: E + 1
          active_set_E_ = E + 1.0;
          for (int which = 0; which < 2; ++which)
            {
              const double Ev = which == 0 ? E + 1.0 : E;
              const double mu = Ev / (2.0 * (1 + nu)), lambda = (2 * nu * mu) / (1.0 - 2 * nu);
              std::vector<double> &out = which == 0 ? lame_assembly : lame_energy;
              out[(size_t) (2 * c)] = lambda;
              out[(size_t) (2 * c + 1)] = mu;
            }
        }
      fm.cell_lame = lame_assembly.data ();
      fm.cell_lame_energy = lame_energy.data ();
    }
  const int rc = pf_create_forest (&fm, &params_, device, &ctx_);
  pf_check (ctx_, rc);
  // set_newton_bc(): u = 0 on every boundary face (cracks.cc:2575-2583, 2686-2694); block layout [u | phi]
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (long long n = 0; n < nn; ++n)
    for (int d = 0; d < dim; ++d)
      {
        lo[d] = std::min (lo[d], f.coordinates ()[(size_t) (dim * n + d)]);
        hi[d] = std::max (hi[d], f.coordinates ()[(size_t) (dim * n + d)]);
      }
  std::vector<uint8_t> dirichlet ((size_t) nn * ncomp, 0), none ((size_t) nn * ncomp, 0);
  for (long long n = 0; n < nn; ++n)
    {
      bool on_boundary = false;
      for (int d = 0; d < dim; ++d)
        {
          const double x = f.coordinates ()[(size_t) (dim * n + d)];
          on_boundary = on_boundary || x == lo[d] || x == hi[d];
        }
      if (on_boundary && !miehe ())
        for (int d = 0; d < dim; ++d)
          dirichlet[(size_t) (dim * n + d)] = 1;
    }
  (void) nv;
  if (miehe ())
    {
      // set_newton_bc() of the Miehe tests (cracks.cc:2584-2625) on the unit square with the slit
      forest_top_nodes_.clear ();
      forest_top_cells_.clear ();
      for (long long n = 0; n < nn; ++n)
        {
          const double x = f.coordinates ()[(size_t) (2 * n)], y = f.coordinates ()[(size_t) (2 * n + 1)];
          const bool top = y == 1.0, bottom = y == 0.0, left = x == 0.0, right = x == 1.0;
          bool cx = false, cy = false;
          if (miehe_kind () == 1)
            {
              cy = bottom || top;
              cx = top;
            }
          else
            {
              const bool lower_slit = y == 0.5 && x >= 0.5 && !f.is_upper_slit_copy (n); // boundary id 4
              cy = left || right || bottom || top || lower_slit;
              cx = bottom || top;
            }
          dirichlet[(size_t) (2 * n)] = cx;
          dirichlet[(size_t) (2 * n + 1)] = cy;
          if (top)
            forest_top_nodes_.push_back (n);
        }
      for (long long c = 0; c < nc; ++c)
        if (f.coordinates ()[(size_t) (2 * f.connectivity ()[(size_t) (4 * c + 2)] + 1)] == 1.0)
          forest_top_cells_.push_back (c);
    }
  pf_check (ctx_, pf_set_constraints (ctx_, dirichlet.data (), none.data ()));
  // the reference hands these systems to AMG; Jacobi-GMRES wants a long basis
  pf_check (ctx_, pf_set_krylov_dim (ctx_, 300));
  gmres_max_iterations = std::max (gmres_max_iterations, 3000);
}

// refine_mesh(), strategy `phase field` during the local pre-refinement (cracks.cc:3971-3995, 4108-4116): cells that
// hold an initial phase-field value below the threshold, up to the level cap
void
FracturePhaseFieldProblem::forest_refine_phase_field_on_initial_values ()
{
  const Forest &f = *forest_;
  const int dim = f.dim (), nv = 1 << dim;
  const int cap = (int) (n_global_pre_refine + n_refinement_cycles + n_local_pre_refine);
  const std::vector<double> ic = forest_initial_values ();
  const long long nn = f.n_nodes ();
  std::vector<char> flags ((size_t) f.n_cells (), 0);
  for (long long c = 0; c < f.n_cells (); ++c)
    {
      if (f.cells ()[(size_t) c].level >= cap)
        continue;
      for (int v = 0; v < nv; ++v)
        if (ic[(size_t) (dim * nn + f.connectivity ()[(size_t) (c * nv + v)])] < value_phase_field_for_refinement)
          flags[(size_t) c] = 1;
    }
  forest_->refine (flags);
}

// BoundaryTensionTest / BoundaryShearTest (cracks.cc:780-797, 845-861) on the forest's nodes, block layout
std::vector<double>
FracturePhaseFieldProblem::forest_miehe_boundary_values (double t) const
{
  const long long nn = forest_->n_nodes ();
  std::vector<double> b ((size_t) nn * 3, 0.0);
  for (const long long n : forest_top_nodes_)
    {
      if (miehe_kind () == 1)
        b[(size_t) (2 * n + 1)] = t;
      else
        b[(size_t) (2 * n)] = -t;
    }
  return b;
}

// refine_mesh(), strategy `phase field` (cracks.cc:3971-3995, 4108-4159): flag cells holding a phase-field dof
// below the threshold (not beyond the level cap), refine with 2:1 balance, interpolate solution / old / old_old
// to the new forest and rebuild the device context.  Returns false if nothing was flagged.
bool
FracturePhaseFieldProblem::forest_refine_phase_field_and_transfer ()
{
  // refine_mesh(), strategy "phase field" (cracks.cc:3971-3995) with the level cap (4108-4116) and the
  // SolutionTransfer of solution / old / old_old (4137-4159); dimension-generic: the Miehe cases (2-D) and
  // `multiple het` (3-D, every step of every non-sneddon case, cracks.cc:4419-4431)
  const Forest coarse = *forest_;
  const int nc = dim_ + 1, nv = 1 << dim_;
  const long long nn_old = coarse.n_nodes (), nd_old = nn_old * nc;
  std::vector<double> blk[3];
  for (int w = 0; w < 3; ++w)
    {
      blk[w].resize ((size_t) nd_old);
      pf_check (ctx_, pf_get_state (ctx_, w, blk[w].data ()));
    }
  const int cap = (int) (n_global_pre_refine + n_refinement_cycles + n_local_pre_refine);
  std::vector<char> flags ((size_t) coarse.n_cells (), 0);
  bool any = false;
  for (long long c = 0; c < coarse.n_cells (); ++c)
    {
      if (coarse.cells ()[(size_t) c].level >= cap)
        continue;
      for (int v = 0; v < nv; ++v)
        if (blk[0][(size_t) (dim_ * nn_old + coarse.connectivity ()[(size_t) (nv * c + v)])] < value_phase_field_for_refinement)
          flags[(size_t) c] = 1;
      any = any || flags[(size_t) c];
    }
  if (!any)
    return false;
  forest_->refine (flags);
  setup_system ();
  const long long nn_new = forest_->n_nodes ();
  std::vector<double> out[3];
  std::vector<double> nodal_old ((size_t) nd_old), nodal_new ((size_t) nn_new * nc);
  for (int w = 0; w < 3; ++w)
    {
      // block layout [u | phi] <-> node-major for the transfer
      for (long long i = 0; i < nn_old; ++i)
        {
          for (int d = 0; d < dim_; ++d)
            nodal_old[(size_t) (nc * i + d)] = blk[w][(size_t) (dim_ * i + d)];
          nodal_old[(size_t) (nc * i + dim_)] = blk[w][(size_t) (dim_ * nn_old + i)];
        }
      forest_->transfer (coarse, nodal_old.data (), nodal_new.data (), nc);
      out[w].resize ((size_t) nn_new * nc);
      for (long long i = 0; i < nn_new; ++i)
        {
          for (int d = 0; d < dim_; ++d)
            out[w][(size_t) (dim_ * i + d)] = nodal_new[(size_t) (nc * i + d)];
          out[w][(size_t) (dim_ * nn_new + i)] = nodal_new[(size_t) (nc * i + dim_)];
        }
    }
  pf_check (ctx_, pf_set_state (ctx_, out[0].data (), out[1].data (), out[2].data (), old_timestep, old_old_timestep, 0,
                                func_pressure.value (time)));
  return true;
}

void
FracturePhaseFieldProblem::forest_prerefine ()
{
  if (hetero ())
    forest_refine_phase_field_on_initial_values ();
  else
    forest_refine_fixed_preref_sneddon ();
}

// the initial values of the test case at the forest's nodes, block layout [u | phi]
std::vector<double>
FracturePhaseFieldProblem::forest_initial_values () const
{
  if (!hetero ())
    return forest_initial_sneddon ();
  const Forest &f = *forest_;
  const long long nn = f.n_nodes ();
  std::vector<double> b ((size_t) nn * 4, 0.0);
  for (long long n = 0; n < nn; ++n)
    b[(size_t) (3 * nn + n)] = initial_multiple_het_3d (&f.coordinates ()[(size_t) (3 * n)], min_cell_diameter);
  return b;
}

// InitialValuesSneddon<2> (cracks.cc:381-406) at the forest's nodes, block layout
std::vector<double>
FracturePhaseFieldProblem::forest_initial_sneddon () const
{
  const Forest &f = *forest_;
  const long long nn = f.n_nodes ();
  std::vector<double> b ((size_t) nn * 3, 0.0);
  for (long long n = 0; n < nn; ++n)
    {
      const double x = f.coordinates ()[(size_t) (2 * n)], y = f.coordinates ()[(size_t) (2 * n + 1)];
      const bool broken = (x * x <= 1.0) && (std::abs (2.0 * y) <= 2.0 * min_cell_diameter);
      b[(size_t) (2 * nn + n)] = broken ? 0.0 : 1.0;
    }
  return b;
}

} // namespace cracks
