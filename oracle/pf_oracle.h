/*
 * pf_oracle.h -- CPU oracle for the (u,phi) hot path of tjhei/cracks.
 *
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (cracks_b200/) never links or calls it.
 *
 * Parity status: PINNED.  oracle/newton_oracle.py drives these functions
 * through the reference's active-set Newton loop (cracks.cc:2780-2994) and
 * reproduces tests/sneddon_3d_1.mpirun=4.statistics (all four rows, bulk and
 * crack energy, <= 1e-8 relative), the initial residual 6.744161e+01 and the
 * TCV 0.0399535 of tests/sneddon_3d_1.mpirun=4.output; see
 * tests/test_oracle_golden.py and tests/golden/sneddon_3d_1.json.
 * The reference itself cannot be compiled here (needs deal.II, Trilinos,
 * p4est, MPI), so there is no oracle/_ref.
 */
#ifndef PF_ORACLE_H
#define PF_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct
{
  int dim;          /* 2 or 3 */
  int n[3];         /* cells per direction (subdivided_hyper_rectangle + refine_global) */
  double h[3];      /* cell edge lengths */
  double origin[3]; /* lower corner */
} pfo_mesh;

typedef struct
{
  double lambda, mu;       /* Lame coefficients (cracks.cc:1507-1510) */
  double G_c;              /* fracture toughness */
  double kappa, eps;       /* constant_k, alpha_eps (cracks.cc:3876-3882) */
  double pressure;         /* func_pressure(time) (cracks.cc:2145) */
  double alpha_biot;       /* 0 in the reference (cracks.cc:1497) */
  double dt_old, dt_oldold;/* old_timestep, old_old_timestep */
  int use_old_timestep_pf; /* cracks.cc:2276 */
} pfo_params;

#define PFO_DECL(D) \
  void pfo_residual_##D (const pfo_mesh *, const pfo_params *, const double *sol, const double *old, \
                         const double *oldold, const unsigned char *constrained, double *r_pde, double *r_total); \
  long pfo_csr_nnz_##D (const pfo_mesh *); \
  void pfo_csr_pattern_##D (const pfo_mesh *, long *rowptr, int *col); \
  void pfo_assemble_jacobian_##D (const pfo_mesh *, const pfo_params *, const double *sol, const double *old, \
                                  const double *oldold, const unsigned char *constrained, \
                                  const long *rowptr, const int *col, double *val); \
  void pfo_apply_jacobian_##D (const pfo_mesh *, const pfo_params *, const double *sol, const double *old, \
                               const double *oldold, const unsigned char *constrained, const double *x, double *y); \
  void pfo_lumped_mass_##D (const pfo_mesh *, double *mass); \
  void pfo_energy_##D (const pfo_mesh *, const pfo_params *, const double *sol, double *bulk, double *crack); \
  double pfo_tcv_##D (const pfo_mesh *, const double *sol); \
  double pfo_cod_##D (const pfo_mesh *, const double *sol, double eval_line, long *n_faces); \
  long pfo_active_set_##D (const pfo_mesh *, double c_scale, const double *r_total, const double *mass, \
                           const double *old, double *sol, const int *cycle, unsigned char *active, long *n_cycling); \
  void pfo_spmv_##D (long nrows, const long *rowptr, const int *col, const double *val, const double *x, double *y);

PFO_DECL (2d)
PFO_DECL (3d)

int pfo_num_threads (void);

#ifdef __cplusplus
}
#endif
#endif
