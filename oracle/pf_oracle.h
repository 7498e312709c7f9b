/*
 * pf_oracle.h -- CPU oracle for the (u,phi) hot path of tjhei/cracks.
 *
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (cracks_b200/) never links or calls it.
 *
 * Parity status: PINNED against every golden the reference ships for this path except the
 * gmsh three-point case (tests/test_oracle_golden.py, test_oracle_miehe.py, test_oracle_adaptive.py;
 * fixtures under tests/golden/ with the scripts that transcribed them):
 *   sneddon_3d_1 (uniform 3-D), sneddon_2d_1 (hanging nodes), miehe_shear_2 + its 2-rank variant
 *   (stress split), miehe_shear_1 and miehe_tension_adaptive_1 (split / predictor-corrector
 *   refinement), hetero_3d_1 (3-D hanging nodes, bitmap coefficient), the Catch cases of the
 *   2x2 eigen routine.  oracle/newton_oracle.py and oracle/adaptive_oracle.py drive these C
 *   functions through the reference's time loop and active-set Newton iteration.
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
  int slit;         /* 2-D only: 1 = the unit_slit.inp topology (meshes/unit_slit.inp, cracks.cc:1202-1205):
                       the nodes on the line y = origin + h*n[1]/2 with x-index > n[0]/2 are doubled; the cells
                       above the line use the copies (appended after the regular nodes), the cells below the
                       originals.  The two sides are not connected: that is the pre-existing crack. */
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
  int split;               /* decompose_stress_matrix > 0 && timestep_number > 0 (cracks.cc:2294, 2338); 2-D only */
  double d_rhs, d_mat;     /* decompose_stress_rhs / decompose_stress_matrix (cracks.cc:1568-1569) */
} pfo_params;

/* General (locally refined) mesh of axis-aligned Q1 cells, as produced by the reference's
 * refine_mesh() on the box meshes (cracks.cc:3895-4163): explicit connectivity, one edge-length
 * vector per cell.  Hanging-node constraints are resolved by the caller (oracle/newton_oracle.py),
 * these functions return the raw, unconstrained cell sums. */
typedef struct
{
  long n_cells, n_nodes;
  const long *cells;     /* [n_cells][2^dim] node numbers, vertex order lexicographic (x fastest) */
  const double *cell_h;  /* [n_cells][dim] edge lengths */
  const double *cell_lame; /* NULL, or [n_cells][2] = (lambda, mu) per cell: `test case = multiple het`
                              recomputes the Lame coefficients from E(cell centre) in every cell
                              (cracks.cc:2207-2216; compute_energy uses other values, 3646-3656) */
} pfo_gmesh;

#define PFO_DECL(D) \
  void pfo_g_residual_##D (const pfo_gmesh *, const pfo_params *, const double *sol, const double *old, \
                           const double *oldold, double *r_raw); \
  void pfo_g_cell_matrices_##D (const pfo_gmesh *, const pfo_params *, const double *sol, const double *old, \
                                const double *oldold, double *mats /* [n_cells][ndpc][ndpc], row = test */); \
  void pfo_g_lumped_mass_##D (const pfo_gmesh *, double *mass); \
  void pfo_g_functionals_##D (const pfo_gmesh *, const pfo_params *, const double *sol, double *bulk_crack_tcv); \
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
  long pfo_n_nodes_##D (const pfo_mesh *); \
  void pfo_cell_nodes_##D (const pfo_mesh *, long *cells /* [n_cells][2^dim] */); \
  void pfo_node_coords_##D (const pfo_mesh *, double *xyz /* [n_nodes][dim] */); \
  void pfo_spmv_##D (long nrows, const long *rowptr, const int *col, const double *val, const double *x, double *y);

PFO_DECL (2d)
PFO_DECL (3d)

/* 2-D only (the reference's split and load functional are 2-D, cracks.cc:1923-2120, 3728-3816) */
void pfo_eigen_2x2 (const double *m /* row-major 2x2 */, double *ev1, double *ev2, double *P /* row-major */);
void pfo_decompose_stress_2d (const double *E, const double *E_lin, double lambda, double mu, int derivative,
                              double *s_plus, double *s_minus);
void pfo_load_2d (const pfo_mesh *, const pfo_params *, const double *sol, double *load /* [2] */);

int pfo_num_threads (void);
void pfo_set_num_threads (int n);

#ifdef __cplusplus
}
#endif
#endif
