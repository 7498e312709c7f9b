/*
 * cracks_b200.h -- C ABI of the B200-native (u,phi) hot path of tjhei/cracks.
 *
 * The reference (cracks.cc) has no plugin/FFI interface; its de-facto seam is
 * deal.II's duck-typed linear operator (`A.vmult(dst, src)`) consumed by
 * SolverGMRES at cracks.cc:2764-2771, fed by assemble_system() at
 * cracks.cc:2917.  Each entry point below names the reference code it
 * replaces.  All functions return PF_OK (0) or a negative pf_status; no C++
 * exception crosses this boundary.  Plain pointers and sizes only.
 *
 * Mesh: the uniform box of GridGenerator::subdivided_hyper_rectangle +
 * refine_global (cracks.cc:1248-1253, 1534), Q1 elements, nodes numbered
 * lexicographically (x fastest).
 *
 * Host vector layout ("block layout", cracks.cc:1587-1590, 1657-1674):
 *   [ u-block | phi-block ],  u-block = dim doubles per node (node-major),
 *   phi-block = one double per node; length n_dofs = (dim+1) * n_nodes.
 * Constraint masks use the same indexing, one byte per dof (0 / 1).
 *
 * Device-resident calls (suffix _dev) take pointers to device memory in the
 * library's internal layout: node-major interleaved, (dim+1) doubles per node
 * ([ux,uy,uz,phi] = 32 B per 3-D node), restricted to the calling rank's slab
 * of node planes (see pf_local_layout).  They never touch host memory and
 * enqueue on the context's stream.
 *
 * Threading: one host thread per context (like one MPI rank, cracks.cc:4587);
 * calls on one context are not re-entrant.  With nranks > 1 every call marked
 * [collective] must be entered by all ranks in the same order.
 */
#ifndef CRACKS_B200_H
#define CRACKS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum
{
  PF_OK = 0,
  PF_BAD_ARG = -1,
  PF_CUDA_ERROR = -2,
  PF_NCCL_ERROR = -3,
  PF_NO_CONVERGENCE = -4, /* SolverControl::NoConvergence, cracks.cc:2987 / GMRES 2762-2771 */
  PF_NUMERIC = -5,        /* non-finite values; the abort() of cracks.cc:1732-1736 */
  PF_UNSUPPORTED = -6
} pf_status;

typedef struct pf_ctx pf_ctx;

typedef struct
{
  int dim;          /* 2 or 3 (Global parameters / Dimension, cracks.cc:4641) */
  int n[3];         /* cells per direction: 10 * 2^(global pre-refinement) for Sneddon */
  double h[3];      /* cell edge lengths */
  double origin[3]; /* lower corner of the box */
  int slit;         /* 2-D, single rank: 1 = the topology of meshes/unit_slit.inp after refine_global
                       (Miehe tests, cracks.cc:1202-1205): the nodes on the line y = origin + h n[1]/2 with
                       x-index > n[0]/2 are doubled, the cells above the line use the copies, which are
                       numbered after the (n[0]+1)(n[1]+1) regular nodes.  0 for the box meshes. */
} pf_mesh;

typedef struct
{
  double lambda, mu;  /* Lame coefficients, cracks.cc:1507-1510 */
  double G_c;         /* Fracture toughness G_c */
  double kappa;       /* constant_k  = K reg(h),   cracks.cc:3879 */
  double eps;         /* alpha_eps   = Eps reg(h), cracks.cc:3881 */
  double alpha_biot;  /* 0 in the reference, cracks.cc:1497 */
} pf_params;

/* Slab decomposition of the node planes (last coordinate) over ranks. */
typedef struct
{
  int64_t n_nodes_global;
  int64_t n_nodes_plane;   /* nodes in one plane of the slowest coordinate */
  int plane_begin;         /* first local node plane (global index), including the lower ghost */
  int plane_end;           /* one past the last local node plane, including the upper ghost */
  int owned_begin;         /* planes [owned_begin, owned_end) are owned by this rank */
  int owned_end;
  int ncomp;               /* dim + 1 */
} pf_local_layout;

/* ---- life cycle --------------------------------------------------------- */

/* Replaces setup_system()'s allocation of matrix/vectors (cracks.cc:1579-1680):
 * allocates all device state for one rank.  nccl_id is a 128-byte
 * ncclUniqueId (from pf_nccl_unique_id on rank 0) or NULL when nranks == 1. */
int pf_create (const pf_mesh *mesh, const pf_params *params, int device,
               int rank, int nranks, const void *nccl_id, pf_ctx **out);
int pf_destroy (pf_ctx *ctx);
const char *pf_last_error (const pf_ctx *ctx);
int pf_nccl_unique_id (void *id128);
int pf_get_layout (const pf_ctx *ctx, pf_local_layout *out);
/* The slab decomposition itself, without touching a GPU (pure host): which
 * node planes / cell layers `rank` of `nranks` holds, owns and evaluates.
 * Stands in for the p4est partition of the uniform forest (cracks.cc:1083,
 * 1180); cell_* pointers may be NULL. */
int pf_slab_layout (const pf_mesh *mesh, int rank, int nranks, pf_local_layout *out, int *cell_begin,
                    int *cell_end, int *own_cell_begin, int *own_cell_end);
/* A locally refined
 * mesh as flat tables, the output of the host forest (cracks_b200/host/forest.h) that stands in for
 * the reference's p4est triangulation after refine_mesh() (cracks.cc:3895-4163).  All cells are
 * axis-aligned; a cell's level selects its edge lengths.  Hanging nodes (make_hanging_node_constraints,
 * 1630-1634) are constrained to the mean of their 2 (edge) or 4 (face) parents.  Jacobi preconditioner;
 * several GPUs: pf_create_forest_distributed.  Dirichlet rows come from pf_set_constraints, initial values from pf_set_state. */
typedef struct
{
  int dim;
  int64_t n_cells, n_nodes;
  const int64_t *conn;        /* [n_cells][2^dim] node numbers, vertex order lexicographic (x fastest) */
  const uint8_t *cell_level;  /* [n_cells] */
  int n_levels;
  const double *level_h;      /* [n_levels][dim] edge lengths of the cells of each level */
  int64_t n_hanging;
  const int64_t *hanging;     /* [n_hanging][5]: node, parent 0..3 (-1 = unused) */
  const double *cell_lame;        /* NULL, or [n_cells][2] = (lambda, mu) used by the assembly (cracks.cc:2207-2216) */
  const double *cell_lame_energy; /* NULL (= cell_lame), or the values compute_energy uses (3646-3656) */
} pf_forest_mesh;
int pf_create_forest (const pf_forest_mesh *mesh, const pf_params *params, int device, pf_ctx **out);
/* The same on `nranks` GPUs (one process per GPU, nccl_id from pf_nccl_unique_id of rank 0).  Stands in for the
 * p4est partition of the reference (parallel::distributed::Triangulation, cracks.cc:1083, 1180): the cells, given
 * identically on every rank (the host forest orders them by lower corner), are cut into `nranks` contiguous
 * equal-count ranges the way p4est cuts its Morton curve; rank r evaluates range r, results do not depend on the cut.  Nodal vectors are replicated on every rank and the per-rank cell
 * sums meet in ONE all-reduce per operator application / residual / diagonal / functional (the reference:
 * ghost import + compress(add), cracks.cc:2147-2154, 2470-2475).  Every [collective] call must be entered by all
 * ranks; host vectors are the complete vectors on every rank. */
int pf_create_forest_distributed (const pf_forest_mesh *mesh, const pf_params *params, int device, int rank,
                                  int nranks, const void *nccl_id, pf_ctx **out);

/* One level of the geometric multigrid hierarchy that replaces the reference's ML AMG set-up
 * (cracks.cc:2477-2497) as `rank` of `nranks` sees it. */
typedef struct
{
  int n[3];                /* cells per direction of this level */
  int replicated;          /* 1: this level is held completely by every rank (no decomposition) */
  int mode_below;          /* level below: 0 = none (coarsest), 1 = same z-slab decomposition, 2 = replicated */
  pf_local_layout layout;  /* node planes this rank holds / owns on this level */
  int inject_begin, inject_end;     /* coarse planes this rank fills when the state is injected from the level above */
  int restrict_begin, restrict_end; /* coarse planes this rank restricts the residual of the level above into */
} pf_mg_level;
/* Pure host function (no GPU): the hierarchy and transfer ranges pf_setup_jacobian uses, dim 3.
 * Levels are ordered fine to coarse; *n_levels <= max_levels. */
int pf_mg_hierarchy (const pf_mesh *mesh, int rank, int nranks, pf_mg_level *levels, int max_levels, int *n_levels);
int64_t pf_n_dofs (const pf_ctx *ctx);
/* the CUDA stream all work of this context is enqueued on (a cudaStream_t) */
void *pf_stream (const pf_ctx *ctx);
int pf_synchronize (pf_ctx *ctx);

/* ---- state and constraints ---------------------------------------------- */

/* The three ghosted vectors assemble_system() imports (cracks.cc:2147-2154),
 * the time-step sizes of the pf_extra extrapolation (2268-2277), the pressure
 * func_pressure(time) (2145).  Host pointers, block layout.  old / oldold may
 * be NULL to keep the previous ones.  [collective] */
int pf_set_state (pf_ctx *ctx, const double *sol, const double *old, const double *oldold,
                  double dt_old, double dt_oldold, int use_old_timestep_pf, double pressure);
int pf_get_solution (pf_ctx *ctx, double *sol);
/* which = 0: solution, 1: old_solution, 2: old_old_solution (what SolutionTransfer carries to a refined
 * mesh, cracks.cc:4137-4159); host buffer in block layout */
int pf_get_state (pf_ctx *ctx, int which, double *out);
/* in-place update of the linearisation point: sol += alpha * dx (line search, cracks.cc:2944) */
int pf_update_solution (pf_ctx *ctx, double alpha);
int pf_set_params (pf_ctx *ctx, const pf_params *params);

/* constraints_update = Dirichlet rows (set_newton_bc, cracks.cc:2709-2714)
 * united with the active set (2878-2881); both masks in block layout, either
 * may be NULL (= keep).  */
int pf_set_constraints (pf_ctx *ctx, const uint8_t *dirichlet_mask, const uint8_t *active_mask);
/* Dirichlet rows of the Sneddon / 3-D cases: all u components on every face
 * (cracks.cc:2575-2583, 2686-2694), built on the device. */
int pf_set_dirichlet_all_faces (pf_ctx *ctx);

/* ---- the hot path ------------------------------------------------------- */

/* assemble_nl_residual() + constraints_update.set_zero + l2_norm
 * (cracks.cc:2507-2512, 2790-2794, 2946-2949).  r_pde / r_total are host
 * buffers in block layout and may be NULL.  [collective] */
int pf_residual (pf_ctx *ctx, double *r_pde, double *r_total, double *l2_norm);

/* Linearise at the current state: what assemble_system(false) does before
 * solve() (cracks.cc:2917) minus the sparse matrix -- it computes the
 * operator diagonal (used for constrained rows and by the preconditioner).
 * Must be called after pf_set_state / pf_set_constraints and before
 * pf_apply_jacobian / pf_solve.  [collective] */
int pf_setup_jacobian (pf_ctx *ctx);

/* y = J(U) x with constraints_update applied the way
 * AffineConstraints::distribute_local_to_global does (rows and columns of
 * constrained dofs eliminated, positive diagonal kept): the vmult of
 * system_pde_matrix at cracks.cc:2770.  Host buffers, block layout.  [collective] */
int pf_apply_jacobian (pf_ctx *ctx, const double *x, double *y);
/* same, device-resident, internal layout, local slab (ghost planes of x are
 * refreshed by the call).  [collective] */
int pf_apply_jacobian_dev (pf_ctx *ctx, double *x_dev, double *y_dev);

/* Preconditioner of pf_solve, the stand-in for the two ML AMG hierarchies of
 * cracks.cc:2477-2497: kind 0 = Jacobi, kind 1 = matrix-free geometric
 * multigrid V-cycle on the 10*2^l mesh family with Chebyshev-Jacobi smoothing
 * of the given degree and smoothing range (default 1, 2, 20; kind 2 = same with
 * the exact 27-point operator in the smoother).  Dim 3 box meshes, any number of
 * ranks: a level keeps the z-slab decomposition while the slabs stay aligned with
 * the coarse cells and is replicated on every rank below that (pf_mg_hierarchy
 * describes the levels).  kind 3 = kind 1, and the V-cycle also on 2-D box meshes
 * and on the unit square with the slit of the Miehe tests (single rank; n -> n/2
 * down to 2 x 2 cells, the slit is kept on every level; a smoothing range of 8 is
 * the robust choice there, 20 stalls GMRES once the active set is irregular).
 * With kinds 1 and 2, 2-D meshes use Jacobi; forest meshes always do.  Takes
 * effect at the next pf_setup_jacobian. */
int pf_set_preconditioner (pf_ctx *ctx, int kind, int cheb_degree, double cheb_ratio);
/* Precision of the multigrid V-cycle: 64 (default) or 32.  With 32 the smoother
 * operator, the Chebyshev steps and the grid transfers of every level run in
 * FP32 on float copies of the linearisation state (made by pf_setup_jacobian);
 * the Krylov vectors are converted at the preconditioner boundary and the outer
 * GMRES, the Jacobian apply (pf_apply_jacobian) and all residuals stay FP64, so
 * converged results are unchanged -- like the ML AMG of cracks.cc:2477-2497, the
 * preconditioner only influences #LinIts.  Opt-in (also PF_MG_FP32=1 in the
 * environment at pf_create); takes effect at the next pf_setup_jacobian. */
int pf_set_multigrid_precision (pf_ctx *ctx, int bits);
/* z = M^-1 v, one application of the preconditioner pf_solve uses (the vmult of
 * BlockDiagonalPreconditioner, cracks.cc:2717-2740); host buffers, block layout,
 * n_dofs doubles each.  pf_setup_jacobian must have been called. */
int pf_apply_preconditioner (pf_ctx *ctx, const double *v, double *z);
/* Precision of the Krylov operator (the Jacobian inside pf_solve / pf_apply_jacobian*): 64 = exact FP64 (default),
 * 32 = inexact Newton, the same 27-point evaluation in FP32 on FP64 vectors (3-D box meshes with cubic cells).
 * Residuals (cracks.cc:2393-2432), the active-set test, energies and the outer GMRES vectors stay FP64, so the
 * Newton fixed point and every parity metric are unchanged.  Takes effect at the next pf_setup_jacobian. */
int pf_set_jacobian_precision (pf_ctx *ctx, int bits);
/* Deterministic scatter (3-D box meshes): the tiled kernels (operator, smoother operator, residual) and the diagonal are
 * launched colour by colour -- 8 launches whose tiles share no node -- so every nodal sum is formed in the same order in
 * every run and a Newton history is reproducible bit for bit on a given number of ranks.  Default off: one launch,
 * red.global.add in scheduler order, results equal up to the last bits (which can move the round-off-determined
 * active-set history of the reference's algorithm, cracks.cc:2863, never the converged step). */
int pf_set_deterministic (pf_ctx *ctx, int on);
/* Smoother operator of the multigrid V-cycle with (1, default) or without (0) the (phi,u) block.  Without it the
 * preconditioner is block diagonal like the reference's BlockDiagonalPreconditioner (cracks.cc:2717-2740) and a
 * smoother application neither stages nor interpolates the state U.  The Krylov operator is not affected. */
int pf_set_multigrid_coupling (pf_ctx *ctx, int coupled);
/* pf_solve as two block stages instead of one GMRES on the whole system.  The reference zeroes the linearised
 * stresses for phi trial functions (cracks.cc:2333-2337), so block (u,phi) of its Jacobian is identically zero and
 * J dx = b is  A du = b_u  followed by  B dphi = b_phi - C du.  Each stage runs the same preconditioned GMRES with the
 * other block's dofs treated as constrained; together they end at |b - J dx| <= tol like the monolithic solve
 * (cracks.cc:2762-2771).  The u equation is linear in u and independent of phi within a time step (phi~ is
 * extrapolated from the old time steps, cracks.cc:2262-2277), so after the first Newton step the u stage is skipped
 * as long as |b_u| is below its share of the tolerance, and the phi stage evaluates only the (phi,phi) block
 * (pf_apply3d_phi.cuh) -- the active-set iteration of a time step then costs scalar solves.  3-D box meshes.
 * Default on; PF_BLOCK_SOLVE=0/1 sets the default at pf_create, 0 gives the monolithic GMRES. */
int pf_set_block_solve (pf_ctx *ctx, int on);
int pf_get_block_solve (pf_ctx *ctx); /* 1 / 0: the setting in force (library default or PF_BLOCK_SOLVE) */
/* Counters since pf_create: out[0] = pf_solve calls that ran as stages, out[1] = of which with a u stage (the others
 * found |b_u| below its tolerance), out[2] / out[3] = GMRES iterations of the u / phi stages. */
int pf_get_block_solve_stats (pf_ctx *ctx, int64_t *out);
/* Tests: restrict the operator to one block (0 = whole system, 1 = u block, 2 = phi block) for pf_apply_jacobian /
 * pf_apply_preconditioner; needs pf_set_block_solve(ctx, 1) and pf_setup_jacobian before. */
int pf_debug_set_block (pf_ctx *ctx, int block);
/* The multigrid V-cycle as ONE CUDA graph launch (captured after every pf_setup_jacobian, halo exchanges and NCCL
 * calls included): on several GPUs a cycle is bound by the host's launch rate (about 200 launches and 20-30 NCCL calls
 * on levels of a few cell layers per rank).  Default off; PF_MG_GRAPH=1 switches it on at pf_create.  [collective] */
int pf_set_multigrid_graph (pf_ctx *ctx, int on);
/* Restart length of GMRES (deal.II's SolverGMRES default keeps 28 basis vectors,
 * cracks.cc:2764; this library's default is 30).  Ill-conditioned small 2-D
 * problems, which the reference hands to a sparse direct solver (2750-2759),
 * want a basis as large as the iteration count. */
int pf_set_krylov_dim (pf_ctx *ctx, int m);

/* diag(J) as used for constrained rows and the Jacobi/Chebyshev smoother */
int pf_jacobian_diagonal (pf_ctx *ctx, double *diag);

/* assemble_diag_mass_matrix(), cracks.cc:2514-2562; one double per node */
int pf_lumped_mass (pf_ctx *ctx, double *mass);

/* New active set, cracks.cc:2822-2899 + cycle detection 2901-2907, using the
 * r_total of the last pf_residual: a phi dof is active iff
 *   r_total/m + c (phi - phi_old) > 0   or it switched >= 5 times.
 * Resets phi to phi_old on the set, installs the set into the constraints.
 * active_mask (one byte per node, may be NULL) receives the set.  [collective] */
int pf_active_set_update (pf_ctx *ctx, double c, uint8_t *active_mask,
                          int64_t *n_active, int64_t *n_cycling, int *changed);
/* The current active set without changing it (one byte per node, 1 = active): what
 * output_results() writes as the "active_set" field, cracks.cc:3193-3208. */
int pf_get_active_set (pf_ctx *ctx, uint8_t *active_mask);
int pf_active_set_reset (pf_ctx *ctx);

/* solve(): GMRES(<= max_it, tol_rel * ||r_pde||_2) on J dx = r_pde with the
 * matrix-free operator, then constraints_update.distribute (constrained
 * entries of dx = 0); cracks.cc:2744-2777.  dx (host, block layout) may be
 * NULL: the update stays on the device for pf_update_solution.
 * Returns PF_NO_CONVERGENCE like SolverControl.  [collective] */
int pf_solve (pf_ctx *ctx, double tol_rel, int max_it, double *dx, int *n_it);

/* ---- functionals (parity metrics) --------------------------------------- */
int pf_energy (pf_ctx *ctx, double *bulk, double *crack);   /* cracks.cc:3615-3701 */
int pf_tcv (pf_ctx *ctx, double *tcv);                      /* cracks.cc:3553-3589 */
/* compute_cod(eval_line), cracks.cc:3452-3549: 0.5 int u.grad(phi) over the mesh faces on the plane
 * x = eval_line; n_faces = 0 means no mesh face lies on that plane (the reference returns -1e300). */
int pf_cod (pf_ctx *ctx, double eval_line, double *value, int64_t *n_faces);
int pf_project_phase_field (pf_ctx *ctx);                   /* cracks.cc:3109-3137 */
int pf_interpolate_sneddon (pf_ctx *ctx, double h_diam);    /* InitialValuesSneddon, cracks.cc:381-406 */

/* ---- 2-D Miehe tests (configs 2 and 4): slit mesh, stress split, load ------ */

/* decompose_stress(): Miehe's tensile/compressive split and its linearisation
 * (cracks.cc:1923-2120, 1691-1737).  active = (decompose_stress_matrix > 0 &&
 * timestep_number > 0), the condition of cracks.cc:2294 and 2338; the two
 * factors are "Decompose stress in rhs / matrix" (1568-1569).  2-D only, like
 * the reference. */
int pf_set_stress_split (pf_ctx *ctx, int active, double decompose_rhs, double decompose_matrix);
/* set_boundary_conditions() for `miehe tension` (kind 1) and `miehe shear`
 * (kind 2), cracks.cc:2584-2625: rebuilds the Dirichlet bits of the constraint
 * mask on the device and, if set_values != 0, writes the boundary values at
 * `time` into the solution (set_initial_bc, 2700-2707).  Needs mesh.slit. */
int pf_dirichlet_miehe (pf_ctx *ctx, int kind, double time, int set_values);
/* InitialValuesTensionOrShear / InitialValuesNoCrack (cracks.cc:679-691, 727-737):
 * u = 0, phi = 1; old = oldold = solution. */
int pf_interpolate_unbroken (pf_ctx *ctx);
/* compute_load(), cracks.cc:3728-3816: traction integral over boundary id 3
 * (top edge), undegraded stress, load_x already multiplied by -1 (3789). */
int pf_load (pf_ctx *ctx, double *load_x, double *load_y);
/* set_initial_bc() for arbitrary Dirichlet data (cracks.cc:2700-2707): writes values[dof] (host, block
 * layout) into the solution on every displacement dof whose Dirichlet bit is set (pf_set_constraints). */
int pf_set_dirichlet_values (pf_ctx *ctx, const double *values);
/* compute_load() on a forest mesh (see pf_create_forest): `cells` lists the cells whose top
 * edge lies on boundary id 3. */
int pf_load_cells (pf_ctx *ctx, const int64_t *cells, int64_t n_cells, double *load_x, double *load_y);
/* min over the owned phase-field dofs: the indicator refine_mesh() tests
 * against "value phase field for refinement" (cracks.cc:3971-3995).  [collective] */
int pf_phase_field_min (pf_ctx *ctx, double *phi_min);
/* time-step bookkeeping on the device: oldold <- old <- sol (cracks.cc:4302-4303);
 * returns ||old - sol||_inf of the step just finished (4478-4483) */
int pf_advance_timestep (pf_ctx *ctx);
int pf_timestep_difference (pf_ctx *ctx, double *linfty);
int pf_restore_old_solution (pf_ctx *ctx);                  /* solution = old_solution, cracks.cc:4340 */
/* time-step sizes of the extrapolation, use_old_timestep_pf and Pressure(time)
 * without touching the device-resident vectors (cracks.cc:2145, 2268-2277) */
int pf_set_time_parameters (pf_ctx *ctx, double dt_old, double dt_oldold, int use_old_timestep_pf,
                            double pressure);
/* line search bookkeeping on the device (cracks.cc:2922, 2955-2956) */
int pf_save_solution (pf_ctx *ctx);            /* saved_solution = solution */
int pf_restore_saved_solution (pf_ctx *ctx);   /* solution = saved_solution */
int pf_scale_update (pf_ctx *ctx, double factor); /* newton_update *= line_search_damping */

/* ---- device-resident helpers for benchmarks / device-side solvers -------- */
int pf_device_vector (pf_ctx *ctx, double **out);            /* local slab, internal layout */
int pf_device_vector_free (pf_ctx *ctx, double *v);
int pf_upload (pf_ctx *ctx, const double *host_block, double *dev);   /* block -> internal */
int pf_download (pf_ctx *ctx, const double *dev, double *host_block); /* internal -> block */
/* number of kernel launches issued on this context so far */
int64_t pf_launch_count (const pf_ctx *ctx);
/* CUDA-event timing of the dominant kernel (tiled 3-D apply) on the context's
 * stream: enable, run, then read the summed milliseconds and launch count */
int pf_profile_enable (pf_ctx *ctx, int on);
int pf_profile_read (pf_ctx *ctx, double *total_ms, int64_t *count);
/* pinned host memory for callers that want full PCIe bandwidth */
int pf_host_alloc (void **out, size_t bytes);
int pf_host_free (void *p);
/* testing aids, per context (every level of its multigrid hierarchy follows):
 * route the 3-D apply / residual through the dimension-generic thread-per-cell kernels (the independent second
 * implementation the parity tests compare the tiled kernels with) */
int pf_debug_force_generic (pf_ctx *ctx, int on);
/* do not use the cubic-cell (isotropic scale) specialisation of the tiled apply */
int pf_debug_disable_iso (pf_ctx *ctx, int on);
/* select a tuning variant of the tiled 3-D apply kernel; 16 = the default.  Other numbers exist only in a library
 * built with `make TUNING=1` (-DPF_TUNING_VARIANTS) and return PF_UNSUPPORTED otherwise */
int pf_debug_set_variant (pf_ctx *ctx, int variant);

#ifdef __cplusplus
}
#endif
#endif /* CRACKS_B200_H */
