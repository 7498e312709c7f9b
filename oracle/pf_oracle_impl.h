/*
 * pf_oracle_impl.h -- body of the CPU oracle, compiled once per DIM (2 and 3).
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * hot path (tjhei/cracks, cracks.cc) used as the checker for the CUDA path
 * and as the timed CPU baseline.  Nothing under cracks_b200/ may call it.
 *
 * Included by pf_oracle.c with DIM and SUF(name) defined.
 *
 * Conventions restated from deal.II as used by the reference:
 *   - Q1 elements, vertex order lexicographic with x fastest,
 *     local dof i = vertex*(DIM+1) + component (FESystem(FE_Q(1)^dim, FE_Q(1)),
 *     cracks.cc:980-983),
 *   - QGauss(3) tensor rule on [0,1]^dim, x fastest (cracks.cc:2156),
 *   - uniform box mesh from subdivided_hyper_rectangle (cracks.cc:1248-1253),
 *     cells and nodes numbered lexicographically with x fastest,
 *   - vectors are node-major interleaved: dof = node*(DIM+1) + component.
 */

#define NC (DIM + 1)               /* components per node: u_0..u_{dim-1}, phi */
#define NV (1 << DIM)              /* vertices per cell */
#define NDPC (NV * NC)             /* dofs per cell: 12 (2D) / 32 (3D) */
#define NQ1 3                      /* Gauss points per direction = fe.degree+2 */
#if DIM == 2
#define NQ 9
#else
#define NQ 27
#endif

/* shape function tables on the unit cell for one (hx,hy,hz) */
typedef struct
{
  double N[NQ][NV];         /* shape value of vertex v at q */
  double dN[NQ][NV][DIM];   /* physical gradient of vertex v at q */
  double JxW[NQ];
} SUF (fe_tab);

static void SUF (fe_init) (SUF (fe_tab) * t, const double *h)
{
  const double g = 0.5 * sqrt (3.0 / 5.0);
  const double xi[3] = { 0.5 - g, 0.5, 0.5 + g };
  const double w[3] = { 5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0 };
  double vol = 1.0;
  for (int d = 0; d < DIM; ++d)
    vol *= h[d];
  for (int q = 0; q < NQ; ++q)
    {
      int qi[3] = { q % 3, (q / 3) % 3, q / 9 };
      double wq = vol;
      for (int d = 0; d < DIM; ++d)
        wq *= w[qi[d]];
      t->JxW[q] = wq;
      for (int v = 0; v < NV; ++v)
        {
          double val = 1.0;
          for (int d = 0; d < DIM; ++d)
            {
              const int b = (v >> d) & 1;
              val *= b ? xi[qi[d]] : 1.0 - xi[qi[d]];
            }
          t->N[q][v] = val;
          for (int e = 0; e < DIM; ++e)
            {
              double gr = 1.0;
              for (int d = 0; d < DIM; ++d)
                {
                  const int b = (v >> d) & 1;
                  if (d == e)
                    gr *= (b ? 1.0 : -1.0) / h[d];
                  else
                    gr *= b ? xi[qi[d]] : 1.0 - xi[qi[d]];
                }
              t->dN[q][v][e] = gr;
            }
        }
    }
}

static inline long SUF (n_nodes) (const pfo_mesh * m)
{
  long n = 1;
  for (int d = 0; d < DIM; ++d)
    n *= (long) m->n[d] + 1;
  return n;
}

static inline long SUF (n_cells) (const pfo_mesh * m)
{
  long n = 1;
  for (int d = 0; d < DIM; ++d)
    n *= (long) m->n[d];
  return n;
}

/* global node numbers of the 2^dim vertices of cell c */
static inline void SUF (cell_nodes) (const pfo_mesh * m, long c, long *nodes)
{
  long ci[3] = { 0, 0, 0 };
  long rem = c;
  for (int d = 0; d < DIM; ++d)
    {
      ci[d] = rem % m->n[d];
      rem /= m->n[d];
    }
  const long sx = 1, sy = (long) m->n[0] + 1;
#if DIM == 3
  const long sz = sy * ((long) m->n[1] + 1);
#endif
  for (int v = 0; v < NV; ++v)
    {
      long id = (ci[0] + (v & 1)) * sx + (ci[1] + ((v >> 1) & 1)) * sy;
#if DIM == 3
      id += (ci[2] + ((v >> 2) & 1)) * sz;
#endif
      nodes[v] = id;
    }
}

/* quadrature-point state shared by residual and Jacobian; restates
 * cracks.cc:2222-2306 (no stress split: sigma+ = sigma, sigma- = 0). */
typedef struct
{
  double pf, pf_extra, grad_pf[DIM];
  double grad_u[DIM][DIM], E[DIM][DIM], tr_E, div_u;
  double sp[DIM][DIM];          /* stress_term_plus */
} SUF (qstate);

static void
SUF (eval_state) (const SUF (fe_tab) * t, int q, const pfo_params * p,
                  const double *lsol, const double *lold, const double *loo,
                  SUF (qstate) * s)
{
  double pf = 0, old_pf = 0, oo_pf = 0;
  memset (s, 0, sizeof (*s));
  for (int v = 0; v < NV; ++v)
    {
      const double N = t->N[q][v];
      pf += N * lsol[v * NC + DIM];
      old_pf += N * lold[v * NC + DIM];
      oo_pf += N * loo[v * NC + DIM];
      for (int e = 0; e < DIM; ++e)
        {
          s->grad_pf[e] += t->dN[q][v][e] * lsol[v * NC + DIM];
          for (int c = 0; c < DIM; ++c)
            s->grad_u[c][e] += t->dN[q][v][e] * lsol[v * NC + c];
        }
    }
  s->pf = pf;
  /* extrapolation, cracks.cc:2268-2277.  The reference writes the factor as
   * (time-(time-dt_o-dt_oo)) / (time-dt_o-(time-dt_o-dt_oo)) = (dt_o+dt_oo)/dt_oo */
  double pf_extra = oo_pf + (p->dt_old + p->dt_oldold) / p->dt_oldold * (old_pf - oo_pf);
  if (pf_extra <= 0.0)
    pf_extra = 0.0;
  if (pf_extra >= 1.0)
    pf_extra = 1.0;
  if (p->use_old_timestep_pf)
    pf_extra = old_pf;
  s->pf_extra = pf_extra;
  s->tr_E = 0;
  for (int a = 0; a < DIM; ++a)
    {
      for (int b = 0; b < DIM; ++b)
        s->E[a][b] = 0.5 * (s->grad_u[a][b] + s->grad_u[b][a]);
      s->tr_E += s->E[a][a];
    }
  s->div_u = s->tr_E;            /* Tensors::get_divergence_u, cracks.cc:330-349 */
  for (int a = 0; a < DIM; ++a)
    for (int b = 0; b < DIM; ++b)
      s->sp[a][b] = (a == b ? p->lambda * s->tr_E : 0.0) + 2 * p->mu * s->E[a][b];
}

static void
SUF (gather) (const long *nodes, const double *vec, double *loc)
{
  for (int v = 0; v < NV; ++v)
    for (int c = 0; c < NC; ++c)
      loc[v * NC + c] = vec[nodes[v] * NC + c];
}

/* local_rhs of cracks.cc:2393-2432 (sign as in the reference: rhs = -F) */
static void
SUF (cell_rhs) (const SUF (fe_tab) * t, const pfo_params * p,
                const double *lsol, const double *lold, const double *loo,
                double *local_rhs)
{
  const double k = p->kappa, Gc = p->G_c, eps = p->eps;
  const double pr = p->pressure, ab = p->alpha_biot;
  for (int i = 0; i < NDPC; ++i)
    local_rhs[i] = 0;
  for (int q = 0; q < NQ; ++q)
    {
      SUF (qstate) s;
      SUF (eval_state) (t, q, p, lsol, lold, loo, &s);
      double spE = 0;
      for (int a = 0; a < DIM; ++a)
        for (int b = 0; b < DIM; ++b)
          spE += s.sp[a][b] * s.E[a][b];
      const double g = (1.0 - k) * s.pf_extra * s.pf_extra + k;
      for (int v = 0; v < NV; ++v)
        {
          for (int c = 0; c < DIM; ++c)
            {
              /* phi_i_grads_u has only row c non-zero: (grad N_v) */
              double sc = 0;
              for (int e = 0; e < DIM; ++e)
                sc += g * s.sp[c][e] * t->dN[q][v][e];
              const double div_lin = t->dN[q][v][c];
              local_rhs[v * NC + c] -=
                (sc - (ab - 1.0) * pr * s.pf_extra * s.pf_extra * div_lin) * t->JxW[q];
            }
          const double N = t->N[q][v];
          double gg = 0;
          for (int e = 0; e < DIM; ++e)
            gg += s.grad_pf[e] * t->dN[q][v][e];
          local_rhs[v * NC + DIM] -=
            ((1.0 - k) * spE * s.pf * N - Gc / eps * (1.0 - s.pf) * N + Gc * eps * gg
             - 2.0 * (ab - 1.0) * pr * s.pf * s.div_u * N) * t->JxW[q];
        }
    }
}

/* local_matrix(j,i) of cracks.cc:2308-2389, loop-for-loop (q, i, j). */
static void
SUF (cell_matrix) (const SUF (fe_tab) * t, const pfo_params * p,
                   const double *lsol, const double *lold, const double *loo,
                   double *local_matrix /* [NDPC][NDPC], row = test j */ )
{
  const double k = p->kappa, Gc = p->G_c, eps = p->eps;
  const double pr = p->pressure, ab = p->alpha_biot;
  for (int i = 0; i < NDPC * NDPC; ++i)
    local_matrix[i] = 0;
  for (int q = 0; q < NQ; ++q)
    {
      SUF (qstate) s;
      SUF (eval_state) (t, q, p, lsol, lold, loo, &s);
      double spE = 0;
      for (int a = 0; a < DIM; ++a)
        for (int b = 0; b < DIM; ++b)
          spE += s.sp[a][b] * s.E[a][b];
      const double g = (1.0 - k) * s.pf_extra * s.pf_extra + k;
      for (int i = 0; i < NDPC; ++i)
        {
          const int vi = i / NC, ci = i % NC;
          /* trial function i: value / gradient per extractor */
          double gu[DIM][DIM];   /* phi_i_grads_u */
          double Npf = 0, gpf[DIM];
          memset (gu, 0, sizeof (gu));
          for (int e = 0; e < DIM; ++e)
            gpf[e] = 0;
          if (ci < DIM)
            for (int e = 0; e < DIM; ++e)
              gu[ci][e] = t->dN[q][vi][e];
          else
            {
              Npf = t->N[q][vi];
              for (int e = 0; e < DIM; ++e)
                gpf[e] = t->dN[q][vi][e];
            }
          double EL[DIM][DIM], trEL = 0, divL = 0;
          for (int a = 0; a < DIM; ++a)
            {
              for (int b = 0; b < DIM; ++b)
                EL[a][b] = 0.5 * (gu[a][b] + gu[b][a]);
              trEL += EL[a][a];
              divL += gu[a][a];
            }
          double spL[DIM][DIM];  /* stress_term_plus_LinU */
          for (int a = 0; a < DIM; ++a)
            for (int b = 0; b < DIM; ++b)
              spL[a][b] = (ci < DIM) ? ((a == b ? p->lambda * trEL : 0.0) + 2 * p->mu * EL[a][b]) : 0.0;
          double spL_E = 0, sp_EL = 0;
          for (int a = 0; a < DIM; ++a)
            for (int b = 0; b < DIM; ++b)
              {
                spL_E += spL[a][b] * s.E[a][b];
                sp_EL += s.sp[a][b] * EL[a][b];
              }
          for (int j = 0; j < NDPC; ++j)
            {
              const int vj = j / NC, cj = j % NC;
              if (cj < DIM)
                {
                  double sc = 0;
                  for (int e = 0; e < DIM; ++e)
                    sc += g * spL[cj][e] * t->dN[q][vj][e];
                  local_matrix[j * NDPC + i] += 1.0 * sc * t->JxW[q];
                }
              else
                {
                  const double Nj = t->N[q][vj];
                  double gg = 0;
                  for (int e = 0; e < DIM; ++e)
                    gg += gpf[e] * t->dN[q][vj][e];
                  local_matrix[j * NDPC + i] +=
                    ((1 - k) * (spL_E + sp_EL) * s.pf * Nj
                     + (1 - k) * spE * Npf * Nj
                     + Gc / eps * Npf * Nj
                     + Gc * eps * gg
                     - 2.0 * (ab - 1.0) * pr * (s.pf * divL + Npf * s.div_u) * Nj) * t->JxW[q];
                }
            }
        }
    }
}

/* ---- residual: cracks.cc:2129-2475 with residual_only = true ------------
 * r_total = hanging-node-only distribution (none on uniform meshes)
 * r_pde   = constraints_update distribution: constrained rows dropped      */
void
SUF (pfo_residual) (const pfo_mesh * m, const pfo_params * p,
                    const double *sol, const double *old, const double *oldold,
                    const unsigned char *constrained, double *r_pde, double *r_total)
{
  SUF (fe_tab) t;
  SUF (fe_init) (&t, m->h);
  const long nn = SUF (n_nodes) (m), ncell = SUF (n_cells) (m);
  for (long i = 0; i < nn * NC; ++i)
    r_total[i] = 0;
  for (int colour = 0; colour < NV; ++colour)
    {
#pragma omp parallel for schedule(static)
      for (long c = 0; c < ncell; ++c)
        {
          long ci[3] = { 0, 0, 0 }, rem = c;
          for (int d = 0; d < DIM; ++d)
            {
              ci[d] = rem % m->n[d];
              rem /= m->n[d];
            }
          int mycol = 0;
          for (int d = 0; d < DIM; ++d)
            mycol |= (int) (ci[d] & 1) << d;
          if (mycol != colour)
            continue;
          long nodes[NV];
          double ls[NDPC], lo[NDPC], loo[NDPC], rhs[NDPC];
          SUF (cell_nodes) (m, c, nodes);
          SUF (gather) (nodes, sol, ls);
          SUF (gather) (nodes, old, lo);
          SUF (gather) (nodes, oldold, loo);
          SUF (cell_rhs) (&t, p, ls, lo, loo, rhs);
          for (int v = 0; v < NV; ++v)
            for (int cc = 0; cc < NC; ++cc)
              r_total[nodes[v] * NC + cc] += rhs[v * NC + cc];
        }
    }
  if (r_pde)
    for (long i = 0; i < nn * NC; ++i)
      r_pde[i] = (constrained && constrained[i]) ? 0.0 : r_total[i];
}

/* ---- CSR pattern of the Q1 vector-valued stencil (full coupling) --------- */
static inline int
SUF (neighbours) (const pfo_mesh * m, long node, long *nb)
{
  long ni[3] = { 0, 0, 0 }, dims[3] = { 1, 1, 1 }, rem = node;
  for (int d = 0; d < DIM; ++d)
    {
      dims[d] = (long) m->n[d] + 1;
      ni[d] = rem % dims[d];
      rem /= dims[d];
    }
  int cnt = 0;
#if DIM == 3
  for (long dz = -1; dz <= 1; ++dz)
#else
  const long dz = 0;
#endif
    for (long dy = -1; dy <= 1; ++dy)
      for (long dx = -1; dx <= 1; ++dx)
        {
          const long x = ni[0] + dx, y = ni[1] + dy, z = ni[2] + dz;
          if (x < 0 || x >= dims[0] || y < 0 || y >= dims[1] || z < 0 || z >= dims[2])
            continue;
          nb[cnt++] = x + dims[0] * (y + dims[1] * z);
        }
  return cnt;
}

long
SUF (pfo_csr_nnz) (const pfo_mesh * m)
{
  const long nn = SUF (n_nodes) (m);
  long nnz = 0;
  for (long n = 0; n < nn; ++n)
    {
      long nb[27];
      nnz += (long) SUF (neighbours) (m, n, nb) * NC * NC;
    }
  return nnz;
}

void
SUF (pfo_csr_pattern) (const pfo_mesh * m, long *rowptr, int *col)
{
  const long nn = SUF (n_nodes) (m);
  long pos = 0;
  for (long n = 0; n < nn; ++n)
    {
      long nb[27];
      const int cnt = SUF (neighbours) (m, n, nb);
      for (int c = 0; c < NC; ++c)
        {
          rowptr[n * NC + c] = pos;
          for (int k = 0; k < cnt; ++k)
            for (int cc = 0; cc < NC; ++cc)
              col[pos++] = (int) (nb[k] * NC + cc);
        }
    }
  rowptr[nn * NC] = pos;
}

static inline long
SUF (csr_find) (const long *rowptr, const int *col, long row, long c)
{
  long lo = rowptr[row], hi = rowptr[row + 1] - 1;
  while (lo <= hi)
    {
      const long mid = (lo + hi) / 2;
      if (col[mid] == c)
        return mid;
      if (col[mid] < c)
        lo = mid + 1;
      else
        hi = mid - 1;
    }
  return -1;
}

/* ---- Jacobian assembly into CSR: cracks.cc:2200-2468 + the semantics of
 * AffineConstraints::distribute_local_to_global for homogeneous constraints
 * without entries (Dirichlet, active set): constrained rows and columns are
 * eliminated, and each cell adds |local(i,i)| (or the cell's average |diag|
 * if that is zero) to the diagonal of a constrained dof.  This library
 * behaviour (deal.II 9.5, affine_constraints.templates.h) is restated from
 * its documentation; the source is not in the container. */
void
SUF (pfo_assemble_jacobian) (const pfo_mesh * m, const pfo_params * p,
                             const double *sol, const double *old, const double *oldold,
                             const unsigned char *constrained,
                             const long *rowptr, const int *col, double *val)
{
  SUF (fe_tab) t;
  SUF (fe_init) (&t, m->h);
  const long ncell = SUF (n_cells) (m);
  const long nn = SUF (n_nodes) (m);
  for (long i = 0; i < rowptr[nn * NC]; ++i)
    val[i] = 0;
  /* 2^dim colours so that OpenMP threads never touch the same row */
  for (int colour = 0; colour < NV; ++colour)
    {
#pragma omp parallel for schedule(static)
      for (long c = 0; c < ncell; ++c)
        {
          long ci[3] = { 0, 0, 0 }, rem = c;
          for (int d = 0; d < DIM; ++d)
            {
              ci[d] = rem % m->n[d];
              rem /= m->n[d];
            }
          int mycol = 0;
          for (int d = 0; d < DIM; ++d)
            mycol |= (int) (ci[d] & 1) << d;
          if (mycol != colour)
            continue;
          long nodes[NV];
          double ls[NDPC], lo[NDPC], loo[NDPC], M[NDPC * NDPC];
          SUF (cell_nodes) (m, c, nodes);
          SUF (gather) (nodes, sol, ls);
          SUF (gather) (nodes, old, lo);
          SUF (gather) (nodes, oldold, loo);
          SUF (cell_matrix) (&t, p, ls, lo, loo, M);
          double avg = 0;
          for (int i = 0; i < NDPC; ++i)
            avg += fabs (M[i * NDPC + i]);
          avg /= NDPC;
          for (int j = 0; j < NDPC; ++j)
            {
              const long row = nodes[j / NC] * NC + j % NC;
              const int crow = constrained && constrained[row];
              if (crow)
                {
                  const double d = fabs (M[j * NDPC + j]);
                  val[SUF (csr_find) (rowptr, col, row, row)] += (d != 0 ? d : avg);
                  continue;
                }
              for (int i = 0; i < NDPC; ++i)
                {
                  const long cc = nodes[i / NC] * NC + i % NC;
                  if (constrained && constrained[cc])
                    continue;
                  val[SUF (csr_find) (rowptr, col, row, cc)] += M[j * NDPC + i];
                }
            }
        }
    }
}

/* matrix-free CPU apply with the same constraint semantics; used to cross-check
 * the CSR path and as an oracle at sizes where the CSR would not fit. */
void
SUF (pfo_apply_jacobian) (const pfo_mesh * m, const pfo_params * p,
                          const double *sol, const double *old, const double *oldold,
                          const unsigned char *constrained, const double *x, double *y)
{
  SUF (fe_tab) t;
  SUF (fe_init) (&t, m->h);
  const long ncell = SUF (n_cells) (m);
  const long nn = SUF (n_nodes) (m);
  for (long i = 0; i < nn * NC; ++i)
    y[i] = 0;
  for (int colour = 0; colour < NV; ++colour)
    {
#pragma omp parallel for schedule(static)
      for (long c = 0; c < ncell; ++c)
        {
          long ci[3] = { 0, 0, 0 }, rem = c;
          for (int d = 0; d < DIM; ++d)
            {
              ci[d] = rem % m->n[d];
              rem /= m->n[d];
            }
          int mycol = 0;
          for (int d = 0; d < DIM; ++d)
            mycol |= (int) (ci[d] & 1) << d;
          if (mycol != colour)
            continue;
          long nodes[NV];
          double ls[NDPC], lo[NDPC], loo[NDPC], lx[NDPC], M[NDPC * NDPC];
          SUF (cell_nodes) (m, c, nodes);
          SUF (gather) (nodes, sol, ls);
          SUF (gather) (nodes, old, lo);
          SUF (gather) (nodes, oldold, loo);
          SUF (gather) (nodes, x, lx);
          SUF (cell_matrix) (&t, p, ls, lo, loo, M);
          double avg = 0;
          for (int i = 0; i < NDPC; ++i)
            avg += fabs (M[i * NDPC + i]);
          avg /= NDPC;
          for (int j = 0; j < NDPC; ++j)
            {
              const long row = nodes[j / NC] * NC + j % NC;
              if (constrained && constrained[row])
                {
                  const double d = fabs (M[j * NDPC + j]);
                  y[row] += (d != 0 ? d : avg) * lx[j];
                  continue;
                }
              double acc = 0;
              for (int i = 0; i < NDPC; ++i)
                {
                  const long cc = nodes[i / NC] * NC + i % NC;
                  if (constrained && constrained[cc])
                    continue;
                  acc += M[j * NDPC + i] * lx[i];
                }
              y[row] += acc;
            }
        }
    }
}

/* ---- phi-block lumped mass, cracks.cc:2514-2562: QGaussLobatto(2) = vertex
 * rule, so the local entry of vertex v is N_v(v)^2 * JxW = vol / 2^dim. */
void
SUF (pfo_lumped_mass) (const pfo_mesh * m, double *mass /* per node */ )
{
  const long nn = SUF (n_nodes) (m), ncell = SUF (n_cells) (m);
  double vol = 1;
  for (int d = 0; d < DIM; ++d)
    vol *= m->h[d];
  for (long i = 0; i < nn; ++i)
    mass[i] = 0;
  for (long c = 0; c < ncell; ++c)
    {
      long nodes[NV];
      SUF (cell_nodes) (m, c, nodes);
      for (int v = 0; v < NV; ++v)
        mass[nodes[v]] += 1.0 * 1.0 * (vol / NV);
    }
}

/* ---- energies, cracks.cc:3615-3701 (note (1+kappa), line 3677) ---------- */
void
SUF (pfo_energy) (const pfo_mesh * m, const pfo_params * p, const double *sol,
                  double *bulk, double *crack)
{
  SUF (fe_tab) t;
  SUF (fe_init) (&t, m->h);
  const long ncell = SUF (n_cells) (m);
  double eb = 0, ec = 0;
  for (long c = 0; c < ncell; ++c)
    {
      long nodes[NV];
      double ls[NDPC];
      SUF (cell_nodes) (m, c, nodes);
      SUF (gather) (nodes, sol, ls);
      for (int q = 0; q < NQ; ++q)
        {
          double pf = 0, gpf[DIM], gu[DIM][DIM];
          memset (gpf, 0, sizeof (gpf));
          memset (gu, 0, sizeof (gu));
          for (int v = 0; v < NV; ++v)
            {
              pf += t.N[q][v] * ls[v * NC + DIM];
              for (int e = 0; e < DIM; ++e)
                {
                  gpf[e] += t.dN[q][v][e] * ls[v * NC + DIM];
                  for (int cc = 0; cc < DIM; ++cc)
                    gu[cc][e] += t.dN[q][v][e] * ls[v * NC + cc];
                }
            }
          double E[DIM][DIM], trE = 0, trE2 = 0, gg = 0;
          for (int a = 0; a < DIM; ++a)
            {
              for (int b = 0; b < DIM; ++b)
                E[a][b] = 0.5 * (gu[a][b] + gu[b][a]);
              trE += E[a][a];
              gg += gpf[a] * gpf[a];
            }
          for (int a = 0; a < DIM; ++a)
            for (int b = 0; b < DIM; ++b)
              trE2 += E[a][b] * E[b][a];
          const double psi = 0.5 * p->lambda * trE * trE + p->mu * trE2;
          eb += ((1 + p->kappa) * pf * pf + p->kappa) * psi * t.JxW[q];
          ec += p->G_c / 2.0 * ((pf - 1) * (pf - 1) / p->eps + p->eps * gg) * t.JxW[q];
        }
    }
  *bulk = eb;
  *crack = ec;
}

/* ---- total crack volume, cracks.cc:3553-3589 ----------------------------- */
double
SUF (pfo_tcv) (const pfo_mesh * m, const double *sol)
{
  SUF (fe_tab) t;
  SUF (fe_init) (&t, m->h);
  const long ncell = SUF (n_cells) (m);
  double tcv = 0;
  for (long c = 0; c < ncell; ++c)
    {
      long nodes[NV];
      double ls[NDPC];
      SUF (cell_nodes) (m, c, nodes);
      SUF (gather) (nodes, sol, ls);
      for (int q = 0; q < NQ; ++q)
        {
          double u[DIM], gpf[DIM];
          memset (u, 0, sizeof (u));
          memset (gpf, 0, sizeof (gpf));
          for (int v = 0; v < NV; ++v)
            for (int e = 0; e < DIM; ++e)
              {
                u[e] += t.N[q][v] * ls[v * NC + e];
                gpf[e] += t.dN[q][v][e] * ls[v * NC + DIM];
              }
          double dot = 0;
          for (int e = 0; e < DIM; ++e)
            dot += u[e] * gpf[e];
          tcv += dot * t.JxW[q];
        }
    }
  return tcv;
}

/* ---- primal-dual active set update, cracks.cc:2822-2899 (uniform mesh: no
 * hanging nodes).  active[node] is rewritten; sol's phi is reset to old on the
 * new active set.  cycle[node] counts inactive->... transitions (2901-2907);
 * the caller updates it.  Returns the number of active dofs. */
long
SUF (pfo_active_set) (const pfo_mesh * m, double c_scale, const double *r_total,
                      const double *mass, const double *old, double *sol,
                      const int *cycle, unsigned char *active, long *n_cycling)
{
  const long nn = SUF (n_nodes) (m);
  long cnt = 0, ncyc = 0;
  for (long n = 0; n < nn; ++n)
    {
      const double old_value = old[n * NC + DIM];
      const double new_value = sol[n * NC + DIM];
      const double gap = new_value - old_value;
      const int cyc = cycle ? cycle[n] : 0;
      active[n] = 0;
      if (r_total[n * NC + DIM] / mass[n] + c_scale * gap <= 0.0 && cyc < 5)
        continue;
      if (cyc >= 5)
        ++ncyc;
      sol[n * NC + DIM] = old_value;
      active[n] = 1;
      ++cnt;
    }
  if (n_cycling)
    *n_cycling = ncyc;
  return cnt;
}

/* ---- crack opening displacement on the plane x = eval_line,
 * cracks.cc:3452-3549: over every cell face lying on that plane (visited from
 * both neighbouring cells, hence the final / 2), QGauss<dim-1>(3),
 * sum of 0.5 * u . grad(phi) * JxW.  *n_faces counts the visited faces. */
double
SUF (pfo_cod) (const pfo_mesh * m, const double *sol, double eval_line, long *n_faces)
{
  const double gq = 0.5 * sqrt (3.0 / 5.0);
  const double xi[3] = { 0.5 - gq, 0.5, 0.5 + gq };
  const double w[3] = { 5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0 };
  const long ncell = SUF (n_cells) (m);
  double cod = 0;
  long faces = 0;
  for (long c = 0; c < ncell; ++c)
    {
      const long cx = c % m->n[0];
      for (int side = 0; side < 2; ++side)
        {
          const double fx = m->origin[0] + (cx + side) * m->h[0];
          if (!(fx < eval_line + 1e-8 && fx > eval_line - 1e-8))
            continue;
          ++faces;
          long nodes[NV];
          double ls[NDPC];
          SUF (cell_nodes) (m, c, nodes);
          SUF (gather) (nodes, sol, ls);
#if DIM == 3
          for (int qz = 0; qz < 3; ++qz)
#else
          const int qz = 0;
#endif
            for (int qy = 0; qy < 3; ++qy)
              {
                const double pt[3] = { (double) side, xi[qy], xi[qz] };
                double JxW = m->h[1] * w[qy];
#if DIM == 3
                JxW *= m->h[2] * w[qz];
#endif
                double u[DIM], gpf[DIM];
                memset (u, 0, sizeof (u));
                memset (gpf, 0, sizeof (gpf));
                for (int v = 0; v < NV; ++v)
                  {
                    double N = 1.0;
                    for (int d = 0; d < DIM; ++d)
                      N *= ((v >> d) & 1) ? pt[d] : 1.0 - pt[d];
                    for (int e = 0; e < DIM; ++e)
                      {
                        double gr = 1.0;
                        for (int d = 0; d < DIM; ++d)
                          {
                            const int b = (v >> d) & 1;
                            gr *= (d == e) ? (b ? 1.0 : -1.0) / m->h[d] : (b ? pt[d] : 1.0 - pt[d]);
                          }
                        u[e] += N * ls[v * NC + e];
                        gpf[e] += gr * ls[v * NC + DIM];
                      }
                  }
                double dot = 0;
                for (int e = 0; e < DIM; ++e)
                  dot += u[e] * gpf[e];
                cod += 0.5 * dot * JxW;
              }
        }
    }
  if (n_faces)
    *n_faces = faces;
  return cod / 2.0;
}

void
SUF (pfo_spmv) (long nrows, const long *rowptr, const int *col, const double *val,
                const double *x, double *y)
{
#pragma omp parallel for schedule(static)
  for (long r = 0; r < nrows; ++r)
    {
      double acc = 0;
      for (long k = rowptr[r]; k < rowptr[r + 1]; ++k)
        acc += val[k] * x[col[k]];
      y[r] = acc;
    }
}

#undef NC
#undef NV
#undef NDPC
#undef NQ1
#undef NQ
