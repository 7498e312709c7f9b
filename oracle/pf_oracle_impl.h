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
#if DIM == 2
  if (m->slit)
    n += m->n[0] / 2;           /* the doubled nodes of the slit, appended */
#endif
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
#if DIM == 2
  /* unit_slit.inp (see pf_oracle.h): the row of cells above the slit line uses the doubled nodes */
  if (m->slit && ci[1] == m->n[1] / 2)
    for (int v = 0; v < 2; ++v)
      {
        const long ix = ci[0] + (v & 1);
        if (ix > m->n[0] / 2)
          nodes[v] = ((long) m->n[0] + 1) * ((long) m->n[1] + 1) + (ix - m->n[0] / 2 - 1);
      }
#endif
}

long
SUF (pfo_n_nodes) (const pfo_mesh * m)
{
  return SUF (n_nodes) (m);
}

void
SUF (pfo_cell_nodes) (const pfo_mesh * m, long *cells)
{
  const long ncell = SUF (n_cells) (m);
  for (long c = 0; c < ncell; ++c)
    SUF (cell_nodes) (m, c, cells + c * NV);
}

void
SUF (pfo_node_coords) (const pfo_mesh * m, double *xyz)
{
  const long ncell = SUF (n_cells) (m);
  for (long c = 0; c < ncell; ++c)
    {
      long nodes[NV], ci[3] = { 0, 0, 0 }, rem = c;
      for (int d = 0; d < DIM; ++d)
        {
          ci[d] = rem % m->n[d];
          rem /= m->n[d];
        }
      SUF (cell_nodes) (m, c, nodes);
      for (int v = 0; v < NV; ++v)
        for (int d = 0; d < DIM; ++d)
          xyz[nodes[v] * DIM + d] = m->origin[d] + m->h[d] * (double) (ci[d] + ((v >> d) & 1));
    }
}

#if DIM == 2
/* ---- closed-form 2x2 symmetric eigen-decomposition, cracks.cc:1691-1737 ----
 * P = [v1 v2] (columns).  The reference abort()s if the vectors are not
 * orthogonal (1732-1736); here P is filled with NaN in that case. */
void
pfo_eigen_2x2 (const double *m, double *ev1, double *ev2, double *P)
{
  const double m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
  double v1[2], v2[2];
  if (fabs (m01) < 1e-10 * fabs (m00) || fabs (m01) < 1e-10 * fabs (m11))
    {
      *ev1 = m00;
      v1[0] = 1;
      v1[1] = 0;
      *ev2 = m11;
      v2[0] = 0;
      v2[1] = 1;
    }
  else
    {
      const double sq = sqrt ((m00 - m11) * (m00 - m11) + 4.0 * m01 * m10);
      *ev1 = 0.5 * ((m00 + m11) + sq);
      *ev2 = 0.5 * ((m00 + m11) - sq);
      v1[0] = 1.0 / (sqrt (1 + (*ev1 - m00) / m01 * (*ev1 - m00) / m01));
      v1[1] = (*ev1 - m00) / (m01 * (sqrt (1 + (*ev1 - m00) / m01 * (*ev1 - m00) / m01)));
      v2[0] = 1.0 / (sqrt (1 + (*ev2 - m00) / m01 * (*ev2 - m00) / m01));
      v2[1] = (*ev2 - m00) / (m01 * (sqrt (1 + (*ev2 - m00) / m01 * (*ev2 - m00) / m01)));
    }
  P[0] = v1[0];
  P[1] = v2[0];
  P[2] = v1[1];
  P[3] = v2[1];
  if (v1[0] * v2[0] + v1[1] * v2[1] > 1.0e-6)
    P[0] = P[1] = P[2] = P[3] = NAN;
}

static inline void
SUF (mat2_mul3) (const double *A, const double *B, const double *C, int transpose_c, double *out)
{
  /* out = A * B * (transpose_c ? C^T : C), row-major 2x2 */
  double AB[4];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j)
      AB[i * 2 + j] = A[i * 2] * B[j] + A[i * 2 + 1] * B[2 + j];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j)
      out[i * 2 + j] = transpose_c ? AB[i * 2] * C[j * 2] + AB[i * 2 + 1] * C[j * 2 + 1]
                                   : AB[i * 2] * C[j] + AB[i * 2 + 1] * C[2 + j];
}

/* ---- Miehe tensile/compressive split and its linearisation, cracks.cc:1923-2120.
 * E, E_lin row-major 2x2.  The derivative branch divides by E[0][1] and by the
 * discriminant without guards, exactly like the reference (1992-2006). */
void
pfo_decompose_stress_2d (const double *E, const double *EL, double lambda, double mu, int derivative,
                         double *sp, double *sm)
{
  const double tr_E = E[0] + E[3], tr_EL = EL[0] + EL[3];
  double e1, e2, P[4];
  pfo_eigen_2x2 (E, &e1, &e2, P);
  const double e1p = fmax (0.0, e1), e2p = fmax (0.0, e2);
  const double Lp[4] = { e1p, 0.0, 0.0, e2p };
  if (!derivative)
    {
      double Ep[4];
      SUF (mat2_mul3) (P, Lp, P, 1, Ep);
      const double trp = fmax (0.0, tr_E);
      for (int i = 0; i < 4; ++i)
        {
          const double id = (i == 0 || i == 3) ? 1.0 : 0.0;
          sp[i] = lambda * trp * id + 2 * mu * Ep[i];
          sm[i] = lambda * (tr_E - trp) * id + 2 * mu * (E[i] - Ep[i]);
        }
      return;
    }
  const double E00 = E[0], E01 = E[1], E10 = E[2], E11 = E[3];
  const double L00 = EL[0], L01 = EL[1], L10 = EL[2], L11 = EL[3];
  const double disk = sqrt (E01 * E10 + (E00 - E11) * (E00 - E11) / 4.0);
  const double e1L = 0.5 * tr_EL + 1.0 / (2.0 * disk) * (L01 * E10 + E01 * L10 + (E00 - E11) * (L00 - L11) / 2.0);
  const double e2L = 0.5 * tr_EL - 1.0 / (2.0 * disk) * (L01 * E10 + E01 * L10 + (E00 - E11) * (L00 - L11) / 2.0);
  const double q1 = (e1 - E00) / E01, q2 = (e2 - E00) / E01;
  const double n1 = 1.0 / (sqrt (1 + q1 * q1)), n2 = 1.0 / (sqrt (1 + q2 * q2));
  const double dq1 = ((e1L - L00) * E01 - (e1 - E00) * L01) / (E01 * E01);
  const double dq2 = ((e2L - L00) * E01 - (e2 - E00) * L01) / (E01 * E01);
  const double n1L = -1.0 * (1.0 / (1.0 + q1 * q1) * 1.0 / (2.0 * sqrt (1.0 + q1 * q1)) * (2.0 * q1) * dq1);
  const double n2L = -1.0 * (1.0 / (1.0 + q2 * q2) * 1.0 / (2.0 * sqrt (1.0 + q2 * q2)) * (2.0 * q2) * dq2);
  /* product rule on normalisation and vector entries (2031-2063) */
  const double v1L[2] = { n1 * 0.0 + n1L * 1.0, n1 * dq1 + n1L * q1 };
  const double v2L[2] = { n2 * 0.0 + n2L * 1.0, n2 * dq2 + n2L * q2 };
  const double PL[4] = { v1L[0], v2L[0], v1L[1], v2L[1] };
  const double e1pL = (e1 < 0.0) ? 0.0 : e1L;   /* 2080-2094: keyed on the sign of the eigenvalue of E */
  const double e2pL = (e2 < 0.0) ? 0.0 : e2L;
  const double LpL[4] = { e1pL, 0.0, 0.0, e2pL };
  double T1[4], T2[4], T3[4], EpL[4];
  SUF (mat2_mul3) (PL, Lp, P, 1, T1);
  SUF (mat2_mul3) (P, LpL, P, 1, T2);
  SUF (mat2_mul3) (P, Lp, PL, 1, T3);
  for (int i = 0; i < 4; ++i)
    EpL[i] = T1[i] + T2[i] + T3[i];
  const double trpL = (tr_E < 0.0) ? 0.0 : tr_EL;
  for (int i = 0; i < 4; ++i)
    {
      const double id = (i == 0 || i == 3) ? 1.0 : 0.0;
      sp[i] = lambda * trpL * id + 2 * mu * EpL[i];
      sm[i] = lambda * (tr_EL - trpL) * id + 2 * mu * (EL[i] - EpL[i]);
    }
}
#endif

/* quadrature-point state shared by residual and Jacobian; restates
 * cracks.cc:2222-2306 (no stress split: sigma+ = sigma, sigma- = 0). */
typedef struct
{
  double pf, pf_extra, grad_pf[DIM];
  double grad_u[DIM][DIM], E[DIM][DIM], tr_E, div_u;
  double sp[DIM][DIM];          /* stress_term_plus */
  double sm[DIM][DIM];          /* stress_term_minus (zero without the split) */
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
#if DIM == 2
  if (p->split)                  /* cracks.cc:2294-2300 */
    {
      const double zero[4] = { 0, 0, 0, 0 };
      pfo_decompose_stress_2d (&s->E[0][0], zero, p->lambda, p->mu, 0, &s->sp[0][0], &s->sm[0][0]);
    }
#endif
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
              if (p->split)      /* + decompose_stress_rhs * sigma^- : grad(phi_i), cracks.cc:2407 */
                for (int e = 0; e < DIM; ++e)
                  sc += p->d_rhs * s.sm[c][e] * t->dN[q][v][e];
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
          double smL[DIM][DIM];  /* stress_term_minus_LinU */
          for (int a = 0; a < DIM; ++a)
            for (int b = 0; b < DIM; ++b)
              {
                spL[a][b] = (ci < DIM) ? ((a == b ? p->lambda * trEL : 0.0) + 2 * p->mu * EL[a][b]) : 0.0;
                smL[a][b] = 0.0;
              }
#if DIM == 2
          if (p->split && ci < DIM)   /* cracks.cc:2338-2345 */
            pfo_decompose_stress_2d (&s.E[0][0], &EL[0][0], p->lambda, p->mu, 1, &spL[0][0], &smL[0][0]);
#endif
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
                  if (p->split)  /* + decompose_stress_matrix * sigma^-_LinU : grad(phi_j), cracks.cc:2362 */
                    for (int e = 0; e < DIM; ++e)
                      sc += p->d_mat * smL[cj][e] * t->dN[q][vj][e];
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

#if DIM == 2
/* ---- load on boundary id 3 (the top edge y = max of unit_slit.inp),
 * cracks.cc:3728-3816: int sigma(u) n ds with QGauss<1>(3), sigma undegraded;
 * load[0] *= -1 (3789). */
void
pfo_load_2d (const pfo_mesh * m, const pfo_params * p, const double *sol, double *load)
{
  const double gq = 0.5 * sqrt (3.0 / 5.0);
  const double xi[3] = { 0.5 - gq, 0.5, 0.5 + gq };
  const double w[3] = { 5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0 };
  double lv[2] = { 0, 0 };
  const long cy = m->n[1] - 1;
  for (long cx = 0; cx < m->n[0]; ++cx)
    {
      long nodes[NV];
      double ls[NDPC];
      SUF (cell_nodes) (m, cx + cy * m->n[0], nodes);
      SUF (gather) (nodes, sol, ls);
      for (int q = 0; q < 3; ++q)
        {
          const double pt[2] = { xi[q], 1.0 };
          double gu[2][2] = { {0, 0}, {0, 0} };
          for (int v = 0; v < NV; ++v)
            for (int e = 0; e < 2; ++e)
              {
                double gr = 1.0;
                for (int d = 0; d < 2; ++d)
                  {
                    const int b = (v >> d) & 1;
                    gr *= (d == e) ? (b ? 1.0 : -1.0) / m->h[d] : (b ? pt[d] : 1.0 - pt[d]);
                  }
                for (int c = 0; c < 2; ++c)
                  gu[c][e] += gr * ls[v * NC + c];
              }
          const double E01 = 0.5 * (gu[0][1] + gu[1][0]), tr = gu[0][0] + gu[1][1];
          const double s01 = 2 * p->mu * E01, s11 = p->lambda * tr + 2 * p->mu * gu[1][1];
          const double JxW = m->h[0] * w[q];
          /* stress * normal with n = (0, 1) */
          lv[0] += s01 * JxW;
          lv[1] += s11 * JxW;
        }
    }
  load[0] = -1.0 * lv[0];
  load[1] = lv[1];
}
#endif

/* ---- general (locally refined) meshes: raw cell sums, constraints resolved by the caller ---- */
static inline pfo_params
SUF (cell_params) (const pfo_gmesh * m, const pfo_params * p, long c)
{
  pfo_params q = *p;
  if (m->cell_lame)
    {
      q.lambda = m->cell_lame[2 * c];
      q.mu = m->cell_lame[2 * c + 1];
    }
  return q;
}

void
SUF (pfo_g_residual) (const pfo_gmesh * m, const pfo_params * p, const double *sol, const double *old,
                      const double *oldold, double *r_raw)
{
  for (long i = 0; i < m->n_nodes * NC; ++i)
    r_raw[i] = 0;
  for (long c = 0; c < m->n_cells; ++c)
    {
      SUF (fe_tab) t;
      SUF (fe_init) (&t, m->cell_h + c * DIM);
      const long *nodes = m->cells + c * NV;
      double ls[NDPC], lo[NDPC], loo[NDPC], rhs[NDPC];
      SUF (gather) (nodes, sol, ls);
      SUF (gather) (nodes, old, lo);
      SUF (gather) (nodes, oldold, loo);
      const pfo_params pc = SUF (cell_params) (m, p, c);
      SUF (cell_rhs) (&t, &pc, ls, lo, loo, rhs);
      for (int v = 0; v < NV; ++v)
        for (int cc = 0; cc < NC; ++cc)
          r_raw[nodes[v] * NC + cc] += rhs[v * NC + cc];
    }
}

void
SUF (pfo_g_cell_matrices) (const pfo_gmesh * m, const pfo_params * p, const double *sol, const double *old,
                           const double *oldold, double *mats)
{
#pragma omp parallel for schedule(static)
  for (long c = 0; c < m->n_cells; ++c)
    {
      SUF (fe_tab) t;
      SUF (fe_init) (&t, m->cell_h + c * DIM);
      const long *nodes = m->cells + c * NV;
      double ls[NDPC], lo[NDPC], loo[NDPC];
      SUF (gather) (nodes, sol, ls);
      SUF (gather) (nodes, old, lo);
      SUF (gather) (nodes, oldold, loo);
      const pfo_params pc = SUF (cell_params) (m, p, c);
      SUF (cell_matrix) (&t, &pc, ls, lo, loo, mats + c * NDPC * NDPC);
    }
}

void
SUF (pfo_g_lumped_mass) (const pfo_gmesh * m, double *mass)
{
  for (long i = 0; i < m->n_nodes; ++i)
    mass[i] = 0;
  for (long c = 0; c < m->n_cells; ++c)
    {
      double vol = 1;
      for (int d = 0; d < DIM; ++d)
        vol *= m->cell_h[c * DIM + d];
      for (int v = 0; v < NV; ++v)
        mass[m->cells[c * NV + v]] += 1.0 * 1.0 * (vol / NV);
    }
}

/* bulk energy, crack energy (cracks.cc:3663-3681) and TCV (3585-3586) */
void
SUF (pfo_g_functionals) (const pfo_gmesh * m, const pfo_params * p, const double *sol, double *out)
{
  double eb = 0, ec = 0, tcv = 0;
  for (long c = 0; c < m->n_cells; ++c)
    {
      SUF (fe_tab) t;
      SUF (fe_init) (&t, m->cell_h + c * DIM);
      double ls[NDPC];
      SUF (gather) (m->cells + c * NV, sol, ls);
      const pfo_params pc = SUF (cell_params) (m, p, c);
      for (int q = 0; q < NQ; ++q)
        {
          double pf = 0, gpf[DIM], u[DIM], gu[DIM][DIM];
          memset (gpf, 0, sizeof (gpf));
          memset (u, 0, sizeof (u));
          memset (gu, 0, sizeof (gu));
          for (int v = 0; v < NV; ++v)
            {
              pf += t.N[q][v] * ls[v * NC + DIM];
              for (int e = 0; e < DIM; ++e)
                {
                  u[e] += t.N[q][v] * ls[v * NC + e];
                  gpf[e] += t.dN[q][v][e] * ls[v * NC + DIM];
                  for (int cc = 0; cc < DIM; ++cc)
                    gu[cc][e] += t.dN[q][v][e] * ls[v * NC + cc];
                }
            }
          double trE = 0, trE2 = 0, gg = 0, ug = 0;
          for (int a = 0; a < DIM; ++a)
            {
              trE += gu[a][a];
              gg += gpf[a] * gpf[a];
              ug += u[a] * gpf[a];
              for (int b = 0; b < DIM; ++b)
                {
                  const double E = 0.5 * (gu[a][b] + gu[b][a]);
                  trE2 += E * E;
                }
            }
          const double psi = 0.5 * pc.lambda * trE * trE + pc.mu * trE2;
          eb += ((1 + p->kappa) * pf * pf + p->kappa) * psi * t.JxW[q];
          ec += p->G_c / 2.0 * ((pf - 1) * (pf - 1) / p->eps + p->eps * gg) * t.JxW[q];
          tcv += ug * t.JxW[q];
        }
    }
  out[0] = eb;
  out[1] = ec;
  out[2] = tcv;
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
