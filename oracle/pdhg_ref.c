/* TEST INFRASTRUCTURE ONLY -- plain-C CPU restatement of kernel K1's algorithm (reflected,
 * restarted Halpern PDHG in fp64 on the Ruiz/Pock-Chambolle scaled LP).  The reference has no
 * counterpart: its node LPs live inside CPXmipopt (reference src/aira.cpp:480), so LP parity is
 * UNPINNED by the reference; objective values are cross-checked against
 * HiGHS in tests/.  This file is the checker for the CUDA kernel's arithmetic and, run over all
 * host cores with OpenMP, the "port" CPU baseline of bench.py.  Never linked into the product.
 *
 * LP (min-form):  min c.x  s.t.  lo <= K x <= hi (dense K, m x n),  l <= x <= u.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define ST_CONVERGED 0
#define ST_CUTOFF 1
#define ST_ITERLIMIT 2
#define ST_INFEASIBLE 3

typedef struct {
  int n, m;
  double *S;      /* scaled matrix m x n */
  double *dr, *dc;
  double eta;
  int *ptr, *col;   /* CSR of the nonzeros of S (the CPU port exploits sparsity like the kernel) */
  double *val;
} scaled_t;

static double clampd(double v, double a, double b) { return fmin(fmax(v, a), b); }

/* Ruiz (10 passes, sqrt of max-abs) then Pock-Chambolle (alpha = 1), spectral norm by power iteration */
static void scale_model(int m, int n, const double* K, scaled_t* s) {
  s->n = n; s->m = m;
  s->S = (double*)malloc(sizeof(double) * m * n);
  s->dr = (double*)malloc(sizeof(double) * m);
  s->dc = (double*)malloc(sizeof(double) * n);
  memcpy(s->S, K, sizeof(double) * m * n);
  for (int i = 0; i < m; ++i) s->dr[i] = 1.0;
  for (int j = 0; j < n; ++j) s->dc[j] = 1.0;
  double* r = (double*)malloc(sizeof(double) * m);
  double* c = (double*)malloc(sizeof(double) * n);
  for (int pass = 0; pass < 11; ++pass) {
    for (int i = 0; i < m; ++i) r[i] = 0;
    for (int j = 0; j < n; ++j) c[j] = 0;
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < n; ++j) {
        double a = fabs(s->S[(size_t)i * n + j]);
        if (pass < 10) { if (a > r[i]) r[i] = a; if (a > c[j]) c[j] = a; }
        else { r[i] += a; c[j] += a; }
      }
    for (int i = 0; i < m; ++i) r[i] = r[i] > 0 ? sqrt(r[i]) : 1.0;
    for (int j = 0; j < n; ++j) c[j] = c[j] > 0 ? sqrt(c[j]) : 1.0;
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < n; ++j) s->S[(size_t)i * n + j] /= (r[i] * c[j]);
    for (int i = 0; i < m; ++i) s->dr[i] /= r[i];
    for (int j = 0; j < n; ++j) s->dc[j] /= c[j];
  }
  double* v = (double*)malloc(sizeof(double) * n);
  double* u = (double*)malloc(sizeof(double) * m);
  double* w = (double*)malloc(sizeof(double) * n);
  for (int j = 0; j < n; ++j) v[j] = 1.0 / sqrt((double)n);
  double sig = 0;
  for (int it = 0; it < 400; ++it) {
    for (int i = 0; i < m; ++i) { double t = 0; for (int j = 0; j < n; ++j) t += s->S[(size_t)i * n + j] * v[j]; u[i] = t; }
    for (int j = 0; j < n; ++j) w[j] = 0;
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) w[j] += s->S[(size_t)i * n + j] * u[i];
    double nw = 0; for (int j = 0; j < n; ++j) nw += w[j] * w[j]; nw = sqrt(nw);
    if (nw == 0) break;
    double ns = sqrt(nw);
    for (int j = 0; j < n; ++j) v[j] = w[j] / nw;
    if (fabs(ns - sig) < 1e-10 * ns && it > 20) { sig = ns; break; }
    sig = ns;
  }
  if (sig <= 0) sig = 1.0;
  s->eta = 0.98 / sig;
  free(r); free(c); free(v); free(u); free(w);
  int nnz = 0;
  for (size_t q = 0; q < (size_t)m * n; ++q) nnz += s->S[q] != 0.0;
  s->ptr = (int*)malloc(sizeof(int) * (m + 1));
  s->col = (int*)malloc(sizeof(int) * (nnz > 0 ? nnz : 1));
  s->val = (double*)malloc(sizeof(double) * (nnz > 0 ? nnz : 1));
  nnz = 0;
  for (int i = 0; i < m; ++i) {
    s->ptr[i] = nnz;
    for (int j = 0; j < n; ++j)
      if (s->S[(size_t)i * n + j] != 0.0) { s->col[nnz] = j; s->val[nnz] = s->S[(size_t)i * n + j]; ++nnz; }
  }
  s->ptr[m] = nnz;
}

/* one node LP on a pre-scaled model; lo/hi/l/u/c are UNSCALED, per node */
static void solve_node(const scaled_t* s, const double* c_un, const double* lo_un, const double* hi_un,
                       const double* l_un, const double* u_un, double eps, int max_iter, int check_every,
                       int fixed_iters, int norm_every, double cutoff, double cutoff_slack,
                       double* x_out, double* y_out, double* pobj_out, double* lb_out, int* iters_out, int* status_out) {
  const int n = s->n, m = s->m;
  const double eta = s->eta;
  double* buf = (double*)calloc((size_t)(7 * n + 8 * m), sizeof(double));
  double *x = buf, *xa = x + n, *xbar = xa + n, *l = xbar + n, *u = l + n, *c = u + n;
  double *y = c + n, *ya = y + m, *yt = ya + m, *sx = yt + m, *sxa = sx + m, *sxt = sxa + m, *lo = sxt + m, *hi = lo + m, *g = hi + m;
  double c2 = 0, obj_upper = 0, bn2 = 0, bn2_un = 0;
  for (int j = 0; j < n; ++j) {
    l[j] = l_un[j] / s->dc[j]; u[j] = u_un[j] / s->dc[j];
    c[j] = c_un[j] * s->dc[j];
    x[j] = clampd(0.0, l[j], u[j]); xa[j] = x[j];
    c2 += c[j] * c[j];
    obj_upper += fmax(c[j] * l[j], c[j] * u[j]);
  }
  for (int i = 0; i < m; ++i) {
    lo[i] = isinf(lo_un[i]) ? lo_un[i] : lo_un[i] * s->dr[i];
    hi[i] = isinf(hi_un[i]) ? hi_un[i] : hi_un[i] * s->dr[i];
    double t = !isinf(hi[i]) ? hi[i] : (!isinf(lo[i]) ? lo[i] : 0.0);
    bn2 += t * t;
    double tu = !isinf(hi_un[i]) ? hi_un[i] : (!isinf(lo_un[i]) ? lo_un[i] : 0.0);
    bn2_un += tu * tu;
    double q = 0;
    for (int e = s->ptr[i]; e < s->ptr[i + 1]; ++e) q += s->val[e] * x[s->col[e]];
    sx[i] = q; sxa[i] = q;
  }
  double w = (c2 > 0 && bn2 > 0) ? sqrt(c2 / bn2) : 1.0;
  double tau = eta / w, sigma = eta * w;
  const double kkt_bden = 1.0 + sqrt(bn2_un);
  int kk = 0, it = 0, status = ST_ITERLIMIT;
  double r0 = 0, rprev = -1.0, best_lb = -HUGE_VAL, pobj = 0;
  const int iter_cap = fixed_iters > 0 ? fixed_iters : max_iter;
  for (;;) {
    ++it;
    const int norm_it = (kk == 0) || (kk % norm_every == 0);
    double dx2 = 0;
    for (int j = 0; j < n; ++j) g[j] = 0;
    for (int i = 0; i < m; ++i) {
      const double yi = y[i];
      if (yi != 0.0) for (int e = s->ptr[i]; e < s->ptr[i + 1]; ++e) g[s->col[e]] += s->val[e] * yi;
    }
    for (int j = 0; j < n; ++j) {
      double xt = clampd(x[j] - tau * (c[j] - g[j]), l[j], u[j]);
      xbar[j] = 2.0 * xt - x[j];
      dx2 += (xt - x[j]) * (xt - x[j]);
    }
    double dy2 = 0, cross = 0;
    for (int i = 0; i < m; ++i) {
      double q = 0;
      if (!(isinf(lo[i]) && isinf(hi[i])))
        for (int e = s->ptr[i]; e < s->ptr[i + 1]; ++e) q += s->val[e] * xbar[s->col[e]];
      double sxti = 0.5 * (q + sx[i]);
      double v = y[i] / sigma - q;
      double yti = sigma * (v - clampd(v, -hi[i], -lo[i]));
      yt[i] = yti; sxt[i] = sxti;
      double dy = yti - y[i];
      dy2 += dy * dy; cross += dy * (sxti - sx[i]);
    }
    int restart = 0;
    if (norm_it) {
      double fp = fmax(0.0, (w / eta) * dx2 - 2.0 * cross + dy2 / (eta * w));   /* squared fixed-point error */
      if (kk == 0) r0 = fp;
      else if (fp <= 0.04 * r0 || (fp <= 0.64 * r0 && rprev >= 0.0 && fp > rprev) || 25 * kk >= 9 * it) restart = 1;
      rprev = fp;
    }
    int stop = 0;
    if ((fixed_iters <= 0 && (it % check_every) == 0) || (fixed_iters > 0 && it >= iter_cap)) {
      double po = 0, dcol = 0, drow = 0, pres2 = 0, fcol = 0, fabs_sum = 0;
      for (int j = 0; j < n; ++j) g[j] = 0;
      for (int i = 0; i < m; ++i) {
        const double yi = yt[i];
        if (yi != 0.0) for (int e = s->ptr[i]; e < s->ptr[i + 1]; ++e) g[s->col[e]] += s->val[e] * yi;
      }
      for (int j = 0; j < n; ++j) {
        double r = c[j] - g[j];
        po += c[j] * 0.5 * (xbar[j] + x[j]);
        dcol += (r > 0) ? r * l[j] : r * u[j];
        double fk = (g[j] < 0) ? -g[j] * l[j] : -g[j] * u[j];       /* the same bound with the objective dropped (Farkas) */
        fcol += fk; fabs_sum += fabs(fk);
      }
      for (int i = 0; i < m; ++i) {
        double rt = 0;
        if (yt[i] > 0) rt = yt[i] * lo[i]; else if (yt[i] < 0) rt = yt[i] * hi[i];
        drow += rt; fabs_sum += fabs(rt);
        double viol = fmax(0.0, fmax(sxt[i] - hi[i], lo[i] - sxt[i])) / s->dr[i];
        if (!(isinf(lo[i]) && isinf(hi[i]))) pres2 += viol * viol;
      }
      pobj = po;
      double dobj = dcol + drow;
      if (fixed_iters > 0) best_lb = dobj;
      else {
        if (dobj > best_lb) best_lb = dobj;
        double gap = fabs(pobj - dobj);
        double rel = fmax(sqrt(pres2) / kkt_bden, gap / (1.0 + fabs(pobj) + fabs(dobj)));
        if (best_lb >= cutoff - cutoff_slack) { status = ST_CUTOFF; stop = 1; }
        /* Farkas certificate: fcol + drow > 0 proves infeasibility (the K1 kernels test the same quantity) */
        else if (fcol + drow > 1e-9 * fabs_sum + 1e-9 || best_lb > obj_upper + 1e-6 * (1.0 + fabs(obj_upper))) { status = ST_INFEASIBLE; stop = 1; }
        else if (rel <= eps) { status = ST_CONVERGED; stop = 1; }
      }
    }
    if (it >= iter_cap) stop = 1;
    if (stop) break;
    if (restart) {
      double dxn = 0, dyn = 0;
      for (int j = 0; j < n; ++j) { double xn = 0.5 * (xbar[j] + x[j]); dxn += (xn - xa[j]) * (xn - xa[j]); x[j] = xn; xa[j] = xn; }
      for (int i = 0; i < m; ++i) { dyn += (yt[i] - ya[i]) * (yt[i] - ya[i]); y[i] = yt[i]; ya[i] = yt[i]; sx[i] = sxt[i]; sxa[i] = sxt[i]; }
      dxn = sqrt(dxn); dyn = sqrt(dyn);
      if (dxn > 1e-10 && dyn > 1e-10) w = exp(0.5 * log(dyn / dxn) + 0.5 * log(w));
      tau = eta / w; sigma = eta * w;
      kk = 0; rprev = -1.0;
    } else {
      double a = (double)(kk + 1) / (double)(kk + 2), c1 = 1.0 - a;
      for (int j = 0; j < n; ++j) x[j] = a * xbar[j] + c1 * xa[j];
      for (int i = 0; i < m; ++i) { y[i] = a * (2.0 * yt[i] - y[i]) + c1 * ya[i]; sx[i] = a * (2.0 * sxt[i] - sx[i]) + c1 * sxa[i]; }
      ++kk;
    }
  }
  for (int j = 0; j < n; ++j) if (x_out) x_out[j] = 0.5 * (xbar[j] + x[j]) * s->dc[j];
  for (int i = 0; i < m; ++i) if (y_out) y_out[i] = yt[i] * s->dr[i];
  *pobj_out = pobj; *lb_out = best_lb; *iters_out = it; *status_out = status;
  free(buf);
}

typedef struct {
  const scaled_t* s;
  int B, m, n;
  const double *c, *lo, *hi, *l, *u;
  double eps; int max_iter, check_every, fixed_iters, norm_every; double cutoff, cutoff_slack;
  double *x_out, *pobj, *lb; int *iters, *status;
  atomic_int next;
} job_t;

static void* worker_main(void* arg) {
  job_t* j = (job_t*)arg;
  for (;;) {
    int b = atomic_fetch_add(&j->next, 1);
    if (b >= j->B) break;
    const int n = j->n, m = j->m;
    solve_node(j->s, j->c + (size_t)b * n, j->lo + (size_t)b * m, j->hi + (size_t)b * m, j->l + (size_t)b * n,
               j->u + (size_t)b * n, j->eps, j->max_iter, j->check_every, j->fixed_iters, j->norm_every, j->cutoff,
               j->cutoff_slack, j->x_out ? j->x_out + (size_t)b * n : NULL, NULL, j->pobj + b, j->lb + b,
               j->iters + b, j->status + b);
  }
  return NULL;
}

/* Batch entry point.  K: dense m x n (shared); per node b: c[b][n], lo[b][m], hi[b][m], l[b][n], u[b][n]
 * (use +-INFINITY for free sides).  Nodes are solved by `threads` pthreads (<=0: all online cores).
 * Returns the number of threads used. */
int pdhg_ref_batch(int m, int n, const double* K, int B, const double* c, const double* lo, const double* hi,
                   const double* l, const double* u, double eps, int max_iter, int check_every, int fixed_iters,
                   int norm_every, double cutoff, double cutoff_slack, int threads,
                   double* x_out, double* pobj, double* lb, int* iters, int* status) {
  scaled_t s;
  scale_model(m, n, K, &s);
  if (norm_every < 1) norm_every = 1;
  int nt = threads > 0 ? threads : (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (nt < 1) nt = 1;
  if (nt > B) nt = B > 0 ? B : 1;
  job_t j = {&s, B, m, n, c, lo, hi, l, u, eps, max_iter, check_every, fixed_iters, norm_every, cutoff, cutoff_slack,
             x_out, pobj, lb, iters, status, 0};
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nt);
  for (int t = 1; t < nt; ++t) pthread_create(&th[t], NULL, worker_main, &j);
  worker_main(&j);
  for (int t = 1; t < nt; ++t) pthread_join(th[t], NULL);
  free(th);
  free(s.S); free(s.dr); free(s.dc); free(s.ptr); free(s.col); free(s.val);
  return nt;
}
