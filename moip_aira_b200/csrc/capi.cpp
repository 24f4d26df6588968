// Model-side entry points of the C ABI (no device code): replaces Problem::Problem and the
// CPXreadcopyprob / CPXget* calls of reference src/problem.cpp:12-154, :157-340.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "solver.h"

extern "C" int moip_model_load(const char* path, moip_model** out) {
  if (!path || !out) return MOIP_ERR_ARG;
  moip_model* m = new moip_model();
  std::string err;
  int rc = moip::load_model(path, m->M, err);
  if (rc) {
    std::fprintf(stderr, "moip_b200: %s: %s\n", path, err.c_str());
    delete m;
    return rc;
  }
  *out = m;
  return MOIP_OK;
}

extern "C" void moip_model_free(moip_model* m) { delete m; }

extern "C" int moip_model_get_info(const moip_model* m, moip_model_info* info) {
  if (!m || !info) return MOIP_ERR_ARG;
  const moip::Model& M = m->M;
  info->n = M.n; info->ms = M.ms; info->k = M.k; info->nnz = (int)M.a_val.size();
  info->sense = M.sense; info->all_binary = M.all_binary ? 1 : 0; info->m = M.m;
  info->mask_words = (M.n + 15) / 16;
  return MOIP_OK;
}

extern "C" int moip_model_objcoef(const moip_model* m, int obj, double* out_n) {
  if (!m || !out_n || obj < 0 || obj >= m->M.k) return MOIP_ERR_ARG;
  std::memcpy(out_n, m->M.objcoef.data() + (size_t)obj * m->M.n, sizeof(double) * m->M.n);
  return MOIP_OK;
}

extern "C" int moip_model_dense(const moip_model* m, double* a_ms_n, char* row_sense_ms, double* rhs_ms,
                                double* lb_n, double* ub_n, uint8_t* is_int_n) {
  if (!m) return MOIP_ERR_ARG;
  const moip::Model& M = m->M;
  if (a_ms_n) {
    std::memset(a_ms_n, 0, sizeof(double) * (size_t)M.ms * M.n);
    for (int i = 0; i < M.ms; ++i)
      for (int q = M.a_ptr[i]; q < M.a_ptr[i + 1]; ++q) a_ms_n[(size_t)i * M.n + M.a_col[q]] += M.a_val[q];
  }
  for (int i = 0; i < M.ms; ++i) {
    if (row_sense_ms) row_sense_ms[i] = M.row_sense[i];
    if (rhs_ms) rhs_ms[i] = M.rhs[i];
  }
  for (int j = 0; j < M.n; ++j) {
    if (lb_n) lb_n[j] = M.lb[j];
    if (ub_n) ub_n[j] = M.ub[j];
    if (is_int_n) is_int_n[j] = M.is_int[j];
  }
  return MOIP_OK;
}

extern "C" int moip_model_colname(const moip_model* m, int j, char* buf, int buflen) {
  if (!m || !buf || buflen < 1 || j < 0 || j >= m->M.n) return MOIP_ERR_ARG;
  std::snprintf(buf, (size_t)buflen, "%s", m->M.names[j].c_str());
  return MOIP_OK;
}

extern "C" int moip_model_selfcheck(const moip_model* mm, int* kernel_path, char* msg, int msglen) {
  if (!mm) return MOIP_ERR_ARG;
  const moip::Model& M = mm->M;
  const int n = M.n, ms = M.ms, k = M.k, m = M.m;
  auto fail = [&](const std::string& why) {
    if (msg && msglen > 0) std::snprintf(msg, (size_t)msglen, "%s", why.c_str());
    return MOIP_ERR_LIMIT;
  };
  if (msg && msglen > 0) msg[0] = 0;
  // ---- kernel the dispatcher (csrc/k1_reg.cu) will pick
  const bool small = M.fast_ok && M.msS == 0 && M.ell2_w == 0 && M.KD >= 3 && M.KD <= 5 && n <= 64 && m == M.KD;
  const size_t need = (size_t)5 * n + (size_t)8 * m;
  int path = M.reg_ok ? 2 : small ? 3 : M.fast_ok ? 1 : ((need + 24 * 8) * sizeof(double) > 200 * 1024 ? 4 : 0);
  if (kernel_path) *kernel_path = path;
  // ---- row order and scalings
  if ((int)M.krow.size() != m || (int)M.dr_k.size() != m) return fail("kernel row order has the wrong length");
  std::vector<char> seen(m, 0);
  for (int r2 = 0; r2 < m; ++r2) {
    const int i = M.krow[r2];
    if (i < 0 || i >= m || seen[i]) return fail("kernel row order is not a permutation");
    seen[i] = 1;
    if (M.dr_k[r2] != M.dr[i]) return fail("row scaling not carried into kernel order");
  }
  if (M.msS + M.KD != m || M.KD != k + M.nL) return fail("short / dense row split inconsistent");
  // ---- dense block: D2[d][j] = dr * K * dc for the k objectives and the long rows
  const double sgn = M.sense == 0 ? 1.0 : -1.0;
  for (int d = 0; d < M.KD; ++d) {
    const int i = M.krow[M.msS + d];
    std::vector<double> rowv(n, 0.0);
    if (i >= ms) for (int j = 0; j < n; ++j) rowv[j] = sgn * M.objcoef[(size_t)(i - ms) * n + j];
    else for (int q = M.a_ptr[i]; q < M.a_ptr[i + 1]; ++q) rowv[M.a_col[q]] += M.a_val[q];
    for (int j = 0; j < n; ++j) {
      const double want = rowv[j] * M.dr[i] * M.dc[j], got = M.D2[(size_t)d * n + j];
      if (std::fabs(want - got) > 1e-12 * (1.0 + std::fabs(want))) return fail("dense block does not reproduce the scaled matrix");
    }
  }
  if (!M.fast_ok) return MOIP_OK;
  // ---- packed column records (k1_fast / k1_reg): values and row ids of the short rows
  const int E = M.ell2_w, U = M.col_units;
  long long nnz_short = 0;
  for (int r2 = 0; r2 < M.msS; ++r2) nnz_short += M.a_ptr[M.krow[r2] + 1] - M.a_ptr[M.krow[r2]];
  long long found = 0;
  std::vector<char> slot_used(M.reg_ok ? (size_t)M.msS * M.RWP + 1 : 1, 0);
  const int LPR = 1 << M.reg_lpr_log2;
  if (M.reg_ok && M.msS > 0) {
    if ((M.RWP / 2) % 2 != 1 || M.RWP % 2) return fail("RWP/2 must be odd (bank-conflict-free 16-byte row reads)");
    if (8 * LPR * M.reg_trips > M.RWP || 8 * LPR * M.reg_trips < M.RW) return fail("row span does not cover the widest short row");
  }
  for (int j = 0; j < n; ++j) {
    const double* rec = M.colrec.data() + (size_t)j * U;
    const double* rec2 = M.colrec2.data() + (size_t)j * U;
    const int32_t* ids = reinterpret_cast<const int32_t*>(rec + E + M.KD);
    const uint32_t* ids2 = reinterpret_cast<const uint32_t*>(rec2 + E + M.KD);
    for (int d = 0; d < M.KD; ++d)
      if (rec[E + d] != M.D2[(size_t)d * n + j] || rec2[E + d] != rec[E + d]) return fail("dense values of a column record differ from D2");
    for (int e = 0; e < E; ++e) {
      const double v = rec[e];
      if (v != M.ellT2_val[(size_t)e * n + j] || rec2[e] != v) return fail("short-row value of a column record differs");
      if (v == 0.0) {
        if (M.reg_ok && (ids2[e] >> 16) != 2048u + (unsigned)(M.msS * M.RWP) * 8u) return fail("padding entry does not point at the dummy slot");
        continue;
      }
      ++found;
      const int r2 = ids[e];
      if (r2 < 0 || r2 >= M.msS) return fail("row id of a column record out of range");
      const int i = M.krow[r2];
      double a = 0; bool present = false;
      for (int q = M.a_ptr[i]; q < M.a_ptr[i + 1]; ++q) if (M.a_col[q] == j) { a += M.a_val[q]; present = true; }
      if (!present || std::fabs(a * M.dr[i] * M.dc[j] - v) > 1e-12 * (1.0 + std::fabs(v))) return fail("column record entry is not the scaled matrix entry");
      if (M.reg_ok) {
        const unsigned yoff = ids2[e] & 0xffffu, poff = ids2[e] >> 16;
        if (yoff != (unsigned)r2 * 8u) return fail("dual offset of a column record is wrong");
        if (poff < 2048u || (poff - 2048u) % 8u) return fail("product offset misaligned");
        const unsigned slot = (poff - 2048u) / 8u;
        if (slot / (unsigned)M.RWP != (unsigned)r2 || slot % (unsigned)M.RWP >= (unsigned)(8 * LPR * M.reg_trips)) return fail("product slot outside the span its row owners read");
        if (slot_used[slot]) return fail("two entries share a product slot");
        slot_used[slot] = 1;
      }
    }
  }
  if (found != nnz_short) return fail("column records do not hold every short-row nonzero exactly once");
  return MOIP_OK;
}

extern "C" const char* moip_version(void) { return "moip_b200 0.1 (sm_100a)"; }
