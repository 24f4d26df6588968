// Model-side entry points of the C ABI (no device code): replaces Problem::Problem and the
// CPXreadcopyprob / CPXget* calls of reference src/problem.cpp:12-154, :157-340.
#include <cstdio>
#include <cstring>

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

extern "C" const char* moip_version(void) { return "moip_b200 0.1 (sm_100a)"; }
