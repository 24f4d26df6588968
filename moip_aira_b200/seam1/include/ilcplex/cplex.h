/*
 * ilcplex/cplex.h -- link-level stand-in for the part of the CPLEX Callable Library that moip_aira binds
 * ("Seam 1").  NOT IBM's header: it declares only the 23 entry points and 10 macros the
 * reference's src/aira.cpp and src/problem.cpp use, with the argument lists of their call sites, so that the
 * UNMODIFIED reference sources compile (-I moip_aira_b200/seam1/include) and link against
 * libcplex_moip_b200.so (cpx_shim.cpp), which forwards every call to the B200 solver core behind
 * include/moip_b200.h.  Each declaration cites the reference call site it serves.
 *
 * Conventions kept from the callable library: every call returns an int status, 0 = OK; CPXopenCPLEX /
 * CPXcreateprob return handles and write a status through a pointer; handles are owned by the caller; output
 * arrays are caller-allocated; one environment + one problem per worker thread (src/aira.cpp:541-585), each used
 * by its owner only.
 */
#ifndef MOIP_B200_SEAM1_CPLEX_H
#define MOIP_B200_SEAM1_CPLEX_H

/* the reference relies on these being dragged in by cplex.h (memcpy, abs/round, exit; SURVEY hazard 7.3-5) */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

struct moip_cpxenv;
struct moip_cpxlp;
typedef struct moip_cpxenv* CPXENVptr;
typedef const struct moip_cpxenv* CPXCENVptr;
typedef struct moip_cpxlp* CPXLPptr;
typedef const struct moip_cpxlp* CPXCLPptr;

/* must stay the literal: (int)CPX_INFBOUND is constant-folded by the reference (src/aira.cpp:802-803, :1961-1980) */
#define CPX_INFBOUND 1.0E+20
#define CPX_MIN 1
#define CPX_MAX (-1)
#define CPX_ON 1
#define CPX_OFF 0
/* MIP solution statuses compared after CPXgetstat (src/aira.cpp:410, :490, :632, :840) */
#define CPXMIP_OPTIMAL 101
#define CPXMIP_INFEASIBLE 103
#define CPXMIP_INForUNBD 119
/* parameters the reference sets (src/aira.cpp:236-246, :422, :570-579); accepted and recorded, they do not
 * change results: the solve behind CPXmipopt is exact and deterministic */
#define CPX_PARALLEL_DETERMINISTIC 1
#define CPXPARAM_Parallel 1109
#define CPXPARAM_Threads 1067
#define CPX_PARAM_SCRIND 1035
#define CPXPARAM_MIP_Tolerances_MIPGap 2009

/* src/aira.cpp:226, :564 */
CPXENVptr CPXopenCPLEX(int* status_p);
/* src/aira.cpp:317 */
int CPXcloseCPLEX(CPXENVptr* env_p);
/* src/problem.cpp:32, :160 */
CPXLPptr CPXcreateprob(CPXCENVptr env, int* status_p, const char* probname);
/* src/aira.cpp:311 */
int CPXfreeprob(CPXCENVptr env, CPXLPptr* lp_p);
/* src/problem.cpp:40, :168 -- extended .lp (objectives = last k rows, count = last rhs) or multi-N-row .mop */
int CPXreadcopyprob(CPXCENVptr env, CPXLPptr lp, const char* filename, const char* filetype);
/* src/problem.cpp:47-49, :175-176; src/aira.cpp:376, :464 */
int CPXgetnumcols(CPXCENVptr env, CPXCLPptr lp);
int CPXgetnumrows(CPXCENVptr env, CPXCLPptr lp);
int CPXgetnumnz(CPXCENVptr env, CPXCLPptr lp);
/* src/problem.cpp:54 */
int CPXgetrhs(CPXCENVptr env, CPXCLPptr lp, double* rhs, int begin, int end);
/* src/problem.cpp:87 */
int CPXgetrows(CPXCENVptr env, CPXCLPptr lp, int* nzcnt_p, int* rmatbeg, int* rmatind, double* rmatval,
               int rmatspace, int* surplus_p, int begin, int end);
/* src/problem.cpp:119, :297 */
int CPXgetobjsen(CPXCENVptr env, CPXCLPptr lp);
/* src/problem.cpp:141, :330 */
int CPXchgsense(CPXCENVptr env, CPXLPptr lp, int cnt, const int* indices, const char* sense);
/* src/problem.cpp:148; src/aira.cpp:383, :474 */
int CPXchgrhs(CPXCENVptr env, CPXLPptr lp, int cnt, const int* indices, const double* values);
/* src/problem.cpp:183 */
int CPXgetcolname(CPXCENVptr env, CPXCLPptr lp, char** name, char* namestore, int storespace, int* surplus_p,
                  int begin, int end);
/* src/problem.cpp:321 */
int CPXaddrows(CPXCENVptr env, CPXLPptr lp, int ccnt, int rcnt, int nzcnt, const double* rhs, const char* sense,
               const int* rmatbeg, const int* rmatind, const double* rmatval, char** colname, char** rowname);
/* src/aira.cpp:378, :469 */
int CPXchgobj(CPXCENVptr env, CPXLPptr lp, int cnt, const int* indices, const double* values);
/* src/aira.cpp:394 */
int CPXchgobjsen(CPXCENVptr env, CPXLPptr lp, int maxormin);
/* src/aira.cpp:236-246, :570-579 */
int CPXsetintparam(CPXENVptr env, int whichparam, int newvalue);
/* src/aira.cpp:422, :502 */
int CPXsetdblparam(CPXENVptr env, int whichparam, double newvalue);
/* src/aira.cpp:400, :423, :480, :503 -- the hot call: one exact single-objective IP on the GPU */
int CPXmipopt(CPXCENVptr env, CPXLPptr lp);
/* src/aira.cpp:409, :425, :489, :505 */
int CPXgetstat(CPXCENVptr env, CPXCLPptr lp);
/* src/aira.cpp:413, :429, :493, :509 */
int CPXgetobjval(CPXCENVptr env, CPXCLPptr lp, double* objval_p);
/* src/aira.cpp:439, :521 */
int CPXgetx(CPXCENVptr env, CPXCLPptr lp, double* x, int begin, int end);

#ifdef __cplusplus
}
#endif
#endif /* MOIP_B200_SEAM1_CPLEX_H */
