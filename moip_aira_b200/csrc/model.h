// Host-side model: what the reference keeps in Problem + the opaque CPXLPptr
// (reference src/problem.h:11-21, src/env.h:6-10), rebuilt without CPLEX, plus the
// derived images the kernels need (exact integer image, scaled LP image).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace moip {

constexpr double kInf = 1.0e20;  // CPX_INFBOUND (reference src/problem.cpp:126)

struct Model {
  // ---- as read from the file ------------------------------------------------------------
  std::string path;
  int n = 0, ms = 0, k = 0;
  int sense = 0;                       // 0 MIN, 1 MAX (reference src/sense.h)
  std::vector<std::string> names;      // column names, order of first appearance
  // structural rows in CSR (file order)
  std::vector<int> a_ptr, a_col;
  std::vector<double> a_val;
  std::vector<char> row_sense;         // 'L','G','E'
  std::vector<double> rhs;             // ms
  std::vector<double> objcoef;         // k*n dense, Problem::objcoef
  std::vector<double> lb, ub;          // column bounds (kInf = none)
  std::vector<uint8_t> is_int;

  // ---- exact integer image (K2/K4) -------------------------------------------------------
  // row i:  lo_i <= sum a_ij x_j <= hi_i  with int64 data; columns lbI..ubI (int32)
  std::vector<int64_t> ai_val;         // same pattern as a_ptr/a_col
  std::vector<int64_t> ri_lo, ri_hi;   // ms (INT64_MIN/MAX = free side)
  std::vector<int64_t> ci;             // k*n objective coefficients
  std::vector<int32_t> lbI, ubI;       // implied-finite integer bounds
  bool all_binary = false;
  bool int_infeasible = false;         // an equality row with non-integral rhs etc.

  // ---- scaled LP image (K1): K = [A ; sgn*C], S = Dr K Dc ---------------------------------
  int m = 0;                           // ms + k
  std::vector<double> dr, dc;          // row / column scalings (x = dc*xs, y = dr*ys)
  // structural part of S^T in column-ELL: entry e of column j at [e*n + j]
  int ell_w = 0;
  std::vector<double> ellT_val;
  std::vector<int> ellT_row;
  // structural part of S in CSR
  std::vector<double> s_val;           // pattern = a_ptr/a_col
  // dense objective block D = Dr[ms..] * sgn*C * Dc  (k*n)
  std::vector<double> D;
  std::vector<double> s_lo, s_hi;      // scaled structural row bounds (+-inf as +-HUGE_VAL)
  // ---- fast-kernel image: rows reordered as [short structural | k objective | long structural]
  bool fast_ok = false;                // false -> generic kernel (too many long rows / wide rows)
  int msS = 0, nL = 0, KD = 0, RW = 0, ell2_w = 0;
  std::vector<int> krow;               // kernel row -> original row (ms + o for objective o)
  std::vector<double> dr_k;            // [m] row scaling in kernel order
  std::vector<double> lo_k, hi_k;      // [m] scaled row bounds in kernel order (objective rows: per node)
  std::vector<double> rowell_val;      // [RW][msS] short rows of S, row-ELL
  std::vector<int> rowell_col;
  std::vector<double> ellT2_val;       // [ell2_w][n] short rows of S^T, column-ELL (row ids in kernel order)
  std::vector<int> ellT2_row;
  std::vector<double> D2;              // [KD][n] dense rows: k objectives then the long structural rows
  // packed images the fast kernel actually reads (16-byte vector loads, one address per column/entry)
  int col_units = 0;                   // 8-byte units per column record (even)
  std::vector<double> colrec;          // [n][col_units]: ell2_w values | KD dense values | ell2_w row ids (int32 pairs)
  std::vector<double> rowrec;          // [RW][msS][2]: {value, column id (int32 in the low half)}
  // register-resident kernel: products S_ij*xbar_j are scattered by the column pass into prod[row*RWP+pos]
  bool reg_ok = false;
  int RWP = 0;                         // padded short-row width (RWP/2 odd: bank-conflict-free 16-byte row reads)
  int reg_lpr_log2 = 0;                // lanes per row owner group (log2)
  int reg_trips = 0;                   // 64-byte steps per lane over its share of a short row
  std::vector<double> colrec2;         // like colrec, ids packed as (product byte offset << 16 | dual byte offset)
  double eta = 0;                      // 0.99 / ||S||_2
  double norm_row_bounds2 = 0;         // sum of squares of finite scaled structural bounds
};

// Loads .lp / .mop (dispatch on extension like reference src/problem.cpp:16-26).
// Returns 0 on success; on failure fills err.
int load_model(const std::string& path, Model& out, std::string& err);
// Builds integer + scaled images; called by load_model.
int finalize_model(Model& m, std::string& err);

}  // namespace moip
