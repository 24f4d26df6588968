// CPLEX-free loader for moip_aira's inputs.  Rebuilds the state the reference reaches through
// CPXreadcopyprob + Problem::read_lp_problem / read_mop_problem (reference src/problem.cpp:29-154,
// :157-340): the last k rows of an .lp are the objectives, k is the RHS of the very last row,
// the dummy objective line only carries the sense; every `N` row of a .mop is an objective.
#include "model.h"

#include <algorithm>
#include <cctype>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <unordered_map>

namespace moip {
namespace {

struct Tok {
  enum Kind { Cmp, Sign, Num, Id, Colon } kind;
  std::string s;
  double v = 0;
};

std::string lower(std::string s) {
  for (auto& c : s) c = (char)std::tolower((unsigned char)c);
  return s;
}

bool id_start(char c) {
  return std::isalpha((unsigned char)c) || std::strchr("_!\"#$%&()/,;?@'`{}|~", c) != nullptr;
}
bool id_char(char c) { return id_start(c) || std::isdigit((unsigned char)c) || c == '.' || c == '[' || c == ']'; }

bool tokenize(const std::string& t, std::vector<Tok>& out, std::string& err) {
  size_t i = 0;
  while (i < t.size()) {
    char c = t[i];
    if (std::isspace((unsigned char)c)) { ++i; continue; }
    Tok tk;
    if (c == '<' || c == '>' || c == '=') {
      tk.kind = Tok::Cmp;
      tk.s = std::string(1, c);
      ++i;
      if (i < t.size() && (t[i] == '=' || t[i] == '<' || t[i] == '>') && t[i] != c) { tk.s += t[i]; ++i; }
      else if (i < t.size() && t[i] == '=' && c != '=') { tk.s += '='; ++i; }
    } else if (c == '+' || c == '-') {
      tk.kind = Tok::Sign; tk.s = std::string(1, c); ++i;
    } else if (std::isdigit((unsigned char)c) || c == '.') {
      char* end = nullptr;
      tk.kind = Tok::Num;
      tk.v = std::strtod(t.c_str() + i, &end);
      if (end == t.c_str() + i) { err = "bad number near '" + t.substr(i, 20) + "'"; return false; }
      i = end - t.c_str();
    } else if (c == ':') {
      tk.kind = Tok::Colon; ++i;
    } else if (id_start(c)) {
      size_t j = i + 1;
      while (j < t.size() && id_char(t[j])) ++j;
      tk.kind = Tok::Id; tk.s = t.substr(i, j - i); i = j;
    } else {
      err = "unexpected character '" + std::string(1, c) + "'";
      return false;
    }
    out.push_back(tk);
  }
  return true;
}

// returns section keyword (normalised) if the line starts with one; rest = remainder of the line
std::string section_of(const std::string& line, std::string& rest) {
  static const char* kws[] = {"minimize", "minimise", "maximize", "maximise", "minimum", "maximum",
                              "subject to", "such that", "s.t.", "st.", "binaries", "binary", "bounds",
                              "bound", "generals", "general", "integers", "integer", "min", "max", "st",
                              "bin", "gen", "int", "end"};
  std::string l = lower(line);
  // collapse runs of blanks so that "subject   to" matches
  std::string c;
  for (size_t i = 0; i < l.size(); ++i) {
    if (std::isspace((unsigned char)l[i])) { if (!c.empty() && c.back() != ' ') c += ' '; }
    else c += l[i];
  }
  for (const char* kw : kws) {
    size_t L = std::strlen(kw);
    if (c.compare(0, L, kw) == 0 && (c.size() == L || c[L] == ' ')) {
      // find where the keyword ends in the original line
      size_t pos = 0, matched = 0;
      while (pos < line.size() && matched < L) {
        if (std::isspace((unsigned char)line[pos])) { if (kw[matched] == ' ') { ++matched; while (pos < line.size() && std::isspace((unsigned char)line[pos])) ++pos; continue; } ++pos; continue; }
        ++matched; ++pos;
      }
      rest = pos < line.size() ? line.substr(pos) : "";
      return kw;
    }
  }
  return "";
}

struct Row { std::map<int, double> coef; char sense; double rhs; };

int read_lp(const std::string& path, Model& M, std::string& err) {
  std::ifstream in(path);
  if (!in) { err = "cannot open " + path; return 2; }
  std::vector<std::pair<std::string, std::string>> sections;
  std::string line, kind, text;
  bool have = false;
  while (std::getline(in, line)) {
    size_t bs = line.find('\\');
    if (bs != std::string::npos) line.erase(bs);
    size_t a = line.find_first_not_of(" \t\r\n");
    if (a == std::string::npos) continue;
    line = line.substr(a);
    std::string rest;
    std::string kw = section_of(line, rest);
    if (!kw.empty()) {
      if (have) sections.emplace_back(kind, text);
      kind = kw; text = rest; have = true;
    } else {
      text += " " + line;
    }
  }
  if (have) sections.emplace_back(kind, text);

  std::unordered_map<std::string, int> index;
  auto col = [&](const std::string& nm) {
    auto it = index.find(nm);
    if (it != index.end()) return it->second;
    int j = (int)M.names.size();
    index[nm] = j; M.names.push_back(nm);
    return j;
  };
  std::vector<Row> rows;
  std::vector<std::string> binaries, generals;
  std::vector<std::vector<Tok>> bound_toks;
  int sense = -1;
  for (auto& sec : sections) {
    const std::string& kd = sec.first;
    std::vector<Tok> tk;
    if (kd.compare(0, 3, "min") == 0) { sense = 0; continue; }
    if (kd.compare(0, 3, "max") == 0) { sense = 1; continue; }
    if (kd == "end") break;
    if (!tokenize(sec.second, tk, err)) return 2;
    if (kd == "subject to" || kd == "such that" || kd == "s.t." || kd == "st" || kd == "st.") {
      Row cur; double sign = 1; bool have_num = false; double num = 0;
      for (size_t i = 0; i < tk.size(); ++i) {
        const Tok& t = tk[i];
        if (t.kind == Tok::Id && i + 1 < tk.size() && tk[i + 1].kind == Tok::Colon) { ++i; continue; }
        if (t.kind == Tok::Sign) { if (t.s == "-") sign = -sign; }
        else if (t.kind == Tok::Num) { num = t.v; have_num = true; }
        else if (t.kind == Tok::Id) {
          cur.coef[col(t.s)] += sign * (have_num ? num : 1.0);
          sign = 1; have_num = false;
        } else if (t.kind == Tok::Cmp) {
          double rs = 1; ++i;
          while (i < tk.size() && tk[i].kind == Tok::Sign) { if (tk[i].s == "-") rs = -rs; ++i; }
          if (i >= tk.size() || tk[i].kind != Tok::Num) { err = "constraint right-hand side must be a number"; return 2; }
          cur.rhs = rs * tk[i].v;
          cur.sense = (t.s[0] == '<' || t.s == "=<") ? 'L' : (t.s[0] == '>' || t.s == "=>") ? 'G' : 'E';
          rows.push_back(cur);
          cur = Row(); sign = 1; have_num = false;
        }
      }
    } else if (kd == "binary" || kd == "binaries" || kd == "bin" ||
               kd == "integers" || kd == "integer" || kd == "int") {
      // `integers` means binary in the shipped KP examples (pinned by 3KP10.out / 4KP10.out)
      for (auto& t : tk) if (t.kind == Tok::Id) binaries.push_back(t.s);
    } else if (kd == "general" || kd == "generals" || kd == "gen") {
      for (auto& t : tk) if (t.kind == Tok::Id) generals.push_back(t.s);
    } else if (kd == "bound" || kd == "bounds") {
      bound_toks.push_back(tk);
    }
  }
  if (sense < 0) { err = "no objective sense line (Minimize/Maximize)"; return 2; }
  if (rows.empty()) { err = "no constraint rows"; return 2; }
  int k = (int)rows.back().rhs;                    // reference src/problem.cpp:54-61
  if (k < 1 || k > (int)rows.size() || (double)k != rows.back().rhs) {
    err = "RHS of the last row is not a valid objective count"; return 2;
  }
  M.sense = sense;
  M.k = k;
  M.n = (int)M.names.size();
  M.ms = (int)rows.size() - k;
  M.lb.assign(M.n, 0.0);
  M.ub.assign(M.n, kInf);
  M.is_int.assign(M.n, 0);
  for (auto& nm : binaries) {
    auto it = index.find(nm);
    if (it == index.end()) { err = "unknown column in binary section: " + nm; return 2; }
    M.ub[it->second] = 1.0; M.is_int[it->second] = 1;
  }
  for (auto& nm : generals) {
    auto it = index.find(nm);
    if (it == index.end()) { err = "unknown column in general section: " + nm; return 2; }
    M.is_int[it->second] = 1;
  }
  for (auto& tk : bound_toks) {
    size_t i = 0;
    auto number = [&](size_t& i, double& v) -> bool {
      double s = 1;
      while (i < tk.size() && tk[i].kind == Tok::Sign) { if (tk[i].s == "-") s = -s; ++i; }
      if (i >= tk.size()) return false;
      if (tk[i].kind == Tok::Id && (lower(tk[i].s) == "inf" || lower(tk[i].s) == "infinity")) { v = s * kInf; ++i; return true; }
      if (tk[i].kind != Tok::Num) return false;
      v = s * tk[i].v; ++i; return true;
    };
    while (i < tk.size()) {
      if (tk[i].kind == Tok::Id) {
        auto it = index.find(tk[i].s);
        if (it == index.end()) { err = "unknown column in bounds: " + tk[i].s; return 2; }
        int j = it->second; ++i;
        if (i < tk.size() && tk[i].kind == Tok::Id && lower(tk[i].s) == "free") { M.lb[j] = -kInf; M.ub[j] = kInf; ++i; continue; }
        if (i >= tk.size() || tk[i].kind != Tok::Cmp) { err = "bad bounds line"; return 2; }
        std::string op = tk[i].s; ++i; double v;
        if (!number(i, v)) { err = "bad bounds value"; return 2; }
        if (op[0] == '<' || op == "=<") M.ub[j] = v; else if (op[0] == '>' || op == "=>") M.lb[j] = v; else M.lb[j] = M.ub[j] = v;
      } else {
        double v;
        if (!number(i, v) || i >= tk.size() || tk[i].kind != Tok::Cmp) { err = "bad bounds line"; return 2; }
        ++i;
        if (i >= tk.size() || tk[i].kind != Tok::Id || !index.count(tk[i].s)) { err = "bad bounds line"; return 2; }
        int j = index[tk[i].s]; ++i;
        M.lb[j] = v;
        if (i < tk.size() && tk[i].kind == Tok::Cmp) { ++i; double v2; if (!number(i, v2)) { err = "bad bounds value"; return 2; } M.ub[j] = v2; }
      }
    }
  }
  M.a_ptr.assign(1, 0);
  for (int i = 0; i < M.ms; ++i) {
    for (auto& kv : rows[i].coef) if (kv.second != 0.0) { M.a_col.push_back(kv.first); M.a_val.push_back(kv.second); }
    M.a_ptr.push_back((int)M.a_col.size());
    M.row_sense.push_back(rows[i].sense);
    M.rhs.push_back(rows[i].rhs);
  }
  M.objcoef.assign((size_t)k * M.n, 0.0);
  for (int o = 0; o < k; ++o)
    for (auto& kv : rows[M.ms + o].coef) M.objcoef[(size_t)o * M.n + kv.first] = kv.second;
  return 0;
}

int read_mop(const std::string& path, Model& M, std::string& err) {
  std::ifstream in(path);
  if (!in) { err = "cannot open " + path; return 2; }
  std::string line, sect;
  std::vector<std::string> obj_rows, con_rows;
  std::unordered_map<std::string, int> oi, ci, index;
  std::vector<char> con_sense;
  struct Ent { int col; std::string row; double v; };
  std::vector<Ent> ents;
  std::unordered_map<std::string, double> rhs;
  std::vector<uint8_t> isint, has_bound;
  bool in_int = false;
  M.sense = 0;
  while (std::getline(in, line)) {
    if (line.empty() || line[0] == '*') continue;
    std::istringstream ss(line);
    std::vector<std::string> f;
    for (std::string w; ss >> w;) f.push_back(w);
    if (f.empty()) continue;
    if (!std::isspace((unsigned char)line[0])) {
      sect = f[0];
      for (auto& c : sect) c = (char)std::toupper((unsigned char)c);
      if (sect == "OBJSENSE" && f.size() > 1) M.sense = (std::toupper((unsigned char)f[1][0]) == 'M' && std::toupper((unsigned char)f[1][1]) == 'A') ? 1 : 0;
      if (sect == "ENDATA") break;
      continue;
    }
    if (sect == "OBJSENSE") {
      M.sense = (f[0].size() > 1 && std::toupper((unsigned char)f[0][1]) == 'A') ? 1 : 0;
    } else if (sect == "ROWS") {
      if (f.size() < 2) continue;
      if (f[0] == "N") { oi[f[1]] = (int)obj_rows.size(); obj_rows.push_back(f[1]); }
      else { ci[f[1]] = (int)con_rows.size(); con_rows.push_back(f[1]); con_sense.push_back(f[0][0]); }
    } else if (sect == "COLUMNS") {
      if (f.size() >= 3 && f[1] == "'MARKER'") { in_int = (f[2] == "'INTORG'"); continue; }
      auto it = index.find(f[0]);
      int j;
      if (it == index.end()) {
        j = (int)M.names.size(); index[f[0]] = j; M.names.push_back(f[0]);
        isint.push_back(in_int ? 1 : 0); has_bound.push_back(0);
        M.lb.push_back(0.0); M.ub.push_back(kInf);
      } else j = it->second;
      for (size_t q = 1; q + 1 < f.size(); q += 2) ents.push_back({j, f[q], std::atof(f[q + 1].c_str())});
    } else if (sect == "RHS") {
      for (size_t q = 1; q + 1 < f.size(); q += 2) rhs[f[q]] = std::atof(f[q + 1].c_str());
    } else if (sect == "BOUNDS") {
      if (f.size() < 3) continue;
      auto it = index.find(f[2]);
      if (it == index.end()) { err = "unknown column in BOUNDS: " + f[2]; return 2; }
      int j = it->second; has_bound[j] = 1;
      double v = f.size() > 3 ? std::atof(f[3].c_str()) : 0.0;
      const std::string& t = f[0];
      if (t == "LO") M.lb[j] = v; else if (t == "UP") M.ub[j] = v;
      else if (t == "FX") M.lb[j] = M.ub[j] = v; else if (t == "PL") M.ub[j] = kInf;
      else if (t == "MI") M.lb[j] = -kInf; else if (t == "FR") { M.lb[j] = -kInf; M.ub[j] = kInf; }
      else if (t == "BV") { M.lb[j] = 0; M.ub[j] = 1; isint[j] = 1; }
      else if (t == "LI") { M.lb[j] = v; isint[j] = 1; } else if (t == "UI") { M.ub[j] = v; isint[j] = 1; }
    } else if (sect == "RANGES") {
      err = "RANGES section is not supported"; return 4;
    }
  }
  M.n = (int)M.names.size();
  M.k = (int)obj_rows.size();
  M.ms = (int)con_rows.size();
  if (M.k < 1) { err = "no objective (N) rows"; return 2; }
  for (int j = 0; j < M.n; ++j) if (isint[j] && !has_bound[j]) M.ub[j] = 1.0;  // MPS marker default
  M.is_int = isint;
  std::vector<std::map<int, double>> rows(M.ms);
  M.objcoef.assign((size_t)M.k * M.n, 0.0);
  for (auto& e : ents) {
    auto io = oi.find(e.row);
    if (io != oi.end()) { M.objcoef[(size_t)io->second * M.n + e.col] = (double)(int)e.v; continue; }  // `signed int val` (:261-264)
    auto ic = ci.find(e.row);
    if (ic != ci.end()) rows[ic->second][e.col] += e.v;
  }
  M.a_ptr.assign(1, 0);
  for (int i = 0; i < M.ms; ++i) {
    for (auto& kv : rows[i]) if (kv.second != 0.0) { M.a_col.push_back(kv.first); M.a_val.push_back(kv.second); }
    M.a_ptr.push_back((int)M.a_col.size());
    M.row_sense.push_back(con_sense[i]);
    auto it = rhs.find(con_rows[i]);
    M.rhs.push_back(it == rhs.end() ? 0.0 : it->second);
  }
  return 0;
}

bool to_int_scaled(double v, double scale, int64_t& out) {
  double s = v * scale;
  double r = std::nearbyint(s);
  if (std::fabs(s - r) > 1e-7 * std::max(1.0, std::fabs(s))) return false;
  if (std::fabs(r) > 4.0e15) return false;
  out = (int64_t)r;
  return true;
}

}  // namespace

int finalize_model(Model& M, std::string& err) {
  const int n = M.n, ms = M.ms, k = M.k;
  if (k > 4) { err = "at most 4 objectives are supported (reference src/aira.cpp:230-233)"; return 4; }
  if (n < 1) { err = "model has no columns"; return 2; }
  for (int j = 0; j < n; ++j)
    if (!M.is_int[j]) { err = "continuous column '" + M.names[j] + "': only pure integer programs are supported"; return 4; }
  // ---- exact integer image ------------------------------------------------------------------
  M.ci.assign((size_t)k * n, 0);
  for (size_t q = 0; q < M.objcoef.size(); ++q)
    if (!to_int_scaled(M.objcoef[q], 1.0, M.ci[q])) { err = "objective coefficients must be integers"; return 4; }
  M.ai_val.assign(M.a_val.size(), 0);
  M.ri_lo.assign(ms, INT64_MIN);
  M.ri_hi.assign(ms, INT64_MAX);
  for (int i = 0; i < ms; ++i) {
    double scale = 1.0; bool ok = false;
    for (int d = 0; d <= 6 && !ok; ++d, scale *= 10.0) {
      ok = true;
      for (int q = M.a_ptr[i]; q < M.a_ptr[i + 1]; ++q) { int64_t t; if (!to_int_scaled(M.a_val[q], scale, t)) { ok = false; break; } }
      if (ok) break;
    }
    if (!ok) { err = "row " + std::to_string(i) + ": coefficients are not decimal fractions with <= 6 digits"; return 4; }
    for (int q = M.a_ptr[i]; q < M.a_ptr[i + 1]; ++q) to_int_scaled(M.a_val[q], scale, M.ai_val[q]);
    double b = M.rhs[i] * scale;
    char s = M.row_sense[i];
    if (s == 'L' || s == 'E') M.ri_hi[i] = (int64_t)std::floor(b + 1e-9 * std::max(1.0, std::fabs(b)));
    if (s == 'G' || s == 'E') M.ri_lo[i] = (int64_t)std::ceil(b - 1e-9 * std::max(1.0, std::fabs(b)));
    if (s == 'E' && M.ri_lo[i] > M.ri_hi[i]) M.int_infeasible = true;   // non-integral equality
  }
  M.lbI.assign(n, 0); M.ubI.assign(n, 0);
  std::vector<int64_t> lo(n), hi(n);
  const int64_t BIG = (int64_t)1 << 40;
  for (int j = 0; j < n; ++j) {
    lo[j] = M.lb[j] <= -kInf ? -BIG : (int64_t)std::ceil(M.lb[j] - 1e-9);
    hi[j] = M.ub[j] >= kInf ? BIG : (int64_t)std::floor(M.ub[j] + 1e-9);
  }
  // implied bounds by activity propagation (needed for e.g. the unbounded knapsack .mop)
  for (int pass = 0; pass < 20; ++pass) {
    bool changed = false;
    for (int i = 0; i < ms; ++i) {
      // min / max activity with infinite contributions counted
      int64_t mn = 0, mx = 0; int mn_inf = 0, mx_inf = 0;
      for (int q = M.a_ptr[i]; q < M.a_ptr[i + 1]; ++q) {
        int64_t a = M.ai_val[q]; int j = M.a_col[q];
        int64_t l = lo[j], h = hi[j];
        if (a > 0) { if (l <= -BIG) ++mn_inf; else mn += a * l; if (h >= BIG) ++mx_inf; else mx += a * h; }
        else if (a < 0) { if (h >= BIG) ++mn_inf; else mn += a * h; if (l <= -BIG) ++mx_inf; else mx += a * l; }
      }
      for (int q = M.a_ptr[i]; q < M.a_ptr[i + 1]; ++q) {
        int64_t a = M.ai_val[q]; int j = M.a_col[q];
        if (a == 0) continue;
        if (M.ri_hi[i] != INT64_MAX) {   // sum <= hi : bound using the others' min activity
          int64_t mine; bool mine_inf;
          if (a > 0) { mine_inf = lo[j] <= -BIG; mine = mine_inf ? 0 : a * lo[j]; }
          else { mine_inf = hi[j] >= BIG; mine = mine_inf ? 0 : a * hi[j]; }
          if (mn_inf - (mine_inf ? 1 : 0) == 0) {
            int64_t rest = mn - mine, room = M.ri_hi[i] - rest;
            if (a > 0) { int64_t nb = (int64_t)std::floor((double)room / (double)a + 1e-12); if (nb < hi[j]) { hi[j] = nb; changed = true; } }
            else { int64_t nb = (int64_t)std::ceil((double)room / (double)a - 1e-12); if (nb > lo[j]) { lo[j] = nb; changed = true; } }
          }
        }
        if (M.ri_lo[i] != INT64_MIN) {   // sum >= lo : bound using the others' max activity
          int64_t mine; bool mine_inf;
          if (a > 0) { mine_inf = hi[j] >= BIG; mine = mine_inf ? 0 : a * hi[j]; }
          else { mine_inf = lo[j] <= -BIG; mine = mine_inf ? 0 : a * lo[j]; }
          if (mx_inf - (mine_inf ? 1 : 0) == 0) {
            int64_t rest = mx - mine, need = M.ri_lo[i] - rest;
            if (a > 0) { int64_t nb = (int64_t)std::ceil((double)need / (double)a - 1e-12); if (nb > lo[j]) { lo[j] = nb; changed = true; } }
            else { int64_t nb = (int64_t)std::floor((double)need / (double)a + 1e-12); if (nb < hi[j]) { hi[j] = nb; changed = true; } }
          }
        }
      }
    }
    if (!changed) break;
  }
  M.all_binary = true;
  for (int j = 0; j < n; ++j) {
    if (lo[j] <= -BIG || hi[j] >= BIG || hi[j] > INT32_MAX / 4 || lo[j] < INT32_MIN / 4) {
      err = "integer column '" + M.names[j] + "' has no finite implied bounds"; return 4;
    }
    if (lo[j] > hi[j]) M.int_infeasible = true;
    M.lbI[j] = (int32_t)lo[j]; M.ubI[j] = (int32_t)hi[j];
    if (!(lo[j] >= 0 && hi[j] <= 1)) M.all_binary = false;
  }
  // ---- scaled LP image: K = [A ; sgn*C] (m x n), Ruiz (10 passes) + Pock-Chambolle ----------
  const int m = ms + k;
  M.m = m;
  if ((double)m * n > 6.0e7) { err = "model too large for the dense scaling pass"; return 5; }
  const double sgn = M.sense == 0 ? 1.0 : -1.0;
  std::vector<double> K((size_t)m * n, 0.0);
  for (int i = 0; i < ms; ++i)
    for (int q = M.a_ptr[i]; q < M.a_ptr[i + 1]; ++q) K[(size_t)i * n + M.a_col[q]] += M.a_val[q];
  for (int o = 0; o < k; ++o)
    for (int j = 0; j < n; ++j) K[(size_t)(ms + o) * n + j] = sgn * M.objcoef[(size_t)o * n + j];
  M.dr.assign(m, 1.0); M.dc.assign(n, 1.0);
  std::vector<double> r(m), c(n);
  auto apply = [&]() {
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) K[(size_t)i * n + j] /= (r[i] * c[j]);
    for (int i = 0; i < m; ++i) M.dr[i] /= r[i];
    for (int j = 0; j < n; ++j) M.dc[j] /= c[j];
  };
  for (int it = 0; it < 10; ++it) {
    std::fill(r.begin(), r.end(), 0.0); std::fill(c.begin(), c.end(), 0.0);
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) { double a = std::fabs(K[(size_t)i * n + j]); r[i] = std::max(r[i], a); c[j] = std::max(c[j], a); }
    for (auto& v : r) v = v > 0 ? std::sqrt(v) : 1.0;
    for (auto& v : c) v = v > 0 ? std::sqrt(v) : 1.0;
    apply();
  }
  std::fill(r.begin(), r.end(), 0.0); std::fill(c.begin(), c.end(), 0.0);
  for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) { double a = std::fabs(K[(size_t)i * n + j]); r[i] += a; c[j] += a; }
  for (auto& v : r) v = v > 0 ? std::sqrt(v) : 1.0;
  for (auto& v : c) v = v > 0 ? std::sqrt(v) : 1.0;
  apply();
  // spectral norm by power iteration on K^T K
  {
    std::vector<double> v(n, 1.0 / std::sqrt((double)n)), u(m), w(n);
    double sig = 0;
    for (int it = 0; it < 400; ++it) {
      for (int i = 0; i < m; ++i) { double s = 0; for (int j = 0; j < n; ++j) s += K[(size_t)i * n + j] * v[j]; u[i] = s; }
      for (int j = 0; j < n; ++j) w[j] = 0;
      for (int i = 0; i < m; ++i) { double ui = u[i]; if (ui != 0) for (int j = 0; j < n; ++j) w[j] += K[(size_t)i * n + j] * ui; }
      double nw = 0; for (double x : w) nw += x * x; nw = std::sqrt(nw);
      if (nw == 0) break;
      double ns = std::sqrt(nw);
      for (int j = 0; j < n; ++j) v[j] = w[j] / nw;
      if (std::fabs(ns - sig) < 1e-10 * ns && it > 20) { sig = ns; break; }
      sig = ns;
    }
    if (sig <= 0) sig = 1.0;
    M.eta = 0.98 / sig;
  }
  // structural part: scaled CSR values + column-ELL of the transpose
  M.s_val.resize(M.a_val.size());
  std::vector<int> cnt(n, 0);
  for (int i = 0; i < ms; ++i)
    for (int q = M.a_ptr[i]; q < M.a_ptr[i + 1]; ++q) { M.s_val[q] = M.a_val[q] * M.dr[i] * M.dc[M.a_col[q]]; ++cnt[M.a_col[q]]; }
  M.ell_w = 0;
  for (int j = 0; j < n; ++j) M.ell_w = std::max(M.ell_w, cnt[j]);
  M.ellT_val.assign((size_t)M.ell_w * n, 0.0);
  M.ellT_row.assign((size_t)M.ell_w * n, 0);
  std::fill(cnt.begin(), cnt.end(), 0);
  for (int i = 0; i < ms; ++i)
    for (int q = M.a_ptr[i]; q < M.a_ptr[i + 1]; ++q) {
      int j = M.a_col[q], e = cnt[j]++;
      M.ellT_val[(size_t)e * n + j] = M.s_val[q];
      M.ellT_row[(size_t)e * n + j] = i;
    }
  M.D.assign((size_t)k * n, 0.0);
  for (int o = 0; o < k; ++o) for (int j = 0; j < n; ++j) M.D[(size_t)o * n + j] = K[(size_t)(ms + o) * n + j];
  M.s_lo.assign(ms, -HUGE_VAL); M.s_hi.assign(ms, HUGE_VAL);
  M.norm_row_bounds2 = 0;
  for (int i = 0; i < ms; ++i) {
    char s = M.row_sense[i];
    if (s == 'L' || s == 'E') { M.s_hi[i] = M.rhs[i] * M.dr[i]; }
    if (s == 'G' || s == 'E') { M.s_lo[i] = M.rhs[i] * M.dr[i]; }
    double b = M.rhs[i] * M.dr[i];
    M.norm_row_bounds2 += b * b;
  }
  // ---- fast-kernel image ----------------------------------------------------------------------
  {
    // small models with few rows (knapsacks): every row rides in the column pass, which is what k1_small.cu needs
    const bool all_dense = n <= 64 && k + ms <= 5;
    const int long_thr = all_dense ? 0 : std::max(32, n / 8);
    std::vector<int> shortr, longr;
    for (int i = 0; i < ms; ++i) ((M.a_ptr[i + 1] - M.a_ptr[i]) > long_thr ? longr : shortr).push_back(i);
    M.msS = (int)shortr.size(); M.nL = (int)longr.size(); M.KD = k + M.nL;
    M.krow.clear();
    for (int i : shortr) M.krow.push_back(i);
    for (int o = 0; o < k; ++o) M.krow.push_back(ms + o);
    for (int i : longr) M.krow.push_back(i);
    M.dr_k.assign(m, 1.0); M.lo_k.assign(m, -HUGE_VAL); M.hi_k.assign(m, HUGE_VAL);
    for (int r2 = 0; r2 < m; ++r2) {
      const int i = M.krow[r2];
      M.dr_k[r2] = M.dr[i];
      if (i < ms) { M.lo_k[r2] = M.s_lo[i]; M.hi_k[r2] = M.s_hi[i]; }
    }
    M.RW = 0;
    for (int i : shortr) M.RW = std::max(M.RW, M.a_ptr[i + 1] - M.a_ptr[i]);
    M.rowell_val.assign((size_t)std::max(1, M.RW) * std::max(1, M.msS), 0.0);
    M.rowell_col.assign((size_t)std::max(1, M.RW) * std::max(1, M.msS), 0);
    std::vector<int> ccnt(n, 0);
    for (int r2 = 0; r2 < M.msS; ++r2) {
      const int i = shortr[r2];
      int e = 0;
      for (int q = M.a_ptr[i]; q < M.a_ptr[i + 1]; ++q, ++e) {
        M.rowell_val[(size_t)e * M.msS + r2] = M.s_val[q];
        M.rowell_col[(size_t)e * M.msS + r2] = M.a_col[q];
        ++ccnt[M.a_col[q]];
      }
    }
    M.ell2_w = 0;
    for (int j = 0; j < n; ++j) M.ell2_w = std::max(M.ell2_w, ccnt[j]);
    M.ellT2_val.assign((size_t)std::max(1, M.ell2_w) * n, 0.0);
    M.ellT2_row.assign((size_t)std::max(1, M.ell2_w) * n, 0);
    std::fill(ccnt.begin(), ccnt.end(), 0);
    for (int r2 = 0; r2 < M.msS; ++r2) {
      const int i = shortr[r2];
      for (int q = M.a_ptr[i]; q < M.a_ptr[i + 1]; ++q) {
        const int j = M.a_col[q], e = ccnt[j]++;
        M.ellT2_val[(size_t)e * n + j] = M.s_val[q];
        M.ellT2_row[(size_t)e * n + j] = r2;
      }
    }
    M.D2.assign((size_t)M.KD * n, 0.0);
    for (int o = 0; o < k; ++o) for (int j = 0; j < n; ++j) M.D2[(size_t)o * n + j] = M.D[(size_t)o * n + j];
    for (int t = 0; t < M.nL; ++t) {
      const int i = longr[t];
      for (int q = M.a_ptr[i]; q < M.a_ptr[i + 1]; ++q) M.D2[(size_t)(k + t) * n + M.a_col[q]] += M.s_val[q];
    }
    M.fast_ok = M.KD <= 6 && M.ell2_w <= 2 && m <= 128;
    const int E = M.ell2_w;
    const int uc = E + M.KD + (E + 1) / 2;
    M.col_units = (uc + 1) & ~1;
    M.colrec.assign((size_t)n * M.col_units, 0.0);
    for (int j = 0; j < n; ++j) {
      double* rec = M.colrec.data() + (size_t)j * M.col_units;
      for (int e = 0; e < E; ++e) rec[e] = M.ellT2_val[(size_t)e * n + j];
      for (int d = 0; d < M.KD; ++d) rec[E + d] = M.D2[(size_t)d * n + j];
      int32_t* ids = reinterpret_cast<int32_t*>(rec + E + M.KD);
      for (int e = 0; e < E; ++e) ids[e] = M.ellT2_row[(size_t)e * n + j];
    }
    {
      // register-resident kernel (k1_reg.cuh): LPR (1 or 2) lanes own a short row and add up the products the
      // column pass scattered into prod[row*RWP + pos]; each lane reads 64 contiguous bytes per step (4 x 16 B)
      // and steps by LPR*64 bytes.  The top 16*(KD+2) threads of the CTA add up the dense rows and the two
      // restart norms instead.  RWP/2 odd puts the 8 lanes of a quarter-warp into 8 distinct 16-byte banks;
      // positions inside a row are free, so they are dealt greedily such that the 16 columns a half-warp
      // stores together hit distinct 8-byte banks too.  Offsets are bytes from the start of dynamic shared
      // memory (ysh at 0, prod at 2048: k1reg::kProdBase).
      const int nt_min = n <= 256 ? 128 : 256;
      const int budget = nt_min - 16 * (M.KD + 2);            // threads left for the short rows
      const int lg = (2 * M.msS <= budget) ? 1 : 0;
      const int LPR = 1 << lg;
      const int trips = std::max(1, (M.RW + 8 * LPR - 1) / (8 * LPR));
      const int span = 8 * LPR * trips;                        // positions the row owners read
      int rwp = span;
      while (rwp % 4 != 2) ++rwp;
      const unsigned prod_base = 2048;
      M.RWP = M.RW > 0 ? rwp : 0;
      M.reg_lpr_log2 = lg;
      M.reg_trips = M.RW > 0 ? trips : 0;
      M.reg_ok = M.fast_ok && M.KD >= 2 && M.KD <= 5 && (E == 0 || E == 2) && n > 64 && n <= 1024 && M.msS <= budget &&
                 m <= 256 && prod_base + ((long long)M.msS * M.RWP + 2) * 8 < 60000;
      M.colrec2 = M.colrec;
      if (M.reg_ok && E > 0) {
        std::vector<std::vector<char>> used(M.msS, std::vector<char>(std::max(1, span), 0));
        const unsigned dummy = prod_base + (unsigned)(M.msS * M.RWP) * 8u;
        for (int e = 0; e < E; ++e)
          for (int g0 = 0; g0 < n; g0 += 16) {
            unsigned banks = 0;
            for (int j = g0; j < std::min(n, g0 + 16); ++j) {
              double* rec = M.colrec2.data() + (size_t)j * M.col_units;
              int32_t* ids = reinterpret_cast<int32_t*>(rec + E + M.KD);
              const int r2 = M.ellT2_row[(size_t)e * n + j];
              if (M.ellT2_val[(size_t)e * n + j] == 0.0) { ids[e] = (int32_t)(dummy << 16); continue; }   // padding entry
              int pick = -1, fallback = -1;
              for (int q = 0; q < span; ++q) {
                if (used[r2][q]) continue;
                if (fallback < 0) fallback = q;
                if (!((banks >> ((r2 * M.RWP + q) & 15)) & 1u)) { pick = q; break; }
              }
              if (pick < 0) pick = fallback;
              used[r2][pick] = 1;
              banks |= 1u << ((r2 * M.RWP + pick) & 15);
              const unsigned off = prod_base + (unsigned)(r2 * M.RWP + pick) * 8u;
              ids[e] = (int32_t)((off << 16) | ((unsigned)r2 * 8u));
            }
          }
      }
    }
    M.rowrec.assign((size_t)std::max(1, M.RW) * std::max(1, M.msS) * 2, 0.0);
    for (int e = 0; e < M.RW; ++e)
      for (int r2 = 0; r2 < M.msS; ++r2) {
        double* rec = M.rowrec.data() + ((size_t)e * M.msS + r2) * 2;
        rec[0] = M.rowell_val[(size_t)e * M.msS + r2];
        int32_t* id = reinterpret_cast<int32_t*>(rec + 1);
        id[0] = M.rowell_col[(size_t)e * M.msS + r2]; id[1] = 0;
      }
  }
  return 0;
}

int load_model(const std::string& path, Model& out, std::string& err) {
  out = Model();
  out.path = path;
  int rc;
  auto ends = [&](const char* suf) { size_t L = std::strlen(suf); return path.size() > L && path.compare(path.size() - L, L, suf) == 0; };
  if (ends(".lp")) rc = read_lp(path, out, err);
  else if (ends(".mop")) rc = read_mop(path, out, err);
  else { err = "unknown file type (need .lp or .mop)"; return 2; }
  if (rc) return rc;
  return finalize_model(out, err);
}

}  // namespace moip
