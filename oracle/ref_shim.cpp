// TEST INFRASTRUCTURE ONLY -- thin extern "C" window onto the reference's own
// Solutions/Result classes (compiled from /root/reference/src where they lie, see
// oracle/Makefile).  Nothing here re-implements reference logic; it only forwards.
#include <vector>
#include "solutions.h"   // reference src/solutions.h (found via -I at build time)

extern "C" {

void* refsol_create(int k) { return new Solutions(k); }
void refsol_destroy(void* h) { delete static_cast<Solutions*>(h); }

void refsol_insert(void* h, const double* ip, const int* result, int infeasible) {
  static_cast<Solutions*>(h)->insert(ip, result, infeasible != 0);
}

// index (insertion order) of the record Solutions::find returns, -1 = nullptr
int refsol_find(void* h, const double* ip, int sense_is_max) {
  Solutions* s = static_cast<Solutions*>(h);
  const Result* r = s->find(ip, sense_is_max ? MAX : MIN);
  if (!r) return -1;
  int i = 0;
  for (auto it = s->begin(); it != s->end(); ++it, ++i)
    if (*it == r) return i;
  return -2;
}

// Q queries, returns first-match index per query (used for timing the reference scan)
void refsol_find_batch(void* h, int q, int k, const double* ips, int sense_is_max, int* out) {
  for (int i = 0; i < q; ++i) out[i] = refsol_find(h, ips + (size_t)i * k, sense_is_max);
}

// timing entry: the reference's own Solutions::find on q queries, nothing else; returns the number of hits
int refsol_find_count(void* h, int q, int k, const double* ips, int sense_is_max) {
  Solutions* s = static_cast<Solutions*>(h);
  int hits = 0;
  for (int i = 0; i < q; ++i) hits += s->find(ips + (size_t)i * k, sense_is_max ? MAX : MIN) != nullptr;
  return hits;
}

int refsol_size(void* h) {
  Solutions* s = static_cast<Solutions*>(h);
  int i = 0;
  for (auto it = s->begin(); it != s->end(); ++it) ++i;
  return i;
}

// sort_unique, then copy the feasible rows out (k ints each); returns the row count
int refsol_sort_unique(void* h, int k, int* rows, int cap) {
  Solutions* s = static_cast<Solutions*>(h);
  s->sort_unique();
  int n = 0;
  for (auto it = s->begin(); it != s->end(); ++it) {
    if ((*it)->infeasible) continue;
    if (n < cap) for (int j = 0; j < k; ++j) rows[(size_t)n * k + j] = (*it)->result[j];
    ++n;
  }
  return n;
}

}  // extern "C"
