// GPU smoke of moip_pool_synergistic_front (cooperative workers) without Python: prints the front of every .lp given for
// W = 1..k workers plus the pool's counters.  Build + run (profiles/r01_coop_smoke_gpu.log was made this way):
//   g++ -O2 -std=c++17 tools/coop_smoke.cpp -o /tmp/coop_smoke -Lmoip_aira_b200 -lmoip_b200 -Wl,-rpath,$PWD/moip_aira_b200
//   /tmp/coop_smoke Examples/3KP10.lp Examples/3AP05.lp ...
#include <cstdio>
#include <vector>
#include "../include/moip_b200.h"
int main(int argc, char** argv) {
  for (int a = 1; a < argc; ++a) {
    moip_model* m = nullptr;
    if (moip_model_load(argv[a], &m)) { std::printf("%s load failed\n", argv[a]); continue; }
    moip_model_info info;
    moip_model_get_info(m, &info);
    moip_pool* p = nullptr;
    int rc = moip_pool_create(m, 0, info.k, &p);
    if (rc) { std::printf("%s pool_create rc=%d\n", argv[a], rc); return 1; }
    for (int w = 1; w <= info.k; ++w) {
      std::vector<int> rows(4096 * info.k);
      int n = 0;
      rc = moip_pool_synergistic_front(p, w, rows.data(), 4096, &n);
      std::printf("%s W=%d rc=%d n=%d :", argv[a], w, rc, n);
      for (int i = 0; i < n && i < 4096; ++i) {
        std::printf(" (");
        for (int j = 0; j < info.k; ++j) std::printf("%d%s", rows[i * info.k + j], j + 1 < info.k ? "," : ")");
      }
      std::printf("\n");
      std::fflush(stdout);
    }
    moip_stats st;
    moip_pool_stats(p, &st);
    std::printf("%s stats: ips=%lld nodes=%lld lps=%lld launches=%lld\n", argv[a], (long long)st.ip_solved,
                (long long)st.bb_nodes, (long long)st.node_lps, (long long)st.kernel_launches);
    moip_pool_destroy(p);
    moip_model_free(m);
  }
  return 0;
}
