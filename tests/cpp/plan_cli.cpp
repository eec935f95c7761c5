// CLI around the pure partition-planning functions of finitevolume.jl_b200/csrc/host_util.h, so that the CPU test suite can
// compare them with their Python mirrors (finitevolume.jl_b200/distributed.py) on random cases.
//   plan_cli planes <n1> <P> <dirichlet_ends>         -> "lo hi" per rank
//   plan_cli faces_before <n1> <n2> <n3> <i1> <i2> <i3>
//   plan_cli halo   (stdin: P, then per rank: start nf nhalo col...)  -> per rank: peers / send_counts / recv_counts / send_dst / send_rows
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../finitevolume.jl_b200/csrc/host_util.h"

int main(int argc, char **argv) {
  if (argc >= 5 && !strcmp(argv[1], "planes")) {
    std::vector<int64_t> lo, hi;
    fvb::slab_planes(atoll(argv[2]), atoi(argv[3]), atoi(argv[4]) != 0, lo, hi);
    for (size_t r = 0; r < lo.size(); ++r) std::printf("%lld %lld\n", (long long)lo[r], (long long)hi[r]);
    return 0;
  }
  if (argc >= 8 && !strcmp(argv[1], "faces_before")) {
    std::printf("%lld\n", (long long)fvb::regulargrid_faces_before(atoll(argv[2]), atoll(argv[3]), atoll(argv[4]), atoll(argv[5]),
                                                                    atoll(argv[6]), atoll(argv[7])));
    return 0;
  }
  if (argc >= 2 && !strcmp(argv[1], "halo")) {
    int P = 0;
    if (std::scanf("%d", &P) != 1) return 2;
    std::vector<int64_t> start((size_t)P), nf((size_t)P);
    std::vector<std::vector<int64_t>> halo((size_t)P);
    for (int r = 0; r < P; ++r) {
      long long s, n, nh;
      if (std::scanf("%lld %lld %lld", &s, &n, &nh) != 3) return 2;
      start[(size_t)r] = s; nf[(size_t)r] = n;
      for (long long k = 0; k < nh; ++k) { long long c; if (std::scanf("%lld", &c) != 1) return 2; halo[(size_t)r].push_back(c); }
    }
    std::vector<fvb::RankHaloPlan> plans;
    const char *why = fvb::plan_halo_exchange(start, nf, halo, plans);
    if (why[0]) { std::printf("error %s\n", why); return 0; }
    for (auto &pl : plans) {
      auto line = [](const char *tag, auto &v) { std::printf("%s", tag); for (auto x : v) std::printf(" %lld", (long long)x); std::printf("\n"); };
      line("peers", pl.peers); line("send_counts", pl.send_counts); line("recv_counts", pl.recv_counts);
      line("send_dst", pl.send_dst); line("send_rows", pl.send_rows);
    }
    return 0;
  }
  return 2;
}
