// Host-only harness (no kernel launch, runs without a GPU): prints the shared-memory layout the TMA
// diagonal SpMV would use for the offsets given on the command line, as one JSON object.
// Built with nvcc and checked against an addressing model by tests/test_host_logic.py.
//   usage: dump_dia_tma_layout <unit 0|1> <o_0> [<o_1> ...]
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "../../finitevolume.jl_b200/csrc/dia_tma.cuh"

namespace fvb {
thread_local std::string g_last_error;
int set_error(int code, const std::string &msg) { g_last_error = msg; return code; }
}  // namespace fvb

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  const bool unit = atoi(argv[1]) != 0;
  const int K = argc - 2;
  if (K > fvb::kDiaMaxOff) return 2;
  int64_t off[fvb::kDiaMaxOff] = {};
  for (int k = 0; k < K; ++k) off[k] = atoll(argv[2 + k]);
  const fvb::DiaTmaLayout L = fvb::dia_tma_layout(K, off, unit);
  auto arr = [&](const char *name, const int *v) {
    std::printf("\"%s\": [%d, %d, %d, %d], ", name, v[0], v[1], v[2], v[3]);
  };
  std::printf("{\"T\": %d, \"near_max\": %d, \"margin\": %d, \"xc\": %d, \"dg\": %d, ", fvb::kDiaTmaTile, fvb::kDiaNearMax,
              L.margin, L.xc, L.dg);
  arr("un", L.un); arr("xl", L.xl); arr("xu", L.xu); arr("ul", L.ul); arr("uu", L.uu);
  std::printf("\"stage_doubles\": %d, \"stages\": %d, \"bar_off\": %d, \"tx_bytes\": %u, \"reach\": %lld, \"smem_bytes\": %zu, "
              "\"smem_max\": %d}\n", L.stage_doubles, L.stages, L.bar_off, L.tx_bytes, (long long)L.reach,
              fvb::dia_tma_smem_bytes(L), fvb::kDiaTmaSmemMax);
  return 0;
}
