// Host build of the kernel bodies with PK_COUNT_ITERS: the solvers call pk_count_nfree(nfree, tag)
// at their decision points; this harness histograms the calls.  Source of the statistics quoted in
// DESIGN.md (share of instances per QP path, factorisations per instance, changed columns between
// consecutive factorisations).  Test / analysis tool only.
#include <cstdio>
#include <map>
static std::map<std::pair<int, int>, long> g_hist;
extern "C" void pk_count_nfree(int nfree, int tag) { g_hist[{tag, nfree}]++; }
extern "C" void pk_count_dump() {
  for (auto& kv : g_hist) printf("tag %5d value %3d : %ld\n", kv.first.first, kv.first.second, kv.second);
  g_hist.clear();
}
#define PK_COUNT_ITERS 1
#include "../../tests/hostsim/hostsim.cpp"
