// compiled core of tools/smsp_sim.py (same model; see that file)
// fast port of tools/smsp_sim.py's core loop: reads "pipe occ stall" triples for the flattened dynamic sequence from stdin
#include <cstdio>
#include <vector>
#include <algorithm>
int main(int argc, char** argv) {
  int warps = atoi(argv[1]); int policy = atoi(argv[2]);  // 0 gto 1 lrr
  int n; if (scanf("%d", &n) != 1) return 1;
  std::vector<int> pipe(n), occ(n), stall(n);
  for (int i = 0; i < n; ++i) if (scanf("%d %d %d", &pipe[i], &occ[i], &stall[i]) != 3) return 1;
  std::vector<int> pc(warps, 0); std::vector<long> ready(warps, 0);
  long pf[4] = {0, 0, 0, 0}, busy[4] = {0, 0, 0, 0};
  long t = 0; int last = 0, done = 0; int rr = 0;
  while (done < warps) {
    bool issued = false;
    for (int k = 0; k < warps + 1 && !issued; ++k) {
      int w;
      if (policy == 0) { if (k == 0) w = last; else { w = k - 1; if (w == last) continue; } }
      else { if (k == warps) break; w = (rr + k) % warps; }
      if (pc[w] >= n || ready[w] > t) continue;
      int i = pc[w]; int p = pipe[i];
      if (pf[p] > t) continue;
      pf[p] = t + occ[i]; busy[p] += occ[i];
      ready[w] = t + stall[i]; pc[w]++;
      if (pc[w] >= n) done++;
      last = w; rr = (w + 1) % warps; issued = true;
    }
    t++;
  }
  printf("%ld %ld %ld\n", t, busy[0], busy[1]);
}
