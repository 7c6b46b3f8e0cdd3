// TEST INFRASTRUCTURE ONLY: exercises pgdrive_b200/csrc/pgd_hostpool.h (the thread pool of pgd_step_host) under
// ThreadSanitizer / AddressSanitizer (tools/sanitize_host_builds.sh).  Every item of every job must run exactly once,
// across sleep / wake cycles, with more threads than cores.
#include <cstdio>
#include <vector>

#include "../pgdrive_b200/csrc/pgd_hostpool.h"

int main() {
  for (int workers : {1, 2, 4, 12}) {
    HostPool pool(workers - 1);
    std::vector<long> sum(257, 0);  // plain memory written by whoever takes the item: a race would be reported
    for (int r = 0; r < 400; ++r) {
      if (r % 5 == 0) pool.begin();
      const int items = (r * 37) % 257;
      const std::function<void(int)> job = [&](int i) { sum[(size_t)i] += i + r; };
      pool.run(items, job);
      if (r % 5 == 4) pool.end();
    }
    pool.end();
    long total = 0;
    for (long v : sum) total += v;
    long want = 0;
    for (int r = 0; r < 400; ++r)
      for (int i = 0; i < (r * 37) % 257; ++i) want += i + r;
    if (total != want) {
      printf("hostpool check FAILED: %ld != %ld with %d workers\n", total, want, workers);
      return 1;
    }
  }
  printf("hostpool check ok\n");
  return 0;
}
