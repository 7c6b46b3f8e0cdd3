/* TEST INFRASTRUCTURE ONLY.  Host (g++) build of the device map generator's source
 * (pgdrive_b200/csrc/pgd_mapgen.cuh, pgd_rng.cuh, pgd_dd.cuh are written for host + device) so that its logic can
 * be checked in a container without a GPU against the reference-pinned Python reset path and against numpy's RNG.
 * The product never loads this library: the product path is the sm_100a kernel in pgdrive_b200/csrc/pgd_mapgen.cu,
 * and tests/test_device_mapgen.py (-m gpu) checks that the kernel's output equals this build's bit for bit.
 */
#include <stdlib.h>
#include <string.h>

#include "../pgdrive_b200/csrc/pgd_mapgen.cuh"

using namespace pgdgen;

extern "C" {

uint64_t pgd_host_hash_seed(uint64_t v) { return hash_seed(v); }

/* ops: 0 = randint(0, arg), 1 = random_sample, 2 = choice(p = probs[0..arg)), 3 = shuffle of arange(arg) (writes arg
 * values), 4 = interval(arg) */
int pgd_host_rng_script(uint64_t seed, const int32_t* ops, const int32_t* args, int n_ops, const double* probs,
                        double* out) {
  MT* mt = (MT*)malloc(sizeof(MT));
  mt_seeded(mt, seed);
  int k = 0;
  for (int i = 0; i < n_ops; ++i) {
    switch (ops[i]) {
      case 0: out[k++] = (double)mt_randint(mt, (uint32_t)args[i]); break;
      case 1: out[k++] = mt_double(mt); break;
      case 2: out[k++] = (double)mt_choice_p(mt, probs, args[i]); break;
      case 3: {
        int n = args[i];
        int* a = (int*)malloc(sizeof(int) * n);
        for (int j = 0; j < n; ++j) a[j] = j;
        for (int j = n - 1; j >= 1; --j) {
          int t = (int)mt_interval(mt, (uint32_t)j);
          int tmp = a[j]; a[j] = a[t]; a[t] = tmp;
        }
        for (int j = 0; j < n; ++j) out[k++] = a[j];
        free(a);
        break;
      }
      default: out[k++] = (double)mt_interval(mt, (uint32_t)args[i]); break;
    }
  }
  free(mt);
  return k;
}

double pgd_host_sin(double x) { return cr_sin(x); }
double pgd_host_cos(double x) { return cr_cos(x); }
double pgd_host_atan2(double y, double x) { return cr_atan2(y, x); }
double pgd_host_atan(double x) { return cr_atan(x); }

int pgd_hostgen(uint64_t seed, const GenConfig* cfg, const GenCaps* caps, PgdMap* map, PgdLane* lanes, PgdRoad* roads,
                PgdBox* boxes, int32_t* cell_start, int32_t* cell_entries, PgdEpisode* episode, PgdSlot* slots,
                int32_t* route_nodes, int32_t* route_roads, int32_t* counts, int32_t* sequence) {
  GenScratch s;
  s.lanes = (GLane*)malloc(sizeof(GLane) * caps->lanes);
  s.roads = (GRoad*)malloc(sizeof(GRoad) * caps->roads);
  s.blocks = (GBlock*)malloc(sizeof(GBlock) * caps->blocks);
  s.boxes = (GBox*)malloc(sizeof(GBox) * caps->boxes);
  s.queue = (int32_t*)malloc(sizeof(int32_t) * 2 * caps->queue);
  s.cand = (int32_t*)malloc(sizeof(int32_t) * (3 * caps->cand + 4 * caps->roads));
  s.mt = (MT*)malloc(sizeof(MT) * 3);
  GenOut out;
  memset(&out, 0, sizeof(out));
  out.map = map; out.lanes = lanes; out.roads = roads; out.boxes = boxes;
  out.cell_start = cell_start; out.cell_entries = cell_entries;
  out.episode = episode; out.slots = slots; out.route_nodes = route_nodes; out.route_roads = route_roads;
  out.counts = counts; out.sequence = sequence;
  int rc = generate_one(seed, *cfg, *caps, s, out);
  free(s.lanes); free(s.roads); free(s.blocks); free(s.boxes); free(s.queue); free(s.cand); free(s.mt);
  return rc;
}

int pgd_host_sizes(int32_t* out) {
  out[0] = (int32_t)sizeof(GenConfig);
  out[1] = (int32_t)sizeof(GenCaps);
  return 2;
}
}
