/* c_abi_driver.c — exercises libjspace_b200.so through the C ABI exactly the way the Julia wrapper
 * (j-space.jl_b200/julia/JSpaceB200.jl) does: same calls, same argument types, same order, same
 * ownership (one jsb_graph handle shared by the input graph and every result graph, freed once).
 * Test infrastructure; built and run by tests/test_c_driver_gpu.py.  No Python, no ctypes: plain C
 * against include/jspace_b200.h.
 *
 *   c_abi_driver <libdir-unused>            runs the sequence, prints "C-ABI-DRIVER OK <checksum>"
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jspace_b200.h"

#define CHECK(call)                                                                         \
  do {                                                                                      \
    int _rc = (call);                                                                       \
    if (_rc != JSB_OK) {                                                                    \
      fprintf(stderr, "%s:%d %s -> %d: %s\n", __FILE__, __LINE__, #call, _rc, jsb_last_error()); \
      exit(1);                                                                              \
    }                                                                                       \
  } while (0)

static unsigned long long mix(unsigned long long h, unsigned long long v) {
  h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
  return h;
}

/* simulate_evolution method 1 on one stream; mirrors run_one + materialise of the wrapper */
static unsigned long long run_method1(jsb_graph* g, unsigned long long seed, long long stream, int n_cell) {
  double tb[1] = {30.0}, rb[1] = {0.5}, ts[2] = {10.0, 20.0};
  jsb_params p;
  memset(&p, 0, sizeof p);
  p.size = (uint32_t)sizeof p;
  p.method = JSB_METHOD_RANDOM;
  p.Tf = 60.0; p.rate_birth = 0.2; p.rate_death = 0.01; p.rate_migration = 0.01;
  p.mu_driver = 0.01; p.adv_mean = 0.4; p.adv_std = 0.1;
  p.rule = JSB_RULE_CONTACT;
  p.n_start_cells = n_cell;
  p.n_bottleneck = 1; p.t_bottleneck = tb; p.ratio_bottleneck = rb;
  p.n_snapshots = 2; p.t_snapshot = ts;
  jsb_run_opts o;
  memset(&o, 0, sizeof o);
  o.size = (uint32_t)sizeof o;
  o.flags = JSB_OUT_EVENTS | JSB_OUT_LATTICE | JSB_OUT_SERIES;
  o.seed = seed;
  o.g_begin = stream; o.g_end = stream + 1;   /* the wrapper's stream counter */
  jsb_result* res = NULL;
  CHECK(jsb_run(g, &p, 1, stream + 1, &o, &res));

  const jsb_summary* sum = NULL;
  CHECK(jsb_result_summaries(res, &sum));
  const jsb_event* ev = NULL; int64_t ne = 0;
  CHECK(jsb_result_events(res, 0, &ev, &ne));
  const jsb_clone* cl = NULL; int32_t nc = 0;
  CHECK(jsb_result_clones(res, 0, &cl, &nc));
  const uint16_t* lat = NULL; const uint32_t* ids = NULL;
  CHECK(jsb_result_lattice(res, 0, &lat, &ids));
  const jsb_snapshot* sn = NULL; int32_t nsn = 0;
  CHECK(jsb_result_snapshots(res, 0, &sn, &nsn));

  unsigned long long h = 0;
  if ((int64_t)sum->n_events != ne || (int32_t)sum->n_clones != nc) { fprintf(stderr, "summary/accessor mismatch\n"); exit(1); }
  for (int64_t i = 0; i < ne; ++i) { h = mix(h, ev[i].cell); h = mix(h, ev[i].type); h = mix(h, ev[i].site); }
  for (int32_t i = 0; i < nc; ++i) { h = mix(h, cl[i].count); h = mix(h, cl[i].label); }
  int64_t nv = jsb_graph_nv(g), alive = 0;
  for (int64_t v = 0; v < nv; ++v) { alive += lat[v] != 0; h = mix(h, ids[v]); }
  if (alive != sum->n_alive) { fprintf(stderr, "lattice/summary mismatch\n"); exit(1); }
  for (int32_t i = 0; i < nsn; ++i) h = mix(h, sn[i].n_events_before);
  /* n_cell_alive / CA_subpop series (what Start.jl:105-116 writes), reduced on the device */
  const jsb_series_row* ser = NULL; int64_t nser = 0; const int64_t* before = NULL; int32_t nsn2 = 0;
  CHECK(jsb_result_series(res, 0, &ser, &nser, &before, &nsn2));
  if (nsn2 != nsn) { fprintf(stderr, "series/snapshot count mismatch\n"); exit(1); }
  if (nser > 0 && (ser[nser - 1].n_alive & ~JSB_SERIES_MORE) != sum->n_alive) { fprintf(stderr, "series does not end at n_alive\n"); exit(1); }
  for (int64_t i = 0; i < nser; ++i) { h = mix(h, ser[i].n_alive); h = mix(h, ((unsigned long long)ser[i].gain << 16) | ser[i].loss); }
  /* genealogy of the alive cells (what sampling_phylogentic_relation needs) */
  uint32_t* sample = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(alive > 0 ? alive : 1));
  int64_t ns = 0;
  for (int64_t v = 0; v < nv && ns < 10; ++v) if (lat[v]) sample[ns++] = ids[v];
  jsb_relation* rel = NULL; int64_t nrel = 0;
  CHECK(jsb_result_genealogy(res, 0, sample, ns, &rel, &nrel));
  for (int64_t i = 0; i < nrel; ++i) h = mix(h, rel[i].father);
  jsb_free(rel);
  free(sample);
  jsb_result_free(res);   /* the wrapper frees the result before returning */
  return h;
}

int main(void) {
  if (jsb_abi_version() != JSB_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }
  if (jsb_device_count() < 1) { fprintf(stderr, "no CUDA device\n"); return 77; }
  unsigned long long h = 0;

  /* spatial_graph(row, col, seed; dim, n_cell) -> one handle */
  jsb_graph* g = NULL;
  CHECK(jsb_graph_lattice(2, 40, 40, &g));
  const int64_t* cand = NULL; int64_t ncand = 0;
  CHECK(jsb_graph_seed_candidates(g, &cand, &ncand));
  h = mix(h, (unsigned long long)cand[0]);
  /* two simulate_evolution calls on the SAME graph object (successive streams), as a script does */
  unsigned long long a = run_method1(g, 1234, 0, 1);
  unsigned long long b = run_method1(g, 1234, 1, 1);
  unsigned long long a2 = run_method1(g, 1234, 0, 1);   /* same stream again: must reproduce */
  if (a != a2) { fprintf(stderr, "stream 0 did not reproduce\n"); return 1; }
  h = mix(mix(h, a), b);
  /* a second graph alive at the same time (CSR path), results interleaved */
  int64_t rowptr[5] = {0, 1, 3, 5, 6};
  int32_t colidx[6] = {2, 1, 3, 2, 4, 3};
  jsb_graph* g2 = NULL;
  CHECK(jsb_graph_csr(4, rowptr, colidx, &g2));
  int64_t one = 2;
  CHECK(jsb_graph_set_seed_candidates(g2, &one, 1));
  h = mix(h, run_method1(g2, 99, 0, 1));
  h = mix(h, run_method1(g, 1234, 2, 3));   /* n_cell > 1 */
  /* error path: the message must be readable and nothing may leak */
  jsb_run_opts bad;
  memset(&bad, 0, sizeof bad);
  bad.size = 3;
  jsb_result* r = NULL;
  jsb_params p0;
  memset(&p0, 0, sizeof p0);
  if (jsb_run(g, &p0, 1, 1, &bad, &r) == JSB_OK || r != NULL || strlen(jsb_last_error()) == 0) { fprintf(stderr, "bad opts accepted\n"); return 1; }
  /* the wrapper's single finalizer: each handle freed exactly once */
  jsb_graph_free(g2);
  jsb_graph_free(g);
  printf("C-ABI-DRIVER OK %016llx\n", h);
  return 0;
}
