// jsb_api.cu — host side of libjspace_b200.so: the C ABI of include/jspace_b200.h.
//
// Graph construction and seeding rules (src/J_Space.jl:43-163), parameter validation, device
// workspace/outputs, the kernel launch, and re-materialisation of the reference's outputs
// (event log `df`, clone table set_mut_pop/α, final lattice, ordered lists).
// There is no CPU simulation path in this library.

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <map>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "jsb_device.cuh"

namespace {

thread_local std::string g_err;

// Pinned host blocks are kept between calls: pinning gigabytes costs seconds, the event logs of an ensemble are
// that large, and callers typically run ensembles back to back.  One block per concurrent jsb_run shard (a call
// with devices[] runs one shard per device, each with its own outputs), at most kPinKeep of them.
// No more blocks exist (cached + handed out) than were ever in use at the same time, and never more than kPinKeep
// are cached: a caller that runs one ensemble at a time keeps one block.
constexpr int kPinKeep = 8;
std::mutex g_pin_mu;
struct PinBlock { void* p; size_t bytes; };
std::vector<PinBlock> g_pin;
int g_pin_out = 0, g_pin_out_max = 1;

// JSB_TIMING=1: phase timings of jsb_run on stderr
struct PhaseTimer {
  bool on;
  std::chrono::steady_clock::time_point t0;
  PhaseTimer() : on(std::getenv("JSB_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void mark(const char* what) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[jsb_run] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

// the smallest cached block of at least `bytes`, else a fresh one
void* pinned_take(size_t bytes, size_t* got) {
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    int best = -1;
    for (int i = 0; i < (int)g_pin.size(); ++i)
      if (g_pin[i].bytes >= bytes && (best < 0 || g_pin[i].bytes < g_pin[best].bytes)) best = i;
    ++g_pin_out;
    g_pin_out_max = std::max(g_pin_out_max, g_pin_out);
    if (best >= 0) {
      void* p = g_pin[best].p;
      *got = g_pin[best].bytes;
      g_pin.erase(g_pin.begin() + best);
      return p;
    }
  }
  // ensembles run back to back produce logs of similar but not equal size: leave headroom so the
  // cached buffer fits the next call (pinning gigabytes takes seconds)
  size_t want = bytes + bytes / 4;
  want = (want + (size_t(256) << 20) - 1) & ~((size_t(256) << 20) - 1);
  void* p = nullptr;
  if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      std::lock_guard<std::mutex> lk(g_pin_mu);
      --g_pin_out;
      return nullptr;
    }
  }
  *got = want;
  return p;
}

// the largest cached block, whatever its size (streamed outputs write into it while the kernel is still running)
void* pinned_take_cached(size_t* got) {
  std::lock_guard<std::mutex> lk(g_pin_mu);
  int best = -1;
  for (int i = 0; i < (int)g_pin.size(); ++i)
    if (best < 0 || g_pin[i].bytes > g_pin[best].bytes) best = i;
  if (best < 0) { *got = 0; return nullptr; }
  ++g_pin_out;
  g_pin_out_max = std::max(g_pin_out_max, g_pin_out);
  void* p = g_pin[best].p;
  *got = g_pin[best].bytes;
  g_pin.erase(g_pin.begin() + best);
  return p;
}

void pinned_give(void* p, size_t bytes) {
  if (!p) return;
  std::vector<void*> drop;
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    if (g_pin_out > 0) --g_pin_out;
    g_pin.push_back({p, bytes});
    while ((int)g_pin.size() > kPinKeep || (int)g_pin.size() + g_pin_out > g_pin_out_max) {  // drop the smallest
      int worst = 0;
      for (int i = 1; i < (int)g_pin.size(); ++i)
        if (g_pin[i].bytes < g_pin[worst].bytes) worst = i;
      drop.push_back(g_pin[worst].p);
      g_pin.erase(g_pin.begin() + worst);
    }
  }
  for (void* d : drop) cudaFreeHost(d);
}

void pinned_drop_all() {
  std::vector<PinBlock> all;
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    all.swap(g_pin);
  }
  for (auto& b : all) cudaFreeHost(b.p);
}

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      int _c = (_e == cudaErrorMemoryAllocation) ? JSB_ENOMEM : JSB_ECUDA;                    \
      rc = fail(_c, "%s failed: %s", #expr, cudaGetErrorString(_e));                          \
      goto done;                                                                              \
    }                                                                                         \
  } while (0)

}  // namespace

// ------------------------------------------------------------------------------------------------
// Graph
// ------------------------------------------------------------------------------------------------
struct jsb_graph {
  int kind = 0;  // 0 lattice, 1 csr
  uint32_t n1 = 0, n2 = 0, n3 = 0;
  int64_t nv = 0, ne = 0;
  int64_t lattice_row = 0, lattice_col = 0;
  std::vector<uint32_t> rowptr, colidx;  // csr, 0-based
  bool have_cand = false;
  std::vector<int64_t> cand;  // 1-based
  // One copy per device, created on first use under `mu` (a graph may be shared by threads that run on
  // different devices); the seed candidates are re-uploaded when they change (cand_version).
  struct DevCopy {
    uint32_t *d_rowptr = nullptr, *d_colidx = nullptr, *d_cand = nullptr;
    uint32_t d_ncand = 0;
    uint64_t cand_version = ~0ull;
  };
  std::mutex mu;
  std::map<int, DevCopy> dev;
  uint64_t cand_version = 0;

  void drop_devices() {
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& kv : dev) {
      cudaSetDevice(kv.first);
      cudaFree(kv.second.d_rowptr);
      cudaFree(kv.second.d_colidx);
      cudaFree(kv.second.d_cand);
    }
    dev.clear();
    cudaSetDevice(cur);
  }
};

namespace {

// neighbours of 0-based vertex v, ascending, into out[]; returns degree
int host_neighbors(const jsb_graph& g, uint32_t v, uint32_t* out, int cap) {
  if (g.kind == 0) {
    uint32_t x = v % g.n1, yz = v / g.n1, y = yz % g.n2, z = yz / g.n2;
    uint32_t plane = g.n1 * g.n2;
    int d = 0;
    auto put = [&](uint32_t w) { if (d < cap) out[d] = w; ++d; };
    if (z > 0) put(v - plane);
    if (y > 0) put(v - g.n1);
    if (x > 0) put(v - 1);
    if (x + 1 < g.n1) put(v + 1);
    if (y + 1 < g.n2) put(v + g.n1);
    if (z + 1 < g.n3) put(v + plane);
    return d;
  }
  uint32_t a = g.rowptr[v], b = g.rowptr[v + 1];
  for (uint32_t e = a; e < b; ++e)
    if ((int)(e - a) < cap) out[e - a] = g.colidx[e];
  return (int)(b - a);
}

// Graphs.closeness_centrality(g) (normalize = true) and the argmax set, src/J_Space.jl:83-85.
// One BFS per vertex, vertices split over host threads.
void closeness_argmax(const jsb_graph& g, std::vector<int64_t>& out) {
  const int64_t nv = g.nv;
  std::vector<double> c((size_t)nv, 0.0);
  unsigned nthreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
  if (nv < 256) nthreads = 1;
  std::atomic<int64_t> next{0};
  auto worker = [&]() {
    std::vector<int32_t> dist((size_t)nv), queue((size_t)nv);
    uint32_t nb[64];
    std::vector<uint32_t> big;
    for (;;) {
      int64_t u = next.fetch_add(1);
      if (u >= nv) break;
      std::fill(dist.begin(), dist.end(), -1);
      int64_t head = 0, tail = 0, reached = 0;
      double sigma = 0.0;
      queue[tail++] = (int32_t)u;
      dist[u] = 0;
      while (head < tail) {
        int32_t v = queue[head++];
        ++reached;
        sigma += (double)dist[v];
        const uint32_t* p = nb;
        int d = host_neighbors(g, (uint32_t)v, nb, 64);
        if (d > 64) {
          big.resize(d);
          host_neighbors(g, (uint32_t)v, big.data(), d);
          p = big.data();
        }
        for (int i = 0; i < d; ++i) {
          uint32_t w = p[i];
          if (dist[w] < 0) { dist[w] = dist[v] + 1; queue[tail++] = (int32_t)w; }
        }
      }
      double l = (double)(reached - 1);
      if (sigma > 0) {
        double cu = l / sigma;
        cu *= l / (double)(nv - 1);
        c[u] = cu;
      }
    }
  };
  std::vector<std::thread> th;
  for (unsigned i = 1; i < nthreads; ++i) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
  double best = c[0];
  for (int64_t u = 1; u < nv; ++u) best = std::max(best, c[u]);
  out.clear();
  for (int64_t u = 0; u < nv; ++u)
    if (c[u] == best) out.push_back(u + 1);
}

int finish_csr(jsb_graph* g, int64_t nv, const std::vector<std::vector<uint32_t>>* adj, const int64_t* rowptr,
               const int32_t* colidx) {
  if (nv < 1 || nv >= (int64_t(1) << 28)) return fail(JSB_ECAPACITY, "graph has %lld vertices; supported: 1..2^28-1", (long long)nv);
  g->kind = 1;
  g->nv = nv;
  g->rowptr.assign((size_t)nv + 1, 0);
  if (adj) {
    for (int64_t v = 0; v < nv; ++v) g->rowptr[v + 1] = g->rowptr[v] + (uint32_t)(*adj)[v].size();
    g->colidx.reserve(g->rowptr[nv]);
    for (int64_t v = 0; v < nv; ++v) g->colidx.insert(g->colidx.end(), (*adj)[v].begin(), (*adj)[v].end());
  } else {
    if (rowptr[0] != 0) return fail(JSB_EINVAL, "rowptr[0] must be 0");
    for (int64_t v = 0; v < nv; ++v) {
      if (rowptr[v + 1] < rowptr[v]) return fail(JSB_EINVAL, "rowptr is not non-decreasing at vertex %lld", (long long)v + 1);
      if (rowptr[v + 1] > 0xFFFFFFF0ll) return fail(JSB_ECAPACITY, "too many edges");
      g->rowptr[v + 1] = (uint32_t)rowptr[v + 1];
    }
    g->colidx.resize(g->rowptr[nv]);
    for (int64_t v = 0; v < nv; ++v) {
      int64_t prev = 0;
      for (int64_t e = rowptr[v]; e < rowptr[v + 1]; ++e) {
        int64_t w = colidx[e];
        if (w < 1 || w > nv) return fail(JSB_EINVAL, "colidx[%lld] = %lld is outside 1..nv", (long long)e, (long long)w);
        if (w == v + 1) return fail(JSB_EINVAL, "self-loop at vertex %lld is not supported", (long long)w);
        if (w <= prev) return fail(JSB_EINVAL, "neighbours of vertex %lld are not strictly ascending", (long long)v + 1);
        prev = w;
        g->colidx[e] = (uint32_t)(w - 1);
      }
    }
  }
  g->ne = (int64_t)g->colidx.size() / 2;
  return JSB_OK;
}

}  // namespace

extern "C" {

int jsb_abi_version(void) { return JSB_ABI_VERSION; }

int jsb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* jsb_last_error(void) { return g_err.c_str(); }

int jsb_graph_lattice(int dim, int64_t row, int64_t col, jsb_graph** out) {
  if (!out) return fail(JSB_EINVAL, "out is null");
  *out = nullptr;
  if (row < 1 || col < 1) return fail(JSB_EINVAL, "row and col must be >= 1 (got %lld, %lld)", (long long)row, (long long)col);
  if ((double)row * (double)col >= 268435456.0) return fail(JSB_ECAPACITY, "lattice has more than 2^28-1 sites");
  jsb_graph* g = new (std::nothrow) jsb_graph();
  if (!g) return fail(JSB_ENOMEM, "out of host memory");
  g->kind = 0;
  g->lattice_row = row;
  g->lattice_col = col;
  if (dim == 2) {  // Graphs.grid((row, col, 1)) src/J_Space.jl:54
    g->n1 = (uint32_t)row; g->n2 = (uint32_t)col; g->n3 = 1;
  } else {         // dim == 3, and the "dim not correct, it set 3" branch :50-51 / :58-59
    int64_t d = (int64_t)std::trunc(std::pow((double)(row * col), 1.0 / 3.0)) + 1;
    if ((double)d * d * d >= 268435456.0) { delete g; return fail(JSB_ECAPACITY, "lattice has more than 2^28-1 sites"); }
    g->n1 = g->n2 = g->n3 = (uint32_t)d;
  }
  g->nv = (int64_t)g->n1 * g->n2 * g->n3;
  g->ne = (int64_t)(g->n1 - 1) * g->n2 * g->n3 + (int64_t)g->n1 * (g->n2 - 1) * g->n3 + (int64_t)g->n1 * g->n2 * (g->n3 - 1);
  // seed vertex, src/J_Space.jl:131-133 (uses the row/col arguments even for dim 3)
  double r = (double)row, c = (double)col;
  double v = (g->nv % 2 == 0) ? (c * std::floor(r / 2.0)) + (std::floor(c / 2.0) + 1.0) : std::floor((double)g->nv / 2.0) + 1.0;
  int64_t v0 = (int64_t)v;
  if ((double)v0 == v && v0 >= 1 && v0 <= g->nv) g->cand.push_back(v0);
  g->have_cand = true;
  *out = g;
  return JSB_OK;
}

int jsb_graph_csr(int64_t nv, const int64_t* rowptr, const int32_t* colidx, jsb_graph** out) {
  if (!out) return fail(JSB_EINVAL, "out is null");
  *out = nullptr;
  if (!rowptr || (!colidx && nv > 0 && rowptr[nv] > 0)) return fail(JSB_EINVAL, "rowptr/colidx is null");
  jsb_graph* g = new (std::nothrow) jsb_graph();
  if (!g) return fail(JSB_ENOMEM, "out of host memory");
  int rc = finish_csr(g, nv, nullptr, rowptr, colidx);
  if (rc != JSB_OK) { delete g; return rc; }
  *out = g;
  return JSB_OK;
}

int jsb_graph_dense_file(const char* path, jsb_graph** out) {
  if (!out) return fail(JSB_EINVAL, "out is null");
  *out = nullptr;
  if (!path) return fail(JSB_EINVAL, "path is null");
  FILE* f = std::fopen(path, "r");
  if (!f) return fail(JSB_EIO, "cannot open adjacency matrix '%s'", path);
  std::vector<long long> vals;
  long long x;
  int got;
  while ((got = std::fscanf(f, "%lld", &x)) == 1) vals.push_back(x);
  bool trailing = (got != EOF);
  std::fclose(f);
  if (trailing) return fail(JSB_EINVAL, "'%s': non-integer token in adjacency matrix", path);
  int64_t n = (int64_t)std::llround(std::sqrt((double)vals.size()));
  if (n < 1 || n * n != (int64_t)vals.size()) return fail(JSB_EINVAL, "'%s': %zu entries do not form a square matrix", path, vals.size());
  std::vector<std::vector<uint32_t>> adj((size_t)n);
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j < n; ++j) {
      bool a = vals[i * n + j] != 0, b = vals[j * n + i] != 0;
      if (a != b) return fail(JSB_EINVAL, "'%s': adjacency matrix is not symmetric at (%lld,%lld)", path, (long long)i + 1, (long long)j + 1);
      if (a && i == j) return fail(JSB_EINVAL, "'%s': self-loop at vertex %lld is not supported", path, (long long)i + 1);
      if (a) adj[i].push_back((uint32_t)j);
    }
  jsb_graph* g = new (std::nothrow) jsb_graph();
  if (!g) return fail(JSB_ENOMEM, "out of host memory");
  int rc = finish_csr(g, n, &adj, nullptr, nullptr);
  if (rc != JSB_OK) { delete g; return rc; }
  *out = g;
  return JSB_OK;
}

int64_t jsb_graph_nv(const jsb_graph* g) { return g ? g->nv : 0; }
int64_t jsb_graph_ne(const jsb_graph* g) { return g ? g->ne : 0; }

int jsb_graph_seed_candidates(jsb_graph* g, const int64_t** sites, int64_t* n) {
  if (!g || !sites || !n) return fail(JSB_EINVAL, "null argument");
  if (!g->have_cand) {
    closeness_argmax(*g, g->cand);
    g->have_cand = true;
    g->cand_version++;
  }
  *sites = g->cand.data();
  *n = (int64_t)g->cand.size();
  return JSB_OK;
}

int jsb_graph_set_seed_candidates(jsb_graph* g, const int64_t* sites, int64_t n) {
  if (!g || (n > 0 && !sites) || n < 0) return fail(JSB_EINVAL, "null argument");
  for (int64_t i = 0; i < n; ++i)
    if (sites[i] < 1 || sites[i] > g->nv) return fail(JSB_EINVAL, "seed candidate %lld is outside 1..nv", (long long)sites[i]);
  std::lock_guard<std::mutex> lk(g->mu);
  g->cand.assign(sites, sites + n);
  g->have_cand = true;
  g->cand_version++;
  return JSB_OK;
}

int jsb_graph_neighbors(const jsb_graph* g, int64_t site, int64_t* out, int cap) {
  if (!g || site < 1 || site > g->nv) return fail(JSB_EINVAL, "site outside 1..nv");
  std::vector<uint32_t> tmp((size_t)std::max(cap, 8));
  int d = host_neighbors(*g, (uint32_t)(site - 1), tmp.data(), (int)tmp.size());
  if (d > (int)tmp.size()) { tmp.resize(d); host_neighbors(*g, (uint32_t)(site - 1), tmp.data(), d); }
  for (int i = 0; i < d && i < cap; ++i) out[i] = (int64_t)tmp[i] + 1;
  return d;
}

void jsb_graph_free(jsb_graph* g) {
  if (!g) return;
  g->drop_devices();
  delete g;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Results
// ------------------------------------------------------------------------------------------------
struct jsb_result {
  int64_t n_local = 0, n_points = 0, reps_per_point = 0, g_begin = 0;
  uint32_t nv = 0;
  uint32_t flags = 0;
  int max_snap = 0;
  std::vector<jsb_summary> summaries;
  std::vector<long long> clone_off;
  std::vector<jsb_clone> clones;
  // one pinned host block (cached between calls): [events | lattice clone | lattice cell | lists]
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  jsb_event* events = nullptr;  // replicate-major
  std::vector<uint64_t> ev_off;  // n_local + 1: first row of every replicate (streamed outputs: in finish order)
  std::vector<uint64_t> ev_cnt;  // n_local: rows of every replicate
  std::vector<uint32_t> lat_slot; // streamed outputs: lattice / lists of replicate i sit in slot lat_slot[i] (its finish rank); empty: slot i
  uint16_t* lat_clone = nullptr;
  uint32_t* lat_cell = nullptr;
  uint32_t* lists = nullptr;
  std::vector<jsb_snapshot> snaps;
  std::vector<int32_t> snap_count;
  std::vector<jsb_series_row> series;   // JSB_OUT_SERIES: rows of every replicate, replicate-major
  std::vector<uint64_t> ser_off;        // n_local + 1
  std::vector<int64_t> ser_snap;        // [n_local][max_snap] rows before every snapshot
  double kernel_ms = 0.0;
  int64_t launches = 0;
  // a multi-device call: one result per device shard (this object then only holds the merged summaries)
  std::vector<jsb_result*> parts;
  std::vector<int64_t> part_off;  // first local index of every part
  ~jsb_result();
};

jsb_result::~jsb_result() {
  pinned_give(pinned, pinned_bytes);
  for (jsb_result* p : parts) delete p;
}

namespace {
// the shard that holds local replicate i of a (possibly multi-device) result; i becomes the index inside it
inline jsb_result* part_of(jsb_result* r, int64_t& i) {
  if (r->parts.empty()) return r;
  size_t k = (size_t)(std::upper_bound(r->part_off.begin(), r->part_off.end(), i) - r->part_off.begin()) - 1;
  i -= r->part_off[k];
  return r->parts[k];
}
inline const jsb_result* part_of(const jsb_result* r, int64_t& i) { return part_of(const_cast<jsb_result*>(r), i); }
}  // namespace

namespace jsb {

// Gathers the pool chunks of every replicate into one replicate-major array (coalesced 16-byte
// copies), so that a single device-to-host copy moves exactly the rows that exist.
__global__ void gather_log_kernel(const jsb_event* __restrict__ pool, const int32_t* __restrict__ owner,
                                  const uint32_t* __restrict__ seq, const unsigned long long* __restrict__ ev_off,
                                  uint32_t n_used, jsb_event* __restrict__ out) {
  for (uint32_t c = blockIdx.x; c < n_used; c += gridDim.x) {
    const int32_t o = owner[c];
    const unsigned long long begin = ev_off[o], end = ev_off[o + 1];
    const unsigned long long first = begin + (unsigned long long)seq[c] * kLogChunk;
    if (first >= end) continue;
    const unsigned long long cnt = (end - first) < (unsigned long long)kLogChunk ? (end - first) : (unsigned long long)kLogChunk;
    const uint4* src = reinterpret_cast<const uint4*>(pool + (size_t)c * kLogChunk);
    uint4* dst = reinterpret_cast<uint4*>(out + first);
    for (unsigned long long i = threadIdx.x; i < cnt * 2; i += blockDim.x) dst[i] = src[i];
  }
}

// JSB_OUT_SERIES: the packed log of one replicate reduced to its series rows (include/jspace_b200.h), one thread
// per replicate; WRITE = false only counts.  Rows flagged JSB_EVF_GROUP_START open one push: a Death row takes a
// cell from its clone, the (Duplicate | Mutation, Duplicate) pair that closes a birth gives one to the clone of its
// first row, a Migration changes nothing; excision Deaths are one push written as one row per cell.
template <bool WRITE>
__global__ void series_kernel(const jsb_event* __restrict__ ev, const unsigned long long* __restrict__ ev_off,
                              const unsigned long long* __restrict__ ev_cnt, const uint32_t* __restrict__ n0,
                              const jsb_snapshot* __restrict__ snaps, int max_snap, long long n_local,
                              unsigned long long* __restrict__ ser_cnt, const unsigned long long* __restrict__ ser_off,
                              jsb_series_row* __restrict__ out, long long* __restrict__ snap_rows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_local) return;
  const jsb_event* e = ev + ev_off[i];
  const unsigned long long m = ev_cnt[i];
  jsb_series_row* o = WRITE ? out + ser_off[i] : nullptr;
  unsigned long long rows = 0;
  uint32_t n = n0[i];
  int si = 0;
  unsigned long long k = 0;
  while (k < m) {
    // the group [k, ge)
    unsigned long long ge = k + 1;
    while (ge < m && !(e[ge].flags & JSB_EVF_GROUP_START)) ++ge;
    if (WRITE && snaps)
      while (si < max_snap && snaps[(size_t)i * max_snap + si].index != 0xFFFFFFFFu &&
             snaps[(size_t)i * max_snap + si].n_events_before <= k) snap_rows[(size_t)i * max_snap + si++] = (long long)rows;
    if (e[k].flags & JSB_EVF_EXCISION) {
      for (unsigned long long q = k; q < ge; ++q) {
        --n;
        if (WRITE) { jsb_series_row r; r.n_alive = n | (q + 1 < ge ? JSB_SERIES_MORE : 0u); r.gain = 0; r.loss = e[q].clone; o[rows] = r; }
        ++rows;
      }
    } else {
      uint16_t gain = 0, loss = 0;
      for (unsigned long long q = k; q < ge; ++q)
        if (e[q].type == JSB_EV_DEATH) { loss = e[q].clone; --n; }
      if (ge - k >= 2 && e[ge - 1].type == JSB_EV_DUPLICATE && (e[ge - 2].type == JSB_EV_DUPLICATE || e[ge - 2].type == JSB_EV_MUTATION)) {
        gain = e[ge - 2].clone;
        ++n;
      }
      if (WRITE) { jsb_series_row r; r.n_alive = n; r.gain = gain; r.loss = loss; o[rows] = r; }
      ++rows;
    }
    k = ge;
  }
  if (WRITE && snaps)
    while (si < max_snap && snaps[(size_t)i * max_snap + si].index != 0xFFFFFFFFu) snap_rows[(size_t)i * max_snap + si++] = (long long)rows;
  if (!WRITE) ser_cnt[i] = rows;
}

}  // namespace jsb

namespace {

int validate_params(const jsb_graph* g, const jsb_params& p, int64_t idx) {
  if (p.size != sizeof(jsb_params)) return fail(JSB_EINVAL, "points[%lld].size = %u, expected %zu (ABI mismatch)", (long long)idx, p.size, sizeof(jsb_params));
  if (p.method != JSB_METHOD_RANDOM && p.method != JSB_METHOD_TREE) return fail(JSB_EINVAL, "points[%lld].method must be 1 or 2", (long long)idx);
  if (p.rule < 0 || p.rule > 3) return fail(JSB_EINVAL, "points[%lld].rule must be 0..3", (long long)idx);
  auto bad = [](double x) { return !(x >= 0.0) || std::isinf(x); };
  if (std::isnan(p.Tf)) return fail(JSB_EINVAL, "points[%lld].Tf is NaN", (long long)idx);
  if (bad(p.rate_death) || bad(p.rate_migration) || bad(p.mu_driver)) return fail(JSB_EINVAL, "points[%lld]: rates must be finite and >= 0", (long long)idx);
  if (p.method == JSB_METHOD_RANDOM && (bad(p.rate_birth) || std::isnan(p.adv_mean) || bad(p.adv_std)))
    return fail(JSB_EINVAL, "points[%lld]: rate_birth/adv_std must be finite and >= 0", (long long)idx);
  if (p.n_start_cells < 0 || p.n_start_cells > g->nv) return fail(JSB_EINVAL, "points[%lld].n_start_cells must be in 0..nv", (long long)idx);
  if (p.n_bottleneck < 0 || p.n_snapshots < 0) return fail(JSB_EINVAL, "points[%lld]: negative array length", (long long)idx);
  if (p.n_bottleneck > 0 && (!p.t_bottleneck || !p.ratio_bottleneck)) return fail(JSB_EINVAL, "points[%lld]: t_bottleneck/ratio_bottleneck is null", (long long)idx);
  for (int i = 0; i < p.n_bottleneck; ++i)
    if (!(p.ratio_bottleneck[i] >= 0.0 && p.ratio_bottleneck[i] <= 1.0))
      return fail(JSB_EINVAL, "points[%lld].ratio_bottleneck[%d] must be in [0,1] (the reference's sample() throws otherwise)", (long long)idx, i);
  if (p.n_snapshots > 0 && !p.t_snapshot) return fail(JSB_EINVAL, "points[%lld].t_snapshot is null", (long long)idx);
  if (p.method == JSB_METHOD_TREE) {
    if (p.n_tree_nodes < 1 || p.n_tree_nodes > jsb::kMaxClones) return fail(JSB_EINVAL, "points[%lld].n_tree_nodes must be in 1..1000", (long long)idx);
    if (p.n_tree_edges < 0 || (p.n_tree_edges > 0 && (!p.edge_parent || !p.edge_child))) return fail(JSB_EINVAL, "points[%lld]: edge list is null", (long long)idx);
    if (!p.node_rate) return fail(JSB_EINVAL, "points[%lld].node_rate is null", (long long)idx);
    if (p.root_node < 0 || p.root_node >= p.n_tree_nodes) return fail(JSB_EINVAL, "points[%lld].root_node out of range", (long long)idx);
    for (int e = 0; e < p.n_tree_edges; ++e)
      if (p.edge_parent[e] < 0 || p.edge_parent[e] >= p.n_tree_nodes || p.edge_child[e] < 0 || p.edge_child[e] >= p.n_tree_nodes)
        return fail(JSB_EINVAL, "points[%lld]: edge %d names a node outside 0..n_tree_nodes-1", (long long)idx, e);
    for (int i = 0; i < p.n_tree_nodes; ++i)
      if (bad(p.node_rate[i])) return fail(JSB_EINVAL, "points[%lld].node_rate[%d] must be finite and >= 0", (long long)idx, i);
    if (bad(p.root_rate)) return fail(JSB_EINVAL, "points[%lld].root_rate must be finite and >= 0", (long long)idx);
  }
  return JSB_OK;
}

// The copy of the graph on `device` (the current device), created / refreshed under the graph's lock.
int upload_graph(jsb_graph* g, int device, jsb_graph::DevCopy* out) {
  int rc = JSB_OK;
  std::lock_guard<std::mutex> lk(g->mu);
  jsb_graph::DevCopy& d = g->dev[device];
  if (g->kind == 1 && !d.d_rowptr) {
    CUDA_TRY(cudaMalloc(&d.d_rowptr, sizeof(uint32_t) * g->rowptr.size()));
    CUDA_TRY(cudaMalloc(&d.d_colidx, sizeof(uint32_t) * std::max<size_t>(1, g->colidx.size())));
    CUDA_TRY(cudaMemcpy(d.d_rowptr, g->rowptr.data(), sizeof(uint32_t) * g->rowptr.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.d_colidx, g->colidx.data(), sizeof(uint32_t) * g->colidx.size(), cudaMemcpyHostToDevice));
  }
  if (d.cand_version != g->cand_version) {
    cudaFree(d.d_cand);
    d.d_cand = nullptr;
    std::vector<uint32_t> c0(g->cand.size());
    for (size_t i = 0; i < c0.size(); ++i) c0[i] = (uint32_t)(g->cand[i] - 1);
    CUDA_TRY(cudaMalloc(&d.d_cand, sizeof(uint32_t) * std::max<size_t>(1, c0.size())));
    CUDA_TRY(cudaMemcpy(d.d_cand, c0.data(), sizeof(uint32_t) * c0.size(), cudaMemcpyHostToDevice));
    d.d_ncand = (uint32_t)c0.size();
    d.cand_version = g->cand_version;
  }
  *out = d;
done:
  return rc;
}

// Per-device state: one lock that serialises jsb_run on a device (launch configuration of the kernels and the
// cached buffers below are per device), and the large device buffers kept between calls.
// (the small per-call tables are cached too: every cudaMalloc / cudaFree / cudaHostAlloc takes the driver lock, and a
// monitoring tool polling the same GPU — nvidia-smi in bench.py — can hold that lock for tens of milliseconds)
enum { BUF_WS, BUF_LOG, BUF_CHUNK_OWNER, BUF_CHUNK_SEQ, BUF_GATHER, BUF_OUT_CLONE, BUF_OUT_CELL, BUF_OUT_LISTS, BUF_CLONES, BUF_SUSP,
       BUF_POINTS, BUF_DPOOL, BUF_IPOOL, BUF_SUM, BUF_COUNTERS, BUF_CLONE_OFF, BUF_CHUNK_NEXT, BUF_SNAPS, BUF_FIN, BUF_FIN_CTR,
       BUF_CHUNK_PREV, BUF_PRIO, BUF_ORDER, BUF_N };
struct DeviceState {
  std::mutex mu;
  struct Buf { void* p = nullptr; size_t bytes = 0; } buf[BUF_N];
  void* h_scratch = nullptr;   // small pinned block (finish list of the streamed outputs)
  size_t h_scratch_bytes = 0;
  cudaStream_t stream = nullptr, stream2 = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // answers of driver queries that do not change (asked once: a monitoring tool polling the GPU holds the lock they need)
  bool have_prop = false, stack_ok = false;
  cudaDeviceProp prop;
};
DeviceState g_devs[64];

// a cached device buffer of at least `bytes` (grown with 1/8 headroom); nullptr + *err on failure
void* dev_buf(DeviceState& ds, int slot, size_t bytes, cudaError_t* err) {
  DeviceState::Buf& b = ds.buf[slot];
  *err = cudaSuccess;
  if (b.p && b.bytes >= bytes) return b.p;
  if (b.p) { cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
  size_t want = bytes + bytes / 8;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) { cudaGetLastError(); want = bytes; e = cudaMalloc(&b.p, want); }
  if (e != cudaSuccess) { cudaGetLastError(); b.p = nullptr; *err = e; return nullptr; }
  b.bytes = want;
  return b.p;
}
#define DEV_BUF(ptr, type, slot, bytes)                                                       \
  do {                                                                                        \
    cudaError_t _e;                                                                           \
    ptr = (type)dev_buf(ds, slot, (bytes), &_e);                                              \
    if (!ptr) {                                                                               \
      rc = fail(_e == cudaErrorMemoryAllocation ? JSB_ENOMEM : JSB_ECUDA, "device allocation of %zu bytes failed: %s", (size_t)(bytes), cudaGetErrorString(_e)); \
      goto done;                                                                              \
    }                                                                                         \
  } while (0)

}  // namespace

namespace {
int run_single(jsb_graph* g, const jsb_params* points, int64_t n_points, int64_t reps_per_point, const jsb_run_opts* opts,
               int device, int64_t g_begin, int64_t g_end, jsb_result** out);
}

extern "C" {

int jsb_run(jsb_graph* g, const jsb_params* points, int64_t n_points, int64_t reps_per_point,
            const jsb_run_opts* opts, jsb_result** out) {
  using namespace jsb;
  if (!out) return fail(JSB_EINVAL, "out is null");
  *out = nullptr;
  if (!g || !points || !opts) return fail(JSB_EINVAL, "null argument");
  if (opts->size != sizeof(jsb_run_opts)) return fail(JSB_EINVAL, "opts.size = %u, expected %zu (ABI mismatch)", opts->size, sizeof(jsb_run_opts));
  if (n_points < 1 || reps_per_point < 1) return fail(JSB_EINVAL, "n_points and reps_per_point must be >= 1");
  if ((double)n_points * (double)reps_per_point > 4294967295.0) return fail(JSB_ECAPACITY, "more than 2^32-1 replicates");
  const int64_t g_total = n_points * reps_per_point;
  int64_t g_begin = opts->g_begin, g_end = opts->g_end == 0 ? g_total : opts->g_end;
  if ((opts->flags & JSB_MODE_REJECTION_FREE) && (opts->flags & JSB_OUT_LISTS)) return fail(JSB_EINVAL, "JSB_OUT_LISTS is not available with JSB_MODE_REJECTION_FREE (no ordered lists exist in that mode)");
  if ((opts->flags & JSB_MODE_REJECTION_FREE) && g->kind == 1) {
    uint32_t maxdeg = 0;
    for (int64_t v = 0; v < g->nv; ++v) maxdeg = std::max(maxdeg, g->rowptr[v + 1] - g->rowptr[v]);
    if (maxdeg > 31) return fail(JSB_ECAPACITY, "JSB_MODE_REJECTION_FREE supports vertex degrees up to 31 (graph has %u)", maxdeg);
  }
  if (g_begin < 0 || g_end > g_total || g_begin > g_end) return fail(JSB_EINVAL, "replicate range [%lld,%lld) is outside [0,%lld)", (long long)g_begin, (long long)g_end, (long long)g_total);
  bool need_closeness = false;
  for (int64_t i = 0; i < n_points; ++i) {
    int rcv = validate_params(g, points[i], i);
    if (rcv != JSB_OK) return rcv;
    if (points[i].n_start_cells == 1) need_closeness = true;
  }
  int ndev = jsb_device_count();
  if (ndev <= 0) return fail(JSB_ENODEVICE, "no CUDA device available; libjspace_b200 has no CPU fallback");
  std::vector<int> devs;
  if (opts->n_devices > 0) {
    if (!opts->devices) return fail(JSB_EINVAL, "opts.devices is null");
    for (int k = 0; k < opts->n_devices; ++k) {
      const int d = opts->devices[k];
      if (d < 0 || d >= ndev || d >= 64) return fail(JSB_EINVAL, "device %d is outside 0..%d", d, ndev - 1);
      if (std::find(devs.begin(), devs.end(), d) != devs.end()) return fail(JSB_EINVAL, "device %d is listed twice", d);
      devs.push_back(d);
    }
  } else {
    if (opts->device < 0 || opts->device >= ndev || opts->device >= 64) return fail(JSB_EINVAL, "device %d is outside 0..%d", opts->device, ndev - 1);
    devs.push_back(opts->device);
  }
  if (need_closeness) {
    // exact closeness centrality is one BFS per vertex on the host, O(V*E): seconds for 10^4 vertices, hours for
    // 10^6.  Large graphs must come with their seed candidates (jsb_graph_set_seed_candidates).
    std::unique_lock<std::mutex> lk(g->mu);
    if (!g->have_cand) {
      if (g->nv > 50000) return fail(JSB_EINVAL, "graph has %lld vertices and no seed candidates: the closeness arg-max of src/J_Space.jl:83-85 is O(V*E) on the host; call jsb_graph_set_seed_candidates (or jsb_graph_seed_candidates once, to compute and cache it)", (long long)g->nv);
      lk.unlock();
      const int64_t* sv;
      int64_t sn;
      jsb_graph_seed_candidates(g, &sv, &sn);
    }
  }
  if (devs.size() == 1 || g_end - g_begin < 2) return run_single(g, points, n_points, reps_per_point, opts, devs[0], g_begin, g_end, out);

  // ---- several devices: contiguous shards, one host thread per device ----------------------------------
  const int nd = (int)std::min<int64_t>((int64_t)devs.size(), g_end - g_begin);
  std::vector<jsb_result*> parts((size_t)nd, nullptr);
  std::vector<int> rcs((size_t)nd, JSB_OK);
  std::vector<std::string> errs((size_t)nd);
  std::vector<int64_t> cut((size_t)nd + 1);
  const int64_t total = g_end - g_begin, base = total / nd, extra = total % nd;
  cut[0] = g_begin;
  for (int k = 0; k < nd; ++k) cut[k + 1] = cut[k] + base + (k < extra ? 1 : 0);
  {
    std::vector<std::thread> th;
    for (int k = 0; k < nd; ++k)
      th.emplace_back([&, k]() {
        rcs[k] = run_single(g, points, n_points, reps_per_point, opts, devs[k], cut[k], cut[k + 1], &parts[k]);
        if (rcs[k] != JSB_OK) errs[k] = g_err;  // g_err is thread-local
      });
    for (auto& t : th) t.join();
  }
  for (int k = 0; k < nd; ++k)
    if (rcs[k] != JSB_OK) {
      for (jsb_result* p : parts) delete p;
      return fail(rcs[k], "device %d: %s", devs[k], errs[k].c_str());
    }
  jsb_result* res = new (std::nothrow) jsb_result();
  if (!res) { for (jsb_result* p : parts) delete p; return fail(JSB_ENOMEM, "out of host memory"); }
  res->n_local = total; res->n_points = n_points; res->reps_per_point = reps_per_point; res->g_begin = g_begin;
  res->nv = (uint32_t)g->nv; res->flags = opts->flags;
  for (int k = 0; k < nd; ++k) {
    res->part_off.push_back(cut[k] - g_begin);
    res->summaries.insert(res->summaries.end(), parts[k]->summaries.begin(), parts[k]->summaries.end());
    res->kernel_ms = std::max(res->kernel_ms, parts[k]->kernel_ms);  // the devices ran side by side
    res->launches += parts[k]->launches;
    res->max_snap = std::max(res->max_snap, parts[k]->max_snap);
  }
  res->parts = parts;
  *out = res;
  return JSB_OK;
}

void jsb_trim(void) {
  int cur = 0;
  cudaGetDevice(&cur);
  const int ndev = jsb_device_count();
  for (int d = 0; d < ndev && d < 64; ++d) {
    std::lock_guard<std::mutex> lk(g_devs[d].mu);
    bool any = g_devs[d].stream || g_devs[d].h_scratch;
    for (auto& b : g_devs[d].buf) any = any || b.p;
    if (!any) continue;
    cudaSetDevice(d);
    for (auto& b : g_devs[d].buf) { cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
    DeviceState& t = g_devs[d];
    if (t.h_scratch) { cudaFreeHost(t.h_scratch); t.h_scratch = nullptr; t.h_scratch_bytes = 0; }
    if (t.stream) { cudaStreamDestroy(t.stream); t.stream = nullptr; }
    if (t.stream2) { cudaStreamDestroy(t.stream2); t.stream2 = nullptr; }
    if (t.ev0) { cudaEventDestroy(t.ev0); t.ev0 = nullptr; }
    if (t.ev1) { cudaEventDestroy(t.ev1); t.ev1 = nullptr; }
  }
  cudaSetDevice(cur);
  pinned_drop_all();
}

}  // extern "C"

namespace {

int run_single(jsb_graph* g, const jsb_params* points, int64_t n_points, int64_t reps_per_point, const jsb_run_opts* opts,
               const int device, const int64_t g_begin, const int64_t g_end, jsb_result** out) {
  using namespace jsb;
  *out = nullptr;
  const int64_t n_local = g_end - g_begin;
  int rc = JSB_OK;
  PhaseTimer pt;
  jsb_result* res = new (std::nothrow) jsb_result();
  if (!res) return fail(JSB_ENOMEM, "out of host memory");
  res->n_local = n_local;
  res->n_points = n_points;
  res->reps_per_point = reps_per_point;
  res->g_begin = g_begin;
  res->nv = (uint32_t)g->nv;
  res->flags = opts->flags;

  // device buffers
  DevPoint* d_points = nullptr;
  double* d_dpool = nullptr;
  int32_t* d_ipool = nullptr;
  uint8_t* d_ws = nullptr;
  jsb_summary* d_sum = nullptr;
  jsb_clone* d_clones = nullptr;
  unsigned long long* d_counters = nullptr;  // [0] work counter, [1] clone pool next, [2] entries into decoupled mode
  long long* d_clone_off = nullptr;
  jsb_event *d_log = nullptr, *d_gather = nullptr;
  uint32_t* d_chunk_next = nullptr;
  unsigned long long* d_ser_meta = nullptr;  // JSB_OUT_SERIES scratch
  jsb_series_row* d_series = nullptr;
  uint4* d_fin = nullptr;               // [n_local] fin_list (streamed outputs)
  uint32_t* d_chunk_prev = nullptr;
  unsigned long long* d_fin_ctr = nullptr;
  cudaStream_t stream2 = nullptr;
  void* h_fin = nullptr;                // pinned: fin_list copy + counter + chunk_next
  bool streamed = false;                // the event log reached the host while the kernel was running
  std::vector<uint4> fin_host;
  int32_t* d_chunk_owner = nullptr;
  uint32_t* d_chunk_seq = nullptr;
  unsigned long long* d_ev_off = nullptr;
  uint64_t ev_total = 0;
  uint32_t ev_used_chunks = 0;
  jsb_snapshot* d_snaps = nullptr;
  uint16_t* d_out_clone = nullptr;
  uint32_t *d_out_cell = nullptr, *d_out_lists = nullptr;
  uint32_t* d_order = nullptr;
  float* d_prio = nullptr;
  unsigned long long* d_trace = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t stream = nullptr;
  int prev_dev = 0;
  cudaGetDevice(&prev_dev);
  jsb_graph::DevCopy gd;
  DeviceState& ds = g_devs[device];

  if (n_local == 0) { *out = res; return JSB_OK; }

  std::unique_lock<std::mutex> dev_lock(ds.mu);  // one jsb_run at a time per device
  {
    CUDA_TRY(cudaSetDevice(device));
    rc = upload_graph(g, device, &gd);
    if (rc != JSB_OK) goto done;
    if (!ds.stream) CUDA_TRY(cudaStreamCreate(&ds.stream));
    stream = ds.stream;
    if (!ds.stack_ok) {  // the kernels need 8 KB of stack per thread; never lower what the host application has set
      size_t cur_stack = 0;
      CUDA_TRY(cudaDeviceGetLimit(&cur_stack, cudaLimitStackSize));
      if (cur_stack < 8192) CUDA_TRY(cudaDeviceSetLimit(cudaLimitStackSize, 8192));
      ds.stack_ok = true;
    }

    // ---- sweep points ---------------------------------------------------------------------
    std::vector<DevPoint> hp((size_t)n_points);
    std::vector<double> dpool(1, 0.0);
    std::vector<int32_t> ipool(1, 0);
    int max_snap = 0;
    for (int64_t i = 0; i < n_points; ++i) {
      const jsb_params& p = points[i];
      DevPoint& d = hp[i];
      std::memset(&d, 0, sizeof d);
      d.Tf = p.Tf; d.rate_birth = p.rate_birth; d.rate_death = p.rate_death; d.rate_migration = p.rate_migration;
      d.mu = p.mu_driver; d.adv_mean = p.adv_mean; d.adv_std = p.adv_std; d.root_rate = p.root_rate;
      d.max_iter = p.max_iterations;
      d.method = (int32_t)p.method; d.rule = p.rule; d.n_start = p.n_start_cells;
      d.n_bott = p.n_bottleneck; d.n_snap = p.n_snapshots;
      d.quirks = p.quirk_flags;
      d.bott_off = (int32_t)dpool.size();
      dpool.insert(dpool.end(), p.t_bottleneck, p.t_bottleneck + p.n_bottleneck);
      dpool.insert(dpool.end(), p.ratio_bottleneck, p.ratio_bottleneck + p.n_bottleneck);
      d.snap_off = (int32_t)dpool.size();
      dpool.insert(dpool.end(), p.t_snapshot, p.t_snapshot + p.n_snapshots);
      max_snap = std::max(max_snap, p.n_snapshots);
      if (p.method == JSB_METHOD_TREE) {
        d.n_nodes = p.n_tree_nodes; d.n_edges = p.n_tree_edges; d.root_node = p.root_node;
        d.rate_off = (int32_t)dpool.size();
        dpool.insert(dpool.end(), p.node_rate, p.node_rate + p.n_tree_nodes);
        d.edge_off = (int32_t)ipool.size();
        ipool.insert(ipool.end(), p.edge_parent, p.edge_parent + p.n_tree_edges);
        ipool.insert(ipool.end(), p.edge_child, p.edge_child + p.n_tree_edges);
      }
    }
    res->max_snap = max_snap;
    DEV_BUF(d_points, DevPoint*, BUF_POINTS, sizeof(DevPoint) * hp.size());
    DEV_BUF(d_dpool, double*, BUF_DPOOL, sizeof(double) * dpool.size());
    DEV_BUF(d_ipool, int32_t*, BUF_IPOOL, sizeof(int32_t) * ipool.size());
    CUDA_TRY(cudaMemcpyAsync(d_points, hp.data(), sizeof(DevPoint) * hp.size(), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(d_dpool, dpool.data(), sizeof(double) * dpool.size(), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(d_ipool, ipool.data(), sizeof(int32_t) * ipool.size(), cudaMemcpyHostToDevice, stream));

    // ---- launch geometry ------------------------------------------------------------------
    if (!ds.have_prop) {
      CUDA_TRY(cudaGetDeviceProperties(&ds.prop, device));
      ds.have_prop = true;
    }
    const cudaDeviceProp& prop = ds.prop;
    RunArgs a;
    std::memset(&a, 0, sizeof a);
    a.g.nv = (uint32_t)g->nv;
    a.site_bytes = g->nv <= 65535 ? 2u : 4u;  // list lengths share the site type: nv itself must fit
    const bool fast_mode = (opts->flags & JSB_MODE_REJECTION_FREE) != 0;
    int wpb = 1, wpb_max = 1;
    if ((fast_mode ? fast_prepare(a, opts->warps_per_sm, &wpb) : ssa_prepare(a, opts->warps_per_sm, &wpb)) != 0) { rc = fail(JSB_ECUDA, "cannot configure shared memory for the SSA kernel"); goto done; }
    // one block per SM; spread small ensembles over all SMs before stacking warps.  In the exact
    // mode warps without a replicate help the busy warps of their block (lookahead windows), so a
    // small ensemble still gets up to 1 + kMaxHelp warps per replicate.
    // How many warps per SM?  Two things favour few: shared memory and L1 share 256 KB and the carve-out comes
    // in steps (..., 132, 164, 196, 228 KB), and L1 caches the member lists (hit rate 57 % at 60 KB, 91 % at
    // 124 KB); and a block of at most eight warps may use 255 registers per thread instead of 168 (ssa_kernel's
    // second instantiation: no spill traffic).  Measured, exact kernel: C2 (4096 x 200x200, latency-bound) 7
    // warps/132 KB 1.74 s, 8 warps/164 KB 1.73 s, 6 warps 1.82 s, and with 168 registers 7 warps 1.92 s, 10
    // warps/196 KB 2.06 s, 12 warps/228 KB 2.20 s; a throughput-bound sweep (65 536 replicates of 100x100) 8
    // warps 5.37 s, 12 warps 5.63 s, 10 warps 5.80 s.  So: as many warps as fit the 132 KB step, at most eight.
    // The rejection-free kernel was only measured on latency-bound ensembles (C2: 12 warps 1.38 s, 8 warps
    // 1.24 s, 7 warps 1.25 s) and keeps every warp that fits for large sweeps.
    if (opts->warps_per_sm == 0 && (!fast_mode || (uint64_t)n_local <= 16ull * 10ull * (uint64_t)prop.multiProcessorCount)) {
      const int w132 = (int)((132u * 1024u - 1024u) / a.smem_per_warp);
      const int w164 = (int)((164u * 1024u - 1024u) / a.smem_per_warp);
      const int w196 = (int)((196u * 1024u - 1024u) / a.smem_per_warp);
      int want = w132 >= 6 ? w132 : (w164 >= 6 ? w164 : (w196 >= 6 ? w196 : wpb));
      if (!fast_mode && want > 8) want = 8;
      if (want < wpb) {
        if ((fast_mode ? fast_prepare(a, want, &wpb) : ssa_prepare(a, want, &wpb)) != 0) { rc = fail(JSB_ECUDA, "cannot configure shared memory for the SSA kernel"); goto done; }
      }
    }
    wpb_max = wpb;
    const bool helpers = !fast_mode && !std::getenv("JSB_NO_HELPERS");
    int grid;
    {
      const int64_t sms = prop.multiProcessorCount;
      const int64_t per_sm = (n_local + sms - 1) / sms;
      if (fast_mode || !helpers) {
        if (per_sm < wpb) wpb = (int)std::max<int64_t>(1, per_sm);
        grid = (int)std::min<int64_t>((n_local + wpb - 1) / wpb, sms);
      } else {
        if ((1 + kMaxHelp) * per_sm < wpb) wpb = (int)((1 + kMaxHelp) * per_sm);
        grid = (int)std::min<int64_t>(n_local, sms);
        a.static_map = (uint64_t)n_local < (uint64_t)grid * wpb ? 1u : 0u;
        a.clock_max_busy = 3;
        if (const char* e = std::getenv("JSB_CLOCK_MAX_BUSY")) a.clock_max_busy = (uint32_t)std::atoi(e);  // tuning hook
        a.helpers = std::getenv("JSB_NO_CLOCK") ? 3u : 1u;  // bit 1: lookahead only, no clock helper (debug)
      }
      if (!fast_mode) {
        a.dec_min_clones = kDecMinClones;
        if (const char* e = std::getenv("JSB_DEC_MIN_CLONES")) a.dec_min_clones = (uint32_t)std::max(1, std::atoi(e));  // tuning hook
        // Self-served deferred batches (no idle warp to serve them): measured slower than the exact fold in the
        // throughput phase of C2, so opt-in; batches are always handed to a server warp when one is attached.
      }
    }
    if (const char* e = std::getenv("JSB_DEBUG_GRID")) {  // debug: pack the replicates on fewer SMs (contention measurements)
      grid = std::max(1, std::atoi(e));
      wpb = (int)std::min<int64_t>(wpb_max, (n_local + grid - 1) / grid);
      a.static_map = 0;
    }
    const uint64_t slots = (uint64_t)grid * wpb;

    // sum of the block capacities after a compaction < 4*n_alive + kMinListCap*S, plus < kSeg entries of
    // alignment in front of every block of 2*kSeg entries and more (those hold > kSeg/2 elements each)
    uint32_t arena_cap = (uint32_t)std::min<uint64_t>(6ull * g->nv + 2ull * kMinListCap * kTab, 0xFFFFFF00ull);
    if (const char* dbg = std::getenv("JSB_DEBUG_ARENA_ENTRIES")) {  // test hook: force compactions
      uint64_t v = std::strtoull(dbg, nullptr, 10);
      arena_cap = (uint32_t)std::max<uint64_t>(v, 4096);  // too small -> JSB_ST_INTERNAL, never UB
    }
    const uint32_t site_bytes = a.site_bytes;
    const WsLayout L = ws_layout((uint32_t)g->nv, arena_cap, site_bytes);
    DEV_BUF(d_ws, uint8_t*, BUF_WS, L.total * slots);

    // ---- outputs --------------------------------------------------------------------------
    DEV_BUF(d_sum, jsb_summary*, BUF_SUM, sizeof(jsb_summary) * n_local);
    DEV_BUF(d_counters, unsigned long long*, BUF_COUNTERS, sizeof(unsigned long long) * 3);
    CUDA_TRY(cudaMemsetAsync(d_counters, 0, sizeof(unsigned long long) * 3, stream));
    unsigned long long clone_cap = std::min<unsigned long long>((unsigned long long)n_local * kMaxClones, 8ull << 20);
    DEV_BUF(d_clones, jsb_clone*, BUF_CLONES, sizeof(jsb_clone) * clone_cap);
    DEV_BUF(d_clone_off, long long*, BUF_CLONE_OFF, sizeof(long long) * n_local);
    uint32_t n_chunks = 0;
    const bool ret_series = (opts->flags & JSB_OUT_SERIES) != 0 && !fast_mode;
    const bool ret_events = (opts->flags & JSB_OUT_EVENTS) != 0;
    const bool log_needed = ret_events || ret_series;  // the series is a reduction of the log, done on the device
    if (log_needed) {
      uint64_t want = opts->log_pool_bytes ? opts->log_pool_bytes : (1ull << 30);
      if (ds.buf[BUF_LOG].bytes < want) {  // a pool of that size is not cached yet: what the device can spare
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        want = std::min<uint64_t>(want, std::max<uint64_t>((uint64_t)(free_b + ds.buf[BUF_LOG].bytes) / 3, ds.buf[BUF_LOG].bytes));
      }
      n_chunks = (uint32_t)std::max<uint64_t>(1, want / (sizeof(jsb_event) * kLogChunk));
      DEV_BUF(d_log, jsb_event*, BUF_LOG, sizeof(jsb_event) * kLogChunk * (size_t)n_chunks);
      DEV_BUF(d_chunk_next, uint32_t*, BUF_CHUNK_NEXT, sizeof(uint32_t));
      CUDA_TRY(cudaMemsetAsync(d_chunk_next, 0, sizeof(uint32_t), stream));
      DEV_BUF(d_chunk_owner, int32_t*, BUF_CHUNK_OWNER, sizeof(int32_t) * n_chunks);
      DEV_BUF(d_chunk_seq, uint32_t*, BUF_CHUNK_SEQ, sizeof(uint32_t) * n_chunks);
    }
    if (max_snap > 0) {
      DEV_BUF(d_snaps, jsb_snapshot*, BUF_SNAPS, sizeof(jsb_snapshot) * (size_t)max_snap * n_local);
      CUDA_TRY(cudaMemsetAsync(d_snaps, 0xFF, sizeof(jsb_snapshot) * (size_t)max_snap * n_local, stream));
    }
    if (opts->flags & JSB_OUT_LATTICE) {
      DEV_BUF(d_out_clone, uint16_t*, BUF_OUT_CLONE, sizeof(uint16_t) * (size_t)g->nv * n_local);
      DEV_BUF(d_out_cell, uint32_t*, BUF_OUT_CELL, sizeof(uint32_t) * (size_t)g->nv * n_local);
    }
    if (opts->flags & JSB_OUT_LISTS) DEV_BUF(d_out_lists, uint32_t*, BUF_OUT_LISTS, sizeof(uint32_t) * 2 * (size_t)g->nv * n_local);

    a.g.kind = g->kind; a.g.n1 = g->n1; a.g.n2 = g->n2; a.g.n3 = g->n3; a.g.nv = (uint32_t)g->nv;
    a.g.rowptr = gd.d_rowptr; a.g.colidx = gd.d_colidx;
    if (g->kind == 0) {
      // magic = ceil(2^(28+L) / d), L = ceil(log2 d): (v * magic) >> (28+L) == v / d for every v < 2^28
      auto magic = [](uint32_t d, uint32_t& m, uint32_t& sh) {
        uint32_t L = 0;
        while ((1ull << L) < d) ++L;
        sh = 28 + L;
        m = (uint32_t)(((1ull << sh) + d - 1) / d);
      };
      magic(g->n1, a.g.m1, a.g.s1);
      magic(g->n2, a.g.m2, a.g.s2);
    }
    a.points = d_points; a.dpool = d_dpool; a.ipool = d_ipool;
    a.seed_cand = gd.d_cand; a.n_cand = gd.d_ncand;
    a.flags = opts->flags | (log_needed ? JSB_OUT_EVENTS : 0u); a.seed = opts->seed; a.g_begin = g_begin; a.n_local = n_local; a.reps_per_point = reps_per_point;
    a.ws = d_ws; a.ws_stride = L.total; a.arena_cap = arena_cap; a.site_bytes = site_bytes; a.L = L;
    a.summaries = d_sum; a.clone_pool = d_clones; a.clone_pool_cap = clone_cap; a.clone_pool_next = d_counters + 1;
    a.clone_off = d_clone_off;
    a.log_pool = d_log; a.n_chunks = n_chunks; a.chunk_next = d_chunk_next; a.chunk_owner = d_chunk_owner; a.chunk_seq = d_chunk_seq;
    a.snaps = d_snaps; a.max_snap = max_snap;
    a.out_clone = d_out_clone; a.out_cell = d_out_cell; a.out_lists = d_out_lists;
    a.work_counter = d_counters;
    // Streamed outputs (exact kernel, event log requested; see the loop after the launch): needs the pinned host
    // block and the gather buffer of an earlier call, because the sizes are only known afterwards.
    const size_t b_lc0 = (opts->flags & JSB_OUT_LATTICE) ? ((sizeof(uint16_t) * (size_t)g->nv * n_local + 255) & ~size_t(255)) : 0;
    const size_t b_li0 = (opts->flags & JSB_OUT_LATTICE) ? ((sizeof(uint32_t) * (size_t)g->nv * n_local + 255) & ~size_t(255)) : 0;
    const size_t b_ls0 = (opts->flags & JSB_OUT_LISTS) ? ((sizeof(uint32_t) * 2 * (size_t)g->nv * n_local + 255) & ~size_t(255)) : 0;
    bool want_stream = ret_events && !fast_mode && !std::getenv("JSB_NO_STREAM_OUT") && ds.buf[BUF_GATHER].p;
    jsb_event* h_events = nullptr;
    if (want_stream) {
      res->pinned = pinned_take_cached(&res->pinned_bytes);
      const size_t fixed = b_lc0 + b_li0 + b_ls0;
      if (res->pinned && res->pinned_bytes > fixed + (size_t(64) << 20)) {
        h_events = (jsb_event*)((uint8_t*)res->pinned + fixed);
        d_gather = (jsb_event*)ds.buf[BUF_GATHER].p;
        a.gather_out = d_gather;
        a.gather_cap = std::min<uint64_t>((res->pinned_bytes - fixed) / sizeof(jsb_event), ds.buf[BUF_GATHER].bytes / sizeof(jsb_event));
        DEV_BUF(d_fin, uint4*, BUF_FIN, sizeof(uint4) * (size_t)n_local);
        DEV_BUF(d_fin_ctr, unsigned long long*, BUF_FIN_CTR, sizeof(unsigned long long));
        DEV_BUF(d_chunk_prev, uint32_t*, BUF_CHUNK_PREV, sizeof(uint32_t) * n_chunks);
        CUDA_TRY(cudaMemsetAsync(d_fin, 0xFF, sizeof(uint4) * (size_t)n_local, stream));
        CUDA_TRY(cudaMemsetAsync(d_fin_ctr, 0, sizeof(unsigned long long), stream));
        a.fin_list = d_fin; a.fin_ctr = d_fin_ctr; a.chunk_prev = d_chunk_prev;
      } else {
        want_stream = false;
      }
    }

    const char* trace_path = std::getenv("JSB_TRACE");  // debug: per-replicate start/end/SM/iterations, raw u64[4]
    if (trace_path) {
      CUDA_TRY(cudaMalloc(&d_trace, sizeof(unsigned long long) * 4 * n_local));
      a.trace = d_trace;
    }
    // every allocation and driver query of the pilot pass happens before the timed region (a cudaMalloc between the
    // two launches stalls on the driver lock when several processes share the box)
    const bool use_pilot = !fast_mode && (uint64_t)n_local > slots && (uint64_t)n_local <= 16 * slots && !std::getenv("JSB_NO_PILOT");
    uint8_t* d_susp = nullptr;
    uint32_t* d_susp_state = nullptr;
    const uint64_t susp_stride = (L.total + susp_extra_bytes(site_bytes) + 255) & ~uint64_t(255);
    if (use_pilot) {
      DEV_BUF(d_prio, float*, BUF_PRIO, sizeof(float) * n_local);
      DEV_BUF(d_order, uint32_t*, BUF_ORDER, sizeof(uint32_t) * n_local);
      // Suspending pilot (jsb_kernels.cu, susp_save / susp_restore): when the parked states of every replicate fit
      // comfortably into device memory the pilot pass is the first stretch of the real run (all outputs on) and the
      // full pass resumes it, so the pilot costs nothing but two workspace copies per replicate.  Otherwise the
      // pilot pass is thrown away as before.  JSB_NO_SUSPEND=1 forces the old behaviour.
      {
        const uint64_t need = susp_stride * (uint64_t)n_local + 4ull * n_local + 256;
        size_t have = ds.buf[BUF_SUSP].bytes;
        size_t free_b = 0, total_b = 0;
        if (need > have) cudaMemGetInfo(&free_b, &total_b);
        if (!std::getenv("JSB_NO_SUSPEND") && (need <= have || need <= (uint64_t)(free_b + have) / 2)) {
          uint8_t* blk = nullptr;
          DEV_BUF(blk, uint8_t*, BUF_SUSP, need);
          d_susp = blk;
          d_susp_state = reinterpret_cast<uint32_t*>(blk + susp_stride * (uint64_t)n_local);
          CUDA_TRY(cudaMemsetAsync(d_susp_state, 0, 4ull * n_local, stream));
        }
      }
    }
    pt.mark("setup + allocations");
    if (!ds.ev0) CUDA_TRY(cudaEventCreate(&ds.ev0));
    if (!ds.ev1) CUDA_TRY(cudaEventCreate(&ds.ev1));
    ev0 = ds.ev0; ev1 = ds.ev1;
    CUDA_TRY(cudaEventRecord(ev0, stream));
    // Longest-first scheduling for ensembles of a few replicates per warp slot: a short pilot pass (8192
    // iterations per replicate; its work estimate has rank correlation ~0.97 with the final iteration
    // counts) orders the full pass longest first, and the kernel deals the heaviest replicates out one
    // per SM.  The kernel time is bounded by the latency of the heaviest replicates, so they must start
    // first and must not share an SM (measured on the C2 ensemble: 2.4-2.7 s -> 2.1-2.2 s, pilot
    // included; without the spreading the order alone gained nothing; pilot lengths 2048 / 8192 / 32768 /
    // 131072 iterations: 1.77 / 1.73 / 1.76 / 1.93 s on the final build).  JSB_NO_PILOT=1 disables it.
    if (use_pilot) {
      RunArgs pa = a;
      // depth: half an iteration per site, 4096..32768 (C2, 200x200: 20 000; measured on the C2 ensemble,
      // three seeds: 8192 -> 1.76 s, 16384 -> 1.72 s, 32768 -> 1.71 s, 65536 -> 1.75 s, pilot included)
      pa.pilot_iters = (uint64_t)std::min<int64_t>(32768, std::max<int64_t>(4096, g->nv / 2));
      if (const char* e = std::getenv("JSB_PILOT_ITERS")) pa.pilot_iters = (uint32_t)std::max(256, std::atoi(e));  // tuning hook
      pa.prio = d_prio;
      pa.trace = nullptr;
      if (d_susp) {
        pa.susp = d_susp; pa.susp_stride = susp_stride; pa.susp_state = d_susp_state;
      } else {
        pa.flags = 0;
        pa.clone_pool = nullptr;
        pa.out_clone = nullptr; pa.out_cell = nullptr; pa.out_lists = nullptr; pa.snaps = nullptr;
      }
      int prc = launch_ssa(pa, grid, wpb, stream);
      if (prc != 0) { rc = fail(JSB_ECUDA, "pilot launch failed: %s", cudaGetErrorString((cudaError_t)prc)); goto done; }
      res->launches += 1;
      std::vector<float> prio((size_t)n_local);
      CUDA_TRY(cudaMemcpyAsync(prio.data(), d_prio, sizeof(float) * n_local, cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
      std::vector<uint32_t> order((size_t)n_local);
      for (int64_t i = 0; i < n_local; ++i) order[i] = (uint32_t)i;
      std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return prio[x] > prio[y]; });
      CUDA_TRY(cudaMemcpyAsync(d_order, order.data(), sizeof(uint32_t) * n_local, cudaMemcpyHostToDevice, stream));
      {  // the first grid*wpb ranks are dealt out statically by the kernel; the counter serves the rest
        const unsigned long long first = slots;
        CUDA_TRY(cudaMemcpyAsync(d_counters, &first, sizeof first, cudaMemcpyHostToDevice, stream));
      }
      CUDA_TRY(cudaStreamSynchronize(stream));  // `order` and `first` are host temporaries
      a.order = d_order;
      if (d_susp) { a.susp = d_susp; a.susp_stride = susp_stride; a.susp_state = d_susp_state; }
    }
    {
      int lrc = fast_mode ? launch_fast(a, grid, wpb, stream) : launch_ssa(a, grid, wpb, stream);
      if (lrc != 0) { rc = fail(JSB_ECUDA, "kernel launch failed: %s", cudaGetErrorString((cudaError_t)lrc)); goto done; }
    }
    res->launches += 1;
    CUDA_TRY(cudaEventRecord(ev1, stream));
    // ---- streamed outputs ---------------------------------------------------------------------------------------
    // The kernel time of an ensemble ends in a tail of a few heavy replicates while most SMs and both copy engines
    // idle, and the event logs are gigabytes.  A replicate that finishes packs its own log into the gather buffer
    // at rows it draws with one atomic, then publishes itself in fin_list; this thread polls that list on a second
    // stream and starts the device-to-host copy of every completed run of ranks (contiguous rows), so that when
    // the kernel ends only the last few logs are left to copy.  When the buffers of the earlier call turn out too
    // small for this one, everything is gathered and copied after the kernel as before.
    if (want_stream) {
      if (!ds.stream2) CUDA_TRY(cudaStreamCreateWithFlags(&ds.stream2, cudaStreamNonBlocking));
      stream2 = ds.stream2;
      {
        const size_t hb = sizeof(uint4) * (size_t)n_local + 64;
        if (ds.h_scratch_bytes < hb) {
          if (ds.h_scratch) { cudaFreeHost(ds.h_scratch); ds.h_scratch = nullptr; ds.h_scratch_bytes = 0; }
          CUDA_TRY(cudaHostAlloc(&ds.h_scratch, hb + hb / 4, cudaHostAllocDefault));
          ds.h_scratch_bytes = hb + hb / 4;
        }
        h_fin = ds.h_scratch;
      }
      uint4* h_list = (uint4*)h_fin;
      unsigned long long* h_ctr = (unsigned long long*)(h_list + n_local);
      uint64_t processed = 0;
      bool ok = true;
      const uint64_t min_batch = (uint64_t)std::max<int64_t>(8, n_local / 128);
      for (;;) {
        const cudaError_t q = cudaEventQuery(ev1);
        if (q != cudaSuccess && q != cudaErrorNotReady) { rc = fail(JSB_ECUDA, "simulation kernel failed: %s", cudaGetErrorString(q)); goto done; }
        const bool finished = q == cudaSuccess;
        CUDA_TRY(cudaMemcpyAsync(h_ctr, d_fin_ctr, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream2));
        CUDA_TRY(cudaStreamSynchronize(stream2));
        const uint64_t cnt = std::min<uint64_t>(*h_ctr >> 40, (uint64_t)n_local);
        if (cnt > processed && (finished || cnt - processed >= min_batch)) {
          CUDA_TRY(cudaMemcpyAsync(h_list + processed, d_fin + processed, sizeof(uint4) * (cnt - processed), cudaMemcpyDeviceToHost, stream2));
          CUDA_TRY(cudaStreamSynchronize(stream2));
          uint64_t k = processed;
          while (k < cnt && h_list[k].x != 0xFFFFFFFFu) {  // an entry follows its rank once the log is packed
            if (h_list[k].x & 0x80000000u) ok = false;      // no room in the buffers of the earlier call
            ++k;
          }
          if (!ok) break;
          if (k > processed) {
            const uint64_t e0 = ((uint64_t)h_list[processed].w << 32) | h_list[processed].z;
            const uint64_t e1 = (((uint64_t)h_list[k - 1].w << 32) | h_list[k - 1].z) + h_list[k - 1].y;
            if (e1 > e0) CUDA_TRY(cudaMemcpyAsync(h_events + e0, d_gather + e0, sizeof(jsb_event) * (e1 - e0), cudaMemcpyDeviceToHost, stream2));
            {  // lattices and lists of these ranks: slots [processed, k) are contiguous
              uint8_t* hb = (uint8_t*)res->pinned;
              const size_t nvz = (size_t)g->nv, r0 = (size_t)processed, nr = (size_t)(k - processed);
              if (opts->flags & JSB_OUT_LATTICE) {
                CUDA_TRY(cudaMemcpyAsync(hb + sizeof(uint16_t) * nvz * r0, d_out_clone + nvz * r0, sizeof(uint16_t) * nvz * nr, cudaMemcpyDeviceToHost, stream2));
                CUDA_TRY(cudaMemcpyAsync(hb + b_lc0 + sizeof(uint32_t) * nvz * r0, d_out_cell + nvz * r0, sizeof(uint32_t) * nvz * nr, cudaMemcpyDeviceToHost, stream2));
              }
              if (opts->flags & JSB_OUT_LISTS)
                CUDA_TRY(cudaMemcpyAsync(hb + b_lc0 + b_li0 + sizeof(uint32_t) * 2 * nvz * r0, d_out_lists + 2 * nvz * r0, sizeof(uint32_t) * 2 * nvz * nr, cudaMemcpyDeviceToHost, stream2));
            }
            processed = k;
          }
        }
        if (finished && processed == cnt) {
          if (cnt != (uint64_t)n_local) ok = false;  // every replicate of the call publishes itself
          break;
        }
        if (!finished) std::this_thread::sleep_for(std::chrono::microseconds(300));
      }
      CUDA_TRY(cudaStreamSynchronize(stream2));
      if (ok) {
        streamed = true;
        fin_host.assign(h_list, h_list + n_local);
        ev_total = *h_ctr & ((1ull << 40) - 1ull);
      }
    }
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (want_stream) {
      // the kernel wrote lattices and lists by finish rank: the slot of every replicate, streamed or not
      if (!streamed) {
        fin_host.resize((size_t)n_local);
        CUDA_TRY(cudaMemcpy(fin_host.data(), d_fin, sizeof(uint4) * (size_t)n_local, cudaMemcpyDeviceToHost));
      }
      res->lat_slot.assign((size_t)n_local, 0);
      for (int64_t k = 0; k < n_local; ++k) res->lat_slot[fin_host[(size_t)k].x & 0x7FFFFFFFu] = (uint32_t)k;
    }
    {
      float ms = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&ms, ev0, ev1));
      res->kernel_ms = ms;
    }

    if (trace_path) {
      std::vector<unsigned long long> tr((size_t)n_local * 4);
      CUDA_TRY(cudaMemcpy(tr.data(), d_trace, sizeof(unsigned long long) * tr.size(), cudaMemcpyDeviceToHost));
      if (FILE* f = std::fopen(trace_path, "wb")) { std::fwrite(tr.data(), sizeof(unsigned long long), tr.size(), f); std::fclose(f); }
    }
    pt.mark("simulation kernel");
    // ---- copy back ------------------------------------------------------------------------
    res->summaries.resize((size_t)n_local);
    CUDA_TRY(cudaMemcpy(res->summaries.data(), d_sum, sizeof(jsb_summary) * n_local, cudaMemcpyDeviceToHost));
    {
      unsigned long long counters[3];
      CUDA_TRY(cudaMemcpy(counters, d_counters, sizeof counters, cudaMemcpyDeviceToHost));
      if (std::getenv("JSB_TIMING")) std::fprintf(stderr, "[jsb] entries into decoupled mode: %llu\n", counters[2]);
      unsigned long long used = std::min(counters[1], clone_cap);
      res->clones.resize((size_t)used);
      res->clone_off.resize((size_t)n_local);
      CUDA_TRY(cudaMemcpy(res->clones.data(), d_clones, sizeof(jsb_clone) * used, cudaMemcpyDeviceToHost));
      CUDA_TRY(cudaMemcpy(res->clone_off.data(), d_clone_off, sizeof(long long) * n_local, cudaMemcpyDeviceToHost));
    }
    if (streamed) {
      res->ev_off.assign((size_t)n_local + 1, 0);
      res->ev_cnt.assign((size_t)n_local, 0);
      for (int64_t k = 0; k < n_local; ++k) {
        const uint4 e = fin_host[(size_t)k];
        res->ev_off[e.x] = ((uint64_t)e.w << 32) | e.z;
        res->ev_cnt[e.x] = e.y;
      }
    } else if (log_needed) {
      uint32_t used_chunks = 0;
      CUDA_TRY(cudaMemcpy(&used_chunks, d_chunk_next, sizeof(uint32_t), cudaMemcpyDeviceToHost));
      used_chunks = std::min(used_chunks, n_chunks);
      std::vector<int32_t> owner(used_chunks);
      CUDA_TRY(cudaMemcpy(owner.data(), d_chunk_owner, sizeof(int32_t) * used_chunks, cudaMemcpyDeviceToHost));
      std::vector<uint32_t> chunks_of((size_t)n_local, 0);
      for (uint32_t c = 0; c < used_chunks; ++c) chunks_of[owner[c]]++;
      res->ev_off.assign((size_t)n_local + 1, 0);
      for (int64_t i = 0; i < n_local; ++i) {
        uint64_t logged = std::min<uint64_t>(res->summaries[i].n_events, (uint64_t)chunks_of[i] * kLogChunk);
        res->ev_off[i + 1] = res->ev_off[i] + logged;
      }
      res->ev_cnt.resize((size_t)n_local);
      for (int64_t i = 0; i < n_local; ++i) res->ev_cnt[i] = res->ev_off[i + 1] - res->ev_off[i];
      ev_total = res->ev_off[n_local];
      ev_used_chunks = used_chunks;
    }
    {
      // one pinned block for every bulk output: [lattice clone | lattice cell | lists | events]
      auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
      const size_t b_ev = ret_events ? up(sizeof(jsb_event) * ev_total) : 0;
      const size_t b_lc = b_lc0, b_li = b_li0, b_ls = b_ls0;
      const size_t need = b_ev + b_lc + b_li + b_ls;
      if (!streamed) {
        if (res->pinned) { pinned_give(res->pinned, res->pinned_bytes); res->pinned = nullptr; res->pinned_bytes = 0; }
        if (need > 0) {
          res->pinned = pinned_take(need, &res->pinned_bytes);
          if (!res->pinned) { rc = fail(JSB_ENOMEM, "cannot pin %llu bytes of host memory for the outputs", (unsigned long long)need); goto done; }
        }
      }
      if (res->pinned) {
        uint8_t* base = (uint8_t*)res->pinned;
        res->lat_clone = (uint16_t*)base;
        res->lat_cell = (uint32_t*)(base + b_lc);
        res->lists = (uint32_t*)(base + b_lc + b_li);
        res->events = ret_events ? (jsb_event*)(base + b_lc + b_li + b_ls) : nullptr;
      }
      if (ev_total > 0 && !streamed) {
        DEV_BUF(d_gather, jsb_event*, BUF_GATHER, sizeof(jsb_event) * ev_total);
        CUDA_TRY(cudaMalloc(&d_ev_off, sizeof(unsigned long long) * (n_local + 1)));
        CUDA_TRY(cudaMemcpyAsync(d_ev_off, res->ev_off.data(), sizeof(unsigned long long) * (n_local + 1), cudaMemcpyHostToDevice, stream));
        int gblocks = (int)std::min<uint32_t>(ev_used_chunks, (uint32_t)prop.multiProcessorCount * 8);
        gather_log_kernel<<<gblocks, 256, 0, stream>>>(d_log, d_chunk_owner, d_chunk_seq, d_ev_off, ev_used_chunks, d_gather);
        CUDA_TRY(cudaGetLastError());
        res->launches += 1;
        if (ret_events) CUDA_TRY(cudaMemcpyAsync(res->events, d_gather, sizeof(jsb_event) * ev_total, cudaMemcpyDeviceToHost, stream));
      }
      if (ret_series) {
        // the packed log (d_gather: streamed or gathered just now) reduced to series rows on the device
        const size_t nl = (size_t)n_local;
        std::vector<uint32_t> n0(nl);
        for (size_t i = 0; i < nl; ++i) {
          const int32_t ns = points[(g_begin + (int64_t)i) / reps_per_point].n_start_cells;
          n0[i] = (uint32_t)(ns < 0 ? 0 : ns);
        }
        CUDA_TRY(cudaMalloc(&d_ser_meta, sizeof(unsigned long long) * 4 * nl + sizeof(uint32_t) * nl + sizeof(long long) * nl * (size_t)std::max(1, max_snap)));
        unsigned long long* d_off = d_ser_meta;
        unsigned long long* d_cnt = d_off + nl;
        unsigned long long* d_scnt = d_cnt + nl;
        unsigned long long* d_soff = d_scnt + nl;
        long long* d_snap_rows = reinterpret_cast<long long*>(d_soff + nl);
        uint32_t* d_n0 = reinterpret_cast<uint32_t*>(d_snap_rows + nl * (size_t)std::max(1, max_snap));
        CUDA_TRY(cudaMemcpyAsync(d_off, res->ev_off.data(), sizeof(unsigned long long) * nl, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(d_cnt, res->ev_cnt.data(), sizeof(unsigned long long) * nl, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(d_n0, n0.data(), sizeof(uint32_t) * nl, cudaMemcpyHostToDevice, stream));
        const int sblocks = (int)((nl + 63) / 64);
        const jsb_event* d_packed = ev_total > 0 ? d_gather : nullptr;
        series_kernel<false><<<sblocks, 64, 0, stream>>>(d_packed, d_off, d_cnt, d_n0, nullptr, 0, n_local, d_scnt, nullptr, nullptr, nullptr);
        CUDA_TRY(cudaGetLastError());
        std::vector<unsigned long long> scnt(nl);
        CUDA_TRY(cudaMemcpyAsync(scnt.data(), d_scnt, sizeof(unsigned long long) * nl, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        res->ser_off.assign(nl + 1, 0);
        for (size_t i = 0; i < nl; ++i) res->ser_off[i + 1] = res->ser_off[i] + scnt[i];
        const uint64_t ser_total = res->ser_off[nl];
        res->series.resize((size_t)ser_total);
        res->ser_snap.assign(nl * (size_t)std::max(1, max_snap), 0);
        CUDA_TRY(cudaMalloc(&d_series, sizeof(jsb_series_row) * std::max<uint64_t>(ser_total, 1)));
        CUDA_TRY(cudaMemcpyAsync(d_soff, res->ser_off.data(), sizeof(unsigned long long) * nl, cudaMemcpyHostToDevice, stream));
        series_kernel<true><<<sblocks, 64, 0, stream>>>(d_packed, d_off, d_cnt, d_n0, max_snap > 0 ? d_snaps : nullptr, max_snap, n_local, nullptr,
                                                        d_soff, d_series, d_snap_rows);
        CUDA_TRY(cudaGetLastError());
        res->launches += 2;
        if (ser_total) CUDA_TRY(cudaMemcpyAsync(res->series.data(), d_series, sizeof(jsb_series_row) * ser_total, cudaMemcpyDeviceToHost, stream));
        if (max_snap > 0) CUDA_TRY(cudaMemcpyAsync(res->ser_snap.data(), d_snap_rows, sizeof(long long) * nl * (size_t)max_snap, cudaMemcpyDeviceToHost, stream));
      }
      if ((opts->flags & JSB_OUT_LATTICE) && !streamed) {
        CUDA_TRY(cudaMemcpyAsync(res->lat_clone, d_out_clone, sizeof(uint16_t) * (size_t)g->nv * n_local, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaMemcpyAsync(res->lat_cell, d_out_cell, sizeof(uint32_t) * (size_t)g->nv * n_local, cudaMemcpyDeviceToHost, stream));
      }
      if ((opts->flags & JSB_OUT_LISTS) && !streamed)
        CUDA_TRY(cudaMemcpyAsync(res->lists, d_out_lists, sizeof(uint32_t) * 2 * (size_t)g->nv * n_local, cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
    }
    pt.mark("summaries + clones + log");
    if (max_snap > 0) {
      res->snaps.resize((size_t)max_snap * n_local);
      CUDA_TRY(cudaMemcpy(res->snaps.data(), d_snaps, sizeof(jsb_snapshot) * res->snaps.size(), cudaMemcpyDeviceToHost));
      res->snap_count.assign((size_t)n_local, 0);
      for (int64_t i = 0; i < n_local; ++i) {
        int c = 0;
        while (c < max_snap && res->snaps[(size_t)i * max_snap + c].index != 0xFFFFFFFFu) ++c;
        res->snap_count[i] = c;
      }
    }
    pt.mark("lattice + lists");
  }

done:
  // (every buffer, stream and event a normal call uses belongs to the device cache; what is freed here are the
  // tables of the rarer paths: gather after the kernel, series, trace)
  if (stream) cudaStreamSynchronize(stream);   // nothing of this call is in flight when the lock is released
  if (stream2) cudaStreamSynchronize(stream2);
  if (d_ser_meta) cudaFree(d_ser_meta);
  if (d_series) cudaFree(d_series);
  if (d_ev_off) cudaFree(d_ev_off);
  if (d_trace) cudaFree(d_trace);
  cudaSetDevice(prev_dev);
  pt.mark("free");
  if (rc != JSB_OK) { delete res; return rc; }
  *out = res;
  return JSB_OK;
}

}  // namespace

extern "C" {

int jsb_result_kernel_ms(const jsb_result* r, double* elapsed_ms) {
  if (!r || !elapsed_ms) return fail(JSB_EINVAL, "null argument");
  *elapsed_ms = r->kernel_ms;
  return JSB_OK;
}

int jsb_result_launches(const jsb_result* r, int64_t* n) {
  if (!r || !n) return fail(JSB_EINVAL, "null argument");
  *n = r->launches;
  return JSB_OK;
}

int64_t jsb_result_count(const jsb_result* r) { return r ? r->n_local : 0; }

int jsb_result_summaries(const jsb_result* r, const jsb_summary** p) {
  if (!r || !p) return fail(JSB_EINVAL, "null argument");
  *p = r->summaries.data();
  return JSB_OK;
}

#define CHECK_INDEX(r, i)                                                                          \
  if (!(r)) return fail(JSB_EINVAL, "result is null");                                             \
  if ((i) < 0 || (i) >= (r)->n_local) return fail(JSB_EINVAL, "replicate index %lld is outside 0..%lld", (long long)(i), (long long)(r)->n_local - 1)

int jsb_result_clones(const jsb_result* r, int64_t i, const jsb_clone** p, int32_t* n) {
  CHECK_INDEX(r, i);
  if (!p || !n) return fail(JSB_EINVAL, "null argument");
  r = part_of(r, i);
  long long off = r->clone_off[i];
  if (off < 0) return fail(JSB_ECAPACITY, "clone pool overflowed; replicate %lld has no clone table", (long long)i);
  *p = r->clones.data() + off;
  *n = (int32_t)r->summaries[i].n_clones;
  return JSB_OK;
}

int jsb_result_events(jsb_result* r, int64_t i, const jsb_event** p, int64_t* n) {
  CHECK_INDEX(r, i);
  if (!p || !n) return fail(JSB_EINVAL, "null argument");
  if (!(r->flags & JSB_OUT_EVENTS)) return fail(JSB_EINVAL, "run did not request JSB_OUT_EVENTS");
  r = part_of(r, i);
  *p = r->events ? r->events + r->ev_off[i] : nullptr;
  *n = (int64_t)r->ev_cnt[i];
  return JSB_OK;
}

int jsb_result_lattice(const jsb_result* r, int64_t i, const uint16_t** clone, const uint32_t** cell_id) {
  CHECK_INDEX(r, i);
  if (!(r->flags & JSB_OUT_LATTICE)) return fail(JSB_EINVAL, "run did not request JSB_OUT_LATTICE");
  r = part_of(r, i);
  const size_t slot = r->lat_slot.empty() ? (size_t)i : (size_t)r->lat_slot[(size_t)i];
  if (clone) *clone = r->lat_clone + slot * r->nv;
  if (cell_id) *cell_id = r->lat_cell + slot * r->nv;
  return JSB_OK;
}

int jsb_result_list(jsb_result* r, int64_t i, int32_t list, const uint32_t** p, int64_t* n) {
  CHECK_INDEX(r, i);
  if (!p || !n) return fail(JSB_EINVAL, "null argument");
  if (!(r->flags & JSB_OUT_LISTS)) return fail(JSB_EINVAL, "run did not request JSB_OUT_LISTS");
  r = part_of(r, i);
  const jsb_summary& s = r->summaries[i];
  if (list < 0 || list > (int32_t)s.n_clones) return fail(JSB_EINVAL, "list %d is outside 0..n_clones", list);
  const uint32_t* base = r->lists + (r->lat_slot.empty() ? (size_t)i : (size_t)r->lat_slot[(size_t)i]) * 2 * r->nv;
  if (list == 0) { *p = base; *n = s.n_alive; return JSB_OK; }
  const jsb_clone* c;
  int32_t nc;
  int rc = jsb_result_clones(r, i, &c, &nc);
  if (rc != JSB_OK) return rc;
  uint64_t off = 0;
  for (int32_t k = 0; k < list - 1; ++k) off += c[k].count;
  *p = base + r->nv + off;
  *n = c[list - 1].count;
  return JSB_OK;
}

int jsb_result_snapshots(const jsb_result* r, int64_t i, const jsb_snapshot** p, int32_t* n) {
  CHECK_INDEX(r, i);
  if (!p || !n) return fail(JSB_EINVAL, "null argument");
  r = part_of(r, i);
  if (r->max_snap == 0) { *p = nullptr; *n = 0; return JSB_OK; }
  *p = r->snaps.data() + (size_t)i * r->max_snap;
  *n = r->snap_count[i];
  return JSB_OK;
}

int jsb_result_series(const jsb_result* r, int64_t i, const jsb_series_row** p, int64_t* n,
                      const int64_t** rows_before_snapshot, int32_t* n_snapshots) {
  CHECK_INDEX(r, i);
  if (!p || !n) return fail(JSB_EINVAL, "null argument");
  if (!(r->flags & JSB_OUT_SERIES)) return fail(JSB_EINVAL, "run did not request JSB_OUT_SERIES");
  r = part_of(r, i);
  if (r->ser_off.empty()) return fail(JSB_EINVAL, "the rejection-free mode keeps no series");
  *p = r->series.data() + r->ser_off[i];
  *n = (int64_t)(r->ser_off[i + 1] - r->ser_off[i]);
  if (rows_before_snapshot && n_snapshots) {
    *rows_before_snapshot = r->max_snap ? r->ser_snap.data() + (size_t)i * r->max_snap : nullptr;
    *n_snapshots = r->max_snap ? r->snap_count[i] : 0;
  }
  return JSB_OK;
}

int jsb_result_histogram(const jsb_result* r, int64_t point, int kind, uint64_t out[64]) {
  if (!r || !out) return fail(JSB_EINVAL, "null argument");
  if (point < 0 || point >= r->n_points) return fail(JSB_EINVAL, "point %lld is outside 0..%lld", (long long)point, (long long)r->n_points - 1);
  if (kind < 0 || kind > 2) return fail(JSB_EINVAL, "kind must be 0..2");
  std::memset(out, 0, sizeof(uint64_t) * 64);
  const int64_t lo = std::max<int64_t>(point * r->reps_per_point, r->g_begin);
  const int64_t hi = std::min<int64_t>((point + 1) * r->reps_per_point, r->g_begin + r->n_local);
  const uint64_t width = std::max<uint64_t>(1, ((uint64_t)r->nv + 63) / 64);
  for (int64_t gi = lo; gi < hi; ++gi) {
    const jsb_summary& s = r->summaries[gi - r->g_begin];
    uint64_t bin;
    if (kind == 0) bin = s.n_alive / width;
    else if (kind == 1) bin = s.n_clones;
    else { bin = 0; uint64_t v = s.n_iterations; while (v > 1) { v >>= 1; ++bin; } }
    out[std::min<uint64_t>(bin, 63)]++;
  }
  return JSB_OK;
}

void jsb_result_free(jsb_result* r) { delete r; }

// create_matrix_relational, src/sampling_phylogentic_relation_and_genotype.jl:60-104: walk the
// Duplicate/Mutation rows backwards; a row whose Cell is currently in the sample list emits
// (Father, Child, Time, Subpop_Child, Event) and replaces the child by its father.
int jsb_result_genealogy(jsb_result* r, int64_t i, const uint32_t* sample_cells, int64_t n_samples,
                         jsb_relation** rows, int64_t* n_rows) {
  CHECK_INDEX(r, i);
  if (!rows || !n_rows || (n_samples > 0 && !sample_cells)) return fail(JSB_EINVAL, "null argument");
  if (!(r->flags & JSB_OUT_EVENTS)) return fail(JSB_EINVAL, "run did not request JSB_OUT_EVENTS");
  r = part_of(r, i);
  const jsb_event* ev = r->events ? r->events + r->ev_off[i] : nullptr;
  const int64_t n = (int64_t)r->ev_cnt[i];
  // every cell id appears as `cell` in at most one Duplicate/Mutation row (ids are unique), so
  // the backward scan is a pointer chase: mark wanted ids, visit rows from the end.
  std::vector<uint8_t> wanted((size_t)r->summaries[i].next_cell_id + 1, 0);
  for (int64_t k = 0; k < n_samples; ++k) {
    if (sample_cells[k] >= wanted.size()) return fail(JSB_EINVAL, "sample cell id %u was never created", sample_cells[k]);
    wanted[sample_cells[k]] = 1;
  }
  std::vector<jsb_relation> outv;
  for (int64_t k = n - 1; k >= 0; --k) {
    const jsb_event& e = ev[k];
    if (e.type != JSB_EV_DUPLICATE && e.type != JSB_EV_MUTATION) continue;  // :121-122
    if (!wanted[e.cell]) continue;
    wanted[e.cell] = 0;
    wanted[e.parent] = 1;
    jsb_relation rel;
    std::memset(&rel, 0, sizeof rel);
    rel.time = e.time;
    rel.father = e.parent;
    rel.child = e.cell;
    // :86-96 — the reference looks the subpopulation up from mutations[end], which for a
    // Mutation row is the FATHER's genotype (the new driver stripped), else the cell's own.
    rel.subpop_child = e.clone;
    if (e.type == JSB_EV_MUTATION && r->clone_off[i] >= 0) rel.subpop_child = r->clones[(size_t)r->clone_off[i] + e.clone - 1].parent;
    rel.type = e.type;
    outv.push_back(rel);
  }
  jsb_relation* buf = (jsb_relation*)std::malloc(sizeof(jsb_relation) * std::max<size_t>(1, outv.size()));
  if (!buf) return fail(JSB_ENOMEM, "out of host memory");
  std::memcpy(buf, outv.data(), sizeof(jsb_relation) * outv.size());
  *rows = buf;
  *n_rows = (int64_t)outv.size();
  return JSB_OK;
}

void jsb_free(void* p) { std::free(p); }

}  // extern "C"
