// jspace_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain, sequential C++ restatement of the spatial clonal-dynamics hot path of
// BIMIB-DISCo/J-Space.jl, written to be read side by side with the reference:
//   graph build + seeding   src/J_Space.jl:43-163
//   initial lists           src/J_Space.jl:296-314, 542-550
//   event handlers          src/J_Space.jl:320-416 (cell_birth, random drivers),
//                           :419-537 (cell_birth, driver tree), :556-568, :574-586, :592-613
//   main loops              src/J_Space.jl:638-842 (method 1), :848-1104 (method 2)
// Each function below cites the lines it follows.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path
//  and cannot be executed here (no Julia in the image), so this restatement is
// pinned only against (a) the Random123 known-answer vectors for Philox4x32-10, (b) the
// statistical invariants of the reference's shipped fileoutput/DriverList.csv, and (c) its own
// invariants.  Random numbers do NOT come from Julia's MersenneTwister: both this oracle and
// the CUDA kernels implement the counter-based "draw contract" of DESIGN.md §3 (Philox keyed
// by (seed, replicate, iteration, draw site)), which a Julia AbstractRNG shim can serve to the
// unmodified reference loop.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this file's library.  The product (libjspace_b200.so) never links or calls it.
//
// Third-party arithmetic restated here because it lives outside /root/reference
// (Manifest.toml pins): Julia 1.7.3 Base `sum` (sequential left fold for n < 1024) and
// `cumsum` (Base.accumulate_pairwise!, blocks of 128); Distributions 0.25.73
// `rand(Exponential(θ)) = θ·randexp`, `rand(Normal(μ,σ)) = μ + σ·randn`; StatsBase 0.33.21
// `sample(rng, a, k; replace=false)` branch structure; Graphs 1.7.3 `grid`,
// `cartesian_product`, `closeness_centrality`.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off; no fast-math — fp64 results must be
// the IEEE results of the written expression order).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

// -------------------------------------------------------------------------------------------
// Records (own definitions; layouts equal include/jspace_b200.h so tests can memcmp)
// -------------------------------------------------------------------------------------------
struct OrcEvent {
  double time;
  uint32_t cell, parent, site, site2;
  uint16_t clone, label;
  uint8_t type, flags;
  uint16_t pad;
};
static_assert(sizeof(OrcEvent) == 32, "event record is 32 bytes");
enum { EV_DUPLICATE = 0, EV_MUTATION = 1, EV_DEATH = 2, EV_MIGRATION = 3 };
enum { EVF_GROUP_START = 1, EVF_DISPLACED = 2, EVF_EXCISION = 4 };

struct OrcClone {
  double alpha;
  uint32_t count;
  uint16_t parent, label;
};
static_assert(sizeof(OrcClone) == 16, "clone record is 16 bytes");

struct OrcSummary {
  uint64_t n_iterations, n_events;
  double t_final;
  uint32_t n_effective, n_births, n_deaths, n_migrations, n_displaced, n_excised, n_alive,
      n_clones, next_cell_id;
  int32_t status;
};
static_assert(sizeof(OrcSummary) == 64, "summary record is 64 bytes");
enum { ST_OK = 0, ST_NONTERMINATING = 1, ST_CLONE_CAPACITY = 2, ST_LOG_OVERFLOW = 3, ST_MAX_ITER = 4 };

struct OrcSnapshot {
  double time;
  uint64_t iteration, n_events_before;
  uint32_t n_alive, index;
};
static_assert(sizeof(OrcSnapshot) == 32, "snapshot record is 32 bytes");

struct OrcParams {  // same field order as jsb_params
  uint32_t size, method;
  double Tf, rate_birth, rate_death, rate_migration, mu_driver, adv_mean, adv_std;
  int32_t rule, n_start_cells, n_bottleneck, n_snapshots;
  const double *t_bottleneck, *ratio_bottleneck, *t_snapshot;
  int32_t n_tree_nodes, n_tree_edges;
  const int32_t *edge_parent, *edge_child;
  const double* node_rate;
  int32_t root_node;
  uint32_t quirk_flags;
  double root_rate;
  uint64_t max_iterations;
};
enum { RULE_CONTACT = 0, RULE_VOTER = 1, RULE_HVOTER = 2, RULE_NONE = 3 };
enum { QUIRK_EXCISION_INTENT = 1 };

#include "julia_mt.hpp"

// -------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  Checked against the Random123
// known-answer vectors in tests/test_oracle_rng.py.
// -------------------------------------------------------------------------------------------
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; ++round) {
    if (round > 0) {
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// -------------------------------------------------------------------------------------------
// Draw contract.  One Philox block = two 64-bit words (a, b):
//   key = (seed_lo, seed_hi)
//   ctr = (iter_lo, iter_hi[15:0] | attempt<<16, stream, site<<28 | seq)
// Loop iterations are numbered from 1; setup draws use iteration 0.
// -------------------------------------------------------------------------------------------
enum DrawSite {
  DS_TIME_CLASS = 0,  // a: D1 randexp (src/J_Space.jl:699), b: D2 class uniform (:721)
  DS_CELL_NBR = 1,    // a: D3 member index (:727/:746/:762), b: D4 neighbour index (:730/:764)
  DS_DRIVER_COIN = 4, // a: D7 driver uniform (:347/:447), b: coin (:391/:512)
  DS_LABEL = 5,       // a: D8 label rank (:371)
  DS_NORMAL = 6,      // a,b: D9 polar-method pair, seq = attempt (:566)
  DS_EXCISION = 7,    // a: index draws of StatsBase.sample (:820-823), seq = draw number
  DS_SEED_TIE = 9,    // a: closeness tie-break (:86), iteration 0
  DS_SEED_POS = 10    // a: n_cell > 1 positions (:109/:154), seq = cell number, iteration 0
};

struct Rng {
  uint64_t seed;
  uint32_t stream;
  void block(uint64_t iter, uint32_t attempt, uint32_t site, uint32_t seq, uint64_t& a,
             uint64_t& b) const {
    uint32_t ctr[4] = {(uint32_t)iter, (uint32_t)((iter >> 32) & 0xFFFFu) | (attempt << 16), stream,
                       (site << 28) | (seq & 0x0FFFFFFFu)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    philox4x32_10(ctr, key, o);
    a = (uint64_t)o[0] | ((uint64_t)o[1] << 32);
    b = (uint64_t)o[2] | ((uint64_t)o[3] << 32);
  }
  uint64_t word(uint64_t iter, uint32_t attempt, uint32_t site, uint32_t seq, int half) const {
    uint64_t a, b;
    block(iter, attempt, site, seq, a, b);
    return half ? b : a;
  }
  // Uniform index in [0, s): Lemire's nearly-divisionless method on 64-bit words; a rejected
  // word is replaced by the same (site, seq, half) word of the next attempt.
  uint64_t index(uint64_t iter, uint32_t site, uint32_t seq, int half, uint64_t s) const {
    uint32_t attempt = 0;
    uint64_t w = word(iter, attempt, site, seq, half);
    unsigned __int128 m = (unsigned __int128)w * s;
    uint64_t l = (uint64_t)m;
    if (l < s) {
      uint64_t t = (0 - s) % s;
      while (l < t) {
        ++attempt;
        w = word(iter, attempt, site, seq, half);
        m = (unsigned __int128)w * s;
        l = (uint64_t)m;
      }
    }
    return (uint64_t)(m >> 64);
  }
};

static inline double u01_53(uint64_t w) { return (double)(w >> 11) * 0x1.0p-53; }            // [0,1)
static inline double u01_open(uint64_t w) { return ((double)(w >> 12) + 0.5) * 0x1.0p-52; }  // (0,1)

// Natural logarithm for positive normal doubles: the FreeBSD/musl single-path form of the
// fdlibm algorithm (Sun Microsystems, freely distributable).  Written with explicit
// left-to-right association and no fused multiply-add so that g++ -ffp-contract=off and
// nvcc -fmad=false produce identical bits.
static double contract_log(double x) {
  static const double ln2_hi = 0x1.62e42fee00000p-1, ln2_lo = 0x1.a39ef35793c76p-33,
                      Lg1 = 0x1.5555555555593p-1, Lg2 = 0x1.999999997fa04p-2,
                      Lg3 = 0x1.2492494229359p-2, Lg4 = 0x1.c71c51d8e78afp-3,
                      Lg5 = 0x1.7466496cb03dep-3, Lg6 = 0x1.39a09d078c69fp-3,
                      Lg7 = 0x1.2f112df3e5244p-3;
  uint64_t bits;
  std::memcpy(&bits, &x, 8);
  uint32_t hx = (uint32_t)(bits >> 32);
  uint32_t lx = (uint32_t)bits;
  int k = 0;
  hx += 0x3ff00000u - 0x3fe6a09eu;
  k += (int)(hx >> 20) - 0x3ff;
  hx = (hx & 0x000fffffu) + 0x3fe6a09eu;
  bits = ((uint64_t)hx << 32) | lx;
  std::memcpy(&x, &bits, 8);
  double f = x - 1.0;
  double hfsq = (0.5 * f) * f;
  double s = f / (2.0 + f);
  double z = s * s;
  double w = z * z;
  double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  double R = t2 + t1;
  double dk = (double)k;
  return (((s * (hfsq + R) + dk * ln2_lo) - hfsq) + f) + dk * ln2_hi;
}

extern "C" double orc_log(double x) { return contract_log(x); }
extern "C" void orc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  philox4x32_10(ctr, key, out);
}

// -------------------------------------------------------------------------------------------
// Graph (1-based vertices, ascending neighbour lists like Graphs.jl SimpleGraph.fadjlist)
// -------------------------------------------------------------------------------------------
struct Graph {
  int64_t nv = 0;
  std::vector<std::vector<int32_t>> nbr;  // nbr[v], v in 1..nv
  int64_t lattice_row = 0, lattice_col = 0;
  bool is_lattice = false;
};

static void add_edge_sorted(Graph& g, int32_t a, int32_t b) {
  auto ins = [](std::vector<int32_t>& v, int32_t x) {
    auto it = std::lower_bound(v.begin(), v.end(), x);
    if (it == v.end() || *it != x) v.insert(it, x);
  };
  ins(g.nbr[a], b);
  ins(g.nbr[b], a);
}

static Graph path_graph(int64_t n) {  // Graphs.path_graph
  Graph g;
  g.nv = n;
  g.nbr.assign(n + 1, {});
  for (int64_t i = 1; i < n; ++i) add_edge_sorted(g, (int32_t)i, (int32_t)(i + 1));
  return g;
}

// Graphs.cartesian_product(g, h): id(i, j) = (i-1)*nv(h) + j
static Graph cartesian_product(const Graph& g, const Graph& h) {
  Graph z;
  z.nv = g.nv * h.nv;
  z.nbr.assign(z.nv + 1, {});
  auto id = [&](int64_t i, int64_t j) { return (int32_t)((i - 1) * h.nv + j); };
  for (int64_t i1 = 1; i1 <= g.nv; ++i1)
    for (int32_t i2 : g.nbr[i1])
      if (i1 < i2)
        for (int64_t j = 1; j <= h.nv; ++j) add_edge_sorted(z, id(i1, j), id(i2, j));
  for (int64_t j1 = 1; j1 <= h.nv; ++j1)
    for (int32_t j2 : h.nbr[j1])
      if (j1 < j2)
        for (int64_t i = 1; i <= g.nv; ++i) add_edge_sorted(z, id(i, j1), id(i, j2));
  return z;
}

// Graphs.grid(dims): g = path(dims[1]); for d in dims[2:end]: g = cartesian_product(path(d), g)
static Graph grid(const std::vector<int64_t>& dims) {
  Graph g = path_graph(dims[0]);
  for (size_t k = 1; k < dims.size(); ++k) g = cartesian_product(path_graph(dims[k]), g);
  return g;
}

// spatial_graph(row, col, seed; dim)  src/J_Space.jl:43-64
static Graph* lattice(int dim, int64_t row, int64_t col) {
  Graph* g = new Graph();
  if (dim == 2) {
    *g = grid({row, col, 1});                                                   // :54
  } else {
    int64_t d = (int64_t)std::trunc(std::pow((double)(row * col), 1.0 / 3.0)) + 1;  // :50 / :58
    *g = grid({d, d, d});                                                       // :51 / :59
  }
  g->is_lattice = true;
  g->lattice_row = row;
  g->lattice_col = col;
  return g;
}

// Graphs.closeness_centrality(g; normalize=true) + argmax set, src/J_Space.jl:83-85
static std::vector<int64_t> closeness_argmax(const Graph& g) {
  std::vector<double> c(g.nv + 1, 0.0);
  std::vector<int32_t> dist(g.nv + 1), queue(g.nv + 1);
  for (int64_t u = 1; u <= g.nv; ++u) {
    std::fill(dist.begin(), dist.end(), -1);
    int64_t head = 0, tail = 0;
    queue[tail++] = (int32_t)u;
    dist[u] = 0;
    int64_t reached = 0;
    double sigma = 0.0;  // sum of finite distances (integers: exact in fp64)
    while (head < tail) {
      int32_t v = queue[head++];
      ++reached;
      sigma += (double)dist[v];
      for (int32_t w : g.nbr[v])
        if (dist[w] < 0) {
          dist[w] = dist[v] + 1;
          queue[tail++] = w;
        }
    }
    double l = (double)(reached - 1);
    if (sigma > 0) {
      c[u] = l / sigma;
      double n = l / (double)(g.nv - 1);
      c[u] *= n;
    }
  }
  double best = c[1];
  for (int64_t u = 2; u <= g.nv; ++u) best = std::max(best, c[u]);
  std::vector<int64_t> out;
  for (int64_t u = 1; u <= g.nv; ++u)
    if (c[u] == best) out.push_back(u);
  return out;
}

// -------------------------------------------------------------------------------------------
// Simulation state
// -------------------------------------------------------------------------------------------
struct Sim {
  const Graph& G;
  OrcParams P;
  Rng rng;
  bool faithful_scans;  // true: `pos ∈ cs_alive` and filter! are linear scans, as in the reference

  // MetaGraph vertex props: :id (0 = none), :Subpop.  :Fit == alpha[:Subpop], :mutation ==
  // genotype of clone :Subpop (both invariants of the reference's handlers).
  std::vector<uint32_t> id;
  std::vector<int32_t> subpop;

  std::vector<int32_t> cs_alive;                // src/J_Space.jl:655
  std::vector<std::vector<int32_t>> ca_subpop;  // :683
  std::vector<double> alpha;                    // :664 / :881
  std::vector<OrcClone> set_mut;                // :660 (genotype = chain of labels via parent)
  std::vector<OrcEvent> df;                     // :661
  std::vector<OrcSnapshot> snaps;
  std::vector<uint32_t> series_n;               // list_len_node_occ, :656-658
  bool used_label[1001];
  int n_used = 0;

  int64_t n_cs_alive = 0;
  double t_curr = 0.0;
  uint64_t iter = 0;
  uint32_t next_id = 1;
  OrcSummary sum{};
  // Reference-stream mode (julia_mt.hpp): when set, every draw is taken from the reference's own
  // MersenneTwister in the reference's call order instead of the counter-based draw contract.
  jmt::MersenneTwister* mt = nullptr;
  std::vector<int64_t> mt_randn_pos;  // pinning aid: where in the raw dSFMT output every randn() call started
  const jmt::IntSet* label_set = nullptr;  // slot layout of Set(1:1000), src/J_Space.jl:366
  // benchmark control (bench.py's time-boxed CPU arm): iterations done so far / stop request
  volatile uint64_t* progress = nullptr;
  volatile int* stop = nullptr;

  Sim(const Graph& g, const OrcParams& p, uint64_t seed, uint32_t stream, bool faithful)
      : G(g), P(p), rng{seed, stream}, faithful_scans(faithful) {
    id.assign(g.nv + 1, 0);
    subpop.assign(g.nv + 1, 0);
    std::memset(used_label, 0, sizeof used_label);
  }

  bool in_alive(int32_t pos) const {  // `pos ∈ cs_alive`
    if (faithful_scans) return std::find(cs_alive.begin(), cs_alive.end(), pos) != cs_alive.end();
    return id[pos] != 0;
  }
  static void filter_out(std::vector<int32_t>& v, int32_t x) {  // filter!(e -> e != x, v)
    v.erase(std::remove(v.begin(), v.end(), x), v.end());
  }
  void log(uint8_t type, uint8_t flags, double time, uint32_t cell, uint32_t parent, uint32_t site,
           uint32_t site2, uint16_t clone, uint16_t label) {
    OrcEvent e{};
    e.time = time; e.cell = cell; e.parent = parent; e.site = site; e.site2 = site2;
    e.clone = clone; e.label = label; e.type = type; e.flags = flags; e.pad = 0;
    df.push_back(e);
  }
  void push_series() { series_n.push_back((uint32_t)n_cs_alive); }
  // rand(seed, 1:s) / rand(seed, v::Vector) index, 0-based
  uint64_t draw_index(uint32_t site, uint32_t seq, int half, uint64_t s) {
    return mt ? mt->rand_index(s) : rng.index(iter, site, seq, half, s);
  }

  // initialize_graph, src/J_Space.jl:79-119 (generic) and :122-163 (lattice)
  void seed_cells(const std::vector<int64_t>& candidates) {
    int n_cell = P.n_start_cells;
    std::vector<int32_t> seeded;  // in uuid1 call order
    if (n_cell == 1) {
      int64_t v0;
      if (G.is_lattice) {
        // :131-133 (v₀ is a Float64 compared with Int vertex numbers)
        double row = (double)G.lattice_row, col = (double)G.lattice_col;
        double v = (G.nv % 2 == 0) ? (col * std::floor(row / 2.0)) + (std::floor(col / 2.0) + 1.0)
                                   : std::floor((double)G.nv / 2.0) + 1.0;
        v0 = (int64_t)v;
        if ((double)v0 != v || v0 < 1 || v0 > G.nv) v0 = 0;  // `i == v₀` never true
      } else {
        // :83-86  v₀ = rand(seed, pos_value)
        uint64_t j = mt ? mt->rand_index(candidates.size())  // rand(seed, pos_value) always draws
                        : (candidates.size() > 1 ? rng.index(0, DS_SEED_TIE, 0, 0, candidates.size()) : 0);
        v0 = candidates[j];
      }
      if (v0) seeded.push_back((int32_t)v0);
    } else if (n_cell > 1) {
      // :105-112 / :151-157  positions = Set(1:nv); pos = rand(seed, positions); delete!
      // contract: the j-th smallest remaining position, j uniform
      std::vector<int32_t> chosen_sorted;
      if (mt) {  // reference stream: id = uuid1(seed); pos = rand(seed, positions::Set); delete!(positions, pos)
        jmt::IntSet positions = jmt::IntSet::range(G.nv, mt->var);
        std::vector<char> gone((size_t)G.nv + 1, 0);
        for (int i = 0; i < n_cell; ++i) {
          mt->uuid1();
          int64_t pos = mt->rand_set(positions, reinterpret_cast<const bool*>(gone.data()));
          gone[pos] = 1;
          seeded.push_back((int32_t)pos);
        }
      }
      for (int i = 0; i < n_cell && !mt; ++i) {
        uint64_t remaining = (uint64_t)G.nv - (uint64_t)i;
        uint64_t j = rng.index(0, DS_SEED_POS, (uint32_t)i, 0, remaining);
        int64_t pos = (int64_t)j + 1;  // rank among remaining, 1-based
        for (int32_t c : chosen_sorted)
          if (c <= pos) ++pos;
        chosen_sorted.insert(std::lower_bound(chosen_sorted.begin(), chosen_sorted.end(), (int32_t)pos),
                             (int32_t)pos);
        seeded.push_back((int32_t)pos);
      }
    }
    for (int32_t v : seeded) {
      if (mt && n_cell == 1) mt->uuid1();
      id[v] = next_id++;  // id = uuid1(seed)
      subpop[v] = 1;      // :mutation => 1 ; :Subpop by :675-680
    }
    // cells_alive(G) :296-300 — filter_vertices iterates vertices in ascending order
    for (int64_t v = 1; v <= G.nv; ++v)
      if (id[v]) cs_alive.push_back((int32_t)v);
    n_cs_alive = (int64_t)cs_alive.size();
    push_series();  // :658
    // set_mut_pop = unique(get_drivermut(G)) :660 ; α = [rate_birth] :664 / [row 1 rate] :881
    if (n_cs_alive > 0) {
      OrcClone c{};
      c.alpha = (P.method == 2) ? P.root_rate : P.rate_birth;
      c.parent = 0;
      c.label = (P.method == 2) ? (uint16_t)(P.root_node + 1) : 1;
      set_mut.push_back(c);
      alpha.push_back(c.alpha);
      ca_subpop.push_back(cs_alive);  // cells_alive_subpop :542-550, ascending
      if (P.method != 2) { used_label[1] = true; n_used = 1; }
    } else {
      alpha.push_back((P.method == 2) ? P.root_rate : P.rate_birth);
    }
  }

  // cell_death, src/J_Space.jl:574-586
  void cell_death(int32_t cell, double time, uint8_t flags) {
    log(EV_DEATH, flags, time, id[cell], 0, (uint32_t)cell, 0, (uint16_t)subpop[cell], 0);
    id[cell] = 0;
    subpop[cell] = 0;
  }

  // migration_cell, src/J_Space.jl:592-613
  void migration_cell(int32_t cell, int32_t pos, double time) {
    uint32_t cid = id[cell];
    int32_t sp = subpop[cell];
    id[cell] = 0; subpop[cell] = 0;
    id[pos] = cid; subpop[pos] = sp;
    log(EV_MIGRATION, EVF_GROUP_START, time, cid, 0, (uint32_t)pos, (uint32_t)cell, (uint16_t)sp, 0);
  }

  // randn by the polar method on contract draws
  double randn() {
    if (mt) return mt->randn();
    for (uint32_t attempt = 0;; ++attempt) {
      uint64_t a, b;
      rng.block(iter, 0, DS_NORMAL, attempt, a, b);
      double v1 = 2.0 * u01_53(a) - 1.0;
      double v2 = 2.0 * u01_53(b) - 1.0;
      double s = v1 * v1 + v2 * v2;
      if (s >= 1.0 || s == 0.0) continue;
      return v1 * std::sqrt((-2.0 * contract_log(s)) / s);
    }
  }

  // cell_birth, both methods: src/J_Space.jl:320-416 and :419-537.  Returns false when the
  // replicate must stop (label space exhausted, reference would throw at :371).
  bool cell_birth(int32_t cell, int32_t pos, double time, int idx) {
    // The two uniforms of this birth (driver test and placement coin) are counter-keyed, so
    // they can be read before the displacement below without changing anything.
    uint64_t wa = 0, wb = 0;
    double r;
    if (mt) {  // reference order: (displacement draws nothing) id, id2 = uuid1 x2 :340-341, then r :347
      mt->uuid1();
      mt->uuid1();
      r = mt->rand_float();
    } else {
      rng.block(iter, 0, DS_DRIVER_COIN, 0, wa, wb);
      r = u01_53(wa);  // :347 / :447
    }
    int32_t sp = subpop[cell];

    int child_node = -1;  // method 2: eligible tree child
    bool driver = false;
    if (P.method == 2) {
      if (r <= P.mu_driver) {  // :448
        int last = set_mut[sp - 1].label - 1;  // muts (String) or muts[end]  :450-454
        for (int e = 0; e < P.n_tree_edges; ++e) {  // findall over edge_list[:,1], file order
          if (P.edge_parent[e] != last) continue;
          int cand = P.edge_child[e];  // :457
          // `new_mut ∉ set_mut` :466 — new_mut = genotype(sp) + cand, and genotypes are unique,
          // so it is present iff some clone has parent sp and last driver cand
          bool present = false;
          for (size_t s = 0; s < set_mut.size(); ++s)
            if (set_mut[s].parent == sp && set_mut[s].label - 1 == cand) { present = true; break; }
          if (!present) { child_node = cand; break; }
        }
      }
      driver = child_node >= 0;  // :476
    } else {
      driver = !(r > P.mu_driver);  // :348
    }
    // Clone table full (1000 entries = every label of Set(1:1000) used): the reference throws at
    // :371 (rand on an empty Set).  Contract: stop the replicate before touching any state.
    if (driver && set_mut.size() >= 1000) { sum.status = ST_CLONE_CAPACITY; return false; }

    uint8_t first_flag = EVF_GROUP_START;
    if (id[pos] != 0) {  // :335-338 / :434-437
      filter_out(ca_subpop[subpop[pos] - 1], pos);
      cell_death(pos, time, EVF_DISPLACED | first_flag);
      first_flag = 0;
      sum.n_displaced++;
    }
    uint32_t nid = next_id++;   // id  = uuid1(seed) :340
    uint32_t nid2 = next_id++;  // id2 = uuid1(seed) :341
    uint32_t parent = id[cell];

    if (!driver) {
      // :351-363 / :477-490
      log(EV_DUPLICATE, first_flag, time, nid, parent, (uint32_t)cell, 0, (uint16_t)sp, 0);
      log(EV_DUPLICATE, 0, time, nid2, parent, (uint32_t)pos, 0, (uint16_t)sp, 0);
      id[cell] = nid;
      id[pos] = nid2; subpop[pos] = sp;
      ca_subpop[idx - 1].push_back(pos);
      return true;
    }

    uint16_t new_label;
    double new_alpha;
    if (P.method == 2) {
      new_label = (uint16_t)(child_node + 1);
      new_alpha = P.node_rate[child_node];  // :503-505
    } else {
      // :366-371  possible_mut = Set(1:1000) minus used; new_drive = rand(seed, possible_mut)
      int n_free = 1000 - n_used;
      int lab = 0;
      if (mt) {
        lab = (int)mt->rand_set(*label_set, used_label);  // rand(seed, possible_mut::Set{Int}) :371
      } else {
        uint64_t j = rng.index(iter, DS_LABEL, 0, 0, (uint64_t)n_free);  // rank among unused
        for (int m = 1; m <= 1000; ++m)
          if (!used_label[m]) {
            if (j == 0) { lab = m; break; }
            --j;
          }
      }
      used_label[lab] = true;
      ++n_used;
      new_label = (uint16_t)lab;
      // new_alpha_driver :556-568: α[idx(parent genotype)] + abs(rand(Normal(avg, std)))
      double z = randn();
      new_alpha = alpha[sp - 1] + std::fabs(P.adv_mean + P.adv_std * z);
    }
    alpha.push_back(new_alpha);
    OrcClone c{};
    c.alpha = new_alpha; c.parent = (uint16_t)sp; c.label = new_label;
    set_mut.push_back(c);  // :384 / :506
    int new_sp = (int)set_mut.size();

    bool coin = (mt ? mt->rand_float() : u01_53(wb)) < 0.5;  // :391 / :512
    // :387-388 / :508-509 (rows are pushed before the coin; their content does not depend on it
    // except for the vertex we additionally record)
    log(EV_MUTATION, first_flag, time, nid2, parent, (uint32_t)(coin ? pos : cell), 0, (uint16_t)new_sp,
        new_label);
    log(EV_DUPLICATE, 0, time, nid, parent, (uint32_t)(coin ? cell : pos), 0, (uint16_t)sp, 0);
    if (coin) {
      // :392-400 / :513-521
      id[cell] = nid;
      id[pos] = nid2; subpop[pos] = new_sp;
      ca_subpop.push_back({pos});
    } else {
      // :402-412 / :523-533
      id[pos] = nid; subpop[pos] = sp;
      id[cell] = nid2; subpop[cell] = new_sp;
      ca_subpop[idx - 1].push_back(pos);
      filter_out(ca_subpop[idx - 1], cell);
      ca_subpop.push_back({cell});
    }
    return true;
  }

  // StatsBase.sample(rng, a, k; replace=false) (0.33.21 sample! branch structure), drawing
  // through the contract's excision site.  Returns the sampled vertices in sample order.
  std::vector<int32_t> sample_noreplace(const std::vector<int32_t>& a, int64_t k) {
    std::vector<int32_t> x((size_t)k);
    int64_t n = (int64_t)a.size();
    uint32_t q = 0;
    auto draw = [&](uint64_t s) { return (int64_t)(mt ? mt->rand_index(s) : rng.index(iter, DS_EXCISION, q++, 0, s)); };
    if (k == 0) return x;
    if (k == 1) {
      x[0] = a[draw((uint64_t)n)];
    } else if (k == 2) {  // samplepair
      int64_t i1 = draw((uint64_t)n) + 1;
      int64_t i2 = draw((uint64_t)(n - 1)) + 1;
      x[0] = a[i1 - 1];
      x[1] = a[(i2 == i1 ? n : i2) - 1];
    } else if (n < k * 24) {  // fisher_yates_sample!
      std::vector<int64_t> inds((size_t)n);
      for (int64_t i = 0; i < n; ++i) inds[i] = i;
      for (int64_t i = 0; i < k; ++i) {
        int64_t j = i + draw((uint64_t)(n - i));  // rand(i:n)
        std::swap(inds[i], inds[j]);
        x[i] = a[inds[i]];
      }
    } else {  // self_avoid_sample!
      std::vector<char> seen((size_t)n, 0);
      for (int64_t i = 0; i < k; ++i) {
        int64_t idx = draw((uint64_t)n);
        while (seen[idx]) idx = draw((uint64_t)n);
        seen[idx] = 1;
        x[i] = a[idx];
      }
    }
    return x;
  }

  // The while loop of simulate_evolution, src/J_Space.jl:687-840 / :904-1066
  void run() {
    const int64_t nvG = G.nv;
    int id_bottleneck = (P.n_bottleneck > 0) ? 1 : 0;  // :668-672
    int time_index = 1;                                // :652
    std::vector<double> a_sub, prob_vet, prob_cum;

    while ((t_curr < P.Tf && n_cs_alive > 0) || (P.rate_death == 0 && n_cs_alive == nvG)) {
      if (!(t_curr < P.Tf && n_cs_alive > 0)) {  // only the second clause holds: never exits
        sum.status = ST_NONTERMINATING;
        break;
      }
      if (P.max_iterations && iter >= P.max_iterations) { sum.status = ST_MAX_ITER; break; }
      if (stop && *stop) { sum.status = ST_MAX_ITER; break; }
      ++iter;
      if (progress) *progress = iter;
      const int S = (int)alpha.size();
      a_sub.resize(S);
      for (int i = 0; i < S; ++i) a_sub[i] = alpha[i] * (double)ca_subpop[i].size();  // :692-694
      double birth = a_sub[0];  // sum(): sequential left fold (Base mapreduce, n < 1024)
      for (int i = 1; i < S; ++i) birth = birth + a_sub[i];
      double death = P.rate_death * (double)n_cs_alive;     // :696
      double M = P.rate_migration * (double)n_cs_alive;     // :697
      double lambda = birth + death + M;                    // :698
      uint64_t wa = 0, wb = 0;
      double t_event;
      if (mt) {
        t_event = mt->randexp() * (1.0 / lambda);  // rand(seed, Exponential(1 / λ), 1)[1] :699
      } else {
        rng.block(iter, 0, DS_TIME_CLASS, 0, wa, wb);
        t_event = (-contract_log(u01_open(wa))) * (1.0 / lambda);  // :699 θ·randexp
      }
      double t_old = t_curr;                                // :700
      t_curr += t_event;                                    // :701

      // snapshot :708-714
      if (P.n_snapshots > 0 && time_index <= P.n_snapshots && t_curr > P.t_snapshot[time_index - 1]) {
        OrcSnapshot s{};
        s.time = t_curr; s.iteration = iter; s.n_events_before = df.size();
        s.n_alive = (uint32_t)n_cs_alive; s.index = (uint32_t)(time_index - 1);
        snaps.push_back(s);
        ++time_index;
      }

      // prob_vet = vcat(Aₙ, Bₙ, Mₙ) ; prob_cum = cumsum(prob_vet)  :703-718
      prob_vet.resize(S + 2);
      for (int i = 0; i < S; ++i) prob_vet[i] = a_sub[i] / lambda;
      prob_vet[S] = death / lambda;
      prob_vet[S + 1] = M / lambda;
      prob_cum.resize(S + 2);
      accumulate_pairwise(prob_vet, prob_cum);

      double k = mt ? mt->rand_float() : u01_53(wb);  // :721
      int min = 0;            // findfirst(k .<= prob_cum) :722-723 (1-based; 0 = nothing)
      for (int i = 0; i < S + 2; ++i)
        if (k <= prob_cum[i]) { min = i + 1; break; }

      bool stop = false;
      if (min == 0) {
        // reference: `nothing` -> MethodError.  Contract: null iteration.
      } else if (min == S + 2) {  // Migration :725-742
        int64_t x = (int64_t)draw_index(DS_CELL_NBR, 0, 0, (uint64_t)n_cs_alive);
        int32_t cell = cs_alive[x];
        const auto& nb = G.nbr[cell];
        if (!nb.empty()) {
          int32_t pos = nb[draw_index(DS_CELL_NBR, 0, 1, nb.size())];
          if (!in_alive(pos)) {
            int idx = subpop[cell];
            migration_cell(cell, pos, t_curr);
            cs_alive.push_back(pos);
            ca_subpop[idx - 1].push_back(pos);
            filter_out(cs_alive, cell);
            filter_out(ca_subpop[idx - 1], cell);
            push_series();
            sum.n_migrations++;
          }
        }
      } else if (min == S + 1) {  // Death :744-759
        int64_t x = (int64_t)draw_index(DS_CELL_NBR, 0, 0, (uint64_t)n_cs_alive);
        int32_t cell = cs_alive[x];
        int idx = subpop[cell];
        cell_death(cell, t_curr, EVF_GROUP_START);
        filter_out(cs_alive, cell);
        filter_out(ca_subpop[idx - 1], cell);
        n_cs_alive -= 1;
        push_series();
        sum.n_deaths++;
      } else {  // Birth :761-810
        const auto& members = ca_subpop[min - 1];
        if (!members.empty()) {  // rand(seed, 1:0) would throw; contract: null iteration
          int64_t x = (int64_t)draw_index(DS_CELL_NBR, 0, 0, members.size());
          int32_t cell = members[x];
          const auto& nb = G.nbr[cell];
          if (!nb.empty()) {
            int32_t pos = nb[draw_index(DS_CELL_NBR, 0, 1, nb.size())];
            bool rule = true;  // :767-786
            if (P.rule == RULE_CONTACT) {
              if (in_alive(pos)) rule = false;
            } else if (P.rule == RULE_VOTER) {
              if (in_alive(pos) && subpop[cell] == subpop[pos]) rule = false;
            } else if (P.rule == RULE_HVOTER) {
              if (in_alive(pos) && alpha[subpop[cell] - 1] <= alpha[subpop[pos] - 1]) rule = false;
            }
            if (rule) {
              bool was_alive = in_alive(pos);
              if (!cell_birth(cell, pos, t_curr, min)) {
                stop = true;
              } else {
                if (!was_alive) {  // :803-806 (cell_birth never changes membership of pos)
                  cs_alive.push_back(pos);
                  n_cs_alive += 1;
                }
                push_series();
                sum.n_births++;
              }
            }
          }
        }
      }
      if (stop) break;

      // excision ("bottleneck") :813-838 / :1036-1065
      if (id_bottleneck != 0) {
        double tb = P.t_bottleneck[id_bottleneck - 1];
        bool fired = false;
        if (tb > t_old && tb <= t_curr) {
          double ratio = P.ratio_bottleneck[id_bottleneck - 1];
          double n_cells_taked = (double)n_cs_alive * ratio;
          int64_t k_take = (int64_t)std::trunc(n_cells_taked);
          std::vector<int32_t> taken = sample_noreplace(cs_alive, k_take);
          bool first = true;
          for (int32_t ct : taken) {
            int idx = subpop[ct];
            cell_death(ct, tb, (uint8_t)(EVF_EXCISION | (first ? EVF_GROUP_START : 0)));
            first = false;
            filter_out(cs_alive, ct);
            filter_out(ca_subpop[idx - 1], ct);
            n_cs_alive -= 1;
            sum.n_excised++;
          }
          push_series();
          fired = true;
        }
        if (P.quirk_flags & QUIRK_EXCISION_INTENT) {
          if (fired) id_bottleneck = (id_bottleneck < P.n_bottleneck) ? id_bottleneck + 1 : 0;
        } else if (P.method == 2) {
          if (id_bottleneck < P.n_bottleneck) id_bottleneck += 1;  // :1062-1064, every iteration
        }  // method 1: never advanced
      }
    }
    sum.n_iterations = iter;
    sum.n_events = df.size();
    sum.t_final = t_curr;
    sum.n_effective = (uint32_t)series_n.size() - 1;
    sum.n_alive = (uint32_t)n_cs_alive;
    sum.n_clones = (uint32_t)set_mut.size();
    sum.next_cell_id = next_id;
    for (size_t i = 0; i < set_mut.size(); ++i) set_mut[i].count = (uint32_t)ca_subpop[i].size();
  }

  // Base.accumulate_pairwise!(add_sum, c, v) — what cumsum(::Vector{Float64}) runs (Julia 1.7.3)
  static double acc_pairwise_rec(const std::vector<double>& v, std::vector<double>& c, double s, size_t i1,
                                 size_t n) {
    double s_;
    if (n < 128) {
      s_ = v[i1];
      c[i1] = s + s_;
      for (size_t i = i1 + 1; i < i1 + n; ++i) {
        s_ = s_ + v[i];
        c[i] = s + s_;
      }
    } else {
      size_t n2 = n >> 1;
      s_ = acc_pairwise_rec(v, c, s, i1, n2);
      s_ = s_ + acc_pairwise_rec(v, c, s + s_, i1 + n2, n - n2);
    }
    return s_;
  }
  static void accumulate_pairwise(const std::vector<double>& v, std::vector<double>& c) {
    size_t n = v.size();
    if (n == 0) return;
    c[0] = v[0];
    if (n == 1) return;
    acc_pairwise_rec(v, c, v[0], 1, n - 1);
  }
};

// -------------------------------------------------------------------------------------------
// C interface for ctypes (tests, smoke, cpu_baseline)
// -------------------------------------------------------------------------------------------
extern "C" {

void* orc_graph_lattice(int dim, int64_t row, int64_t col) { return lattice(dim, row, col); }

void* orc_graph_csr(int64_t nv, const int64_t* rowptr, const int32_t* colidx) {
  Graph* g = new Graph();
  g->nv = nv;
  g->nbr.assign(nv + 1, {});
  for (int64_t v = 1; v <= nv; ++v)
    for (int64_t e = rowptr[v - 1]; e < rowptr[v]; ++e) g->nbr[v].push_back(colidx[e]);
  for (int64_t v = 1; v <= nv; ++v) std::sort(g->nbr[v].begin(), g->nbr[v].end());
  return g;
}

// readdlm(path, Int) + the evident intent SimpleGraph(adj), src/J_Space.jl:67-71
void* orc_graph_dense_file(const char* path) {
  FILE* f = std::fopen(path, "r");
  if (!f) return nullptr;
  std::vector<long> vals;
  long x;
  while (std::fscanf(f, "%ld", &x) == 1) vals.push_back(x);
  std::fclose(f);
  int64_t n = (int64_t)std::llround(std::sqrt((double)vals.size()));
  if (n * n != (int64_t)vals.size()) return nullptr;
  Graph* g = new Graph();
  g->nv = n;
  g->nbr.assign(n + 1, {});
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j < n; ++j)
      if (vals[i * n + j] != 0 && i != j) add_edge_sorted(*g, (int32_t)(i + 1), (int32_t)(j + 1));
  return g;
}

int64_t orc_graph_nv(void* g) { return ((Graph*)g)->nv; }
int orc_graph_neighbors(void* g, int64_t v, int64_t* out, int cap) {
  const auto& nb = ((Graph*)g)->nbr[v];
  for (int i = 0; i < (int)nb.size() && i < cap; ++i) out[i] = nb[i];
  return (int)nb.size();
}
int64_t orc_graph_closeness_argmax(void* g, int64_t* out, int64_t cap) {
  auto v = closeness_argmax(*(Graph*)g);
  for (int64_t i = 0; i < (int64_t)v.size() && i < cap; ++i) out[i] = v[i];
  return (int64_t)v.size();
}
void orc_graph_free(void* g) { delete (Graph*)g; }

// One replicate.  `candidates` = seed candidates for generic graphs with n_cell == 1 (may be
// null for lattices).  faithful_scans != 0 keeps the reference's O(n) membership tests.
void* orc_run(void* g, const OrcParams* p, uint64_t seed, uint32_t stream, const int64_t* candidates,
              int64_t n_candidates, int faithful_scans) {
  Graph* G = (Graph*)g;
  Sim* s = new Sim(*G, *p, seed, stream, faithful_scans != 0);
  std::vector<int64_t> cand;
  if (candidates) cand.assign(candidates, candidates + n_candidates);
  if (!G->is_lattice && p->n_start_cells == 1 && cand.empty()) cand = closeness_argmax(*G);
  s->seed_cells(cand);
  s->run();
  return s;
}
// Same, for bench.py's time-boxed CPU arm: *progress is updated every iteration, and the run stops
// (status MAX_ITER, partial summary) as soon as *stop is non-zero.
void* orc_run_ctl(void* g, const OrcParams* p, uint64_t seed, uint32_t stream, const int64_t* candidates,
                  int64_t n_candidates, int faithful_scans, volatile uint64_t* progress, volatile int* stop) {
  Graph* G = (Graph*)g;
  Sim* s = new Sim(*G, *p, seed, stream, faithful_scans != 0);
  s->progress = progress;
  s->stop = stop;
  std::vector<int64_t> cand;
  if (candidates) cand.assign(candidates, candidates + n_candidates);
  if (!G->is_lattice && p->n_start_cells == 1 && cand.empty()) cand = closeness_argmax(*G);
  s->seed_cells(cand);
  s->run();
  return s;
}
// One replicate on the REFERENCE's own random stream: MersenneTwister(mt_seed) restated in julia_mt.hpp, every
// draw in the reference's call order (src/Start.jl:12-13 -> spatial_graph -> simulate_evolution).  `variant`
// = {set_slots, sizehint_mode, hash_float, range_fast, set_ndl, int_fill_tail}; null = the defaults.
void* orc_run_mt(void* g, const OrcParams* p, uint64_t mt_seed, const int32_t* variant, const int64_t* candidates,
                 int64_t n_candidates) {
  Graph* G = (Graph*)g;
  jmt::Variant v;
  if (variant) {
    v.set_slots = variant[0]; v.sizehint_mode = variant[1]; v.hash_float = variant[2] != 0;
    v.range_fast = variant[3] != 0; v.set_ndl = variant[4] != 0; v.int_fill_tail = variant[5]; v.test_float_start = variant[6]; v.refill_pad = variant[7];
  }
  jmt::MersenneTwister mt(mt_seed, v);
  jmt::IntSet labels = jmt::IntSet::range(1000, v);
  Sim* s = new Sim(*G, *p, 0, 0, false);
  s->mt = &mt;
  s->label_set = &labels;
  std::vector<int64_t> cand;
  if (candidates) cand.assign(candidates, candidates + n_candidates);
  if (!G->is_lattice && p->n_start_cells == 1 && cand.empty()) cand = closeness_argmax(*G);
  s->seed_cells(cand);
  s->run();
  s->mt_randn_pos = mt.randn_pos;
  s->mt = nullptr;
  s->label_set = nullptr;
  return s;
}
// the first n doubles rand(MersenneTwister(seed)) returns, and one randn / randexp after them (known-answer checks)
void orc_mt_probe(uint64_t mt_seed, double* floats, int n, double* normal, double* expo) {
  jmt::Variant v;
  jmt::MersenneTwister a(mt_seed, v), b(mt_seed, v), c(mt_seed, v);
  for (int i = 0; i < n; ++i) floats[i] = a.rand_float();
  *normal = b.randn();
  *expo = c.randexp();
}
// the raw dSFMT output of MersenneTwister(seed): n doubles in [1, 2) as bit patterns, in generation order
void orc_mt_raw(uint64_t mt_seed, uint64_t* out, int64_t n) {
  jmt::Variant v;
  jmt::MersenneTwister m(mt_seed, v);
  for (int64_t k = 0; k + 1 < n; k += 2) m.ds.next(out + k);
}
// pinning aid: out[k] = the value randexp() returns when its first raw double is dSFMT output k and the ziggurat
// accepts at once (NaN otherwise: 1.1 % of the calls)
void orc_mt_randexp_scan(uint64_t mt_seed, int64_t n, double* out) {
  jmt::Variant v;
  jmt::MersenneTwister m(mt_seed, v);
  const jmt::Zig& z = jmt::MersenneTwister::zig();
  uint64_t w[2];
  for (int64_t k = 0; k + 1 < n; k += 2) {
    m.ds.next(w);
    for (int h = 0; h < 2; ++h) {
      const uint64_t ri = w[h] & 0x000fffffffffffffULL;
      const unsigned idx = (unsigned)(ri & 0xFF);
      out[k + h] = ri < z.ke[idx] ? (double)ri * z.we[idx] : std::nan("");
    }
  }
}
// rand(MersenneTwister(seed), 1:n[k]) for k = 0 .. count-1 from ONE generator (0-based results)
void orc_mt_range_seq(uint64_t mt_seed, const int64_t* n, int64_t count, int64_t* out) {
  jmt::Variant v;
  jmt::MersenneTwister m(mt_seed, v);
  for (int64_t k = 0; k < count; ++k) out[k] = (int64_t)m.rand_index((uint64_t)n[k]);
}
// Base.hash(::Int64) as restated in julia_mt.hpp (float_form = 0: hash_64_64(x), the form of Julia 1.x for which
// hash(1) == 0x5bca7c69b794f8ce)
uint64_t orc_julia_hash(int64_t x, int float_form) { return jmt::julia_hash_int(x, float_form != 0); }
// slot layout of Set(1:n) under the Variant knobs: out[i] = key in slot i (0 = empty); returns the number of slots
int64_t orc_intset_layout(int64_t n, const int32_t* variant, int64_t* out, int64_t cap) {
  jmt::Variant v;
  v.set_slots = variant[0]; v.sizehint_mode = variant[1]; v.hash_float = variant[2] != 0;
  jmt::IntSet s = jmt::IntSet::range(n, v);
  const int64_t sz = (int64_t)s.slots.size();
  for (int64_t i = 0; i < sz && i < cap; ++i) out[i] = s.slots[i] == 1 ? s.keys[i] : 0;
  return sz;
}
// n successive randexp() (kind 0) / randn() (kind 1) / rand() (kind 2) values of ONE MersenneTwister(seed)
void orc_mt_seq(uint64_t mt_seed, int kind, int64_t n, double* out) {
  jmt::Variant v;
  jmt::MersenneTwister m(mt_seed, v);
  for (int64_t k = 0; k < n; ++k) out[k] = kind == 0 ? m.randexp() : (kind == 1 ? m.randn() : m.rand_float());
}
// the first n values rand(MersenneTwister(seed), UInt64) returns (integer cache of julia_mt.hpp)
void orc_mt_u64(uint64_t mt_seed, int64_t n, uint64_t* out) {
  jmt::Variant v;
  jmt::MersenneTwister m(mt_seed, v);
  for (int64_t k = 0; k < n; ++k) out[k] = m.rand_u64();
}
// pinning aid: out[k] = the value randn() returns when its first raw double is dSFMT output k of
// MersenneTwister(seed) and the ziggurat accepts at once (NaN otherwise: 1.2 % of the calls)
void orc_mt_randn_scan(uint64_t mt_seed, int64_t n, double* out) {
  jmt::Variant v;
  jmt::MersenneTwister m(mt_seed, v);
  const jmt::Zig& z = jmt::MersenneTwister::zig();
  uint64_t w[2];
  for (int64_t k = 0; k + 1 < n; k += 2) {
    m.ds.next(w);
    for (int h = 0; h < 2; ++h) {
      const uint64_t r = w[h] & 0x000fffffffffffffULL;
      const int64_t rabs = (int64_t)(r >> 1);
      const unsigned idx = (unsigned)(rabs & 0xFF);
      out[k + h] = (uint64_t)rabs < z.ki[idx] ? (double)((r & 1) ? -rabs : rabs) * z.wi[idx] : std::nan("");
    }
  }
}
int64_t orc_mt_randn_pos(void* s, const int64_t** p) {
  Sim* S = (Sim*)s;
  *p = S->mt_randn_pos.data();
  return (int64_t)S->mt_randn_pos.size();
}
const OrcSummary* orc_summary(void* s) { return &((Sim*)s)->sum; }
int64_t orc_events(void* s, const OrcEvent** p) {
  Sim* S = (Sim*)s;
  *p = S->df.data();
  return (int64_t)S->df.size();
}
int64_t orc_clones(void* s, const OrcClone** p) {
  Sim* S = (Sim*)s;
  *p = S->set_mut.data();
  return (int64_t)S->set_mut.size();
}
int64_t orc_snapshots(void* s, const OrcSnapshot** p) {
  Sim* S = (Sim*)s;
  *p = S->snaps.data();
  return (int64_t)S->snaps.size();
}
int64_t orc_series(void* s, const uint32_t** p) {
  Sim* S = (Sim*)s;
  *p = S->series_n.data();
  return (int64_t)S->series_n.size();
}
// final lattice: clone id (uint16) and cell id (uint32) per vertex, index 0 = vertex 1
void orc_lattice(void* s, uint16_t* clone, uint32_t* cell_id) {
  Sim* S = (Sim*)s;
  for (int64_t v = 1; v <= S->G.nv; ++v) {
    clone[v - 1] = (uint16_t)S->subpop[v];
    cell_id[v - 1] = S->id[v];
  }
}
// list 0 = cs_alive, list c >= 1 = ca_subpop[c]
int64_t orc_list(void* s, int32_t list, const int32_t** p) {
  Sim* S = (Sim*)s;
  const std::vector<int32_t>& v = list == 0 ? S->cs_alive : S->ca_subpop[list - 1];
  *p = v.data();
  return (int64_t)v.size();
}
void orc_free(void* s) { delete (Sim*)s; }

}  // extern "C"
