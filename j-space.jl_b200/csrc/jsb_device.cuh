// jsb_device.cuh — shared host/device declarations of the SSA engine (internal, not ABI).
#pragma once
#include <cstdint>

#include "jspace_b200.h"

namespace jsb {

constexpr int kMaxClones = 1000;   // Set(1:1000), src/J_Space.jl:366
constexpr int kTab = 1024;         // clone-table stride
constexpr int kSmemTable = 1024;   // class boundaries (S+2 <= 1002), padded with +inf, per-warp smem slice
constexpr double kGuardEps = 1e-11; // class-selection guard band (relative to λ), DESIGN.md §4
constexpr int kLogChunk = 1024;    // records per event-log chunk (32 KB)
constexpr int kMinListCap = 32;    // initial capacity of a ca_subpop list in the arena
#ifndef JSB_MAXWARPS
#define JSB_MAXWARPS 12
#endif
constexpr uint32_t kDecMinClones = 64;  // decoupled mode: below this many clones the exact fold is cheaper than the hand-off (measured)
constexpr int kMaxLook = 4;  // lookahead depth: helper warps evaluating later windows for one busy warp
constexpr int kMaxHelp = kMaxLook + 1;  // plus the clock helper (exact fold and time chain off the critical path)
constexpr int kMaxWarpsPerBlock = JSB_MAXWARPS;  // one block per SM; warps per block chosen at launch

// Draw sites of the draw contract; ctr.w = site << 28 | seq
enum DrawSite : uint32_t {
  DS_TIME_CLASS = 0,
  DS_CELL_NBR = 1,
  DS_DRIVER_COIN = 4,
  DS_LABEL = 5,
  DS_NORMAL = 6,
  DS_EXCISION = 7,
  DS_SEED_TIE = 9,
  DS_SEED_POS = 10
};

struct DevGraph {
  int kind;  // 0 = implicit lattice (n1 fastest), 1 = CSR
  uint32_t n1, n2, n3;
  uint32_t nv;
  uint32_t m1, s1, m2, s2;  // v / n1 == (v * m1) >> s1 and yz / n2 == (yz * m2) >> s2 for v < 2^28
  const uint32_t* rowptr;  // [nv+1]
  const uint32_t* colidx;  // 0-based neighbour ids, ascending per row
};

struct DevPoint {
  double Tf, rate_birth, rate_death, rate_migration, mu, adv_mean, adv_std, root_rate;
  uint64_t max_iter;
  int32_t method, rule, n_start, n_bott, n_snap, n_nodes, n_edges, root_node;
  uint32_t quirks;
  int32_t bott_off;  // into dpool: t_bottleneck[n_bott] then ratio[n_bott]
  int32_t snap_off;  // into dpool: t_snapshot[n_snap]
  int32_t rate_off;  // into dpool: node_rate[n_nodes]
  int32_t edge_off;  // into ipool: edge_parent[n_edges] then edge_child[n_edges]
  int32_t pad;
};

// Workspace layout per slot (all offsets 16-byte aligned); computed identically on host and device.
struct WsLayout {
  uint64_t clone, cellid, A, a_head, l_head, lseg, arena, inds, samp, c_cnt, c_cap, c_alpha, c_parent, c_label, g_w, g_c, ftree,
      total;
};
struct RunArgs {
  DevGraph g;
  const DevPoint* points;
  const double* dpool;
  const int32_t* ipool;
  const uint32_t* seed_cand;  // 0-based candidate seed vertices (n_cell == 1)
  uint32_t n_cand;
  uint32_t flags;
  uint64_t seed;
  int64_t g_begin;
  int64_t n_local;
  int64_t reps_per_point;
  // per-slot workspace
  uint8_t* ws;
  uint64_t ws_stride;
  uint32_t arena_cap;  // entries
  uint32_t site_bytes; // 2 when nv <= 65535, else 4 (storage type of the ordered lists and of the per-clone tables)
  uint32_t occ_words;  // per-warp shared occupancy bitmask words (0 = lattice too large, use global)
  uint32_t smem_per_warp;  // bytes of dynamic shared memory per warp
  WsLayout L;          // workspace offsets (computed once on the host)
  // outputs
  jsb_summary* summaries;
  jsb_clone* clone_pool;
  unsigned long long clone_pool_cap;
  unsigned long long* clone_pool_next;
  long long* clone_off;  // per replicate, -1 = pool overflow
  jsb_event* log_pool;
  uint32_t n_chunks;
  uint32_t* chunk_next;
  int32_t* chunk_owner;
  uint32_t* chunk_seq;
  jsb_snapshot* snaps;
  int32_t max_snap;
  uint16_t* out_clone;
  uint32_t* out_cell;
  uint32_t* out_lists;  // [n_local][2][nv]: cs_alive then concatenated ca_subpop
  unsigned long long* work_counter;
  uint32_t fast_smem_tree;  // rejection-free mode: upper tree levels live in shared memory
  uint32_t helpers;         // exact mode: idle warps of a block evaluate the next window for busy ones
  uint32_t dec_min_clones;  // decoupled mode only for replicates with at least this many clones
  uint32_t clock_max_busy;  // a clock helper is attached only while at most this many replicates run on the SM
  uint32_t static_map;      // fewer replicates than warps: replicate idx is owned by block idx % grid
  // optional longest-first scheduling: a pilot pass runs every replicate for
  // `pilot_iters` iterations and writes a work estimate; the full pass then takes replicates in `order`
  const uint32_t* order;  // nullptr: natural order
  uint64_t pilot_iters;   // > 0: pilot pass (no outputs except prio)
  unsigned long long* trace;  // JSB_TRACE: per replicate {start ns, end ns, sm<<8|warp, iterations}
  float* prio;
  // Suspending pilot: the pilot pass runs with every output on, parks the state of each replicate that reaches
  // `pilot_iters` in `susp` (its slot workspace + registers + shared tables) and finishes the others; the full
  // pass resumes the parked ones.  Same draws, same order: the trajectory does not know it was interrupted.
  // Streamed outputs: a finished replicate takes (finish rank, first row of its log in the packed output) with one
  // atomic on fin_ctr = rank << 40 | rows, copies its log chunks to gather_out[first row ...] (walking chunk_prev
  // back from its last chunk) and then publishes fin_list[rank] = {local index (bit 31: not packed, no room), rows
  // stored, first row} as a single 16-byte store over a 0xFF-filled array.  The host copies the packed rows of
  // completed ranks to the host while the kernel is still running (jsb_api.cu).
  uint4* fin_list;
  unsigned long long* fin_ctr;
  jsb_event* gather_out;          // packed output (device); a finished replicate copies its own chunks there
  unsigned long long gather_cap;  // rows that fit (gather buffer and host block of an earlier call)
  uint32_t* chunk_prev;           // [n_chunks] previous chunk of the same replicate (kNone for its first)
  uint8_t* susp;            // [n_local][susp_stride]; nullptr: the pilot pass is thrown away (old behaviour)
  uint64_t susp_stride;     // ws_stride + susp_extra_bytes(site_bytes)
  uint32_t* susp_state;     // [n_local] 0 = not started, 1 = parked, 2 = finished in the pilot pass
};
__host__ __device__ inline uint64_t susp_extra_bytes(uint32_t site_bytes) { return 256 + 128 + 2ull * kTab * site_bytes; }

__host__ __device__ inline uint64_t jsb_align16(uint64_t x) { return (x + 15) & ~uint64_t(15); }
__host__ __device__ inline WsLayout ws_layout(uint32_t nv, uint32_t arena_cap, uint32_t site_bytes) {
  WsLayout L;
  uint64_t o = 0;
  L.clone = o;    o = jsb_align16(o + 2ull * nv);
  L.cellid = o;   o = jsb_align16(o + 4ull * nv);
  L.A = o;        o = jsb_align16(o + (uint64_t)site_bytes * ((uint64_t)nv + 1 + 512));  // tiered vector: whole segments of 256
  L.a_head = o;   o = jsb_align16(o + 2ull * (nv / 256 + 4));
  L.l_head = o;   o = jsb_align16(o + 2ull * (arena_cap / 256 + 8));  // segment heads of the tiered ca_subpop blocks
  L.lseg = o;     o = jsb_align16(o + (uint64_t)site_bytes * nv);  // per vertex: segment of its clone's list (hint)
  L.arena = o;    o = jsb_align16(o + (uint64_t)site_bytes * arena_cap + 32);
  L.inds = o;     o = jsb_align16(o + 4ull * nv);
  L.samp = o;     o = jsb_align16(o + 4ull * nv);
  L.c_cnt = o;    o = jsb_align16(o + 4ull * kTab);  // rejection-free mode: clone sizes (exact mode keeps them in smem)
  L.c_cap = o;    o = jsb_align16(o + 4ull * kTab);
  L.c_alpha = o;  o = jsb_align16(o + 8ull * kTab);
  L.c_parent = o; o = jsb_align16(o + 2ull * kTab);
  L.c_label = o;  o = jsb_align16(o + 2ull * kTab);
  L.g_w = o;      o = jsb_align16(o + 8ull * (kTab + 8));
  L.g_c = o;      o = jsb_align16(o + 8ull * (kTab + 8));
  L.ftree = o;    o = jsb_align16(o + 8ull * ((uint64_t)nv / 31 + 256));  // fast mode: upper tree levels (geometric sum)
  L.total = (o + 255) & ~uint64_t(255);
  return L;
}

// host-callable launch wrapper (jsb_kernels.cu)
int launch_ssa(const RunArgs& args, int grid_blocks, int warps_per_block, void* stream);
int fast_prepare(RunArgs& args, int max_warps_per_sm, int* warps_per_block);
int launch_fast(const RunArgs& args, int grid_blocks, int warps_per_block, void* stream);
int ssa_prepare(RunArgs& args, int max_warps_per_sm, int* warps_per_block);  // fills occ_words / smem_per_warp

}  // namespace jsb
