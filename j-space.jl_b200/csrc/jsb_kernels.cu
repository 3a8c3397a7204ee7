// jsb_kernels.cu — the ensemble SSA kernel (sm_100a).
//
// One independent replicate of the reference's `simulate_evolution` loop
// (src/J_Space.jl:687-840 method 1, :904-1066 method 2) per WARP; persistent warps pull
// replicate numbers from a global counter (replicate run lengths differ by orders of magnitude).
//
// How a warp runs a strictly sequential algorithm with 32 lanes:
//   * The draw contract is counter-based (Philox4x32-10 keyed by seed, stream, iteration, draw
//     site), so the random words of iterations it+1..it+32 are computed by 32 lanes at once.
//   * 79-97 % of the reference's iterations are null events (target site occupied) that do not
//     change any state.  The warp therefore evaluates 32 consecutive iterations speculatively
//     against the current state (class selection, member lookup, neighbour, occupancy test), finds
//     the first lane whose iteration changes state (ballot + ffs), commits all earlier
//     null iterations in one go, applies that one event cooperatively (vectorised list shifts,
//     searches, log stores) and re-speculates from the next iteration.
//   * Event times are a sequential fp64 chain (t += θ·e); it is run over the committed lanes
//     only, and per-lane times are materialised only when the batch may touch Tf, an excision
//     time or a snapshot time.
//   * Event-class selection.  The reference recomputes, every iteration, w_i = α_i·n_i, their
//     sequential sum, λ, the vector w/λ and its pairwise cumsum, then findfirst(k .<= prob_cum).
//     Here the sequential prefix sums B_i = fl(B_{i-1} + w_i) are kept (they ARE the reference's
//     `sum`, so λ and θ are bit-exact) and re-folded only from the first clone whose size changed.
//     A class is selected by comparing k·λ with B_i; because |prob_cum_i − B_i/λ| < 4·(S+2)·2^-53,
//     the choice is provably the reference's whenever k·λ is farther than ε·λ (ε = 1e-11) from
//     both neighbouring boundaries.  Otherwise (≈ 1e-8 of iterations) the exact prob_cum table is
//     built in the reference's order and used instead.
// Everything that must match the reference bit-for-bit (sum order of Σα·n, the pairwise cumsum,
// insertion-ordered lists with stable deletes) is kept in the reference's order; see DESIGN.md.
//
// fp64 determinism: this translation unit is compiled with -fmad=false and uses only IEEE
// + - * / sqrt in a fixed association, so event times and fitness values are bit-identical to
// the CPU oracle's.

#include <cuda_runtime.h>

#include <cstdio>

#include "jsb_device.cuh"

namespace jsb {

#define FULL 0xffffffffu
constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint32_t kSeg = 256;  // tiered-vector segment length (cs_alive and every long ca_subpop list)

// Optional phase profiler (build with -DJSB_PROFILE; debug library only): per-replicate cycle
// totals per phase, printed by lane 0 for the first replicates.
#ifdef JSB_PROFILE
enum { PH_REFOLD, PH_DRAWS, PH_CLASS, PH_EVAL, PH_TIME, PH_BIRTH, PH_DEATH, PH_MIGR, PH_EXC, PH_N };
#define PROF_DECL long long prof_c[PH_N] = {0}; long long prof_n[PH_N] = {0}; long long prof_t0 = 0; long long prof_fold = 0;
#define PROF_T0() prof_t0 = clock64()
#define PROF_ADD(ph) do { long long _t = clock64(); prof_c[ph] += _t - prof_t0; prof_n[ph]++; prof_t0 = _t; } while (0)
__device__ long long g_sub[8];  // sub-phase cycle totals of one profiled replicate (lane 0)
#define SUB_T0() long long _s0 = clock64()
#define SUB_ADD(k) do { long long _s1 = clock64(); if ((threadIdx.x & 31) == 0) g_sub[k] += _s1 - _s0; _s0 = _s1; } while (0)
#else
#define PROF_DECL
#define PROF_T0()
#define PROF_ADD(ph)
#define SUB_T0()
#define SUB_ADD(k)
#endif

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 and the draw contract
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t o[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    if (r > 0) {
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

struct Key {
  uint32_t k0, k1, stream;
};

__device__ __forceinline__ void draw_block(const Key& key, uint64_t iter, uint32_t attempt, uint32_t site,
                                           uint32_t seq, uint64_t& a, uint64_t& b) {
  uint32_t o[4];
  philox4x32_10((uint32_t)iter, ((uint32_t)(iter >> 32) & 0xFFFFu) | (attempt << 16), key.stream,
                (site << 28) | (seq & 0x0FFFFFFFu), key.k0, key.k1, o);
  a = (uint64_t)o[0] | ((uint64_t)o[1] << 32);
  b = (uint64_t)o[2] | ((uint64_t)o[3] << 32);
}

// Out-of-line copy for everything outside the speculative window (code size: the inlined Philox is
// ~120 instructions per call site and the kernel is instruction-fetch sensitive).
__device__ __noinline__ void draw_block_nl(Key key, uint64_t iter, uint32_t attempt, uint32_t site, uint32_t seq,
                                           uint64_t& a, uint64_t& b) {
  draw_block(key, iter, attempt, site, seq, a, b);
}

// Rejection branch of Lemire's bounded draw (probability s / 2^64 per draw): out of line.
__device__ __noinline__ uint32_t bounded_retry(Key key, uint64_t lo, uint64_t hi, uint32_t s, uint64_t iter,
                                               uint32_t site, uint32_t seq, int half) {
  uint64_t t = (0ull - (uint64_t)s) % (uint64_t)s;
  uint32_t attempt = 0;
  while (lo < t) {
    ++attempt;
    uint64_t a, b;
    draw_block_nl(key, iter, attempt, site, seq, a, b);
    uint64_t w = half ? b : a;
    hi = __umul64hi(w, (uint64_t)s);
    lo = w * (uint64_t)s;
  }
  return (uint32_t)hi;
}

// Lemire's nearly-divisionless bounded draw; `w` is attempt 0 of (iter, site, seq, half).
__device__ __forceinline__ uint32_t bounded(const Key& key, uint64_t w, uint32_t s, uint64_t iter, uint32_t site,
                                            uint32_t seq, int half) {
  uint64_t hi = __umul64hi(w, (uint64_t)s);
  uint64_t lo = w * (uint64_t)s;
  if (lo < (uint64_t)s) return bounded_retry(key, lo, hi, s, iter, site, seq, half);
  return (uint32_t)hi;
}

__device__ __forceinline__ double u01_53(uint64_t w) { return (double)(w >> 11) * 0x1.0p-53; }
__device__ __forceinline__ double u01_open(uint64_t w) { return ((double)(w >> 12) + 0.5) * 0x1.0p-52; }

// Natural log for positive normal doubles, the single-path fdlibm/musl form; the association
// below is part of the draw contract.
__device__ __forceinline__ double contract_log(double x) {
  const double ln2_hi = 0x1.62e42fee00000p-1, ln2_lo = 0x1.a39ef35793c76p-33, Lg1 = 0x1.5555555555593p-1,
               Lg2 = 0x1.999999997fa04p-2, Lg3 = 0x1.2492494229359p-2, Lg4 = 0x1.c71c51d8e78afp-3,
               Lg5 = 0x1.7466496cb03dep-3, Lg6 = 0x1.39a09d078c69fp-3, Lg7 = 0x1.2f112df3e5244p-3;
  uint32_t hx = (uint32_t)__double2hiint(x);
  uint32_t lx = (uint32_t)__double2loint(x);
  hx += 0x3ff00000u - 0x3fe6a09eu;
  int k = (int)(hx >> 20) - 0x3ff;
  hx = (hx & 0x000fffffu) + 0x3fe6a09eu;
  x = __hiloint2double((int)hx, (int)lx);
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

// ---------------------------------------------------------------------------------------------
// Explicit shared-memory accesses for the hot reads.  The per-warp slices are addressed by 32-bit
// shared-window offsets kept in registers; going through generic pointers made ptxas re-derive the
// window base (S2UR SR_CgaCtaId + LDC + IMAD) in front of every access.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long lds_u64(uint32_t a) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
// global load that asks L1 to keep the line (the per-replicate fitness table is re-read after every event)
__device__ __forceinline__ double ldg_keep_f64(const double* p) {
  double v;
  asm volatile("ld.global.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void lds_v2f64(uint32_t a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts_v2f64(uint32_t a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
// Sequential left fold acc <- fl(acc + v[k]), k = 0..count-1, over doubles in shared memory at the
// 16-byte aligned address `sa`; with STORE the running value replaces v[k] (written by `writer`
// lanes only; every lane runs the same chain).  A dependent fp64 add issues every ~8 cycles on
// sm_100 and a shared-memory load takes ~30, so the loop is double-buffered in groups of eight:
// the next group is requested before the current one is added (measured 10.7 cycles per element
// against 14-28 for simpler loop shapes, scripts/micro/fold2.cu).  The prefetch address is clamped
// to the last group instead of being predicated, which keeps the body free of branches.
#define JSB_LOAD8(v, o)                                                                            \
  {                                                                                                \
    lds_v2f64((o), v[0], v[1]); lds_v2f64((o) + 16u, v[2], v[3]);                                  \
    lds_v2f64((o) + 32u, v[4], v[5]); lds_v2f64((o) + 48u, v[6], v[7]);                            \
  }
#define JSB_CHAIN8(v, o)                                                                           \
  {                                                                                                \
    const double p0 = acc + v[0]; const double p1 = p0 + v[1]; if (STORE && writer) sts_v2f64((o), p0, p1);        \
    const double p2 = p1 + v[2]; const double p3 = p2 + v[3]; if (STORE && writer) sts_v2f64((o) + 16u, p2, p3);   \
    const double p4 = p3 + v[4]; const double p5 = p4 + v[5]; if (STORE && writer) sts_v2f64((o) + 32u, p4, p5);   \
    const double p6 = p5 + v[6]; acc = p6 + v[7]; if (STORE && writer) sts_v2f64((o) + 48u, p6, acc);              \
  }
template <bool STORE>
__device__ __forceinline__ double fold_chain(uint32_t sa, uint32_t count, double acc, bool writer) {
  const uint32_t ngrp = count >> 3;
  if (ngrp) {
    const uint32_t last_g = sa + (ngrp - 1u) * 64u;
    uint32_t o = sa;
    double a[8], b[8];
    JSB_LOAD8(a, o);
    for (uint32_t g = 0;;) {
      { const uint32_t on = min(o + 64u, last_g); JSB_LOAD8(b, on); }
      JSB_CHAIN8(a, o);
      if (++g == ngrp) break;
      o += 64u;
      { const uint32_t on = min(o + 64u, last_g); JSB_LOAD8(a, on); }
      JSB_CHAIN8(b, o);
      if (++g == ngrp) break;
      o += 64u;
    }
  }
  for (uint32_t k = ngrp << 3; k < count; ++k) {
    acc = acc + lds_f64(sa + k * 8u);
    if (STORE && writer) sts_f64(sa + k * 8u, acc);
  }
  return acc;
}
#undef JSB_LOAD8
#undef JSB_CHAIN8

template <typename T>
__device__ __forceinline__ uint32_t lds_tab(uint32_t base, uint32_t i);
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
  return v;
}
// entry i of a per-clone table of site-typed values
template <>
__device__ __forceinline__ uint32_t lds_tab<uint16_t>(uint32_t base, uint32_t i) { return lds_u16(base + i * 2u); }
template <>
__device__ __forceinline__ uint32_t lds_tab<uint32_t>(uint32_t base, uint32_t i) { return lds_u32(base + i * 4u); }

// ---------------------------------------------------------------------------------------------
// Graph access (0-based vertices on the device; the ABI adds 1)
// ---------------------------------------------------------------------------------------------
// Implicit lattice of Graphs.grid((n1,n2,n3)): vertex = x + y*n1 + z*n1*n2, neighbours ascending:
// z-1, y-1, x-1, x+1, y+1, z+1.
__device__ __forceinline__ uint32_t fast_div(uint32_t v, uint32_t magic, uint32_t shift) {
  return (uint32_t)(((unsigned long long)v * magic) >> shift);  // exact for v < 2^28 (jsb_api.cu)
}
__device__ __forceinline__ uint32_t lattice_mask(const DevGraph& g, uint32_t v) {
  uint32_t yz = fast_div(v, g.m1, g.s1);
  uint32_t x = v - yz * g.n1;
  uint32_t z = fast_div(yz, g.m2, g.s2);
  uint32_t y = yz - z * g.n2;
  return (uint32_t)(z > 0) | ((uint32_t)(y > 0) << 1) | ((uint32_t)(x > 0) << 2) | ((uint32_t)(x + 1 < g.n1) << 3) |
         ((uint32_t)(y + 1 < g.n2) << 4) | ((uint32_t)(z + 1 < g.n3) << 5);
}
__device__ __forceinline__ uint32_t lattice_pick(const DevGraph& g, uint32_t v, uint32_t mask, uint32_t j) {
  // the j-th (0-based) set bit of mask (at most six bits are set; j < popc(mask))
  uint32_t m = mask;
#pragma unroll
  for (uint32_t k = 0; k < 5u; ++k)
    if (k < j) m &= m - 1u;
  const int bit = __ffs(m) - 1;
  // bits 0..5 = -plane, -n1, -1, +1, +n1, +plane
  const int dist = bit < 3 ? 2 - bit : bit - 3;  // 0: x, 1: y, 2: z
  const int step = dist == 0 ? 1 : (dist == 1 ? (int)g.n1 : (int)(g.n1 * g.n2));
  return (uint32_t)((int)v + (bit < 3 ? -step : step));
}

// ---------------------------------------------------------------------------------------------
// Per-replicate context.  T = storage type of the ordered lists (uint16_t when nv <= 65535).
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Ctx {
  const RunArgs* A;
  const DevPoint* P;
  const double *tb, *ratio, *tsnap, *node_rate;
  const int32_t *edge_parent, *edge_child;
  uint16_t* clone;
  uint32_t* cellid;
  T *Alist, *arena;
  uint16_t* a_head;         // tiered-vector segment heads of cs_alive
  uint16_t* l_head;         // tiered-vector segment heads of the ca_subpop blocks, indexed by arena offset / kSeg
  T* lseg;                  // per vertex: the segment of its clone's list that holds it (position hint, see list_remove_value)
  uint32_t *inds, *samp;
  T *c_off8, *c_len;        // per-warp shared slices [kTab]: arena offset / 8 and length of every clone's list,
                            // stored in the site type (u16 lattices: 4 KB instead of 8 KB per warp)
  __device__ __forceinline__ uint32_t off(uint32_t i) const { return (uint32_t)c_off8[i] << 3; }
  __device__ __forceinline__ void set_off(uint32_t i, uint32_t v) {
    if (v & 7u) __trap();  // blocks start on multiples of 8 entries (capacities are; compaction rounds to 8)
    c_off8[i] = (T)(v >> 3);
  }
  uint32_t* c_cap;          // global; capacities are powers of two >= kMinListCap
  double* c_alpha;
  uint16_t *c_parent, *c_label;
  double *g_w, *g_c;        // global-memory scratch of the exact prob_cum fallback
  double* B;                // per-warp shared slice: class boundaries [kSmemTable]
  uint32_t* occ;            // per-warp shared occupancy bitmask (nullptr: test clone[] in global)
  uint32_t sB, sLen, sOff, sOcc;  // the same slices as 32-bit shared-window addresses (sOcc = 0: none)
  uint32_t sDt;             // 32 doubles: per-lane time increments of the current window
  uint32_t search_top;      // power of two >= S+2: first probe distance of the class search
  Key key;
  int64_t local_idx;
  int lane;
  bool track_ids;           // cell ids are only maintained when the log or the lattice is wanted
  uint8_t* ws_base;         // this warp slot's workspace (suspend / resume copies)
};

struct Rep {
  double t, theta, lambda, guard;
  uint64_t it, nlog;
  uint64_t nstored;     // rows in the log pool when it ran out (log_on false); otherwise nlog rows are stored
  uint32_t n, S, next_id, arena_top, cur_chunk, used_w;
  uint32_t dirty_from;  // first clone index whose size changed since the last refold (kNone = clean)
  int bott_id, snap_idx, status;
  uint32_t n_eff, n_births, n_deaths, n_migr, n_disp, n_exc;
  bool log_on, exact_valid;
};

static_assert(sizeof(Rep) <= 256, "susp_save parks Rep in a 256-byte block");

__device__ __forceinline__ void mark_dirty(Rep& r, uint32_t clone_idx) {
  r.dirty_from = clone_idx < r.dirty_from ? clone_idx : r.dirty_from;
}

// shared-memory occupancy bitmask (bit set <=> clone[site] != 0); single-lane updates
__device__ __forceinline__ void occ_set(uint32_t* occ, uint32_t site) { if (occ) occ[site >> 5] |= 1u << (site & 31); }
__device__ __forceinline__ void occ_clear(uint32_t* occ, uint32_t site) { if (occ) occ[site >> 5] &= ~(1u << (site & 31)); }

// ---------------------------------------------------------------------------------------------
// Event log: 32-byte records in pool chunks, rows in the reference's push! order
// ---------------------------------------------------------------------------------------------
// 16-byte streaming copy by one warp (workspace parking, packing a finished log)
__device__ __forceinline__ void warp_copy16(uint8_t* dst, const uint8_t* src, uint64_t bytes, int lane) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  const uint64_t n = (bytes + 15) >> 4;
  uint64_t i = (uint64_t)lane;
  for (; i + 96 < n; i += 128) {
    const uint4 a = __ldcs(s + i), b = __ldcs(s + i + 32), c = __ldcs(s + i + 64), e = __ldcs(s + i + 96);
    __stcs(d + i, a); __stcs(d + i + 32, b); __stcs(d + i + 64, c); __stcs(d + i + 96, e);
  }
  for (; i < n; i += 32) __stcs(d + i, __ldcs(s + i));
}
template <typename T>
__device__ __noinline__ void log_open_chunk(Ctx<T>& X, Rep& r) {
  uint32_t c = 0;
  if (X.lane == 0) {
    c = atomicAdd(X.A->chunk_next, 1u);
    if (c < X.A->n_chunks) {
      X.A->chunk_owner[c] = (int32_t)X.local_idx;
      X.A->chunk_seq[c] = (uint32_t)(r.nlog / kLogChunk);
      if (X.A->chunk_prev) X.A->chunk_prev[c] = r.nlog ? r.cur_chunk : kNone;
    }
  }
  c = __shfl_sync(FULL, c, 0);
  if (c >= X.A->n_chunks) {
    r.log_on = false;
    r.nstored = r.nlog;
    r.status = JSB_ST_LOG_OVERFLOW;
    return;
  }
  r.cur_chunk = c;
}

__device__ __forceinline__ void log_store(jsb_event* dst, int half, double time, uint32_t cell, uint32_t parent,
                                          uint32_t site1, uint32_t site2, uint32_t clone, uint32_t label,
                                          uint32_t type, uint32_t flags) {
  // two 16-byte stores = one full 32-byte sector
  uint4 v;
  if (half == 0) {
    v.x = (uint32_t)__double2loint(time);
    v.y = (uint32_t)__double2hiint(time);
    v.z = cell;
    v.w = parent;
  } else {
    v.x = site1;
    v.y = site2;
    v.z = (clone & 0xFFFFu) | (label << 16);
    v.w = (type & 0xFFu) | ((flags & 0xFFu) << 8);
  }
  reinterpret_cast<uint4*>(dst)[half] = v;
}

// Out-of-line row writer (only reached when the log is requested); returns the chunk in use or
// kNone when the pool is exhausted.
__device__ __noinline__ uint32_t log_emit(const RunArgs* A, int64_t local_idx, int lane, uint64_t nlog, uint32_t cur_chunk,
                                          uint32_t type, uint32_t flags, double time, uint32_t cell, uint32_t parent,
                                          uint32_t site1, uint32_t site2, uint32_t clone, uint32_t label) {
  const uint32_t slot = (uint32_t)(nlog % kLogChunk);
  if (slot == 0) {
    uint32_t c = 0;
    if (lane == 0) {
      c = atomicAdd(A->chunk_next, 1u);
      if (c < A->n_chunks) {
        A->chunk_owner[c] = (int32_t)local_idx;
        A->chunk_seq[c] = (uint32_t)(nlog / kLogChunk);
        if (A->chunk_prev) A->chunk_prev[c] = nlog ? cur_chunk : kNone;
      }
    }
    c = __shfl_sync(FULL, c, 0);
    if (c >= A->n_chunks) return kNone;
    cur_chunk = c;
  }
  if (lane < 2) {
    jsb_event* dst = A->log_pool + (size_t)cur_chunk * kLogChunk + slot;
    log_store(dst, lane, time, cell, parent, site1, site2, clone, label, type, flags);
  }
  return cur_chunk;
}

template <typename T>
__device__ __forceinline__ void log_row(Ctx<T>& X, Rep& r, uint32_t type, uint32_t flags, double time, uint32_t cell,
                                        uint32_t parent, uint32_t site0, uint32_t site2_1, uint32_t clone,
                                        uint32_t label) {
  if (r.log_on) {
    const uint32_t slot = (uint32_t)(r.nlog % kLogChunk);
    if (slot != 0) {  // inside the current chunk: two 16-byte stores, no call
      if (X.lane < 2) {
        jsb_event* dst = X.A->log_pool + (size_t)r.cur_chunk * kLogChunk + slot;
        log_store(dst, X.lane, time, cell, parent, site0 + 1, site2_1, clone, label, type, flags);
      }
    } else {          // first record of a chunk: claim one from the pool
      const uint32_t c = log_emit(X.A, X.local_idx, X.lane, r.nlog, r.cur_chunk, type, flags, time, cell, parent,
                                  site0 + 1, site2_1, clone, label);
      if (c == kNone) { r.log_on = false; r.nstored = r.nlog; r.status = JSB_ST_LOG_OVERFLOW; }
      else r.cur_chunk = c;
    }
  }
  r.nlog++;
}

// ---------------------------------------------------------------------------------------------
// Ordered lists.  cs_alive is one array; the ca_subpop lists live in an arena as contiguous
// blocks (offset, len, cap) so that `list[x]` is a single load.  filter!(e -> e != v, list) is a
// vectorised search + left shift (stable): 16-byte loads, several in flight per lane.  Blocks
// start on 16-byte boundaries and have capacities that are multiples of 16 bytes.
// ---------------------------------------------------------------------------------------------
template <typename T> struct Vec;
template <> struct Vec<uint32_t> {
  static constexpr uint32_t EPV = 4;
  static __device__ __forceinline__ uint4 shift1(uint4 v, uint32_t nx) { return make_uint4(v.y, v.z, v.w, nx); }
  static __device__ __forceinline__ uint32_t match(uint4 v, uint32_t value) {  // bitmask over the 4 elements
    return (uint32_t)(v.x == value) | ((uint32_t)(v.y == value) << 1) | ((uint32_t)(v.z == value) << 2) |
           ((uint32_t)(v.w == value) << 3);
  }
};
template <> struct Vec<uint16_t> {
  static constexpr uint32_t EPV = 8;
  static __device__ __forceinline__ uint4 shift1(uint4 v, uint32_t nx) {
    return make_uint4(__funnelshift_r(v.x, v.y, 16), __funnelshift_r(v.y, v.z, 16), __funnelshift_r(v.z, v.w, 16),
                      __funnelshift_r(v.w, nx, 16));
  }
  static __device__ __forceinline__ uint32_t match(uint4 v, uint32_t value) {
    uint32_t m = 0;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      m |= (uint32_t)((w[i] & 0xFFFFu) == value) << (2 * i);
      m |= (uint32_t)((w[i] >> 16) == value) << (2 * i + 1);
    }
    return m;
  }
};

// index of `value` in list[0..len) (unique), kNone when absent
template <typename T>
__device__ __noinline__ uint32_t list_find(const T* list, uint32_t len, uint32_t value, int lane) {
  constexpr uint32_t EPV = Vec<T>::EPV;
  constexpr int U = 4;
  const uint4* lv = reinterpret_cast<const uint4*>(list);
  const uint32_t ngroups = (len + EPV - 1) / EPV;
  for (uint32_t gb = 0; gb < ngroups; gb += 32 * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint32_t g = gb + u * 32 + lane;
      v[u] = make_uint4(kNone, kNone, kNone, kNone);
      if (g < ngroups) v[u] = lv[g];
    }
    uint32_t hit = kNone;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint32_t g = gb + u * 32 + lane;
      uint32_t m = Vec<T>::match(v[u], value);
      if (g < ngroups && m) {
        // the lowest matching element; a stale equal value beyond len can only sit above it
        uint32_t e = g * EPV + (uint32_t)__ffs(m) - 1u;
        if (e < len) hit = e;
      }
    }
    uint32_t bal = __ballot_sync(FULL, hit != kNone);
    if (bal) return __shfl_sync(FULL, hit, __ffs(bal) - 1);
  }
  return kNone;
}

// remove list[idx], keep order.  May read up to 16 bytes past list[len) (allocation slack).
template <typename T>
__device__ __noinline__ void list_remove_at(T* list, uint32_t len, uint32_t idx, int lane) {
  constexpr uint32_t EPV = Vec<T>::EPV;
  constexpr int U = 4;  // 16-byte loads in flight per lane; measured on C2: 4 and 2 equal, 8 +2.5 %, 12 +5 %, 16 +13 % kernel time
  if (idx + 1 >= len) return;
  const uint32_t last_dst = len - 2;
  const uint32_t g0 = idx / EPV, g1 = last_dst / EPV;
  {  // head: destination elements of group g0 (at most EPV-1 of them)
    uint32_t head_end = (g0 + 1) * EPV < len - 1 ? (g0 + 1) * EPV : len - 1;
    uint32_t e = idx + (uint32_t)lane;
    T v = 0;
    bool act = e < head_end;
    if (act) v = list[e + 1];
    __syncwarp();
    if (act) list[e] = v;
    __syncwarp();
  }
  uint4* lv = reinterpret_cast<uint4*>(list);
  for (uint32_t gb = g0 + 1; gb <= g1; gb += 32 * U) {
    uint4 v[U];
    uint32_t nx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint32_t g = gb + u * 32 + lane;
      if (g <= g1) {
        v[u] = lv[g];
        nx[u] = (uint32_t)list[(g + 1) * EPV];
      }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint32_t g = gb + u * 32 + lane;
      if (g <= g1) lv[g] = Vec<T>::shift1(v[u], nx[u]);
    }
    __syncwarp();
  }
}

__device__ __forceinline__ uint32_t round_cap(uint32_t c) { return (c + 7u) & ~7u; }
__device__ __forceinline__ uint32_t gc_cap(uint32_t len) {  // power of two >= max(min, 2*len)
  uint32_t c = kMinListCap;
  while (c < len * 2) c <<= 1;
  return c;
}

// ---------------------------------------------------------------------------------------------
// cs_alive as a tiered vector (Goodrich & Kloss): fixed segments of kSeg elements, each a circular
// buffer with its own head; every segment but the last is full.  Element p lives in segment
// p / kSeg at offset (head + p % kSeg) mod kSeg, so `cs_alive[x]` stays O(1), and the stable delete
// filter!(e -> e != cell, cs_alive) moves at most kSeg elements inside one segment plus ONE element
// per following segment (its front becomes the previous segment's back) instead of shifting the
// whole tail: ~n/2 elements (40 KB at n = 20 000) -> a few hundred bytes.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ uint32_t a_slot(const Ctx<T>& X, uint32_t p) {
  const uint32_t s = p / kSeg;
  return s * kSeg + ((X.a_head[s] + (p & (kSeg - 1u))) & (kSeg - 1u));
}
template <typename T>
__device__ __forceinline__ uint32_t a_get(const Ctx<T>& X, uint32_t p) { return (uint32_t)X.Alist[a_slot(X, p)]; }

// push!(cs_alive, site) with n elements present (single lane)
template <typename T>
__device__ __forceinline__ void a_append(const Ctx<T>& X, uint32_t n, uint32_t site) {
  const uint32_t s = n / kSeg;
  if ((n & (kSeg - 1u)) == 0) X.a_head[s] = 0;
  X.Alist[s * kSeg + ((X.a_head[s] + (n & (kSeg - 1u))) & (kSeg - 1u))] = (T)site;
}

// remove logical element p of n, keep order (whole warp).  `hint` (may be null): per-vertex segment numbers,
// kept up to date for the one element per later segment that changes segment.
template <typename T>
__device__ __noinline__ void a_remove(T* A, uint16_t* head, uint32_t n, uint32_t p, int lane, T* hint) {
  const uint32_t s = p / kSeg, i = p & (kSeg - 1u);
  const uint32_t ls = (n - 1) / kSeg;                       // last segment
  const uint32_t cnt = (s == ls) ? ((n - 1) & (kSeg - 1u)) + 1u : kSeg;
  const uint32_t hs = head[s];
  T* seg = A + s * kSeg;
  // inside segment s: logical j <- j+1 for j in [i, cnt-1)
  for (uint32_t base = i; base + 1 < cnt; base += 32) {
    const uint32_t j = base + lane;
    T v = 0;
    const bool act = j + 1 < cnt;
    if (act) v = seg[(hs + j + 1) & (kSeg - 1u)];
    __syncwarp();
    if (act) seg[(hs + j) & (kSeg - 1u)] = v;
    __syncwarp();
  }
  // every later segment hands its front to the back of its predecessor and rotates by one
  uint32_t prev_head_carry = hs;  // head of segment t-1 before this call, for the first lane
  for (uint32_t base = s + 1; base <= ls; base += 32) {
    const uint32_t t = base + lane;
    const bool act = t <= ls;
    uint32_t ht = 0, hp = 0;
    T v = 0;
    if (act) {
      ht = head[t];
      hp = (lane == 0) ? prev_head_carry : (uint32_t)head[t - 1];
      v = A[t * kSeg + ht];
    }
    const uint32_t last_ht = __shfl_sync(FULL, ht, 31);
    __syncwarp();
    if (act) {
      // back slot of t-1: segment s kept its head (slot head+kSeg-1); later ones rotated, so their
      // new back is their old front slot
      const uint32_t slot = (t - 1 == s) ? ((hs + kSeg - 1u) & (kSeg - 1u)) : hp;
      A[(t - 1) * kSeg + slot] = v;
      head[t] = (uint16_t)((ht + 1u) & (kSeg - 1u));
      if (hint) hint[v] = (T)(t - 1u);
    }
    prev_head_carry = last_ht;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// The ca_subpop lists, tiered as well.  A list of at most kSeg elements is a flat array (any 16-byte
// aligned place of the arena).  A longer one lives in a block of capacity >= 2*kSeg that starts on a
// multiple of kSeg entries of the arena, and is a tiered vector whose segment heads sit in
// l_head[arena offset / kSeg + segment]: `ca_subpop[i][x]` (:763) stays O(1) — one extra load of the
// head — and filter!(e -> e != v, ca_subpop[i]) (:739,755; :411) moves at most kSeg elements plus one
// per following segment instead of shifting the tail of an 80 KB list.  Invariant: whenever a list has
// len <= kSeg its elements are stored flat from the block start (a list that shrinks to kSeg elements
// is rotated back once), so small lists never look at a head.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ uint32_t l_get(const Ctx<T>& X, uint32_t off, uint32_t len, uint32_t x) {
  if (len <= kSeg) return (uint32_t)X.arena[off + x];
  const uint32_t s = x / kSeg;
  const uint32_t h = X.l_head[off / kSeg + s];
  return (uint32_t)X.arena[off + s * kSeg + ((h + (x & (kSeg - 1u))) & (kSeg - 1u))];
}

// logical position of `value` in a tiered list of len > kSeg elements (unique), kNone when absent.
// Scans the raw storage of the segments in use; in the last (partial) segment a slot only counts when
// it is one of the cnt live ones (a stale copy of the same vertex may sit in a dead slot).
template <typename T>
__device__ __noinline__ uint32_t tlist_find(const T* list, const uint16_t* head, uint32_t len, uint32_t value, int lane) {
  constexpr uint32_t EPV = Vec<T>::EPV;
  constexpr int U = 4;
  const uint4* lv = reinterpret_cast<const uint4*>(list);
  const uint32_t ls = (len - 1u) / kSeg;                      // last segment
  const uint32_t cnt_last = ((len - 1u) & (kSeg - 1u)) + 1u;  // live elements in it
  const uint32_t ngroups = (ls + 1u) * (kSeg / EPV);
  for (uint32_t gb = 0; gb < ngroups; gb += 32 * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint32_t g = gb + u * 32 + lane;
      v[u] = make_uint4(kNone, kNone, kNone, kNone);
      if (g < ngroups) v[u] = lv[g];
    }
    uint32_t hit = kNone;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint32_t g = gb + u * 32 + lane;
      uint32_t m = g < ngroups ? Vec<T>::match(v[u], value) : 0u;
      while (m) {
        const uint32_t e = g * EPV + (uint32_t)__ffs(m) - 1u;  // raw element index
        m &= m - 1u;
        const uint32_t sg = e / kSeg;
        const uint32_t j = (e - (uint32_t)head[sg]) & (kSeg - 1u);  // logical offset inside the segment
        if (sg < ls || j < cnt_last) hit = sg * kSeg + j;
      }
    }
    uint32_t bal = __ballot_sync(FULL, hit != kNone);
    if (bal) return __shfl_sync(FULL, hit, __ffs(bal) - 1);
  }
  return kNone;
}

// the same inside ONE segment (raw storage `seg`, head h, cnt live elements): logical offset or kNone
template <typename T>
__device__ __noinline__ uint32_t tseg_find(const T* seg, uint32_t h, uint32_t cnt, uint32_t value, int lane) {
  constexpr uint32_t EPV = Vec<T>::EPV;
  constexpr uint32_t NG = kSeg / EPV;  // 16-byte groups per segment: 32 (u16) or 64 (u32)
  const uint4* lv = reinterpret_cast<const uint4*>(seg);
  uint32_t hit = kNone;
#pragma unroll
  for (uint32_t u = 0; u < NG / 32u; ++u) {
    const uint32_t g = u * 32u + (uint32_t)lane;
    uint32_t m = Vec<T>::match(lv[g], value);
    while (m) {
      const uint32_t e = g * EPV + (uint32_t)__ffs(m) - 1u;
      m &= m - 1u;
      const uint32_t j = (e - h) & (kSeg - 1u);
      if (j < cnt) hit = j;
    }
  }
  const uint32_t bal = __ballot_sync(FULL, hit != kNone);
  return bal ? __shfl_sync(FULL, hit, __ffs(bal) - 1) : kNone;
}

// rotate the first segment of a block so that its head is 0 (the list just shrank to kSeg elements)
template <typename T>
__device__ __noinline__ void seg_normalise(T* seg, uint16_t* head, int lane) {
  const uint32_t h = head[0];
  __syncwarp();
  if (h != 0) {
    T v[kSeg / 32];
#pragma unroll
    for (uint32_t k = 0; k < kSeg / 32; ++k) v[k] = seg[(h + (uint32_t)lane + 32u * k) & (kSeg - 1u)];
    __syncwarp();
#pragma unroll
    for (uint32_t k = 0; k < kSeg / 32; ++k) seg[(uint32_t)lane + 32u * k] = v[k];
    if (lane == 0) head[0] = 0;
  }
  __syncwarp();
}

__device__ __forceinline__ uint32_t align_seg(uint32_t o) { return (o + (kSeg - 1u)) & ~(kSeg - 1u); }

// Arena compaction, in place, order inside every list untouched.  Pass 1 squeezes the blocks down in
// ascending-offset order, each stored flat (cap = len rounded to 16 bytes; a tiered list goes through
// the scratch array, its segments being rotations); the final offsets follow from one sequential walk
// in that order (block i gets cap = max(min, 2*len) rounded up to a power of two, blocks of 2*kSeg and
// more start on a multiple of kSeg); pass 2 spreads the blocks out again from the last to the first.
// Afterwards every list is flat from its block start, so every head is 0.  Visited blocks are tagged in
// the top bit of cap.  O(S^2/32) bookkeeping — only runs when clone turnover has fragmented the arena.
template <typename T>
__device__ __noinline__ void arena_gc(Ctx<T>& X, Rep& r) {
  const int lane = X.lane;
  // scratch: the tables of the exact prob_cum fallback (2*(kTab+8) words each; invalidated below) hold the
  // visiting order of pass 1 and the final offsets — the clone count is not bounded by the lattice size
  uint32_t* order = reinterpret_cast<uint32_t*>(X.g_w);
  uint32_t* fin = reinterpret_cast<uint32_t*>(X.g_c);
  uint32_t* lin = X.inds;    // a tiered list passes through here (len <= nv)
  r.exact_valid = false;
  uint32_t top = 0;
  for (uint32_t step = 0; step < r.S; ++step) {
    uint32_t best = kNone, besti = kNone;
    for (uint32_t i = lane; i < r.S; i += 32) {
      uint32_t o = X.off(i);
      if (!(X.c_cap[i] & 0x80000000u) && o < best) { best = o; besti = i; }
    }
    for (int d = 16; d > 0; d >>= 1) {
      uint32_t ob = __shfl_xor_sync(FULL, best, d), oi = __shfl_xor_sync(FULL, besti, d);
      if (ob < best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    const uint32_t i = besti, l = X.c_len[i], src = best;
    if (l > kSeg) {  // tiered -> flat through the scratch array
      for (uint32_t k = lane; k < l; k += 32) lin[k] = l_get(X, src, l, k);
      __syncwarp();
      for (uint32_t k = lane; k < l; k += 32) X.arena[top + k] = (T)lin[k];
      __syncwarp();
    } else if (top != src) {
      for (uint32_t b = 0; b < l; b += 32) {  // dst < src: ascending chunks are safe
        uint32_t k = b + lane;
        T v = 0;
        if (k < l) v = X.arena[src + k];
        __syncwarp();
        if (k < l) X.arena[top + k] = v;
        __syncwarp();
      }
    }
    if (lane == 0) {
      X.set_off(i, top);
      X.c_cap[i] = round_cap(l) | 0x80000000u;
      order[step] = i;
    }
    __syncwarp();
    top += round_cap(l);
  }
  // final offsets, in the same order
  uint32_t total = 0;
  if (lane == 0) {
    for (uint32_t k = 0; k < r.S; ++k) {
      const uint32_t cap = gc_cap(X.c_len[order[k]]);
      if (cap >= 2u * kSeg) total = align_seg(total);
      fin[k] = total;
      total += cap;
    }
  }
  total = __shfl_sync(FULL, total, 0);
  __syncwarp();
  if (total > X.A->arena_cap) {  // cannot happen when arena_cap >= 6*nv + kMinListCap*kTab
    for (uint32_t i = lane; i < r.S; i += 32) X.c_cap[i] &= 0x7FFFFFFFu;
    __syncwarp();
    r.arena_top = top;
    r.status = JSB_ST_INTERNAL;
    return;
  }
  for (uint32_t k = r.S; k-- > 0;) {
    const uint32_t i = order[k], l = X.c_len[i], src = X.off(i), noff = fin[k];
    if (noff != src && l > 0) {
      for (uint32_t cb = (l + 31) / 32; cb-- > 0;) {  // dst > src: descending chunks are safe
        uint32_t q = cb * 32 + lane;
        T v = 0;
        if (q < l) v = X.arena[src + q];
        __syncwarp();
        if (q < l) X.arena[noff + q] = v;
        __syncwarp();
      }
    }
    if (lane == 0) {
      X.set_off(i, noff);
      X.c_cap[i] = gc_cap(l);
    }
    __syncwarp();
  }
  for (uint32_t q = lane; q < X.A->arena_cap / kSeg + 2u; q += 32) X.l_head[q] = 0;
  __syncwarp();
  r.arena_top = total;
}

// slow path of list_append: list c is full
template <typename T>
__device__ __noinline__ bool list_grow(Ctx<T>& X, Rep& r, uint32_t c) {
  uint32_t len = X.c_len[c], cap = X.c_cap[c];
  uint32_t ncap = cap * 2;
  uint32_t dst = ncap >= 2u * kSeg ? align_seg(r.arena_top) : r.arena_top;
  if (dst + ncap > X.A->arena_cap) {
    arena_gc(X, r);
    if (r.status == JSB_ST_INTERNAL) return false;
    len = X.c_len[c];
    cap = X.c_cap[c];
    if (len < cap) return true;
    ncap = cap * 2;
    dst = ncap >= 2u * kSeg ? align_seg(r.arena_top) : r.arena_top;
    if (dst + ncap > X.A->arena_cap) { r.status = JSB_ST_INTERNAL; return false; }
  }
  // the list is full (len == cap): a tiered block is copied segment for segment with its heads, a flat one
  // as it is (and, when the new block is a tiered one, becomes its first segment with head 0)
  const uint32_t src = X.off(c);
  for (uint32_t b = 0; b < len; b += 32) {
    uint32_t k = b + X.lane;
    if (k < len) X.arena[dst + k] = X.arena[src + k];
  }
  if (ncap >= 2u * kSeg) {
    const uint32_t nseg_old = cap >= 2u * kSeg ? cap / kSeg : 0u;
    for (uint32_t q = X.lane; q < ncap / kSeg; q += 32)
      X.l_head[dst / kSeg + q] = q < nseg_old ? X.l_head[src / kSeg + q] : (uint16_t)0;
  }
  __syncwarp();
  if (X.lane == 0) {
    X.set_off(c, dst);
    X.c_cap[c] = ncap;
  }
  __syncwarp();
  r.arena_top = dst + ncap;
  return true;
}

template <typename T>
__device__ __forceinline__ bool list_append(Ctx<T>& X, Rep& r, uint32_t c, uint32_t site) {
  {
    // capacities are powers of two >= kMinListCap, so a list can only be full when its length is
    // one; only then is the capacity (global memory) looked at
    const uint32_t len0 = X.c_len[c];
    if (len0 >= (uint32_t)kMinListCap && (len0 & (len0 - 1)) == 0 && len0 >= X.c_cap[c]) {
      Ctx<T> xc = X; Rep rc = r;
      const bool ok = list_grow(xc, rc, c);
      r = rc;
      if (!ok) return false;
    }
  }
  if (X.lane == 0) {
    const uint32_t len = X.c_len[c], off = X.off(c);
    if (len < kSeg) {
      X.arena[off + len] = (T)site;
    } else {  // the block holds >= 2*kSeg entries and starts on a multiple of kSeg
      const uint32_t sg = len / kSeg, hi = off / kSeg + sg;
      if ((len & (kSeg - 1u)) == 0) X.l_head[hi] = 0;
      X.arena[off + sg * kSeg + ((X.l_head[hi] + (len & (kSeg - 1u))) & (kSeg - 1u))] = (T)site;
    }
    X.lseg[site] = (T)(len / kSeg);
    X.c_len[c] = (T)(len + 1);
  }
  __syncwarp();
  return true;
}

// filter!(e -> e != site, ca_subpop[c]).  The reference scans the list for the vertex; here lseg[site]
// names the segment that holds it (every operation that moves an element to another segment keeps it up to
// date: appends, the one-per-segment hand-over of a tiered delete, the excision rebuild), so only kSeg
// entries are searched.  A wrong hint would still be found by the full scan, and is reported as an
// internal error so that the tests see it.
template <typename T>
__device__ __forceinline__ void list_remove_value(Ctx<T>& X, Rep& r, uint32_t c, uint32_t site) {
  const uint32_t len = X.c_len[c], off = X.off(c);
  T* list = X.arena + off;
  if (len <= kSeg) {
    uint32_t idx = list_find<T>(list, len, site, X.lane);
    if (idx != kNone) {
      list_remove_at<T>(list, len, idx, X.lane);
      if (X.lane == 0) X.c_len[c] = (T)(len - 1);
      __syncwarp();
    }
  } else {
    uint16_t* head = X.l_head + off / kSeg;
    const uint32_t ls = (len - 1u) / kSeg;
    uint32_t sg = (uint32_t)X.lseg[site];
    uint32_t p = kNone;
    if (sg <= ls) {
      const uint32_t cnt = sg == ls ? ((len - 1u) & (kSeg - 1u)) + 1u : kSeg;
      const uint32_t j = tseg_find<T>(list + sg * kSeg, head[sg], cnt, site, X.lane);
      if (j != kNone) p = sg * kSeg + j;
    }
    if (p == kNone) {
      p = tlist_find<T>(list, head, len, site, X.lane);
      if (p != kNone) r.status = JSB_ST_INTERNAL;  // stale hint
    }
    if (p != kNone) {
      a_remove<T>(list, head, len, p, X.lane, X.lseg);
      if (X.lane == 0) X.c_len[c] = (T)(len - 1);
      __syncwarp();
      if (len - 1u == kSeg) seg_normalise<T>(list, head, X.lane);
    }
  }
}

template <typename T>
__device__ __noinline__ bool new_clone(Ctx<T>& X, Rep& r, uint32_t parent1, uint32_t label, double alpha,
                                       uint32_t first_site) {
  const uint32_t c = r.S;
  if (r.arena_top + kMinListCap > X.A->arena_cap) {
    arena_gc(X, r);
    if (r.status == JSB_ST_INTERNAL) return false;
    if (r.arena_top + kMinListCap > X.A->arena_cap) { r.status = JSB_ST_INTERNAL; return false; }
  }
  if (X.lane == 0) {
    X.set_off(c, r.arena_top);
    X.c_len[c] = (T)(1);
    X.c_cap[c] = kMinListCap;
    X.c_alpha[c] = alpha;
    X.c_parent[c] = (uint16_t)parent1;
    X.c_label[c] = (uint16_t)label;
    X.arena[r.arena_top] = (T)first_site;
    X.lseg[first_site] = 0;
  }
  __syncwarp();
  r.arena_top += kMinListCap;
  r.S = c + 1;
  mark_dirty(r, c);
  return true;
}

// ---------------------------------------------------------------------------------------------
// Class boundaries: B_i = fl(B_{i-1} + α_i·n_i) (the reference's sequential `sum`, :692-695),
// B_S = fl(birth + death), B_{S+1} = λ (:696-698).  Re-folded from the first changed clone.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void refold(Ctx<T>& X, Rep& r) {
  const uint32_t S = r.S;
  const uint32_t from = r.dirty_from < S ? r.dirty_from : S;
  const int lane = X.lane;
  double* __restrict__ B = X.B;
  const double* __restrict__ alpha = X.c_alpha;
  // phase A: B[i] <- w_i = α_i·n_i for every i >= from.  α lives in global memory (L2 latency with
  // the shared-memory carve-out this kernel uses), so sixteen loads per lane are put in flight at once and the lines are asked to stay in L1.
  const uint32_t sB = X.sB, sLen = X.sLen;
  SUB_T0();
  if (S - from <= 32u) {
    // few clones to re-weigh (every event of a sweep with a tiny driver rate, the early phase of every replicate):
    // one load per lane instead of the sixteen-deep pipeline below
    const uint32_t i = from + (uint32_t)lane;
    if (i < S) sts_f64(sB + i * 8u, ldg_keep_f64(alpha + i) * (double)lds_tab<T>(sLen, i));  // :693
  } else
  for (uint32_t base = from; base < S; base += 512u) {
    double a[16];
#pragma unroll
    for (uint32_t k = 0; k < 16u; ++k) {
      const uint32_t i = base + k * 32u + (uint32_t)lane;
      a[k] = 0.0;
      if (i < S) a[k] = ldg_keep_f64(alpha + i);
    }
#pragma unroll
    for (uint32_t k = 0; k < 16u; ++k) {
      const uint32_t i = base + k * 32u + (uint32_t)lane;
      if (i < S) sts_f64(sB + i * 8u, a[k] * (double)lds_tab<T>(sLen, i));  // :693
    }
  }
  __syncwarp();
  SUB_ADD(0);
  // phase B: in-place sequential left fold (the reference's `sum`).  Every lane runs the chain
  // (the total is needed warp-wide); lane 0 alone writes the prefixes back.  The warp is converged
  // here, so a lane never reads an entry lane 0 has already overwritten: loads of a group are
  // issued by the whole warp before lane 0's stores of that group.
  double acc = from ? lds_f64(sB + (from - 1u) * 8u) : 0.0;  // 0.0 + w_0 == w_0 exactly
  uint32_t i = from;
  const bool writer = lane == 0;
  if (i < S && (i & 1u)) {  // align the chain to 16 bytes
    acc = acc + lds_f64(sB + i * 8u);
    if (writer) sts_f64(sB + i * 8u, acc);
    ++i;
  }
  if (i < S) acc = fold_chain<true>(sB + i * 8u, S - i, acc, writer);
  __syncwarp();
  SUB_ADD(1);
  const double death = X.P->rate_death * (double)r.n;     // :696
  const double M = X.P->rate_migration * (double)r.n;     // :697
  const double bd = acc + death;
  const double lambda = bd + M;                           // :698
  if (lane == 0) {
    B[S] = bd;
    B[S + 1] = lambda;
  }
  {
    uint32_t top = X.search_top;
    while (top < S + 2) top <<= 1;
    X.search_top = top;
  }
  r.lambda = lambda;
  r.theta = 1.0 / lambda;                                 // Exponential(1 / λ) :699
  r.guard = ((X.P->quirks & JSB_QUIRK_DEBUG_WIDE_GUARD) ? 1e-3 : kGuardEps) * lambda;
  r.exact_valid = false;
  r.dirty_from = kNone;
  __syncwarp();
  SUB_ADD(2);
}

// Exact prob_cum = cumsum(vcat(α_subpop ./ λ, death/λ, M/λ)) (:703-718) in Julia's
// accumulate_pairwise! order, stored as its running maximum so that a lower bound equals
// findfirst(k .<= prob_cum).  Only built when some k falls inside the guard band.
// Depth-bounded form of the recursion (n <= 1001 needs at most 3 levels of halving before every
// block is shorter than 128), so the stack size is known at compile time.
template <int DEPTH>
__device__ __forceinline__ double acc_pairwise(double* __restrict__ v, double* __restrict__ c, double s, int i1, int n,
                                               double& mx, int lane) {
  if (DEPTH == 0 || n < 128) {
    double s_ = v[i1];
    double cv = s + s_;
    mx = cv > mx ? cv : mx;
    if (lane == 0) c[i1] = mx;
    for (int i = i1 + 1; i < i1 + n; ++i) {
      s_ = s_ + v[i];
      cv = s + s_;
      mx = cv > mx ? cv : mx;
      if (lane == 0) c[i] = mx;
    }
    return s_;
  }
  int n2 = n >> 1;
  double sl = acc_pairwise<(DEPTH > 0 ? DEPTH - 1 : 0)>(v, c, s, i1, n2, mx, lane);
  double sr = acc_pairwise<(DEPTH > 0 ? DEPTH - 1 : 0)>(v, c, s + sl, i1 + n2, n - n2, mx, lane);
  return sl + sr;
}

template <typename T>
__device__ __noinline__ void build_exact_table(Ctx<T>& X, Rep& r) {
  const int S = (int)r.S;
  double* __restrict__ w = X.g_w;
  double* __restrict__ c = X.g_c;
  const double lambda = r.lambda;
  for (int i = X.lane; i < S; i += 32) w[i] = (X.c_alpha[i] * (double)X.c_len[i]) / lambda;  // Aₙ :703
  if (X.lane == 0) {
    w[S] = (X.P->rate_death * (double)r.n) / lambda;          // Bₙ :704
    w[S + 1] = (X.P->rate_migration * (double)r.n) / lambda;  // Mₙ :705
  }
  __syncwarp();
  double v1 = w[0];
  if (X.lane == 0) c[0] = v1;
  double mx = v1;
  static_assert(kMaxClones + 1 < (128 << 4), "pairwise depth");
  acc_pairwise<4>(w, c, v1, 1, S + 1, mx, X.lane);
  __syncwarp();
  r.exact_valid = true;
}

// ---------------------------------------------------------------------------------------------
// One speculative evaluation of iteration q against the current state (per lane)
// ---------------------------------------------------------------------------------------------
struct Eval {
  uint32_t cls;   // 0..S-1 birth of clone cls+1, S death, S+1 migration, S+2 none
  uint32_t x;     // index into the member list
  uint32_t cell;  // 0-based vertex
  uint32_t pos;   // 0-based vertex
  uint32_t occ;   // clone id at pos (0 = empty)
  bool effective;
};

// Class search: first i with !(B[i] < key).  The table is padded with +inf up to kSmemTable, all
// values are non-negative doubles, so the IEEE bit patterns order like the numbers and the search
// is a branch-free sequence of integer compares on ld.shared (an fp64 compare costs a full fp64
// pipe round trip on this part).  `top` = power of two >= S+2.
__device__ __forceinline__ uint32_t class_search(uint32_t sB, uint32_t top, double key) {
  const unsigned long long k = (unsigned long long)__double_as_longlong(key);
  uint32_t lo = 0;  // number of leading entries known to be < key
  for (uint32_t step = top; step > 0; step >>= 1) {
    const unsigned long long v = lds_u64(sB + (lo + step - 1u) * 8u);
    lo = v < k ? lo + step : lo;
  }
  return lo;
}

// same search on a global-memory table of n entries (exact prob_cum fallback; rare)
__device__ __forceinline__ uint32_t lower_bound(const double* tab, uint32_t n, double key) {
  const unsigned long long* t = reinterpret_cast<const unsigned long long*>(tab);
  const unsigned long long k = (unsigned long long)__double_as_longlong(key);
  uint32_t lo = 0;
#pragma unroll
  for (uint32_t step = 512; step > 0; step >>= 1) {
    uint32_t probe = lo + step;
    if (probe <= n && t[probe - 1] < k) lo = probe;
  }
  return lo;
}

// Phase 1 of a speculative evaluation: class -> member list -> x -> the dividing/dying/moving
// cell.  Written branch-light so that the sub-batches of one lane interleave (ILP).
template <typename T>
__device__ __forceinline__ void eval_member(const Ctx<T>& X, const Rep& r, uint64_t q, uint32_t cls, uint64_t w_cell,
                                            uint32_t& x, uint32_t& cell, bool& valid) {
  const uint32_t S = r.S;
  const bool is_birth = cls < S;
  const uint32_t ci = is_birth ? cls : 0u;
  const uint32_t off = lds_tab<T>(X.sOff, ci) << 3, cl = lds_tab<T>(X.sLen, ci);
  const uint32_t llen = is_birth ? cl : r.n;
  // cls >= S+2: findfirst -> nothing in the reference; llen == 0: rand(1:0) throws.  Contract:
  // both are null iterations.
  valid = (cls < S + 2) && (llen > 0);
  x = 0;
  cell = 0;
  if (valid) {
    x = bounded(X.key, w_cell, llen, q, DS_CELL_NBR, 0, 0);  // :727 / :746 / :762
    cell = is_birth ? l_get(X, off, cl, x) : a_get(X, x);
  }
}

// Phase 2: neighbour draw, occupancy test, contact rule.  Returns whether the iteration changes
// state.
template <typename T>
__device__ __forceinline__ bool eval_target(const Ctx<T>& X, const Rep& r, uint64_t q, uint32_t cls, uint32_t cell,
                                            bool valid, uint64_t w_nbr, uint32_t& pos) {
  const uint32_t S = r.S;
  pos = 0;
  if (!valid) return false;
  if (cls == S) return true;  // death :744-759 needs no neighbour
  const DevGraph& g = X.A->g;
  uint32_t deg, mask = 0, row0 = 0;
  if (g.kind == 0) { mask = lattice_mask(g, cell); deg = (uint32_t)__popc(mask); }
  else { row0 = g.rowptr[cell]; deg = g.rowptr[cell + 1] - row0; }
  if (deg == 0) return false;  // rand on an empty neighbour list throws; contract: null iteration
  const uint32_t j = bounded(X.key, w_nbr, deg, q, DS_CELL_NBR, 0, 1);  // :730 / :764
  pos = (g.kind == 0) ? lattice_pick(g, cell, mask, j) : g.colidx[row0 + j];
  const int model = X.P->rule;
  uint32_t occ;
  if (X.occ) {
    // `pos ∈ cs_alive` from the shared bitmask; the clone id itself is only needed when a voter
    // rule may displace the occupant
    const bool occupied = (lds_u32(X.sOcc + (pos >> 5) * 4u) >> (pos & 31)) & 1u;
    occ = occupied ? ((model == JSB_RULE_CONTACT || cls == S + 1) ? 1u : (uint32_t)X.clone[pos]) : 0u;
  } else {
    occ = X.clone[pos];
  }
  if (cls == S + 1) return occ == 0;  // migration :732
  bool rule = true;                   // :767-786
  if (model == JSB_RULE_CONTACT) rule = (occ == 0);
  else if (model == JSB_RULE_VOTER) rule = !(occ != 0 && occ == cls + 1);
  else if (model == JSB_RULE_HVOTER) rule = !(occ != 0 && X.c_alpha[cls] <= X.c_alpha[occ - 1]);
  return rule;
}

// One speculative window of U sub-batches: lane l of sub-batch u evaluates iteration
// it + 1 + 32u + l against the current state.  effm[u] = ballot of state-changing lanes.
// Measured on B200 (C2 ensemble): windows of 2-4 sub-batches per lane cut single-replicate latency
// by only ~8 % and cost 30-60 % throughput in wasted speculation (ptxas does not interleave the
// sub-batches; a 2-wide window used only after an all-null window was 25 % slower overall), so the
// window stays at one sub-batch; the multi-sub-batch path is kept for experiments.
constexpr int kMaxSub = 4;
constexpr bool kAdaptiveWindow = false;
template <typename T, int U>
__device__ __forceinline__ bool batch_eval(Ctx<T>& X, Rep& r, bool force_exact, uint32_t wmask, double (&ev)[kMaxSub],
                                           uint32_t (&effm)[kMaxSub], uint32_t (&cls)[kMaxSub], uint32_t (&xx)[kMaxSub],
                                           uint32_t (&cell)[kMaxSub], uint32_t (&pos)[kMaxSub], bool defer_exact) {
  const int lane = X.lane;
  const uint32_t nb = r.S + 2;
  uint64_t wc[U], wd[U];
  bool unc[U];
  bool unc_any = false;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint64_t q = r.it + 1 + (uint64_t)(u * 32 + lane);
    uint64_t wa, wb;
    draw_block(X.key, q, 0, DS_TIME_CLASS, 0, wa, wb);
    draw_block(X.key, q, 0, DS_CELL_NBR, 0, wc[u], wd[u]);
    ev[u] = -contract_log(u01_open(wa));  // the unit exponential of :699; the caller scales it by 1/λ
    // event class: k·λ against the folded boundaries
    const double key = u01_53(wb) * r.lambda;          // :721
    const uint32_t c = class_search(X.sB, X.search_top, key);
    const uint32_t cc = c < nb ? c : nb - 1;
    const double hi_b = lds_f64(X.sB + cc * 8u), lo_b = lds_f64(X.sB + (cc > 0 ? cc - 1 : 0) * 8u);
    unc[u] = force_exact || c >= nb || (hi_b - key <= r.guard) || (cc > 0 && key - lo_b <= r.guard);
    unc_any = unc_any || unc[u];
    cls[u] = c;
  }
  if (__ballot_sync(FULL, unc_any) & wmask) {  // some k is inside the guard band: exact prob_cum table
    if (defer_exact) return false;  // decoupled mode: the caller fetches the exact total rate first
    if (!r.exact_valid) { Ctx<T> xc = X; Rep rc = r; build_exact_table(xc, rc); r = rc; }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (unc[u]) {
        const uint64_t q = r.it + 1 + (uint64_t)(u * 32 + lane);
        uint64_t wa, wb;
        draw_block(X.key, q, 0, DS_TIME_CLASS, 0, wa, wb);
        cls[u] = lower_bound(X.g_c, nb, u01_53(wb));  // findfirst(k .<= prob_cum) :722-723
      }
    }
  }
  bool valid[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint64_t q = r.it + 1 + (uint64_t)(u * 32 + lane);
    eval_member(X, r, q, cls[u], wc[u], xx[u], cell[u], valid[u]);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint64_t q = r.it + 1 + (uint64_t)(u * 32 + lane);
    const bool eff = eval_target(X, r, q, cls[u], cell[u], valid[u], wd[u], pos[u]);
    effm[u] = __ballot_sync(FULL, eff) & wmask;
  }
  return true;
}

// sequential t += dt over lanes 0..last of one sub-batch (:701).  The increments are staged in
// shared memory so that the chain is nothing but dependent fp64 adds (broadcast loads, issued
// ahead); through shuffles every step also paid the shuffle latency.
__device__ __forceinline__ double time_chain_at(uint32_t sDt, double t, uint32_t last) {
  return fold_chain<false>(sDt, last + 1u, t, false);
}
__device__ __forceinline__ double time_chain(uint32_t sDt, int lane, double t, double dt, uint32_t last) {
  sts_f64(sDt + (uint32_t)lane * 8u, dt);
  __syncwarp();
  t = time_chain_at(sDt, t, last);
  __syncwarp();
  return t;
}

// ---------------------------------------------------------------------------------------------
// Event application (whole warp cooperates; control flow is warp-uniform)
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void apply_migration(Ctx<T>& X, Rep& r, const Eval& e) {  // :733-741, migration_cell :592-613
  uint32_t cid = X.track_ids ? X.cellid[e.cell] : 0u;
  uint32_t cl = X.clone[e.cell];
  __syncwarp();
  if (X.lane == 0) {
    X.clone[e.cell] = 0;
    X.clone[e.pos] = (uint16_t)cl;
    if (X.track_ids) X.cellid[e.pos] = cid;
    a_append(X, r.n, e.pos);  // push!(cs_alive, pos)
    occ_clear(X.occ, e.cell);
    occ_set(X.occ, e.pos);
  }
  __syncwarp();
  log_row(X, r, JSB_EV_MIGRATION, JSB_EVF_GROUP_START, r.t, cid, 0, e.pos, e.cell + 1, cl, 0);
  list_append(X, r, cl - 1, e.pos);                   // push!(ca_subpop[idx], pos)
  a_remove<T>(X.Alist, X.a_head, r.n + 1, e.x, X.lane, nullptr);  // filter!(e -> e != cell, cs_alive)
  list_remove_value(X, r, cl - 1, e.cell);               // filter!(e -> e != cell, ca_subpop[idx])
  r.n_eff++;
  r.n_migr++;
}

template <typename T>
__device__ __forceinline__ uint32_t apply_death(Ctx<T>& X, Rep& r, const Eval& e) {  // :746-759, cell_death :574-586; returns the clone
  uint32_t cid = X.track_ids ? X.cellid[e.cell] : 0u;
  uint32_t cl = X.clone[e.cell];
  __syncwarp();
  log_row(X, r, JSB_EV_DEATH, JSB_EVF_GROUP_START, r.t, cid, 0, e.cell, 0, cl, 0);
  if (X.lane == 0) { X.clone[e.cell] = 0; occ_clear(X.occ, e.cell); }
  __syncwarp();
  a_remove<T>(X.Alist, X.a_head, r.n, e.x, X.lane, nullptr);
  list_remove_value(X, r, cl - 1, e.cell);
  r.n -= 1;
  r.n_eff++;
  r.n_deaths++;
  mark_dirty(r, cl - 1);
  return cl;
}

// j-th (0-based) unused label of 1..1000; lane l owns labels 32l..32l+31
__device__ __noinline__ uint32_t pick_label(uint32_t used_w, uint32_t j, int lane) {
  uint32_t freew = ~used_w;
  uint32_t cnt = (uint32_t)__popc(freew);
  uint32_t incl = cnt;
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(FULL, incl, d);
    if (lane >= d) incl += o;
  }
  uint32_t excl = incl - cnt;
  bool mine = (j >= excl) && (j < incl);
  uint32_t lab = 0;
  if (mine) {
    uint32_t k = j - excl;  // k-th set bit of freew
    uint32_t m = freew;
    for (uint32_t q = 0; q < k; ++q) m &= m - 1;
    uint32_t bit = (uint32_t)__ffs(m) - 1u;
    lab = (uint32_t)lane * 32u + bit;
  }
  uint32_t who = __ballot_sync(FULL, mine);
  return __shfl_sync(FULL, lab, __ffs(who) - 1);
}

// randn by the polar method on contract draws (D9)
__device__ __noinline__ double contract_randn(Key key, uint64_t iter) {
  for (uint32_t attempt = 0;; ++attempt) {
    uint64_t a, b;
    draw_block_nl(key, iter, 0, DS_NORMAL, attempt, a, b);
    double v1 = 2.0 * u01_53(a) - 1.0;
    double v2 = 2.0 * u01_53(b) - 1.0;
    double s = v1 * v1 + v2 * v2;
    if (s >= 1.0 || s == 0.0) continue;
    return v1 * sqrt((-2.0 * contract_log(s)) / s);
  }
}

// method 2: first child of the genotype's last driver, in edge-list order, whose genotype is not
// in set_mut yet (:450-473); -1 when none
template <typename T>
__device__ __noinline__ int tree_child(const Ctx<T>& X, const Rep& r, uint32_t sp) {
  const DevPoint& P = *X.P;
  int child_node = -1;
  int last = (int)X.c_label[sp - 1] - 1;
  for (int ed = 0; ed < P.n_edges && child_node < 0; ++ed) {
    if (X.edge_parent[ed] != last) continue;
    int cand = X.edge_child[ed];
    bool present = false;  // new_mut ∈ set_mut :466
    for (uint32_t s0 = X.lane; s0 < r.S; s0 += 32)
      if (X.c_parent[s0] == sp && (int)X.c_label[s0] - 1 == cand) present = true;
    present = __any_sync(FULL, present);
    if (!present) child_node = cand;
  }
  return child_node;
}

// cell_birth :320-416 / :419-537 plus the caller's bookkeeping :803-808.  Returns false when the
// replicate must stop.
template <typename T>
__device__ __forceinline__ bool apply_birth(Ctx<T>& X, Rep& r, const Eval& e) {
  const DevPoint& P = *X.P;
  const uint32_t sp = e.cls + 1;  // :Subpop of the dividing cell (== min)
  uint64_t wa, wb;
  draw_block_nl(X.key, r.it, 0, DS_DRIVER_COIN, 0, wa, wb);
  const double rr = u01_53(wa);  // :347 / :447
  int child_node = -1;
  bool driver;
  if (P.method == JSB_METHOD_TREE) {
    if (rr <= P.mu) { Ctx<T> xc = X; Rep rc = r; child_node = tree_child(xc, rc, sp); }  // :448
    driver = child_node >= 0;                            // :476
  } else {
    driver = !(rr > P.mu);  // :348
  }
  if (driver && r.S >= (uint32_t)kMaxClones) {  // reference throws at :371
    r.status = JSB_ST_CLONE_CAPACITY;
    return false;
  }
  uint32_t first_flag = JSB_EVF_GROUP_START;
  if (e.occ != 0) {  // :335-338 displacement (voter rules); e.occ was read by the caller
    uint32_t pid = X.track_ids ? X.cellid[e.pos] : 0u;
    __syncwarp();
    list_remove_value(X, r, e.occ - 1, e.pos);
    mark_dirty(r, e.occ - 1);
    log_row(X, r, JSB_EV_DEATH, JSB_EVF_DISPLACED | first_flag, r.t, pid, 0, e.pos, 0, e.occ, 0);
    first_flag = 0;
    r.n_disp++;
  }
  const uint32_t nid = r.next_id, nid2 = r.next_id + 1;  // uuid1 ×2 :340-341
  r.next_id += 2;
  const uint32_t parent = X.track_ids ? X.cellid[e.cell] : 0u;
  __syncwarp();
  if (!driver) {  // :351-363 / :477-490
    log_row(X, r, JSB_EV_DUPLICATE, first_flag, r.t, nid, parent, e.cell, 0, sp, 0);
    log_row(X, r, JSB_EV_DUPLICATE, 0, r.t, nid2, parent, e.pos, 0, sp, 0);
    if (X.lane == 0) {
      if (X.track_ids) { X.cellid[e.cell] = nid; X.cellid[e.pos] = nid2; }
      X.clone[e.pos] = (uint16_t)sp;
    }
    __syncwarp();
    if (!list_append(X, r, sp - 1, e.pos)) return false;
    mark_dirty(r, sp - 1);
  } else {
    uint32_t label;
    double nalpha;
    if (P.method == JSB_METHOD_TREE) {
      label = (uint32_t)child_node + 1;
      nalpha = X.node_rate[child_node];  // :503-505
    } else {
      uint32_t n_free = 1000u - r.S;  // one label per clone, the founder holds label 1
      uint64_t a, b;
      draw_block_nl(X.key, r.it, 0, DS_LABEL, 0, a, b);
      uint32_t j = bounded(X.key, a, n_free, r.it, DS_LABEL, 0, 0);  // :371
      label = pick_label(r.used_w, j, X.lane);
      if ((uint32_t)X.lane == (label >> 5)) r.used_w |= 1u << (label & 31u);
      double z = contract_randn(X.key, r.it);  // :565-566
      nalpha = X.c_alpha[sp - 1] + fabs(P.adv_mean + P.adv_std * z);
    }
    const uint32_t nsp = r.S + 1;
    const bool coin = u01_53(wb) < 0.5;  // :391 / :512
    log_row(X, r, JSB_EV_MUTATION, first_flag, r.t, nid2, parent, coin ? e.pos : e.cell, 0, nsp, label);
    log_row(X, r, JSB_EV_DUPLICATE, 0, r.t, nid, parent, coin ? e.cell : e.pos, 0, sp, 0);
    if (coin) {  // :392-400
      if (X.lane == 0) {
        if (X.track_ids) { X.cellid[e.cell] = nid; X.cellid[e.pos] = nid2; }
        X.clone[e.pos] = (uint16_t)nsp;
      }
      __syncwarp();
      { Ctx<T> xc = X; Rep rc = r; const bool ok = new_clone(xc, rc, sp, label, nalpha, e.pos); r = rc; if (!ok) return false; }
    } else {  // :402-412
      if (X.lane == 0) {
        if (X.track_ids) { X.cellid[e.pos] = nid; X.cellid[e.cell] = nid2; }
        X.clone[e.pos] = (uint16_t)sp;
        X.clone[e.cell] = (uint16_t)nsp;
      }
      __syncwarp();
      if (!list_append(X, r, sp - 1, e.pos)) return false;
      list_remove_value(X, r, sp - 1, e.cell);
      { Ctx<T> xc = X; Rep rc = r; const bool ok = new_clone(xc, rc, sp, label, nalpha, e.cell); r = rc; if (!ok) return false; }
    }
  }
  if (e.occ == 0) {  // :803-806
    if (X.lane == 0) { a_append(X, r.n, e.pos); occ_set(X.occ, e.pos); }
    __syncwarp();
    r.n += 1;
    mark_dirty(r, r.S);  // death and migration weights changed (B_S, B_{S+1})
  }
  r.n_eff++;
  r.n_births++;
  return true;
}

// Excision block :813-838 / :1036-1061: StatsBase.sample(rng, cs_alive, k; replace=false), then
// cell_death for each sampled vertex in sample order, logged at time tb.
template <typename T>
__device__ __noinline__ void apply_excision(Ctx<T>& X, Rep& r, double tb, double ratio) {
  const uint32_t n = r.n;
  const double taked = (double)n * ratio;  // :819
  long long kk = (long long)trunc(taked);
  if (kk < 0) kk = 0;
  if (kk > (long long)n) kk = (long long)n;  // reference: sample() throws
  const uint32_t k = (uint32_t)kk;
  uint32_t* inds = X.inds;
  uint32_t* samp = X.samp;
  const int lane = X.lane;
  if (k == 1) {
    uint64_t a, b;
    draw_block_nl(X.key, r.it, 0, DS_EXCISION, 0, a, b);
    uint32_t i = bounded(X.key, a, n, r.it, DS_EXCISION, 0, 0);
    if (lane == 0) samp[0] = a_get(X, i);
  } else if (k == 2) {  // samplepair
    uint64_t a, b;
    draw_block_nl(X.key, r.it, 0, DS_EXCISION, 0, a, b);
    uint32_t i1 = bounded(X.key, a, n, r.it, DS_EXCISION, 0, 0) + 1;
    draw_block_nl(X.key, r.it, 0, DS_EXCISION, 1, a, b);
    uint32_t i2 = bounded(X.key, a, n - 1, r.it, DS_EXCISION, 1, 0) + 1;
    if (lane == 0) {
      samp[0] = a_get(X, i1 - 1);
      samp[1] = a_get(X, (i2 == i1 ? n : i2) - 1);
    }
  } else if (k > 2 && (unsigned long long)n < (unsigned long long)k * 24ull) {  // fisher_yates_sample!
    for (uint32_t i = lane; i < n; i += 32) inds[i] = i;
    __syncwarp();
    for (uint32_t base = 0; base < k; base += 32) {
      uint32_t i = base + lane;
      uint32_t jmine = 0;
      if (i < k) {
        uint64_t a, b;
        draw_block_nl(X.key, r.it, 0, DS_EXCISION, i, a, b);
        jmine = i + bounded(X.key, a, n - i, r.it, DS_EXCISION, i, 0);  // rand(i:n)
      }
      uint32_t cnt = (k - base) < 32u ? (k - base) : 32u;
      for (uint32_t l = 0; l < cnt; ++l) {  // the swaps are sequential
        uint32_t j = __shfl_sync(FULL, jmine, (int)l);
        if (lane == 0) {
          uint32_t ii = base + l;
          uint32_t tj = inds[j], ti = inds[ii];
          inds[j] = ti;
          inds[ii] = tj;
          samp[ii] = a_get(X, tj);
        }
      }
      __syncwarp();
    }
  } else if (k > 2) {  // self_avoid_sample!
    for (uint32_t i = lane; i < n; i += 32) inds[i] = 0;
    __syncwarp();
    if (lane == 0) {
      uint32_t q = 0;
      for (uint32_t i = 0; i < k; ++i) {
        uint32_t idx;
        for (;;) {
          uint64_t a, b;
          draw_block_nl(X.key, r.it, 0, DS_EXCISION, q, a, b);
          idx = bounded(X.key, a, n, r.it, DS_EXCISION, q, 0);
          ++q;
          if (!inds[idx]) break;
        }
        inds[idx] = 1;
        samp[i] = a_get(X, idx);
      }
    }
  }
  __syncwarp();
  // deaths in sample order; rows written in parallel inside each log chunk
  uint32_t done = 0;
  while (done < k) {
    uint32_t m = k - done;
    if (r.log_on) {
      uint32_t slot = (uint32_t)(r.nlog % kLogChunk);
      if (slot == 0) log_open_chunk(X, r);
      uint32_t room = kLogChunk - slot;
      if (m > room) m = room;
    }
    for (uint32_t b = 0; b < m; b += 32) {
      uint32_t i = b + lane;
      if (i < m) {
        uint32_t site = samp[done + i];
        uint32_t cid = X.track_ids ? X.cellid[site] : 0u;
        uint32_t cl = X.clone[site];
        if (r.log_on) {
          jsb_event* dst = X.A->log_pool + (size_t)r.cur_chunk * kLogChunk + (uint32_t)(r.nlog % kLogChunk) + i;
          uint32_t fl = JSB_EVF_EXCISION | ((done + i) == 0 ? JSB_EVF_GROUP_START : 0);
          log_store(dst, 0, tb, cid, 0, site + 1, 0, cl, 0, JSB_EV_DEATH, fl);
          log_store(dst, 1, tb, cid, 0, site + 1, 0, cl, 0, JSB_EV_DEATH, fl);
        }
      }
    }
    r.nlog += m;
    done += m;
  }
  __syncwarp();
  for (uint32_t i = lane; i < k; i += 32) X.clone[samp[i]] = 0;
  __syncwarp();
  if (X.occ && lane == 0)
    for (uint32_t i = 0; i < k; ++i) occ_clear(X.occ, samp[i]);
  __syncwarp();
  // stable compaction of cs_alive and of every ca_subpop list (== the sequence of filter! calls).
  // cs_alive is rebuilt as a fresh tiered vector (all heads 0) through the scratch array.
  {
    uint32_t wpos = 0;
    for (uint32_t base = 0; base < n; base += 32) {
      uint32_t i = base + lane;
      uint32_t v = 0;
      bool keep = false;
      if (i < n) { v = a_get(X, i); keep = X.clone[v] != 0; }
      uint32_t m = __ballot_sync(FULL, keep);
      if (keep) inds[wpos + __popc(m & ((1u << lane) - 1u))] = v;
      wpos += __popc(m);
    }
    __syncwarp();
    for (uint32_t i = lane; i < wpos; i += 32) X.Alist[i] = (T)inds[i];
    for (uint32_t sg = lane; sg * kSeg <= wpos; sg += 32) X.a_head[sg] = 0;
    __syncwarp();
  }
  for (uint32_t c = 0; c < r.S; ++c) {
    const uint32_t len = X.c_len[c], off = X.off(c);
    T* list = X.arena + off;
    uint32_t wpos = 0;
    if (len <= kSeg) {  // flat: in place (a kept element never moves up)
      for (uint32_t base = 0; base < len; base += 32) {
        uint32_t i = base + lane;
        T v = 0;
        bool keep = false;
        if (i < len) { v = list[i]; keep = X.clone[v] != 0; }
        uint32_t m = __ballot_sync(FULL, keep);
        __syncwarp();
        if (keep) list[wpos + __popc(m & ((1u << lane) - 1u))] = v;
        wpos += __popc(m);
        __syncwarp();
      }
    } else {            // tiered: through the scratch array, rebuilt flat with every head 0
      for (uint32_t base = 0; base < len; base += 32) {
        uint32_t i = base + lane;
        uint32_t v = 0;
        bool keep = false;
        if (i < len) { v = l_get(X, off, len, i); keep = X.clone[v] != 0; }
        uint32_t m = __ballot_sync(FULL, keep);
        if (keep) inds[wpos + __popc(m & ((1u << lane) - 1u))] = v;
        wpos += __popc(m);
      }
      __syncwarp();
      for (uint32_t i = lane; i < wpos; i += 32) { list[i] = (T)inds[i]; X.lseg[inds[i]] = (T)(i / kSeg); }
      for (uint32_t sg = lane; sg <= (len - 1u) / kSeg; sg += 32) X.l_head[off / kSeg + sg] = 0;
      __syncwarp();
    }
    if (lane == 0) X.c_len[c] = (T)(wpos);
  }
  __syncwarp();
  r.n = n - k;
  r.n_exc += k;
  r.n_eff++;  // one series push :835-836
  r.dirty_from = 0;
}

// ---------------------------------------------------------------------------------------------
// Seeding: initialize_graph :79-163 and the initial lists :655-683
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __noinline__ void seed_replicate(Ctx<T>& X, Rep& r) {
  const DevPoint& P = *X.P;
  const DevGraph& g = X.A->g;
  const int lane = X.lane;
  for (uint32_t i = lane; i < (g.nv + 1) / 2; i += 32) reinterpret_cast<uint32_t*>(X.clone)[i] = 0;
  if (X.occ)
    for (uint32_t i = lane; i < X.A->occ_words; i += 32) X.occ[i] = 0;
  for (uint32_t i = lane; i < (uint32_t)kSmemTable; i += 32) X.B[i] = __longlong_as_double(0x7ff0000000000000ll);  // +inf pad
  for (uint32_t sg = lane; sg * kSeg <= (uint32_t)(P.n_start > 1 ? P.n_start : 1); sg += 32) X.a_head[sg] = 0;
  for (uint32_t q = lane; q < X.A->arena_cap / kSeg + 2u; q += 32) X.l_head[q] = 0;
  __syncwarp();
  uint32_t n0 = 0;
  if (P.n_start == 1) {
    if (X.A->n_cand > 0) {
      uint32_t j = 0;
      if (X.A->n_cand > 1) {
        uint64_t a, b;
        draw_block_nl(X.key, 0, 0, DS_SEED_TIE, 0, a, b);
        j = bounded(X.key, a, X.A->n_cand, 0, DS_SEED_TIE, 0, 0);  // :86
      }
      uint32_t v0 = X.A->seed_cand[j];
      if (lane == 0) {
        X.clone[v0] = 1;
        X.cellid[v0] = 1;
        X.Alist[0] = (T)v0;
        X.lseg[v0] = 0;
        occ_set(X.occ, v0);
      }
      n0 = 1;
    }
  } else if (P.n_start > 1) {
    // :105-112 / :151-157: the i-th cell takes the j-th smallest remaining vertex; ids in draw
    // order, lists in ascending vertex order.  samp[] holds the chosen vertices sorted.
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int i = 0; i < P.n_start; ++i) {
        uint64_t a, b;
        draw_block_nl(X.key, 0, 0, DS_SEED_POS, (uint32_t)i, a, b);
        uint32_t pos = bounded(X.key, a, g.nv - (uint32_t)i, 0, DS_SEED_POS, (uint32_t)i, 0);  // 0-based rank
        uint32_t ins = 0;
        for (uint32_t q = 0; q < cnt; ++q)
          if (X.samp[q] <= pos) { ++pos; ins = q + 1; }
        for (uint32_t q = cnt; q > ins; --q) X.samp[q] = X.samp[q - 1];
        X.samp[ins] = pos;
        ++cnt;
        X.clone[pos] = 1;
        X.cellid[pos] = (uint32_t)i + 1;
        occ_set(X.occ, pos);
      }
      for (uint32_t q = 0; q < cnt; ++q) { X.Alist[q] = (T)X.samp[q]; X.lseg[X.samp[q]] = (T)(q / kSeg); }
    }
    n0 = (uint32_t)P.n_start;
  }
  __syncwarp();
  r.n = n0;
  r.next_id = n0 + 1;
  r.S = 0;
  r.arena_top = 0;
  r.used_w = 0;
  if (lane == 0) r.used_w = 0x3u;               // label 0 does not exist, label 1 is the founder
  if (lane == 31) r.used_w = 0xFFFFFE00u;       // labels 1001..1023 do not exist
  if (n0 > 0) {
    uint32_t cap = kMinListCap;
    while (cap < n0) cap *= 2;
    if (lane == 0) {
      X.set_off(0, 0);
      X.c_len[0] = (T)(n0);
      X.c_cap[0] = cap;
      X.c_alpha[0] = (P.method == JSB_METHOD_TREE) ? P.root_rate : P.rate_birth;  // :664 / :881
      X.c_parent[0] = 0;
      X.c_label[0] = (P.method == JSB_METHOD_TREE) ? (uint16_t)(P.root_node + 1) : (uint16_t)1;
    }
    for (uint32_t i = lane; i < n0; i += 32) X.arena[i] = X.Alist[i];
    r.arena_top = cap;
    r.S = 1;
  }
  __syncwarp();
}

template <typename T>
__device__ __noinline__ void write_outputs(Ctx<T>& X, Rep& r) {
  const RunArgs& A = *X.A;
  const int lane = X.lane;
  const uint32_t nv = A.g.nv;
  __syncwarp();
  if (lane == 0) {
    jsb_summary s;
    s.n_iterations = r.it;
    s.n_events = r.nlog;
    s.t_final = r.t;
    s.n_effective = r.n_eff;
    s.n_births = r.n_births;
    s.n_deaths = r.n_deaths;
    s.n_migrations = r.n_migr;
    s.n_displaced = r.n_disp;
    s.n_excised = r.n_exc;
    s.n_alive = r.n;
    s.n_clones = r.S;
    s.next_cell_id = r.next_id;
    s.status = r.status;
    A.summaries[X.local_idx] = s;
  }
  // streamed outputs: finish rank and first log row, one atomic; lattice and lists go to slot `rank` so that the
  // host can copy completed runs of ranks as contiguous blocks
  unsigned long long fin_stored = 0, fin_old = 0;
  if (A.fin_list) {
    if (A.flags & JSB_OUT_EVENTS) fin_stored = r.log_on ? r.nlog : r.nstored;
    if (lane == 0) fin_old = atomicAdd(A.fin_ctr, (1ull << 40) + fin_stored);
    fin_old = __shfl_sync(FULL, fin_old, 0);
  }
  const size_t out_slot = A.fin_list ? (size_t)(fin_old >> 40) : (size_t)X.local_idx;
  long long off = -1;
  if (A.clone_pool) {
    unsigned long long o = 0;
    if (lane == 0) o = atomicAdd(A.clone_pool_next, (unsigned long long)r.S);
    o = __shfl_sync(FULL, o, 0);
    if (o + r.S <= A.clone_pool_cap) {
      off = (long long)o;
      for (uint32_t i = lane; i < r.S; i += 32) {
        jsb_clone c;
        c.alpha = X.c_alpha[i];
        c.count = X.c_len[i];
        c.parent = X.c_parent[i];
        c.label = X.c_label[i];
        A.clone_pool[o + i] = c;
      }
    }
    if (lane == 0) A.clone_off[X.local_idx] = off;
  }
  if (A.out_clone) {
    uint16_t* oc = A.out_clone + out_slot * nv;
    uint32_t* oi = A.out_cell + out_slot * nv;
    for (uint32_t v = lane; v < nv; v += 32) {
      uint16_t cl = X.clone[v];
      oc[v] = cl;
      oi[v] = cl ? X.cellid[v] : 0u;
    }
  }
  if (A.out_lists) {
    uint32_t* la = A.out_lists + out_slot * 2 * nv;
    uint32_t* ll = la + nv;
    for (uint32_t i = lane; i < r.n; i += 32) la[i] = a_get(X, i) + 1;
    uint32_t wpos = 0;
    for (uint32_t c = 0; c < r.S; ++c) {
      const uint32_t len = X.c_len[c], off = X.off(c);
      for (uint32_t i = lane; i < len; i += 32) ll[wpos + i] = l_get(X, off, len, i) + 1;
      wpos += len;
    }
  }
  __syncwarp();
  if (A.fin_list) {
    const unsigned long long stored = fin_stored, old = fin_old;
    const unsigned long long first = old & ((1ull << 40) - 1ull);
    const uint32_t rank = (uint32_t)(old >> 40);
    bool packed = stored == 0;
    if (stored > 0 && A.gather_out && first + stored <= A.gather_cap) {
      // pack this replicate's chunks (last one first, following chunk_prev) into the output at its rows
      uint32_t c = r.cur_chunk;
      for (long long q = (long long)((stored - 1) / kLogChunk); q >= 0; --q) {
        const unsigned long long done = (unsigned long long)q * kLogChunk;
        const unsigned long long cnt = stored - done < (unsigned long long)kLogChunk ? stored - done : (unsigned long long)kLogChunk;
        warp_copy16(reinterpret_cast<uint8_t*>(A.gather_out + first + done),
                    reinterpret_cast<const uint8_t*>(A.log_pool + (size_t)c * kLogChunk), cnt * sizeof(jsb_event), lane);
        c = A.chunk_prev[c];
      }
      packed = true;
    }
    __threadfence();  // everything this replicate wrote is visible before its entry is
    __syncwarp();
    if (lane == 0)
      A.fin_list[rank] = make_uint4((uint32_t)X.local_idx | (packed ? 0u : 0x80000000u), (uint32_t)stored, (uint32_t)first,
                                    (uint32_t)(first >> 32));
  }
}

// Per-lane (t_old, t_curr) of the candidate iterations and everything that can happen besides
// the first state-changing event: loop exit at Tf, max_iterations, excision, snapshot.  Only
// reached when the batch's time span touches one of those (a handful of times per replicate).
struct SlowOut {
  uint32_t j;
  bool stop, maxit, do_event, do_snap, do_bott;
  double t_prev, t_cur, tb, ratio;
};

template <typename T>
__device__ __noinline__ SlowOut slow_scan(const Ctx<T>& X, const Rep& r, double dt, bool effective, uint32_t jE,
                                          uint32_t wmask, uint64_t q) {
  const DevPoint& P = *X.P;
  const int lane = X.lane;
  const bool intent = (P.quirks & JSB_QUIRK_EXCISION_INTENT) != 0;
  double my_prev = r.t, my_t = r.t, run = r.t;
  for (uint32_t l = 0; l <= jE; ++l) {
    double prev = run;
    run = run + __shfl_sync(FULL, dt, (int)l);  // :701
    if ((uint32_t)lane == l) { my_prev = prev; my_t = run; }
  }
  const bool in_range = (uint32_t)lane <= jE;
  const bool stop_l = in_range && !(my_prev < P.Tf);  // loop test :687 (n > 0 holds inside a batch)
  const bool maxit_l = in_range && P.max_iter && (r.it + (uint64_t)lane >= P.max_iter);
  bool bott_l = false;
  double tb_l = 0.0, ratio_l = 0.0;
  if (in_range && P.n_bott > 0) {
    int id;
    if (intent) id = r.bott_id;
    else if (P.method == JSB_METHOD_TREE) id = (int)(q < (uint64_t)P.n_bott ? q : (uint64_t)P.n_bott);  // :1062-1064
    else id = 1;  // method 1 never advances id_bottleneck
    if (id > 0) {
      tb_l = X.tb[id - 1];
      ratio_l = X.ratio[id - 1];
      bott_l = tb_l > my_prev && tb_l <= my_t;  // :814-815
    }
  }
  const bool snap_l = in_range && r.snap_idx < P.n_snap && my_t > X.tsnap[r.snap_idx];  // :708-710
  const bool special = stop_l || maxit_l || bott_l || snap_l || (in_range && effective);
  const uint32_t smask = __ballot_sync(FULL, special) & wmask;
  SlowOut o;
  if (smask == 0) {  // nothing special after all: commit the whole range
    o.j = jE; o.stop = o.maxit = o.do_event = o.do_snap = o.do_bott = false;
    o.t_prev = 0.0; o.t_cur = __shfl_sync(FULL, my_t, (int)jE); o.tb = o.ratio = 0.0;
    return o;
  }
  const int j = __ffs(smask) - 1;
  o.j = (uint32_t)j;
  o.stop = __shfl_sync(FULL, (int)stop_l, j) != 0;
  o.maxit = __shfl_sync(FULL, (int)maxit_l, j) != 0;
  o.do_event = __shfl_sync(FULL, (int)effective, j) != 0;
  o.do_snap = __shfl_sync(FULL, (int)snap_l, j) != 0;
  o.do_bott = __shfl_sync(FULL, (int)bott_l, j) != 0;
  o.t_prev = __shfl_sync(FULL, my_prev, j);
  o.t_cur = __shfl_sync(FULL, my_t, j);
  o.tb = __shfl_sync(FULL, tb_l, j);
  o.ratio = __shfl_sync(FULL, ratio_l, j);
  return o;
}

// ---------------------------------------------------------------------------------------------
// Warp-cooperative lookahead.  A warp that has no replicate left (the tail of an ensemble, or an
// ensemble smaller than the machine) attaches to a busy warp of its block; lookahead helper k
// evaluates the window of 32 iterations k+1 windows ahead of the owner's, against the same state,
// while the owner evaluates the current one.  The owner commits every window up to and including
// the first one with a state-changing iteration (up to 160 iterations per evaluation latency
// instead of 32).  The owner never changes any state while a request is outstanding, and the
// helpers only ever read; all sit in the same block, so the mailbox lives in shared memory.
// Results are identical by construction (same evaluation code, same state) — the parity tests run
// with helpers active.
// ---------------------------------------------------------------------------------------------
struct HelpRes {
  uint32_t ack;                             // id of the last request this helper finished
  uint32_t usable, effm, cls, x, cell, pos;  // first state-changing lane of the helper's window
  uint32_t pad;
  double dt[32];
  double sum_dt, pad2;  // dt[0] + ... + dt[first state-changing lane] (all 32 when there is none), any order
};
// Channel to the "clock" helper of a replicate (decoupled mode, see clock_serve).
struct ClockChan {
  uint32_t attached, warp;   // a clock helper serves this replicate; its warp slot (its shared slice holds the ring)
  uint32_t posted, done;     // sequence numbers: last message written by the owner / finished by the helper
  uint32_t claimed, pad[3];  // an idle warp has taken the clock role (attached follows once it is ready)
  double t_exact, lambda_exact;  // exact time and total rate after message `done`
  double t_init, lambda_init;    // payload of an INIT message
};
struct ClockMsg {
  unsigned long long it_first;  // the round committed iterations it_first+1 .. it_first+m
  uint32_t kind, m;             // 0 INIT, 1 ROUND
  uint32_t from, S, n;          // state after the round's event: fold from clone `from` (kNone: unchanged)
  uint32_t n_rows, n_upd, pad;
  uint32_t rows[3];             // log records of the event (chunk * kLogChunk + slot): their time is the round's end
  uint32_t upd_idx[3];          // clones whose size changed ...
  int32_t upd_kind[3];          // ... by +1 / -1, or 2 = new clone of size 1
  uint32_t pad2;
};
static_assert(sizeof(ClockMsg) == 80, "ring slots are 16-byte multiples");
constexpr uint32_t kClockRing = 8;
constexpr uint32_t kClockResync = 1024;  // events between two exact re-bases of the owner's approximate table
struct HelpMail {
  uint32_t word;          // (request id << 8) | helpers asked; written last by the owner
  uint32_t nhelp, done;   // helpers attached (only grows); the owner has no work left
  uint32_t S, n, search_top, stream, pad;
  unsigned long long it_base;  // lookahead helper k evaluates iterations it_base + 32(k+1) + 1 .. + 32
  double lambda, guard, theta;
  ClockChan clk;
  HelpRes res[kMaxLook];
};
static_assert(sizeof(HelpMail) % 16 == 0, "mailbox keeps the 16-byte alignment of the smem slices");
static_assert(offsetof(HelpMail, res) % 16 == 0 && sizeof(HelpRes) % 16 == 0 && offsetof(HelpRes, dt) % 16 == 0, "dt is read with 16-byte loads");

template <typename T>
__device__ __noinline__ void helper_serve(Ctx<T> X, volatile HelpMail* M, const uint32_t h) {
  const int lane = X.lane;
  volatile HelpRes* R = &M->res[h];
  for (;;) {
    uint32_t w;
    for (;;) {
      w = M->word;
      // a request is this helper's iff it was counted (h < helpers asked) and is not yet acknowledged
      if ((w & 0xFFu) > h && R->ack != (w >> 8)) break;
      if (M->done) return;
      __nanosleep(20);
    }
    __threadfence_block();
    Rep r;
    r.it = M->it_base + 32ull * (h + 1u); r.S = M->S; r.n = M->n;
    r.lambda = M->lambda; r.guard = M->guard; r.theta = M->theta;
    X.search_top = M->search_top;
    X.key.stream = M->stream;
    const uint64_t q = r.it + 1 + (uint64_t)lane;
    uint64_t wa, wb, wc, wd;
    draw_block(X.key, q, 0, DS_TIME_CLASS, 0, wa, wb);
    draw_block(X.key, q, 0, DS_CELL_NBR, 0, wc, wd);
    const double dt = (-contract_log(u01_open(wa))) * r.theta;
    const uint32_t nb = r.S + 2;
    const double key = u01_53(wb) * r.lambda;
    const uint32_t c = class_search(X.sB, X.search_top, key);
    const uint32_t cc = c < nb ? c : nb - 1;
    const double hi_b = lds_f64(X.sB + cc * 8u), lo_b = lds_f64(X.sB + (cc > 0 ? cc - 1 : 0) * 8u);
    const bool unc = c >= nb || (hi_b - key <= r.guard) || (cc > 0 && key - lo_b <= r.guard);
    const uint32_t anyunc = __ballot_sync(FULL, unc);
    uint32_t x = 0, cell = 0, pos = 0;
    bool valid = false, eff = false;
    if (!anyunc) {  // inside the guard band the owner decides with its exact table
      eval_member(X, r, q, c, wc, x, cell, valid);
      eff = eval_target(X, r, q, c, cell, valid, wd, pos);
    }
    const uint32_t effm = __ballot_sync(FULL, eff);
    R->dt[lane] = dt;
    {  // approximate span of the window up to its first event (decoupled mode only needs the sum)
      const int je = effm ? __ffs(effm) - 1 : 31;
      double sdt = lane <= je ? dt : 0.0;
      for (int d = 16; d > 0; d >>= 1) sdt += __shfl_xor_sync(FULL, sdt, d);
      if (lane == 0) R->sum_dt = sdt;
    }
    if (effm && lane == __ffs(effm) - 1) { R->cls = c; R->x = x; R->cell = cell; R->pos = pos; }
    if (lane == 0) { R->effm = effm; R->usable = anyunc ? 0u : 1u; }
    __syncwarp();
    __threadfence_block();
    if (lane == 0) R->ack = w >> 8;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Decoupled mode: the clock helper.
//
// Two sequential fp64 chains sit on a replicate's critical path after every event: the fold that
// gives the exact total rate λ (`sum`, :692-698) and the time chain t += e·(1/λ) (:699-701).
// Neither decides what happens next, except near Tf / an excision time / a snapshot time: the
// event class is decided by comparing k·λ with the class boundaries, and that decision is the
// same for any boundaries within the guard band of the exact ones (§4 of DESIGN.md; inside the
// band the exact prob_cum table decides).  So when an idle warp is available the owner keeps an
// APPROXIMATE table (each event's ±α added to the suffix in parallel, re-based on the exact one
// every kClockResync events) and an approximate clock, and streams what it did to the clock helper,
// which owns the exact table, replays the exact fold and time chain one round behind, and patches
// the exact event times into the log records.  Whenever the approximate clock comes within a
// relative 1e-9 of a time that matters, or a draw lands in the guard band, the owner waits for
// the helper, takes over its exact state and runs that round on the exact path.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned char* slice_of(const double* my_B, int my_warp, uint32_t smem_per_warp, uint32_t warp) {
  return reinterpret_cast<unsigned char*>(const_cast<double*>(my_B)) + ((long)warp - (long)my_warp) * (long)smem_per_warp;
}
template <typename T>
__device__ __forceinline__ volatile ClockMsg* ring_of(unsigned char* slice) {  // the helper's (idle) c_off8 region
  static_assert(sizeof(T) * kTab >= sizeof(ClockMsg) * kClockRing, "the ring fits the region");
  return reinterpret_cast<volatile ClockMsg*>(slice + sizeof(double) * (kSmemTable + 32) + sizeof(T) * kTab);
}

// XO: the owner's context; `mine`: this (idle) warp's own shared slice, reused as
// [ exact B table | dt staging | exact c_len copy | message ring ].
template <typename T>
__device__ __noinline__ void clock_serve(Ctx<T> XO, volatile HelpMail* M, unsigned char* mine) {
  const int lane = XO.lane;
  const RunArgs& A = *XO.A;
  volatile ClockChan* C = &M->clk;
  volatile ClockMsg* ring = ring_of<T>(mine);
  // the owner's context with the tables redirected to this warp's slice
  Ctx<T> XF = XO;
  XF.B = reinterpret_cast<double*>(mine);
  XF.c_len = reinterpret_cast<T*>(mine + sizeof(double) * (kSmemTable + 32));
  XF.sB = smem_addr(XF.B);
  XF.sDt = XF.sB + 8u * kSmemTable;
  XF.sLen = smem_addr(XF.c_len);
  const uint32_t soB = XO.sB, soLen = XO.sLen;
  double t = 0.0, lambda = 1.0, theta = 1.0;
  uint32_t seq = C->done;
  if (lane == 0) { C->warp = (uint32_t)(threadIdx.x >> 5); __threadfence_block(); C->attached = 1; }
  __syncwarp();
  for (;;) {
    while (C->posted == seq) {
      if (M->done) return;
      __nanosleep(20);
    }
    __threadfence_block();
    ++seq;
    volatile ClockMsg* m = &ring[seq % kClockRing];
    const uint32_t kind = m->kind, S = m->S, n = m->n;
    if (kind == 0) {  // INIT: take over the owner's exact table, sizes, clock
      for (uint32_t i = lane; i < S + 2; i += 32) sts_f64(XF.sB + i * 8u, lds_f64(soB + i * 8u));
      for (uint32_t i = lane; i < S; i += 32) XF.c_len[i] = (T)lds_tab<T>(soLen, i);
      t = C->t_init;
      lambda = C->lambda_init;
      theta = 1.0 / lambda;
      __syncwarp();
    } else {
      // exact time chain over the committed iterations, with the rate of the state they ran in
      const uint32_t cnt = m->m;
      const uint64_t it0 = m->it_first;
      XF.key.stream = M->stream;
      for (uint32_t done = 0; done < cnt; done += 32) {
        const uint32_t left = cnt - done;
        const uint64_t q = it0 + 1 + done + (uint64_t)lane;
        uint64_t wa, wb;
        draw_block(XF.key, q, 0, DS_TIME_CLASS, 0, wa, wb);
        const double dt = (-contract_log(u01_open(wa))) * theta;  // :699
        t = time_chain(XF.sDt, lane, t, dt, left >= 32u ? 31u : left - 1u);  // :701
      }
      const uint32_t n_rows = m->n_rows;
      if (lane < (int)n_rows) *reinterpret_cast<double*>(A.log_pool + m->rows[lane]) = t;
      const uint32_t from = m->from;
      if (from != kNone) {
        const uint32_t n_upd = m->n_upd;
        if (lane == 0) {
          for (uint32_t u = 0; u < n_upd; ++u) {
            const uint32_t idx = m->upd_idx[u];
            const int k = m->upd_kind[u];
            XF.c_len[idx] = (T)(k == 2 ? 1u : (uint32_t)XF.c_len[idx] + (uint32_t)k);
          }
        }
        __syncwarp();
        Rep rf;
        rf.S = S; rf.n = n; rf.dirty_from = from; rf.lambda = lambda; rf.theta = theta; rf.guard = 0.0; rf.exact_valid = false;
        refold(XF, rf);
        lambda = rf.lambda;
        theta = rf.theta;
      }
    }
    if (lane == 0) {
      C->t_exact = t;
      C->lambda_exact = lambda;
      __threadfence_block();
      C->done = seq;
    }
    __syncwarp();
  }
}

// ---- owner side -----------------------------------------------------------------------------
__device__ __forceinline__ void clock_wait_idle(volatile ClockChan* C) {
  while (C->done != C->posted) __nanosleep(20);
  __threadfence_block();
}

// writes one message into the helper's ring (lane 0), waiting for a free slot
__device__ __noinline__ void clock_post(volatile HelpMail* M, volatile ClockMsg* ring, int lane, uint32_t kind,
                                        unsigned long long it_first, uint32_t m, uint32_t from, uint32_t S, uint32_t n,
                                        uint32_t n_rows, uint32_t row0, uint32_t row1, uint32_t row2, uint32_t n_upd,
                                        uint32_t u0, int k0, uint32_t u1, int k1, uint32_t u2, int k2) {
  volatile ClockChan* C = &M->clk;
  uint32_t seq = 0;
  if (lane == 0) {
    seq = C->posted + 1u;
    while (seq - C->done > kClockRing - 1u) __nanosleep(20);
    volatile ClockMsg* q = &ring[seq % kClockRing];
    q->it_first = it_first; q->kind = kind; q->m = m; q->from = from; q->S = S; q->n = n;
    q->n_rows = n_rows; q->rows[0] = row0; q->rows[1] = row1; q->rows[2] = row2;
    q->n_upd = n_upd; q->upd_idx[0] = u0; q->upd_idx[1] = u1; q->upd_idx[2] = u2;
    q->upd_kind[0] = k0; q->upd_kind[1] = k1; q->upd_kind[2] = k2;
    __threadfence_block();  // the event's stores (lists, fitness table, log records) and the message, before the count
    C->posted = seq;
  }
  __syncwarp();
}

// Enter decoupled mode: the owner's table is exact and clean.
template <typename T>
__device__ __noinline__ void dec_enter(Ctx<T>& X, Rep& r, volatile HelpMail* M, volatile ClockMsg* ring) {
  volatile ClockChan* C = &M->clk;
  if (X.lane == 0) {
    C->t_init = r.t; C->lambda_init = r.lambda; M->stream = X.key.stream;
    atomicAdd(X.A->work_counter + 2, 1ull);  // statistics: entries into decoupled mode (JSB_TIMING)
  }
  __syncwarp();
  clock_post(M, ring, X.lane, 0u, r.it, 0u, kNone, r.S, r.n, 0u, 0u, 0u, 0u, 0u, 0u, 0, 0u, 0, 0u, 0);
  clock_wait_idle(C);
}

// Leave decoupled mode: wait for the helper, take its exact table, total rate and clock.
template <typename T>
__device__ __noinline__ void dec_leave(Ctx<T>& X, Rep& r, volatile HelpMail* M, unsigned char* helper_slice) {
  volatile ClockChan* C = &M->clk;
  clock_wait_idle(C);
  const uint32_t sF = smem_addr(helper_slice);
  for (uint32_t i = X.lane; i < r.S + 2; i += 32) sts_f64(X.sB + i * 8u, lds_f64(sF + i * 8u));
  r.t = C->t_exact;
  r.lambda = C->lambda_exact;
  r.theta = 1.0 / r.lambda;  // :699, the same expression the fold uses
  r.guard = ((X.P->quirks & JSB_QUIRK_DEBUG_WIDE_GUARD) ? 1e-3 : kGuardEps) * r.lambda;
  r.exact_valid = false;
  r.dirty_from = kNone;
  __syncwarp();
}

// The approximate counterpart of refold(): every boundary from the changed clones on moves by the
// same ±α (one rounding per entry per event; re-based on the exact table every kClockResync events,
// so the drift stays below 2e-13·λ, two orders of magnitude inside the guard band).
template <typename T>
__device__ __noinline__ void approx_update(Ctx<T>& X, Rep& r, uint32_t S_old, uint32_t i0, double a0, uint32_t i1, double a1,
                                           bool has_new, double new_alpha) {
  const uint32_t sB = X.sB;
  const uint32_t lo = i0 < i1 ? i0 : i1;
  if (lo != kNone) {
    for (uint32_t i = lo + (uint32_t)X.lane; i < S_old; i += 32) {
      double v = lds_f64(sB + i * 8u);
      if (i >= i0) v = v + a0;
      if (i >= i1) v = v + a1;
      // the class search orders the entries by their bit patterns: never negative (true values are >= 0)
      sts_f64(sB + i * 8u, v < 0.0 ? 0.0 : v);
    }
  }
  __syncwarp();
  const uint32_t S = r.S;
  if (X.lane == 0) {
    double acc = lds_f64(sB + (S_old - 1u) * 8u);
    if (has_new) { acc = acc + new_alpha; sts_f64(sB + S_old * 8u, acc); }
    const double bd = acc + X.P->rate_death * (double)r.n;
    sts_f64(sB + S * 8u, bd);
    sts_f64(sB + (S + 1u) * 8u, bd + X.P->rate_migration * (double)r.n);
  }
  __syncwarp();
  const double lambda = lds_f64(sB + (S + 1u) * 8u);
  {
    uint32_t top = X.search_top;
    while (top < S + 2) top <<= 1;
    X.search_top = top;
  }
  r.lambda = lambda;
  r.theta = 1.0 / lambda;
  r.guard = ((X.P->quirks & JSB_QUIRK_DEBUG_WIDE_GUARD) ? 1e-3 : 2.0 * kGuardEps) * lambda;
  r.exact_valid = false;
  r.dirty_from = kNone;
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// One replicate
// ---------------------------------------------------------------------------------------------
// Runs rounds on the approximate clock (see clock_serve) until the exact state is needed again: a
// time that matters comes within reach, a draw lands in the guard band, or the replicate ends.
// On return the state is exact and clean (the helper's table, rate and clock have been taken
// over); returns true when the replicate is over (extinction or an error status).
template <typename T>
__device__ __noinline__ bool run_decoupled(Ctx<T>& Xref, Rep& rref, volatile HelpMail* M, const int my_warp, const double Tf,
                                           const uint64_t max_iter, const double watch_snap, const double watch_tb) {
  Ctx<T> X = Xref;
  Rep r = rref;
  const RunArgs& A = *X.A;
  const DevPoint& P = *X.P;
  const int lane = X.lane;
  uint32_t hw = 0;
  if (lane == 0) hw = M->clk.warp;
  hw = __shfl_sync(FULL, hw, 0);
  unsigned char* fh_slice = slice_of(X.B, my_warp, A.smem_per_warp, hw);
  volatile ClockMsg* ring = ring_of<T>(fh_slice);
  { Ctx<T> xc = X; Rep rc = r; dec_enter(xc, rc, M, ring); }
  r.guard = ((P.quirks & JSB_QUIRK_DEBUG_WIDE_GUARD) ? 1e-3 : 2.0 * kGuardEps) * r.lambda;
  uint32_t dec_events = 0;  // events since the approximate table was re-based on the exact one
  bool finished = false;

  for (;;) {
    if (r.n == 0) { finished = true; break; }  // extinction; the approximate clock never reaches Tf in here
    if (dec_events >= kClockResync) {          // re-base the approximate table and clock
      { Ctx<T> xc = X; Rep rc = r; dec_leave(xc, rc, M, fh_slice); r = rc; }
      { Ctx<T> xc = X; Rep rc = r; dec_enter(xc, rc, M, ring); }
      r.guard = ((P.quirks & JSB_QUIRK_DEBUG_WIDE_GUARD) ? 1e-3 : 2.0 * kGuardEps) * r.lambda;
      dec_events = 0;
    }
    // ---- lookahead requests, as in the exact loop (the rate they get is the approximate one) ----
    uint32_t nposted = 0, reqid = 0;
    {
      if (lane == 0) {
        const uint32_t nh = M->nhelp;
        nposted = nh;
        reqid = ((M->word >> 8) + 1u) & 0xFFFFFFu;
      }
      nposted = __shfl_sync(FULL, nposted, 0);
      reqid = __shfl_sync(FULL, reqid, 0);
      if (nposted) {
        if (nposted > (uint32_t)kMaxLook) nposted = (uint32_t)kMaxLook;
        if (lane == 0) {
          M->S = r.S; M->n = r.n; M->search_top = X.search_top; M->stream = X.key.stream;
          M->it_base = r.it;
          M->lambda = r.lambda; M->guard = r.guard; M->theta = r.theta;
          __threadfence_block();
          M->word = (reqid << 8) | nposted;
        }
        __syncwarp();
      }
    }
    auto wait_acks = [&]() {
      for (uint32_t h = 0; h < nposted; ++h)
        while (M->res[h].ack != reqid) __nanosleep(20);
      __threadfence_block();
    };

    double dtv[kMaxSub];
    uint32_t effm[kMaxSub], cls[kMaxSub], xx[kMaxSub], cell[kMaxSub], pos[kMaxSub];
    if (!batch_eval<T, 1>(X, r, false, FULL, dtv, effm, cls, xx, cell, pos, true)) {
      wait_acks();  // a draw inside the guard band: the exact loop decides this round
      break;
    }
    dtv[0] = dtv[0] * r.theta;
    bool found = effm[0] != 0;
    uint32_t jE = found ? (uint32_t)__ffs(effm[0]) - 1u : 31u;

    // exact end time of a span lies in [lo, hi] of the approximate one (drift between re-bases << 1e-9)
    auto may_touch = [&](double t1a, uint64_t n_commit) -> bool {
      const double hi = t1a * (1.0 + 1e-9) + 1e-12;
      const double lo = r.t * (1.0 - 1e-9) - 1e-12;
      bool sl = !(hi < Tf);
      if (max_iter && r.it + n_commit > max_iter) sl = true;
      if (hi > watch_snap) sl = true;
      if (watch_tb > lo && watch_tb <= hi) sl = true;
      return sl;
    };
    double span = (uint32_t)lane <= jE ? dtv[0] : 0.0;
    for (int d = 16; d > 0; d >>= 1) span += __shfl_xor_sync(FULL, span, d);
    double t1 = r.t + span;
    if (may_touch(t1, (uint64_t)jE + 1u)) {
      wait_acks();
      break;
    }
    uint64_t adv = 0;
    int from_helper = -1;
    if (nposted) {
      wait_acks();
      for (uint32_t h = 0; h < nposted && !found; ++h) {
        volatile HelpRes* R = &M->res[h];
        if (!R->usable) break;
        const uint32_t effm2 = R->effm;
        const uint32_t jE2 = effm2 ? (uint32_t)__ffs(effm2) - 1u : 31u;
        const double t2 = t1 + R->sum_dt;
        if (may_touch(t2, adv + 32u + jE2 + 1u)) break;
        adv += 32u;
        t1 = t2;
        jE = jE2;
        found = effm2 != 0;
        from_helper = (int)h;
      }
    }
    const uint64_t it_first = r.it;
    const uint32_t m_commit = (uint32_t)adv + jE + 1u;
    r.it += m_commit;
    r.t = t1;  // approximate; the log records written below get the exact time from the helper

    // ---- the event, and what it does to the clone sizes -----------------------------------------
    bool keep_going = true;
    const uint32_t S_old = r.S, chunk0 = r.cur_chunk;
    const uint64_t nlog0 = r.nlog;
    uint32_t u_i0 = kNone, u_i1 = kNone;  // existing clones whose size changes (by -1 / +1)
    double u_a0 = 0.0, u_a1 = 0.0;        // the signed change of their class weight
    if (found) {
      Eval ej;
      if (from_helper >= 0) {
        volatile HelpRes* R = &M->res[from_helper];
        ej.cls = R->cls; ej.x = R->x; ej.cell = R->cell; ej.pos = R->pos;
      } else {
        ej.cls = __shfl_sync(FULL, cls[0], (int)jE);
        ej.x = __shfl_sync(FULL, xx[0], (int)jE);
        ej.cell = __shfl_sync(FULL, cell[0], (int)jE);
        ej.pos = __shfl_sync(FULL, pos[0], (int)jE);
      }
      ej.occ = (ej.cls < r.S && P.rule != JSB_RULE_CONTACT) ? (uint32_t)X.clone[ej.pos] : 0u;
      ej.effective = true;
      if (ej.cls == r.S + 1) apply_migration(X, r, ej);
      else if (ej.cls == r.S) {
        const uint32_t cl = apply_death(X, r, ej);
        u_i0 = cl - 1u; u_a0 = -X.c_alpha[cl - 1u];
      } else {
        // requested before the event so that the loads are back when it is done
        u_i1 = ej.cls; u_a1 = X.c_alpha[ej.cls];
        if (ej.occ) { u_i0 = ej.occ - 1u; u_a0 = -X.c_alpha[ej.occ - 1u]; }
        keep_going = apply_birth(X, r, ej);
      }
    }
    // ---- tell the clock helper what this round did; move the approximate table -----------------
    const bool has_new = r.S > S_old;
    if (has_new) { u_i1 = kNone; u_a1 = 0.0; }  // a driver birth leaves the parent clone's size unchanged
    if (u_i0 == u_i1) { u_i0 = u_i1 = kNone; u_a0 = u_a1 = 0.0; }  // a cell displaced by its own clone: no change
    uint32_t n_rows = r.log_on ? (uint32_t)(r.nlog - nlog0) : 0u;
    if (n_rows > 3u) n_rows = 0u;  // cannot happen (at most displaced + two daughters)
    uint32_t rows[3] = {0u, 0u, 0u};
#pragma unroll
    for (uint32_t k = 0; k < 3u; ++k) {
      const uint64_t pr = nlog0 + k;
      const uint32_t ch = (pr / kLogChunk == (r.nlog - 1) / kLogChunk) ? r.cur_chunk : chunk0;
      rows[k] = ch * kLogChunk + (uint32_t)(pr % kLogChunk);
    }
    const uint32_t from = keep_going ? r.dirty_from : kNone;
    uint32_t n_upd = 0, ui[3] = {0u, 0u, 0u};
    int uk[3] = {0, 0, 0};
    if (u_i0 != kNone) { ui[n_upd] = u_i0; uk[n_upd] = -1; ++n_upd; }
    if (u_i1 != kNone) { ui[n_upd] = u_i1; uk[n_upd] = 1; ++n_upd; }
    if (has_new) { ui[n_upd] = S_old; uk[n_upd] = 2; ++n_upd; }
    clock_post(M, ring, lane, 1u, it_first, m_commit, from, r.S, r.n, n_rows, rows[0], rows[1], rows[2], n_upd, ui[0], uk[0],
               ui[1], uk[1], ui[2], uk[2]);
    if (from != kNone) {
      const double new_alpha = has_new ? X.c_alpha[S_old] : 0.0;
      Ctx<T> xc = X; Rep rc = r;
      approx_update(xc, rc, S_old, u_i0, u_a0, u_i1, u_a1, has_new, new_alpha);
      r = rc; X.search_top = xc.search_top;
      ++dec_events;
    }
    if (!keep_going || r.status == JSB_ST_INTERNAL) { finished = true; break; }
  }
  { Ctx<T> xc = X; Rep rc = r; dec_leave(xc, rc, M, fh_slice); r = rc; }
  Xref.search_top = X.search_top;
  rref = r;
  return finished;
}

// ---------------------------------------------------------------------------------------------
// Suspending pilot (RunArgs::susp).  A replicate that reaches the pilot depth is parked: the live part of its
// slot workspace, the Rep registers, the per-lane label bitmap and the shared per-clone tables go to
// susp[local_idx]; the full pass copies them back into whatever slot picks the replicate up, rebuilds the
// occupancy bitmask from clone[] and re-folds the class boundaries from clone 0 (the sequential fold is
// deterministic), and carries on at iteration it + 1.  Draws are addressed by iteration number, so the
// trajectory is the one an uninterrupted run produces.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __noinline__ void susp_copy(const Ctx<T>& X, uint32_t arena_top, bool save) {
  const RunArgs& A = *X.A;
  uint8_t* park = A.susp + (uint64_t)X.local_idx * A.susp_stride;
  uint8_t* ws = X.ws_base;
  const uint64_t live = A.L.arena + (uint64_t)arena_top * sizeof(T);   // clone, cellid, cs_alive, heads, hints, arena in use
  const uint64_t tab0 = A.L.c_cnt, tab1 = A.L.g_w;                     // per-clone tables kept in global memory
  if (save) { warp_copy16(park, ws, live, X.lane); warp_copy16(park + tab0, ws + tab0, tab1 - tab0, X.lane); }
  else { warp_copy16(ws, park, live, X.lane); warp_copy16(ws + tab0, park + tab0, tab1 - tab0, X.lane); }
}
template <typename T>
__device__ __noinline__ void susp_save(const Ctx<T>& X, const Rep& r) {
  const RunArgs& A = *X.A;
  __syncwarp();
  susp_copy(X, r.arena_top, true);
  uint8_t* blk = A.susp + (uint64_t)X.local_idx * A.susp_stride + A.ws_stride;
  if (X.lane == 0) *reinterpret_cast<Rep*>(blk) = r;
  reinterpret_cast<uint32_t*>(blk + 256)[X.lane] = r.used_w;
  T* tl = reinterpret_cast<T*>(blk + 384);
  for (uint32_t i = X.lane; i < r.S; i += 32) { tl[i] = X.c_len[i]; tl[kTab + i] = X.c_off8[i]; }
  __threadfence();
  __syncwarp();
  if (X.lane == 0) A.susp_state[X.local_idx] = 1u;
}
template <typename T>
__device__ __noinline__ void susp_restore(Ctx<T>& X, Rep& r) {
  const RunArgs& A = *X.A;
  const uint8_t* blk = A.susp + (uint64_t)X.local_idx * A.susp_stride + A.ws_stride;
  r = *reinterpret_cast<const Rep*>(blk);
  r.used_w = reinterpret_cast<const uint32_t*>(blk + 256)[X.lane];
  susp_copy(X, r.arena_top, false);
  const T* tl = reinterpret_cast<const T*>(blk + 384);
  for (uint32_t i = X.lane; i < r.S; i += 32) { X.c_len[i] = tl[i]; X.c_off8[i] = tl[kTab + i]; }
  for (uint32_t i = X.lane; i < (uint32_t)kSmemTable; i += 32) X.B[i] = __longlong_as_double(0x7ff0000000000000ll);  // +inf pad
  __syncwarp();  // the workspace copy above is visible to the whole warp
  if (X.occ) {
    const uint32_t nv = A.g.nv;
    for (uint32_t w = X.lane; w < A.occ_words; w += 32) {
      uint32_t bits = 0;
      const uint32_t s0 = w * 32u;
      for (uint32_t k = 0; k < 32u && s0 + k < nv; k += 2) {  // clone[] is u16: two sites per 32-bit load (nv + 1 entries are allocated)
        const uint32_t two = *reinterpret_cast<const uint32_t*>(X.clone + s0 + k);
        bits |= (uint32_t)((two & 0xFFFFu) != 0u) << k;
        if (s0 + k + 1 < nv) bits |= (uint32_t)((two >> 16) != 0u) << (k + 1);
      }
      X.occ[w] = bits;
    }
  }
  r.dirty_from = 0;
  r.exact_valid = false;
  // the pilot depth is not an outcome; a log pool that ran out during the pilot pass is
  r.status = ((A.flags & JSB_OUT_EVENTS) && !r.log_on) ? JSB_ST_LOG_OVERFLOW : JSB_ST_OK;
  X.search_top = 2;
  __syncwarp();
}

template <typename T>
__device__ __forceinline__ void run_replicate(Ctx<T>& X, volatile HelpMail* M, const int my_warp) {
  const RunArgs& A = *X.A;
  const DevPoint& P = *X.P;
  const int lane = X.lane;
  Rep r;
  r.t = 0.0; r.theta = 0.0; r.lambda = 0.0; r.guard = 0.0; r.it = 0; r.nlog = 0; r.nstored = 0; r.cur_chunk = 0;
  r.bott_id = P.n_bott > 0 ? 1 : 0;  // :668-672
  r.snap_idx = 0;
  r.status = JSB_ST_OK;
  r.n_eff = r.n_births = r.n_deaths = r.n_migr = r.n_disp = r.n_exc = 0;
  r.log_on = (A.flags & JSB_OUT_EVENTS) != 0 && (A.pilot_iters == 0 || A.susp != nullptr);
  r.exact_valid = false;
  r.dirty_from = 0;
  X.search_top = 2;
  if (A.susp && A.pilot_iters == 0 && A.susp_state[X.local_idx] == 1u) {  // parked by the pilot pass: carry on
    Ctx<T> xc = X; Rep rc = r; susp_restore(xc, rc); r = rc; X.search_top = xc.search_top;
  } else {
    Ctx<T> xc = X; Rep rc = r; seed_replicate(xc, rc); r = rc;
  }

  const bool intent = (P.quirks & JSB_QUIRK_EXCISION_INTENT) != 0;
  const bool force_exact = (P.quirks & JSB_QUIRK_DEBUG_EXACT_CLASSES) != 0;
  const uint32_t width = (P.quirks & JSB_QUIRK_DEBUG_SERIAL) ? 1u : 32u;
  const uint32_t wmask = width == 32 ? FULL : ((1u << width) - 1u);
  const uint32_t nv = A.g.nv;
  uint32_t Ucur = 1;  // sub-batches per speculative window (1, 2 or 4), adapted to the run length
  if (force_exact || width != 32u) M = nullptr;

  // Times that end a fast window, kept in registers (they live in global memory): the next snapshot
  // time, and the excision time the reference's index arithmetic is looking at (:813-838 method 1
  // never advances it, :1062-1064 method 2 advances it every iteration up to the last entry).
  const double Tf = P.Tf;
  const uint64_t max_iter = P.max_iter;
  const int n_bott = P.n_bott;
  const bool tree_quirk = n_bott > 0 && !intent && P.method == JSB_METHOD_TREE;
  const double kNever = __longlong_as_double(0x7ff0000000000000ll);  // +inf
  double watch_snap = kNever, watch_tb = -1.0;
  auto reload_watch = [&]() {
    watch_snap = r.snap_idx < P.n_snap ? X.tsnap[r.snap_idx] : kNever;
    watch_tb = -1.0;  // never inside (r.t, t1]: event times are positive
    if (n_bott > 0) {
      if (intent) { if (r.bott_id > 0) watch_tb = X.tb[r.bott_id - 1]; }
      else if (P.method == JSB_METHOD_TREE) watch_tb = X.tb[n_bott - 1];
      else watch_tb = X.tb[0];
    }
  };
  reload_watch();

  const bool dec_allowed = M != nullptr && (A.helpers & 2u) == 0 && A.pilot_iters == 0;
  const uint32_t dec_min_clones = (P.quirks & JSB_QUIRK_DEBUG_DECOUPLE) ? 1u : A.dec_min_clones;
  uint64_t dec_hold = 0;  // no (re-)entry into decoupled mode before this iteration

  PROF_DECL
  for (;;) {  // exact loop <-> decoupled stretches
  bool want_dec = false;
  for (;;) {
    if (!(r.t < Tf && r.n > 0)) {  // loop test :687-688
      if (P.rate_death == 0.0 && r.n == nv) r.status = JSB_ST_NONTERMINATING;
      break;
    }
    PROF_T0();
#ifdef JSB_PROFILE
    if (r.dirty_from != kNone) prof_fold += (long long)r.S - (long long)(r.dirty_from < r.S ? r.dirty_from : r.S);
#endif
    if (r.dirty_from != kNone) { refold(X, r); PROF_ADD(PH_REFOLD); }

    // ---- a clock helper is attached: run on the approximate clock until something needs the exact one
    if (dec_allowed && r.S >= dec_min_clones && r.it >= dec_hold && M->clk.attached != 0 &&
        !(tree_quirk && r.it + 1 < (uint64_t)n_bott) &&
        !(max_iter && r.it + 4096u >= max_iter)) {
      want_dec = true;  // handled outside this loop: the call must not shape its register allocation
      break;
    }

    // ---- ask the attached helper warps for the windows after this one ---------------------------
    uint32_t nposted = 0, reqid = 0;
    if (M) {
      if (lane == 0) {
        const uint32_t nh = M->nhelp;  // lookahead helpers attached
        nposted = nh;
        reqid = ((M->word >> 8) + 1u) & 0xFFFFFFu;
      }
      nposted = __shfl_sync(FULL, nposted, 0);
      reqid = __shfl_sync(FULL, reqid, 0);
      if (nposted) {
        if (nposted > (uint32_t)kMaxLook) nposted = (uint32_t)kMaxLook;
        if (lane == 0) {
          M->S = r.S; M->n = r.n; M->search_top = X.search_top; M->stream = X.key.stream;
          M->it_base = r.it;
          M->lambda = r.lambda; M->guard = r.guard; M->theta = r.theta;
          __threadfence_block();  // the state written by the last event and this request, before the word
          M->word = (reqid << 8) | nposted;
        }
        __syncwarp();
      }
    }

    // ---- speculative window: Ucur sub-batches of 32 iterations ----------------------------------
    double dtv[kMaxSub];
    uint32_t effm[kMaxSub], cls[kMaxSub], xx[kMaxSub], cell[kMaxSub], pos[kMaxSub];
    if (!kAdaptiveWindow || Ucur == 1) batch_eval<T, 1>(X, r, force_exact, wmask, dtv, effm, cls, xx, cell, pos, false);
    else if (Ucur == 2) batch_eval<T, 2>(X, r, force_exact, wmask, dtv, effm, cls, xx, cell, pos, false);
    else batch_eval<T, 4>(X, r, force_exact, wmask, dtv, effm, cls, xx, cell, pos, false);
#pragma unroll
    for (int u = 0; u < kMaxSub; ++u) dtv[u] = dtv[u] * r.theta;  // :699
    PROF_ADD(PH_EVAL);

    // first state-changing iteration of the window (sub-batch ustar, lane jE)
    uint32_t ustar = Ucur - 1, jE = width - 1u;
    bool found = false;
#pragma unroll
    for (int u = kMaxSub - 1; u >= 0; --u)
      if ((uint32_t)u < Ucur && effm[u]) { ustar = (uint32_t)u; jE = (uint32_t)__ffs(effm[u]) - 1u; found = true; }

    // ---- time chain over the iterations that may commit -----------------------------------------
    double tend = r.t;
#pragma unroll
    for (int u = 0; u < kMaxSub; ++u)
      if ((uint32_t)u <= ustar && (uint32_t)u < Ucur) tend = time_chain(X.sDt, lane, tend, dtv[u], (uint32_t)u == ustar ? jE : 31u);

    // ---- can anything other than that one event happen in (t, tend]? ---------------------------
    auto touches = [&](double t1, uint64_t n_commit) -> bool {
      bool sl = !(t1 < Tf);
      if (max_iter && r.it + n_commit > max_iter) sl = true;
      if (t1 > watch_snap) sl = true;
      if (tree_quirk && r.it + 1 < (uint64_t)n_bott) sl = true;
      if (watch_tb > r.t && watch_tb <= t1) sl = true;
      return sl;
    };
    bool slow = touches(tend, (uint64_t)ustar * 32u + jE + 1u);
    if (slow && Ucur > 1) {  // retry with the first sub-batch only
      ustar = 0;
      found = effm[0] != 0;
      jE = found ? (uint32_t)__ffs(effm[0]) - 1u : 31u;
      tend = time_chain(X.sDt, lane, r.t, dtv[0], jE);
      slow = touches(tend, (uint64_t)jE + 1u);
      Ucur = 1;
    } else if (kAdaptiveWindow && width == 32u) {
      // adapt the window to the observed run length
      if (!found) Ucur = Ucur < (uint32_t)kMaxSub ? Ucur * 2 : (uint32_t)kMaxSub;
      else Ucur = ustar == 0 ? 1u : (ustar == 1 ? 2u : (uint32_t)kMaxSub);
    }

    uint32_t j = jE;
    bool do_event = found;
    bool do_snap = false, do_bott = false;
    double tb_fire = 0.0, ratio_fire = 0.0;
    int from_helper = -1;

    if (nposted) {
      // every helper must be finished before anything changes (they read the lists) and before a
      // window is used; they have been running in parallel with this warp's own evaluation
      for (uint32_t h = 0; h < nposted; ++h)
        while (M->res[h].ack != reqid) __nanosleep(20);
      __threadfence_block();
      uint64_t adv = 0;  // full null windows committed ahead of the one that ends the round
      for (uint32_t h = 0; h < nposted && !found && !slow; ++h) {
        volatile HelpRes* R = &M->res[h];
        if (!R->usable) break;
        const uint32_t effm2 = R->effm;
        const uint32_t jE2 = effm2 ? (uint32_t)__ffs(effm2) - 1u : 31u;
        const double tend2 = time_chain_at(smem_addr(const_cast<double*>(&R->dt[0])), tend, jE2);
        if (touches(tend2, adv + 32u + jE2 + 1u)) break;  // something besides that one event: redo it the slow way
        adv += 32u;
        tend = tend2;
        j = jE = jE2;
        ustar = 0;
        found = do_event = effm2 != 0;
        from_helper = (int)h;
      }
      r.it += adv;
    }

    if (!slow) {
      r.it += (uint64_t)ustar * 32u + jE + 1u;
      r.t = tend;
    } else {
      SlowOut o;
      const bool eff0 = (effm[0] >> lane) & 1u;
      { Ctx<T> xc = X; Rep rc = r; o = slow_scan(xc, rc, dtv[0], eff0, jE, wmask, r.it + 1 + (uint64_t)lane); }
      j = o.j;
      if (o.stop || o.maxit) {
        r.it += j;
        r.t = o.t_prev;
        if (o.stop) { if (P.rate_death == 0.0 && r.n == nv) r.status = JSB_ST_NONTERMINATING; }
        else r.status = JSB_ST_MAX_ITER;
        break;
      }
      r.it += j + 1;
      r.t = o.t_cur;
      do_event = o.do_event;
      do_snap = o.do_snap;
      do_bott = o.do_bott;
      tb_fire = o.tb;
      ratio_fire = o.ratio;
    }

    PROF_ADD(PH_TIME);
    if (do_snap) {  // :708-714
      if (lane == 0 && A.snaps) {
        jsb_snapshot s;
        s.time = r.t; s.iteration = r.it; s.n_events_before = r.nlog; s.n_alive = r.n;
        s.index = (uint32_t)r.snap_idx;
        A.snaps[(size_t)X.local_idx * A.max_snap + r.snap_idx] = s;
      }
      r.snap_idx++;
      reload_watch();
    }
    bool keep_going = true;
    if (do_event) {
      Eval ej;
      ej.cls = ej.x = ej.cell = ej.pos = 0;
      if (from_helper >= 0) { volatile HelpRes* R = &M->res[from_helper]; ej.cls = R->cls; ej.x = R->x; ej.cell = R->cell; ej.pos = R->pos; }
#pragma unroll
      for (int u = 0; u < kMaxSub; ++u)
        if (from_helper < 0 && (uint32_t)u == ustar) {
          ej.cls = __shfl_sync(FULL, cls[u], (int)j);
          ej.x = __shfl_sync(FULL, xx[u], (int)j);
          ej.cell = __shfl_sync(FULL, cell[u], (int)j);
          ej.pos = __shfl_sync(FULL, pos[u], (int)j);
        }
      // the occupant's clone id matters only when a voter rule displaces it
      ej.occ = (ej.cls < r.S && P.rule != JSB_RULE_CONTACT) ? (uint32_t)X.clone[ej.pos] : 0u;
      ej.effective = true;
      if (ej.cls == r.S + 1) { apply_migration(X, r, ej); PROF_ADD(PH_MIGR); }
      else if (ej.cls == r.S) { apply_death(X, r, ej); PROF_ADD(PH_DEATH); }
      else { keep_going = apply_birth(X, r, ej); PROF_ADD(PH_BIRTH); }
    }
    if (!keep_going) break;
    if (do_bott) {
      { Ctx<T> xc = X; Rep rc = r; apply_excision(xc, rc, tb_fire, ratio_fire); r = rc; }
      PROF_ADD(PH_EXC);
      if (intent) { r.bott_id = (r.bott_id < P.n_bott) ? r.bott_id + 1 : 0; reload_watch(); }
    }
    if (r.status == JSB_ST_INTERNAL) break;
  }
  if (!want_dec) break;
  {
    // ---- a clock helper is attached: run on the approximate clock until something needs the exact one
    Ctx<T> xc = X; Rep rc = r;
    const bool finished = run_decoupled(xc, rc, M, my_warp, Tf, max_iter, watch_snap, watch_tb);
    r = rc; X.search_top = xc.search_top;
    dec_hold = r.it + 1024u;
    if (finished) break;
  }
  }
#ifdef JSB_PROFILE
  if (lane == 0 && (X.local_idx < 2 || r.it > 6000000ull)) {
    printf("sub-phases: refold A %lld B %lld tail %lld\n", g_sub[0], g_sub[1], g_sub[2]);
    printf("rep %lld it %llu rounds %lld eff %u S %u n %u | refold %lld/%lld (elems %lld) draws %lld/%lld class %lld eval %lld time %lld | birth %lld/%lld death %lld/%lld migr %lld/%lld exc %lld\n",
           (long long)X.local_idx, (unsigned long long)r.it, prof_n[PH_EVAL], r.n_eff, r.S, r.n, prof_c[PH_REFOLD], prof_n[PH_REFOLD], prof_fold,
           prof_c[PH_DRAWS], prof_n[PH_DRAWS], prof_c[PH_CLASS], prof_c[PH_EVAL], prof_c[PH_TIME], prof_c[PH_BIRTH],
           prof_n[PH_BIRTH], prof_c[PH_DEATH], prof_n[PH_DEATH], prof_c[PH_MIGR], prof_n[PH_MIGR], prof_c[PH_EXC]);
  }
#endif
  if (A.pilot_iters) {
    // work estimate for longest-first scheduling: iterations left if the total rate stayed as it is
    const uint64_t user_max = A.points[(A.g_begin + X.local_idx) / A.reps_per_point].max_iter;
    const bool capped = r.status == JSB_ST_MAX_ITER && !(user_max && r.it >= user_max);  // stopped by the pilot depth only
    if (lane == 0) A.prio[X.local_idx] = (A.susp ? capped : (r.n > 0 && r.t < P.Tf)) ? (float)(r.lambda * (P.Tf - r.t)) : 0.0f;
    if (!A.susp) return;
    if (capped) { Ctx<T> xc = X; Rep rc = r; rc.status = JSB_ST_OK; susp_save(xc, rc); return; }
    if (lane == 0) A.susp_state[X.local_idx] = 2u;  // over before the pilot depth: outputs below, nothing left for the full pass
  }
  { Ctx<T> xc = X; Rep rc = r; write_outputs(xc, rc); }
}

// Per-warp shared-memory slice:
//   [ B: kSmemTable doubles | dt staging: 32 doubles | c_len: kTab sites | c_off8: kTab sites | DevPoint | occupancy bits ]
// ("site" = the u16/u32 storage type of the ordered lists)
__host__ __device__ inline size_t smem_base_bytes(uint32_t site_bytes) {
  return sizeof(double) * (kSmemTable + 32) + 2 * (size_t)site_bytes * kTab + ((sizeof(DevPoint) + 15) & ~size_t(15));
}
// the helper mailbox sits at the end of every warp's slice
__device__ __forceinline__ volatile HelpMail* mail_of(unsigned char* s_dyn, uint32_t smem_per_warp, int warp) {
  return reinterpret_cast<volatile HelpMail*>(s_dyn + (size_t)(warp + 1) * smem_per_warp - sizeof(HelpMail));
}

// context of warp slot `warp` of this block (a warp's own slot, or the slot of the warp it helps)
template <typename T>
__device__ __forceinline__ Ctx<T> make_ctx(const RunArgs& A, unsigned char* s_dyn, int warp, int lane, DevPoint** sP_out) {
  const uint64_t slot = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  const WsLayout& L = A.L;
  uint8_t* base = A.ws + slot * A.ws_stride;
  unsigned char* sw = s_dyn + (size_t)warp * A.smem_per_warp;
  Ctx<T> X;
  X.A = &A;
  X.lane = lane;
  X.ws_base = base;
  X.clone = reinterpret_cast<uint16_t*>(base + L.clone);
  X.cellid = reinterpret_cast<uint32_t*>(base + L.cellid);
  X.Alist = reinterpret_cast<T*>(base + L.A);
  X.arena = reinterpret_cast<T*>(base + L.arena);
  X.a_head = reinterpret_cast<uint16_t*>(base + L.a_head);
  X.l_head = reinterpret_cast<uint16_t*>(base + L.l_head);
  X.lseg = reinterpret_cast<T*>(base + L.lseg);
  X.inds = reinterpret_cast<uint32_t*>(base + L.inds);
  X.samp = reinterpret_cast<uint32_t*>(base + L.samp);
  X.c_cap = reinterpret_cast<uint32_t*>(base + L.c_cap);
  X.c_alpha = reinterpret_cast<double*>(base + L.c_alpha);
  X.c_parent = reinterpret_cast<uint16_t*>(base + L.c_parent);
  X.c_label = reinterpret_cast<uint16_t*>(base + L.c_label);
  X.g_w = reinterpret_cast<double*>(base + L.g_w);
  X.g_c = reinterpret_cast<double*>(base + L.g_c);
  X.B = reinterpret_cast<double*>(sw);
  X.c_len = reinterpret_cast<T*>(sw + sizeof(double) * (kSmemTable + 32));
  X.c_off8 = X.c_len + kTab;
  DevPoint* sP = reinterpret_cast<DevPoint*>(X.c_off8 + kTab);
  X.occ = A.occ_words ? reinterpret_cast<uint32_t*>(sw + smem_base_bytes((uint32_t)sizeof(T))) : nullptr;
  X.sB = smem_addr(X.B);
  X.sDt = X.sB + 8u * kSmemTable;
  X.sLen = smem_addr(X.c_len);
  X.sOff = smem_addr(X.c_off8);
  X.sOcc = X.occ ? smem_addr(X.occ) : 0u;
  // opaque to the compiler from here on: otherwise ptxas re-derives the shared-window addresses (S2R
  // SR_CgaCtaId + cvta, ~4 % of all issued instructions in the profile) wherever it is short of registers
  asm volatile("" : "+r"(X.sB), "+r"(X.sDt), "+r"(X.sLen), "+r"(X.sOff), "+r"(X.sOcc));
  X.search_top = 2;
  X.key.k0 = (uint32_t)A.seed;
  X.key.k1 = (uint32_t)(A.seed >> 32);
  X.key.stream = 0;
  X.track_ids = (A.flags & (JSB_OUT_EVENTS | JSB_OUT_LATTICE)) != 0;
  X.P = sP;
  X.local_idx = 0;
  X.tb = X.ratio = X.tsnap = X.node_rate = nullptr;
  X.edge_parent = X.edge_child = nullptr;
  *sP_out = sP;
  return X;
}

// MAXT = largest block the instantiation is launched with.  Latency-bound ensembles run at most eight warps per
// SM (see jsb_run), which leaves 255 registers per thread instead of 168: the 8-warp instantiation has no spills.
template <typename T, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) ssa_kernel(const __grid_constant__ RunArgs A) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  DevPoint* sP;
  Ctx<T> X = make_ctx<T>(A, s_dyn, warp, lane, &sP);
  volatile HelpMail* myM = mail_of(s_dyn, A.smem_per_warp, warp);
  if (lane == 0) {
    myM->word = 0; myM->nhelp = 0; myM->done = 0;
    myM->clk.attached = 0; myM->clk.warp = 0; myM->clk.posted = 0; myM->clk.done = 0; myM->clk.claimed = 0;
    for (int h = 0; h < kMaxLook; ++h) myM->res[h].ack = 0;
  }
  __syncthreads();  // the only block-wide barrier: mailboxes are initialised before anyone looks

  for (uint32_t round = 0;; ++round) {
    unsigned long long idx = 0;
    if (A.static_map) {
      // fewer replicates than warps: block b owns replicates b, b + grid, ...; warp w takes the w-th
      idx = (unsigned long long)blockIdx.x + (unsigned long long)warp * gridDim.x;
      if (round > 0) idx = (unsigned long long)A.n_local;
    } else if (A.order && round == 0) {
      // longest-first order: the heaviest replicates are dealt out one per SM (rank k -> block k mod
      // grid), so that the long tails end up on different SMs; the work counter starts after them
      // (boustrophedon: the SM that holds the heaviest replicate gets the lightest of every second row)
      const unsigned b = (warp & 1) ? gridDim.x - 1u - blockIdx.x : blockIdx.x;
      idx = (unsigned long long)b + (unsigned long long)warp * gridDim.x;
    } else {
      if (lane == 0) idx = atomicAdd(A.work_counter, 1ull);
      idx = __shfl_sync(FULL, idx, 0);
    }
    if ((long long)idx >= A.n_local) break;
    if (A.order) idx = A.order[idx];
    if (A.susp && A.pilot_iters == 0 && A.susp_state[idx] == 2u) continue;  // finished inside the pilot pass
    const int64_t gidx = A.g_begin + (int64_t)idx;
    X.local_idx = (int64_t)idx;
    X.key.stream = (uint32_t)gidx;
    {  // stage this replicate's sweep point in shared memory
      const uint32_t* src = reinterpret_cast<const uint32_t*>(A.points + (gidx / A.reps_per_point));
      uint32_t* dst = reinterpret_cast<uint32_t*>(sP);
      for (uint32_t i = lane; i < sizeof(DevPoint) / 4; i += 32) dst[i] = src[i];
      __syncwarp();
      if (A.pilot_iters && lane == 0 && (sP->max_iter == 0 || sP->max_iter > A.pilot_iters)) sP->max_iter = A.pilot_iters;
      __syncwarp();
    }
    const DevPoint* P = sP;
    X.P = P;
    X.tb = A.dpool + P->bott_off;
    X.ratio = X.tb + P->n_bott;
    X.tsnap = A.dpool + P->snap_off;
    X.node_rate = A.dpool + P->rate_off;
    X.edge_parent = A.ipool + P->edge_off;
    X.edge_child = X.edge_parent + P->n_edges;
    unsigned long long t_start = 0;
    if (A.trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    run_replicate<T>(X, A.helpers ? myM : nullptr, warp);
    if (A.trace && lane == 0) {
      unsigned long long t_end;
      uint32_t smid;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      unsigned long long* tr = A.trace + 4ull * idx;
      tr[0] = t_start; tr[1] = t_end; tr[2] = ((unsigned long long)smid << 8) | (unsigned)warp;
      tr[3] = A.summaries[idx].n_iterations;
    }
  }

  if (!A.helpers) return;
  // No replicate left for this warp: help the busy warps of the block.  A clock helper is as busy as
  // the warp it serves, so that role is only taken when few replicates are left on this SM (it would
  // slow the others down as much as it speeds one up); lookahead helpers are nearly free.
  if (lane == 0) myM->done = 1;
  __threadfence_block();
  __syncwarp();
  for (;;) {
    int target = -1;
    uint32_t role = 0;  // 0 = clock, 1 + k = lookahead window k
    if (lane == 0) {
      while (target < 0) {
        int busy = 0, best = -1, cbest = -1;
        uint32_t best_n = (uint32_t)kMaxLook;
        for (int w = 0; w < nwarps; ++w) {
          if (w == warp) continue;
          volatile HelpMail* m = mail_of(s_dyn, A.smem_per_warp, w);
          if (m->done) continue;
          ++busy;
          if (cbest < 0 && m->clk.claimed == 0) cbest = w;
          const uint32_t nh = m->nhelp;
          if (nh < best_n) { best_n = nh; best = w; }
        }
        if (cbest >= 0 && busy <= (int)A.clock_max_busy) {
          volatile HelpMail* m = mail_of(s_dyn, A.smem_per_warp, cbest);
          if (atomicCAS(const_cast<uint32_t*>(&m->clk.claimed), 0u, 1u) == 0u) { target = cbest; role = 0; }
          continue;
        }
        if (best < 0) break;
        volatile HelpMail* m = mail_of(s_dyn, A.smem_per_warp, best);
        if (atomicCAS(const_cast<uint32_t*>(&m->nhelp), best_n, best_n + 1u) == best_n) { target = best; role = 1u + best_n; }
      }
    }
    target = __shfl_sync(FULL, target, 0);
    role = __shfl_sync(FULL, role, 0);
    if (target < 0) break;
    DevPoint* oP;
    Ctx<T> XO = make_ctx<T>(A, s_dyn, target, lane, &oP);
    if (role == 0u) clock_serve<T>(XO, mail_of(s_dyn, A.smem_per_warp, target), s_dyn + (size_t)warp * A.smem_per_warp);
    else helper_serve<T>(XO, mail_of(s_dyn, A.smem_per_warp, target), role - 1u);
  }
}

// Chooses the shared-memory layout (tables + occupancy bitmask when it fits) and the number of
// warps per block (one block per SM).
// (device attributes and the function attributes last set are remembered per device: jsb_run serialises the calls
// on a device, and every driver query skipped is one less wait on a lock that monitoring tools hold)
struct PrepCache { int max_optin = 0, sm_max = 0, bytes = -1, pct = -1, fast_bytes = -1, fast_pct = -1; };
static PrepCache g_prep[64];

int ssa_prepare(RunArgs& args, int max_warps_per_sm, int* warps_per_block) {
  const uint32_t words = (args.g.nv + 31) / 32;
  int dev = 0;
  cudaGetDevice(&dev);
  PrepCache& pc = g_prep[dev & 63];
  if (pc.max_optin == 0) {
    if (cudaDeviceGetAttribute(&pc.max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return -1;
    cudaDeviceGetAttribute(&pc.sm_max, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  }
  const int max_optin = pc.max_optin;
  const size_t with_occ = ((smem_base_bytes(args.site_bytes) + 4ull * words + 15) & ~size_t(15)) + sizeof(HelpMail);
  const size_t without = ((smem_base_bytes(args.site_bytes) + 15) & ~size_t(15)) + sizeof(HelpMail);
  int cap = kMaxWarpsPerBlock;
  if (max_warps_per_sm > 0 && max_warps_per_sm < cap) cap = max_warps_per_sm;
  int w_occ = (int)((size_t)max_optin / with_occ);
  int w_no = (int)((size_t)max_optin / without);
  if (w_occ > cap) w_occ = cap;
  if (w_no > cap) w_no = cap;
  // keep the bitmask unless it costs more than half of the resident warps
  bool use_occ = w_occ >= 1 && w_occ * 2 >= w_no;
  int w = use_occ ? w_occ : w_no;
  if (w < 1) return -1;
  args.occ_words = use_occ ? words : 0;
  args.smem_per_warp = (uint32_t)(use_occ ? with_occ : without);
  const int bytes = (int)((size_t)args.smem_per_warp * w);
  (void)bytes;
  if (pc.bytes != max_optin) {  // once per device: the largest opt-in size, every launch geometry fits under it
    pc.bytes = -1;
    if (cudaFuncSetAttribute(ssa_kernel<uint16_t, kMaxWarpsPerBlock * 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(ssa_kernel<uint32_t, kMaxWarpsPerBlock * 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(ssa_kernel<uint16_t, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(ssa_kernel<uint32_t, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin) != cudaSuccess) return -1;
    pc.bytes = max_optin;
  }
  *warps_per_block = w;
  return 0;
}

int launch_ssa(const RunArgs& args, int grid_blocks, int warps_per_block, void* stream) {
  const size_t smem = (size_t)args.smem_per_warp * warps_per_block;
  {
    // ask for no more shared-memory carve-out than the block needs: what is left of the 256 KB is the L1 that
    // caches the member lists (hit rate ~57 %; every KB of carve-out taken from it was measurable).  Set when the
    // geometry changes, not per launch.
    int dev = 0;
    cudaGetDevice(&dev);
    PrepCache& pc = g_prep[dev & 63];
    // rounded down: the driver never goes below what the block needs, and rounding up can skip a step
    int pct = pc.sm_max > 0 ? (int)(((long long)(smem + 1024) * 100) / pc.sm_max) : 100;
    if (pct > 100) pct = 100;
    if (pct != pc.pct) {
      cudaFuncSetAttribute(ssa_kernel<uint16_t, kMaxWarpsPerBlock * 32>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
      cudaFuncSetAttribute(ssa_kernel<uint32_t, kMaxWarpsPerBlock * 32>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
      cudaFuncSetAttribute(ssa_kernel<uint16_t, 256>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
      cudaFuncSetAttribute(ssa_kernel<uint32_t, 256>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
      pc.pct = pct;
    }
  }
  const bool small_block = warps_per_block <= 8;
  if (args.site_bytes == 2) {
    if (small_block) ssa_kernel<uint16_t, 256><<<grid_blocks, warps_per_block * 32, smem, (cudaStream_t)stream>>>(args);
    else ssa_kernel<uint16_t, kMaxWarpsPerBlock * 32><<<grid_blocks, warps_per_block * 32, smem, (cudaStream_t)stream>>>(args);
  } else {
    if (small_block) ssa_kernel<uint32_t, 256><<<grid_blocks, warps_per_block * 32, smem, (cudaStream_t)stream>>>(args);
    else ssa_kernel<uint32_t, kMaxWarpsPerBlock * 32><<<grid_blocks, warps_per_block * 32, smem, (cudaStream_t)stream>>>(args);
  }
  return (int)cudaGetLastError();
}


// =================================================================================================
// Rejection-free mode (JSB_MODE_REJECTION_FREE)
//
// Same stochastic law as the reference loop, different algorithm.  In the reference a cell of clone
// c is picked at rate α_c, a uniform neighbour is drawn and the event is rejected unless the contact
// rule admits it (src/J_Space.jl:762-786); 79-99 % of iterations are such null events.  By the
// thinning property the state-changing events form a process in which cell v fires at rate
//     p_v = α_c·adm_v/deg_v  (birth into a uniformly chosen admissible neighbour)
//         + rate_death
//         + rate_migration·emp_v/deg_v  (move to a uniformly chosen empty neighbour)
// and time advances by Exp(Σ_v p_v) between them.  p_v sits in a 32-ary sum tree: leaves (one fp64
// per site) in HBM, every upper level in shared memory; a warp selects an event by one 32-wide
// scan + ballot per level and updates the ≤ deg+1 sites around a changed site with shared-memory
// atomics.  Upper levels are re-summed from the leaves every kFastRebuild events (rounding drift).
//
// What the reference does at Tf, at an excision time and for snapshots depends on the (invisible)
// null iterations; that is reproduced in law: given the next state-changing event at t' and a
// boundary τ in (t, t'), the event belongs to the loop iteration that crosses τ with probability
// exp(−ν (t' − τ)), ν = λ_reference − Σ p_v being the null-iteration rate.
// =================================================================================================
constexpr int kFastLevels = 7;        // 32^6 > 2^28 sites, plus the single-entry root
__constant__ double kInvDeg[33] = {0.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10,
                                   1.0 / 11, 1.0 / 12, 1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19,
                                   1.0 / 20, 1.0 / 21, 1.0 / 22, 1.0 / 23, 1.0 / 24, 1.0 / 25, 1.0 / 26, 1.0 / 27, 1.0 / 28,
                                   1.0 / 29, 1.0 / 30, 1.0 / 31, 1.0 / 32};  // x/deg as x * kInvDeg[deg] (no fp64 division)
constexpr uint32_t kFastRebuild = 2048;       // events between re-summations of the upper levels
constexpr uint32_t kFastRefreshLeaves = 8;   // every this many re-summations the leaves are recomputed too
enum : uint32_t { DS_F_TIME_SELECT = 12, DS_F_KIND_NBR = 13, DS_F_BOUNDARY = 11 };

struct FCtx {
  const RunArgs* A;
  const DevPoint* P;
  const double *tb, *ratio, *tsnap, *node_rate;
  const int32_t *edge_parent, *edge_child;
  uint16_t* clone;
  uint32_t* cellid;
  uint32_t *inds, *samp;
  double* lvl[kFastLevels];  // lvl[0] = per-site rates, lvl[k][i] = sum of lvl[k-1][32i .. 32i+31]
  uint32_t lvl_n[kFastLevels];
  int top;                   // lvl[top] is the root: one entry = the total rate (top = 0 only when nv == 1)
  uint32_t* c_cnt;
  double* c_alpha;
  uint16_t *c_parent, *c_label;
  Key key;
  int64_t local_idx;
  int lane;
  int lane_delta;            // lattice: vertex offset of this lane's direction slot (lanes 0..5)
  bool track_ids;
};

struct FRep {
  double t, wsum;  // wsum = Σ α_c·n_c (the birth part of the reference's λ)
  uint64_t it, nlog;
  uint32_t n, S, next_id, cur_chunk, used_w, since_rebuild;
  int bott_id, snap_idx, status;
  uint32_t n_births, n_deaths, n_migr, n_disp, n_exc;
  bool log_on;
};

// Neighbourhood of site s in "slots": lattice slot = direction bit (z-1, y-1, x-1, x+1, y+1, z+1),
// CSR slot = position in the row (degree <= 32 in this mode).  One gather: the site's own clone and
// the clones on all neighbours are loaded together.
struct FHood {
  uint32_t own;       // clone on s (0 = empty)
  uint32_t present;   // slots that exist
  uint32_t empty;     // slots whose neighbour is empty
  uint32_t cw[6];     // lattice: clone per direction (statically indexed: registers)
  uint32_t cwl[32];   // CSR: clone per row position (dynamically indexed: local memory)
};

__device__ __forceinline__ uint32_t f_slot_site(const FCtx& X, uint32_t s, uint32_t slot) {
  const DevGraph& g = X.A->g;
  if (g.kind == 0) {
    const int plane = (int)(g.n1 * g.n2), n1 = (int)g.n1;
    const int d = slot == 0 ? -plane : slot == 1 ? -n1 : slot == 2 ? -1 : slot == 3 ? 1 : slot == 4 ? n1 : plane;
    return (uint32_t)((int)s + d);
  }
  return g.colidx[g.rowptr[s] + slot];
}

__device__ __forceinline__ void f_gather(const FCtx& X, uint32_t s, FHood& h) {
  const DevGraph& g = X.A->g;
  h.own = X.clone[s];
  h.empty = 0;
  if (g.kind == 0) {
    const uint32_t mask = lattice_mask(g, s);
    const int plane = (int)(g.n1 * g.n2), n1 = (int)g.n1;
    const int delta[6] = {-plane, -n1, -1, 1, n1, plane};
#pragma unroll
    for (int b = 0; b < 6; ++b) h.cw[b] = ((mask >> b) & 1u) ? (uint32_t)X.clone[(uint32_t)((int)s + delta[b])] : 0u;
#pragma unroll
    for (int b = 0; b < 6; ++b)
      if (((mask >> b) & 1u) && h.cw[b] == 0) h.empty |= 1u << b;
    h.present = mask;
    return;
  }
  const uint32_t row0 = g.rowptr[s];
  uint32_t deg = g.rowptr[s + 1] - row0;
  if (deg > 32) deg = 32;
  h.present = deg == 32 ? FULL : ((1u << deg) - 1u);
  for (uint32_t base = 0; base < deg; base += 8) {
    uint32_t nb[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) nb[q] = base + q < deg ? g.colidx[row0 + base + q] : 0u;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (base + q < deg) {
        const uint32_t c = X.clone[nb[q]];
        h.cwl[base + q] = c;
        if (c == 0) h.empty |= 1u << (base + q);
      }
  }
}

// may a cell of clone c divide onto a neighbour holding clone cw (0 = empty)?  :767-786
__device__ __forceinline__ bool f_admissible(const FCtx& X, uint32_t c, uint32_t cw) {
  if (cw == 0) return true;
  const int model = X.P->rule;
  if (model == JSB_RULE_CONTACT) return false;
  if (model == JSB_RULE_VOTER) return cw != c;
  if (model == JSB_RULE_HVOTER) return !(X.c_alpha[c - 1] <= X.c_alpha[cw - 1]);
  return true;
}

// slots a cell of clone h.own may divide onto
__device__ __forceinline__ uint32_t f_adm_mask(const FCtx& X, const FHood& h) {
  if (h.own == 0) return 0;  // an empty site divides nowhere (and has no fitness to compare)
  if (X.P->rule == JSB_RULE_CONTACT) return h.empty;
  uint32_t m = 0;
  if (X.A->g.kind == 0) {
#pragma unroll
    for (int b = 0; b < 6; ++b)
      if (((h.present >> b) & 1u) && f_admissible(X, h.own, h.cw[b])) m |= 1u << b;
  } else {
    for (uint32_t b = 0; b < 32 && ((h.present >> b) & 1u); ++b)
      if (f_admissible(X, h.own, h.cwl[b])) m |= 1u << b;
  }
  return m;
}

// rate of a site from its gathered neighbourhood
__device__ __forceinline__ double f_rate(const FCtx& X, const FHood& h, uint32_t adm_mask, double& bpart) {
  bpart = 0.0;
  if (h.own == 0) return 0.0;
  const uint32_t deg = (uint32_t)__popc(h.present);
  double mpart = 0.0;
  if (deg > 0) {
    bpart = (X.c_alpha[h.own - 1] * (double)__popc(adm_mask)) * kInvDeg[deg];
    mpart = (X.P->rate_migration * (double)__popc(h.empty)) * kInvDeg[deg];
  }
  return bpart + X.P->rate_death + mpart;
}

// add `delta` to a tree entry: native shared-memory reduction when the level lives in shared memory
__device__ __forceinline__ void f_tree_add(double* p, bool in_smem, double delta) {
  if (in_smem) asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(smem_addr(p)), "d"(delta) : "memory");
  else atomicAdd(p, delta);
}

// Warp-wide view of the neighbourhood of site x: lane = slot (lattice direction / CSR row position,
// degree <= 31 in this mode), lane 31 = the site itself.  Clones and leaf rates are loaded in ONE
// round trip.
struct FWarpHood {
  uint32_t u;        // neighbour site of this lane's slot
  uint32_t cu;       // its clone (0 = empty)
  double leaf;       // its current rate
  bool present;      // the slot exists
  uint32_t own;      // clone on x (broadcast)
  double own_leaf;   // rate of x (broadcast)
  uint32_t deg;      // degree of x
};

__device__ __forceinline__ uint32_t f_site_degree(const FCtx& X, uint32_t u) {
  const DevGraph& g = X.A->g;
  if (g.kind == 0) return (uint32_t)__popc(lattice_mask(g, u));
  return g.rowptr[u + 1] - g.rowptr[u];
}

__device__ __forceinline__ FWarpHood f_warp_gather(const FCtx& X, uint32_t x) {
  const DevGraph& g = X.A->g;
  const int lane = X.lane;
  FWarpHood h;
  h.u = x;
  h.cu = 0;
  h.leaf = 0.0;
  if (g.kind == 0) {
    const uint32_t mask = lattice_mask(g, x);
    h.present = lane < 6 && ((mask >> lane) & 1u);
    h.u = (uint32_t)((int)x + X.lane_delta);
  } else {
    const uint32_t row0 = g.rowptr[x], deg = g.rowptr[x + 1] - row0;
    h.present = (uint32_t)lane < deg && lane < 31;
    if (h.present) h.u = g.colidx[row0 + lane];
  }
  uint32_t c = 0;
  double lf = 0.0;
  if (h.present) { c = X.clone[h.u]; lf = X.lvl[0][h.u]; }
  if (lane == 31) { c = X.clone[x]; lf = X.lvl[0][x]; }
  h.own = __shfl_sync(FULL, c, 31);
  h.own_leaf = __shfl_sync(FULL, lf, 31);
  if (h.present) { h.cu = c; h.leaf = lf; }
  h.deg = (uint32_t)__popc(__ballot_sync(FULL, h.present));
  return h;
}

// add `delta` to every ancestor of site s
__device__ __forceinline__ void f_push_up(const FCtx& X, uint32_t s, double delta) {
  const bool in_smem = X.A->fast_smem_tree != 0;
  uint32_t i = s;
#pragma unroll
  for (int k = 1; k < kFastLevels; ++k) {
    if (k <= X.top) {
      i >>= 5;
      f_tree_add(&X.lvl[k][i], in_smem, delta);
    }
  }
}

// The clone on site x changed from a to b (0 = empty); clone[x] == b is already stored and h is a
// fresh gather of x's neighbourhood.  Each occupied neighbour u moves by
//     Δp_u = (α_u·(A(c_u,b) − A(c_u,a)) + m·([b=0] − [a=0])) / deg_u
// (one lane per neighbour, no further loads), x itself is recomputed from the gathered clones.
__device__ __forceinline__ void f_apply_change(const FCtx& X, uint32_t x, uint32_t a, uint32_t b, const FWarpHood& h) {
  const int lane = X.lane;
  const double mig = X.P->rate_migration;
  const bool contact = X.P->rule == JSB_RULE_CONTACT;
  if (h.present && h.cu != 0) {
    const int demp = (int)(b == 0) - (int)(a == 0);
    const int dadm = contact ? demp : (int)f_admissible(X, h.cu, b) - (int)f_admissible(X, h.cu, a);
    if (dadm != 0 || demp != 0) {
      const double delta = (X.c_alpha[h.cu - 1] * (double)dadm + mig * (double)demp) * kInvDeg[f_site_degree(X, h.u)];
      X.lvl[0][h.u] = h.leaf + delta;
      f_push_up(X, h.u, delta);
    }
  }
  const uint32_t emp = (uint32_t)__popc(__ballot_sync(FULL, h.present && h.cu == 0));
  const uint32_t adm = contact ? emp : (uint32_t)__popc(__ballot_sync(FULL, h.present && b != 0 && f_admissible(X, b, h.cu)));
  if (lane == 31) {
    double p = 0.0;
    if (b != 0) {
      p = X.P->rate_death;
      if (h.deg > 0) p = (X.c_alpha[b - 1] * (double)adm) * kInvDeg[h.deg] + X.P->rate_death + (mig * (double)emp) * kInvDeg[h.deg];
    }
    X.lvl[0][x] = p;
    const double delta = p - h.own_leaf;
    if (delta != 0.0) f_push_up(X, x, delta);
  }
  __syncwarp();
}

// debug (JSB_QUIRK_DEBUG_VERIFY_TREE): every leaf around x must equal the rate recomputed from scratch
__device__ __noinline__ bool f_verify_around(const FCtx& X, uint32_t x) {
  const FWarpHood h = f_warp_gather(X, x);
  bool bad = false;
  if (h.present || X.lane == 31) {
    const uint32_t s = X.lane == 31 ? x : h.u;
    FHood hh;
    f_gather(X, s, hh);
    double bp;
    const double want = f_rate(X, hh, f_adm_mask(X, hh), bp);
    const double have = X.lvl[0][s];
    bad = fabs(want - have) > 1e-9 * (1.0 + fabs(want));
  }
  return __any_sync(FULL, bad);
}

// warp reduction of one value per lane (fp64)
__device__ __forceinline__ double f_warp_sum(double v) {
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
  return v;
}

// upper levels from the leaves (exact re-summation); four rows of 32 children in flight
__device__ __noinline__ void f_rebuild_upper(const FCtx& X) {
  for (int k = 1; k <= X.top; ++k) {
    const uint32_t nk = X.lvl_n[k], nc = X.lvl_n[k - 1];
    const double* child = X.lvl[k - 1];
    for (uint32_t i = 0; i < nk; i += 4) {
      double v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t c = (i + q) * 32u + (uint32_t)X.lane;
        v[q] = c < nc ? child[c] : 0.0;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] += __shfl_xor_sync(FULL, v[q], d);
      }
      if (X.lane == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (i + q < nk) X.lvl[k][i + q] = v[q];
      }
    }
    __syncwarp();
  }
}

// every leaf from the lattice state, then the upper levels
__device__ __noinline__ void f_rebuild_all(const FCtx& X) {
  const uint32_t nv = X.A->g.nv;
  for (uint32_t s = X.lane; s < nv; s += 32) {
    FHood h;
    f_gather(X, s, h);
    double b;
    X.lvl[0][s] = f_rate(X, h, f_adm_mask(X, h), b);
  }
  __syncwarp();
  f_rebuild_upper(X);
}

// select the site whose cumulative rate interval contains `target`; resid = offset inside the site.
// The 32-wide prefix scans run in fp32 (a relative 1e-7 perturbation of the selection probabilities,
// far below what any ensemble statistic resolves; fp64 adds cost ~6x the latency on this part);
// the tree itself and the total rate stay fp64.
__device__ __forceinline__ bool f_select(const FCtx& X, double target64, uint32_t& site, double& resid) {
  uint32_t idx = 0;
  float target = (float)target64;
#pragma unroll
  for (int k = kFastLevels - 2; k >= 0; --k) {
    if (k >= X.top && X.top > 0) continue;  // start below the root
    const uint32_t base = idx * 32u;
    const uint32_t c = base + (uint32_t)X.lane;
    const double val64 = c < X.lvl_n[k] ? X.lvl[k][c] : 0.0;
    const float val = (float)val64;
    float incl = val;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float o = __shfl_up_sync(FULL, incl, d);
      if (X.lane >= d) incl += o;
    }
    uint32_t m = __ballot_sync(FULL, (incl > target) && (val64 > 0.0));
    int j;
    if (m) {
      j = __ffs(m) - 1;
    } else {  // rounding pushed the target past the last child: take the last non-empty one
      m = __ballot_sync(FULL, val64 > 0.0);
      if (m == 0) return false;  // stale sum over an empty subtree: caller re-sums and retries
      j = 31 - __clz(m);
    }
    const float vj = __shfl_sync(FULL, val, j);
    const float ej = __shfl_sync(FULL, incl, j) - vj;
    target -= ej;
    if (target < 0.0f) target = 0.0f;
    if (!(target < vj)) target = vj * 0.999999f;
    idx = base + (uint32_t)j;
  }
  site = idx;
  resid = (double)target;
  return true;
}

__device__ __forceinline__ void f_log(const FCtx& X, FRep& r, uint32_t type, uint32_t flags, double time, uint32_t cell,
                                      uint32_t parent, uint32_t site0, uint32_t site2_1, uint32_t clone, uint32_t label) {
  if (r.log_on) {
    const uint32_t slot = (uint32_t)(r.nlog % kLogChunk);
    if (slot != 0) {  // inside the current chunk: two 16-byte stores, no call
      if (X.lane < 2) {
        jsb_event* dst = X.A->log_pool + (size_t)r.cur_chunk * kLogChunk + slot;
        log_store(dst, X.lane, time, cell, parent, site0 + 1, site2_1, clone, label, type, flags);
      }
    } else {          // first record of a chunk: claim one from the pool
      const uint32_t c = log_emit(X.A, X.local_idx, X.lane, r.nlog, r.cur_chunk, type, flags, time, cell, parent,
                                  site0 + 1, site2_1, clone, label);
      if (c == kNone) { r.log_on = false; r.status = JSB_ST_LOG_OVERFLOW; }
      else r.cur_chunk = c;
    }
  }
  r.nlog++;
}

// remove the cell of clone c on site v (death, displacement, excision)
__device__ __forceinline__ void f_kill(const FCtx& X, FRep& r, uint32_t v, uint32_t c, double time, uint32_t flags) {
  const uint32_t cid = X.track_ids ? X.cellid[v] : 0u;
  __syncwarp();
  f_log(X, r, JSB_EV_DEATH, flags, time, cid, 0, v, 0, c, 0);
  if (X.lane == 0) X.clone[v] = 0;
  __syncwarp();
  r.wsum -= X.c_alpha[c - 1];
}

// cell_birth (:320-416 / :419-537) without the rejection test; v divides onto w
__device__ __forceinline__ bool f_birth(const FCtx& X, FRep& r, uint32_t v, uint32_t w, uint32_t sp, uint32_t occ,
                                     bool& mutant_stays) {
  const DevPoint& P = *X.P;
  mutant_stays = false;
  uint64_t wa, wb;
  draw_block_nl(X.key, r.it, 0, DS_DRIVER_COIN, 0, wa, wb);
  const double rr = u01_53(wa);
  int child_node = -1;
  bool driver;
  if (P.method == JSB_METHOD_TREE) {
    if (rr <= P.mu) {
      const int last = (int)X.c_label[sp - 1] - 1;
      for (int ed = 0; ed < P.n_edges && child_node < 0; ++ed) {
        if (X.edge_parent[ed] != last) continue;
        const int cand = X.edge_child[ed];
        bool present = false;
        for (uint32_t s0 = X.lane; s0 < r.S; s0 += 32)
          if (X.c_parent[s0] == sp && (int)X.c_label[s0] - 1 == cand) present = true;
        present = __any_sync(FULL, present);
        if (!present) child_node = cand;
      }
    }
    driver = child_node >= 0;
  } else {
    driver = !(rr > P.mu);
  }
  if (driver && r.S >= (uint32_t)kMaxClones) { r.status = JSB_ST_CLONE_CAPACITY; return false; }
  uint32_t first_flag = JSB_EVF_GROUP_START;
  if (occ != 0) {
    f_kill(X, r, w, occ, r.t, JSB_EVF_DISPLACED | first_flag);
    first_flag = 0;
    r.n_disp++;
    r.n -= 1;
  }
  const uint32_t nid = r.next_id, nid2 = r.next_id + 1;
  r.next_id += 2;
  const uint32_t parent = X.track_ids ? X.cellid[v] : 0u;
  __syncwarp();
  if (!driver) {
    f_log(X, r, JSB_EV_DUPLICATE, first_flag, r.t, nid, parent, v, 0, sp, 0);
    f_log(X, r, JSB_EV_DUPLICATE, 0, r.t, nid2, parent, w, 0, sp, 0);
    if (X.lane == 0) {
      if (X.track_ids) { X.cellid[v] = nid; X.cellid[w] = nid2; }
      X.clone[w] = (uint16_t)sp;
    }
    r.wsum += X.c_alpha[sp - 1];
  } else {
    uint32_t label;
    double nalpha;
    if (P.method == JSB_METHOD_TREE) {
      label = (uint32_t)child_node + 1;
      nalpha = X.node_rate[child_node];
    } else {
      uint64_t a, b;
      draw_block_nl(X.key, r.it, 0, DS_LABEL, 0, a, b);
      const uint32_t j = bounded(X.key, a, 1000u - r.S, r.it, DS_LABEL, 0, 0);
      label = pick_label(r.used_w, j, X.lane);
      if ((uint32_t)X.lane == (label >> 5)) r.used_w |= 1u << (label & 31u);
      nalpha = X.c_alpha[sp - 1] + fabs(P.adv_mean + P.adv_std * contract_randn(X.key, r.it));
    }
    const uint32_t nsp = r.S + 1;
    const bool coin = u01_53(wb) < 0.5;
    f_log(X, r, JSB_EV_MUTATION, first_flag, r.t, nid2, parent, coin ? w : v, 0, nsp, label);
    f_log(X, r, JSB_EV_DUPLICATE, 0, r.t, nid, parent, coin ? v : w, 0, sp, 0);
    if (X.lane == 0) {
      X.c_alpha[r.S] = nalpha;
      X.c_parent[r.S] = (uint16_t)sp;
      X.c_label[r.S] = (uint16_t)label;
      if (coin) {  // mutant on the new site
        if (X.track_ids) { X.cellid[v] = nid; X.cellid[w] = nid2; }
        X.clone[w] = (uint16_t)nsp;
      } else {     // mutant stays, the wild-type daughter moves out
        if (X.track_ids) { X.cellid[w] = nid; X.cellid[v] = nid2; }
        X.clone[w] = (uint16_t)sp;
        X.clone[v] = (uint16_t)nsp;
      }
    }
    mutant_stays = !coin;
    r.S = nsp;
    r.wsum += nalpha;
  }
  __syncwarp();
  r.n += 1;
  r.n_births++;
  return true;
}

// excision: k = trunc(n·ratio) cells uniformly without replacement, deaths logged at tb (:813-838)
__device__ __noinline__ void f_excision(const FCtx& X, FRep& r, double tb, double ratio) {
  const uint32_t nv = X.A->g.nv;
  const int lane = X.lane;
  long long kk = (long long)trunc((double)r.n * ratio);
  if (kk < 0) kk = 0;
  if (kk > (long long)r.n) kk = (long long)r.n;
  const uint32_t k = (uint32_t)kk;
  // alive sites, ascending, into inds[]
  uint32_t cnt = 0;
  for (uint32_t base = 0; base < nv; base += 32) {
    const uint32_t s = base + lane;
    const bool alive = s < nv && X.clone[s] != 0;
    const uint32_t m = __ballot_sync(FULL, alive);
    if (alive) X.inds[cnt + __popc(m & ((1u << lane) - 1u))] = s;
    cnt += __popc(m);
  }
  __syncwarp();
  // partial Fisher-Yates: k uniform draws without replacement
  for (uint32_t base = 0; base < k; base += 32) {
    const uint32_t i = base + lane;
    uint32_t jm = 0;
    if (i < k) {
      uint64_t a, b;
      draw_block(X.key, r.it, 0, DS_EXCISION, i, a, b);
      jm = i + bounded(X.key, a, cnt - i, r.it, DS_EXCISION, i, 0);
    }
    const uint32_t c2 = (k - base) < 32u ? (k - base) : 32u;
    for (uint32_t l = 0; l < c2; ++l) {
      const uint32_t j = __shfl_sync(FULL, jm, (int)l);
      if (lane == 0) {
        const uint32_t ii = base + l;
        const uint32_t tj = X.inds[j], ti = X.inds[ii];
        X.inds[j] = ti;
        X.inds[ii] = tj;
      }
    }
    __syncwarp();
  }
  for (uint32_t i = 0; i < k; ++i)
    f_kill(X, r, X.inds[i], (uint32_t)X.clone[X.inds[i]], tb, JSB_EVF_EXCISION | (i == 0 ? JSB_EVF_GROUP_START : 0));
  r.n -= k;
  r.n_exc += k;
  __syncwarp();
  f_rebuild_all(X);
  r.since_rebuild = 0;
}

__device__ __noinline__ void f_seed(const FCtx& X, FRep& r) {
  const DevPoint& P = *X.P;
  const DevGraph& g = X.A->g;
  const int lane = X.lane;
  for (uint32_t i = lane; i < (g.nv + 1) / 2; i += 32) reinterpret_cast<uint32_t*>(X.clone)[i] = 0;
  for (uint32_t i = lane; i < g.nv; i += 32) X.lvl[0][i] = 0.0;
  for (int k = 1; k <= X.top; ++k)
    for (uint32_t i = lane; i < X.lvl_n[k]; i += 32) X.lvl[k][i] = 0.0;
  __syncwarp();
  uint32_t n0 = 0;
  if (P.n_start == 1) {
    if (X.A->n_cand > 0) {
      uint32_t j = 0;
      if (X.A->n_cand > 1) {
        uint64_t a, b;
        draw_block_nl(X.key, 0, 0, DS_SEED_TIE, 0, a, b);
        j = bounded(X.key, a, X.A->n_cand, 0, DS_SEED_TIE, 0, 0);
      }
      const uint32_t v0 = X.A->seed_cand[j];
      if (lane == 0) { X.clone[v0] = 1; if (X.track_ids) X.cellid[v0] = 1; }
      n0 = 1;
    }
  } else if (P.n_start > 1) {
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int i = 0; i < P.n_start; ++i) {
        uint64_t a, b;
        draw_block_nl(X.key, 0, 0, DS_SEED_POS, (uint32_t)i, a, b);
        uint32_t pos = bounded(X.key, a, g.nv - (uint32_t)i, 0, DS_SEED_POS, (uint32_t)i, 0);
        uint32_t ins = 0;
        for (uint32_t q = 0; q < cnt; ++q)
          if (X.samp[q] <= pos) { ++pos; ins = q + 1; }
        for (uint32_t q = cnt; q > ins; --q) X.samp[q] = X.samp[q - 1];
        X.samp[ins] = pos;
        ++cnt;
        X.clone[pos] = 1;
        if (X.track_ids) X.cellid[pos] = (uint32_t)i + 1;
      }
    }
    n0 = (uint32_t)P.n_start;
  }
  __syncwarp();
  const double a0 = (P.method == JSB_METHOD_TREE) ? P.root_rate : P.rate_birth;
  if (lane == 0) {
    X.c_alpha[0] = a0;
    X.c_parent[0] = 0;
    X.c_label[0] = (P.method == JSB_METHOD_TREE) ? (uint16_t)(P.root_node + 1) : (uint16_t)1;
  }
  __syncwarp();
  r.n = n0;
  r.next_id = n0 + 1;
  r.S = n0 > 0 ? 1 : 0;
  r.wsum = a0 * (double)n0;
  r.used_w = 0;
  if (lane == 0) r.used_w = 0x3u;
  if (lane == 31) r.used_w = 0xFFFFFE00u;
  if (n0 > 0) f_rebuild_all(X);
  r.since_rebuild = 0;
}

__device__ __noinline__ void f_outputs(const FCtx& X, const FRep& r) {
  const RunArgs& A = *X.A;
  const int lane = X.lane;
  const uint32_t nv = A.g.nv;
  __syncwarp();
  if (lane == 0) {
    jsb_summary s;
    s.n_iterations = (uint64_t)r.n_births + r.n_deaths + r.n_migr;  // state-changing events only
    s.n_events = r.nlog;
    s.t_final = r.t;
    s.n_effective = r.n_births + r.n_deaths + r.n_migr + (r.n_exc ? 1u : 0u);
    s.n_births = r.n_births;
    s.n_deaths = r.n_deaths;
    s.n_migrations = r.n_migr;
    s.n_displaced = r.n_disp;
    s.n_excised = r.n_exc;
    s.n_alive = r.n;
    s.n_clones = r.S;
    s.next_cell_id = r.next_id;
    s.status = r.status;
    A.summaries[X.local_idx] = s;
  }
  // clone sizes are not tracked during the run: count them on the final lattice
  for (uint32_t i = lane; i < r.S; i += 32) X.c_cnt[i] = 0;
  __syncwarp();
  for (uint32_t v = lane; v < nv; v += 32) {
    const uint32_t cl = X.clone[v];
    if (cl) atomicAdd(&X.c_cnt[cl - 1], 1u);
  }
  __syncwarp();
  long long off = -1;
  if (A.clone_pool) {
    unsigned long long o = 0;
    if (lane == 0) o = atomicAdd(A.clone_pool_next, (unsigned long long)r.S);
    o = __shfl_sync(FULL, o, 0);
    if (o + r.S <= A.clone_pool_cap) {
      off = (long long)o;
      for (uint32_t i = lane; i < r.S; i += 32) {
        jsb_clone c;
        c.alpha = X.c_alpha[i];
        c.count = X.c_cnt[i];
        c.parent = X.c_parent[i];
        c.label = X.c_label[i];
        A.clone_pool[o + i] = c;
      }
    }
    if (lane == 0) A.clone_off[X.local_idx] = off;
  }
  if (A.out_clone) {
    uint16_t* oc = A.out_clone + (size_t)X.local_idx * nv;
    uint32_t* oi = A.out_cell + (size_t)X.local_idx * nv;
    for (uint32_t v = lane; v < nv; v += 32) {
      const uint16_t cl = X.clone[v];
      oc[v] = cl;
      oi[v] = cl ? X.cellid[v] : 0u;
    }
  }
  __syncwarp();
}

// apply the state-changing event that fires on site v (resid = offset inside v's rate interval)
__device__ __forceinline__ bool f_event(const FCtx& X, FRep& r, uint32_t v, double resid) {
  const DevPoint& P = *X.P;
  const int lane = X.lane;
  const FWarpHood hv = f_warp_gather(X, v);  // round trip: v, its neighbours, their rates
  const uint32_t c = hv.own;
  const uint32_t emp_mask = __ballot_sync(FULL, hv.present && hv.cu == 0);
  const uint32_t adm_mask = P.rule == JSB_RULE_CONTACT ? emp_mask : __ballot_sync(FULL, hv.present && f_admissible(X, c, hv.cu));
  const double bpart = (X.c_alpha[c - 1] * (double)__popc(adm_mask)) * kInvDeg[hv.deg];
  int kind = resid < bpart ? 0 : (resid < bpart + P.rate_death ? 1 : 2);  // birth | death | migration
  if (kind == 0 && adm_mask == 0) kind = 1;  // rounding at a part boundary
  if (kind == 2 && emp_mask == 0) kind = 1;
  bool ok = true;
  uint32_t w = v;
  if (kind == 1) {  // cell_death :574-586
    f_kill(X, r, v, c, r.t, JSB_EVF_GROUP_START);
    r.n -= 1;
    r.n_deaths++;
    f_apply_change(X, v, c, 0, hv);
  } else {
    const uint32_t pool = kind == 0 ? adm_mask : emp_mask;
    uint32_t want = 0;
    if (__popc(pool) > 1) {  // a second Philox block only when there is a choice
      uint64_t wc, wd;
      draw_block(X.key, r.it, 0, DS_F_KIND_NBR, 0, wc, wd);
      want = bounded(X.key, wd, (uint32_t)__popc(pool), r.it, DS_F_KIND_NBR, 0, 1);
    }
    uint32_t m = pool;
    for (uint32_t q = 0; q < want; ++q) m &= m - 1;  // the want-th admissible / empty neighbour
    const int slot = __ffs(m) - 1;
    w = __shfl_sync(FULL, hv.u, slot);
    const uint32_t occ = __shfl_sync(FULL, hv.cu, slot);
    if (kind == 0) {
      bool mutant_stays;
      ok = f_birth(X, r, v, w, c, occ, mutant_stays);  // stores the new clones of w (and v)
      if (ok) {
        // w: occ -> its new clone.  If the mutant stayed on v, v still counts as clone c here and
        // changes in a second step, against a fresh gather (w has changed meanwhile).
        const uint32_t cw_new = X.clone[w];
        const uint32_t cv_new = mutant_stays ? (uint32_t)X.clone[v] : c;
        if (mutant_stays && lane == 0) X.clone[v] = (uint16_t)c;
        __syncwarp();
        const FWarpHood hw = f_warp_gather(X, w);
        f_apply_change(X, w, occ, cw_new, hw);
        if (mutant_stays) {
          if (lane == 0) X.clone[v] = (uint16_t)cv_new;
          __syncwarp();
          const FWarpHood hv2 = f_warp_gather(X, v);
          f_apply_change(X, v, c, cv_new, hv2);
        }
      }
    } else {  // migration_cell :592-613
      const uint32_t cid = X.track_ids ? X.cellid[v] : 0u;
      __syncwarp();
      if (lane == 0) X.clone[v] = 0;
      __syncwarp();
      f_apply_change(X, v, c, 0, hv);  // w is empty in hv: untouched by this step
      if (lane == 0) {
        X.clone[w] = (uint16_t)c;
        if (X.track_ids) X.cellid[w] = cid;
      }
      __syncwarp();
      f_log(X, r, JSB_EV_MIGRATION, JSB_EVF_GROUP_START, r.t, cid, 0, w, v + 1, c, 0);
      r.n_migr++;
      const FWarpHood hw = f_warp_gather(X, w);
      f_apply_change(X, w, 0, c, hw);
    }
  }
  r.since_rebuild++;
  if (ok && (P.quirks & JSB_QUIRK_DEBUG_VERIFY_TREE)) {
    FCtx xc = X;
    if (f_verify_around(xc, v) || f_verify_around(xc, w)) { r.status = JSB_ST_INTERNAL; ok = false; }
  }
  return ok;
}

// time of the null iteration that crosses tau: first point of a rate-nu process after tau,
// conditioned to fall before t_new
__device__ __forceinline__ double f_null_crossing(double tau, double t_new, double nu, double u) {
  if (!(nu > 0.0)) return tau;
  const double x = tau + (-log1p(-u * (1.0 - exp(-nu * (t_new - tau))))) / nu;
  return x < t_new ? x : tau;
}

__device__ __forceinline__ void f_run(const FCtx& X) {
  const RunArgs& A = *X.A;
  const DevPoint& P = *X.P;
  const int lane = X.lane;
  const uint32_t nv = A.g.nv;
  const bool intent = (P.quirks & JSB_QUIRK_EXCISION_INTENT) != 0;
  const double dm = P.rate_death + P.rate_migration;
  FRep r;
  r.t = 0.0; r.wsum = 0.0; r.it = 0; r.nlog = 0; r.cur_chunk = 0;
  // which excision entry can fire: the reference's method 1 never advances its index (only entry 1),
  // method 2 advances it every iteration (in effect only the last entry); intent: one after another
  r.bott_id = P.n_bott > 0 ? ((intent || P.method != JSB_METHOD_TREE) ? 1 : P.n_bott) : 0;
  r.snap_idx = 0;
  r.status = JSB_ST_OK;
  r.n_births = r.n_deaths = r.n_migr = r.n_disp = r.n_exc = 0;
  r.log_on = (A.flags & JSB_OUT_EVENTS) != 0;
  { FCtx xc = X; FRep rc = r; f_seed(xc, rc); r = rc; }

#ifdef JSB_PROFILE
  long long fp[6] = {0, 0, 0, 0, 0, 0}, ft0 = clock64();
#define FPROF(i) do { long long _t = clock64(); fp[i] += _t - ft0; ft0 = _t; } while (0)
#else
#define FPROF(i)
#endif
  uint32_t n_resum = 0;
  while (r.t < P.Tf && r.n > 0) {  // loop test :687
    if (P.max_iter && r.it >= P.max_iter) { r.status = JSB_ST_MAX_ITER; break; }
    FPROF(5);
    if (r.since_rebuild >= kFastRebuild) {
      FCtx xc = X;
      if (++n_resum % kFastRefreshLeaves == 0) f_rebuild_all(xc); else f_rebuild_upper(xc);
      r.since_rebuild = 0;
    }
    FPROF(0);
    const double Lam = X.lvl[X.top][0];  // the root of the sum tree
    if (!(Lam > 0.0)) {  // nothing can change any more: the reference idles through null events to Tf
      r.t = P.Tf;
      break;
    }
    r.it += 1;
    uint64_t wa, wb;
    draw_block(X.key, r.it, 0, DS_F_TIME_SELECT, 0, wa, wb);
    // (2) statistical mode: the library log is enough (no bit-exactness contract here)
    const double t_new = r.t + (-log(u01_open(wa))) / Lam;
    const double nu = fmax(r.wsum + dm * (double)r.n - Lam, 0.0);  // rate of the reference's null iterations
    FPROF(1);

    // an excision time inside (t, t_new]?
    bool excise_after = false;
    double tbv = 0.0, ratio = 0.0;
    int next_bott = r.bott_id;
    if (r.bott_id > 0) {
      tbv = X.tb[r.bott_id - 1];
      if (tbv > r.t && tbv <= t_new) {
        ratio = X.ratio[r.bott_id - 1];
        next_bott = intent ? ((r.bott_id < P.n_bott) ? r.bott_id + 1 : 0) : 0;
        uint64_t we, wf;
        draw_block_nl(X.key, r.it, 0, DS_F_BOUNDARY, 0, we, wf);
        if (u01_53(we) < exp(-nu * (t_new - tbv))) {
          excise_after = true;  // the event itself is the iteration that crosses tb: event, then excision
        } else {
          // a null iteration crosses tb: excision on the current state; the clock restarts there
          r.t = f_null_crossing(tbv, t_new, nu, u01_53(wf));
          r.bott_id = next_bott;
          { FCtx xc = X; FRep rc = r; f_excision(xc, rc, tbv, ratio); r = rc; }
          continue;
        }
      }
    }
    // Tf inside (t, t_new]: the event happens only if it is the iteration that crosses Tf
    if (!excise_after && !(t_new < P.Tf)) {
      uint64_t we, wf;
      draw_block_nl(X.key, r.it, 0, DS_F_BOUNDARY, 1, we, wf);
      if (!(u01_53(we) < exp(-nu * (t_new - P.Tf)))) {
        r.t = f_null_crossing(P.Tf, t_new, nu, u01_53(wf));
        if (r.t < P.Tf) r.t = P.Tf;
        break;
      }
    }
    // snapshots are taken before the event of the crossing iteration (:708-714)
    while (r.snap_idx < P.n_snap && t_new > X.tsnap[r.snap_idx]) {
      if (lane == 0 && A.snaps) {
        jsb_snapshot s;
        s.time = t_new; s.iteration = r.it; s.n_events_before = r.nlog; s.n_alive = r.n;
        s.index = (uint32_t)r.snap_idx;
        A.snaps[(size_t)X.local_idx * A.max_snap + r.snap_idx] = s;
      }
      r.snap_idx++;
    }
    uint32_t v;
    double resid;
    if (!f_select(X, u01_53(wb) * Lam, v, resid)) {  // stale sums over an emptied subtree: re-sum, retry
      { FCtx xc = X; f_rebuild_upper(xc); }
      r.since_rebuild = 0;
      continue;  // (r.it moved on: the retry uses fresh draws; a pending excision is decided again)
    }
    FPROF(2);
    r.t = t_new;
    if (!f_event(X, r, v, resid)) break;
    FPROF(3);
    if (excise_after) {
      r.bott_id = next_bott;
      { FCtx xc = X; FRep rc = r; f_excision(xc, rc, tbv, ratio); r = rc; }
    }
  }
#ifdef JSB_PROFILE
  if (lane == 0 && (X.local_idx < 2 || r.n_births > 60000))
    printf("fast rep %lld events %u S %u n %u | rebuild %lld draws %lld select %lld event %lld other %lld\n", (long long)X.local_idx,
           r.n_births + r.n_deaths + r.n_migr, r.S, r.n, fp[0], fp[1], fp[2], fp[3], fp[5]);
#endif
  if (r.status == JSB_ST_OK && P.rate_death == 0.0 && r.n == nv && !(r.t < P.Tf)) r.status = JSB_ST_NONTERMINATING;
  { FCtx xc = X; FRep rc = r; f_outputs(xc, rc); }
}

// (a 255-register instantiation for blocks of at most eight warps, as for ssa_kernel, changes nothing here: 1.26 s either way)
__global__ void __launch_bounds__(384, 1) fast_kernel(const __grid_constant__ RunArgs A) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint64_t slot = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  const WsLayout& L = A.L;
  uint8_t* base = A.ws + slot * A.ws_stride;
  unsigned char* sw = s_dyn + (size_t)warp * A.smem_per_warp;
  FCtx X;
  X.A = &A;
  X.lane = lane;
  X.clone = reinterpret_cast<uint16_t*>(base + L.clone);
  X.cellid = reinterpret_cast<uint32_t*>(base + L.cellid);
  X.inds = reinterpret_cast<uint32_t*>(base + L.inds);
  X.samp = reinterpret_cast<uint32_t*>(base + L.samp);
  X.c_cnt = reinterpret_cast<uint32_t*>(base + L.c_cnt);
  X.c_parent = reinterpret_cast<uint16_t*>(base + L.c_parent);
  X.c_label = reinterpret_cast<uint16_t*>(base + L.c_label);
  DevPoint* sP = reinterpret_cast<DevPoint*>(sw);
  constexpr size_t kFastFixed = ((sizeof(DevPoint) + 15) & ~size_t(15)) + sizeof(double) * kTab;
  X.c_alpha = reinterpret_cast<double*>(sw + ((sizeof(DevPoint) + 15) & ~size_t(15)));  // α per clone: shared
  // tree levels: leaves in the (otherwise unused) list arena, upper levels in shared memory when they fit
  X.lvl[0] = reinterpret_cast<double*>(base + L.arena);
  X.lvl_n[0] = A.g.nv;
  double* up = A.fast_smem_tree ? reinterpret_cast<double*>(sw + kFastFixed)
                                : reinterpret_cast<double*>(base + L.ftree);
  int k = 0;
  while (X.lvl_n[k] > 1 && k + 1 < kFastLevels) {
    X.lvl_n[k + 1] = (X.lvl_n[k] + 31) / 32;
    X.lvl[k + 1] = up;
    up += X.lvl_n[k + 1];
    ++k;
  }
  X.top = k;
  for (int q = k + 1; q < kFastLevels; ++q) { X.lvl[q] = nullptr; X.lvl_n[q] = 0; }
  {
    const int plane = (int)(A.g.n1 * A.g.n2), n1 = (int)A.g.n1;
    X.lane_delta = lane == 0 ? -plane : lane == 1 ? -n1 : lane == 2 ? -1 : lane == 3 ? 1 : lane == 4 ? n1 : lane == 5 ? plane : 0;
  }
  X.key.k0 = (uint32_t)A.seed ^ 0xFA57FA57u;  // a stream family of its own
  X.key.k1 = (uint32_t)(A.seed >> 32);
  X.track_ids = (A.flags & (JSB_OUT_EVENTS | JSB_OUT_LATTICE)) != 0;

  for (;;) {
    unsigned long long idx = 0;
    if (lane == 0) idx = atomicAdd(A.work_counter, 1ull);
    idx = __shfl_sync(FULL, idx, 0);
    if ((long long)idx >= A.n_local) break;
    const int64_t gidx = A.g_begin + (int64_t)idx;
    X.local_idx = (int64_t)idx;
    X.key.stream = (uint32_t)gidx;
    {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(A.points + (gidx / A.reps_per_point));
      uint32_t* dst = reinterpret_cast<uint32_t*>(sP);
      __syncwarp();
      for (uint32_t i = lane; i < sizeof(DevPoint) / 4; i += 32) dst[i] = src[i];
      __syncwarp();
    }
    const DevPoint* P = sP;
    X.P = P;
    X.tb = A.dpool + P->bott_off;
    X.ratio = X.tb + P->n_bott;
    X.tsnap = A.dpool + P->snap_off;
    X.node_rate = A.dpool + P->rate_off;
    X.edge_parent = A.ipool + P->edge_off;
    X.edge_child = X.edge_parent + P->n_edges;
    f_run(X);
  }
}

static size_t fast_upper_entries(uint32_t nv) {
  size_t total = 0;
  uint32_t n = nv;
  while (n > 1) { n = (n + 31) / 32; total += n; }
  return total;
}

int fast_prepare(RunArgs& args, int max_warps_per_sm, int* warps_per_block) {
  int dev = 0;
  cudaGetDevice(&dev);
  int max_optin = 0;
  if (cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return -1;
  const size_t point_bytes = ((sizeof(DevPoint) + 15) & ~size_t(15)) + sizeof(double) * kTab;  // point + α table
  const size_t tree_bytes = 8 * fast_upper_entries(args.g.nv);
  int cap = 12;
  if (max_warps_per_sm > 0 && max_warps_per_sm < cap) cap = max_warps_per_sm;
  size_t per_warp = (point_bytes + tree_bytes + 15) & ~size_t(15);
  bool smem_tree = per_warp * 4 <= (size_t)max_optin;  // at least 4 warps per SM with the tree in smem
  if (!smem_tree) per_warp = point_bytes;
  int w = (int)((size_t)max_optin / per_warp);
  if (w > cap) w = cap;
  if (w < 1) return -1;
  args.fast_smem_tree = smem_tree ? 1u : 0u;
  args.smem_per_warp = (uint32_t)per_warp;
  if (cudaFuncSetAttribute(fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * w)) != cudaSuccess) return -1;
  {  // no more carve-out than needed: the rest of the 256 KB is L1 (see ssa_prepare)
    int sm_max = 0;
    cudaDeviceGetAttribute(&sm_max, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    int pct = sm_max > 0 ? (int)(((long long)(per_warp * w + 1024) * 100) / sm_max) : 100;
    if (pct > 100) pct = 100;
    cudaFuncSetAttribute(fast_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  }
  *warps_per_block = w;
  return 0;
}

int launch_fast(const RunArgs& args, int grid_blocks, int warps_per_block, void* stream) {
  fast_kernel<<<grid_blocks, warps_per_block * 32, (size_t)args.smem_per_warp * warps_per_block, (cudaStream_t)stream>>>(args);
  return (int)cudaGetLastError();
}

}  // namespace jsb
