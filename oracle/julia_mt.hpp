// julia_mt.hpp — the reference's OWN random stream, restated (TEST INFRASTRUCTURE ONLY).
//
// The reference threads one `MersenneTwister(seed)` (src/Start.jl:12-13) through every stage.  This file
// restates, from the published algorithms and the Julia 1.7.3 standard library (the version pinned by the
// reference's Manifest.toml:3), exactly what that object does for the calls on the dynamics path, so that
// the oracle can be run on the reference's own stream and compared with the one output of this path the
// reference ships: fileoutput/DriverList.csv (src/Start.jl:149-161).  Nothing here is reachable from the
// product; the CUDA path and the draw contract do not use it.
//
//   dSFMT-19937 (Saito & Matsumoto, dSFMT 2.2.x: do_recursion, init_by_array, period certification)
//   Random.MersenneTwister: the 1002-double `vals` cache, the 501-UInt128 `ints` cache and its refill
//       mt_setfull!(r, ::Type{<:BitInteger}) of Julia 1.6+ (stdlib/Random/src/RNGs.jl; pinned by the manual's
//       bitrand(MersenneTwister(1234), 10)), mt_pop! for UInt64 / UInt128
//   rand(r) = pop - 1.0; rand(r, UInt52Raw()) = bits of the popped double
//   rand(r, a:b) for one draw: SamplerRangeFast (RNGs.jl: `Sampler(::Type{MersenneTwister}, r, ::Val{1})`);
//       inside rand(::Dict/::Set): SamplerRangeNDL on UInt64 draws (generation.jl, `Val(Inf)`)
//   randn / randexp: the ziggurat of stdlib/Random/src/normal.jl (256 strips; tables regenerated here by the
//       published construction, randmtzig.c `create_ziggurat_tables`)
//   rand(r, ::Set{Int}): uniform slot of the underlying Dict until a filled one (generation.jl), with the
//       Dict's slot layout (base/dict.jl: hashindex, linear probing, rehash!, sizehint!) and
//       hash(::Int64) = hash_64_64(3|x| + bits(Float64(x))) (base/float.jl)
//   UUIDs.uuid1(rng): one rand(rng, UInt128)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace jmt {

typedef unsigned __int128 u128;

// ---- knobs for the (few) places where the recollection of the 1.7.3 sources is uncertain; the pinning
// script tries the alternatives and reports which combination reproduces DriverList.csv -------------------
struct Variant {
  int set_slots = 0;        // slots of Set(1:1000): 0 = follow sizehint!/growth rules (sizehint_mode below), else forced
  int sizehint_mode = 2;    // union!'s sizehint!(s, n): 2 = reserves cld(3n, 2), 1 = reserves n, 0 = no sizehint! call
  bool hash_float = false;  // hash(::Int64): hash_64_64(x) (false; hash(1) == 0x5bca7c69b794f8ce as the Julia 1.x REPL prints) or the older base/float.jl form (true)
  bool range_fast = true;   // single rand(rng, a:b): SamplerRangeFast (true) or SamplerRangeNDL (false)
  bool set_ndl = true;      // slot draws inside rand(::Set): NDL on UInt64 (true) or Fast on UInt52Raw (false)
  int int_fill_tail = 2;    // doubles taken from the scalar cache at the end of a big array fill
  int refill_pad = 0;         // pinning aid: scalar doubles consumed by an integer-cache refill, relative to the model
  int test_float_start = -1;  // pinning aid: >= 0 -> the scalar doubles are raw[test_float_start...] contiguously and the
                              // integer cache draws from a separate generator (never touches the scalar stream)
};

// ---- dSFMT-19937 ------------------------------------------------------------------------------------------
struct DSFMT {
  static constexpr int N = 191, POS1 = 117, SL1 = 19, SR = 12;
  static constexpr uint64_t MSK1 = 0x000ffafffffffb3fULL, MSK2 = 0x000ffdfffc90fffdULL;
  static constexpr uint64_t FIX1 = 0x90014964b32f4329ULL, FIX2 = 0x3b8d12ac548a7c7aULL;
  static constexpr uint64_t PCV1 = 0x3d84e1ac0dc82880ULL;
  uint64_t st[N + 1][2];
  int pos = 0;  // next state word to regenerate

  static uint32_t f1(uint32_t x) { return (x ^ (x >> 27)) * 1664525u; }
  static uint32_t f2(uint32_t x) { return (x ^ (x >> 27)) * 1566083941u; }

  void init_by_array(const uint32_t* key, int key_length) {
    const int size = (N + 1) * 4, lag = 11, mid = (size - lag) / 2;
    uint32_t p[(N + 1) * 4];  // the state as 32-bit words (little endian: idxof(i) = i)
    std::memset(p, 0x8b, sizeof p);
    int count = key_length + 1 > size ? key_length + 1 : size;
    uint32_t r = f1(p[0] ^ p[mid % size] ^ p[(size - 1) % size]);
    p[mid % size] += r;
    r += (uint32_t)key_length;
    p[(mid + lag) % size] += r;
    p[0] = r;
    count--;
    int i = 1, j = 0;
    for (; j < count && j < key_length; ++j) {
      r = f1(p[i] ^ p[(i + mid) % size] ^ p[(i + size - 1) % size]);
      p[(i + mid) % size] += r;
      r += key[j] + (uint32_t)i;
      p[(i + mid + lag) % size] += r;
      p[i] = r;
      i = (i + 1) % size;
    }
    for (; j < count; ++j) {
      r = f1(p[i] ^ p[(i + mid) % size] ^ p[(i + size - 1) % size]);
      p[(i + mid) % size] += r;
      r += (uint32_t)i;
      p[(i + mid + lag) % size] += r;
      p[i] = r;
      i = (i + 1) % size;
    }
    for (j = 0; j < size; ++j) {
      r = f2(p[i] + p[(i + mid) % size] + p[(i + size - 1) % size]);
      p[(i + mid) % size] ^= r;
      r -= (uint32_t)i;
      p[(i + mid + lag) % size] ^= r;
      p[i] = r;
      i = (i + 1) % size;
    }
    std::memcpy(st, p, sizeof p);
    for (int k = 0; k < N; ++k)  // initial_mask
      for (int h = 0; h < 2; ++h) st[k][h] = (st[k][h] & 0x000FFFFFFFFFFFFFULL) | 0x3FF0000000000000ULL;
    // period certification
    uint64_t t0 = st[N][0] ^ FIX1, t1 = st[N][1] ^ FIX2;
    uint64_t inner = (t0 & PCV1) ^ (t1 & 1ULL);
    for (int s = 32; s > 0; s >>= 1) inner ^= inner >> s;
    if ((inner & 1) == 0) st[N][1] ^= 1;  // PCV2 & 1 == 1
    pos = 0;
  }

  // next 128-bit output = two doubles in [1, 2) (bit patterns), in generation order
  void next(uint64_t out[2]) {
    uint64_t* a = st[pos];
    uint64_t* b = st[(pos + POS1) % N];
    uint64_t* lung = st[N];
    const uint64_t t0 = a[0], t1 = a[1], L0 = lung[0], L1 = lung[1];
    lung[0] = (t0 << SL1) ^ (L1 >> 32) ^ (L1 << 32) ^ b[0];
    lung[1] = (t1 << SL1) ^ (L0 >> 32) ^ (L0 << 32) ^ b[1];
    a[0] = (lung[0] >> SR) ^ (lung[0] & MSK1) ^ t0;
    a[1] = (lung[1] >> SR) ^ (lung[1] & MSK2) ^ t1;
    out[0] = a[0];
    out[1] = a[1];
    pos = (pos + 1) % N;
  }
  // dsfmt_fill_array_close1_open2: n (even) doubles, as bit patterns
  void fill(uint64_t* dst, int n) {
    for (int k = 0; k < n; k += 2) next(dst + k);
  }
};

// ---- ziggurat tables (normal.jl's ki/wi/fi and ke/we/fe), regenerated in 50-digit arithmetic from the two
// published constants by make_zig_tables.py (the strip areas follow from them) ---------------------------------
#include "zig_tables.inc"
struct Zig {
  const uint64_t *ki = ZKI, *ke = ZKE;
  const double *wi = ZWI, *fi = ZFI, *we = ZWE, *fe = ZFE;
  static constexpr double NOR_R = 3.6541528853610087963519472518, NOR_INV_R = 1.0 / 3.6541528853610087963519472518;
  static constexpr double EXP_R = 7.6971174701310497140446280481;
};

// ---- Base.Dict slot layout for a Set{Int} built by Set(1:n) -------------------------------------------------
static inline uint64_t hash_64_64(uint64_t a) {
  a = ~a + (a << 21);
  a = a ^ (a >> 24);
  a = a + (a << 3) + (a << 8);
  a = a ^ (a >> 14);
  a = a + (a << 2) + (a << 4);
  a = a ^ (a >> 28);
  a = a + (a << 31);
  return a;
}
static inline uint64_t julia_hash_int(int64_t x, bool float_form) {
  if (!float_form) return hash_64_64((uint64_t)x);
  const double f = (double)x;
  uint64_t fb;
  std::memcpy(&fb, &f, 8);
  const uint64_t a = (uint64_t)(x < 0 ? -x : x);
  return hash_64_64(3 * a + fb);  // hx(a, b, h = 0)
}

struct IntSet {  // Dict{Int,Nothing} as far as slot positions go
  std::vector<uint8_t> slots;  // 0 empty, 1 filled, 2 deleted
  std::vector<int64_t> keys;
  int64_t count = 0, ndel = 0, maxprobe = 0;
  bool hash_float = true;
  static int64_t tablesz(int64_t x) {
    if (x < 16) return 16;
    int64_t s = 1;
    while (s < x) s <<= 1;
    return s;
  }
  explicit IntSet(bool hf) : slots(16, 0), keys(16, 0), hash_float(hf) {}
  int64_t hashindex(int64_t key, int64_t sz) const { return (int64_t)(julia_hash_int(key, hash_float) & (uint64_t)(sz - 1)) + 1; }
  void rehash(int64_t newsz) {
    newsz = tablesz(newsz);
    std::vector<uint8_t> olds = slots;
    std::vector<int64_t> oldk = keys;
    const int64_t sz = (int64_t)olds.size();
    slots.assign((size_t)newsz, 0);
    keys.assign((size_t)newsz, 0);
    int64_t cnt = 0, mp = 0;
    for (int64_t i = 1; i <= sz; ++i)
      if (olds[i - 1] == 1) {
        const int64_t k = oldk[i - 1];
        int64_t index0 = hashindex(k, newsz), index = index0;
        while (slots[index - 1] != 0) index = (index & (newsz - 1)) + 1;
        const int64_t probe = (index - index0) & (newsz - 1);
        if (probe > mp) mp = probe;
        slots[index - 1] = 1;
        keys[index - 1] = k;
        ++cnt;
      }
    count = cnt;
    ndel = 0;
    maxprobe = mp;
  }
  void sizehint(int64_t newsz, bool one_point_five) {
    const int64_t oldsz = (int64_t)slots.size();
    if (one_point_five) newsz = (3 * newsz + 1) / 2;
    if (newsz <= oldsz) return;
    newsz = std::max(newsz, (oldsz * 5) >> 2);
    rehash(newsz);
  }
  // ht_keyindex2!: > 0 found, < 0 insertion slot
  int64_t keyindex2(int64_t key) {
    for (;;) {
      const int64_t sz = (int64_t)keys.size();
      int64_t iter = 0, index = hashindex(key, sz), avail = 0;
      bool broke = false;
      for (;;) {
        if (slots[index - 1] == 0) return avail < 0 ? avail : -index;
        if (slots[index - 1] == 2) {
          if (avail == 0) avail = -index;
        } else if (keys[index - 1] == key) {
          return index;
        }
        index = (index & (sz - 1)) + 1;
        ++iter;
        if (iter > maxprobe) { broke = true; break; }
      }
      (void)broke;
      if (avail < 0) return avail;
      const int64_t maxallowed = std::max<int64_t>(16, sz >> 6);
      while (iter < maxallowed) {
        if (slots[index - 1] != 1) { maxprobe = iter; return -index; }
        index = (index & (sz - 1)) + 1;
        ++iter;
      }
      rehash(count > 64000 ? sz * 2 : sz * 4);
    }
  }
  void push(int64_t key) {
    int64_t index = keyindex2(key);
    if (index > 0) return;
    index = -index;
    slots[index - 1] = 1;
    keys[index - 1] = key;
    ++count;
    const int64_t sz = (int64_t)keys.size();
    if (ndel >= ((3 * sz) >> 2) || count * 3 > sz * 2) rehash(count > 64000 ? count * 2 : count * 4);
  }
  // Set(1:n): union!(Set{Int}(), 1:n) = sizehint!(s, n); push! each
  static IntSet range(int64_t n, const Variant& v) {
    IntSet s(v.hash_float);
    if (v.set_slots > 0) s.rehash(v.set_slots);
    else if (v.sizehint_mode > 0) s.sizehint(n, v.sizehint_mode == 2);
    for (int64_t k = 1; k <= n; ++k) s.push(k);
    return s;
  }
};

// ---- Random.MersenneTwister ---------------------------------------------------------------------------------
struct MersenneTwister {
  static constexpr int CACHE_F = 1002, CACHE_I_BYTES = 501 * 16;
  DSFMT ds;
  uint64_t vals[CACHE_F];
  u128 ints[501];
  int idxF = CACHE_F, idxI = 0;
  Variant var;
  // bookkeeping for the pinning script: position in the raw dSFMT output of the next scalar double
  int64_t raw_made = 0, vals_base = 0;
  int64_t raw_pos() const { return idxF == CACHE_F ? raw_made : vals_base + idxF; }
  std::vector<int64_t> randn_pos;  // raw position of the first double of every randn() call
  static const Zig& zig() { static const Zig z; return z; }

  MersenneTwister(uint64_t seed, const Variant& v) : var(v) {
    uint32_t key[2];
    int n = 0;
    do { key[n++] = (uint32_t)seed; seed >>= 32; } while (seed != 0);  // make_seed
    ds.init_by_array(key, n);
    if (v.test_float_start >= 0) {
      uint64_t tmp[2];
      for (int k = 0; k + 1 < v.test_float_start; k += 2) ds.next(tmp);
      raw_made = v.test_float_start & ~1;
      if (v.test_float_start & 1) { gen_rand(); idxF = 1; }
    }
  }
  // -- the vals cache: doubles in [1, 2), as bit patterns
  void gen_rand() { vals_base = raw_made; ds.fill(vals, CACHE_F); raw_made += CACHE_F; idxF = 0; }
  uint64_t pop_bits() {
    if (idxF == CACHE_F) gen_rand();
    return vals[idxF++];
  }
  double rand_float() {  // rand(rng) in [0, 1)
    const uint64_t b = pop_bits();
    double d;
    std::memcpy(&d, &b, 8);
    return d - 1.0;
  }
  uint64_t raw52() { return pop_bits(); }  // rand(rng, UInt52Raw()): the caller masks
  // -- rand!(r, A::UnsafeView{Float64}, CloseOpen12) on n doubles
  void fill_floats(uint64_t* A, int n) {
    const int n2 = (n - var.int_fill_tail) / 2 * 2;
    if (n2 < 382) {  // _rand_max383!: straight from the scalar cache
      for (int k = 0; k < n; ++k) A[k] = pop_bits();
      return;
    }
    ds.fill(A, n2);                                  // dsfmt_fill_array_close1_open2!, bypasses the cache
    raw_made += n2;
    for (int k = n2; k < n; ++k) A[k] = pop_bits();  // the tail comes from the scalar path
  }
  // -- mt_setfull!(r, ::Type{<:BitInteger}) of Julia 1.6+ (stdlib/Random/src/RNGs.jl): dSFMT randomizes 52 of the 64
  // bits of a word, so 627 UInt128 (1254 doubles, ONE fill_array! straight from the generator: the scalar cache is
  // not touched) are condensed into 501 fully random ones — word 501+j donates 12 bits to each half of words
  // 4j .. 4j+3 (shifts 48, 36, 24, 12), word 626 to word 500.  Pinned by the Julia manual's
  // `bitrand(MersenneTwister(1234), 10)` (tests/test_reference_pin.py).
  void fill_ints() {
    if (var.test_float_start >= 0) {  // pinning aid: content irrelevant, scalar stream untouched
      for (int k = 0; k < 501; ++k) ints[k] = ((u128)0x9E3779B97F4A7C15ULL * (u128)(k + 1 + raw_made)) << 17;
      idxI = CACHE_I_BYTES;
      return;
    }
    static thread_local uint64_t F[1254];
    ds.fill(F, 1254);
    raw_made += 1254;
    u128 A[627];
    for (int k = 0; k < 627; ++k) A[k] = (u128)F[2 * k] | ((u128)F[2 * k + 1] << 64);
    int k = 501, n = 0;
    while (n != 500) {
      const u128 u = A[k++];
      A[n++] ^= u << 48;
      A[n++] ^= u << 36;
      A[n++] ^= u << 24;
      A[n++] ^= u << 12;
    }
    A[500] ^= A[626] << 48;
    for (int q = 0; q < 501; ++q) ints[q] = A[q];
    idxI = CACHE_I_BYTES;
  }
  uint64_t rand_u64() {
    if ((idxI >> 3) < 1) fill_ints();
    idxI -= 8;
    const int i = idxI;
    const u128 x = ints[i >> 4];
    return (uint64_t)(x >> ((i & 15) * 8));
  }
  u128 rand_u128() {
    if ((idxI >> 4) < 1) fill_ints();
    const int idx = idxI >> 4;
    idxI = (idx << 4) - 16;
    return ints[idx - 1];
  }
  void uuid1() { (void)rand_u128(); }  // UUIDs.uuid1(rng): the value only seeds clock_seq/node bits

  // -- rand(rng, 1:n), 0-based result
  uint64_t range_ndl(uint64_t s) {
    uint64_t x = rand_u64();
    u128 m = (u128)x * s;
    uint64_t l = (uint64_t)m;
    if (l < s) {
      const uint64_t t = (0 - s) % s;
      while (l < t) {
        x = rand_u64();
        m = (u128)x * s;
        l = (uint64_t)m;
      }
    }
    return (uint64_t)(m >> 64);
  }
  uint64_t range_fast(uint64_t n) {  // SamplerRangeFast{UInt64}: m = n - 1
    const uint64_t m = n - 1;
    const unsigned bw = m ? 64u - (unsigned)__builtin_clzll(m) : 0u;
    const uint64_t mask = bw >= 64 ? ~0ull : ((1ull << bw) - 1ull);
    for (;;) {
      const uint64_t x = (bw <= 52 ? raw52() : rand_u64()) & mask;
      if (x <= m) return x;
    }
  }
  uint64_t rand_index(uint64_t n) { return var.range_fast ? range_fast(n) : range_ndl(n); }

  // -- randexp / randn (normal.jl)
  double randexp() {
    const Zig& z = zig();
    for (;;) {
      const uint64_t ri = raw52() & 0x000fffffffffffffULL;
      const unsigned idx = (unsigned)(ri & 0xFF);
      const double x = (double)ri * z.we[idx];
      if (ri < z.ke[idx]) return x;
      if (idx == 0) return Zig::EXP_R - std::log(rand_float());
      if ((z.fe[idx - 1] - z.fe[idx]) * rand_float() + z.fe[idx] < std::exp(-x)) return x;
    }
  }
  double randn() {
    const Zig& z = zig();
    randn_pos.push_back(raw_pos());
    for (;;) {
      const uint64_t r = raw52() & 0x000fffffffffffffULL;
      const int64_t rabs = (int64_t)(r >> 1);
      const unsigned idx = (unsigned)(rabs & 0xFF);
      const double x = (double)((r & 1) ? -rabs : rabs) * z.wi[idx];
      if ((uint64_t)rabs < z.ki[idx]) return x;
      if (idx == 0) {
        for (;;) {
          const double xx = -Zig::NOR_INV_R * std::log(rand_float());
          const double yy = -std::log(rand_float());
          if (yy + yy > xx * xx) return ((rabs >> 8) & 1) ? -Zig::NOR_R - xx : Zig::NOR_R + xx;
        }
      }
      if ((z.fi[idx - 1] - z.fi[idx]) * rand_float() + z.fi[idx] < std::exp(-0.5 * x * x)) return x;
    }
  }
  // -- rand(rng, s::Set{Int}) where s = Set(1:n) minus the keys flagged in `removed` (1-based flags)
  int64_t rand_set(const IntSet& s, const bool* removed) {
    const uint64_t sz = (uint64_t)s.slots.size();
    for (;;) {
      const uint64_t i = var.set_ndl ? range_ndl(sz) : range_fast(sz);  // 0-based slot
      if (s.slots[i] == 1 && !removed[s.keys[i]]) return s.keys[i];
    }
  }
};

}  // namespace jmt
