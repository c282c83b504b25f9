// psum_plan.h -- host-side packing and pass planning for the Pauli-sum engine (no CUDA here).
//
// Replaces the reference's container logic only in the sense of "pack once":
//   WeightedPauliOperator.simplify (legacy/weighted_pauli_operator.py:341-400)  -> merge_terms
// Everything else (grouping by x-mask, free-bit tiling, pattern compaction) has no reference
// counterpart: the reference materialises a CSR matrix instead (legacy/op_converter.py:120-123).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <unordered_map>
#include <vector>

namespace psum {

static inline int popc64(uint64_t v) { return __builtin_popcountll(v); }
static inline uint64_t low_mask(int n) { return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }

// compress the bits of v selected by mask (ascending) into the low bits
static inline uint64_t pext64(uint64_t v, uint64_t mask) {
  uint64_t out = 0;
  int j = 0;
  for (uint64_t m = mask; m; m &= m - 1, ++j) {
    int b = __builtin_ctzll(m);
    out |= ((v >> b) & 1ull) << j;
  }
  return out;
}

struct MergedTerm {  // one label Pauli after merging duplicates; c = coeff * (-i)^ypow
  uint64_t x, z;
  double re, im;
};

struct CompTerm {  // real-valued component term of one shard/part: value * (im ? i : 1)
  uint64_t xl, zl;  // local x / z masks
  double val;
  uint8_t im;
};

// ---- device-facing records (layout shared with the kernels) --------------------------------
// A pass keeps 2^RB amplitudes per thread in registers (RB = 3 or 4 "register bits" = tile bits
// 5 .. 5+RB-1).  x-groups of a pass whose masks differ only in register bits form a BUNDLE: the
// 2^RB partner amplitudes a thread needs are the same set for every member (the register part of
// the mask is a permutation inside the thread), so a bundle costs ONE set of partner loads.
struct BundleRec {    // 16 bytes
  uint32_t xo;        // tile-local xor with the register bits cleared; bit31 = remote flag
  uint32_t meta;      // bits 0..7 member count, bits 8..15 hbit, bits 16..31 first member (group
                      // index inside the chunk).  hbit = 1 + tile bit (warp-uniform, i.e. >= 5+RB)
                      // of the highest set bit of xo for local bundles, else 0: the Hermitian
                      // <psi|H|psi> kernels process such a bundle only for amplitudes with that bit
                      // clear and double it (pair i, i^x)
  uint64_t xlo;       // local x mask with the register QUBITS cleared (partner base of remote bundles)
};
static_assert(sizeof(BundleRec) == 16, "BundleRec layout");

struct GroupRec {     // 48 bytes
  uint32_t xt;        // tile-local xor (x gathered onto the tile index); bit31 = remote flag
  uint32_t meta;      // bits 0..15 first pattern (relative to the chunk), 16..18 entry count,
                      // bit 19 = slot 1 of a fast group is imaginary,
                      // bit 20 = has imaginary patterns, bit 21 = flat (one real class-0 entry),
                      // bits 22..27 = variant: v < NA      fast, slot-1 class 0,  xr = v
                      //                        NA <= v < 2NA fast, slot-1 class = xr = v - NA
                      //                        63          general (class entry lists)
                      //   (xr = register bits of xt; "fast" = <= 2 patterns, slot 0 real class 0)
                      // bits 28..31 = hbit of the group's bundle (kept for the plan checker)
  uint32_t zt0, zt1;  // in-tile z-patterns of the first two patterns (inline copy)
  double cf0, cf1;    // their folded coefficients: written in SHARED memory by the fold
  uint64_t xl;        // full local x mask (partner index = i ^ xl)
  uint16_t ent[4];    // non-empty classes: (cls << 11) | count ; cls = im*NA + ((zt >> 5) & (NA-1))
  // variant 62 (quadratic group): bytes 8..31 (zt0 .. cf1) hold instead u8 seg[17] (pattern offsets of
  // the segments, seg[nseg] = pattern count), then at byte 30 a u16 offset (in doubles) of the
  // group's tables inside the chunk's quad arena; ent[] is unused
};
static_assert(sizeof(GroupRec) == 48, "GroupRec layout");

struct PatRec {       // 16 bytes, staged in shared memory per chunk (cf is written by the fold)
  double cf;
  uint32_t zt;
  uint32_t pad;
};

struct ChunkRec {  // 64 bytes
  int32_t g0;  // first group
  int32_t ng;  // groups in the chunk
  int32_t p0, np;
  int32_t b0, nb;  // bundles of the chunk: they cover the groups [ns, ng) in bundle order
  uint16_t ns;     // the first ns groups are "singles": local fast groups that are alone in their
                   // bundle.  They need no bundle record and are sorted by s = im*NA + xr so that each
                   // (im, xr) runs its own straight-line loop with a compile-time register permutation
  uint16_t nq;     // quadratic groups in the chunk (their tables are rebuilt per tile)
  uint32_t q0;     // bits 0..30: their group indices (inside the chunk) are quad_idx[q0 .. q0 + nq);
                   // bit 31: some pattern of the chunk sums more than kLongTerms terms (warp-cooperative fold)
  uint8_t s_cnt[32];  // singles per s
};
static_assert(sizeof(ChunkRec) == 64, "ChunkRec layout");

constexpr int kMaxTileBits = 13;
constexpr int kRegBitLo = 5;     // tile bits 5 .. 5+RB-1 live in registers (2^RB amplitudes per thread)
constexpr int kVarGeneral = 63;
constexpr int kVarQuad = 62;
constexpr int kLongTerms = 24;   // a pattern with more terms is folded by a whole warp
// "Quadratic" groups: every in-tile z-pattern is real and touches at most two tile bits (ZZ couplings,
// Z X Z decorations, ...).  D(lane, w, r) of such a group is a quadratic form in the signs of the tile
// bits; it separates into small tables over the lane bits and over the warp/round bits that are
// rebuilt per tile, so a thread evaluates ~14 table reads instead of one record per pattern:
//   class {}      A0  = T_L[lane] + T_W[w] + sum_j s_j(w) C_j[lane]        (j over the warp/round bits)
//   class {i}     A_i = TL_i[lane] + TW_i[w]                               (i over the register bits)
//   class {i,i'}  K[mask]
// Table layout of one group (doubles): (1 + nW + RB) lane tables of 32, (1 + RB) w tables of 2^nW,
// K[2^RB].  The group's patterns are sorted into segments, one per table (seg offsets in the record):
//   0: lane-only (incl. the constant)  1: w-only  2+j: {lane bit, w bit j}  2+nW+i: {reg i} (+ lane bit)
//   2+nW+RB+i: {w bit, reg i}  2+nW+2RB: {reg i, reg i'}
constexpr int kQuadMinPatterns = 6;
#ifndef PSUM_QUAD_DOUBLES
#define PSUM_QUAD_DOUBLES 2048
#endif
constexpr int kQuadDoubles = PSUM_QUAD_DOUBLES;   // shared-memory arena of the quad tables of a chunk
inline int quad_table_doubles(int L, int rb) {
  const int nW = L - kRegBitLo - rb;
  return (1 + nW + rb) * 32 + (1 + rb) * (1 << nW) + (1 << rb);
}
#ifndef PSUM_PAT_CHUNK
#define PSUM_PAT_CHUNK 768
#endif
#ifndef PSUM_GRP_CHUNK
#define PSUM_GRP_CHUNK 128
#endif
constexpr int kPatChunk = PSUM_PAT_CHUNK;  // patterns staged in shared memory at a time
constexpr int kGrpChunk = PSUM_GRP_CHUNK;  // groups staged in shared memory at a time

// Staging capacities of one chunk.  A pass runs either as two CTAs per SM (tiles up to 2^12, the
// small capacities, two rounds per tile) or as ONE big CTA per SM (twice the threads, one round
// per tile up to 2^12, most of the SM's shared memory for the plan): dense passes -- many groups,
// quadratic-group tables -- then fit one chunk, so the fold and the table build run once per tile.
// The pattern / group capacities are compile-time constants of the kernel configuration (static
// shared-memory offsets); only the quad arena, which sits behind the tile, has a run-time size.
constexpr int kPatCapSmall = PSUM_PAT_CHUNK, kGrpCapSmall = PSUM_GRP_CHUNK;
constexpr int kPatCapBig = 2048, kGrpCapBig = 192;   // <= 255 groups: per-(im, xr) single counters are bytes
struct ChunkCaps {
  int pat = kPatCapSmall, grp = kGrpCapSmall, quad = kQuadDoubles;
  bool big = false;
};
inline ChunkCaps small_caps() { return ChunkCaps(); }
inline ChunkCaps big_caps(int L) {   // the CTA's shared memory (tile + staging) must stay below 227 KB
  ChunkCaps c;
  c.pat = kPatCapBig;
  c.grp = kGrpCapBig;
  c.quad = L >= 13 ? 4096 : 8192;
  c.big = true;
  return c;
}

struct PassHost {
  int L = 0;                   // tile bits
  int rb = 3;                  // register bits (2^rb amplitudes per thread)
  ChunkCaps caps;              // staging capacities this pass was chunked for (caps.big: one CTA per SM)
  uint64_t free_mask = 0;      // over local qubits
  uint8_t fbit[16] = {0};      // qubit of tile bit j: 0..4 lane bits, 5..5+rb-1 register bits, then warp/round
  bool has_im = false;         // any imaginary component pattern in this pass
  int hi_bit = 0;              // highest local qubit the pass combines (free set or any member x-mask)
  int64_t n_flat = 0;          // groups whose D is the same for all register amplitudes
  int64_t n_fast = 0;          // groups served by a fast variant (real, <= 2 inline patterns)
  std::vector<BundleRec> bundles;
  std::vector<uint16_t> quad_idx;   // group index (inside its chunk) of every quadratic group
  int64_t n_singles = 0;       // fast groups alone in their bundle (no bundle record)
  int64_t n_remote_bundles = 0, n_general = 0, n_general_patterns = 0;   // for the cost model
  int64_t n_quad = 0;          // quadratic groups (table evaluation)
  std::vector<GroupRec> groups;
  std::vector<uint16_t> pat_zt;
  std::vector<uint16_t> pat_inl;   // (group index in chunk) * 2 + slot for inlined patterns, else 0xffff
  std::vector<uint32_t> pat_tptr;  // size patterns+1
  std::vector<uint64_t> term_zb;
  std::vector<double> term_val;
  std::vector<ChunkRec> chunks;
  int64_t n_xgroups = 0, n_remote = 0;
};

struct SimpleTerm {  // for the streaming kernel A: complex coefficient per (x,z)
  uint64_t z;
  double re, im;
};
struct SimpleGroup {
  uint64_t xl;
  int32_t t0, nt;
};

struct PartHost {  // all terms whose global x-part is g
  int g = 0;
  std::vector<CompTerm> comp;          // component terms (rank sign folded in)
  std::vector<PassHost> passes;        // tiled plan (kernel B)
  std::vector<SimpleGroup> sgroups;    // streaming plan (kernel A)
  std::vector<SimpleTerm> sterms;
  int64_t n_xgroups = 0;
};

struct PlanOptions {
  int tile_bits = 12;
  int low_bits = 3;    // lowest qubits always free: 2^3 amplitudes = one 128-byte line per run
  int min_cover = 3;
  int plan_tries = 64; // randomised greedy restarts of the pass planner; the cheapest plan wins (64 tries
                       // take ~20 ms for 10 000 terms and find 12 passes for C4 where 8 tries find 14)
  int late_bits = 0;   // top local qubits kept out of the early passes (see choose_passes)
  int late_split = 0;  // 1: groups that touch ONE late qubit get passes that free only that late qubit, so
                       // their tiles complete chunk pair by chunk pair while the state arrives (run_ready)
  int reg_bits = 3;    // 3: 8 amplitudes per thread, 256-thread CTAs; 4: 16 per thread, 128-thread CTAs
                       // (4 per thread at 64 registers and twice the warps was measured: 378 vs 305 ms on C4)
  int share_weight = 100;  // percent: weight of the partner-load term when the register qubits are chosen
  int min_bundle = 2;  // a bundle of fewer members, all of them fast, is split into singles (own loads,
                       // straight-line loops): sharing only pays when enough members use one load set
  int cta_mode = 0;    // 0 auto (a pass that needs several chunks or holds quadratic groups runs as one
                       // big CTA per SM), 1 always two CTAs per SM, 2 always one big CTA
  int kernel = 0;  // 0 auto, 1 streaming only, 2 tiled
};

// -------------------------------------------------------------------------------------------
inline std::vector<MergedTerm> merge_terms(int64_t T, const uint64_t* x, const uint64_t* z,
                                           const uint8_t* ypow, const double* c) {
  std::vector<MergedTerm> out;
  std::unordered_map<uint64_t, std::vector<int>> index;  // hash(x,z) -> candidates
  out.reserve(T);
  for (int64_t k = 0; k < T; ++k) {
    double re = c[2 * k], im = c[2 * k + 1];
    // multiply by (-i)^ypow
    switch (ypow ? (ypow[k] & 3) : 0) {
      case 1: { double t = re; re = im; im = -t; } break;       // (a+bi)(-i) = b - ai
      case 2: re = -re; im = -im; break;
      case 3: { double t = re; re = -im; im = t; } break;       // (a+bi)(i) = -b + ai
      default: break;
    }
    uint64_t key = x[k] * 0x9E3779B97F4A7C15ull ^ (z[k] + 0x7F4A7C15ull + (x[k] << 6));
    auto& cand = index[key];
    bool found = false;
    for (int id : cand)
      if (out[id].x == x[k] && out[id].z == z[k]) {
        out[id].re += re;
        out[id].im += im;
        found = true;
        break;
      }
    if (!found) {
      cand.push_back((int)out.size());
      out.push_back({x[k], z[k], re, im});
    }
  }
  std::vector<MergedTerm> kept;
  kept.reserve(out.size());
  for (auto& t : out)
    if (t.re != 0.0 || t.im != 0.0) kept.push_back(t);  // chop(0.0): drop exact zeros
  return kept;
}

// Greedy choice of free-bit sets.  masks = distinct local x-masks of one part.
// Returns for each pass its free mask and the indices (into masks) it serves; groups that fit no
// pass ("remote": partner amplitudes read from global memory) are appended to the LAST pass.
struct PassChoice {
  uint64_t free_mask;
  std::vector<int> members;
};

inline std::vector<PassChoice> choose_passes(const std::vector<uint64_t>& masks, int n_loc, int L,
                                             int c, int min_cover, uint64_t seed = 0, int late_bits = 0,
                                             int late_split = 0) {
  std::vector<PassChoice> passes;
  uint64_t rng = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
  auto noise = [&]() {  // seed 0 -> deterministic greedy; otherwise scores are jittered by <= 40 %
    if (seed == 0) return 1.0;
    rng ^= rng << 13;
    rng ^= rng >> 7;
    rng ^= rng << 17;
    return 1.0 + 0.4 * (double)(rng >> 11) * (1.0 / 9007199254740992.0);
  };
  L = std::min(L, n_loc);
  c = std::min(c, L);
  const uint64_t low = low_mask(c);
  const uint64_t all = low_mask(n_loc);
  // "late" qubits: the top late_bits local qubits.  Stage 1 plans the groups that do not touch
  // them with free sets that avoid them, so those passes only ever combine amplitudes of one
  // 2^(n_loc - late_bits) chunk -- they can run chunk by chunk while the state is still arriving
  // from the host (psum_expect_host).  Stage 2 plans everything else.
  late_bits = std::max(0, std::min(late_bits, n_loc - L));
  const uint64_t late = all & ~low_mask(n_loc - late_bits);
  std::vector<int> remote, stage1, stage2;
  for (int i = 0; i < (int)masks.size(); ++i) {
    if (popc64(masks[i] & ~low) > L - c) remote.push_back(i);
    else if (masks[i] & late) stage2.push_back(i);
    else stage1.push_back(i);
  }
  auto fill_mask = [&](uint64_t S) {  // pad with the lowest unused bits up to L bits
    for (int b = 0; b < n_loc && popc64(S) < L; ++b) S |= (1ull << b);
    return S;
  };
  auto run_stage = [&](std::vector<int> uncovered, uint64_t allowed, std::vector<int>& leftover) {
    while (!uncovered.empty()) {
      uint64_t S = low;
      int remaining = L - c;
      while (remaining > 0) {
        double score[64];
        for (int b = 0; b < 64; ++b) score[b] = 0.0;
        bool any = false;
        for (int i : uncovered) {
          uint64_t need = masks[i] & ~S;
          int nn = popc64(need);
          if (nn == 0 || nn > remaining || (need & ~allowed)) continue;
          any = true;
          double w = 1.0 / (double)(nn * nn);
          for (uint64_t m = need; m; m &= m - 1) score[__builtin_ctzll(m)] += w;
        }
        if (!any) break;
        int best = -1;
        for (int b = 0; b < n_loc; ++b) score[b] *= noise();
        for (int b = 0; b < n_loc; ++b)
          if (!(S >> b & 1ull) && (allowed >> b & 1ull) && (best < 0 || score[b] > score[best])) best = b;
        if (best < 0 || score[best] <= 0.0) break;
        S |= 1ull << best;
        --remaining;
      }
      S = fill_mask(S) & all;
      PassChoice pc;
      pc.free_mask = S;
      std::vector<int> rest;
      for (int i : uncovered) {
        if ((masks[i] & ~S) == 0) pc.members.push_back(i);
        else rest.push_back(i);
      }
      if ((int)pc.members.size() < min_cover && !passes.empty()) {
        // not worth a pass of its own (a pass re-reads psi and read-modify-writes y)
        for (int i : uncovered) leftover.push_back(i);
        return;
      }
      passes.push_back(std::move(pc));
      uncovered.swap(rest);
    }
  };
  if (late) {
    std::vector<int> left1;
    run_stage(stage1, all & ~late, left1);
    for (int i : left1) stage2.push_back(i);
    if (late_split) {
      // stage 1.5: one late qubit at a time, lowest first (its chunk pairs complete earliest)
      for (int t = n_loc - late_bits; t < n_loc; ++t) {
        std::vector<int> mine, rest, leftt;
        for (int i : stage2) ((masks[i] & late) == (1ull << t) ? mine : rest).push_back(i);
        if (mine.empty()) continue;
        run_stage(mine, (all & ~late) | (1ull << t), leftt);
        stage2.swap(rest);
        for (int i : leftt) stage2.push_back(i);
      }
    }
  } else {
    for (int i : stage1) stage2.push_back(i);
    std::sort(stage2.begin(), stage2.end());
  }
  run_stage(stage2, all, remote);
  if (passes.empty()) {
    PassChoice pc;
    pc.free_mask = fill_mask(low) & all;
    passes.push_back(pc);
  }
  // groups no pass serves read their partners from global memory; they ride on the LAST pass so
  // that the earlier (possibly chunk-local) passes stay chunk-local
  for (int i : remote) passes.back().members.push_back(i);
  return passes;
}

// Build the device-facing arrays of one pass.
inline PassHost build_pass(const std::vector<CompTerm>& comp,
                           const std::vector<std::vector<int>>& terms_of_mask,
                           const std::vector<uint64_t>& masks, const PassChoice& pc, int n_loc, int rb,
                           const ChunkCaps caps = ChunkCaps(), int min_bundle = 2, int share_weight = 100) {
  const double share_w = 0.01 * share_weight;
  PassHost P;
  P.caps = caps;
  P.free_mask = pc.free_mask;
  P.L = popc64(pc.free_mask);
  P.rb = rb;
  const int NA = 1 << rb;
  const uint64_t S = pc.free_mask;
  // ---- roles of the free bits -------------------------------------------------------------
  // tile bits 0..2 = the three lowest free qubits (128-byte lines / conflict-free LDS.128);
  // register bits (tile bits 5..5+rb-1) = the rb free qubits that minimise the shared-memory
  // traffic of the pass: x-groups whose masks differ only in register qubits share one set of
  // partner loads (a bundle), and a group whose z-patterns avoid the register qubits has one D
  // for all amplitudes of a thread ("flat"); the rest ascending by how often x-masks flip them.
  std::vector<int> fb;
  for (int b = 0; b < n_loc; ++b)
    if (S >> b & 1ull) fb.push_back(b);
  {
    const int L = (int)fb.size();
    std::vector<int> cand(fb.begin() + std::min(3, L), fb.end());
    // distinct (im, z & S) per member group
    std::vector<std::vector<uint64_t>> zsets;
    zsets.reserve(pc.members.size());
    for (int mi : pc.members) {
      std::vector<uint64_t> zs;
      for (int t : terms_of_mask[mi]) zs.push_back(((comp[t].zl & S) << 1) | (uint64_t)comp[t].im);
      std::sort(zs.begin(), zs.end());
      zs.erase(std::unique(zs.begin(), zs.end()), zs.end());
      zsets.push_back(std::move(zs));
    }
    std::vector<int> reg;
    const int nreg = std::min<int>(rb, (int)cand.size());
    if ((int)cand.size() > nreg) {
      std::vector<int> pick(nreg), best;
      for (int i = 0; i < nreg; ++i) pick[i] = i;
      double best_cost = -1.0;
      std::vector<uint64_t> keys, cls;
      while (true) {
        uint64_t rm = 0;
        for (int i : pick) rm |= 1ull << cand[i];
        keys.clear();
        // cost of the pass in SM cycles per warp and round under this choice of register qubits
        // (shared-memory pipe: 1 wavefront / cycle; issue: 4 instructions / cycle):
        //   partner load set      4 NA wavefronts
        //   fast group            ~12   (two inline slots; slot-1 class must be 0 or the x-mask's register part)
        //   quadratic group       ~30   (separable tables; independent of the classes)
        //   general group         ~8 + 9 per class entry + 3 per pattern
        double gcost = 0.0;
        for (size_t m = 0; m < pc.members.size(); ++m) {
          const uint64_t xl = masks[pc.members[m]];
          keys.push_back(xl & ~rm);   // local and remote groups alike: equal outside the register qubits
          const std::vector<uint64_t>& zs = zsets[m];
          bool all_real = true, le2 = true;
          for (uint64_t zk : zs) {
            if (zk & 1ull) all_real = false;
            if (popc64(zk >> 1) > 2) le2 = false;
          }
          if (all_real && le2 && (int)zs.size() >= kQuadMinPatterns) {
            gcost += 30.0;
            continue;
          }
          cls.clear();
          for (uint64_t zk : zs) cls.push_back(((zk >> 1) & rm) << 1 | (zk & 1ull));
          bool fast = false;
          if (zs.size() <= 2) {
            // slot 0 must be a real class-0 pattern (or absent); slot 1: class 0 or the register part of x
            const uint64_t xr = xl & rm;
            int n_r0 = 0, n_ok1 = 0;
            for (uint64_t ck : cls) {
              if (ck == 0) ++n_r0;
              else if ((ck >> 1) == 0 || (ck >> 1) == xr) ++n_ok1;
            }
            fast = (zs.size() == 1) ? (n_r0 + n_ok1 == 1) : (n_r0 >= 1 && n_r0 + n_ok1 == 2);
          }
          if (fast) {
            gcost += 12.0;
          } else {
            std::sort(cls.begin(), cls.end());
            const double ncls = (double)(std::unique(cls.begin(), cls.end()) - cls.begin());
            gcost += 8.0 + 9.0 * ncls + 3.0 * (double)zs.size();
          }
        }
        std::sort(keys.begin(), keys.end());
        const double nb = (double)(std::unique(keys.begin(), keys.end()) - keys.begin());
        const double cost = share_w * 4.0 * NA * nb + gcost;
        if (best_cost < 0.0 || cost < best_cost) {
          best_cost = cost;
          best = pick;
        }
        int i = nreg - 1;
        while (i >= 0 && pick[i] == (int)cand.size() - nreg + i) --i;
        if (i < 0) break;
        ++pick[i];
        for (int j = i + 1; j < nreg; ++j) pick[j] = pick[j - 1] + 1;
      }
      for (int i : best) reg.push_back(cand[i]);
    } else {
      reg = cand;
    }
    std::sort(reg.begin(), reg.end());
    std::vector<int> rest;
    for (int b : cand)
      if (std::find(reg.begin(), reg.end(), b) == reg.end()) rest.push_back(b);
    // the qubits most often flipped by this pass's x-masks go to the warp-uniform tile bits,
    // where the Hermitian expectation kernels can skip the partner half of a bundle
    std::vector<double> xhits(64, 0.0);
    for (int mi : pc.members)
      for (uint64_t m = masks[mi] & S; m; m &= m - 1) xhits[__builtin_ctzll(m)] += 1.0;
    std::stable_sort(rest.begin(), rest.end(), [&](int a, int b) { return xhits[a] < xhits[b]; });
    int j = 0;
    for (int i = 0; i < std::min(3, L); ++i) P.fbit[j++] = (uint8_t)fb[i];
    size_t ri = 0;
    while (j < 5 && ri < rest.size()) P.fbit[j++] = (uint8_t)rest[ri++];
    for (int b : reg) P.fbit[j++] = (uint8_t)b;
    while (ri < rest.size()) P.fbit[j++] = (uint8_t)rest[ri++];
  }
  auto gather = [&](uint64_t v) {  // qubit mask -> tile-index mask
    uint32_t out = 0;
    for (int j = 0; j < P.L; ++j) out |= (uint32_t)((v >> P.fbit[j]) & 1ull) << j;
    return out;
  };
  const uint32_t reg_tmask = (uint32_t)(NA - 1) << kRegBitLo;
  uint64_t reg_qmask = 0;
  for (int j = kRegBitLo; j < kRegBitLo + rb && j < P.L; ++j) reg_qmask |= 1ull << P.fbit[j];
  struct Key {
    uint8_t im;
    uint32_t zt;
    int term;
  };
  struct PatTmp { int cls; uint32_t zt; std::vector<int> terms; };
  struct SubGroup { GroupRec G; std::vector<PatTmp> pats; int slot_of[2]; int var; bool remote; uint64_t xlo; };
  std::vector<SubGroup> subs;
  {
    uint64_t u = S;
    for (int mi : pc.members) u |= masks[mi];
    P.hi_bit = u ? 63 - __builtin_clzll(u) : 0;
  }
  for (int mi : pc.members) {
    const uint64_t xl = masks[mi];
    const bool remote = (xl & ~S) != 0;
    P.n_xgroups++;
    if (remote) P.n_remote++;
    std::vector<Key> keys;
    keys.reserve(terms_of_mask[mi].size());
    for (int t : terms_of_mask[mi])
      keys.push_back({comp[t].im, gather(comp[t].zl), t});
    auto cls_of = [&](const Key& k) { return (int)k.im * NA + (int)((k.zt >> kRegBitLo) & (uint32_t)(NA - 1)); };
    std::stable_sort(keys.begin(), keys.end(), [&](const Key& a, const Key& b) {
      int ca = cls_of(a), cb = cls_of(b);
      if (ca != cb) return ca < cb;
      return a.zt < b.zt;
    });
    // patterns = runs of equal (im, zt)
    std::vector<PatTmp> pats;
    for (int k = 0; k < (int)keys.size();) {
      int e = k;
      PatTmp pt{cls_of(keys[k]), keys[k].zt, {}};
      while (e < (int)keys.size() && keys[e].im == keys[k].im && keys[e].zt == keys[k].zt)
        pt.terms.push_back(keys[e++].term);
      pats.push_back(std::move(pt));
      k = e;
    }
    // quadratic group?  (all patterns real, at most two tile bits each)
    {
      bool quad = (int)pats.size() >= kQuadMinPatterns && pats.size() <= 255 && quad_table_doubles(P.L, rb) <= caps.quad &&
                  P.L - kRegBitLo - rb <= 5;   // a w table (2^nW entries) is filled by one warp, lane = index
      for (const PatTmp& pt : pats)
        if (pt.cls >= NA || __builtin_popcount(pt.zt) > 2) quad = false;
      if (quad) {
        const int nW = P.L - kRegBitLo - rb;
        auto seg_of = [&](uint32_t zt) {
          const uint32_t zl = zt & 31u, zr = (zt >> kRegBitLo) & (uint32_t)(NA - 1), zw = zt >> (kRegBitLo + rb);
          const int nl = __builtin_popcount(zl), nr = __builtin_popcount(zr), nw = __builtin_popcount(zw);
          if (nr == 0) {
            if (nw == 0) return 0;
            if (nl == 0) return 1;
            return 2 + __builtin_ctz(zw);
          }
          if (nr == 1) return (nw == 0 ? 2 + nW : 2 + nW + rb) + __builtin_ctz(zr);
          return 2 + nW + 2 * rb;
          (void)nl;
        };
        std::stable_sort(pats.begin(), pats.end(), [&](const PatTmp& a, const PatTmp& b) { return seg_of(a.zt) < seg_of(b.zt); });
        SubGroup sg;
        GroupRec& G = sg.G;
        std::memset(&G, 0, sizeof(G));
        G.xt = gather(xl) | (remote ? 0x80000000u : 0u);
        G.xl = xl;
        sg.remote = remote;
        sg.xlo = xl & ~reg_qmask;
        sg.var = kVarQuad;
        sg.slot_of[0] = 0;
        sg.slot_of[1] = 1;
        G.meta = (uint32_t)kVarQuad << 22;
        uint8_t* area = reinterpret_cast<uint8_t*>(&G) + 8;
        const int nseg = 2 + nW + 2 * rb + 1;
        size_t q = 0;
        for (int sgi = 0; sgi <= nseg; ++sgi) {
          while (q < pats.size() && seg_of(pats[q].zt) < sgi) ++q;
          area[sgi] = (uint8_t)q;
        }
        sg.pats = pats;
        P.n_quad++;
        subs.push_back(std::move(sg));
        continue;
      }
    }
    // sub-groups: each holds <= 4 non-empty classes, <= caps.pat patterns and <= 2047 patterns
    // per class.  Patterns are in class order, so cut greedily.
    size_t p = 0;
    while (p < pats.size()) {
      SubGroup sg;
      GroupRec& G = sg.G;
      std::memset(&G, 0, sizeof(G));
      G.xt = gather(xl) | (remote ? 0x80000000u : 0u);
      G.xl = xl;
      sg.remote = remote;
      sg.xlo = xl & ~reg_qmask;
      int count = 0, nent = 0;
      bool has_im = false;
      size_t q = p;
      while (q < pats.size() && count < caps.pat) {
        const int cls = pats[q].cls;
        if (nent == 0 || (G.ent[nent - 1] >> 11) != cls) {
          if (nent == 4) break;
          G.ent[nent++] = (uint16_t)(cls << 11);
        } else if ((G.ent[nent - 1] & 0x7ff) == 0x7ff) {
          break;
        }
        G.ent[nent - 1]++;
        if (cls >= NA) has_im = true;
        ++count;
        ++q;
      }
      const bool flat = (nent == 1) && (G.ent[0] >> 11) == 0;
      G.meta = ((uint32_t)nent << 16) | (has_im ? (1u << 20) : 0u) | (flat ? (1u << 21) : 0u);
      // fast variants: at most two patterns; slot 0 of the inline copy holds a real class-0 pattern
      // (or nothing), slot 1 the other one -- real or imaginary, its class 0 or equal to the
      // register part of the x-mask (the XX + YY pair groups, single X/Y-type terms)
      const int xr = (int)((G.xt & reg_tmask) >> kRegBitLo);
      const bool first_r0 = pats[p].cls == 0;
      const bool two_slot = count == 1 || (count == 2 && first_r0);
      sg.slot_of[0] = 0;
      sg.slot_of[1] = 1;
      sg.var = kVarGeneral;
      if (two_slot) {
        const int j1 = (count == 1 && first_r0) ? -1 : count - 1;   // pattern that goes to slot 1
        const int c1 = j1 >= 0 ? pats[p + j1].cls : 0;
        const int zr1 = c1 & (NA - 1);
        if (zr1 == 0 || zr1 == xr) {
          sg.var = (zr1 == 0 ? 0 : NA) + xr;
          if (count == 1 && !first_r0) sg.slot_of[0] = 1;
          for (int j = 0; j < count; ++j) {
            if (sg.slot_of[j] == 0) G.zt0 = pats[p + j].zt;
            else G.zt1 = pats[p + j].zt;
          }
          if (c1 >= NA) G.meta |= 1u << 19;                         // slot 1 is imaginary
        }
      }
      if (sg.var == kVarGeneral) {
        G.zt0 = pats[p].zt;
        G.zt1 = count > 1 ? pats[p + 1].zt : 0u;
      }
      G.meta |= (uint32_t)sg.var << 22;
      if (has_im) P.has_im = true;
      if (flat) P.n_flat++;
      if (sg.var != kVarGeneral) P.n_fast++;
      else {
        P.n_general++;
        P.n_general_patterns += count;
      }
      sg.pats.assign(pats.begin() + p, pats.begin() + q);
      subs.push_back(std::move(sg));
      p = q;
    }
  }
  // bundle order: local before remote, then by the mask outside the register qubits, then by variant
  std::stable_sort(subs.begin(), subs.end(), [](const SubGroup& a, const SubGroup& b) {
    if (a.remote != b.remote) return !a.remote;
    if (a.xlo != b.xlo) return a.xlo < b.xlo;
    return a.var < b.var;
  });
  // runs of equal (remote, xlo) = bundles
  struct Run { size_t s0, s1; int np, nq; };
  std::vector<Run> runs;
  const int qd = quad_table_doubles(P.L, rb);
  for (size_t si = 0; si < subs.size();) {
    size_t se = si;
    int np = 0, nq = 0;
    while (se < subs.size() && se - si < 255 && subs[se].remote == subs[si].remote && subs[se].xlo == subs[si].xlo &&
           np + (int)subs[se].pats.size() <= caps.pat && (int)(se - si) < caps.grp &&
           (nq + (subs[se].var == kVarQuad ? 1 : 0)) * qd <= caps.quad) {
      np += (int)subs[se].pats.size();
      nq += subs[se].var == kVarQuad ? 1 : 0;
      ++se;
    }
    runs.push_back({si, se, np, nq});
    si = se;
  }
  // chunks: consecutive runs up to the staging capacity.  Inside a chunk the singles (a local fast
  // group alone in its run) come first, sorted by (im, xr); the other runs follow with a bundle record each
  P.pat_tptr.push_back(0);
  size_t r0 = 0;
  int g0 = 0, p0 = 0;
  while (r0 < runs.size()) {
    size_t r1 = r0;
    int np = 0, ng = 0, nqc = 0;
    while (r1 < runs.size() && ng + (int)(runs[r1].s1 - runs[r1].s0) <= caps.grp && np + runs[r1].np <= caps.pat &&
           (nqc + runs[r1].nq) * qd <= caps.quad) {
      np += runs[r1].np;
      ng += (int)(runs[r1].s1 - runs[r1].s0);
      nqc += runs[r1].nq;
      ++r1;
    }
    ChunkRec cr;
    std::memset(&cr, 0, sizeof(cr));
    cr.b0 = (int32_t)P.bundles.size();
    cr.q0 = (uint32_t)P.quad_idx.size();
    auto s_index = [&](const SubGroup& sg) { return (int)((sg.G.meta >> 19) & 1u) * NA + (sg.var & (NA - 1)); };
    std::vector<size_t> singles, multi;
    // a run whose members are all local fast groups and that is shorter than min_bundle is emitted as
    // singles (each member loads its own partners); runs are then referenced by (run, member) pairs
    struct Single { size_t r, q; };
    std::vector<Single> single_members;
    for (size_t r = r0; r < r1; ++r) {
      bool all_fast = !subs[runs[r].s0].remote;
      for (size_t q = runs[r].s0; q < runs[r].s1; ++q)
        if (subs[q].var >= 2 * NA) all_fast = false;
      if (all_fast && (int)(runs[r].s1 - runs[r].s0) < std::max(2, min_bundle)) {
        for (size_t q = runs[r].s0; q < runs[r].s1; ++q) single_members.push_back({r, q});
      } else {
        multi.push_back(r);
      }
    }
    for (const Single& sm : single_members) singles.push_back(sm.q);
    std::stable_sort(singles.begin(), singles.end(), [&](size_t a, size_t b) {
      return s_index(subs[a]) < s_index(subs[b]);
    });
    int pl = 0, gl = 0;
    int nq_emitted = 0;
    bool has_long = false;
    auto emit_group = [&](SubGroup& sg, uint32_t hbit) {
      sg.G.meta |= (uint32_t)pl | (hbit << 28);
      const bool is_quad = sg.var == kVarQuad;
      if (is_quad) {
        const uint16_t off = (uint16_t)(nq_emitted * qd);
        std::memcpy(reinterpret_cast<uint8_t*>(&sg.G) + 30, &off, 2);
        P.quad_idx.push_back((uint16_t)gl);
        ++nq_emitted;
      }
      for (size_t j = 0; j < sg.pats.size(); ++j) {
        if ((int)sg.pats[j].terms.size() > kLongTerms) has_long = true;
        // bit 15 of the stored pattern marks patterns of a Hermitian-paired bundle (doubled by the
        // fold of the Hermitian expectation kernels, masked off by every fold)
        P.pat_zt.push_back((uint16_t)(sg.pats[j].zt | (hbit ? 0x8000u : 0u)));
        P.pat_inl.push_back(j < 2 && !is_quad ? (uint16_t)(gl * 2 + sg.slot_of[j]) : (uint16_t)0xffff);
        for (int t : sg.pats[j].terms) {
          P.term_zb.push_back(comp[t].zl & ~S);
          P.term_val.push_back(comp[t].val);
        }
        P.pat_tptr.push_back((uint32_t)P.term_val.size());
        ++pl;
      }
      P.groups.push_back(sg.G);
      ++gl;
    };
    auto hbit_of = [&](const SubGroup& sg) -> uint32_t {
      const uint32_t xo = (sg.G.xt & 0x7fffffffu) & ~reg_tmask;
      if (sg.remote || xo == 0u) return 0u;
      const int msb = 31 - __builtin_clz(xo);
      return msb >= kRegBitLo + rb ? (uint32_t)(msb + 1) : 0u;
    };
    for (size_t q : singles) {
      SubGroup& sg = subs[q];
      cr.s_cnt[s_index(sg)]++;
      emit_group(sg, hbit_of(sg));
    }
    cr.ns = (uint16_t)singles.size();
    P.n_singles += (int64_t)singles.size();
    for (size_t r : multi) {
      const size_t si = runs[r].s0, se = runs[r].s1;
      BundleRec B;
      const uint32_t hbit = hbit_of(subs[si]);
      B.xo = ((subs[si].G.xt & 0x7fffffffu) & ~reg_tmask) | (subs[si].remote ? 0x80000000u : 0u);
      B.meta = (uint32_t)(se - si) | (hbit << 8) | ((uint32_t)gl << 16);
      B.xlo = subs[si].xlo;
      P.bundles.push_back(B);
      if (subs[si].remote) P.n_remote_bundles++;
      for (size_t q = si; q < se; ++q) emit_group(subs[q], hbit);
    }
    cr.g0 = g0;
    cr.ng = gl;
    cr.p0 = p0;
    cr.np = pl;
    cr.nb = (int32_t)P.bundles.size() - cr.b0;
    cr.nq = (uint16_t)nq_emitted;
    if (has_long) cr.q0 |= 0x80000000u;
    P.chunks.push_back(cr);
    g0 += gl;
    p0 += pl;
    r0 = r1;
  }
  return P;
}

// Split merged terms into parts by global x-part for (rank, world) and plan each part.
inline std::vector<PartHost> build_parts(const std::vector<MergedTerm>& terms, int n, int p_bits,
                                         int rank, const PlanOptions& opt) {
  const int n_loc = n - p_bits;
  const uint64_t lm = low_mask(n_loc);
  std::map<int, PartHost> parts;
  for (const auto& t : terms) {
    int g = p_bits ? (int)(t.x >> n_loc) : 0;
    uint64_t zg = p_bits ? (t.z >> n_loc) : 0;
    double sgn = (popc64((uint64_t)rank & zg) & 1) ? -1.0 : 1.0;
    PartHost& P = parts[g];
    P.g = g;
    if (t.re != 0.0) P.comp.push_back({t.x & lm, t.z & lm, sgn * t.re, 0});
    if (t.im != 0.0) P.comp.push_back({t.x & lm, t.z & lm, sgn * t.im, 1});
  }
  std::vector<PartHost> out;
  for (auto& kv : parts) {
    PartHost& P = kv.second;
    // distinct local x masks, sorted: diagonal first, then by value
    std::map<uint64_t, int> idx;
    std::vector<uint64_t> masks;
    std::vector<std::vector<int>> tom;
    for (int t = 0; t < (int)P.comp.size(); ++t) {
      auto it = idx.find(P.comp[t].xl);
      if (it == idx.end()) {
        it = idx.emplace(P.comp[t].xl, (int)masks.size()).first;
        masks.push_back(P.comp[t].xl);
        tom.emplace_back();
      }
      tom[it->second].push_back(t);
    }
    P.n_xgroups = (int64_t)masks.size();
    // streaming plan (kernel A): merge re/im components of the same (x,z) back to complex
    {
      std::vector<int> order(masks.size());
      for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
      std::sort(order.begin(), order.end(), [&](int a, int b) { return masks[a] < masks[b]; });
      for (int mi : order) {
        SimpleGroup sg{masks[mi], (int32_t)P.sterms.size(), 0};
        std::map<uint64_t, int> zidx;
        for (int t : tom[mi]) {
          auto it = zidx.find(P.comp[t].zl);
          if (it == zidx.end()) {
            it = zidx.emplace(P.comp[t].zl, (int)P.sterms.size()).first;
            P.sterms.push_back({P.comp[t].zl, 0.0, 0.0});
          }
          if (P.comp[t].im) P.sterms[it->second].im += P.comp[t].val;
          else P.sterms[it->second].re += P.comp[t].val;
        }
        sg.nt = (int32_t)P.sterms.size() - sg.t0;
        P.sgroups.push_back(sg);
      }
    }
    // tiled plan (kernel B)
    if (n_loc >= 11 && opt.kernel != 1) {
      int L = std::max(11, std::min({opt.tile_bits, kMaxTileBits, n_loc}));
      // cost of a plan in units of 16 N bytes of HBM traffic: a pass reads psi and reads+writes
      // y (3 units), a remote group re-reads one partner vector (1 unit)
      std::vector<PassChoice> pcs;
      long best_cost = -1;
      // the restarts share a deterministic work budget (mask visits of the greedy loops), so that an
      // operator with tens of thousands of distinct x-masks is still planned in about a second
      double work = 0.0;
      constexpr double kPlanWorkBudget = 4e8;
      for (int t = 0; t < std::max(1, opt.plan_tries); ++t) {
        if (t >= 4 && work > kPlanWorkBudget) break;
        std::vector<PassChoice> cand = choose_passes(masks, n_loc, L, opt.low_bits, opt.min_cover, (uint64_t)t, opt.late_bits, opt.late_split);
        work += (double)masks.size() * (double)cand.size() * (double)std::max(1, L - opt.low_bits);
        long n_remote = 0;
        for (auto& pc : cand)
          for (int mi : pc.members)
            if (masks[mi] & ~pc.free_mask) ++n_remote;
        const long cost = 3 * (long)cand.size() + n_remote;
        if (best_cost < 0 || cost < best_cost) {
          best_cost = cost;
          pcs.swap(cand);
        }
      }
      for (auto& pc : pcs) {
        const int rb = opt.reg_bits == 4 ? 4 : 3;
        // one big CTA needs a full round of 2^12 amplitudes: tiles of 2^11 always run as two CTAs
        const int Lp = popc64(pc.free_mask);
        const bool can_big = Lp >= 12;
        PassHost ph = build_pass(P.comp, tom, masks, pc, n_loc, rb,
                                 ((opt.cta_mode == 2 && can_big) || Lp >= 13) ? big_caps(Lp) : small_caps(), opt.min_bundle, opt.share_weight);
        if (opt.cta_mode == 0 && can_big && !ph.caps.big && (ph.chunks.size() > 1 || ph.n_quad > 0))
          ph = build_pass(P.comp, tom, masks, pc, n_loc, rb, big_caps(Lp), opt.min_bundle, opt.share_weight);
        P.passes.push_back(std::move(ph));
      }
    }
    out.push_back(std::move(P));
  }
  return out;
}

}  // namespace psum
