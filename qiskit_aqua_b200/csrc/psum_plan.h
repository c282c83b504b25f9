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
struct GroupRec {     // 48 bytes
  uint32_t xt;        // tile-local xor (x gathered onto the tile index); bit31 = remote flag
  uint32_t meta;      // bits 0..15 first pattern (relative to the chunk), 16..18 entry count,
                      // bit 20 = has imaginary patterns, bit 21 = flat (one real class-0 entry),
                      // bits 24..27 = class (im*8 + zr) of the inline pattern in slot 1 (kind A groups)
                      // bits 28..31 = hbit: 1 + tile bit (>= 8, i.e. warp-uniform) of the highest
                      //               set bit of xt for local off-diagonal groups, else 0.  The
                      //               Hermitian <psi|H|psi> kernels process such a group only for
                      //               amplitudes with that bit clear and double it (pair i, i^x)
  uint32_t zt0, zt1;  // in-tile z-patterns of the first two patterns (inline copy)
  double cf0, cf1;    // their folded coefficients: written in SHARED memory by the fold
  uint64_t xl;        // full local x mask (partner index = i ^ xl)
  uint16_t ent[4];    // non-empty classes: (cls << 12) | count ; cls = im*8 + ((zt >> 5) & 7)
};
static_assert(sizeof(GroupRec) == 48, "GroupRec layout");

struct PatRec {       // 16 bytes, staged in shared memory per chunk (cf is written by the fold)
  double cf;
  uint32_t zt;
  uint32_t pad;
};

struct ChunkRec {  // 32 bytes
  int32_t g0;  // first group
  int32_t ng;  // low 16 bits: groups in the chunk; high 16 bits: how many of them (the first
               // ones) are "kind A": local, real, <= 2 patterns with slot 0 of class 0 -> fast loop
  int32_t p0, np;
  uint8_t a_cnt[16];  // kind A groups are sorted by the class of slot 1 (0..7 real, 8..15
                      // imaginary): count per class
};
static_assert(sizeof(ChunkRec) == 32, "ChunkRec layout");

constexpr int kMaxTileBits = 13;
constexpr int kRegBitLo = 5;     // tile bits 5..7 live in registers (8 amplitudes per thread)
#ifndef PSUM_PAT_CHUNK
#define PSUM_PAT_CHUNK 1024
#endif
#ifndef PSUM_GRP_CHUNK
#define PSUM_GRP_CHUNK 128
#endif
constexpr int kPatChunk = PSUM_PAT_CHUNK;  // patterns staged in shared memory at a time
constexpr int kGrpChunk = PSUM_GRP_CHUNK;  // groups staged in shared memory at a time

struct PassHost {
  int L = 0;                   // tile bits
  uint64_t free_mask = 0;      // over local qubits
  uint8_t fbit[16] = {0};      // qubit of tile bit j: 0..4 lane bits, 5..7 register bits, 8.. warp/round
  bool has_im = false;         // any imaginary component pattern in this pass
  int hi_bit = 0;              // highest local qubit the pass combines (free set or any member x-mask)
  int64_t n_flat = 0;          // groups whose D is the same for the 8 register amplitudes
  int64_t n_kind_a = 0;        // of those: local, <= 2 patterns (fast loop)
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
  int low_bits = 4;
  int min_cover = 3;
  int plan_tries = 8;  // randomised greedy restarts of the pass planner; the cheapest plan wins
  int late_bits = 0;   // top local qubits kept out of the early passes (see choose_passes)
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
                                             int c, int min_cover, uint64_t seed = 0, int late_bits = 0) {
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
                           const std::vector<uint64_t>& masks, const PassChoice& pc, int n_loc) {
  PassHost P;
  P.free_mask = pc.free_mask;
  P.L = popc64(pc.free_mask);
  const uint64_t S = pc.free_mask;
  // ---- roles of the free bits -------------------------------------------------------------
  // tile bits 0..2 = the three lowest free qubits (128-byte lines / conflict-free LDS.128),
  // tile bits 5..7 (register bits) = the free qubits least often touched by a z-mask of this
  // pass, so that most groups have the same D for all 8 amplitudes of a thread ("flat"),
  // the rest ascending.
  {
    std::vector<int> fb;
    for (int b = 0; b < n_loc; ++b)
      if (S >> b & 1ull) fb.push_back(b);
    const int L = (int)fb.size();
    std::vector<double> hits(64, 0.0);
    for (int mi : pc.members)
      for (int t : terms_of_mask[mi])
        for (uint64_t m = comp[t].zl & S; m; m &= m - 1) hits[__builtin_ctzll(m)] += 1.0;
    std::vector<int> cand(fb.begin() + std::min(3, L), fb.end());
    std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) { return hits[a] < hits[b]; });
    std::vector<int> reg(cand.begin(), cand.begin() + std::min<size_t>(3, cand.size()));
    std::sort(reg.begin(), reg.end());
    std::vector<int> rest;
    for (int b : cand)
      if (std::find(reg.begin(), reg.end(), b) == reg.end()) rest.push_back(b);
    // the qubits most often flipped by this pass's x-masks go to the warp-uniform tile bits (8..),
    // where the Hermitian expectation kernels can skip the partner half of a group
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
  struct Key {
    uint8_t im;
    uint32_t zt;
    int term;
  };
  struct PatTmp { int cls; uint32_t zt; std::vector<int> terms; };
  struct SubGroup { GroupRec G; std::vector<PatTmp> pats; bool kind_a; int slot_of[2]; };
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
    auto cls_of = [](const Key& k) { return (int)k.im * 8 + (int)((k.zt >> kRegBitLo) & 7); };
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
    // sub-groups: each holds <= 4 non-empty classes, <= kPatChunk patterns and <= 4095 patterns
    // per class.  Patterns are in class order, so cut greedily.
    size_t p = 0;
    while (p < pats.size()) {
      SubGroup sg;
      GroupRec& G = sg.G;
      std::memset(&G, 0, sizeof(G));
      G.xt = gather(xl) | (remote ? 0x80000000u : 0u);
      G.xl = xl;
      int count = 0, nent = 0;
      bool has_im = false;
      size_t q = p;
      while (q < pats.size() && count < kPatChunk) {
        const int cls = pats[q].cls;
        if (nent == 0 || (G.ent[nent - 1] >> 12) != cls) {
          if (nent == 4) break;
          G.ent[nent++] = (uint16_t)(cls << 12);
        } else if ((G.ent[nent - 1] & 0xfff) == 0xfff) {
          break;
        }
        G.ent[nent - 1]++;
        if (cls >= 8) has_im = true;
        ++count;
        ++q;
      }
      const bool flat = (nent == 1) && (G.ent[0] >> 12) == 0;
      G.meta = ((uint32_t)nent << 16) | (has_im ? (1u << 20) : 0u) | (flat ? (1u << 21) : 0u);
      {
        const uint32_t xt_only = G.xt & 0x7fffffffu;
        if (!remote && xt_only != 0u) {
          const int msb = 31 - __builtin_clz(xt_only);
          if (msb >= 8) G.meta |= (uint32_t)(msb + 1) << 28;
        }
      }
      // kind A: local, real, at most two patterns, the first (if two) of class 0.  Slot 0 of the
      // inline copy holds the class-0 pattern (or nothing), slot 1 the other one with its class.
      sg.kind_a = !remote && count <= 2 && (count == 1 || pats[p].cls == 0);
      sg.slot_of[0] = 0;
      sg.slot_of[1] = 1;
      if (sg.kind_a) {
        if (count == 1 && pats[p].cls != 0) sg.slot_of[0] = 1;   // lone non-flat pattern -> slot 1
        for (int j = 0; j < count; ++j) {
          if (sg.slot_of[j] == 0) G.zt0 = pats[p + j].zt;
          else {
            G.zt1 = pats[p + j].zt;
            G.meta |= (uint32_t)(pats[p + j].cls & 15) << 24;
          }
        }
      } else {
        G.zt0 = pats[p].zt;
        G.zt1 = count > 1 ? pats[p + 1].zt : 0u;
      }
      if (has_im) P.has_im = true;
      if (flat) P.n_flat++;
      if (sg.kind_a) P.n_kind_a++;
      sg.pats.assign(pats.begin() + p, pats.begin() + q);
      subs.push_back(std::move(sg));
      p = q;
    }
  }
  // chunks: consecutive sub-groups up to the staging capacity; inside a chunk kind A goes first
  P.pat_tptr.push_back(0);
  size_t s0 = 0;
  int g0 = 0, p0 = 0;
  while (s0 < subs.size()) {
    size_t s1 = s0;
    int np = 0;
    while (s1 < subs.size() && (int)(s1 - s0) < kGrpChunk && np + (int)subs[s1].pats.size() <= kPatChunk) {
      np += (int)subs[s1].pats.size();
      ++s1;
    }
    std::stable_sort(subs.begin() + s0, subs.begin() + s1, [](const SubGroup& a, const SubGroup& b) {
      const int ka = a.kind_a ? (int)((a.G.meta >> 24) & 15u) : 16, kb = b.kind_a ? (int)((b.G.meta >> 24) & 15u) : 16;
      return ka < kb;
    });
    ChunkRec cr;
    std::memset(&cr, 0, sizeof(cr));
    int ng_a = 0, pl = 0;
    for (size_t si = s0; si < s1; ++si) {
      SubGroup& sg = subs[si];
      if (sg.kind_a) {
        ++ng_a;
        cr.a_cnt[(sg.G.meta >> 24) & 15u]++;
      }
      sg.G.meta |= (uint32_t)pl;
      for (size_t j = 0; j < sg.pats.size(); ++j) {
        // bit 15 of the stored pattern marks patterns of a Hermitian-paired group (doubled by the
        // fold of the Hermitian expectation kernels, masked off by every fold)
        P.pat_zt.push_back((uint16_t)(sg.pats[j].zt | ((sg.G.meta >> 28) ? 0x8000u : 0u)));
        P.pat_inl.push_back(j < 2 ? (uint16_t)((si - s0) * 2 + sg.slot_of[j]) : (uint16_t)0xffff);
        for (int t : sg.pats[j].terms) {
          P.term_zb.push_back(comp[t].zl & ~S);
          P.term_val.push_back(comp[t].val);
        }
        P.pat_tptr.push_back((uint32_t)P.term_val.size());
        ++pl;
      }
      P.groups.push_back(sg.G);
    }
    cr.g0 = g0;
    cr.ng = (int32_t)((s1 - s0) | ((size_t)ng_a << 16));
    cr.p0 = p0;
    cr.np = np;
    P.chunks.push_back(cr);
    g0 += (int)(s1 - s0);
    p0 += np;
    s0 = s1;
  }
  (void)n_loc;
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
      for (int t = 0; t < std::max(1, opt.plan_tries); ++t) {
        std::vector<PassChoice> cand = choose_passes(masks, n_loc, L, opt.low_bits, opt.min_cover, (uint64_t)t, opt.late_bits);
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
      for (auto& pc : pcs) P.passes.push_back(build_pass(P.comp, tom, masks, pc, n_loc));
    }
    out.push_back(std::move(P));
  }
  return out;
}

}  // namespace psum
