// psum_kernels.cuh -- sm_100a device code of the Pauli-sum engine.
//
//   y[i] = sum_g D_g(i) * src[i ^ x_g],   D_g(i) = sum_{k in g} c'_k (-1)^{popc(i & z_k)}
//
// Kernel B (k_pass) is the product path.  One launch = one "pass": a free set S of L qubits
// defines tiles (the 2^L amplitudes that differ only in S).  A CTA stages one tile of the source
// vector in shared memory with cp.async, and every x-group whose mask lies inside S reads its
// partner amplitudes from that tile -- psi is read from HBM once per pass, not once per group.
// Signs are never evaluated per (term, amplitude):
//   * bits outside S:   folded into the coefficient once per (tile, term)      ("fold")
//   * equal in-tile z-patterns of one group are merged into one pattern          ("compaction")
//   * register bits = tile bits 5..5+RB-1 (the 2^RB amplitudes a thread owns): patterns are summed
//     per class zr = (zt >> 5) & (2^RB - 1) and added with compile-time signs; a group whose
//     patterns are all class 0 ("flat") has one D for all amplitudes of a thread   ("class sums")
//   * the remaining tile bits: one popc per (pattern, thread)
// Partner sharing: x-groups whose masks differ only in register bits form a bundle; the 2^RB
// partner amplitudes are loaded ONCE per bundle and every member applies its own compile-time
// register permutation pv[r ^ xr] -- the shared-memory pipe, which bounds this kernel, sees one
// load set per bundle instead of one per group.
// Per-tile overhead: a pass whose records fit one staging chunk copies its group / bundle records -- and,
// for real-only passes, its component terms -- to shared memory ONCE per CTA; per tile only the fold
// (signs of the bits outside S -> coefficients) and the quadratic-group tables are redone, overlapped with
// the cp.async tile load; the old y of the first round is requested before the wait for the tile.
// Kernel A (k_stream) is the plain streaming form used for n_loc < 11 qubits.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "psum_plan.h"

namespace psum {

constexpr int kThreads = 256;   // block size of the small vector kernels

struct PassParams {
  const BundleRec* bundles;
  const GroupRec* groups;
  const uint16_t* pat_zt;
  const uint16_t* pat_inl;
  const uint32_t* pat_tptr;
  const uint64_t* term_zb;
  const double* term_val;
  const ChunkRec* chunks;
  const uint16_t* quad_idx;
  int nchunks;
  int L;
  int n_nonfree;
  int accumulate;
  int quad_cap;   // doubles of the quad-table arena behind the tile
  int term_cache; // > 0: a one-chunk pass keeps its `term_cache` component terms (zb, val) in shared memory behind
                  // the quad arena (copied once per CTA): the per-tile fold then reads them at shared-memory
                  // latency instead of ~L2 latency
  int batch;      // > 1: `batch` source vectors (stride batch_stride amplitudes) share the staged plan of each
                  // tile: the fold and the table build run once per tile, not once per state
                  // (expectation-only launches of one-chunk passes; partial[b * gridDim.x + block])
  uint64_t batch_stride;
  uint64_t ntiles;
  uint64_t tile_begin, tile_end;  // tile range of this launch (whole pass: 0 .. ntiles)
  uint64_t roff[16];  // byte offset (16 * amplitude index) of register amplitude r inside a tile
  uint8_t fbit[16];   // qubit of tile bit j (0..4 lane, 5..5+RB-1 register, then warp/round)
  uint8_t nbit[64];   // positions of the non-free local bits, ascending
};

// shared memory of one k_pass CTA (bytes), L tile bits
// layout: [offHi 64 u64][red 32 double2][baseT 9*64 u64][group records][bundle records][patterns] at
// compile-time offsets (big = one CTA per SM), then the tile (2^L amplitudes), then the quad arena
constexpr int kChunkSlots = 8;   // chunk records mirrored in shared memory (read every round); one more slot is
                                 // refilled on the fly for the chunks of a pass beyond the first kChunkSlots
constexpr int kMaxBatch = 32;    // states per batched launch (per-warp accumulators in shared memory)
__host__ __device__ constexpr size_t pass_static_bytes_c(bool big) {
  return 8 * 64 + 16 * 32 + 8 * 9 * 64 + sizeof(ChunkRec) * (kChunkSlots + 1) + sizeof(GroupRec) * ((big ? kGrpCapBig : kGrpCapSmall) + 1) +
         sizeof(BundleRec) * ((big ? kGrpCapBig : kGrpCapSmall) + 1) + sizeof(PatRec) * (big ? kPatCapBig : kPatCapSmall);
}
__host__ __device__ constexpr size_t pass_smem_bytes_c(int L, bool big, int quad_cap, int batch = 1, int term_cache = 0) {
  return pass_static_bytes_c(big) + ((size_t)16 << L) + 8 * (size_t)quad_cap +
         (batch > 1 ? (size_t)16 * 32 * kMaxBatch : 0) +   // [warp][state] expectation accumulators (<= 32 warps)
         (size_t)16 * (size_t)term_cache;                  // cached component terms: zb[term_cache], val[term_cache]
}

__device__ __forceinline__ double flip_sign(double v, int s) {
  return __hiloint2double(__double2hiint(v) ^ (s << 31), __double2loint(v));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
// loads through 32-bit shared-space addresses: the hot loops keep one base register per array instead
// of re-deriving the CTA's shared window (S2R SR_CgaCtaId) for every generic pointer
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double2 lds_d2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double lds_d(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// deterministic block reduction of a complex value; result valid in thread 0
__device__ __forceinline__ double2 block_reduce(double2 v, double2* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_down_sync(0xffffffffu, v.x, o);
    v.y += __shfl_down_sync(0xffffffffu, v.y, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    v = l < nw ? scratch[l] : make_double2(0.0, 0.0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v.x += __shfl_down_sync(0xffffffffu, v.x, o);
      v.y += __shfl_down_sync(0xffffffffu, v.y, o);
    }
  }
  return v;
}

// ------------------------------------------------------------------------------------------
// Kernel B: one pass over all tiles of a free set.
// ------------------------------------------------------------------------------------------
#define PSUM_REP32(M)                                                                        \
  M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11) M(12) M(13) M(14) M(15)      \
  M(16) M(17) M(18) M(19) M(20) M(21) M(22) M(23) M(24) M(25) M(26) M(27) M(28) M(29) M(30) M(31)

// D[(C / NA) * NA + r] += (-1)^{popc(r & C)} s  for r = 0..NA-1, C a compile-time class (im*NA + zr)
template <int NA, int C, int ND>
__device__ __forceinline__ void add_class(double (&D)[ND], double s) {
  if constexpr (C < ND) {
#pragma unroll
    for (int r = 0; r < NA; ++r) {
      if (__popc(r & (C & (NA - 1))) & 1) D[(C / NA) * NA + r] -= s;
      else D[(C / NA) * NA + r] += s;
    }
  }
}

// fast members: yacc[r] += d_r * pv[PRE ? r : r ^ XR], d_r = dp, or dm where popc(r & XR) is odd
// (PRE: the loads already applied the register permutation)
template <int NA, int XR, bool PRE>
__device__ __forceinline__ void fast_fma(double2 (&yacc)[NA], const double2 (&pv)[NA], double dp, double dm) {
  if constexpr (XR < NA) {
#pragma unroll
    for (int r = 0; r < NA; ++r) {
      const double d = (__popc(r & XR) & 1) ? dm : dp;
      const double2 v = pv[PRE ? r : (r ^ XR)];
      yacc[r].x = fma(d, v.x, yacc[r].x);
      yacc[r].y = fma(d, v.y, yacc[r].y);
    }
  }
}
// the same with an imaginary slot 1: yacc[r] += (s0 + i d_r) * pv[..], d_r = da / db
template <int NA, int XR, bool PRE>
__device__ __forceinline__ void fast_fma_im(double2 (&yacc)[NA], const double2 (&pv)[NA], double s0, double da,
                                            double db) {
  if constexpr (XR < NA) {
#pragma unroll
    for (int r = 0; r < NA; ++r) {
      const double d = (__popc(r & XR) & 1) ? db : da;
      const double2 v = pv[PRE ? r : (r ^ XR)];
      yacc[r].x = fma(s0, v.x, fma(-d, v.y, yacc[r].x));
      yacc[r].y = fma(s0, v.y, fma(d, v.x, yacc[r].y));
    }
  }
}

// general members: yacc[r] += (D[r] + i D[NA + r]) * pv[r ^ XR]
template <int NA, int XR, bool CPLX, int ND>
__device__ __forceinline__ void perm_fma(double2 (&yacc)[NA], const double2 (&pv)[NA], const double (&D)[ND]) {
  if constexpr (XR < NA) {
#pragma unroll
    for (int r = 0; r < NA; ++r) {
      if constexpr (CPLX) {
        yacc[r].x = fma(D[r], pv[r ^ XR].x, fma(-D[NA + r], pv[r ^ XR].y, yacc[r].x));
        yacc[r].y = fma(D[r], pv[r ^ XR].y, fma(D[NA + r], pv[r ^ XR].x, yacc[r].y));
      } else {
        yacc[r].x = fma(D[r], pv[r ^ XR].x, yacc[r].x);
        yacc[r].y = fma(D[r], pv[r ^ XR].y, yacc[r].y);
      }
    }
  }
}

// sum of the (tile-folded, thread-signed) coefficients of `cnt` patterns starting at p
__device__ __forceinline__ double pattern_sum(uint32_t pat_s, int& p, int cnt, uint32_t t0) {
  double s = 0.0;
#pragma unroll 4
  for (int j = 0; j < cnt; ++j, ++p) {
    const uint4 rec = lds_u4(pat_s + 16u * (uint32_t)p);
    s += __hiloint2double((int)(rec.y ^ ((uint32_t)(__popc(rec.z & t0)) << 31)), (int)rec.x);
  }
  return s;
}

//   STORE : y (+)= contribution.  With `accumulate` the accumulators start from the old y, so a
//           pass costs one read and one write of y and the last pass sees the final H.psi
//   EXPECT: STORE mode  -> sum conj(bra_i) * y_i of the FINAL y (host enables it on the last
//                          launch only);  no-STORE mode -> sum conj(bra_i) * contribution_i
//   IM    : the pass has imaginary component patterns (complex D); real-only passes carry half
//           the class accumulators
//   RB,NT : 2^RB amplitudes per thread, NT threads per CTA.  (3,256) and (4,128): two CTAs per SM,
//           tiles up to 2^12;  (3,512) and (4,256): one CTA per SM, 2^13 tiles
//   bra == nullptr -> the bra is the staged tile itself (<psi|H|psi> on the local part)
//   HERM  : (expectation only, Hermitian H, bra == source) a local off-diagonal bundle whose x has
//           its highest tile bit among the warp-uniform bits is evaluated only on the half of the
//           amplitudes with that bit clear, with doubled coefficients: the pair (i, i^x)
//           contributes z + conj(z) = 2 Re z.  The imaginary part (identically zero for a
//           Hermitian operator) is returned as exactly 0.
template <bool STORE, bool EXPECT, bool IM, int RB, int NT, bool HERM = false>
__global__ void __launch_bounds__(NT, (NT << RB) <= 2048 ? 2 : 1)
k_pass(const double2* __restrict__ src, const double2* __restrict__ bra, double2* __restrict__ y,
       const PassParams P, double2* __restrict__ partial) {
  constexpr int NA = 1 << RB;
  constexpr int ND = IM ? 2 * NA : NA;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int L = P.L;
  const uint32_t tile_amps = 1u << L;
  constexpr bool BIG = (NT << RB) >= 4096;
  constexpr int PCAP = BIG ? kPatCapBig : kPatCapSmall, GCAP = BIG ? kGrpCapBig : kGrpCapSmall;
  uint64_t* offHi = reinterpret_cast<uint64_t*>(smem_raw);
  double2* red = reinterpret_cast<double2*>(offHi + 64);
  uint64_t* baseT = reinterpret_cast<uint64_t*>(red + 32);  // [9][64] deposit tables
  uint4* chk4 = reinterpret_cast<uint4*>(baseT + 9 * 64);   // the first kChunkSlots chunk records (4 x uint4 each)
  uint4* grp4 = chk4 + 4 * (kChunkSlots + 1);               // 3 x uint4 per group record
  uint4* bun4 = grp4 + 3 * (GCAP + 1);                      // 1 x uint4 per bundle record
  PatRec* pat = reinterpret_cast<PatRec*>(bun4 + (GCAP + 1));
  double2* tile = reinterpret_cast<double2*>(pat + PCAP);
  double* qtab = reinterpret_cast<double*>(tile + tile_amps);   // tables of the chunk's quadratic groups
  double2* wacc = reinterpret_cast<double2*>(qtab + P.quad_cap);   // [warp][kMaxBatch], batched launches only
  const int nbatch = (!STORE && EXPECT && P.batch > 1) ? P.batch : 1;
  // the shared-memory term cache is compiled into the real-only variants (the host enables it only for
  // passes without imaginary patterns): in the IM variants the second fold path costs registers the
  // round loops need (measured: C5 +2 %)
  constexpr bool TCACHE = !IM;
  const uint4* pat4 = reinterpret_cast<const uint4*>(pat);
  const int nW = L - 5 - RB;                                 // warp/round tile bits
  const uint32_t tile_s = smem_addr(tile), grp_s = smem_addr(grp4), bun_s = smem_addr(bun4), pat_s = smem_addr(pat),
                 qtab_s = smem_addr(qtab), chk_s = smem_addr(chk4);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // deposit tables: offset of tile element e = offLo[e & 127] | offHi[e >> 7]
  uint64_t offLo = 0;
#pragma unroll
  for (int j = 0; j < 7; ++j) offLo |= (uint64_t)((tid >> j) & 1) << P.fbit[j];
  const uint64_t offLane = offLo & ((1ull << P.fbit[0]) | (1ull << P.fbit[1]) | (1ull << P.fbit[2]) |
                                    (1ull << P.fbit[3]) | (1ull << P.fbit[4]));
  if (tid < 64) {
    uint64_t o = 0;
    for (int j = 7; j < L; ++j) o |= (uint64_t)((tid >> (j - 7)) & 1) << P.fbit[j];
    offHi[tid] = o;
  }
  // tile id -> base index: 6 bits of the tile id per table lookup
  const int nbt = (P.n_nonfree + 5) / 6;
  for (int e = tid; e < nbt * 64; e += NT) {
    const int k = e >> 6, v = e & 63;
    uint64_t o = 0;
    for (int j = 0; j < 6 && 6 * k + j < P.n_nonfree; ++j) o |= (uint64_t)((v >> j) & 1) << P.nbit[6 * k + j];
    baseT[e] = o;
  }
  for (int e = tid; e < 4 * (P.nchunks < kChunkSlots ? P.nchunks : kChunkSlots); e += NT)
    chk4[e] = reinterpret_cast<const uint4*>(P.chunks)[e];
  if (nbatch > 1)
    for (int e = tid; e < (NT / 32) * kMaxBatch; e += NT) wacc[e] = make_double2(0.0, 0.0);
  __syncthreads();
  const int nrounds = (int)(tile_amps / (NT * NA));
  const int nload = (int)(tile_amps / NT);
  const bool one_chunk = (P.nchunks == 1);

  double2 eacc = make_double2(0.0, 0.0);

  // Stage a chunk: copy its bundle and group records to shared memory, then fold -- bits outside
  // the free set become a per-tile sign of each term; equal in-tile patterns of a group are summed
  // into one coefficient, stored in pat[] and (first two patterns of a group) inline in its record.
  // copy_records = false: the records of this chunk are already in shared memory (a one-chunk pass copies
  // them once per CTA; every tile's fold overwrites all the coefficient slots it uses)
  auto stage_chunk = [&](const ChunkRec ch, uint64_t base, bool copy_records) {
    if (copy_records) {
      for (int g = tid; g < ch.ng * 3; g += NT)
        grp4[g] = reinterpret_cast<const uint4*>(P.groups + ch.g0)[g];
      for (int b = tid; b < ch.nb; b += NT)
        bun4[b] = reinterpret_cast<const uint4*>(P.bundles + ch.b0)[b];
      if (tid < 3) grp4[ch.ng * 3 + tid] = make_uint4(0u, 0u, 0u, 0u);   // read-ahead slots
      if (tid == 3) bun4[ch.nb] = make_uint4(0u, 0u, 0u, 0u);
    }
    // long term lists (a pattern that collects, e.g., every Z_a Z_b with both qubits outside the
    // tile) are deferred to a queue and folded by a whole warp each: one lane looping over a hundred
    // terms would stall its warp -- and with it the CTA -- at the next barrier
    int* lq_count = reinterpret_cast<int*>(red);
    uint16_t* lq = reinterpret_cast<uint16_t*>(red) + 8;
    constexpr int kLongCap = 120;
    const bool has_long = (ch.q0 >> 31) != 0;
    // term cache (optional, behind the quad arena and the batch accumulators): addresses derived here so
    // that they do not occupy registers in the round loops
    const uint32_t tc_zb_s = qtab_s + 8u * (uint32_t)P.quad_cap + (P.batch > 1 ? 16u * 32u * kMaxBatch : 0u);
    const uint32_t tc_val_s = tc_zb_s + 8u * (uint32_t)P.term_cache;
    if (has_long && tid == 4) *lq_count = 0;
    if (copy_records || has_long) __syncthreads();
    auto put_pattern = [&](int p, int gp, double v) {
      const uint32_t zraw = P.pat_zt[gp];
      if (HERM && (zraw & 0x8000u)) v += v;
      uint4 rec;
      rec.x = (uint32_t)__double2loint(v);
      rec.y = (uint32_t)__double2hiint(v);
      rec.z = zraw & 0x7fffu;
      rec.w = 0;
      reinterpret_cast<uint4*>(pat)[p] = rec;
      const uint32_t inl = P.pat_inl[gp];
      if (inl != 0xffffu)
        reinterpret_cast<double*>(grp4 + 3 * (inl >> 1) + 1)[inl & 1u] = v;
    };
    for (int p = tid; p < ch.np; p += NT) {
      const int gp = ch.p0 + p;
      const uint32_t ta = P.pat_tptr[gp], tb = P.pat_tptr[gp + 1];
      if (has_long && tb - ta > (uint32_t)kLongTerms) {
        const int k = atomicAdd(lq_count, 1);
        if (k < kLongCap) {
          lq[k] = (uint16_t)p;
          continue;
        }
      }
      double v = 0.0;
      if (TCACHE && P.term_cache) {
#pragma unroll 4
        for (uint32_t t = ta; t < tb; ++t) {
          const uint2 zb = lds_u2(tc_zb_s + 8u * t);
          v += flip_sign(lds_d(tc_val_s + 8u * t), (__popc(zb.x & (uint32_t)base) + __popc(zb.y & (uint32_t)(base >> 32))) & 1);
        }
      } else {
#pragma unroll 4
        for (uint32_t t = ta; t < tb; ++t)
          v += flip_sign(__ldg(P.term_val + t), __popcll(base & __ldg(P.term_zb + t)) & 1);
      }
      put_pattern(p, gp, v);
    }
    if (has_long) {
      __syncthreads();
      const int nlong = *lq_count < kLongCap ? *lq_count : kLongCap;
      for (int li = warp; li < nlong; li += NT / 32) {
        const int p = (int)lq[li], gp = ch.p0 + p;
        const uint32_t ta = P.pat_tptr[gp], tb = P.pat_tptr[gp + 1];
        double v = 0.0;
        if (TCACHE && P.term_cache) {
          for (uint32_t t = ta + (uint32_t)lane; t < tb; t += 32u) {
            const uint2 zb = lds_u2(tc_zb_s + 8u * t);
            v += flip_sign(lds_d(tc_val_s + 8u * t), (__popc(zb.x & (uint32_t)base) + __popc(zb.y & (uint32_t)(base >> 32))) & 1);
          }
        } else {
          for (uint32_t t = ta + (uint32_t)lane; t < tb; t += 32u)
            v += flip_sign(P.term_val[t], __popcll(base & P.term_zb[t]) & 1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) put_pattern(p, gp, v);
      }
    }
    if (ch.nq) {
      // quadratic groups: rebuild the separable tables of this tile from the folded patterns.
      // One task = one table of one group = the signed sum of one pattern segment (psum_plan.h) for
      // every table index; a task runs on one warp, lane = table index, so the pattern loop is uniform.
      __syncthreads();
      const int nlt = 1 + nW + RB, ntab = nlt + 1 + RB + 1;
      const uint32_t ntab_inv = 65536u / (uint32_t)ntab + 1u;   // task / ntab by a multiplication (task < 2^12)
      for (int task = warp; task < (int)ch.nq * ntab; task += NT / 32) {
        const int qi = (int)(((uint32_t)task * ntab_inv) >> 16), tb = task - qi * ntab;
        const int g = (int)P.quad_idx[(ch.q0 & 0x7fffffffu) + qi];
        const uint32_t meta = reinterpret_cast<const uint32_t*>(grp4 + 3 * g)[1];
        const uint8_t* area = reinterpret_cast<const uint8_t*>(grp4 + 3 * g) + 8;
        double* qt = qtab + (int)*reinterpret_cast<const uint16_t*>(area + 22);
        const int pfirst = (int)(meta & 0xffffu);
        int seg, dst, cnt;
        uint32_t M;
        bool filt = false;
        if (tb < nlt) {                    // lane tables: T_L, C_j, TL_i
          seg = tb == 0 ? 0 : tb + 1;
          M = (uint32_t)lane;
          dst = tb * 32 + lane;
          cnt = 32;
        } else if (tb < nlt + 1 + RB) {    // w tables: T_W, TW_i
          const int t = tb - nlt;
          seg = t == 0 ? 1 : 2 + nW + RB + (t - 1);
          M = (uint32_t)lane << (5 + RB);
          dst = nlt * 32 + (t << nW) + lane;
          cnt = 1 << nW;
        } else {                           // K[mask of two register bits]
          seg = 2 + nW + 2 * RB;
          M = 0u;
          filt = true;
          dst = nlt * 32 + ((1 + RB) << nW) + lane;
          cnt = NA;
        }
        double v = 0.0;
        const int pa = pfirst + (int)area[seg], pb = pfirst + (int)area[seg + 1];
#pragma unroll 4
        for (int p = pa; p < pb; ++p) {
          const uint4 rec = pat4[p];
          const double c = __hiloint2double((int)(rec.y ^ ((uint32_t)__popc(rec.z & M) << 31)), (int)rec.x);
          if (!filt || (int)((rec.z >> 5) & (uint32_t)(NA - 1)) == lane) v += c;
        }
        if (lane < cnt) qt[dst] = v;
      }
    }
  };

  // one-chunk pass: its group / bundle records are the same for every tile -- copied once
  if (one_chunk) {
    ChunkRec ch0;
    uint4* cw = reinterpret_cast<uint4*>(&ch0);
#pragma unroll
    for (int q = 0; q < 4; ++q) cw[q] = lds_u4(chk_s + 16u * q);
    for (int g = tid; g < ch0.ng * 3; g += NT)
      grp4[g] = reinterpret_cast<const uint4*>(P.groups + ch0.g0)[g];
    for (int b = tid; b < ch0.nb; b += NT)
      bun4[b] = reinterpret_cast<const uint4*>(P.bundles + ch0.b0)[b];
    if (tid < 3) grp4[ch0.ng * 3 + tid] = make_uint4(0u, 0u, 0u, 0u);   // read-ahead slots
    if (tid == 3) bun4[ch0.nb] = make_uint4(0u, 0u, 0u, 0u);
    if (TCACHE) {
      uint64_t* tc_zb = reinterpret_cast<uint64_t*>(wacc + (P.batch > 1 ? 32 * kMaxBatch : 0));
      double* tc_val = reinterpret_cast<double*>(tc_zb + P.term_cache);
      for (int t = tid; t < P.term_cache; t += NT) {
        tc_zb[t] = P.term_zb[t];
        tc_val[t] = P.term_val[t];
      }
    }
    // (made visible by the barrier at the top of the first tile)
  }

  for (uint64_t tile_id = P.tile_begin + blockIdx.x; tile_id < P.tile_end; tile_id += gridDim.x) {
    uint64_t base = 0;
    for (int k = 0; k < nbt; ++k) base |= baseT[k * 64 + (int)((tile_id >> (6 * k)) & 63ull)];

   for (int bb = 0; bb < nbatch; ++bb) {
    const double2* __restrict__ srcb = src + (uint64_t)bb * P.batch_stride;
    __syncthreads();  // previous tile fully consumed
    for (int j = 0; j < nload; ++j) {
      const int e = j * NT + tid;
      const uint64_t idx = base | offLo | offHi[e >> 7];
      cp_async16(&tile[e], &srcb[idx]);
      // the old y of this tile is needed when the rounds start: pull its lines into L2 now
      // (one prefetch per 128-byte line: 8 amplitudes)
      if (STORE && P.accumulate && (tid & 7) == 0)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(y + idx));
    }
    // the old y of the first round is loaded NOW, so that its latency hides behind the fold and the wait
    // for the tile instead of stalling every warp of the CTA at once right after the barrier (real-only
    // passes; with imaginary patterns the longer live range costs more than it hides: measured)
    constexpr bool EARLY_Y = !IM;
    double2 yacc[NA];
    auto init_acc = [&](int rnd) {
      const uint32_t t0 = (uint32_t)rnd * (NT * NA) + (uint32_t)warp * (32u * NA) + (uint32_t)lane;
      const char* ybase = reinterpret_cast<const char*>(y + (base | offLane | offHi[t0 >> 7]));
#pragma unroll
      for (int r = 0; r < NA; ++r)
        yacc[r] = (STORE && P.accumulate) ? *reinterpret_cast<const double2*>(ybase + P.roff[r])
                                          : make_double2(0.0, 0.0);
    };
    if (EARLY_Y) init_acc(0);
    if (one_chunk && bb == 0) {   // the fold overlaps the tile load; shared by the batch
      ChunkRec ch0;
      uint4* cw = reinterpret_cast<uint4*>(&ch0);
#pragma unroll
      for (int q = 0; q < 4; ++q) cw[q] = lds_u4(chk_s + 16u * q);
      stage_chunk(ch0, base, false);
    }
    cp_async_wait_all();
    __syncthreads();

    for (int rnd = 0; rnd < nrounds; ++rnd) {
      const uint32_t t0 = (uint32_t)rnd * (NT * NA) + (uint32_t)warp * (32u * NA) + (uint32_t)lane;
      const uint64_t i0 = base | offLane | offHi[t0 >> 7];   // register bits of t0 are clear
      if (!EARLY_Y || rnd > 0) init_acc(rnd);

      for (int c = 0; c < P.nchunks; ++c) {
        // the chunk record always comes from shared memory (a global fallback inside this loop is
        // hoisted by the compiler and costs every round a long-scoreboard stall): chunks beyond the
        // mirrored ones go through the spare slot, refilled behind the barrier that staging needs anyway
        const int slot = c < kChunkSlots ? c : kChunkSlots;
        if (!one_chunk) {
          __syncthreads();  // everyone is done with the previous chunk's pat / grp (and the spare slot)
          if (c >= kChunkSlots) {
            if (tid < 4) chk4[4 * kChunkSlots + tid] = reinterpret_cast<const uint4*>(P.chunks + c)[tid];
            __syncthreads();
          }
        }
        ChunkRec ch;
        {
          uint4* cw = reinterpret_cast<uint4*>(&ch);
#pragma unroll
          for (int q = 0; q < 4; ++q) cw[q] = lds_u4(chk_s + 64u * (uint32_t)slot + 16u * q);
        }
        if (!one_chunk) {
          stage_chunk(ch, base, true);
          __syncthreads();
        }
        // ---- singles: a fast group alone in its bundle; one straight-line loop per (im, xr) with
        // the register permutation folded into the load offsets -----------------------------------
        {
          int g = 0;
#pragma unroll
          for (int s = 0; s < ND; ++s) {
            constexpr int kUnroll = RB == 3 ? 2 : 1;
            const int gend = g + (int)ch.s_cnt[s];
            const int XR = s & (NA - 1);   // compile-time after unrolling
#pragma unroll(kUnroll)
            for (; g < gend; ++g) {
              const uint4 h0 = lds_u4(grp_s + 48u * (uint32_t)g), hc = lds_u4(grp_s + 48u * (uint32_t)g + 16u);
              if (HERM && (h0.y >> 28) && ((t0 >> ((h0.y >> 28) - 1u)) & 1u)) continue;  // partner half
              const uint32_t pa = tile_s + 16u * ((t0 ^ h0.x) & ~((uint32_t)(NA - 1) << 5));
              double2 pv[NA];
#pragma unroll
              for (int r = 0; r < NA; ++r) pv[r] = lds_d2(pa + 512u * (uint32_t)(r ^ XR));
              const uint32_t q1 = (uint32_t)__popc(h0.w & t0);
              const uint32_t sg = (h0.y >> (22 + RB)) & 1u;   // slot-1 class = xr: odd r flip its sign
              const double s0 = __hiloint2double((int)(hc.y ^ ((uint32_t)__popc(h0.z & t0) << 31)), (int)hc.x);
              const double s1a = __hiloint2double((int)(hc.w ^ (q1 << 31)), (int)hc.z);
              const double s1b = __hiloint2double((int)(hc.w ^ ((q1 + sg) << 31)), (int)hc.z);
              if (s < NA) {
#define PSUM_SINGLE_CASE(V) if (XR == V) fast_fma<NA, V, true>(yacc, pv, s0 + s1a, s0 + s1b);
                PSUM_REP32(PSUM_SINGLE_CASE)
#undef PSUM_SINGLE_CASE
              } else {
#define PSUM_SINGLE_CASE(V) if (XR == V) fast_fma_im<NA, V, true>(yacc, pv, s0, s1a, s1b);
                PSUM_REP32(PSUM_SINGLE_CASE)
#undef PSUM_SINGLE_CASE
              }
            }
          }
        }
        // ---- bundles: one set of partner amplitudes, a register permutation per member ----------
        {
          for (int b = 0; b < ch.nb; ++b) {
            const uint4 B = lds_u4(bun_s + 16u * (uint32_t)b);
            int g = (int)(B.y >> 16);
            const int gend = g + (int)(B.y & 0xffu);
            if (HERM) {  // partner half of a paired bundle (warp-uniform)
              const uint32_t hb = (B.y >> 8) & 0xffu;
              if (hb && ((t0 >> (hb - 1u)) & 1u)) continue;
            }
            double2 pv[NA];
            if (!(B.x >> 31)) {
              const uint32_t pa = tile_s + 16u * (t0 ^ B.x);   // register bits clear on both sides
#pragma unroll
              for (int r = 0; r < NA; ++r) pv[r] = lds_d2(pa + 512u * (uint32_t)r);
            } else {
              const char* pp = reinterpret_cast<const char*>(srcb + (i0 ^ ((uint64_t)B.z | ((uint64_t)B.w << 32))));
#pragma unroll
              for (int r = 0; r < NA; ++r) pv[r] = __ldg(reinterpret_cast<const double2*>(pp + P.roff[r]));
            }
            for (; g < gend; ++g) {
              const uint4 h0 = lds_u4(grp_s + 48u * (uint32_t)g), hc = lds_u4(grp_s + 48u * (uint32_t)g + 16u);
              const uint32_t var = (h0.y >> 22) & 63u;
              const uint32_t xr = (h0.x >> 5) & (uint32_t)(NA - 1);
              if (var < 2u * NA) {
                // fast member: <= 2 patterns, coefficients inline in the record
                const uint32_t q1 = (uint32_t)__popc(h0.w & t0);
                const uint32_t sg = var >> RB;
                const double s0 = __hiloint2double((int)(hc.y ^ ((uint32_t)__popc(h0.z & t0) << 31)), (int)hc.x);
                const double s1a = __hiloint2double((int)(hc.w ^ (q1 << 31)), (int)hc.z);
                const double s1b = __hiloint2double((int)(hc.w ^ ((q1 + sg) << 31)), (int)hc.z);
                if (IM && ((h0.y >> 19) & 1u)) {
                  switch (xr) {
#define PSUM_FAST_CASE(V) case V: fast_fma_im<NA, V, false>(yacc, pv, s0, s1a, s1b); break;
                    PSUM_REP32(PSUM_FAST_CASE)
#undef PSUM_FAST_CASE
                  }
                } else {
                  const double dp = s0 + s1a, dm = s0 + s1b;
                  switch (xr) {
#define PSUM_FAST_CASE(V) case V: fast_fma<NA, V, false>(yacc, pv, dp, dm); break;
                    PSUM_REP32(PSUM_FAST_CASE)
#undef PSUM_FAST_CASE
                  }
                }
              } else if (var == (uint32_t)kVarQuad) {
                // quadratic member: D from the separable tables of this tile
                const uint32_t qt = qtab_s + 8u * (hc.w >> 16) + 8u * (uint32_t)lane;   // lane tables: [table][lane]
                const uint32_t w = t0 >> (5 + RB);
                const uint32_t qw = qtab_s + 8u * (hc.w >> 16) + 256u * (uint32_t)(1 + nW + RB) + 8u * w;   // w tables
                const uint32_t wsz8 = 8u << nW;
                double v[NA];
#pragma unroll
                for (int m = 0; m < NA; ++m) v[m] = 0.0;
                double a0 = lds_d(qt) + lds_d(qw);
#pragma unroll
                for (int j = 0; j < 5; ++j)
                  if (j < nW) a0 += flip_sign(lds_d(qt + 256u * (uint32_t)(1 + j)), (int)((w >> j) & 1u));
                v[0] = a0;
#pragma unroll
                for (int i = 0; i < RB; ++i)
                  v[1 << i] = lds_d(qt + 256u * (uint32_t)(1 + nW + i)) + lds_d(qw + wsz8 * (uint32_t)(1 + i));
#pragma unroll
                for (int m = 0; m < NA; ++m)
                  if (__popc(m) == 2) v[m] = lds_d(qw - 8u * w + wsz8 * (uint32_t)(1 + RB) + 8u * (uint32_t)m);
                // Walsh-Hadamard over the register bits: D[r] = sum_m (-1)^{popc(m & r)} v[m]
#pragma unroll
                for (int i = 0; i < RB; ++i) {
#pragma unroll
                  for (int m = 0; m < NA; ++m) {
                    if (!(m & (1 << i))) {
                      const double a = v[m], b = v[m | (1 << i)];
                      v[m] = a + b;
                      v[m | (1 << i)] = a - b;
                    }
                  }
                }
                switch (xr) {
#define PSUM_PERM_CASE(V) case V: perm_fma<NA, V, false, NA>(yacc, pv, v); break;
                  PSUM_REP32(PSUM_PERM_CASE)
#undef PSUM_PERM_CASE
                }
              } else {
                // general member: class entry lists
                const uint4 h2 = lds_u4(grp_s + 48u * (uint32_t)g + 32u);
                int p = (int)(h0.y & 0xffffu);
                if ((h0.y >> 21) & 1u) {
                  // flat: one real class-0 entry -> the same D for all amplitudes of the thread
                  const double sf = pattern_sum(pat_s, p, (int)(h2.z & 0x7ffu), t0);
                  switch (xr) {
#define PSUM_FLAT_CASE(V) case V: fast_fma<NA, V, false>(yacc, pv, sf, sf); break;
                    PSUM_REP32(PSUM_FLAT_CASE)
#undef PSUM_FLAT_CASE
                  }
                } else {
                  const int nent = (int)((h0.y >> 16) & 7u);
                  uint64_t ents = (uint64_t)h2.z | ((uint64_t)h2.w << 32);
                  double D[ND];
#pragma unroll
                  for (int r = 0; r < ND; ++r) D[r] = 0.0;
                  for (int e = 0; e < nent; ++e) {
                    const uint32_t ent = (uint32_t)ents & 0xffffu;
                    ents >>= 16;
                    const int cnt = (int)(ent & 0x7ffu);
                    double sc;
                    if (cnt == 1) {   // the common case: one pattern per class entry
                      const uint4 rec = lds_u4(pat_s + 16u * (uint32_t)p);
                      sc = __hiloint2double((int)(rec.y ^ ((uint32_t)(__popc(rec.z & t0)) << 31)), (int)rec.x);
                      ++p;
                    } else {
                      sc = pattern_sum(pat_s, p, cnt, t0);
                    }
                    const uint32_t zr = (ent >> 11) & (uint32_t)(NA - 1);
                    if (IM && (ent >> 11) >= (uint32_t)NA) {
                      switch (zr) {
#define PSUM_CLASS_CASE(C) case C: add_class<NA, (C < NA ? NA + C : ND), ND>(D, sc); break;
                        PSUM_REP32(PSUM_CLASS_CASE)
#undef PSUM_CLASS_CASE
                      }
                    } else {
                      switch (zr) {
#define PSUM_CLASS_CASE(C) case C: add_class<NA, (C < NA ? C : ND), ND>(D, sc); break;
                        PSUM_REP32(PSUM_CLASS_CASE)
#undef PSUM_CLASS_CASE
                      }
                    }
                  }
                  if (IM && ((h0.y >> 20) & 1u)) {
                    switch (xr) {
#define PSUM_PERM_CASE(V) case V: perm_fma<NA, V, IM, ND>(yacc, pv, D); break;
                      PSUM_REP32(PSUM_PERM_CASE)
#undef PSUM_PERM_CASE
                    }
                  } else {
                    switch (xr) {
#define PSUM_PERM_CASE(V) case V: perm_fma<NA, V, false, ND>(yacc, pv, D); break;
                      PSUM_REP32(PSUM_PERM_CASE)
#undef PSUM_PERM_CASE
                    }
                  }
                }
              }
            }
          }
        }
      }
      // epilogue of the round: y and/or the expectation partial
#pragma unroll
      for (int r = 0; r < NA; ++r) {
        if (EXPECT) {
          const double2 b = bra ? __ldg(reinterpret_cast<const double2*>(
                                      reinterpret_cast<const char*>(bra + i0) + P.roff[r]))
                                : lds_d2(tile_s + 16u * (t0 | ((uint32_t)r << 5)));
          eacc.x += b.x * yacc[r].x + b.y * yacc[r].y;
          eacc.y += b.x * yacc[r].y - b.y * yacc[r].x;
        }
        if (STORE)
          *reinterpret_cast<double2*>(reinterpret_cast<char*>(y + i0) + P.roff[r]) = yacc[r];
      }
    }
    if (nbatch > 1) {
      // this state's share of the tile goes to the warp's own accumulator (fixed order: deterministic)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        eacc.x += __shfl_down_sync(0xffffffffu, eacc.x, o);
        eacc.y += __shfl_down_sync(0xffffffffu, eacc.y, o);
      }
      if (lane == 0) {
        double2 a = wacc[warp * kMaxBatch + bb];
        a.x += eacc.x;
        a.y += eacc.y;
        wacc[warp * kMaxBatch + bb] = a;
      }
      eacc = make_double2(0.0, 0.0);
    }
   }
  }
  if (EXPECT) {
    if (nbatch > 1) {
      __syncthreads();
      for (int bb = tid; bb < nbatch; bb += NT) {
        double2 tot = make_double2(0.0, 0.0);
        for (int w = 0; w < NT / 32; ++w) {
          tot.x += wacc[w * kMaxBatch + bb].x;
          tot.y += wacc[w * kMaxBatch + bb].y;
        }
        if (HERM) tot.y = 0.0;
        partial[(size_t)bb * gridDim.x + blockIdx.x] = tot;
      }
      return;
    }
    if (HERM) eacc.y = 0.0;  // exactly zero for a Hermitian operator; the paired evaluation keeps 2 Re z
    double2 tot = block_reduce(eacc, red);
    if (tid == 0) partial[blockIdx.x] = tot;
  }
}

#ifdef PSUM_SMALL_KERNELS   // defined by psum.cu only: the kernels below are not templates
// ------------------------------------------------------------------------------------------
// Kernel A: streaming form, one amplitude per thread (grid-stride).  Any n_loc >= 0.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_stream(const double2* __restrict__ src, const double2* __restrict__ bra, double2* __restrict__ y,
         const SimpleGroup* __restrict__ groups, int ngroups, const SimpleTerm* __restrict__ terms,
         uint64_t n_amps, int accumulate, double2* __restrict__ partial) {
  __shared__ double2 red[kThreads / 32];
  double2 eacc = make_double2(0.0, 0.0);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps;
       i += (uint64_t)gridDim.x * blockDim.x) {
    double2 acc = make_double2(0.0, 0.0);
    for (int g = 0; g < ngroups; ++g) {
      const SimpleGroup G = groups[g];
      double dre = 0.0, dim = 0.0;
      for (int t = G.t0; t < G.t0 + G.nt; ++t) {
        const SimpleTerm T = terms[t];
        const int s = __popcll(i & T.z) & 1;
        dre += flip_sign(T.re, s);
        dim += flip_sign(T.im, s);
      }
      const double2 p = src[i ^ G.xl];
      acc.x += dre * p.x - dim * p.y;
      acc.y += dre * p.y + dim * p.x;
    }
    if (y && accumulate) {  // the expectation below is then <bra|y> of the final y
      const double2 old = y[i];
      acc.x += old.x;
      acc.y += old.y;
    }
    if (partial) {
      const double2 b = bra[i];
      eacc.x += b.x * acc.x + b.y * acc.y;
      eacc.y += b.x * acc.y - b.y * acc.x;
    }
    if (y) y[i] = acc;
  }
  if (partial) {
    double2 tot = block_reduce(eacc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
  }
}

// dense matrix of the sum (row-major N x N, zero-initialised by the caller): H[i][i ^ x_g] = D_g(i).
// One thread per (row, group).  Small n only (to_matrix / the k >= 2^n - 1 branch of the eigensolver).
__global__ void __launch_bounds__(kThreads)
k_to_dense(double2* __restrict__ H, const SimpleGroup* __restrict__ groups, int ngroups,
           const SimpleTerm* __restrict__ terms, uint64_t n_amps) {
  const uint64_t total = n_amps * (uint64_t)ngroups;
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = e % n_amps;
    const SimpleGroup G = groups[e / n_amps];
    double dre = 0.0, dim = 0.0;
    for (int t = G.t0; t < G.t0 + G.nt; ++t) {
      const SimpleTerm T = terms[t];
      const int s = __popcll(i & T.z) & 1;
      dre += flip_sign(T.re, s);
      dim += flip_sign(T.im, s);
    }
    H[i * n_amps + (i ^ G.xl)] = make_double2(dre, dim);
  }
}

// Pauli decomposition of a dense matrix (inverse of k_to_dense): for every x-mask the generalised
// diagonal d_x[i] = M[i][i ^ x] is Walsh-Hadamard transformed over i,
//   W[x][z] = 2^-n sum_i (-1)^{popc(i & z)} M[i][i ^ x]      ( = weight(x, z) * (-i)^{popc(x & z)} ).
// One CTA per x; the row lives in shared memory (n <= 12: 64 KB).
__global__ void __launch_bounds__(kThreads)
k_pauli_decompose(const double2* __restrict__ M, double2* __restrict__ Wout, int n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* row = reinterpret_cast<double2*>(smem_raw);
  const uint64_t N = 1ull << n;
  for (uint64_t x = blockIdx.x; x < N; x += gridDim.x) {
    __syncthreads();
    for (uint64_t i = threadIdx.x; i < N; i += blockDim.x) row[i] = M[i * N + (i ^ x)];
    __syncthreads();
    for (uint64_t h = 1; h < N; h <<= 1) {
      for (uint64_t p = threadIdx.x; p < N / 2; p += blockDim.x) {
        const uint64_t lo = ((p & ~(h - 1)) << 1) | (p & (h - 1)), hi = lo | h;
        const double2 a = row[lo], b = row[hi];
        row[lo] = make_double2(a.x + b.x, a.y + b.y);
        row[hi] = make_double2(a.x - b.x, a.y - b.y);
      }
      __syncthreads();
    }
    const double inv = 1.0 / (double)N;
    for (uint64_t z = threadIdx.x; z < N; z += blockDim.x)
      Wout[x * N + z] = make_double2(row[z].x * inv, row[z].y * inv);
  }
}

// batched expectation: out[b] (+)= sum_i partial[b * cols + i]   (one block per state)
__global__ void __launch_bounds__(kThreads)
k_accum_batch(const double2* __restrict__ partial, int cols, double2* __restrict__ out, int accumulate) {
  __shared__ double2 red[kThreads / 32];
  const double2* row = partial + (size_t)blockIdx.x * cols;
  double2 v = make_double2(0.0, 0.0);
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    v.x += row[i].x;
    v.y += row[i].y;
  }
  double2 tot = block_reduce(v, red);
  if (threadIdx.x == 0) {
    if (accumulate) {
      tot.x += out[blockIdx.x].x;
      tot.y += out[blockIdx.x].y;
    }
    out[blockIdx.x] = tot;
  }
}

// sum `count` complex partials in a fixed order -> out[0]  (single block)
__global__ void __launch_bounds__(kThreads)
k_reduce_partials(const double2* __restrict__ partial, int count, double2* __restrict__ out,
                  int accumulate) {
  __shared__ double2 red[kThreads / 32];
  double2 v = make_double2(0.0, 0.0);
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    v.x += partial[i].x;
    v.y += partial[i].y;
  }
  double2 tot = block_reduce(v, red);
  if (threadIdx.x == 0) {
    if (accumulate) {
      tot.x += out[0].x;
      tot.y += out[0].y;
    }
    out[0] = tot;
  }
}

// ------------------------------------------------------------------------------------------
// vector helpers (Lanczos, synthetic states)
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// uniform in [-0.5, 0.5), exactly representable -> bit-identical on host and device
__host__ __device__ __forceinline__ double counter_uniform(uint64_t seed, uint64_t ctr) {
  const uint64_t h = splitmix64(seed ^ splitmix64(ctr));
  return (double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

__global__ void k_fill_state(double2* psi, uint64_t n_amps, uint64_t first, uint64_t seed) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t gi = first + i;
    psi[i] = make_double2(counter_uniform(seed, 2 * gi), counter_uniform(seed, 2 * gi + 1));
  }
}

__global__ void __launch_bounds__(kThreads)
k_dot(const double2* __restrict__ a, const double2* __restrict__ b, uint64_t n_amps,
      double2* __restrict__ partial) {
  __shared__ double2 red[kThreads / 32];
  double2 v = make_double2(0.0, 0.0);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const double2 p = a[i], q = b[i];
    v.x += p.x * q.x + p.y * q.y;
    v.y += p.x * q.y - p.y * q.x;
  }
  double2 tot = block_reduce(v, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

// y = y * s  (complex scalar from device memory: sdev[0] if sdev else (sre, sim))
__global__ void k_scale(double2* a, uint64_t n_amps, double sre, double sim) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const double2 p = a[i];
    a[i] = make_double2(p.x * sre - p.y * sim, p.x * sim + p.y * sre);
  }
}

__global__ void k_axpy(double2* __restrict__ y, const double2* __restrict__ x, uint64_t n_amps,
                       double are, double aim) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const double2 p = x[i];
    double2 q = y[i];
    q.x += are * p.x - aim * p.y;
    q.y += are * p.y + aim * p.x;
    y[i] = q;
  }
}

// Lanczos fused update: w = w - alpha*v - beta*vprev ; partial = |w|^2
__global__ void __launch_bounds__(kThreads)
k_lanczos_update(double2* __restrict__ w, const double2* __restrict__ v,
                 const double2* __restrict__ vprev, uint64_t n_amps, double alpha, double beta,
                 double2* __restrict__ partial) {
  __shared__ double2 red[kThreads / 32];
  double2 acc = make_double2(0.0, 0.0);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps;
       i += (uint64_t)gridDim.x * blockDim.x) {
    double2 q = w[i];
    const double2 p = v[i];
    q.x -= alpha * p.x;
    q.y -= alpha * p.y;
    if (vprev) {
      const double2 o = vprev[i];
      q.x -= beta * o.x;
      q.y -= beta * o.y;
    }
    w[i] = q;
    acc.x += q.x * q.x + q.y * q.y;
  }
  double2 tot = block_reduce(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

// out[j] = <basis_j | w> for j in [0,m): one block row per j (gridDim.y = m)
__global__ void __launch_bounds__(kThreads)
k_multi_dot(const double2* __restrict__ basis, uint64_t stride, const double2* __restrict__ w,
            uint64_t n_amps, double2* __restrict__ partial) {
  __shared__ double2 red[kThreads / 32];
  const double2* a = basis + (uint64_t)blockIdx.y * stride;
  double2 v = make_double2(0.0, 0.0);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const double2 p = a[i], q = w[i];
    v.x += p.x * q.x + p.y * q.y;
    v.y += p.x * q.y - p.y * q.x;
  }
  double2 tot = block_reduce(v, red);
  if (threadIdx.x == 0) partial[(uint64_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
}

// w -= sum_j coef[j] * basis_j   (coef on device, m vectors)
__global__ void k_multi_axpy(double2* __restrict__ w, const double2* __restrict__ basis,
                             uint64_t stride, int m, const double2* __restrict__ coef,
                             uint64_t n_amps, double sign) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps;
       i += (uint64_t)gridDim.x * blockDim.x) {
    double2 q = w[i];
    for (int j = 0; j < m; ++j) {
      const double2 c = coef[j];
      const double2 p = basis[(uint64_t)j * stride + i];
      q.x += sign * (c.x * p.x - c.y * p.y);
      q.y += sign * (c.x * p.y + c.y * p.x);
    }
    w[i] = q;
  }
}

// reduce m rows of `cols` partials each -> out[j]
__global__ void __launch_bounds__(kThreads)
k_reduce_rows(const double2* __restrict__ partial, int cols, double2* __restrict__ out) {
  __shared__ double2 red[kThreads / 32];
  const double2* row = partial + (uint64_t)blockIdx.x * cols;
  double2 v = make_double2(0.0, 0.0);
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    v.x += row[i].x;
    v.y += row[i].y;
  }
  double2 tot = block_reduce(v, red);
  if (threadIdx.x == 0) out[blockIdx.x] = tot;
}

// ------------------------------------------------------------------------------------------
// gate-level state preparation (ansatz -> psi on the device, "next" row 1 of SURVEY 8f)
// ------------------------------------------------------------------------------------------
struct Gate2x2 {
  double2 u00, u01, u10, u11;
};
// psi <- (U on qubit q) psi ; one thread per amplitude pair
__global__ void k_gate_1q(double2* __restrict__ psi, uint64_t n_pairs, int q, Gate2x2 U) {
  const uint64_t low = (1ull << q) - 1ull;
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs;
       p += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i0 = ((p & ~low) << 1) | (p & low);
    const uint64_t i1 = i0 | (1ull << q);
    const double2 a = psi[i0], b = psi[i1];
    double2 r0, r1;
    r0.x = U.u00.x * a.x - U.u00.y * a.y + U.u01.x * b.x - U.u01.y * b.y;
    r0.y = U.u00.x * a.y + U.u00.y * a.x + U.u01.x * b.y + U.u01.y * b.x;
    r1.x = U.u10.x * a.x - U.u10.y * a.y + U.u11.x * b.x - U.u11.y * b.y;
    r1.y = U.u10.x * a.y + U.u10.y * a.x + U.u11.x * b.y + U.u11.y * b.x;
    psi[i0] = r0;
    psi[i1] = r1;
  }
}
// controlled-X (swap) or controlled-Z (sign) on the quarter of the amplitudes with control = 1
__global__ void k_gate_cx_cz(double2* __restrict__ psi, uint64_t n_quads, int c, int t, int is_cz) {
  const int lo = c < t ? c : t, hi = c < t ? t : c;
  const uint64_t mlo = (1ull << lo) - 1ull, mhi = (1ull << (hi - 1)) - 1ull;
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_quads;
       p += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t i = ((p & ~mlo) << 1) | (p & mlo);     // insert a 0 at bit lo
    i = ((i & ~((mhi << 1) | 1ull)) << 1) | (i & ((mhi << 1) | 1ull));  // insert a 0 at bit hi
    const uint64_t i10 = i | (1ull << c);           // control 1, target 0
    const uint64_t i11 = i10 | (1ull << t);
    if (is_cz) {
      double2 v = psi[i11];
      psi[i11] = make_double2(-v.x, -v.y);
    } else {
      const double2 a = psi[i10], b = psi[i11];
      psi[i10] = b;
      psi[i11] = a;
    }
  }
}
// A block of gates whose qubits fit one tile: the tile (the 2^L amplitudes that differ only in the
// block's free qubits) is staged in shared memory once, every gate of the block is applied there, and
// the tile is written back -- one HBM sweep per BLOCK instead of one per gate.
struct GateRec {      // 80 bytes
  int32_t kind;       // 0: 2x2 unitary u on tile bit a; 1: CX (control b, target a); 2: CZ (a, b)
  int32_t a, b;
  int32_t pad;
  double u[8];        // row-major {u00, u01, u10, u11} as (re, im) pairs
};
struct GateBlockParams {
  int L, n_nonfree;
  uint64_t ntiles;
  uint8_t fbit[16];   // qubit of tile bit j (ascending; the three lowest qubits come first)
  uint8_t nbit[64];   // positions of the non-free qubits, ascending
};
__global__ void __launch_bounds__(kThreads, 2)
k_gate_block(double2* __restrict__ psi, const GateBlockParams P, const GateRec* __restrict__ gates, int ngates) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* tile = reinterpret_cast<double2*>(smem_raw);
  const int L = P.L, tid = threadIdx.x;
  const uint32_t tile_amps = 1u << L;
  uint64_t* offT = reinterpret_cast<uint64_t*>(tile + tile_amps);   // [2][64]: tile index -> offset, 6 bits per lookup
  uint64_t* baseT = offT + 128;                                     // [9][64]: tile id -> base
  for (int e = tid; e < 128; e += kThreads) {
    const int k = e >> 6, v = e & 63;
    uint64_t o = 0;
    for (int j = 0; j < 6 && 6 * k + j < L; ++j) o |= (uint64_t)((v >> j) & 1) << P.fbit[6 * k + j];
    offT[e] = o;
  }
  const int nbt = (P.n_nonfree + 5) / 6;
  for (int e = tid; e < nbt * 64; e += kThreads) {
    const int k = e >> 6, v = e & 63;
    uint64_t o = 0;
    for (int j = 0; j < 6 && 6 * k + j < P.n_nonfree; ++j) o |= (uint64_t)((v >> j) & 1) << P.nbit[6 * k + j];
    baseT[e] = o;
  }
  __syncthreads();
  for (uint64_t tile_id = blockIdx.x; tile_id < P.ntiles; tile_id += gridDim.x) {
    uint64_t base = 0;
    for (int k = 0; k < nbt; ++k) base |= baseT[k * 64 + (int)((tile_id >> (6 * k)) & 63ull)];
    __syncthreads();
    for (uint32_t e = tid; e < tile_amps; e += kThreads)
      cp_async16(&tile[e], &psi[base | offT[e & 63u] | offT[64 + (e >> 6)]]);
    cp_async_wait_all();
    __syncthreads();
    for (int k = 0; k < ngates; ++k) {
      const GateRec G = gates[k];
      if (G.kind == 0) {
        const uint32_t a = (uint32_t)G.a, lowm = (1u << a) - 1u;
        for (uint32_t p = tid; p < tile_amps / 2; p += kThreads) {
          const uint32_t i0 = ((p & ~lowm) << 1) | (p & lowm), i1 = i0 | (1u << a);
          const double2 v0 = tile[i0], v1 = tile[i1];
          double2 r0, r1;
          r0.x = G.u[0] * v0.x - G.u[1] * v0.y + G.u[2] * v1.x - G.u[3] * v1.y;
          r0.y = G.u[0] * v0.y + G.u[1] * v0.x + G.u[2] * v1.y + G.u[3] * v1.x;
          r1.x = G.u[4] * v0.x - G.u[5] * v0.y + G.u[6] * v1.x - G.u[7] * v1.y;
          r1.y = G.u[4] * v0.y + G.u[5] * v0.x + G.u[6] * v1.y + G.u[7] * v1.x;
          tile[i0] = r0;
          tile[i1] = r1;
        }
      } else {
        const uint32_t lo = (uint32_t)(G.a < G.b ? G.a : G.b), hi = (uint32_t)(G.a < G.b ? G.b : G.a);
        const uint32_t mlo = (1u << lo) - 1u, mhi = (1u << hi) - 1u;
        for (uint32_t p = tid; p < tile_amps / 4; p += kThreads) {
          uint32_t i = ((p & ~mlo) << 1) | (p & mlo);            // insert a 0 at bit lo
          i = ((i & ~mhi) << 1) | (i & mhi);                      // insert a 0 at bit hi
          const uint32_t i11 = i | (1u << G.a) | (1u << G.b);
          if (G.kind == 1) {                                      // CX: swap target where control = 1
            const uint32_t i10 = i | (1u << G.b);
            const double2 v = tile[i10];
            tile[i10] = tile[i11];
            tile[i11] = v;
          } else {                                                // CZ
            const double2 v = tile[i11];
            tile[i11] = make_double2(-v.x, -v.y);
          }
        }
      }
      __syncthreads();
    }
    for (uint32_t e = tid; e < tile_amps; e += kThreads)
      psi[base | offT[e & 63u] | offT[64 + (e >> 6)]] = tile[e];
  }
}

__global__ void k_set_basis_state(double2* __restrict__ psi, uint64_t n_amps, uint64_t index) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps;
       i += (uint64_t)gridDim.x * blockDim.x)
    psi[i] = make_double2(i == index ? 1.0 : 0.0, 0.0);
}

// diagonal of an all-Z sum: d[i] = sum_k re_k (-1)^{popc(i & z_k)}
__global__ void k_diag(double* __restrict__ d, const SimpleTerm* __restrict__ terms, int nterms,
                       uint64_t n_amps, uint64_t first_index_sign_mask_unused) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps;
       i += (uint64_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int t = 0; t < nterms; ++t) v += flip_sign(terms[t].re, __popcll(i & terms[t].z) & 1);
    d[i] = v;
  }
}

// lexicographic (value, index) minimum strictly greater than (lo_val, lo_idx)
struct ValIdx {
  double v;
  uint64_t i;
};
__device__ __forceinline__ bool vi_less(const ValIdx& a, const ValIdx& b) {
  return a.v < b.v || (a.v == b.v && a.i < b.i);
}
__global__ void __launch_bounds__(kThreads)
k_argmin_after(const double* __restrict__ d, uint64_t n_amps, ValIdx lo, int have_lo,
               ValIdx* __restrict__ partial) {
  __shared__ ValIdx red[kThreads / 32];
  ValIdx best{1.0 / 0.0, ~0ull};
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps;
       i += (uint64_t)gridDim.x * blockDim.x) {
    ValIdx c{d[i], i};
    if (have_lo && !vi_less(lo, c)) continue;
    if (vi_less(c, best)) best = c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ValIdx other{__shfl_down_sync(0xffffffffu, best.v, o), __shfl_down_sync(0xffffffffu, best.i, o)};
    if (vi_less(other, best)) best = other;
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int j = 1; j < kThreads / 32; ++j)
      if (vi_less(red[j], best)) best = red[j];
    partial[blockIdx.x] = best;
  }
}

#endif  // PSUM_SMALL_KERNELS

}  // namespace psum
