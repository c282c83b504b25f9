// psum_kpass_inst.cu -- instantiates k_pass for one (RB, NT) configuration (see psum_kpass.h).
#include "psum_kpass.h"

#ifndef PSUM_INST_RB
#error "compile with -DPSUM_INST_RB=3|4 -DPSUM_INST_NT=128|256|512"
#endif

namespace psum {

#define PSUM_CAT2(a, b, c, d) a##b##c##d
#define PSUM_NAME(prefix, RB, NT) PSUM_CAT2(prefix, RB, _, NT)

cudaError_t PSUM_NAME(kpass_launch_, PSUM_INST_RB, PSUM_INST_NT)(int mode, int grid, size_t smem, cudaStream_t st,
                                                              const double2* src, const double2* bra, double2* y,
                                                              const PassParams& pp, double2* partial) {
  constexpr int RB = PSUM_INST_RB, NT = PSUM_INST_NT;
#define GO(S, E, I, H) k_pass<S, E, I, RB, NT, H><<<grid, NT, smem, st>>>(src, bra, y, pp, partial)
  switch (mode) {
    case kModeStore: GO(true, false, false, false); break;
    case kModeStore | kModeIm: GO(true, false, true, false); break;
    case kModeStore | kModeExpect: GO(true, true, false, false); break;
    case kModeStore | kModeExpect | kModeIm: GO(true, true, true, false); break;
    case kModeExpect: GO(false, true, false, false); break;
    case kModeExpect | kModeIm: GO(false, true, true, false); break;
    case kModeExpect | kModeHerm: GO(false, true, false, true); break;
    case kModeExpect | kModeIm | kModeHerm: GO(false, true, true, true); break;
    default: return cudaErrorInvalidValue;
  }
#undef GO
  return cudaGetLastError();
}

cudaError_t PSUM_NAME(kpass_set_smem_, PSUM_INST_RB, PSUM_INST_NT)(int max_smem) {
  constexpr int RB = PSUM_INST_RB, NT = PSUM_INST_NT;
  cudaError_t e;
#define SET(S, E, I, H)                                                                                        \
  e = cudaFuncSetAttribute(k_pass<S, E, I, RB, NT, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem); \
  if (e != cudaSuccess) return e;
  SET(true, false, false, false) SET(true, false, true, false) SET(true, true, false, false)
  SET(true, true, true, false) SET(false, true, false, false) SET(false, true, true, false)
  SET(false, true, false, true) SET(false, true, true, true)
#undef SET
  return cudaSuccess;
}

}  // namespace psum
