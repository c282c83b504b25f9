// psum_kpass.h -- host-callable launchers of the k_pass instantiations.  Every (RB, NT) configuration
// lives in its own translation unit (psum_kpass_inst.cu compiled with -DPSUM_INST_RB/-DPSUM_INST_NT) so
// that the configurations compile in parallel.
#pragma once
#include <cuda_runtime.h>

#include "psum_kernels.cuh"

namespace psum {

// mode bits of a k_pass launch
constexpr int kModeStore = 1, kModeExpect = 2, kModeIm = 4, kModeHerm = 8;

#define PSUM_DECL_KPASS(RB, NT)                                                                          \
  cudaError_t kpass_launch_##RB##_##NT(int mode, int grid, size_t smem, cudaStream_t st, const double2* src, \
                                       const double2* bra, double2* y, const PassParams& pp, double2* partial); \
  cudaError_t kpass_set_smem_##RB##_##NT(int max_smem);
PSUM_DECL_KPASS(3, 256)
PSUM_DECL_KPASS(3, 512)
PSUM_DECL_KPASS(4, 128)
PSUM_DECL_KPASS(4, 256)
#undef PSUM_DECL_KPASS

// one entry point over the compiled (register bits, threads) configurations
inline cudaError_t kpass_launch(int rb, int threads, int mode, int grid, size_t smem, cudaStream_t st, const double2* src,
                                const double2* bra, double2* y, const PassParams& pp, double2* partial) {
  if (rb == 4)
    return threads == 256 ? kpass_launch_4_256(mode, grid, smem, st, src, bra, y, pp, partial)
                          : kpass_launch_4_128(mode, grid, smem, st, src, bra, y, pp, partial);
  return threads == 512 ? kpass_launch_3_512(mode, grid, smem, st, src, bra, y, pp, partial)
                        : kpass_launch_3_256(mode, grid, smem, st, src, bra, y, pp, partial);
}

}  // namespace psum
