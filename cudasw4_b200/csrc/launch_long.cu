// Launchers of the CTA-wide array kernels (kernels_s16_long.cuh, kernels_s32_long.cuh) and of the one-subject-per-warp
// exact 32-bit kernel (kernels_s32.cuh).
#include <cstdlib>
#include "launch.hpp"

namespace sw4 {

cudaError_t launch_s16_long(const S16LongParams& prm, int grid, cudaStream_t stream) {
    static bool configured[64] = {};
    cudaError_t e = ensure_smem_attr(sw_s16_long_kernel<0>, s16_long_smem_bytes(kLongMaxWarps), configured);
    if (e != cudaSuccess) return e;
    sw_s16_long_kernel<0><<<grid, prm.warps * 32, s16_long_smem_bytes(prm.warps), stream>>>(prm);
    return cudaGetLastError();
}

cudaError_t launch_s32_long(const S32LongParams& prm, int grid, cudaStream_t stream) {
    static bool configured[64] = {};
    cudaError_t e = ensure_smem_attr(sw_s32_long_kernel<0>, s32_long_smem_bytes(kLongMaxWarps), configured);
    if (e != cudaSuccess) return e;
    sw_s32_long_kernel<0><<<grid, prm.warps * 32, s32_long_smem_bytes(prm.warps), stream>>>(prm);
    return cudaGetLastError();
}

cudaError_t launch_s32(const S32Params& prm, int blocks, cudaStream_t stream) {
    sw_s32_kernel<0><<<blocks, kS32Threads, 0, stream>>>(prm);
    return cudaGetLastError();
}

}  // namespace sw4
