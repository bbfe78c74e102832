// Instantiations + launcher of the full-warp one-row-per-step packed kernel (kernels_s16_wide.cuh).
#include <cstdlib>
#include "launch.hpp"

namespace sw4 {

template <int R, bool MULTI>
static cudaError_t launch_one(const S16WideParams& prm, int grid, cudaStream_t stream) {
    static bool configured[64] = {};
    cudaError_t e = ensure_smem_attr(sw_s16_wide_kernel<R, MULTI>, s16_wide_smem_bytes<R>(), configured);
    if (e != cudaSuccess) return e;
    return launch_clustered(sw_s16_wide_kernel<R, MULTI>, prm, grid, kS16Threads, s16_wide_smem_bytes<R>(), stream);
}

cudaError_t launch_s16_wide(int R, bool multi, const S16WideParams& prm, int grid, cudaStream_t stream) {
    if (multi) return R == 32 ? launch_one<32, true>(prm, grid, stream) : cudaErrorInvalidValue;
    switch (R) {
        case 18: return launch_one<18, false>(prm, grid, stream);
        case 20: return launch_one<20, false>(prm, grid, stream);
        case 22: return launch_one<22, false>(prm, grid, stream);
        case 24: return launch_one<24, false>(prm, grid, stream);
        case 26: return launch_one<26, false>(prm, grid, stream);
        case 28: return launch_one<28, false>(prm, grid, stream);
        case 30: return launch_one<30, false>(prm, grid, stream);
        case 32: return launch_one<32, false>(prm, grid, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sw4
