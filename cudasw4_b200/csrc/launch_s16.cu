// Instantiations + launcher of the two-rows-per-step packed kernel (kernels_s16.cuh), one per length class.
#include <cstdlib>
#include "launch.hpp"

namespace sw4 {

template <int R>
static cudaError_t launch_one(const S16Params& prm, int grid, cudaStream_t stream) {
    static bool configured[64] = {};
    cudaError_t e = ensure_smem_attr(sw_s16_kernel<R>, s16_smem_bytes<R>(), configured);
    if (e != cudaSuccess) return e;
    return launch_clustered(sw_s16_kernel<R>, prm, grid, kS16Threads, s16_smem_bytes<R>(), stream);
}

cudaError_t launch_s16(int R, const S16Params& prm, int grid, cudaStream_t stream) {
    switch (R) {
        case 4: return launch_one<4>(prm, grid, stream);
        case 6: return launch_one<6>(prm, grid, stream);
        case 8: return launch_one<8>(prm, grid, stream);
        case 10: return launch_one<10>(prm, grid, stream);
        case 12: return launch_one<12>(prm, grid, stream);
        case 14: return launch_one<14>(prm, grid, stream);
        case 16: return launch_one<16>(prm, grid, stream);
        case 18: return launch_one<18>(prm, grid, stream);
        case 20: return launch_one<20>(prm, grid, stream);
        case 22: return launch_one<22>(prm, grid, stream);
        case 24: return launch_one<24>(prm, grid, stream);
        case 26: return launch_one<26>(prm, grid, stream);
        case 28: return launch_one<28>(prm, grid, stream);
        case 30: return launch_one<30>(prm, grid, stream);
        case 32: return launch_one<32>(prm, grid, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sw4
