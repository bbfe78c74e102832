// Instantiations + launcher of the multi-segment form of the two-rows-per-step packed kernel (kernels_s16.cuh, MULTI):
// subjects longer than 16 x R columns, i.e. the 576..1024-column classes as two segments and the long class.
// SW4_GAPS selects the gap-score set this translation unit instantiates.
#include <cstdlib>
#include "launch.hpp"

#ifndef SW4_GAPS
#define SW4_GAPS 0
#endif

namespace sw4 {

template <int R>
static cudaError_t launch_one(const S16Params& prm, int grid, cudaStream_t stream) {
    static bool configured[64] = {};
    auto kernel = sw_s16_kernel<R, true, SW4_GAPS>;
    cudaError_t e = ensure_smem_attr(kernel, s16_smem_bytes<R, true>(), configured);
    if (e != cudaSuccess) return e;
    return launch_clustered(kernel, prm, grid, kS16Threads, s16_smem_bytes<R, true>(), stream);
}

#define SW4_CAT2(a, b) a##b
#define SW4_CAT(a, b) SW4_CAT2(a, b)
cudaError_t SW4_CAT(launch_s16_multi_gaps, SW4_GAPS)(int R, const S16Params& prm, int grid, cudaStream_t stream) {
    switch (R) {
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

#if SW4_GAPS == 0
cudaError_t launch_s16_multi_gaps1(int R, const S16Params& prm, int grid, cudaStream_t stream);
cudaError_t launch_s16_multi_gaps2(int R, const S16Params& prm, int grid, cudaStream_t stream);
cudaError_t launch_s16_multi_gaps3(int R, const S16Params& prm, int grid, cudaStream_t stream);
cudaError_t launch_s16_multi(int R, const S16Params& prm, int grid, cudaStream_t stream) {
    static const bool generic = getenv("SW4_NO_GAP_SETS") != nullptr;
    switch (generic ? 0 : s16_gap_set_for(prm.gop2, prm.gex2)) {
        case 1: return launch_s16_multi_gaps1(R, prm, grid, stream);
        case 2: return launch_s16_multi_gaps2(R, prm, grid, stream);
        case 3: return launch_s16_multi_gaps3(R, prm, grid, stream);
        default: return launch_s16_multi_gaps0(R, prm, grid, stream);
    }
}
#endif

}  // namespace sw4
