// Instantiations + launcher of the two-rows-per-step packed kernel (kernels_s16.cuh), one per length class and gap set.
// SW4_GAPS selects the gap-score set this translation unit instantiates (one unit per set so that they build in parallel).
#include <cstdlib>
#include "launch.hpp"

#ifndef SW4_GAPS
#define SW4_GAPS 0
#endif

namespace sw4 {

template <int R>
static cudaError_t launch_one(const S16Params& prm, int grid, cudaStream_t stream) {
    static bool configured[64] = {};
    auto kernel = sw_s16_kernel<R, false, SW4_GAPS>;
    cudaError_t e = ensure_smem_attr(kernel, s16_smem_bytes<R>(), configured);
    if (e != cudaSuccess) return e;
    return launch_clustered(kernel, prm, grid, kS16Threads, s16_smem_bytes<R>(), stream);
}

#define SW4_CAT2(a, b) a##b
#define SW4_CAT(a, b) SW4_CAT2(a, b)
cudaError_t SW4_CAT(launch_s16_gaps, SW4_GAPS)(int R, const S16Params& prm, int grid, cudaStream_t stream) {
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

#if SW4_GAPS == 0
cudaError_t launch_s16_gaps1(int R, const S16Params& prm, int grid, cudaStream_t stream);
cudaError_t launch_s16_gaps2(int R, const S16Params& prm, int grid, cudaStream_t stream);
cudaError_t launch_s16_gaps3(int R, const S16Params& prm, int grid, cudaStream_t stream);
// picks the instantiation whose immediates equal the requested gap scores, else the run-time one
cudaError_t launch_s16(int R, const S16Params& prm, int grid, cudaStream_t stream) {
    static const bool generic = getenv("SW4_NO_GAP_SETS") != nullptr;
    switch (generic ? 0 : s16_gap_set_for(prm.gop2, prm.gex2)) {
        case 1: return launch_s16_gaps1(R, prm, grid, stream);
        case 2: return launch_s16_gaps2(R, prm, grid, stream);
        case 3: return launch_s16_gaps3(R, prm, grid, stream);
        default: return launch_s16_gaps0(R, prm, grid, stream);
    }
}
#endif

}  // namespace sw4
