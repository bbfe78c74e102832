// Exact 32-bit twin of the CTA-wide array kernel (kernels_s16_long.cuh): the W warps of a CTA work on one stream of
// 512-column blocks of ONE subject at a time, so a 35 k x 35 k re-scoring takes ceil(69 / W) periods instead of 69.
//
// Role: exact re-scoring of the subjects whose packed 16-bit score reached the overflow threshold.
// Replaces the reference's NW_local_affine_multi_pass_dpx_s32 + device-side launcher
// (src/dpx_s32_kernels.cuh:273-968, 1180-1290), which gives each such subject to a single warp-sized block.
// The work list lives in device memory (filled by the 16-bit kernels of the same scan); the grid is launched at full
// width and CTAs beyond the list length retire at once. Profile: one plane prof[s][row] = M[q_row][s] as int32.
#pragma once
#include "kernels_s16_long.cuh"
#include "kernels_s32.cuh"

namespace sw4 {

struct S32LongParams {
    const uint8_t* chars;        // residue codes, makedb layout (shard-local)
    const size_t* offsets;
    const int32_t* lengths;
    const int32_t* list;         // local subject indices to score
    const int* listCountPtr;     // number of entries (device memory)
    int* ticket;
    int warps;                   // W
    int ringSlots;               // S
    const int32_t* prof;         // [21][profStride] M[q_p][s]; rows p >= qlen hold kNegS32
    int profStride;
    int qlen;
    int period;
    int gop, gex;
    int32_t* scores;
    int2* border;                // [gridDim.x][borderStride]
    int borderStride;
};

static inline int s32_long_smem_bytes(int warps) {
    return 21 * (s16_long_ring_slots(warps) + 32) * 4 + warps * kLongFifoRows * 8 + 2 * kLongMaxWarps * 16;
}

__device__ __forceinline__ void long_ring_fill_s32(uint32_t base, int rowWords, int S, const S32LongParams& prm, int slot0, int p0) {
    for (int id = threadIdx.x; id < 21 * 4; id += blockDim.x) {
        const int c = id & 3, s = id >> 2;
        const int32_t* src = prm.prof + (size_t)s * prm.profStride + p0 + 4 * c;
        const uint32_t dst = base + (s * rowWords + slot0 + 4 * c) * 4;
        cp_async16(dst, src);
        if (slot0 == 0) cp_async16(dst + S * 4, src);
    }
}

// GAPS: gap-score set compiled in as instruction immediates (s16_gap_set in kernels_s16.cuh; 0 = run-time values)
template <int GAPS = 0>
__global__ void __launch_bounds__(kLongMaxWarps * 32, 1) sw_s32_long_kernel(const S32LongParams prm) {
    constexpr int R = kLongR;
    extern __shared__ __align__(16) unsigned char smem[];
    const int count = *prm.listCountPtr;
    if ((int)blockIdx.x >= count) return;  // every CTA that stays takes at least one subject's worth of tickets
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int W = prm.warps, S = prm.ringSlots, P = prm.period;
    const int rowWords = S + 32;
    const uint32_t ringBase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t fifoBase = ringBase + 21 * rowWords * 4;
    volatile int4* desc = reinterpret_cast<volatile int4*>(smem + 21 * rowWords * 4 + W * kLongFifoRows * 8);
    const uint32_t fifoIn = fifoBase + w * (kLongFifoRows * 8);
    const uint32_t fifoOut = fifoIn + kLongFifoRows * 8;
    int2* border = prm.border + (size_t)blockIdx.x * prm.borderStride;
    const int gop = GAPS > 0 ? (int)(short)(s16_gap_set(GAPS).x & 0xffffu) : prm.gop;
    const int gex = GAPS > 0 ? (int)(short)(s16_gap_set(GAPS).y & 0xffffu) : prm.gex;

    for (int i = threadIdx.x; i < 21 * rowWords; i += blockDim.x) reinterpret_cast<int*>(smem)[i] = kNegS32;
    __syncthreads();
    long_ring_fill_s32(ringBase, rowWords, S, prm, 0, 0);
    cp_async_commit();

    uint32_t a[R];  // ring byte address of this column's profile row (lane phase folded in)
    int Hp[R], F[R];
    int mx = 0, Elast = kNegS32, HinPrev = 0;
#pragma unroll
    for (int j = 0; j < R; j++) { a[j] = ringBase; Hp[j] = 0; F[j] = kNegS32; }
    const int skew = kLongLag * w + lane;
    int p = skew == 0 ? 0 : P - skew;
    int xs = skew == 0 ? 0 : S - skew;
    const int pRestart = lane == 0 ? 0 : P - lane;
    bool alive = true, haveWork = false, useBorder = false;
    int subj = -1, periodIndex = 0;
    // warp 0 only: where the CTA's block stream stands
    int curSubj = -1, curBlock = 0, curLeft = 0;
    bool streamEnded = false;

    auto restart = [&]() {
        __syncwarp();
        if (haveWork) {
            int r = mx;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r = max(r, __shfl_xor_sync(0xffffffffu, r, o));
            // the subject's first block replaces the saturated 16-bit score, later blocks (which finish later) raise it
            if (lane == 0) {
                if (useBorder) atomicMax(prm.scores + subj, r);
                else atomicExch(prm.scores + subj, r);
            }
        }
        if (w == 0) {
            for (int i = 0; i < W; i++) {
                int4 dd = make_int4(0, -1, 0, 0);
                bool first = false;
                if (!streamEnded && curLeft == 0) {
                    int item = 0;
                    if (lane == 0) item = atomicAdd(prm.ticket, 1);
                    item = __shfl_sync(0xffffffffu, item, 0);
                    if (item >= count) {
                        streamEnded = true;
                    } else {
                        curSubj = prm.list[item];
                        curLeft = max(1, (prm.lengths[curSubj] + kLongBlockCols - 1) / kLongBlockCols);
                        curBlock = 0;
                        first = true;
                    }
                }
                if (!streamEnded) {
                    dd = make_int4(curBlock, curSubj, 0, 1 | (first ? 2 : 0));
                    curBlock++;
                    curLeft--;
                }
                if (lane == 0) {
                    volatile int4* slot = desc + (periodIndex & 1) * kLongMaxWarps + i;
                    slot->x = dd.x; slot->y = dd.y; slot->z = dd.z; slot->w = dd.w;
                }
            }
            __syncwarp();
        }
        int4 d;
        {
            volatile int4* slot = desc + (periodIndex & 1) * kLongMaxWarps + w;
            d.x = slot->x; d.y = slot->y; d.z = slot->z; d.w = slot->w;
        }
        periodIndex++;
        haveWork = (d.w & 1) != 0;
        alive = haveWork;
        if (haveWork) {
            subj = d.y;
            useBorder = (d.w & 2) == 0;
            const int len = prm.lengths[subj];
            const int col0 = d.x * kLongBlockCols + lane * R;
            const uint8_t* s = prm.chars + prm.offsets[subj];  // 4-byte aligned (makedb pads every sequence to 4)
#pragma unroll
            for (int b = 0; b < R / 4; b++) {
                uint32_t word = 0x14141414u;
                if (col0 + 4 * b < len) word = __ldg(reinterpret_cast<const uint32_t*>(s + col0 + 4 * b));
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    uint32_t code = (word >> (8 * h)) & 0xffu;
                    if (col0 + 4 * b + h >= len || code > 20u) code = 20u;
                    a[b * 4 + h] = ringBase + (code * rowWords + xs) * 4;
                }
            }
#pragma unroll
            for (int j = 0; j < R; j++) { Hp[j] = 0; F[j] = kNegS32; }
            HinPrev = 0;
            mx = 0;
        }
        __syncwarp();
    };

    int sfill = kLongBatch % S, pfill = kLongBatch % P;
#pragma unroll 1
    for (int batch = 0;; ++batch) {
        cp_async_wait_all();
        if (!__syncthreads_or(alive)) break;
        long_ring_fill_s32(ringBase, rowWords, S, prm, sfill, pfill);
        if (w == 0 && lane < 8) {
            const int row = pfill + 2 * lane;
            if (row < prm.qlen) cp_async16(fifoBase + (row & (kLongFifoRows - 1)) * 8, border + row);
        }
        cp_async_commit();
        sfill += kLongBatch;
        if (sfill >= S) sfill -= S;
        pfill += kLongBatch;
        if (pfill >= P) pfill -= P;
        if (batch > 0) {
            int delta = kLongBatch * 4;
            xs += kLongBatch;
            if (xs >= S) { xs -= S; delta = (kLongBatch - S) * 4; }
            if (haveWork) {
#pragma unroll
                for (int j = 0; j < R; j++) a[j] += delta;
            }
        }
        static_for<2>([&](auto halfIndex) {
            constexpr int half = decltype(halfIndex)::value;
            if (p == pRestart && alive) restart();
            if (haveWork) {
                static_for<8>([&](auto stepIndex) {
                    constexpr int i = half * 8 + decltype(stepIndex)::value;
                    int Hin = __shfl_up_sync(0xffffffffu, Hp[R - 1], 1);
                    int Ein = __shfl_up_sync(0xffffffffu, Elast, 1);
                    const bool realRow = (unsigned)p < (unsigned)prm.qlen;
                    uint2 bd = make_uint2(0u, 0u);  // only the first lane's real rows read the FIFO (the slot is then final)
                    if (lane == 0 && realRow) bd = lds_u64(fifoIn + (p & (kLongFifoRows - 1)) * 8);
                    if (lane == 0) { Hin = useBorder ? (int)bd.x : 0; Ein = useBorder ? (int)bd.y : kNegS32; }
                    if (!realRow) { Hin = 0; Ein = kNegS32; }
                    int E = Ein;
                    constexpr int kPrefetch = 6;
                    int q0[kPrefetch + 1];
#pragma unroll
                    for (int c = 0; c <= kPrefetch && c < R; c++) q0[c] = (int)lds_u32_imm<i * 4>(a[c]);
                    int d = HinPrev + q0[0];
                    int dPrev = 0;
#pragma unroll
                    for (int j = 0; j < R; j++) {
                        const int n0 = q0[(j + 1) % (kPrefetch + 1)];
                        if (j + 1 + kPrefetch < R) q0[j % (kPrefetch + 1)] = (int)lds_u32_imm<i * 4>(a[j + 1 + kPrefetch]);
                        int dNext = 0;
                        if (j + 1 < R) dNext = Hp[j] + n0;
                        const int h = __vimax3_s32_relu(d, E, F[j]);
                        Hp[j] = h;
                        const int tt = h + gop;
                        E = __viaddmax_s32(E, gex, tt);
                        F[j] = __viaddmax_s32(F[j], gex, tt);
                        if (j & 1) mx = __vimax3_s32(mx, d, dPrev);
                        dPrev = d;
                        d = dNext;
                    }
                    Elast = E;
                    HinPrev = Hin;
                    if (lane == 31 && realRow) {
                        if (w + 1 < W) sts_u64(fifoOut + (p & (kLongFifoRows - 1)) * 8, (uint32_t)Hp[R - 1], (uint32_t)Elast);
                        else border[p] = make_int2(Hp[R - 1], Elast);
                    }
                    if (++p == P) p = 0;
                });
            } else {
                p += 8;
                if (p >= P) p -= P;
            }
        });
    }
}

}  // namespace sw4
