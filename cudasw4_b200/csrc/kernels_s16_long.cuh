// CTA-wide systolic array for the multi-segment class: all warps of a CTA work on ONE stream of 512-column blocks.
//
// The one-warp-per-pair kernels (kernels_s16_wide.cuh, and the reference's multi-pass kernels,
// src/dpx_s16_kernels.cuh:290-762, and its one-block-per-subject 32-bit kernel, src/dpx_s32_kernels.cuh:273-968) walk
// the segments of a long subject one after the other, so one pair of 35 k-residue subjects keeps a single warp busy for
// 35 periods while the rest of the GPU runs dry: a database of long sequences is latency-bound, not throughput-bound.
// Here the W warps of a CTA form one array of 32*W lanes. The CTA consumes a stream of blocks (512 columns of a pair of
// subjects each: item 0 block 0, 1, 2, ..., item 1 block 0, ...); block n of the stream runs on warp n mod W during
// period n div W, and warp w trails warp w-1 by kLongLag = 48 steps (32 lanes + one 16-step batch), so that the border
// column (H, E per query row) of block n-1 reaches block n through a 64-row FIFO in shared memory that only needs the
// CTA barrier the ring refill takes every 16 steps anyway. The hand-over from warp W-1 to warp 0 (next period) goes
// through a per-CTA row array in global memory, fetched back by cp.async one batch ahead. Every warp computes in every
// period no matter where the item boundaries fall, and a pair of B blocks is finished after ceil(B / W) periods
// instead of B.
//
// A block can start anywhere in the array, so the fused-pair profile (441 rows, one ring slot per lane of skew) would
// need 441 x 800 words. This kernel instead keeps two 21-row planes, lo[s][slot] = M[q][s] in the low half and
// hi[s][slot] = M[q][s] << 16, and adds both (two VIADD.16x2 on the FMA pipe; the DPX/ALU count per cell-pair, which
// bounds the kernel, is unchanged). 16 register columns per lane.
#pragma once
#include "kernels_s16.cuh"

namespace sw4 {

constexpr int kLongR = 16;                    // register columns per lane
constexpr int kLongBlockCols = 32 * kLongR;   // columns per warp per period
constexpr int kLongLag = 48;                  // steps between consecutive warps
constexpr int kLongFifoRows = 64;
constexpr int kLongBatch = 16;                // steps between CTA barriers / ring refills
constexpr int kLongMaxWarps = 16;

struct S16LongParams {
    const uint16_t* cols;        // the class's pair-blocks: consecutive fused column codes (s0 + 21*s1), 1024 per block
    const S16Item* items;        // [numItems], firstBlock / numSegments count 1024-column blocks
    const int32_t* lengths;      // [numLocalSubjects]
    int numItems;
    int* ticket;
    int warps;                   // W: 2, 4, 8 or 16 (blockDim.x = 32 W)
    int ringSlots;               // S: multiple of 32, >= 48 W + 16
    const uint32_t* profLo;      // [21][profStride]  (M[q_p][s] & 0xffff), rows p >= qlen hold 0x0000c180
    const uint32_t* profHi;      // [21][profStride]  (M[q_p][s] << 16),   rows p >= qlen hold 0xc1800000
    int profStride;
    int qlen;
    int period;                  // P: multiple of 16, >= qlen + 32 and >= 48 W + 64
    uint32_t gop2, gex2;
    int ovfThreshold, statThreshold;
    int32_t* scores;             // must hold -1 (or any value below every score) for the class's subjects at launch
    int32_t* ovfList;
    int* ovfCount;
    int* statCount;
    unsigned long long* elapsedNs;
    uint2* border;               // [gridDim.x][borderStride]
    int borderStride;            // >= round_up(qlen, 2)
};

static inline int s16_long_ring_slots(int warps) { return (kLongLag * warps + 16 + 31) / 32 * 32; }
static inline int s16_long_smem_bytes(int warps) {
    return 2 * 21 * (s16_long_ring_slots(warps) + 32) * 4 + warps * kLongFifoRows * 8 + 2 * kLongMaxWarps * 16;
}

__device__ __forceinline__ void sts_u64(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 lds_u64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

// ring slots [slot0, slot0 + 16) <- profile rows [p0, p0 + 16) of both planes (slot0, p0 multiples of 16)
__device__ __forceinline__ void long_ring_fill(uint32_t loBase, uint32_t hiBase, int rowWords, int S, const S16LongParams& prm,
                                               int slot0, int p0) {
    for (int id = threadIdx.x; id < 2 * 21 * 4; id += blockDim.x) {
        const int c = id & 3, s = (id >> 2) % 21, plane = id / 84;
        const uint32_t* src = (plane ? prm.profHi : prm.profLo) + (size_t)s * prm.profStride + p0 + 4 * c;
        const uint32_t dst = (plane ? hiBase : loBase) + (s * rowWords + slot0 + 4 * c) * 4;
        cp_async16(dst, src);
        if (slot0 == 0) cp_async16(dst + S * 4, src);  // mirror of slots [0, 16) behind the ring's end
    }
}

// GAPS: gap-score set compiled in as instruction immediates (s16_gap_set in kernels_s16.cuh; 0 = run-time values)
template <int GAPS = 0>
__global__ void __launch_bounds__(kLongMaxWarps * 32, 1) sw_s16_long_kernel(const S16LongParams prm) {
    constexpr int R = kLongR;
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long tStart;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tStart));
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int W = prm.warps, S = prm.ringSlots, P = prm.period;
    const int rowWords = S + 32;
    const uint32_t loBase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t hiBase = loBase + 21 * rowWords * 4;
    const uint32_t fifoBase = hiBase + 21 * rowWords * 4;
    volatile int4* desc = reinterpret_cast<volatile int4*>(smem + 2 * 21 * rowWords * 4 + W * kLongFifoRows * 8);
    const uint32_t fifoIn = fifoBase + w * (kLongFifoRows * 8);
    const uint32_t fifoOut = fifoIn + kLongFifoRows * 8;  // the next warp's input (unused by the last warp)
    uint2* border = prm.border + (size_t)blockIdx.x * prm.borderStride;
    const uint32_t NEG2 = ((uint32_t)(uint16_t)kNegS16 << 16) | (uint16_t)kNegS16;
    const uint32_t gop2 = GAPS > 0 ? s16_gap_set(GAPS).x : prm.gop2, gex2 = GAPS > 0 ? s16_gap_set(GAPS).y : prm.gex2;

    // rows "before time 0" and between two periods are gap rows: the ring starts out as -16000 everywhere
    for (int i = threadIdx.x; i < 21 * rowWords; i += blockDim.x) {
        reinterpret_cast<uint32_t*>(smem)[i] = NEG2 & 0xffffu;
        reinterpret_cast<uint32_t*>(smem)[21 * rowWords + i] = NEG2 & 0xffff0000u;
    }
    __syncthreads();
    long_ring_fill(loBase, hiBase, rowWords, S, prm, 0, 0);
    cp_async_commit();

    uint32_t a0[R], a1[R];  // ring byte addresses of this column's two profile rows (lane phase folded in)
    uint32_t Hp[R], F[R];
    uint32_t mx = 0, Elast = NEG2, HinPrev = 0;
#pragma unroll
    for (int j = 0; j < R; j++) { a0[j] = loBase; a1[j] = hiBase; Hp[j] = 0; F[j] = NEG2; }
    // this lane's row in the period-P schedule: (t - 48 w - lane) mod P; its ring slot at the start of the batch
    const int skew = kLongLag * w + lane;
    int p = skew == 0 ? 0 : P - skew;
    int xs = skew == 0 ? 0 : S - skew;
    const int pRestart = lane == 0 ? 0 : P - lane;
    bool alive = true, haveWork = false, useBorder = false, isLast = false;
    int sub0 = -1, sub1 = -1, periodIndex = 0;
    // warp 0 only: where the CTA's block stream stands
    int curS0 = -1, curS1 = -1, curBlock = 0, curLeft = 0;
    bool streamEnded = false;

    auto restart = [&]() {
        __syncwarp();
        if (haveWork) {  // the block is complete: fold its maximum into the pair's scores
            uint32_t r = mx;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r = __vmaxs2(r, __shfl_xor_sync(0xffffffffu, r, o));
            if (lane == 0) {
                const int lo = (int)(short)(r & 0xffff), hi = (int)(short)(r >> 16);
                if (sub0 >= 0) {
                    const int best = max(atomicMax(prm.scores + sub0, lo), lo);
                    if (isLast) {  // blocks of a pair finish in stream order: this is the pair's final score
                        if (best >= prm.statThreshold && prm.lengths[sub0] <= kStatMaxLength) atomicAdd(prm.statCount, 1);
                        if (best >= prm.ovfThreshold) prm.ovfList[atomicAdd(prm.ovfCount, 1)] = sub0;
                    }
                }
                if (sub1 >= 0) {
                    const int best = max(atomicMax(prm.scores + sub1, hi), hi);
                    if (isLast) {
                        if (best >= prm.statThreshold && prm.lengths[sub1] <= kStatMaxLength) atomicAdd(prm.statCount, 1);
                        if (best >= prm.ovfThreshold) prm.ovfList[atomicAdd(prm.ovfCount, 1)] = sub1;
                    }
                }
            }
        }
        if (w == 0) {  // the first warp of the array deals out this period's W blocks
            for (int i = 0; i < W; i++) {
                int4 dd = make_int4(0, -1, -1, 0);
                bool first = false;
                if (!streamEnded && curLeft == 0) {
                    int item = 0;
                    if (lane == 0) item = atomicAdd(prm.ticket, 1);
                    item = __shfl_sync(0xffffffffu, item, 0);
                    if (item >= prm.numItems) {
                        streamEnded = true;
                    } else {
                        const S16Item it = prm.items[item];
                        curS0 = it.subject0;
                        curS1 = it.subject1;
                        int len = 1;
                        if (curS0 >= 0) len = max(len, prm.lengths[curS0]);
                        if (curS1 >= 0) len = max(len, prm.lengths[curS1]);
                        curLeft = min((len + kLongBlockCols - 1) / kLongBlockCols, it.numSegments * 2);
                        curBlock = it.firstBlock * 2;
                        first = true;
                    }
                }
                if (!streamEnded) {
                    dd = make_int4(curBlock, curS0, curS1, 1 | (first ? 2 : 0) | (curLeft == 1 ? 4 : 0));
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
            sub0 = d.y;
            sub1 = d.z;
            useBorder = (d.w & 2) == 0;
            isLast = (d.w & 4) != 0;
            const uint4* src = reinterpret_cast<const uint4*>(prm.cols + (size_t)d.x * kLongBlockCols + lane * R);
            const uint4 c0 = __ldg(src), c1 = __ldg(src + 1);
            const uint32_t cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
            for (int b = 0; b < R / 2; b++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t f = h ? (cw[b] >> 16) : (cw[b] & 0xffffu);
                    const uint32_t s1 = (f * 3121u) >> 16;  // f / 21 for f < 441
                    const uint32_t s0 = f - 21u * s1;
                    a0[b * 2 + h] = loBase + (s0 * rowWords + xs) * 4;
                    a1[b * 2 + h] = hiBase + (s1 * rowWords + xs) * 4;
                }
            }
#pragma unroll
            for (int j = 0; j < R; j++) { Hp[j] = 0; F[j] = NEG2; }
            HinPrev = 0;
            mx = 0;
        }
        __syncwarp();
    };

    int sfill = kLongBatch % S, pfill = kLongBatch % P;  // next refill: ring slots / profile rows (= warp 0's next rows)
#pragma unroll 1
    for (int batch = 0;; ++batch) {
        cp_async_wait_all();
        if (!__syncthreads_or(alive)) break;
        long_ring_fill(loBase, hiBase, rowWords, S, prm, sfill, pfill);
        // warp 0's left border for the next batch: rows [pfill, pfill + 16) written by the last warp one period ago
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
                for (int j = 0; j < R; j++) { a0[j] += delta; a1[j] += delta; }
            }
        }
        static_for<2>([&](auto halfIndex) {
            constexpr int half = decltype(halfIndex)::value;
            if (p == pRestart && alive) restart();  // warp-uniform
            if (haveWork) {
                static_for<8>([&](auto stepIndex) {
                    constexpr int i = half * 8 + decltype(stepIndex)::value;
                    uint32_t Hin = __shfl_up_sync(0xffffffffu, Hp[R - 1], 1);
                    uint32_t Ein = __shfl_up_sync(0xffffffffu, Elast, 1);
                    const bool realRow = (unsigned)p < (unsigned)prm.qlen;
                    uint2 bd = make_uint2(0u, 0u);  // only the first lane's real rows read the FIFO (the slot is then final)
                    if (lane == 0 && realRow) bd = lds_u64(fifoIn + (p & (kLongFifoRows - 1)) * 8);
                    if (lane == 0) { Hin = useBorder ? bd.x : 0u; Ein = useBorder ? bd.y : NEG2; }
                    if (!realRow) { Hin = 0; Ein = NEG2; }
                    uint32_t E = Ein;
#ifdef SW4_LONG_PREFETCH
                    constexpr int kPrefetch = SW4_LONG_PREFETCH;
#else
                    constexpr int kPrefetch = 6;
#endif
                    uint32_t q0[kPrefetch + 1], q1[kPrefetch + 1];
#pragma unroll
                    for (int c = 0; c <= kPrefetch && c < R; c++) {
                        q0[c] = lds_u32_imm<i * 4>(a0[c]);
                        q1[c] = lds_u32_imm<i * 4>(a1[c]);
                    }
                    uint32_t d = __vadd2(__vadd2(HinPrev, q0[0]), q1[0]);
                    uint32_t dPrev = 0;
#pragma unroll
                    for (int j = 0; j < R; j++) {
                        const uint32_t n0 = q0[(j + 1) % (kPrefetch + 1)], n1 = q1[(j + 1) % (kPrefetch + 1)];
                        if (j + 1 + kPrefetch < R) {
                            q0[j % (kPrefetch + 1)] = lds_u32_imm<i * 4>(a0[j + 1 + kPrefetch]);
                            q1[j % (kPrefetch + 1)] = lds_u32_imm<i * 4>(a1[j + 1 + kPrefetch]);
                        }
                        uint32_t dNext = 0;
                        if (j + 1 < R) dNext = __vadd2(__vadd2(Hp[j], n0), n1);
                        const uint32_t h = __vimax3_s16x2_relu(d, E, F[j]);
                        Hp[j] = h;
                        const uint32_t tt = __vadd2(h, gop2);
                        E = __viaddmax_s16x2(E, gex2, tt);
                        F[j] = __viaddmax_s16x2(F[j], gex2, tt);
                        if (j & 1) mx = __vimax3_s16x2(mx, d, dPrev);
                        dPrev = d;
                        d = dNext;
                    }
                    Elast = E;
                    HinPrev = Hin;
                    if (lane == 31 && realRow) {  // right border of the block: row p is complete
                        if (w + 1 < W) sts_u64(fifoOut + (p & (kLongFifoRows - 1)) * 8, Hp[R - 1], Elast);
                        else border[p] = make_uint2(Hp[R - 1], Elast);
                    }
                    if (++p == P) p = 0;
                });
            } else {
                p += 8;
                if (p >= P) p -= P;
            }
        });
    }
    if (threadIdx.x == 0) {
        unsigned long long tEnd;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tEnd));
        atomicMax(prm.elapsedNs, tEnd - tStart);
        atomicMax(prm.elapsedNs + 32, ~tStart);
        atomicMax(prm.elapsedNs + 64, tEnd);
    }
}

}  // namespace sw4
