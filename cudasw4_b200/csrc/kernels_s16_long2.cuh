// CTA-wide systolic arrays for FEW very long subjects, two query rows per step (round 2).
//
// Same idea as kernels_s16_long.cuh (which holds the design notes): the warps of an array work on ONE stream of
// 512-column blocks (item 0 block 0, 1, 2, ..., item 1 block 0, ...), block n on warp n mod W in period n div W, so that
// a pair of B blocks is finished after ceil(B / W) periods instead of B - a database of a few long sequences is
// latency-bound, not throughput-bound. What changes is the step: every lane advances TWO query rows per step, as in
// kernels_s16.cuh. One 64-bit shared load per plane serves both rows of a column (1 load per cell-pair instead of 2),
// the hand-over shuffles, FIFO accesses and per-step bookkeeping are amortised over twice the cells, and the two rows'
// dependency chains interleave (the one-row array kernel spends a quarter of its issue slots waiting on its single
// chain and a fifth at the CTA barrier, profiles/prof_s16_long_r2.summary.txt).
//
// Geometry: lane l trails lane l-1 by one step (two rows), warp w trails warp w-1 by kLag2 = 40 steps (32 lanes + one
// 8-step batch: the border column of a block reaches the next warp through a 16-entry FIFO in shared memory, and the
// only synchronisation is the CTA barrier the ring refill takes every 8 steps anyway). The ring of profile rows has to
// span the whole array, 80 W + 16 rows, which caps W at 8; a CTA therefore runs 16 / W independent arrays side by side
// on the same ring (they are in phase: the same query rows at the same time), each with its own block stream, FIFOs and
// border rows. Short queries simply get more, shorter arrays (W = 2 at q = 144) instead of a 64-thread CTA.
#pragma once
#include "kernels_s16_long.cuh"

namespace sw4 {

constexpr int kLag2 = 40;            // steps between consecutive warps of an array
constexpr int kFifo2 = 32;           // entries (steps) per warp FIFO: an entry is consumed 9 steps after it was produced and its
                                     // slot rewritten 32 steps after that - always at least one CTA barrier later (with 16 entries
                                     // the rewrite could fall into the very batch of the read: racecheck found it)
constexpr int kLong2Warps = 16;      // warps per CTA (16 / W arrays of W warps)
constexpr int kLong2MaxW = 8;

struct S16Long2Params {
    const uint16_t* cols;        // the class's pair-blocks: consecutive fused column codes (s0 + 21*s1), 1024 per block
    const S16Item* items;        // [numItems], firstBlock / numSegments count 1024-column blocks
    const int32_t* lengths;      // [numLocalSubjects]
    int numItems;
    int* ticket;
    int warps;                   // W: 2, 4 or 8 warps per array
    int ringSlots;               // S (rows): multiple of 32, >= 80 W + 16
    const uint32_t* profLo;      // [21][profStride]  (M[q_p][s] & 0xffff), rows p >= qlen hold 0x0000c180
    const uint32_t* profHi;      // [21][profStride]  (M[q_p][s] << 16),   rows p >= qlen hold 0xc1800000
    int profStride;              // >= 2 * period
    int qlen;
    int period;                  // P in STEPS: multiple of 16, >= ceil(q/2) + 32 and >= 40 W + 16
    uint32_t gop2, gex2;
    int ovfThreshold, statThreshold;
    int32_t* scores;             // must hold -1 (or any value below every score) for the class's subjects at launch
    int32_t* ovfList;
    int* ovfCount;
    int* statCount;
    uint4* border;               // [gridDim.x * arrays][borderStride]: last warp -> first warp of the next period
    int borderStride;            // entries (steps) per array, >= period
};

static inline int s16_long2_ring_slots(int warps) { return (2 * kLag2 * warps + 16 + 31) / 32 * 32; }
static inline int s16_long2_smem_bytes(int warps) {
    return 2 * 21 * (s16_long2_ring_slots(warps) + 32) * 4 + kLong2Warps * kFifo2 * 16 + 2 * kLong2Warps * 16;
}
// W for a query: the largest power of two <= 8 whose array fits the period; 0 = query too short for an array
static inline int s16_long2_warps(int qlen, int* periodOut) {
    const int p0 = ((qlen + 1) / 2 + 32 + 15) / 16 * 16;  // (entries of the last 32 steps of a period are gap steps: never read)
    int w = kLong2MaxW;
    while (w > 1 && kLag2 * w + 16 > p0) w >>= 1;
    if (periodOut) *periodOut = p0;
    return w >= 2 ? w : 0;
}

// ring slots [slot0, slot0 + 16) <- profile rows [p0, p0 + 16) of both planes (slot0, p0 multiples of 16)
__device__ __forceinline__ void long2_ring_fill(uint32_t loBase, uint32_t hiBase, int rowWords, int S, const S16Long2Params& prm,
                                                int slot0, int p0) {
    for (int id = threadIdx.x; id < 2 * 21 * 4; id += blockDim.x) {
        const int c = id & 3, s = (id >> 2) % 21, plane = id / 84;
        const uint32_t* src = (plane ? prm.profHi : prm.profLo) + (size_t)s * prm.profStride + p0 + 4 * c;
        const uint32_t dst = (plane ? hiBase : loBase) + (s * rowWords + slot0 + 4 * c) * 4;
        cp_async16(dst, src);
        if (slot0 == 0) cp_async16(dst + S * 4, src);  // mirror of slots [0, 16) behind the ring's end
    }
}

template <int GAPS = 0>
__global__ void __launch_bounds__(kLong2Warps * 32, 1) sw_s16_long2_kernel(const S16Long2Params prm) {
    constexpr int R = kLongR;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, w16 = threadIdx.x >> 5;
    const int W = prm.warps, S = prm.ringSlots, P = prm.period;
    const int arr = w16 / W, w = w16 - arr * W;   // array inside the CTA, warp inside the array
    const int rowWords = S + 32;
    const uint32_t loBase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t hiBase = loBase + 21 * rowWords * 4;
    const uint32_t fifoBase = hiBase + 21 * rowWords * 4;
    volatile int4* desc = reinterpret_cast<volatile int4*>(smem + 2 * 21 * rowWords * 4 + kLong2Warps * kFifo2 * 16) + arr * 2 * W;
    const uint32_t fifoIn = fifoBase + w16 * (kFifo2 * 16);
    const uint32_t fifoOut = fifoIn + kFifo2 * 16;  // the next warp's input (unused by the last warp of an array)
    uint4* border = prm.border + (size_t)(blockIdx.x * (kLong2Warps / W) + arr) * prm.borderStride;
    const uint32_t NEG2 = ((uint32_t)(uint16_t)kNegS16 << 16) | (uint16_t)kNegS16;
    const uint32_t gop2 = GAPS > 0 ? s16_gap_set(GAPS).x : prm.gop2, gex2 = GAPS > 0 ? s16_gap_set(GAPS).y : prm.gex2;
    const int qSteps = (prm.qlen + 1) >> 1;

    // rows "before time 0" and between two periods are gap rows: the ring starts out as -16000 everywhere
    for (int i = threadIdx.x; i < 21 * rowWords; i += blockDim.x) {
        reinterpret_cast<uint32_t*>(smem)[i] = NEG2 & 0xffffu;
        reinterpret_cast<uint32_t*>(smem)[21 * rowWords + i] = NEG2 & 0xffff0000u;
    }
    __syncthreads();
    long2_ring_fill(loBase, hiBase, rowWords, S, prm, 0, 0);
    cp_async_commit();

    uint32_t a0[R], a1[R];  // ring byte addresses of this column's two profile rows (lane phase folded in)
    uint32_t Hp[R], F[R];
    uint32_t mx = 0, HlastA = 0, ElastA = NEG2, ElastB = NEG2, HinPrevB = 0;
#pragma unroll
    for (int j = 0; j < R; j++) { a0[j] = loBase; a1[j] = hiBase; Hp[j] = 0; F[j] = NEG2; }
    // this lane's step in the period-P schedule: (t - 40 w - lane) mod P; its ring slot (a row index) at the batch start
    const int skew = kLag2 * w + lane;
    int p = skew == 0 ? 0 : P - skew;
    int xs = skew == 0 ? 0 : S - 2 * skew;
    const int pRestart = lane == 0 ? 0 : P - lane;
    bool alive = true, haveWork = false, useBorder = false, isLast = false;
    int sub0 = -1, sub1 = -1, periodIndex = 0;
    // first warp of an array only: where the array's block stream stands
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
                    volatile int4* slot = desc + (periodIndex & 1) * W + i;
                    slot->x = dd.x; slot->y = dd.y; slot->z = dd.z; slot->w = dd.w;
                }
            }
            __syncwarp();
        }
        int4 d;
        {
            volatile int4* slot = desc + (periodIndex & 1) * W + w;
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
            // the hand-over registers restart from the boundary values too (see kernels_s16.cuh): lanes > 0 then need no
            // "is this a real row" test, their gap rows only ever see values of this very block
            HinPrevB = 0; HlastA = 0; ElastA = NEG2; ElastB = NEG2;
            mx = 0;
        }
        __syncwarp();
    };

    // the descriptor of a block is written by its array's first warp one restart (40 w steps) before warp w reads it, with
    // at least one CTA barrier in between (w >= 1: 40 steps = 5 batches); warp 0 reads its own after __syncwarp
    int sfill = 16 % S, pfill = 16 % (2 * P);  // next refill: ring slots / profile rows
#pragma unroll 1
    for (int batch = 0;; ++batch) {
        cp_async_wait_all();
        if (!__syncthreads_or(alive)) break;
        long2_ring_fill(loBase, hiBase, rowWords, S, prm, sfill, pfill);
        // first warp's left border for the NEXT batch: entries written by the array's last warp one period ago
        if (w == 0 && lane < 8) {
            int e = p + 8;   // lane 0's step at the next batch start (lane 0 of warp 0: p is a multiple of 8)
            e = __shfl_sync(0xffu, e, 0);
            if (e >= P) e -= P;
            if (e + lane < qSteps) cp_async16(fifoIn + ((e + lane) & (kFifo2 - 1)) * 16, border + e + lane);
        }
        cp_async_commit();
        sfill += 16;
        if (sfill >= S) sfill -= S;
        pfill += 16;
        if (pfill >= 2 * P) pfill -= 2 * P;
        if (batch > 0) {
            int delta = 16 * 4;
            xs += 16;
            if (xs >= S) { xs -= S; delta = (16 - S) * 4; }
            if (haveWork) {
#pragma unroll
                for (int j = 0; j < R; j++) { a0[j] += delta; a1[j] += delta; }
            }
        }
        if (p == pRestart && alive) restart();  // warp-uniform
        if (haveWork) {
            // hand-over addresses of this batch, all in terms of the first lane's step p0 (a multiple of 8): lane 0 reads
            // FIFO entry (p0 + i) & 31, lane 31 - 31 steps behind - writes entry (p0 - 31 + i) & 31; only its last step of a
            // batch can cross a multiple of 32 or the end of the period
            const int p0 = __shfl_sync(0xffffffffu, p, 0);
            const uint32_t inBase = fifoIn + (p0 & (kFifo2 - 8)) * 16;
            const int nReal = qSteps - p0;  // steps of this batch (counted from 0) that are real rows for the first lane
            int p31 = p0 - 31;
            if (p31 < 0) p31 += P;
            int e7 = p31 + 7;
            if (e7 >= P) e7 -= P;
            const uint32_t outA = fifoOut + (p31 & (kFifo2 - 1)) * 16, out7 = fifoOut + (e7 & (kFifo2 - 1)) * 16;
            uint4* const bA = border + p31;
            uint4* const b7 = border + e7;
            const bool first = lane == 0, last = lane == 31, lastWarp = w + 1 == W;
            const bool take = first && useBorder;
            static_for<8>([&](auto stepIndex) {
                constexpr int i = decltype(stepIndex)::value;
                uint32_t HinA = __shfl_up_sync(0xffffffffu, HlastA, 1);
                uint32_t EinA = __shfl_up_sync(0xffffffffu, ElastA, 1);
                uint32_t HinB = __shfl_up_sync(0xffffffffu, Hp[R - 1], 1);
                uint32_t EinB = __shfl_up_sync(0xffffffffu, ElastB, 1);
                if (first) {
                    // the block's left border for lane 0: the FIFO entry of this step - or the boundary column for an item's
                    // first block (the FIFO then holds another item's values) and for gap steps (their entries are not
                    // bounded by this block's scores, and their slots may be in the middle of being rewritten)
                    HinA = 0u; EinA = NEG2; HinB = 0u; EinB = NEG2;
                    if (take && i < nReal) {
                        const uint4 bv = lds_u128_imm<i * 16>(inBase);
                        HinA = bv.x; EinA = bv.y; HinB = bv.z; EinB = bv.w;
                    }
                }
                {
                    uint32_t E1 = EinA, E2 = EinB;
#ifdef SW4_LONG2_PREFETCH
                    constexpr int kPrefetch = SW4_LONG2_PREFETCH;
#else
                    constexpr int kPrefetch = 4;
#endif
                    uint2 q0[kPrefetch + 1], q1[kPrefetch + 1];
#pragma unroll
                    for (int c = 0; c <= kPrefetch && c < R; c++) {
                        q0[c] = lds_u64_imm<i * 8>(a0[c]);
                        q1[c] = lds_u64_imm<i * 8>(a1[c]);
                    }
                    uint32_t da = __vadd2(__vadd2(HinPrevB, q0[0].x), q1[0].x);  // row a: diagonal = row b of the previous step, previous lane
                    uint32_t db = __vadd2(__vadd2(HinA, q0[0].y), q1[0].y);      // row b: diagonal = row a of this step, previous lane
                    uint32_t ha = 0;
#pragma unroll
                    for (int j = 0; j < R; j++) {
                        const uint2 n0 = q0[(j + 1) % (kPrefetch + 1)], n1 = q1[(j + 1) % (kPrefetch + 1)];
                        if (j + 1 + kPrefetch < R) {
                            q0[j % (kPrefetch + 1)] = lds_u64_imm<i * 8>(a0[j + 1 + kPrefetch]);
                            q1[j % (kPrefetch + 1)] = lds_u64_imm<i * 8>(a1[j + 1 + kPrefetch]);
                        }
                        uint32_t na = 0, nb = 0;
                        if (j + 1 < R) na = __vadd2(__vadd2(Hp[j], n0.x), n1.x);
                        ha = __vimax3_s16x2_relu(da, E1, F[j]);
                        const uint32_t ta = __vadd2(ha, gop2);
                        E1 = __viaddmax_s16x2(E1, gex2, ta);
                        const uint32_t Fa = __viaddmax_s16x2(F[j], gex2, ta);
                        if (j + 1 < R) nb = __vadd2(__vadd2(ha, n0.y), n1.y);
                        const uint32_t hb = __vimax3_s16x2_relu(db, E2, Fa);
                        Hp[j] = hb;
                        const uint32_t tb = __vadd2(hb, gop2);
                        E2 = __viaddmax_s16x2(E2, gex2, tb);
                        F[j] = __viaddmax_s16x2(Fa, gex2, tb);
                        mx = __vimax3_s16x2(mx, da, db);
                        da = na;
                        db = nb;
                    }
                    HlastA = ha;
                    ElastA = E1;
                    ElastB = E2;
                    HinPrevB = HinB;
                }
                if (last) {  // right border of the block for this step's two rows (gap steps included, see above)
                    if (lastWarp) *(i < 7 ? bA + i : b7) = make_uint4(HlastA, ElastA, Hp[R - 1], ElastB);
                    else sts_u128(i < 7 ? outA + i * 16 : out7, HlastA, ElastA, Hp[R - 1], ElastB);
                }
            });
        }
        p += 8;
        if (p >= P) p -= P;
    }
}

}  // namespace sw4
