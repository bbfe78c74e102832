// Full-warp variant of the packed s16x2 Gotoh kernel: ONE query row per step, 32-lane pipelines (G = 32), used for
// subjects of 513..1024 residues (single segment) and for the multi-segment class (1024-column segments, subjects of
// any length). Same ring profile, staging, ticketing and restart logic as kernels_s16.cuh - which holds the design
// notes - but lane l works on row t-l only, so a 32-lane group fits the 64-row ring window (the two-rows-per-step
// kernel is limited to 16-lane groups). F never crosses a segment; the border column (H, E) per query row goes through
// a per-warp array in global memory, 32 rows at a time, coalesced.
// Replaces the reference's multi-pass kernels (src/dpx_s16_kernels.cuh:290-762, 874-965, 1066-1213).
#pragma once
#include "kernels_s16.cuh"

namespace sw4 {

constexpr int kWideFillBatch = 16;          // steps between ring refills / CTA barriers (unrolled)

struct S16WideParams {
    const uint16_t* cols;        // [numBlocks][G*R] fused column codes, lane-major (lane m owns [m*R, m*R+R))
    const S16Item* items;        // [numItems] in the order they should be started
    int numItems;
    int* ticket;                 // zero-initialised work counter (items are handed out dynamically)
    int logG;                    // G = 1 << logG lanes per group
    const uint32_t* profile;     // [441][profStride] positional query profile
    int profStride;
    int qlen;
    int period;                  // P: steps between two alignments of a group; multiple of 8, >= max(32, qlen + G - 1); G >= 8
    uint32_t gop2, gex2;         // gap scores replicated in both halves
    int ovfThreshold;            // running maximum >= this => exact 32-bit re-scoring (25000, reference MAX_ACC_SHORT)
    int statThreshold;           // running maximum >= this => counted in stats.num_overflows (25000, or 2048 for Half2)
    int32_t* scores;             // [numLocalSubjects]
    int32_t* ovfList;            // local subject indices that need the exact 32-bit path
    int* ovfCount;
    int* statCount;
    unsigned long long* elapsedNs;  // max over CTAs of this launch's run time (feedback for the host's SM partition)
    int activeGroups;            // groups per CTA that take work (fewer than all when the class cannot fill its SMs: the
                                 // items are then spread over more SMs and every warp gets a larger share of its scheduler)
    int ctaOffset;               // index of this launch's first CTA within the class (a class may be split in two launches)
    const int32_t* lengths;      // MULTI only: [numLocalSubjects] (subjects longer than kStatMaxLength are not counted)
    uint2* border;               // MULTI only: [gridWarps][borderStride] (H, E) of a segment's last column per query row
    int borderStride;
};

// software-pipeline depth of the substitution loads per instantiation (see s16_prefetch_depth)
template <int R>
__host__ __device__ constexpr int s16_wide_prefetch_depth() {
#ifdef SW4_WIDE_PREFETCH
    return SW4_WIDE_PREFETCH;
#else
    return R == 24 ? 12 : R == 28 ? 14 : 8;  // measured (tools/class_sweep.py): R = 28: 6.49 vs 6.30 TCUPS, R = 24: 6.60 vs 6.53
#endif
}

template <int R>
constexpr int s16_wide_smem_bytes() { return kRingBytes + kS16Warps * 32 * R * 2 + kS16Warps * 8 * kGroupStateInts * 4; }

// Refill ring slots for global steps [x0, x0+16) (x0 % 16 == 0); p0 = x0 mod period.
__device__ __forceinline__ void ring_fill_wide(uint32_t ringBase, const uint32_t* __restrict__ profile, int profStride, int x0,
                                          int p0, int period) {
    const int slot0 = x0 & (kRingSlots - 1);
    for (int id = threadIdx.x; id < kFused * (kWideFillBatch / 4); id += kS16Threads) {
        const int f = id >> 2, c = id & 3;
        int p = p0 + 4 * c;
        if (p >= period) p -= period;
        if (p >= period) p -= period;
        const int slot = slot0 + 4 * c;
        const uint32_t* src = profile + (size_t)f * profStride + p;
        const uint32_t dst = ringBase + (f * kRingStride + 32 + slot) * 4;
        cp_async16(dst, src);
        if (slot >= 32) cp_async16(dst - kRingSlots * 4, src);
    }
    cp_async_commit();
}

// R = register columns per lane. MULTI = the long class: G must be 32 and an item may span several segments.
template <int R, bool MULTI>
__global__ void __launch_bounds__(kS16Threads, 1) sw_s16_wide_kernel(const S16WideParams prm) {
    static_assert(R % 2 == 0, "a lane's staging slice (R u16 codes) must be a whole number of 4-byte words");
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long tStart;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tStart));
    const uint32_t ringBase = (uint32_t)__cvta_generic_to_shared(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int logG = MULTI ? 5 : prm.logG, G = 1 << logG;
    const int g = lane >> logG, m = lane & (G - 1);
    const int P = prm.period;
    const unsigned groupMask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (g << logG));
    const int leader = g << logG;  // lane index of the group's first lane

    // this lane's slice of the group's staging area: R fused u16 column codes of the next pair-block
    const uint32_t stageLane = ringBase + kRingBytes + (warp * 32 + lane) * (R * 2);
    // group state in shared memory: [0] subject0 [1] subject1 [2] segments left after the current one
    // [3] look-ahead valid [4] look-ahead block [5] look-ahead starts a new item [6] look-ahead item index
    volatile int* gs = reinterpret_cast<volatile int*>(smem + kRingBytes + kS16Warps * 32 * R * 2) +
                       (warp * 8 + g) * kGroupStateInts;
    uint2* border = MULTI ? prm.border + (size_t)((blockIdx.x + prm.ctaOffset) * kS16Warps + warp) * prm.borderStride : nullptr;

    // Warps that will never get work (the class was spread over more SMs than it can fill, prm.activeGroups) only keep
    // the CTA's ring going: same barriers and their share of every refill, none of the arithmetic, so the busy warps get
    // their scheduler to themselves.
    if (warp * (32 >> logG) >= prm.activeGroups) {
        const uint32_t NEG2i = ((uint32_t)(uint16_t)kNegS16 << 16) | (uint16_t)kNegS16;
        for (int i = threadIdx.x; i < kFused * kRingStride; i += kS16Threads) reinterpret_cast<uint32_t*>(smem)[i] = NEG2i;
        __syncthreads();
        ring_fill_wide(ringBase, prm.profile, prm.profStride, 0, 0, P);
        int pf = kWideFillBatch % P;
#pragma unroll 1
        for (int batch = 0;; ++batch) {
            cp_async_wait_all();
            if (!__syncthreads_or(false)) break;
            ring_fill_wide(ringBase, prm.profile, prm.profStride, (batch + 1) * kWideFillBatch, pf, P);
            pf += kWideFillBatch;
            while (pf >= P) pf -= P;
        }
        return;
    }

    uint32_t colAddr[R];  // ring byte address of this column's fused-pair row (lane and batch-phase offsets folded in)
    uint32_t Hp[R];       // H of the previous row
    uint32_t F[R];        // F for the next row
    const uint32_t NEG2 = ((uint32_t)(uint16_t)kNegS16 << 16) | (uint16_t)kNegS16;
    uint32_t mx = 0, Elast = NEG2, HinPrev = 0;
#pragma unroll
    for (int j = 0; j < R; j++) { colAddr[j] = ringBase; Hp[j] = 0; F[j] = NEG2; }
    // p = this lane's row in the period-P schedule, (t - lane) mod P. The group restarts (stores the finished pair,
    // loads the next one) when its first lane is at row 0, i.e. when this lane is at row pRestart.
    int p = (lane == 0) ? 0 : P - lane;
    const int pRestart = (m == 0) ? 0 : P - m;
    bool haveWork = false;    // a segment is being computed
    bool useBorder = false;   // MULTI: the current segment continues an item (left border comes from `border`)
    bool alive = (warp * (32 >> logG) + g) < prm.activeGroups;  // the group still has something to compute, finalise or start
    uint2 inBuf = make_uint2(0, NEG2), outBuf = make_uint2(0, 0);

    // look-ahead: fetch the descriptor of the item / segment that follows and start copying its columns
    auto fetch_lookahead = [&](bool continuing, int curBlock) {
        int blk = -1, isNew = 0, item = -1;
        if (!alive) {
            item = prm.numItems;  // this group never takes work
        } else if (continuing) {
            blk = curBlock + 1;
        } else {
            if (m == 0) item = atomicAdd(prm.ticket, 1);
            item = __shfl_sync(groupMask, item, leader);
            if (item < prm.numItems) { blk = prm.items[item].firstBlock; isNew = 1; }
        }
        if (m == 0) { gs[3] = blk >= 0; gs[4] = blk; gs[5] = isNew; gs[6] = item; }
        if (blk >= 0) {
            const unsigned char* src = (const unsigned char*)(prm.cols + (size_t)blk * (G * R)) + m * (R * 2);
            if constexpr ((R * 2) % 16 == 0) {
#pragma unroll
                for (int i = 0; i < R * 2 / 16; i++) cp_async16(stageLane + i * 16, src + i * 16);
            } else if constexpr ((R * 2) % 8 == 0) {
#pragma unroll
                for (int i = 0; i < R * 2 / 8; i++) cp_async8(stageLane + i * 8, src + i * 8);
            } else {
#pragma unroll
                for (int i = 0; i < R * 2 / 4; i++) cp_async4(stageLane + i * 4, src + i * 4);
            }
        }
        cp_async_commit();
        __syncwarp(groupMask);
    };

    // prologue: lanes l > 0 run their first l steps at "negative time" (rows before the query starts): those ring
    // slots must read as gap rows too, so the whole ring starts out as -16000; then the first batch + first pair-blocks.
    for (int i = threadIdx.x; i < kFused * kRingStride; i += kS16Threads) reinterpret_cast<uint32_t*>(smem)[i] = NEG2;
    __syncthreads();
    ring_fill_wide(ringBase, prm.profile, prm.profStride, 0, 0, P);
    fetch_lookahead(false, 0);
    int pfill = kWideFillBatch % P;  // (next fill start) mod P

    // The step offset inside the ring is an instruction immediate: 16 steps are unrolled and the column addresses are
    // advanced by 64 bytes once per batch (ptxas does not fold a uniform register into LDS addresses, and an
    // address add per cell would cost an issue slot per cell-pair).
    uint32_t phaseBase = ringBase + (32 - lane) * 4;  // + 64 bytes per batch, wrapping every 4 batches
#pragma unroll 1
    for (int batch = 0;; ++batch) {
        cp_async_wait_all();
        if (!__syncthreads_or(alive)) break;
        ring_fill_wide(ringBase, prm.profile, prm.profStride, (batch + 1) * kWideFillBatch, pfill, P);
        pfill += kWideFillBatch;
        while (pfill >= P) pfill -= P;
        if (batch > 0) {
            const int delta = (batch & 3) ? kWideFillBatch * 4 : -(kRingSlots - kWideFillBatch) * 4;
            phaseBase += delta;
#pragma unroll
            for (int j = 0; j < R; j++) colAddr[j] += delta;
        }
        static_for<kWideFillBatch>([&](auto stepIndex) {
            constexpr int i = decltype(stepIndex)::value;
            if ((i & 7) == 0 && p == pRestart && alive) {  // group restart: uniform in the group, divergent across groups
                __syncwarp(groupMask);
                int segsLeft = gs[2];
                if (haveWork && segsLeft == 0) {  // the item is complete: reduce the maxima and store the two scores
                    uint32_t r = mx;
                    for (int o = G >> 1; o > 0; o >>= 1) r = __vmaxs2(r, __shfl_xor_sync(groupMask, r, o));
                    if (m == 0) {
                        const int s0 = gs[0], s1 = gs[1];
                        const int lo = (int)(short)(r & 0xffff), hi = (int)(short)(r >> 16);
                        if (s0 >= 0) {
                            if (lo >= prm.statThreshold && (!MULTI || prm.lengths[s0] <= kStatMaxLength)) atomicAdd(prm.statCount, 1);
                            if (lo >= prm.ovfThreshold) prm.ovfList[atomicAdd(prm.ovfCount, 1)] = s0;
                            prm.scores[s0] = lo;
                        }
                        if (s1 >= 0) {
                            if (hi >= prm.statThreshold && (!MULTI || prm.lengths[s1] <= kStatMaxLength)) atomicAdd(prm.statCount, 1);
                            if (hi >= prm.ovfThreshold) prm.ovfList[atomicAdd(prm.ovfCount, 1)] = s1;
                            prm.scores[s1] = hi;
                        }
                    }
                }
                const bool laValid = gs[3] != 0;
                const int laBlk = gs[4];
                const bool laNew = gs[5] != 0;
                haveWork = laValid;
                alive = laValid;
                if (laValid) {
                    if (laNew) {
                        const S16Item it = prm.items[gs[6]];
                        __syncwarp(groupMask);
                        if (m == 0) { gs[0] = it.subject0; gs[1] = it.subject1; gs[2] = it.numSegments - 1; }
                        segsLeft = it.numSegments - 1;
                        mx = 0;
                        useBorder = false;
                    } else {
                        __syncwarp(groupMask);
                        if (m == 0) gs[2] = segsLeft - 1;
                        segsLeft -= 1;
                        useBorder = true;
                    }
                    cp_async_wait_all();
                    __syncwarp(groupMask);
#pragma unroll
                    for (int b = 0; b < R / 2; b++) {
                        const uint32_t w = lds_u32_imm<0>(stageLane + b * 4);
                        colAddr[b * 2 + 0] = phaseBase + (w & 0xffffu) * (kRingStride * 4);
                        colAddr[b * 2 + 1] = phaseBase + (w >> 16) * (kRingStride * 4);
                    }
#pragma unroll
                    for (int j = 0; j < R; j++) { Hp[j] = 0; F[j] = NEG2; }
                    HinPrev = 0;
                    Elast = NEG2;  // hand-over registers start from the boundary values as well (see kernels_s16.cuh)
                    __syncwarp(groupMask);
                    fetch_lookahead(segsLeft > 0, laBlk);
                }
            }
            // systolic hand-over from the previous lane (row p was computed there one step earlier)
            uint32_t Hin = __shfl_up_sync(0xffffffffu, Hp[R - 1], 1);
            uint32_t Ein = __shfl_up_sync(0xffffffffu, Elast, 1);
            // Rows p >= q are "gap rows" between two alignments of the group: they are computed like any other row (no
            // branch => no register shuffling at a merge point) on profile entries of -16000, with the hand-over
            // inputs forced to the boundary values so that nothing leaks into the freshly reset state; they can
            // never raise the running maximum.
            if constexpr (MULTI) {
                const bool realRow = (unsigned)p < (unsigned)prm.qlen;
                // left border of a continued item: 32 rows at a time, coalesced; lane 0 is at row p
                const int p0 = __shfl_sync(0xffffffffu, p, 0);
                if ((p0 & 31) == 0 && p0 < prm.qlen) inBuf = useBorder ? border[p0 + lane] : make_uint2(0, NEG2);
                const uint32_t bH = __shfl_sync(0xffffffffu, inBuf.x, p0 & 31);
                const uint32_t bE = __shfl_sync(0xffffffffu, inBuf.y, p0 & 31);
                if (m == 0) { Hin = bH; Ein = bE; }
                if (!realRow) { Hin = 0; Ein = NEG2; }
            } else {
                if (m == 0) { Hin = 0; Ein = NEG2; }  // gap rows need no forcing: every lane's state is reset at the restart
            }
            {
                uint32_t E = Ein;
                // substitution words are fetched kPrefetch columns ahead of their use (explicit software pipeline, see
                // kernels_s16.cuh)
                constexpr int kPrefetch = s16_wide_prefetch_depth<R>();
                uint32_t sq[kPrefetch + 1];
#pragma unroll
                for (int c = 0; c <= kPrefetch && c < R; c++) sq[c] = lds_u32_imm<i * 4>(colAddr[c]);
                uint32_t d = __vadd2(HinPrev, sq[0]);
                uint32_t dPrev = 0;
#pragma unroll
                for (int j = 0; j < R; j++) {
                    // look-ahead: the next column's diagonal term reads Hp[j] before this column overwrites it
                    const uint32_t sn = sq[(j + 1) % (kPrefetch + 1)];
                    if (j + 1 + kPrefetch < R) sq[j % (kPrefetch + 1)] = lds_u32_imm<i * 4>(colAddr[j + 1 + kPrefetch]);
                    uint32_t dNext = 0;
                    if (j + 1 < R) dNext = __vadd2(Hp[j], sn);
                    const uint32_t h = __vimax3_s16x2_relu(d, E, F[j]);
                    Hp[j] = h;
                    const uint32_t tt = __vadd2(h, prm.gop2);
                    E = __viaddmax_s16x2(E, prm.gex2, tt);
                    F[j] = __viaddmax_s16x2(F[j], prm.gex2, tt);
                    // max over d == max over H: a best local alignment ends on a match, and d having a second use
                    // keeps ptxas from fusing the add into an ALU-pipe VIADDMNMX
                    if (j & 1) mx = __vimax3_s16x2(mx, d, dPrev);
                    dPrev = d;
                    d = dNext;
                }
                Elast = E;
                HinPrev = Hin;
            }
            if constexpr (MULTI) {
                // right border: lane 31 has just finished its row p31; collect 32 rows, then store them coalesced
                const int p31 = __shfl_sync(0xffffffffu, p, 31);
                const uint32_t vH = __shfl_sync(0xffffffffu, Hp[R - 1], 31);
                const uint32_t vE = __shfl_sync(0xffffffffu, Elast, 31);
                if (p31 < prm.qlen && haveWork) {
                    if (lane == (p31 & 31)) outBuf = make_uint2(vH, vE);
                    if ((p31 & 31) == 31 || p31 == prm.qlen - 1) {
                        if (lane <= (p31 & 31)) border[(p31 & ~31) + lane] = outBuf;
                    }
                }
            }
            if (++p == P) p = 0;
        });
    }
    if (threadIdx.x == 0) {
        unsigned long long tEnd;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tEnd));
        atomicMax(prm.elapsedNs, tEnd - tStart);
        atomicMax(prm.elapsedNs + 32, ~tStart);  // earliest CTA start (as a max of the complement)
        atomicMax(prm.elapsedNs + 64, tEnd);     // latest CTA end
    }
}

}  // namespace sw4
