// Inter-sequence Gotoh score kernel, packed s16x2 (two subjects per 32-bit lane register), subjects up to 512 residues.
//
// Replaces the reference's NW_local_affine_{single,multi}_pass_dpx_s16 family (src/dpx_s16_kernels.cuh:290-1357)
// with a different organisation, designed around what the B200 SM measures (tools/ubench/, profiles/ubench_*.txt):
// DPX min/max ops issue on the ALU pipe at 1 warp-inst / 2 clk / scheduler, VIADD.16x2 on the FMA-heavy pipe, shared
// memory serves one conflict-free 128-byte request per clock, and every instruction that is not one of the 3.5 DPX ops
// a cell-pair needs costs close to a full issue slot. So the loop must contain nothing but the recurrence:
//
//  * Half-warp = one 16-stage systolic pipeline. Lane l (0..15 inside its half-warp) works on query rows 2(t-l) and
//    2(t-l)+1 at step t: TWO rows per step, so a column's two substitution words come from ONE 64-bit shared load,
//    F[j]/Hp[j] are read and written once per two rows, the two rows' E chains interleave (ILP 2) and the per-step
//    bookkeeping (hand-over shuffles, counters) is amortised over 2*R cell-pairs. A half-warp is cut into 16/G groups
//    (G = 8 or 16 lanes); a group aligns one *pair-block* (two subjects, G*R columns, R register columns per lane) and
//    restarts on its own schedule every P steps (P = ceil(q/2)+G-1 rounded to 8): no warp-wide fill/drain.
//    Row pairs in [q, 2P) are *gap rows*: computed like any other row (no branch, hence no register shuffling at a merge
//    point) on profile entries of -16000. Whatever they leave in a lane's registers is reset at the group's restart.
//  * Substitution scores come from a *positional* query profile prof[f][row] = (M[q_row][s1] << 16 | M[q_row][s0]) for
//    the fused residue pair f = s0 + 21*s1. A 64-row sliding window lives in shared memory as ring[f][96] (64 slots +
//    32 mirrored), refilled 16 rows ahead with cp.async. Lane l reads rows 2(t-l), 2(t-l)+1: the 16 lanes of a
//    half-warp (one LDS.64 wavefront) always cover 32 different banks - no conflicts whatever the residues are - and
//    the address is (per-column register, fixed for the whole alignment) + (instruction immediate): the 8 steps of a
//    batch are unrolled and the column addresses advance by 64 bytes once per batch. No per-cell address arithmetic.
//  * The device database stores pair-blocks as fused u16 column codes (device_db.cuh), fetched one alignment ahead with
//    cp.async into a per-group staging area; work items are handed out through an atomic ticket.
//  * Subjects longer than 16 x R columns (MULTI instantiations): a group aligns the 16 x R-column segments of a pair one
//    after the other, one period each. The last column's (H, E) of both rows of every step goes to a per-group border
//    array in global memory (one 16-byte store per step by the group's last lane) and comes back as the left border of
//    the next segment through a small double-buffered staging area in shared memory (cp.async one batch ahead, one
//    128-bit shared load per step); F never crosses a segment (it runs along the query), the running maximum carries
//    over. (kernels_s16_wide.cuh holds the older full-warp one-row-per-step form of the same idea.)
//
// Arithmetic (bit-exact vs the oracle): H = max(0, diag+s, E, F); t = H+gop; E' = max(E+gex, t); F' = max(F+gex, t);
// packed modular s16 adds, -16000 as minus infinity; the running maximum is taken over diag+s (equal to max H: a best
// local alignment ends on a match; and the second use keeps ptxas from fusing the add into an ALU-pipe VIADDMNMX).
// A pair whose maximum reaches `ovfThreshold` is re-scored in 32 bit (kernels_s32.cuh): the reference's envelope
// argument.
#pragma once
#include <cstdint>
#include <utility>
#include <cuda_runtime.h>

namespace sw4 {

#ifndef SW4_S16_THREADS
#define SW4_S16_THREADS 512                // development switch (kernel-variant sweeps)
#endif
constexpr int kS16Threads = SW4_S16_THREADS;  // 16 warps, one CTA per SM
constexpr int kS16Warps = kS16Threads / 32;
constexpr int kRingSlots = 64;             // query rows held in the ring
constexpr int kRingStride = 96;            // words per fused-pair row: 32 mirrored + 64 live slots
constexpr int kFused = 441;                // 21 x 21 residue pairs
constexpr int kBatchSteps = 8;             // steps between ring refills / CTA barriers (unrolled)
constexpr int kBatchRows = 2 * kBatchSteps;
constexpr int kRingBytes = kFused * kRingStride * 4;
constexpr short kNegS16 = -16000;

// One unit of work for a group: a pair of subjects occupying `numSegments` consecutive pair-blocks (1 for every
// subject that fits the class's G*R columns; > 1 only in the multi-segment class, where segment s holds columns
// [s*G*R, (s+1)*G*R) and the last column's (H, E) of every query row is handed to the next segment through `border`).
struct S16Item {
    int subject0, subject1;   // local subject index of the low / high half (-1 = none)
    int firstBlock;           // index of the first pair-block in `cols`
    int numSegments;
};

struct S16Params {
    const uint16_t* cols;        // [numBlocks][G*R] fused column codes, lane-major (lane m owns [m*R, m*R+R))
    const S16Item* items;        // [numItems] in the order they should be started (read by the MULTI instantiations only)
    int numItems;
    int firstSubject, numSubjects;  // single-segment classes: item k = subjects (firstSubject + 2k, + 2k + 1), pair-block k -
                                 // derived on the spot instead of two dependent global loads at every restart
    int* ticket;                 // zero-initialised work counter (items are handed out dynamically)
    int logG;                    // G = 1 << logG lanes per group, 8 or 16
    const uint32_t* profile;     // [441][profStride] positional query profile, rows >= qlen hold -16000
    int profStride;
    int qlen;
    int period;                  // P: steps between two alignments of a group; multiple of 8, >= max(16, ceil(q/2) + G - 1)
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
    // MULTI instantiations only (subjects longer than one 16 x R segment):
    const int32_t* lengths;      // [numLocalSubjects] (statistics rule; segments that are padding only are skipped)
    uint4* border;               // [gridGroups][borderStride]: (H, E) of a segment's last column for the two rows of every step
    int borderStride;            // entries per group, >= period + 16
    int* borderSlots;            // [numBorderSlots] zero-initialised flags: a CTA owns the border arrays of the slot it holds
    int numBorderSlots;          //   (several classes run concurrently on one border buffer; one CTA fits per SM)
    int blockScale;              // the class stores blocks of blockScale * 16 * R columns (2 for the 32-lane layouts)
};

template <int N, class Fn, int... Is>
__device__ __forceinline__ void static_for_impl(Fn&& fn, std::integer_sequence<int, Is...>) {
    (fn(std::integral_constant<int, Is>{}), ...);
}
// compile-time unrolled loop: fn(std::integral_constant<int, i>) for i in [0, N)
template <int N, class Fn>
__device__ __forceinline__ void static_for(Fn&& fn) {
    static_for_impl<N>(fn, std::make_integer_sequence<int, N>{});
}

template <int IMM>
__device__ __forceinline__ uint32_t lds_u32_imm(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
    return v;
}
template <int IMM>
__device__ __forceinline__ uint2 lds_u64_imm(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(addr), "n"(IMM));
    return v;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// the reference counts "overflows" only in its partition 34 (240 < len <= 8000, src/cudasw4.cuh:2152-2169): longer
// subjects go straight to its 32-bit kernel
constexpr int kStatMaxLength = 8000;
constexpr int kGroupStateInts = 8;  // per-group bookkeeping kept in shared memory (touched at restarts only)

// Software-pipeline depth of the substitution loads (columns fetched ahead of their use) per instantiation: ptxas'
// schedule is sensitive to it (tools/class_sweep.py measures every class; SW4_S16_PREFETCH overrides for such sweeps).
template <int R>
__host__ __device__ constexpr int s16_prefetch_depth() {
#ifdef SW4_S16_PREFETCH
    return SW4_S16_PREFETCH;
#else
    // measured per instantiation (B200, q = 1000, G = 16, depths 5..16): R = 20: 6 -> 7.04 vs 6.53 TCUPS at 8..14;
    // R = 24: 14 -> 6.94 vs 6.76; R = 30: 14 -> 7.02 vs 6.75; R = 32: 7 -> 6.78 vs 6.69; elsewhere 10 is at the top
    return R == 20 ? 6 : (R == 24 || R == 28 || R == 30) ? 14 : R == 32 ? 7 : 10;
#endif
}

// The same refill with everything that does not change from batch to batch precomputed once per thread and parked in
// shared memory (round 2; ptxas re-derives such values from the thread index every batch rather than spend two
// registers on them, ~60 instructions per thread and batch): thread t copies the 16-byte piece c = t & 3 of the rows
// f = (t >> 2) + 128 k, k = 0..3. plan[t] = (ring byte address of slot 0 of row f0 / piece c, byte offset of the same
// piece at position 0 inside the profile).
constexpr int kFillPlanBytes = kS16Threads * 8;
__device__ __forceinline__ void ring_fill_plan(uint32_t planBase, uint32_t ringBase, int profStride) {
    const int f0 = threadIdx.x >> 2, c = threadIdx.x & 3;
    const uint32_t dst0 = ringBase + (f0 * kRingStride + 32 + 4 * c) * 4;
    const uint32_t srcOff0 = ((uint32_t)f0 * (uint32_t)profStride + 4u * c) * 4u;
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(planBase + threadIdx.x * 8), "r"(dst0), "r"(srcOff0) : "memory");
}
__device__ __forceinline__ void ring_fill_fast(uint32_t planBase, const uint32_t* __restrict__ profile, int profStride, int x0, int p0) {
    constexpr int kRowsPerRound = kS16Threads / 4;                             // fused-pair rows covered by one round of all threads
    constexpr int kRounds = (kFused + kRowsPerRound - 1) / kRowsPerRound;      // 4 with 512 threads
    uint32_t dst, srcOff;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(dst), "=r"(srcOff) : "r"(planBase + threadIdx.x * 8));
    const int slot0 = x0 & (kRingSlots - 1);
    const bool mirror = slot0 >= 32;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(profile) + (srcOff + (uint32_t)p0 * 4u);
    const uint32_t rowStep = (uint32_t)kRowsPerRound * (uint32_t)profStride * 4u;
    dst += slot0 * 4;
#pragma unroll
    for (int k = 0; k < kRounds; k++) {
        if (k < kRounds - 1 || threadIdx.x < (kFused - (kRounds - 1) * kRowsPerRound) * 4) {
            cp_async16(dst + k * (kRowsPerRound * kRingStride * 4), src);
            if (mirror) cp_async16(dst + k * (kRowsPerRound * kRingStride * 4) - kRingSlots * 4, src);
        }
        src += rowStep;
    }
    cp_async_commit();
}

constexpr int kBorderStageBytes = 2 * kBatchSteps * 16;  // per group: two batches of (H_a, E_a, H_b, E_b) per step
template <int R, bool MULTI = false>
__host__ __device__ constexpr int s16_smem_bytes() {
    return kRingBytes + kS16Warps * 32 * R * 2 + kS16Warps * 4 * kGroupStateInts * 4 + (MULTI ? kS16Warps * 2 * kBorderStageBytes : 0) +
           kS16Threads * 8;  // ... + the refill plan (ring_fill_plan)
}

// Refill the ring slots of query rows [x0, x0+16) (x0 % 16 == 0); p0 = x0 mod periodRows (periodRows % 16 == 0).
__device__ __forceinline__ void ring_fill(uint32_t ringBase, const uint32_t* __restrict__ profile, int profStride, int x0,
                                          int p0) {
    const int slot0 = x0 & (kRingSlots - 1);
    for (int id = threadIdx.x; id < kFused * (kBatchRows / 4); id += kS16Threads) {
        const int f = id >> 2, c = id & 3;
        const int slot = slot0 + 4 * c;
        const uint32_t* src = profile + (size_t)f * profStride + p0 + 4 * c;
        const uint32_t dst = ringBase + (f * kRingStride + 32 + slot) * 4;
        cp_async16(dst, src);
        if (slot >= 32) cp_async16(dst - kRingSlots * 4, src);
    }
    cp_async_commit();
}

__device__ __forceinline__ uint4 lds_u128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
template <int IMM>
__device__ __forceinline__ uint4 lds_u128_imm(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "n"(IMM));
    return v;
}
__device__ __forceinline__ void sts_u128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Gap-score sets with their own instantiations (gop, gex replicated in both halves): 0 = run-time values,
// 1 = -11 / -1 (BLOSUM62's default and what the reference's align always runs),
// 2 = -13 / -2 (the documented BLOSUM45 / BLOSUM50 default), 3 = -10 / -1 (BLOSUM80).
constexpr int kS16GapSets = 4;
__host__ __device__ constexpr uint2 s16_gap_set(int gaps) {
    return gaps == 1 ? uint2{0xfff5fff5u, 0xffffffffu} : gaps == 2 ? uint2{0xfff3fff3u, 0xfffefffeu}
         : gaps == 3 ? uint2{0xfff6fff6u, 0xffffffffu} : uint2{0u, 0u};
}
inline int s16_gap_set_for(uint32_t gop2, uint32_t gex2) {
    for (int g = 1; g < kS16GapSets; g++)
        if (s16_gap_set(g).x == gop2 && s16_gap_set(g).y == gex2) return g;
    return 0;
}

// R = register columns per lane. MULTI = items may span several 16 x R-column segments (G must be 16).
// GAPS = gap-score set compiled into the instructions (0 = read prm.gop2 / prm.gex2).
template <int R, bool MULTI = false, int GAPS = 0>
__global__ void __launch_bounds__(kS16Threads, 1) sw_s16_kernel(const S16Params prm) {
    static_assert(R % 2 == 0, "a lane's staging slice (R u16 codes) must be a whole number of 4-byte words");
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long tStart;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tStart));
    const uint32_t ringBase = (uint32_t)__cvta_generic_to_shared(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int hl = lane & 15;                          // lane inside its half-warp = pipeline stage
    const int logG = prm.logG, G = 1 << logG;
    const int g = lane >> logG, m = lane & (G - 1);    // group inside the warp, lane inside the group
    const int groupsPerWarp = 32 >> logG;
    const int P = prm.period;                          // steps
    const int periodRows = 2 * P;
    const unsigned groupMask = ((G == 16 ? 0xffffu : 0xffu)) << (g << logG);
    const int leader = g << logG;                      // lane index of the group's first lane
    const uint32_t NEG2 = ((uint32_t)(uint16_t)kNegS16 << 16) | (uint16_t)kNegS16;
    // Gap scores: instruction immediates in the specialised instantiations (GAPS > 0), registers otherwise. An immediate
    // saves one register-file read on three of the 5.5 recurrence instructions of a cell-pair, and operand delivery is
    // what this loop is short of (measured on the 1 M x 256 benchmark: 7.18 -> 7.66 TCUPS).
    const uint32_t SW4_GOP2 = GAPS > 0 ? s16_gap_set(GAPS).x : prm.gop2;
    const uint32_t SW4_GEX2 = GAPS > 0 ? s16_gap_set(GAPS).y : prm.gex2;

    // MULTI: the CTA takes a free slot of the border buffer (the length classes of a scan run concurrently and share it;
    // a CTA that finds none waits for one: slot holders never wait for anything, so this cannot deadlock)
    int borderSlot = 0;
    if constexpr (MULTI) {
        __shared__ int sBorderSlot;
        if (threadIdx.x == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            int i = (int)(smid % (unsigned)prm.numBorderSlots);
            while (atomicCAS(prm.borderSlots + i, 0, 1) != 0) {
                if (++i == prm.numBorderSlots) i = 0;
                __nanosleep(200);
            }
            sBorderSlot = i;
        }
        __syncthreads();
        borderSlot = sBorderSlot;
    }

    // Warps that will never get work (the class was spread over more SMs than it can fill, prm.activeGroups) only keep
    // the CTA's ring going: same barriers and their share of every refill, none of the arithmetic, so the busy warps get
    // their scheduler to themselves.
    if (warp * groupsPerWarp >= prm.activeGroups) {
        for (int i = threadIdx.x; i < kFused * kRingStride; i += kS16Threads) reinterpret_cast<uint32_t*>(smem)[i] = NEG2;
        __syncthreads();
        ring_fill(ringBase, prm.profile, prm.profStride, 0, 0);
        int pf = kBatchRows % periodRows;
#pragma unroll 1
        for (int batch = 0;; ++batch) {
            cp_async_wait_all();
            if (!__syncthreads_or(false)) break;
            ring_fill(ringBase, prm.profile, prm.profStride, (batch + 1) * kBatchRows, pf);
            pf += kBatchRows;
            if (pf >= periodRows) pf -= periodRows;
        }
        return;
    }

    // this lane's slice of the group's staging area: R fused u16 column codes of the next pair-block
    const uint32_t stageLane = ringBase + kRingBytes + (warp * 32 + lane) * (R * 2);
    // group state in shared memory: [0] subject0 [1] subject1 [2] segments left after the current one
    // [3] look-ahead valid [4] look-ahead block [5] look-ahead starts a new item [6] look-ahead item index
    volatile int* gs = reinterpret_cast<volatile int*>(smem + kRingBytes + kS16Warps * 32 * R * 2) +
                       (warp * 4 + g) * kGroupStateInts;

    // MULTI: border rows of this group in global memory (entry e <-> leader step e - 16, so that the last lane, which
    // trails the leader by 15 steps, never indexes below the array) and its staging area in shared memory
    uint4* borderBase = nullptr;
    uint32_t bstage = 0;
    if constexpr (MULTI) {
        const size_t groupGlobal = (size_t)borderSlot * (kS16Warps * 2) + warp * 2 + g;
        borderBase = prm.border + groupGlobal * (size_t)prm.borderStride + 16;
        bstage = ringBase + kRingBytes + kS16Warps * 32 * R * 2 + kS16Warps * 4 * kGroupStateInts * 4 + (warp * 2 + g) * kBorderStageBytes;
    }
    bool useBorder = false;       // MULTI: the segment being computed continues an item (its left border is in `border`)
    bool nextUseBorder = false;   // MULTI: the same for the look-ahead segment
    int pl = 0;                   // MULTI: the group leader's step p at the start of the current batch

    uint32_t colAddr[R];  // ring byte address of this column's fused-pair row (lane and batch-phase offsets folded in)
    uint32_t Hp[R];       // H of the lower row (b) of the previous step
    uint32_t F[R];        // F for the upper row (a) of the next step
    uint32_t mx = 0, HlastA = 0, ElastA = NEG2, ElastB = NEG2, HinPrevB = 0;
#pragma unroll
    for (int j = 0; j < R; j++) { colAddr[j] = ringBase; Hp[j] = 0; F[j] = NEG2; }
    // p = this lane's step in the period-P schedule, (t - hl) mod P: it works on query rows 2p and 2p+1. The group
    // restarts (stores the finished pair, loads the next one) when its first lane is at step 0, i.e. when this lane is at
    // step pRestart. Restarts can only fall on the first step of a batch (P and the group offsets are multiples of 8).
    int p = (hl == 0) ? 0 : P - hl;
    const int pRestart = (m == 0) ? 0 : P - m;
    bool haveWork = false;    // a segment is being computed
    bool alive = (warp * groupsPerWarp + g) < prm.activeGroups;  // the group still has something to compute or start

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
            if (item < prm.numItems) { blk = MULTI ? prm.items[item].firstBlock * prm.blockScale : item; isNew = 1; }
        }
        if (m == 0) { gs[3] = blk >= 0; gs[4] = blk; gs[5] = isNew; gs[6] = item; }
        if constexpr (MULTI) nextUseBorder = blk >= 0 && !isNew;
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
    constexpr int kSmemBytes = s16_smem_bytes<R, MULTI>();
    const uint32_t fillPlan = ringBase + kSmemBytes - kFillPlanBytes;
    ring_fill_plan(fillPlan, ringBase, prm.profStride);
    ring_fill_fast(fillPlan, prm.profile, prm.profStride, 0, 0);
    fetch_lookahead(false, 0);
    int pfill = kBatchRows % periodRows;  // (first row of the next fill) mod periodRows
    if constexpr (MULTI) {  // both staging buffers start out as the boundary column (H = 0, E = -inf)
        if (m < kBatchSteps) { sts_u128(bstage + m * 16, 0, NEG2, 0, NEG2); sts_u128(bstage + 128 + m * 16, 0, NEG2, 0, NEG2); }
    }
    bool writeBorder = false;  // MULTI: another segment of the item follows, so the last lane stores its column

    uint32_t phaseBase = ringBase + (32 - 2 * hl) * 4;  // + 64 bytes (16 rows) per batch, wrapping every 4 batches
#pragma unroll 1
    for (int batch = 0;; ++batch) {
        cp_async_wait_all();
        if (!__syncthreads_or(alive)) break;
        ring_fill_fast(fillPlan, prm.profile, prm.profStride, (batch + 1) * kBatchRows, pfill);
        pfill += kBatchRows;
        if (pfill >= periodRows) pfill -= periodRows;
        if (batch > 0) {
            const int delta = (batch & 3) ? kBatchRows * 4 : -(kRingSlots - kBatchRows) * 4;
            phaseBase += delta;
#pragma unroll
            for (int j = 0; j < R; j++) colAddr[j] += delta;
        }
        if (p == pRestart && alive) {  // group restart: uniform in the group, divergent across groups
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
            if constexpr (MULTI) writeBorder = false;
            if (laValid) {
                if (laNew) {
                    S16Item it;
                    if constexpr (MULTI) {
                        it = prm.items[gs[6]];
                    } else {
                        const int k = gs[6];
                        it.subject0 = prm.firstSubject + 2 * k;
                        it.subject1 = (2 * k + 1 < prm.numSubjects) ? it.subject0 + 1 : -1;
                        it.firstBlock = k;
                        it.numSegments = 1;
                    }
                    int nseg = it.numSegments;
                    if constexpr (MULTI) {
                        nseg *= prm.blockScale;
                        if (prm.blockScale > 1) {  // trailing segments that hold padding only are not computed
                            const int len0 = prm.lengths[it.subject0], len1 = it.subject1 >= 0 ? prm.lengths[it.subject1] : 0;
                            nseg = max(1, min(nseg, (max(len0, len1) + G * R - 1) / (G * R)));
                        }
                    }
                    __syncwarp(groupMask);
                    if (m == 0) { gs[0] = it.subject0; gs[1] = it.subject1; gs[2] = nseg - 1; }
                    segsLeft = nseg - 1;
                    mx = 0;
                } else {  // the next segment of the same item (MULTI only)
                    __syncwarp(groupMask);
                    if (m == 0) gs[2] = segsLeft - 1;
                    segsLeft -= 1;
                }
                if constexpr (MULTI) { useBorder = !laNew; writeBorder = segsLeft > 0; }
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
                // the hand-over registers start from the boundary values too: between its restart and its first real row a
                // lane runs up to G-1 gap rows whose inputs then come from reset registers of its left neighbour (which
                // is in the same state), so nothing of the previous alignment can leak into the new one and the steps need
                // no "is this a real row" test
                HinPrevB = 0; HlastA = 0; ElastA = NEG2; ElastB = NEG2;
                __syncwarp(groupMask);
                fetch_lookahead(segsLeft > 0, laBlk);
            }
        }
        uint32_t bstageRd = 0;
        uint4* borderWr = nullptr;
        bool wb15 = false;
        if constexpr (MULTI) {
            // left border of the NEXT batch's steps -> the other staging buffer (cp.async, or the boundary column when that
            // batch belongs to the first segment of an item); the next batch starts a new segment when the leader wraps
            const int plNext = (pl + kBatchSteps >= P) ? 0 : pl + kBatchSteps;
            const bool ub = (plNext == 0) ? nextUseBorder : useBorder;
            if (m < kBatchSteps) {
                // Entries of steps beyond the query are never taken from memory: the last lane only ever stores entries
                // up to P - 16 and whatever older scans (longer queries, other items) left behind there is not bounded by
                // this item's scores - as input of the first lane's gap rows it could leak into the running maximum.
                const uint32_t dst = bstage + ((batch + 1) & 1) * (kBatchSteps * 16) + m * 16;
                if (ub && 2 * (plNext + m) < prm.qlen) cp_async16(dst, borderBase + plNext + m);
                else sts_u128(dst, 0, NEG2, 0, NEG2);
            }
            cp_async_commit();
            bstageRd = bstage + (batch & 1) * (kBatchSteps * 16);
            borderWr = borderBase + (pl - (G - 1));  // the last lane trails the leader by G - 1 steps
            wb15 = writeBorder && m == G - 1;
        }
        static_for<kBatchSteps>([&](auto stepIndex) {
            constexpr int i = decltype(stepIndex)::value;
            // systolic hand-over from the previous lane: it worked on this row pair one step earlier
            uint32_t HinA = __shfl_up_sync(0xffffffffu, HlastA, 1);
            uint32_t EinA = __shfl_up_sync(0xffffffffu, ElastA, 1);
            uint32_t HinB = __shfl_up_sync(0xffffffffu, Hp[R - 1], 1);
            uint32_t EinB = __shfl_up_sync(0xffffffffu, ElastB, 1);
            // the first lane of a group takes the boundary column (H = 0, E = -inf) instead of its neighbour's values
            if constexpr (MULTI) {
                // left border of this step's two rows (a group-wide broadcast load: predicating it on the first lane costs
                // registers - more spills in the R >= 28 instantiations - for nothing measurable)
                const uint4 bv = lds_u128_imm<i * 16>(bstageRd);
                if (m == 0) { HinA = bv.x; EinA = bv.y; HinB = bv.z; EinB = bv.w; }
            } else {
                if (m == 0) { HinA = 0; EinA = NEG2; HinB = 0; EinB = NEG2; }
            }
            {
                uint32_t E1 = EinA, E2 = EinB;
                // substitution words are fetched kPrefetch columns ahead of their use: an explicit software pipeline, because
                // how far ptxas hoists the loads on its own varies with unrelated source changes (measured on the 1M x 256
                // benchmark: distance 1..4: 6.47, 5: 6.69, 8..16: 6.85 TCUPS for R = 32; 6.82 -> 6.89 for R = 16)
                constexpr int kPrefetch = s16_prefetch_depth<R>();
                uint2 sq[kPrefetch + 1];
#pragma unroll
                for (int c = 0; c <= kPrefetch && c < R; c++) sq[c] = lds_u64_imm<i * 8>(colAddr[c]);
                uint32_t da = __vadd2(HinPrevB, sq[0].x);  // row a: diagonal = row b of the previous step, previous lane
                uint32_t db = __vadd2(HinA, sq[0].y);      // row b: diagonal = row a of this step, previous lane
                uint32_t ha = 0;
#pragma unroll
                for (int j = 0; j < R; j++) {
                    // look-ahead: the next column's diagonal terms are formed before Hp[j] is overwritten
                    const uint2 sn = sq[(j + 1) % (kPrefetch + 1)];
                    if (j + 1 + kPrefetch < R) sq[j % (kPrefetch + 1)] = lds_u64_imm<i * 8>(colAddr[j + 1 + kPrefetch]);
                    uint32_t na = 0, nb = 0;
                    if (j + 1 < R) na = __vadd2(Hp[j], sn.x);
                    ha = __vimax3_s16x2_relu(da, E1, F[j]);
                    const uint32_t ta = __vadd2(ha, SW4_GOP2);
                    E1 = __viaddmax_s16x2(E1, SW4_GEX2, ta);
                    const uint32_t Fa = __viaddmax_s16x2(F[j], SW4_GEX2, ta);
                    if (j + 1 < R) nb = __vadd2(ha, sn.y);
                    const uint32_t hb = __vimax3_s16x2_relu(db, E2, Fa);
                    Hp[j] = hb;
                    const uint32_t tb = __vadd2(hb, SW4_GOP2);
                    E2 = __viaddmax_s16x2(E2, SW4_GEX2, tb);
                    F[j] = __viaddmax_s16x2(Fa, SW4_GEX2, tb);
                    mx = __vimax3_s16x2(mx, da, db);
                    da = na;
                    db = nb;
                }
                HlastA = ha;
                ElastA = E1;
                ElastB = E2;
                HinPrevB = HinB;
            }
            if constexpr (MULTI) {
                if (wb15) borderWr[i] = make_uint4(HlastA, ElastA, Hp[R - 1], ElastB);  // right border of this step's two rows
            }
        });
        p += kBatchSteps;  // P and the lanes' offsets to their group leader are multiples of the batch
        if (p >= P) p -= P;
        if constexpr (MULTI) pl = (pl + kBatchSteps >= P) ? 0 : pl + kBatchSteps;
    }
    if (threadIdx.x == 0) {
        unsigned long long tEnd;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tEnd));
        atomicMax(prm.elapsedNs, tEnd - tStart);
        atomicMax(prm.elapsedNs + 32, ~tStart);  // earliest CTA start (as a max of the complement)
        atomicMax(prm.elapsedNs + 64, tEnd);     // latest CTA end
        if constexpr (MULTI) {  // every warp is past the last barrier: nobody touches the border arrays any more
            __threadfence();
            atomicExch(prm.borderSlots + borderSlot, 0);
        }
    }
}

}  // namespace sw4
