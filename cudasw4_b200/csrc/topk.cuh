// Device-side top-k selection under the total order (score descending, database id ascending).
//
// Replaces the reference's "write 8 bytes per subject, radix-sort everything in chunks of 1e6, merge"
// (src/cudasw4.cuh:1365-1401, src/util.cuh:159-192) by a two-pass selection that reads the int32 score array once
// per pass and never materialises more than k candidates per block:
//   pass 1  every block owns a contiguous range of subjects. It finds the exact k-th largest score T of its range with
//           two shared-memory histograms (score >> 8, then score & 255 inside the deciding bin) and emits all
//           entries above T plus the first (k - #above) entries equal to T in index order: exactly its local top-k.
//   pass 2  one block bitonic-sorts the <= 8192 surviving (score, index) keys and writes the k best with global ids.
// Shard-local indices are ascending in global id, so "first in index order" is the reference's tie rule
// (ascending DB id) without the 1e6-chunk artefact.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sw4 {

constexpr int kTopkThreads = 1024;
constexpr int kTopkHiBins = 4096;     // score >> 8, clipped: exact up to scores of 2^20
constexpr int kTopkMaxCandidates = 8192;

struct TopkCand { int32_t score; int32_t index; };

__device__ __forceinline__ int topk_hi(int s) { return min(max(s, 0) >> 8, kTopkHiBins - 1); }


// Block-wide: given hist[0..nbins) in shared memory, find the highest bin b such that (count of entries in bins > b) < k
// <= (count in bins >= b), or b = 0 when even all bins together hold fewer than k. Returns (b, count above b) in
// sBin/sAbove. Every thread owns nbins/blockDim consecutive bins; the suffix sums come from a reverse block scan.
__device__ __forceinline__ void topk_find_bin(const int* hist, int nbins, int k, int baseAbove, int* warpTotals, int* sBin,
                                              int* sAbove) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nthreads = blockDim.x;
    const int per = (nbins + nthreads - 1) / nthreads;
    // thread t owns bins [lo, hi) counted from the TOP: logical index i <-> bin nbins-1-i
    const int lo = tid * per, hi = min(nbins, lo + per);
    int mine = 0;
    for (int i = lo; i < hi; i++) mine += hist[nbins - 1 - i];
    // inclusive prefix over threads (in top-down order)
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warpTotals[w] = incl;
    if (tid == 0) { *sBin = 0; *sAbove = -1; }
    __syncthreads();
    int before = baseAbove;
    for (int x = 0; x < w; x++) before += warpTotals[x];
    const int exclusive = before + incl - mine;  // entries in bins above this thread's range
    if (exclusive < k && exclusive + mine >= k) {   // the crossing is inside this thread's bins (exactly one thread)
        int above = exclusive;
        for (int i = lo; i < hi; i++) {
            const int c = hist[nbins - 1 - i];
            if (above + c >= k) { *sBin = nbins - 1 - i; *sAbove = above; break; }
            above += c;
        }
    }
    __syncthreads();
    if (*sAbove < 0) {  // fewer than k entries in total: bin 0 decides, everything above it is kept
        if (tid == 0) {
            int total = baseAbove;
            for (int x = 0; x < (nthreads + 31) / 32; x++) total += warpTotals[x];
            *sAbove = total - hist[0];
            *sBin = 0;
        }
        __syncthreads();
    }
}

// scores[begin+i]; indexOf = begin+i (or indices[begin+i] when indices != nullptr, which must be ascending)
__global__ void __launch_bounds__(kTopkThreads) topk_pass1_kernel(const int32_t* __restrict__ scores,
                                                                  const int32_t* __restrict__ indices, long long n,
                                                                  int k, TopkCand* __restrict__ out) {
    __shared__ int hist[kTopkHiBins];
    __shared__ int warpTotals[kTopkThreads / 32];
    __shared__ int sBin, sAbove, sRunning, sEmitted;
    const int tid = threadIdx.x;
    const long long per = (n + gridDim.x - 1) / gridDim.x;
    const long long begin = per * blockIdx.x;
    const long long end = min(n, begin + per);
    TopkCand* myOut = out + (size_t)blockIdx.x * k;
    for (int i = tid; i < k; i += blockDim.x) myOut[i] = TopkCand{-1, 0x7fffffff};
    const long long cnt = end - begin;
    if (cnt <= 0) return;

    // ---- level 1: histogram of score >> 8 ----
    for (int i = tid; i < kTopkHiBins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (long long i = begin + tid; i < end; i += blockDim.x) atomicAdd(&hist[topk_hi(scores[i])], 1);
    __syncthreads();
    topk_find_bin(hist, kTopkHiBins, k, 0, warpTotals, &sBin, &sAbove);  // bin b decides (or everything fits: b == 0)
    const int bin = sBin;
    const int aboveBin = sAbove;
    // ---- level 2: exact threshold inside the deciding bin ----
    // (the top bin also holds clipped scores >= 2^20; then the low byte is not the full story - handled below by
    //  treating every entry of a clipped top bin as "equal" and letting pass 2 order them exactly)
    __syncthreads();
    for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const bool clippedBin = (bin == kTopkHiBins - 1);
    for (long long i = begin + tid; i < end; i += blockDim.x) {
        const int s = scores[i];
        if (topk_hi(s) == bin) atomicAdd(&hist[clippedBin ? 0 : (max(s, 0) & 255)], 1);
    }
    __syncthreads();
    topk_find_bin(hist, 256, k, aboveBin, warpTotals, &sBin, &sAbove);
    if (tid == 0) { sRunning = 0; sEmitted = 0; }
    __syncthreads();
    const int T = clippedBin ? (bin << 8) : ((bin << 8) | sBin);  // k-th largest score of this range (or lower bound)
    const int need = k - sAbove;                                  // how many entries == T (>= T if clipped) to keep
    // ---- emit: everything above T (unordered), then the first `need` entries equal to T in index order ----
    for (long long base = begin; base < end; base += blockDim.x) {
        const long long i = base + tid;
        int s = -1;
        if (i < end) s = scores[i];
        const bool isAbove = !clippedBin && i < end && s > T;
        const bool isEq = i < end && (clippedBin ? (s >= T) : (s == T));
        if (isAbove) {
            const int pos = atomicAdd(&sEmitted, 1);
            myOut[pos] = TopkCand{s, indices ? indices[i] : (int32_t)i};
        }
        const unsigned bal = __ballot_sync(0xffffffffu, isEq);
        const int lane = tid & 31, w = tid >> 5;
        if (lane == 0) warpTotals[w] = __popc(bal);
        __syncthreads();
        int before = sRunning;
        for (int x = 0; x < w; x++) before += warpTotals[x];
        const int rank = before + __popc(bal & ((1u << lane) - 1u));
        if (isEq && rank < need) myOut[sAbove + rank] = TopkCand{s, indices ? indices[i] : (int32_t)i};
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int x = 0; x < kTopkThreads / 32; x++) tot += warpTotals[x];
            sRunning += tot;
        }
        __syncthreads();
        if (!clippedBin && sRunning >= need && sEmitted >= sAbove) break;  // uniform: shared values after barrier
    }
}

// One block: sort numCand (<= 8192) candidates by (score desc, index asc), write the k best; index -> global id.
__global__ void __launch_bounds__(kTopkThreads) topk_pass2_kernel(const TopkCand* __restrict__ cand, int numCand, int k,
                                                                  const int32_t* __restrict__ globalIds,
                                                                  int32_t* __restrict__ outScores,
                                                                  int32_t* __restrict__ outIds, int* outCount) {
    extern __shared__ unsigned long long keys[];  // pow2 >= numCand
    int n2 = 1;
    while (n2 < numCand) n2 <<= 1;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        unsigned long long key = 0;  // sorts last
        if (i < numCand && cand[i].score >= 0)
            key = ((unsigned long long)(unsigned)(cand[i].score + 1) << 32) | (unsigned)(0x7fffffff - cand[i].index);
        keys[i] = key;
    }
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < n2 / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = keys[lo], b = keys[hi];
                if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
    }
    int valid = 0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const unsigned long long key = (i < n2) ? keys[i] : 0ull;
        if (key != 0) {
            const int idx = 0x7fffffff - (int)(unsigned)(key & 0xffffffffu);
            outScores[i] = (int)(key >> 32) - 1;
            outIds[i] = globalIds ? globalIds[idx] : idx;
        } else {
            outScores[i] = -1;
            outIds[i] = -1;
        }
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < k && i < n2; i++) valid += keys[i] != 0;
        *outCount = valid;
    }
}

}  // namespace sw4
