// Device-resident database layout + per-query preparation kernels.
//
// Replaces the reference's GpuDatabaseAllocation / batch copy plans (src/gpudatabaseallocation.cuh:22-61,
// src/dbbatching.cuh:16-99) and the per-block LUT expansion + query staging (src/dpx_s16_kernels.cuh:55-69,
// src/cudasw4.cuh:1280-1310):
//   * raw shard      chars / offsets / lengths exactly as makedb stores them (used by the exact 32-bit kernel)
//   * pair-blocks    for every length class (G lanes x R columns): consecutive subjects of the length-sorted shard are
//                    paired, and column c of the pair is stored as the fused code f = s0[c] + 21*s1[c] (u16), padded
//                    with the 'other' code 20; one block = G*R*2 contiguous bytes = one coalesced cp.async burst
//   * query profile  prof[f][p] = (M[q_p][s1] << 16) | (M[q_p][s0] & 0xffff) for every query position p, built once per
//                    scan (441 x q words); positions >= q hold -16000 ("gap rows" of the kernel's schedule)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sw4 {

// letters -> residue codes (reference src/convert.cuh:6-34): the 20 upper-case letters in NCBI order, all else 20
__global__ void convert_query_kernel(const char* __restrict__ letters, uint8_t* __restrict__ codes, int n, int padTo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= padTo) return;
    uint8_t c = 20;
    if (i < n) {
        const char a = letters[i];
        const char order[21] = "ARNDCQEGHILKMFPSTWYV";
#pragma unroll
        for (int k = 0; k < 20; k++)
            if (a == order[k]) c = (uint8_t)k;
    }
    codes[i] = c;
}

__global__ void build_profile_kernel(const uint8_t* __restrict__ qcodes, int qlen, const int8_t* __restrict__ matrix,
                                     uint32_t* __restrict__ profile, int stride) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;  // fused pair: s0 = f % 21 (low half), s1 = f / 21 (high half)
    if (p >= stride) return;
    uint32_t v = 0xc180c180u;  // (-16000, -16000)
    if (f >= 483) {
        // rows 483..503: plain int32 plane for the exact 32-bit array kernel (sw_s32_long_kernel)
        v = p < qlen ? (uint32_t)(int32_t)matrix[qcodes[p] * 21 + (f - 483)] : (uint32_t)(-(1 << 28));
    } else if (f >= 441) {
        // rows 441..461: single-residue plane with the score in the low half, rows 462..482: the same in the high half
        // (sw_s16_long_kernel adds one entry of each)
        const int s = (f - 441) % 21;
        const bool high = f >= 462;
        const uint32_t m = p < qlen ? (uint32_t)(uint16_t)(int16_t)matrix[qcodes[p] * 21 + s] : 0xc180u;
        v = high ? (m << 16) : m;
    } else if (p < qlen) {
        const int qc = qcodes[p];
        const int lo = matrix[qc * 21 + f % 21], hi = matrix[qc * 21 + f / 21];
        v = ((uint32_t)(uint16_t)(int16_t)hi << 16) | (uint16_t)(int16_t)lo;
    }
    profile[(size_t)f * stride + p] = v;
}

// One thread per (pair-block, column). blockItem[b] names the work item (a pair of subjects) block b belongs to; the
// block covers columns [seg*columns, (seg+1)*columns) of the pair, seg = b - item.firstBlock.
struct PairItem { int subject0, subject1, firstBlock, numSegments; };  // same layout as sw4::S16Item

__global__ void build_pair_blocks_kernel(const uint8_t* __restrict__ chars, const size_t* __restrict__ offsets,
                                         const int32_t* __restrict__ lengths, const PairItem* __restrict__ items,
                                         const int32_t* __restrict__ blockItem, int numBlocks, int columns,
                                         uint16_t* __restrict__ cols) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)numBlocks * columns;
    if (gid >= total) return;
    const int b = (int)(gid / columns), c = (int)(gid % columns);
    const PairItem it = items[blockItem[b]];
    const long long col = (long long)(b - it.firstBlock) * columns + c;
    int r0 = 20, r1 = 20;
    if (it.subject0 >= 0 && col < lengths[it.subject0]) r0 = chars[offsets[it.subject0] + col];
    if (it.subject1 >= 0 && col < lengths[it.subject1]) r1 = chars[offsets[it.subject1] + col];
    r0 = min(r0, 20); r1 = min(r1, 20);
    cols[gid] = (uint16_t)(r0 + 21 * r1);
}

}  // namespace sw4
