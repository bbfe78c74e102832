// TEST INFRASTRUCTURE ONLY — never linked into the product.
// Thin C-ABI harness around the *reference's own* host code, compiled from the sources where they lie
// under /root/reference (nothing is copied into this repo). It exposes, for pinning oracle/sw_oracle.cpp:
//   * the reference's private scalar CPU Gotoh (src/cudasw4.cuh:2331-2392, BLOSUM62 only),
//   * the four live 21x21 substitution tables (src/types.hpp:29-201),
//   * the letter->code map (src/convert.cuh:6-34),
//   * the PseudoDB generator (src/dbdata.hpp:219-272),
//   * the 36 length-partition boundaries (src/length_partitions.hpp:75-113).
// Built by oracle/Makefile into oracle/_ref/libref_harness.so (host code only; no kernel is ever launched).
#include <algorithm>
#include <array>
#include <cstring>
#include <iostream>
#include <memory>
#include <numeric>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include <fstream>
#include <map>
#include <future>
#include <mutex>
#include <thread>
#include <chrono>
#include <functional>
#include <limits>
#include <cassert>
#include <cstdint>
#include <string_view>
#include <type_traits>
#include <omp.h>
#include <thrust/device_vector.h>
#include <thrust/sort.h>
#include <thrust/merge.h>
#include <cub/cub.cuh>

#define private public
#define protected public
#include "cudasw4.cuh"
#undef private
#undef protected

extern "C" {

int ref_cpu_gotoh_blosum62(const char* q_codes, const char* s_codes, int qlen, int slen, int gop, int gex){
    // the member touches no object state; call it on raw aligned storage to avoid the GPU-requiring constructor
    alignas(alignof(cudasw4::CudaSW4)) static unsigned char storage[sizeof(cudasw4::CudaSW4)];
    auto* self = reinterpret_cast<cudasw4::CudaSW4*>(storage);
    return self->affine_local_DP_host_protein_blosum62_converted(q_codes, s_codes, qlen, slen, gop, gex);
}

// All subjects of a makedb-layout block against one query with the reference's own scalar routine, one OpenMP task per
// subject: the structure of its computeAllScoresCPU_blosum62 (src/cudasw4.cuh:767-796). Returns the threads used.
// (bench.py's CPU arm: "kind": "reference".)
int ref_cpu_scan_blosum62(const char* q_codes, int qlen, const char* chars, const size_t* offsets, const int* lengths, long n,
                          int gop, int gex, int* out, int threads){
    alignas(alignof(cudasw4::CudaSW4)) static unsigned char storage[sizeof(cudasw4::CudaSW4)];
    auto* self = reinterpret_cast<cudasw4::CudaSW4*>(storage);
    if(threads <= 0) threads = omp_get_max_threads();
    #pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
    for(long i = 0; i < n; i++){
        out[i] = self->affine_local_DP_host_protein_blosum62_converted(q_codes, chars + offsets[i], qlen, lengths[i], gop, gex);
    }
    return threads;
}

// type: 45, 50, 62, 80 -> the 21x21 "_20" table the shipped align uses (CAN_USE_FULL_BLOSUM is off)
int ref_blosum_table(int type, signed char* out441){
    auto put = [&](auto arr){ for(int i = 0; i < 441; i++) out441[i] = arr[i]; };
    switch(type){
        case 45: put(cudasw4::BLOSUM45_20::get1D()); return 0;
        case 50: put(cudasw4::BLOSUM50_20::get1D()); return 0;
        case 62: put(cudasw4::BLOSUM62_20::get1D()); return 0;
        case 80: put(cudasw4::BLOSUM80_20::get1D()); return 0;
    }
    return -1;
}

void ref_convert_aa(const char* in, char* out, long n){
    cudasw4::ConvertAA_20 conv;
    for(long i = 0; i < n; i++) out[i] = conv(in[i]);
}

// one pseudo-db subject (codes, padded to a multiple of 4), as PseudoDBdata builds it
long ref_pseudodb(long num, int length, int seed, char* chars_out, long chars_cap, int* lengths_out, size_t* offsets_out){
    cudasw4::PseudoDBdata db(num, length, seed);
    if(long(db.numChars()) > chars_cap) return -1;
    std::memcpy(chars_out, db.chars(), db.numChars());
    std::memcpy(lengths_out, db.lengths(), sizeof(int) * num);
    std::memcpy(offsets_out, db.offsets(), sizeof(size_t) * (num + 1));
    return long(db.numChars());
}

int ref_length_partition_boundaries(int* out, int cap){
    auto b = cudasw4::getLengthPartitionBoundaries();
    int n = int(b.size());
    for(int i = 0; i < n && i < cap; i++) out[i] = b[i];
    return n;
}

}
