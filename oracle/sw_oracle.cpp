// ============================================================================================================
// TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of the reference's hot path. Nothing in the product
// (cudasw4_b200/, include/) links, loads or calls this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may.
//
// Parity status: PINNED.  Checked in tests/test_oracle.py against (1) the reference's own scalar CPU Gotoh compiled
// from /root/reference (oracle/ref_harness.cu -> fixtures in tests/golden/), (2) the reference's substitution
// tables, letter map, PseudoDB generator and partition boundaries dumped the same way, (3) the reference `makedb`
// binary's output files, and (4) the known-answer tables of SURVEY.md §8c.
//
// What each function restates (file:line are relative to /root/reference):
//   sw4o_convert_letters      src/convert.cuh:6-34            (20 upper-case letters -> 0..19, all else -> 20)
//   sw4o_substitution_matrix  src/types.hpp:29-201            (21x21 tables; data in oracle/blosum_data.h)
//   sw4o_gotoh_score          src/cudasw4.cuh:2331-2392       (scalar int32 local Gotoh, score only)
//   sw4o_scan                 src/cudasw4.cuh:767-796         (all subjects, OpenMP over subjects)
//   sw4o_topk                 src/cudasw4.cuh:1365-1401       (sort by score desc; ties by ascending DB id)
//   sw4o_pseudo_subject       src/dbdata.hpp:219-246          (mt19937(seed), uniform_int_distribution(0,19))
//   sw4o_length_partition     src/length_partitions.hpp:75-113, src/cudasw4.cuh:904-926
//   sw4o_gcups                src/cudasw4.cuh:2264-2271
// ============================================================================================================
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "blosum_data.h"

namespace {

const signed char* table_for(int blosum) {
    switch (blosum) {
        case 45: return ORACLE_BLOSUM45;
        case 50: return ORACLE_BLOSUM50;
        case 62: return ORACLE_BLOSUM62;
        case 80: return ORACLE_BLOSUM80;
    }
    return nullptr;
}

const int kBoundaries[36] = {48, 64, 80, 96, 112, 128, 144, 160, 176, 192, 208, 224, 240, 256, 288, 320, 352, 384,
                             416, 448, 480, 512, 576, 640, 704, 768, 832, 896, 960, 1024, 1088, 1152, 1216, 1280,
                             8000, 2147483646};

// Rolling two-row formulation of the textbook recurrence; identical cell values to cudasw4.cuh:2331-2392.
//   E[i][j] = max(E[i][j-1] + gex, H[i][j-1] + gop)   (gap in the query direction, along the subject)
//   F[i][j] = max(F[i-1][j] + gex, H[i-1][j] + gop)
//   H[i][j] = max(0, H[i-1][j-1] + M[q_i][s_j], E[i][j], F[i][j]);  score = max H
int gotoh(const signed char* M, const unsigned char* q, int qlen, const unsigned char* s, int slen, int gop, int gex,
          int* Hrow, int* Frow) {
    const int NEG = -10000;  // the reference's NEGINFINITY (cudasw4.cuh:2339)
    for (int j = 0; j <= slen; j++) { Hrow[j] = 0; Frow[j] = NEG; }
    int best = 0;
    for (int i = 0; i < qlen; i++) {
        const signed char* Mrow = M + 21 * q[i];
        int diag = 0;      // H[i-1][j-1]
        int left = 0;      // H[i][j-1]
        int E = NEG;
        for (int j = 1; j <= slen; j++) {
            const int up = Hrow[j];
            E = std::max(E + gex, left + gop);
            const int F = std::max(Frow[j] + gex, up + gop);
            int h = std::max(0, std::max(diag + Mrow[s[j - 1]], std::max(E, F)));
            diag = up;
            Hrow[j] = h;
            Frow[j] = F;
            left = h;
            if (h > best) best = h;
        }
    }
    return best;
}

}  // namespace

extern "C" {

void sw4o_convert_letters(const char* in, unsigned char* out, long n) {
    static const char order[] = "ARNDCQEGHILKMFPSTWYV";
    unsigned char map[256];
    std::memset(map, 20, sizeof(map));
    for (int i = 0; i < 20; i++) map[(unsigned char)order[i]] = (unsigned char)i;
    for (long i = 0; i < n; i++) out[i] = map[(unsigned char)in[i]];
}

int sw4o_substitution_matrix(int blosum, signed char* out441) {
    const signed char* t = table_for(blosum);
    if (!t) return -1;
    std::memcpy(out441, t, 441);
    return 0;
}

// q, s are residue CODES (0..20).
int sw4o_gotoh_score(int blosum, const unsigned char* q, int qlen, const unsigned char* s, int slen, int gop, int gex) {
    const signed char* t = table_for(blosum);
    if (!t) return -1;
    std::vector<int> H(slen + 1), F(slen + 1);
    return gotoh(t, q, qlen, s, slen, gop, gex, H.data(), F.data());
}

// Score one query against n subjects stored makedb-style: chars (codes) + byte offsets + true lengths.
// threads <= 0 -> all available. Returns the number of threads used.
int sw4o_scan(int blosum, const unsigned char* q, int qlen, const unsigned char* chars, const std::size_t* offsets,
              const std::int32_t* lengths, long n, int gop, int gex, std::int32_t* scores_out, int threads) {
    const signed char* t = table_for(blosum);
    if (!t) return -1;
    int used = 1;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
    used = threads;
#pragma omp parallel num_threads(threads)
#endif
    {
        std::vector<int> H, F;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (long i = 0; i < n; i++) {
            const int slen = lengths[i];
            if ((int)H.size() < slen + 1) { H.resize(slen + 1); F.resize(slen + 1); }
            scores_out[i] = gotoh(t, q, qlen, chars + offsets[i], slen, gop, gex, H.data(), F.data());
        }
    }
    return used;
}

// top-k under (score desc, id asc). ids_out/scores_out hold min(k, n) entries; returns that count.
long sw4o_topk(const std::int32_t* scores, long n, long k, std::int32_t* scores_out, std::int32_t* ids_out) {
    std::vector<std::int32_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    const long m = std::min(k, n);
    std::partial_sort(idx.begin(), idx.begin() + m, idx.end(), [&](std::int32_t a, std::int32_t b) {
        if (scores[a] != scores[b]) return scores[a] > scores[b];
        return a < b;
    });
    for (long i = 0; i < m; i++) { ids_out[i] = idx[i]; scores_out[i] = scores[idx[i]]; }
    return m;
}

// The single subject the reference's PseudoDB replicates: `length` letters -> codes (libstdc++ distributions).
void sw4o_pseudo_subject(int length, int seed, unsigned char* codes_out) {
    static const char letters[] = "ARNDCQEGHILKMFPSTWYV";
    std::mt19937 gen(seed);
    std::uniform_int_distribution<> dist(0, 19);
    std::vector<char> tmp(length);
    for (int i = 0; i < length; i++) tmp[i] = letters[dist(gen)];
    sw4o_convert_letters(tmp.data(), codes_out, length);
}

int sw4o_num_length_partitions() { return 36; }
int sw4o_length_partition_boundary(int i) { return kBoundaries[i]; }
// length k is in partition i iff boundary[i-1] < k <= boundary[i]
int sw4o_length_partition(int length) {
    for (int i = 0; i < 36; i++)
        if (length <= kBoundaries[i]) return i;
    return 35;
}

double sw4o_gcups(double cells, double seconds) { return cells / 1000. / 1000. / 1000. / seconds; }

}  // extern "C"
