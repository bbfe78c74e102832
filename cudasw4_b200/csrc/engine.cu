// Host engine behind the C ABI of include/sw4b200.h: database handling, sharding, per-GPU working sets, the scan
// pipeline (query -> profile -> score kernels -> exact re-scoring -> top-k -> merge) and timing.
// Mirrors the responsibilities of the reference's cudasw4::CudaSW4 (src/cudasw4.cuh:244-2454) with a different
// design: the shard is always device resident in a kernel-ready layout, one persistent launch per length class,
// device-side top-k, and only k (score, id) pairs per GPU ever reach the host.
#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cuda_runtime.h>

#include "../../include/sw4b200.h"
#include "blosum_tables.hpp"
#include "device_db.cuh"
#include "kernels_s16.cuh"
#include "kernels_s16_wide.cuh"
#include "kernels_s16_long.cuh"
#include "kernels_s32.cuh"
#include "kernels_s32_long.cuh"
#include "topk.cuh"

namespace sw4 {

// ---------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------
struct Error {
    int code;
    std::string msg;
};

[[noreturn]] static void fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw Error{code, buf};
}

#define SW4_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            ::sw4::fail(e_ == cudaErrorMemoryAllocation ? SW4_ERR_NOMEM : SW4_ERR_CUDA, "%s failed: %s (%s:%d)", \
                        #expr, cudaGetErrorString(e_), __FILE__, __LINE__);                              \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// length classes of the packed 16-bit kernel: G lanes x R columns, capacity G*R
// ---------------------------------------------------------------------------------------------------------------
// G = 1<<logG lanes x R columns per lane; capacity = G*R columns per pair-block. Subjects up to 512 residues run on the
// two-rows-per-step kernel (G = 8 or 16, kernels_s16.cuh); 513..1024 and the multi-segment class (1024-column segments,
// any length) on the full-warp one-row-per-step variant (G = 32, kernels_s16_wide.cuh).
struct LengthClass { int logG, R, capacity; bool wide, multi; };
static const LengthClass kLengthClasses[] = {
    {3, 4, 32, false, false}, {3, 6, 48, false, false}, {3, 8, 64, false, false}, {3, 10, 80, false, false},
    {3, 12, 96, false, false}, {3, 14, 112, false, false}, {3, 16, 128, false, false}, {3, 18, 144, false, false},
    {3, 20, 160, false, false}, {3, 22, 176, false, false}, {3, 24, 192, false, false}, {3, 26, 208, false, false},
    {3, 28, 224, false, false}, {3, 30, 240, false, false}, {4, 16, 256, false, false}, {4, 18, 288, false, false},
    {4, 20, 320, false, false}, {4, 22, 352, false, false}, {4, 24, 384, false, false}, {4, 26, 416, false, false},
    {4, 28, 448, false, false}, {4, 30, 480, false, false}, {4, 32, 512, false, false}, {5, 18, 576, true, false},
    {5, 20, 640, true, false}, {5, 22, 704, true, false}, {5, 24, 768, true, false}, {5, 26, 832, true, false},
    {5, 28, 896, true, false}, {5, 30, 960, true, false}, {5, 32, 1024, true, false}, {5, 32, 1024, true, true},
};
constexpr int kNumLengthClasses = sizeof(kLengthClasses) / sizeof(kLengthClasses[0]);
constexpr int kS16OverflowThreshold = 25000;  // reference MAX_ACC_SHORT, src/kernels.cuh:5
constexpr int kHalf2Threshold = 2048;         // reference MAX_ACC_HALF2, src/kernels.cuh:4
constexpr int kNumCounters = 8 + 32;           // [0] overflow [1] stat [2],[3] s32 tickets [4] top-k count [8+c] class tickets
constexpr int kShardBlock = 256;              // subjects per interleaving block (even => pairs never straddle)

static const int kRefBoundaries[36] = {48,  64,  80,  96,  112, 128, 144, 160, 176, 192,  208,  224,
                                       240, 256, 288, 320, 352, 384, 416, 448, 480, 512,  576,  640,
                                       704, 768, 832, 896, 960, 1024, 1088, 1152, 1216, 1280, 8000, 2147483646};

// ---------------------------------------------------------------------------------------------------------------
// host database (makedb format)
// ---------------------------------------------------------------------------------------------------------------
struct MappedFile {
    void* ptr = nullptr;
    size_t size = 0;
    MappedFile() = default;
    MappedFile(const MappedFile&) = delete;
    MappedFile& operator=(const MappedFile&) = delete;
    ~MappedFile() { if (ptr && size) munmap(ptr, size); }
    void open(const std::string& path, bool populate) {
        int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) fail(SW4_ERR_IO, "cannot open %s: %s", path.c_str(), strerror(errno));
        struct stat st;
        if (fstat(fd, &st) != 0) { ::close(fd); fail(SW4_ERR_IO, "cannot stat %s", path.c_str()); }
        size = (size_t)st.st_size;
        if (size > 0) {
            ptr = mmap(nullptr, size, PROT_READ, MAP_PRIVATE | (populate ? MAP_POPULATE : 0), fd, 0);
            if (ptr == MAP_FAILED) { ptr = nullptr; ::close(fd); fail(SW4_ERR_IO, "cannot mmap %s", path.c_str()); }
        }
        ::close(fd);
    }
};

struct HostDB {
    const uint8_t* chars = nullptr;
    const size_t* offsets = nullptr;
    const int32_t* lengths = nullptr;
    const char* headers = nullptr;
    const size_t* headerOffsets = nullptr;
    size_t n = 0;
    uint64_t residues = 0;
    int minLen = 0, maxLen = 0;
    // owners
    MappedFile fChars, fOffsets, fLengths, fHeaders, fHeaderOffsets;
    std::vector<uint8_t> vChars;
    std::vector<size_t> vOffsets, vHeaderOffsets;
    std::vector<int32_t> vLengths;
    std::vector<char> vHeaders;

    void finish() {
        residues = 0; minLen = n ? lengths[0] : 0; maxLen = 0;
        for (size_t i = 0; i < n; i++) {
            if (lengths[i] < 0) fail(SW4_ERR_INVALID, "negative sequence length at %zu", i);
            if (i && lengths[i] < lengths[i - 1]) fail(SW4_ERR_INVALID, "database is not sorted by length (id %zu)", i);
            residues += (uint64_t)lengths[i];
            minLen = std::min(minLen, lengths[i]);
            maxLen = std::max(maxLen, lengths[i]);
        }
        if (n > (size_t)0x7ffffffe) fail(SW4_ERR_INVALID, "too many sequences");
    }
};

// ---------------------------------------------------------------------------------------------------------------
// per-GPU working set
// ---------------------------------------------------------------------------------------------------------------
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    void alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        SW4_CUDA(cudaMalloc(&p, count * sizeof(T)));
        n = count;
    }
    void ensure(size_t count) { if (count > n) alloc(count + count / 4); }
    size_t bytes() const { return n * sizeof(T); }
};

struct ClassLayout {
    int cls = 0;               // index into kLengthClasses
    int first = 0, count = 0;  // local subject range
    int numItems = 0, numBlocks = 0;
    double rate = 1.0;         // measured SM-milliseconds per unit of modelled cost (feedback for the SM partition)
    double lastCost = 0;
    int lastGrid = 0;
    DevBuf<uint16_t> cols;
    DevBuf<S16Item> items;
};

constexpr int kMaxClassStreams = 32;
constexpr int kProfileRows = kFused + 63;  // 441 fused-pair rows + two s16 single-residue planes + one int32 plane (array kernels)

struct Shard {
    int device = 0;
    int smCount = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t evStart = nullptr, evK0 = nullptr, evK1 = nullptr, evStop = nullptr, evFork = nullptr;
    cudaStream_t classStreams[kMaxClassStreams] = {}, oddStreams[kMaxClassStreams] = {};
    cudaEvent_t evJoin[kMaxClassStreams] = {}, evJoinOdd[kMaxClassStreams] = {};
    // subject selection
    std::vector<int32_t> globalIds;  // local index -> global id (ascending)
    uint64_t residues = 0;
    size_t n = 0;
    bool uploaded = false;
    // raw shard
    DevBuf<uint8_t> dChars;
    DevBuf<size_t> dOffsets;
    DevBuf<int32_t> dLengths, dGlobalIds;
    std::vector<std::unique_ptr<ClassLayout>> classes;
    size_t numZeroLength = 0;  // the shard's first entries (it is sorted by length)
    // per scan
    DevBuf<int32_t> dScores, dOvfList;
    DevBuf<int> dCounters;  // [0] overflow count, [1] stat count, [2] ticket long, [3] ticket overflow, [4] topk count
    DevBuf<char> dQueryLetters;
    DevBuf<uint8_t> dQueryCodes;
    DevBuf<uint32_t> dProfile;
    DevBuf<int8_t> dMatrix;
    DevBuf<int2> dBorder;
    DevBuf<uint2> dBorderLong;            // per CTA of the long-subject array kernel: border rows between periods
    DevBuf<unsigned long long> dClassNs;  // per length class: run time of its last launch (written by the kernel)
    unsigned long long* hClassNs = nullptr;
    DevBuf<uint2> dBorderWide;  // left/right border columns of the multi-segment class, one row array per warp
    size_t borderWideStride = 0;
    DevBuf<TopkCand> dCand;
    DevBuf<int32_t> dTopScores, dTopIds;
    // pinned host staging
    int queryCapacity = 0, topCapacity = 0;  // what the per-scan scratch below is currently sized for
    size_t borderWarps = 0;
    char* hQuery = nullptr; size_t hQueryCap = 0;
    int32_t* hTop = nullptr; size_t hTopCap = 0;  // scores[k], ids[k], count, ovf, stat
    int launches = 0;

    ~Shard() {
        cudaSetDevice(device);
        if (hQuery) cudaFreeHost(hQuery);
        if (hTop) cudaFreeHost(hTop);
        if (hClassNs) cudaFreeHost(hClassNs);
        if (evStart) cudaEventDestroy(evStart);
        if (evK0) cudaEventDestroy(evK0);
        if (evK1) cudaEventDestroy(evK1);
        if (evStop) cudaEventDestroy(evStop);
        if (evFork) cudaEventDestroy(evFork);
        for (int i = 0; i < kMaxClassStreams; i++) {
            if (evJoin[i]) cudaEventDestroy(evJoin[i]);
            if (evJoinOdd[i]) cudaEventDestroy(evJoinOdd[i]);
            if (classStreams[i]) cudaStreamDestroy(classStreams[i]);
            if (oddStreams[i]) cudaStreamDestroy(oddStreams[i]);
        }
        if (stream) cudaStreamDestroy(stream);
    }
};

// steps (row pairs) between two alignments of a group: ceil(q/2) + G - 1 rounded up to a batch, at least 16 so that
// no lane of a half-warp starts before step 0
// The 256-column class can run as 16 lanes x 16 columns or as 8 lanes x 32 columns on the same pair-blocks (a block is
// 256 consecutive column codes either way). 8 x 32 has 8 fewer fill steps per alignment, 16 x 16 the better steady state
// (fewer live registers): short queries take the former (measured cross-over between q = 1000 and 1500).
static inline LengthClass shapeForQuery(const LengthClass& lc, int qlen) {
    static const int crossover = [] { const char* e = getenv("SW4_CLASS256_CROSSOVER"); return e ? atoi(e) : 1200; }();
    if (lc.capacity == 256 && !lc.wide && qlen < crossover) return LengthClass{3, 32, 256, false, false};
    return lc;
}

static inline int s16Period(int qlen, int G) { return std::max(16, ((qlen + 1) / 2 + G - 1 + 7) / 8 * 8); }
// the wide variant advances one row per step: q + G - 1 rounded up to 8, at least 32
static inline int s16WidePeriod(int qlen, int G) { return std::max(32, (qlen + G - 1 + 7) / 8 * 8); }

template <class Kernel, class Params>
static void launch_clustered(Kernel kernel, const Params& prm, int grid, int smemBytes, cudaStream_t stream);

template <int R, bool MULTI>
static void launch_s16_wide(const S16WideParams& prm, int grid, cudaStream_t stream) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        SW4_CUDA(cudaFuncSetAttribute(sw_s16_wide_kernel<R, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, s16_wide_smem_bytes<R>()));
        configured[dev & 63] = true;
    }
    launch_clustered(sw_s16_wide_kernel<R, MULTI>, prm, grid, s16_wide_smem_bytes<R>(), stream);
}

template <int R>
static void launch_s16(const S16Params& prm, int grid, cudaStream_t stream) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        SW4_CUDA(cudaFuncSetAttribute(sw_s16_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, s16_smem_bytes<R>()));
        configured[dev & 63] = true;
    }
    launch_clustered(sw_s16_kernel<R>, prm, grid, s16_smem_bytes<R>(), stream);
}

template <class Kernel, class Params>
static void launch_clustered(Kernel kernel, const Params& prm, int grid, int smemBytes, cudaStream_t stream) {
    // CTAs are launched as clusters of 2 whenever the grid is even: the two SMs of a TPC share an instruction cache, and
    // with 17 different (large, heavily unrolled) kernels resident at once it pays to give both SMs the same code.
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kS16Threads);
    cfg.dynamicSmemBytes = (size_t)smemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (grid % 2 == 0 && !getenv("SW4_NO_CLUSTER")) ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SW4_CUDA(cudaLaunchKernelEx(&cfg, kernel, prm));
}

struct Engine {
    std::vector<int> deviceIds;
    int numTop = 10;
    int blosum = 62;
    int gop = -11, gex = -1;
    int kernelTypes[4] = {SW4_KERNEL_DPX_S16, SW4_KERNEL_DPX_S16, SW4_KERNEL_DPX_S32, SW4_KERNEL_DPX_S32};
    sw4_mem_config mem{};
    bool verbose = false;
    // scheduling of the length classes (development switches: SW4_SCHED=partition|backfill, SW4_BACKFILL_ITEMS=n)
    bool backfill = [] { const char* e = getenv("SW4_SCHED"); return !(e && std::string(e) == "partition"); }();
    bool useLongKernel = [] { const char* e = getenv("SW4_NO_LONG_KERNEL"); return !e; }();
    int longMinWarps = [] { const char* e = getenv("SW4_LONG_MIN_WARPS"); return e ? std::max(2, atoi(e)) : 2; }();
    int backfillItems = [] { const char* e = getenv("SW4_BACKFILL_ITEMS"); return e ? std::max(1, atoi(e)) : 4; }();
    int shardRank = 0, shardWorld = 1;
    std::unique_ptr<HostDB> db;
    std::vector<std::unique_ptr<Shard>> shards;
    int8_t matrix[441];
    std::string lastError;
    // total timer
    std::chrono::steady_clock::time_point totalStart;
    double totalCells = 0;
    int totalOverflows = 0;

    void buildMatrix() {
        const SubstitutionTriangle* t = nullptr;
        for (const auto& cand : kSubstitutionTriangles)
            if (cand.id == blosum) t = &cand;
        if (!t) fail(SW4_ERR_INVALID, "unsupported substitution matrix blosum%d (45, 50, 62, 80)", blosum);
        for (int r = 0; r < 21; r++)
            for (int c = 0; c < 21; c++) {
                int v = t->low;
                if (r < 20 && c < 20) { const int a = std::max(r, c), b = std::min(r, c); v = t->tri[a * (a + 1) / 2 + b]; }
                matrix[r * 21 + c] = (int8_t)v;
            }
    }

    void checkGaps() const {
        if (gop > 0 || gex > 0 || gop < -4096 || gex < -4096)
            fail(SW4_ERR_INVALID, "gap scores must be in [-4096, 0] (gop=%d gex=%d)", gop, gex);
    }

    // ---- sharding: interleaved blocks of kShardBlock consecutive subjects of the length-sorted database ----
    void assignShards() {
        shards.clear();
        const int perHandle = (int)deviceIds.size();
        const long long totalShards = (long long)perHandle * shardWorld;
        const size_t n = db->n;
        for (int d = 0; d < perHandle; d++) {
            auto sh = std::make_unique<Shard>();
            sh->device = deviceIds[d];
            const long long myShard = (long long)shardRank * perHandle + d;
            const size_t numBlocks = (n + kShardBlock - 1) / kShardBlock;
            for (size_t b = (size_t)myShard; b < numBlocks; b += (size_t)totalShards) {
                const size_t lo = b * kShardBlock, hi = std::min(n, lo + kShardBlock);
                for (size_t i = lo; i < hi; i++) {
                    sh->globalIds.push_back((int32_t)i);
                    sh->residues += (uint64_t)db->lengths[i];
                }
            }
            sh->n = sh->globalIds.size();
            shards.push_back(std::move(sh));
        }
    }

    void initShardDevice(Shard& sh) {
        SW4_CUDA(cudaSetDevice(sh.device));
        cudaDeviceProp prop;
        SW4_CUDA(cudaGetDeviceProperties(&prop, sh.device));
        if (prop.major < 9) fail(SW4_ERR_CUDA, "device %d (%s) has no DPX instructions; sm_100a required", sh.device, prop.name);
        sh.smCount = prop.multiProcessorCount;
        if (!sh.stream) {
            SW4_CUDA(cudaStreamCreateWithFlags(&sh.stream, cudaStreamNonBlocking));
            SW4_CUDA(cudaEventCreate(&sh.evStart));
            SW4_CUDA(cudaEventCreate(&sh.evK0));
            SW4_CUDA(cudaEventCreate(&sh.evK1));
            SW4_CUDA(cudaEventCreate(&sh.evStop));
            SW4_CUDA(cudaEventCreateWithFlags(&sh.evFork, cudaEventDisableTiming));
            for (int i = 0; i < kMaxClassStreams; i++) {
                SW4_CUDA(cudaStreamCreateWithFlags(&sh.classStreams[i], cudaStreamNonBlocking));
                SW4_CUDA(cudaStreamCreateWithFlags(&sh.oddStreams[i], cudaStreamNonBlocking));
                SW4_CUDA(cudaEventCreateWithFlags(&sh.evJoin[i], cudaEventDisableTiming));
                SW4_CUDA(cudaEventCreateWithFlags(&sh.evJoinOdd[i], cudaEventDisableTiming));
            }
        }
    }

    void uploadShard(Shard& sh) {
        initShardDevice(sh);
        const size_t n = sh.n;
        // gather the shard's sequences into one contiguous makedb-style block
        std::vector<size_t> offsets(n + 1, 0);
        std::vector<int32_t> lengths(n);
        for (size_t i = 0; i < n; i++) {
            const size_t g = (size_t)sh.globalIds[i];
            lengths[i] = db->lengths[g];
            offsets[i + 1] = offsets[i] + (((size_t)lengths[i] + 3) / 4) * 4;
        }
        const size_t totalChars = offsets[n];
        size_t freeB = 0, totalB = 0;
        SW4_CUDA(cudaMemGetInfo(&freeB, &totalB));
        const size_t need = totalChars * 3 + n * 40 + ((size_t)512 << 20);
        const size_t limit = std::min(freeB, mem.max_gpu_mem);
        if (need > limit)
            fail(SW4_ERR_NOMEM, "database shard does not fit on device %d: need ~%zu MiB, %zu MiB usable (streaming mode "
                 "is not implemented; use more GPUs)", sh.device, need >> 20, limit >> 20);
        uint8_t* hChars = nullptr;
        SW4_CUDA(cudaMallocHost(&hChars, std::max<size_t>(totalChars, 1)));
        struct PinnedFree { uint8_t* p; ~PinnedFree() { cudaFreeHost(p); } } pinnedFree{hChars};
        {
            const int nt = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
            std::vector<std::thread> th;
            for (int t = 0; t < nt; t++)
                th.emplace_back([&, t] {
                    for (size_t i = (size_t)t; i < n; i += (size_t)nt) {
                        const size_t g = (size_t)sh.globalIds[i];
                        const size_t len = (size_t)lengths[i];
                        uint8_t* dst = hChars + offsets[i];
                        const uint8_t* src = db->chars + db->offsets[g];
                        for (size_t c = 0; c < len; c++) dst[c] = src[c] > 20 ? 20 : src[c];
                        for (size_t c = len; c < offsets[i + 1] - offsets[i]; c++) dst[c] = 20;
                    }
                });
            for (auto& t : th) t.join();
        }
        sh.dChars.alloc(totalChars + 16);
        sh.dOffsets.alloc(n + 1);
        sh.dLengths.alloc(n);
        sh.dGlobalIds.alloc(n);
        sh.dScores.alloc(n);
        sh.dOvfList.alloc(n);
        sh.dCounters.alloc(kNumCounters);
        sh.dClassNs.alloc(96);
        if (!sh.hClassNs) SW4_CUDA(cudaMallocHost(&sh.hClassNs, 96 * sizeof(unsigned long long)));
        sh.dMatrix.alloc(441);
        SW4_CUDA(cudaMemcpyAsync(sh.dChars.p, hChars, totalChars, cudaMemcpyHostToDevice, sh.stream));
        SW4_CUDA(cudaMemcpyAsync(sh.dOffsets.p, offsets.data(), (n + 1) * sizeof(size_t), cudaMemcpyHostToDevice, sh.stream));
        SW4_CUDA(cudaMemcpyAsync(sh.dLengths.p, lengths.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, sh.stream));
        SW4_CUDA(cudaMemcpyAsync(sh.dGlobalIds.p, sh.globalIds.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, sh.stream));
        SW4_CUDA(cudaMemsetAsync(sh.dScores.p, 0, std::max<size_t>(n, 1) * sizeof(int32_t), sh.stream));

        // length classes (the shard is ascending in length): consecutive subjects are paired into work items
        sh.classes.clear();
        size_t pos = std::upper_bound(lengths.begin(), lengths.end(), 0) - lengths.begin();  // length-0 subjects score 0
        sh.numZeroLength = pos;
        std::vector<S16Item> items;
        std::vector<int32_t> blockItem;
        for (int c = 0; c < kNumLengthClasses; c++) {
            const LengthClass& lc = kLengthClasses[c];
            const size_t end = lc.multi ? n : std::upper_bound(lengths.begin(), lengths.end(), lc.capacity) - lengths.begin();
            if (end > pos) {
                auto cl = std::make_unique<ClassLayout>();
                cl->cls = c;
                cl->first = (int)pos;
                cl->count = (int)(end - pos);
                items.clear();
                blockItem.clear();
                for (size_t i = pos; i < end; i += 2) {
                    S16Item it;
                    it.subject0 = (int)i;
                    it.subject1 = (i + 1 < end) ? (int)(i + 1) : -1;
                    const int maxLen = std::max(lengths[i], (i + 1 < end) ? lengths[i + 1] : 0);
                    it.numSegments = lc.multi ? (maxLen + lc.capacity - 1) / lc.capacity : 1;
                    it.firstBlock = 0;
                    items.push_back(it);
                }
                if (lc.multi) std::reverse(items.begin(), items.end());  // longest first: they bound the makespan
                int blk = 0;
                for (size_t k = 0; k < items.size(); k++) {
                    items[k].firstBlock = blk;
                    for (int sgm = 0; sgm < items[k].numSegments; sgm++) blockItem.push_back((int32_t)k);
                    blk += items[k].numSegments;
                }
                cl->numItems = (int)items.size();
                cl->numBlocks = blk;
                const int columns = lc.capacity;
                cl->cols.alloc((size_t)cl->numBlocks * columns);
                cl->items.alloc(items.size());
                DevBuf<int32_t> dBlockItem;
                dBlockItem.alloc(blockItem.size());
                SW4_CUDA(cudaMemcpyAsync(cl->items.p, items.data(), items.size() * sizeof(S16Item), cudaMemcpyHostToDevice, sh.stream));
                SW4_CUDA(cudaMemcpyAsync(dBlockItem.p, blockItem.data(), blockItem.size() * sizeof(int32_t), cudaMemcpyHostToDevice, sh.stream));
                const long long total = (long long)cl->numBlocks * columns;
                build_pair_blocks_kernel<<<(unsigned)((total + 255) / 256), 256, 0, sh.stream>>>(
                    sh.dChars.p, sh.dOffsets.p, sh.dLengths.p, reinterpret_cast<const PairItem*>(cl->items.p), dBlockItem.p,
                    cl->numBlocks, columns, cl->cols.p);
                SW4_CUDA(cudaGetLastError());
                SW4_CUDA(cudaStreamSynchronize(sh.stream));  // dBlockItem and the host vectors go out of scope
                sh.classes.push_back(std::move(cl));
            }
            pos = std::max(pos, end);
        }
        SW4_CUDA(cudaStreamSynchronize(sh.stream));
        sh.uploaded = true;
        if (verbose)
            fprintf(stderr, "[sw4] device %d: %zu subjects, %llu residues, %zu length classes\n", sh.device, n,
                    (unsigned long long)sh.residues, sh.classes.size());
    }

    void upload() {
        if (!db) fail(SW4_ERR_INVALID, "no database set");
        bool fresh = false;
        for (auto& sh : shards)
            if (!sh->uploaded) { uploadShard(*sh); fresh = true; }
        if (fresh) {  // untimed warm-up scan: loads every kernel this shard will launch and sizes the scratch buffers
            const int k = (int)std::min<size_t>((size_t)std::max(numTop, 1), std::max<size_t>(db->n, 1));
            std::string warm;
            for (int i = 0; i < 320; i++) warm.push_back("ARNDCQEGHILKMFPSTWYV"[(i * 7) % 20]);
            for (int rep = 0; rep < 2; rep++) {  // twice: the second pass runs with a measured SM partition
                enqueueAll(warm.data(), (int)warm.size(), std::min(k, kTopkMaxCandidates / 2));
                for (auto& sh : shards) { SW4_CUDA(cudaSetDevice(sh->device)); SW4_CUDA(cudaStreamSynchronize(sh->stream)); updateClassRates(*sh); }
            }
            for (auto& sh : shards) { SW4_CUDA(cudaSetDevice(sh->device)); SW4_CUDA(cudaStreamSynchronize(sh->stream)); updateClassRates(*sh); }
        }
    }

    // ---- one scan on one shard: everything is enqueued on sh.stream ----
    int statThreshold() const { return kernelTypes[0] == SW4_KERNEL_HALF2 ? kHalf2Threshold : kS16OverflowThreshold; }

    void enqueueScan(Shard& sh, const char* query, int qlen, int k) {
        SW4_CUDA(cudaSetDevice(sh.device));
        sh.launches = 0;
        const int qpad = (qlen + 3) / 4 * 4;
        if ((size_t)qlen > sh.hQueryCap) {
            if (sh.hQuery) cudaFreeHost(sh.hQuery);
            sh.hQuery = nullptr;
            sh.hQueryCap = std::max<size_t>((size_t)qlen * 2, 65536);
            SW4_CUDA(cudaMallocHost(&sh.hQuery, sh.hQueryCap));
        }
        const size_t topWords = (size_t)2 * k + 8;
        if (topWords > sh.hTopCap) {
            if (sh.hTop) cudaFreeHost(sh.hTop);
            sh.hTop = nullptr;
            sh.hTopCap = topWords + 64;
            SW4_CUDA(cudaMallocHost(&sh.hTop, sh.hTopCap * sizeof(int32_t)));
        }
        // rows of the positional profile start on 128-byte lines: the ring refill copies 64 contiguous bytes per row, which
        // then always fall into two sectors of one line (a stride of 4 mod 8 words cost 3-4 % on the peak benchmark)
        const int profStride = (qlen + 64 + 31) / 32 * 32;
        // All device scratch is sized here, BEFORE the timed region, for a query capacity that only grows by doubling:
        // cudaMalloc/cudaFree inside the event-bracketed region stall the stream for up to hundreds of milliseconds.
        if (qlen > sh.queryCapacity || k > sh.topCapacity) {
            SW4_CUDA(cudaStreamSynchronize(sh.stream));
            int cap = std::max(sh.queryCapacity, 8192);
            while (cap < qlen) cap *= 2;
            sh.queryCapacity = cap;
            sh.topCapacity = std::max(sh.topCapacity, std::max(k, 64));
            const size_t capStride = (size_t)(cap + 64 + 31) / 32 * 32;
            sh.dQueryLetters.ensure((size_t)cap + 16);
            sh.dQueryCodes.ensure((size_t)cap + 16);
            sh.dProfile.ensure((size_t)kProfileRows * capStride);
            sh.dTopScores.ensure(sh.topCapacity);
            sh.dTopIds.ensure(sh.topCapacity);
            sh.dCand.ensure((size_t)kTopkMaxCandidates);
            const size_t capBorderStride = (size_t)(cap + 31) / 32 * 32 + 32;
            sh.borderWarps = (size_t)sh.smCount * 4 * kS32WarpsPerBlock;
            while (sh.borderWarps * capBorderStride * sizeof(int2) > mem.max_temp_bytes && sh.borderWarps > kS32WarpsPerBlock)
                sh.borderWarps = (sh.borderWarps / 2 + kS32WarpsPerBlock - 1) / kS32WarpsPerBlock * kS32WarpsPerBlock;
            sh.dBorder.ensure(sh.borderWarps * capBorderStride);
            bool anyMulti = false;
            for (auto& cl : sh.classes) anyMulti |= kLengthClasses[cl->cls].multi;
            sh.borderWideStride = capBorderStride;  // rows
            sh.dBorderWide.ensure(anyMulti ? (size_t)sh.smCount * kS16Warps * capBorderStride : 1);
            sh.dBorderLong.ensure(anyMulti ? (size_t)sh.smCount * 8 * capBorderStride : 1);
        }
        memcpy(sh.hQuery, query, (size_t)qlen);

        cudaStream_t st = sh.stream;
        SW4_CUDA(cudaEventRecord(sh.evStart, st));
        SW4_CUDA(cudaMemcpyAsync(sh.dQueryLetters.p, sh.hQuery, (size_t)qlen, cudaMemcpyHostToDevice, st));
        SW4_CUDA(cudaMemcpyAsync(sh.dMatrix.p, matrix, 441, cudaMemcpyHostToDevice, st));
        SW4_CUDA(cudaMemsetAsync(sh.dCounters.p, 0, kNumCounters * sizeof(int), st));
        SW4_CUDA(cudaMemsetAsync(sh.dClassNs.p, 0, 96 * sizeof(unsigned long long), st));
        // every scan starts from "-1 = not scored" so that a subject the kernels missed can never keep an old score;
        // empty subjects (and everything, for an empty query) score 0 by definition
        SW4_CUDA(cudaMemsetAsync(sh.dScores.p, qlen == 0 ? 0 : 0xff, std::max<size_t>(sh.n, 1) * sizeof(int32_t), st));
        if (qlen > 0 && sh.numZeroLength) SW4_CUDA(cudaMemsetAsync(sh.dScores.p, 0, sh.numZeroLength * sizeof(int32_t), st));
        if (qpad > 0) convert_query_kernel<<<(qpad + 255) / 256, 256, 0, st>>>(sh.dQueryLetters.p, sh.dQueryCodes.p, qlen, qpad);
        build_profile_kernel<<<dim3((profStride + 127) / 128, kProfileRows), 128, 0, st>>>(sh.dQueryCodes.p, qlen, sh.dMatrix.p,
                                                                                   sh.dProfile.p, profStride);
        SW4_CUDA(cudaGetLastError());
        sh.launches += 2;
        SW4_CUDA(cudaEventRecord(sh.evK0, st));

        // packed 16-bit classes, longest first
        const uint32_t gop2 = ((uint32_t)(uint16_t)(int16_t)gop << 16) | (uint16_t)(int16_t)gop;
        const uint32_t gex2 = ((uint32_t)(uint16_t)(int16_t)gex << 16) | (uint16_t)(int16_t)gex;
        // All length classes run CONCURRENTLY, each on its own stream with its own share of the SMs (one persistent CTA
        // per SM): the shares are sized so that every class finishes at about the same time, instead of 17 launches that
        // each under-fill the GPU one after the other (the reference: 36 launches over 10 streams, src/cudasw4.cuh:1745).
        const int numClasses = (int)sh.classes.size();
        std::vector<int> grid(numClasses, 0), cap(numClasses, 0);
        std::vector<double> cost(numClasses, 0.0);
        if (qlen > 0 && numClasses > 0) {
            int used = 0;
            for (int ci = 0; ci < numClasses; ci++) {
                const ClassLayout& cl = *sh.classes[ci];
                const LengthClass lc = shapeForQuery(kLengthClasses[cl.cls], qlen);
                const int G = 1 << lc.logG;
                const int groupsPerCta = kS16Warps * (32 >> lc.logG);
                const int period = lc.wide ? s16WidePeriod(qlen, G) : s16Period(qlen, G);
                // a class may spread out to as few as 4 busy warps per SM (one per scheduler) when there are SMs to spare
                cap[ci] = std::max(1, (cl.numItems + 3) / 4);
                cost[ci] = (double)cl.numBlocks / groupsPerCta * period * (lc.wide ? lc.R * 7.3 + 30.0 : lc.R * 14.6 + 40.0) * cl.rate;
                grid[ci] = 1;
                used++;
            }
            if (backfill) {
                // Back-fill scheduling: every class is launched at (up to) full width, longest subjects first, on its own
                // stream; one CTA fits per SM, so a later class's CTAs start as earlier CTAs retire and the dynamic
                // tickets even out the rest. No SM goes idle before the last class runs dry.
                const int width = sh.smCount >= 2 ? (sh.smCount & ~1) : 1;
                for (int ci = 0; ci < numClasses; ci++) {
                    const ClassLayout& cl = *sh.classes[ci];
                    const LengthClass lc = shapeForQuery(kLengthClasses[cl.cls], qlen);
                    const int groupsPerCta = kS16Warps * (32 >> lc.logG);
                    int g = (cl.numItems + groupsPerCta * backfillItems - 1) / (groupsPerCta * backfillItems);
                    g = std::max(1, std::min(width, g));
                    if (g > 1) g = (g + 1) & ~1;
                    grid[ci] = std::min(g, width);
                }
                used = sh.smCount;
            }
            while (used < sh.smCount) {  // next SM goes to the class with the largest remaining load per SM
                int best = -1;
                double bestLoad = 0;
                for (int ci = 0; ci < numClasses; ci++) {
                    if (grid[ci] >= cap[ci]) continue;
                    const double load = cost[ci] / grid[ci];
                    if (load > bestLoad) { bestLoad = load; best = ci; }
                }
                if (best < 0) break;
                grid[best]++;
                used++;
            }
            SW4_CUDA(cudaEventRecord(sh.evFork, st));
        }
        for (int ci = numClasses - 1; ci >= 0 && qlen > 0; ci--) {
            ClassLayout& cl = *sh.classes[ci];
            const LengthClass lc = shapeForQuery(kLengthClasses[cl.cls], qlen);
            const int G = 1 << lc.logG;
            cudaStream_t cst = sh.classStreams[ci % kMaxClassStreams];
            SW4_CUDA(cudaStreamWaitEvent(cst, sh.evFork, 0));
            cl.lastCost = cost[ci] / cl.rate;
            cl.lastGrid = grid[ci];
            const int groupsPerCta = kS16Warps * (32 >> lc.logG);
            const int activeGroups = std::min(groupsPerCta, std::max(1, (cl.numItems + grid[ci] - 1) / grid[ci]));
            auto fillCommon = [&](auto& prm) {
                prm.cols = cl.cols.p;
                prm.items = cl.items.p;
                prm.numItems = cl.numItems;
                prm.ticket = sh.dCounters.p + 8 + cl.cls;
                prm.logG = lc.logG;
                prm.profile = sh.dProfile.p;
                prm.profStride = profStride;
                prm.qlen = qlen;
                prm.gop2 = gop2;
                prm.gex2 = gex2;
                prm.ovfThreshold = kS16OverflowThreshold;
                prm.statThreshold = (lc.capacity > 240) ? statThreshold() : 0x7fffffff;
                prm.scores = sh.dScores.p;
                prm.ovfList = sh.dOvfList.p;
                prm.ovfCount = sh.dCounters.p + 0;
                prm.statCount = sh.dCounters.p + 1;
                prm.elapsedNs = sh.dClassNs.p + cl.cls;
                prm.activeGroups = activeGroups;
            };
            S16Params narrow{};
            S16WideParams wide{};
            if (lc.wide) {
                fillCommon(wide);
                wide.period = s16WidePeriod(qlen, G);
                wide.lengths = sh.dLengths.p;
                wide.border = sh.dBorderWide.p;
                wide.borderStride = (int)sh.borderWideStride;
            } else {
                fillCommon(narrow);
                narrow.period = s16Period(qlen, G);
            }
            auto launchClass = [&](int g, cudaStream_t strm, int ctaOffset) {
                if (lc.wide) {
                    wide.ctaOffset = ctaOffset;
                    if (lc.multi) {
                        launch_s16_wide<32, true>(wide, g, strm);
                    } else {
                        switch (lc.R) {
                            case 18: launch_s16_wide<18, false>(wide, g, strm); break;
                            case 20: launch_s16_wide<20, false>(wide, g, strm); break;
                            case 22: launch_s16_wide<22, false>(wide, g, strm); break;
                            case 24: launch_s16_wide<24, false>(wide, g, strm); break;
                            case 26: launch_s16_wide<26, false>(wide, g, strm); break;
                            case 28: launch_s16_wide<28, false>(wide, g, strm); break;
                            case 30: launch_s16_wide<30, false>(wide, g, strm); break;
                            case 32: launch_s16_wide<32, false>(wide, g, strm); break;
                            default: fail(SW4_ERR_INVALID, "no wide kernel for R=%d", lc.R);
                        }
                    }
                } else {
                    narrow.ctaOffset = ctaOffset;
                    switch (lc.R) {
                        case 4: launch_s16<4>(narrow, g, strm); break;
                        case 6: launch_s16<6>(narrow, g, strm); break;
                        case 8: launch_s16<8>(narrow, g, strm); break;
                        case 10: launch_s16<10>(narrow, g, strm); break;
                        case 12: launch_s16<12>(narrow, g, strm); break;
                        case 14: launch_s16<14>(narrow, g, strm); break;
                        case 16: launch_s16<16>(narrow, g, strm); break;
                        case 18: launch_s16<18>(narrow, g, strm); break;
                        case 20: launch_s16<20>(narrow, g, strm); break;
                        case 22: launch_s16<22>(narrow, g, strm); break;
                        case 24: launch_s16<24>(narrow, g, strm); break;
                        case 26: launch_s16<26>(narrow, g, strm); break;
                        case 28: launch_s16<28>(narrow, g, strm); break;
                        case 30: launch_s16<30>(narrow, g, strm); break;
                        case 32: launch_s16<32>(narrow, g, strm); break;
                        default: fail(SW4_ERR_INVALID, "no kernel for R=%d", lc.R);
                    }
                }
                sh.launches++;
            };
            // the multi-segment class runs on the CTA-wide array kernel whenever the query is long enough to keep
            // at least two warps of an array busy (kernels_s16_long.cuh)
            int longWarps = 0;
            if (lc.multi && useLongKernel) {
                const int p0 = (qlen + 32 + 15) / 16 * 16;
                longWarps = kLongMaxWarps;
                while (longWarps > 1 && kLongLag * longWarps + 64 > p0) longWarps >>= 1;
                if (longWarps < longMinWarps) longWarps = 0;
                if (longWarps) {
                    S16LongParams lp{};
                    lp.cols = cl.cols.p;
                    lp.items = cl.items.p;
                    lp.lengths = sh.dLengths.p;
                    lp.numItems = cl.numItems;
                    lp.ticket = sh.dCounters.p + 8 + cl.cls;
                    lp.warps = longWarps;
                    lp.ringSlots = s16_long_ring_slots(longWarps);
                    lp.profLo = sh.dProfile.p + (size_t)kFused * profStride;
                    lp.profHi = sh.dProfile.p + (size_t)(kFused + 21) * profStride;
                    lp.profStride = profStride;
                    lp.qlen = qlen;
                    lp.period = p0;
                    lp.gop2 = gop2;
                    lp.gex2 = gex2;
                    lp.ovfThreshold = kS16OverflowThreshold;
                    lp.statThreshold = statThreshold();
                    lp.scores = sh.dScores.p;
                    lp.ovfList = sh.dOvfList.p;
                    lp.ovfCount = sh.dCounters.p + 0;
                    lp.statCount = sh.dCounters.p + 1;
                    lp.elapsedNs = sh.dClassNs.p + cl.cls;
                    lp.border = sh.dBorderLong.p;
                    lp.borderStride = (int)sh.borderWideStride;
                    const int smemBytes = s16_long_smem_bytes(longWarps);
                    static bool configured[64] = {};
                    if (!configured[sh.device & 63]) {
                        SW4_CUDA(cudaFuncSetAttribute(sw_s16_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      s16_long_smem_bytes(kLongMaxWarps)));
                        configured[sh.device & 63] = true;
                    }
                    const int ctasPerSm = std::max(1, std::min(kLongMaxWarps / longWarps, (227 * 1024) / (smemBytes + 1024)));
                    const int g = std::max(1, std::min(cl.numItems, std::min(sh.smCount * ctasPerSm, sh.smCount * 8)));
                    cl.lastGrid = g;
                    sw_s16_long_kernel<<<g, longWarps * 32, smemBytes, cst>>>(lp);
                    SW4_CUDA(cudaGetLastError());
                    sh.launches++;
                    SW4_CUDA(cudaEventRecord(sh.evJoin[ci % kMaxClassStreams], cst));
                    SW4_CUDA(cudaStreamWaitEvent(st, sh.evJoin[ci % kMaxClassStreams], 0));
                    continue;
                }
            }
            // even part as clusters of 2 (same code on both SMs of a TPC), an odd leftover CTA on a second stream;
            // both launches share the class's ticket counter and timing slot
            const int evenPart = grid[ci] & ~1;
            if (evenPart > 0) launchClass(evenPart, cst, 0);
            SW4_CUDA(cudaEventRecord(sh.evJoin[ci % kMaxClassStreams], cst));
            SW4_CUDA(cudaStreamWaitEvent(st, sh.evJoin[ci % kMaxClassStreams], 0));
            if (grid[ci] & 1) {
                cudaStream_t ost = sh.oddStreams[ci % kMaxClassStreams];
                SW4_CUDA(cudaStreamWaitEvent(ost, sh.evFork, 0));
                launchClass(1, ost, evenPart);
                SW4_CUDA(cudaEventRecord(sh.evJoinOdd[ci % kMaxClassStreams], ost));
                SW4_CUDA(cudaStreamWaitEvent(st, sh.evJoinOdd[ci % kMaxClassStreams], 0));
            }
        }

        // exact 32-bit path: long subjects, then whatever saturated in 16 bit
        const int borderStride = (qlen + 31) / 32 * 32 + 32;
        auto launchS32 = [&](const int32_t* list, const int* countPtr, int countHost, int* ticket, bool countStats) {
            int blocks = (int)(sh.borderWarps / kS32WarpsPerBlock);
            if (!countPtr) blocks = std::max(1, std::min(blocks, (countHost + kS32WarpsPerBlock - 1) / kS32WarpsPerBlock));
            S32Params p{};
            p.chars = sh.dChars.p; p.offsets = sh.dOffsets.p; p.lengths = sh.dLengths.p;
            p.list = list; p.listCountPtr = countPtr; p.listCountHost = countHost;
            p.query = sh.dQueryCodes.p; p.qlen = qlen; p.matrix = sh.dMatrix.p; p.gop = gop; p.gex = gex;
            p.border = sh.dBorder.p; p.borderStride = borderStride; p.ticket = ticket; p.scores = sh.dScores.p;
            p.statThreshold = countStats ? statThreshold() : 0x7fffffff;
            p.statCount = sh.dCounters.p + 1;
            sw_s32_kernel<<<blocks, kS32Threads, 0, st>>>(p);
            SW4_CUDA(cudaGetLastError());
            sh.launches++;
        };
        if (qlen > 0 && !sh.classes.empty()) {
            // long queries (the only ones that can saturate 16 bits) re-score on the CTA-wide array, 16 warps per subject
            const int p0 = (qlen + 32 + 15) / 16 * 16;
            if (useLongKernel && kLongLag * kLongMaxWarps + 64 <= p0) {
                S32LongParams lp{};
                lp.chars = sh.dChars.p; lp.offsets = sh.dOffsets.p; lp.lengths = sh.dLengths.p;
                lp.list = sh.dOvfList.p; lp.listCountPtr = sh.dCounters.p + 0; lp.ticket = sh.dCounters.p + 3;
                lp.warps = kLongMaxWarps;
                lp.ringSlots = s16_long_ring_slots(kLongMaxWarps);
                lp.prof = reinterpret_cast<const int32_t*>(sh.dProfile.p + (size_t)(kFused + 42) * profStride);
                lp.profStride = profStride; lp.qlen = qlen; lp.period = p0; lp.gop = gop; lp.gex = gex;
                lp.scores = sh.dScores.p;
                lp.border = reinterpret_cast<int2*>(sh.dBorderLong.p);
                lp.borderStride = (int)sh.borderWideStride;
                const int smemBytes = s32_long_smem_bytes(kLongMaxWarps);
                static bool configured[64] = {};
                if (!configured[sh.device & 63]) {
                    SW4_CUDA(cudaFuncSetAttribute(sw_s32_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
                    configured[sh.device & 63] = true;
                }
                sw_s32_long_kernel<<<sh.smCount, kLongMaxWarps * 32, smemBytes, st>>>(lp);
                SW4_CUDA(cudaGetLastError());
                sh.launches++;
            } else {
                launchS32(sh.dOvfList.p, sh.dCounters.p + 0, 0, sh.dCounters.p + 3, false);
            }
        }
        SW4_CUDA(cudaEventRecord(sh.evK1, st));

        // top-k
        const long long n = (long long)sh.n;
        if (k > 0 && n > 0) {
            int blocks = (int)std::min<long long>(std::min<long long>(sh.smCount, kTopkMaxCandidates / k), (n + 4095) / 4096);
            blocks = std::max(blocks, 1);
            topk_pass1_kernel<<<blocks, kTopkThreads, 0, st>>>(sh.dScores.p, nullptr, n, k, sh.dCand.p);
            const int numCand = blocks * k;
            int n2 = 1;
            while (n2 < numCand) n2 <<= 1;
            static bool pass2Configured[64] = {};
            if (!pass2Configured[sh.device & 63]) {
                SW4_CUDA(cudaFuncSetAttribute(topk_pass2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTopkMaxCandidates * 8));
                pass2Configured[sh.device & 63] = true;
            }
            topk_pass2_kernel<<<1, kTopkThreads, (size_t)n2 * 8, st>>>(sh.dCand.p, numCand, k, sh.dGlobalIds.p, sh.dTopScores.p,
                                                                       sh.dTopIds.p, sh.dCounters.p + 4);
            SW4_CUDA(cudaGetLastError());
            sh.launches += 2;
            SW4_CUDA(cudaMemcpyAsync(sh.hTop, sh.dTopScores.p, (size_t)k * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            SW4_CUDA(cudaMemcpyAsync(sh.hTop + k, sh.dTopIds.p, (size_t)k * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        }
        SW4_CUDA(cudaMemcpyAsync(sh.hTop + 2 * k, sh.dCounters.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
        SW4_CUDA(cudaMemcpyAsync(sh.hClassNs, sh.dClassNs.p, 96 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        SW4_CUDA(cudaEventRecord(sh.evStop, st));
    }

    // One host thread per GPU issues that GPU's launches (the reference drives all GPUs from a single thread in lock-step
    // phases, src/cudasw4.cuh:1509-2259: ~35 launches x 8 GPUs back to back before the last GPU starts).
    void enqueueAll(const char* query, int qlen, int k) {
        if (shards.size() == 1) { enqueueScan(*shards[0], query, qlen, k); return; }
        std::vector<std::thread> workers;
        std::vector<Error> errors(shards.size(), Error{SW4_OK, ""});
        for (size_t i = 1; i < shards.size(); i++)
            workers.emplace_back([&, i] {
                try { enqueueScan(*shards[i], query, qlen, k); }
                catch (const Error& e) { errors[i] = e; }
                catch (const std::exception& e) { errors[i] = Error{SW4_ERR_INVALID, e.what()}; }
            });
        try { enqueueScan(*shards[0], query, qlen, k); }
        catch (const Error& e) { errors[0] = e; }
        for (auto& w : workers) w.join();
        for (auto& e : errors)
            if (e.code != SW4_OK) throw e;
    }

    // feedback for the SM partition: how many SM-milliseconds a unit of modelled cost really took in the last scan
    static void updateClassRates(Shard& sh) {
        double norm = 0;
        int cnt = 0;
        const bool debug = getenv("SW4_DEBUG_PARTITION") != nullptr;
        unsigned long long t0 = ~0ull;
        if (debug && sh.hClassNs)
            for (auto& clp : sh.classes)
                if (sh.hClassNs[32 + clp->cls]) t0 = std::min(t0, ~sh.hClassNs[32 + clp->cls]);
        for (auto& clp : sh.classes) {
            ClassLayout& cl = *clp;
            if (cl.lastGrid == 0 || cl.lastCost <= 0 || !sh.hClassNs) continue;
            const double ms = (double)sh.hClassNs[cl.cls] * 1e-6;
            if (ms <= 0) continue;
            const double r = (double)ms * cl.lastGrid / cl.lastCost;
            if (debug)
                fprintf(stderr, "[sw4] class %2d (G=%2d R=%2d%s) items %7d blocks %7d grid %3d  longest CTA %.3f ms  first start %.3f last end %.3f ms  rate %.4g -> %.4g\n",
                        cl.cls, 1 << kLengthClasses[cl.cls].logG, kLengthClasses[cl.cls].R, kLengthClasses[cl.cls].multi ? " multi" : "",
                        cl.numItems, cl.numBlocks, cl.lastGrid, ms, (double)(~sh.hClassNs[32 + cl.cls] - t0) * 1e-6,
                        (double)(sh.hClassNs[64 + cl.cls] - t0) * 1e-6, cl.rate, r);
            cl.rate = (cl.rate == 1.0) ? r : 0.5 * cl.rate + 0.5 * r;
            cl.lastGrid = 0;
            norm += cl.rate;
            cnt++;
        }
        (void)norm; (void)cnt;
    }

    void scan(const char* query, int qlen, int32_t* outScores, int32_t* outIds, int32_t* outCount, sw4_stats* stats) {
        if (!db) fail(SW4_ERR_INVALID, "no database set");
        if (qlen < 0 || (qlen > 0 && !query)) fail(SW4_ERR_INVALID, "invalid query");
        if (qlen > (1 << 24)) fail(SW4_ERR_INVALID, "query too long (%d)", qlen);
        checkGaps();
        upload();
        const size_t nTotal = db->n;
        const int k = (int)std::min<size_t>((size_t)std::max(numTop, 0), nTotal);
        // Result lists longer than the device selection keeps (4096) are rare (the reference's CLI default is 10): they take
        // the slow but exact route of copying all scores to the host (the reference's CUDASW_DEBUG_CHECK_CORRECTNESS mode).
        const bool hostSelect = k > kTopkMaxCandidates / 2;
        enqueueAll(query, qlen, hostSelect ? 0 : k);
        double seconds = 0, kernelSeconds = 0;
        int overflows = 0, launches = 0;
        struct Entry { int32_t score, id; };
        std::vector<Entry> merged;
        for (auto& shp : shards) {
            Shard& sh = *shp;
            SW4_CUDA(cudaSetDevice(sh.device));
            SW4_CUDA(cudaStreamSynchronize(sh.stream));
            float ms = 0, kms = 0;
            updateClassRates(sh);
            SW4_CUDA(cudaEventElapsedTime(&ms, sh.evStart, sh.evStop));
            SW4_CUDA(cudaEventElapsedTime(&kms, sh.evK0, sh.evK1));
            seconds = std::max(seconds, (double)ms * 1e-3);
            kernelSeconds = std::max(kernelSeconds, (double)kms * 1e-3);
            const int kDev = hostSelect ? 0 : k;
            const int* counters = sh.hTop + 2 * kDev;
            overflows += counters[1];
            launches += sh.launches;
            if (hostSelect) {
                std::vector<int32_t> all(sh.n);
                SW4_CUDA(cudaMemcpy(all.data(), sh.dScores.p, sh.n * sizeof(int32_t), cudaMemcpyDeviceToHost));
                for (size_t i = 0; i < sh.n; i++) merged.push_back(Entry{all[i], sh.globalIds[i]});
            } else {
                const int cnt = (k > 0 && sh.n > 0) ? counters[4] : 0;
                for (int i = 0; i < cnt; i++) merged.push_back(Entry{sh.hTop[i], sh.hTop[k + i]});
            }
        }
        const int got = (int)std::min<size_t>((size_t)k, merged.size());
        std::partial_sort(merged.begin(), merged.begin() + got, merged.end(), [](const Entry& a, const Entry& b) {
            if (a.score != b.score) return a.score > b.score;
            return a.id < b.id;
        });
        for (int i = 0; i < got; i++) { outScores[i] = merged[i].score; outIds[i] = merged[i].id; }
        if (outCount) *outCount = got;
        uint64_t residues = 0;
        for (auto& sh : shards) residues += sh->residues;
        const double cells = (double)residues * (double)qlen;
        totalCells += cells;
        totalOverflows += overflows;
        if (stats) {
            stats->num_overflows = overflows;
            stats->seconds = seconds;
            stats->gcups = seconds > 0 ? cells / 1e9 / seconds : 0;
            stats->kernel_seconds = kernelSeconds;
            stats->cells = cells;
            stats->kernel_launches = launches;
        }
    }
};

}  // namespace sw4

// =================================================================================================================
// C ABI
// =================================================================================================================
struct sw4_handle {
    sw4::Engine eng;
};

static std::string g_globalError;

template <class Fn>
static int guarded(sw4_handle* h, Fn&& fn) {
    try {
        fn();
        return SW4_OK;
    } catch (const sw4::Error& e) {
        (h ? h->eng.lastError : g_globalError) = e.msg;
        return e.code;
    } catch (const std::bad_alloc&) {
        (h ? h->eng.lastError : g_globalError) = "out of host memory";
        return SW4_ERR_NOMEM;
    } catch (const std::exception& e) {
        (h ? h->eng.lastError : g_globalError) = e.what();
        return SW4_ERR_INVALID;
    }
}

extern "C" {

const char* sw4_version(void) { return "sw4b200 0.1 sm_100a"; }

const char* sw4_last_error(const sw4_handle* h) { return h ? h->eng.lastError.c_str() : g_globalError.c_str(); }

int sw4_create(const int* device_ids, int num_devices, int num_top, int blosum, int gop, int gex, const sw4_mem_config* mem,
               int verbose, sw4_handle** out) {
    if (!out) return SW4_ERR_INVALID;
    *out = nullptr;
    sw4_handle* h = nullptr;
    int rc = guarded(nullptr, [&] {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            sw4::fail(SW4_ERR_CUDA, "no usable CUDA device (%s); this engine has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        h = new sw4_handle();
        sw4::Engine& eng = h->eng;
        if (device_ids && num_devices > 0) {
            for (int i = 0; i < num_devices; i++) {
                if (device_ids[i] < 0 || device_ids[i] >= count) sw4::fail(SW4_ERR_INVALID, "invalid device id %d", device_ids[i]);
                eng.deviceIds.push_back(device_ids[i]);
            }
        } else {
            for (int i = 0; i < count; i++) eng.deviceIds.push_back(i);  // all visible GPUs, like src/main.cu:110-128
        }
        eng.numTop = num_top;
        eng.blosum = blosum;
        eng.gop = gop > 0 ? -gop : gop;
        eng.gex = gex > 0 ? -gex : gex;
        eng.verbose = verbose != 0;
        eng.mem.max_batch_bytes = 128ull << 20;
        eng.mem.max_batch_sequences = 10000000;
        eng.mem.max_temp_bytes = 4ull << 30;
        eng.mem.max_gpu_mem = SIZE_MAX;
        if (mem) eng.mem = *mem;
        if (eng.mem.max_temp_bytes == 0) eng.mem.max_temp_bytes = 4ull << 30;
        eng.buildMatrix();
        eng.checkGaps();
        if (num_top < 0) sw4::fail(SW4_ERR_INVALID, "num_top must be >= 0");
    });
    if (rc != SW4_OK) { delete h; return rc; }
    *out = h;
    return SW4_OK;
}

int sw4_destroy(sw4_handle* h) {
    delete h;
    return SW4_OK;
}

int sw4_set_gap_scores(sw4_handle* h, int gop, int gex) {
    if (!h) return SW4_ERR_INVALID;
    return guarded(h, [&] {
        const int a = gop > 0 ? -gop : gop, b = gex > 0 ? -gex : gex;  // src/cudasw4.cuh:539-550 negates positives
        const int oa = h->eng.gop, ob = h->eng.gex;
        h->eng.gop = a; h->eng.gex = b;
        try { h->eng.checkGaps(); } catch (...) { h->eng.gop = oa; h->eng.gex = ob; throw; }
    });
}

int sw4_set_num_top(sw4_handle* h, int num_top) {
    if (!h) return SW4_ERR_INVALID;
    return guarded(h, [&] {
        if (num_top < 0) sw4::fail(SW4_ERR_INVALID, "num_top must be >= 0");
        h->eng.numTop = num_top;
    });
}

int sw4_set_blosum(sw4_handle* h, int blosum) {
    if (!h) return SW4_ERR_INVALID;
    return guarded(h, [&] {
        const int old = h->eng.blosum;
        h->eng.blosum = blosum;
        try { h->eng.buildMatrix(); } catch (...) { h->eng.blosum = old; h->eng.buildMatrix(); throw; }
    });
}

int sw4_set_kernel_types(sw4_handle* h, int single_pass, int many_pass_small, int many_pass_large, int overflow) {
    if (!h) return SW4_ERR_INVALID;
    return guarded(h, [&] {
        const int v[4] = {single_pass, many_pass_small, many_pass_large, overflow};
        for (int x : v)
            if (x < 0 || x > 3) sw4::fail(SW4_ERR_INVALID, "invalid kernel type %d", x);
        // validity rules of src/cudasw4.cuh:841-855
        if (many_pass_large != SW4_KERNEL_FLOAT && many_pass_large != SW4_KERNEL_DPX_S32)
            sw4::fail(SW4_ERR_INVALID, "many_pass_large must be Float or DPXs32");
        if (overflow != SW4_KERNEL_FLOAT && overflow != SW4_KERNEL_DPX_S32)
            sw4::fail(SW4_ERR_INVALID, "overflow type must be Float or DPXs32");
        for (int i = 0; i < 4; i++) h->eng.kernelTypes[i] = v[i];
    });
}

int sw4_set_shard(sw4_handle* h, int rank, int world) {
    if (!h) return SW4_ERR_INVALID;
    return guarded(h, [&] {
        if (world < 1 || rank < 0 || rank >= world) sw4::fail(SW4_ERR_INVALID, "invalid shard %d of %d", rank, world);
        if (h->eng.db) sw4::fail(SW4_ERR_INVALID, "sw4_set_shard must be called before a database is set");
        h->eng.shardRank = rank;
        h->eng.shardWorld = world;
    });
}

int sw4_set_database_files(sw4_handle* h, const char* db_prefix, int prefetch) {
    if (!h || !db_prefix) return SW4_ERR_INVALID;
    return guarded(h, [&] {
        auto db = std::make_unique<sw4::HostDB>();
        const std::string p = std::string(db_prefix) + "0";
        db->fChars.open(p + "chars", prefetch != 0);
        db->fOffsets.open(p + "offsets", true);
        db->fLengths.open(p + "lengths", true);
        db->fHeaders.open(p + "headers", false);
        db->fHeaderOffsets.open(p + "headeroffsets", false);
        db->n = db->fLengths.size / sizeof(int32_t);
        if (db->fOffsets.size != (db->n + 1) * sizeof(size_t) || db->fHeaderOffsets.size != (db->n + 1) * sizeof(size_t))
            sw4::fail(SW4_ERR_IO, "inconsistent database files for prefix %s", db_prefix);
        db->chars = (const uint8_t*)db->fChars.ptr;
        db->offsets = (const size_t*)db->fOffsets.ptr;
        db->lengths = (const int32_t*)db->fLengths.ptr;
        db->headers = (const char*)db->fHeaders.ptr;
        db->headerOffsets = (const size_t*)db->fHeaderOffsets.ptr;
        if (db->n && db->offsets[db->n] > db->fChars.size) sw4::fail(SW4_ERR_IO, "chars file too small for prefix %s", db_prefix);
        db->finish();
        h->eng.db = std::move(db);
        h->eng.assignShards();
    });
}

int sw4_set_database_memory(sw4_handle* h, const char* chars, const size_t* offsets, const int32_t* lengths, const char* headers,
                            const size_t* header_offsets, size_t num_sequences) {
    if (!h) return SW4_ERR_INVALID;
    return guarded(h, [&] {
        if (num_sequences && (!chars || !offsets || !lengths)) sw4::fail(SW4_ERR_INVALID, "null database arrays");
        auto db = std::make_unique<sw4::HostDB>();
        db->chars = (const uint8_t*)chars;
        db->offsets = offsets;
        db->lengths = lengths;
        db->headers = headers;
        db->headerOffsets = headers ? header_offsets : nullptr;
        db->n = num_sequences;
        db->finish();
        h->eng.db = std::move(db);
        h->eng.assignShards();
    });
}

int sw4_set_pseudo_database(sw4_handle* h, size_t num_sequences, int length, int seed) {
    if (!h) return SW4_ERR_INVALID;
    return guarded(h, [&] {
        if (length < 0) sw4::fail(SW4_ERR_INVALID, "negative length");
        auto db = std::make_unique<sw4::HostDB>();
        // reference src/dbdata.hpp:219-246: one random subject, replicated
        static const char letters[] = "ARNDCQEGHILKMFPSTWYV";
        std::mt19937 gen(seed);
        std::uniform_int_distribution<> dist(0, 19);
        std::vector<uint8_t> one(length);
        for (int i = 0; i < length; i++) {
            const char c = letters[dist(gen)];
            one[i] = (uint8_t)(strchr(letters, c) - letters);
        }
        const size_t padded = ((size_t)length + 3) / 4 * 4;
        db->vChars.assign(num_sequences * padded, 20);
        db->vOffsets.resize(num_sequences + 1);
        db->vLengths.assign(num_sequences, length);
        db->vHeaders.assign(num_sequences, 'H');
        db->vHeaderOffsets.resize(num_sequences + 1);
        for (size_t i = 0; i < num_sequences; i++) {
            db->vOffsets[i] = i * padded;
            db->vHeaderOffsets[i] = i;
            if (length) memcpy(db->vChars.data() + i * padded, one.data(), (size_t)length);
        }
        db->vOffsets[num_sequences] = num_sequences * padded;
        db->vHeaderOffsets[num_sequences] = num_sequences;
        db->chars = db->vChars.data();
        db->offsets = db->vOffsets.data();
        db->lengths = db->vLengths.data();
        db->headers = db->vHeaders.data();
        db->headerOffsets = db->vHeaderOffsets.data();
        db->n = num_sequences;
        db->finish();
        h->eng.db = std::move(db);
        h->eng.assignShards();
    });
}

int sw4_upload_database(sw4_handle* h) {
    if (!h) return SW4_ERR_INVALID;
    return guarded(h, [&] { h->eng.upload(); });
}

int sw4_scan(sw4_handle* h, const char* query, int32_t query_length, int32_t* out_scores, int32_t* out_ids, int32_t* out_count,
             sw4_stats* stats) {
    if (!h) return SW4_ERR_INVALID;
    return guarded(h, [&] {
        if (h->eng.numTop > 0 && (!out_scores || !out_ids)) sw4::fail(SW4_ERR_INVALID, "null output arrays");
        h->eng.scan(query, query_length, out_scores, out_ids, out_count, stats);
    });
}

int sw4_last_scan_all_scores(sw4_handle* h, int32_t* out_scores, int32_t* out_ids, size_t capacity, size_t* out_count) {
    if (!h) return SW4_ERR_INVALID;
    return guarded(h, [&] {
        size_t total = 0;
        for (auto& sh : h->eng.shards) total += sh->n;
        if (capacity < total) sw4::fail(SW4_ERR_INVALID, "capacity %zu < %zu", capacity, total);
        size_t pos = 0;
        for (auto& shp : h->eng.shards) {
            sw4::Shard& sh = *shp;
            if (!sh.uploaded) sw4::fail(SW4_ERR_INVALID, "no scan has run yet");
            SW4_CUDA(cudaSetDevice(sh.device));
            SW4_CUDA(cudaMemcpy(out_scores + pos, sh.dScores.p, sh.n * sizeof(int32_t), cudaMemcpyDeviceToHost));
            if (out_ids) memcpy(out_ids + pos, sh.globalIds.data(), sh.n * sizeof(int32_t));
            pos += sh.n;
        }
        if (out_count) *out_count = total;
    });
}

int sw4_reference_header(const sw4_handle* h, int32_t id, const char** ptr, size_t* len) {
    if (!h || !h->eng.db || id < 0 || (size_t)id >= h->eng.db->n || !ptr || !len) return SW4_ERR_INVALID;
    const sw4::HostDB& db = *h->eng.db;
    if (!db.headers) { *ptr = ""; *len = 0; return SW4_OK; }
    *ptr = db.headers + db.headerOffsets[id];
    *len = db.headerOffsets[id + 1] - db.headerOffsets[id];
    return SW4_OK;
}

int sw4_reference_length(const sw4_handle* h, int32_t id, int32_t* len) {
    if (!h || !h->eng.db || id < 0 || (size_t)id >= h->eng.db->n || !len) return SW4_ERR_INVALID;
    *len = h->eng.db->lengths[id];
    return SW4_OK;
}

int sw4_reference_sequence(const sw4_handle* h, int32_t id, char* out, size_t capacity, size_t* len) {
    if (!h || !h->eng.db || id < 0 || (size_t)id >= h->eng.db->n) return SW4_ERR_INVALID;
    const sw4::HostDB& db = *h->eng.db;
    const size_t L = (size_t)db.lengths[id];
    if (len) *len = L;
    if (!out) return SW4_OK;
    if (capacity < L) return SW4_ERR_INVALID;
    static const char inv[] = "ARNDCQEGHILKMFPSTWYV-";  // src/convert.cuh:36-64
    const uint8_t* s = db.chars + db.offsets[id];
    for (size_t i = 0; i < L; i++) out[i] = inv[s[i] > 20 ? 20 : s[i]];
    return SW4_OK;
}

int sw4_total_timer_start(sw4_handle* h) {
    if (!h) return SW4_ERR_INVALID;
    h->eng.totalStart = std::chrono::steady_clock::now();
    h->eng.totalCells = 0;
    h->eng.totalOverflows = 0;
    return SW4_OK;
}

int sw4_total_timer_stop(sw4_handle* h, sw4_stats* stats) {
    if (!h || !stats) return SW4_ERR_INVALID;
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - h->eng.totalStart).count();
    memset(stats, 0, sizeof(*stats));
    stats->seconds = s;
    stats->cells = h->eng.totalCells;
    stats->gcups = s > 0 ? h->eng.totalCells / 1e9 / s : 0;
    stats->num_overflows = h->eng.totalOverflows;
    return SW4_OK;
}

int sw4_get_db_info(const sw4_handle* h, sw4_db_info* info) {
    if (!h || !info || !h->eng.db) return SW4_ERR_INVALID;
    const sw4::HostDB& db = *h->eng.db;
    memset(info, 0, sizeof(*info));
    info->num_sequences = db.n;
    info->num_residues = db.residues;
    info->min_length = db.minLen;
    info->max_length = db.maxLen;
    size_t pos = 0;
    for (int p = 0; p < 36; p++) {  // membership b[i-1] < len <= b[i], src/cudasw4.cuh:904-926
        const size_t end = std::upper_bound(db.lengths + pos, db.lengths + db.n, sw4::kRefBoundaries[p]) - db.lengths;
        info->partition_counts[p] = end - pos;
        pos = end;
    }
    info->shard_rank = h->eng.shardRank;
    info->shard_world = h->eng.shardWorld;
    for (auto& sh : h->eng.shards) { info->shard_sequences += sh->n; info->shard_residues += sh->residues; }
    return SW4_OK;
}

}  // extern "C"
