// Pipe-throughput micro-benchmark for the instructions the Gotoh recurrence is made of.
// Defines the integer/DPX roofline denominator: warp-instructions per clock per SM
// for VIADD.16x2, VIADDMNMX.S16x2, VIMNMX3.S16x2.RELU, IMAD, mixes of them, and LDS.32 under bank conflicts.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_pipes tools/ubench/pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

constexpr int ILP = 8;
constexpr int ITERS = 4096;

template<int OP>
__device__ __forceinline__ void step(unsigned (&x)[ILP], unsigned a, unsigned b, int one){
    #pragma unroll
    for(int i=0;i<ILP;i++){
        if constexpr (OP==0) x[i] = __vadd2(x[i], a);                          // VIADD.16x2
        if constexpr (OP==1) x[i] = __viaddmax_s16x2(x[i], a, b);              // VIADDMNMX.S16x2
        if constexpr (OP==2) x[i] = __vimax3_s16x2_relu(x[i], a, b);           // VIMNMX3.S16x2.RELU
        if constexpr (OP==3) x[i] = x[i] * one + a;                            // IMAD (fma pipe)
        if constexpr (OP==4) x[i] = (x[i] ^ a) & b;                            // LOP3 (alu)
        if constexpr (OP==5) { x[i] = __viaddmax_s16x2(x[i], a, b); x[i] = x[i]*one + a; } // 1 DPX : 1 IMAD
        if constexpr (OP==6) { x[i] = __viaddmax_s16x2(x[i], a, b); x[i] = __vimax3_s16x2_relu(x[i], b, a); x[i] = x[i]*one + a; } // 2 DPX : 1 IMAD
        if constexpr (OP==7) { x[i] = __viaddmax_s16x2(x[i], a, b); x[i] = __vadd2(x[i], a);} // DPX + VIADD16x2 (same pipe?)
        if constexpr (OP==8) x[i] = __vmaxs2(x[i], a);                         // VIMNMX.S16x2
        if constexpr (OP==9) { x[i] = __vadd2(x[i], a); x[i] = x[i]*one + b; } // VIADD + IMAD
        if constexpr (OP==10) x[i] = __viaddmax_s32(x[i], a, b);               // VIADDMNMX s32
        if constexpr (OP==11) x[i] = __vimax3_s32_relu(x[i], a, b);            // VIMNMX3 s32 relu
        if constexpr (OP==12) x[i] = x[i] + a;                                 // IADD (whatever ptxas picks)
        if constexpr (OP==13) { x[i] = __viaddmax_s16x2(x[i], a, b); x[i] = __vimax3_s16x2_relu(x[i], b, a); x[i] = __viaddmax_s16x2(x[i], b, a); x[i] = x[i]*one + a; x[i] = x[i]*one + b;} // 3 DPX : 2 IMAD
    }
}
static const int OPS_PER_STEP[] = {1,1,1,1,1,2,3,2,1,2,1,1,1,5,1};
static const char* NAMES[] = {"VIADD.16x2","VIADDMNMX.S16x2","VIMNMX3.S16x2.RELU","IMAD","LOP3","DPX+IMAD(1:1)","DPX+DPX+IMAD(2:1)","DPX+VIADD16x2","VIMNMX.S16x2","VIADD16x2+IMAD","VIADDMNMX.S32","VIMNMX3.S32.RELU","IADD","3DPX+2IMAD",""};

template<int OP>
__global__ void __launch_bounds__(1024) kern(unsigned* out, unsigned a, unsigned b, int one, long long* cycles){
    unsigned x[ILP];
    #pragma unroll
    for(int i=0;i<ILP;i++) x[i] = threadIdx.x * 7 + i;
    long long t0 = clock64();
    #pragma unroll 1
    for(int it=0; it<ITERS; it++){
        step<OP>(x,a,b,one);
        step<OP>(x,b,a,one);
    }
    long long t1 = clock64();
    unsigned s=0;
    #pragma unroll
    for(int i=0;i<ILP;i++) s ^= x[i];
    out[blockIdx.x*blockDim.x+threadIdx.x] = s;
    if(threadIdx.x==0) cycles[blockIdx.x] = t1-t0;
}

// LDS.32 with a controlled conflict degree: lane l reads word (l % distinct)*stride... random pattern from table
__global__ void __launch_bounds__(1024) lds_kern(unsigned* out, const int* pattern, int npat, long long* cycles){
    extern __shared__ unsigned sm[];
    for(int i=threadIdx.x;i<9261;i+=blockDim.x) sm[i]=i*2654435761u;
    __syncthreads();
    int idx[8];
    for(int i=0;i<8;i++) idx[i] = pattern[(threadIdx.x + i*1024) % npat];
    unsigned acc=0;
    long long t0=clock64();
    #pragma unroll 1
    for(int it=0; it<ITERS; it++){
        #pragma unroll
        for(int i=0;i<8;i++){ unsigned v = sm[idx[i]]; acc += v; idx[i] = (idx[i] + (v&0)) ; }
    }
    long long t1=clock64();
    out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
    if(threadIdx.x==0) cycles[blockIdx.x]=t1-t0;
}

template<int OP>
void run(int threads, int blocksPerSm, int nsm, unsigned* d_out, long long* d_cyc){
    int blocks = nsm*blocksPerSm;
    cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<OP><<<blocks,threads>>>(d_out, 0x00010002u, 0x00030001u, 1, d_cyc); // warm
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    kern<OP><<<blocks,threads>>>(d_out, 0x00010002u, 0x00030001u, 1, d_cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms,e0,e1));
    long long cyc; CK(cudaMemcpy(&cyc,d_cyc,8,cudaMemcpyDeviceToHost));
    double warpinst = double(threads/32)*blocksPerSm*ILP*2.0*ITERS*OPS_PER_STEP[OP];
    printf("%-22s thr=%4d bps=%d  cycles(block0)=%lld  warp-inst/clk/SM=%.3f  lanes/clk/SM=%.1f  ms=%.3f  => clk=%.0f MHz\n",
        NAMES[OP], threads, blocksPerSm, cyc, warpinst/cyc, warpinst*32/cyc, ms, cyc/(ms*1e3));
}

int main(){
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
    printf("device %s sms=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int nsm=p.multiProcessorCount;
    unsigned* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, sizeof(unsigned)*nsm*4*1024)); CK(cudaMalloc(&d_cyc, 8*nsm*4));
    for(int threads : {256, 512, 1024}){
        run<0>(threads,1,nsm,d_out,d_cyc); run<1>(threads,1,nsm,d_out,d_cyc); run<2>(threads,1,nsm,d_out,d_cyc);
        run<8>(threads,1,nsm,d_out,d_cyc); run<3>(threads,1,nsm,d_out,d_cyc); run<4>(threads,1,nsm,d_out,d_cyc);
        run<12>(threads,1,nsm,d_out,d_cyc); run<10>(threads,1,nsm,d_out,d_cyc); run<11>(threads,1,nsm,d_out,d_cyc);
        run<5>(threads,1,nsm,d_out,d_cyc); run<6>(threads,1,nsm,d_out,d_cyc); run<7>(threads,1,nsm,d_out,d_cyc);
        run<9>(threads,1,nsm,d_out,d_cyc); run<13>(threads,1,nsm,d_out,d_cyc);
    }
    // LDS conflicts
    int hpat[8192];
    int* d_pat; CK(cudaMalloc(&d_pat, sizeof(hpat)));
    CK(cudaFuncSetAttribute(lds_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 9261*4));
    for(int mode=0; mode<4; mode++){
        srand(1);
        for(int i=0;i<8192;i++){
            if(mode==0) hpat[i] = i%32;                 // conflict-free
            if(mode==1) hpat[i] = rand()%9261;          // random over the 21^3 LUT
            if(mode==2) hpat[i] = (i%8)*441 + 7;        // 8 distinct (broadcast groups)
            if(mode==3) hpat[i] = (i%32)*32;            // 32-way conflict
        }
        CK(cudaMemcpy(d_pat,hpat,sizeof(hpat),cudaMemcpyHostToDevice));
        for(int threads : {256,1024}){
            lds_kern<<<nsm,threads,9261*4>>>(d_out,d_pat,8192,d_cyc);
            CK(cudaDeviceSynchronize());
            long long cyc; CK(cudaMemcpy(&cyc,d_cyc,8,cudaMemcpyDeviceToHost));
            double wi = double(threads/32)*8.0*ITERS;
            printf("LDS.32 mode=%d thr=%4d cycles=%lld  warp-LDS/clk/SM=%.3f\n", mode, threads, cyc, wi/cyc);
        }
    }
    return 0;
}
