// Dependent-chain latency of the recurrence's instructions (one warp, ILP 1): cycles per instruction.
#include <cstdio>
#include <cuda_runtime.h>
template<int OP> __global__ void lat(unsigned* out, unsigned a, unsigned b, long long* cyc){
    unsigned x = threadIdx.x, y = threadIdx.x*3;
    long long t0 = clock64();
    #pragma unroll 1
    for(int it=0; it<1024; it++){
        #pragma unroll
        for(int i=0;i<16;i++){
            if constexpr (OP==0) x = __viaddmax_s16x2(x, a, b);
            if constexpr (OP==1) x = __vimax3_s16x2_relu(x, a, b);
            if constexpr (OP==2) x = __vadd2(x, a);
            if constexpr (OP==3) { x = __vimax3_s16x2_relu(x, a, b); x = __vadd2(x, a); x = __viaddmax_s16x2(y, b, x); } // current chain: ALU, FMA, ALU
            if constexpr (OP==4) { x = __vimax3_s16x2_relu(x, a, b); x = __viaddmax_s16x2(x, a, y); }                     // variant D chain: ALU, ALU
            if constexpr (OP==5) x = __vimax3_s16x2(x, a, b);
            if constexpr (OP==6) x = x * a + b;  // IMAD
            if constexpr (OP==7) x = __viaddmax_s32(x, a, b);
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + y;
    if(threadIdx.x==0) *cyc = t1 - t0;
}
template<int OP> void run(const char* name, int n){
    unsigned* o; long long* c; cudaMalloc(&o, 4096); cudaMalloc(&c, 8);
    lat<OP><<<1,32>>>(o, 0x00010001u, 0x00020003u, c); cudaDeviceSynchronize();
    lat<OP><<<1,32>>>(o, 0x00010001u, 0x00020003u, c); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-40s %.2f cycles per instruction (%.2f per chain unit)\n", name, double(h)/(1024.0*16*n), double(h)/(1024.0*16));
}
int main(){
    run<0>("VIADDMNMX.S16x2 dependent", 1); run<1>("VIMNMX3.S16x2.RELU dependent", 1); run<5>("VIMNMX3.S16x2 dependent", 1);
    run<2>("VIADD.16x2 dependent", 1); run<6>("IMAD dependent", 1); run<7>("VIADDMNMX.S32 dependent", 1);
    run<3>("chain H->t->E (ALU,FMA,ALU)", 3); run<4>("chain H->E (ALU,ALU)", 2);
    return 0;
}
