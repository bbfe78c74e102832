// Recurrence body in isolation, part 3: two independent alignments interleaved per lane (ILP 2) vs one.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template<int R,int NA> __global__ void __launch_bounds__(512,1) body(unsigned* out, const unsigned* in, unsigned gop2, unsigned gex2, int steps, long long* cyc){
    extern __shared__ unsigned sm[];
    for(int i=threadIdx.x;i<441*96;i+=blockDim.x) sm[i]=0x00010002u*(i%7);
    __syncthreads();
    unsigned Hp[NA][R], F[NA][R], col[NA][R];
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm);
    for(int a=0;a<NA;a++) for(int j=0;j<R;j++){ Hp[a][j]=0; F[a][j]=0xc180c180u; col[a][j]= base + ((in[(threadIdx.x*R*NA+a*R+j)%16384]%441)*96 + (32-(threadIdx.x&31)))*4; }
    unsigned mx[NA], E[NA], diag[NA];
    for(int a=0;a<NA;a++){ mx[a]=0; E[a]=0xc180c180u; diag[a]=0; }
    long long t0=clock64();
    #pragma unroll 1
    for(int t=0;t<steps;t++){
        unsigned d[NA], dPrev[NA];
        #pragma unroll
        for(int a=0;a<NA;a++){ unsigned s0; asm volatile("ld.shared.u32 %0,[%1];":"=r"(s0):"r"(col[a][0])); d[a]=__vadd2(diag[a],s0); dPrev[a]=0; }
        #pragma unroll
        for(int j=0;j<R;j++){
            #pragma unroll
            for(int a=0;a<NA;a++){
                unsigned dNext=0;
                if(j+1<R){ unsigned s; asm volatile("ld.shared.u32 %0,[%1+4];":"=r"(s):"r"(col[a][j+1])); dNext=__vadd2(Hp[a][j],s); }
                unsigned h=__vimax3_s16x2_relu(d[a],E[a],F[a][j]); Hp[a][j]=h; unsigned tt=__vadd2(h,gop2);
                E[a]=__viaddmax_s16x2(E[a],gex2,tt); F[a][j]=__viaddmax_s16x2(F[a][j],gex2,tt);
                if(j&1) mx[a]=__vimax3_s16x2(mx[a],d[a],dPrev[a]);
                dPrev[a]=d[a]; d[a]=dNext;
            }
        }
        #pragma unroll
        for(int a=0;a<NA;a++) diag[a]=Hp[a][R-1]^E[a];
    }
    long long t1=clock64();
    unsigned acc=0; for(int a=0;a<NA;a++){ acc^=mx[a]^E[a]; for(int j=0;j<R;j++) acc^=Hp[a][j]^F[a][j]; }
    out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
    if(threadIdx.x==0) cyc[blockIdx.x]=t1-t0;
}
template<int R,int NA> void run(const char* name, unsigned* out, unsigned* in, long long* cyc){
    const int steps=2000; const int smem=441*96*4;
    cudaFuncSetAttribute(body<R,NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    body<R,NA><<<148,512,smem>>>(out,in,0xfff5fff5u,0xffffffffu,steps,cyc); cudaDeviceSynchronize();
    body<R,NA><<<148,512,smem>>>(out,in,0xfff5fff5u,0xffffffffu,steps,cyc);
    cudaError_t e=cudaDeviceSynchronize(); if(e!=cudaSuccess){printf("err %s\n",cudaGetErrorString(e));return;}
    long long h; cudaMemcpy(&h,cyc,8,cudaMemcpyDeviceToHost);
    printf("%-24s R=%2d NA=%d  %.2f cycles per cell-pair per scheduler\n", name, R, NA, double(h)/(4.0*steps*R*NA));
}
int main(){
    unsigned *out,*in; long long* cyc; cudaMalloc(&out,148*1024*4); cudaMalloc(&in,16384*4); cudaMalloc(&cyc,148*8);
    unsigned* h=(unsigned*)malloc(16384*4); for(int i=0;i<16384;i++) h[i]=(unsigned)rand(); cudaMemcpy(in,h,16384*4,cudaMemcpyHostToDevice);
    run<32,1>("one chain",out,in,cyc); run<16,2>("two chains",out,in,cyc); run<8,4>("four chains",out,in,cyc); run<16,1>("one chain R16",out,in,cyc); run<12,2>("two chains R12",out,in,cyc);
    return 0;
}
