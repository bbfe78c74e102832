// Recurrence body in isolation, part 2: effect of occupancy (R columns x threads), of the LDS, and of instruction order.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template<int R,int THR,bool USE_LDS,int ORDER> __global__ void __launch_bounds__(THR,1) body(unsigned* out, const unsigned* in, unsigned gop2, unsigned gex2, int steps, long long* cyc){
    extern __shared__ unsigned sm[];
    for(int i=threadIdx.x;i<441*96;i+=blockDim.x) sm[i]=0x00010002u*(i%7);
    __syncthreads();
    unsigned Hp[R], F[R], col[R];
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm);
    for(int j=0;j<R;j++){ Hp[j]=0; F[j]=0xc180c180u; col[j]= USE_LDS ? base + ((in[(threadIdx.x*R+j)%16384]%441)*96 + (32-(threadIdx.x&31)))*4 : in[(threadIdx.x*R+j)%16384]&0x00070007; }
    unsigned mx=0, E=0xc180c180u, diag=0;
    long long t0=clock64();
    #pragma unroll 1
    for(int t=0;t<steps;t++){
        unsigned dPrev=0;
        unsigned s0; if(USE_LDS) asm volatile("ld.shared.u32 %0,[%1];":"=r"(s0):"r"(col[0])); else s0=col[0];
        unsigned d=__vadd2(diag,s0);
        #pragma unroll
        for(int j=0;j<R;j++){
            unsigned dNext=0;
            if(j+1<R){ unsigned s; if(USE_LDS) asm volatile("ld.shared.u32 %0,[%1+4];":"=r"(s):"r"(col[j+1])); else s=col[j+1]; dNext=__vadd2(Hp[j],s); }
            unsigned h=__vimax3_s16x2_relu(d,E,F[j]); Hp[j]=h; unsigned tt=__vadd2(h,gop2);
            if(ORDER==0){ E=__viaddmax_s16x2(E,gex2,tt); F[j]=__viaddmax_s16x2(F[j],gex2,tt); }
            else { F[j]=__viaddmax_s16x2(F[j],gex2,tt); E=__viaddmax_s16x2(E,gex2,tt); }
            if(j&1) mx=__vimax3_s16x2(mx,d,dPrev);
            dPrev=d; d=dNext;
        }
        diag=Hp[R-1]^E;
    }
    long long t1=clock64();
    unsigned acc=mx^E; for(int j=0;j<R;j++) acc^=Hp[j]^F[j];
    out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
    if(threadIdx.x==0) cyc[blockIdx.x]=t1-t0;
}
template<int R,int THR,bool L,int O> void run(const char* name, unsigned* out, unsigned* in, long long* cyc){
    const int steps=2000; const int smem=441*96*4;
    cudaFuncSetAttribute(body<R,THR,L,O>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    body<R,THR,L,O><<<148,THR,smem>>>(out,in,0xfff5fff5u,0xffffffffu,steps,cyc); cudaDeviceSynchronize();
    body<R,THR,L,O><<<148,THR,smem>>>(out,in,0xfff5fff5u,0xffffffffu,steps,cyc);
    cudaError_t e=cudaDeviceSynchronize(); if(e!=cudaSuccess){printf("err %s\n",cudaGetErrorString(e));return;}
    long long h; cudaMemcpy(&h,cyc,8,cudaMemcpyDeviceToHost);
    printf("%-40s R=%2d thr=%4d  %.2f cycles per cell-pair per scheduler\n", name, R, THR, double(h)/((THR/128.0)*steps*R));
}
int main(){
    unsigned *out,*in; long long* cyc; cudaMalloc(&out,148*1024*4); cudaMalloc(&in,16384*4); cudaMalloc(&cyc,148*8);
    unsigned* h=(unsigned*)malloc(16384*4); for(int i=0;i<16384;i++) h[i]=(unsigned)rand(); cudaMemcpy(in,h,16384*4,cudaMemcpyHostToDevice);
    run<32,512,true,0>("baseline",out,in,cyc);
    run<32,512,false,0>("no LDS",out,in,cyc);
    run<32,512,true,1>("F before E",out,in,cyc);
    run<32,256,true,0>("half occupancy",out,in,cyc);
    run<16,1024,true,0>("R=16, 1024 threads",out,in,cyc);
    run<16,512,true,0>("R=16, 512 threads",out,in,cyc);
    run<24,640,true,0>("R=24, 640 threads",out,in,cyc);
    run<8,1024,true,0>("R=8, 1024 threads",out,in,cyc);
    return 0;
}
