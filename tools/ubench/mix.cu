// Throughput of instruction MIXES (no tight dependencies): 7 DPX ops per 4 "add" ops of different kinds.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
constexpr int ITERS=2048;
template<int KIND> __global__ void __launch_bounds__(1024) k(unsigned* out, unsigned a, unsigned b, float fa, long long* cyc){
    unsigned x[8]; float f[4]; unsigned y[4];
    for(int i=0;i<8;i++) x[i]=threadIdx.x*7+i; for(int i=0;i<4;i++){ f[i]=threadIdx.x*0.5f+i; y[i]=threadIdx.x*3+i; }
    extern __shared__ unsigned sm[]; sm[threadIdx.x]=threadIdx.x; __syncthreads();
    long long t0=clock64();
    #pragma unroll 1
    for(int it=0;it<ITERS;it++){
        // 7 DPX
        x[0]=__viaddmax_s16x2(x[0],a,b); x[1]=__vimax3_s16x2_relu(x[1],a,b); x[2]=__viaddmax_s16x2(x[2],b,a); x[3]=__vimax3_s16x2_relu(x[3],b,a);
        x[4]=__viaddmax_s16x2(x[4],a,b); x[5]=__vimax3_s16x2(x[5],a,b); x[6]=__viaddmax_s16x2(x[6],b,a);
        // 4 adds
        #pragma unroll
        for(int i=0;i<4;i++){
            if constexpr(KIND==1) y[i]=__vadd2(y[i],a);
            if constexpr(KIND==2) y[i]=y[i]*a+b;           // IMAD
            if constexpr(KIND==3) y[i]=(y[i]^a)+b;         // LOP3+IADD-ish (ALU)
            if constexpr(KIND==4) f[i]=f[i]+fa;            // FADD
            if constexpr(KIND==5) { __half2 h=*(__half2*)&y[i]; h=__hadd2(h,*(__half2*)&a); y[i]=*(unsigned*)&h; } // HADD2
            if constexpr(KIND==6) f[i]=fmaf(f[i],fa,fa);   // FFMA
            if constexpr(KIND==7) { y[i]=__vadd2(y[i],a); }
        }
        if constexpr(KIND==7){ x[7]+=sm[(x[7]&1023)]; y[0]+=sm[(y[0]&1023)]; } // + 2 LDS (dependent addr but hidden by occupancy)
    }
    long long t1=clock64();
    unsigned s=0; for(int i=0;i<8;i++) s^=x[i]; for(int i=0;i<4;i++) s^=y[i]^__float_as_uint(f[i]);
    out[blockIdx.x*blockDim.x+threadIdx.x]=s; if(threadIdx.x==0) cyc[blockIdx.x]=t1-t0;
}
template<int KIND> void run(const char* n,int thr,unsigned* o,long long* c){
    k<KIND><<<148,thr,4096>>>(o,0x00010002u,0x00030001u,1.0001f,c); cudaDeviceSynchronize(); k<KIND><<<148,thr,4096>>>(o,0x00010002u,0x00030001u,1.0001f,c); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h,c,8,cudaMemcpyDeviceToHost);
    printf("%-28s thr=%4d  %.2f cycles per (7 DPX + 4 x) group per scheduler => %.2f per 3.5 DPX\n",n,thr,double(h)/(ITERS*(thr/128.0)), double(h)/(ITERS*(thr/128.0))/2);
}
int main(){ unsigned* o; long long* c; cudaMalloc(&o,148*1024*4); cudaMalloc(&c,148*8);
  for(int thr: {512,1024}){
    run<0>("7 DPX alone",thr,o,c); run<1>("+4 VIADD.16x2",thr,o,c); run<2>("+4 IMAD",thr,o,c); run<3>("+4 (LOP3,IADD) ALU",thr,o,c);
    run<4>("+4 FADD",thr,o,c); run<5>("+4 HADD2",thr,o,c); run<6>("+4 FFMA",thr,o,c); run<7>("+4 VIADD.16x2 +2 LDS",thr,o,c);
  } return 0; }
