// Does instruction ORDER matter for the 3.5 DPX : 2 VIADD.16x2 (: 1 LDS) mix? Independent chains, explicit orders.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS=2048;
#define D1(i) x[i]=__viaddmax_s16x2(x[i],a,b);
#define D2(i) x[i]=__vimax3_s16x2_relu(x[i],b,a);
#define V(i)  y[i]=__vadd2(y[i],a);
#define L(i)  z[i]+=sm[(threadIdx.x+i*32)&1023];
template<int ORD> __global__ void __launch_bounds__(1024) k(unsigned* out, unsigned a, unsigned b, long long* cyc){
    unsigned x[8],y[4],z[2]; for(int i=0;i<8;i++)x[i]=threadIdx.x*7+i; for(int i=0;i<4;i++)y[i]=threadIdx.x*3+i; z[0]=z[1]=0;
    extern __shared__ unsigned sm[]; sm[threadIdx.x]=threadIdx.x; __syncthreads();
    long long t0=clock64();
    #pragma unroll 1
    for(int it=0;it<ITERS;it++){
        if constexpr(ORD==0){ D1(0) D2(1) D1(2) D2(3) D1(4) D2(5) D1(6) V(0) V(1) V(2) V(3) }                 // grouped
        if constexpr(ORD==1){ D1(0) V(0) D2(1) D1(2) V(1) D2(3) D1(4) V(2) D2(5) D1(6) V(3) }                 // interleaved
        if constexpr(ORD==2){ D1(0) V(0) D2(1) L(0) D1(2) V(1) D2(3) D1(4) V(2) D2(5) L(1) D1(6) V(3) }       // interleaved + 2 LDS
        if constexpr(ORD==3){ D1(0) D2(1) D1(2) D2(3) D1(4) D2(5) D1(6) V(0) V(1) V(2) V(3) L(0) L(1) }       // grouped + 2 LDS
        if constexpr(ORD==4){ D1(0) D2(1) D1(2) D2(3) D1(4) D2(5) D1(6) }                                     // DPX only
        if constexpr(ORD==5){ D1(0) D2(1) D1(2) D2(3) D1(4) D2(5) D1(6) L(0) L(1) }                           // DPX + LDS
    }
    long long t1=clock64();
    unsigned s=z[0]^z[1]; for(int i=0;i<8;i++)s^=x[i]; for(int i=0;i<4;i++)s^=y[i];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s; if(threadIdx.x==0)cyc[blockIdx.x]=t1-t0;
}
template<int ORD> void run(const char* n,int thr,unsigned* o,long long* c){
    k<ORD><<<148,thr,4096>>>(o,0x00010002u,0x00030001u,c); cudaDeviceSynchronize(); k<ORD><<<148,thr,4096>>>(o,0x00010002u,0x00030001u,c); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h,c,8,cudaMemcpyDeviceToHost);
    printf("%-30s thr=%4d  %.2f cycles per 3.5 DPX (+2 VIADD [+1 LDS]) per scheduler\n",n,thr,double(h)/(ITERS*(thr/128.0))/2);
}
int main(){ unsigned* o; long long* c; cudaMalloc(&o,148*1024*4); cudaMalloc(&c,148*8);
  for(int thr: {512,1024}){ run<4>("DPX only",thr,o,c); run<5>("DPX + LDS",thr,o,c); run<0>("grouped",thr,o,c); run<1>("interleaved",thr,o,c); run<3>("grouped + LDS",thr,o,c); run<2>("interleaved + LDS",thr,o,c); }
  return 0; }
