// Which combinations of FMA-heavy ops and shared-memory loads slow the DPX stream down? (follow-up of order.cu)
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS=2048;
#define D1(i) x[i]=__viaddmax_s16x2(x[i],a,b);
#define D2(i) x[i]=__vimax3_s16x2_relu(x[i],b,a);
#define DPX7 D1(0) D2(1) D1(2) D2(3) D1(4) D2(5) D1(6)
#define V(i)  y[i]=__vadd2(y[i],a);
#define IM(i) y[i]=y[i]*a+b;
#define FA(i) f[i]=f[i]+fa;
#define L(i)  z[i]+=sm[(threadIdx.x+i*32)&1023];
#define L64(i) { uint2 v=((uint2*)sm)[(threadIdx.x+i*32)&511]; z[i]+=v.x^v.y; }
#define SH(i) z[i]+=__shfl_up_sync(0xffffffffu,z[i],1);
template<int C> __global__ void __launch_bounds__(1024) k(unsigned* out, unsigned a, unsigned b, float fa, long long* cyc){
    unsigned x[8],y[4],z[2]; float f[4]; for(int i=0;i<8;i++)x[i]=threadIdx.x*7+i; for(int i=0;i<4;i++){y[i]=threadIdx.x*3+i; f[i]=i+threadIdx.x;} z[0]=z[1]=1;
    extern __shared__ unsigned sm[]; sm[threadIdx.x]=threadIdx.x; __syncthreads();
    long long t0=clock64();
    #pragma unroll 1
    for(int it=0;it<ITERS;it++){
        DPX7
        if constexpr(C==1){ V(0) V(1) V(2) V(3) L(0) L(1) }
        if constexpr(C==2){ IM(0) IM(1) IM(2) IM(3) L(0) L(1) }
        if constexpr(C==3){ FA(0) FA(1) FA(2) FA(3) L(0) L(1) }
        if constexpr(C==4){ V(0) V(1) L(0) L(1) }
        if constexpr(C==5){ V(0) V(1) V(2) V(3) L(0) }
        if constexpr(C==6){ V(0) V(1) V(2) V(3) L64(0) }
        if constexpr(C==7){ V(0) V(1) V(2) V(3) SH(0) SH(1) }
        if constexpr(C==8){ L(0) L(1) }
        if constexpr(C==9){ V(0) V(1) V(2) V(3) }
        if constexpr(C==10){ V(0) V(1) V(2) V(3) z[0]+=sm[(threadIdx.x)&1023]; z[1]+=sm[(threadIdx.x+7)&1023]; }   // loop-invariant addresses (no address math)
    }
    long long t1=clock64();
    unsigned s=z[0]^z[1]; for(int i=0;i<8;i++)s^=x[i]; for(int i=0;i<4;i++)s^=y[i]^__float_as_uint(f[i]);
    out[blockIdx.x*blockDim.x+threadIdx.x]=s; if(threadIdx.x==0)cyc[blockIdx.x]=t1-t0;
}
template<int C> void run(const char* n,unsigned* o,long long* c){
    for(int thr: {512,1024}){
    k<C><<<148,thr,4096>>>(o,0x00010002u,0x00030001u,1.5f,c); cudaDeviceSynchronize(); k<C><<<148,thr,4096>>>(o,0x00010002u,0x00030001u,1.5f,c); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h,c,8,cudaMemcpyDeviceToHost);
    printf("%-36s thr=%4d  %.2f cycles per 3.5 DPX unit per scheduler\n",n,thr,double(h)/(ITERS*(thr/128.0))/2); }
}
int main(){ unsigned* o; long long* c; cudaMalloc(&o,148*1024*4); cudaMalloc(&c,148*8);
  run<8>("7 DPX + 2 LDS",o,c); run<9>("7 DPX + 4 VIADD",o,c); run<1>("7 DPX + 4 VIADD + 2 LDS",o,c); run<2>("7 DPX + 4 IMAD + 2 LDS",o,c); run<3>("7 DPX + 4 FADD + 2 LDS",o,c);
  run<4>("7 DPX + 2 VIADD + 2 LDS",o,c); run<5>("7 DPX + 4 VIADD + 1 LDS",o,c); run<6>("7 DPX + 4 VIADD + 1 LDS.64",o,c); run<7>("7 DPX + 4 VIADD + 2 SHFL",o,c); run<10>("7 DPX + 4 VIADD + 2 LDS (fixed addr)",o,c);
  return 0; }
