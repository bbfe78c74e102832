// Does DPX throughput depend on how many distinct register operands an instruction reads? (register-file bandwidth)
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N=8, ITERS=2048;
template<int OP,int PAT> __global__ void __launch_bounds__(1024) k(unsigned* out, unsigned a, unsigned b, long long* cyc){
    unsigned x[N],y[N],z[N];
    for(int i=0;i<N;i++){x[i]=threadIdx.x+i; y[i]=threadIdx.x*3+i; z[i]=threadIdx.x*5+i;}
    long long t0=clock64();
    #pragma unroll 1
    for(int it=0;it<ITERS;it++){
        #pragma unroll
        for(int r=0;r<2;r++)
        #pragma unroll
        for(int i=0;i<N;i++){
            unsigned p = PAT>=1 ? y[(i+r)%N] : a;
            unsigned q = PAT>=2 ? z[(i+r*3)%N] : b;
            if constexpr(OP==0) x[i]=__vimax3_s16x2_relu(x[i],p,q);
            if constexpr(OP==1) x[i]=__viaddmax_s16x2(x[i],p,q);
            if constexpr(OP==2) x[i]=__vmaxs2(x[i],p);
            if constexpr(OP==3) x[i]=__vadd2(x[i],p);
        }
    }
    long long t1=clock64();
    unsigned s=0; for(int i=0;i<N;i++) s^=x[i]^y[i]^z[i];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s; if(threadIdx.x==0) cyc[blockIdx.x]=t1-t0;
}
template<int OP,int PAT> void run(const char* n,int thr,unsigned* o,long long* c){
    k<OP,PAT><<<148,thr>>>(o,1,2,c); cudaDeviceSynchronize(); k<OP,PAT><<<148,thr>>>(o,1,2,c); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h,c,8,cudaMemcpyDeviceToHost);
    double wi=double(thr/32)*ITERS*2*N; printf("%-34s thr=%4d  %.3f warp-inst/clk/SM\n",n,thr,wi/h);
}
int main(){ unsigned* o; long long* c; cudaMalloc(&o,148*1024*4); cudaMalloc(&c,148*8);
  for(int thr: {512,1024}){
    run<0,0>("VIMNMX3.RELU x,a,b (1 varying)",thr,o,c); run<0,1>("VIMNMX3.RELU x,y,b (2 varying)",thr,o,c); run<0,2>("VIMNMX3.RELU x,y,z (3 varying)",thr,o,c);
    run<1,0>("VIADDMNMX x,a,b",thr,o,c); run<1,1>("VIADDMNMX x,y,b",thr,o,c); run<1,2>("VIADDMNMX x,y,z",thr,o,c);
    run<2,0>("VIMNMX x,a",thr,o,c); run<2,1>("VIMNMX x,y",thr,o,c);
    run<3,0>("VIADD.16x2 x,a",thr,o,c); run<3,1>("VIADD.16x2 x,y",thr,o,c);
  } return 0; }
