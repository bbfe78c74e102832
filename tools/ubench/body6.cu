// Two query rows per step per lane (one LDS.64 serves both rows of a column; F and Hp are touched once per two rows)
// versus the current one-row step. Same arithmetic per cell.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template<int R,int ROWS> __global__ void __launch_bounds__(512,1) body(unsigned* out, const unsigned* in, unsigned gop2, unsigned gex2, int steps, long long* cyc){
    extern __shared__ unsigned sm[];
    for(int i=threadIdx.x;i<441*96;i+=blockDim.x) sm[i]=0x00010002u*(i%7);
    __syncthreads();
    unsigned Hp[R], F[R], col[R];
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm);
    for(int j=0;j<R;j++){ Hp[j]=0; F[j]=0xc180c180u; col[j]= base + ((in[(threadIdx.x*R+j)%16384]%441)*96 + 2*(16-(threadIdx.x&15)))*4; }
    unsigned mx=0, E1=0xc180c180u, E2=0xc180c180u, diag1=0, diag2=0;
    long long t0=clock64();
    #pragma unroll 1
    for(int t=0;t<steps;t++){
        if constexpr (ROWS==1){
            unsigned dPrev=0; unsigned s0; asm volatile("ld.shared.u32 %0,[%1];":"=r"(s0):"r"(col[0]));
            unsigned d=__vadd2(diag1,s0);
            #pragma unroll
            for(int j=0;j<R;j++){
                unsigned dNext=0;
                if(j+1<R){ unsigned s; asm volatile("ld.shared.u32 %0,[%1+8];":"=r"(s):"r"(col[j+1])); dNext=__vadd2(Hp[j],s); }
                unsigned h=__vimax3_s16x2_relu(d,E1,F[j]); Hp[j]=h; unsigned tt=__vadd2(h,gop2);
                E1=__viaddmax_s16x2(E1,gex2,tt); F[j]=__viaddmax_s16x2(F[j],gex2,tt);
                if(j&1) mx=__vimax3_s16x2(mx,d,dPrev);
                dPrev=d; d=dNext;
            }
            diag1=Hp[R-1]^E1;
        } else {
            // rows a (upper) and b (lower) of this step; hPrevA = H[a][j-1] is the diagonal of (b, j)
            unsigned sa, sb; asm volatile("ld.shared.v2.u32 {%0,%1},[%2];":"=r"(sa),"=r"(sb):"r"(col[0]));
            unsigned da=__vadd2(diag1,sa), db=__vadd2(diag2,sb);
            #pragma unroll
            for(int j=0;j<R;j++){
                unsigned na=0, nb=0, sa2=0, sb2=0;
                if(j+1<R){ asm volatile("ld.shared.v2.u32 {%0,%1},[%2+8];":"=r"(sa2),"=r"(sb2):"r"(col[j+1])); na=__vadd2(Hp[j],sa2); }
                unsigned ha=__vimax3_s16x2_relu(da,E1,F[j]); unsigned ta=__vadd2(ha,gop2);
                E1=__viaddmax_s16x2(E1,gex2,ta); unsigned Fa=__viaddmax_s16x2(F[j],gex2,ta);
                if(j+1<R) nb=__vadd2(ha,sb2);               // diagonal of (b, j+1) is H[a][j]
                unsigned hb=__vimax3_s16x2_relu(db,E2,Fa); Hp[j]=hb; unsigned tb=__vadd2(hb,gop2);
                E2=__viaddmax_s16x2(E2,gex2,tb); F[j]=__viaddmax_s16x2(Fa,gex2,tb);
                mx=__vimax3_s16x2(mx,da,db);
                da=na; db=nb;
            }
            diag1=Hp[R-1]^E1; diag2=diag1^E2;
        }
    }
    long long t1=clock64();
    unsigned acc=mx^E1^E2; for(int j=0;j<R;j++) acc^=Hp[j]^F[j];
    out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
    if(threadIdx.x==0) cyc[blockIdx.x]=t1-t0;
}
template<int R,int ROWS> void run(const char* name, unsigned* out, unsigned* in, long long* cyc){
    const int steps=2000; const int smem=441*96*4;
    cudaFuncSetAttribute(body<R,ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    body<R,ROWS><<<148,512,smem>>>(out,in,0xfff5fff5u,0xffffffffu,steps,cyc); cudaDeviceSynchronize();
    body<R,ROWS><<<148,512,smem>>>(out,in,0xfff5fff5u,0xffffffffu,steps,cyc);
    cudaError_t e=cudaDeviceSynchronize(); if(e!=cudaSuccess){printf("err %s\n",cudaGetErrorString(e));return;}
    long long h; cudaMemcpy(&h,cyc,8,cudaMemcpyDeviceToHost);
    printf("%-28s R=%2d rows/step=%d  %.2f cycles per cell-pair per scheduler\n", name, R, ROWS, double(h)/(4.0*steps*R*ROWS));
}
int main(){
    unsigned *out,*in; long long* cyc; cudaMalloc(&out,148*1024*4); cudaMalloc(&in,16384*4); cudaMalloc(&cyc,148*8);
    unsigned* h=(unsigned*)malloc(16384*4); for(int i=0;i<16384;i++) h[i]=(unsigned)rand(); cudaMemcpy(in,h,16384*4,cudaMemcpyHostToDevice);
    run<32,1>("one row per step",out,in,cyc); run<32,2>("two rows per step",out,in,cyc); run<16,2>("two rows per step",out,in,cyc); run<24,2>("two rows per step",out,in,cyc);
    return 0;
}
