#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void up(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 r; asm("add.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ u64 mul2(u64 a, u64 b){ u64 r; asm("mul.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}

// MODE 0: scalar FFMA x16 ; 1: FFMA2 x8 ; 2: FADD2 x8 ; 3: scalar FADD x16 ; 4: complex twiddle mul packed (FMUL2+FFMA2 swizzled) x8 ; 5: complex twiddle scalar x8
// 6: FFMA2 x8 + 8 IADD (int alu) ; 7: FFMA x16 + 8 IADD
template<int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s, float c){
  float x[16];
  for(int i=0;i<16;i++) x[i]=threadIdx.x*0.001f+i;
  int q[8]; for(int i=0;i<8;i++) q[i]=threadIdx.x+i;
  if(MODE==0 || MODE==7){
    for(int it=0;it<iters;it++){
#pragma unroll
      for(int i=0;i<16;i++) x[i]=fmaf(x[i],s,c);
      if(MODE==7){
#pragma unroll
      for(int i=0;i<8;i++) q[i]=(q[i]^it)+i;
      }
    }
  } else if (MODE==1 || MODE==6) {
    u64 p[8]; for(int i=0;i<8;i++) p[i]=pk(x[2*i],x[2*i+1]);
    u64 ss=pk(s,s), cc=pk(c,c);
    for(int it=0;it<iters;it++){
#pragma unroll
      for(int i=0;i<8;i++) p[i]=fma2(p[i],ss,cc);
      if(MODE==6){
#pragma unroll
      for(int i=0;i<8;i++) q[i]=(q[i]^it)+i;
      }
    }
    for(int i=0;i<8;i++) up(p[i],x[2*i],x[2*i+1]);
  } else if (MODE==2) {
    u64 p[8]; for(int i=0;i<8;i++) p[i]=pk(x[2*i],x[2*i+1]);
    u64 cc=pk(c,0.5f*c);
    for(int it=0;it<iters;it++){
#pragma unroll
      for(int i=0;i<8;i++) p[i]=add2(p[i],cc);
    }
    for(int i=0;i<8;i++) up(p[i],x[2*i],x[2*i+1]);
  } else if (MODE==3) {
    for(int it=0;it<iters;it++){
#pragma unroll
      for(int i=0;i<16;i++) x[i]=x[i]+s;
    }
  } else if (MODE==4) {
    u64 p[8]; for(int i=0;i<8;i++) p[i]=pk(x[2*i],x[2*i+1]);
    for(int it=0;it<iters;it++){
#pragma unroll
      for(int i=0;i<8;i++){ float a,b; up(p[i],a,b); u64 t=mul2(p[i],pk(c,c)); p[i]=fma2(pk(b,a),pk(s,-s),t);}
    }
    for(int i=0;i<8;i++) up(p[i],x[2*i],x[2*i+1]);
  } else if (MODE==5) {
    for(int it=0;it<iters;it++){
#pragma unroll
      for(int i=0;i<8;i++){ float a=x[2*i], b=x[2*i+1]; x[2*i]=fmaf(a,c,b*s); x[2*i+1]=fmaf(b,c,-a*s);}
    }
  }
  float acc=0; for(int i=0;i<16;i++) acc+=x[i];
  int qa=0; for(int i=0;i<8;i++) qa+=q[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc+qa;
}
template<int M> void run(float* d,int iters,int ctas){
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best=1e9;
  for(int rep=0;rep<3;rep++){
    cudaEventRecord(a);
    k<M><<<148*ctas,256>>>(d,iters,0.9999f,0.70710678f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms;
  }
  double lane_ops=(double)148*ctas*256*iters*16;
  // per SMSP per cycle warp-level "scalar-equivalent" ops: lane_ops/32 / (148*4) / (ms*1e-3*1.965e9)
  double per=lane_ops/32/(148*4)/(best*1e-3*1.965e9);
  printf("mode %d ctas/SM %d: %.3f ms  scalar-equivalent warp-ops per SMSP-cycle (at 1965 MHz): %.3f\n",M,ctas,best,per);
}
int main(){
  float* d; cudaMalloc(&d, 148*8*256*4);
  int iters=20000;
  for(int ctas=2; ctas<=8; ctas*=2){
    run<0>(d,iters,ctas); run<1>(d,iters,ctas); run<2>(d,iters,ctas); run<3>(d,iters,ctas);
    run<4>(d,iters,ctas); run<5>(d,iters,ctas); run<6>(d,iters,ctas); run<7>(d,iters,ctas);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
