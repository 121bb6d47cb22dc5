// microbenchmark: FFMA vs FFMA2 issue throughput on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template<int MODE> __global__ void k(float* out, int iters){
  float a[16]; unsigned long long p[8];
  for(int q=0;q<16;q++) a[q]=threadIdx.x*0.001f+q;
  for(int q=0;q<8;q++){ float2 t=make_float2(a[2*q],a[2*q+1]); p[q]=*reinterpret_cast<unsigned long long*>(&t);}
  float b=1.0001f, c=0.5f; float2 b2=make_float2(b,b), c2=make_float2(c,c);
  unsigned long long pb=*reinterpret_cast<unsigned long long*>(&b2), pc=*reinterpret_cast<unsigned long long*>(&c2);
  for(int it=0; it<iters; it++){
    if(MODE==0){
#pragma unroll
      for(int q=0;q<16;q++) a[q]=fmaf(a[q],b,c);
    } else {
#pragma unroll
      for(int q=0;q<8;q++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[q]) : "l"(pb), "l"(pc));
    }
  }
  float s=0; for(int q=0;q<16;q++) s+=a[q];
  for(int q=0;q<8;q++){ float2 t=*reinterpret_cast<float2*>(&p[q]); s+=t.x+t.y; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){
  float* out; cudaMalloc(&out, 148*8*256*4*4);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters=20000; int blocks=148*8;
  for(int mode=0; mode<2; mode++){
    for(int rep=0; rep<2; rep++){
      cudaEventRecord(e0);
      if(mode==0) k<0><<<blocks,256>>>(out,iters); else k<1><<<blocks,256>>>(out,iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms,e0,e1);
      double fma = (double)blocks*256*iters*16;
      printf("mode %d (%s): %.3f ms, %.2f TFMA/s = %.1f TFLOP/s\n", mode, mode?"FFMA2":"FFMA", ms, fma/ms*1e-9, 2*fma/ms*1e-9);
    }
  }
  return 0;
}
