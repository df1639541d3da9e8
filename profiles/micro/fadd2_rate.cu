// Microbenchmark: issue/pipe rate of FADD vs packed FADD2 (add.rn.f32x2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fadd2_rate fadd2_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a,float b){u64 r; asm("mov.b64 %0,{%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(u64 v,float&a,float&b){asm("mov.b64 {%0,%1},%2;":"=f"(a),"=f"(b):"l"(v));}
__device__ __forceinline__ u64 add2(u64 a,u64 b){u64 r; asm volatile("add.rn.f32x2 %0,%1,%2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ float add1(float a,float b){float r; asm volatile("add.rn.f32 %0,%1,%2;":"=f"(r):"f"(a),"f"(b)); return r;}

constexpr int ITERS = 4096;
// MODE 0: 8 scalar FADD chains; 1: 8 FADD2 chains (pair operand); 2: 8 FADD2 chains (broadcast operand)
// MODE 3: 8 FADD2 + 8 LOP3 per iter; 4: 8 FADD + 8 LOP3; 5: 16 scalar FADD chains
template<int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed, uint32_t iseed){
  float a[16]; u64 p[8]; uint32_t x[8];
  #pragma unroll
  for(int i=0;i<16;++i) a[i]=seed*(i+1)+threadIdx.x;
  #pragma unroll
  for(int i=0;i<8;++i){ p[i]=pk(a[2*i],a[2*i+1]); x[i]=iseed+i*threadIdx.x; }
  float b=seed*0.5f; u64 bp=pk(b,b*3.f); u64 bb=pk(b,b);
  for(int it=0;it<ITERS;++it){
    if(MODE==0){
      #pragma unroll
      for(int i=0;i<8;++i) a[i]=add1(a[i],b);
    } else if(MODE==5){
      #pragma unroll
      for(int i=0;i<16;++i) a[i]=add1(a[i],b);
    } else if(MODE==1){
      #pragma unroll
      for(int i=0;i<8;++i) p[i]=add2(p[i],bp);
    } else if(MODE==2){
      #pragma unroll
      for(int i=0;i<8;++i) p[i]=add2(p[i],bb);
    } else if(MODE==3){
      #pragma unroll
      for(int i=0;i<8;++i){ p[i]=add2(p[i],bp); asm volatile("lop3.b32 %0,%0,%1,%2,0x96;":"+r"(x[i]):"r"(iseed),"r"(it)); }
    } else if(MODE==4){
      #pragma unroll
      for(int i=0;i<8;++i){ a[i]=add1(a[i],b); asm volatile("lop3.b32 %0,%0,%1,%2,0x96;":"+r"(x[i]):"r"(iseed),"r"(it)); }
    }
  }
  float s=0; 
  #pragma unroll
  for(int i=0;i<16;++i) s+=a[i];
  #pragma unroll
  for(int i=0;i<8;++i){ float u,v; upk(p[i],u,v); s+=u+v+__uint_as_float(x[i]&0x3fffffff); }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int MODE> void run(const char* name,int per_iter_instr,int lanes_per_instr){
  float* out; cudaMalloc(&out, 148*8*256*4);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int w=0;w<2;++w) k<MODE><<<148*8,256>>>(out,1.0f,7u);
  cudaEventRecord(e0);
  const int reps=10;
  for(int r=0;r<reps;++r) k<MODE><<<148*8,256>>>(out,1.0f,7u);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); ms/=reps;
  double warp_instr = (double)148*8*8*ITERS*per_iter_instr;   // warps * iters * instr
  int clk_khz; cudaDeviceGetAttribute(&clk_khz,cudaDevAttrClockRate,0);
  double cyc = ms*1e-3*clk_khz*1e3;
  printf("%-28s %.3f ms  %.3f warp-instr/clk/SMSP (at max clock)  %.2f fp32-lane-ops/clk/SM\n",name,ms,
         warp_instr/cyc/(148*4), warp_instr*32*lanes_per_instr/ (double)per_iter_instr * (MODE==3||MODE==4?0.5:1.0) /cyc/148);
  cudaFree(out);
}
int main(){
  run<0>("FADD x8",8,1);
  run<5>("FADD x16",16,1);
  run<1>("FADD2 pair x8",8,2);
  run<2>("FADD2 bcast x8",8,2);
  run<3>("FADD2 x8 + LOP3 x8",16,2);
  run<4>("FADD x8 + LOP3 x8",16,1);
  return 0;
}
