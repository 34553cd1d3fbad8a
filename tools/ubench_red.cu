// ubench_red.cu -- microbenchmark (not product code): throughput of the accumulation primitives the
// gradVol scatter can be built from, on one B200.  Prints G lane-ops/s and GB/s per pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_red ubench_red.cu && ./ubench_red
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void red1(float*p,float v){asm volatile("red.global.add.f32 [%0], %1;"::"l"(p),"f"(v):"memory");}
__device__ __forceinline__ void red2(float*p,float a,float b){asm volatile("red.global.add.v2.f32 [%0], {%1,%2};"::"l"(p),"f"(a),"f"(b):"memory");}
__device__ __forceinline__ void red4(float*p,float a,float b,float c,float d){asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"::"l"(p),"f"(a),"f"(b),"f"(c),"f"(d):"memory");}

// mode 0: scalar red, lanes consecutive; 1: v2 consecutive; 2: v4 consecutive; 3: scalar, random addr;
// 4: scalar, consecutive but only lanes<8 active ; 5: v4, only lanes <8 active (same bytes as mode 0 per instr)
// 6: plain st (reference) scalar consecutive; 7: plain st v4
// each thread does ITER ops; footprint = `span` floats (wraps) so that the target stays L2-resident or not.
template<int MODE>
__global__ void k_red(float* g, size_t span_mask, int iters){
  const size_t tid = (size_t)blockIdx.x*blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const size_t warp = tid >> 5;
  unsigned rnd = (unsigned)tid*2654435761u + 12345u;
  for(int i=0;i<iters;i++){
    // each (warp, i) owns a distinct 128-float window; windows of different i are far apart
    size_t base = ((warp*131 + (size_t)i*7919) * 128) & span_mask;
    if(MODE==0){ red1(g + base + lane, 1.0f); }
    else if(MODE==1){ red2(g + base + 2*lane, 1.0f, 1.0f); }
    else if(MODE==2){ red4(g + base + 4*lane, 1.0f,1.0f,1.0f,1.0f); }
    else if(MODE==3){ rnd = rnd*1664525u+1013904223u; red1(g + ((size_t)(rnd>>4) & span_mask), 1.0f); }
    else if(MODE==4){ if(lane<8) red1(g + base + lane, 1.0f); }
    else if(MODE==5){ if(lane<8) red4(g + base + 4*lane, 1.0f,1.0f,1.0f,1.0f); }
    else if(MODE==6){ g[base+lane] = 1.0f; }
    else if(MODE==7){ *reinterpret_cast<float4*>(g+base+4*lane) = make_float4(1,1,1,1); }
    else if(MODE==8){ // scalar red, lanes consecutive, 8 different rows per iteration (like the 8 corners)
      #pragma unroll
      for(int k=0;k<8;k++) red1(g + ((base + (size_t)k*4096 + lane + (k&1)) & span_mask), 1.0f);
    }
  }
}

// shared memory accumulate: mode 0 atomicAdd float (CAS loop), 1 plain RMW (ld, add, st) conflict-free, 2 int atomicAdd
template<int MODE>
__global__ void k_smem(float* out, int iters){
  extern __shared__ float s[];
  const int n = 8192;
  for(int i=threadIdx.x;i<n;i+=blockDim.x) s[i]=0;
  __syncthreads();
  const int lane = threadIdx.x&31, w = threadIdx.x>>5;
  float* mine = s + w*1024; // private 1024-float window per warp (8 warps)
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int k=0;k<8;k++){
      int a = ((i*37 + k*96) & 991) + lane; if(a>1023) a-=1024;
      if(MODE==0) atomicAdd(&mine[a], 1.0f);
      else if(MODE==1){ float v = mine[a]; mine[a] = v + 1.0f; __syncwarp(); }
      else if(MODE==2) atomicAdd(reinterpret_cast<int*>(&mine[a]), 1);
    }
  }
  __syncthreads();
  float acc=0; for(int i=threadIdx.x;i<n;i+=blockDim.x) acc+=s[i];
  if(acc==-1.f) out[0]=acc;
}

template<int MODE> void run_red(const char* name, float* g, size_t span, int lanes_active, int floats_per_lane){
  const int threads=256, blocks=148*8, iters = (MODE==8)?32:256;
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_red<MODE><<<blocks,threads>>>(g, span-1, iters); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_red<MODE><<<blocks,threads>>>(g, span-1, iters);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms,e0,e1));
  double warp_instr = (double)blocks*threads/32*iters*((MODE==8)?8:1);
  double laneops = warp_instr*lanes_active;
  double bytes = laneops*floats_per_lane*4;
  printf("%-44s span %6.0f MiB: %7.3f ms  %8.2f G lane-ops/s  %8.1f GB/s  %6.2f cyc/warp-instr/SM@1.9GHz\n", name, span*4.0/1048576, ms,
         laneops/ms/1e6, bytes/ms/1e6, ms*1e-3*1.9e9*148/warp_instr);
}
template<int MODE> void run_smem(const char* name, float* g){
  const int threads=256, blocks=148*4, iters=2048;
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_smem<MODE><<<blocks,threads,32768>>>(g, iters); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_smem<MODE><<<blocks,threads,32768>>>(g, iters);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms,e0,e1));
  double warp_instr = (double)blocks*threads/32*iters*8;
  printf("%-44s %7.3f ms  %8.2f G lane-ops/s  %6.2f cyc/warp-op/SM@1.9GHz\n", name, ms, warp_instr*32/ms/1e6, ms*1e-3*1.9e9*148/warp_instr);
}
int main(){
  float* g; size_t big = (size_t)1<<28; // 1 GiB of floats = 2^28 floats
  CK(cudaMalloc(&g, big*4)); CK(cudaMemset(g,0,big*4));
  for(size_t span : {(size_t)1<<22 /*16 MiB: L2 resident*/, (size_t)1<<28 /*1 GiB*/}){
    run_red<0>("red.f32   32 lanes consecutive", g, span, 32, 1);
    run_red<1>("red.v2    32 lanes consecutive", g, span, 32, 2);
    run_red<2>("red.v4    32 lanes consecutive", g, span, 32, 4);
    run_red<3>("red.f32   32 lanes random", g, span, 32, 1);
    run_red<4>("red.f32    8 lanes consecutive", g, span, 8, 1);
    run_red<5>("red.v4     8 lanes consecutive", g, span, 8, 4);
    run_red<6>("st.f32    32 lanes consecutive", g, span, 32, 1);
    run_red<7>("st.v4     32 lanes consecutive", g, span, 32, 4);
    run_red<8>("red.f32   8 rows x 32 lanes (corner pattern)", g, span, 32, 1);
  }
  run_smem<0>("smem atomicAdd(float) CAS loop, no conflicts", g);
  run_smem<1>("smem plain RMW + syncwarp, no conflicts", g);
  run_smem<2>("smem atomicAdd(int), no conflicts", g);
  return 0;
}
