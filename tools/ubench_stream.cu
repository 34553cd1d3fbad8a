// ubench_stream.cu -- microbenchmark (not product code): what does a kernel with the forward sampler's
// HBM access pattern (16 B/voxel streamed in, 4 B/voxel streamed out, + 4 B/voxel volume) achieve when it
// does no gathers and no arithmetic?  Gives the realistic ceiling for k_sampler_fwd_packed.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__device__ __forceinline__ float4 ldna(const float4*p){float4 r; asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];":"=f"(r.x),"=f"(r.y),"=f"(r.z),"=f"(r.w):"l"(p)); return r;}

template<int VPT, int MODE>  // MODE 0: grid->out ; 1: + read vol[j] (coalesced) ; 2: + 8 coalesced volume reads (identity gather pattern)
__global__ void __launch_bounds__(256) k(const float4* __restrict__ g, const float* __restrict__ vol, float* __restrict__ out, long long n, int W, int HW){
  long long j0 = (long long)blockIdx.x*256*VPT + threadIdx.x;
  float4 v[VPT];
  #pragma unroll
  for(int i=0;i<VPT;i++){ long long j = min(j0+i*256, n-1); v[i]=ldna(g+j); }
  #pragma unroll
  for(int i=0;i<VPT;i++){
    long long j = j0+i*256;
    float r = v[i].x+v[i].y+v[i].z;
    if(MODE==1){ r += __ldg(vol + min(j,n-1)); }
    if(MODE==2){
      long long jj = min(j, n-1-HW-W-1);
      const float* p = vol + jj;
      r += __ldg(p) + __ldg(p+1) + __ldg(p+W) + __ldg(p+W+1) + __ldg(p+HW) + __ldg(p+HW+1) + __ldg(p+HW+W) + __ldg(p+HW+W+1);
    }
    if(j<n) __stcs(out+j, r);
  }
}
template<int VPT,int MODE> void run(const char* name, const float4*g, const float*vol, float*out, long long n){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  unsigned blocks = (unsigned)((n + 256*VPT-1)/(256*VPT));
  for(int i=0;i<3;i++) k<VPT,MODE><<<blocks,256>>>(g,vol,out,n,64,4096);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for(int i=0;i<10;i++) k<VPT,MODE><<<blocks,256>>>(g,vol,out,n,64,4096);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); ms/=10;
  double bytes = (double)n*(16+4+(MODE?4:0));
  printf("%-40s VPT=%d: %7.1f us  %7.1f GB/s (%.1f%% of 6538)\n", name, VPT, ms*1e3, bytes/ms/1e6, bytes/ms/1e6/65.383);
}
__device__ __forceinline__ void red1(float*p,float v){asm volatile("red.global.add.f32 [%0], %1;"::"l"(p),"f"(v));}
// backward pattern.  MODE 0: grid(16)+go(4) -> ggrid(16) ; 1: + 8 coalesced volume reads ; 2: + 8 coalesced REDs into gv ; 3: REDs but no ggrid store
// 4: 4 REDs (x-merged pattern) + everything
template<int VPT, int MODE>
__global__ void __launch_bounds__(256) kb(const float4* __restrict__ g, const float* __restrict__ go, const float* __restrict__ vol, float* __restrict__ gv, float4* __restrict__ gg, long long n, int W, int HW){
  long long j0 = (long long)blockIdx.x*256*VPT + threadIdx.x;
  float4 v[VPT]; float o[VPT];
  #pragma unroll
  for(int i=0;i<VPT;i++){ long long j = min(j0+i*256, n-1); v[i]=ldna(g+j); o[i]=__ldg(go+j); }
  #pragma unroll
  for(int i=0;i<VPT;i++){
    long long j = j0+i*256;
    float r = v[i].x+v[i].y+v[i].z+o[i];
    long long jj = min(j, n-1-HW-W-1);
    if(MODE>=1){ const float* p = vol + jj; r += __ldg(p) + __ldg(p+1) + __ldg(p+W) + __ldg(p+W+1) + __ldg(p+HW) + __ldg(p+HW+1) + __ldg(p+HW+W) + __ldg(p+HW+W+1); }
    if(MODE==2 || MODE==3){ float* p = gv + jj; if(j<n){ red1(p,r); red1(p+1,r); red1(p+W,r); red1(p+W+1,r); red1(p+HW,r); red1(p+HW+1,r); red1(p+HW+W,r); red1(p+HW+W+1,r);} }
    if(MODE==4){ float* p = gv + jj; if(j<n){ red1(p+1,r); red1(p+W+1,r); red1(p+HW+1,r); red1(p+HW+W+1,r);} }
    if(MODE!=3 && j<n) __stcs(gg+j, make_float4(r,r,r,0.f));
    if(MODE==3 && r==123.456f && j<n) __stcs(gg+j, make_float4(r,r,r,0.f));
  }
}
template<int VPT,int MODE> void runb(const char* name, const float4*g, const float*go, const float*vol, float*gv, float4*gg, long long n, bool memset_gv){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  unsigned blocks = (unsigned)((n + 256*VPT-1)/(256*VPT));
  for(int i=0;i<3;i++) kb<VPT,MODE><<<blocks,256>>>(g,go,vol,gv,gg,n,64,4096);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for(int i=0;i<10;i++){ if(memset_gv) CK(cudaMemsetAsync(gv,0,n*4)); kb<VPT,MODE><<<blocks,256>>>(g,go,vol,gv,gg,n,64,4096); }
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); ms/=10;
  printf("%-58s VPT=%d: %7.1f us\n", name, VPT, ms*1e3);
}
int main(){
  long long n = 64LL*64*64*64;
  float4* g; float* vol; float* out;
  CK(cudaMalloc(&g, n*16)); CK(cudaMalloc(&vol, n*4)); CK(cudaMalloc(&out, n*4));
  CK(cudaMemset(g,0,n*16)); CK(cudaMemset(vol,0,n*4));
  run<1,0>("grid->out", g,vol,out,n); run<2,0>("grid->out", g,vol,out,n); run<4,0>("grid->out", g,vol,out,n);
  run<1,1>("grid+vol->out", g,vol,out,n); run<2,1>("grid+vol->out", g,vol,out,n); run<4,1>("grid+vol->out", g,vol,out,n);
  run<1,2>("grid+8 coalesced gathers->out", g,vol,out,n); run<2,2>("grid+8 coalesced gathers->out", g,vol,out,n); run<4,2>("grid+8 coalesced gathers->out", g,vol,out,n);
  { float* go; float* gv; float4* gg; CK(cudaMalloc(&go,n*4)); CK(cudaMalloc(&gv,n*4)); CK(cudaMalloc(&gg,n*16)); CK(cudaMemset(go,0,n*4));
    runb<2,0>("bwd: grid+go -> ggrid", g,go,vol,gv,gg,n,false);
    runb<2,1>("bwd: grid+go+8 gathers -> ggrid", g,go,vol,gv,gg,n,false);
    runb<2,2>("bwd: grid+go+8 gathers -> ggrid + 8 REDs", g,go,vol,gv,gg,n,false);
    runb<2,2>("bwd: grid+go+8 gathers -> ggrid + 8 REDs + memset", g,go,vol,gv,gg,n,true);
    runb<2,4>("bwd: grid+go+8 gathers -> ggrid + 4 REDs + memset", g,go,vol,gv,gg,n,true);
    runb<2,3>("bwd: grid+go+8 gathers -> 8 REDs + memset (no ggrid)", g,go,vol,gv,gg,n,true);
    runb<1,2>("bwd: grid+go+8 gathers -> ggrid + 8 REDs + memset", g,go,vol,gv,gg,n,true);
    runb<4,2>("bwd: grid+go+8 gathers -> ggrid + 8 REDs + memset", g,go,vol,gv,gg,n,true);
  }
  // plain copy for reference
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaMemcpyAsync(vol, out, n*4, cudaMemcpyDeviceToDevice)); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); for(int i=0;i<10;i++) CK(cudaMemcpyAsync(g, g + n/2, n*8, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); ms/=10; printf("cudaMemcpy D2D 128 MiB: %.1f us  %.1f GB/s (r+w)\n", ms*1e3, (double)n*16/ms/1e6);
  return 0;
}
