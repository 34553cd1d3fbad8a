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
int main(){
  long long n = 64LL*64*64*64;
  float4* g; float* vol; float* out;
  CK(cudaMalloc(&g, n*16)); CK(cudaMalloc(&vol, n*4)); CK(cudaMalloc(&out, n*4));
  CK(cudaMemset(g,0,n*16)); CK(cudaMemset(vol,0,n*4));
  run<1,0>("grid->out", g,vol,out,n); run<2,0>("grid->out", g,vol,out,n); run<4,0>("grid->out", g,vol,out,n);
  run<1,1>("grid+vol->out", g,vol,out,n); run<2,1>("grid+vol->out", g,vol,out,n); run<4,1>("grid+vol->out", g,vol,out,n);
  run<1,2>("grid+8 coalesced gathers->out", g,vol,out,n); run<2,2>("grid+8 coalesced gathers->out", g,vol,out,n); run<4,2>("grid+8 coalesced gathers->out", g,vol,out,n);
  // plain copy for reference
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaMemcpyAsync(vol, out, n*4, cudaMemcpyDeviceToDevice)); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); for(int i=0;i<10;i++) CK(cudaMemcpyAsync(g, g + n/2, n*8, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); ms/=10; printf("cudaMemcpy D2D 128 MiB: %.1f us  %.1f GB/s (r+w)\n", ms*1e3, (double)n*16/ms/1e6);
  return 0;
}
