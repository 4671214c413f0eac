#include <cstdio>
#include <vector>
#include <cstdlib>
#include "../../eao-fusion_b200/csrc/orb_kernels.cuh"
#include "../../oracle/cv_primitives.h"
__device__ __forceinline__ int arc_naive(const uint8_t* p){
  const int v=p[0]; int d[16];
#pragma unroll
  for(int k=0;k<16;++k) d[k]=(int)p[RING_OFF(k,FAST_TP)]-v;
  int best=-256;
#pragma unroll
  for(int k=0;k<16;++k){ int mn=d[k],mx=d[k];
#pragma unroll
    for(int j=1;j<9;++j){ mn=min(mn,d[(k+j)&15]); mx=max(mx,d[(k+j)&15]); }
    best=max(best,max(mn,-mx)); }
  return best; }
// packed: lo half = d, hi half = -d ; per-halfword signed min/max
__device__ __forceinline__ int arc_packed(const uint8_t* p){
  const int v=p[0]; unsigned w[16];
#pragma unroll
  for(int k=0;k<16;++k){ int d=(int)p[RING_OFF(k,FAST_TP)]-v; w[k]=((unsigned)d&0xffffu)|((unsigned)(-d)<<16); }
  unsigned m2[16],m4[16];
#pragma unroll
  for(int k=0;k<16;++k) m2[k]=__vmins2(w[k],w[(k+1)&15]);
#pragma unroll
  for(int k=0;k<16;++k) m4[k]=__vmins2(m2[k],m2[(k+2)&15]);
  unsigned best=0x80008000u;
#pragma unroll
  for(int k=0;k<16;++k){ unsigned m9=__vimin3_s16x2(m4[k],m4[(k+4)&15],w[(k+8)&15]); best=__vmaxs2(best,m9); }
  int lo=(short)(best&0xffff), hi=(short)(best>>16);
  return max(lo,hi); }
template<int V> __global__ void k(const unsigned char* in, int* out, int n){
  __shared__ unsigned char t[72*8];
  for(int it=0; it<n; ++it){
    for(int i=threadIdx.x;i<72*8;i+=blockDim.x) t[i]=in[it*72*8+i];
    __syncthreads();
    if(threadIdx.x<40){ const uint8_t* p=t+3*72+10+threadIdx.x; out[it*40+threadIdx.x]= V==0? eaof::arc_best(p): V==1? arc_naive(p): arc_packed(p);} 
    __syncthreads();
  }
}
int main(){ int n=200; std::vector<unsigned char> h(n*72*8); srand(1); for(size_t i=0;i<h.size();i++) h[i]= (i/ (72*8))&1 ? rand()&255 : 100+rand()%60;
 unsigned char* d; int* o; cudaMalloc(&d,h.size()); cudaMalloc(&o,n*40*4); cudaMemcpy(d,h.data(),h.size(),cudaMemcpyHostToDevice);
 for(int v=0;v<3;v++){ if(v==0) k<0><<<1,64>>>(d,o,n); else if(v==1) k<1><<<1,64>>>(d,o,n); else k<2><<<1,64>>>(d,o,n);
 printf("%s\n",cudaGetErrorString(cudaDeviceSynchronize()));
 std::vector<int> ho(n*40); cudaMemcpy(ho.data(),o,n*40*4,cudaMemcpyDeviceToHost);
 int bad=0; for(int it=0;it<n;it++)for(int j=0;j<40;j++){ int b=cvprim::fast_arc_best(h.data()+it*72*8+3*72+10+j,72); if(b!=ho[it*40+j]){ if(bad<3)printf("gpu %d host %d\n",ho[it*40+j],b); bad++; } }
 printf("variant %d: bad %d of %d\n",v,bad,n*40); } }
