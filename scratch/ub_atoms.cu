// microbench: shared-memory atomic throughput on sm_100a (scratch, not product)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p){return (uint32_t)__cvta_generic_to_shared(p);}
template<int MODE>
__global__ void __launch_bounds__(1024,1) k(const uint32_t* idx, int n_iter, uint32_t* out, long long* cyc, int cells, int active_mod){
  extern __shared__ uint32_t tab[];
  for(int i=threadIdx.x;i<cells;i+=blockDim.x) tab[i]=0;
  __syncthreads();
  const uint32_t base=smem_u32(tab);
  // each thread preloads 16 indices
  uint32_t ix[16];
  for(int j=0;j<16;j++) ix[j]=idx[(blockIdx.x*1024+threadIdx.x)*16+j]%cells;
  bool act = (threadIdx.x % active_mod)==0;
  uint32_t acc=0;
  __syncthreads();
  long long t0=clock64();
  for(int it=0;it<n_iter;it++){
#pragma unroll
    for(int j=0;j<16;j++){
      uint32_t a=base+ix[j]*4;
      if(MODE==0){ // atom with return
        uint32_t old; if(act){asm volatile("atom.shared.add.u32 %0,[%1],%2;":"=r"(old):"r"(a),"r"(720720u):"memory"); acc+=old;}
      } else if(MODE==1){ // red
        if(act) asm volatile("red.shared.add.u32 [%0],%1;"::"r"(a),"r"(720720u):"memory");
      } else if(MODE==2){ // plain lds+sts (racy) for reference
        uint32_t v; asm volatile("ld.shared.u32 %0,[%1];":"=r"(v):"r"(a)); asm volatile("st.shared.u32 [%0],%1;"::"r"(a),"r"(v+1):"memory");
      } else if(MODE==3){ // lds only gather
        uint32_t v; asm volatile("ld.shared.u32 %0,[%1];":"=r"(v):"r"(a)); acc+=v;
      } else if(MODE==4){ // lds u16 gather
        unsigned short v; asm volatile("ld.shared.u16 %0,[%1];":"=h"(v):"r"(base+ix[j]*2)); acc+=v;
      }
    }
  }
  __syncthreads();
  long long t1=clock64();
  if(threadIdx.x==0) cyc[blockIdx.x]=t1-t0;
  out[blockIdx.x*1024+threadIdx.x]=acc+tab[threadIdx.x%cells];
}
int main(){
  const int G=148; int n_iter=200;
  uint32_t *idx,*out; long long* cyc;
  size_t n=(size_t)G*1024*16;
  uint32_t* h=(uint32_t*)malloc(n*4);
  cudaMalloc(&idx,n*4); cudaMalloc(&out,G*1024*4); cudaMalloc(&cyc,G*8);
  long long hc[G];
  const char* names[]={"atom(ret)","red","lds+sts","lds32 gather","lds16 gather"};
  for(int pat=0;pat<3;pat++){
    // pat0: uniform random, pat1: zipf-ish (square of uniform), pat2: all same
    uint64_t s=88172645463325252ull;
    for(size_t i=0;i<n;i++){ s^=s<<13; s^=s>>7; s^=s<<17; double u=(s>>11)*(1.0/9007199254740992.0);
      h[i]= pat==0? (uint32_t)(u*3001) : pat==1? (uint32_t)(u*u*u*u*3001) : 7; }
    cudaMemcpy(idx,h,n*4,cudaMemcpyHostToDevice);
    for(int am=1;am<=2;am++)
    for(int mode=0;mode<5;mode++){
      for(int rep=0;rep<2;rep++){
        if(mode==0) k<0><<<G,1024,3001*4>>>(idx,n_iter,out,cyc,3001,am);
        if(mode==1) k<1><<<G,1024,3001*4>>>(idx,n_iter,out,cyc,3001,am);
        if(mode==2) k<2><<<G,1024,3001*4>>>(idx,n_iter,out,cyc,3001,am);
        if(mode==3) k<3><<<G,1024,3001*4>>>(idx,n_iter,out,cyc,3001,am);
        if(mode==4) k<4><<<G,1024,3001*4>>>(idx,n_iter,out,cyc,3001,am);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(hc,cyc,G*8,cudaMemcpyDeviceToHost);
      double c=0; for(int i=0;i<G;i++) c+=hc[i]; c/=G;
      double ops=(double)n_iter*16*1024; // lane-ops per SM (incl. inactive)
      printf("pat%d active1/%d %-14s cyc/warp-instr %.2f  cyc/active-lane %.3f\n",pat,am,names[mode],c/(ops/32),c/(ops/am));
    }
  }
  printf("%s\n",cudaGetErrorString(cudaGetLastError()));
}
