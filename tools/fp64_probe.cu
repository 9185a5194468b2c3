// FP64 pipe probe for B200 (sm_100a): peak DFMA vs DMMA.8x8x4 throughput, register-resident.
// Used once to pick the roofline denominator and the GEMM inner instruction. Not product code.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int NACC>
__global__ void __launch_bounds__(1024) k_dmma(double* out, const double* in, int iters){
  double a=in[threadIdx.x&31], b=in[32+(threadIdx.x&31)];
  double c[NACC][2];
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=0;c[i][1]=0;}
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]),"+d"(c[i][1]) : "d"(a),"d"(b));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC>
__global__ void __launch_bounds__(1024) k_dfma(double* out, const double* in, int iters){
  double a=in[threadIdx.x&31], b=in[32+(threadIdx.x&31)];
  double c[NACC];
  #pragma unroll
  for(int i=0;i<NACC;i++) c[i]=i;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) c[i]=fma(a,c[i],b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// Do DMMA and DFMA share the FP64 units?  Even warps issue DMMA, odd warps DFMA (MODE 0), or every warp interleaves
// both (MODE 1).  If the mixed run takes ~max(t_dmma, t_dfma) the pipes are concurrent; ~sum means shared.
template<int MODE>
__global__ void __launch_bounds__(1024) k_mixed(double* out, const double* in, int it_mma, int it_fma){
  double a=in[threadIdx.x&31], b=in[32+(threadIdx.x&31)];
  double c[8][2]; double f[16];
  #pragma unroll
  for(int i=0;i<8;i++){c[i][0]=0;c[i][1]=0;}
  #pragma unroll
  for(int i=0;i<16;i++) f[i]=i;
  const int warp=threadIdx.x>>5;
  if(MODE==0){
    if(((warp>>2)&1)==0){
      for(int it=0;it<it_mma;it++){
        #pragma unroll
        for(int i=0;i<8;i++)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]),"+d"(c[i][1]) : "d"(a),"d"(b));
      }
    } else {
      for(int it=0;it<it_fma;it++){
        #pragma unroll
        for(int i=0;i<16;i++) f[i]=fma(a,f[i],b);
      }
    }
  } else {
    for(int it=0;it<it_mma;it++){
      #pragma unroll
      for(int i=0;i<8;i++){
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]),"+d"(c[i][1]) : "d"(a),"d"(b));
        f[2*i]=fma(a,f[2*i],b); f[2*i+1]=fma(a,f[2*i+1],b);
      }
    }
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<8;i++) s+=c[i][0]+c[i][1];
  #pragma unroll
  for(int i=0;i<16;i++) s+=f[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<typename F> float timeit(F f){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); return ms;
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  printf("device %s SMs %d clock %d kHz\n",p.name,p.multiProcessorCount,p.clockRate);
  double *in,*out; CK(cudaMalloc(&in,1024)); CK(cudaMalloc(&out,sizeof(double)*148*8*1024));
  double h[64]; for(int i=0;i<64;i++) h[i]=1e-3*i; CK(cudaMemcpy(in,h,512,cudaMemcpyHostToDevice));
  int nsm=p.multiProcessorCount;
  const int iters=20000;
  for(int threads: {128,256,512,1024}){
    for(int bps: {1,2}){
      if(threads*bps>2048) continue;
      int grid=nsm*bps;
      {float ms=timeit([&]{k_dmma<8><<<grid,threads>>>(out,in,iters);});
       double fl=(double)grid*(threads/32)*iters*8.0*512.0; printf("DMMA884 acc8 thr %4d bps %d: %.2f TFLOP/s (%.2f ms)\n",threads,bps,fl/ms*1e-9,ms);}
      {float ms=timeit([&]{k_dmma<16><<<grid,threads>>>(out,in,iters);});
       double fl=(double)grid*(threads/32)*iters*16.0*512.0; printf("DMMA884 acc16 thr %4d bps %d: %.2f TFLOP/s (%.2f ms)\n",threads,bps,fl/ms*1e-9,ms);}
      {float ms=timeit([&]{k_dfma<16><<<grid,threads>>>(out,in,iters);});
       double fl=(double)grid*threads*(double)iters*16.0*2.0; printf("DFMA    acc16 thr %4d bps %d: %.2f TFLOP/s (%.2f ms)\n",threads,bps,fl/ms*1e-9,ms);}
    }
  }

  {
    // per warp: it_mma*8 DMMA (16 clk each at peak) vs it_fma*16 DFMA (2 clk each): equal pipe time when it_fma = 4*it_mma
    const int threads=512, grid=nsm, im=20000, ifm=80000;
    float t_m=timeit([&]{k_mixed<0><<<grid,threads>>>(out,in,im,0);});
    float t_f=timeit([&]{k_mixed<0><<<grid,threads>>>(out,in,0,ifm);});
    float t_b=timeit([&]{k_mixed<0><<<grid,threads>>>(out,in,im,ifm);});
    printf("MIXED warp-specialised (8 DMMA warps + 8 DFMA warps / SM): dmma only %.2f ms, dfma only %.2f ms, both %.2f ms\n",t_m,t_f,t_b);
    float t_i=timeit([&]{k_mixed<1><<<grid,threads>>>(out,in,im,0);});
    float t_m16=timeit([&]{k_dmma<8><<<grid,threads>>>(out,in,im);});
    printf("MIXED interleaved (16 warps, 8 DMMA + 16 DFMA per iter): %.2f ms; same DMMA alone %.2f ms\n",t_i,t_m16);
  }
  CK(cudaGetLastError());
  return 0;
}
