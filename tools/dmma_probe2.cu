// DMMA.8x8x4 issue-pattern probe for B200 (sm_100a): does the FP64 tensor pipe keep its register-only peak when every
// DMMA brings fresh A / B operand registers, as in the row kernels' k-step (2 A fragments x 4 B fragments -> 8 DMMAs)?
// Not product code.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe2 dmma_probe2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// MODE 0: one a, one b for all 8 accumulators.  MODE 1: a[2] x b[4] (fixed registers).  MODE 2: NS register sets of
// a[2] x b[4], one per k-step (operands change every k-step).  MODE 3: operands re-read from shared memory every k-step.
template <int MODE, int NS>
__global__ void __launch_bounds__(256, 2) k(double* out, const double* in, int iters) {
  __shared__ double sh[NS * 6 * 32 + 64];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NS * 6 * 32 + 64; i += blockDim.x) sh[i] = in[i & 63] + i;
  __syncthreads();
  double a[NS][2], b[NS][4], acc[2][4][2];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
#pragma unroll
    for (int i = 0; i < 2; ++i) a[s][i] = in[(lane + 3 * s + i) & 63];
#pragma unroll
    for (int i = 0; i < 4; ++i) b[s][i] = in[(lane + 5 * s + 7 * i + 1) & 63];
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0; acc[i][j][1] = 0; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (MODE == 3) {
        const double* p = sh + ((it + s) & (NS - 1)) * 6 * 32 + lane;
        a[s][0] = p[0]; a[s][1] = p[32];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[s][j] = p[64 + 32 * j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (MODE == 0) dmma(acc[i][j][0], acc[i][j][1], a[0][0], b[0][0]);
          else if (MODE == 1) dmma(acc[i][j][0], acc[i][j][1], a[0][i], b[0][j]);
          else dmma(acc[i][j][0], acc[i][j][1], a[s][i], b[s][j]);
        }
    }
  }
  double sum = 0;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) sum += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); CK(cudaDeviceSynchronize());
  cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int nsm = p.multiProcessorCount;
  double *in, *out; CK(cudaMalloc(&in, 64 * 8)); CK(cudaMalloc(&out, (size_t)nsm * 4 * 1024 * 8));
  double h[64]; for (int i = 0; i < 64; ++i) h[i] = 1e-3 * (i + 1);
  CK(cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice));
  const int iters = 4096;
#define RUN(MODE, NS, CTAS) { float ms = timeit([&] { k<MODE, NS><<<nsm * CTAS, 256>>>(out, in, iters / NS); }); \
    double fl = (double)nsm * CTAS * 8 * (iters / NS) * NS * 8 * 512.0; \
    printf("mode %d sets %d ctas/SM %d: %.2f TFLOP/s (%.3f ms)\n", MODE, NS, CTAS, fl / ms / 1e9, ms); }
  RUN(0, 1, 2) RUN(1, 1, 2) RUN(2, 2, 2) RUN(2, 4, 2) RUN(3, 4, 2)
  RUN(0, 1, 1) RUN(1, 1, 1) RUN(2, 4, 1) RUN(3, 4, 1)
  return 0;
}
