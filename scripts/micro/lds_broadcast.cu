// Microbenchmark: shared-memory wavefronts of 64-bit / 128-bit loads with few distinct addresses per warp
// (broadcast patterns), B200.  Decides whether a "2D slice" contraction (every lane of a row reads the
// same word) is cheaper than line ownership (all lanes distinct).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 lds_broadcast.cu -o lds_broadcast && ./lds_broadcast
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double *out, int iters, int stride_words, int group) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // word index read by this lane: lanes in the same group of `group` lanes read the same word
  int w = (lane / group) * stride_words;
  if (group < 0) w = (lane % (-group)) * stride_words;       // periodic pattern: word = (lane mod g) * stride
  double acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int u = 0; u < 16; ++u) acc += sm[(w + u * 33 + it) & 4095];
    } else {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const double2 v = *reinterpret_cast<const double2 *>(&sm[((w + u * 34 + it * 2) & 4094)]);
        acc += v.x + v.y;
      }
    }
  }
  long long t1 = clock64();
  if (acc == 1.2345) out[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)(t1 - t0);
}

int main() {
  double *out; cudaMalloc(&out, 16);
  const int iters = 2000;
  struct Case { const char *name; int mode, stride, group; } cases[] = {
    {"LDS.64 all distinct, stride 1", 0, 1, 1},
    {"LDS.64 groups of 7 (5 distinct), stride 7", 0, 7, 7},
    {"LDS.64 groups of 7, stride 1", 0, 1, 7},
    {"LDS.64 groups of 2 (16 distinct), stride 1", 0, 1, 2},
    {"LDS.64 groups of 2 (16 distinct), stride 2 (32 banks twice)", 0, 2, 2},
    {"LDS.64 groups of 4 (8 distinct), stride 1", 0, 1, 4},
    {"LDS.64 all same word", 0, 0, 32},
    {"LDS.64 distinct stride 2 (2-way conflict)", 0, 2, 1},
    {"LDS.64 word = lane %  8", 0, 1, -8},
    {"LDS.64 word = lane %  7", 0, 1, -7},
    {"LDS.64 word = lane % 16", 0, 1, -16},
    {"LDS.64 word = lane %  4", 0, 1, -4},
    {"LDS.64 word = lane %  5", 0, 1, -5},
    {"LDS.64 word = lane %  9", 0, 1, -9},
    {"LDS.64 word = (lane % 8) * 9", 0, 9, -8},
    {"LDS.64 word = (lane / 8) * 7", 0, 7, 8},
    {"LDS.64 word = (lane / 8) * 8", 0, 8, 8},
    {"LDS.64 word = (lane / 8) * 9", 0, 9, 8},
    {"LDS.64 word = (lane / 16) * 9", 0, 9, 16},
    {"LDS.64 word = (lane / 6) * 7", 0, 7, 6},
    {"LDS.64 word = (lane / 10) * 9", 0, 9, 10},
    {"LDS.64 word = (lane / 5)", 0, 1, 5},
    {"LDS.64 word = (lane / 3)", 0, 1, 3},
    {"LDS.128 all distinct stride 2", 1, 2, 1},
    {"LDS.128 groups of 7, stride 8", 1, 8, 7},
    {"LDS.128 groups of 4 (8 distinct) stride 2", 1, 2, 4},
    {"LDS.128 all same", 1, 0, 32},
  };
  for (auto &c : cases) {
    for (int warps : {4, 16}) {
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        if (c.mode == 0) k<0><<<148, warps * 32, 4096 * 8>>>(out, iters, c.stride, c.group);
        else k<1><<<148, warps * 32, 4096 * 8>>>(out, iters, c.stride, c.group);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      double h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      const double lds = (double)iters * 16 * warps;          // warp-level LDS instructions per SM
      printf("%-62s warps/SM %2d: %.3f clk per warp-LDS (per SM)\n", c.name, warps, h[1] / lds);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
