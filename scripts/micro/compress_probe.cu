// Does L2 compute-data-compression (cuMemCreate with CU_MEM_ALLOCATION_COMP_GENERIC) help the "dst = 0, then
// red.add into it" pattern of vmult?  Zero lines compress; the probe times, on plain and on compressible memory:
//   memset(0), a streaming read of the zeros, red.add of random values into every 3rd double of the zeroed buffer,
//   and a streaming write of incompressible data.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o compress_probe compress_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s; cuGetErrorString(r_, &s); printf("%s failed: %s\n", #x, s); exit(1); } } while (0)
#define RT(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void read_kernel(const double *p, size_t n, double *out) {
  double s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += p[i];
  if (s == 12345.678) *out = s;
}
__global__ void red_kernel(double *p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (i % 3 == 0) atomicAdd(p + i, 1.0 + 1e-9 * (double)i);
    else p[i] = 0.5 + 1e-9 * (double)i;
}
__global__ void write_kernel(double *p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = 1.0 / (1.0 + (double)(i * 2654435761u % 1000003));
}

static float timed(void (*f)(double *, size_t, double *), double *p, size_t n, double *aux, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(p, n, aux); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f(p, n, aux);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}
static void do_memset(double *p, size_t n, double *) { cudaMemsetAsync(p, 0, n * 8); }
static void do_read(double *p, size_t n, double *aux) { read_kernel<<<148 * 16, 256>>>(p, n, aux); }
static void do_zero_red(double *p, size_t n, double *) { cudaMemsetAsync(p, 0, n * 8); red_kernel<<<148 * 16, 256>>>(p, n); }
static void do_write(double *p, size_t n, double *) { write_kernel<<<148 * 16, 256>>>(p, n); }

int main() {
  RT(cudaSetDevice(0));
  RT(cudaFree(0));
  CUdevice dev; CK(cuDeviceGet(&dev, 0));
  int comp = 0;
  CK(cuDeviceGetAttribute(&comp, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, dev));
  printf("generic compression supported: %d\n", comp);
  const size_t n = (size_t)148035889;          // the bench's vector
  size_t bytes = n * 8;
  double *plain; RT(cudaMalloc(&plain, bytes));
  double *aux; RT(cudaMalloc(&aux, 8));
  double *cmp = nullptr;
  if (comp) {
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
    size_t gran = 0;
    CK(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
    const size_t sz = (bytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h;
    CK(cuMemCreate(&h, sz, &prop, 0));
    CUmemAllocationProp got = {};
    CK(cuMemGetAllocationPropertiesFromHandle(&got, h));
    printf("allocation compressionType: %d (granularity %zu)\n", (int)got.allocFlags.compressionType, gran);
    CUdeviceptr va;
    CK(cuMemAddressReserve(&va, sz, 0, 0, 0));
    CK(cuMemMap(va, sz, 0, h, 0));
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CK(cuMemSetAccess(va, sz, &acc, 1));
    cmp = reinterpret_cast<double *>(va);
  }
  const double gb = bytes / 1e9;
  for (int which = 0; which < (cmp ? 2 : 1); ++which) {
    double *p = which ? cmp : plain;
    const char *name = which ? "compressible" : "plain       ";
    const float t_set = timed(do_memset, p, n, aux, 10);
    const float t_read0 = timed(do_read, p, n, aux, 10);
    const float t_zr = timed(do_zero_red, p, n, aux, 10);
    const float t_write = timed(do_write, p, n, aux, 10);
    const float t_read1 = timed(do_read, p, n, aux, 10);
    printf("%s  memset %.3f ms (%.0f GB/s)  read zeros %.3f ms (%.0f GB/s)  memset+red/store %.3f ms  write data %.3f ms (%.0f GB/s)  read data %.3f ms (%.0f GB/s)\n",
           name, t_set, gb / t_set * 1e3, t_read0, gb / t_read0 * 1e3, t_zr, t_write, gb / t_write * 1e3, t_read1, gb / t_read1 * 1e3);
  }
  return 0;
}
