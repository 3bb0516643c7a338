// Microbenchmark (2 GPUs, one process): how fast can SM-issued stores push a packed halo into a peer's memory over NVLink?
//   * copy engine (cudaMemcpyPeerAsync)            * kernel, coalesced 8-byte stores (the publish blocks of the RHS kernels)
//   * kernel, 16-byte stores                       * kernel, gather through an index (98 % dense, ascending) + 8-byte stores
//   * unaligned destination (8-byte aligned only)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o p2p_store_bench p2p_store_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int ILP>
__global__ void push8(const double* __restrict__ src, double* __restrict__ dst, long long n) {
  const long long nt = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (ILP - 1) * nt < n; i += ILP * nt) {
    double v[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) v[k] = src[i + k * nt];
#pragma unroll
    for (int k = 0; k < ILP; ++k) dst[i + k * nt] = v[k];
  }
  for (; i < n; i += nt) dst[i] = src[i];
}
__global__ void push16(const double2* __restrict__ src, double2* __restrict__ dst, long long n2) {
  const long long nt = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += nt) dst[i] = src[i];
}
template <int ILP>
__global__ void push_idx(const int* __restrict__ idx, const double* __restrict__ src, double* __restrict__ dst, long long n) {
  const long long nt = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (ILP - 1) * nt < n; i += ILP * nt) {
    int j[ILP]; double v[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) j[k] = idx[i + k * nt];
#pragma unroll
    for (int k = 0; k < ILP; ++k) v[k] = src[j[k]];
#pragma unroll
    for (int k = 0; k < ILP; ++k) dst[i + k * nt] = v[k];
  }
  for (; i < n; i += nt) dst[i] = src[idx[i]];
}

int main() {
  int nd = 0; CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
  CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0));
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
  for (long long mb : {8LL, 40LL, 300LL}) {
    const long long n = mb * 1000000 / 8;
    double *src, *dst; int* idx;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&dst, (n + 16) * 8));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&src, (n * 51 / 50 + 16) * 8)); CK(cudaMalloc(&idx, n * 4));
    std::vector<int> h(n);
    for (long long i = 0; i < n; ++i) h[i] = (int)(i + i / 50);      // 98 % dense, ascending
    CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto timeit = [&](const char* name, auto f) {
      for (int w = 0; w < 3; ++w) f();
      CK(cudaDeviceSynchronize());
      cudaEventRecord(a);
      const int reps = 20;
      for (int r = 0; r < reps; ++r) f();
      cudaEventRecord(b); CK(cudaEventSynchronize(b));
      float ms; cudaEventElapsedTime(&ms, a, b);
      printf("%4lld MB  %-46s %8.1f us  %7.1f GB/s\n", mb, name, ms / reps * 1e3, mb * 1e6 / (ms / reps * 1e-3) / 1e9);
    };
    timeit("copy engine cudaMemcpyPeerAsync", [&] { CK(cudaMemcpyPeerAsync(dst, 1, src, 0, n * 8, 0)); });
    for (int grid : {592, 1184, 2368, 4736}) {
      char nm[96];
      snprintf(nm, sizeof nm, "8-byte stores ILP4, grid %d x 128", grid);
      timeit(nm, [&] { push8<4><<<grid, 128>>>(src, dst, n); });
    }
    timeit("8-byte stores ILP8, grid 2368 x 128", [&] { push8<8><<<2368, 128>>>(src, dst, n); });
    timeit("8-byte stores ILP4, grid 2368 x 128, dst + 8 B", [&] { push8<4><<<2368, 128>>>(src, dst + 1, n); });
    timeit("16-byte stores, grid 2368 x 128", [&] { push16<<<2368, 128>>>((const double2*)src, (double2*)dst, n / 2); });
    timeit("index + gather + 8-byte stores ILP4, grid 2368", [&] { push_idx<4><<<2368, 128>>>(idx, src, dst, n); });
    timeit("index + gather + 8-byte stores ILP8, grid 2368", [&] { push_idx<8><<<2368, 128>>>(idx, src, dst, n); });
    CK(cudaFree(src)); CK(cudaFree(idx)); CK(cudaSetDevice(1)); CK(cudaFree(dst)); CK(cudaSetDevice(0));
  }
  return 0;
}
