// Microbenchmark: does the size of the L1 data cache (= what the shared-memory carve-out leaves of the SM's 256 KB) bound the
// rate of random 8-byte gathers from an L2-resident table?  Same gather kernel as tools/gather_bench.cu (8 M gathers, 8 MB
// table, 128-thread blocks, ILP 8), run (a) with the preferred carve-out forced to 0 / 25 / 50 / 75 / 100 % while the kernel
// uses no shared memory, (b) with 0 .. 14 KB of dynamic shared memory per block (what a panel-staging kernel allocates),
// (c) with ld.global.nc.L1::no_allocate gathers under the same carve-outs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_capacity_gather_bench l1_capacity_gather_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int BLOCK = 128;
template <int ILP, int MODE>
__global__ void __launch_bounds__(BLOCK) gather_kernel(const int* __restrict__ idx, const double* __restrict__ tab, double* __restrict__ out, long long n) {
  extern __shared__ double dyn[];
  const long long base = ((long long)blockIdx.x * BLOCK) * ILP + threadIdx.x;
  double acc = 0.0;
  int ii[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) { const long long j = base + (long long)k * BLOCK; ii[k] = j < n ? idx[j] : 0; }
#pragma unroll
  for (int k = 0; k < ILP; ++k) {
    double v;
    if (MODE == 0) v = tab[ii[k]];
    else asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(tab + ii[k]));
    acc += v;
  }
  if (acc == 1.2345e300) dyn[threadIdx.x] = acc;     // keeps the dynamic allocation referenced
  out[(long long)blockIdx.x * BLOCK + threadIdx.x] = acc;
}

template <int ILP, int MODE>
float run(const int* idx, const double* tab, double* out, long long n, int carveout, int dyn_bytes) {
  auto k = gather_kernel<ILP, MODE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
  const int grid = (int)((n + (long long)BLOCK * ILP - 1) / ((long long)BLOCK * ILP));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 3; ++w) k<<<grid, BLOCK, dyn_bytes>>>(idx, tab, out, n);
  cudaEventRecord(a);
  const int reps = 20;
  for (int r = 0; r < reps; ++r) k<<<grid, BLOCK, dyn_bytes>>>(idx, tab, out, n);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  return ms / reps * 1e3f;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("device %s, %d SMs, %d MHz, shared memory per SM %zu KB\n", prop.name, prop.multiProcessorCount, prop.clockRate / 1000, prop.sharedMemPerMultiprocessor / 1024);
  const long long n = 8000000, tabn = 1000000;
  std::vector<int> h(n);
  unsigned long long s = 88172645463325252ULL;
  for (long long i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % tabn); }
  int* idx; double *tab, *out;
  cudaMalloc(&idx, n * 4); cudaMalloc(&tab, tabn * 8); cudaMalloc(&out, n * 8);
  cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice); cudaMemset(tab, 0, tabn * 8);
  const double clk = prop.clockRate * 1e3;
  auto rate = [&](float us) { return n / (us * 1e-6) / clk / prop.multiProcessorCount; };
  for (int carve : {0, 25, 50, 75, 100}) {
    const float a8 = run<8, 0>(idx, tab, out, n, carve, 0), a4 = run<4, 0>(idx, tab, out, n, carve, 0), b8 = run<8, 1>(idx, tab, out, n, carve, 0);
    printf("preferred carve-out %3d %%, no shared memory used : ld.global ILP4 %5.1f us (%.2f per clk per SM) ILP8 %5.1f us (%.2f) | nc.L1::no_allocate ILP8 %5.1f us (%.2f)\n",
           carve, a4, rate(a4), a8, rate(a8), b8, rate(b8));
  }
  for (int kb : {0, 2, 4, 6, 8, 10, 12, 14}) {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gather_kernel<8, 0>, BLOCK, kb * 1024);
    const float a8 = run<8, 0>(idx, tab, out, n, -1, kb * 1024), b8 = run<8, 1>(idx, tab, out, n, -1, kb * 1024);
    printf("dynamic shared memory %2d KB per block (%2d blocks = %3d KB per SM), default carve-out : ld.global ILP8 %5.1f us (%.2f per clk per SM) | nc.L1::no_allocate ILP8 %5.1f us (%.2f)\n",
           kb, nb, nb * kb, a8, rate(a8), b8, rate(b8));
  }
  return 0;
}
