// Microbenchmark 3 (round 2): does streaming the index / parameter arrays from DRAM (instead of a warm L2) slow the random
// gathers down?  8 M random 8-byte gathers from an 8 MB table; per gather a coalesced 4-byte index and an 8-byte parameter
// are streamed and an 8-byte result per 8 gathers is written.  "warm": the same 96 MB of streams every repetition (they
// stay in the 126 MB L2); "cold": every repetition reads a different 96 MB slice of a 1.5 GB pool.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cold_gather_bench cold_gather_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int ILP, bool PARAM>
__global__ void k(const int* __restrict__ idx, const double* __restrict__ par, const double* __restrict__ tab, double* __restrict__ out, long long n) {
  long long base = ((long long)blockIdx.x * blockDim.x) * ILP + threadIdx.x;
  int ii[ILP]; double pp[ILP];
#pragma unroll
  for (int q = 0; q < ILP; ++q) { long long j = base + (long long)q * blockDim.x; ii[q] = j < n ? __ldcs(idx + j) : 0; pp[q] = (PARAM && j < n) ? __ldcs(par + j) : 1.0; }
  double acc = 0.0;
#pragma unroll
  for (int q = 0; q < ILP; ++q) acc += pp[q] * tab[ii[q]];
  out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  const long long n = 8000000, tabn = 1000000;
  const int NSL = 16;
  int* idx; double *par, *tab, *out;
  CK(cudaMalloc(&idx, n * 4 * NSL)); CK(cudaMalloc(&par, n * 8 * NSL)); CK(cudaMalloc(&tab, tabn * 8)); CK(cudaMalloc(&out, n * 8));
  std::vector<int> h(n);
  unsigned long long s = 88172645463325252ULL;
  for (long long i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % tabn); }
  for (int c = 0; c < NSL; ++c) CK(cudaMemcpy(idx + c * n, h.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(par, 0, n * 8 * NSL)); CK(cudaMemset(tab, 0, tabn * 8));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int param = 0; param < 2; ++param)
    for (int cold = 0; cold < 2; ++cold)
      for (int block : {128, 256}) {
        const int ILP = 8;
        int grid = (int)((n + (long long)block * ILP - 1) / ((long long)block * ILP));
        auto launch = [&](int r) {
          const int c = cold ? r % NSL : 0;
          if (param) k<ILP, true><<<grid, block>>>(idx + c * n, par + c * n, tab, out, n);
          else k<ILP, false><<<grid, block>>>(idx + c * n, par + c * n, tab, out, n);
        };
        for (int w = 0; w < 4; ++w) launch(w);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(a);
        const int reps = 32;
        for (int r = 0; r < reps; ++r) launch(r);
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("8 M random gathers (8 MB table) + 4 B index%s per gather, streams %s, block %d: %.1f us = %.0f Ggather/s = %.2f per clk per SM\n",
               param ? " + 8 B parameter" : "", cold ? "COLD (DRAM)" : "warm (L2)", block, ms / reps * 1e3, n / (ms / reps * 1e6), n / (ms / reps * 1e-3) / (148 * 1.965e9));
      }
  return 0;
}
