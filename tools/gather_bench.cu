// Microbenchmark: ceiling for random 8-byte gathers out of an L2-resident table on B200.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int ILP, int MODE>
__global__ void gather_kernel(const int* __restrict__ idx, const double* __restrict__ tab, double* __restrict__ out, long long n) {
  long long base = ((long long)blockIdx.x * blockDim.x) * ILP + threadIdx.x;
  double acc = 0.0;
  int ii[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) { long long j = base + (long long)k * blockDim.x; ii[k] = j < n ? idx[j] : 0; }
#pragma unroll
  for (int k = 0; k < ILP; ++k) {
    double v;
    if (MODE == 0) v = tab[ii[k]];
    else if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(tab + ii[k]));
    else asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(tab + ii[k]));
    acc += v;
  }
  out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int ILP, int MODE>
float run(const int* idx, const double* tab, double* out, long long n, int block) {
  long long per_block = (long long)block * ILP;
  int grid = (int)((n + per_block - 1) / per_block);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 3; ++w) gather_kernel<ILP, MODE><<<grid, block>>>(idx, tab, out, n);
  cudaEventRecord(a);
  const int reps = 20;
  for (int r = 0; r < reps; ++r) gather_kernel<ILP, MODE><<<grid, block>>>(idx, tab, out, n);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main(int argc, char** argv) {
  const long long n = 8000000;            // gathers per launch (= directed entries of config 2)
  for (long long tabn : {1000000LL, 16000000LL}) {   // 8 MB table (L2 resident) / 128 MB table
    for (int pattern = 0; pattern < 3; ++pattern) {
      std::vector<int> h(n);
      unsigned long long s = 88172645463325252ULL;
      for (long long i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        if (pattern == 0) h[i] = (int)(s % tabn);                                   // uniform random
        else if (pattern == 1) h[i] = (int)((i / 8 * 9973 + (s % 64)) % tabn);        // rows of 8 entries within a 512 B window
        else h[i] = (int)(i % tabn);                                                  // fully coalesced
      }
      int* idx; double *tab, *out;
      cudaMalloc(&idx, n * 4); cudaMalloc(&tab, tabn * 8); cudaMalloc(&out, n * 8);
      cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice); cudaMemset(tab, 0, tabn * 8);
      const char* pn[] = {"random", "windowed", "coalesced"};
      for (int block : {128, 256, 512}) {
        float t1 = run<1, 0>(idx, tab, out, n, block), t4 = run<4, 0>(idx, tab, out, n, block), t8 = run<8, 0>(idx, tab, out, n, block);
        float t8nc = run<8, 1>(idx, tab, out, n, block), t8cg = run<8, 2>(idx, tab, out, n, block), t16 = run<16, 0>(idx, tab, out, n, block);
        printf("table %3lld MB %-9s block %3d: ILP1 %.1f us | ILP4 %.1f | ILP8 %.1f | ILP16 %.1f | ILP8 nc.noalloc %.1f | ILP8 cg %.1f  => best %.1f Ggather/s\n",
               tabn * 8 / 1000000, pn[pattern], block, t1 * 1e3, t4 * 1e3, t8 * 1e3, t16 * 1e3, t8nc * 1e3, t8cg * 1e3,
               n / (1e6 * fminf(fminf(fminf(t1, t4), fminf(t8, t16)), fminf(t8nc, t8cg))));
      }
      cudaFree(idx); cudaFree(tab); cudaFree(out);
    }
  }
  return 0;
}
