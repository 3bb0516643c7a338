// Microbenchmark 2 (round 2): what can lower the cost of the 2E random neighbour reads below one L1TEX wavefront each?
//   (1) LDG gathers whose lanes are SORTED by address inside a per-CTA tile of T entries (lanes of one warp instruction
//       then share 128-byte lines: fewer wavefronts).  T = 13.5 K ... 1 M entries over the 8 MB table.
//   (2) random 8-byte LDS / STS on plain local shared memory (the cost of staging gathered values through shared memory)
//   (3) random 8-byte REMOTE stores st.shared::cluster (push instead of pull), cluster sizes 2..16
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sorted_gather_bench sorted_gather_bench.cu
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static unsigned long long rng_s = 88172645463325252ULL;
static inline unsigned long long rnd() { rng_s ^= rng_s << 13; rng_s ^= rng_s >> 7; rng_s ^= rng_s << 17; return rng_s; }

// ---- (1) LDG gather, consecutive threads = consecutive entries of the index stream ------------------------------------
template <int ILP>
__global__ void gather_kernel(const int* __restrict__ idx, const double* __restrict__ tab, double* __restrict__ out, long long n) {
  long long base = ((long long)blockIdx.x * blockDim.x) * ILP + threadIdx.x;
  double acc = 0.0;
  int ii[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) { long long j = base + (long long)k * blockDim.x; ii[k] = j < n ? idx[j] : 0; }
#pragma unroll
  for (int k = 0; k < ILP; ++k) acc += tab[ii[k]];
  out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
float time_it(F f, int reps = 20) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 3; ++w) f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int r = 0; r < reps; ++r) f();
  cudaEventRecord(b); CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

// ---- (2) local shared memory: random LDS.64 / STS.64 ------------------------------------------------------------------
template <int ILP, int STORE>
__global__ void smem_kernel(const unsigned* __restrict__ idx, double* __restrict__ out, long long cnt, int per_cta) {
  extern __shared__ double tab[];
  for (int i = threadIdx.x; i < per_cta; i += blockDim.x) tab[i] = (double)(i & 1023);
  __syncthreads();
  const unsigned* my = idx + (long long)blockIdx.x * cnt;
  double acc = 0.0;
  for (long long j0 = threadIdx.x; j0 < cnt; j0 += (long long)blockDim.x * ILP) {
    unsigned e[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) { long long j = j0 + (long long)k * blockDim.x; e[k] = j < cnt ? my[j] : 0u; }
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      if (STORE) tab[e[k]] = (double)e[k];
      else acc += tab[e[k]];
    }
  }
  __syncthreads();
  if (STORE) acc = tab[threadIdx.x];
  out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---- (3) remote stores into the shared memory of the cluster ------------------------------------------------------------
template <int ILP>
__global__ void dsmem_store_kernel(const unsigned* __restrict__ idx, double* __restrict__ out, long long cnt, int per_cta) {
  extern __shared__ double tab[];
  cg::cluster_group cl = cg::this_cluster();
  for (int i = threadIdx.x; i < per_cta; i += blockDim.x) tab[i] = 0.0;
  cl.sync();
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(tab);
  const unsigned* my = idx + (long long)blockIdx.x * cnt;
  for (long long j0 = threadIdx.x; j0 < cnt; j0 += (long long)blockDim.x * ILP) {
    unsigned e[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) { long long j = j0 + (long long)k * blockDim.x; e[k] = j < cnt ? my[j] : 0u; }
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      unsigned r = e[k] >> 16, off = e[k] & 0xffffu, ra;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(sbase + off * 8u), "r"(r));
      asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"((double)off) : "memory");
    }
  }
  cl.sync();
  out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = tab[threadIdx.x];
}

int main() {
  const long long n = 8000000;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  const double clk = prop.clockRate * 1e3;
  printf("device %s, %d SMs, %.0f MHz\n", prop.name, nsm, clk / 1e6);
  double* out; CK(cudaMalloc(&out, (size_t)n * 8));

  // (1) sorted tiles
  {
    const long long tabn = 1000000;
    double* tab; int* idx; CK(cudaMalloc(&tab, tabn * 8)); CK(cudaMalloc(&idx, n * 4)); CK(cudaMemset(tab, 0, tabn * 8));
    std::vector<int> h(n);
    for (long long T : {0LL, 512LL, 13500LL, 27000LL, 54000LL, 108000LL, 216000LL, 1000000LL, 8000000LL}) {
      for (long long i = 0; i < n; ++i) h[i] = (int)(rnd() % tabn);
      if (T > 0) for (long long a = 0; a < n; a += T) std::sort(h.begin() + a, h.begin() + std::min(n, a + T));
      // distinct 128-byte lines per 32 consecutive entries (= wavefronts per warp instruction)
      double lines = 0; long long groups = 0;
      for (long long a = 0; a + 32 <= n; a += 32) {
        int cntl = 0; int last = -1;
        std::vector<int> l(32); for (int k = 0; k < 32; ++k) l[k] = h[a + k] >> 4;
        std::sort(l.begin(), l.end());
        for (int k = 0; k < 32; ++k) if (l[k] != last) { ++cntl; last = l[k]; }
        lines += cntl; ++groups;
      }
      CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
      for (int block : {128, 256}) {
        float t4 = time_it([&] { gather_kernel<4><<<(int)((n + block * 4 - 1) / (block * 4)), block>>>(idx, tab, out, n); });
        float t8 = time_it([&] { gather_kernel<8><<<(int)((n + block * 8 - 1) / (block * 8)), block>>>(idx, tab, out, n); });
        float t = std::min(t4, t8);
        printf("LDG gather, tiles of %8lld entries sorted by address: %.2f lines per warp instruction | block %3d: ILP4 %.1f us ILP8 %.1f us => %.0f Ggather/s, %.2f per clk per SM\n",
               T, lines / groups, block, t4 * 1e3, t8 * 1e3, n / (t * 1e6), n / (t * 1e-3) / (nsm * clk));
      }
    }
    cudaFree(tab); cudaFree(idx);
  }
  // (2) local shared memory
  {
    const int per_cta = 16384;   // 128 KB
    std::vector<unsigned> h(n);
    for (long long i = 0; i < n; ++i) h[i] = (unsigned)(rnd() % per_cta);
    unsigned* idx; CK(cudaMalloc(&idx, n * 4)); CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
    const long long cnt = n / nsm;
    auto k0 = smem_kernel<8, 0>; auto k1 = smem_kernel<8, 1>;
    CK(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, per_cta * 8));
    CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, per_cta * 8));
    for (int block : {512, 1024}) {
      float tl = time_it([&] { k0<<<nsm, block, per_cta * 8>>>(idx, out, cnt, per_cta); });
      float ts = time_it([&] { k1<<<nsm, block, per_cta * 8>>>(idx, out, cnt, per_cta); });
      printf("local shared memory 128 KB, random 8-byte, block %4d: LDS %.1f us = %.2f per clk per SM | STS %.1f us = %.2f per clk per SM  (incl. coalesced 4-byte index stream)\n",
             block, tl * 1e3, n / (tl * 1e-3) / (nsm * clk), ts * 1e3, n / (ts * 1e-3) / (nsm * clk));
    }
    // sorted-ish (ascending within a warp: the staging writes of the sorted gather are NOT sorted; the reads by row are)
    cudaFree(idx);
  }
  // (3) remote stores
  {
    const int per_cta = 16384;
    for (int csize : {1, 2, 4, 8, 16}) {
      const long long tabn = (long long)per_cta * csize;
      std::vector<unsigned> h(n);
      for (long long i = 0; i < n; ++i) { unsigned long long t = rnd() % (unsigned long long)tabn; h[i] = (unsigned)((t / per_cta) << 16 | (t % per_cta)); }
      unsigned* idx; CK(cudaMalloc(&idx, n * 4)); CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
      auto kern = dsmem_store_kernel<8>;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, per_cta * 8));
      if (csize > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      cudaLaunchConfig_t cfg = {};
      int grid = nsm / csize * csize;
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = per_cta * 8;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int maxcl = 0; cudaOccupancyMaxActiveClusters(&maxcl, kern, &cfg);
      if (maxcl * csize < grid) { grid = maxcl * csize; cfg.gridDim = dim3(grid); }
      long long cnt = n / grid;
      float t = time_it([&] { CK(cudaLaunchKernelEx(&cfg, kern, (const unsigned*)idx, out, cnt, per_cta)); });
      printf("remote stores st.shared::cluster.f64, cluster %2d, grid %3d: %.1f us = %.0f Gstore/s = %.2f per clk per SM (148)\n", csize, grid, t * 1e3, n / (t * 1e6), n / (t * 1e-3) / (148 * clk));
      cudaFree(idx);
    }
  }
  return 0;
}
