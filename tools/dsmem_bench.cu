// Microbenchmark: where can 8 M random 8-byte reads of an 8 MB table be served fastest on B200?
//   (a) LDG through L1TEX from the L2-resident table (the round-1 kernels; see gather_bench.cu)
//   (b) ld.shared::cluster from a table column-blocked over the shared memory of a thread-block cluster (DSMEM),
//       cluster sizes 1 (plain local shared memory), 2, 4, 8, 16
//   (c) both at once (are the two paths additive?)
// Every variant reads a coalesced 4-byte index stream and writes one 8-byte partial sum per thread, like gather_bench.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_bench dsmem_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ double ld_cluster(unsigned saddr, unsigned rank) {
  unsigned ra; double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(saddr), "r"(rank));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra));
  return v;
}

// idx entries: (owner CTA rank << 16) | offset (doubles) inside that CTA's slice.  Each CTA walks cnt entries.
// FRAC_LDG of every 8 entries (0..8) go to the global table instead (additivity test).
template <int ILP, int LDG8>
__global__ void dsmem_gather(const unsigned* __restrict__ idx, const double* __restrict__ gtab, double* __restrict__ out,
                             long long cnt, int per_cta, int csize) {
  extern __shared__ double tab[];
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank();
  for (int i = threadIdx.x; i < per_cta; i += blockDim.x) tab[i] = gtab[(size_t)rank * per_cta + i];
  cl.sync();
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(tab);
  const unsigned* my = idx + (long long)blockIdx.x * cnt;
  double acc = 0.0;
  for (long long j0 = threadIdx.x; j0 < cnt; j0 += (long long)blockDim.x * ILP) {
    unsigned e[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) { long long j = j0 + (long long)k * blockDim.x; e[k] = j < cnt ? my[j] : 0u; }
    double v[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      unsigned r = e[k] >> 16, off = e[k] & 0xffffu;
      if (k < LDG8) v[k] = gtab[(size_t)r * per_cta + off];
      else v[k] = ld_cluster(sbase + off * 8u, r);
    }
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc += v[k];
  }
  out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = acc;
  cl.sync();   // nobody leaves while a peer may still read its slice
}

template <int ILP, int LDG8>
float run(int csize, int nsm_use, int block, int per_cta, const unsigned* idx, const double* gtab, double* out, long long n, double* chk) {
  int grid = nsm_use / csize * csize;
  long long cnt = n / grid;
  size_t smem = (size_t)per_cta * 8;
  auto kern = dsmem_gather<ILP, LDG8>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (csize > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int maxcl = -1;
  cudaOccupancyMaxActiveClusters(&maxcl, kern, &cfg);
  if (maxcl * csize < grid) {   // not all clusters co-resident: shrink the grid to what fits (keeps the one-wave model)
    grid = maxcl * csize; cfg.gridDim = dim3(grid); cnt = n / grid;
  }
  if (grid == 0) return -1.f;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 3; ++w) CK(cudaLaunchKernelEx(&cfg, kern, idx, gtab, out, cnt, per_cta, csize));
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  const int reps = 20;
  for (int r = 0; r < reps; ++r) CK(cudaLaunchKernelEx(&cfg, kern, idx, gtab, out, cnt, per_cta, csize));
  cudaEventRecord(b); CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b);
  std::vector<double> h((size_t)grid * block);
  CK(cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost));
  double s = 0; for (double x : h) s += x; *chk = s;
  printf("  [grid %d = %d clusters (max active %d), %lld entries/CTA]", grid, grid / csize, maxcl, cnt);
  return ms / reps;
}

// fill-only variant: how long does loading the column block into shared memory take (no gathers)
int main(int argc, char** argv) {
  const long long n = 8000000;
  int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  const int nsm = prop.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, (size_t)nsm * 1024 * 8));
  for (int per_cta : {16384, 25600}) {     // 128 KB and 200 KB of table per CTA
    for (int csize : {1, 2, 4, 8, 16}) {
      long long tabn = (long long)per_cta * csize;
      std::vector<unsigned> h(n);
      std::vector<double> ht(tabn);
      for (long long i = 0; i < tabn; ++i) ht[i] = (double)(i % 1000);
      unsigned long long s = 88172645463325252ULL;
      for (long long i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        unsigned long long t = s % (unsigned long long)tabn;
        h[i] = (unsigned)((t / per_cta) << 16 | (t % per_cta));
      }
      unsigned* idx; double* gtab;
      CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&gtab, tabn * 8));
      CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(gtab, ht.data(), tabn * 8, cudaMemcpyHostToDevice));
      for (int block : {512, 1024}) {
        double c;
        float t;
        t = run<4, 0>(csize, nsm, block, per_cta, idx, gtab, out, n, &c);
        printf(" slice %3d KB cluster %2d block %4d ILP4 all-DSMEM : %.1f us  %.0f Ggather/s  %.2f per clk per SM (148)  chk %.0f\n", per_cta * 8 / 1024, csize, block, t * 1e3, n / (t * 1e6), n / (t * 1e-3) / (148.0 * prop.clockRate * 1e3), c);
        t = run<8, 0>(csize, nsm, block, per_cta, idx, gtab, out, n, &c);
        printf(" slice %3d KB cluster %2d block %4d ILP8 all-DSMEM : %.1f us  %.0f Ggather/s  %.2f per clk per SM (148)  chk %.0f\n", per_cta * 8 / 1024, csize, block, t * 1e3, n / (t * 1e6), n / (t * 1e-3) / (148.0 * prop.clockRate * 1e3), c);
        t = run<8, 2>(csize, nsm, block, per_cta, idx, gtab, out, n, &c);
        printf(" slice %3d KB cluster %2d block %4d ILP8 2/8 LDG   : %.1f us  %.0f Ggather/s  %.2f per clk per SM (148)  chk %.0f\n", per_cta * 8 / 1024, csize, block, t * 1e3, n / (t * 1e6), n / (t * 1e-3) / (148.0 * prop.clockRate * 1e3), c);
        t = run<8, 4>(csize, nsm, block, per_cta, idx, gtab, out, n, &c);
        printf(" slice %3d KB cluster %2d block %4d ILP8 4/8 LDG   : %.1f us  %.0f Ggather/s  %.2f per clk per SM (148)  chk %.0f\n", per_cta * 8 / 1024, csize, block, t * 1e3, n / (t * 1e6), n / (t * 1e-3) / (148.0 * prop.clockRate * 1e3), c);
        t = run<8, 8>(csize, nsm, block, per_cta, idx, gtab, out, n, &c);
        printf(" slice %3d KB cluster %2d block %4d ILP8 all LDG   : %.1f us  %.0f Ggather/s  %.2f per clk per SM (148)  chk %.0f\n", per_cta * 8 / 1024, csize, block, t * 1e3, n / (t * 1e6), n / (t * 1e-3) / (148.0 * prop.clockRate * 1e3), c);
      }
      cudaFree(idx); cudaFree(gtab);
    }
  }
  return 0;
}
