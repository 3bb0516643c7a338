// Microbenchmark: can 8 M random 8-byte gathers out of an L2-resident 8 MB table leave the LSU tag stage (0.84 per clock per
// SM, profiles/r02a_gather_bench_raw.txt) by taking another path into the SM?
//   ldg     ld.global through the LSU (the baseline)
//   tex     tex1Dfetch<int2> through the TEX pipe of L1TEX (texture object over the linear table)
//   ldgsts  cp.async 8 B (LDGSTS) into shared memory, then a conflict-free LDS
//   bulk    cp.async.bulk 16 B (UBLKCP, the TMA engine) per thread into shared memory, mbarrier completion
//   mixes   alternate gathers over two paths
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o path_gather_bench path_gather_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int BLOCK = 128;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// MODE 0 ldg, 1 tex, 2 ldg+tex alternating, 3 ldgsts, 4 ldgsts + ldg alternating
template <int ILP, int MODE>
__global__ void __launch_bounds__(BLOCK) gather_kernel(const int* __restrict__ idx, const double* __restrict__ tab, cudaTextureObject_t tex,
                                                       double* __restrict__ out, long long n) {
  __shared__ double panel[(MODE >= 3) ? ILP * BLOCK : 1];
  const long long base = ((long long)blockIdx.x * BLOCK) * ILP + threadIdx.x;
  double acc = 0.0;
  int ii[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) { const long long j = base + (long long)k * BLOCK; ii[k] = j < n ? idx[j] : 0; }
  if constexpr (MODE <= 2) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      double v;
      if (MODE == 0 || (MODE == 2 && (k & 1))) v = tab[ii[k]];
      else { const int2 w = tex1Dfetch<int2>(tex, ii[k]); v = __hiloint2double(w.y, w.x); }
      acc += v;
    }
  } else {
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      if (MODE == 3 || !(k & 1))
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(&panel[k * BLOCK + threadIdx.x])), "l"(tab + ii[k]) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if constexpr (MODE == 4) {
#pragma unroll
      for (int k = 1; k < ILP; k += 2) acc += tab[ii[k]];
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int k = 0; k < ILP; ++k)
      if (MODE == 3 || !(k & 1)) acc += panel[k * BLOCK + threadIdx.x];
  }
  out[(long long)blockIdx.x * BLOCK + threadIdx.x] = acc;
}

// TMA path: every thread issues cp.async.bulk of the aligned 16-byte pair that holds its element; one mbarrier per block.
// LDGMIX: of ILP gathers, every second one goes through LDG instead.
template <int ILP, bool LDGMIX>
__global__ void __launch_bounds__(BLOCK) bulk_kernel(const int* __restrict__ idx, const double* __restrict__ tab, double* __restrict__ out, long long n) {
  __shared__ __align__(16) double2 panel[ILP * BLOCK];
  __shared__ __align__(8) unsigned long long bar;
  const long long base = ((long long)blockIdx.x * BLOCK) * ILP + threadIdx.x;
  int ii[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) { const long long j = base + (long long)k * BLOCK; ii[k] = j < n ? idx[j] : 0; }
  constexpr int NB = LDGMIX ? (ILP + 1) / 2 : ILP;       // bulk copies per thread
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(NB * BLOCK * 16) : "memory");
  __syncthreads();
#pragma unroll
  for (int k = 0; k < ILP; ++k) {
    if (!LDGMIX || !(k & 1)) {
      const double* src = tab + (ii[k] & ~1);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                   ::"r"(smem_u32(&panel[k * BLOCK + threadIdx.x])), "l"(src), "r"(smem_u32(&bar)) : "memory");
    }
  }
  double acc = 0.0;
  if constexpr (LDGMIX) {
#pragma unroll
    for (int k = 1; k < ILP; k += 2) acc += tab[ii[k]];
  }
  {   // wait for phase 0
    unsigned done = 0;
    while (!done) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
  }
#pragma unroll
  for (int k = 0; k < ILP; ++k) {
    if (!LDGMIX || !(k & 1)) {
      const double2 w = panel[k * BLOCK + threadIdx.x];
      acc += (ii[k] & 1) ? w.y : w.x;
    }
  }
  out[(long long)blockIdx.x * BLOCK + threadIdx.x] = acc;
}

template <class F>
float time_it(F launch) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 3; ++w) launch();
  cudaEventRecord(a);
  const int reps = 20;
  for (int r = 0; r < reps; ++r) launch();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  return ms / reps * 1e3f;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("device %s, %d SMs, %d MHz\n", prop.name, prop.multiProcessorCount, prop.clockRate / 1000);
  const long long n = 8000000, tabn = 1000000;
  std::vector<int> h(n);
  std::vector<double> ht(tabn);
  unsigned long long s = 88172645463325252ULL;
  for (long long i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % tabn); }
  for (long long i = 0; i < tabn; ++i) ht[i] = (double)(i % 1000) * 0.001;
  double ref = 0.0;
  for (long long i = 0; i < n; ++i) ref += ht[h[i]];
  int* idx; double *tab, *out;
  cudaMalloc(&idx, n * 4); cudaMalloc(&tab, tabn * 8); cudaMalloc(&out, n * 8);
  cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(tab, ht.data(), tabn * 8, cudaMemcpyHostToDevice);
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = tab; rd.res.linear.desc = cudaCreateChannelDesc<int2>(); rd.res.linear.sizeInBytes = tabn * 8;
  cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
  cudaTextureObject_t tex = 0;
  cudaError_t te = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
  if (te != cudaSuccess) printf("texture object: %s\n", cudaGetErrorString(te));
  const double clk = prop.clockRate * 1e3;
  auto report = [&](const char* name, int ilp, float us) {
    std::vector<double> ho((size_t)((n + ilp * BLOCK - 1) / (ilp * BLOCK)) * BLOCK);
    cudaMemcpy(ho.data(), out, ho.size() * 8, cudaMemcpyDeviceToHost);
    double sum = 0.0; for (double v : ho) sum += v;
    printf("%-28s ILP%-2d : %6.1f us  %5.0f Ggather/s  %.2f per clk per SM   %s\n", name, ilp, us, n / (us * 1e3), n / (us * 1e-6) / clk / prop.multiProcessorCount,
           fabs(sum - ref) <= 1e-6 * fabs(ref) ? "ok" : "WRONG SUM");
  };
#define RUN(NAME, ILP, MODE) { const int grid = (int)((n + (long long)BLOCK * ILP - 1) / ((long long)BLOCK * ILP)); \
    cudaMemset(out, 0, n * 8); report(NAME, ILP, time_it([&] { gather_kernel<ILP, MODE><<<grid, BLOCK>>>(idx, tab, tex, out, n); })); }
#define RUNB(NAME, ILP, MIX) { const int grid = (int)((n + (long long)BLOCK * ILP - 1) / ((long long)BLOCK * ILP)); \
    cudaMemset(out, 0, n * 8); report(NAME, ILP, time_it([&] { bulk_kernel<ILP, MIX><<<grid, BLOCK>>>(idx, tab, out, n); })); }
  RUN("ldg", 4, 0) RUN("ldg", 8, 0)
  if (te == cudaSuccess) { RUN("tex1Dfetch<int2>", 4, 1) RUN("tex1Dfetch<int2>", 8, 1) RUN("ldg + tex alternating", 8, 2) RUN("ldg + tex alternating", 16, 2) }
  RUN("ldgsts 8 B -> smem", 4, 3) RUN("ldgsts 8 B -> smem", 8, 3) RUN("ldgsts 8 B -> smem", 16, 3)
  RUN("ldgsts + ldg alternating", 8, 4) RUN("ldgsts + ldg alternating", 16, 4)
  RUNB("cp.async.bulk 16 B (TMA)", 4, false) RUNB("cp.async.bulk 16 B (TMA)", 8, false)
  RUNB("bulk + ldg alternating", 8, true) RUNB("bulk + ldg alternating", 16, true)
  return 0;
}
