// Can ANY kernel become resident on an SM beside the persistent tcgen05 GEMM?  (diagnostic, one B200)
//
// Launches the C2 layer-0 weight-gradient GEMM ([416 x 65536] x [65536 x 256], through the C-ABI: dr_dense_bwd) on
// stream A and, right behind it, a dummy kernel on stream B whose CTAs record %globaltimer when they start and then
// spin for a few microseconds.  A 1-thread stamp kernel before / after the GEMM on stream A gives the GEMM's own
// start / end on the same clock.  Reported per dummy configuration (threads, registers via launch bounds, dynamic
// shared memory, carveout preference): how many dummy CTAs started BEFORE the GEMM ended.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_coresidency tools/probe_coresidency.cu \
//        -I include -L deep_recommenders_b200/lib -ldeeprec_b200 -Xlinker -rpath=deep_recommenders_b200/lib
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "deeprec_b200.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__global__ void stamp(unsigned long long* out) { *out = gtime(); }

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) dummy(unsigned long long* starts, unsigned* smids, int spin_ns) {
  extern __shared__ unsigned char sm[];
  const unsigned long long t0 = gtime();
  if (threadIdx.x == 0) {
    unsigned s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    starts[blockIdx.x] = t0;
    smids[blockIdx.x] = s;
    if (spin_ns < 0) sm[0] = 1;       // keep the dynamic shared memory referenced
  }
  while (gtime() - t0 < (unsigned long long)spin_ns) {}
}

int main(int argc, char** argv) {
  const int64_t M = 65536; const int K = 416, N = 256;
  int share = argc > 1 ? atoi(argv[1]) : 1;
  dr_tune_set("tc_dw_share", share);
  float *x, *g, *gw, *w;
  CK(cudaMalloc(&x, M * K * 4)); CK(cudaMalloc(&g, M * N * 4)); CK(cudaMalloc(&gw, K * N * 4)); CK(cudaMalloc(&w, K * N * 4));
  CK(cudaMemset(x, 0, M * K * 4)); CK(cudaMemset(g, 0, M * N * 4)); CK(cudaMemset(w, 0, K * N * 4));
  cudaStream_t A, B;
  CK(cudaStreamCreateWithFlags(&A, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&B, cudaStreamNonBlocking));
  unsigned long long *st, *t01; unsigned* smid;
  const int maxc = 148 * 16;
  CK(cudaMalloc(&st, maxc * 8)); CK(cudaMalloc(&smid, maxc * 4)); CK(cudaMalloc(&t01, 16));
  cudaEvent_t go; CK(cudaEventCreateWithFlags(&go, cudaEventDisableTiming));
  for (int i = 0; i < 3; ++i)
    if (int rc = dr_dense_bwd(x, w, nullptr, g, M, K, N, 0, nullptr, nullptr, gw, nullptr, A)) { printf("dr_dense_bwd rc=%d %s\n", rc, dr_last_error()); return 1; }
  CK(cudaDeviceSynchronize());

  struct Cfg { const char* name; int threads, minb, smem, carve, ctas_per_sm; };
  std::vector<Cfg> cfgs = {
      {"t128 r<=255 smem0 ", 128, 1, 0, -1, 1},       {"t128 r<=255 smem0 carve100", 128, 1, 0, 100, 1},
      {"t256 r<=64  smem0 ", 256, 4, 0, -1, 2},       {"t256 r<=64  smem0 carve100", 256, 4, 0, 100, 2},
      {"t256 r<=64  smem2K carve100", 256, 4, 2048, 100, 2}, {"t256 r<=64  smem8K carve100", 256, 4, 8192, 100, 1},
      {"t64  r<=255 smem0 carve100", 64, 1, 0, 100, 1},       {"t256 r<=32 smem0 carve100 x4", 256, 8, 0, 100, 4},
  };
  for (const Cfg& c : cfgs) {
    void (*k)(unsigned long long*, unsigned*, int) =
        c.threads == 128 ? dummy<128, 1> : (c.threads == 64 ? dummy<64, 1> : (c.minb == 8 ? dummy<256, 8> : dummy<256, 4>));
    if (c.carve >= 0) CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, c.carve));
    const int ctas = 148 * c.ctas_per_sm;
    CK(cudaMemset(st, 0, maxc * 8));
    CK(cudaDeviceSynchronize());
    stamp<<<1, 1, 0, A>>>(t01);
    CK(cudaEventRecord(go, A));
    dr_dense_bwd(x, w, nullptr, g, M, K, N, 0, nullptr, nullptr, gw, nullptr, A);
    stamp<<<1, 1, 0, A>>>(t01 + 1);
    CK(cudaStreamWaitEvent(B, go, 0));
    k<<<ctas, c.threads, c.smem, B>>>(st, smid, 5000);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> hs(ctas); unsigned long long ht[2];
    CK(cudaMemcpy(hs.data(), st, ctas * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ht, t01, 16, cudaMemcpyDeviceToHost));
    int before = 0; std::sort(hs.begin(), hs.end());
    for (auto s : hs) if (s < ht[1]) ++before;
    printf("{\"share\": %d, \"dummy\": \"%s\", \"ctas\": %d, \"gemm_us\": %.1f, \"dummy_ctas_started_before_gemm_end\": %d, "
           "\"first_dummy_start_us\": %.1f, \"median_dummy_start_us\": %.1f}\n",
           share, c.name, ctas, (ht[1] - ht[0]) / 1e3, before, ((double)hs[0] - (double)ht[0]) / 1e3,
           ((double)hs[ctas / 2] - (double)ht[0]) / 1e3);
  }
  return 0;
}
