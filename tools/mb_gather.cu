// Standalone microbenchmark (round 2): how fast can one B200 gather the C2 working set -- 65536 x 26 random 64-byte
// rows out of 26 M -- and how many DRAM bytes does each way of asking for a row cost?
//
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 tools/mb_gather.cu -o tools/bin/mb_gather
//   run:   tools/bin/mb_gather [--iters N] [--only name] [--stride 16|32]
//   ncu:   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum tools/bin/mb_gather --iters 1
//
// Variants (every one verified against a naive reference kernel before it is timed):
//   reg<F>      lane group of 4 owns one example, 13 independent 16-B loads per lane in flight (the round-1 kernel's
//               scheme); F = load flavour: nc.L1::no_allocate / + .L2::64B / + .L2::128B / + .L2::256B / plain / .cg
//   cpasync     cp.async (LDGSTS) 16 B straight into a shared-memory tile (no registers held while in flight), compute
//               from shared memory, the stacked rows leave as ONE cp.async.bulk store per tile
//   bulk        one 64-B cp.async.bulk (TMA, non-tensor) per row, mbarrier completion
//   gather4     cp.async.bulk.tensor.2d ... tile::gather4: one TMA instruction fetches 4 rows by index
// Not product code: the winner moves into deep_recommenders_b200/csrc/embed_fm.cu.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#define CK(x)                                                                                       \
  do {                                                                                              \
    cudaError_t e_ = (x);                                                                           \
    if (e_ != cudaSuccess) {                                                                        \
      fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));          \
      exit(2);                                                                                      \
    }                                                                                               \
  } while (0)

constexpr int S = 26, D = 16, EX = 8, LK = EX * S;   // lookups per warp tile = 208
constexpr int64_t ROWS = 1000000;                   // rows per table

struct P {
  const float* arena;
  int stride;          // floats per row: 16 (64-B rows) or 32 (128-B fused rows [emb | w | pad])
  const int64_t* ids;  // [B, S]
  int64_t B;
  float* stack;        // [B, S*D] or nullptr
  float* sum;          // [B, D]
  float* lin;          // [B] (sum of in-row weights, stride 32 only) or nullptr
  int* err;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------ init + reference
__global__ void k_init(float* arena, int stride, int64_t total_rows) {
  int64_t n = total_rows * stride;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / stride;
    int c = (int)(i % stride);
    float v = 0.f;
    if (c < D) v = (float)((r * 31 + c * 7) % 1001) * 0.001f - 0.5f;
    else if (c == D) v = (float)((r * 13) % 503) * 0.002f - 0.5f;
    arena[i] = v;
  }
}

__global__ void k_ids(int64_t* ids, int64_t n, uint64_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t x = (uint64_t)i * 0x9E3779B97F4A7C15ull + seed;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
    int64_t id = (int64_t)(x % (uint64_t)ROWS);
    if ((x >> 40) % 50 == 0) id = -1;                 // 2 % OOV -> zero row
    ids[i] = id;
  }
}

__global__ void k_ref(P p) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= p.B * 4) return;
  int64_t b = t >> 2;
  int c = (int)(t & 3);
  float4 a = make_float4(0, 0, 0, 0);
  float lin = 0.f;
  for (int s = 0; s < S; ++s) {
    int64_t id = p.ids[b * S + s];
    float4 v = make_float4(0, 0, 0, 0);
    if ((uint64_t)id < (uint64_t)ROWS) {
      const float* row = p.arena + ((int64_t)s * ROWS + id) * p.stride;
      v = *reinterpret_cast<const float4*>(row + c * 4);
      if (c == 0 && p.lin) lin += row[D];
    }
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    if (p.stack) *reinterpret_cast<float4*>(p.stack + (b * S + s) * D + c * 4) = v;
  }
  *reinterpret_cast<float4*>(p.sum + b * D + c * 4) = a;
  if (c == 0 && p.lin) p.lin[b] = lin;
}

// ------------------------------------------------------------------------------------------ reg<FLAVOR>
template <int F>
__device__ __forceinline__ float4 ld16(const float* p) {
  float4 r;
  if (F == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (F == 1) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (F == 2) asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (F == 3) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (F == 4) asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (F == 5) asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (F == 6) asm volatile("ld.global.L1::evict_first.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
template <int F>
__device__ __forceinline__ float ld4(const float* p) {
  float r;
  if (F == 1 || F == 6) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.f32 %0, [%1];" : "=f"(r) : "l"(p));
  else asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

template <int F, int U>
__global__ void __launch_bounds__(256) k_reg(P p) {
  __shared__ int64_t sids[8][LK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int64_t ntiles = (p.B + EX - 1) / EX;
  const bool want_lin = p.lin != nullptr && c == 0;
  for (int64_t tile = blockIdx.x * 8 + warp; tile < ntiles; tile += (int64_t)gridDim.x * 8) {
    const int64_t b0 = tile * EX;
    const int nex = (int)min((int64_t)EX, p.B - b0);
    for (int i = lane; i < nex * S; i += 32) sids[warp][i] = __ldg(p.ids + b0 * S + i);
    __syncwarp();
    const bool ok = g < nex;
    const int64_t b = b0 + g;
    float4 a = make_float4(0, 0, 0, 0);
    float lin = 0.f;
    for (int s0 = 0; s0 < S; s0 += U) {
      float4 v[U];
      float w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int s = s0 + u;
        v[u] = make_float4(0, 0, 0, 0);
        w[u] = 0.f;
        if (s < S && ok) {
          const int64_t id = sids[warp][g * S + s];
          if ((uint64_t)id < (uint64_t)ROWS) {
            const float* row = p.arena + ((int64_t)s * ROWS + id) * p.stride;
            v[u] = ld16<F>(row + c * 4);
            if (want_lin) w[u] = ld4<F>(row + D);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int s = s0 + u;
        if (s < S) {
          a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w;
          lin += w[u];
          if (p.stack && ok) *reinterpret_cast<float4*>(p.stack + (b * S + s) * D + c * 4) = v[u];
        }
      }
    }
    if (ok) *reinterpret_cast<float4*>(p.sum + b * D + c * 4) = a;
    if (ok && want_lin) p.lin[b] = lin;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------ staged variants
// Per-warp tile: EX examples x S rows x 64 B = 13312 B, flat in lookup order j = e*S + s, i.e. exactly the bytes
// stack[b0 .. b0+EX) -- one contiguous cp.async.bulk store.  wbuf: the in-row weights of the tile's lookups.
constexpr int TILE_BYTES = LK * 64;                 // 13312
constexpr int WBUF_BYTES = LK * 16;                 // 3328 (16 B per lookup so the TMA variants can use it too)
constexpr int BUF_BYTES = TILE_BYTES + WBUF_BYTES;  // 16640

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a mis-programmed TMA must not hang the box
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err) {
  for (int i = 0; i < (1 << 22); ++i)
    if (mbar_try(bar, parity)) return true;
  atomicExch(err, 1);
  return false;
}
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t sdst, const CUtensorMap* tm, int col, int r0, int r1, int r2, int r3,
                                            uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
               " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(sdst), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

// compute phase shared by the staged kernels: the warp walks the tile's examples; for one example its 32 lanes read
// 26 rows x 4 chunks = 104 float4 contiguously (conflict free), lanes with equal (lane & 3) combine by xor-shuffles.
template <int WB>   // bytes per wbuf entry (4: cpasync, 16: TMA variants)
__device__ __forceinline__ void tile_compute(const P& p, const unsigned char* tile, const unsigned char* wbuf, int64_t b0,
                                             int nex, int lane) {
  for (int e = 0; e < nex; ++e) {
    float4 a = make_float4(0, 0, 0, 0);
    const float4* src = reinterpret_cast<const float4*>(tile + e * (S * 64));
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = it * 32 + lane;
      if (idx < S * 4) {
        const float4 v = src[idx];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
    }
    float lin = 0.f;
    if (p.lin && lane < S) lin = *reinterpret_cast<const float*>(wbuf + (e * S + lane) * WB);
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      a.x += __shfl_xor_sync(0xffffffffu, a.x, o); a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
      a.z += __shfl_xor_sync(0xffffffffu, a.z, o); a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
    }
    if (p.lin) {
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) lin += __shfl_xor_sync(0xffffffffu, lin, o);
    }
    if (lane < 4) *reinterpret_cast<float4*>(p.sum + (b0 + e) * D + lane * 4) = a;
    if (p.lin && lane == 0) p.lin[b0 + e] = lin;
  }
}

// ---- cp.async (LDGSTS) staging ---------------------------------------------------------------------------------
template <int WARPS, int NBUF>
__global__ void __launch_bounds__(WARPS * 32) k_cpasync(P p) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int WBB = LK * 4;
  constexpr int BB = TILE_BYTES + WBB + 64;   // + pad: keeps 128-B alignment of the next tile (13312 + 832 + 64 = 14208 = 111 * 128)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c = lane & 3, q = lane >> 2;
  unsigned char* mybuf = smem + (size_t)warp * NBUF * BB;
  const int64_t ntiles = (p.B + EX - 1) / EX;
  const int64_t w0 = (int64_t)blockIdx.x * WARPS + warp, nw = (int64_t)gridDim.x * WARPS;
  const bool want_lin = p.lin != nullptr;

  auto issue = [&](int64_t tile, int buf) {
    unsigned char* tb = mybuf + buf * BB;
    const int64_t b0 = tile * EX;
    const int n = (int)min((int64_t)LK, (p.B - b0) * S);
    // lane owns lookups j = lane + 32 k: coalesced id reads, global row index (or -1) kept in registers
    int row[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int j = lane + 32 * k;
      row[k] = -1;
      if (j < n) {
        const int64_t id = __ldg(p.ids + b0 * S + j);
        const int s = j % S;
        if ((uint64_t)id < (uint64_t)ROWS) row[k] = (int)(s * ROWS + id);
      }
    }
    if (want_lin) {
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const int j = lane + 32 * k;
        if (j < LK) cp_async4(smem_u32(tb + TILE_BYTES + j * 4), p.arena + (int64_t)max(row[k], 0) * p.stride + D, row[k] >= 0);
      }
    }
    // instruction i moves rows 8i .. 8i+7 (lane = row-in-octet q, chunk c): 8 whole rows per LDGSTS
#pragma unroll
    for (int i = 0; i < S; ++i) {
      const int r = __shfl_sync(0xffffffffu, row[i >> 2], ((i & 3) << 3) + q);
      cp_async16(smem_u32(tb + (i * 8 + q) * 64 + c * 16), p.arena + (int64_t)max(r, 0) * p.stride + c * 4, r >= 0);
    }
    cp_async_commit();
  };

  int it = 0;
  // prologue
#pragma unroll
  for (int k = 0; k < NBUF - 1; ++k) {
    if (w0 + k * nw < ntiles) issue(w0 + k * nw, k);
    else cp_async_commit();
  }
  for (int64_t tile = w0; tile < ntiles; tile += nw, ++it) {
    const int buf = it % NBUF;
    const int64_t nxt = tile + (NBUF - 1) * nw;
    // the buffer refilled now was stored by the previous iteration: its bulk store must have read it
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
    if (nxt < ntiles) issue(nxt, (it + NBUF - 1) % NBUF);
    else cp_async_commit();
    cp_async_wait<NBUF - 1>();
    fence_async_smem();
    __syncwarp();
    unsigned char* tb = mybuf + buf * BB;
    const int64_t b0 = tile * EX;
    const int nex = (int)min((int64_t)EX, p.B - b0);
    if (p.stack && lane == 0) {
      bulk_store(p.stack + b0 * S * D, smem_u32(tb), (uint32_t)nex * S * 64);
      bulk_commit();
    }
    tile_compute<4>(p, tb, tb + TILE_BYTES, b0, nex, lane);
    __syncwarp();
  }
  if (lane == 0) bulk_wait0();
}

// ---- one 64-B TMA bulk copy per row -----------------------------------------------------------------------------
// MODE 0: cp.async.bulk per row (+ 16-B bulk copy of [w | pad]); MODE 1: tile::gather4, 4 rows per instruction
template <int WARPS, int NBUF, int MODE>
__global__ void __launch_bounds__(WARPS * 32) k_tma(P p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int BB = BUF_BYTES;     // 16640 = 130 * 128
  __shared__ __align__(8) uint64_t bars[WARPS * NBUF];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* mybuf = smem + (size_t)warp * NBUF * BB;
  const int64_t ntiles = (p.B + EX - 1) / EX;
  const int64_t w0 = (int64_t)blockIdx.x * WARPS + warp, nw = (int64_t)gridDim.x * WARPS;
  const bool want_lin = p.lin != nullptr;
  const int64_t total_rows = (int64_t)S * ROWS;
  if (lane == 0)
    for (int k = 0; k < NBUF; ++k) mbar_init(smem_u32(&bars[warp * NBUF + k]), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();

  auto issue = [&](int64_t tile, int buf) {
    unsigned char* tb = mybuf + buf * BB;
    const uint32_t bar = smem_u32(&bars[warp * NBUF + buf]);
    const int64_t b0 = tile * EX;
    const int n = (int)min((int64_t)LK, (p.B - b0) * S);
    if (MODE == 0) {
      int row[7];
      int nvalid = 0;
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const int j = lane + 32 * k;
        row[k] = -1;
        if (j < n) {
          const int64_t id = __ldg(p.ids + b0 * S + j);
          if ((uint64_t)id < (uint64_t)ROWS) { row[k] = (int)((j % S) * ROWS + id); ++nvalid; }
        }
      }
      int tot = nvalid;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)tot * (want_lin ? 80u : 64u));
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const int j = lane + 32 * k;
        if (j < LK) {
          if (row[k] >= 0) {
            const float* src = p.arena + (int64_t)row[k] * p.stride;
            bulk_load(smem_u32(tb + j * 64), src, 64, bar);
            if (want_lin) bulk_load(smem_u32(tb + TILE_BYTES + j * 16), src + D, 16, bar);
          } else {
            float4 z = make_float4(0, 0, 0, 0);
            *reinterpret_cast<float4*>(tb + j * 64) = z; *reinterpret_cast<float4*>(tb + j * 64 + 16) = z;
            *reinterpret_cast<float4*>(tb + j * 64 + 32) = z; *reinterpret_cast<float4*>(tb + j * 64 + 48) = z;
            *reinterpret_cast<float4*>(tb + TILE_BYTES + j * 16) = z;
          }
        }
      }
    } else {
      // lane owns quads m = lane + 32 k' (52 quads per tile): 4 consecutive lookups -> one gather4
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)(LK / 4) * 256u);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int m = lane + 32 * k;
        if (m < LK / 4) {
          int r[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const int j = m * 4 + x;
            r[x] = (int)total_rows;            // out of bounds -> TMA zero fill
            if (j < n) {
              const int64_t id = __ldg(p.ids + b0 * S + j);
              if ((uint64_t)id < (uint64_t)ROWS) r[x] = (int)((j % S) * ROWS + id);
            }
          }
          tma_gather4(smem_u32(tb + m * 256), &tmap, 0, r[0], r[1], r[2], r[3], bar);
        }
      }
    }
  };
  (void)issue;

  int it = 0;
#pragma unroll
  for (int k = 0; k < NBUF - 1; ++k)
    if (w0 + k * nw < ntiles) issue(w0 + k * nw, k);
  for (int64_t tile = w0; tile < ntiles; tile += nw, ++it) {
    const int buf = it % NBUF;
    const int64_t nxt = tile + (NBUF - 1) * nw;
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
    if (nxt < ntiles) issue(nxt, (it + NBUF - 1) % NBUF);
    const uint32_t bar = smem_u32(&bars[warp * NBUF + buf]);
    if (!mbar_wait(bar, (uint32_t)((it / NBUF) & 1), p.err)) return;
    if (MODE == 0) fence_async_smem();    // the zero rows were generic-proxy stores
    __syncwarp();
    unsigned char* tb = mybuf + buf * BB;
    const int64_t b0 = tile * EX;
    const int nex = (int)min((int64_t)EX, p.B - b0);
    if (p.stack && lane == 0) {
      bulk_store(p.stack + b0 * S * D, smem_u32(tb), (uint32_t)nex * S * 64);
      bulk_commit();
    }
    tile_compute<16>(p, tb, tb + TILE_BYTES, b0, nex, lane);
    __syncwarp();
  }
  if (lane == 0) bulk_wait0();
}



// ------------------------------------------------------------------------------------------ pair mapping
// 8 lanes per example cover TWO adjacent slots (2j, 2j+1): their 8 x 16 B = one whole, line-aligned 128-B line of the
// stacked output, so every store instruction writes 4 full lines instead of 8 half lines.  4 examples per warp,
// 13 iterations (all 26 slots of the 4 examples in flight at U = 13).  ST: 0 plain, 1 st.global.cs, 2 L1::no_allocate
template <int ST>
__device__ __forceinline__ void st16(float* p, float4 v) {
  if (ST == 0) *reinterpret_cast<float4*>(p) = v;
  if (ST == 1) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  if (ST == 2) asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int F, int ST, int U, bool PAIR>
__global__ void __launch_bounds__(256) k_reg2(P p) {
  constexpr int EXW = PAIR ? 4 : 8;          // examples per warp tile
  constexpr int NIT = PAIR ? S / 2 : S;      // iterations over slots (slot pairs)
  __shared__ int srow[8][EXW * S];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = PAIR ? lane >> 3 : lane >> 2, h = PAIR ? (lane >> 2) & 1 : 0, c = lane & 3;
  const int64_t ntiles = (p.B + EXW - 1) / EXW;
  const bool want_lin = p.lin != nullptr && c == 0;
  for (int64_t tile = blockIdx.x * 8 + warp; tile < ntiles; tile += (int64_t)gridDim.x * 8) {
    const int64_t b0 = tile * EXW;
    const int nex = (int)min((int64_t)EXW, p.B - b0);
    for (int i = lane; i < EXW * S; i += 32) {
      int r = -1;
      if (i < nex * S) {
        const int64_t id = __ldg(p.ids + b0 * S + i);
        if ((uint64_t)id < (uint64_t)ROWS) r = (int)((i % S) * ROWS + id);
      }
      srow[warp][i] = r;
    }
    __syncwarp();
    const bool ok = g < nex;
    const int64_t b = b0 + g;
    float4 a = make_float4(0, 0, 0, 0);
    float lin = 0.f;
    for (int j0 = 0; j0 < NIT; j0 += U) {
      float4 v[U];
      float w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int s = PAIR ? 2 * (j0 + u) + h : j0 + u;
        v[u] = make_float4(0, 0, 0, 0);
        w[u] = 0.f;
        if (j0 + u < NIT) {
          const int r = srow[warp][g * S + s];
          if (r >= 0) {
            const float* row = p.arena + (int64_t)r * p.stride;
            v[u] = ld16<F>(row + c * 4);
            if (want_lin) w[u] = ld4<F>(row + D);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int s = PAIR ? 2 * (j0 + u) + h : j0 + u;
        if (j0 + u < NIT) {
          a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w;
          lin += w[u];
          if (p.stack && ok) st16<ST>(p.stack + (b * S + s) * D + c * 4, v[u]);
        }
      }
    }
    if (PAIR) {   // even-slot + odd-slot partial sums
      a.x += __shfl_xor_sync(0xffffffffu, a.x, 4); a.y += __shfl_xor_sync(0xffffffffu, a.y, 4);
      a.z += __shfl_xor_sync(0xffffffffu, a.z, 4); a.w += __shfl_xor_sync(0xffffffffu, a.w, 4);
      lin += __shfl_xor_sync(0xffffffffu, lin, 4);
    }
    if (ok && h == 0) *reinterpret_cast<float4*>(p.sum + b * D + c * 4) = a;
    if (ok && h == 0 && want_lin) p.lin[b] = lin;
    __syncwarp();
  }
}


// ------------------------------------------------------------------------------------------ L2 software prefetch
// Hypothesis under test: the register gather is capped by outstanding L1 misses per SM (~256 lines x ~1 us DRAM latency),
// not by DRAM.  prefetch.global.L2 is fire-and-forget (no register, no L1 miss entry): every warp prefetches the rows of
// its NEXT tile into L2 while it processes the current one, whose loads then are L2 hits (~0.3 us).
template <int PF>
__device__ __forceinline__ void pf_l2(const void* p) {
  if (PF == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  if (PF == 2) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p));
}
template <int F, int PF, int U, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_pf(P p) {
  __shared__ int srow[WARPS][2][LK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int64_t ntiles = (p.B + EX - 1) / EX;
  const int64_t nw = (int64_t)gridDim.x * WARPS;
  const bool want_lin = p.lin != nullptr && c == 0;
  int64_t idr[7];
  auto load_ids = [&](int64_t t) {
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int j = lane + 32 * k;
      idr[k] = -1;
      if (t < ntiles && j < LK && t * EX * S + j < p.B * S) idr[k] = __ldg(p.ids + t * EX * S + j);
    }
  };
  auto emit = [&](int buf) {
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int j = lane + 32 * k;
      if (j < LK) {
        int r = -1;
        if ((uint64_t)idr[k] < (uint64_t)ROWS) r = (int)((j % S) * ROWS + idr[k]);
        srow[warp][buf][j] = r;
        if (PF && r >= 0) pf_l2<PF>(p.arena + (int64_t)r * p.stride);
      }
    }
  };
  int64_t tile = (int64_t)blockIdx.x * WARPS + warp;
  load_ids(tile);
  emit(0);
  load_ids(tile + nw);
  for (int it = 0; tile < ntiles; tile += nw, ++it) {
    emit((it + 1) & 1);            // rows of the NEXT tile: staged + prefetched into L2
    load_ids(tile + 2 * nw);       // ids two tiles ahead travel during this iteration
    __syncwarp();
    const int* rows = srow[warp][it & 1];
    const int64_t b0 = tile * EX;
    const int nex = (int)min((int64_t)EX, p.B - b0);
    const bool ok = g < nex;
    const int64_t b = b0 + g;
    float4 a = make_float4(0, 0, 0, 0);
    float lin = 0.f;
    for (int s0 = 0; s0 < S; s0 += U) {
      float4 v[U];
      float w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int s = s0 + u;
        v[u] = make_float4(0, 0, 0, 0);
        w[u] = 0.f;
        if (s < S) {
          const int r = rows[g * S + s];
          if (r >= 0) {
            const float* row = p.arena + (int64_t)r * p.stride;
            v[u] = ld16<F>(row + c * 4);
            if (want_lin) w[u] = ld4<F>(row + D);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int s = s0 + u;
        if (s < S) {
          a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w;
          lin += w[u];
          if (p.stack && ok) *reinterpret_cast<float4*>(p.stack + (b * S + s) * D + c * 4) = v[u];
        }
      }
    }
    if (ok) *reinterpret_cast<float4*>(p.sum + b * D + c * 4) = a;
    if (ok && want_lin) p.lin[b] = lin;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------ hybrid (round 2)
// Rows arrive through REGISTER loads (the fastest way to ask for a row, see profiles/mb_gather_r02_*.jsonl), the stacked
// output leaves through a per-warp shared-memory tile and ONE cp.async.bulk store per tile (13 KB contiguous) instead of
// 26 x 8 scattered 64-B register stores.  V8: 256-bit loads -- lanes 0,1 of a 4-lane group fetch the two halves of the
// 64-B row, lane 2 the [w | pad] sector of the same 128-B line IN THE SAME INSTRUCTION (one L2 request per lookup).
// DYN: warps draw tiles from an atomic counter (no 3-vs-4-tiles tail).
__device__ __forceinline__ void ld32(const float* p, float (&r)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "l"(p));
}
__device__ unsigned int g_tile_counter;

template <int F, bool V8, bool DYN, int WARPS, int U>
__global__ void __launch_bounds__(WARPS * 32) k_hyb(P p) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int WB = TILE_BYTES + LK * 4;       // stacked tile + int32 row indices
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  unsigned char* tile_s = smem + (size_t)warp * WB;
  int* srow = reinterpret_cast<int*>(tile_s + TILE_BYTES);
  const int64_t ntiles = (p.B + EX - 1) / EX;
  const bool want_lin = p.lin != nullptr;
  int64_t tile = DYN ? 0 : (int64_t)blockIdx.x * WARPS + warp;
  const int64_t nw = (int64_t)gridDim.x * WARPS;
  bool first = true;
  for (;;) {
    if (DYN) {
      unsigned int t = 0;
      if (lane == 0) t = atomicAdd(&g_tile_counter, 1u);
      tile = __shfl_sync(0xffffffffu, t, 0);
    }
    if (tile >= ntiles) break;
    const int64_t b0 = tile * EX;
    const int nex = (int)min((int64_t)EX, p.B - b0);
    // the previous tile's bulk store must have read the tile buffer before it is overwritten
    if (!first && p.stack && lane == 0) bulk_wait_read0();
    first = false;
    for (int i = lane; i < LK; i += 32) {
      int r = -1;
      if (i < nex * S) {
        const int64_t id = __ldg(p.ids + b0 * S + i);
        if ((uint64_t)id < (uint64_t)ROWS) r = (int)((i % S) * ROWS + id);
      }
      srow[i] = r;
    }
    __syncwarp();
    const bool ok = g < nex;
    const int64_t b = b0 + g;
    float lin = 0.f;
    if (!V8) {
      float4 a = make_float4(0, 0, 0, 0);
      for (int s0 = 0; s0 < S; s0 += U) {
        float4 v[U];
        float w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int s = s0 + u;
          v[u] = make_float4(0, 0, 0, 0);
          w[u] = 0.f;
          if (s < S) {
            const int r = srow[g * S + s];
            if (r >= 0) {
              const float* row = p.arena + (int64_t)r * p.stride;
              v[u] = ld16<F>(row + c * 4);
              if (want_lin && c == 0) w[u] = ld4<F>(row + D);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int s = s0 + u;
          if (s < S) {
            a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w;
            lin += w[u];
            if (p.stack) *reinterpret_cast<float4*>(tile_s + (g * S + s) * 64 + c * 16) = v[u];
          }
        }
      }
      if (ok) *reinterpret_cast<float4*>(p.sum + b * D + c * 4) = a;
      if (ok && want_lin && c == 0) p.lin[b] = lin;
    } else {
      float a[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = 0.f;
      const bool ld_lane = c < 2 || (c == 2 && want_lin);
      for (int s0 = 0; s0 < S; s0 += U) {
        float v[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int s = s0 + u;
#pragma unroll
          for (int k = 0; k < 8; ++k) v[u][k] = 0.f;
          if (s < S && ld_lane) {
            const int r = srow[g * S + s];
            if (r >= 0) ld32(p.arena + (int64_t)r * p.stride + c * 8, v[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int s = s0 + u;
          if (s < S) {
            if (c < 2) {
#pragma unroll
              for (int k = 0; k < 8; ++k) a[k] += v[u][k];
              if (p.stack) {
                float4* d = reinterpret_cast<float4*>(tile_s + (g * S + s) * 64 + c * 32);
                d[0] = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
                d[1] = make_float4(v[u][4], v[u][5], v[u][6], v[u][7]);
              }
            } else {
              lin += v[u][0];
            }
          }
        }
      }
      if (ok && c < 2) {
        float4* d = reinterpret_cast<float4*>(p.sum + b * D + c * 8);
        d[0] = make_float4(a[0], a[1], a[2], a[3]);
        d[1] = make_float4(a[4], a[5], a[6], a[7]);
      }
      if (ok && want_lin && c == 2) p.lin[b] = lin;
    }
    if (p.stack) {
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        bulk_store(p.stack + b0 * S * D, smem_u32(tile_s), (uint32_t)nex * S * 64);
        bulk_commit();
      }
    }
    __syncwarp();
    if (!DYN) tile += nw;
  }
  if (p.stack && lane == 0) bulk_wait0();
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(CUtensorMap* tm, const float* base, int64_t rows, int stride, int box_rows, CUtensorMapL2promotion prom) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)stride, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)stride * 4};
  cuuint32_t box[2] = {16u, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1u, 1u};
  CUresult r = ((EncodeTiledFn)fn)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, prom,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fprintf(stderr, "cuTensorMapEncodeTiled -> %d\n", (int)r);
  return r == CUDA_SUCCESS;
}

static int g_sms = 148;   // --sms N: size every grid for N SMs (per-SM cap vs DRAM experiment)
struct Ctx {
  P p[4];
  float *ref_stack, *ref_sum, *ref_lin;
  int iters;
  int64_t B;
  int stride;
};

template <typename L>
static void run(const char* name, Ctx& c, bool stack, bool lin, L launch, double alg_bytes_per_ex) {
  P p = c.p[0];
  if (!stack) p.stack = nullptr;
  if (!lin) p.lin = nullptr;
  CK(cudaMemset(p.sum, 0xff, c.B * D * 4));
  if (p.stack) CK(cudaMemset(p.stack, 0xff, c.B * S * D * 4));
  if (p.lin) CK(cudaMemset(p.lin, 0xff, c.B * 4));
  CK(cudaMemset(p.err, 0, 4));
  launch(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("{\"variant\": \"%s\", \"error\": \"%s\"}\n", name, cudaGetErrorString(e)); exit(3); }
  int herr = 0;
  CK(cudaMemcpy(&herr, p.err, 4, cudaMemcpyDeviceToHost));
  // verify against the reference outputs of batch 0
  size_t ns = (size_t)c.B * S * D, nu = (size_t)c.B * D;
  std::vector<float> a(ns), b(ns);
  bool stack_ok = true, sum_ok = true, lin_ok = true;
  if (p.stack) {
    CK(cudaMemcpy(a.data(), p.stack, ns * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), c.ref_stack, ns * 4, cudaMemcpyDeviceToHost));
    stack_ok = memcmp(a.data(), b.data(), ns * 4) == 0;
  }
  CK(cudaMemcpy(a.data(), p.sum, nu * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.data(), c.ref_sum, nu * 4, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < nu; ++i) if (!(fabsf(a[i] - b[i]) <= 1e-4f)) { sum_ok = false; break; }
  if (p.lin) {
    CK(cudaMemcpy(a.data(), p.lin, c.B * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), c.ref_lin, c.B * 4, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < c.B; ++i) if (!(fabsf(a[i] - b[i]) <= 1e-4f)) { lin_ok = false; break; }
  }
  float ms = 0.f;
  if (c.iters > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) { P q = c.p[(i + 1) & 3]; if (!stack) q.stack = nullptr; if (!lin) q.lin = nullptr; launch(q); }
    CK(cudaEventRecord(e0));
    for (int i = 0; i < c.iters; ++i) { P q = c.p[i & 3]; if (!stack) q.stack = nullptr; if (!lin) q.lin = nullptr; launch(q); }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= c.iters;
  }
  double us = ms * 1e3;
  printf("{\"variant\": \"%s\", \"stride_B\": %d, \"stack\": %d, \"lin\": %d, \"us\": %.2f, \"alg_GBs\": %.1f, "
         "\"ok\": {\"stack\": %s, \"sum\": %s, \"lin\": %s, \"timeout\": %s}}\n",
         name, c.stride * 4, (int)stack, (int)lin, us, us > 0 ? alg_bytes_per_ex * c.B / us * 1e-3 : 0.0,
         stack_ok ? "true" : "false", sum_ok ? "true" : "false", lin_ok ? "true" : "false", herr ? "true" : "false");
  fflush(stdout);
}

int main(int argc, char** argv) {
  int iters = 20, stride = 32;
  std::string only;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--iters")) iters = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--stride")) stride = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--only")) only = argv[++i];
    else if (!strcmp(argv[i], "--sms")) g_sms = atoi(argv[++i]);
  }
  const int64_t B = 65536, total_rows = (int64_t)S * ROWS;
  Ctx c;
  c.iters = iters; c.B = B; c.stride = stride;
  float* arena;
  CK(cudaMalloc(&arena, (size_t)total_rows * stride * 4));
  k_init<<<148 * 8, 256>>>(arena, stride, total_rows);
  CK(cudaDeviceSynchronize());
  int* err;
  CK(cudaMalloc(&err, 4));
  for (int k = 0; k < 4; ++k) {
    int64_t* ids;
    CK(cudaMalloc(&ids, B * S * 8));
    k_ids<<<148 * 4, 256>>>(ids, B * S, 1234567ull * (k + 1));
    P& p = c.p[k];
    p.arena = arena; p.stride = stride; p.ids = ids; p.B = B; p.err = err;
    CK(cudaMalloc(&p.stack, B * S * D * 4));
    CK(cudaMalloc(&p.sum, B * D * 4));
    CK(cudaMalloc(&p.lin, B * 4));
  }
  CK(cudaMalloc(&c.ref_stack, B * S * D * 4));
  CK(cudaMalloc(&c.ref_sum, B * D * 4));
  CK(cudaMalloc(&c.ref_lin, B * 4));
  {
    P r = c.p[0];
    r.stack = c.ref_stack; r.sum = c.ref_sum; r.lin = stride > D ? c.ref_lin : nullptr;
    k_ref<<<(int)((B * 4 + 255) / 256), 256>>>(r);
    CK(cudaDeviceSynchronize());
  }
  const bool has_lin = stride > D;
  // algorithmic bytes / example: ids 8S + rows 64S (+ 4S weights) + stack 64S + sum 64 (+ 4 logit)
  auto alg = [&](bool stack, bool lin) { return (double)S * (8 + 64 + (lin ? 4 : 0)) + (stack ? S * 64 : 0) + 64 + (lin ? 4 : 0); };
  auto want = [&](const char* n) { return only.empty() || strncmp(n, only.c_str(), only.size()) == 0; };   // prefix match

#define RUN_REG(F, U, NAME)                                                                                       \
  if (want(NAME)) {                                                                                               \
    int occ = 0;                                                                                                  \
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_reg<F, U>, 256, 0));                                 \
    fprintf(stderr, "%s: %d CTAs / SM\n", NAME, occ);                                                             \
    for (int st = 0; st < 2; ++st)                                                                                \
      for (int li = 0; li < (has_lin ? 2 : 1); ++li)                                                              \
        run(NAME, c, st, li, [&](const P& p) { k_reg<F, U><<<g_sms * occ, 256>>>(p); }, alg(st, li));               \
  }
  RUN_REG(0, 13, "reg_nc_na")
  RUN_REG(1, 13, "reg_nc_na_L2_64B")
  RUN_REG(2, 13, "reg_nc_na_L2_128B")
  RUN_REG(3, 13, "reg_nc_na_L2_256B")
  RUN_REG(4, 13, "reg_plain")
  RUN_REG(5, 13, "reg_cg")
  RUN_REG(6, 13, "reg_evict_first_L2_64B")
  RUN_REG(0, 26, "reg_nc_na_u26")

#define RUN_CPA(W, NB, NAME)                                                                                      \
  if (want(NAME)) {                                                                                               \
    size_t sm = (size_t)W * NB * (TILE_BYTES + LK * 4 + 64);                                                      \
    CK(cudaFuncSetAttribute(k_cpasync<W, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));             \
    int occ = 0;                                                                                                  \
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cpasync<W, NB>, W * 32, sm));                        \
    fprintf(stderr, "%s: %zu B smem / CTA, %d CTAs / SM\n", NAME, sm, occ);                                       \
    for (int st = 0; st < 2; ++st)                                                                                \
      for (int li = 0; li < (has_lin ? 2 : 1); ++li)                                                              \
        run(NAME, c, st, li, [&](const P& p) { k_cpasync<W, NB><<<g_sms * occ, W * 32, sm>>>(p); }, alg(st, li));   \
  }
  RUN_CPA(4, 2, "cpasync_w4_b2")
  RUN_CPA(2, 2, "cpasync_w2_b2")
  RUN_CPA(4, 3, "cpasync_w4_b3")
  RUN_CPA(8, 2, "cpasync_w8_b2")
  RUN_CPA(2, 4, "cpasync_w2_b4")



#define RUN_REG2(F, ST, U, PAIR, CARVE, DUMMY, NAME)                                                              \
  if (want(NAME)) {                                                                                               \
    if (CARVE >= 0) CK(cudaFuncSetAttribute(k_reg2<F, ST, U, PAIR>, cudaFuncAttributePreferredSharedMemoryCarveout, CARVE)); \
    if (DUMMY) CK(cudaFuncSetAttribute(k_reg2<F, ST, U, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, DUMMY)); \
    int occ = 0;                                                                                                  \
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_reg2<F, ST, U, PAIR>, 256, DUMMY));                  \
    fprintf(stderr, "%s: %d CTAs / SM\n", NAME, occ);                                                             \
    for (int st = 0; st < 2; ++st)                                                                                \
      for (int li = (has_lin ? 1 : 0); li < (has_lin ? 2 : 1); ++li)                                              \
        run(NAME, c, st, li, [&](const P& p) { k_reg2<F, ST, U, PAIR><<<g_sms * occ, 256, DUMMY>>>(p); }, alg(st, li)); \
  }
  RUN_REG2(0, 0, 13, false, -1, 0, "r2_base")
  RUN_REG2(0, 0, 13, true, -1, 0, "r2_pair")
  RUN_REG2(0, 1, 13, true, -1, 0, "r2_pair_cs")
  RUN_REG2(0, 2, 13, true, -1, 0, "r2_pair_na")
  RUN_REG2(0, 1, 13, false, -1, 0, "r2_base_cs")
  RUN_REG2(1, 0, 13, true, -1, 0, "r2_pair_L2_64B")
  RUN_REG2(0, 0, 7, true, -1, 0, "r2_pair_u7")
  RUN_REG2(0, 0, 13, false, 0, 0, "r2_base_carve0")
  RUN_REG2(0, 0, 13, false, -1, 100 * 1024, "r2_base_dummy100k")
  RUN_REG2(0, 0, 13, true, 0, 0, "r2_pair_carve0")


#define RUN_PF(F, PF, U, W, CPS, NAME)                                                                            \
  if (want(NAME)) {                                                                                               \
    int occ = 0;                                                                                                  \
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pf<F, PF, U, W>, W * 32, 0));                        \
    if (CPS > 0 && occ > CPS) occ = CPS;                                                                          \
    fprintf(stderr, "%s: %d CTAs / SM\n", NAME, occ);                                                             \
    for (int st = 0; st < 2; ++st)                                                                                \
      for (int li = (has_lin ? 1 : 0); li < (has_lin ? 2 : 1); ++li)                                              \
        run(NAME, c, st, li, [&](const P& p) { k_pf<F, PF, U, W><<<g_sms * occ, W * 32>>>(p); }, alg(st, li));      \
  }
  RUN_PF(0, 0, 13, 8, 0, "pf0_w8")            // no prefetch: baseline of this code shape
  RUN_PF(0, 1, 13, 8, 0, "pf1_w8")
  RUN_PF(0, 1, 13, 8, 1, "pf1_w8_c1")         // 1 CTA / SM: 31 MB of prefetched lines in flight
  RUN_PF(0, 2, 13, 8, 0, "pf2_w8")
  RUN_PF(0, 1, 7, 8, 0, "pf1_w8_u7")
  RUN_PF(0, 1, 13, 4, 0, "pf1_w4")
  RUN_PF(0, 1, 26, 8, 1, "pf1_w8_u26_c1")
  RUN_PF(1, 1, 13, 8, 0, "pf1_w8_L2_64B")

#define RUN_HYB(F, V8, DYN, W, U, NAME)                                                                           \
  if (want(NAME) && (!(V8) || stride == 32)) {                                                                    \
    size_t sm = (size_t)W * (TILE_BYTES + LK * 4);                                                                \
    CK(cudaFuncSetAttribute(k_hyb<F, V8, DYN, W, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));      \
    int occ = 0;                                                                                                  \
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_hyb<F, V8, DYN, W, U>, W * 32, sm));                 \
    fprintf(stderr, "%s: %zu B smem / CTA, %d CTAs / SM\n", NAME, sm, occ);                                       \
    for (int st = 0; st < 2; ++st)                                                                                \
      for (int li = 0; li < (has_lin ? 2 : 1); ++li)                                                              \
        run(NAME, c, st, li, [&](const P& p) {                                                                    \
          if (DYN) { unsigned int z = 0; CK(cudaMemcpyToSymbolAsync(g_tile_counter, &z, 4, 0, cudaMemcpyHostToDevice, 0)); } \
          k_hyb<F, V8, DYN, W, U><<<g_sms * occ, W * 32, sm>>>(p); }, alg(st, li));                                 \
  }
  RUN_HYB(0, false, false, 8, 13, "hyb_v4_nc")
  RUN_HYB(1, false, false, 8, 13, "hyb_v4_L2_64B")
  RUN_HYB(0, false, true, 8, 13, "hyb_v4_nc_dyn")
  RUN_HYB(1, false, true, 8, 13, "hyb_v4_L2_64B_dyn")
  RUN_HYB(0, false, true, 8, 9, "hyb_v4_nc_dyn_u9")
  RUN_HYB(0, false, true, 8, 7, "hyb_v4_nc_dyn_u7")
  RUN_HYB(0, true, false, 8, 13, "hyb_v8_w8")
  RUN_HYB(0, true, false, 12, 13, "hyb_v8_w12")
  RUN_HYB(0, true, true, 12, 13, "hyb_v8_w12_dyn")
  RUN_HYB(0, true, true, 14, 9, "hyb_v8_w14_u9_dyn")
  RUN_HYB(0, true, true, 16, 7, "hyb_v8_w16_u7_dyn")
  RUN_HYB(0, true, true, 8, 13, "hyb_v8_w8_dyn")

  CUtensorMap tm1, tm4;
  bool m1 = make_map(&tm1, arena, total_rows, stride, 1, CU_TENSOR_MAP_L2_PROMOTION_NONE);
  bool m4 = make_map(&tm4, arena, total_rows, stride, 4, CU_TENSOR_MAP_L2_PROMOTION_NONE);
#define RUN_TMA(W, NB, MODE, MAP, NAME, LINMAX)                                                                   \
  if (want(NAME)) {                                                                                               \
    size_t sm = (size_t)W * NB * BUF_BYTES;                                                                       \
    CK(cudaFuncSetAttribute(k_tma<W, NB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));           \
    int occ = 0;                                                                                                  \
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_tma<W, NB, MODE>, W * 32, sm));                      \
    fprintf(stderr, "%s: %zu B smem / CTA, %d CTAs / SM\n", NAME, sm, occ);                                       \
    for (int st = 0; st < 2; ++st)                                                                                \
      for (int li = 0; li < LINMAX; ++li)                                                                         \
        run(NAME, c, st, li, [&](const P& p) { k_tma<W, NB, MODE><<<g_sms * occ, W * 32, sm>>>(p, MAP); }, alg(st, li)); \
  }
  RUN_TMA(4, 2, 0, tm1, "bulk64_w4_b2", (has_lin ? 2 : 1))
  RUN_TMA(2, 3, 0, tm1, "bulk64_w2_b3", (has_lin ? 2 : 1))
  if (m1) { RUN_TMA(4, 2, 1, tm1, "gather4_box1_w4_b2", 1) }
  if (m1) { RUN_TMA(2, 3, 1, tm1, "gather4_box1_w2_b3", 1) }
  return 0;
}
