// Opt-in variant of the fused gather + first-order + FM forward (rows E + L + F) that STAGES THE ROWS THROUGH TMA INTO
// SHARED MEMORY: `cp.async.bulk.tensor.2d ... tile::gather4` (UTMALDG.2D.GATHER4) fetches four table rows by index per
// instruction into a per-warp tile, completion on an mbarrier, the FM sums are computed from shared memory and the stacked
// rows leave as one `cp.async.bulk` store per tile.  Same reference lines as embed_fm.cu
// (keras/models/ranking/fm.py:23-37, deepfm.py:36-46).
//
// Why it is NOT the default: measured on B200 at C2 (tools/mb_gather.cu, profiles/mb_gather_r02_*.jsonl and the bench
// lines of this kernel) the TMA path is slower than register loads for this access pattern -- a gather4 costs the SM's
// TMA unit ~46 cycles per 4 rows, the register path is bound by the DRAM's random-row rate before that -- and for skewed
// (Zipf) ids the hot rows are L2 hits already.  It exists so that the comparison is a product measurement, not a claim.
//
// Scope: the fused row layout of embedding.py (one 128-B line [emb D | w | pad] per row, all tables in one arena),
// D = 16, S <= 32.  Everything else uses dr_embed_fm_fwd.
#include <cuda.h>
#include "common.cuh"

namespace dr {

__device__ __forceinline__ uint32_t tsm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void t_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void t_gather4(uint32_t sdst, const CUtensorMap* tm, int col, int r0, int r1, int r2, int r3,
                                          uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
               " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(sdst), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
__device__ __forceinline__ void t_bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}

constexpr int TMA_EX = 8;          // examples per warp tile
constexpr int TMA_WARPS = 4;
constexpr int TMA_NBUF = 2;
constexpr int TMA_D = 16;

struct EmbedTmaParams {
  const float* arena;
  int64_t total_rows;
  const int64_t* slot_offsets;
  const int64_t* rows;
  const void* ids;
  const float* bias;
  int64_t B;
  int S;
  int64_t row_stride;
  float* out_stack;
  float* out_sum;
  float* out_logit;
};

template <typename IdT>
__global__ void __launch_bounds__(TMA_WARPS * 32) embed_fm_fwd_tma_kernel(const EmbedTmaParams p,
                                                                           const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[TMA_WARPS * TMA_NBUF];
  __shared__ int64_t s_off[32], s_rows[32];
  const int S = p.S, LK = TMA_EX * S;              // lookups per tile (a multiple of 4)
  const int tile_bytes = LK * 64, buf_bytes = (tile_bytes + LK * 4 + 127) / 128 * 128;   // rows + first-order weights, 128-B multiple (TMA destination alignment)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* mybuf = smem + (size_t)warp * TMA_NBUF * buf_bytes;
  for (int i = threadIdx.x; i < S; i += blockDim.x) { s_off[i] = p.slot_offsets[i]; s_rows[i] = p.rows[i]; }
  if (lane == 0)
    for (int k = 0; k < TMA_NBUF; ++k) t_mbar_init(tsm_u32(&bars[warp * TMA_NBUF + k]), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const IdT* __restrict__ ids = reinterpret_cast<const IdT*>(p.ids);
  const int64_t ntiles = (p.B + TMA_EX - 1) / TMA_EX;
  const int64_t w0 = (int64_t)blockIdx.x * TMA_WARPS + warp, nw = (int64_t)gridDim.x * TMA_WARPS;
  const float bias = p.bias ? __ldg(p.bias) : 0.f;
  const int oob = (int)p.total_rows;               // a row index outside the tensor: TMA fills zeros (OOV id -> zero row)

  auto issue = [&](int64_t tile, int buf) {
    unsigned char* tb = mybuf + buf * buf_bytes;
    const uint32_t bar = tsm_u32(&bars[warp * TMA_NBUF + buf]);
    const int64_t b0 = tile * TMA_EX;
    const int n = (int)min((int64_t)LK, (p.B - b0) * S);
    if (lane == 0) t_mbar_expect_tx(bar, (uint32_t)(LK / 4) * 256u);
    __syncwarp();
    for (int m = lane; m < LK / 4; m += 32) {      // quad m = lookups 4m .. 4m+3 of the tile -> one gather4
      int r[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int j = m * 4 + x;
        r[x] = oob;
        float w = 0.f;
        if (j < n) {
          const int s = j % S;
          const int64_t id = (int64_t)__ldg(ids + b0 * S + j);
          if ((uint64_t)id < (uint64_t)s_rows[s]) {
            r[x] = (int)(s_off[s] + id);
            w = ldg_nc_na_f32(p.arena + (size_t)r[x] * p.row_stride + TMA_D);     // in-row first-order weight
          }
        }
        *reinterpret_cast<float*>(tb + tile_bytes + j * 4) = w;
      }
      t_gather4(tsm_u32(tb + m * 256), &tmap, 0, r[0], r[1], r[2], r[3], bar);
    }
  };

  int it = 0;
  if (w0 < ntiles) issue(w0, 0);
  for (int64_t tile = w0; tile < ntiles; tile += nw, ++it) {
    const int buf = it % TMA_NBUF;
    const int64_t nxt = tile + nw;
    // the buffer refilled now was stored from by the previous iteration: its bulk store must have read it
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    if (nxt < ntiles) issue(nxt, (it + 1) % TMA_NBUF);
    t_mbar_wait(tsm_u32(&bars[warp * TMA_NBUF + buf]), (uint32_t)((it / TMA_NBUF) & 1));
    __syncwarp();                                   // the first-order weights were ordinary shared-memory stores
    unsigned char* tb = mybuf + buf * buf_bytes;
    const int64_t b0 = tile * TMA_EX;
    const int nex = (int)min((int64_t)TMA_EX, p.B - b0);
    if (p.out_stack && lane == 0) {
      t_bulk_store(p.out_stack + b0 * S * TMA_D, tsm_u32(tb), (uint32_t)nex * S * 64);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    // FM sums from shared memory: for one example the 32 lanes read its S rows x 4 chunks contiguously (conflict free),
    // lanes with equal (lane & 3) combine by xor-shuffles
    for (int e = 0; e < nex; ++e) {
      float4 a = f4_zero(), q = f4_zero();
      const float4* src = reinterpret_cast<const float4*>(tb + e * (S * 64));
      for (int idx = lane; idx < S * 4; idx += 32) {
        const float4 v = src[idx];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
      }
      float lin = 0.f;
      for (int s = lane; s < S; s += 32) lin += *reinterpret_cast<const float*>(tb + tile_bytes + (e * S + s) * 4);
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, o); a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
        a.z += __shfl_xor_sync(0xffffffffu, a.z, o); a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
        q.x += __shfl_xor_sync(0xffffffffu, q.x, o); q.y += __shfl_xor_sync(0xffffffffu, q.y, o);
        q.z += __shfl_xor_sync(0xffffffffu, q.z, o); q.w += __shfl_xor_sync(0xffffffffu, q.w, o);
      }
      lin = group_sum<32>(lin);
      float t = (a.x * a.x - q.x) + (a.y * a.y - q.y) + (a.z * a.z - q.z) + (a.w * a.w - q.w);
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      if (p.out_sum && lane < 4) stg4(p.out_sum + (b0 + e) * TMA_D + lane * 4, a);
      if (p.out_logit && lane == 0) p.out_logit[b0 + e] = (bias + lin) + 0.5f * t;
    }
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeTiledFnE)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace dr

using namespace dr;

extern "C" int dr_embed_fm_fwd_tma(const float* arena, int64_t total_rows, const int64_t* slot_offsets,
                                   const int64_t* rows, const void* ids, int id_bytes, const float* bias, int64_t B, int S,
                                   int D, int64_t row_stride, float* out_stack, float* out_sum, float* out_logit,
                                   void* stream) {
  DR_REQUIRE(arena && slot_offsets && rows && ids, DR_EINVAL, "dr_embed_fm_fwd_tma: null pointer");
  DR_REQUIRE(B >= 0 && S >= 1, DR_EINVAL, "dr_embed_fm_fwd_tma: B=%lld S=%d", (long long)B, S);
  DR_REQUIRE(id_bytes == 8 || id_bytes == 4, DR_EINVAL, "dr_embed_fm_fwd_tma: id_bytes=%d (need 4 or 8)", id_bytes);
  DR_REQUIRE(D == TMA_D && S <= 32 && row_stride >= D + 4 && row_stride % 4 == 0 && total_rows >= 1 &&
                 total_rows < ((int64_t)1 << 31) - 1,
             DR_ENOTSUP, "dr_embed_fm_fwd_tma: this variant serves fused rows with D=16, S<=32 (got D=%d S=%d stride=%lld)", D, S,
             (long long)row_stride);
  DR_REQUIRE(aligned16(arena) && (!out_stack || aligned16(out_stack)) && (!out_sum || aligned16(out_sum)), DR_EALIGN,
             "dr_embed_fm_fwd_tma: arena / out_stack / out_sum not 16-B aligned");
  if (B == 0) return DR_OK;
  static EncodeTiledFnE enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    DR_REQUIRE(e == cudaSuccess && q == cudaDriverEntryPointSuccess && fn, DR_ENOTSUP,
               "dr_embed_fm_fwd_tma: cuTensorMapEncodeTiled entry point not available");
    enc = (EncodeTiledFnE)fn;
  }
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)row_stride, (cuuint64_t)total_rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_stride * 4};
  cuuint32_t box[2] = {(cuuint32_t)TMA_D, 1u};       // gather4: four 1-row boxes of D floats per instruction
  cuuint32_t es[2] = {1u, 1u};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(arena), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DR_REQUIRE(r == CUDA_SUCCESS, DR_EINVAL, "dr_embed_fm_fwd_tma: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  EmbedTmaParams p{};
  p.arena = arena; p.total_rows = total_rows; p.slot_offsets = slot_offsets; p.rows = rows; p.ids = ids; p.bias = bias;
  p.B = B; p.S = S; p.row_stride = row_stride; p.out_stack = out_stack; p.out_sum = out_sum; p.out_logit = out_logit;
  const size_t smem = (size_t)TMA_WARPS * TMA_NBUF * (((size_t)TMA_EX * S * 64 + (size_t)TMA_EX * S * 4 + 127) / 128 * 128);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ntiles = (B + TMA_EX - 1) / TMA_EX;
  if (id_bytes == 8) {
    auto k = embed_fm_fwd_tma_kernel<int64_t>;
    DR_CUDA_CALL(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    DR_CUDA_CALL(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, TMA_WARPS * 32, smem));
    int64_t ctas = (int64_t)kNumSMs * (occ > 0 ? occ : 1);
    if (ctas * TMA_WARPS > ntiles) ctas = (ntiles + TMA_WARPS - 1) / TMA_WARPS;
    k<<<(unsigned)ctas, TMA_WARPS * 32, smem, st>>>(p, tm);
  } else {
    auto k = embed_fm_fwd_tma_kernel<int32_t>;
    DR_CUDA_CALL(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    DR_CUDA_CALL(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, TMA_WARPS * 32, smem));
    int64_t ctas = (int64_t)kNumSMs * (occ > 0 ? occ : 1);
    if (ctas * TMA_WARPS > ntiles) ctas = (ntiles + TMA_WARPS - 1) / TMA_WARPS;
    k<<<(unsigned)ctas, TMA_WARPS * 32, smem, st>>>(p, tm);
  }
  DR_CUDA_LAUNCH_CHECK("dr_embed_fm_fwd_tma");
  return DR_OK;
}
