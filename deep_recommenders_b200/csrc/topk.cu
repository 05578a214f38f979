// SURVEY.md 8(f) #3 -- row-wise top-k, the selection step of the retrieval indexes
// (keras/models/retrieval/factorized_top_k.py: BruteForce.call :330-334 `tf.math.top_k(scores, k)`,
// Streaming.call :196-226 per-batch top_k then the merge top_k over [state | batch], _exclude :58-62).
// tf.math.top_k contract: values sorted descending; among equal values the lower index comes first.
// One warp per row, k selection passes over the row: each pass takes the largest (value, -index) strictly
// below the previous pick, so the result is exactly that order with no shared-memory heap.  Cost k * nc / 32
// loads per lane from L1/L2 (the row is read once from HBM); meant for the merge widths of the retrieval path
// (k <= a few hundred).  The score matrix may be a column window of a wider buffer (ld = row pitch in floats).
#include <float.h>
#include "common.cuh"

namespace dr {

__global__ void __launch_bounds__(256) topk_rows_kernel(const float* __restrict__ scores, int64_t nq, int64_t nc,
                                                         int64_t ld, int k, float* __restrict__ out_vals,
                                                         int32_t* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= nq) return;   // whole warps leave together: row is warp-uniform
  const float* sr = scores + (size_t)row * ld;
  float prev_v = INFINITY;
  int64_t prev_i = -1;
  for (int r = 0; r < k; ++r) {
    float best_v = -INFINITY;
    int64_t best_i = INT64_MAX;
    for (int64_t j = lane; j < nc; j += 32) {
      const float v = __ldg(sr + j);
      // NaN never compares: it is skipped, as a row of finite scores is the contract
      const bool below = (v < prev_v) || (v == prev_v && j > prev_i);
      if (below && (v > best_v || (v == best_v && j < best_i))) { best_v = v; best_i = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
    }
    prev_v = best_v; prev_i = best_i;
    if (lane == 0) {
      const bool valid = best_i != INT64_MAX;
      out_idx[(size_t)row * k + r] = valid ? (int32_t)best_i : -1;
      out_vals[(size_t)row * k + r] = valid ? best_v : -INFINITY;
    }
  }
}

}  // namespace dr

using namespace dr;

extern "C" int dr_topk_rows(const float* scores, int64_t nq, int64_t nc, int64_t ld, int k, float* out_vals,
                            int32_t* out_idx, void* stream) {
  DR_REQUIRE(nq >= 0 && nc >= 1 && ld >= nc, DR_EINVAL, "dr_topk_rows: nq=%lld nc=%lld ld=%lld", (long long)nq,
             (long long)nc, (long long)ld);
  // tf.math.top_k raises InvalidArgument "input must have at least k columns"
  DR_REQUIRE(k >= 1 && k <= nc, DR_EINVAL, "dr_topk_rows: input must have at least k columns (k=%d, columns=%lld)", k,
             (long long)nc);
  DR_REQUIRE(nc < ((int64_t)1 << 31), DR_EINVAL, "dr_topk_rows: nc=%lld does not fit the int32 index output",
             (long long)nc);
  if (nq == 0) return DR_OK;
  DR_REQUIRE(scores && out_vals && out_idx, DR_EINVAL, "dr_topk_rows: null pointer");
  const int64_t ctas = (nq * 32 + 255) / 256;
  topk_rows_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(scores, nq, nc, ld, k, out_vals, out_idx);
  DR_CUDA_LAUNCH_CHECK("dr_topk_rows");
  return DR_OK;
}
