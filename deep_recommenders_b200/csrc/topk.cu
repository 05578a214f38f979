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

// Variant 2 (default for k > 4): one pass over the row.  The warp keeps the current top-k as a list sorted by
// (value desc, index asc) in shared memory and a threshold tau = its k-th value; the row is scanned 32 columns at
// a time in ascending index order, and only values STRICTLY above tau are inserted (an equal value further right
// loses the tie, as in tf.math.top_k).  After the first k columns insertions become rare (about k ln(nc / k) in
// total for unordered data), so the cost is nc / 32 loads per lane -- each score is read exactly once.
__global__ void __launch_bounds__(256) topk_select_kernel(const float* __restrict__ scores, int64_t nq, int64_t nc,
                                                           int64_t ld, int k, int kpad, float* __restrict__ out_vals,
                                                           int32_t* __restrict__ out_idx) {
  extern __shared__ __align__(16) unsigned char topk_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* sv = reinterpret_cast<float*>(topk_smem) + (size_t)w * kpad;
  int32_t* si = reinterpret_cast<int32_t*>(topk_smem + (size_t)wpb * kpad * sizeof(float)) + (size_t)w * kpad;
  const int64_t row = (int64_t)blockIdx.x * wpb + w;
  if (row >= nq) return;   // warp-uniform
  const float* sr = scores + (size_t)row * ld;
  int count = 0;
  float tau = -INFINITY;
  for (int64_t j0 = 0; j0 < nc; j0 += 32) {
    const int64_t j = j0 + lane;
    float v = j < nc ? __ldg(sr + j) : -INFINITY;
    if (v != v) v = -INFINITY;   // NaN ranks last
    const bool cand = j < nc && (count < k || v > tau);
    unsigned bal = __ballot_sync(0xffffffffu, cand);
    while (bal) {   // all quantities below are warp-uniform
      const int src = __ffs(bal) - 1;
      bal &= bal - 1;
      const float cv = __shfl_sync(0xffffffffu, v, src);
      if (count == k && !(cv > tau)) continue;   // the threshold rose since the ballot
      int pos = 0;                                // elements that stay in front: value >= cv
      for (int e = lane; e < count; e += 32) pos += sv[e] >= cv ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) pos += __shfl_xor_sync(0xffffffffu, pos, o);
      const int last = count < k ? count : k - 1;   // highest slot written by the shift
      for (int top = last; top > pos; top -= 32) {  // move (pos, last] one slot right, highest chunk first
        const int e = top - lane;
        const bool act = e > pos;
        float tv = 0.f;
        int32_t ti = 0;
        if (act) { tv = sv[e - 1]; ti = si[e - 1]; }
        __syncwarp();
        if (act) { sv[e] = tv; si[e] = ti; }
        __syncwarp();
      }
      if (lane == 0) { sv[pos] = cv; si[pos] = (int32_t)(j0 + src); }
      __syncwarp();
      if (count < k) ++count;
      if (count == k) tau = sv[k - 1];
    }
  }
  for (int r = lane; r < k; r += 32) {
    const bool valid = r < count;
    out_vals[(size_t)row * k + r] = valid ? sv[r] : -INFINITY;
    out_idx[(size_t)row * k + r] = valid ? si[r] : -1;
  }
}

int g_tune_topk_variant = 0;   // 0 = auto, 1 = k selection passes, 2 = single-pass warp select

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
  const bool select = g_tune_topk_variant == 2 || (g_tune_topk_variant == 0 && k > 4);
  if (select && k <= 4096) {
    const int kpad = (k + 31) & ~31;
    int wpb = (int)(40960 / ((size_t)kpad * 8));   // warps per CTA that fit 40 KB of lists
    if (wpb > 8) wpb = 8;
    if (wpb < 1) wpb = 1;
    const size_t smem = (size_t)wpb * kpad * 8;
    const int64_t ctas = (nq + wpb - 1) / wpb;
    topk_select_kernel<<<(unsigned)ctas, wpb * 32, smem, (cudaStream_t)stream>>>(scores, nq, nc, ld, k, kpad, out_vals,
                                                                               out_idx);
    DR_CUDA_LAUNCH_CHECK("dr_topk_rows(select)");
    return DR_OK;
  }
  const int64_t ctas = (nq * 32 + 255) / 256;
  topk_rows_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(scores, nq, nc, ld, k, out_vals, out_idx);
  DR_CUDA_LAUNCH_CHECK("dr_topk_rows");
  return DR_OK;
}
