// SURVEY.md 8(f) #1 proper: Adam for the embedding tables INSIDE the backward scatter path (one kernel), row-sparse.
//
// Reference call sites: examples/train_deepfm_on_movielens_keras.py:44 (tf.keras.optimizers.Adam()),
// examples/train_fm_on_movielens_estimator.py:51 (tf.train.AdamOptimizer(0.01)).
//
// Adam is not linear in the gradient, so the row gradient must be complete (summed over every lookup of that row in
// the batch) before (p, m, v) move.  One kernel does both phases with a per-row countdown:
//   0. dr_embed_adam_count (tiny, runs early in the step on a side stream): count[row] += 1 per lookup.
//   1. dr_embed_fm_bwd_adam: every lookup adds its gradient row into the row's state block with vector atomics
//      (red.global.add.v4.f32), fences, and decrements count[row]; the lane group that takes the count to zero is the
//      LAST contributor of that row: it reads the finished gradient, applies TensorFlow's ApplyAdam functor to the
//      row's 16-B chunks of (p, m, v) -- the arithmetic of dr_adam_step, bit for bit -- and clears the gradient, so the
//      state block is ready for the next step.  No dense gradient arena sweep, no second kernel, no per-row stamp.
// Per-row state lives in ONE block of `stride` floats ([g D | m D | v D | g_w m_w v_w count | pad], 128-B multiple,
// same DRAM page): a touched row costs the parameter line plus this block.
//
// Semantics: rows the batch did not touch keep p, m, v (tfa.optimizers.LazyAdam).  tf.keras Adam additionally decays m, v
// of EVERY row each step and moves every row by its decayed momentum (dr_adam_step over the arena is that exact form, at
// 7 arena sweeps per step); this is the documented deviation of the row-sparse form.  On the rows a step touches the
// update is TensorFlow's functor exactly.
#include "common.cuh"

namespace dr {

__device__ __forceinline__ void adam_functor(float& p, float g, float& m, float& v, float lr_t, float omb1, float omb2,
                                             float eps) {
  // training_ops ApplyAdam: m += (g - m)(1 - b1); v += (g*g - v)(1 - b2); var -= lr_t * m / (sqrt(v) + eps)
  m += (g - m) * omb1;
  v += (g * g - v) * omb2;
  p -= (m * lr_t) / (sqrtf(v) + eps);
}

__host__ __device__ inline int adam_state_stride(int D) { return (3 * D + 4 + 31) / 32 * 32; }

__device__ __forceinline__ float4 ld_cg4(const float* p) {
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ float ld_cg1(const float* p) {
  float r;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
  return r;
}

template <typename IdT>
__global__ void __launch_bounds__(256) embed_adam_count_kernel(const IdT* __restrict__ ids, int64_t n, int S,
                                                                const int64_t* __restrict__ rows,
                                                                const int64_t* __restrict__ slot_offsets,
                                                                float* __restrict__ state, int SS, int cnt_off) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int s = (int)(i % S);
    const int64_t id = (int64_t)__ldg(ids + i);
    if ((uint64_t)id < (uint64_t)__ldg(rows + s))
      atomicAdd(reinterpret_cast<int*>(state + (size_t)(__ldg(slot_offsets + s) + id) * SS + cnt_off), 1);
  }
}

struct AdamBwdParams {
  const void* ids;
  const int64_t* rows;
  const int64_t* slot_offsets;
  const float* stack;
  const float* sum_e;
  const float* g_logit;
  const float* g_stack;
  int64_t B;
  int S, D;
  float* const* table_ptrs;
  float* const* lin_ptrs;
  int64_t row_stride, lin_stride;
  int lin_in_row;
  float* state;
  int SS;
  float* g_bias;
  const float* lr_t_dev;
  float omb1, omb2, eps;
};

// Slot-parallel mapping of embed_fm_bwd_sp_kernel: a warp walks examples, its lane groups take consecutive slots.
template <int LPR, typename IdT>
__global__ void __launch_bounds__(256) embed_fm_bwd_adam_kernel(const AdamBwdParams p) {
  constexpr int SPW = 32 / LPR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float s_bias_part[8];
  const int S = p.S, D = p.D;
  float** s_tab = reinterpret_cast<float**>(smem_raw);
  float** s_lin = s_tab + S;
  int64_t* s_rows = reinterpret_cast<int64_t*>(s_lin + S);
  int64_t* s_off = s_rows + S;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    s_tab[i] = p.table_ptrs[i];
    s_lin[i] = p.lin_ptrs ? p.lin_ptrs[i] : nullptr;
    s_rows[i] = p.rows[i];
    s_off[i] = p.slot_offsets[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp_in_cta = threadIdx.x >> 5, warps_per_cta = blockDim.x >> 5;
  const int c = lane % LPR, sg = lane / LPR;
  const bool chunk_ok = (c * 4) < D;
  const bool has_fm = p.g_logit != nullptr, has_gs = p.g_stack != nullptr;
  const bool has_w = has_fm && (p.lin_in_row || p.lin_ptrs != nullptr);     // a first-order weight exists and gets a gradient
  const IdT* __restrict__ ids = reinterpret_cast<const IdT*>(p.ids);
  const int64_t warp0 = (int64_t)blockIdx.x * warps_per_cta + warp_in_cta;
  const int64_t nwarps = (int64_t)gridDim.x * warps_per_cta;
  const float lr_t = __ldg(p.lr_t_dev);
  const float omb1 = p.omb1, omb2 = p.omb2, eps = p.eps;
  const int SS = p.SS, o_m = D, o_v = 2 * D, o_s = 3 * D;
  float bias_acc = 0.f;

  for (int64_t b = warp0; b < p.B; b += nwarps) {
    const float gl = has_fm ? __ldg(p.g_logit + b) : 0.f;
    if (lane == 0) bias_acc += gl;
    float4 sum = f4_zero();
    if (has_fm && chunk_ok) {
      if (p.sum_e) {
        sum = ldg4(p.sum_e + (size_t)b * D + c * 4);
      } else {
        for (int s = 0; s < S; ++s) sum = f4_add(sum, ldg4(p.stack + ((size_t)b * S + s) * D + c * 4));
      }
    }
    const IdT* my_ids = ids + (size_t)b * S;
    const size_t ex0 = (size_t)b * S * D + c * 4;
    for (int s0 = 0; s0 < S; s0 += SPW) {       // trip count is warp-uniform: every lane reaches the barriers below
      const int s = s0 + sg;
      int64_t id = -1;
      float4 e = f4_zero(), gs = f4_zero();
      if (s < S) {
        id = (int64_t)__ldg(my_ids + s);
        if (chunk_ok) {
          if (has_fm) e = ldg_nc_na(p.stack + ex0 + (size_t)s * D);
          if (has_gs) gs = ldg_nc_na(p.g_stack + ex0 + (size_t)s * D);
        }
      }
      const bool valid = s < S && (uint64_t)id < (uint64_t)s_rows[s];
      float* st = nullptr;
      if (valid) {
        st = p.state + (size_t)(s_off[s] + id) * SS;
        if (chunk_ok) {
          float4 d;
          d.x = fmaf(gl, sum.x - e.x, gs.x);
          d.y = fmaf(gl, sum.y - e.y, gs.y);
          d.z = fmaf(gl, sum.z - e.z, gs.z);
          d.w = fmaf(gl, sum.w - e.w, gs.w);
          red_add_v4(st + c * 4, d);
        }
        if (c == 0 && has_w) red_add_f32(st + o_s, gl);
      }
      __threadfence();     // this lane's contributions are performed before the group leader announces them
      __syncwarp();
      int old = 0;
      if (valid && c == 0) old = atomicSub(reinterpret_cast<int*>(st + o_s + 3), 1);
      old = __shfl_sync(0xffffffffu, old, sg * LPR);
      if (valid && old == 1) {     // last contributor of this row in this batch: the gradient is complete
        __threadfence();
        float* prow = s_tab[s] + (size_t)id * p.row_stride;
        if (chunk_ok) {
          const float4 g = ld_cg4(st + c * 4);
          float4 m = ld_cg4(st + o_m + c * 4), v = ld_cg4(st + o_v + c * 4);
          float4 w = *reinterpret_cast<const float4*>(prow + c * 4);
          adam_functor(w.x, g.x, m.x, v.x, lr_t, omb1, omb2, eps);
          adam_functor(w.y, g.y, m.y, v.y, lr_t, omb1, omb2, eps);
          adam_functor(w.z, g.z, m.z, v.z, lr_t, omb1, omb2, eps);
          adam_functor(w.w, g.w, m.w, v.w, lr_t, omb1, omb2, eps);
          *reinterpret_cast<float4*>(prow + c * 4) = w;
          stg4(st + o_m + c * 4, m);
          stg4(st + o_v + c * 4, v);
          stg4(st + c * 4, f4_zero());
        }
        if (c == 0 && has_w) {
          float* wp = p.lin_in_row ? prow + D : s_lin[s] + (size_t)id * p.lin_stride;
          float w = *wp, m = ld_cg1(st + o_s + 1), v = ld_cg1(st + o_s + 2);
          const float g = ld_cg1(st + o_s);
          adam_functor(w, g, m, v, lr_t, omb1, omb2, eps);
          *wp = w;
          st[o_s + 1] = m;
          st[o_s + 2] = v;
          st[o_s] = 0.f;
        }
      }
    }
  }
  if (p.g_bias && has_fm) {
    bias_acc = group_sum<32>(bias_acc);
    if (lane == 0) s_bias_part[warp_in_cta] = bias_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < warps_per_cta; ++w) t += s_bias_part[w];
      red_add_f32(p.g_bias, t);
    }
  }
}

static int lpr_of(int D) {
  int chunks = D / 4, lpr = 1;
  while (lpr < chunks) lpr <<= 1;
  return lpr;
}

}  // namespace dr

using namespace dr;

extern "C" int dr_embed_adam_state_stride(int D) {
  if (D < 4 || D > 128 || (D & 3)) return -1;
  return adam_state_stride(D);
}

extern "C" int dr_embed_adam_count(const void* ids, int id_bytes, int64_t B, int S, int D, const int64_t* rows,
                                   const int64_t* slot_offsets, float* state, void* stream) {
  DR_REQUIRE(B >= 0 && S >= 1 && S <= 4096, DR_EINVAL, "dr_embed_adam_count: B=%lld S=%d", (long long)B, S);
  DR_REQUIRE(D >= 4 && D <= 128 && D % 4 == 0, DR_EINVAL, "dr_embed_adam_count: D=%d unsupported", D);
  DR_REQUIRE(id_bytes == 8 || id_bytes == 4, DR_EINVAL, "dr_embed_adam_count: id_bytes=%d (need 4 or 8)", id_bytes);
  if (B == 0) return DR_OK;
  DR_REQUIRE(ids && rows && slot_offsets && state, DR_EINVAL, "dr_embed_adam_count: null pointer");
  const int64_t n = B * S;
  int64_t ctas = (n + 255) / 256;
  if (ctas > (int64_t)kNumSMs * 16) ctas = (int64_t)kNumSMs * 16;
  const int SS = adam_state_stride(D);
  cudaStream_t st = (cudaStream_t)stream;
  if (id_bytes == 8)
    embed_adam_count_kernel<int64_t><<<(unsigned)ctas, 256, 0, st>>>((const int64_t*)ids, n, S, rows, slot_offsets, state, SS, 3 * D + 3);
  else
    embed_adam_count_kernel<int32_t><<<(unsigned)ctas, 256, 0, st>>>((const int32_t*)ids, n, S, rows, slot_offsets, state, SS, 3 * D + 3);
  DR_CUDA_LAUNCH_CHECK("dr_embed_adam_count");
  return DR_OK;
}

extern "C" int dr_embed_fm_bwd_adam(const void* ids, int id_bytes, const int64_t* rows, const int64_t* slot_offsets,
                                    const float* stack, const float* sum_e, const float* g_logit, const float* g_stack,
                                    int64_t B, int S, int D, int64_t row_stride, int64_t lin_stride, int flags,
                                    float* const* table_ptrs, float* const* lin_ptrs, float* state, float* g_bias,
                                    const float* lr_t_dev, float beta1, float beta2, float eps, void* stream) {
  DR_REQUIRE(B >= 0 && S >= 1 && S <= 4096, DR_EINVAL, "dr_embed_fm_bwd_adam: B=%lld S=%d", (long long)B, S);
  DR_REQUIRE(D >= 4 && D <= 128 && D % 4 == 0, DR_EINVAL, "dr_embed_fm_bwd_adam: D=%d unsupported", D);
  DR_REQUIRE(id_bytes == 8 || id_bytes == 4, DR_EINVAL, "dr_embed_fm_bwd_adam: id_bytes=%d (need 4 or 8)", id_bytes);
  if (B == 0) return DR_OK;
  DR_REQUIRE(ids && rows && slot_offsets && table_ptrs && state && lr_t_dev, DR_EINVAL, "dr_embed_fm_bwd_adam: null pointer");
  DR_REQUIRE(g_logit || g_stack, DR_EINVAL, "dr_embed_fm_bwd_adam: both g_logit and g_stack are NULL");
  DR_REQUIRE(!g_logit || stack, DR_EINVAL, "dr_embed_fm_bwd_adam: g_logit given but stack is NULL");
  DR_REQUIRE((!stack || aligned16(stack)) && (!g_stack || aligned16(g_stack)) && (!sum_e || aligned16(sum_e)) &&
                 aligned16(state),
             DR_EALIGN, "dr_embed_fm_bwd_adam: stack / g_stack / sum_e / state not 16-B aligned");
  if (row_stride == 0) row_stride = D;
  if (lin_stride == 0) lin_stride = 1;
  const int lin_in_row = (flags & DR_EMBED_LIN_IN_ROW) ? 1 : 0;
  DR_REQUIRE(row_stride >= D + (lin_in_row ? 4 : 0) && row_stride % 4 == 0 && lin_stride >= 1, DR_EINVAL,
             "dr_embed_fm_bwd_adam: bad strides row=%lld lin=%lld", (long long)row_stride, (long long)lin_stride);
  DR_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, DR_EINVAL,
             "dr_embed_fm_bwd_adam: beta1=%g beta2=%g eps=%g out of range", (double)beta1, (double)beta2, (double)eps);
  AdamBwdParams p{};
  p.ids = ids; p.rows = rows; p.slot_offsets = slot_offsets; p.stack = stack; p.sum_e = sum_e; p.g_logit = g_logit;
  p.g_stack = g_stack; p.B = B; p.S = S; p.D = D; p.table_ptrs = table_ptrs; p.lin_ptrs = lin_in_row ? nullptr : lin_ptrs;
  p.row_stride = row_stride; p.lin_stride = lin_stride; p.lin_in_row = lin_in_row; p.state = state;
  p.SS = adam_state_stride(D); p.g_bias = g_bias; p.lr_t_dev = lr_t_dev;
  p.omb1 = 1.f - beta1; p.omb2 = 1.f - beta2; p.eps = eps;
  const int threads = 256, warps = threads / 32;
  const size_t smem = (size_t)S * (sizeof(void*) * 2 + sizeof(int64_t) * 2);
  int64_t ctas = (B + warps - 1) / warps;
  if (ctas > (int64_t)kNumSMs * 8) ctas = (int64_t)kNumSMs * 8;
  cudaStream_t st = (cudaStream_t)stream;
#define DR_ADAM(L)                                                                                              \
  do {                                                                                                          \
    if (id_bytes == 8) embed_fm_bwd_adam_kernel<L, int64_t><<<(unsigned)ctas, threads, smem, st>>>(p);          \
    else embed_fm_bwd_adam_kernel<L, int32_t><<<(unsigned)ctas, threads, smem, st>>>(p);                        \
  } while (0)
  switch (lpr_of(D)) {
    case 1: DR_ADAM(1); break;
    case 2: DR_ADAM(2); break;
    case 4: DR_ADAM(4); break;
    case 8: DR_ADAM(8); break;
    case 16: DR_ADAM(16); break;
    default: DR_ADAM(32); break;
  }
#undef DR_ADAM
  DR_CUDA_LAUNCH_CHECK("dr_embed_fm_bwd_adam");
  return DR_OK;
}
