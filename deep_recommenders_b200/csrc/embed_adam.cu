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

__host__ __device__ inline int adam_state_stride(int D) { return (3 * D + 5 + 31) / 32 * 32; }   // + count + step stamp

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

// ---- exact tf.keras Adam for rows that were NOT touched for a while (EXACT mode) ---------------------------------------
// TensorFlow's Adam is dense: in a step that does not touch a row, m <- m - m(1-b1), v <- v - v(1-b2) and the row still
// moves by -lr_j m / (sqrt(v) + eps).  The row-sparse form replays those skipped steps when the row is next touched (or
// flushed): per-row stamp = last step whose update the row has seen; lr_j comes from the history dr_adam_advance_hist
// keeps (ring of `hist_len` bias-corrected rates), or is recomputed when the step fell out of the ring.  Steps beyond
// kCatchUpMax after the stamp cannot move the row any more (m has decayed by 0.9^2048, far below fp32's range relative to
// v): they only decay v, in closed form.
constexpr int kCatchUpMax = 2048;
struct AdamHist {
  const float* lr_hist;
  int hist_len;
  float lr, b1, b2;
};
__device__ __forceinline__ float adam_lr_at(const AdamHist& h, int64_t j, int64_t t_now) {
  if (h.lr_hist && t_now - j < h.hist_len) return __ldg(h.lr_hist + (j % h.hist_len));
  return h.lr * sqrtf(1.f - powf(h.b2, (float)j)) / (1.f - powf(h.b1, (float)j));
}
// N elements of one lane: replay steps (stamp, upto] with zero gradient
template <int N>
__device__ __forceinline__ void adam_catch_up(float (&p)[N], float (&m)[N], float (&v)[N], int64_t stamp, int64_t upto,
                                              int64_t t_now, const AdamHist& h, float omb1, float omb2, float eps) {
  if (stamp <= 0 || upto <= stamp) return;                    // never updated (m = v = 0): nothing to replay
  const int64_t gap = upto - stamp;
  const int n = (int)(gap < kCatchUpMax ? gap : kCatchUpMax);
  for (int k = 1; k <= n; ++k) {
    const float lr_j = adam_lr_at(h, stamp + k, t_now);
#pragma unroll
    for (int e = 0; e < N; ++e) {
      m[e] = fmaf(-m[e], omb1, m[e]);
      v[e] = fmaf(-v[e], omb2, v[e]);
      p[e] -= __fdividef(m[e] * lr_j, sqrtf(v[e]) + eps);
    }
  }
  if (gap > n) {
    const float dm = powf(1.f - omb1, (float)(gap - n)), dv = powf(1.f - omb2, (float)(gap - n));
#pragma unroll
    for (int e = 0; e < N; ++e) { m[e] *= dm; v[e] *= dv; }
  }
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


// EXACT, before the forward: count the batch's lookups per row (as embed_adam_count_kernel) AND bring every row the batch
// touches up to date with step t-1, so that the forward reads what tf.keras Adam's dense update would have left there.
// One lane group per lookup; the lookup that raises a row's count from 0 is the row's first arriver and replays the
// row's pending steps (stamp, t-1]; the others only count.
template <int LPR, typename IdT>
__global__ void __launch_bounds__(256) embed_adam_prepare_kernel(const IdT* __restrict__ ids, int64_t n_lookups, int S, int D,
                                                                  const int64_t* __restrict__ rows,
                                                                  const int64_t* __restrict__ slot_offsets,
                                                                  float* const* __restrict__ table_ptrs,
                                                                  float* const* __restrict__ lin_ptrs, int64_t row_stride,
                                                                  int64_t lin_stride, int lin_in_row,
                                                                  float* __restrict__ state, int SS,
                                                                  const int64_t* __restrict__ step_dev, AdamHist hist,
                                                                  float omb1, float omb2, float eps) {
  constexpr int G = 32 / LPR;
  const int lane = threadIdx.x & 31, c = lane % LPR, grp = lane / LPR;
  const int64_t t_now = *step_dev;
  const int o_m = D, o_v = 2 * D, o_s = 3 * D;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t base = warp * G; base < n_lookups; base += nwarps * G) {
    const int64_t L = base + grp;
    int64_t id = -1;
    int s = 0;
    if (L < n_lookups) {
      s = (int)(L % S);
      id = (int64_t)__ldg(ids + L);
      if ((uint64_t)id >= (uint64_t)__ldg(rows + s)) id = -1;
    }
    float* st = id >= 0 ? state + (size_t)(__ldg(slot_offsets + s) + id) * SS : nullptr;
    int first = 0, stamp_i = 0;
    if (id >= 0 && c == 0) {
      first = atomicAdd(reinterpret_cast<int*>(st + o_s + 3), 1) == 0;
      if (first) stamp_i = __float_as_int(ld_cg1(st + o_s + 4));
    }
    first = __shfl_sync(0xffffffffu, first, grp * LPR);
    stamp_i = __shfl_sync(0xffffffffu, stamp_i, grp * LPR);
    const int64_t stamp = stamp_i;
    if (first && stamp > 0 && stamp < t_now - 1) {
      float* prow = table_ptrs[s] + (size_t)id * row_stride;
      if (c * 4 < D) {
        float4 m = ld_cg4(st + o_m + c * 4), v = ld_cg4(st + o_v + c * 4);
        float4 w = *reinterpret_cast<const float4*>(prow + c * 4);
        float pa[4] = {w.x, w.y, w.z, w.w}, ma[4] = {m.x, m.y, m.z, m.w}, va[4] = {v.x, v.y, v.z, v.w};
        adam_catch_up<4>(pa, ma, va, stamp, t_now - 1, t_now, hist, omb1, omb2, eps);
        *reinterpret_cast<float4*>(prow + c * 4) = make_float4(pa[0], pa[1], pa[2], pa[3]);
        stg4(st + o_m + c * 4, make_float4(ma[0], ma[1], ma[2], ma[3]));
        stg4(st + o_v + c * 4, make_float4(va[0], va[1], va[2], va[3]));
      }
      if (c == 0) {
        if (lin_in_row || lin_ptrs) {
          float* wp = lin_in_row ? prow + D : lin_ptrs[s] + (size_t)id * lin_stride;
          float pa[1] = {*wp}, ma[1] = {ld_cg1(st + o_s + 1)}, va[1] = {ld_cg1(st + o_s + 2)};
          adam_catch_up<1>(pa, ma, va, stamp, t_now - 1, t_now, hist, omb1, omb2, eps);
          *wp = pa[0]; st[o_s + 1] = ma[0]; st[o_s + 2] = va[0];
        }
        st[o_s + 4] = __int_as_float((int)(t_now - 1));
      }
    }
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
  const int64_t* step_dev;     // EXACT: current step t (dr_adam_advance_hist)
  AdamHist hist;
};

// Slot-parallel mapping of embed_fm_bwd_sp_kernel: a warp walks examples, its lane groups take consecutive slots.
template <int LPR, typename IdT, bool EXACT>
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
  const int64_t t_now = EXACT ? *p.step_dev : 0;
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
      int stamp_i = 0;             // EXACT: the row's step stamp, read by the group leader, broadcast while the warp is converged
      if (EXACT) {
        if (valid && c == 0 && old == 1) stamp_i = __float_as_int(ld_cg1(st + o_s + 4));
        stamp_i = __shfl_sync(0xffffffffu, stamp_i, sg * LPR);
      }
      if (valid && old == 1) {     // last contributor of this row in this batch: the gradient is complete
        __threadfence();
        float* prow = s_tab[s] + (size_t)id * p.row_stride;
        const int64_t stamp = stamp_i;
        if (chunk_ok) {
          const float4 g = ld_cg4(st + c * 4);
          float4 m = ld_cg4(st + o_m + c * 4), v = ld_cg4(st + o_v + c * 4);
          float4 w = *reinterpret_cast<const float4*>(prow + c * 4);
          if (EXACT) {     // replay the steps this row sat out, then take this step's update
            float pa[4] = {w.x, w.y, w.z, w.w}, ma[4] = {m.x, m.y, m.z, m.w}, va[4] = {v.x, v.y, v.z, v.w};
            adam_catch_up<4>(pa, ma, va, stamp, t_now - 1, t_now, p.hist, omb1, omb2, eps);
            w = make_float4(pa[0], pa[1], pa[2], pa[3]); m = make_float4(ma[0], ma[1], ma[2], ma[3]);
            v = make_float4(va[0], va[1], va[2], va[3]);
          }
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
          if (EXACT) {
            float pa[1] = {w}, ma[1] = {m}, va[1] = {v};
            adam_catch_up<1>(pa, ma, va, stamp, t_now - 1, t_now, p.hist, omb1, omb2, eps);
            w = pa[0]; m = ma[0]; v = va[0];
          }
          adam_functor(w, g, m, v, lr_t, omb1, omb2, eps);
          *wp = w;
          st[o_s + 1] = m;
          st[o_s + 2] = v;
          st[o_s] = 0.f;
        }
        if (EXACT && c == 0) st[o_s + 4] = __int_as_float((int)t_now);
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

// EXACT: bring EVERY row up to date with step `t_now` (before reading the tables for evaluation / a checkpoint):
// one lane group per table row, rows whose stamp is already t_now are skipped.
template <int LPR>
__global__ void __launch_bounds__(256) embed_adam_flush_kernel(float* const* __restrict__ table_ptrs,
                                                                float* const* __restrict__ lin_ptrs,
                                                                const int64_t* __restrict__ rows,
                                                                const int64_t* __restrict__ slot_offsets, int S, int D,
                                                                int64_t row_stride, int64_t lin_stride, int lin_in_row,
                                                                float* __restrict__ state, int SS, int64_t total_rows,
                                                                const int64_t* __restrict__ step_dev, AdamHist hist,
                                                                float omb1, float omb2, float eps) {
  const int lane = threadIdx.x & 31, c = lane % LPR, grp = lane / LPR;
  constexpr int G = 32 / LPR;
  const int64_t t_now = *step_dev;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int o_m = D, o_v = 2 * D, o_s = 3 * D;
  for (int64_t base = warp * G; base < total_rows; base += nwarps * G) {
    const int64_t r = base + grp;
    const bool in = r < total_rows;
    float* st = state + (size_t)(in ? r : 0) * SS;
    const int64_t stamp = in ? (int64_t)__float_as_int(st[o_s + 4]) : 0;   // every lane reads its group's stamp ...
    __syncwarp();                                                         // ... before any lane of the warp writes one
    if (!in || stamp <= 0 || stamp >= t_now) continue;
    int s = 0;                                             // table of global row r (S is small: linear scan)
    while (s + 1 < S && __ldg(slot_offsets + s + 1) <= r) ++s;
    const int64_t id = r - __ldg(slot_offsets + s);
    float* prow = table_ptrs[s] + (size_t)id * row_stride;
    if (c * 4 < D) {
      float4 m = *reinterpret_cast<float4*>(st + o_m + c * 4), v = *reinterpret_cast<float4*>(st + o_v + c * 4);
      float4 w = *reinterpret_cast<float4*>(prow + c * 4);
      float pa[4] = {w.x, w.y, w.z, w.w}, ma[4] = {m.x, m.y, m.z, m.w}, va[4] = {v.x, v.y, v.z, v.w};
      adam_catch_up<4>(pa, ma, va, stamp, t_now, t_now, hist, omb1, omb2, eps);
      *reinterpret_cast<float4*>(prow + c * 4) = make_float4(pa[0], pa[1], pa[2], pa[3]);
      *reinterpret_cast<float4*>(st + o_m + c * 4) = make_float4(ma[0], ma[1], ma[2], ma[3]);
      *reinterpret_cast<float4*>(st + o_v + c * 4) = make_float4(va[0], va[1], va[2], va[3]);
    }
    if (c == 0 && (lin_in_row || lin_ptrs)) {
      float* wp = lin_in_row ? prow + D : lin_ptrs[s] + (size_t)id * lin_stride;
      float pa[1] = {*wp}, ma[1] = {st[o_s + 1]}, va[1] = {st[o_s + 2]};
      adam_catch_up<1>(pa, ma, va, stamp, t_now, t_now, hist, omb1, omb2, eps);
      *wp = pa[0]; st[o_s + 1] = ma[0]; st[o_s + 2] = va[0];
    }
    if (c == 0) st[o_s + 4] = __int_as_float((int)t_now);
  }
}

__global__ void adam_advance_hist_kernel(int64_t* step, float lr, float b1, float b2, float* lr_t, float* lr_hist,
                                         int hist_len) {
  const int64_t t = *step + 1;
  *step = t;
  const double p1 = pow((double)b1, (double)t), p2 = pow((double)b2, (double)t);
  const float v = (float)((double)lr * sqrt(1.0 - p2) / (1.0 - p1));
  *lr_t = v;
  if (lr_hist && hist_len > 0) lr_hist[t % hist_len] = v;
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

static int embed_fm_bwd_adam_impl(const char* fn, const void* ids, int id_bytes, const int64_t* rows,
                                  const int64_t* slot_offsets, const float* stack, const float* sum_e,
                                  const float* g_logit, const float* g_stack, int64_t B, int S, int D,
                                  int64_t row_stride, int64_t lin_stride, int flags, float* const* table_ptrs,
                                  float* const* lin_ptrs, float* state, float* g_bias, const float* lr_t_dev, float beta1,
                                  float beta2, float eps, const int64_t* step_dev, const float* lr_hist, int hist_len,
                                  float lr, void* stream) {
  DR_REQUIRE(B >= 0 && S >= 1 && S <= 4096, DR_EINVAL, "%s: B=%lld S=%d", fn, (long long)B, S);
  DR_REQUIRE(D >= 4 && D <= 128 && D % 4 == 0, DR_EINVAL, "%s: D=%d unsupported", fn, D);
  DR_REQUIRE(id_bytes == 8 || id_bytes == 4, DR_EINVAL, "%s: id_bytes=%d (need 4 or 8)", fn, id_bytes);
  if (B == 0) return DR_OK;
  DR_REQUIRE(ids && rows && slot_offsets && table_ptrs && state && lr_t_dev, DR_EINVAL, "%s: null pointer", fn);
  DR_REQUIRE(g_logit || g_stack, DR_EINVAL, "%s: both g_logit and g_stack are NULL", fn);
  DR_REQUIRE(!g_logit || stack, DR_EINVAL, "%s: g_logit given but stack is NULL", fn);
  DR_REQUIRE((!stack || aligned16(stack)) && (!g_stack || aligned16(g_stack)) && (!sum_e || aligned16(sum_e)) &&
                 aligned16(state),
             DR_EALIGN, "%s: stack / g_stack / sum_e / state not 16-B aligned", fn);
  if (row_stride == 0) row_stride = D;
  if (lin_stride == 0) lin_stride = 1;
  const int lin_in_row = (flags & DR_EMBED_LIN_IN_ROW) ? 1 : 0;
  DR_REQUIRE(row_stride >= D + (lin_in_row ? 4 : 0) && row_stride % 4 == 0 && lin_stride >= 1, DR_EINVAL,
             "%s: bad strides row=%lld lin=%lld", fn, (long long)row_stride, (long long)lin_stride);
  DR_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, DR_EINVAL,
             "%s: beta1=%g beta2=%g eps=%g out of range", fn, (double)beta1, (double)beta2, (double)eps);
  const bool exact = step_dev != nullptr;
  AdamBwdParams p{};
  p.ids = ids; p.rows = rows; p.slot_offsets = slot_offsets; p.stack = stack; p.sum_e = sum_e; p.g_logit = g_logit;
  p.g_stack = g_stack; p.B = B; p.S = S; p.D = D; p.table_ptrs = table_ptrs; p.lin_ptrs = lin_in_row ? nullptr : lin_ptrs;
  p.row_stride = row_stride; p.lin_stride = lin_stride; p.lin_in_row = lin_in_row; p.state = state;
  p.SS = adam_state_stride(D); p.g_bias = g_bias; p.lr_t_dev = lr_t_dev;
  p.omb1 = 1.f - beta1; p.omb2 = 1.f - beta2; p.eps = eps;
  p.step_dev = step_dev; p.hist = AdamHist{lr_hist, hist_len, lr, beta1, beta2};
  const int threads = 256, warps = threads / 32;
  const size_t smem = (size_t)S * (sizeof(void*) * 2 + sizeof(int64_t) * 2);
  int64_t ctas = (B + warps - 1) / warps;
  if (ctas > (int64_t)kNumSMs * 8) ctas = (int64_t)kNumSMs * 8;
  cudaStream_t st = (cudaStream_t)stream;
#define DR_ADAM2(L, IdT)                                                                                        \
  do {                                                                                                          \
    if (exact) embed_fm_bwd_adam_kernel<L, IdT, true><<<(unsigned)ctas, threads, smem, st>>>(p);                \
    else embed_fm_bwd_adam_kernel<L, IdT, false><<<(unsigned)ctas, threads, smem, st>>>(p);                     \
  } while (0)
#define DR_ADAM(L) do { if (id_bytes == 8) DR_ADAM2(L, int64_t); else DR_ADAM2(L, int32_t); } while (0)
  switch (lpr_of(D)) {
    case 1: DR_ADAM(1); break;
    case 2: DR_ADAM(2); break;
    case 4: DR_ADAM(4); break;
    case 8: DR_ADAM(8); break;
    case 16: DR_ADAM(16); break;
    default: DR_ADAM(32); break;
  }
#undef DR_ADAM
#undef DR_ADAM2
  DR_CUDA_LAUNCH_CHECK(fn);
  return DR_OK;
}

extern "C" int dr_embed_fm_bwd_adam(const void* ids, int id_bytes, const int64_t* rows, const int64_t* slot_offsets,
                                    const float* stack, const float* sum_e, const float* g_logit, const float* g_stack,
                                    int64_t B, int S, int D, int64_t row_stride, int64_t lin_stride, int flags,
                                    float* const* table_ptrs, float* const* lin_ptrs, float* state, float* g_bias,
                                    const float* lr_t_dev, float beta1, float beta2, float eps, void* stream) {
  return embed_fm_bwd_adam_impl("dr_embed_fm_bwd_adam", ids, id_bytes, rows, slot_offsets, stack, sum_e, g_logit, g_stack, B,
                                S, D, row_stride, lin_stride, flags, table_ptrs, lin_ptrs, state, g_bias, lr_t_dev, beta1,
                                beta2, eps, nullptr, nullptr, 0, 0.f, stream);
}

extern "C" int dr_embed_fm_bwd_adam_tf(const void* ids, int id_bytes, const int64_t* rows, const int64_t* slot_offsets,
                                       const float* stack, const float* sum_e, const float* g_logit, const float* g_stack,
                                       int64_t B, int S, int D, int64_t row_stride, int64_t lin_stride, int flags,
                                       float* const* table_ptrs, float* const* lin_ptrs, float* state, float* g_bias,
                                       const int64_t* step_dev, const float* lr_t_dev, const float* lr_hist, int hist_len,
                                       float lr, float beta1, float beta2, float eps, void* stream) {
  DR_REQUIRE(step_dev, DR_EINVAL, "dr_embed_fm_bwd_adam_tf: step_dev is NULL");
  DR_REQUIRE(hist_len >= 0 && (hist_len == 0 || lr_hist), DR_EINVAL, "dr_embed_fm_bwd_adam_tf: bad lr history");
  return embed_fm_bwd_adam_impl("dr_embed_fm_bwd_adam_tf", ids, id_bytes, rows, slot_offsets, stack, sum_e, g_logit, g_stack,
                                B, S, D, row_stride, lin_stride, flags, table_ptrs, lin_ptrs, state, g_bias, lr_t_dev, beta1,
                                beta2, eps, step_dev, lr_hist, hist_len, lr, stream);
}

extern "C" int dr_adam_advance_hist(int64_t* step_dev, float lr, float beta1, float beta2, float* lr_t_dev,
                                    float* lr_hist, int hist_len, void* stream) {
  DR_REQUIRE(step_dev && lr_t_dev, DR_EINVAL, "dr_adam_advance_hist: null pointer");
  DR_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && hist_len >= 0, DR_EINVAL,
             "dr_adam_advance_hist: beta1=%g beta2=%g hist_len=%d out of range", (double)beta1, (double)beta2, hist_len);
  adam_advance_hist_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev, lr, beta1, beta2, lr_t_dev, lr_hist, hist_len);
  DR_CUDA_LAUNCH_CHECK("dr_adam_advance_hist");
  return DR_OK;
}


extern "C" int dr_embed_adam_prepare(const void* ids, int id_bytes, int64_t B, int S, int D, const int64_t* rows,
                                     const int64_t* slot_offsets, int64_t row_stride, int64_t lin_stride, int flags,
                                     float* const* table_ptrs, float* const* lin_ptrs, float* state,
                                     const int64_t* step_dev, const float* lr_hist, int hist_len, float lr, float beta1,
                                     float beta2, float eps, void* stream) {
  DR_REQUIRE(B >= 0 && S >= 1 && S <= 4096, DR_EINVAL, "dr_embed_adam_prepare: B=%lld S=%d", (long long)B, S);
  DR_REQUIRE(D >= 4 && D <= 128 && D % 4 == 0, DR_EINVAL, "dr_embed_adam_prepare: D=%d unsupported", D);
  DR_REQUIRE(id_bytes == 8 || id_bytes == 4, DR_EINVAL, "dr_embed_adam_prepare: id_bytes=%d (need 4 or 8)", id_bytes);
  if (B == 0) return DR_OK;
  DR_REQUIRE(ids && rows && slot_offsets && table_ptrs && state && step_dev, DR_EINVAL, "dr_embed_adam_prepare: null pointer");
  if (row_stride == 0) row_stride = D;
  if (lin_stride == 0) lin_stride = 1;
  const int lin_in_row = (flags & DR_EMBED_LIN_IN_ROW) ? 1 : 0;
  const AdamHist h{lr_hist, hist_len, lr, beta1, beta2};
  const int lpr = lpr_of(D);
  const int64_t n = B * S;
  int64_t ctas = (n * lpr + 255) / 256;
  if (ctas > (int64_t)kNumSMs * 16) ctas = (int64_t)kNumSMs * 16;
  cudaStream_t st = (cudaStream_t)stream;
#define DR_PREP2(L, IdT)                                                                                               \
  embed_adam_prepare_kernel<L, IdT><<<(unsigned)ctas, 256, 0, st>>>((const IdT*)ids, n, S, D, rows, slot_offsets, table_ptrs, \
                                                                  lin_in_row ? nullptr : lin_ptrs, row_stride, lin_stride, \
                                                                  lin_in_row, state, adam_state_stride(D), step_dev, h,  \
                                                                  1.f - beta1, 1.f - beta2, eps)
#define DR_PREP(L) do { if (id_bytes == 8) DR_PREP2(L, int64_t); else DR_PREP2(L, int32_t); } while (0)
  switch (lpr) {
    case 1: DR_PREP(1); break;
    case 2: DR_PREP(2); break;
    case 4: DR_PREP(4); break;
    case 8: DR_PREP(8); break;
    case 16: DR_PREP(16); break;
    default: DR_PREP(32); break;
  }
#undef DR_PREP
#undef DR_PREP2
  DR_CUDA_LAUNCH_CHECK("dr_embed_adam_prepare");
  return DR_OK;
}

extern "C" int dr_embed_adam_flush(const int64_t* rows, const int64_t* slot_offsets, int S, int D, int64_t total_rows,
                                   int64_t row_stride, int64_t lin_stride, int flags, float* const* table_ptrs,
                                   float* const* lin_ptrs, float* state, const int64_t* step_dev, const float* lr_hist,
                                   int hist_len, float lr, float beta1, float beta2, float eps, void* stream) {
  DR_REQUIRE(rows && slot_offsets && table_ptrs && state && step_dev, DR_EINVAL, "dr_embed_adam_flush: null pointer");
  DR_REQUIRE(S >= 1 && S <= 4096 && D >= 4 && D <= 128 && D % 4 == 0 && total_rows >= 0, DR_EINVAL,
             "dr_embed_adam_flush: S=%d D=%d total_rows=%lld", S, D, (long long)total_rows);
  if (total_rows == 0) return DR_OK;
  if (row_stride == 0) row_stride = D;
  if (lin_stride == 0) lin_stride = 1;
  const int lin_in_row = (flags & DR_EMBED_LIN_IN_ROW) ? 1 : 0;
  const AdamHist h{lr_hist, hist_len, lr, beta1, beta2};
  const int lpr = lpr_of(D);
  int64_t ctas = (total_rows * lpr + 255) / 256;
  if (ctas > (int64_t)kNumSMs * 16) ctas = (int64_t)kNumSMs * 16;
  cudaStream_t st = (cudaStream_t)stream;
#define DR_FLUSH(L)                                                                                                 \
  embed_adam_flush_kernel<L><<<(unsigned)ctas, 256, 0, st>>>(table_ptrs, lin_in_row ? nullptr : lin_ptrs, rows, slot_offsets, S, \
                                                           D, row_stride, lin_stride, lin_in_row, state, adam_state_stride(D), \
                                                           total_rows, step_dev, h, 1.f - beta1, 1.f - beta2, eps)
  switch (lpr) {
    case 1: DR_FLUSH(1); break;
    case 2: DR_FLUSH(2); break;
    case 4: DR_FLUSH(4); break;
    case 8: DR_FLUSH(8); break;
    case 16: DR_FLUSH(16); break;
    default: DR_FLUSH(32); break;
  }
#undef DR_FLUSH
  DR_CUDA_LAUNCH_CHECK("dr_embed_adam_flush");
  return DR_OK;
}
