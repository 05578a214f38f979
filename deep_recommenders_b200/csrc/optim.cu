// SURVEY.md 8(f) #1 -- Adam, the optimizer both reference examples train with
// (examples/train_deepfm_on_movielens_keras.py:44 tf.keras.optimizers.Adam(); examples/train_fm_on_movielens_estimator.py:51
// tf.train.AdamOptimizer(0.01)).  TensorFlow applies Adam DENSELY even to IndexedSlices gradients: m and v of
// every row decay every step and every row of the variable moves (optimizer_v2/adam.py _resource_apply_sparse,
// training/adam.py _apply_sparse_shared: `m_t = assign(m, m * beta1)` over the whole variable, then a
// scatter_add of the touched rows).  dr_adam_step is that dense update in one pass over the parameter arena;
// the row-sparse gradients are accumulated beforehand by dr_embed_fm_bwd / dr_scatter_add with scale = 1.
// HBM-bound: 4 reads + 3 (or 4, zero_grad) writes of n floats per step.
#include "common.cuh"

namespace dr {

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float lr_t, float omb1, float omb2,
                                          float eps) {
  // TF's ApplyAdam functor: m += (g - m)(1 - b1); v += (g*g - v)(1 - b2); var -= lr_t * m / (sqrt(v) + eps)
  m += (g - m) * omb1;
  v += (g * g - v) * omb2;
  p -= (m * lr_t) / (sqrtf(v) + eps);
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                    float lr_t, float omb1, float omb2, float eps, int zero_grad,
                                                    int vec_ok) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n4 = vec_ok ? n / 4 : 0;
  for (int64_t i = tid; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_elem(pp.x, gg.x, mm.x, vv.x, lr_t, omb1, omb2, eps);
    adam_elem(pp.y, gg.y, mm.y, vv.y, lr_t, omb1, omb2, eps);
    adam_elem(pp.z, gg.z, mm.z, vv.z, lr_t, omb1, omb2, eps);
    adam_elem(pp.w, gg.w, mm.w, vv.w, lr_t, omb1, omb2, eps);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = f4_zero();
  }
  for (int64_t i = n4 * 4 + tid; i < n; i += stride) {
    float pp = p[i], mm = m[i], vv = v[i];
    adam_elem(pp, g[i], mm, vv, lr_t, omb1, omb2, eps);
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (zero_grad) g[i] = 0.f;
  }
}

}  // namespace dr

using namespace dr;

extern "C" int dr_adam_step(float* p, float* g, float* m, float* v, int64_t n, float lr_t, float beta1, float beta2,
                            float eps, int zero_grad, void* stream) {
  DR_REQUIRE(n >= 0, DR_EINVAL, "dr_adam_step: n=%lld < 0", (long long)n);
  if (n == 0) return DR_OK;
  DR_REQUIRE(p && g && m && v, DR_EINVAL, "dr_adam_step: null pointer");
  DR_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, DR_EINVAL,
             "dr_adam_step: beta1=%g beta2=%g eps=%g out of range", (double)beta1, (double)beta2, (double)eps);
  const int vec_ok = aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v);
  int64_t ctas = (n / 4 + 255) / 256;
  if (ctas > kNumSMs * 8) ctas = kNumSMs * 8;
  if (ctas < 1) ctas = 1;
  adam_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr_t, 1.f - beta1, 1.f - beta2, eps,
                                                                 zero_grad, vec_ok);
  DR_CUDA_LAUNCH_CHECK("dr_adam_step");
  return DR_OK;
}
