// Fused AdamW over one flat fp32 parameter / gradient bucket (sm_100a).            SURVEY.md §8 f3
//
// Replaces `optim.step()` of torch.optim.AdamW(weight_decay=1e-7) with the per-step learning
// rate of transduction_model.py:178-189,210.  The model's parameters, gradients (the
// data-parallel all-reduce bucket), exp_avg and exp_avg_sq are each ONE contiguous buffer, so the
// whole update is a single streaming pass: 16 B read + 12 B written per element (HBM-bound,
// 1.49 GB at 53.3 M parameters), instead of ~100 per-tensor launches.  The 1/world_size of the
// gradient mean is folded in (the separate scaling pass over the bucket disappears).
//
// The learning rate and the step counter live in device cells: a captured CUDA graph replays the
// update with whatever the schedule wrote into `lr_cell`, and `adamw_tick` advances the step
// (bias corrections) at execution time.
#include "ssb_common.cuh"

namespace {

__global__ void adamw_tick_kernel(int64_t* step_cell) { *step_cell += 1; }

__global__ void __launch_bounds__(256)
adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                  float* __restrict__ v, int64_t n4, int64_t n, const float* __restrict__ lr_cell,
                  const int64_t* __restrict__ step_cell, float beta1, float beta2, float eps,
                  float weight_decay, float grad_scale) {
  const float lr = __ldg(lr_cell);
  const double step = (double)__ldg(step_cell);
  // torch.optim.AdamW (single-tensor path) arithmetic, bias corrections in double like the host
  const float bc1 = (float)(1.0 - pow((double)beta1, step));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, step));
  const float step_size = lr / bc1;
  const float decay = 1.f - lr * weight_decay;
  const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= grad_scale;
    pp *= decay;
    mm = mm + omb1 * (gg - mm);                 // exp_avg.lerp_(grad, 1 - beta1)
    vv = vv * beta2 + omb2 * gg * gg;           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp -= step_size * (mm / denom);
  };
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y);
    upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail (n % 4 elements)
  const int64_t t = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) upd(p[t], g[t], m[t], v[t]);
}

}  // namespace

extern "C" {

int ssb_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_cell,
                   int64_t* step_cell, float beta1, float beta2, float eps, float weight_decay,
                   float grad_scale, void* stream) {
  if (n == 0) return SSB_OK;
  SSB_REQUIRE(p && g && m && v && lr_cell && step_cell && n > 0, "adamw_flat: null argument");
  SSB_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0,
              "adamw_flat: buffers must be 16 B aligned");
  SSB_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f,
              "adamw_flat: bad hyper-parameters");
  cudaStream_t st = (cudaStream_t)stream;
  adamw_tick_kernel<<<1, 1, 0, st>>>(step_cell);
  SSB_LAUNCH_CHECK("adamw_tick");
  const int64_t n4 = n / 4;
  const int64_t want = (n4 + 255) / 256;
  const int64_t cap = (int64_t)ssb::num_sms() * 8;
  const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
  adamw_flat_kernel<<<grid, 256, 0, st>>>(p, g, m, v, n4, n, lr_cell, step_cell, beta1, beta2, eps,
                                          weight_decay, grad_scale);
  SSB_LAUNCH_CHECK("adamw_flat");
  return SSB_OK;
}

}  // extern "C"
