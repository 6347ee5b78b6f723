// One-launch Adam over every parameter of both fields (train.py:934, 2313-2325 build torch.optim.Adam over
// get_optparam_groups and call .step() once per iteration).  The train step keeps all gradients in ONE flat buffer that every
// p.grad aliases (dynrf_b200/backward.py), with each view using its parameter's own dense strides, so element i of a
// parameter's storage pairs with flat_grad[offset + i]; the moments live in two more flat buffers of the same shape.
// torch's foreach / capturable Adam costs ~150 small launches per step on the 66 parameter tensors of this model.
#include "dynrf_common.cuh"

namespace dynrf {

constexpr int kAdamThreads = 256;
constexpr int kAdamPerBlock = kAdamThreads * 8;

struct AdamTable {
  const unsigned long long* param_ptr;  // [T] device pointers (float*)
  const long long* offset;              // [2 T]: start, end of every parameter in the flat buffers (starts are aligned, so padded)
  const int* blk0;                      // [T + 1] first block of each tensor
  int T;
};

__global__ void __launch_bounds__(kAdamThreads)
k_adam_flat(AdamTable tb, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v, float* step_p /*[2]: step, blocks done*/,
            const float* __restrict__ lr_p /*[T] per-tensor learning rate*/, float beta1, float beta2, float eps) {
  // tensor of this block: binary search in blk0
  int lo = 0, hi = tb.T;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((int)blockIdx.x >= __ldg(tb.blk0 + mid)) lo = mid;
    else hi = mid;
  }
  const int t = lo;
  const long long off = __ldg(tb.offset + 2 * t), n = __ldg(tb.offset + 2 * t + 1) - off;
  float* __restrict__ p = reinterpret_cast<float*>(__ldg(tb.param_ptr + t));
  const float step = *reinterpret_cast<volatile float*>(step_p) + 1.0f;   // this update's step number
  const float bc1 = 1.0f - powf(beta1, step), bc2 = 1.0f - powf(beta2, step);
  const float step_size = __ldg(lr_p + t) / bc1;
  const float inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
  const long long base = (long long)(blockIdx.x - __ldg(tb.blk0 + t)) * kAdamPerBlock;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const long long i = base + (long long)k * kAdamThreads + threadIdx.x;
    if (i < n) {
      const float g = grad[off + i];
      const float mm = beta1 * m[off + i] + (1.0f - beta1) * g;     // exp_avg.lerp_(grad, 1 - beta1)
      const float vv = beta2 * v[off + i] + (1.0f - beta2) * g * g; // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      m[off + i] = mm;
      v[off + i] = vv;
      const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
      p[i] -= step_size * (mm / denom);
    }
  }
  // the last block to finish publishes the new step number: every other block has read the old one by then
  if (threadIdx.x == 0) {
    unsigned int* done = reinterpret_cast<unsigned int*>(step_p + 1);
    __threadfence();
    if (atomicAdd(done, 1u) == gridDim.x - 1) {
      *done = 0u;
      step_p[0] = step;
    }
  }
}

}  // namespace dynrf

using namespace dynrf;

extern "C" {

int dynrf_adam_blocks(long long numel) { return (int)((numel + kAdamPerBlock - 1) / kAdamPerBlock); }

/* param_ptr [T] (device array of float* as uint64), offset [2T] (int64: start, end of each parameter), blk0 [T+1] (int32; blk0[t+1]-blk0[t] =
 * dynrf_adam_blocks(numel_t)), lr [T], step [2] (step[0]: float update count, incremented by the call; step[1]: uint32 scratch,
 * zero-initialised by the caller once), grad / m / v flat buffers. */
int dynrf_adam_flat(const uint64_t* param_ptr, const int64_t* offset, const int32_t* blk0, int T, int total_blocks, const float* grad,
                    float* m, float* v, float* step, const float* lr, float beta1, float beta2, float eps, void* stream) {
  if (!param_ptr || !offset || !blk0 || !grad || !m || !v || !step || !lr || T <= 0 || total_blocks <= 0) return DYNRF_E_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  AdamTable tb{reinterpret_cast<const unsigned long long*>(param_ptr), reinterpret_cast<const long long*>(offset), blk0, T};
  k_adam_flat<<<total_blocks, kAdamThreads, 0, s>>>(tb, grad, m, v, step, lr, beta1, beta2, eps);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
