// Fused multi-tensor AdamW (decoupled weight decay), the optimizer the reference trains with:
//   torch.optim.AdamW(model.parameters(), lr=args.lr, weight_decay=1e-2)            (main_eDOS.py:93, main_phDOS.py:90)
// One launch updates up to kMaxTensors parameter tensors: the tensor pointers travel in the kernel-parameter space (no
// device-side pointer table, no host->device copy), blockIdx.y selects the tensor, blockIdx.x a 4096-element chunk.
//   p <- p (1 - lr wd);  m <- m + (1 - b1)(g - m);  v <- b2 v + (1 - b2) g^2;  p <- p - (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// in exactly the operation order of torch's single-tensor implementation.  HBM-bound: 28 bytes per parameter.
#include <math.h>
#include "common.cuh"

namespace dost {

constexpr int kMaxTensors = 32;
constexpr int kChunk = 4096;

struct AdamWPack {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  float* m[kMaxTensors];
  float* v[kMaxTensors];
  long long n[kMaxTensors];
};

__global__ void __launch_bounds__(256) adamw_kernel(const AdamWPack pk, float lr, float beta1, float beta2, float eps, float wd,
                                                    float step_size, float inv_bc2_sqrt) {
  const int t = blockIdx.y;
  const long long n = pk.n[t];
  const long long beg = (long long)blockIdx.x * kChunk;
  if (beg >= n) return;
  const long long end = min64(n, beg + kChunk);
  float* __restrict__ p = pk.p[t];
  const float* __restrict__ g = pk.g[t];
  float* __restrict__ m = pk.m[t];
  float* __restrict__ v = pk.v[t];
  const float decay = 1.f - lr * wd;
  for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const float gi = g[i];
    const float pi = p[i] * decay;
    const float mi = m[i] + (1.f - beta1) * (gi - m[i]);
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;
    const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
}

}  // namespace dost

using namespace dost;

extern "C" int dost_adamw_step(int ntensors, void* const* params, const void* const* grads, void* const* exp_avg,
                               void* const* exp_avg_sq, const long long* numel, double lr, double beta1, double beta2, double eps,
                               double weight_decay, long long step, dost_stream_t stream) {
  DOST_REQUIRE(ntensors >= 0 && params && grads && exp_avg && exp_avg_sq && numel && step >= 1, "adamw_step: bad args");
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1), inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  cudaStream_t st = (cudaStream_t)stream;
  for (int t0 = 0; t0 < ntensors; t0 += kMaxTensors) {
    AdamWPack pk;
    const int nt = (ntensors - t0) < kMaxTensors ? (ntensors - t0) : kMaxTensors;
    long long maxn = 0;
    for (int i = 0; i < kMaxTensors; ++i) {
      const int j = t0 + (i < nt ? i : 0);
      pk.p[i] = (float*)params[j];
      pk.g[i] = (const float*)grads[j];
      pk.m[i] = (float*)exp_avg[j];
      pk.v[i] = (float*)exp_avg_sq[j];
      pk.n[i] = i < nt ? numel[j] : 0;
      DOST_REQUIRE(i >= nt || (pk.p[i] && pk.g[i] && pk.m[i] && pk.v[i]), "adamw_step: null tensor %d", j);
      if (pk.n[i] > maxn) maxn = pk.n[i];
    }
    if (maxn == 0) continue;
    dim3 grid((unsigned)((maxn + kChunk - 1) / kChunk), nt);
    adamw_kernel<<<grid, 256, 0, st>>>(pk, (float)lr, (float)beta1, (float)beta2, (float)eps, (float)weight_decay, step_size,
                                       inv_bc2_sqrt);
    int rc = check_launch("adamw_step");
    if (rc != DOST_OK) return rc;
  }
  return DOST_OK;
}
