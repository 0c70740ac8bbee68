// DOS loss forward/backward (main_eDOS.py:111-123, main_phDOS.py:109-114), fused and deterministic.
#include "common.cuh"

namespace dost {

// One block per (crystal, branch): sse[branch*B + b] = sum_t (y - pred)^2, fixed-order reduction.
template <typename T>
__global__ void __launch_bounds__(128) loss_sse_kernel(const T* __restrict__ pg, const T* __restrict__ ps,
                                                       const T* __restrict__ y, int clamp, int B, int Tn,
                                                       T* __restrict__ sse) {
  __shared__ T red[4];
  const int b = blockIdx.x, br = blockIdx.y;
  const T* p = (br == 0 ? pg : ps) + (long long)b * Tn;
  const T* yy = y + (long long)b * Tn;
  T acc = T(0);
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) {
    T yt = yy[t];
    if (clamp && yt < T(0)) yt = T(0);
    const T d = yt - p[t];
    acc = fma(d, d, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) sse[br * B + b] = (red[0] + red[1]) + (red[2] + red[3]);
}

// mode 0: saved[br*B+b] = rmse_b; loss = mean_b rmse_g + beta * mean_b rmse_s.
// mode 1: saved[br] = sqrt(sum_b sse / (B*T)); loss = saved[0] + beta * saved[1].
template <typename T>
__global__ void __launch_bounds__(256) loss_final_kernel(const T* __restrict__ sse, int mode, T beta, int B, int Tn,
                                                         T* __restrict__ loss, T* __restrict__ saved) {
  __shared__ T red[2][256];
  for (int br = 0; br < 2; ++br) {
    T acc = T(0);
    // strided but fixed assignment of crystals to threads -> deterministic
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
      const T v = sse[br * B + b];
      if (mode == 0) {
        const T r = sqrt(v / T(Tn));
        saved[br * B + b] = r;
        acc += r;
      } else {
        acc += v;
      }
    }
    red[br][threadIdx.x] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    T tot[2];
    for (int br = 0; br < 2; ++br) {
      T s = T(0);
      for (int i = 0; i < 256; ++i) s += red[br][i];
      tot[br] = s;
    }
    if (mode == 0) {
      *loss = tot[0] / T(B) + beta * (tot[1] / T(B));
    } else {
      const T r0 = sqrt(tot[0] / (T(B) * T(Tn))), r1 = sqrt(tot[1] / (T(B) * T(Tn)));
      saved[0] = r0;
      saved[1] = r1;
      *loss = r0 + beta * r1;
    }
  }
}

template <typename T>
__global__ void loss_bwd_kernel(const T* __restrict__ pg, const T* __restrict__ ps, const T* __restrict__ y,
                                int mode, int clamp, T beta, int B, int Tn, const T* __restrict__ saved,
                                const T* __restrict__ gl, T* __restrict__ dpg, T* __restrict__ dps) {
  const long long n = (long long)B * Tn;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = (int)(i / Tn);
  const T g = gl ? *gl : T(1);
  T yt = y[i];
  if (clamp && yt < T(0)) yt = T(0);
  // d rmse / d pred = -(y - p) / (T * rmse)  [per crystal]   or  -(y - p) / (B*T*rmse)  [batch-wide]
  const T rg = (mode == 0) ? saved[b] : saved[0];
  const T rs = (mode == 0) ? saved[B + b] : saved[1];
  const T denom = T(B) * T(Tn);
  dpg[i] = g * (pg[i] - yt) / (denom * rg);
  dps[i] = g * beta * (ps[i] - yt) / (denom * rs);
}

template <typename T>
static int run_loss_fwd(int mode, const void* pg, const void* ps, const void* y, double beta, int B, int Tn, void* loss,
                        void* saved, void* sse, cudaStream_t st) {
  dim3 grid(B, 2);
  loss_sse_kernel<T><<<grid, 128, 0, st>>>((const T*)pg, (const T*)ps, (const T*)y, mode == 0, B, Tn, (T*)sse);
  int rc = check_launch("loss_sse");
  if (rc != DOST_OK) return rc;
  loss_final_kernel<T><<<1, 256, 0, st>>>((const T*)sse, mode, (T)beta, B, Tn, (T*)loss, (T*)saved);
  return check_launch("loss_final");
}

}  // namespace dost

using namespace dost;

// saved must hold 4*B elements: [0,2B) rmse values (mode 0) / [0,2) (mode 1); [2B,4B) scratch for the sse partials.
extern "C" int dost_loss_fwd(int dtype, int mode, const void* pred_g, const void* pred_s, const void* y, double beta,
                             int B, int T, void* loss, void* saved, dost_stream_t stream) {
  DOST_REQUIRE(pred_g && pred_s && y && loss && saved && B > 0 && T > 0, "loss_fwd: bad args");
  DOST_REQUIRE(mode == 0 || mode == 1, "loss_fwd: mode must be 0 (eDOS) or 1 (phonon)");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DOST_F32)
    return run_loss_fwd<float>(mode, pred_g, pred_s, y, beta, B, T, loss, saved, (float*)saved + 2 * (size_t)B, st);
  if (dtype == DOST_F64)
    return run_loss_fwd<double>(mode, pred_g, pred_s, y, beta, B, T, loss, saved, (double*)saved + 2 * (size_t)B, st);
  set_error("loss_fwd: unsupported dtype %d", dtype);
  return DOST_ERR_UNSUPPORTED;
}

extern "C" int dost_loss_bwd(int dtype, int mode, const void* pred_g, const void* pred_s, const void* y, double beta,
                             int B, int T, const void* saved, const void* grad_loss, void* d_pred_g, void* d_pred_s,
                             dost_stream_t stream) {
  DOST_REQUIRE(pred_g && pred_s && y && saved && d_pred_g && d_pred_s && B > 0 && T > 0, "loss_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * T;
  const int blocks = ceil_div(n, 256);
  if (dtype == DOST_F32)
    loss_bwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)pred_g, (const float*)pred_s, (const float*)y, mode,
                                                   mode == 0, (float)beta, B, T, (const float*)saved,
                                                   (const float*)grad_loss, (float*)d_pred_g, (float*)d_pred_s);
  else if (dtype == DOST_F64)
    loss_bwd_kernel<double><<<blocks, 256, 0, st>>>((const double*)pred_g, (const double*)pred_s, (const double*)y,
                                                    mode, mode == 0, (double)beta, B, T, (const double*)saved,
                                                    (const double*)grad_loss, (double*)d_pred_g, (double*)d_pred_s);
  else {
    set_error("loss_bwd: unsupported dtype %d", dtype);
    return DOST_ERR_UNSUPPORTED;
  }
  return check_launch("loss_bwd");
}

// ------------------------------------------------------------------------------------------------ evaluation metrics
// utils.test (utils.py:61-112) evaluates with batch_size = 1: per crystal, targets and predictions clamped at 0, then
// MSE, RMSE, MAE and R^2 = 1 - SSE / sum (y - mean y)^2 over its T energies; the epoch numbers are the means over the
// crystals.  One block per crystal writes (mse, rmse, mae, r2); dost_eval_metrics reduces them in a fixed order.
namespace dost {

template <typename T>
__global__ void __launch_bounds__(128) eval_crystal_kernel(const T* __restrict__ pred, const T* __restrict__ y, int clamp, int Tn,
                                                           T* __restrict__ per) {
  // `clamp`: eDOS (utils.py:75-76 clamps BOTH the target and the prediction at 0); phonon (utils.py:127-131) clamps neither.
  __shared__ T red[4][4];
  __shared__ T mean_s;
  const int b = blockIdx.x;
  const T* p = pred + (long long)b * Tn;
  const T* yy = y + (long long)b * Tn;
  T sse = T(0), sae = T(0), sy = T(0);
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) {
    T yt = yy[t], pt = p[t];
    if (clamp && yt < T(0)) yt = T(0);
    if (clamp && pt < T(0)) pt = T(0);
    const T d = yt - pt;
    sse = fma(d, d, sse);
    sae += (d < T(0) ? -d : d);
    sy += yt;
  }
  sse = warp_sum(sse); sae = warp_sum(sae); sy = warp_sum(sy);
  if ((threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    red[0][w] = sse; red[1][w] = sae; red[2][w] = sy;
  }
  __syncthreads();
  if (threadIdx.x == 0) mean_s = ((red[2][0] + red[2][1]) + (red[2][2] + red[2][3])) / T(Tn);
  __syncthreads();
  // total sum of squares around the mean in a second pass (sum y^2 - (sum y)^2 / T cancels in fp32)
  const T mean = mean_s;
  T sst = T(0);
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) {
    T yt = yy[t];
    if (clamp && yt < T(0)) yt = T(0);
    const T d = yt - mean;
    sst = fma(d, d, sst);
  }
  sst = warp_sum(sst);
  if ((threadIdx.x & 31) == 0) red[3][threadIdx.x >> 5] = sst;
  __syncthreads();
  if (threadIdx.x == 0) {
    T v[4];
    for (int k = 0; k < 4; ++k) v[k] = (red[k][0] + red[k][1]) + (red[k][2] + red[k][3]);
    const T mse = v[0] / T(Tn);
    per[4 * b + 0] = mse;
    per[4 * b + 1] = sqrt(mse);
    per[4 * b + 2] = v[1] / T(Tn);
    // sklearn.metrics.r2_score of the flattened crystal (utils.py:20-23); a constant target (sst == 0) scores 1 for a
    // perfect prediction and 0 otherwise (sklearn's force_finite), not -inf / NaN
    per[4 * b + 3] = (v[3] > T(0)) ? T(1) - v[0] / v[3] : (v[0] == T(0) ? T(1) : T(0));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) eval_mean_kernel(const T* __restrict__ per, int B, T* __restrict__ out) {
  __shared__ double red[4][256];
  for (int k = 0; k < 4; ++k) {
    double acc = 0.0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) acc += (double)per[4 * b + k];
    red[k][threadIdx.x] = acc;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int i = 0; i < 256; ++i) s += red[threadIdx.x][i];
    out[threadIdx.x] = (T)(s / B);
  }
}

template <typename T>
static int run_eval(const void* pred, const void* y, int clamp_pred, int B, int Tn, void* per, void* mean, cudaStream_t st) {
  eval_crystal_kernel<T><<<B, 128, 0, st>>>((const T*)pred, (const T*)y, clamp_pred, Tn, (T*)per);
  int rc = check_launch("eval_metrics per crystal");
  if (rc != DOST_OK || !mean) return rc;
  eval_mean_kernel<T><<<1, 256, 0, st>>>((const T*)per, B, (T*)mean);
  return check_launch("eval_metrics mean");
}

}  // namespace dost

extern "C" int dost_eval_metrics(int dtype, const void* pred, const void* y, int clamp_pred, int B, int T, void* per_crystal, void* mean,
                                 dost_stream_t stream) {
  DOST_REQUIRE(pred && y && per_crystal && B > 0 && T > 0, "eval_metrics: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DOST_F32) return dost::run_eval<float>(pred, y, clamp_pred, B, T, per_crystal, mean, st);
  if (dtype == DOST_F64) return dost::run_eval<double>(pred, y, clamp_pred, B, T, per_crystal, mean, st);
  dost::set_error("eval_metrics: unsupported dtype %d", dtype);
  return DOST_ERR_UNSUPPORTED;
}
