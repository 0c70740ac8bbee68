// Row-wise streaming kernels: LayerNorm(+PReLU) forward/backward, fixed-order column sums, PReLU backward.
// All HBM-bound; one warp per row, lane-strided (coalesced) accesses, deterministic two-stage reductions.
#include "common.cuh"

namespace dost {

template <typename T> __device__ __forceinline__ T rsqrt_t(T v);
template <> __device__ __forceinline__ float rsqrt_t<float>(float v) { return 1.0f / sqrtf(v); }
template <> __device__ __forceinline__ double rsqrt_t<double>(double v) { return 1.0 / sqrt(v); }

constexpr int kRowWarps = 8;

// ------------------------------------------------------------------ LayerNorm forward (+ optional PReLU)
template <typename T, int NPL>
__global__ void __launch_bounds__(kRowWarps * 32) ln_fwd_kernel(const T* __restrict__ x, long long ldx,
                                                                const T* __restrict__ gamma,
                                                                const T* __restrict__ beta,
                                                                const T* __restrict__ slope_p, T* __restrict__ y,
                                                                T* __restrict__ stats, long long M, int W) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T gam[NPL], bet[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int h = lane + 32 * i;
    gam[i] = (h < W) ? __ldg(gamma + h) : T(0);
    bet[i] = (h < W) ? __ldg(beta + h) : T(0);
  }
  const bool has_act = slope_p != nullptr;
  const T slope = has_act ? __ldg(slope_p) : T(0);
  const T invW = T(1) / T(W);
  for (long long r = blockIdx.x * (long long)kRowWarps + warp; r < M; r += (long long)gridDim.x * kRowWarps) {
    const T* xr = x + r * ldx;
    T v[NPL];
    T s = T(0);
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int h = lane + 32 * i;
      v[i] = (h < W) ? xr[h] : T(0);
      s += v[i];
    }
    const T mean = warp_sum(s) * invW;
    T q = T(0);
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int h = lane + 32 * i;
      const T d = (h < W) ? v[i] - mean : T(0);
      q += d * d;
    }
    const T rstd = rsqrt_t<T>(warp_sum(q) * invW + T(1e-5));
    T* yr = y + r * (long long)W;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int h = lane + 32 * i;
      if (h < W) {
        T o = (v[i] - mean) * rstd * gam[i] + bet[i];
        if (has_act) o = (o > T(0)) ? o : slope * o;
        yr[h] = o;
      }
    }
    if (lane == 0 && stats) {
      stats[2 * r] = mean;
      stats[2 * r + 1] = rstd;
    }
  }
}

// ------------------------------------------------------------------ LayerNorm backward (+ PReLU)
// ws layout: [nblocks][2*W + 1] = (dgamma partial, dbeta partial, dslope partial)
template <typename T, int NPL>
__global__ void __launch_bounds__(kRowWarps * 32) ln_bwd_kernel(const T* __restrict__ dy, long long ld_dy,
                                                                const T* __restrict__ x, long long ldx,
                                                                const T* __restrict__ stats,
                                                                const T* __restrict__ gamma,
                                                                const T* __restrict__ beta,
                                                                const T* __restrict__ slope_p, T* __restrict__ dx,
                                                                T* __restrict__ ws, long long M, int W,
                                                                long long rows_per_block) {
  __shared__ T red[kRowWarps];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);  // [kRowWarps][2*W]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T gam[NPL], bet[NPL], dgam[NPL], dbet[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int h = lane + 32 * i;
    gam[i] = (h < W) ? __ldg(gamma + h) : T(0);
    bet[i] = (h < W) ? __ldg(beta + h) : T(0);
    dgam[i] = T(0);
    dbet[i] = T(0);
  }
  const bool has_act = slope_p != nullptr;
  const T slope = has_act ? __ldg(slope_p) : T(0);
  double dsl = 0.0;  // the PReLU slope gradient is one scalar summed over M*W terms: accumulate it in fp64
  const T invW = T(1) / T(W);
  const long long rbeg = blockIdx.x * rows_per_block;
  const long long rend = min(M, rbeg + rows_per_block);
  for (long long r = rbeg + warp; r < rend; r += kRowWarps) {
    const T mean = stats[2 * r], rstd = stats[2 * r + 1];
    const T* xr = x + r * ldx;
    const T* gr = dy + r * ld_dy;
    T xh[NPL], dxh[NPL];
    T s1 = T(0), s2 = T(0);
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int h = lane + 32 * i;
      if (h < W) {
        xh[i] = (xr[h] - mean) * rstd;
        T g = gr[h];
        if (has_act) {
          const T o = xh[i] * gam[i] + bet[i];
          if (!(o > T(0))) {
            dsl += (double)g * (double)o;
            g *= slope;
          }
        }
        dgam[i] += g * xh[i];
        dbet[i] += g;
        dxh[i] = g * gam[i];
        s1 += dxh[i];
        s2 += dxh[i] * xh[i];
      } else {
        xh[i] = T(0);
        dxh[i] = T(0);
      }
    }
    s1 = warp_sum(s1) * invW;
    s2 = warp_sum(s2) * invW;
    T* dxr = dx + r * (long long)W;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int h = lane + 32 * i;
      if (h < W) dxr[h] = rstd * (dxh[i] - s1 - xh[i] * s2);
    }
  }
  // combine the warps of this block in a fixed order
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int h = lane + 32 * i;
    if (h < W) {
      sm[(warp * 2 + 0) * W + h] = dgam[i];
      sm[(warp * 2 + 1) * W + h] = dbet[i];
    }
  }
  dsl = warp_sum(dsl);
  if (lane == 0) red[warp] = (T)dsl;
  __syncthreads();
  T* wsb = ws + (long long)blockIdx.x * (2 * W + 1);
  for (int c = threadIdx.x; c < 2 * W; c += blockDim.x) {
    const int which = c / W, h = c % W;
    T s = T(0);
#pragma unroll
    for (int w = 0; w < kRowWarps; ++w) s += sm[(w * 2 + which) * W + h];
    wsb[c] = s;
  }
  if (threadIdx.x == 0) {
    T s = T(0);
#pragma unroll
    for (int w = 0; w < kRowWarps; ++w) s += red[w];
    wsb[2 * W] = s;
  }
}

// out_k[c] = sum_b ws[b][off_k + c], blocks visited in order (deterministic)
template <typename T>
__global__ void reduce_partials_kernel(const T* __restrict__ ws, int nblk, int ncols, int W, T* __restrict__ o0,
                                       T* __restrict__ o1, T* __restrict__ o2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  T s = T(0);
  for (int b = 0; b < nblk; ++b) s += ws[(long long)b * ncols + c];
  if (c < W) {
    if (o0) o0[c] = s;
  } else if (c < 2 * W) {
    if (o1) o1[c - W] = s;
  } else {
    if (o2) o2[c - 2 * W] = s;
  }
}

// ------------------------------------------------------------------ column sums
template <typename T>
__global__ void __launch_bounds__(256) colsum_stage1_kernel(const T* __restrict__ x, long long ld, long long M,
                                                            long long W, long long rows_per_chunk,
                                                            T* __restrict__ dst) {
  __shared__ T sm[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const long long col = blockIdx.x * 32LL + cx;
  const long long rbeg = blockIdx.y * rows_per_chunk, rend = min(M, rbeg + rows_per_chunk);
  T s = T(0);
  if (col < W) {
    long long r = rbeg + ry;
    T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
    for (; r + 24 < rend; r += 32) {
      s0 += x[r * ld + col];
      s1 += x[(r + 8) * ld + col];
      s2 += x[(r + 16) * ld + col];
      s3 += x[(r + 24) * ld + col];
    }
    for (; r < rend; r += 8) s0 += x[r * ld + col];
    s = (s0 + s1) + (s2 + s3);
  }
  sm[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && col < W) {
    T t = T(0);
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sm[k][cx];
    dst[(long long)blockIdx.y * W + col] = t;
  }
}

template <typename T>
__global__ void colsum_stage2_kernel(const T* __restrict__ ws, int nchunks, long long W, T* __restrict__ out) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= W) return;
  T s = T(0);
  for (int b = 0; b < nchunks; ++b) s += ws[(long long)b * W + c];
  out[c] = s;
}

static inline int colsum_chunks(long long M, long long W) {
  const long long gx = (W + 31) / 32;
  long long nch = (4LL * kNumSMs + gx - 1) / gx;
  const long long maxch = (M + 63) / 64;
  if (nch > maxch) nch = maxch;
  if (nch < 1) nch = 1;
  if (nch > 65535) nch = 65535;
  return (int)nch;
}

// ------------------------------------------------------------------ PReLU backward
template <typename T>
__global__ void __launch_bounds__(256) prelu_bwd_kernel(const T* __restrict__ da, const T* __restrict__ z,
                                                        const T* __restrict__ slope_p, T* __restrict__ dz,
                                                        T* __restrict__ ws, long long n, long long per_block) {
  __shared__ double red[8];
  const T slope = __ldg(slope_p);
  const long long beg = blockIdx.x * per_block, end = min(n, beg + per_block);
  double acc = 0.0;
  for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const T zz = z[i], g = da[i];
    if (zz > T(0)) {
      dz[i] = g;
    } else {
      dz[i] = g * slope;
      acc += (double)g * (double)zz;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    ws[blockIdx.x] = (T)s;
  }
}

template <typename T>
__global__ void sum_small_kernel(const T* __restrict__ ws, int n, T* __restrict__ out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    T s = T(0);
    for (int i = 0; i < n; ++i) s += ws[i];
    *out = s;
  }
}

static inline int npl_for(int W) {
  int npl = (W + 31) / 32, p = 1;
  while (p < npl) p <<= 1;
  return p;
}

static inline int ln_bwd_blocks(long long M) {
  long long nb = (M + kRowWarps - 1) / kRowWarps;
  if (nb > 4LL * kNumSMs) nb = 4LL * kNumSMs;
  if (nb < 1) nb = 1;
  return (int)nb;
}

template <typename T>
static int run_ln_fwd(const void* x, long long ldx, const void* gamma, const void* beta, const void* slope, void* y,
                      void* stats, long long M, int W, cudaStream_t st) {
  const int npl = npl_for(W);
  int blocks = (int)min64((M + kRowWarps - 1) / kRowWarps, 16LL * kNumSMs);
  if (blocks < 1) blocks = 1;
#define DOST_LN_FWD(NPL)                                                                                       \
  case NPL:                                                                                                    \
    ln_fwd_kernel<T, NPL><<<blocks, kRowWarps * 32, 0, st>>>((const T*)x, ldx, (const T*)gamma, (const T*)beta, \
                                                             (const T*)slope, (T*)y, (T*)stats, M, W);         \
    break;
  switch (npl) {
    DOST_LN_FWD(1) DOST_LN_FWD(2) DOST_LN_FWD(4) DOST_LN_FWD(8) DOST_LN_FWD(16) DOST_LN_FWD(32)
    default:
      set_error("ln_fwd: width %d > 1024 unsupported", W);
      return DOST_ERR_UNSUPPORTED;
  }
#undef DOST_LN_FWD
  return check_launch("ln_fwd");
}

template <typename T>
static int run_ln_bwd(const void* dy, long long ld_dy, const void* x, long long ldx, const void* stats,
                      const void* gamma, const void* beta, const void* slope, void* dx, void* dgamma, void* dbeta,
                      void* dslope, long long M, int W, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int npl = npl_for(W);
  const int blocks = ln_bwd_blocks(M);
  const size_t need = sizeof(T) * (size_t)blocks * (2 * W + 1);
  if (!workspace || workspace_bytes < need) {
    set_error("ln_bwd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return DOST_ERR_WORKSPACE;
  }
  const long long rpb = (M + blocks - 1) / blocks;
  const size_t smem = sizeof(T) * (size_t)kRowWarps * 2 * W;
#define DOST_LN_BWD(NPL)                                                                                          \
  case NPL:                                                                                                       \
    if (smem > 48 * 1024)                                                                                         \
      cudaFuncSetAttribute(ln_bwd_kernel<T, NPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    ln_bwd_kernel<T, NPL><<<blocks, kRowWarps * 32, smem, st>>>((const T*)dy, ld_dy, (const T*)x, ldx,            \
                                                                (const T*)stats, (const T*)gamma, (const T*)beta, \
                                                                (const T*)slope, (T*)dx, (T*)workspace, M, W, rpb); \
    break;
  switch (npl) {
    DOST_LN_BWD(1) DOST_LN_BWD(2) DOST_LN_BWD(4) DOST_LN_BWD(8) DOST_LN_BWD(16) DOST_LN_BWD(32)
    default:
      set_error("ln_bwd: width %d > 1024 unsupported", W);
      return DOST_ERR_UNSUPPORTED;
  }
#undef DOST_LN_BWD
  int rc = check_launch("ln_bwd");
  if (rc != DOST_OK) return rc;
  const int ncols = 2 * W + 1;
  reduce_partials_kernel<T><<<ceil_div(ncols, 128), 128, 0, st>>>((const T*)workspace, blocks, ncols, W, (T*)dgamma,
                                                                  (T*)dbeta, (T*)dslope);
  return check_launch("ln_bwd reduce");
}

template <typename T>
static int run_colsum(const void* x, long long ld, long long M, long long W, void* out, void* workspace,
                      size_t workspace_bytes, cudaStream_t st) {
  const int nch = colsum_chunks(M, W);
  const long long rpc = (M + nch - 1) / nch;
  dim3 grid((unsigned)((W + 31) / 32), nch);
  if (nch == 1) {
    colsum_stage1_kernel<T><<<grid, 256, 0, st>>>((const T*)x, ld, M, W, rpc, (T*)out);
    return check_launch("colsum");
  }
  const size_t need = sizeof(T) * (size_t)nch * W;
  if (!workspace || workspace_bytes < need) {
    set_error("colsum: workspace too small (%zu < %zu)", workspace_bytes, need);
    return DOST_ERR_WORKSPACE;
  }
  colsum_stage1_kernel<T><<<grid, 256, 0, st>>>((const T*)x, ld, M, W, rpc, (T*)workspace);
  int rc = check_launch("colsum stage1");
  if (rc != DOST_OK) return rc;
  colsum_stage2_kernel<T><<<ceil_div(W, 256), 256, 0, st>>>((const T*)workspace, nch, W, (T*)out);
  return check_launch("colsum stage2");
}

static inline int prelu_blocks(long long n) {
  long long nb = (n + 4095) / 4096;
  if (nb > 4LL * kNumSMs) nb = 4LL * kNumSMs;
  if (nb < 1) nb = 1;
  return (int)nb;
}

template <typename T>
static int run_prelu_bwd(const void* da, const void* z, const void* slope, void* dz, void* dslope, long long n,
                         void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int blocks = prelu_blocks(n);
  if (!workspace || workspace_bytes < sizeof(T) * (size_t)blocks) {
    set_error("prelu_bwd: workspace too small");
    return DOST_ERR_WORKSPACE;
  }
  const long long per = (n + blocks - 1) / blocks;
  prelu_bwd_kernel<T><<<blocks, 256, 0, st>>>((const T*)da, (const T*)z, (const T*)slope, (T*)dz, (T*)workspace, n, per);
  int rc = check_launch("prelu_bwd");
  if (rc != DOST_OK) return rc;
  sum_small_kernel<T><<<1, 32, 0, st>>>((const T*)workspace, blocks, (T*)dslope);
  return check_launch("prelu_bwd reduce");
}

}  // namespace dost

using namespace dost;

#define DOST_DISPATCH(dtype, CALL_F32, CALL_F64, name)   \
  if ((dtype) == DOST_F32) return CALL_F32;              \
  if ((dtype) == DOST_F64) return CALL_F64;              \
  set_error(name ": unsupported dtype %d", (int)(dtype)); \
  return DOST_ERR_UNSUPPORTED;

extern "C" int dost_ln_fwd(int dtype, const void* x, long long ldx, const void* gamma, const void* beta,
                           const void* prelu_slope, void* y, void* stats, long long M, int W, dost_stream_t stream) {
  if (M == 0) return DOST_OK;
  DOST_REQUIRE(x && gamma && beta && y && M > 0 && W > 0, "ln_fwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  DOST_DISPATCH(dtype, run_ln_fwd<float>(x, ldx, gamma, beta, prelu_slope, y, stats, M, W, st),
                run_ln_fwd<double>(x, ldx, gamma, beta, prelu_slope, y, stats, M, W, st), "ln_fwd")
}

extern "C" size_t dost_ln_bwd_workspace_bytes(int dtype, long long M, int W) {
  return (dtype == DOST_F64 ? 8 : 4) * (size_t)ln_bwd_blocks(M) * (2 * W + 1);
}

extern "C" int dost_ln_bwd(int dtype, const void* dy, long long ld_dy, const void* x, long long ldx,
                           const void* stats, const void* gamma, const void* beta, const void* prelu_slope, void* dx,
                           void* dgamma, void* dbeta, void* dslope, long long M, int W, void* workspace,
                           size_t workspace_bytes, dost_stream_t stream) {
  DOST_REQUIRE(dy && x && stats && gamma && beta && dx && M > 0 && W > 0, "ln_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  DOST_DISPATCH(dtype,
                run_ln_bwd<float>(dy, ld_dy, x, ldx, stats, gamma, beta, prelu_slope, dx, dgamma, dbeta, dslope, M, W,
                                  workspace, workspace_bytes, st),
                run_ln_bwd<double>(dy, ld_dy, x, ldx, stats, gamma, beta, prelu_slope, dx, dgamma, dbeta, dslope, M, W,
                                   workspace, workspace_bytes, st),
                "ln_bwd")
}

extern "C" size_t dost_colsum_workspace_bytes(int dtype, long long M, long long W) {
  const int nch = colsum_chunks(M, W);
  return nch == 1 ? 0 : (dtype == DOST_F64 ? 8 : 4) * (size_t)nch * W;
}

extern "C" int dost_colsum(int dtype, const void* x, long long ld, long long M, long long W, void* out,
                           void* workspace, size_t workspace_bytes, dost_stream_t stream) {
  DOST_REQUIRE(x && out && M > 0 && W > 0, "colsum: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  DOST_DISPATCH(dtype, run_colsum<float>(x, ld, M, W, out, workspace, workspace_bytes, st),
                run_colsum<double>(x, ld, M, W, out, workspace, workspace_bytes, st), "colsum")
}

extern "C" size_t dost_prelu_bwd_workspace_bytes(int dtype, long long n) {
  return (dtype == DOST_F64 ? 8 : 4) * (size_t)prelu_blocks(n);
}

extern "C" int dost_prelu_bwd(int dtype, const void* da, const void* z, const void* slope, void* dz, void* dslope,
                              long long n, void* workspace, size_t workspace_bytes, dost_stream_t stream) {
  DOST_REQUIRE(da && z && slope && dz && dslope && n > 0, "prelu_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  DOST_DISPATCH(dtype, run_prelu_bwd<float>(da, z, slope, dz, dslope, n, workspace, workspace_bytes, st),
                run_prelu_bwd<double>(da, z, slope, dz, dslope, n, workspace, workspace_bytes, st), "prelu_bwd")
}
