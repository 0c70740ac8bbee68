// fp32 row kernels of the tensor-core path: LayerNorm (+PReLU) forward/backward with 16-byte vector accesses that
// write their results directly as bf16 hi/lo operand planes for dost_gemm_bf16 (no separate conversion pass), and
// column sums over planes (bias gradients).  One warp per row, W = 128 * NV columns (NV float4 per lane), fixed-order
// two-stage reductions (deterministic).  HBM-bound: fwd reads 4 B and writes 4 B (planes) per element.
#include <cuda_bf16.h>
#include "common.cuh"

namespace dost {
namespace rbf {

constexpr int kWarps = 8;

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void split4(const float4& v, uint2& hi, uint2& lo) {
  hi.x = pack_bf16(v.x, v.y);
  hi.y = pack_bf16(v.z, v.w);
  lo.x = pack_bf16(v.x - __uint_as_float(hi.x << 16), v.y - __uint_as_float(hi.x & 0xFFFF0000u));
  lo.y = pack_bf16(v.z - __uint_as_float(hi.y << 16), v.w - __uint_as_float(hi.y & 0xFFFF0000u));
}
__device__ __forceinline__ float4 unsplit4(const uint2& hi, const uint2& lo) {
  float4 v;
  v.x = __uint_as_float(hi.x << 16) + __uint_as_float(lo.x << 16);
  v.y = __uint_as_float(hi.x & 0xFFFF0000u) + __uint_as_float(lo.x & 0xFFFF0000u);
  v.z = __uint_as_float(hi.y << 16) + __uint_as_float(lo.y << 16);
  v.w = __uint_as_float(hi.y & 0xFFFF0000u) + __uint_as_float(lo.y & 0xFFFF0000u);
  return v;
}

// ---------------------------------------------------------------------------------------------- LayerNorm forward
// y = LN(x) * gamma + beta, optionally PReLU; outputs: fp32 y (optional), planes (optional), stats (mean, rstd).
// Optional gather-add prologue (split-weight message passing): x[r] += ga[ia[r]] + gb[ib[r]] (rows of width W, pitch
// ldg), and the sum is written back to x (it is the LayerNorm input the backward needs).
template <int NV>
__global__ void __launch_bounds__(kWarps * 32) ln_fwd_kernel(float* __restrict__ x, long long ldx, const float* __restrict__ ga,
                                                             const int* __restrict__ ia, const float* __restrict__ gb,
                                                             const int* __restrict__ ib, long long ldg,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const float* __restrict__ slope_p, float* __restrict__ y,
                                                             __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                             long long ldp, float* __restrict__ stats, long long M) {
  constexpr int W = 128 * NV;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 gam[NV], bet[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    gam[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    bet[i] = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
  }
  const bool has_act = slope_p != nullptr;
  const float slope = has_act ? __ldg(slope_p) : 0.f;
  const float invW = 1.f / W;
  // gather indices of this warp's NEXT row are requested one iteration ahead: the gathers then depend on a value that is
  // already there instead of adding an index round trip to every row's latency chain
  const long long rstep = (long long)gridDim.x * kWarps;
  long long r = blockIdx.x * (long long)kWarps + warp;
  long long ra_n = 0, rb_n = 0;
  if (ga && r < M) {
    ra_n = __ldg(ia + r);
    rb_n = __ldg(ib + r);
  }
  for (; r < M; r += rstep) {
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      v[i] = *(reinterpret_cast<const float4*>(x + r * ldx) + lane + 32 * i);
    if (ga) {
      const long long ra = ra_n, rb = rb_n;
      if (r + rstep < M) {
        ra_n = __ldg(ia + r + rstep);
        rb_n = __ldg(ib + r + rstep);
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(ga + ra * ldg) + lane + 32 * i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(gb + rb * ldg) + lane + 32 * i);
        v[i].x += a.x + b.x; v[i].y += a.y + b.y; v[i].z += a.z + b.z; v[i].w += a.w + b.w;
        *(reinterpret_cast<float4*>(x + r * ldx) + lane + 32 * i) = v[i];
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * invW;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * invW + 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o;
      o.x = (v[i].x - mean) * rstd * gam[i].x + bet[i].x;
      o.y = (v[i].y - mean) * rstd * gam[i].y + bet[i].y;
      o.z = (v[i].z - mean) * rstd * gam[i].z + bet[i].z;
      o.w = (v[i].w - mean) * rstd * gam[i].w + bet[i].w;
      if (has_act) {
        o.x = (o.x > 0.f) ? o.x : slope * o.x;
        o.y = (o.y > 0.f) ? o.y : slope * o.y;
        o.z = (o.z > 0.f) ? o.z : slope * o.z;
        o.w = (o.w > 0.f) ? o.w : slope * o.w;
      }
      if (y) reinterpret_cast<float4*>(y + r * (long long)W)[lane + 32 * i] = o;
      if (hi) {
        uint2 h, l;
        split4(o, h, l);
        reinterpret_cast<uint2*>(hi + r * ldp)[lane + 32 * i] = h;
        if (lo) reinterpret_cast<uint2*>(lo + r * ldp)[lane + 32 * i] = l;
      }
    }
    if (lane == 0 && stats) {
      stats[2 * r] = mean;
      stats[2 * r + 1] = rstd;
    }
  }
}

// ---------------------------------------------------------------------------------------------- LayerNorm backward
// dx = LN'(dy) (+ dres); also dgamma, dbeta, dslope and (optionally) the column sums of the stored dx (the bias gradient of the
// Linear that feeds the LayerNorm).  dx is written as fp32 and/or as planes.
// ws layout: [nblocks][3 * W + 1] = (dgamma, dbeta, dxsum, dslope) partials.
template <int NV>
__global__ void __launch_bounds__(kWarps * 32) ln_bwd_kernel(const float* __restrict__ dy, long long ld_dy,
                                                             const float* __restrict__ x, long long ldx,
                                                             const float* __restrict__ stats, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ slope_p,
                                                             const float* __restrict__ dres, long long ld_dres,
                                                             float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_hi,
                                                             __nv_bfloat16* __restrict__ dx_lo, long long ldp,
                                                             float* __restrict__ ws, long long M, long long rows_per_block) {
  constexpr int W = 128 * NV;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sm = reinterpret_cast<float*>(smem_raw);  // [kWarps][3 * W]
  __shared__ double red[kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 gam[NV], bet[NV], dgam[NV], dbet[NV], dxs[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    gam[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    bet[i] = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
    dgam[i] = dbet[i] = dxs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const bool has_act = slope_p != nullptr;
  const float slope = has_act ? __ldg(slope_p) : 0.f;
  double dsl = 0.0;
  const float invW = 1.f / W;
  const long long rbeg = blockIdx.x * rows_per_block;
  const long long rend = min(M, rbeg + rows_per_block);
  for (long long r = rbeg + warp; r < rend; r += kWarps) {
    const float mean = __ldg(stats + 2 * r), rstd = __ldg(stats + 2 * r + 1);
    float4 xh[NV], g[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      xh[i] = __ldg(reinterpret_cast<const float4*>(x + r * ldx) + lane + 32 * i);
      g[i] = __ldg(reinterpret_cast<const float4*>(dy + r * ld_dy) + lane + 32 * i);
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float* xv = &xh[i].x;
      float* gv = &g[i].x;
      const float* gm = &gam[i].x;
      const float* bt = &bet[i].x;
      float* dg = &dgam[i].x;
      float* db = &dbet[i].x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xv[j] = (xv[j] - mean) * rstd;
        float gg = gv[j];
        if (has_act) {
          const float o = xv[j] * gm[j] + bt[j];
          if (!(o > 0.f)) {
            dsl += (double)gg * (double)o;
            gg *= slope;
          }
        }
        dg[j] += gg * xv[j];
        db[j] += gg;
        gv[j] = gg * gm[j];
        s1 += gv[j];
        s2 += gv[j] * xv[j];
      }
    }
    s1 = warp_sum(s1) * invW;
    s2 = warp_sum(s2) * invW;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o;
      o.x = rstd * (g[i].x - s1 - xh[i].x * s2);
      o.y = rstd * (g[i].y - s1 - xh[i].y * s2);
      o.z = rstd * (g[i].z - s1 - xh[i].z * s2);
      o.w = rstd * (g[i].w - s1 - xh[i].w * s2);
      if (dres) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(dres + r * ld_dres) + lane + 32 * i);
        o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
      }
      dxs[i].x += o.x; dxs[i].y += o.y; dxs[i].z += o.z; dxs[i].w += o.w;      // column sums of the STORED value (incl. dres)
      if (dx) reinterpret_cast<float4*>(dx + r * (long long)W)[lane + 32 * i] = o;
      if (dx_hi) {
        uint2 h, l;
        split4(o, h, l);
        reinterpret_cast<uint2*>(dx_hi + r * ldp)[lane + 32 * i] = h;
        if (dx_lo) reinterpret_cast<uint2*>(dx_lo + r * ldp)[lane + 32 * i] = l;
      }
    }
  }
  // combine the warps of this block in a fixed order
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    reinterpret_cast<float4*>(sm + (warp * 3 + 0) * W)[lane + 32 * i] = dgam[i];
    reinterpret_cast<float4*>(sm + (warp * 3 + 1) * W)[lane + 32 * i] = dbet[i];
    reinterpret_cast<float4*>(sm + (warp * 3 + 2) * W)[lane + 32 * i] = dxs[i];
  }
  dsl = warp_sum(dsl);
  if (lane == 0) red[warp] = dsl;
  __syncthreads();
  float* wsb = ws + (long long)blockIdx.x * (3 * W + 1);
  for (int c = threadIdx.x; c < 3 * W; c += blockDim.x) {
    const int which = c / W, h = c % W;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += sm[(w * 3 + which) * W + h];
    wsb[c] = s;
  }
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w];
    wsb[3 * W] = (float)s;
  }
}

// out_k[c] = sum_b ws[b][k * W + c] in a fixed order (deterministic); k = 0: dgamma, 1: dbeta, 2: dxsum, 3: dslope.
// Block = 32 columns x 8 partial-row lanes: lane ry sums partial blocks ry, ry + 8, ... and the 8 lanes are combined
// through shared memory in order.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ ws, int nblk, int W, float* __restrict__ o0,
                                                              float* __restrict__ o1, float* __restrict__ o2,
                                                              float* __restrict__ o3) {
  __shared__ float sm[8][33];
  const int ncols = 3 * W + 1;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < ncols)
    for (int b = ry; b < nblk; b += 8) s += ws[(long long)b * ncols + c];
  sm[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < ncols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sm[k][cx];
    float* dst = (c < W) ? o0 : (c < 2 * W ? o1 : (c < 3 * W ? o2 : o3));
    if (dst) dst[c < 3 * W ? c % W : 0] = t;
  }
}

// ---------------------------------------------------------------------------------------------- column sums of planes
// stage 1: block (32 column groups of 8) x 8 row lanes; partial sums per row chunk in a fixed order.
__global__ void __launch_bounds__(256) colsum_planes_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                                            long long ld, long long M, int W, long long rows_per_chunk,
                                                            float* __restrict__ dst) {
  __shared__ float sm[8][32][9];
  const int cg = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cg) * 8;
  const long long rbeg = blockIdx.y * rows_per_chunk, rend = min(M, rbeg + rows_per_chunk);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < W) {
    for (long long r = rbeg + ry; r < rend; r += 8) {
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi + r * ld + col));
      const uint32_t* hp = &h.x;
      if (lo) {
        const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo + r * ld + col));
        const uint32_t* lp = &l.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[2 * j] += __uint_as_float(hp[j] << 16) + __uint_as_float(lp[j] << 16);
          acc[2 * j + 1] += __uint_as_float(hp[j] & 0xFFFF0000u) + __uint_as_float(lp[j] & 0xFFFF0000u);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[2 * j] += __uint_as_float(hp[j] << 16);
          acc[2 * j + 1] += __uint_as_float(hp[j] & 0xFFFF0000u);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[ry][cg][j] = acc[j];
  __syncthreads();
  if (ry == 0 && col < W) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += sm[k][cg][j];
      if (col + j < W) dst[(long long)blockIdx.y * W + col + j] = t;
    }
  }
}

__global__ void __launch_bounds__(256) colsum_stage2_kernel(const float* __restrict__ ws, int nchunks, int W, float* __restrict__ out) {
  __shared__ float sm[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < W)
    for (int b = ry; b < nchunks; b += 8) s += ws[(long long)b * W + c];
  sm[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < W) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sm[k][cx];
    out[c] = t;
  }
}

// ---------------------------------------------------------------------------------------------- Linear(H -> 1)
// out_layer (DOSTransformer.py:75,89): one dot product per energy token.  A 128 x 128 GEMM tile would be 127/128 padding;
// these are streaming kernels (one warp per row, 16-byte loads): forward reads x once, backward reads x once and writes dx.
template <int NV>   // K = 128 * NV
__global__ void __launch_bounds__(kWarps * 32) rowdot_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, float* __restrict__ out, long long M) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 wv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) wv[i] = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * i);
  const float b = bias ? __ldg(bias) : 0.f;
  for (long long r = blockIdx.x * (long long)kWarps + warp; r < M; r += (long long)gridDim.x * kWarps) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ldx) + lane + 32 * i);
      s += (v.x * wv[i].x + v.y * wv[i].y) + (v.z * wv[i].z + v.w * wv[i].w);
    }
    s = warp_sum(s);
    if (lane == 0) out[r] = s + b;
  }
}

// dx[m, :] = dout[m] w;  ws[block][0..K) = partial dw = sum_m dout[m] x[m, :],  ws[block][K] = partial db = sum_m dout[m]
template <int NV>
__global__ void __launch_bounds__(kWarps * 32) rowdot_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ x,
                                                                 long long ldx, const float* __restrict__ w, float* __restrict__ dx,
                                                                 float* __restrict__ ws, long long M, long long rows_per_block) {
  constexpr int K = 128 * NV;
  __shared__ __align__(16) float sm[kWarps][K];
  __shared__ float sb[kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 wv[NV], acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    wv[i] = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * i);
    acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float bsum = 0.f;
  const long long rbeg = blockIdx.x * rows_per_block, rend = min(M, rbeg + rows_per_block);
  for (long long r = rbeg + warp; r < rend; r += kWarps) {
    const float g = __ldg(dout + r);
    bsum += g;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ldx) + lane + 32 * i);
      acc[i].x = fmaf(g, v.x, acc[i].x); acc[i].y = fmaf(g, v.y, acc[i].y);
      acc[i].z = fmaf(g, v.z, acc[i].z); acc[i].w = fmaf(g, v.w, acc[i].w);
      if (dx) reinterpret_cast<float4*>(dx + r * (long long)K)[lane + 32 * i] = make_float4(g * wv[i].x, g * wv[i].y, g * wv[i].z, g * wv[i].w);
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) reinterpret_cast<float4*>(sm[warp])[lane + 32 * i] = acc[i];
  if (lane == 0) sb[warp] = bsum;
  __syncthreads();
  float* wsb = ws + (long long)blockIdx.x * (K + 1);
  for (int c = threadIdx.x; c < K; c += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kWarps; ++k) t += sm[k][c];
    wsb[c] = t;
  }
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kWarps; ++k) t += sb[k];
    wsb[K] = t;
  }
}

static inline int bwd_blocks(long long M) {
  long long nb = (M + kWarps - 1) / kWarps;
  if (nb > 2LL * kNumSMs) nb = 2LL * kNumSMs;     // 2 resident blocks per SM; fewer partial rows for the second stage
  if (nb < 1) nb = 1;
  return (int)nb;
}
static inline int colsum_chunks(long long M, int W) {
  const long long gx = (W + 255) / 256;
  long long nch = (4LL * kNumSMs + gx - 1) / gx;
  const long long maxch = (M + 63) / 64;
  if (nch > maxch) nch = maxch;
  if (nch < 1) nch = 1;
  return (int)nch;
}

}  // namespace rbf
}  // namespace dost

using namespace dost;

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int dost_ln_fwd_planes(float* x, long long ldx, const float* ga, const int32_t* ia, const float* gb, const int32_t* ib,
                                  long long ldg, const float* gamma, const float* beta, const float* prelu_slope, float* y,
                                  void* hi, void* lo, long long ldp, float* stats, long long M, int W, dost_stream_t stream) {
  if (M == 0) return DOST_OK;
  DOST_REQUIRE(x && gamma && beta && (y || hi) && M > 0, "ln_fwd_planes: bad args");
  DOST_REQUIRE(W % 128 == 0 && W <= 1024 && (W == 128 || W == 256 || W == 512 || W == 1024),
               "ln_fwd_planes: W must be 128, 256, 512 or 1024 (got %d)", W);
  DOST_REQUIRE(al16(x) && ldx % 4 == 0 && al16(gamma) && al16(beta) && al16(y) && al16(hi) && al16(lo) && ldp % 4 == 0,
               "ln_fwd_planes: operands must be 16-byte aligned");
  DOST_REQUIRE(!ga || (gb && ia && ib && al16(ga) && al16(gb) && ldg % 4 == 0), "ln_fwd_planes: bad gather-add operands");
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)min64((M + rbf::kWarps - 1) / rbf::kWarps, 16LL * kNumSMs);
#define DOST_LNF(NV)                                                                                                     \
  rbf::ln_fwd_kernel<NV><<<blocks, rbf::kWarps * 32, 0, st>>>(x, ldx, ga, ia, gb, ib, ldg, gamma, beta, prelu_slope, y, (__nv_bfloat16*)hi, \
                                                             (__nv_bfloat16*)lo, ldp, stats, M)
  switch (W) {
    case 128: DOST_LNF(1); break;
    case 256: DOST_LNF(2); break;
    case 512: DOST_LNF(4); break;
    default: DOST_LNF(8); break;
  }
#undef DOST_LNF
  return check_launch("ln_fwd_planes");
}

extern "C" size_t dost_ln_bwd_planes_workspace_bytes(long long M, int W) {
  return sizeof(float) * (size_t)rbf::bwd_blocks(M) * (3 * W + 1);
}

extern "C" int dost_ln_bwd_planes(const float* dy, long long ld_dy, const float* x, long long ldx, const float* stats,
                                  const float* gamma, const float* beta, const float* prelu_slope, const float* dres,
                                  long long ld_dres, float* dx, void* dx_hi, void* dx_lo, long long ldp, float* dgamma,
                                  float* dbeta, float* dslope, float* dxsum, long long M, int W, void* workspace,
                                  size_t workspace_bytes, dost_stream_t stream) {
  DOST_REQUIRE(dy && x && stats && gamma && beta && (dx || dx_hi) && M > 0, "ln_bwd_planes: bad args");
  DOST_REQUIRE(W == 128 || W == 256 || W == 512 || W == 1024, "ln_bwd_planes: W must be 128, 256, 512 or 1024 (got %d)", W);
  DOST_REQUIRE(al16(dy) && ld_dy % 4 == 0 && al16(x) && ldx % 4 == 0 && al16(gamma) && al16(beta) && al16(dres) &&
                   ld_dres % 4 == 0 && al16(dx) && al16(dx_hi) && al16(dx_lo) && ldp % 4 == 0,
               "ln_bwd_planes: operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = rbf::bwd_blocks(M);
  const size_t need = sizeof(float) * (size_t)blocks * (3 * W + 1);
  if (!workspace || workspace_bytes < need) {
    set_error("ln_bwd_planes: workspace too small (%zu < %zu)", workspace_bytes, need);
    return DOST_ERR_WORKSPACE;
  }
  const long long rpb = (M + blocks - 1) / blocks;
  const size_t smem = sizeof(float) * (size_t)rbf::kWarps * 3 * W;
#define DOST_LNB(NV)                                                                                                        \
  {                                                                                                                         \
    if (smem + 2048 > 48 * 1024) cudaFuncSetAttribute(rbf::ln_bwd_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    rbf::ln_bwd_kernel<NV><<<blocks, rbf::kWarps * 32, smem, st>>>(dy, ld_dy, x, ldx, stats, gamma, beta, prelu_slope, dres,  \
                                                                  ld_dres, dx, (__nv_bfloat16*)dx_hi, (__nv_bfloat16*)dx_lo, \
                                                                  ldp, (float*)workspace, M, rpb);                          \
  }
  switch (W) {
    case 128: DOST_LNB(1) break;
    case 256: DOST_LNB(2) break;
    case 512: DOST_LNB(4) break;
    default: DOST_LNB(8) break;
  }
#undef DOST_LNB
  int rc = check_launch("ln_bwd_planes");
  if (rc != DOST_OK) return rc;
  const int ncols = 3 * W + 1;
  rbf::reduce_partials_kernel<<<ceil_div(ncols, 32), 256, 0, st>>>((const float*)workspace, blocks, W, dgamma, dbeta, dxsum, dslope);
  return check_launch("ln_bwd_planes reduce");
}

extern "C" size_t dost_colsum_planes_workspace_bytes(long long M, int W) {
  return sizeof(float) * (size_t)rbf::colsum_chunks(M, W) * W;
}

extern "C" int dost_colsum_planes(const void* hi, const void* lo, long long ld, long long M, int W, float* out, void* workspace,
                                  size_t workspace_bytes, dost_stream_t stream) {
  DOST_REQUIRE(hi && out && M > 0 && W > 0 && ld % 8 == 0 && al16(hi) && al16(lo), "colsum_planes: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const int nch = rbf::colsum_chunks(M, W);
  const size_t need = sizeof(float) * (size_t)nch * W;
  if (!workspace || workspace_bytes < need) {
    set_error("colsum_planes: workspace too small (%zu < %zu)", workspace_bytes, need);
    return DOST_ERR_WORKSPACE;
  }
  const long long rpc = (M + nch - 1) / nch;
  dim3 grid((unsigned)((W + 255) / 256), nch);
  rbf::colsum_planes_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)hi, (const __nv_bfloat16*)lo, ld, M, W, rpc,
                                                  (float*)workspace);
  int rc = check_launch("colsum_planes stage1");
  if (rc != DOST_OK) return rc;
  rbf::colsum_stage2_kernel<<<ceil_div(W, 32), 256, 0, st>>>((const float*)workspace, nch, W, out);
  return check_launch("colsum_planes stage2");
}

extern "C" int dost_rowdot_fwd(const float* x, long long ldx, const float* w, const float* bias, float* out, long long M, int K,
                               dost_stream_t stream) {
  DOST_REQUIRE(x && w && out && M > 0 && (K == 128 || K == 256 || K == 512) && al16(x) && al16(w) && ldx % 4 == 0, "rowdot_fwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)min64((M + rbf::kWarps - 1) / rbf::kWarps, 16LL * kNumSMs);
  if (K == 128) rbf::rowdot_fwd_kernel<1><<<blocks, rbf::kWarps * 32, 0, st>>>(x, ldx, w, bias, out, M);
  else if (K == 256) rbf::rowdot_fwd_kernel<2><<<blocks, rbf::kWarps * 32, 0, st>>>(x, ldx, w, bias, out, M);
  else rbf::rowdot_fwd_kernel<4><<<blocks, rbf::kWarps * 32, 0, st>>>(x, ldx, w, bias, out, M);
  return check_launch("rowdot_fwd");
}

extern "C" size_t dost_rowdot_bwd_workspace_bytes(long long M, int K) { return sizeof(float) * (size_t)rbf::bwd_blocks(M) * (K + 1); }

extern "C" int dost_rowdot_bwd(const float* dout, const float* x, long long ldx, const float* w, float* dx, float* dwb, long long M,
                               int K, void* workspace, size_t workspace_bytes, dost_stream_t stream) {
  DOST_REQUIRE(dout && x && w && dwb && M > 0 && (K == 128 || K == 256 || K == 512) && al16(x) && al16(w) && al16(dx) && ldx % 4 == 0,
               "rowdot_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = rbf::bwd_blocks(M);
  const size_t need = sizeof(float) * (size_t)blocks * (K + 1);
  if (!workspace || workspace_bytes < need) {
    set_error("rowdot_bwd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return DOST_ERR_WORKSPACE;
  }
  const long long rpb = (M + blocks - 1) / blocks;
  float* ws = (float*)workspace;
  if (K == 128) rbf::rowdot_bwd_kernel<1><<<blocks, rbf::kWarps * 32, 0, st>>>(dout, x, ldx, w, dx, ws, M, rpb);
  else if (K == 256) rbf::rowdot_bwd_kernel<2><<<blocks, rbf::kWarps * 32, 0, st>>>(dout, x, ldx, w, dx, ws, M, rpb);
  else rbf::rowdot_bwd_kernel<4><<<blocks, rbf::kWarps * 32, 0, st>>>(dout, x, ldx, w, dx, ws, M, rpb);
  int rc = check_launch("rowdot_bwd");
  if (rc != DOST_OK) return rc;
  // partial rows are K + 1 wide: the fixed-order column reduction gives dw (columns 0..K-1) and db (column K)
  rbf::colsum_stage2_kernel<<<ceil_div(K + 1, 32), 256, 0, st>>>(ws, blocks, K + 1, dwb);
  return check_launch("rowdot_bwd reduce");
}
