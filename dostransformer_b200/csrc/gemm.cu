// FMA-pipe GEMM family with gather/concat prologue and fused epilogue (fp32 / fp64).
//
//   C[m,n] = epi( sum_k A(m,k) * B(n,k) )
//
// A and B are described by access modes rather than transposes so the same kernel serves the forward
// Linear (A: activation rows, possibly gathered/concatenated; B: weight [N,K]), the input gradient
// (B: weight read N-contiguous), the weight gradient (both operands read along the row index, reduction
// over rows, optional row gather on the reduction index, deterministic split-K) and the batched
// attention contractions.  See include/dost.h for the reference call sites this replaces.
//
// Tile: (16*TM) x (16*TN) x 16 with TM = TN = 2 * (16 bytes / sizeof(T)); 256 threads; each thread owns a
// 2x2 arrangement of VxV sub-tiles (bank-conflict-free vector reads of the shared tiles); global loads are
// register-staged one k-tile ahead of the math (double-buffered shared memory, one barrier per k-tile).
#include "gemm_common.cuh"

namespace dost {

template <typename T> __device__ __forceinline__ typename VecOf<T>::type vzero();
template <> __device__ __forceinline__ float4 vzero<float>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <> __device__ __forceinline__ double2 vzero<double>() { return make_double2(0.0, 0.0); }

__device__ __forceinline__ void unpack(const float4& v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ void unpack(const double2& v, double* o) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ float4 pack(const float* o) { return make_float4(o[0], o[1], o[2], o[3]); }
__device__ __forceinline__ double2 pack(const double* o) { return make_double2(o[0], o[1]); }

// Load up to V consecutive elements starting at p; elements >= nvalid are zero.
template <typename T>
__device__ __forceinline__ typename VecOf<T>::type load_vec(const T* p, int nvalid, bool vec_ok) {
  constexpr int V = VecOf<T>::N;
  using Vec = typename VecOf<T>::type;
  if (nvalid >= V && vec_ok) return __ldg(reinterpret_cast<const Vec*>(p));
  T tmp[V];
#pragma unroll
  for (int j = 0; j < V; ++j) tmp[j] = (j < nvalid) ? __ldg(p + j) : T(0);
  return pack(tmp);
}

template <typename T, bool A_MC, bool B_MC>
__global__ void __launch_bounds__(256, 2) gemm_kernel(const GemmDev<T> g) {
  constexpr int V = VecOf<T>::N;
  using Vec = typename VecOf<T>::type;
  constexpr int TM = 2 * V, TN = 2 * V;
  constexpr int BM = 16 * TM, BN = 16 * TN, BK = 16;
  constexpr int MQ = BM / V;  // vectors per smem row (32 for both dtypes)
  constexpr int NQ = BN / V;

  __shared__ __align__(16) T As[2][BK][BM];
  __shared__ __align__(16) T Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  int kbeg = 0, kend = g.K;
  long long aoff = 0, boff = 0, coff = 0;
  if (g.zmode == 1) {
    aoff = (long long)blockIdx.z * g.a_bstride;
    boff = (long long)blockIdx.z * g.b_bstride;
    coff = (long long)blockIdx.z * g.c_bstride;
  } else if (g.zmode == 2) {
    kbeg = blockIdx.z * g.kchunk;
    kend = min(g.K, kbeg + g.kchunk);
  }
  const int nkt = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;

  // ---- per-thread load bookkeeping (2 vector items per operand per k-tile)
  const T* arow[2] = {nullptr, nullptr};
  const T* brow[2] = {nullptr, nullptr};
  int a_seg_cur = -1;
  if (!B_MC) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int i = tid + it * 256;
      const int n = n0 + (i % BN);
      brow[it] = (n < g.N) ? g.b.base + boff + (long long)n * g.b.ld : nullptr;
    }
  }

  Vec ra[2], rb[2];

  auto load_tiles = [&](int kt) {
    const int k0 = kbeg + kt * BK;
    // ---------------- A
    if (A_MC) {
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int i = tid + it * 256;
        const int mq = i % MQ, k = i / MQ;
        const int kk = k0 + k, m = m0 + mq * V;
        if (kk < kend && m < g.M) {
          const T* p = g.a[0].base + aoff + (long long)kk * g.a[0].ld + m;
          ra[it] = load_vec<T>(p, g.M - m, g.a[0].vec_ok);
        } else {
          ra[it] = vzero<T>();
        }
      }
    } else {
      int s = 0;
      while (s + 1 < g.a_nseg && k0 >= g.a[s].kend) ++s;
      if (s != a_seg_cur) {
        a_seg_cur = s;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int i = tid + it * 256;
          const int m = m0 + (i % BM);
          if (m < g.M) {
            long long r = m / g.a[s].div;
            if (g.a[s].idx) r = __ldg(g.a[s].idx + r);
            arow[it] = g.a[s].base + aoff + r * g.a[s].ld;
          } else {
            arow[it] = nullptr;
          }
        }
      }
      const int kstart = (s == 0) ? 0 : g.a[s - 1].kend;
      const int klim = min(kend, g.a[s].kend);
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int i = tid + it * 256;
        const int kk = k0 + (i / BM) * V;
        const int nvalid = klim - kk;
        if (arow[it] && nvalid > 0) ra[it] = load_vec<T>(arow[it] + (kk - kstart), nvalid, g.a[s].vec_ok);
        else ra[it] = vzero<T>();
      }
    }
    // ---------------- B
    if (B_MC) {
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int i = tid + it * 256;
        const int nq = i % NQ, k = i / NQ;
        const int kk = k0 + k, n = n0 + nq * V;
        if (kk < kend && n < g.N) {
          long long r = kk / g.b.div;
          if (g.b.idx) r = __ldg(g.b.idx + r);
          const T* p = g.b.base + boff + r * g.b.ld + n;
          rb[it] = load_vec<T>(p, g.N - n, g.b.vec_ok);
        } else {
          rb[it] = vzero<T>();
        }
      }
    } else {
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int i = tid + it * 256;
        const int kk = k0 + (i / BN) * V;
        const int nvalid = kend - kk;
        if (brow[it] && nvalid > 0) rb[it] = load_vec<T>(brow[it] + kk, nvalid, g.b.vec_ok);
        else rb[it] = vzero<T>();
      }
    }
  };

  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int i = tid + it * 256;
      if (A_MC) {
        const int mq = i % MQ, k = i / MQ;
        *reinterpret_cast<Vec*>(&As[buf][k][mq * V]) = ra[it];
      } else {
        const int row = i % BM, kq = i / BM;
        T tmp[V];
        unpack(ra[it], tmp);
#pragma unroll
        for (int j = 0; j < V; ++j) As[buf][kq * V + j][row] = tmp[j];
      }
      if (B_MC) {
        const int nq = i % NQ, k = i / NQ;
        *reinterpret_cast<Vec*>(&Bs[buf][k][nq * V]) = rb[it];
      } else {
        const int row = i % BN, kq = i / BN;
        T tmp[V];
        unpack(rb[it], tmp);
#pragma unroll
        for (int j = 0; j < V; ++j) Bs[buf][kq * V + j][row] = tmp[j];
      }
    }
  };

  T acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = T(0);

  if (nkt > 0) {
    load_tiles(0);
    store_tiles(0);
  }
  __syncthreads();

  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nkt) load_tiles(kt + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      T a[TM], b[TN];
      unpack(*reinterpret_cast<const Vec*>(&As[buf][kk][ty * V]), a);
      unpack(*reinterpret_cast<const Vec*>(&As[buf][kk][BM / 2 + ty * V]), a + V);
      unpack(*reinterpret_cast<const Vec*>(&Bs[buf][kk][tx * V]), b);
      unpack(*reinterpret_cast<const Vec*>(&Bs[buf][kk][BN / 2 + tx * V]), b + V);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nkt) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---------------- epilogue
  if (g.zmode == 2) {
    T* ws = g.ws + (long long)blockIdx.z * g.M * g.N;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ((i < V) ? ty * V + i : BM / 2 + ty * V + (i - V));
      if (m >= g.M) continue;
#pragma unroll
      for (int jg = 0; jg < 2; ++jg) {
        const int n = n0 + jg * (BN / 2) + tx * V;
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (n + j < g.N) ws[(long long)m * g.N + n + j] = acc[i][jg * V + j];
      }
    }
    return;
  }

  T pslope = T(0);
  if (g.act == DOST_ACT_PRELU) pslope = __ldg(g.prelu_slope);
  else if (g.act == DOST_ACT_LEAKY) pslope = g.act_slope;

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ((i < V) ? ty * V + i : BM / 2 + ty * V + (i - V));
    if (m >= g.M) continue;
#pragma unroll
    for (int jg = 0; jg < 2; ++jg) {
      const int n = n0 + jg * (BN / 2) + tx * V;
      if (n >= g.N) continue;
      T v[V];
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = acc[i][jg * V + j];
      const bool full = g.epi_vec && (n + V <= g.N);
      if (full) {
        if (g.bias) {
          T t[V];
          unpack(__ldg(reinterpret_cast<const Vec*>(g.bias + n)), t);
#pragma unroll
          for (int j = 0; j < V; ++j) v[j] += t[j];
        }
        if (g.out_pre) *reinterpret_cast<Vec*>(g.out_pre + coff + (long long)m * g.ld_pre + n) = pack(v);
        if (g.act != DOST_ACT_NONE) {
#pragma unroll
          for (int j = 0; j < V; ++j) v[j] = (v[j] > T(0)) ? v[j] : pslope * v[j];
        }
        if (g.dact_saved) {
          T t[V];
          unpack(__ldg(reinterpret_cast<const Vec*>(g.dact_saved + coff + (long long)m * g.ld_dact + n)), t);
#pragma unroll
          for (int j = 0; j < V; ++j) v[j] *= (t[j] > T(0)) ? T(1) : g.dact_slope;
        }
        if (g.residual) {
          T t[V];
          unpack(__ldg(reinterpret_cast<const Vec*>(g.residual + coff + (long long)m * g.ld_res + n)), t);
#pragma unroll
          for (int j = 0; j < V; ++j) v[j] += t[j];
        }
        Vec* op = reinterpret_cast<Vec*>(g.out + coff + (long long)m * g.ldc + n);
        if (g.accumulate) {
          T t[V];
          unpack(*op, t);
#pragma unroll
          for (int j = 0; j < V; ++j) v[j] += t[j];
        }
        *op = pack(v);
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          if (n + j >= g.N) continue;
          T x = v[j];
          if (g.bias) x += __ldg(g.bias + n + j);
          if (g.out_pre) g.out_pre[coff + (long long)m * g.ld_pre + n + j] = x;
          if (g.act != DOST_ACT_NONE) x = (x > T(0)) ? x : pslope * x;
          if (g.dact_saved) x *= (__ldg(g.dact_saved + coff + (long long)m * g.ld_dact + n + j) > T(0)) ? T(1) : g.dact_slope;
          if (g.residual) x += __ldg(g.residual + coff + (long long)m * g.ld_res + n + j);
          T* op = g.out + coff + (long long)m * g.ldc + n + j;
          if (g.accumulate) x += *op;
          *op = x;
        }
      }
    }
  }
}

// Fixed-order reduction of split-K partials: out[m,n] (+)= sum_z ws[z][m][n].
template <typename T>
__global__ void splitk_reduce_kernel(const T* __restrict__ ws, T* __restrict__ out, long long ldc, int M, int N,
                                     int splits, int accumulate) {
  const long long total = (long long)M * N;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    T s = T(0);
    for (int z = 0; z < splits; ++z) s += ws[(long long)z * total + i];
    const long long m = i / N, n = i % N;
    T* op = out + m * ldc + n;
    *op = accumulate ? (*op + s) : s;
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// precision: 0 = FMA pipe (fp32/fp64), 1 = tcgen05 bf16x3 (fp32 parity), 2 = tcgen05 bf16.
static inline bool use_tensor_cores(const GemmDev<float>& g, int precision, bool amc, bool bmc) {
  return precision > 0 && gemm_tc_supported(g, amc, bmc);
}
static inline bool use_tensor_cores(const GemmDev<double>&, int, bool, bool) { return false; }
static inline int launch_tc(const GemmDev<float>& g, int precision, bool amc, bool bmc, int batch, int split, cudaStream_t st) {
  return launch_gemm_tc(g, precision, amc, bmc, batch, split, st);
}
static inline int launch_tc(const GemmDev<double>&, int, bool, bool, int, int, cudaStream_t) { return DOST_ERR_UNSUPPORTED; }

template <typename T>
static int run_gemm(const dost_gemm_t* h, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  constexpr int V = VecOf<T>::N;
  constexpr int BM = 32 * V, BN = 32 * V;
  GemmDev<T> g;
  g.M = h->M; g.N = h->N; g.K = h->K;
  const int batch = h->batch < 1 ? 1 : h->batch;
  const int split = h->split_k < 1 ? 1 : h->split_k;
  DOST_REQUIRE(h->M > 0 && h->N > 0 && h->K >= 0, "gemm: bad shape M=%d N=%d K=%d", h->M, h->N, h->K);
  DOST_REQUIRE(!(batch > 1 && split > 1), "gemm: batch and split_k are exclusive");
  DOST_REQUIRE(h->out != nullptr, "gemm: out is null");
  DOST_REQUIRE(h->precision >= 0 && h->precision <= 2, "gemm: precision must be 0 (FMA), 1 (bf16x3) or 2 (bf16)");
  g.a_nseg = (h->a_mode == DOST_MC) ? 1 : h->a_nseg;
  DOST_REQUIRE(g.a_nseg >= 1 && g.a_nseg <= 3, "gemm: a_nseg must be 1..3");
  int kacc = 0;
  for (int s = 0; s < g.a_nseg; ++s) {
    const dost_seg_t& sg = h->a[s];
    DOST_REQUIRE(sg.base != nullptr, "gemm: A segment %d base is null", s);
    const int width = (h->a_mode == DOST_MC) ? h->K : sg.width;
    DOST_REQUIRE(width > 0, "gemm: A segment %d has width %d", s, width);
    if (g.a_nseg > 1) DOST_REQUIRE(width % 16 == 0, "gemm: concatenated A segments need widths %% 16 == 0 (got %d)", width);
    kacc += width;
    g.a[s].base = (const T*)sg.base;
    g.a[s].ld = sg.ld;
    g.a[s].idx = sg.idx;
    g.a[s].div = sg.div < 1 ? 1 : sg.div;
    g.a[s].kend = kacc;
    bool ok = aligned16(sg.base) && (sg.ld % V == 0) && (h->a_bstride % V == 0);
    if (h->a_mode == DOST_KC) ok = ok && true;  // k offsets inside a segment are multiples of V by construction
    g.a[s].vec_ok = ok ? 1 : 0;
    DOST_REQUIRE(batch == 1 || (sg.idx == nullptr && g.a[s].div == 1), "gemm: row maps not allowed when batched");
  }
  for (int s = g.a_nseg; s < 3; ++s) g.a[s] = g.a[0];
  DOST_REQUIRE(kacc == h->K, "gemm: A segment widths sum to %d, K=%d", kacc, h->K);
  DOST_REQUIRE(h->b.base != nullptr, "gemm: B base is null");
  g.b.base = (const T*)h->b.base;
  g.b.ld = h->b.ld;
  g.b.idx = h->b.idx;
  g.b.div = h->b.div < 1 ? 1 : h->b.div;
  g.b.kend = h->K;
  g.b.vec_ok = (aligned16(h->b.base) && (h->b.ld % V == 0) && (h->b_bstride % V == 0)) ? 1 : 0;
  DOST_REQUIRE(h->b_mode == DOST_MC || (h->b.idx == nullptr && g.b.div == 1), "gemm: B row map needs MC mode");
  g.a_bstride = h->a_bstride; g.b_bstride = h->b_bstride; g.c_bstride = h->c_bstride;
  g.bias = (const T*)h->bias;
  g.act = h->act;
  g.act_slope = (h->act == DOST_ACT_RELU) ? T(0) : (T)h->act_slope;
  g.prelu_slope = (const T*)h->prelu_slope;
  DOST_REQUIRE(h->act != DOST_ACT_PRELU || h->prelu_slope, "gemm: PReLU needs a slope pointer");
  g.out_pre = (T*)h->out_pre; g.ld_pre = h->ld_pre;
  g.dact_saved = (const T*)h->dact_saved; g.ld_dact = h->ld_dact;
  g.dact_slope = (h->dact_kind == DOST_ACT_RELU) ? T(0) : (T)h->dact_slope;
  g.residual = (const T*)h->residual; g.ld_res = h->ld_res;
  g.out = (T*)h->out; g.ldc = h->ldc;
  g.accumulate = h->accumulate;
  bool ev = aligned16(h->out) && (h->ldc % V == 0) && (h->c_bstride % V == 0);
  if (h->bias) ev = ev && aligned16(h->bias);
  if (h->out_pre) ev = ev && aligned16(h->out_pre) && (h->ld_pre % V == 0);
  if (h->dact_saved) ev = ev && aligned16(h->dact_saved) && (h->ld_dact % V == 0);
  if (h->residual) ev = ev && aligned16(h->residual) && (h->ld_res % V == 0);
  g.epi_vec = ev ? 1 : 0;
  g.zmode = batch > 1 ? 1 : (split > 1 ? 2 : 0);
  g.kchunk = 0;
  g.ws = nullptr;
  int gz = batch;
  if (split > 1) {
    DOST_REQUIRE(!h->bias && h->act == DOST_ACT_NONE && !h->out_pre && !h->dact_saved && !h->residual,
                 "gemm: split_k supports only plain (accumulating) stores");
    const size_t need = sizeof(T) * (size_t)split * h->M * h->N;
    if (!workspace || workspace_bytes < need) {
      set_error("gemm: split_k workspace too small (%zu < %zu)", workspace_bytes, need);
      return DOST_ERR_WORKSPACE;
    }
    int kchunk = (h->K + split - 1) / split;
    kchunk = ((kchunk + 15) / 16) * 16;
    g.kchunk = kchunk;
    g.ws = (T*)workspace;
    gz = split;   // slices past K are empty: they contribute exact zeros to the fixed-order reduction
  }
  const bool amc = h->a_mode == DOST_MC, bmc = h->b_mode == DOST_MC;
  int rc;
  if (use_tensor_cores(g, h->precision, amc, bmc)) {
    rc = launch_tc(g, h->precision, amc, bmc, batch, split, st);
  } else {
    dim3 grid(ceil_div(h->M, BM), ceil_div(h->N, BN), gz);
    DOST_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm: grid too large");
    if (amc && bmc) gemm_kernel<T, true, true><<<grid, 256, 0, st>>>(g);
    else if (amc) gemm_kernel<T, true, false><<<grid, 256, 0, st>>>(g);
    else if (bmc) gemm_kernel<T, false, true><<<grid, 256, 0, st>>>(g);
    else gemm_kernel<T, false, false><<<grid, 256, 0, st>>>(g);
    rc = check_launch("gemm");
  }
  if (rc != DOST_OK) return rc;
  if (split > 1) {
    const long long total = (long long)h->M * h->N;
    int blocks = min(ceil_div(total, 256), kNumSMs * 8);
    splitk_reduce_kernel<T><<<blocks, 256, 0, st>>>(g.ws, g.out, g.ldc, h->M, h->N, split, h->accumulate);
    rc = check_launch("gemm split-k reduce");
  }
  return rc;
}

}  // namespace dost

extern "C" size_t dost_gemm_workspace_bytes(const dost_gemm_t* g) {
  if (!g || g->split_k <= 1) return 0;
  const size_t es = g->dtype == DOST_F64 ? 8 : 4;
  return es * (size_t)g->split_k * g->M * g->N;
}

extern "C" int dost_gemm(const dost_gemm_t* g, void* workspace, size_t workspace_bytes, dost_stream_t stream) {
  DOST_REQUIRE(g != nullptr, "gemm: null descriptor");
  if (g->dtype == DOST_F32) return dost::run_gemm<float>(g, workspace, workspace_bytes, (cudaStream_t)stream);
  if (g->dtype == DOST_F64) return dost::run_gemm<double>(g, workspace, workspace_bytes, (cudaStream_t)stream);
  dost::set_error("gemm: unsupported dtype %d", g->dtype);
  return DOST_ERR_UNSUPPORTED;
}
