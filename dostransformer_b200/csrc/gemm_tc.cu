// tcgen05 (5th-gen tensor core) GEMM for sm_100a with the same gather/concat prologue and fused epilogue as the
// FMA-pipe kernel in gemm.cu.
//
//   C[m,n] = epi( sum_k A(m,k) * B(n,k) ),  A/B fp32 in HBM, bf16 operands in shared memory, fp32 accumulate in TMEM.
//
// Precision modes
//   bf16x3 (NSPLIT = 3): x = hi + lo with hi = bf16(x), lo = bf16(x - hi); the product is accumulated as
//                        hi*hi + hi*lo + lo*hi (three tcgen05.mma per k-step into the same TMEM accumulator).
//                        Relative error per product ~2^-16.
//   bf16   (NSPLIT = 1): operands rounded to bf16 once.
//
// Structure (one CTA per SM, persistent over output tiles, 416 threads):
//   warps 0-3   epilogue : tcgen05.ld the 128 x BN fp32 accumulator (one TMEM lane = one output row per thread),
//                          bias / activation / residual / ... , direct fp32 stores
//   warp  4     MMA      : one elected thread issues tcgen05.mma (M=128, N=BN, K=16), tcgen05.commit -> mbarriers
//   warps 5-12  producers: ld.global fp32 (gathered / concatenated rows) -> cvt to bf16 hi(/lo) -> st.shared in the
//                          UMMA canonical SWIZZLE_128B layout (K-major when the reduction index is contiguous in
//                          HBM, MN-major otherwise: no transposition in either case) -> fence.proxy.async -> mbarrier
// Pipelines: smem full/empty ring (producers <-> MMA), double-buffered TMEM accumulator full/empty (MMA <-> epilogue).
// The operands of this kernel are fp32 in HBM, so TMA (which cannot convert) is not used; the conversion is what the
// producer warps are for, and it is what bounds the kernel (L1TEX wavefronts of the LDG/STS traffic, ~0.2 PFLOP/s).
// Every large GEMM of the model therefore runs on gemm_bf.cu (operands pre-split into bf16 planes by their producers,
// pure TMA -> tcgen05); this kernel remains for operands with row maps (gathers, broadcasts) and odd widths.
#include <cuda_bf16.h>
#include "gemm_common.cuh"
#include "tc_ptx.cuh"

namespace dost {
namespace tc {

constexpr int BM = 128, BK = 64;
constexpr int kEpiWarps = 4, kProdWarps = 8;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;  // 416
constexpr int kMaxStages = 4;
constexpr int kSmemBudget = 200 * 1024;

struct Sched {
  int m_tiles, n_tiles, z_count, total_tiles;
};

using namespace ptx;      // tcgen05 / mbarrier wrappers and the UMMA descriptors (tc_ptx.cuh)

// MN-major operands here are 64-row blocks (BK = 64): LBO = 8192 B.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, bool mn_major) { return make_smem_desc(saddr, mn_major ? 8192u : 0u); }
__device__ __forceinline__ uint32_t make_idesc(int n, bool a_mn, bool b_mn) { return ptx::make_idesc(BM, n, a_mn, b_mn); }
__device__ __forceinline__ void fence_proxy_async() { fence_async_smem(); }

// 4 fp32 -> 4 bf16 (hi, 8 bytes) and, if SPLIT, the 4 bf16 residuals (lo) of the error-compensated split.
template <bool SPLIT>
__device__ __forceinline__ void convert4(const float4& x, uint2& hi, uint2& lo) {
  hi.x = pack_bf16(x.x, x.y);
  hi.y = pack_bf16(x.z, x.w);
  if (SPLIT) {
    lo.x = pack_bf16(x.x - __uint_as_float(hi.x << 16), x.y - __uint_as_float(hi.x & 0xFFFF0000u));
    lo.y = pack_bf16(x.z - __uint_as_float(hi.y << 16), x.w - __uint_as_float(hi.y & 0xFFFF0000u));
  }
}

__device__ __forceinline__ void sts64(uint32_t addr, uint2 v) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}

// Producer-side view of the CTA's flattened sequence of (output tile, k-tile) work items.
//
// Per k-tile every producer thread moves NAI + NBI items; an item is 4 consecutive fp32 in HBM (one LDG.128, lanes of a
// warp contiguous: a warp-wide load covers 512 contiguous bytes = 4 full lines) = one 8-byte half of a 16-byte bf16
// chunk of the SWIZZLE_128B operand tile in shared memory (STS.64, conflict-free per half-warp).
//   K-major  operand (reduction index contiguous in HBM): item i -> row i / 16, k = 4 * (i % 16)
//   MN-major operand (row index contiguous in HBM)      : item i -> k = i / (ROWS / 4), rows 4 * (i % (ROWS / 4))
// with i = pt + 256 * j, j = 0 .. items-1.  Everything that does not change from one k-tile to the next (tile
// origin, gathered row numbers, segment of a concatenated A operand, shared-memory offsets) is resolved once per
// tile / segment / thread; the per-item work on the fast path is one predicate, one IMAD.WIDE, one LDG.128, the
// conversion and two STS.64.
template <int BN, bool A_MC, bool B_MC>
struct Producer {
  static constexpr int NAI = BM * BK / 4 / kProdThreads;   // 8
  static constexpr int NBI = BN * BK / 4 / kProdThreads;   // 16 / 8 / 4
  static constexpr int NI = NAI + NBI;
  static constexpr int BW4 = BN / 4, BKS = kProdThreads / BW4;   // MN-major B: float4 per k-row, k-rows per item step
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;

  const GemmDev<float>& g;
  const Sched& sch;
  const int pt;
  // thread constants: element offset along the contiguous HBM axis (c) and index along the other axis (r) of item 0
  int a_c, a_r, b_c, b_r;
  uint32_t a_soff, b_soff;      // shared-memory byte offset of item 0 inside the operand tile
  // position
  int tile, kt, nkt, m0, n0, kbeg, kend, k0;
  // A operand
  int seg, kstart, klim;
  int arow[A_MC ? 1 : NAI];
  const float* abase;
  int ald;
  bool a_fast, a_pred;
  const float* pa;
  // B operand
  const float* bbase;
  int bld;
  bool b_fast, b_pred, b_gather;
  const float* pb;

  __device__ __forceinline__ Producer(const GemmDev<float>& g_, const Sched& s_, int pt_) : g(g_), sch(s_), pt(pt_) {
    if (!A_MC) {
      a_c = (pt & 15) * 4;
      a_r = pt >> 4;
      a_soff = (a_r >> 3) * 1024 + (a_r & 7) * 128 + (((((pt & 15) >> 1) ^ (a_r & 7))) << 4) + (pt & 1) * 8;
    } else {
      a_c = pt >> 5;              // k row inside the k-tile (+ 8 j)
      a_r = (pt & 31) * 4;        // first of the 4 consecutive rows
      const int rc = (pt & 31) >> 1;
      a_soff = (rc >> 3) * 8192 + a_c * 128 + (((rc & 7) ^ a_c) << 4) + (pt & 1) * 8;
    }
    if (!B_MC) {
      b_c = (pt & 15) * 4;
      b_r = pt >> 4;
      b_soff = (b_r >> 3) * 1024 + (b_r & 7) * 128 + (((((pt & 15) >> 1) ^ (b_r & 7))) << 4) + (pt & 1) * 8;
    } else {
      b_c = pt / BW4;
      b_r = (pt % BW4) * 4;
      const int rc = (pt % BW4) >> 1;
      b_soff = (rc >> 3) * 8192 + (b_c >> 3) * 1024 + (b_c & 7) * 128 + (((rc & 7) ^ (b_c & 7)) << 4) + (pt & 1) * 8;
    }
    bld = (int)g.b.ld;
    b_gather = B_MC && (g.b.idx != nullptr || g.b.div != 1);
  }

  __device__ __forceinline__ bool valid() const { return tile < sch.total_tiles; }

  // shared-memory byte offset of item j relative to item 0 (compile-time for unrolled j), applied to `off0`
  __device__ __forceinline__ static uint32_t a_item_off(uint32_t off0, int j) { return off0 + j * (A_MC ? 1024 : 2048); }
  __device__ __forceinline__ static uint32_t b_item_off(uint32_t off0, int j) {
    if (!B_MC) return off0 + j * 2048;
    if (BKS == 8) return off0 + j * 1024;
    if (BKS == 16) return off0 + j * 2048;
    // BKS == 4 (BN = 256): k = kw + 4 j toggles bit 2 of (k & 7) with the parity of j
    return (off0 ^ ((j & 1) * 0x40)) + (j & 1) * 512 + (j >> 1) * 1024;
  }

  __device__ __forceinline__ void enter_segment() {     // K-major A: resolve the rows of this thread's items
    kstart = (seg == 0) ? 0 : g.a[seg - 1].kend;
    klim = min(kend, g.a[seg].kend);
    abase = g.a[seg].base + (g.zmode == 1 ? (long long)(tile / (sch.m_tiles * sch.n_tiles)) * g.a_bstride : 0);
    ald = (int)g.a[seg].ld;
    if (!A_MC) {
      const int adiv = g.a[seg].div;
      const int* aidx = g.a[seg].idx;
#pragma unroll
      for (int j = 0; j < NAI; ++j) {
        const int m = m0 + a_r + 16 * j;
        int row = -1;
        if (m < g.M) {
          row = (adiv == 1) ? m : m / adiv;
          if (aidx) row = __ldg(aidx + row);
        }
        arow[j] = row;
      }
    }
  }

  __device__ __forceinline__ void enter_ktile() {
    k0 = kbeg + kt * BK;
    if (!A_MC) {
      if (k0 >= klim && seg + 1 < g.a_nseg) {
        ++seg;
        enter_segment();
      }
      a_fast = g.a[seg].vec_ok != 0 && ((klim - k0) & 3) == 0;
      a_pred = k0 + a_c < klim;
      pa = abase + (k0 - kstart + a_c);
    } else {
      a_fast = g.a[0].vec_ok != 0 && (g.M & 3) == 0;
      a_pred = m0 + a_r < g.M;
      pa = abase + (long long)(k0 + a_c) * ald + (m0 + a_r);
    }
    if (!B_MC) {
      b_fast = g.b.vec_ok != 0 && ((kend - k0) & 3) == 0;
      b_pred = k0 + b_c < kend;
      pb = bbase + (long long)(n0 + b_r) * bld + (k0 + b_c);
    } else {
      b_fast = g.b.vec_ok != 0 && (g.N & 3) == 0;
      b_pred = n0 + b_r < g.N;
      pb = b_gather ? bbase + (n0 + b_r) : bbase + (long long)(k0 + b_c) * bld + (n0 + b_r);
    }
  }

  __device__ __forceinline__ void seek() {     // first tile at or after `tile` with a non-empty reduction range
    const int tiles_per_z = sch.m_tiles * sch.n_tiles;
    kt = 0;
    nkt = 0;
    while (tile < sch.total_tiles) {
      const int z = tile / tiles_per_z, rem = tile - z * tiles_per_z;
      const int mt = rem / sch.n_tiles, nt = rem - mt * sch.n_tiles;
      m0 = mt * BM;
      n0 = nt * BN;
      kbeg = 0;
      kend = g.K;
      if (g.zmode == 2) {
        kbeg = z * g.kchunk;
        kend = min(g.K, kbeg + g.kchunk);
      }
      nkt = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;
      if (nkt > 0) {
        bbase = g.b.base + (g.zmode == 1 ? (long long)z * g.b_bstride : 0);
        seg = 0;
        if (!A_MC)
          while (seg + 1 < g.a_nseg && kbeg >= g.a[seg].kend) ++seg;
        enter_segment();
        enter_ktile();
        return;
      }
      tile += gridDim.x;
    }
  }

  __device__ __forceinline__ void next() {
    if (++kt < nkt) {
      enter_ktile();
      return;
    }
    tile += gridDim.x;
    seek();
  }

  // ---- fast path: aligned float4 items, all four elements in range whenever the item is in range
  __device__ __forceinline__ const float* fast_ptr(int j) const {
    if (j < NAI) {
      if (!A_MC) return (a_pred && arow[A_MC ? 0 : j] >= 0) ? pa + (long long)arow[A_MC ? 0 : j] * ald : nullptr;
      return (a_pred && k0 + a_c + 8 * j < kend) ? pa + (long long)(8 * j) * ald : nullptr;
    }
    const int jb = j - NAI;
    if (!B_MC) return (b_pred && n0 + b_r + 16 * jb < g.N) ? pb + (long long)(16 * jb) * bld : nullptr;
    const int kk = k0 + b_c + BKS * jb;
    if (!(b_pred && kk < kend)) return nullptr;
    if (!b_gather) return pb + (long long)(BKS * jb) * bld;
    const int bdiv = g.b.div;
    int row = (bdiv == 1) ? kk : kk / bdiv;
    const int* bidx = g.b.idx;
    if (bidx) row = __ldg(bidx + row);
    return pb + (long long)row * bld;
  }

  // ---- general path: ragged edges / unaligned operands
  __device__ __forceinline__ void slow_ptr(int j, const float*& p, int& nvalid) const {
    p = nullptr;
    if (j < NAI) {
      if (!A_MC) {
        const int kk = k0 + a_c;
        nvalid = klim - kk;
        if (arow[A_MC ? 0 : j] >= 0 && nvalid > 0) p = abase + (long long)arow[A_MC ? 0 : j] * ald + (kk - kstart);
      } else {
        const int kk = k0 + a_c + 8 * j, m = m0 + a_r;
        nvalid = g.M - m;
        if (kk < kend && nvalid > 0) p = abase + (long long)kk * ald + m;
      }
    } else {
      const int jb = j - NAI;
      if (!B_MC) {
        const int n = n0 + b_r + 16 * jb, kk = k0 + b_c;
        nvalid = kend - kk;
        if (n < g.N && nvalid > 0) p = bbase + (long long)n * bld + kk;
      } else {
        const int kk = k0 + b_c + BKS * jb, n = n0 + b_r;
        nvalid = g.N - n;
        if (kk < kend && nvalid > 0) {
          const int bdiv = g.b.div;
          int row = (bdiv == 1) ? kk : kk / bdiv;
          const int* bidx = g.b.idx;
          if (bidx) row = __ldg(bidx + row);
          p = bbase + (long long)row * bld + n;
        }
      }
    }
  }

  // Loads items [J0, J0 + N) of the current k-tile.  Pointers first (the row-index loads of gathered operands
  // overlap), then the data loads back to back.
  template <int J0, int N>
  __device__ __forceinline__ void load(float4 (&v)[N]) const {
    if (a_fast && b_fast) {
      const float* p[N];
#pragma unroll
      for (int j = 0; j < N; ++j) p[j] = fast_ptr(J0 + j);
#pragma unroll
      for (int j = 0; j < N; ++j)
        v[j] = (p[j] != nullptr) ? __ldg(reinterpret_cast<const float4*>(p[j])) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float* p;
        int nv;
        slow_ptr(J0 + j, p, nv);
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p != nullptr) {
          if (nv > 0) v[j].x = __ldg(p);
          if (nv > 1) v[j].y = __ldg(p + 1);
          if (nv > 2) v[j].z = __ldg(p + 2);
          if (nv > 3) v[j].w = __ldg(p + 3);
        }
      }
    }
  }

  template <bool SPLIT, int J0, int N>
  __device__ __forceinline__ void store(uint32_t stage_addr, const float4 (&v)[N]) const {
    const uint32_t sa = stage_addr, sb = stage_addr + (SPLIT ? 2 : 1) * A_BYTES;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const bool is_a = (J0 + j) < NAI;
      const uint32_t addr = is_a ? sa + a_item_off(a_soff, J0 + j) : sb + b_item_off(b_soff, J0 + j - NAI);
      uint2 hi, lo;
      convert4<SPLIT>(v[j], hi, lo);
      sts64(addr, hi);
      if (SPLIT) sts64(addr + (is_a ? A_BYTES : B_BYTES), lo);
    }
  }
};

template <int NSPLIT, int BN, bool A_MC, bool B_MC>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const GemmDev<float> g, const Sched sch) {
  constexpr bool SPLIT = NSPLIT == 3;
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * (A_BYTES + B_BYTES);
  constexpr int NSTAGE = (kSmemBudget / STAGE_BYTES) < kMaxStages ? (kSmemBudget / STAGE_BYTES) : kMaxStages;
  static_assert(NSTAGE >= 2, "need at least two smem stages");
  constexpr int ACC_COLS = BN;  // fp32 columns per accumulator buffer

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // manual 1024-byte alignment (SWIZZLE_128B atoms)
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) unsigned long long bars[2 * kMaxStages + 4];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kMaxStages]);
  const uint32_t accf0 = smem_u32(&bars[2 * kMaxStages]), acce0 = smem_u32(&bars[2 * kMaxStages + 2]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full0 + 8 * s, kProdWarps);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accf0 + 8 * b, 1);
      mbar_init(acce0 + 8 * b, kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == kEpiWarps) {  // the MMA warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int kred = (g.zmode == 2) ? g.kchunk : g.K;   // reduction length handled by one tile (upper bound)
  const int tiles_per_z = sch.m_tiles * sch.n_tiles;

  if (warp >= kEpiWarps + 1) {
    // ============================================================== producers
    // Software-pipelined in quarters of a k-tile over two register buffers: while one quarter is converted and
    // stored, the loads of the next quarter (possibly of the next k-tile or output tile) are already in flight, so
    // every producer thread keeps NI/4 .. NI/2 independent LDG.128 outstanding (the loop is L2-latency bound
    // otherwise).
    using Prod = Producer<BN, A_MC, B_MC>;
    constexpr int NI = Prod::NI, Q = NI / 4;
    static_assert(NI % 4 == 0, "items per thread must split into four groups");
    Prod pr(g, sch, threadIdx.x - (kEpiWarps + 1) * 32);
    int stage = 0;
    uint32_t phase = 0;
    float4 b0[Q], b1[Q];
    pr.tile = blockIdx.x;
    pr.seek();
    bool more = pr.valid();
    if (more) pr.template load<0, Q>(b0);
    while (more) {
      pr.template load<Q, Q>(b1);
      mbar_wait(empty0 + 8 * stage, phase ^ 1);
      const uint32_t sbase = smem_u32(smem) + stage * STAGE_BYTES;
      pr.template store<SPLIT, 0, Q>(sbase, b0);
      pr.template load<2 * Q, Q>(b0);
      pr.template store<SPLIT, Q, Q>(sbase, b1);
      pr.template load<3 * Q, Q>(b1);
      pr.template store<SPLIT, 2 * Q, Q>(sbase, b0);
      pr.next();
      more = pr.valid();
      if (more) pr.template load<0, Q>(b0);
      pr.template store<SPLIT, 3 * Q, Q>(sbase, b1);
      fence_proxy_async();      // make the generic-proxy stores visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(full0 + 8 * stage);
      if (++stage == NSTAGE) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == kEpiWarps) {
    // ============================================================== MMA issuer
    const uint32_t idesc = make_idesc(BN, A_MC, B_MC);
    constexpr uint32_t A_KSTEP = A_MC ? (2048 >> 4) : (32 >> 4);   // descriptor start-address advance per K = 16
    constexpr uint32_t B_KSTEP = B_MC ? (2048 >> 4) : (32 >> 4);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t acc_phase[2] = {0, 0};
    int local = 0;
    for (int tile = blockIdx.x; tile < sch.total_tiles; tile += gridDim.x, ++local) {
      const int z = tile / tiles_per_z;
      int kbeg = 0, kend = g.K;
      if (g.zmode == 2) {
        kbeg = z * g.kchunk;
        kend = min(g.K, kbeg + g.kchunk);
      }
      const int nkt = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;   // trailing split-K slices may be empty
      const int buf = local & 1;
      mbar_wait(acce0 + 8 * buf, acc_phase[buf] ^ 1);     // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * ACC_COLS;
      for (int kt = 0; kt < nkt; ++kt) {
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sA = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sB = sA + (SPLIT ? 2 : 1) * A_BYTES;
          const uint64_t dAhi = make_desc(sA, A_MC), dBhi = make_desc(sB, B_MC);
          const uint64_t dAlo = make_desc(sA + A_BYTES, A_MC), dBlo = make_desc(sB + B_BYTES, B_MC);
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint64_t a_adv = static_cast<uint64_t>(ks * A_KSTEP), b_adv = static_cast<uint64_t>(ks * B_KSTEP);
            const uint32_t first = (kt > 0 || ks > 0) ? 1u : 0u;
            if (SPLIT) {
              umma_f16(tmem_d, dAlo + a_adv, dBhi + b_adv, idesc, first);
              umma_f16(tmem_d, dAhi + a_adv, dBlo + b_adv, idesc, 1u);
              umma_f16(tmem_d, dAhi + a_adv, dBhi + b_adv, idesc, 1u);
            } else {
              umma_f16(tmem_d, dAhi + a_adv, dBhi + b_adv, idesc, first);
            }
          }
          umma_commit(empty0 + 8 * stage);                 // smem slot free once these MMAs retire
          if (kt == nkt - 1) umma_commit(accf0 + 8 * buf);  // accumulator complete
        }
        __syncwarp();
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (nkt == 0 && lane == 0) umma_commit(accf0 + 8 * buf);
      acc_phase[buf] ^= 1;
    }
  } else {
    // ============================================================== epilogue (warps 0-3 <-> TMEM lanes 32w..32w+31)
    uint32_t acc_phase[2] = {0, 0};
    int local = 0;
    float pslope = 0.f;
    if (g.act == DOST_ACT_PRELU) pslope = __ldg(g.prelu_slope);
    else if (g.act == DOST_ACT_LEAKY) pslope = g.act_slope;
    for (int tile = blockIdx.x; tile < sch.total_tiles; tile += gridDim.x, ++local) {
      const int z = tile / tiles_per_z, rem = tile - z * tiles_per_z;
      const int mt = rem / sch.n_tiles, nt = rem - mt * sch.n_tiles;
      const int m = mt * BM + warp * 32 + lane, n0 = nt * BN;
      const int buf = local & 1;
      int kbeg = 0, kend = g.K;
      if (g.zmode == 2) {
        kbeg = z * g.kchunk;
        kend = min(g.K, kbeg + g.kchunk);
      }
      const bool empty_k = kend <= kbeg;
      const long long coff = (g.zmode == 1) ? (long long)z * g.c_bstride : 0;
      mbar_wait(accf0 + 8 * buf, acc_phase[buf]);
      acc_phase[buf] ^= 1;
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + buf * ACC_COLS;
      const bool row_ok = m < g.M;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr0 + c0, r);
        tmem_ld_wait();
        if (c0 + 32 >= BN) {       // last read of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acce0 + 8 * buf);
        }
        // Only `row_ok` differs between lanes.  No lane may leave this iteration early: the tcgen05.ld above is a
        // warp-collective (.sync.aligned) instruction, so the warp reconverges explicitly at the end of every chunk.
        if (row_ok && n0 + c0 < g.N) {
        if (g.zmode == 2) {
          float* ws = g.ws + ((long long)z * g.M + m) * g.N + n0 + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c0 + j < g.N) ws[j] = empty_k ? 0.f : __uint_as_float(r[j]);
        } else {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const int n = n0 + c0 + j4 * 4;
          if (n >= g.N) break;
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = empty_k ? 0.f : __uint_as_float(r[j4 * 4 + j]);
          if (g.epi_vec && n + 4 <= g.N) {
            if (g.bias) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(g.bias + n));
              v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
            }
            if (g.out_pre)
              *reinterpret_cast<float4*>(g.out_pre + coff + (long long)m * g.ld_pre + n) = make_float4(v[0], v[1], v[2], v[3]);
            if (g.act != DOST_ACT_NONE) {
#pragma unroll
              for (int j = 0; j < 4; ++j) v[j] = (v[j] > 0.f) ? v[j] : pslope * v[j];
            }
            if (g.dact_saved) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(g.dact_saved + coff + (long long)m * g.ld_dact + n));
              v[0] *= (t.x > 0.f) ? 1.f : g.dact_slope;
              v[1] *= (t.y > 0.f) ? 1.f : g.dact_slope;
              v[2] *= (t.z > 0.f) ? 1.f : g.dact_slope;
              v[3] *= (t.w > 0.f) ? 1.f : g.dact_slope;
            }
            if (g.residual) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(g.residual + coff + (long long)m * g.ld_res + n));
              v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
            }
            float4* op = reinterpret_cast<float4*>(g.out + coff + (long long)m * g.ldc + n);
            if (g.accumulate) {
              const float4 t = *op;
              v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
            }
            *op = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (n + j >= g.N) continue;
              float x = v[j];
              if (g.bias) x += __ldg(g.bias + n + j);
              if (g.out_pre) g.out_pre[coff + (long long)m * g.ld_pre + n + j] = x;
              if (g.act != DOST_ACT_NONE) x = (x > 0.f) ? x : pslope * x;
              if (g.dact_saved) x *= (__ldg(g.dact_saved + coff + (long long)m * g.ld_dact + n + j) > 0.f) ? 1.f : g.dact_slope;
              if (g.residual) x += __ldg(g.residual + coff + (long long)m * g.ld_res + n + j);
              float* op = g.out + coff + (long long)m * g.ldc + n + j;
              if (g.accumulate) x += *op;
              *op = x;
            }
          }
        }
        }  // zmode
        }  // row_ok
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
  (void)kred;
}

template <int NSPLIT, int BN, bool A_MC, bool B_MC>
static int launch_one(const GemmDev<float>& g, const Sched& sch, cudaStream_t st) {
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = (NSPLIT == 3 ? 2 : 1) * (A_BYTES + B_BYTES);
  constexpr int NSTAGE = (kSmemBudget / STAGE_BYTES) < kMaxStages ? (kSmemBudget / STAGE_BYTES) : kMaxStages;
  const int smem = NSTAGE * STAGE_BYTES + 1024;
  auto kern = gemm_tc_kernel<NSPLIT, BN, A_MC, B_MC>;
  static PerDevice cfg_once;
  if (bool* cfg_flag = cfg_once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("gemm_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return DOST_ERR_LAUNCH;
    }
    *cfg_flag = true;
  }
  const int sms = sm_count();
  const int grid = sch.total_tiles < sms ? sch.total_tiles : sms;
  kern<<<grid, kThreads, smem, st>>>(g, sch);
  return check_launch("gemm_tc");
}

template <int NSPLIT, int BN>
static int launch_modes(const GemmDev<float>& g, const Sched& sch, bool a_mc, bool b_mc, cudaStream_t st) {
  if (a_mc && b_mc) return launch_one<NSPLIT, BN, true, true>(g, sch, st);
  if (a_mc) return launch_one<NSPLIT, BN, true, false>(g, sch, st);
  if (b_mc) return launch_one<NSPLIT, BN, false, true>(g, sch, st);
  return launch_one<NSPLIT, BN, false, false>(g, sch, st);
}

}  // namespace tc

bool gemm_tc_supported(const GemmDev<float>& g, bool a_mc, bool b_mc) {
  (void)b_mc;
  // concatenated A segments must not straddle a 64-wide k-tile
  if (!a_mc && g.a_nseg > 1) {
    int prev = 0;
    for (int s = 0; s < g.a_nseg; ++s) {
      if ((g.a[s].kend - prev) % tc::BK != 0) return false;
      prev = g.a[s].kend;
    }
  }
  // tiny problems stay on the FMA pipe (a 128 x BN tile would be mostly padding)
  return g.M >= 64 && g.N >= 16 && (long long)g.M * g.N * (long long)g.K >= (1LL << 21);
}

int launch_gemm_tc(const GemmDev<float>& g, int precision, bool a_mc, bool b_mc, int batch, int split, cudaStream_t st) {
  tc::Sched sch;
  const int bn = g.N <= 64 ? 64 : (g.N <= 128 ? 128 : 256);
  sch.m_tiles = (g.M + tc::BM - 1) / tc::BM;
  sch.n_tiles = (g.N + bn - 1) / bn;
  sch.z_count = batch > 1 ? batch : (split > 1 ? split : 1);
  const long long total = (long long)sch.m_tiles * sch.n_tiles * sch.z_count;
  if (total > 0x7fffffffLL) {
    set_error("gemm_tc: too many tiles");
    return DOST_ERR_ARG;
  }
  sch.total_tiles = (int)total;
  if (precision == 1) {
    if (bn == 64) return tc::launch_modes<3, 64>(g, sch, a_mc, b_mc, st);
    if (bn == 128) return tc::launch_modes<3, 128>(g, sch, a_mc, b_mc, st);
    return tc::launch_modes<3, 256>(g, sch, a_mc, b_mc, st);
  }
  if (bn == 64) return tc::launch_modes<1, 64>(g, sch, a_mc, b_mc, st);
  if (bn == 128) return tc::launch_modes<1, 128>(g, sch, a_mc, b_mc, st);
  return tc::launch_modes<1, 256>(g, sch, a_mc, b_mc, st);
}

}  // namespace dost
