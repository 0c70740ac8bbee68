// tcgen05 (5th-gen tensor core) GEMM for sm_100a with the same gather/concat prologue and fused epilogue as the
// FMA-pipe kernel in gemm.cu.
//
//   C[m,n] = epi( sum_k A(m,k) * B(n,k) ),  A/B fp32 in HBM, bf16 operands in shared memory, fp32 accumulate in TMEM.
//
// Precision modes
//   bf16x3 (NSPLIT = 3): x = hi + lo with hi = bf16(x), lo = bf16(x - hi); the product is accumulated as
//                        hi*hi + hi*lo + lo*hi (three tcgen05.mma per k-step into the same TMEM accumulator).
//                        Relative error per product ~2^-16: this is the fp32-parity path.
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
// Activations are fp32 in HBM, so TMA (which cannot convert) is not used for the operands; the conversion is what the
// producer warps are for.
#include <cuda_bf16.h>
#include "gemm_common.cuh"

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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, SWIZZLE_128B, version 1 (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   [0,14) start >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
// K-major : 8-row x 128-byte atoms, SBO = 1024 B between 8-row groups, LBO unused (1).
// MN-major: 64(mn) x 8(k) atoms of 1024 B, SBO = 1024 B between k-groups, LBO = 8192 B between 64-row blocks (BK = 64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, bool mn_major) {
  uint64_t d = static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(mn_major ? (8192 >> 4) : 1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor, kind::f16 (InstrDescriptor): c_format F32 (1) @4, a/b format BF16 (1) @7/@10,
// a_major @15, b_major @16 (0 = K-major, 1 = MN-major), N>>3 @17, M>>4 @24.
__device__ __forceinline__ uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo_elem, float hi_elem) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);   // .x = lo_elem (low 16 bits)
  return *reinterpret_cast<uint32_t*>(&v);
}

// 8 fp32 -> 8 bf16 (hi) and, if SPLIT, 8 bf16 residuals (lo).
template <bool SPLIT>
__device__ __forceinline__ void convert8(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_bf16(x[2 * i], x[2 * i + 1]);
    if (SPLIT) {
      const float h0 = __uint_as_float(h[i] << 16), h1 = __uint_as_float(h[i] & 0xFFFF0000u);
      l[i] = pack_bf16(x[2 * i] - h0, x[2 * i + 1] - h1);
    }
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  if (SPLIT) lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void load8(const float* p, int nvalid, bool vec_ok, float (&x)[8]) {
  if (p != nullptr && nvalid >= 8 && vec_ok) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
    x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = (p != nullptr && j < nvalid) ? __ldg(p + j) : 0.f;
  }
}

// Byte offset of the 16-byte chunk holding elements (row r, k = 8*kc .. 8*kc+7) in a K-major SWIZZLE_128B tile.
__device__ __forceinline__ uint32_t off_kmajor(int r, int kc) {
  return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((kc ^ (r & 7)) << 4));
}
// Byte offset of the 16-byte chunk holding elements (rows 8*rc .. 8*rc+7, reduction index k) in an MN-major tile.
__device__ __forceinline__ uint32_t off_mnmajor(int rc, int k) {
  return static_cast<uint32_t>((rc >> 3) * 8192 + (k >> 3) * 1024 + (k & 7) * 128 + (((rc & 7) ^ (k & 7)) << 4));
}

// Fills one operand tile (ROWS x 64 bf16, SWIZZLE_128B) from fp32 global memory.  `addr(i0, i1, p, nvalid)` returns the
// source pointer of an item (8 consecutive fp32) or nullptr, and how many of the 8 are in range.
//   K-major  (MN_MAJOR = false): item = (row r, 16-byte chunk kc): i0 = r, i1 = kc, 8 consecutive k
//   MN-major (MN_MAJOR = true) : item = (row chunk rc, k)        : i0 = rc, i1 = k, 8 consecutive rows
template <bool SPLIT, int ROWS, bool MN_MAJOR, typename AddrFn>
__device__ __forceinline__ void fill_tile(unsigned char* s_hi, unsigned char* s_lo, int pt, AddrFn addr, bool vec_ok) {
  constexpr int kItems = ROWS * 8 / kProdThreads;   // items per thread
  constexpr int kBatch = 4;
  static_assert(kItems % kBatch == 0 || kItems < kBatch, "item count must be a multiple of the batch");
  constexpr int NB = kItems < kBatch ? kItems : kBatch;
#pragma unroll
  for (int b0 = 0; b0 < kItems; b0 += NB) {
    const float* p[NB];
    int nv[NB];
    uint32_t off[NB];
    float4 v0[NB], v1[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int i = pt + (b0 + j) * kProdThreads;
      int i0, i1;
      if (!MN_MAJOR) {
        i1 = i & 7;
        i0 = i >> 3;
        off[j] = off_kmajor(i0, i1);
      } else {
        constexpr int RC = ROWS / 8;
        i0 = i % RC;
        i1 = i / RC;
        off[j] = off_mnmajor(i0, i1);
      }
      addr(i0, i1, p[j], nv[j]);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const bool fast = (p[j] != nullptr) && vec_ok && nv[j] >= 8;
      v0[j] = fast ? __ldg(reinterpret_cast<const float4*>(p[j])) : make_float4(0.f, 0.f, 0.f, 0.f);
      v1[j] = fast ? __ldg(reinterpret_cast<const float4*>(p[j]) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      float x[8] = {v0[j].x, v0[j].y, v0[j].z, v0[j].w, v1[j].x, v1[j].y, v1[j].z, v1[j].w};
      const bool fast = (p[j] != nullptr) && vec_ok && nv[j] >= 8;
      if (!fast && p[j] != nullptr) {      // ragged edge or unaligned rows: scalar loads
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = (e < nv[j]) ? __ldg(p[j] + e) : 0.f;
      }
      uint4 hi, lo;
      convert8<SPLIT>(x, hi, lo);
      *reinterpret_cast<uint4*>(s_hi + off[j]) = hi;
      if (SPLIT) *reinterpret_cast<uint4*>(s_lo + off[j]) = lo;
    }
  }
}

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
    const int pt = threadIdx.x - (kEpiWarps + 1) * 32;  // 0..255
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < sch.total_tiles; tile += gridDim.x) {
      const int z = tile / tiles_per_z, rem = tile - z * tiles_per_z;
      const int mt = rem / sch.n_tiles, nt = rem - mt * sch.n_tiles;
      const int m0 = mt * BM, n0 = nt * BN;
      int kbeg = 0, kend = g.K;
      long long aoff = 0, boff = 0;
      if (g.zmode == 1) {
        aoff = (long long)z * g.a_bstride;
        boff = (long long)z * g.b_bstride;
      } else if (g.zmode == 2) {
        kbeg = z * g.kchunk;
        kend = min(g.K, kbeg + g.kchunk);
      }
      const int nkt = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;   // trailing split-K slices may be empty
      for (int kt = 0; kt < nkt; ++kt) {
        const int k0 = kbeg + kt * BK;
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        unsigned char* sA = smem + stage * STAGE_BYTES;
        unsigned char* sAlo = sA + A_BYTES;
        unsigned char* sB = sA + (SPLIT ? 2 : 1) * A_BYTES;
        unsigned char* sBlo = sB + B_BYTES;
        // Loads are issued in batches of kBatch items (2 x LDG.128 each) before any conversion so that every producer
        // thread keeps 2*kBatch independent 16-byte loads in flight (the loop is latency-bound otherwise).
        // ---------------------------------------------------------------- A tile: 128 rows x 64 k
        if (!A_MC) {
          int s = 0;
          while (s + 1 < g.a_nseg && k0 >= g.a[s].kend) ++s;
          const int kstart = (s == 0) ? 0 : g.a[s - 1].kend;
          const int klim = min(kend, g.a[s].kend);
          const bool vok = g.a[s].vec_ok != 0;
          const float* abase = g.a[s].base + aoff;
          const long long ald = g.a[s].ld;
          const int* aidx = g.a[s].idx;
          const int adiv = g.a[s].div;
          fill_tile<SPLIT, BM, false>(sA, sAlo, pt, [&](int r, int kc, const float*& p, int& nvalid) {
            const int m = m0 + r, kk = k0 + kc * 8;
            nvalid = klim - kk;
            p = nullptr;
            if (m < g.M && kk < klim) {
              long long row = m / adiv;
              if (aidx) row = __ldg(aidx + row);
              p = abase + row * ald + (kk - kstart);
            }
          }, vok);
        } else {
          const bool vok = g.a[0].vec_ok != 0;
          const float* abase = g.a[0].base + aoff;
          const long long ald = g.a[0].ld;
          fill_tile<SPLIT, BM, true>(sA, sAlo, pt, [&](int rc, int k, const float*& p, int& nvalid) {
            const int m = m0 + rc * 8, kk = k0 + k;
            nvalid = g.M - m;
            p = (m < g.M && kk < kend) ? abase + (long long)kk * ald + m : nullptr;
          }, vok);
        }
        // ---------------------------------------------------------------- B tile: BN rows x 64 k
        if (!B_MC) {
          const bool vok = g.b.vec_ok != 0;
          const float* bbase = g.b.base + boff;
          const long long bld = g.b.ld;
          fill_tile<SPLIT, BN, false>(sB, sBlo, pt, [&](int r, int kc, const float*& p, int& nvalid) {
            const int n = n0 + r, kk = k0 + kc * 8;
            nvalid = kend - kk;
            p = (n < g.N && kk < kend) ? bbase + (long long)n * bld + kk : nullptr;
          }, vok);
        } else {
          const bool vok = g.b.vec_ok != 0;
          const float* bbase = g.b.base + boff;
          const long long bld = g.b.ld;
          const int* bidx = g.b.idx;
          const int bdiv = g.b.div;
          fill_tile<SPLIT, BN, true>(sB, sBlo, pt, [&](int rc, int k, const float*& p, int& nvalid) {
            const int n = n0 + rc * 8, kk = k0 + k;
            nvalid = g.N - n;
            p = nullptr;
            if (n < g.N && kk < kend) {
              long long row = kk / bdiv;
              if (bidx) row = __ldg(bidx + row);
              p = bbase + row * bld + n;
            }
          }, vok);
        }
        fence_proxy_async();      // make the generic-proxy stores visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(full0 + 8 * stage);
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
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
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("gemm_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return DOST_ERR_LAUNCH;
    }
    configured = true;
  }
  const int grid = sch.total_tiles < kNumSMs ? sch.total_tiles : kNumSMs;
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
