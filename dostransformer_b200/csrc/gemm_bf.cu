// TMA-fed tcgen05 GEMM over bf16 operand planes (sm_100a).
//
//   C[m,n] = epi( sum_k A(m,k) * B(n,k) ),   fp32 accumulation in TMEM.
//
// Operands live in HBM as bf16 "planes": hi = bf16(x) and, for the error-compensated fp32-parity mode (bf16x3),
// lo = bf16(x - hi).  The planes are written by the kernels that produce the activations (dost_split_planes, the
// LayerNorm kernels, this kernel's own epilogue), so the GEMM main loop is pure TMA -> shared memory -> tcgen05.mma:
// no thread touches operand data.  (gemm_tc.cu converts fp32 operands on the fly instead and stays for gathered /
// batched operands; its producers are bound by the L1TEX pipe at ~0.2 PFLOP/s.)
//
// Warp roles (320 threads, one persistent CTA per SM):
//   warps 0-7  epilogue: warp w owns TMEM lanes 32 (w % 4) .. +31 and the column half w / 4 of the accumulator:
//              tcgen05.ld 32x32 fp32 chunks -> XOR-swizzled shared-memory transpose -> row-contiguous (coalesced) bias /
//              activation / act' / residual / fp32 store and optional bf16 hi/lo plane store for the next GEMM
//   warp  8    MMA issuer (one elected lane): 1 (bf16) or 3 (bf16x3: lo*hi + hi*lo + hi*hi) tcgen05.mma per K=16 step
//   warp  9    TMA producer (one elected lane): cp.async.bulk.tensor.2d with SWIZZLE_128B boxes, mbarrier complete_tx
// K-major operands (reduction index contiguous) use one {64 k, ROWS} box per plane and k-tile; MN-major operands (row
// index contiguous: the transposed operands of the weight-gradient GEMMs) use ROWS/64 boxes of {64 rows, 64 k}.  Both
// land in the canonical UMMA SWIZZLE_128B layouts, so nothing is ever transposed in memory.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace dost {
namespace bf {

constexpr int BM = 128, BK = 64;
constexpr int kMaxStages = 6;
constexpr int kEpiBytes = 8 * 4096;             // staging: 8 epilogue warps x one 32 x 32 fp32 tile, or 16 warps x one 32 x 16 tile
constexpr int kSmemBudget = 225 * 1024;         // stages + staging (+ 1 KB alignment slack <= 227 KB)

struct Maps {
  CUtensorMap a_hi[3], a_lo[3], b_hi, b_lo;
  CUtensorMap b_hi2, b_lo2;      // K-major B with a 128-row box (CTA pair: each CTA stages half of the 256 output columns)
  // TMA-store epilogue (Params::tma_epi): fp32 out {32 columns, 32 rows} boxes, SWIZZLE_128B; bf16 planes {32, 32}, SWIZZLE_64B
  CUtensorMap c_out, c_hi, c_lo;
  CUtensorMap c_pre;             // the pre-activation copy (out_pre), same box as c_out
};

struct Params {
  int M, N, K;
  int a_nseg, a_kend[3];
  int a_mc, b_mc;
  int zmode, kchunk;            // 0: single, 1: batched (z = batch index), 2: split-K (raw partials to ws)
  long long c_bstride, res_bstride;   // batched: element strides of out / residual between batches
  const int* b_rowoff;          // batched, ragged B: rows of problem z start at b_rowoff[z] (instead of a batch stride)
  const int* c_rowoff;          // batched, ragged out: rows of problem z are written at c_rowoff[z] + m ...
  const int* c_rowlim;          // ... for m < c_rowlim[z] only
  int m_tiles, n_tiles, total_tiles;
  const float* bias;
  const float* rowbias; long long ld_rowbias; int rowbias_div;
  int act; float act_slope; const float* prelu_slope;
  float* out_pre; long long ld_pre;
  const __nv_bfloat16* dact_hi; long long ld_dact; float dact_slope;
  const float* residual; long long ld_res;
  float* out; long long ldc; int accumulate;
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo; long long ld_op;
  float* ws;
  float* colpart;               // [ceil(M / 32)][N] per-32-row column sums of the stored value (bias gradients)
  unsigned int* out_gate;       // [N / 32][ld_gate] words (word-major: a warp's 32 rows are 32 consecutive words): bit j of word
                                // (w, m) = pre-activation value of element (m, 32 w + j) > 0 (written by the TMA-store epilogue)
  const unsigned int* dact_gate;   // the same bits read back as the act' mask
  long long ld_gate;
  int tma_epi;                  // epilogue variant: values finished in the accumulator's own layout (thread = row), staged in
                                // the TMA-swizzled layout and written with cp.async.bulk.tensor stores (no second register pass)
};

using namespace ptx;      // tcgen05 / TMA / mbarrier wrappers (tc_ptx.cuh)

// MN-major operands of this kernel come in 64-row TMA boxes: 8192 B between 64-element blocks along MN
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, bool mn_major) { return make_smem_desc(saddr, mn_major ? 8192u : 0u); }

// Predicated global accesses for the epilogue's row tails: a guarded `if (row_ok) *p = v;` compiles to a branch around its
// own basic block (address arithmetic included), eight of them per feature; these stay straight-line code.
__device__ __forceinline__ void stg128_if(float* p, float a, float b, float c, float d, bool ok) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q st.global.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(p), "f"(a),
               "f"(b), "f"(c), "f"(d), "r"((uint32_t)ok)
               : "memory");
}
__device__ __forceinline__ void stg64_if(void* p, uint32_t a, uint32_t b, bool ok) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q st.global.v2.b32 [%0], {%1, %2};\n\t}" ::"l"(p), "r"(a), "r"(b),
               "r"((uint32_t)ok)
               : "memory");
}
__device__ __forceinline__ float4 ldg128_nc_if(const float* p, bool ok) {      // read-only data (residual, row bias)
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
               : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
               : "l"(p), "r"((uint32_t)ok)
               : "memory");
  return v;
}
__device__ __forceinline__ float4 ldg128_if(const float* p, bool ok) {         // data this kernel also writes (accumulate)
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q ld.global.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
               : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
               : "l"(p), "r"((uint32_t)ok)
               : "memory");
  return v;
}
__device__ __forceinline__ uint2 ldg64_nc_if(const void* p, bool ok) {
  uint2 v = make_uint2(0u, 0u);
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q ld.global.nc.v2.b32 {%0, %1}, [%2];\n\t}"
               : "+r"(v.x), "+r"(v.y)
               : "l"(p), "r"((uint32_t)ok)
               : "memory");
  return v;
}
__device__ __forceinline__ uint4 ldg128u_nc_if(const void* p, bool ok) {
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];\n\t}"
               : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w)
               : "l"(p), "r"((uint32_t)ok)
               : "memory");
  return v;
}
// ---- TMA store (shared -> global) of one box, bulk async-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1), "r"(src)
               : "memory");
}

struct TileInfo {
  int m0, n0, kbeg, kend, nkt, z;
};
// CTA2: a tile is 256 x BN, owned by a CTA pair; CTA `rank` holds rows m0 .. m0+127 (m0 includes 128 * rank).
template <int BN, bool CTA2>
__device__ __forceinline__ TileInfo tile_info(const Params& p, int tile, int rank) {
  TileInfo t;
  const int per_z = p.m_tiles * p.n_tiles;
  t.z = tile / per_z;
  const int rem = tile - t.z * per_z;
  const int mt = rem / p.n_tiles, nt = rem - mt * p.n_tiles;
  t.m0 = mt * (CTA2 ? 2 * BM : BM) + rank * BM;
  t.n0 = nt * BN;
  t.kbeg = 0;
  t.kend = p.K;
  if (p.zmode == 2) {
    t.kbeg = t.z * p.kchunk;
    t.kend = min(p.K, t.kbeg + p.kchunk);
  }
  t.nkt = (t.kend > t.kbeg) ? (t.kend - t.kbeg + BK - 1) / BK : 0;
  return t;
}

// EW = number of epilogue warps: 8 (32-column chunks; both epilogue variants) or 16 (TMA-store epilogue only, 16-column
// chunks: twice the warps per scheduler to hide the chunk's latency chain, <= 112 registers per thread).
template <int NSPLIT, int BN, bool CTA2, int EW = 8>
__global__ void __launch_bounds__((EW + 2) * 32, 1) gemm_bf_kernel(const __grid_constant__ Maps maps, const Params p) {
  constexpr int kEpiWarps = EW;
  constexpr bool SPLIT = NSPLIT == 3;
  constexpr int BNL = CTA2 ? BN / 2 : BN;          // B rows (output columns) this CTA stages per k-tile
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BNL * BK * 2;
  constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * (A_BYTES + B_BYTES);
  constexpr int NSTAGE_RAW = (kSmemBudget - kEpiBytes) / STAGE_BYTES;
  constexpr int NSTAGE = NSTAGE_RAW < kMaxStages ? NSTAGE_RAW : kMaxStages;
  static_assert(NSTAGE >= 2, "need at least two smem stages");

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_smem = reinterpret_cast<float*>(smem + NSTAGE * STAGE_BYTES);
  __shared__ __align__(8) unsigned long long bars[2 * kMaxStages + 4];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kMaxStages]);
  const uint32_t accf0 = smem_u32(&bars[2 * kMaxStages]), acce0 = smem_u32(&bars[2 * kMaxStages + 2]);
  // CTA pair: rank 0 (leader) issues the MMAs for both CTAs; the full / accumulator-empty barriers that the MMA warp
  // waits on live in the leader and are signalled remotely by the peer's TMA loads / epilogue warps; the empty /
  // accumulator-full barriers are per CTA and signalled by multicast tcgen05.commit.
  const int rank = CTA2 ? (int)cluster_ctarank() : 0;
  const int tile0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accf0 + 8 * b, 1);
      mbar_init(acce0 + 8 * b, CTA2 ? 2 * kEpiWarps : kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == kEpiWarps) {
    if (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const bool a_mc = p.a_mc != 0, b_mc = p.b_mc != 0;

  if (warp == kEpiWarps + 1) {
    // ============================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < p.total_tiles; tile += tile_step) {
        const TileInfo t = tile_info<BN, CTA2>(p, tile, rank);
        const int nb0 = t.n0 + rank * BNL;       // first B row (output column) staged by this CTA
        int seg = 0;
        for (int kt = 0; kt < t.nkt; ++kt) {
          const int k0 = t.kbeg + kt * BK;
          while (seg + 1 < p.a_nseg && k0 >= p.a_kend[seg]) ++seg;
          const int kseg = k0 - (seg == 0 ? 0 : p.a_kend[seg - 1]);   // k inside the segment's own planes
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          // CTA pair: both CTAs' loads complete on the leader's barrier, which expects the bytes of both
          const uint32_t bar = CTA2 ? mapa_rank(full0 + 8 * stage, 0) : full0 + 8 * stage;
          if (!CTA2) mbar_expect_tx(bar, STAGE_BYTES);
          else if (rank == 0) mbar_expect_tx(full0 + 8 * stage, 2 * STAGE_BYTES);
          const uint32_t sA = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sB = sA + (SPLIT ? 2 : 1) * A_BYTES;
          const int zb = (p.zmode == 1) ? t.z : 0;
          auto load = [&](uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2) {
            if (CTA2) tma_load_3d_2cta(dst, m, c0, c1, c2, bar);
            else tma_load_3d(dst, m, c0, c1, c2, bar);
          };
          // ragged B: every problem reads the same 2-D plane at its own row offset (rows past the problem's own belong
          // to the next problem; the caller masks them)
          const int brow = p.b_rowoff ? __ldg(p.b_rowoff + t.z) : 0;
          const int zbb = p.b_rowoff ? 0 : zb;
          if (!a_mc) {
            load(sA, &maps.a_hi[seg], kseg, t.m0, zb);
            if (SPLIT) load(sA + A_BYTES, &maps.a_lo[seg], kseg, t.m0, zb);
          } else {
#pragma unroll
            for (int b = 0; b < BM / 64; ++b) {
              load(sA + b * 8192, &maps.a_hi[0], t.m0 + 64 * b, k0, zb);
              if (SPLIT) load(sA + A_BYTES + b * 8192, &maps.a_lo[0], t.m0 + 64 * b, k0, zb);
            }
          }
          if (!b_mc) {
            load(sB, CTA2 ? &maps.b_hi2 : &maps.b_hi, k0, nb0 + brow, zbb);
            if (SPLIT) load(sB + B_BYTES, CTA2 ? &maps.b_lo2 : &maps.b_lo, k0, nb0 + brow, zbb);
          } else {
#pragma unroll
            for (int b = 0; b < BNL / 64; ++b) {
              load(sB + b * 8192, &maps.b_hi, nb0 + 64 * b, k0 + brow, zbb);
              if (SPLIT) load(sB + B_BYTES + b * 8192, &maps.b_lo, nb0 + 64 * b, k0 + brow, zbb);
            }
          }
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == kEpiWarps) {
    // ============================================================== MMA issuer
    const uint32_t idesc = make_idesc(CTA2 ? 2 * BM : BM, BN, a_mc, b_mc);
    const uint32_t a_kstep = a_mc ? (2048 >> 4) : (32 >> 4);   // descriptor start-address advance per K = 16
    const uint32_t b_kstep = b_mc ? (2048 >> 4) : (32 >> 4);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t acc_phase[2] = {0, 0};
    int local = 0;
    for (int tile = tile0; tile < p.total_tiles && rank == 0; tile += tile_step, ++local) {
      const TileInfo t = tile_info<BN, CTA2>(p, tile, rank);
      const int buf = local & 1;
      mbar_wait(acce0 + 8 * buf, acc_phase[buf] ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * BN;
      for (int kt = 0; kt < t.nkt; ++kt) {
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sA = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sB = sA + (SPLIT ? 2 : 1) * A_BYTES;
          const uint64_t dAhi = make_desc(sA, a_mc), dBhi = make_desc(sB, b_mc);
          const uint64_t dAlo = make_desc(sA + A_BYTES, a_mc), dBlo = make_desc(sB + B_BYTES, b_mc);
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint64_t a_adv = static_cast<uint64_t>(ks * a_kstep), b_adv = static_cast<uint64_t>(ks * b_kstep);
            const uint32_t first = (kt > 0 || ks > 0) ? 1u : 0u;
            auto mma = [&](uint64_t da, uint64_t db, uint32_t acc) {
              if (CTA2) umma_f16_2cta(tmem_d, da, db, idesc, acc);
              else umma_f16(tmem_d, da, db, idesc, acc);
            };
            if (SPLIT) {
              mma(dAlo + a_adv, dBhi + b_adv, first);
              mma(dAhi + a_adv, dBlo + b_adv, 1u);
              mma(dAhi + a_adv, dBhi + b_adv, 1u);
            } else {
              mma(dAhi + a_adv, dBhi + b_adv, first);
            }
          }
          if (CTA2) {
            umma_commit_2cta(empty0 + 8 * stage);
            if (kt == t.nkt - 1) umma_commit_2cta(accf0 + 8 * buf);
          } else {
            umma_commit(empty0 + 8 * stage);
            if (kt == t.nkt - 1) umma_commit(accf0 + 8 * buf);
          }
        }
        __syncwarp();
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (t.nkt == 0 && lane == 0) {
        if (CTA2) umma_commit_2cta(accf0 + 8 * buf);
        else umma_commit(accf0 + 8 * buf);
      }
      acc_phase[buf] ^= 1;
    }
  } else {
    // ============================================================== epilogue (8 warps)
    uint32_t acc_phase[2] = {0, 0};
    int local = 0;
    float pslope = 0.f;
    if (p.act == DOST_ACT_PRELU) pslope = __ldg(p.prelu_slope);
    else if (p.act == DOST_ACT_LEAKY) pslope = p.act_slope;
    constexpr int NPART = EW / 4;                          // column parts of the accumulator (one per warp of a lane quarter)
    constexpr int CH = BN / NPART;                         // accumulator columns handled by this warp
    constexpr int CW = EW == 16 ? 16 : 32;                 // columns per chunk (one tcgen05.ld, one staging tile, one store)
    constexpr int STG_BYTES = kEpiBytes / EW;              // this warp's staging tile: 32 rows x CW fp32, or hi + lo planes
    const int quad = warp & 3, half = warp >> 2;
    const uint32_t stg = smem_u32(epi_smem) + warp * STG_BYTES;
    const int rsub = lane >> 3, cj = lane & 7;             // after the transpose: rows rsub + 4 i, 16-byte column chunk cj
    const uint32_t acce_remote0 = CTA2 ? mapa_rank(acce0, 0) : 0u;   // the leader's accumulator-empty barriers
    // Feature switches and the hot pointers / strides live in registers for the whole kernel: read from the parameter
    // bank inside the chunk loop, every `if (p.x)` is a constant load + dependent branch (~a dozen serialised
    // round trips per chunk with only two epilogue warps per scheduler to hide them).  The empty asm statements keep the
    // compiler from rematerialising the loads.
    enum : uint32_t { F_BIAS = 1, F_ROWBIAS = 2, F_PRE = 4, F_ACT = 8, F_DACT = 16, F_RES = 32, F_OUT = 64, F_ACC = 128,
                      F_HI = 256, F_LO = 512, F_COLPART = 1024, F_SPLITK = 2048, F_ROWLIM = 4096, F_TMA = 8192, F_GATE_OUT = 16384, F_GATE_IN = 32768 };
    uint32_t feat = (p.bias ? F_BIAS : 0u) | (p.rowbias ? F_ROWBIAS : 0u) | (p.out_pre ? F_PRE : 0u) |
                    (p.act != DOST_ACT_NONE ? F_ACT : 0u) | (p.dact_hi ? F_DACT : 0u) | (p.residual ? F_RES : 0u) |
                    (p.out ? F_OUT : 0u) | (p.accumulate ? F_ACC : 0u) | (p.out_hi ? F_HI : 0u) | (p.out_lo ? F_LO : 0u) |
                    (p.colpart ? F_COLPART : 0u) | (p.zmode == 2 ? F_SPLITK : 0u) | (p.c_rowlim ? F_ROWLIM : 0u) |
                    (p.tma_epi ? F_TMA : 0u) | (p.out_gate ? F_GATE_OUT : 0u) | (p.dact_gate ? F_GATE_IN : 0u);
    int Mv = p.M, Nv = p.N;
    long long ldc_v = p.ldc, ldop_v = p.ld_op, ldres_v = p.ld_res;
    const float* bias_v = p.bias;
    const float* res_v = p.residual;
    float* out_v = p.out;
    __nv_bfloat16* hi_v = p.out_hi;
    __nv_bfloat16* lo_v = p.out_lo;
    asm volatile("" : "+r"(feat), "+r"(Mv), "+r"(Nv));
    asm volatile("" : "+l"(ldc_v), "+l"(ldop_v), "+l"(ldres_v));
    asm volatile("" : "+l"(bias_v), "+l"(res_v), "+l"(out_v), "+l"(hi_v), "+l"(lo_v));
    for (int tile = tile0; tile < p.total_tiles; tile += tile_step, ++local) {
      const TileInfo t = tile_info<BN, CTA2>(p, tile, rank);
      const int buf = local & 1;
      // TMA-store epilogue: the side inputs of a chunk (residual, act' mask) do not depend on the accumulator.  They are
      // requested one chunk ahead - the first chunk's before this warp even waits for the MMAs of the tile - so their DRAM
      // latency (~1-2 us under load, once per chunk per warp otherwise) hides behind the wait and the previous chunk.
      uint4 pre[CW / 4];   // one buffer for both kinds (a launch has a residual OR an act' mask on this path, never both)
      const int mrow_t = t.m0 + quad * 32 + lane;
      const long long zres = (p.zmode == 1) ? (long long)t.z * p.res_bstride : 0;
      auto prefetch_side = [&](int c0n) {
        const int n0 = t.n0 + half * CH + c0n;
        const bool ok = mrow_t < Mv;
        if (feat & F_RES) {
          const float* rp = res_v + zres + (long long)mrow_t * ldres_v + n0;
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) pre[j] = ldg128u_nc_if(rp + 4 * j, ok && n0 + 4 * j < Nv);
        } else if (feat & F_DACT) {
          const __nv_bfloat16* dp = p.dact_hi + (long long)mrow_t * p.ld_dact + n0;
#pragma unroll
          for (int j = 0; j < CW / 8; ++j) pre[j] = ldg128u_nc_if(dp + 8 * j, ok && n0 + 8 * j < Nv);
        }
      };
      // 1-bit gates: one word per row and 32 columns (16-column chunks: the word is shared by two consecutive chunks)
      uint32_t gate_w = 0u;
      auto prefetch_gate = [&](int c0n) {
        const int n0 = t.n0 + half * CH + c0n;
        gate_w = 0u;
        if (mrow_t < Mv && n0 < Nv) gate_w = __ldg(p.dact_gate + (long long)(n0 >> 5) * p.ld_gate + mrow_t);
      };
      if ((feat & F_TMA) && (feat & (F_RES | F_DACT))) prefetch_side(0);
      if (feat & F_GATE_IN) prefetch_gate(0);
      mbar_wait(accf0 + 8 * buf, acc_phase[buf]);
      acc_phase[buf] ^= 1;
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * BN + half * CH;
      const int mbase = t.m0 + quad * 32 + rsub;           // this thread's first row after the transpose
      const int nbase = t.n0 + half * CH + cj * 4;
      uint32_t r[CW];
      tmem_ld_chunk<CW>(taddr0, r);
      if ((feat & F_TMA) || EW == 16) {
        // ---------------------------------------------------------------------------------- TMA-store epilogue
        // Everything happens in the accumulator's own layout (tcgen05.ld 32x32b: thread = row, CW consecutive columns
        // in registers): bias / row bias / activation / act' mask / residual are applied there, the finished values
        // are written once into this warp's staging tile in the TMA swizzle, and one elected lane issues the
        // cp.async.bulk.tensor store.  TMA clips rows >= M and columns >= N.  No transposed second register pass, no
        // per-thread global address arithmetic or predicated stores.
        const int mrow = t.m0 + quad * 32 + lane;
        const bool row_ok = mrow < Mv;
        const bool warp_rows_ok = t.m0 + quad * 32 < Mv;
        uint32_t gate_acc = 0u;
        // fp32 values of this chunk -> staging tile (rows of CW * 4 bytes in the TMA swizzle of that pitch: 128 B: row & 7,
        // 64 B: (row >> 1) & 3) -> one bulk-tensor store: kind 0 = out, 1 = out_pre, 2 = out += (reduction store)
        auto stage_f32 = [&](const float (&val)[CW], int kind, int n0s) {
          if (lane == 0) bulk_wait_read0();                  // the previous store has finished READING the tile
          __syncwarp();
          const uint32_t swf = CW == 32 ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
          for (int j = 0; j < CW / 4; ++j)
            sts128(stg + lane * (CW * 4) + ((j ^ swf) << 4), __float_as_uint(val[4 * j]), __float_as_uint(val[4 * j + 1]),
                   __float_as_uint(val[4 * j + 2]), __float_as_uint(val[4 * j + 3]));
          fence_async_smem();                                // generic-proxy smem writes -> visible to the async proxy (TMA)
          __syncwarp();
          if (lane == 0) {
            const int mw = t.m0 + quad * 32;
            const int zc = (p.zmode == 1) ? t.z : 0;        // batched: the box never spans two problems (rows >= M clipped)
            if (kind == 2) tma_reduce_add_3d(&maps.c_out, stg, n0s, mw, zc);
            else tma_store_3d(kind == 1 ? &maps.c_pre : &maps.c_out, stg, n0s, mw, zc);
            bulk_commit();
          }
        };
#pragma unroll 1
        for (int c0 = 0; c0 < CH; c0 += CW) {
          const int n0 = t.n0 + half * CH + c0;            // first column of the chunk (warp-uniform)
          const bool live = warp_rows_ok && n0 < Nv;

          tmem_ld_wait();
          float v[CW];
#pragma unroll
          for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(r[j]);
          if (c0 + CW < CH) {
            tmem_ld_chunk<CW>(taddr0 + c0 + CW, r);        // next chunk streams in while this one is processed
          } else {                                         // last read of this accumulator: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CTA2) mbar_arrive_cluster(acce_remote0 + 8 * buf);
              else mbar_arrive(acce0 + 8 * buf);
            }
          }
          if (!live) continue;
          if (feat & F_BIAS) {                             // the same CW values for every lane: broadcast loads
#pragma unroll
            for (int j = 0; j < CW / 4; ++j) {
              const float4 b = ldg128_nc_if(bias_v + n0 + 4 * j, n0 + 4 * j < Nv);
              v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          }
          if (feat & F_ROWBIAS) {
            const float* rb = p.rowbias + (long long)(mrow / p.rowbias_div) * p.ld_rowbias + n0;
#pragma unroll
            for (int j = 0; j < CW / 4; ++j) {
              const float4 b = ldg128_nc_if(rb + 4 * j, row_ok && n0 + 4 * j < Nv);
              v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          }
          if (feat & F_PRE) stage_f32(v, 1, n0);
          if (feat & F_GATE_OUT) {
            uint32_t bits = 0u;
#pragma unroll
            for (int j = 0; j < CW; ++j) bits |= (v[j] > 0.f ? 1u : 0u) << j;
            if (CW == 32) {
              if (row_ok) p.out_gate[(long long)(n0 >> 5) * p.ld_gate + mrow] = bits;
            } else {                                        // two 16-column chunks per word
              if (n0 & 16) {
                if (row_ok) p.out_gate[(long long)(n0 >> 5) * p.ld_gate + mrow] = gate_acc | (bits << 16);
              } else {
                gate_acc = bits;
              }
            }
          }
          if (feat & F_ACT) {
#pragma unroll
            for (int j = 0; j < CW; ++j) v[j] = (v[j] > 0.f) ? v[j] : pslope * v[j];
          }
          if (feat & F_GATE_IN) {
            const float ds = p.dact_slope;
            const uint32_t gsh = gate_w >> (CW == 32 ? 0 : (n0 & 16));
#pragma unroll
            for (int j = 0; j < CW; ++j) v[j] = ((gsh >> j) & 1u) ? v[j] : ds * v[j];
            if (c0 + CW < CH && (CW == 32 || (n0 & 16))) prefetch_gate(c0 + CW);
          }
          if (feat & F_DACT) {
            const float ds = p.dact_slope;
#pragma unroll
            for (int j = 0; j < CW / 8; ++j) {
              const uint32_t w[4] = {pre[j].x, pre[j].y, pre[j].z, pre[j].w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float m0 = __uint_as_float(w[q] << 16), m1 = __uint_as_float(w[q] & 0xFFFF0000u);
                v[8 * j + 2 * q] = (m0 > 0.f) ? v[8 * j + 2 * q] : ds * v[8 * j + 2 * q];
                v[8 * j + 2 * q + 1] = (m1 > 0.f) ? v[8 * j + 2 * q + 1] : ds * v[8 * j + 2 * q + 1];
              }
            }
          }
          if (feat & F_RES) {
#pragma unroll
            for (int j = 0; j < CW / 4; ++j) {
              v[4 * j] += __uint_as_float(pre[j].x); v[4 * j + 1] += __uint_as_float(pre[j].y);
              v[4 * j + 2] += __uint_as_float(pre[j].z); v[4 * j + 3] += __uint_as_float(pre[j].w);
            }
          }
          // `pre` is consumed: the next chunk's side inputs fly during the staging / store of this chunk and the first
          // half of the next one
          if ((feat & (F_RES | F_DACT)) && c0 + CW < CH) prefetch_side(c0 + CW);
          // One staging tile per warp, used once per kind of output of this chunk (pre-activation copy above, fp32 result,
          // bf16 planes): every use first waits until the previous store has READ the tile.
          if (feat & F_OUT) stage_f32(v, (feat & F_ACC) ? 2 : 0, n0);
          if (feat & F_HI) {
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
            uint32_t hi[CW / 2];
#pragma unroll
            for (int j = 0; j < CW / 2; ++j) hi[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
            const uint32_t sw = CW == 32 ? ((lane >> 1) & 3) : ((lane >> 2) & 1);
#pragma unroll
            for (int c = 0; c < CW / 8; ++c)
              sts128(stg + lane * (CW * 2) + ((c ^ sw) << 4), hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
            if (feat & F_LO) {
#pragma unroll
              for (int c = 0; c < CW / 8; ++c) {
                uint32_t lo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const int j = 4 * c + q;
                  lo[q] = pack_bf16(v[2 * j] - __uint_as_float(hi[j] << 16), v[2 * j + 1] - __uint_as_float(hi[j] & 0xFFFF0000u));
                }
                sts128(stg + 64 * CW + lane * (CW * 2) + ((c ^ sw) << 4), lo[0], lo[1], lo[2], lo[3]);
              }
            }
            fence_async_smem();                              // generic-proxy smem writes -> visible to the async proxy (TMA)
            __syncwarp();
            if (lane == 0) {
              const int mw = t.m0 + quad * 32;
              tma_store_3d(&maps.c_hi, stg, n0, mw, 0);
              if (feat & F_LO) tma_store_3d(&maps.c_lo, stg + 64 * CW, n0, mw, 0);
              bulk_commit();
            }
          }
          if (feat & F_COLPART) {
            // column sums over this warp's 32 rows by recursive halving across the lanes (fixed order): each level sends
            // half of the values to the partner lane; once a single value is left (CW = 16: at the last level) the level
            // is a plain pairwise add.  Afterwards lane L holds the sum of column L (CW = 32) or L >> 1 (CW = 16).
#pragma unroll
            for (int j = 0; j < CW; ++j) v[j] = row_ok ? v[j] : 0.f;
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) {
              const int hn = CW == 32 ? sft : sft / 2;     // values left after this level (a function of the unrolled loop
              if (hn >= 1) {                                // variable only: v[] must stay in registers)
                const bool upper = (lane & sft) != 0;
#pragma unroll
                for (int i = 0; i < hn; ++i) {
                  const float send = upper ? v[i] : v[i + hn];
                  const float keep = upper ? v[i + hn] : v[i];
                  v[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
                }
              } else {
                v[0] += __shfl_xor_sync(0xffffffffu, v[0], sft);
              }
            }
            const int col = CW == 32 ? lane : (lane >> 1);
            if ((CW == 32 || (lane & 1) == 0) && n0 + col < Nv)
              p.colpart[(long long)((t.m0 + quad * 32) >> 5) * p.N + n0 + col] = v[0];
          }
        }
        continue;      // next tile
      }
      if constexpr (EW == 8) {
#pragma unroll 1
      for (int c0 = 0; c0 < CH; c0 += 32) {
        const int n = nbase + c0;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);       // requested before the staging round trip, which hides its latency
        if ((feat & F_BIAS) && n < Nv) b4 = __ldg(reinterpret_cast<const float4*>(bias_v + n));
        uint2 sgn[8];
        if (feat & F_DACT) {                               // activation-derivative mask, also requested early
          const int mlim0 = (feat & F_ROWLIM) ? min(Mv, __ldg(p.c_rowlim + t.z)) : Mv;
          const __nv_bfloat16* dp = p.dact_hi + (long long)mbase * p.ld_dact + n;
          const long long st4 = 4 * p.ld_dact;
#pragma unroll
          for (int i = 0; i < 8; ++i) sgn[i] = ldg64_nc_if(dp + i * st4, mbase + 4 * i < mlim0 && n < Nv);
        }
        tmem_ld_wait();
        // ---- stage: lane = accumulator row; 16-byte chunk j goes to column chunk j ^ (row & 7) (conflict-free)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(stg + lane * 128 + ((j ^ (lane & 7)) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        if (c0 + 32 < CH) {
          tmem_ld32(taddr0 + c0 + 32, r);                  // next chunk streams in while this one is processed
        } else {                                           // last read of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CTA2) mbar_arrive_cluster(acce_remote0 + 8 * buf);
            else mbar_arrive(acce0 + 8 * buf);
          }
        }
        __syncwarp();
        float v[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = rsub + 4 * i;
          const float4 q = lds128(stg + rl * 128 + ((cj ^ (rl & 7)) << 4));
          v[i][0] = q.x; v[i][1] = q.y; v[i][2] = q.z; v[i][3] = q.w;
        }
        __syncwarp();                                      // staging tile may be overwritten by the next chunk
        // warp-uniform skip only (the column sums below shuffle across the whole warp); a thread whose 4 columns lie past
        // N (N % 4 == 0: all in or all out) stays in the loop with every access predicated off
        if (t.n0 + half * CH + c0 >= Nv || t.m0 + quad * 32 >= Mv) continue;
        bool mok[8];
        const int mlim = (feat & F_ROWLIM) ? min(Mv, __ldg(p.c_rowlim + t.z)) : Mv;
        const bool nok = n < Nv;
#pragma unroll
        for (int i = 0; i < 8; ++i) mok[i] = nok && mbase + 4 * i < mlim;
        if (feat & F_SPLITK) {
          float* ws = p.ws + ((long long)t.z * Mv + mbase) * Nv + n;
          const bool empty = t.nkt == 0;                    // an empty K slice contributes zeros (its TMEM is stale)
#pragma unroll
          for (int i = 0; i < 8; ++i)
            stg128_if(ws + (long long)(4 * i) * Nv, empty ? 0.f : v[i][0], empty ? 0.f : v[i][1], empty ? 0.f : v[i][2],
                      empty ? 0.f : v[i][3], mok[i]);
          continue;
        }
        if (feat & F_BIAS) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { v[i][0] += b4.x; v[i][1] += b4.y; v[i][2] += b4.z; v[i][3] += b4.w; }
        }
        if (feat & F_ROWBIAS) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 q = ldg128_nc_if(p.rowbias + (long long)((mbase + 4 * i) / p.rowbias_div) * p.ld_rowbias + n, mok[i]);
            v[i][0] += q.x; v[i][1] += q.y; v[i][2] += q.z; v[i][3] += q.w;
          }
        }
        if (feat & F_PRE) {
          float* op = p.out_pre + (long long)mbase * p.ld_pre + n;
          const long long st4 = 4 * p.ld_pre;
#pragma unroll
          for (int i = 0; i < 8; ++i) stg128_if(op + i * st4, v[i][0], v[i][1], v[i][2], v[i][3], mok[i]);
        }
        if (feat & F_ACT) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) v[i][j] = (v[i][j] > 0.f) ? v[i][j] : pslope * v[i][j];
        }
        if (feat & F_DACT) {
          const float ds = p.dact_slope;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // the bf16 bit pattern, moved to the top half of a word, is the same number as an fp32: compare that with 0
            // (-0 and +0 take the slope like every non-positive input; the masks come from activations, never NaN)
            const float m[4] = {__uint_as_float(sgn[i].x << 16), __uint_as_float(sgn[i].x & 0xFFFF0000u),
                                __uint_as_float(sgn[i].y << 16), __uint_as_float(sgn[i].y & 0xFFFF0000u)};
#pragma unroll
            for (int j = 0; j < 4; ++j) v[i][j] = (m[j] > 0.f) ? v[i][j] : ds * v[i][j];
          }
        }
        if (feat & F_RES) {
          const float* rp = res_v + (p.zmode == 1 ? (long long)t.z * p.res_bstride : 0) + (long long)mbase * ldres_v + n;
          const long long st4 = 4 * ldres_v;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 q = ldg128_nc_if(rp + i * st4, mok[i]);
            v[i][0] += q.x; v[i][1] += q.y; v[i][2] += q.z; v[i][3] += q.w;
          }
        }
        if (feat & F_OUT) {
          float* op = out_v + (p.c_rowoff ? (long long)__ldg(p.c_rowoff + t.z) * ldc_v
                                          : (p.zmode == 1 ? (long long)t.z * p.c_bstride : 0)) + (long long)mbase * ldc_v + n;
          const long long st4 = 4 * ldc_v;
          if (feat & F_ACC) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 q = ldg128_if(op + i * st4, mok[i]);
              v[i][0] += q.x; v[i][1] += q.y; v[i][2] += q.z; v[i][3] += q.w;
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) stg128_if(op + i * st4, v[i][0], v[i][1], v[i][2], v[i][3], mok[i]);
        }
        if (feat & F_HI) {
          __nv_bfloat16* hp = hi_v + (long long)mbase * ldop_v + n;
          const long long st4 = 4 * ldop_v;
          uint2 hi[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            hi[i].x = pack_bf16(v[i][0], v[i][1]);
            hi[i].y = pack_bf16(v[i][2], v[i][3]);
            stg64_if(hp + i * st4, hi[i].x, hi[i].y, mok[i]);
          }
          if (feat & F_LO) {
            __nv_bfloat16* lp = lo_v + (long long)mbase * ldop_v + n;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t lx = pack_bf16(v[i][0] - __uint_as_float(hi[i].x << 16), v[i][1] - __uint_as_float(hi[i].x & 0xFFFF0000u));
              const uint32_t ly = pack_bf16(v[i][2] - __uint_as_float(hi[i].y << 16), v[i][3] - __uint_as_float(hi[i].y & 0xFFFF0000u));
              stg64_if(lp + i * st4, lx, ly, mok[i]);
            }
          }
        }
        if (feat & F_COLPART) {       // column sums of this warp's 32 rows, fixed order: the thread's 8 rows, then xor-8, xor-16
          float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) cs[j] += mok[i] ? v[i][j] : 0.f;     // rows past M / the ragged limit do not count
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 8);
            cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 16);
          }
          if (lane < 8 && nok)
            *reinterpret_cast<float4*>(p.colpart + (long long)((t.m0 + quad * 32) >> 5) * p.N + n) = make_float4(cs[0], cs[1], cs[2], cs[3]);
        }
      }
      }   // EW == 8: register epilogue
    }
    if (((feat & F_TMA) || EW == 16) && lane == 0) bulk_wait0();   // the staging tile must outlive the last store's read
  }

  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();     // the peer may still be signalling this CTA's barriers / reading its operands
  if (warp == kEpiWarps) {
    tc_fence_after();
    if (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// Fixed-order reduction of split-K partials (same contract as gemm.cu): out[m,n] (+)= sum_z ws[z][m][n].
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out, long long ldc, int M, int N,
                                     int splits, int accumulate) {
  const long long total4 = (long long)M * N / 4;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total4; i += stride) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < splits; ++z) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(ws + (long long)z * M * N) + i);
      s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
    }
    const long long e = i * 4, m = e / N, n = e % N;
    float4* op = reinterpret_cast<float4*>(out + m * ldc + n);
    if (accumulate) {
      const float4 q = *op;
      s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
    }
    *op = s;
  }
}

// fp32 [rows, cols] (ld) -> bf16 planes [rows, ldp]: hi = bf16(x), lo = bf16(x - hi) (optional).  Columns in
// [cols, ldp) are zero-filled so that padded planes can be used as GEMM operands directly.
__global__ void split_planes_kernel(const float* __restrict__ x, long long ld, long long rows, int cols,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int ldp, int vec_ok) {
  const int chunks = ldp / 8;                      // 8 elements (16 bytes of bf16) per thread step
  const long long total = rows * chunks;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const long long r = i / chunks;
    const int c = static_cast<int>(i - r * chunks) * 8;
    float v[8];
    const float* src = x + r * ld + c;
    if (vec_ok && c + 8 <= cols) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c + j < cols) ? __ldg(src + j) : 0.f;
    }
    uint4 h, l;
    uint32_t* hp = &h.x;
    uint32_t* lp = &l.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      hp[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
      lp[j] = pack_bf16(v[2 * j] - __uint_as_float(hp[j] << 16), v[2 * j + 1] - __uint_as_float(hp[j] & 0xFFFF0000u));
    }
    *reinterpret_cast<uint4*>(hi + r * ldp + c) = h;
    if (lo) *reinterpret_cast<uint4*>(lo + r * ldp + c) = l;
  }
}

// The same conversion for up to kMultiSplit tensors in ONE launch (the weights of the model, once per step): the tensor
// descriptors travel in the kernel-parameter space, blockIdx.y selects the tensor.
constexpr int kMultiSplit = 24;
struct MultiSplit {
  const float* x[kMultiSplit];
  __nv_bfloat16* hi[kMultiSplit];
  __nv_bfloat16* lo[kMultiSplit];
  long long ld[kMultiSplit];
  long long rows[kMultiSplit];
  int cols[kMultiSplit];
  int ldp[kMultiSplit];
};
__global__ void __launch_bounds__(256) split_planes_multi_kernel(const MultiSplit ms) {
  const int t = blockIdx.y;
  const float* __restrict__ x = ms.x[t];
  __nv_bfloat16* __restrict__ hi = ms.hi[t];
  __nv_bfloat16* __restrict__ lo = ms.lo[t];
  const long long ld = ms.ld[t], rows = ms.rows[t];
  const int cols = ms.cols[t], ldp = ms.ldp[t];
  const int vec_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0 && ld % 4 == 0) ? 1 : 0;
  const int chunks = ldp / 8;
  const long long total = rows * chunks;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const long long r = i / chunks;
    const int c = static_cast<int>(i - r * chunks) * 8;
    float v[8];
    const float* src = x + r * ld + c;
    if (vec_ok && c + 8 <= cols) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c + j < cols) ? __ldg(src + j) : 0.f;
    }
    uint4 h, l;
    uint32_t* hp = &h.x;
    uint32_t* lp = &l.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      hp[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
      lp[j] = pack_bf16(v[2 * j] - __uint_as_float(hp[j] << 16), v[2 * j + 1] - __uint_as_float(hp[j] & 0xFFFF0000u));
    }
    *reinterpret_cast<uint4*>(hi + r * ldp + c) = h;
    if (lo) *reinterpret_cast<uint4*>(lo + r * ldp + c) = l;
  }
}

// Same conversion, plus the column sums of x (the bias gradient when x is the gradient of a Linear's output): one pass
// over x instead of two.  Block = 32 column chunks (256 columns) x 8 row lanes over a contiguous row range; the row
// lanes are combined through shared memory and the row ranges by colsum_stage2_kernel, both in a fixed order.
__global__ void __launch_bounds__(256) split_planes_colsum_kernel(const float* __restrict__ x, long long ld, long long rows, int cols,
                                                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                                  int ldp, int vec_ok, long long rows_per_chunk,
                                                                  float* __restrict__ partial) {
  __shared__ float sm[8][32][9];
  const int cg = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cg) * 8;
  const long long rbeg = blockIdx.y * rows_per_chunk, rend = min(rows, rbeg + rows_per_chunk);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < ldp) {
    for (long long r = rbeg + ry; r < rend; r += 8) {
      float v[8];
      const float* src = x + r * ld + c;
      if (vec_ok && c + 8 <= cols) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (c + j < cols) ? __ldg(src + j) : 0.f;
      }
      uint4 h, l;
      uint32_t* hp = &h.x;
      uint32_t* lp = &l.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hp[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
        lp[j] = pack_bf16(v[2 * j] - __uint_as_float(hp[j] << 16), v[2 * j + 1] - __uint_as_float(hp[j] & 0xFFFF0000u));
      }
      *reinterpret_cast<uint4*>(hi + r * ldp + c) = h;
      if (lo) *reinterpret_cast<uint4*>(lo + r * ldp + c) = l;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[ry][cg][j] = acc[j];
  __syncthreads();
  if (ry == 0 && c < cols) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += sm[k][cg][j];
      if (c + j < cols) partial[(long long)blockIdx.y * cols + c + j] = t;
    }
  }
}

// 32 columns x 32 row lanes per block: lane ry sums chunks ry, ry + 32, ... (four independent loads in flight), lanes are
// combined in order (deterministic).  Only ceil(W / 32) blocks exist, so each keeps as many loads in flight as it can.
__global__ void __launch_bounds__(1024) colsum_stage2_kernel(const float* __restrict__ ws, int nchunks, int W, float* __restrict__ out,
                                                             int per_y) {
  // blockIdx.y reduces chunks [y * per_y, min(nchunks, (y + 1) * per_y)) into out[y][W] (a single y: the final sums)
  __shared__ float sm[32][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int cbeg = blockIdx.y * per_y;
  const int cnt = min(per_y, nchunks - cbeg);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < W) {
    const float* src = ws + (long long)cbeg * W + c;
    int b = ry;
    for (; b + 96 < cnt; b += 128) {
      s0 += src[(long long)b * W];
      s1 += src[(long long)(b + 32) * W];
      s2 += src[(long long)(b + 64) * W];
      s3 += src[(long long)(b + 96) * W];
    }
    for (; b < cnt; b += 32) s0 += src[(long long)b * W];
  }
  sm[ry][cx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ry == 0 && c < W) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += sm[k][cx];
    out[(long long)blockIdx.y * W + c] = t;
  }
}

// Column sums of `nchunks` partial rows in a fixed order.  Many partial rows (one per 32 output rows of a large GEMM) are
// first reduced by kMidRows row groups in parallel (only ceil(W / 32) blocks would otherwise walk the whole table).
constexpr int kMidRows = 32;
static int colsum_reduce(const float* ws, int nchunks, int W, float* mid, float* out, cudaStream_t st) {
  if (nchunks > 8 * kMidRows && mid) {
    const int per = (nchunks + kMidRows - 1) / kMidRows;
    const int ny = (nchunks + per - 1) / per;
    colsum_stage2_kernel<<<dim3(ceil_div(W, 32), ny), 1024, 0, st>>>(ws, nchunks, W, mid, per);
    count_launch();
    colsum_stage2_kernel<<<dim3(ceil_div(W, 32), 1), 1024, 0, st>>>(mid, ny, W, out, ny);
  } else {
    colsum_stage2_kernel<<<dim3(ceil_div(W, 32), 1), 1024, 0, st>>>(ws, nchunks, W, out, nchunks);
  }
  return check_launch("column sums");
}

static inline int split_colsum_chunks(long long rows, int ldp) {
  const long long gx = (ldp / 8 + 31) / 32;
  long long nch = (3LL * kNumSMs + gx - 1) / gx;
  const long long maxch = (rows + 31) / 32;
  if (nch > maxch) nch = maxch;
  if (nch < 1) nch = 1;
  return (int)nch;
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}

// bf16 tensor [batch][outer rows][inner elements contiguous], row pitch ld and batch pitch bstride elements;
// box = {64 inner, box_outer rows, 1 batch}.
static int make_map(CUtensorMap* m, const void* base, long long inner, long long outer, long long ld, int box_outer,
                    long long batch = 1, long long bstride = 0) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("gemm_bf16: cuTensorMapEncodeTiled is not available");
    return DOST_ERR_UNSUPPORTED;
  }
  if (batch <= 1 || bstride <= 0) {
    batch = 1;
    bstride = outer * ld;
  }
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bstride * 2};
  cuuint32_t box[3] = {64u, (cuuint32_t)box_outer, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_bf16: cuTensorMapEncodeTiled failed (%d) base=%p inner=%lld outer=%lld ld=%lld box_outer=%d", (int)r, base,
              inner, outer, ld, box_outer);
    return DOST_ERR_ARG;
  }
  return DOST_OK;
}

// Output map for the TMA-store epilogue: [batch][outer rows][inner elements contiguous], row pitch ld and batch pitch
// bstride elements, box {box_cols, 32, 1} (a box never spans two problems of a batch).
static int make_out_map(CUtensorMap* m, const void* base, bool is_f32, int box_cols, long long inner, long long outer, long long ld,
                        long long batch = 1, long long bstride = 0) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("gemm_bf16: cuTensorMapEncodeTiled is not available");
    return DOST_ERR_UNSUPPORTED;
  }
  const int esz = is_f32 ? 4 : 2;
  if (batch <= 1 || bstride <= 0) {
    batch = 1;
    bstride = outer * ld;
  }
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * esz, (cuuint64_t)bstride * esz};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, 32u, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  const int row_bytes = box_cols * esz;        // 128 / 64 / 32: the staging tile is written in the swizzle of its row pitch
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                  : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(m, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_bf16: cuTensorMapEncodeTiled (output) failed (%d) base=%p inner=%lld outer=%lld ld=%lld", (int)r, base, inner,
              outer, ld);
    return DOST_ERR_ARG;
  }
  return DOST_OK;
}

// The TMA-store epilogue (DOST_GEMM_TMA_EPI=0 disables it): one kind of output (fp32 or planes), plain single problem.
static bool tma_epi_enabled() {      // read per call: the parity test flips it between two launches of one process
  const char* e = getenv("DOST_GEMM_TMA_EPI");
  return !(e && e[0] == '0');
}

static bool epi16_enabled(bool colsum) {
  const char* e = getenv("DOST_GEMM_EPI16");
  if (e && e[0] == '0') return false;
  if (e && e[0] == '2') return true;
  return !colsum;
}

template <int NSPLIT, int BN, bool CTA2, int EW = 8>
static int launch(const Maps& maps, const Params& p, cudaStream_t st) {
  constexpr int kThreads = (EW + 2) * 32;
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = (CTA2 ? BN / 2 : BN) * BK * 2;
  constexpr int STAGE_BYTES = (NSPLIT == 3 ? 2 : 1) * (A_BYTES + B_BYTES);
  constexpr int NSTAGE_RAW = (kSmemBudget - kEpiBytes) / STAGE_BYTES;
  constexpr int NSTAGE = NSTAGE_RAW < kMaxStages ? NSTAGE_RAW : kMaxStages;
  const int smem = NSTAGE * STAGE_BYTES + kEpiBytes + 1024;
  auto kern = gemm_bf_kernel<NSPLIT, BN, CTA2, EW>;
  static PerDevice cfg_once;
  if (bool* cfg_flag = cfg_once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("gemm_bf16: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return DOST_ERR_LAUNCH;
    }
    *cfg_flag = true;
  }
  if (!CTA2) {
    const int sms = sm_count();
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    kern<<<grid, kThreads, smem, st>>>(maps, p);
    return check_launch("gemm_bf16");
  }
  // CTA pairs: clusters of 2 along x, one pair per 256-row tile, persistent over the tiles
  const int half_sms = sm_count() / 2;
  const int pairs = p.total_tiles < half_sms ? p.total_tiles : half_sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, maps, p);
  if (e != cudaSuccess) {
    set_error("gemm_bf16: cluster launch failed: %s", cudaGetErrorString(e));
    return DOST_ERR_LAUNCH;
  }
  return check_launch("gemm_bf16 (CTA pairs)");
}

// CTA pairs (tcgen05 cta_group::2, 256 x 256 tiles) for the large problems; DOST_GEMM_2CTA=0 disables them.
static bool use_cta_pairs(int M, int N) {
  static const int enabled = [] {
    const char* e = getenv("DOST_GEMM_2CTA");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return enabled && M > 128 && N > 128;
}

static inline bool al16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

static int run(const dost_gemm_bf16_t* h, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  DOST_REQUIRE(h->M > 0 && h->N > 0 && h->K > 0, "gemm_bf16: bad shape M=%d N=%d K=%d", h->M, h->N, h->K);
  DOST_REQUIRE(h->precision == DOST_PREC_BF16X3 || h->precision == DOST_PREC_BF16, "gemm_bf16: precision must be bf16x3 or bf16");
  DOST_REQUIRE(h->N % 4 == 0, "gemm_bf16: N must be a multiple of 4 (got %d)", h->N);
  const bool split3 = h->precision == DOST_PREC_BF16X3;
  const bool a_mc = h->a_mode == DOST_MC, b_mc = h->b_mode == DOST_MC;
  const int nseg = a_mc ? 1 : h->a_nseg;
  DOST_REQUIRE(nseg >= 1 && nseg <= 3, "gemm_bf16: a_nseg must be 1..3");
  const int bn = h->N <= 64 ? 64 : (h->N <= 128 ? 128 : 256);
  const bool pairs = use_cta_pairs(h->M, h->N);
  const int batch = h->batch < 1 ? 1 : h->batch;
  DOST_REQUIRE(!(batch > 1 && h->split_k > 1), "gemm_bf16: batch and split_k are exclusive");
  DOST_REQUIRE(batch == 1 || (h->a_nseg <= 1 && !h->rowbias && !h->out_pre && !h->dact_hi && !h->out_hi),
               "gemm_bf16: batched problems support bias / activation / residual / fp32 stores only");
  DOST_REQUIRE(batch == 1 || (h->a_bstride % 8 == 0 && h->b_bstride % 8 == 0 && h->c_bstride % 4 == 0 && h->res_bstride % 4 == 0),
               "gemm_bf16: batch strides must keep 16-byte alignment");

  Maps maps;
  Params p;
  p.M = h->M; p.N = h->N; p.K = h->K;
  p.a_nseg = nseg;
  p.a_mc = a_mc; p.b_mc = b_mc;
  int kacc = 0;
  for (int s = 0; s < nseg; ++s) {
    const dost_planes_t& pl = h->a[s];
    DOST_REQUIRE(pl.hi && (!split3 || pl.lo), "gemm_bf16: A segment %d planes missing", s);
    DOST_REQUIRE(al16(pl.hi) && al16(pl.lo) && pl.ld % 8 == 0, "gemm_bf16: A segment %d planes must be 16-byte aligned, ld %% 8 == 0", s);
    const int width = a_mc ? h->K : pl.width;
    if (nseg > 1) DOST_REQUIRE(width % BK == 0, "gemm_bf16: concatenated A segments need widths %% 64 == 0 (got %d)", width);
    kacc += width;
    p.a_kend[s] = kacc;
    int rc;
    if (!a_mc) {       // [M rows, width] K-major: box {64 k, 128 rows}
      const long long rows = batch > 1 ? h->M : pl.rows;
      rc = make_map(&maps.a_hi[s], pl.hi, width, rows, pl.ld, BM, batch, h->a_bstride);
      if (rc == DOST_OK && split3) rc = make_map(&maps.a_lo[s], pl.lo, width, rows, pl.ld, BM, batch, h->a_bstride);
    } else {           // [K rows, M] MN-major: box {64 m, 64 k}
      rc = make_map(&maps.a_hi[s], pl.hi, h->M, h->K, pl.ld, 64, batch, h->a_bstride);
      if (rc == DOST_OK && split3) rc = make_map(&maps.a_lo[s], pl.lo, h->M, h->K, pl.ld, 64, batch, h->a_bstride);
    }
    if (rc != DOST_OK) return rc;
    if (!split3) maps.a_lo[s] = maps.a_hi[s];
  }
  for (int s = nseg; s < 3; ++s) {
    maps.a_hi[s] = maps.a_hi[0];
    maps.a_lo[s] = maps.a_lo[0];
    p.a_kend[s] = kacc;
  }
  DOST_REQUIRE(kacc == h->K, "gemm_bf16: A segment widths sum to %d, K=%d", kacc, h->K);
  {
    const dost_planes_t& pl = h->b;
    DOST_REQUIRE(pl.hi && (!split3 || pl.lo), "gemm_bf16: B planes missing");
    DOST_REQUIRE(al16(pl.hi) && al16(pl.lo) && pl.ld % 8 == 0, "gemm_bf16: B planes must be 16-byte aligned, ld %% 8 == 0");
    int rc;
    // one 2-D plane of pl.rows rows, per-problem row offsets.  Also with a single problem (a batch of one crystal): the
    // map must end at the stored rows, so that the K / N padding reads zeros instead of whatever follows the allocation
    const bool ragged_b = h->b_rowoff != nullptr;
    if (ragged_b) {
      DOST_REQUIRE(pl.rows > 0, "gemm_bf16: ragged B needs the total number of stored rows");
      if (!b_mc) {
        rc = make_map(&maps.b_hi, pl.hi, h->K, pl.rows, pl.ld, bn);
        if (rc == DOST_OK && split3) rc = make_map(&maps.b_lo, pl.lo, h->K, pl.rows, pl.ld, bn);
        if (rc == DOST_OK && pairs) rc = make_map(&maps.b_hi2, pl.hi, h->K, pl.rows, pl.ld, 128);
        if (rc == DOST_OK && pairs && split3) rc = make_map(&maps.b_lo2, pl.lo, h->K, pl.rows, pl.ld, 128);
      } else {
        rc = make_map(&maps.b_hi, pl.hi, h->N, pl.rows, pl.ld, 64);
        if (rc == DOST_OK && split3) rc = make_map(&maps.b_lo, pl.lo, h->N, pl.rows, pl.ld, 64);
      }
    } else if (!b_mc) {       // [N rows, K] K-major: box {64 k, bn rows} (and {64 k, 128 rows} for CTA pairs)
      const long long nrows = (pl.rows > 0 && pl.rows < h->N) ? pl.rows : h->N;   // N may be padded past the stored rows
      rc = make_map(&maps.b_hi, pl.hi, h->K, nrows, pl.ld, bn, batch, h->b_bstride);
      if (rc == DOST_OK && split3) rc = make_map(&maps.b_lo, pl.lo, h->K, nrows, pl.ld, bn, batch, h->b_bstride);
      if (rc == DOST_OK && pairs) rc = make_map(&maps.b_hi2, pl.hi, h->K, nrows, pl.ld, 128, batch, h->b_bstride);
      if (rc == DOST_OK && pairs && split3) rc = make_map(&maps.b_lo2, pl.lo, h->K, nrows, pl.ld, 128, batch, h->b_bstride);
    } else {           // [K rows, N] MN-major
      rc = make_map(&maps.b_hi, pl.hi, h->N, h->K, pl.ld, 64, batch, h->b_bstride);
      if (rc == DOST_OK && split3) rc = make_map(&maps.b_lo, pl.lo, h->N, h->K, pl.ld, 64, batch, h->b_bstride);
    }
    if (rc != DOST_OK) return rc;
    if (!split3) maps.b_lo = maps.b_hi;
    if (!pairs || b_mc) maps.b_hi2 = maps.b_hi;
    if (!pairs || b_mc || !split3) maps.b_lo2 = maps.b_hi2;
  }
  const int split = h->split_k < 1 ? 1 : h->split_k;
  p.zmode = split > 1 ? 2 : (batch > 1 ? 1 : 0);
  p.c_bstride = h->c_bstride;
  p.res_bstride = h->res_bstride;
  p.b_rowoff = h->b_rowoff;      // honoured for a single problem too (index 0): the row limit keeps the padded rows of a
  p.c_rowoff = h->c_rowoff;      // one-crystal batch from being written past the output
  p.c_rowlim = h->c_rowlim;
  DOST_REQUIRE(!p.c_rowoff || p.c_rowlim, "gemm_bf16: ragged output needs both c_rowoff and c_rowlim");
  DOST_REQUIRE(split == 1 || !(p.b_rowoff || p.c_rowoff || p.c_rowlim), "gemm_bf16: ragged problems cannot be split along K");
  p.kchunk = 0;
  p.ws = nullptr;
  if (split > 1) {
    DOST_REQUIRE(!h->bias && !h->rowbias && h->act == DOST_ACT_NONE && !h->out_pre && !h->dact_hi && !h->residual && !h->out_hi,
                 "gemm_bf16: split_k supports only plain (accumulating) fp32 stores");
    const size_t need = sizeof(float) * (size_t)split * h->M * h->N;
    if (!workspace || workspace_bytes < need) {
      set_error("gemm_bf16: split_k workspace too small (%zu < %zu)", workspace_bytes, need);
      return DOST_ERR_WORKSPACE;
    }
    int kchunk = (h->K + split - 1) / split;
    p.kchunk = ((kchunk + BK - 1) / BK) * BK;     // slices end on k-tile boundaries (TMA fetches whole k-tiles)
    p.ws = (float*)workspace;
  }
  p.m_tiles = pairs ? (h->M + 2 * BM - 1) / (2 * BM) : (h->M + BM - 1) / BM;
  p.n_tiles = (h->N + bn - 1) / bn;
  const long long total = (long long)p.m_tiles * p.n_tiles * (batch > 1 ? batch : split);
  DOST_REQUIRE(total <= 0x7fffffffLL, "gemm_bf16: too many tiles");
  p.total_tiles = (int)total;
  p.bias = h->bias;
  p.rowbias = h->rowbias; p.ld_rowbias = h->ld_rowbias; p.rowbias_div = h->rowbias_div < 1 ? 1 : h->rowbias_div;
  p.act = h->act;
  p.act_slope = (h->act == DOST_ACT_RELU) ? 0.f : h->act_slope;
  p.prelu_slope = h->prelu_slope;
  DOST_REQUIRE(h->act != DOST_ACT_PRELU || h->prelu_slope, "gemm_bf16: PReLU needs a slope pointer");
  p.out_pre = h->out_pre; p.ld_pre = h->ld_pre;
  p.dact_hi = (const __nv_bfloat16*)h->dact_hi; p.ld_dact = h->ld_dact; p.dact_slope = h->dact_slope;
  p.residual = h->residual; p.ld_res = h->ld_res;
  p.out = h->out; p.ldc = h->ldc; p.accumulate = h->accumulate;
  p.out_hi = (__nv_bfloat16*)h->out_hi; p.out_lo = (__nv_bfloat16*)h->out_lo; p.ld_op = h->ld_op;
  p.out_gate = h->out_gate; p.dact_gate = h->dact_gate; p.ld_gate = h->ld_gate;
  DOST_REQUIRE(!(h->out_gate || h->dact_gate) || (h->N % 32 == 0 && h->ld_gate >= h->M && batch == 1 && split == 1 &&
                                                  !(h->dact_gate && h->dact_hi)),
               "gemm_bf16: activation gates need N %% 32 == 0, ld_gate >= M, a single un-split problem and no dact_hi");
  p.colpart = nullptr;
  const int nrb = (h->M + 31) / 32;
  if (h->colsum) {
    DOST_REQUIRE(split == 1 && batch == 1, "gemm_bf16: colsum needs a plain (single, un-split) problem");
    const size_t need = sizeof(float) * ((size_t)nrb + kMidRows) * h->N;
    if (!workspace || workspace_bytes < need) {
      set_error("gemm_bf16: colsum workspace too small (%zu < %zu)", workspace_bytes, need);
      return DOST_ERR_WORKSPACE;
    }
    p.colpart = (float*)workspace;
  }
  DOST_REQUIRE(h->out || h->out_hi, "gemm_bf16: no output");
  DOST_REQUIRE(!h->out || (al16(h->out) && h->ldc % 4 == 0), "gemm_bf16: out must be 16-byte aligned with ldc %% 4 == 0");
  DOST_REQUIRE(!h->bias || al16(h->bias), "gemm_bf16: bias must be 16-byte aligned");
  DOST_REQUIRE(!h->rowbias || (al16(h->rowbias) && h->ld_rowbias % 4 == 0), "gemm_bf16: rowbias alignment");
  DOST_REQUIRE(!h->out_pre || (al16(h->out_pre) && h->ld_pre % 4 == 0), "gemm_bf16: out_pre alignment");
  DOST_REQUIRE(!h->residual || (al16(h->residual) && h->ld_res % 4 == 0), "gemm_bf16: residual alignment");
  DOST_REQUIRE(!h->dact_hi || (((uintptr_t)h->dact_hi & 7) == 0 && h->ld_dact % 4 == 0), "gemm_bf16: dact plane alignment");
  DOST_REQUIRE(!h->out_hi || (((uintptr_t)h->out_hi & 7) == 0 && ((uintptr_t)h->out_lo & 7) == 0 && h->ld_op % 4 == 0),
               "gemm_bf16: output plane alignment");

  // ---- TMA-store epilogue: no split-K / ragged rows; the fp32 result, its pre-activation copy and the bf16 planes go out
  // one after the other through the warp's staging tile (an accumulating fp32 output becomes a TMA reduction store)
  p.tma_epi = 0;
  bool ew16 = false;
  maps.c_out = maps.b_hi;
  maps.c_hi = maps.b_hi;
  maps.c_lo = maps.b_hi;
  maps.c_pre = maps.b_hi;
  if (tma_epi_enabled() && split == 1 && !(h->accumulate && h->out_hi) && !p.c_rowoff && !p.c_rowlim &&
      !(h->out_pre && !h->out) && (batch == 1 || !h->out_hi) && h->M >= 32 && h->N >= 32 && !(h->dact_hi && h->residual) &&
      (!h->dact_hi || (al16(h->dact_hi) && h->ld_dact % 8 == 0))) {
    int rc2 = DOST_OK;
    // CTA-pair launches run 16 epilogue warps on 16-column chunks (DOST_GEMM_EPI16=0: 8 warps on 32-column chunks; =2:
    // also with fused column sums, where 16 warps measured 16 % slower: 0.400 -> 0.464 ms at the FFN's fc2 input gradient)
    ew16 = pairs && epi16_enabled(h->colsum != nullptr);
    const int bc = ew16 ? 16 : 32;
    // every kind of output the launch has: fp32 result, pre-activation copy (same shape), bf16 planes (single problems only)
    if (h->out) rc2 = make_out_map(&maps.c_out, h->out, true, bc, h->N, h->M, h->ldc, batch, h->c_bstride);
    if (rc2 == DOST_OK && h->out_pre) rc2 = make_out_map(&maps.c_pre, h->out_pre, true, bc, h->N, h->M, h->ld_pre);
    if (rc2 == DOST_OK && h->out_hi) {
      if (al16(h->out_hi) && al16(h->out_lo) && h->ld_op % 8 == 0) {
        rc2 = make_out_map(&maps.c_hi, h->out_hi, false, bc, h->N, h->M, h->ld_op);
        if (rc2 == DOST_OK && h->out_lo) rc2 = make_out_map(&maps.c_lo, h->out_lo, false, bc, h->N, h->M, h->ld_op);
      } else {
        rc2 = DOST_ERR_ARG;
      }
    }
    if (rc2 == DOST_OK) p.tma_epi = 1;       // (a shape the encoder rejects simply keeps the register epilogue)
    else ew16 = false;
  }

  if ((h->out_gate || h->dact_gate) && !p.tma_epi) {
    set_error("gemm_bf16: activation gates are implemented by the TMA-store epilogue only (this launch is not eligible)");
    return DOST_ERR_UNSUPPORTED;
  }
  int rc;
  if (pairs && ew16) {
    rc = split3 ? launch<3, 256, true, 16>(maps, p, st) : launch<1, 256, true, 16>(maps, p, st);
  } else if (pairs) {
    rc = split3 ? launch<3, 256, true>(maps, p, st) : launch<1, 256, true>(maps, p, st);
  } else if (split3) {
    rc = bn == 64 ? launch<3, 64, false>(maps, p, st)
                  : (bn == 128 ? launch<3, 128, false>(maps, p, st) : launch<3, 256, false>(maps, p, st));
  } else {
    rc = bn == 64 ? launch<1, 64, false>(maps, p, st)
                  : (bn == 128 ? launch<1, 128, false>(maps, p, st) : launch<1, 256, false>(maps, p, st));
  }
  if (rc != DOST_OK) return rc;
  if (h->colsum) {
    rc = colsum_reduce(p.colpart, nrb, h->N, p.colpart + (size_t)nrb * h->N, h->colsum, st);
    if (rc != DOST_OK) return rc;
  }
  if (split > 1) {
    const long long total4 = (long long)h->M * h->N / 4;
    int blocks = (int)min64((total4 + 255) / 256, (long long)kNumSMs * 8);
    if (blocks < 1) blocks = 1;
    splitk_reduce_kernel<<<blocks, 256, 0, st>>>(p.ws, h->out, h->ldc, h->M, h->N, split, h->accumulate);
    rc = check_launch("gemm_bf16 split-k reduce");
  }
  return rc;
}

}  // namespace bf
}  // namespace dost

extern "C" size_t dost_gemm_bf16_workspace_bytes(const dost_gemm_bf16_t* g) {
  if (!g) return 0;
  if (g->colsum) return sizeof(float) * ((size_t)((g->M + 31) / 32) + dost::bf::kMidRows) * g->N;
  if (g->split_k <= 1) return 0;
  return sizeof(float) * (size_t)g->split_k * g->M * g->N;
}

extern "C" int dost_gemm_bf16(const dost_gemm_bf16_t* g, void* workspace, size_t workspace_bytes, dost_stream_t stream) {
  DOST_REQUIRE(g != nullptr, "gemm_bf16: null descriptor");
  return dost::bf::run(g, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" size_t dost_split_planes_colsum_workspace_bytes(long long rows, int cols, long long ldp) {
  return sizeof(float) * (size_t)dost::bf::split_colsum_chunks(rows, (int)ldp) * cols;
}

extern "C" int dost_split_planes_colsum(const float* x, long long ld, long long rows, int cols, void* hi, void* lo, long long ldp,
                                        float* colsum, void* workspace, size_t workspace_bytes, dost_stream_t stream) {
  DOST_REQUIRE(x && hi && colsum, "split_planes_colsum: null pointer");
  DOST_REQUIRE(rows > 0 && cols > 0 && ldp >= cols && ldp % 8 == 0, "split_planes_colsum: need rows > 0, ldp %% 8 == 0, ldp >= cols");
  DOST_REQUIRE(((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0, "split_planes_colsum: planes must be 16-byte aligned");
  const int nch = dost::bf::split_colsum_chunks(rows, (int)ldp);
  const size_t need = sizeof(float) * (size_t)nch * cols;
  if (!workspace || workspace_bytes < need) {
    dost::set_error("split_planes_colsum: workspace too small (%zu < %zu)", workspace_bytes, need);
    return DOST_ERR_WORKSPACE;
  }
  const int vec_ok = (((uintptr_t)x & 15) == 0 && ld % 4 == 0) ? 1 : 0;
  const long long rpc = (rows + nch - 1) / nch;
  dim3 grid((unsigned)((ldp / 8 + 31) / 32), nch);
  cudaStream_t st = (cudaStream_t)stream;
  dost::bf::split_planes_colsum_kernel<<<grid, 256, 0, st>>>(x, ld, rows, cols, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, (int)ldp, vec_ok,
                                                             rpc, (float*)workspace);
  int rc = dost::check_launch("split_planes_colsum");
  if (rc != DOST_OK) return rc;
  return dost::bf::colsum_reduce((const float*)workspace, nch, cols, nullptr, colsum, st);
}

extern "C" int dost_split_planes_multi(int ntensors, const float* const* x, const long long* ld, const long long* rows, const int* cols,
                                       void* const* hi, void* const* lo, const long long* ldp, dost_stream_t stream) {
  DOST_REQUIRE(ntensors >= 0 && x && ld && rows && cols && hi && lo && ldp, "split_planes_multi: null argument");
  using dost::bf::kMultiSplit;
  for (int t0 = 0; t0 < ntensors; t0 += kMultiSplit) {
    dost::bf::MultiSplit ms;
    const int nt = (ntensors - t0) < kMultiSplit ? (ntensors - t0) : kMultiSplit;
    long long maxwork = 0;
    for (int i = 0; i < kMultiSplit; ++i) {
      const int j = t0 + (i < nt ? i : 0);
      DOST_REQUIRE(i >= nt || (x[j] && hi[j] && rows[j] > 0 && cols[j] > 0 && ldp[j] >= cols[j] && ldp[j] % 8 == 0),
                   "split_planes_multi: tensor %d: need rows > 0, ldp %% 8 == 0, ldp >= cols", j);
      DOST_REQUIRE(i >= nt || (((uintptr_t)hi[j] & 15) == 0 && ((uintptr_t)lo[j] & 15) == 0),
                   "split_planes_multi: tensor %d: planes must be 16-byte aligned", j);
      ms.x[i] = x[j];
      ms.hi[i] = (__nv_bfloat16*)hi[j];
      ms.lo[i] = (__nv_bfloat16*)lo[j];
      ms.ld[i] = ld[j];
      ms.rows[i] = i < nt ? rows[j] : 0;
      ms.cols[i] = cols[j];
      ms.ldp[i] = (int)ldp[j];
      const long long work = ms.rows[i] * (ldp[j] / 8);
      if (work > maxwork) maxwork = work;
    }
    if (maxwork == 0) continue;
    long long bx = (maxwork + 255) / 256;
    if (bx > 64) bx = 64;
    dim3 grid((unsigned)bx, nt);
    dost::bf::split_planes_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ms);
    int rc = dost::check_launch("split_planes_multi");
    if (rc != DOST_OK) return rc;
  }
  return DOST_OK;
}

extern "C" int dost_split_planes(const float* x, long long ld, long long rows, int cols, void* hi, void* lo, long long ldp,
                                 dost_stream_t stream) {
  DOST_REQUIRE(x && hi, "split_planes: null pointer");
  DOST_REQUIRE(rows >= 0 && cols > 0 && ldp >= cols && ldp % 8 == 0, "split_planes: need ldp %% 8 == 0 and ldp >= cols");
  DOST_REQUIRE(((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0, "split_planes: planes must be 16-byte aligned");
  if (rows == 0) return DOST_OK;
  const int vec_ok = (((uintptr_t)x & 15) == 0 && ld % 4 == 0) ? 1 : 0;
  const long long total = rows * (ldp / 8);
  int blocks = (int)dost::min64((total + 255) / 256, (long long)dost::kNumSMs * 16);
  if (blocks < 1) blocks = 1;
  dost::bf::split_planes_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, ld, rows, cols, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo,
                                                                          (int)ldp, vec_ok);
  return dost::check_launch("split_planes");
}
