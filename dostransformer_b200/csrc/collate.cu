// On-device batch assembly from a packed crystal store.
//
// Replaces the CPU DataLoader collate of the launchers (reference main_eDOS.py:54, main_phDOS.py:52-54:
// torch_geometric.loader.DataLoader -> Batch.from_data_list): node / edge tensors of the selected crystals are
// concatenated, edge_index gets each crystal's node offset, batch = repeat_interleave(arange(B), n_b), per-crystal
// fields are stacked / concatenated.
//
// Packed store layout (all resident in HBM, built once per dataset): node rows of crystal c live at rows
// node_ptr[c] .. node_ptr[c+1] of every node table, edge rows at edge_ptr[c] .. edge_ptr[c+1] of every edge table,
// edge_index holds crystal-LOCAL node ids.  A batch is a list of crystal ids; since a crystal's rows are contiguous
// in the store and in the batch, assembly is a segmented copy: one block column per crystal, 16-byte lanes when the
// row size allows, 4-byte lanes otherwise.  Pure HBM traffic: bytes read + bytes written, no atomics, no sorting.
#include "common.cuh"

namespace dost {
namespace {

// One block: out_ptr[b+1] = sum_{i<=b} (all_ptr[ids[i]+1] - all_ptr[ids[i]]) for the node and the edge table, and the
// largest node count (the to_dense_batch padding length, DOSTransformer.py:61) for callers that want it on the device.
__global__ void __launch_bounds__(1024) collate_ptr_kernel(const long long* __restrict__ ids, long long B,
                                                           const long long* __restrict__ node_ptr_all,
                                                           const long long* __restrict__ edge_ptr_all, long long C,
                                                           long long* __restrict__ node_ptr_out,
                                                           long long* __restrict__ edge_ptr_out,
                                                           long long* __restrict__ nmax_out, int* __restrict__ bad) {
  __shared__ long long wsum_n[32], wsum_e[32];
  __shared__ long long carry_n, carry_e, smax[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    carry_n = 0;
    carry_e = 0;
    node_ptr_out[0] = 0;
    if (edge_ptr_out) edge_ptr_out[0] = 0;
  }
  long long lmax = 0;
  __syncthreads();
  for (long long base = 0; base < B; base += blockDim.x) {
    long long i = base + tid;
    long long n = 0, e = 0;
    if (i < B) {
      long long c = ids[i];
      if (c < 0 || c >= C) {
        *bad = 1;
      } else {
        n = node_ptr_all[c + 1] - node_ptr_all[c];
        if (edge_ptr_all) e = edge_ptr_all[c + 1] - edge_ptr_all[c];
      }
    }
    lmax = max64(lmax, n);
    long long sn = n, se = e;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long tn = __shfl_up_sync(0xffffffffu, sn, o), te = __shfl_up_sync(0xffffffffu, se, o);
      if (lane >= o) {
        sn += tn;
        se += te;
      }
    }
    if (lane == 31) {
      wsum_n[warp] = sn;
      wsum_e[warp] = se;
    }
    __syncthreads();
    if (warp == 0) {
      long long vn = wsum_n[lane], ve = wsum_e[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        long long tn = __shfl_up_sync(0xffffffffu, vn, o), te = __shfl_up_sync(0xffffffffu, ve, o);
        if (lane >= o) {
          vn += tn;
          ve += te;
        }
      }
      wsum_n[lane] = vn;
      wsum_e[lane] = ve;
    }
    __syncthreads();
    long long off_n = carry_n + (warp ? wsum_n[warp - 1] : 0), off_e = carry_e + (warp ? wsum_e[warp - 1] : 0);
    if (i < B) {
      node_ptr_out[i + 1] = off_n + sn;
      if (edge_ptr_out) edge_ptr_out[i + 1] = off_e + se;
    }
    __syncthreads();
    if (tid == blockDim.x - 1) {
      carry_n = off_n + sn;
      carry_e = off_e + se;
    }
    __syncthreads();
  }
  if (nmax_out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = max64(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    if (lane == 0) smax[warp] = lmax;
    __syncthreads();
    if (tid == 0) {
      long long m = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = max64(m, smax[w]);
      *nmax_out = m;
    }
  }
}

// Segmented row copy.  blockIdx.y walks the crystals of the batch, blockIdx.x the chunks of one crystal's segment.
// U = uint4 when every row is a multiple of 16 bytes (rows then start 16-byte aligned in both tables), else uint32.
template <typename U>
__global__ void __launch_bounds__(256) seg_copy_kernel(const U* __restrict__ src, const long long* __restrict__ src_ptr,
                                                       const long long* __restrict__ ids,
                                                       const long long* __restrict__ out_ptr, long long B,
                                                       long long row_units, U* __restrict__ dst) {
  for (long long b = blockIdx.y; b < B; b += gridDim.y) {
    const long long c = ids[b];
    long long s0, rows, d0;
    if (src_ptr) {           // ragged table: rows src_ptr[c] .. src_ptr[c+1] go to rows out_ptr[b] ..
      s0 = src_ptr[c];
      rows = src_ptr[c + 1] - s0;
      d0 = out_ptr[b];
    } else {                 // per-crystal table: row c goes to row b
      s0 = c;
      rows = 1;
      d0 = b;
    }
    const U* s = src + s0 * row_units;
    U* d = dst + d0 * row_units;
    const long long n = rows * row_units;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      d[i] = s[i];
  }
}

// edge_index[:, e] = local id + node offset of the crystal in the batch; batch[r] = b.
__global__ void __launch_bounds__(256) collate_index_kernel(const long long* __restrict__ ei_all, long long E_all,
                                                            const long long* __restrict__ node_ptr_all,
                                                            const long long* __restrict__ edge_ptr_all,
                                                            const long long* __restrict__ ids,
                                                            const long long* __restrict__ node_ptr_out,
                                                            const long long* __restrict__ edge_ptr_out, long long B,
                                                            long long E_out, long long* __restrict__ ei_out,
                                                            long long* __restrict__ batch_out) {
  for (long long b = blockIdx.y; b < B; b += gridDim.y) {
    const long long c = ids[b];
    const long long n0 = node_ptr_out[b], nn = node_ptr_out[b + 1] - n0;
    const long long start = (long long)blockIdx.x * blockDim.x + threadIdx.x, step = (long long)gridDim.x * blockDim.x;
    if (batch_out)
      for (long long i = start; i < nn; i += step) batch_out[n0 + i] = b;
    if (ei_out) {
      const long long es = edge_ptr_all[c], ne = edge_ptr_all[c + 1] - es, eo = edge_ptr_out[b];
      for (long long i = start; i < ne; i += step) {
        ei_out[eo + i] = ei_all[es + i] + n0;
        ei_out[E_out + eo + i] = ei_all[E_all + es + i] + n0;
      }
    }
  }
}

dim3 seg_grid(long long B, long long units_total) {
  long long per = B > 0 ? units_total / B : 0;                       // mean units per crystal
  long long gx = max64(1, min64(64, (per + 256 * 4 - 1) / (256 * 4)));  // ~4 units per thread
  long long gy = max64(1, min64(B, 65535));
  return dim3((unsigned)gx, (unsigned)gy, 1);
}

}  // namespace
}  // namespace dost

using namespace dost;

extern "C" int dost_collate_ptr(const int64_t* ids, long long B, const int64_t* node_ptr_all,
                                const int64_t* edge_ptr_all, long long C, int64_t* node_ptr_out, int64_t* edge_ptr_out,
                                int64_t* nmax_out, int32_t* bad_flag, dost_stream_t stream) {
  DOST_REQUIRE(ids && node_ptr_all && node_ptr_out && bad_flag && B >= 0 && C > 0, "collate_ptr: bad args");
  DOST_REQUIRE((edge_ptr_all == nullptr) == (edge_ptr_out == nullptr), "collate_ptr: edge tables must come in pairs");
  collate_ptr_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(
      (const long long*)ids, B, (const long long*)node_ptr_all, (const long long*)edge_ptr_all, C,
      (long long*)node_ptr_out, (long long*)edge_ptr_out, (long long*)nmax_out, bad_flag);
  return check_launch("collate_ptr");
}

extern "C" int dost_collate_rows(const void* src, const int64_t* src_ptr_all, const int64_t* ids,
                                 const int64_t* out_ptr, long long B, long long rows_out, long long row_bytes,
                                 void* dst, dost_stream_t stream) {
  if (B == 0 || rows_out == 0 || row_bytes == 0) return DOST_OK;
  DOST_REQUIRE(src && ids && dst && B > 0 && rows_out > 0, "collate_rows: bad args");
  DOST_REQUIRE((src_ptr_all == nullptr) == (out_ptr == nullptr), "collate_rows: ragged tables need both pointers");
  DOST_REQUIRE(src_ptr_all || rows_out == B, "collate_rows: a per-crystal table yields one row per crystal");
  DOST_REQUIRE(row_bytes > 0 && row_bytes % 4 == 0, "collate_rows: rows must be a multiple of 4 bytes (got %lld)",
               row_bytes);
  const bool wide = row_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(src) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  if (wide) {
    long long ru = row_bytes / 16;
    seg_copy_kernel<uint4><<<seg_grid(B, rows_out * ru), 256, 0, st>>>(
        (const uint4*)src, (const long long*)src_ptr_all, (const long long*)ids, (const long long*)out_ptr, B, ru,
        (uint4*)dst);
  } else {
    long long ru = row_bytes / 4;
    seg_copy_kernel<uint32_t><<<seg_grid(B, rows_out * ru), 256, 0, st>>>(
        (const uint32_t*)src, (const long long*)src_ptr_all, (const long long*)ids, (const long long*)out_ptr, B, ru,
        (uint32_t*)dst);
  }
  return check_launch("collate_rows");
}

extern "C" int dost_collate_index(const int64_t* edge_index_all, long long E_all, const int64_t* node_ptr_all,
                                  const int64_t* edge_ptr_all, const int64_t* ids, const int64_t* node_ptr_out,
                                  const int64_t* edge_ptr_out, long long B, long long N_out, long long E_out,
                                  int64_t* edge_index_out, int64_t* batch_out, dost_stream_t stream) {
  if (B == 0 || (N_out == 0 && E_out == 0)) return DOST_OK;
  DOST_REQUIRE(ids && node_ptr_all && node_ptr_out && B > 0, "collate_index: bad args");
  DOST_REQUIRE(!edge_index_out || (edge_index_all && edge_ptr_all && edge_ptr_out), "collate_index: edge tables missing");
  DOST_REQUIRE(edge_index_out || batch_out, "collate_index: nothing to write");
  if (E_out == 0) edge_index_out = nullptr;
  collate_index_kernel<<<seg_grid(B, (edge_index_out ? E_out : N_out)), 256, 0, (cudaStream_t)stream>>>(
      (const long long*)edge_index_all, E_all, (const long long*)node_ptr_all, (const long long*)edge_ptr_all,
      (const long long*)ids, (const long long*)node_ptr_out, (const long long*)edge_ptr_out, B, E_out,
      (long long*)edge_index_out, (long long*)batch_out);
  return check_launch("collate_index");
}
