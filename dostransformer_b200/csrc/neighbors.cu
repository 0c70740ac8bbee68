// Periodic neighbour lists and the eDOS bond features on the device.
//
// Replaces the offline CPU graph construction of the reference:
//   * phonon: ase.neighbor_list("ijS", cutoff=r_max, self_interaction=True) and the edge vectors
//     pos[dst] - pos[src] + shift @ lattice (utils.py:267-273);
//   * eDOS:   pymatgen Structure.get_all_neighbors(radius=8) sorted by distance, first 12 kept, short lists padded with
//     index 0 / distance radius + 1, Gaussian distance expansion exp(-(d - mu)^2 / 0.2^2), mu = 0, 0.2, ..., 8.0
//     (data/mat2graph.py:162-179, 185, 212-243).
// ASE and pymatgen are not part of the reference tree (and are unpinned), so the ORDER of the list is this library's own
// canonical one, stated here and followed by oracle/neighbors_oracle.py: edges sorted by (centre atom i, neighbour atom j,
// shift Sx, Sy, Sz); k-nearest selection is stable with respect to that order.
//
// Arithmetic contract (fp64, every operation individually rounded - no FMA contraction - so that the CPU oracle reproduces
// the bits):  s_c = (Sx*L[0][c] + Sy*L[1][c]) + Sz*L[2][c];  v_c = (pos[j][c] - pos[i][c]) + s_c;
// d = sqrt((v_0*v_0 + v_1*v_1) + v_2*v_2);  an image is a neighbour iff d < cutoff, the pair (i, i, S = 0) only with
// self_interaction.  The image ranges searched per pair are a superset derived from the fractional offset and the lattice
// plane spacings, so they do not influence the result.
//
// One warp per centre atom; candidates (j, S) are visited in canonical order 32 at a time and compacted with a ballot, so
// the list comes out sorted without a sort.  Two passes (count, exclusive scan by the caller, fill).
#include "common.cuh"

namespace dost {
namespace {

struct Cell {
  double L[9];      // rows = lattice vectors
  double Li[9];     // inverse (columns give fractional coordinates: f = v * Li)
  double inv_h[3];  // 1 / spacing of the lattice planes normal to direction k
};

__device__ __forceinline__ void make_cell(const double* __restrict__ lat, Cell& c) {
#pragma unroll
  for (int k = 0; k < 9; ++k) c.L[k] = lat[k];
  const double* a = c.L;
  const double* b = c.L + 3;
  const double* cc = c.L + 6;
  const double bxc[3] = {b[1] * cc[2] - b[2] * cc[1], b[2] * cc[0] - b[0] * cc[2], b[0] * cc[1] - b[1] * cc[0]};
  const double cxa[3] = {cc[1] * a[2] - cc[2] * a[1], cc[2] * a[0] - cc[0] * a[2], cc[0] * a[1] - cc[1] * a[0]};
  const double axb[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
  const double vol = a[0] * bxc[0] + a[1] * bxc[1] + a[2] * bxc[2];
  const double iv = 1.0 / vol;
  // inverse of the row-vector matrix: columns are (b x c, c x a, a x b) / vol
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    c.Li[r * 3 + 0] = bxc[r] * iv;
    c.Li[r * 3 + 1] = cxa[r] * iv;
    c.Li[r * 3 + 2] = axb[r] * iv;
  }
  c.inv_h[0] = sqrt(bxc[0] * bxc[0] + bxc[1] * bxc[1] + bxc[2] * bxc[2]) * fabs(iv);
  c.inv_h[1] = sqrt(cxa[0] * cxa[0] + cxa[1] * cxa[1] + cxa[2] * cxa[2]) * fabs(iv);
  c.inv_h[2] = sqrt(axb[0] * axb[0] + axb[1] * axb[1] + axb[2] * axb[2]) * fabs(iv);
}

// The contract arithmetic: every operation rounded on its own.
__device__ __forceinline__ double image_dist(const double* __restrict__ L, const double dp[3], int sx, int sy, int sz,
                                             double v[3]) {
  const double fx = (double)sx, fy = (double)sy, fz = (double)sz;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double s = __dadd_rn(__dadd_rn(__dmul_rn(fx, L[c]), __dmul_rn(fy, L[3 + c])), __dmul_rn(fz, L[6 + c]));
    v[c] = __dadd_rn(dp[c], s);
  }
  const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(v[0], v[0]), __dmul_rn(v[1], v[1])), __dmul_rn(v[2], v[2]));
  return __dsqrt_rn(d2);
}

// FILL = false: count[i];  FILL = true: write the edges of atom i at edge_ptr[i]..
template <bool FILL>
__global__ void __launch_bounds__(128) neighbor_kernel(const double* __restrict__ lattice, const double* __restrict__ pos,
                                                       const long long* __restrict__ node_ptr,
                                                       const long long* __restrict__ crystal_of, long long N, double cutoff,
                                                       int self_interaction, long long* __restrict__ count,
                                                       const long long* __restrict__ edge_ptr,
                                                       long long* __restrict__ edge_src, long long* __restrict__ edge_dst,
                                                       long long* __restrict__ edge_shift, double* __restrict__ edge_vec,
                                                       double* __restrict__ edge_len, int local_ids) {
  const int lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= N) return;
  const long long c = crystal_of[i];
  const long long n0 = node_ptr[c], n1 = node_ptr[c + 1];
  Cell cell;
  make_cell(lattice + 9 * c, cell);
  const double pi[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
  long long total = 0;
  long long base = FILL ? edge_ptr[i] : 0;
  for (long long j = n0; j < n1; ++j) {
    double dp[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) dp[k] = __dsub_rn(pos[3 * j + k], pi[k]);
    // fractional offset of the pair and the image range that can come within the cutoff (superset, with slack)
    int lo[3], cnt[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double f = dp[0] * cell.Li[0 * 3 + k] + dp[1] * cell.Li[1 * 3 + k] + dp[2] * cell.Li[2 * 3 + k];
      const double w = cutoff * cell.inv_h[k] + 1e-6;
      lo[k] = (int)floor(-f - w);
      const int hi = (int)ceil(-f + w);
      cnt[k] = hi - lo[k] + 1;
    }
    const int ncand = cnt[0] * cnt[1] * cnt[2];
    for (int q0 = 0; q0 < ncand; q0 += 32) {
      const int q = q0 + lane;
      bool ok = false;
      int sx = 0, sy = 0, sz = 0;
      double v[3] = {0.0, 0.0, 0.0}, d = 0.0;
      if (q < ncand) {
        sz = lo[2] + q % cnt[2];
        const int r = q / cnt[2];
        sy = lo[1] + r % cnt[1];
        sx = lo[0] + r / cnt[1];
        d = image_dist(cell.L, dp, sx, sy, sz, v);
        ok = d < cutoff && (self_interaction || j != i || sx != 0 || sy != 0 || sz != 0);
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (FILL && ok) {
        const long long e = base + total + __popc(m & ((1u << lane) - 1u));
        edge_src[e] = local_ids ? i - n0 : i;
        edge_dst[e] = local_ids ? j - n0 : j;
        if (edge_shift) {
          edge_shift[3 * e] = sx;
          edge_shift[3 * e + 1] = sy;
          edge_shift[3 * e + 2] = sz;
        }
        if (edge_vec) {
          edge_vec[3 * e] = v[0];
          edge_vec[3 * e + 1] = v[1];
          edge_vec[3 * e + 2] = v[2];
        }
        if (edge_len) edge_len[e] = d;
      }
      total += __popc(m);
    }
  }
  if (!FILL && lane == 0) count[i] = total;
}

// k smallest distances of every atom's (sorted-by-(j,S)) list, ties resolved by list position; one warp per atom.
// out_idx [N,k] = neighbour atom (as stored in edge_dst), out_dist [N,k]; short lists padded with (pad_idx, pad_dist).
__global__ void __launch_bounds__(128) knn_select_kernel(const long long* __restrict__ edge_ptr,
                                                         const long long* __restrict__ edge_dst,
                                                         const double* __restrict__ edge_len, long long N, int k,
                                                         long long pad_idx, double pad_dist, long long* __restrict__ out_idx,
                                                         double* __restrict__ out_dist, long long* __restrict__ out_edge) {
  const int lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= N) return;
  const long long e0 = edge_ptr[i], e1 = edge_ptr[i + 1];
  double last_d = -1.0;
  long long last_e = -1;      // the previous pick: the next one is the smallest (d, e) strictly after it
  for (int r = 0; r < k; ++r) {
    double bd = 0.0;
    long long be = -1;
    for (long long e = e0 + lane; e < e1; e += 32) {
      const double d = edge_len[e];
      const bool after = d > last_d || (d == last_d && e > last_e);
      if (after && (be < 0 || d < bd)) {      // within a lane e ascends, so the first minimum wins ties
        bd = d;
        be = e;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd, o);
      const long long oe = __shfl_xor_sync(0xffffffffu, be, o);
      if (oe >= 0 && (be < 0 || od < bd || (od == bd && oe < be))) {
        bd = od;
        be = oe;
      }
    }
    if (lane == 0) {
      out_idx[i * k + r] = be >= 0 ? edge_dst[be] : pad_idx;
      out_dist[i * k + r] = be >= 0 ? bd : pad_dist;
      if (out_edge) out_edge[i * k + r] = be;
    }
    if (be < 0) {            // list exhausted: pad the rest
      if (lane == 0)
        for (int r2 = r + 1; r2 < k; ++r2) {
          out_idx[i * k + r2] = pad_idx;
          out_dist[i * k + r2] = pad_dist;
          if (out_edge) out_edge[i * k + r2] = -1;
        }
      break;
    }
    last_d = bd;
    last_e = be;
  }
}

// GaussianDistance.expand (mat2graph.py:162-179): out[e, f] = float(exp(-(d_e - mu_f)^2 / var^2)), mu_f = dmin + f * step.
__global__ void gaussian_expand_kernel(const double* __restrict__ dist, long long n, double dmin, double step, int nfilt,
                                       double var2, float* __restrict__ out) {
  const long long total = n * nfilt;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long e = t / nfilt;
    const int f = (int)(t - e * nfilt);
    const double mu = __dadd_rn(dmin, __dmul_rn((double)f, step));
    const double u = __dsub_rn(dist[e], mu);
    const double q = __ddiv_rn(-__dmul_rn(u, u), var2);
    out[t] = (float)exp(q);
  }
}

}  // namespace
}  // namespace dost

using namespace dost;

extern "C" int dost_neighbor_count(const double* lattice, const double* pos, const int64_t* node_ptr,
                                   const int64_t* crystal_of, long long N, double cutoff, int self_interaction,
                                   int64_t* count, dost_stream_t stream) {
  if (N == 0) return DOST_OK;
  DOST_REQUIRE(lattice && pos && node_ptr && crystal_of && count && N > 0 && cutoff > 0, "neighbor_count: bad args");
  neighbor_kernel<false><<<ceil_div(N, 4), 128, 0, (cudaStream_t)stream>>>(
      lattice, pos, (const long long*)node_ptr, (const long long*)crystal_of, N, cutoff, self_interaction,
      (long long*)count, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0);
  return check_launch("neighbor_count");
}

extern "C" int dost_neighbor_fill(const double* lattice, const double* pos, const int64_t* node_ptr,
                                  const int64_t* crystal_of, long long N, double cutoff, int self_interaction,
                                  const int64_t* edge_ptr, int local_ids, int64_t* edge_src, int64_t* edge_dst,
                                  int64_t* edge_shift, double* edge_vec, double* edge_len, dost_stream_t stream) {
  if (N == 0) return DOST_OK;
  DOST_REQUIRE(lattice && pos && node_ptr && crystal_of && edge_ptr && edge_src && edge_dst && N > 0 && cutoff > 0,
               "neighbor_fill: bad args");
  neighbor_kernel<true><<<ceil_div(N, 4), 128, 0, (cudaStream_t)stream>>>(
      lattice, pos, (const long long*)node_ptr, (const long long*)crystal_of, N, cutoff, self_interaction, nullptr,
      (const long long*)edge_ptr, (long long*)edge_src, (long long*)edge_dst, (long long*)edge_shift, edge_vec, edge_len,
      local_ids);
  return check_launch("neighbor_fill");
}

extern "C" int dost_knn_select(const int64_t* edge_ptr, const int64_t* edge_dst, const double* edge_len, long long N, int k,
                               long long pad_idx, double pad_dist, int64_t* out_idx, double* out_dist, int64_t* out_edge,
                               dost_stream_t stream) {
  if (N == 0 || k == 0) return DOST_OK;
  DOST_REQUIRE(edge_ptr && edge_dst && edge_len && out_idx && out_dist && N > 0 && k > 0, "knn_select: bad args");
  knn_select_kernel<<<ceil_div(N, 4), 128, 0, (cudaStream_t)stream>>>(
      (const long long*)edge_ptr, (const long long*)edge_dst, edge_len, N, k, pad_idx, pad_dist, (long long*)out_idx,
      out_dist, (long long*)out_edge);
  return check_launch("knn_select");
}

extern "C" int dost_gaussian_expand(const double* dist, long long n, double dmin, double step, int nfilt, double var,
                                    float* out, dost_stream_t stream) {
  if (n == 0 || nfilt == 0) return DOST_OK;
  DOST_REQUIRE(dist && out && n > 0 && nfilt > 0 && step > 0 && var > 0, "gaussian_expand: bad args");
  const long long total = n * nfilt;
  const int blocks = (int)min64(ceil_div(total, 256), (long long)kNumSMs * 16);
  gaussian_expand_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dist, n, dmin, step, nfilt, var * var, out);
  return check_launch("gaussian_expand");
}
