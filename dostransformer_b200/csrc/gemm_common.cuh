// Device-side GEMM descriptor shared by the FMA-pipe kernels (gemm.cu) and the tcgen05 kernels (gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace dost {

template <typename T>
struct SegDev {
  const T* base;
  long long ld;
  const int* idx;
  int div;
  int kend;    // cumulative end of this segment along k
  int vec_ok;  // 16-byte vector loads allowed
};

template <typename T>
struct GemmDev {
  int M, N, K;
  int a_nseg;
  SegDev<T> a[3];
  long long a_bstride;
  SegDev<T> b;
  long long b_bstride;
  const T* bias;
  int act;
  T act_slope;
  const T* prelu_slope;
  T* out_pre;
  long long ld_pre;
  const T* dact_saved;
  long long ld_dact;
  T dact_slope;
  const T* residual;
  long long ld_res;
  T* out;
  long long ldc;
  long long c_bstride;
  int accumulate;
  int zmode;  // 0: single, 1: batched over blockIdx.z, 2: split-K over blockIdx.z (raw partials to ws)
  int kchunk;
  T* ws;
  int epi_vec;
};

// tcgen05 path (gemm_tc.cu): precision 1 = bf16x3 error-compensated (fp32 parity), 2 = plain bf16 operands.
int launch_gemm_tc(const GemmDev<float>& g, int precision, bool a_mc, bool b_mc, int batch, int split, cudaStream_t st);
bool gemm_tc_supported(const GemmDev<float>& g, bool a_mc, bool b_mc);

}  // namespace dost
