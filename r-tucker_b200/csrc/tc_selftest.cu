// Known-answer self test of the tcgen05 building blocks (descriptors, interleaved staging format,
// K-major / MN-major operands, TMEM readback).  D[128, N] = op(A) op(B)^T with TF32 inputs.
//   a_mn = 0: A given as [128][K] (K contiguous)   -> K-major operand
//   a_mn = 1: A given as [K][128] (M contiguous)   -> MN-major operand
//   b_mn likewise for B ([N][K] or [K][N]).
// flags bit0/bit1: swap the roles of LBO/SBO for A/B (used once to pin the descriptor convention).
#include "common.h"
#include "tc.cuh"

namespace {
using namespace rt::tc;

__global__ void __launch_bounds__(128, 1)
tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K,
                   int a_mn, int b_mn, int flags) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int M = 128;
  // staged arrays: X[rows][cols]; for the K-major form rows = M (or N), cols = K; for the MN-major form rows = K
  const int a_rows = a_mn ? K : M, a_cols = a_mn ? M : K;
  const int b_rows = b_mn ? K : N, b_cols = b_mn ? N : K;
  const uint32_t a_CS = (uint32_t)a_rows * 16u, b_CS = (uint32_t)b_rows * 16u, RS = 128u;
  unsigned char* sA = smem;
  unsigned char* sB = smem + (size_t)a_rows * a_cols * 4;
  for (int e = tid; e < 220 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0u;
  __syncthreads();
  for (int e = tid; e < a_rows * a_cols; e += 128) {
    const int r = e / a_cols, c = e % a_cols;
    *reinterpret_cast<uint32_t*>(sA + il_offset(r, c, a_CS, RS)) = to_tf32(A[e]);
  }
  for (int e = tid; e < b_rows * b_cols; e += 128) {
    const int r = e / b_cols, c = e % b_cols;
    *reinterpret_cast<uint32_t*>(sB + il_offset(r, c, b_CS, RS)) = to_tf32(B[e]);
  }
  if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<256>(&tmem_base_slot);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(M, N, a_mn != 0, b_mn != 0);
    for (int ks = 0; ks < K / 8; ++ks) {
      uint32_t a_addr, a_lbo, a_sbo, b_addr, b_lbo, b_sbo;
      if (!a_mn) { a_addr = smem_u32(sA) + ks * 2 * a_CS; a_lbo = a_CS; a_sbo = RS; }
      else       { a_addr = smem_u32(sA) + ks * RS;       a_lbo = RS;   a_sbo = a_CS; }
      if (!b_mn) { b_addr = smem_u32(sB) + ks * 2 * b_CS; b_lbo = b_CS; b_sbo = RS; }
      else       { b_addr = smem_u32(sB) + ks * RS;       b_lbo = RS;   b_sbo = b_CS; }
      if (flags & 1) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; }
      if (flags & 2) { uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
      mma_tf32(tmem, make_desc(a_addr, a_lbo, a_sbo), make_desc(b_addr, b_lbo, b_sbo), idesc, ks > 0);
    }
    mma_commit(&mbar);
  }
  mbar_wait(&mbar, 0);
  fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
    const int row = warp * 32 + (tid & 31);
    for (int j = 0; j < 16; ++j)
      if (c0 + j < N) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace

extern "C" int rt_tc_selftest(const float* A, const float* B, float* D, int N, int K, int a_mn, int b_mn,
                              int flags, void* stream) {
  RT_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 8 == 0 && K > 0, "rt_tc_selftest: bad shape N=%d K=%d", N, K);
  RT_REQUIRE((size_t)(128 + N) * K * 4 <= 200 * 1024, "rt_tc_selftest: operands do not fit in shared memory");
  const size_t smem = 220 * 1024;   // generous: descriptor experiments must never address outside the window
  RT_CHECK_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, a_mn, b_mn, flags);
  RT_LAUNCH_CHECK();
  return 0;
}
