// Known-answer self test of the fp16 tcgen05 building blocks used by score_bce_v3.cu:
//   * kind::f16 MMA with fp16 operands in the interleaved (SWIZZLE_NONE) format of tc.cuh,
//     K-major and MN-major views of the SAME staged buffer;
//   * shared -> global bulk copies, plain and with fp32 add-reduction (cp.reduce.async.bulk).
// D[128, N] = op(A) op(B)^T.   a_mn = 0: A given as [128][K];  a_mn = 1: A given as [K][128];  b likewise.
// flags bit0 / bit1 swap LBO and SBO of A / B (pins the descriptor convention once on hardware).
#include "common.h"
#include "tc.cuh"
#include <cuda_fp16.h>

namespace {
using namespace rt::tc;

__global__ void __launch_bounds__(128, 1)
tc_selftest16_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K,
                     int a_mn, int b_mn, int flags) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int M = 128;
  const int a_rows = a_mn ? K : M, a_cols = a_mn ? M : K;
  const int b_rows = b_mn ? K : N, b_cols = b_mn ? N : K;
  const uint32_t a_CS = (uint32_t)a_rows * 16u, b_CS = (uint32_t)b_rows * 16u, RS = 128u;
  unsigned char* sA = smem;
  unsigned char* sB = smem + (size_t)a_rows * a_cols * 2;
  for (int e = tid; e < 200 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0u;
  __syncthreads();
  for (int e = tid; e < a_rows * a_cols; e += 128) {
    const int r = e / a_cols, c = e % a_cols;
    *reinterpret_cast<__half*>(sA + il16_offset(r, c, a_CS)) = __float2half_rn(A[e]);
  }
  for (int e = tid; e < b_rows * b_cols; e += 128) {
    const int r = e / b_cols, c = e % b_cols;
    *reinterpret_cast<__half*>(sB + il16_offset(r, c, b_CS)) = __float2half_rn(B[e]);
  }
  if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<256>(&tmem_base_slot);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(M, N, a_mn != 0, b_mn != 0);
    for (int ks = 0; ks < K / 16; ++ks) {
      uint32_t a_addr, a_lbo, a_sbo, b_addr, b_lbo, b_sbo;
      if (!a_mn) { a_addr = smem_u32(sA) + ks * 2 * a_CS; a_lbo = a_CS; a_sbo = RS; }
      else       { a_addr = smem_u32(sA) + ks * 2 * RS;   a_lbo = RS;   a_sbo = a_CS; }
      if (!b_mn) { b_addr = smem_u32(sB) + ks * 2 * b_CS; b_lbo = b_CS; b_sbo = RS; }
      else       { b_addr = smem_u32(sB) + ks * 2 * RS;   b_lbo = RS;   b_sbo = b_CS; }
      if (flags & 1) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; }
      if (flags & 2) { uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
      mma_f16(tmem, make_desc(a_addr, a_lbo, a_sbo), make_desc(b_addr, b_lbo, b_sbo), idesc, ks > 0);
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

// out[0:n] = a (plain bulk store) then += b (bulk add-reduction), n floats, n * 4 a multiple of 16
__global__ void bulk_reduce_selftest_kernel(const float* __restrict__ a, const float* __restrict__ b, float* out, int n) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* sa = reinterpret_cast<float*>(smem);
  float* sb = sa + n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { sa[i] = a[i]; sb[i] = b[i]; }
  fence_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    bulk_s2g(out, sa, (uint32_t)n * 4u);
    bulk_commit();
    bulk_wait<0>();
    bulk_s2g_add_f32(out, sb, (uint32_t)n * 4u);
    bulk_commit();
    bulk_wait<0>();
  }
}

}  // namespace

extern "C" int rt_tc_selftest16(const float* A, const float* B, float* D, int N, int K, int a_mn, int b_mn,
                                int flags, void* stream) {
  RT_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K > 0, "rt_tc_selftest16: bad shape N=%d K=%d", N, K);
  RT_REQUIRE((size_t)(128 + N) * K * 2 <= 180 * 1024, "rt_tc_selftest16: operands do not fit in shared memory");
  const size_t smem = 200 * 1024;
  RT_CHECK_CUDA(cudaFuncSetAttribute(tc_selftest16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest16_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, a_mn, b_mn, flags);
  RT_LAUNCH_CHECK();
  return 0;
}

extern "C" int rt_bulk_reduce_selftest(const float* a, const float* b, float* out, int n, void* stream) {
  RT_REQUIRE(n > 0 && n % 4 == 0 && n <= 8192, "rt_bulk_reduce_selftest: n=%d must be a multiple of 4, <= 8192", n);
  bulk_reduce_selftest_kernel<<<1, 128, (size_t)n * 8, (cudaStream_t)stream>>>(a, b, out, n);
  RT_LAUNCH_CHECK();
  return 0;
}

// ---- MMA throughput probe: `reps` back-to-back kind::f16 MMAs (M = 128, N, K = 16 each, `ksteps` distinct
//      k-steps cycled) on operands resident in shared memory; returns the elapsed SM cycles of the issue..commit
//      window.  a_mn / b_mn choose the descriptor view; the contents are irrelevant (zeros). ----
namespace {
__global__ void __launch_bounds__(128, 1)
mma_probe_kernel(int N, int ksteps, int a_mn, int b_mn, int reps, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 200 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0u;
  if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<256>(&tmem_base_slot);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_slot;
  if (tid == 0) {
    const uint32_t a_CS = 128 * 16, b_CS = (uint32_t)N * 16u, RS = 128u;
    const uint32_t sA = smem_u32(smem), sB = sA + 64 * 1024;
    const uint32_t idesc = make_idesc_f16(128, N, a_mn != 0, b_mn != 0);
    // descriptors are built once; a k-step only adds to the 14-bit address field (the uniform datapath that feeds
    // tcgen05.mma has a long latency per dependent operation: rebuilding a descriptor per MMA costs ~200 cycles)
    const uint64_t da0 = a_mn ? make_desc(sA, RS, a_CS) : make_desc(sA, a_CS, RS);
    const uint64_t db0 = b_mn ? make_desc(sB, RS, b_CS) : make_desc(sB, b_CS, RS);
    const uint32_t sa = (a_mn ? 2 * RS : 2 * a_CS) >> 4, sb = (b_mn ? 2 * RS : 2 * b_CS) >> 4;
    (void)ksteps;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < reps; i += 8) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) mma_f16(tmem, da0 + (uint64_t)(ks * sa), db0 + (uint64_t)(ks * sb), idesc, (i | ks) > 0);
    }
    const long long t1 = clock64();
    mma_commit(&mbar);
    mbar_wait(&mbar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}
}  // namespace

extern "C" int rt_mma_probe(int N, int ksteps, int a_mn, int b_mn, int reps, long long* out_dev, void* stream) {
  RT_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && ksteps >= 1 && ksteps <= 8 && reps >= 1, "rt_mma_probe: bad arguments");
  const size_t smem = 200 * 1024;
  RT_CHECK_CUDA(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(N, ksteps, a_mn, b_mn, reps, out_dev);
  RT_LAUNCH_CHECK();
  return 0;
}
