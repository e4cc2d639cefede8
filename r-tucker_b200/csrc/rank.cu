// Filtered ranking by compare-and-count on a dense prediction matrix (HBM-bound).
//
// Replaces filter_predictions (reference src/utils/utils.py:15-22) followed by metrics
// (src/utils/metrics.py:4-8: full descending sort + gather + argmax).  No sort: the rank of the
// target is 1 + #{entries greater} (+ its position among exact ties).  The filter list is
// applied as a sparse correction instead of rewriting P in place.
//
// Algorithmic bytes per launch: 4*B*N (one read of P) + filter list + 12*B of counters.
#include "common.h"

namespace {

constexpr int kThreads = 256;
constexpr int kSegLen = 8192;  // elements of one row handled by one CTA

struct Counts {
  int g, e, eb;
};

__device__ __forceinline__ void count_one(float v, int j, float pt, int t, Counts& c) {
  c.g += (v > pt);
  int eq = (v == pt);
  c.e += eq;
  c.eb += eq & (j < t);
}

__global__ void __launch_bounds__(kThreads)
rank_dense_kernel(const float* __restrict__ P, int64_t ldp, int N,
                  const int32_t* __restrict__ target, const int32_t* __restrict__ flt_off,
                  const int32_t* __restrict__ flt_idx, int32_t* __restrict__ greater,
                  int32_t* __restrict__ equal, int32_t* __restrict__ equal_before) {
  const int b = blockIdx.y;
  const int seg = blockIdx.x;
  const float* row = P + (int64_t)b * ldp;
  const int t = target[b];
  const float pt = __ldg(row + t);
  const int j0 = seg * kSegLen;
  const int j1 = min(N, j0 + kSegLen);
  Counts c{0, 0, 0};

  // split [j0, j1) into scalar head, 16-byte aligned float4 body, scalar tail
  const uintptr_t addr = reinterpret_cast<uintptr_t>(row + j0);
  int head = (int)(((16 - (addr & 15)) & 15) >> 2);
  head = min(head, j1 - j0);
  const int a0 = j0 + head;
  const int nvec = (j1 - a0) >> 2;
  const int tail0 = a0 + (nvec << 2);
  if ((int)threadIdx.x < head) count_one(__ldg(row + j0 + threadIdx.x), j0 + threadIdx.x, pt, t, c);
  if ((int)threadIdx.x < j1 - tail0)
    count_one(__ldg(row + tail0 + threadIdx.x), tail0 + threadIdx.x, pt, t, c);
  const float4* vrow = reinterpret_cast<const float4*>(row + a0);
#pragma unroll 4
  for (int i = threadIdx.x; i < nvec; i += kThreads) {
    const float4 v = __ldcs(vrow + i);
    const int j = a0 + (i << 2);
    count_one(v.x, j, pt, t, c);
    count_one(v.y, j + 1, pt, t, c);
    count_one(v.z, j + 2, pt, t, c);
    count_one(v.w, j + 3, pt, t, c);
  }
  // the target itself was counted as "equal" (unless NaN); remove it
  if (threadIdx.x == 0 && t >= j0 && t < j1) c.e -= (pt == pt);

  // sparse filter correction, done once per row by segment 0
  if (seg == 0) {
    const int f0 = flt_off[b], f1 = flt_off[b + 1];
    const int zg = (0.0f > pt), ze = (0.0f == pt);
    for (int i = f0 + threadIdx.x; i < f1; i += kThreads) {
      const int f = flt_idx[i];
      if (f == t) continue;
      const float v = __ldg(row + f);
      const int eq = (v == pt);
      c.g += zg - (v > pt);
      c.e += ze - eq;
      c.eb += (f < t) ? (ze - eq) : 0;
    }
  }

  __shared__ int sg[kThreads / 32], se[kThreads / 32], sb[kThreads / 32];
  const int g = rt::warp_sum(c.g), e = rt::warp_sum(c.e), eb = rt::warp_sum(c.eb);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sg[w] = g; se[w] = e; sb[w] = eb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tg = 0, te = 0, tb = 0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) { tg += sg[i]; te += se[i]; tb += sb[i]; }
    // integer atomics: order-independent, deterministic
    if (tg) atomicAdd(greater + b, tg);
    if (te) atomicAdd(equal + b, te);
    if (tb) atomicAdd(equal_before + b, tb);
  }
}

// In-place dense filter, exactly reference src/utils/utils.py:18-21: keep the score at the filter
// column, zero scores and targets wherever targets == 1, restore the kept score, set targets to
// one-hot at the filter column.
__global__ void __launch_bounds__(kThreads)
filter_dense_kernel(float* __restrict__ P, int64_t ldp, float* __restrict__ T, int64_t ldt, int N,
                    const int32_t* __restrict__ filter_col) {
  const int b = blockIdx.y;
  const int f = filter_col[b];
  float* prow = P + (int64_t)b * ldp;
  float* trow = T + (int64_t)b * ldt;
  for (int j = blockIdx.x * kThreads + threadIdx.x; j < N; j += gridDim.x * kThreads) {
    const float t = trow[j];
    if (j == f) { trow[j] = 1.0f; continue; }   // prediction kept as is
    if (t == 1.0f) { prow[j] = 0.0f; trow[j] = 0.0f; }
  }
}

}  // namespace

extern "C" int rt_filter_dense(float* P, int64_t ldp, float* T, int64_t ldt, int B, int N,
                               const int32_t* filter_col, void* stream) {
  RT_REQUIRE(B >= 0 && N > 0 && ldp >= N && ldt >= N && B <= 65535, "rt_filter_dense: bad shape");
  if (B == 0) return 0;
  dim3 grid(rt::cdiv(N, kThreads * 8), B);
  filter_dense_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(P, ldp, T, ldt, N, filter_col);
  RT_LAUNCH_CHECK();
  return 0;
}

extern "C" int rt_rank_filtered(const float* P, int64_t ldp, int B, int N, const int32_t* target,
                                const int32_t* flt_off, const int32_t* flt_idx, int32_t* greater,
                                int32_t* equal, int32_t* equal_before, void* stream) {
  RT_REQUIRE(B >= 0 && N > 0 && ldp >= N, "rt_rank_filtered: bad shape B=%d N=%d ldp=%lld", B, N,
             (long long)ldp);
  RT_REQUIRE(B <= 65535, "rt_rank_filtered: B=%d exceeds 65535 rows per call", B);
  cudaStream_t s = (cudaStream_t)stream;
  if (B == 0) return 0;
  RT_CHECK_CUDA(cudaMemsetAsync(greater, 0, sizeof(int32_t) * B, s));
  RT_CHECK_CUDA(cudaMemsetAsync(equal, 0, sizeof(int32_t) * B, s));
  RT_CHECK_CUDA(cudaMemsetAsync(equal_before, 0, sizeof(int32_t) * B, s));
  dim3 grid(rt::cdiv(N, kSegLen), B);
  rank_dense_kernel<<<grid, kThreads, 0, s>>>(P, ldp, N, target, flt_off, flt_idx, greater, equal,
                                              equal_before);
  RT_LAUNCH_CHECK();
  return 0;
}
