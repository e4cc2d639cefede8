// Device-resident epoch: the whole (s, r) -> objects vocabulary of a split
// lives in HBM as CSR; a batch is assembled ON THE DEVICE from a device-side permutation, so an epoch needs
// no host work, no host->device copy and no synchronisation per step.
//
// Replaces, for the fused path, KG_dataset.__getitem__ + the default collate + DataLoader(shuffle=True)
// (reference src/data/Dataset.py:42-53, train.py:75-76,226-236), which build and ship a dense B x n_ent
// target matrix (84 MB per WN18RR batch) every step.
#include "common.h"

namespace {

constexpr int kQPerCta = 8;   // one warp per query

__global__ void __launch_bounds__(256)
epoch_batch_kernel(const int64_t* __restrict__ perm, int lo, int B, const int* __restrict__ feat_all, int fc,
                   const int* __restrict__ off_all, const int* __restrict__ idx_all, int* __restrict__ feat_out,
                   int* __restrict__ off_out, int* __restrict__ idx_out, int cap) {
  __shared__ int red[8];
  __shared__ int cnt8[kQPerCta];
  __shared__ int base_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * kQPerCta;
  // number of targets of all queries before this CTA's first one (B is a few hundred: a CTA-wide sum is cheaper
  // than a second kernel for the scan)
  int part = 0;
  for (int j = tid; j < q0; j += 256) {
    const int64_t it = perm[lo + j];
    part += off_all[it + 1] - off_all[it];
  }
  part = rt::warp_sum(part);
  if (lane == 0) red[warp] = part;
  if (tid < kQPerCta) {
    int c = 0;
    if (q0 + tid < B) { const int64_t it = perm[lo + q0 + tid]; c = off_all[it + 1] - off_all[it]; }
    cnt8[tid] = c;
  }
  __syncthreads();
  if (tid == 0) { int b = 0; for (int w = 0; w < 8; ++w) b += red[w]; base_s = b; }
  __syncthreads();
  const int q = q0 + warp;
  if (q >= B) return;
  int my = base_s;
  for (int u = 0; u < warp; ++u) my += cnt8[u];
  const int64_t it = perm[lo + q];
  const int src = off_all[it];
  const int n = cnt8[warp];
  for (int e = lane; e < n; e += 32)
    if (my + e < cap) idx_out[my + e] = idx_all[src + e];
  if (lane < fc) feat_out[q * fc + lane] = feat_all[it * fc + lane];
  if (lane == 0) {
    off_out[q] = min(my, cap);
    if (q == B - 1) off_out[B] = min(my + n, cap);
  }
}

}  // namespace

extern "C" int rt_epoch_batch(const int64_t* perm, int lo, int B, const int* feat_all, int feat_cols, const int* off_all,
                              const int* idx_all, int* feat_out, int* off_out, int* idx_out, int cap, void* stream) {
  RT_REQUIRE(perm && feat_all && off_all && idx_all && feat_out && off_out && idx_out, "rt_epoch_batch: NULL argument");
  RT_REQUIRE(B > 0 && lo >= 0 && feat_cols >= 1 && feat_cols <= 32 && cap > 0, "rt_epoch_batch: bad arguments");
  epoch_batch_kernel<<<rt::cdiv(B, kQPerCta), 256, 0, (cudaStream_t)stream>>>(perm, lo, B, feat_all, feat_cols, off_all,
                                                                              idx_all, feat_out, off_out, idx_out, cap);
  RT_LAUNCH_CHECK();
  return 0;
}
