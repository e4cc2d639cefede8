// Batched symmetric eigen-decomposition by two-sided BLOCK JACOBI (fp64), one cooperative launch.
//
// Used for the HOSVD / SF-HOSVD inside the retraction (reference: Tucker.round / SFTucker.round
// of tucker_riemopt, called at src/model/asymmetric/optim.py:108 and symmetric/optim.py:55,102),
// where the left singular vectors of each unfolding of the rank-2r core are the eigenvectors of
// its (2r_i x 2r_i) Gram matrix.  The reference calls cuSOLVER gesvd on the unfoldings.
//
// Scheme: blocks of 8 indices; a round-robin tournament pairs the blocks; per round
//   phase A: one warp per block pair diagonalises its 16x16 sub-matrix (scalar cyclic Jacobi in
//            shared memory) and publishes the 16x16 rotation Q;
//   phase B: every 16x16 tile (pair k rows, pair l cols) becomes Q_k^T A_kl Q_l, and V <- V Q.
// grid.sync() separates the phases.  N-independent, latency-bound: reported separately from the
// HBM / tensor rooflines (SURVEY.md section 8d "N-independent serial part").
#include "common.h"
#include <cooperative_groups.h>
#include <math.h>

namespace cg = cooperative_groups;

namespace rt {

constexpr int EB = 8;          // block size
constexpr int EP = 2 * EB;     // pair size (16)
constexpr int ELD = EP + 1;    // padded smem leading dim
constexpr int kMaxProblems = 4;
constexpr int kEigThreads = 512;
constexpr int kEigWarps = kEigThreads / 32;
constexpr int kMaxSweeps = 16;

struct EigProblem {
  const double* A_in;  // [n][n] dense symmetric
  double* w;           // [n] descending
  double* V_out;       // [n][n] columns = eigenvectors
  double* Ap;          // [np][np] padded work copy
  double* Vp;          // [np][np]
  double* J;           // [npairs][EP*EP]
  int* skip;           // [npairs]
  double* scal;        // [0]=norm2, [1+sweep]=off2 of that sweep
  int n, np, nb, npairs;
};

struct EigBatch {
  EigProblem p[kMaxProblems];
  int count;
};

// round-robin tournament (circle method) on nb (even) players: pair k of round t
__device__ __forceinline__ void rr_pair(int nb, int t, int k, int& bi, int& bj) {
  const int m = nb - 1;
  if (k == 0) { bi = m; bj = t % m; }
  else { bi = (t + k) % m; bj = (t - k + m) % m; }
  if (bi > bj) { const int x = bi; bi = bj; bj = x; }
}

__device__ __forceinline__ int pair_index(int bi, int bj, int i) {  // i in [0,16)
  return (i < EB) ? bi * EB + i : bj * EB + (i - EB);
}

// One warp diagonalises the symmetric 16x16 matrix S (smem, ld ELD); Q accumulates rotations.
__device__ void warp_jacobi16(double* S, double* Q, double* cs, int lane) {
  for (int e = lane; e < EP * EP; e += 32) Q[(e / EP) * ELD + (e % EP)] = (e / EP == e % EP) ? 1.0 : 0.0;
  __syncwarp();
  for (int sweep = 0; sweep < 10; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int e = lane; e < EP * EP; e += 32) {
      const int i = e / EP, j = e % EP;
      const double v = S[i * ELD + j];
      if (i == j) dg += v * v; else off += v * v;
    }
    off = rt::warp_sum(off);
    dg = rt::warp_sum(dg);
    if (off <= 1e-30 * dg || off == 0.0) break;
    for (int t = 0; t < EP - 1; ++t) {
      if (lane < EP / 2) {
        int p, q;
        rr_pair(EP, t, lane, p, q);
        const double apq = S[p * ELD + q];
        double c = 1.0, s = 0.0;
        if (apq != 0.0) {
          const double tau = (S[q * ELD + q] - S[p * ELD + p]) / (2.0 * apq);
          const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = 1.0 / sqrt(1.0 + tt * tt);
          s = tt * c;
        }
        cs[lane * 4 + 0] = c;
        cs[lane * 4 + 1] = s;
        cs[lane * 4 + 2] = (double)p;
        cs[lane * 4 + 3] = (double)q;
      }
      __syncwarp();
      // two-sided update, 64 independent 2x2 blocks, 2 per lane
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int blk = lane + 32 * h;
        const int k = blk >> 3, l = blk & 7;
        const double ck = cs[k * 4], sk = cs[k * 4 + 1];
        const int p = (int)cs[k * 4 + 2], q = (int)cs[k * 4 + 3];
        const double cl = cs[l * 4], sl = cs[l * 4 + 1];
        const int u = (int)cs[l * 4 + 2], v = (int)cs[l * 4 + 3];
        const double m00 = S[p * ELD + u], m01 = S[p * ELD + v];
        const double m10 = S[q * ELD + u], m11 = S[q * ELD + v];
        const double t00 = ck * m00 - sk * m10, t01 = ck * m01 - sk * m11;
        const double t10 = sk * m00 + ck * m10, t11 = sk * m01 + ck * m11;
        double n00 = cl * t00 - sl * t01, n01 = sl * t00 + cl * t01;
        double n10 = cl * t10 - sl * t11, n11 = sl * t10 + cl * t11;
        if (k == l) { n01 = 0.0; n10 = 0.0; }
        S[p * ELD + u] = n00; S[p * ELD + v] = n01;
        S[q * ELD + u] = n10; S[q * ELD + v] = n11;
      }
      // Q <- Q R : 16 rows x 8 pairs, 4 per lane
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const int it = lane + 32 * h;
        const int row = it >> 3, l = it & 7;
        const double cl = cs[l * 4], sl = cs[l * 4 + 1];
        const int u = (int)cs[l * 4 + 2], v = (int)cs[l * 4 + 3];
        const double qu = Q[row * ELD + u], qv = Q[row * ELD + v];
        Q[row * ELD + u] = cl * qu - sl * qv;
        Q[row * ELD + v] = sl * qu + cl * qv;
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(kEigThreads, 1)
eig_block_jacobi_kernel(EigBatch batch) {
  extern __shared__ __align__(16) double esm[];
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * kEigWarps + warp;
  const int nwarps = gridDim.x * kEigWarps;
  const int gtid = blockIdx.x * kEigThreads + threadIdx.x;
  const int nthreads = gridDim.x * kEigThreads;
  double* T = esm + warp * (3 * EP * ELD + 32);  // per-warp: T, Qk, Ql, cs
  double* Qk = T + EP * ELD;
  double* Ql = Qk + EP * ELD;
  double* cs = Ql + EP * ELD;

  // ---- phase 0: padded copies, V = I, norms ----
  for (int pi = 0; pi < batch.count; ++pi) {
    const EigProblem& P = batch.p[pi];
    double loc = 0.0;
    for (int e = gtid; e < P.np * P.np; e += nthreads) {
      const int i = e / P.np, j = e - i * P.np;
      double v = 0.0;
      if (i < P.n && j < P.n) v = 0.5 * (P.A_in[(int64_t)i * P.n + j] + P.A_in[(int64_t)j * P.n + i]);
      P.Ap[e] = v;
      P.Vp[e] = (i == j) ? 1.0 : 0.0;
      loc += v * v;
    }
    loc = rt::warp_sum(loc);
    if (lane == 0 && loc != 0.0) atomicAdd(&P.scal[0], loc);
  }
  grid.sync();

  int max_rounds = 0;
  for (int pi = 0; pi < batch.count; ++pi) max_rounds = max(max_rounds, batch.p[pi].nb - 1);
  bool done[kMaxProblems];
  for (int pi = 0; pi < kMaxProblems; ++pi) done[pi] = (pi >= batch.count) || (batch.p[pi].nb < 2);

  for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
    bool all_done = true;
    for (int pi = 0; pi < batch.count; ++pi) all_done = all_done && done[pi];
    if (all_done) break;
    for (int round = 0; round < max_rounds; ++round) {
      // ---- phase A: diagonalise pair sub-matrices ----
      int base = 0;
      for (int pi = 0; pi < batch.count; ++pi) {
        const EigProblem& P = batch.p[pi];
        if (done[pi] || round >= P.nb - 1) continue;
        for (int item = gwarp - base; item < P.npairs; item += nwarps) {
          if (item < 0) continue;
          int bi, bj;
          rr_pair(P.nb, round, item, bi, bj);
          double off = 0.0;
          for (int e = lane; e < EP * EP; e += 32) {
            const int i = e / EP, j = e % EP;
            const double v = P.Ap[(int64_t)pair_index(bi, bj, i) * P.np + pair_index(bi, bj, j)];
            T[i * ELD + j] = v;
            if (i != j) off += v * v;  // whole pair sub-matrix: the diagonal blocks must end up diagonal too
          }
          off = rt::warp_sum(off);
          __syncwarp();
          const bool skip = (off <= 1e-34 * P.scal[0]);
          if (!skip) {
            warp_jacobi16(T, Qk, cs, lane);
            for (int e = lane; e < EP * EP; e += 32)
              P.J[(int64_t)item * EP * EP + e] = Qk[(e / EP) * ELD + (e % EP)];
            // the diagonal tile is now diag(T): write it here, phase B skips k == l
            for (int e = lane; e < EP * EP; e += 32) {
              const int i = e / EP, j = e % EP;
              P.Ap[(int64_t)pair_index(bi, bj, i) * P.np + pair_index(bi, bj, j)] =
                  (i == j) ? T[i * ELD + i] : 0.0;
            }
          }
          if (lane == 0) {
            P.skip[item] = skip ? 1 : 0;
            if (off != 0.0) atomicAdd(&P.scal[1 + sweep], off);
          }
          __syncwarp();
        }
        base = (base + P.npairs) % nwarps;
      }
      grid.sync();
      // ---- phase B: A_kl <- Q_k^T A_kl Q_l (k != l),  V[:, l] <- V[:, l] Q_l ----
      base = 0;
      for (int pi = 0; pi < batch.count; ++pi) {
        const EigProblem& P = batch.p[pi];
        if (done[pi] || round >= P.nb - 1) continue;
        const int nA = P.npairs * P.npairs;
        const int nV = (P.np / EP) * P.npairs;
        for (int item = gwarp - base; item < nA + nV; item += nwarps) {
          if (item < 0) continue;
          const bool isV = item >= nA;
          int k, l;
          if (isV) { k = (item - nA) / P.npairs; l = (item - nA) % P.npairs; }
          else { k = item / P.npairs; l = item % P.npairs; }
          if (!isV && k == l) continue;
          const bool sk = isV ? true : (P.skip[k] != 0);
          const bool sl = P.skip[l] != 0;
          if (sk && sl) continue;
          int bik = 0, bjk = 0, bil, bjl;
          if (!isV) rr_pair(P.nb, round, k, bik, bjk);
          rr_pair(P.nb, round, l, bil, bjl);
          double* M = isV ? P.Vp : P.Ap;
          for (int e = lane; e < EP * EP; e += 32) {
            const int i = e / EP, j = e % EP;
            const int gi = isV ? k * EP + i : pair_index(bik, bjk, i);
            T[i * ELD + j] = M[(int64_t)gi * P.np + pair_index(bil, bjl, j)];
            Ql[i * ELD + j] = sl ? (i == j ? 1.0 : 0.0) : P.J[(int64_t)l * EP * EP + e];
            if (!isV) Qk[i * ELD + j] = sk ? (i == j ? 1.0 : 0.0) : P.J[(int64_t)k * EP * EP + e];
          }
          __syncwarp();
          // X = T Ql : lane -> row i = lane/2, cols (lane&1)*8 .. +7
          const int i = lane >> 1, j0 = (lane & 1) * 8;
          double x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = 0.0;
          for (int m = 0; m < EP; ++m) {
            const double tv = T[i * ELD + m];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = fma(tv, Ql[m * ELD + j0 + j], x[j]);
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) T[i * ELD + j0 + j] = x[j];
          __syncwarp();
          if (!isV) {  // Y = Qk^T X
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = 0.0;
            for (int m = 0; m < EP; ++m) {
              const double qv = Qk[m * ELD + i];
#pragma unroll
              for (int j = 0; j < 8; ++j) x[j] = fma(qv, T[m * ELD + j0 + j], x[j]);
            }
          }
          const int gi = isV ? k * EP + i : pair_index(bik, bjk, i);
#pragma unroll
          for (int j = 0; j < 8; ++j) M[(int64_t)gi * P.np + pair_index(bil, bjl, j0 + j)] = x[j];
          __syncwarp();
        }
        base = (base + nA + nV) % nwarps;
      }
      grid.sync();
    }
    // convergence: off-diagonal mass seen during this sweep (uniform decision: same memory, after sync)
    for (int pi = 0; pi < batch.count; ++pi) {
      if (done[pi]) continue;
      const EigProblem& P = batch.p[pi];
      const double off2 = P.scal[1 + sweep], n2 = P.scal[0];
      if (off2 <= 1e-28 * n2) done[pi] = true;
    }
  }

  // ---- final: sort eigenvalues descending, emit w and V ----
  for (int pi = 0; pi < batch.count; ++pi) {
    const EigProblem& P = batch.p[pi];
    for (int i = gtid; i < P.n; i += nthreads) {
      const double di = P.Ap[(int64_t)i * P.np + i];
      int rank = 0;
      for (int j = 0; j < P.n; ++j) {
        const double dj = P.Ap[(int64_t)j * P.np + j];
        rank += (dj > di) || (dj == di && j < i);
      }
      P.w[rank] = di;
      P.scal[1 + kMaxSweeps + i] = (double)rank;  // reuse scal tail as the permutation
    }
  }
  grid.sync();
  for (int pi = 0; pi < batch.count; ++pi) {
    const EigProblem& P = batch.p[pi];
    for (int e = gtid; e < P.n * P.n; e += nthreads) {
      const int i = e / P.n, j = e - i * P.n;
      const int rank = (int)P.scal[1 + kMaxSweeps + j];
      P.V_out[(int64_t)i * P.n + rank] = P.Vp[(int64_t)i * P.np + j];
    }
  }
}

struct EigLayout {
  int np, nb, npairs;
  size_t off_Ap, off_Vp, off_J, off_skip, off_scal, total;
};

EigLayout eig_layout(int n) {
  EigLayout L;
  L.nb = cdiv(n, EB);
  if (L.nb & 1) L.nb += 1;
  if (L.nb < 2) L.nb = 2;
  L.np = L.nb * EB;
  L.npairs = L.nb / 2;
  size_t o = 0;
  L.off_Ap = o; o += align_up(sizeof(double) * L.np * L.np, 256);
  L.off_Vp = o; o += align_up(sizeof(double) * L.np * L.np, 256);
  L.off_J = o; o += align_up(sizeof(double) * L.npairs * EP * EP, 256);
  L.off_skip = o; o += align_up(sizeof(int) * L.npairs, 256);
  L.off_scal = o; o += align_up(sizeof(double) * (1 + kMaxSweeps + L.np), 256);
  L.total = o;
  return L;
}

size_t eig_ws_bytes(int n) { return eig_layout(n).total; }

// Solve `count` (<= 4) independent problems in one cooperative launch.
int eig_batch(int count, const double* const* A, const int* n, double* const* w, double* const* V,
              void* const* ws, cudaStream_t s) {
  RT_REQUIRE(count >= 1 && count <= kMaxProblems, "eig_batch: count=%d out of range", count);
  EigBatch b{};
  b.count = count;
  int total_items = 0;
  for (int i = 0; i < count; ++i) {
    RT_REQUIRE(n[i] >= 1 && n[i] <= 1024, "eig_batch: n=%d out of range", n[i]);
    EigLayout L = eig_layout(n[i]);
    char* base = (char*)ws[i];
    EigProblem& P = b.p[i];
    P.A_in = A[i]; P.w = w[i]; P.V_out = V[i];
    P.Ap = (double*)(base + L.off_Ap);
    P.Vp = (double*)(base + L.off_Vp);
    P.J = (double*)(base + L.off_J);
    P.skip = (int*)(base + L.off_skip);
    P.scal = (double*)(base + L.off_scal);
    P.n = n[i]; P.np = L.np; P.nb = L.nb; P.npairs = L.npairs;
    RT_CHECK_CUDA(cudaMemsetAsync(P.scal, 0, sizeof(double) * (1 + kMaxSweeps + L.np), s));
    total_items += L.npairs * L.npairs + (L.np / EP) * L.npairs;
  }
  const size_t smem = (size_t)kEigWarps * (3 * EP * ELD + 32) * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    RT_CHECK_CUDA(cudaFuncSetAttribute(eig_block_jacobi_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int grid = cdiv(total_items, kEigWarps);
  const int max_grid = sm_count();  // 1 CTA / SM (launch bounds) => co-resident
  if (grid > max_grid) grid = max_grid;
  if (grid < 1) grid = 1;
  void* args[] = {(void*)&b};
  RT_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)eig_block_jacobi_kernel, dim3(grid),
                                            dim3(kEigThreads), args, smem, s));
  return 0;
}

}  // namespace rt

extern "C" size_t rt_eigh_ws_bytes(int n) { return n > 0 ? rt::eig_ws_bytes(n) : 0; }

extern "C" int rt_eigh(double* A, int n, double* w, double* V, void* ws, void* stream) {
  RT_REQUIRE(ws != nullptr, "rt_eigh: workspace is NULL");
  const double* Ain[1] = {A};
  double* wo[1] = {w};
  double* Vo[1] = {V};
  void* wss[1] = {ws};
  return rt::eig_batch(1, Ain, &n, wo, Vo, wss, (cudaStream_t)stream);
}
