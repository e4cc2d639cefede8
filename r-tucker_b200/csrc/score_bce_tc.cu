// (b) tcgen05 / TMEM variant of the fused 1-N score + sigmoid + label-smoothed BCE + backward kernel.
//
// Same contract as the FFMA kernel in score_bce.cu (reference: src/model/asymmetric/R_TuckER.py:47-48,
// nn.BCELoss at train.py:79,136 on the targets of src/data/Dataset.py:43-52, and their backward), but the
// three B x N x r2 contractions run on the 5th-generation tensor cores in TF32 with fp32 accumulation
// in tensor memory:
//      GEMM1  Z[b,n]   = sum_k Q[b,k]  O[n,k]        D in TMEM cols [256,384)
//      GEMM2  H[b,c]  += sum_n G[b,n]  O[n,c]        D in TMEM cols [256,256+R2P)  (re-uses Z's columns)
//      GEMM3  dO[n,c] += sum_b G[b,n]  Q'[b,c]       D in TMEM cols [0,R2P), accumulates over the batch chunks
// TF32 tcgen05.mma only accepts K-major shared-memory operands (MN-major returns zeros: measured,
// tools/tc_probe2.py), so the operands whose contraction index is not their contiguous index are
// staged TRANSPOSED (O^T, G^T, Q'^T); all staging uses the interleaved core-matrix format of tc.cuh.
// Looser tolerance than the FFMA path: TF32 inputs (10-bit mantissa, round-to-nearest at staging).
//
// Work per launch 6*B*N*r2 flop; algorithmic bytes 8*N*r2.
#include "common.h"
#include "tc.cuh"
#include <math.h>

namespace {
using namespace rt::tc;

constexpr int TB = 128;            // queries per chunk   (MMA M of GEMM1/GEMM2, K of GEMM3)
constexpr int TN = 128;            // entities per tile   (MMA N of GEMM1, K of GEMM2, M of GEMM3)
constexpr int KB = 32;             // r2-columns per staged block (GEMM1 k-block, GEMM2/3 n-block)
constexpr int kThreads = 256;
constexpr uint32_t RS = 128;       // 8 rows x 16 B
constexpr uint32_t CS_G = TB * 16;            // G[b][n]:  column-chunk stride (conflict-free: 32 lanes = 32 rows)
constexpr uint32_t CS_GT = TN * 16 + 16;      // G^T[n][b]: padded so 4-byte transposed stores spread over banks
constexpr uint32_t CS_ROW = TB * 16 + 16;     // staged [128 rows][32 cols] blocks (Q, O), padded
constexpr uint32_t CS_T = KB * 16;            // staged transposed [32 rows][128 cols] blocks (O^T, Q'^T)
constexpr uint32_t G_BYTES = (TN / 4) * CS_G;          // 65536
constexpr uint32_t GT_BYTES = (TB / 4) * CS_GT;        // 66048
constexpr uint32_t HALF_BYTES = (KB / 4) * CS_ROW;     // 16512  (>= (TN/4)*CS_T = 16384)
constexpr uint32_t STAGE_BYTES = 2 * HALF_BYTES;       // 33024
constexpr uint32_t OFF_G = 0, OFF_GT = OFF_G + G_BYTES, OFF_STAGE = OFF_GT + GT_BYTES;
constexpr uint32_t OFF_MASK = OFF_STAGE + 2 * STAGE_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_MASK + TB * 2 * 8;   // 199,680
constexpr int HB_LD = 129;                                // epilogue transpose buffer [128][129] floats in the stage area
static_assert(128 * HB_LD * 4 <= 2 * STAGE_BYTES, "transpose buffer must fit in the stage area");
constexpr uint32_t TM_D3 = 0, TM_Z = 256, TM_D2 = 256;

struct TcArgs {
  const float* q; const float* qp; const float* O;
  int B, r2, n_begin, n_local;
  const int32_t* off; const int32_t* idx;
  float t_pos, t_neg, inv_count;
  double* loss_partial; float* H_ws; float* dO;
  int n_tiles, r2p;
};

// [rows0, rows0+128) x [col0, col0+32) of src (row-major, ld) -> K-major block, zero padded
__device__ __forceinline__ void stage_rows(unsigned char* dst, const float* __restrict__ src, int ld, int row0,
                                           int rows_valid, int col0, int cols_valid, bool vec_ok) {
  for (int e = threadIdx.x; e < TB * (KB / 4); e += kThreads) {
    const int row = e >> 3, ch = e & 7;
    const int gr = row0 + row, gc = col0 + 4 * ch;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row < rows_valid) {
      const float* p = src + (int64_t)gr * ld + gc;
      if (vec_ok && gc + 4 <= col0 + cols_valid) {
        const float4 f = *reinterpret_cast<const float4*>(p);
        v = make_uint4(to_tf32(f.x), to_tf32(f.y), to_tf32(f.z), to_tf32(f.w));
      } else {
        uint32_t t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = (gc + j < col0 + cols_valid) ? to_tf32(p[j]) : 0u;
        v = make_uint4(t[0], t[1], t[2], t[3]);
      }
    }
    *reinterpret_cast<uint4*>(dst + (uint32_t)ch * CS_ROW + (uint32_t)(row >> 3) * RS + (uint32_t)(row & 7) * 16u) = v;
  }
}

// transposed: dst[c][r] for c in [col0, col0+32), r in [row0, row0+128): K-major along r
__device__ __forceinline__ void stage_cols_T(unsigned char* dst, const float* __restrict__ src, int ld, int row0,
                                             int rows_valid, int col0, int cols_valid) {
  for (int e = threadIdx.x; e < KB * (TN / 4); e += kThreads) {
    const int c = e & 31, rch = e >> 5;
    uint32_t t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 4 * rch + j;
      t[j] = (r < rows_valid && c < cols_valid) ? to_tf32(__ldg(src + (int64_t)(row0 + r) * ld + col0 + c)) : 0u;
    }
    *reinterpret_cast<uint4*>(dst + (uint32_t)rch * CS_T + (uint32_t)c * 16u) = make_uint4(t[0], t[1], t[2], t[3]);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
score_tc_kernel(TcArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar_stage[2], bar_z, bar_d;
  __shared__ uint32_t tmem_slot;
  __shared__ double red[kThreads / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;       // TMEM lane quarter / column half of this warp
  unsigned char* sG = smem + OFF_G;
  unsigned char* sGT = smem + OFF_GT;
  unsigned char* sStage[2] = {smem + OFF_STAGE, smem + OFF_STAGE + STAGE_BYTES};
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(smem + OFF_MASK);   // [TB][2]
  float* hbuf = reinterpret_cast<float*>(smem + OFF_STAGE);

  if (tid == 0) {
    mbar_init(&bar_stage[0], 1); mbar_init(&bar_stage[1], 1); mbar_init(&bar_z, 1); mbar_init(&bar_d, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  const int r2 = a.r2, r2p = a.r2p;
  const bool vec_ok = (r2 % 4 == 0);
  const int nkb = (r2 + KB - 1) / KB;          // GEMM1 k-blocks
  const int ncb = r2p / KB + ((r2p % KB) ? 1 : 0);   // GEMM2/3 column blocks (last one may be 16 wide)
  const int n_chunks = (a.B + TB - 1) / TB;
  const uint32_t idesc1 = make_idesc_tf32(TB, TN, false, false);
  uint32_t ph_stage[2] = {0u, 0u}, ph_z = 0u, ph_d = 0u;
  int used[2] = {0, 0};                          // number of un-waited commits per stage (0 or 1)
  double loss_acc = 0.0;
  bool first_tile = true;

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int n0 = tile * TN;
    const int n_valid = min(TN, a.n_local - n0);
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      const int b0 = chunk * TB;
      const int b_valid = min(TB, a.B - b0);
      // ================= GEMM1: Z = Q O^T, k-blocked, 2-stage ring =================
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1;
        if (used[s]) { mbar_wait(&bar_stage[s], ph_stage[s]); ph_stage[s] ^= 1u; used[s] = 0; }
        const int kw = min(KB, r2 - kb * KB);
        stage_rows(sStage[s], a.q, r2, b0, b_valid, kb * KB, kw, vec_ok);
        stage_rows(sStage[s] + HALF_BYTES, a.O, r2, n0, n_valid, kb * KB, kw, vec_ok);
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
          fence_after_sync();
          const uint32_t aA = smem_u32(sStage[s]), aB = aA + HALF_BYTES;
          const int ksteps = (kw + 7) / 8;
          for (int ks = 0; ks < ksteps; ++ks)
            mma_tf32(tmem + TM_Z, make_desc(aA + ks * 2 * CS_ROW, CS_ROW, RS), make_desc(aB + ks * 2 * CS_ROW, CS_ROW, RS),
                     idesc1, (kb | ks) != 0);
          mma_commit(&bar_stage[s]);
          if (kb == nkb - 1) mma_commit(&bar_z);
        }
        used[s] = 1;
      }
      // ---- sparse positives of this (chunk, tile) as 128-bit row masks ----
      mask[tid] = 0ull;                                   // kThreads == 2 * TB
      __syncthreads();
      {
        const int rr = tid >> 1, part = tid & 1;
        const int b = b0 + rr;
        if (b < a.B) {
          unsigned long long m0 = 0ull, m1 = 0ull;
          const int e1 = a.off[b + 1];
          for (int e = a.off[b] + part; e < e1; e += 2) {
            const int o = a.idx[e] - a.n_begin - n0;
            if (o >= 0 && o < 64) m0 |= 1ull << o;
            else if (o >= 64 && o < TN) m1 |= 1ull << (o - 64);
          }
          if (m0) atomicOr(&mask[rr * 2 + 0], m0);
          if (m1) atomicOr(&mask[rr * 2 + 1], m1);
        }
      }
      __syncthreads();
      // ================= epilogue 1: Z -> p, loss, G ; G and G^T staged as TF32 =================
      mbar_wait(&bar_z, ph_z); ph_z ^= 1u;
      fence_after_sync();
      {
        const int rr = quarter * 32 + lane;              // query row of this thread (TMEM lane)
        const int b = b0 + rr;
        const unsigned long long mrow = mask[rr * 2 + half];
        float loss_t = 0.0f;
#pragma unroll 1
        for (int cgp = 0; cgp < 2; ++cgp) {
          const int cbase = half * 64 + cgp * 32;          // entity column within the tile
          uint32_t v[32];
          tmem_ld32(tmem + TM_Z + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cbase, v);
          tmem_ld_wait();
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            uint32_t gq[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int j = j4 * 4 + jj;
              const int cc = cbase + j;
              const bool valid = (b < a.B) && (cc < n_valid);
              const bool pos = (mrow >> (cgp * 32 + j)) & 1ull;
              const float z = __uint_as_float(v[j]);
              const float p = 1.0f / (1.0f + expf(-z));
              const float lp = fmaxf(logf(p), -100.0f);
              const float lq = fmaxf(log1pf(-p), -100.0f);
              const float t = pos ? a.t_pos : a.t_neg;
              float g = 0.0f;
              if (valid) {
                loss_t -= t * lp + (1.0f - t) * lq;
                const float pq = (1.0f - p) * p;
                g = (p - t) / fmaxf(pq, 1e-12f) * a.inv_count * pq;
              }
              gq[jj] = to_tf32(g);
              // G^T[n][b]: row n = cc, col b = rr
              *reinterpret_cast<uint32_t*>(sGT + (uint32_t)(rr >> 2) * CS_GT + (uint32_t)(cc >> 3) * RS +
                                           (uint32_t)(cc & 7) * 16u + (uint32_t)(rr & 3) * 4u) = gq[jj];
            }
            // G[b][n]: row b = rr, cols cbase + 4*j4 .. +3  -> one 16-byte chunk
            const int cc0 = cbase + j4 * 4;
            *reinterpret_cast<uint4*>(sG + (uint32_t)(cc0 >> 2) * CS_G + (uint32_t)(rr >> 3) * RS + (uint32_t)(rr & 7) * 16u) =
                make_uint4(gq[0], gq[1], gq[2], gq[3]);
          }
        }
        loss_acc += (double)loss_t;
      }
      fence_before_sync();
      fence_async_smem();
      __syncthreads();
      // ================= GEMM2 (H chunk) and GEMM3 (dO tile), column-blocked =================
      for (int cb = 0; cb < ncb; ++cb) {
        const int s = (nkb + cb) & 1;
        if (used[s]) { mbar_wait(&bar_stage[s], ph_stage[s]); ph_stage[s] ^= 1u; used[s] = 0; }
        const int cw = min(KB, r2p - cb * KB);             // 32 or 16
        const int cvalid = max(0, min(cw, r2 - cb * KB));
        stage_cols_T(sStage[s], a.O, r2, n0, n_valid, cb * KB, cvalid);                 // O^T [c][n]
        stage_cols_T(sStage[s] + HALF_BYTES, a.qp, r2, b0, b_valid, cb * KB, cvalid);   // Q'^T [c][b]
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
          fence_after_sync();
          const uint32_t idn = make_idesc_tf32(128, cw, false, false);
          const uint32_t aOT = smem_u32(sStage[s]), aQT = aOT + HALF_BYTES;
          const uint32_t aG = smem_u32(sG), aGT = smem_u32(sGT);
          for (int ks = 0; ks < TN / 8; ++ks)      // H[b, c-block] = sum_n G[b,n] O^T[c,n]
            mma_tf32(tmem + TM_D2 + cb * KB, make_desc(aG + ks * 2 * CS_G, CS_G, RS),
                     make_desc(aOT + ks * 2 * CS_T, CS_T, RS), idn, ks != 0);
          for (int ks = 0; ks < TB / 8; ++ks)      // dO[n, c-block] += sum_b G^T[n,b] Q'^T[c,b]
            mma_tf32(tmem + TM_D3 + cb * KB, make_desc(aGT + ks * 2 * CS_GT, CS_GT, RS),
                     make_desc(aQT + ks * 2 * CS_T, CS_T, RS), idn, (chunk | ks) != 0);
          mma_commit(&bar_stage[s]);
          if (cb == ncb - 1) mma_commit(&bar_d);
        }
        used[s] = 1;
      }
      // all MMAs of this chunk done: operands (G, G^T, stages) and D2 are free / ready
      mbar_wait(&bar_d, ph_d); ph_d ^= 1u;
      fence_after_sync();
      for (int s = 0; s < 2; ++s)
        if (used[s]) { mbar_wait(&bar_stage[s], ph_stage[s]); ph_stage[s] ^= 1u; used[s] = 0; }
      // ================= epilogue 2: D2 -> H_ws slice of this CTA (through smem for coalescing) =================
      for (int pass = 0; pass * 128 < r2p; ++pass) {
        const int c0 = pass * 128;
        const int cw = min(128, r2p - c0);
        for (int cblk = half; cblk * 32 < cw; cblk += 2) {
          const int rr = quarter * 32 + lane;
          if (cw - cblk * 32 >= 32) {
            uint32_t v[32];
            tmem_ld32(tmem + TM_D2 + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c0 + cblk * 32), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) hbuf[rr * HB_LD + cblk * 32 + j] = __uint_as_float(v[j]);
          } else {
            uint32_t v[16];
            tmem_ld16(tmem + TM_D2 + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c0 + cblk * 32), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) hbuf[rr * HB_LD + cblk * 32 + j] = __uint_as_float(v[j]);
          }
        }
        fence_before_sync();
        __syncthreads();
        const int cols = min(cw, r2 - c0);
        float* Hc = a.H_ws + ((int64_t)blockIdx.x * a.B + b0) * r2 + c0;
        for (int e = tid; e < b_valid * cols; e += kThreads) {
          const int rr = e / cols, c = e - rr * cols;
          float* p = Hc + (int64_t)rr * r2 + c;
          const float val = hbuf[rr * HB_LD + c];
          *p = first_tile ? val : (*p + val);
        }
        __syncthreads();
      }
      fence_after_sync();
    }
    // ================= tile epilogue: D3 -> dO rows of this tile =================
    for (int pass = 0; pass * 128 < r2p; ++pass) {
      const int c0 = pass * 128;
      const int cw = min(128, r2p - c0);
      for (int cblk = half; cblk * 32 < cw; cblk += 2) {
        const int rr = quarter * 32 + lane;
        if (cw - cblk * 32 >= 32) {
          uint32_t v[32];
          tmem_ld32(tmem + TM_D3 + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c0 + cblk * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) hbuf[rr * HB_LD + cblk * 32 + j] = __uint_as_float(v[j]);
        } else {
          uint32_t v[16];
          tmem_ld16(tmem + TM_D3 + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c0 + cblk * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) hbuf[rr * HB_LD + cblk * 32 + j] = __uint_as_float(v[j]);
        }
      }
      fence_before_sync();
      __syncthreads();
      const int cols = min(cw, r2 - c0);
      for (int e = tid; e < n_valid * cols; e += kThreads) {
        const int rr = e / cols, c = e - rr * cols;
        a.dO[(int64_t)(n0 + rr) * r2 + c0 + c] = hbuf[rr * HB_LD + c];
      }
      __syncthreads();
    }
    fence_after_sync();
    first_tile = false;
  }
  if (first_tile) {   // CTA without a tile: its H slice must still be defined
    float* Hc = a.H_ws + (int64_t)blockIdx.x * a.B * r2;
    for (int64_t e = tid; e < (int64_t)a.B * r2; e += kThreads) Hc[e] = 0.0f;
  }
  double v = rt::warp_sum(loss_acc);
  if (lane == 0) red[warp] = v;
  fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) s += red[w];
    a.loss_partial[blockIdx.x] = s;
  }
  if (warp == 0) tmem_dealloc<512>(tmem);
}

__global__ void reduce_H_tc_kernel(const float* __restrict__ H_ws, int nparts, int64_t count, float* __restrict__ H) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.0f;
  for (int k = 0; k < nparts; ++k) s += H_ws[(int64_t)k * count + i];
  H[i] = s;
}
__global__ void reduce_loss_tc_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += partial[i];
    out[0] = s;
  }
}

int tc_grid(int n_local) {
  const int n_tiles = rt::cdiv(n_local, TN);
  const int sms = rt::sm_count();
  const int waves = rt::cdiv(n_tiles, sms);
  int grid = waves > 0 ? rt::cdiv(n_tiles, waves) : 1;
  return grid < 1 ? 1 : grid;
}

}  // namespace

extern "C" size_t rt_score_bce_tc_ws_bytes(int B, int n_local, int r2) {
  const int grid = tc_grid(n_local);
  return rt::align_up((size_t)grid * B * r2 * sizeof(float), 256) + (size_t)grid * sizeof(double);
}

extern "C" int rt_score_bce_tc(const float* q, const float* qp, const float* O, int B, int r2, int n_begin,
                               int n_local, int n_total, int b_total, const int32_t* tgt_off,
                               const int32_t* tgt_idx, float label_smoothing, double* loss_sum, float* H,
                               float* dO, void* ws, void* stream) {
  RT_REQUIRE(r2 >= 1 && r2 <= 256, "rt_score_bce_tc: r2=%d out of range (1..256)", r2);
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = tc_grid(n_local);
  TcArgs a{};
  a.q = q; a.qp = qp ? qp : q; a.O = O;
  a.B = B; a.r2 = r2; a.n_begin = n_begin; a.n_local = n_local;
  a.off = tgt_off; a.idx = tgt_idx;
  a.t_neg = label_smoothing / (float)n_total;
  a.t_pos = (1.0f - label_smoothing) + a.t_neg;
  a.inv_count = (float)(1.0 / ((double)b_total * (double)n_total));
  a.H_ws = (float*)ws;
  a.loss_partial = (double*)((char*)ws + rt::align_up((size_t)grid * B * r2 * sizeof(float), 256));
  a.dO = dO;
  a.n_tiles = rt::cdiv(n_local, TN);
  a.r2p = (r2 + 15) / 16 * 16;
  RT_CHECK_CUDA(cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  score_tc_kernel<<<grid, kThreads, SMEM_BYTES, s>>>(a);
  RT_LAUNCH_CHECK();
  const int64_t count = (int64_t)B * r2;
  reduce_H_tc_kernel<<<(int)((count + 255) / 256), 256, 0, s>>>(a.H_ws, grid, count, H);
  RT_LAUNCH_CHECK();
  reduce_loss_tc_kernel<<<1, 32, 0, s>>>(a.loss_partial, grid, loss_sum);
  RT_LAUNCH_CHECK();
  return 0;
}
