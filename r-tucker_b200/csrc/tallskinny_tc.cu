// Tall-skinny passes on the tensor cores: tcgen05 TF32 with a 3-product split (x = x_hi + x_lo,
// x.y ~ x_hi.y_hi + x_hi.y_lo + x_lo.y_hi), i.e. fp32-level accuracy (~1e-6 relative) at tensor-core rate.
//
//   (the factor updates Y = a0 X0 + sum_t X_t K_t live in apply_tc.cu)
//   gram :  C[ra, rb] = A[n, ra]^T B[n, rb]                                (Gram products of grad / norm / project)
//
// Both keep their accumulators in tensor memory and stage K-major operands in the interleaved
// core-matrix format of tc.cuh; the small right-hand matrices of `apply` are pre-split and pre-packed
// once per call into the exact shared-memory image, so a CTA stages them with plain 16-byte async copies.
#include "common.h"
#include "tc.cuh"

namespace {
using namespace rt::tc;

constexpr int kThreads = 256;
constexpr int TM = 128;             // rows per CTA tile
constexpr int KB = 32;              // contraction elements per staged block
constexpr uint32_t RS = 128;
constexpr uint32_t CS_X = TM * 16 + 16;                 // [128 rows][32 k] blocks (padded chunk stride)
constexpr uint32_t X_BYTES = (KB / 4) * CS_X;           // 16512
constexpr int RCP_MAX = 256;

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// gram
// ---------------------------------------------------------------------------------------------------
constexpr int GKB = 16;                          // source rows (contraction elements) per staged block of the Gram
constexpr uint32_t CS_GA = 128 * 16 + 16;        // A^T block: [128 rows (a)][16 k (n)]
struct GramTcArgs {
  const float* A; int64_t lda; const float* B; int64_t ldb;
  int n, ra, rb, rbp, mtiles, rows_per_cta;
  float* partial;     // [grid][mtiles*128][rbp] fp32
  uint32_t cs_b, b_bytes;   // B^T block: [rbp rows (b)][32 k], chunk stride rbp*16+16
};

// dst[c][r] for c < ncols (operand rows), r in a block of 32 source rows: hi and lo images
__device__ __forceinline__ void stage_T_split(unsigned char* dst_hi, unsigned char* dst_lo, uint32_t cs,
                                              const float* __restrict__ src, int64_t ld, int row0, int rows_valid,
                                              int col0, int ncols_pad, int ncols_valid) {
  // item = (c, rch): 4 consecutive source rows r = 4*rch .. +3 -> one 16-byte chunk of operand row c
  for (int e = threadIdx.x; e < ncols_pad * (GKB / 4); e += kThreads) {
    const int c = e % ncols_pad, rch = e / ncols_pad;
    float t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 4 * rch + j;
      t[j] = (r < rows_valid && c < ncols_valid) ? __ldg(src + (int64_t)(row0 + r) * ld + col0 + c) : 0.0f;
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_tf32(t[j], h[j], l[j]);
    const uint32_t off = (uint32_t)rch * cs + (uint32_t)(c >> 3) * RS + (uint32_t)(c & 7) * 16u;
    *reinterpret_cast<uint4*>(dst_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(dst_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
gram_tc_kernel(GramTcArgs g) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar_stage[2], bar_done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const uint32_t a_bytes = (GKB / 4) * CS_GA;                     // one M-tile image (hi or lo)
  const uint32_t stage_bytes = 2 * g.mtiles * a_bytes + 2 * g.b_bytes;
  if (tid == 0) { mbar_init(&bar_stage[0], 1); mbar_init(&bar_stage[1], 1); mbar_init(&bar_done, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int r_begin = blockIdx.x * g.rows_per_cta;
  const int r_end = min(g.n, r_begin + g.rows_per_cta);
  const int nblocks = (r_end > r_begin) ? (r_end - r_begin + GKB - 1) / GKB : 0;
  const uint32_t idesc = make_idesc_tf32(128, g.rbp, false, false);
  uint32_t ph[2] = {0u, 0u};
  int used[2] = {0, 0};
  for (int blk = 0; blk < nblocks; ++blk) {
    const int s = blk & 1;
    unsigned char* st = smem + (size_t)s * stage_bytes;
    if (used[s]) { mbar_wait(&bar_stage[s], ph[s]); ph[s] ^= 1u; used[s] = 0; }
    const int row0 = r_begin + blk * GKB;
    const int rows_valid = min(GKB, r_end - row0);
    for (int mt = 0; mt < g.mtiles; ++mt)
      stage_T_split(st + (size_t)(2 * mt) * a_bytes, st + (size_t)(2 * mt + 1) * a_bytes, CS_GA, g.A, g.lda, row0,
                    rows_valid, mt * 128, 128, min(128, g.ra - mt * 128));
    unsigned char* sb = st + (size_t)2 * g.mtiles * a_bytes;
    stage_T_split(sb, sb + g.b_bytes, g.cs_b, g.B, g.ldb, row0, rows_valid, 0, g.rbp, g.rb);
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      const uint32_t aB = smem_u32(sb);
      for (int mt = 0; mt < g.mtiles; ++mt) {
        const uint32_t aAh = smem_u32(st + (size_t)(2 * mt) * a_bytes), aAl = aAh + a_bytes;
        const uint32_t d = tmem + (uint32_t)(mt * 256);
#pragma unroll
        for (int ks = 0; ks < GKB / 8; ++ks) {
          const uint64_t dAh = make_desc(aAh + ks * 2 * CS_GA, CS_GA, RS), dAl = make_desc(aAl + ks * 2 * CS_GA, CS_GA, RS);
          const uint64_t dBh = make_desc(aB + ks * 2 * g.cs_b, g.cs_b, RS);
          const uint64_t dBl = make_desc(aB + g.b_bytes + ks * 2 * g.cs_b, g.cs_b, RS);
          mma_tf32(d, dAl, dBh, idesc, (blk | ks) != 0);
          mma_tf32(d, dAh, dBl, idesc, true);
          mma_tf32(d, dAh, dBh, idesc, true);
        }
      }
      mma_commit(&bar_stage[s]);
      if (blk == nblocks - 1) mma_commit(&bar_done);
    }
    used[s] = 1;
  }
  if (nblocks > 0) { mbar_wait(&bar_done, 0); fence_after_sync(); }
  // ---- partial C of this CTA -> workspace (fp32), rows a = mt*128 + lane row ----
  float* out = g.partial + (size_t)blockIdx.x * g.mtiles * 128 * g.rbp;
  for (int mt = 0; mt < g.mtiles; ++mt) {
    const int arow = mt * 128 + quarter * 32 + lane;
    for (int c0 = half * 32; c0 < g.rbp; c0 += 64) {
      uint32_t v[32];
      if (nblocks > 0) {
        if (g.rbp - c0 >= 32) tmem_ld32(tmem + (uint32_t)(mt * 256) + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
        else {
          uint32_t w[16];
          tmem_ld16(tmem + (uint32_t)(mt * 256) + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, w);
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = w[j]; v[j + 16] = 0u; }
        }
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j < g.rbp) out[(size_t)arow * g.rbp + c0 + j] = __uint_as_float(v[j]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

__global__ void gram_tc_reduce_kernel(const float* __restrict__ partial, int nparts, int mrows, int rbp, int ra, int rb,
                                      double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ra * rb) return;
  const int i = e / rb, j = e - i * rb;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += (double)partial[((size_t)p * mrows + i) * rbp + j];
  out[e] = s;
}

struct GramTcPlan { int grid, rows_per_cta, mtiles, rbp; size_t smem; uint32_t cs_b, b_bytes; };
GramTcPlan gram_tc_plan(int n, int ra, int rb) {
  GramTcPlan p;
  p.mtiles = rt::cdiv(ra, 128);
  p.rbp = (rb + 15) / 16 * 16;
  p.cs_b = (uint32_t)p.rbp * 16u + 16u;
  p.b_bytes = (GKB / 4) * p.cs_b;
  p.smem = 2 * ((size_t)2 * p.mtiles * (GKB / 4) * CS_GA + 2 * p.b_bytes);
  int grid = rt::sm_count();
  int rows = rt::cdiv(n, grid);
  rows = rt::cdiv(rows, GKB) * GKB;
  if (rows < GKB) rows = GKB;
  p.rows_per_cta = rows;
  p.grid = rt::cdiv(n, rows);
  return p;
}

}  // namespace

// ---- C ABI -------------------------------------------------------------------------------------------
extern "C" int rt_gram_tc_supported(int ra, int rb) {
  if (ra < 1 || ra > 256 || rb < 1 || rb > 256) return 0;
  GramTcPlan p = gram_tc_plan(1024, ra, rb);
  return p.smem <= 220 * 1024 ? 1 : 0;
}

extern "C" size_t rt_gram_tc_ws_bytes(int n, int ra, int rb) {
  GramTcPlan p = gram_tc_plan(n > 0 ? n : 1, ra, rb);
  return (size_t)p.grid * p.mtiles * 128 * p.rbp * sizeof(float) + 256;
}

extern "C" int rt_gram_tc(const float* A, int64_t lda, const float* B, int64_t ldb, int n, int ra, int rb,
                          double* out, void* ws, void* stream) {
  RT_REQUIRE(rt_gram_tc_supported(ra, rb), "rt_gram_tc: unsupported shape ra=%d rb=%d", ra, rb);
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) { RT_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * ra * rb, s)); return 0; }
  RT_REQUIRE(ws != nullptr, "rt_gram_tc: workspace is NULL");
  GramTcPlan p = gram_tc_plan(n, ra, rb);
  GramTcArgs g{};
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.n = n; g.ra = ra; g.rb = rb; g.rbp = p.rbp; g.mtiles = p.mtiles;
  g.rows_per_cta = p.rows_per_cta; g.partial = (float*)ws; g.cs_b = p.cs_b; g.b_bytes = p.b_bytes;
  RT_CHECK_CUDA(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  gram_tc_kernel<<<p.grid, kThreads, p.smem, s>>>(g);
  RT_LAUNCH_CHECK();
  gram_tc_reduce_kernel<<<rt::cdiv(ra * rb, 256), 256, 0, s>>>(g.partial, p.grid, p.mtiles * 128, p.rbp, ra, rb, out);
  RT_LAUNCH_CHECK();
  return 0;
}
