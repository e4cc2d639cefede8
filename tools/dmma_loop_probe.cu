// What bounds a DMMA k-loop that takes its fragments straight from L2 (the GEMM unit of csrc/small_exec.cuh, gram_sym.cu)?
// A CTA of 8 warps owns a 32 x 32 tile of C = A B^T (A, B: n x K row-major fp64, L2 resident), the warps split K,
// 16 DMMA m8n8k4 per k-step and warp, operands prefetched DEPTH steps ahead.  Variants: loads + DMMA, DMMA only
// (operands loaded once), loads only.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/dmma_loop_probe.cu -o tools/_probe/dmma_loop_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE, int DEPTH>   // MODE 0: loads + dmma, 1: dmma only, 2: loads only
__global__ void __launch_bounds__(256, 2) loop_probe(const double* __restrict__ A, const double* __restrict__ B, int n, int K, int reps, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int tiles = n / 32;
  const int tm = (blockIdx.x % (tiles * tiles)) / tiles, tn = blockIdx.x % tiles;
  const double* pa[4]; const double* pb[4];
  for (int x = 0; x < 4; ++x) { pa[x] = A + (size_t)(tm * 32 + 8 * x + g) * K + t; pb[x] = B + (size_t)(tn * 32 + 8 * x + g) * K + t; }
  double c[4][4][2];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }
  double ra[DEPTH][4], rb[DEPTH][4];
  const int steps = K / 4;
  double sink = 0.0;
  for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d)
#pragma unroll
      for (int x = 0; x < 4; ++x) { ra[d][x] = __ldcg(pa[x] + 4 * (warp + 8 * d)); rb[d][x] = __ldcg(pb[x] + 4 * (warp + 8 * d)); }
    for (int s0 = warp; s0 < steps; s0 += 8 * DEPTH) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        const int s = s0 + 8 * d;
        if (s >= steps) break;
        const int sp = s + 8 * (DEPTH - 1);
        constexpr int dn_of[8] = {0, 1, 2, 3, 4, 5, 6, 7};
        const int dn = (d + DEPTH - 1) % DEPTH;
        (void)dn_of;
        if (MODE != 1 && sp < steps) {
#pragma unroll
          for (int x = 0; x < 4; ++x) { ra[dn][x] = __ldcg(pa[x] + 4 * sp); rb[dn][x] = __ldcg(pb[x] + 4 * sp); }
        }
        if (MODE != 2) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma(c[i][j][0], c[i][j][1], ra[d][i], rb[d][j]);
        } else {
#pragma unroll
          for (int x = 0; x < 4; ++x) sink += ra[d][x] + rb[d][x];
        }
      }
    }
  }
  double s = sink;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j][0] + c[i][j][1];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
template <int MODE, int DEPTH>
void run(const char* name, const double* A, const double* B, int n, int K, double* out) {
  const int reps = 20, grid = 296;
  loop_probe<MODE, DEPTH><<<grid, 256>>>(A, B, n, K, 1, out);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  loop_probe<MODE, DEPTH><<<grid, 256>>>(A, B, n, K, reps, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double steps_per_sm = 2.0 * reps * (K / 4);          // 2 CTAs per SM, all warps together run K / 4 steps per rep
  const double cyc = ms * 1e-3 * 1.965e9;
  printf("%-34s K=%5d: %7.3f ms  %7.1f clk per k-step per SM (16 DMMA, 2 KB of operands)  = %5.2f clk/DMMA, %5.1f B/clk/SM %s\n", name, K, ms,
         cyc / steps_per_sm, cyc / steps_per_sm / 16.0, 2048.0 / (cyc / steps_per_sm), cudaGetLastError() ? "ERROR" : "");
}
int main() {
  const int n = 224;
  for (int K : {200, 2000}) {
    double *A, *B, *out;
    cudaMalloc(&A, sizeof(double) * n * K); cudaMalloc(&B, sizeof(double) * n * K); cudaMalloc(&out, 8 * 296 * 256);
    cudaMemset(A, 0, sizeof(double) * n * K); cudaMemset(B, 0, sizeof(double) * n * K);
    run<0, 3>("loads + DMMA, prefetch depth 3", A, B, n, K, out);
    run<0, 2>("loads + DMMA, prefetch depth 2", A, B, n, K, out);
    run<0, 4>("loads + DMMA, prefetch depth 4", A, B, n, K, out);
    run<1, 3>("DMMA only", A, B, n, K, out);
    run<2, 3>("loads only", A, B, n, K, out);
    cudaFree(A); cudaFree(B); cudaFree(out);
  }
  return 0;
}
