// fp64 tensor-core throughput by MMA shape on sm_100a (m8n8k4 vs the sm_90+ shapes m16n8k4 / k8 / k16) and plain DFMA.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/dmma_probe.cu -o tools/_probe/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>
template <int SHAPE, int NACC>
__global__ void __launch_bounds__(1024) probe(double* out, int iters, double seed) {
  double c[NACC][4];
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + i);
  for (int i = 0; i < 4; ++i) b[i] = seed * (threadIdx.x * 3 + i);
  for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
      if (SHAPE == 0) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a[0]), "d"(b[0]));
      } else if (SHAPE == 1) {
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                     : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
      } else if (SHAPE == 2) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
      } else if (SHAPE == 3) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                     : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                       "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
      } else if (SHAPE == 5) {   // m8n8k4 with operands that change from one instruction to the next (a real GEMM inner loop)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a[j & 7]), "d"(b[(j >> 1) & 3]));
      } else if (SHAPE == 6) {   // a 4 x 4 register tile: 4 A fragments x 4 B fragments (16 DMMAs per k-step)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a[j >> 2]), "d"(b[j & 3]));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) c[j][i] = fma(a[i], b[i], c[j][i]);
      }
    }
  }
  double s = 0.0;
  for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int SHAPE, int NACC>
void run(const char* name, double flop_per_inst_per_warp, int threads) {
  double* out; cudaMalloc(&out, 148 * 1024 * 8 * 2);
  const int iters = 4000;
  probe<SHAPE, NACC><<<148, threads>>>(out, 10, 1e-9);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<SHAPE, NACC><<<148, threads>>>(out, iters, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double insts = (double)iters * NACC * (threads / 32) * 148;
  printf("%-26s acc=%2d warps/SM=%2d: %8.3f ms  %7.2f TFLOP/s  (%.1f clk per warp-instruction per SM at 1.965 GHz)%s\n", name, NACC, threads / 32, ms,
         insts * flop_per_inst_per_warp / ms * 1e-9, ms * 1e-3 * 1.965e9 / (insts / 148), cudaGetLastError() ? " ERROR" : "");
  cudaFree(out);
}
// conversion rates: F2F.F64.F32 / F2F.F32.F64 (hardware casts) against the integer-pipe versions of csrc/common.h
__device__ __forceinline__ double cv_int(float f) {
  const unsigned int u = __float_as_uint(f);
  const unsigned int ex = u & 0x7f800000u;
  unsigned int hi = (u & 0x80000000u) | (((u & 0x7fffffffu) >> 3) + 0x38000000u), lo = u << 29;
  if (ex == 0u) { if (u & 0x007fffffu) return (double)f; hi = u & 0x80000000u; lo = 0u; } else if (ex == 0x7f800000u) return (double)f;
  return __hiloint2double((int)hi, (int)lo);
}
template <int KIND>
__global__ void __launch_bounds__(1024) cvt_probe(double* out, int iters, float seed) {
  float f[8]; double acc[8];
  for (int i = 0; i < 8; ++i) { f[i] = seed * (threadIdx.x + 1 + i); acc[i] = 0.0; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double d = KIND == 0 ? (double)f[i] : cv_int(f[i]);
      // keep the chain on the integer pipe: fold the bits back into the next input
      const unsigned int h = (unsigned int)__double2hiint(d);
      f[i] = __uint_as_float((__float_as_uint(f[i]) + (h & 1u) + 1u) & 0x3fffffffu | 0x30000000u);
      acc[i] = __hiloint2double(__double2hiint(acc[i]) ^ (int)h, __double2loint(d));
    }
  }
  double s = 0.0;
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND>
void run_cvt(const char* name, int threads) {
  double* out; cudaMalloc(&out, 148 * 1024 * 8);
  const int iters = 2000;
  cvt_probe<KIND><<<148, threads>>>(out, 10, 1.0f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  cvt_probe<KIND><<<148, threads>>>(out, iters, 1.0f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double per_sm = (double)iters * 8 * threads;
  printf("%-28s warps/SM=%2d: %8.3f ms  %6.2f conversions per clock per SM\n", name, threads / 32, ms, per_sm / (ms * 1e-3 * 1.965e9));
  cudaFree(out);
}
int main() {
  run_cvt<0>("fp32 -> fp64 cast (F2F)", 256); run_cvt<0>("fp32 -> fp64 cast (F2F)", 1024);
  run_cvt<1>("fp32 -> fp64 integer pipes", 256); run_cvt<1>("fp32 -> fp64 integer pipes", 1024);
  for (int threads : {256, 512}) {
    if (threads == 256) { run<5, 16>("m8n8k4 rotating operands", 512, 256); run<6, 16>("m8n8k4 4x4 register tile", 512, 256); }
    if (threads == 512) { run<5, 16>("m8n8k4 rotating operands", 512, 512); run<6, 16>("m8n8k4 4x4 register tile", 512, 512); }
  }
  for (int threads : {128, 256, 512, 1024}) {
    if (threads == 128) { run<0, 16>("m8n8k4", 512, 128); run<1, 8>("m16n8k4", 1024, 128); run<2, 8>("m16n8k8", 2048, 128); run<3, 8>("m16n8k16", 4096, 128); run<4, 8>("dfma x4", 256, 128); }
    if (threads == 256) { run<0, 16>("m8n8k4", 512, 256); run<1, 8>("m16n8k4", 1024, 256); run<2, 8>("m16n8k8", 2048, 256); run<3, 8>("m16n8k16", 4096, 256); run<4, 8>("dfma x4", 256, 256); }
    if (threads == 512) { run<0, 16>("m8n8k4", 512, 512); run<1, 8>("m16n8k4", 1024, 512); run<2, 8>("m16n8k8", 2048, 512); run<3, 8>("m16n8k16", 4096, 512); run<4, 8>("dfma x4", 256, 512); }
    if (threads == 1024) { run<0, 16>("m8n8k4", 512, 1024); run<1, 8>("m16n8k4", 1024, 1024); run<2, 8>("m16n8k8", 2048, 1024); run<3, 8>("m16n8k16", 4096, 1024); run<4, 8>("dfma x4", 256, 1024); }
  }
  return 0;
}
