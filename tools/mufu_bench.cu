// Per-SM throughput of the special-function / conversion instructions the epilogues lean on.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>
template <int OP>
__global__ void k(float* out, int iters) {
  float a = threadIdx.x * 1e-3f + 0.1f, b = a + 0.01f, c = a + 0.02f, d = a + 0.03f;
  uint32_t acc = 0;
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(b)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(c)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(d)); }
    if (OP == 1) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d)); }
    if (OP == 2) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(b)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(d)); }
    if (OP == 3) { uint32_t r0, r1; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r0) : "f"(a), "f"(b)); asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r1) : "f"(c), "f"(d)); acc ^= r0 ^ r1; a += 1.f; c += 1.f; }
    if (OP == 4) { a = fmaf(a, 1.0001f, 0.5f); b = fmaf(b, 1.0001f, 0.5f); c = fmaf(c, 1.0001f, 0.5f); d = fmaf(d, 1.0001f, 0.5f); }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + __uint_as_float(acc);
}
template <int OP> void run(const char* name, int ops_per_iter) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 20000;
  k<OP><<<148, 1024>>>(out, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<OP><<<148, 1024>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double ops = 1024.0 * iters * ops_per_iter;            // per SM
  printf("%-22s %8.3f ms  -> %.1f ops/ns/SM  (~%.1f per clk at %.2f GHz nominal)\n", name, ms, ops / (ms * 1e6), ops / (ms * 1e6) / (clk * 1e-6), clk * 1e-6);
  cudaFree(out);
}
int main() {
  run<0>("tanh.approx", 4); run<1>("ex2.approx", 4); run<2>("rcp.approx", 4); run<3>("cvt.bf16x2 (per pair)", 2); run<4>("ffma", 4);
  return 0;
}
