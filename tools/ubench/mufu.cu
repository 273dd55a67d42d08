// Throughput probes for the attention softmax inner loop on sm_100a: MUFU.EX2, bf16x2 pack (F2FP), FFMA, FMNMX, and
// tcgen05.ld bandwidth.  ops/clk/SM = total ops / (elapsed cycles * SMs).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) probe(float* out, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-6f + i;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) a[i] = fmaf(a[i], 1.0001f, 0.5f);
      if (MODE == 2) a[i] = fmaxf(a[i], a[(i + 1) & 7] * 0.5f);
      if (MODE == 3) {  // pack two floats to bf16x2
        __nv_bfloat162 t = __floats2bfloat162_rn(a[i], a[(i + 1) & 7]);
        acc ^= *reinterpret_cast<uint32_t*>(&t);
        a[i] += 1.0f;
      }
      if (MODE == 5) {  // packed half2 exponential: TWO elements per MUFU instruction?
        uint32_t h = __float_as_uint(a[i]);
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h));
        a[i] = __uint_as_float(h);
      }
      if (MODE == 6) {  // packed bf16x2 exponential
        uint32_t h = __float_as_uint(a[i]);
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h));
        a[i] = __uint_as_float(h);
      }
      if (MODE == 7) {  // candidate softmax inner loop, per PAIR of scores: 2 FMNMX + 2 FFMA + pack to half2 + one f16x2 ex2
        const float s0 = a[i], s1 = a[(i + 1) & 7];
        const float x0 = fmaf(s0, 0.125f, -1.f), x1 = fmaf(s1, 0.125f, -1.f);
        __half2 hx = __floats2half2_rn(x0, x1);
        uint32_t h = *reinterpret_cast<uint32_t*>(&hx);
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h));
        acc ^= h;
        a[i] = fmaxf(s0, s1) + 1e-3f;
      }
      if (MODE == 8) {  // current softmax inner loop, per PAIR of scores: 2 FMNMX + 2 FFMA + 2 f32 ex2 + 2 FADD + 1 bf16x2 pack
        const float s0 = a[i], s1 = a[(i + 1) & 7];
        float e0, e1;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(s0, 0.125f, -1.f)));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(s1, 0.125f, -1.f)));
        __nv_bfloat162 t = __floats2bfloat162_rn(e0, e1);
        acc ^= *reinterpret_cast<uint32_t*>(&t);
        a[i] = fmaxf(s0, s1) + (e0 + e1) * 1e-6f;
      }
      if (MODE == 4) {  // ex2 + fma + pack mix like the softmax loop
        float e;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(a[i], 0.125f, -1.f)));
        __nv_bfloat162 t = __floats2bfloat162_rn(e, e);
        acc ^= *reinterpret_cast<uint32_t*>(&t);
        a[i] += e;
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 123.456f || acc == 0x12345) out[0] = s;
}

template <int MODE>
void run(const char* name, int threads) {
  float* out;
  cudaMalloc(&out, 4);
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int iters = 4096;
  probe<MODE><<<sms, threads>>>(out, 16, 1.f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<MODE><<<sms, threads>>>(out, iters, 1.f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)sms * threads * iters * 8;
  printf("%-28s threads/SM %4d: %8.3f ms  -> %6.1f ops/clk/SM at %d MHz nominal (%.1f at 1.965 GHz)\n", name, threads, ms,
         ops / (ms * 1e-3) / (clk * 1e3) / sms, clk / 1000, ops / (ms * 1e-3) / 1.965e9 / sms);
  cudaFree(out);
}

int main() {
  for (int th : {256, 512, 1024}) {
    run<0>("MUFU.EX2", th);
    run<1>("FFMA", th);
    run<2>("FMNMX+FMUL", th);
    run<3>("F2FP.BF16 pack (+FADD)", th);
    run<4>("FFMA+EX2+pack+FADD mix", th);
    run<5>("MUFU.EX2 f16x2 (instr/clk)", th);
    run<6>("MUFU.EX2 bf16x2 (instr/clk)", th);
    run<7>("pair: half2 ex2 loop (pairs)", th);
    run<8>("pair: f32 ex2 loop (pairs)", th);
  }
  return 0;
}
