// tcgen05.ld (TMEM -> registers) throughput probe on sm_100a: 4..16 warps per CTA read 32x32b.x32 chunks back to back.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void __launch_bounds__(512, 1) probe(float* out, int iters, int nwarps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(base + ((warp >> 2) * 128 + c * 32) % 512));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i];
      }
    }
  }
  if (acc == 0x12345678) out[0] = 1.f;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

int main() {
  float* out;
  cudaMalloc(&out, 4);
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int nw : {4, 8, 16}) {
    const int iters = 2048;
    probe<<<sms, 512>>>(out, 8, nw);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<<<sms, 512>>>(out, iters, nw);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double words = (double)sms * nw * 32 * iters * 4 * 32;  // fp32 words read
    printf("tcgen05.ld 32x32b.x32, %2d warps/SM: %.3f ms -> %.1f fp32 words/clk/SM (%.0f B/clk/SM) at 1.965 GHz; err=%s\n", nw, ms,
           words / (ms * 1e-3) / 1.965e9 / sms, 4 * words / (ms * 1e-3) / 1.965e9 / sms, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
