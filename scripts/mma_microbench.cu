// Measures what one SM's tcgen05.mma pipe sustains for the shape k_tc_gemm uses (M=128, N=128, K=16, fp16 -> fp32),
// with all 148 SMs busy, for several issue patterns.  Build + run (B200):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_microbench scripts/mma_microbench.cu && /tmp/mma_microbench
// Output: cycles per MMA (clock64 on the issuing thread, max over SMs) and the implied dense TFLOP/s at the measured
// SM clock.  Used for DESIGN.md's "what bounds k_tc_gemm" section; not part of the product.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// mode 0: SS, one accumulator        1: SS, 4 accumulators round-robin     2: TS, one accumulator
// mode 3: SS, commit after every 4 MMAs (no wait)    4: SS N=256 one accumulator   5: TS, 2 accumulators round-robin
__global__ void __launch_bounds__(64, 1) k_bench(int mode, int reps, long long* cycles) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 64) reinterpret_cast<uint32_t*>(raw)[i] = 0x3C003C00u;  // fp16 1.0
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&bar2), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_slot;
  if (threadIdx.x == 0) {
    const int n = (mode == 4) ? 256 : 128;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sA = base, sB = base + 64 * 1024;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = desc_sw128(sA + kb * 16384 + k * 32), bd = desc_sw128(sB + kb * 16384 + k * 32);
          const int m = kb * 4 + k;
          if (mode == 0 || mode == 3) mma_ss(tm, ad, bd, idesc, 1);
          else if (mode == 1) mma_ss(tm + (m & 3) * 128, ad, bd, idesc, 1);
          else if (mode == 2) mma_ts(tm, tm + 256 + kb * 32 + k * 8, bd, idesc, 1);
          else if (mode == 4) mma_ss(tm, ad, bd, idesc, 1);
          else mma_ts(tm + (m & 1) * 128, tm + 256 + kb * 32 + k * 8, bd, idesc, 1);
        }
        if (mode == 3) tc_commit(smem_u32(&bar2));
      }
    }
    tc_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
  }
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  long long* d;
  cudaMalloc(&d, sms * sizeof(long long));
  const size_t smem = 161 * 1024;
  cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char* names[] = {"SS 128x128x16, 1 accumulator", "SS 128x128x16, 4 accumulators", "TS 128x128x16, 1 accumulator",
                         "SS 128x128x16, commit per 4 MMAs", "SS 128x256x16, 1 accumulator", "TS 128x128x16, 2 accumulators"};
  const int reps = 2000;
  for (int mode = 0; mode < 6; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      k_bench<<<sms, 64, smem>>>(mode, reps, d);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(err)); return 1; }
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      long long h[256]; cudaMemcpy(h, d, sms * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
      const double n = (mode == 4) ? 256 : 128;
      const double mmas = 16.0 * reps, flop = 2.0 * 128 * n * 16 * mmas * sms;
      if (rep == 1)
        printf("%-36s %7.1f clk/MMA (clock64)  kernel %.3f ms  %.0f TFLOP/s dense (wall)\n", names[mode], mx / mmas, ms,
               flop / (ms * 1e-3) / 1e12);
    }
  }
  printf("SMs %d, nominal clock %.0f MHz\n", sms, khz / 1000.0);
  return 0;
}
