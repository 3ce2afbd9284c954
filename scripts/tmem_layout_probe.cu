// tmem_layout_probe.cu -- prints which (lane, column) of tensor memory each thread receives from
// tcgen05.ld.sync.aligned.16x256b.x8 (used by k_tc_gemm's dual-direction epilogue).  Lane l, column c of TMEM is
// first filled with the value l * 1000 + c through the 32x32b shape (thread = lane, register = column).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_layout_probe scripts/tmem_layout_probe.cu && ./tmem_layout_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(uint32_t* out) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s;
  // each of the 4 warps owns lanes 32*warp .. +31; fill 64 columns
  uint32_t v[32];
  for (int half = 0; half < 2; ++half) {
    for (int c = 0; c < 32; ++c) v[c] = (uint32_t)((warp * 32 + lane) * 1000 + half * 32 + c);
    const uint32_t taddr = base + half * 32 + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // read back with 16x256b.x8 at lane offsets 0 and 16 of this warp's quadrant
  for (int lo = 0; lo < 2; ++lo) {
    uint32_t r[32];
    const uint32_t taddr = base + ((uint32_t)(warp * 32 + lo * 16) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) out[((warp * 2 + lo) * 32 + lane) * 32 + i] = r[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64) : "memory");
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 4 * 2 * 32 * 32 * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  static uint32_t h[4 * 2 * 32 * 32];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int w = 0; w < 4; ++w)
    for (int lo = 0; lo < 2; ++lo)
      for (int t = 0; t < 32; ++t)
        for (int i = 0; i < 32; ++i) {
          const uint32_t v = h[((w * 2 + lo) * 32 + t) * 32 + i];
          const int lane = v / 1000, col = v % 1000;
          // expectation (CUTLASS SM100_TMEM_LOAD_16dp256b8x): reg 4n + 2h + e = lane t/4 + 8h, column 8n + 2(t%4) + e
          const int n = i >> 2, hh = (i >> 1) & 1, e2 = i & 1;
          const int want_lane = w * 32 + lo * 16 + t / 4 + 8 * hh, want_col = 8 * n + 2 * (t % 4) + e2;
          if (lane != want_lane || col != want_col) {
            if (bad < 20) printf("warp %d lo %d thread %d reg %d: got lane %d col %d, expected lane %d col %d\n", w, lo, t, i, lane, col, want_lane, want_col);
            ++bad;
          }
        }
  printf("tcgen05.ld.16x256b.x8 layout: %s (%d mismatches)\n", bad ? "DIFFERENT from the expectation" : "as expected", bad);
  for (int i = 0; i < 8; ++i) printf("warp 1, lane-offset 16, thread 5, reg %d -> lane %u col %u\n", i, h[((1 * 2 + 1) * 32 + 5) * 32 + i] / 1000, h[((1 * 2 + 1) * 32 + 5) * 32 + i] % 1000);
  return bad ? 2 : 0;
}
