// Experiment (GPU): discover the (thread, register) -> (TMEM lane, column) mapping of the tcgen05.st shapes by
// writing unique ids and reading them back with tcgen05.ld.32x32b (thread t <- lane 32*w + t, consecutive columns).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(uint32_t* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  const uint32_t wbase = base + ((uint32_t)(warp * 32) << 16);
  for (int shape = 0; shape < 4; ++shape) {
    // clear 16 columns
    {
      uint32_t z = 0xFFFFFFFFu;
      for (int c = 0; c < 16; ++c)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(wbase + c), "r"(z) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    uint32_t r[8];
    for (int j = 0; j < 8; ++j) r[j] = (uint32_t)(lane << 8 | j);
    for (int h = 0; h < 2; ++h) {   // two 16-lane halves
      const uint32_t a = wbase + ((uint32_t)(h * 16) << 16);
      uint32_t q[8];
      for (int j = 0; j < 8; ++j) q[j] = r[j] | (h << 16);
      if (shape == 0) asm volatile("tcgen05.st.sync.aligned.16x64b.x1.b32 [%0], {%1};" ::"r"(a), "r"(q[0]) : "memory");
      if (shape == 1) asm volatile("tcgen05.st.sync.aligned.16x128b.x1.b32 [%0], {%1, %2};" ::"r"(a), "r"(q[0]), "r"(q[1]) : "memory");
      if (shape == 2) asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]) : "memory");
      if (shape == 3) asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]), "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(wbase) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (warp == 0) for (int c = 0; c < 16; ++c) out[(shape * 32 + lane) * 16 + c] = v[c];
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64u) : "memory");
}
int main() {
  uint32_t* d; cudaMalloc(&d, 4 * 32 * 16 * 4);
  k<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  static uint32_t h[4 * 32 * 16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[4] = {"16x64b.x1", "16x128b.x1", "16x256b.x1", "16x256b.x2"};
  for (int s = 0; s < 4; ++s) {
    printf("shape %s: TMEM lane: col0..15 as half.thread.reg (-- = untouched)\n", names[s]);
    for (int l = 0; l < 32; ++l) {
      printf("  lane %2d:", l);
      for (int c = 0; c < 16; ++c) {
        uint32_t v = h[(s * 32 + l) * 16 + c];
        if (v == 0xFFFFFFFFu) printf(" ------");
        else printf(" %u.%02u.%u", v >> 16, (v >> 8) & 0xFF, v & 0xFF);
      }
      printf("\n");
    }
  }
  return 0;
}
