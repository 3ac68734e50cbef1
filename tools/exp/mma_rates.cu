// Micro-benchmark (GPU): how fast can small-N tensor-core instructions be issued on sm_100a?
//   (1) tcgen05.mma kind::tf32, M = 128, N in {16, 32, 64, 128, 256}, A from tensor memory (TS) or shared memory (SS),
//       issued back to back by ONE elected thread (and by two warps at once, separate accumulators): clk per instruction.
//   (2) legacy mma.sync.m16n8k8 tf32 issued by every warp of the SM: instructions per clk per SM.
// Decides the design of the fused small-CNN kernel (MNIST-8: conv2 is a 16-channel GEMM, far below one UMMA tile).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rates mma_rates.cu && ./mma_rates
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW1:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D1;\n\tbra W1;\n\tD1:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint64_t sw128_desc(uint32_t addr) {
  uint64_t d;
  const uint32_t lo = ((addr >> 4) & 0x3FFFu) | (1u << 16);
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(0x40004040u));
  return d;
}
__device__ __forceinline__ uint32_t idesc_tf32(int n, int m) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

// mode: 0 = TS (A in tensor memory), 1 = SS (A in shared memory).  nissue = number of issuing warps (1 or 2).
__global__ void __launch_bounds__(128, 1) umma_rate(int N, int M, int mode, int nissue, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base;            // 128 rows x 128 B = 16 KB
  const uint32_t b_smem = base + 16384;    // 256 rows x 128 B = 32 KB
  const uint32_t bars = base + 16384 + 32768;
  const uint32_t slot = bars + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + 4 * i), "r"(0));
  if (warp == 0) {
    if (lane == 0) { mbar_init(bars, 1); mbar_init(bars + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  long long t0 = 0, t1 = 0;
  if (warp < nissue) {
    const bool leader = elect_one();
    const uint32_t idesc = idesc_tf32(N, M);
    const uint32_t d = tmem + (uint32_t)(warp * 256);       // accumulator columns of this issuer (N <= 256; two issuers: N <= 128 each... columns [0,256) / [256,512) minus A)
    const uint32_t a_t = tmem + 480u;                        // 32 columns of A (garbage) at the top
    const uint64_t bd = sw128_desc(b_smem), ad = sw128_desc(a_smem);
    __syncwarp();
    t0 = clock64();
    if (leader) {
      for (int it = 0; it < iters; it += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          // advance the descriptor's start address by 32 bytes per k-step inside the 128-byte row, as a real loop would
          if (mode == 0) umma_ts(d, a_t + 8u * (u & 3), bd + (uint64_t)(2 * (u & 3)), idesc, 1u);
          else umma_ss(d, ad + (uint64_t)(2 * (u & 3)), bd + (uint64_t)(2 * (u & 3)), idesc, 1u);
        }
      }
      umma_commit(bars + 8u * warp);
    }
    __syncwarp();
    const long long t_issue = clock64();
    mbar_wait(bars + 8u * warp, 0);
    t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[2 * warp] = t_issue - t0; out[2 * warp + 1] = t1 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// legacy path: every warp issues mma.sync m16n8k8 tf32 with NACC independent accumulators
template <int NACC>
__global__ void __launch_bounds__(1024, 1) mma_sync_rate(int iters, long long* out, float* sink) {
  float c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 5, b1 = 11;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  long long* d_out; float* sink;
  cudaMalloc(&d_out, 64); cudaMalloc(&sink, 64);
  long long h[4];
  const int smem = 1024 + 16384 + 32768 + 256;
  cudaFuncSetAttribute(umma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4096;
  printf("tcgen05.mma kind::tf32 M=128, %d instructions back to back per issuer, 148 CTAs\n", iters);
  printf("%-4s %-5s %-7s %12s %12s\n", "mode", "N", "issuers", "clk/mma(iss)", "clk/mma(done)");
  for (int mode = 0; mode < 2; ++mode)
    for (int nissue = 1; nissue <= 2; ++nissue)
      for (int N : {16, 32, 48, 64, 128, 256}) {
        if (nissue == 2 && N > 128) continue;
        for (int rep = 0; rep < 2; ++rep) {
          umma_rate<<<148, 128, smem>>>(N, 128, mode, nissue, iters, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d_out, 32, cudaMemcpyDeviceToHost);
        printf("%-4s %-5d %-7d %12.1f %12.1f\n", mode ? "SS" : "TS", N, nissue, (double)h[0] / iters, (double)h[1] / iters);
      }
  printf("\nmma.sync.m16n8k8 tf32, 148 CTAs\n%-6s %-5s %14s %16s\n", "warps", "nacc", "mma/clk/SM", "TFLOP/s@1.9GHz");
  for (int warps : {4, 8, 16, 32}) {
    const int it2 = 2048;
    for (int nacc : {4, 8}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (nacc == 4) mma_sync_rate<4><<<148, warps * 32>>>(it2, d_out, sink);
        else mma_sync_rate<8><<<148, warps * 32>>>(it2, d_out, sink);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, d_out, 8, cudaMemcpyDeviceToHost);
      const double per_clk = (double)warps * it2 * nacc / (double)h[0];
      printf("%-6d %-5d %14.3f %16.1f\n", warps, nacc, per_clk, per_clk * 2.0 * 16 * 8 * 8 * 148 * 1.9e9 / 1e12);
    }
  }
  return 0;
}
