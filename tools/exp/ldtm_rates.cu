// Micro-benchmark (GPU): how fast does tcgen05.ld drain an accumulator on sm_100a?  The conv kernel's single accumulator
// stage puts the drain (128 lanes x 2*BN columns) between two tiles' MMAs, so: is it bound by tensor memory's read rate
// or by the latency of the loads a warp keeps in flight?
//   W warps (warp w reads lane quarter w % 4), each repeats: DEPTH x tcgen05.ld.32x32b.xX, then tcgen05.wait::ld.
//   Reported: clk per round, clk per instruction, bytes / clk / SM; optionally while another warp issues N = 256 MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_rates ldtm_rates.cu && ./ldtm_rates
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

template <int X> struct Ld;
template <> struct Ld<16> {
  static __device__ __forceinline__ void ld(uint32_t t, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(t) : "memory");
  }
};
template <> struct Ld<32> {
  static __device__ __forceinline__ void ld(uint32_t t, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                   "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(t) : "memory");
  }
};

// W reading warps (of 16), one optional MMA warp (warp 16).  out[0] = max clk over the readers, out[1] = xor of what was read
template <int X, int DEPTH>
__global__ void __launch_bounds__(544, 1) ldtm_rate(int W, int rounds, int with_mma, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t b_smem = base;            // 256 rows x 128 B = 32 KB of zeros
  const uint32_t bars = base + 32768;
  const uint32_t slot = bars + 64;
  __shared__ long long tmax;
  __shared__ int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < 32768 / 4; i += blockDim.x) asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + 4 * i), "r"(0));
  if (threadIdx.x == 0) { tmax = 0; stop = 0; }
  if (warp == 0) {
    if (lane == 0) { mbar_init(bars, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
  uint32_t sink = 0;
  if (warp < W) {
    const uint32_t t0a = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);   // like the epilogue: warps of a quarter take column ranges
    uint32_t r[DEPTH][X];
    const long long t0 = clock64();
    for (int it = 0; it < rounds; ++it) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) Ld<X>::ld(t0a + (uint32_t)(((it * DEPTH + d) * X) & 255), r[d]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int d = 0; d < DEPTH; ++d)
#pragma unroll
        for (int j = 0; j < X; ++j) sink ^= r[d][j];
    }
    const long long t1 = clock64();
    if (lane == 0) atomicMax((unsigned long long*)&tmax, (unsigned long long)(t1 - t0));
    __syncwarp();
    if (lane == 0) atomicAdd(&stop, 1);
  } else if (warp == 16 && with_mma) {
    // N = 256 TS-mode MMAs (A = garbage columns 448.., accumulator columns 256..511 -- the readers stay below 256+64)
    const bool leader = elect_one();
    const uint32_t idesc = idesc_tf32(with_mma, 128);
    const uint64_t bd = sw128_desc(b_smem);
    uint32_t ph = 0;
    while (*(volatile int*)&stop < W) {
      if (leader) {
#pragma unroll
        for (int u = 0; u < 8; ++u) umma_ts(tmem + 256u, tmem + 224u + 8u * (u & 3), bd + (uint64_t)(2 * (u & 3)), idesc, 1u);
        umma_commit(bars);
      }
      __syncwarp();
      mbar_wait(bars, ph);
      ph ^= 1u;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) out[0] = tmax;
  if (sink == 0x12345678u) out[1] = sink;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int X, int DEPTH>
static void run(int W, int with_mma, long long* d_out) {
  const int rounds = 2048 / DEPTH;
  const size_t smem = 32768 + 1024 + 256;
  cudaFuncSetAttribute(ldtm_rate<X, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long h[2] = {0, 0};
  for (int rep = 0; rep < 2; ++rep) {
    ldtm_rate<X, DEPTH><<<1, 544, smem>>>(W, rounds, with_mma, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const double clk_round = (double)h[0] / rounds;
  const double bytes = (double)W * DEPTH * X * 32 * 4;   // per round, all warps
  printf("x%-3d depth %d  warps %-2d  mma N=%-3d  %8.1f clk/round  %7.1f clk/instr/warp  %7.1f B/clk/SM   (a 128 x 256-column drain: %6.0f clk)\n", X, DEPTH, W, with_mma,
         clk_round, clk_round / DEPTH, bytes / clk_round, 131072.0 / (bytes / clk_round));
  fflush(stdout);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  for (int with_mma : {0, 256}) {
    for (int W : {1, 4, 8, 16}) {
      run<16, 1>(W, with_mma, d_out);
      run<16, 2>(W, with_mma, d_out);
      run<16, 4>(W, with_mma, d_out);
      run<32, 1>(W, with_mma, d_out);
      run<32, 2>(W, with_mma, d_out);
    }
  }
  return 0;
}
