// Internal declarations shared by the translation units of libb200rt.so.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200rt.h"

namespace b200 {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
#define B200_FAIL(code, ...)        \
  do {                              \
    ::b200::set_error(__VA_ARGS__); \
    return (code);                  \
  } while (0)
#define B200_CUDA(expr)                                                                        \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      ::b200::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, \
                        cudaGetErrorString(e__));                                              \
      return B200_ECUDA;                                                                       \
    }                                                                                          \
  } while (0)
#define B200_TRY(expr)          \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != 0) return rc__; \
  } while (0)

// ---------------------------------------------------------------- device views
// Physical layout of every tensor: "channels-last rows".  Element (n, c, h, w) of a rank-4 tensor lives
// at p[((n*H + h)*W + w) * ld + c]; ld >= C is the pixel pitch in floats (ld > C for channel views into a
// Concat result and for channel-padded tensors).  Rank-2 [R,K] is N=R, C=K, H=W=1.
struct TView {
  float* p = nullptr;
  int N = 0, C = 0, H = 1, W = 1;
  int ld = 0;
  long long pixels() const { return (long long)N * H * W; }
  long long numel() const { return pixels() * C; }
  bool dense() const { return ld == C; }
};

static inline int round_up4(int c) { return (c + 3) & ~3; }

// ---------------------------------------------------------------- kernel launch parameter blocks
struct ConvArgs {
  const float* x; int N, C, H, W, ldx;       // C = effective (possibly zero-padded) input channels
  const float* w; int M, KH, KW, K, ldw, wc; // weights [M][KH][KW][wc >= C]; K = KH*KW*C, row pitch ldw = KH*KW*wc
  const float* bias;                          // [M] or null
  const float* chan_add;                      // [M] or null (folded Add of a [M,1,1] initializer)
  float* y; int Ho, Wo, ldy;
  int sh, sw, pt, pl;                         // strides, top/left zero padding
  int relu;
  // Walk the output tiles from the last to the first.  The graph executor alternates the direction from launch to
  // launch: a launch then starts on the part of its input that its predecessor wrote last and that is still in the
  // 126 MB L2.  Results do not depend on it (tiles are independent).
  int reverse;
  // Fire expand fusion (planner): the K axis may list the taps in another order -- nibble i of tap_perm = the tap (r * KW + s)
  // at K position i, 0 = natural order -- so that the centre tap comes first; the first skip_m filters (the 1x1 branch) are
  // exact zeros outside the first skip_kb k-blocks, and the tcgen05 kernel does not issue their columns there.
  unsigned long long tap_perm;
  int skip_m, skip_kb;
  // MaxPool 3x3 / stride 2 (zero-fill padding pool_pt / pool_pl, the reference's) fused in front of a pointwise convolution
  // (planner): x, H, W describe the tensor BEFORE the pool, Ho x Wo the pooled map the 1x1 convolution runs on.  tcgen05
  // kernel only: the window rows arrive by TMA and the converter warps take the maximum on their way to tensor memory.
  int pool, pool_pt, pool_pl;
  // Programmatic dependent launch (planner, tcgen05 kernel): the launch may begin while its predecessor on the stream is
  // still running -- its CTAs take SMs as the predecessor's CTAs leave and run their prologue (barriers, tensor-memory
  // allocation, decode tables, first weight tiles); every access to activations comes after griddepcontrol.wait.
  int pdl;
};

struct PoolArgs {
  const float* x; int N, C, H, W, ldx;
  float* y; int Ho, Wo, ldy;
  int kh, kw, sh, sw, pt, pl;                 // zero-fill padding (reference semantics)
  int reverse;                                // as ConvArgs::reverse (strip kernel only)
};

// ---------------------------------------------------------------- kernel launchers (all async on `st`)
// layout.cu
int launch_nchw_to_rows(const float* src_nchw, TView dst, bool zero_pad_lanes, cudaStream_t st);
int launch_rows_to_nchw(TView src, float* dst_nchw, cudaStream_t st);
// NCHW [N,C,H,W] (H, W even) -> 2x2 space-to-depth rows [N, H/2, W/2, 4*C], channel (dy*2+dx)*C + c
// nonfinite (optional): device int set to 1 when the input holds an Inf / NaN (see b200_model's finite guard)
int launch_nchw_to_s2d(const float* src_nchw, int N, int C, int H, int W, float* dst, cudaStream_t st, int* nonfinite = nullptr);
int launch_nonfinite_scan(const float* p, size_t n, int* flag, cudaStream_t st);
int launch_copy_rows(TView src, TView dst, cudaStream_t st);          // same N,C,H,W; pitches may differ
int launch_copy_block(TView src, TView dst, int n0, int h0, int w0, cudaStream_t st);   // src into dst at pixel offset (n0, h0, w0)
int launch_transpose2d(const float* src, int R, int C, float* dst, cudaStream_t st);  // dst[c][r] = src[r][c]
// bandwidth_ops.cu
int launch_relu(TView x, TView y, cudaStream_t st);
int launch_add_channel(TView x, const float* b, TView y, cudaStream_t st);   // y = x + b[c]
int launch_add_same(TView x, TView b, TView y, cudaStream_t st);            // b.N == x.N or b.N == 1
int launch_maxpool(const PoolArgs& a, cudaStream_t st);
int launch_global_avgpool(TView x, TView y, cudaStream_t st);               // y: [N,C,1,1]
int launch_softmax(TView x, float* y /* dense [N, C*H*W] */, cudaStream_t st);
int launch_gap_softmax(TView x, float* y /* dense [N, C] */, cudaStream_t st);
int launch_fill_zero(float* p, size_t n, cudaStream_t st);
// conv_simt.cu -- CUDA-core fp32 implicit GEMM (cross-check path and fallback for shapes tcgen05 does not take)
int launch_conv_simt(const ConvArgs& a, cudaStream_t st);
// conv_tc.cu -- tcgen05 3xTF32 implicit GEMM.  Returns B200_EUNSUPPORTED when the shape is not eligible.
struct TcWeights;  // pre-split (hi, lo) weight tiles + TMA descriptor, built once per weight tensor
int tc_supported(const ConvArgs& a);
// natural_k: keep K in its natural order (producers that write whole rows with tcgen05.st.32x32b, mnist8_fused.cu) instead of
// the 16-group permutation of tcgen05.st.16x256b (conv_tc.cu)
int tc_prepare_weights(const float* w_dev, int M, int K, cudaStream_t st, std::shared_ptr<TcWeights>* out, bool natural_k = false);
int launch_conv_tc(const ConvArgs& a, const TcWeights& w, cudaStream_t st);

// mnist8_fused.cu -- the MNIST-8 graph in two launches (config 5: small-kernel regime)
size_t mnist8_p1_floats(int N);   // floats of the zero-haloed stem output [N][18][18][8] (+ 16 B per image), N rounded up to 8
struct Mnist8StemConsts { float w[25][8]; float bias[8]; float add[8]; };   // stem weights tap-major, Conv bias, folded Add (host copies: kernel parameters)
int launch_mnist8_stem(const float* x, const Mnist8StemConsts& k, float* p1, int N, cudaStream_t st, int* nonfinite = nullptr);
int launch_mnist8_head(const float* p1, const TcWeights& w2, const float* bias2, const float* add2, const float* wm, const float* bm,
                       float* out, int N, cudaStream_t st);

// the same graph in ONE launch: stem warps (CUDA cores) one group ahead of the head (tcgen05) inside each persistent CTA
int launch_mnist8_onepass(const float* x, const Mnist8StemConsts& k, const TcWeights& w2, const float* bias2, const float* add2, const float* wm,
                          const float* bm, float* out, int N, cudaStream_t st, int* nonfinite = nullptr);

// ---------------------------------------------------------------- geometry with the reference's quirks
struct Geo { int Ho, Wo, pt, pb, pl, pr; };
// conv2d / max_pool2d output dims and effective zero padding (convolution_op.rs:293-324,:519-557;
// max_pool_op.rs:215-246,:363-401).  auto_pad already resolved (Conv pad promotion done by the caller).
int ref_geometry(int auto_pad, int H, int W, int kh, int kw, int sh, int sw, const int64_t pads[4], Geo* g);

}  // namespace b200

// ---------------------------------------------------------------- opaque ABI types
struct b200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  std::recursive_mutex mu;
  int64_t launches = 0;
  int sm_count = 0;
  float* l2_flush = nullptr;  // lazily allocated 256 MiB scratch for profile runs
  // Tensors and models keep their context alive: b200_ctx_destroy only marks the context closed and drops the
  // caller's reference; the CUDA resources go when the last handle created on it is freed.
  std::atomic<int> refs{1};
  bool closed = false;
};
namespace b200 {
void ctx_retain(b200_ctx* c);
void ctx_release(b200_ctx* c);   // frees the context when the last reference goes
}

struct b200_tensor {
  b200_ctx* ctx = nullptr;
  int rank = 0;
  int64_t dims[4] = {0, 0, 0, 0};
  b200::TView v;                       // physical view
  std::shared_ptr<void> storage;       // owning allocation (shared with views / aliases)
  bool pad_zeroed = false;             // lanes [C, ld) are known to be zero (set by upload; never for views)
  bool is_view = false;                // channel view of a wider tensor: lanes [C, ld) belong to the siblings
  std::shared_ptr<b200::TcWeights> tc; // cached tcgen05 weight preparation (when used as Conv weights)
  uint64_t version = 0;                // bumped by upload; invalidates `tc`
};
