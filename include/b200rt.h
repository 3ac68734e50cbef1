/*
 * b200rt.h -- C ABI of the B200-native fp32 operator backend (libb200rt.so).
 *
 * This is the drop-in boundary for the hot path of jackperlo/onnx-rusty-inference-engine:
 *   inference()      src/inference_engine/model_inference.rs:29
 *   node_inference() src/inference_engine/model_inference.rs:128 (the op-dispatch match, :138-161)
 *   the ten operators in src/inference_fp32_ops/ (one .rs file each)
 * A Rust `b200rt-sys` crate, the Python entry point
 * (`group17.onnx_make_inference`, lib.rs:14-31) and the C++/Python benches all bind exactly
 * these symbols.  Plain pointers and sizes only; no C++ / torch / CUDA types in any signature
 * (a CUDA stream crosses the boundary as `void*`).
 *
 * Conventions
 *   - Every function returns 0 on success and a negative B200_E* code on failure; it never
 *     unwinds or aborts across the boundary.  b200_last_error() returns a thread-local,
 *     human-readable message for the last failure on the calling thread.  Conditions on which
 *     the reference panics (unknown op model_inference.rs:158, unknown attribute
 *     convolution_op.rs:160 / max_pool_op.rs:111 / concatenate_op.rs:27 / dropout_op.rs:25,
 *     missing `strides` convolution_op.rs:285, ...) are reported as B200_EUNSUPPORTED /
 *     B200_EINVAL instead.
 *   - There is NO CPU fallback: without a usable sm_100 device b200_ctx_create fails.
 *   - Tensors are fp32.  LOGICAL dims follow the reference: rank 4 = NCHW (the store's Array4
 *     slot), rank 2 = [rows, cols] (the Array2 slot), rank 1 / rank 3 for bias-like initializers.
 *     Host buffers passed to upload/download are dense row-major in the logical order, exactly
 *     what the reference's ndarray holds.  The PHYSICAL layout in HBM is the backend's business
 *     (rank-4 tensors are stored channels-last so that Concat is a zero-copy channel-offset view
 *     and Conv is a K-contiguous implicit GEMM).
 *   - All work on a context is ordered on that context's CUDA stream.  A context may be used from
 *     several host threads (calls are serialised by an internal lock), mirroring the reference's
 *     mutex-guarded store (model_inference.rs:30-32).
 */
#ifndef B200RT_H
#define B200RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_EINVAL (-1)       /* bad argument / shape mismatch (reference: unwrap / assert panics) */
#define B200_EUNSUPPORTED (-2) /* op / attribute / rank outside the reference's supported set */
#define B200_ECUDA (-3)        /* CUDA runtime or driver error */
#define B200_ENOMEM (-4)
#define B200_EPARSE (-5)       /* malformed ONNX / TensorProto bytes */
#define B200_ENODEVICE (-6)    /* no sm_100 GPU: the backend refuses to run (no CPU fallback) */

/* auto_pad values.  Conv spells the explicit mode "NOT_SET" (convolution_op.rs:143), MaxPool spells it
 * "NOTSET" (max_pool_op.rs:96); both map to B200_PAD_NOTSET here. */
#define B200_PAD_VALID 0
#define B200_PAD_SAME_UPPER 1
#define B200_PAD_SAME_LOWER 2
#define B200_PAD_NOTSET 3

typedef struct b200_ctx b200_ctx;
typedef struct b200_tensor b200_tensor;
typedef struct b200_model b200_model;

/* ------------------------------------------------------------------ library / context */
const char *b200_last_error(void);
const char *b200_version(void);
/* Number of visible CUDA devices that are sm_100; 0 means the backend cannot run. */
int b200_device_count(void);

/* One context per GPU (replaces the reference's process-wide store + branch threads,
 * model_inference.rs:30-37).  `stream` may be NULL (the context creates its own) or an existing
 * cudaStream_t passed as void* (e.g. torch.cuda.current_stream().cuda_stream). */
int b200_ctx_create(int device, void *stream, b200_ctx **out);
int b200_ctx_destroy(b200_ctx *ctx);
int b200_sync(b200_ctx *ctx);
/* The context's cudaStream_t as void* (its own, or the one passed to b200_ctx_create): lets a host that already owns
 * CUDA work order it against the backend's with events. */
void *b200_ctx_stream(const b200_ctx *ctx);
/* Number of backend kernels launched on this context since creation (CUDA-graph replays count the
 * kernels inside the graph). */
int64_t b200_ctx_launch_count(const b200_ctx *ctx);

/* ------------------------------------------------------------------ tensors (the store's values) */
/* dims: rank 1..4.  Contents are uninitialised. */
int b200_tensor_alloc(b200_ctx *ctx, const int64_t *dims, int rank, b200_tensor **out);
/* host: dense row-major in logical order, n == product(dims) floats. */
int b200_tensor_upload(b200_tensor *t, const float *host, size_t n);
int b200_tensor_download(const b200_tensor *t, float *host, size_t n);
int b200_tensor_rank(const b200_tensor *t);
int b200_tensor_dims(const b200_tensor *t, int64_t *dims_out /* >= 4 entries */);
/* Zero-copy view of channels [c_off, c_off+c_len) of a rank-4 tensor (Concat as views,
 * replacing ndarray::concatenate, concatenate_op.rs:31).  The view keeps the parent alive. */
int b200_tensor_view_channels(b200_tensor *parent, int64_t c_off, int64_t c_len, b200_tensor **out);
int b200_tensor_free(b200_tensor *t);

/* ------------------------------------------------------------------ operators
 * One entry point per reference operator.  `y` is in/out: if *y is NULL the op allocates the output
 * (caller frees it); otherwise *y must already have the right logical dims (it may be a channel view,
 * which is how a producer writes straight into a Concat result). */

typedef struct b200_conv_params {
  int64_t strides[2];   /* required, like the reference (convolution_op.rs:285) */
  int64_t pads[4];      /* ONNX order [h_begin, w_begin, h_end, w_end] (convolution_op.rs:267-278) */
  int64_t dilations[2]; /* must be {1,1} (or {0,0} = absent); >1 is broken upstream, rejected here */
  int64_t group;        /* must be 1 (or 0 = absent) */
  int32_t auto_pad;     /* B200_PAD_*; any pad > 0 promotes VALID to NOTSET (convolution_op.rs:169-173) */
  int32_t fuse_relu;    /* apply Relu in the epilogue (relu_op.rs:31-33 folded into Conv) */
} b200_conv_params;

/* convolution(), convolution_op.rs:94 / conv2d :224.  x [N,C,H,W], w [M,C,kH,kW], bias [M] or NULL,
 * chan_add [M,1,1] or NULL (the MNIST `Add` of a per-channel initializer, add_op.rs:75, folded in). */
int b200_conv2d(b200_ctx *ctx, const b200_tensor *x, const b200_tensor *w, const b200_tensor *bias,
                const b200_tensor *chan_add, const b200_conv_params *p, b200_tensor **y);
int b200_conv2d_out_dims(const int64_t x_dims[4], const int64_t w_dims[4], const b200_conv_params *p,
                         int64_t y_dims[4]);

typedef struct b200_pool_params {
  int64_t kernel[2];  /* required (max_pool_op.rs:100) */
  int64_t strides[2]; /* required (max_pool_op.rs:207) */
  int64_t pads[4];    /* honoured ONLY when auto_pad == NOTSET (max_pool_op.rs:188-201) */
  int32_t auto_pad;   /* default VALID (max_pool_op.rs:88); no pad promotion */
  int32_t reserved;
} b200_pool_params;

/* max_pool(), max_pool_op.rs:65 / max_pool2d :157: zero-fill padding, fold from -FLT_MAX. */
int b200_maxpool2d(b200_ctx *ctx, const b200_tensor *x, const b200_pool_params *p, b200_tensor **y);
int b200_maxpool2d_out_dims(const int64_t x_dims[4], const b200_pool_params *p, int64_t y_dims[4]);

/* relu(), relu_op.rs:11 (rank 4 only upstream; rank 2 accepted here as well). */
int b200_relu(b200_ctx *ctx, const b200_tensor *x, b200_tensor **y);
/* add(), add_op.rs:16: rank-4 x + rank-3 [C,1,1] b (add_op.rs:75) or rank-2 x + rank-2 b of the same
 * shape (add_op.rs:84; the batch-N extension also accepts b = [1,K] broadcast over rows). */
int b200_add(b200_ctx *ctx, const b200_tensor *x, const b200_tensor *b, b200_tensor **y);
/* mul() (ONNX MatMul), mul_op.rs:11: a [R,K] . b [K,N]; bias [1,N] or NULL fuses the following Add. */
int b200_matmul(b200_ctx *ctx, const b200_tensor *a, const b200_tensor *b, const b200_tensor *bias,
                b200_tensor **y);
/* reshape(), reshape_op.rs:16: rank-4 (or rank-2) data -> rank 2 (d0, d1); 0 copies the input dim;
 * elements keep the reference's NCHW memory order (reshape_op.rs:89). */
int b200_reshape(b200_ctx *ctx, const b200_tensor *x, const int64_t *shape, int n_shape, b200_tensor **y);
/* concatenation(), concatenate_op.rs:11: exactly two rank-4 inputs.  Copies only when an input is not
 * already the matching channel view of *y. */
int b200_concat(b200_ctx *ctx, const b200_tensor *a, const b200_tensor *b, int64_t axis, b200_tensor **y);
/* drop_out(), dropout_op.rs:12: identity at inference; *y becomes a zero-copy alias when NULL. */
int b200_dropout(b200_ctx *ctx, const b200_tensor *x, float ratio, b200_tensor **y);
/* global_average_pool(), global_average_pool_op.rs:11: [N,C,H,W] -> [N,C,1,1]. */
int b200_global_avgpool(b200_ctx *ctx, const b200_tensor *x, b200_tensor **y);
/* softmax(), softmax_op.rs:13: flatten to (N, C*H*W), axis 1 -> rank-2 [N, C*H*W]. */
int b200_softmax(b200_ctx *ctx, const b200_tensor *x, b200_tensor **y);

/* ------------------------------------------------------------------ graph level
 * inference() (model_inference.rs:29) with the whole walk on the device: weights uploaded once,
 * activations resident in HBM, Conv+Add+Relu / Concat / Dropout / Reshape fused or elided, the node
 * sequence captured in a CUDA graph per batch size. */
int b200_model_load_onnx(b200_ctx *ctx, const uint8_t *bytes, size_t len, b200_model **out);
int b200_model_load_file(b200_ctx *ctx, const char *path, b200_model **out);
int b200_model_free(b200_model *m);
/* Per-image input dims (C,H,W of the single non-initializer graph input) and output element count. */
int b200_model_io(const b200_model *m, int64_t in_chw[3], int64_t *out_per_image);
/* Host-to-host: uploads `batch` images (NCHW, dense), runs, downloads batch*out_per_image floats. */
int b200_model_run(b200_model *m, const float *host_in, int64_t batch, float *host_out);
/* Pipelined host-to-host: enqueues H2D (its own copy stream), the run, and D2H (a second copy stream) and returns;
 * two batches may be in flight, so the H2D of batch i+1 overlaps the compute of batch i.  host_in / host_out should
 * be pinned and must stay valid until b200_model_sync returns. */
int b200_model_run_async(b200_model *m, const float *host_in, int64_t batch, float *host_out);
int b200_model_sync(b200_model *m);
/* Multi-GPU inference() (model_inference.rs:29 has no counterpart: the reference is single-device): models[i] holds the
 * same ONNX graph on its own context (one context per device).  The batch is split contiguously -- the first
 * batch % n shards get one image more -- each shard goes through b200_model_run_async of its model on its own host
 * thread, and host_out receives the [batch, out_per_image] logits in shard (= image) order.  Weights are replicated;
 * there is no collective on the data path.  Returns the first failing shard's code. */
int b200_model_run_sharded(b200_model *const *models, int n, const float *host_in, int64_t batch, float *host_out);
/* Device-resident: d_in / d_out are device pointers (NCHW dense input, [batch, out_per_image] output)
 * on the context's device; asynchronous on the context's stream. */
int b200_model_run_device(b200_model *m, const float *d_in, int64_t batch, float *d_out);
/* Options: "cuda_graph" (0/1, default 1), "conv_path" (0 = auto, 1 = force CUDA-core fp32 cross-check
 * kernel, 2 = force tcgen05 3xTF32), "fire_fusion" (0/1, default 1: expand1x1 + expand3x3 of a Fire module as one
 * launch when both fit one channel tile), "pool_fusion" (0/1, default 1: a MaxPool 3x3 / 2 whose only consumer is a
 * pointwise convolution with <= 64 filters runs inside that convolution's launch; same bits), "pdl" (0/1, default 1:
 * tcgen05 launches are programmatic dependent launches, their prologue runs under the predecessor's tail), "s2d" (0/1, default 1: a stride-2 stem convolution runs on a 2x2
 * space-to-depth copy of the graph input), "alt_order" (0/1, default 1: launches walk their tiles in alternating
 * directions so that each starts on what its predecessor left in the L2), "fused_cnn" (0 / 1 / 2, default 2: the MNIST-8
 * graph, when the graph matches, node by node / as two fused launches / as one launch), "finite_guard" (0/1, default 1: see below), "verbose" (0/1: print the reference's
 * per-node lines).
 * Finite guard: the tensor-core path splits every value into hi + lo (Inf - Inf = NaN) and some fusions add 0 * x terms, so
 * it reproduces the reference for FINITE inputs; the reference itself keeps an Inf / NaN local to the outputs that really
 * read it (convolution_op.rs:480).  Every run's input stage therefore checks the input: b200_model_run reruns a batch that
 * holds an Inf / NaN on the CUDA-core fp32 plan (the reference's semantics, slower) by itself; b200_model_sync returns
 * B200_EUNSUPPORTED when a batch run through the asynchronous entry points since the last sync held one. */
int b200_model_set_option(b200_model *m, const char *key, int64_t value);
/* Per-launch profile of the last planned batch size: runs each planned launch `iters` times between CUDA
 * events and writes one JSON document into buf.  flush_l2: 0 = repeat each launch back to back, 1 = flush the L2
 * before every timed launch (cold cache), 2 = run the whole launch list in model order `iters` times and time every
 * launch in place (the cache state of a real run).  Format: [{"name":..,"kind":..,"ms":..,"flops":..,"bytes":..}, ...]. */
int b200_model_profile(b200_model *m, int64_t batch, int iters, int flush_l2, char *buf, size_t cap);
/* HBM held by the activation arena planned for `batch` images: with liveness reuse (activations whose launch ranges are
 * disjoint share bytes) and, for comparison, the sum of all activations (what the store of model_inference.rs:30-32, which
 * frees nothing until inference() returns, would hold). */
int b200_model_arena_bytes(b200_model *m, int64_t batch, int64_t *arena_bytes, int64_t *no_reuse_bytes);
/* Number of kernel launches one b200_model_run_device of `batch` images issues. */
int64_t b200_model_launches_per_run(b200_model *m, int64_t batch);

/* Serialized TensorProto (.pb) reader, replacing read_input_data (main.rs:44-53): returns the element
 * count in *n; copies min(*n, cap) floats into out when out != NULL; dims_out gets up to 8 dims. */
int b200_tensorproto_read(const uint8_t *bytes, size_t len, float *out, size_t cap, int64_t *dims_out,
                          int *rank_out, size_t *n);

#ifdef __cplusplus
}
#endif
#endif /* B200RT_H */
