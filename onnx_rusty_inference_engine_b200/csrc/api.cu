// C ABI: library, context, tensors and the ten operator entry points (include/b200rt.h).
// Each operator mirrors one function of src/inference_fp32_ops/*.rs (cited in the header) and launches the
// hand-written kernels of this directory; nothing here computes on the host.
#include <cstdlib>
#include <cstring>

#include "internal.h"

namespace b200 {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ref_geometry(int auto_pad, int H, int W, int kh, int kw, int sh, int sw, const int64_t pads[4], Geo* g) {
  if (sh <= 0 || sw <= 0 || kh <= 0 || kw <= 0) B200_FAIL(B200_EINVAL, "stride/kernel must be positive");
  g->pt = g->pb = g->pl = g->pr = 0;
  switch (auto_pad) {
    case B200_PAD_SAME_UPPER:
    case B200_PAD_SAME_LOWER: {
      // get_padding_size (convolution_op.rs:519-557): usize arithmetic, kernel < stride underflows (panic)
      if (kh < sh || kw < sw) B200_FAIL(B200_EINVAL, "SAME_* with kernel < stride panics in the reference");
      const int ph = (H % sh == 0) ? (kh - sh) : (kh - (H % sh));
      const int pw = (W % sw == 0) ? (kw - sw) : (kw - (W % sw));
      const int t = ph / 2, b = ph - t, l = pw / 2, r = pw - l;
      // the reference returns (bottom, top, right, left) swapped (:547-556): odd extra goes to top/left
      g->pt = b; g->pb = t; g->pl = r; g->pr = l;
      g->Ho = (H + sh - 1) / sh;  // ceil(H / s), convolution_op.rs:297-311
      g->Wo = (W + sw - 1) / sw;
      // the reference's im2col yields (Hp-k)/s+1 rows; it must agree or upstream indexes out of bounds
      if ((H + ph - kh) / sh + 1 != g->Ho || (W + pw - kw) / sw + 1 != g->Wo)
        B200_FAIL(B200_EINVAL, "SAME_* geometry inconsistent in the reference for this shape");
      return 0;
    }
    case B200_PAD_NOTSET: {
      for (int i = 0; i < 4; ++i)
        if (pads[i] < 0) B200_FAIL(B200_EINVAL, "negative pads");
      g->pt = (int)pads[0]; g->pl = (int)pads[1]; g->pb = (int)pads[2]; g->pr = (int)pads[3];
      if (H + g->pt + g->pb < kh || W + g->pl + g->pr < kw) B200_FAIL(B200_EINVAL, "kernel larger than padded input");
      g->Ho = (H - kh + g->pt + g->pb) / sh + 1;
      g->Wo = (W - kw + g->pl + g->pr) / sw + 1;
      return 0;
    }
    case B200_PAD_VALID:
      if (H < kh || W < kw) B200_FAIL(B200_EINVAL, "kernel larger than input");
      g->Ho = (H - kh) / sh + 1;
      g->Wo = (W - kw) / sw + 1;
      return 0;
    default: B200_FAIL(B200_EINVAL, "unknown auto_pad %d", auto_pad);
  }
}

static int alloc_storage(b200_ctx* ctx, size_t bytes, std::shared_ptr<void>* out) {
  void* p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return B200_ENOMEM;
  }
  (void)ctx;
  *out = std::shared_ptr<void>(p, [](void* q) { cudaFree(q); });
  return 0;
}

static int64_t numel_of(const b200_tensor* t) {
  int64_t n = 1;
  for (int i = 0; i < t->rank; ++i) n *= t->dims[i];
  return n;
}

// Physical view for logical dims (see TView in internal.h).
static int make_view(const int64_t* dims, int rank, TView* v) {
  for (int i = 0; i < rank; ++i)
    if (dims[i] < 0 || dims[i] > 0x7fffffff) B200_FAIL(B200_EINVAL, "bad dim %lld", (long long)dims[i]);
  switch (rank) {
    case 4: v->N = (int)dims[0]; v->C = (int)dims[1]; v->H = (int)dims[2]; v->W = (int)dims[3];
            v->ld = v->C >= 3 ? round_up4(v->C) : v->C; break;
    case 3: if (dims[1] != 1 || dims[2] != 1) B200_FAIL(B200_EUNSUPPORTED, "rank-3 tensors must be [C,1,1] (add_op.rs:55-58)");
            v->N = 1; v->C = (int)dims[0]; v->H = v->W = 1; v->ld = v->C; break;
    case 2: v->N = (int)dims[0]; v->C = (int)dims[1]; v->H = v->W = 1; v->ld = v->C; break;
    case 1: v->N = 1; v->C = (int)dims[0]; v->H = v->W = 1; v->ld = v->C; break;
    default: B200_FAIL(B200_EUNSUPPORTED, "rank %d unsupported (utils.rs:146-184 handles 1..4)", rank);
  }
  return 0;
}

static int new_tensor(b200_ctx* ctx, const int64_t* dims, int rank, b200_tensor** out) {
  std::unique_ptr<b200_tensor> t(new b200_tensor());
  t->ctx = ctx;
  t->rank = rank;
  for (int i = 0; i < rank; ++i) t->dims[i] = dims[i];
  B200_TRY(make_view(dims, rank, &t->v));
  const size_t bytes = (size_t)t->v.pixels() * t->v.ld * sizeof(float);
  B200_TRY(alloc_storage(ctx, bytes, &t->storage));
  t->v.p = (float*)t->storage.get();
  ctx_retain(ctx);
  *out = t.release();
  return 0;
}

// *y == NULL: allocate; otherwise verify logical dims.
static int ensure_out(b200_ctx* ctx, b200_tensor** y, const int64_t* dims, int rank) {
  if (!y) B200_FAIL(B200_EINVAL, "output pointer is NULL");
  if (*y == nullptr) return new_tensor(ctx, dims, rank, y);
  if ((*y)->rank != rank) B200_FAIL(B200_EINVAL, "output rank %d != expected %d", (*y)->rank, rank);
  for (int i = 0; i < rank; ++i)
    if ((*y)->dims[i] != dims[i])
      B200_FAIL(B200_EINVAL, "output dim %d is %lld, expected %lld", i, (long long)(*y)->dims[i], (long long)dims[i]);
  (*y)->pad_zeroed = false;
  (*y)->tc.reset();   // about to be overwritten: a cached weight preparation of the old contents is stale
  (*y)->version++;
  return 0;
}

struct Guard {
  b200_ctx* c;
  int prev = -1;
  explicit Guard(b200_ctx* ctx) : c(ctx) {
    c->mu.lock();
    cudaGetDevice(&prev);
    if (prev != c->device) cudaSetDevice(c->device);
  }
  ~Guard() {
    if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
    c->mu.unlock();
  }
};

void ctx_retain(b200_ctx* c) { c->refs.fetch_add(1, std::memory_order_relaxed); }
void ctx_release(b200_ctx* c) {
  if (c->refs.fetch_sub(1, std::memory_order_acq_rel) != 1) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->l2_flush) cudaFree(c->l2_flush);
  if (c->owns_stream) cudaStreamDestroy(c->stream);
  if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
  delete c;
}

int conv_effective_channels(const b200_tensor* x, const b200_tensor* w) {
  const int C = x->v.C;
  if (C % 4 != 0 && C >= 3 && x->pad_zeroed && w->pad_zeroed && x->v.ld == round_up4(C) && w->v.ld == x->v.ld)
    return x->v.ld;  // zero lanes on both operands contribute exactly 0 to every dot product
  return C;
}

}  // namespace b200

using namespace b200;

extern "C" {

const char* b200_last_error(void) { return g_err; }
const char* b200_version(void) { return "b200rt 0.1 (sm_100a)"; }

int b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
  }
  return ok;
}

int b200_ctx_create(int device, void* stream, b200_ctx** out) {
  if (!out) B200_FAIL(B200_EINVAL, "out is NULL");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    B200_FAIL(B200_ENODEVICE, "no CUDA device visible: the B200 backend has no CPU fallback");
  }
  if (device < 0 || device >= n) B200_FAIL(B200_EINVAL, "device %d out of range (0..%d)", device, n - 1);
  cudaDeviceProp p;
  B200_CUDA(cudaGetDeviceProperties(&p, device));
  if (p.major != 10)
    B200_FAIL(B200_ENODEVICE, "device %d is sm_%d%d; this backend is built for sm_100a only", device, p.major, p.minor);
  B200_CUDA(cudaSetDevice(device));
  std::unique_ptr<b200_ctx> c(new b200_ctx());
  c->device = device;
  c->sm_count = p.multiProcessorCount;
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    B200_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->owns_stream = true;
  }
  *out = c.release();
  return 0;
}

int b200_ctx_destroy(b200_ctx* ctx) {
  if (!ctx) return 0;
  {
    // in-flight calls from other threads hold the lock; wait for them, then drain the stream
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (ctx->closed) B200_FAIL(B200_EINVAL, "context destroyed twice");
    ctx->closed = true;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
  }
  ctx_release(ctx);   // tensors / models still alive keep the context until they are freed
  return 0;
}

void* b200_ctx_stream(const b200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int b200_sync(b200_ctx* ctx) {
  if (!ctx) B200_FAIL(B200_EINVAL, "ctx is NULL");
  Guard g(ctx);
  B200_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int64_t b200_ctx_launch_count(const b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------ tensors
int b200_tensor_alloc(b200_ctx* ctx, const int64_t* dims, int rank, b200_tensor** out) {
  if (!ctx || !dims || !out) B200_FAIL(B200_EINVAL, "NULL argument");
  Guard g(ctx);
  return new_tensor(ctx, dims, rank, out);
}

int b200_tensor_upload(b200_tensor* t, const float* host, size_t n) {
  if (!t || !host) B200_FAIL(B200_EINVAL, "NULL argument");
  if ((int64_t)n != numel_of(t)) B200_FAIL(B200_EINVAL, "upload of %zu floats into a tensor of %lld", n, (long long)numel_of(t));
  b200_ctx* ctx = t->ctx;
  Guard g(ctx);
  t->version++;
  t->tc.reset();
  if (n == 0) return 0;
  const TView& v = t->v;
  if (v.dense() && (t->rank != 4 || v.H * v.W == 1 || v.C == 1)) {
    B200_CUDA(cudaMemcpyAsync(v.p, host, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    t->pad_zeroed = !t->is_view;
    return 0;
  }
  float* stage = nullptr;
  B200_CUDA(cudaMalloc(&stage, n * sizeof(float)));
  cudaError_t e = cudaMemcpyAsync(stage, host, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
  int rc = 0;
  if (e == cudaSuccess) {
    // A tensor that owns its rows zero-fills the lanes [C, ld) (the padded 4th lane of a C = 3 input).  A channel
    // view writes exactly its C lanes: [C, ld) of its pitch are the siblings' channels.
    rc = launch_nchw_to_rows(stage, v, /*zero_pad_lanes=*/!t->is_view, ctx->stream);
    ctx->launches++;
    e = cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(stage);
  if (e != cudaSuccess) B200_FAIL(B200_ECUDA, "upload failed: %s", cudaGetErrorString(e));
  if (rc == 0) t->pad_zeroed = !t->is_view;
  return rc;
}

int b200_tensor_download(const b200_tensor* t, float* host, size_t n) {
  if (!t || !host) B200_FAIL(B200_EINVAL, "NULL argument");
  if ((int64_t)n != numel_of(t)) B200_FAIL(B200_EINVAL, "download of %zu floats from a tensor of %lld", n, (long long)numel_of(t));
  b200_ctx* ctx = t->ctx;
  Guard g(ctx);
  if (n == 0) return 0;
  const TView& v = t->v;
  if (v.dense() && (t->rank != 4 || v.H * v.W == 1 || v.C == 1)) {
    B200_CUDA(cudaMemcpyAsync(host, v.p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
  float* stage = nullptr;
  B200_CUDA(cudaMalloc(&stage, n * sizeof(float)));
  int rc = launch_rows_to_nchw(v, stage, ctx->stream);
  ctx->launches++;
  cudaError_t e = cudaMemcpyAsync(host, stage, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(stage);
  if (e != cudaSuccess) B200_FAIL(B200_ECUDA, "download failed: %s", cudaGetErrorString(e));
  return rc;
}

int b200_tensor_rank(const b200_tensor* t) { return t ? t->rank : B200_EINVAL; }

int b200_tensor_dims(const b200_tensor* t, int64_t* dims_out) {
  if (!t || !dims_out) B200_FAIL(B200_EINVAL, "NULL argument");
  for (int i = 0; i < 4; ++i) dims_out[i] = i < t->rank ? t->dims[i] : 1;
  return 0;
}

int b200_tensor_view_channels(b200_tensor* parent, int64_t c_off, int64_t c_len, b200_tensor** out) {
  if (!parent || !out) B200_FAIL(B200_EINVAL, "NULL argument");
  if (parent->rank != 4) B200_FAIL(B200_EINVAL, "channel views need a rank-4 parent");
  if (c_off < 0 || c_len <= 0 || c_off + c_len > parent->dims[1]) B200_FAIL(B200_EINVAL, "channel range out of bounds");
  std::unique_ptr<b200_tensor> t(new b200_tensor());
  t->ctx = parent->ctx;
  t->rank = 4;
  t->dims[0] = parent->dims[0]; t->dims[1] = c_len; t->dims[2] = parent->dims[2]; t->dims[3] = parent->dims[3];
  t->v = parent->v;
  t->v.C = (int)c_len;
  t->v.p = parent->v.p + c_off;
  t->storage = parent->storage;
  t->is_view = true;
  ctx_retain(t->ctx);
  *out = t.release();
  return 0;
}

int b200_tensor_free(b200_tensor* t) {
  if (!t) return 0;
  b200_ctx* ctx = t->ctx;
  {
    Guard g(ctx);
    delete t;
  }
  ctx_release(ctx);
  return 0;
}

// ------------------------------------------------------------------ Conv
static int conv_resolve(const int64_t xd[4], const int64_t wd[4], const b200_conv_params* p, Geo* g, int* auto_pad) {
  if (!p) B200_FAIL(B200_EINVAL, "conv params are NULL");
  if (p->strides[0] <= 0 || p->strides[1] <= 0)
    B200_FAIL(B200_EINVAL, "Conv: `strides` is required (convolution_op.rs:285 unwraps it)");
  if ((p->group != 0 && p->group != 1))
    B200_FAIL(B200_EUNSUPPORTED, "Conv: group=%lld (group>1 is broken upstream, convolution_op.rs:252-258)", (long long)p->group);
  if ((p->dilations[0] > 1) || (p->dilations[1] > 1))
    B200_FAIL(B200_EUNSUPPORTED, "Conv: dilation>1 is broken upstream (convolution_op.rs:590-612)");
  if (xd[1] != wd[1]) B200_FAIL(B200_EINVAL, "Conv: C_in %lld != weight C %lld (convolution_op.rs:252 assert)", (long long)xd[1], (long long)wd[1]);
  int ap = p->auto_pad;
  if (p->pads[0] > 0 || p->pads[1] > 0 || p->pads[2] > 0 || p->pads[3] > 0) ap = B200_PAD_NOTSET;  // :169-173
  *auto_pad = ap;
  return ref_geometry(ap, (int)xd[2], (int)xd[3], (int)wd[2], (int)wd[3], (int)p->strides[0], (int)p->strides[1], p->pads, g);
}

int b200_conv2d_out_dims(const int64_t x_dims[4], const int64_t w_dims[4], const b200_conv_params* p, int64_t y_dims[4]) {
  Geo g; int ap;
  B200_TRY(conv_resolve(x_dims, w_dims, p, &g, &ap));
  y_dims[0] = x_dims[0]; y_dims[1] = w_dims[0]; y_dims[2] = g.Ho; y_dims[3] = g.Wo;
  return 0;
}

int b200_conv2d(b200_ctx* ctx, const b200_tensor* x, const b200_tensor* w, const b200_tensor* bias,
                const b200_tensor* chan_add, const b200_conv_params* p, b200_tensor** y) {
  if (!ctx || !x || !w || !y) B200_FAIL(B200_EINVAL, "NULL argument");
  if (x->rank != 4 || w->rank != 4) B200_FAIL(B200_EINVAL, "Conv operands must be rank 4 (convolution_op.rs:101,111)");
  Geo g; int ap;
  B200_TRY(conv_resolve(x->dims, w->dims, p, &g, &ap));
  const int M = (int)w->dims[0];
  if (bias && (bias->rank != 1 || bias->dims[0] != M)) B200_FAIL(B200_EINVAL, "Conv bias must be [M] (convolution_op.rs:711)");
  if (chan_add && (chan_add->v.numel() != M)) B200_FAIL(B200_EINVAL, "chan_add must hold M values");
  if (!w->v.dense() && !(w->v.ld == round_up4(w->v.C))) B200_FAIL(B200_EINVAL, "weights must be a whole tensor, not a view");
  Guard gd(ctx);
  int64_t yd[4] = {x->dims[0], M, g.Ho, g.Wo};
  B200_TRY(ensure_out(ctx, y, yd, 4));
  ConvArgs a{};
  const int Ceff = conv_effective_channels(x, w);
  a.x = x->v.p; a.N = x->v.N; a.C = Ceff; a.H = x->v.H; a.W = x->v.W; a.ldx = x->v.ld;
  a.w = w->v.p; a.M = M; a.KH = (int)w->dims[2]; a.KW = (int)w->dims[3];
  a.K = a.KH * a.KW * Ceff; a.wc = w->v.ld; a.ldw = a.KH * a.KW * a.wc;
  a.bias = bias ? bias->v.p : nullptr;
  a.chan_add = chan_add ? chan_add->v.p : nullptr;
  a.y = (*y)->v.p; a.Ho = g.Ho; a.Wo = g.Wo; a.ldy = (*y)->v.ld;
  a.sh = (int)p->strides[0]; a.sw = (int)p->strides[1]; a.pt = g.pt; a.pl = g.pl;
  a.relu = p->fuse_relu ? 1 : 0;
  // B200_CONV_PATH=1 forces the CUDA-core cross-check kernel for the per-op entry point (tests / debugging)
  static const int forced_path = [] { const char* e = getenv("B200_CONV_PATH"); return e ? atoi(e) : 0; }();
  if (forced_path != 1 && tc_supported(a) == 0) {
    b200_tensor* wm = const_cast<b200_tensor*>(w);
    if (!wm->tc) { B200_TRY(tc_prepare_weights(a.w, a.M, a.K, ctx->stream, &wm->tc)); ctx->launches++; }
    B200_TRY(launch_conv_tc(a, *wm->tc, ctx->stream));
  } else {
    B200_TRY(launch_conv_simt(a, ctx->stream));
  }
  ctx->launches++;
  return 0;
}

// ------------------------------------------------------------------ MaxPool
static int pool_resolve(const int64_t xd[4], const b200_pool_params* p, Geo* g) {
  if (!p) B200_FAIL(B200_EINVAL, "pool params are NULL");
  if (p->kernel[0] <= 0 || p->kernel[1] <= 0) B200_FAIL(B200_EINVAL, "MaxPool: `kernel_shape` is required (max_pool_op.rs:100)");
  if (p->strides[0] <= 0 || p->strides[1] <= 0) B200_FAIL(B200_EINVAL, "MaxPool: `strides` is required (max_pool_op.rs:207)");
  // no pad promotion here: with auto_pad absent (VALID) the pads attribute is ignored (max_pool_op.rs:88,188)
  return ref_geometry(p->auto_pad, (int)xd[2], (int)xd[3], (int)p->kernel[0], (int)p->kernel[1], (int)p->strides[0],
                      (int)p->strides[1], p->pads, g);
}

int b200_maxpool2d_out_dims(const int64_t x_dims[4], const b200_pool_params* p, int64_t y_dims[4]) {
  Geo g;
  B200_TRY(pool_resolve(x_dims, p, &g));
  y_dims[0] = x_dims[0]; y_dims[1] = x_dims[1]; y_dims[2] = g.Ho; y_dims[3] = g.Wo;
  return 0;
}

int b200_maxpool2d(b200_ctx* ctx, const b200_tensor* x, const b200_pool_params* p, b200_tensor** y) {
  if (!ctx || !x || !y) B200_FAIL(B200_EINVAL, "NULL argument");
  if (x->rank != 4) B200_FAIL(B200_EINVAL, "MaxPool input must be rank 4");
  Geo g;
  B200_TRY(pool_resolve(x->dims, p, &g));
  Guard gd(ctx);
  int64_t yd[4] = {x->dims[0], x->dims[1], g.Ho, g.Wo};
  B200_TRY(ensure_out(ctx, y, yd, 4));
  PoolArgs a{};
  a.x = x->v.p; a.N = x->v.N; a.C = x->v.C; a.H = x->v.H; a.W = x->v.W; a.ldx = x->v.ld;
  a.y = (*y)->v.p; a.Ho = g.Ho; a.Wo = g.Wo; a.ldy = (*y)->v.ld;
  a.kh = (int)p->kernel[0]; a.kw = (int)p->kernel[1]; a.sh = (int)p->strides[0]; a.sw = (int)p->strides[1];
  a.pt = g.pt; a.pl = g.pl;
  B200_TRY(launch_maxpool(a, ctx->stream));
  ctx->launches++;
  return 0;
}

// ------------------------------------------------------------------ Relu / Add
int b200_relu(b200_ctx* ctx, const b200_tensor* x, b200_tensor** y) {
  if (!ctx || !x || !y) B200_FAIL(B200_EINVAL, "NULL argument");
  if (x->rank != 4 && x->rank != 2) B200_FAIL(B200_EINVAL, "Relu input must be rank 4 (relu_op.rs:16) or rank 2");
  Guard gd(ctx);
  B200_TRY(ensure_out(ctx, y, x->dims, x->rank));
  B200_TRY(launch_relu(x->v, (*y)->v, ctx->stream));
  ctx->launches++;
  return 0;
}

int b200_add(b200_ctx* ctx, const b200_tensor* x, const b200_tensor* b, b200_tensor** y) {
  if (!ctx || !x || !b || !y) B200_FAIL(B200_EINVAL, "NULL argument");
  Guard gd(ctx);
  if (x->rank == 4) {
    if (b->rank != 3 || b->dims[0] != x->dims[1])
      B200_FAIL(B200_EINVAL, "Add: rank-4 input needs a [C,1,1] second operand (add_op.rs:55-58,75)");
    B200_TRY(ensure_out(ctx, y, x->dims, 4));
    B200_TRY(launch_add_channel(x->v, b->v.p, (*y)->v, ctx->stream));
  } else if (x->rank == 2) {
    if (b->rank != 2 || b->dims[1] != x->dims[1] || (b->dims[0] != x->dims[0] && b->dims[0] != 1))
      B200_FAIL(B200_EINVAL, "Add: rank-2 operands must have the same shape (add_op.rs:84)");
    B200_TRY(ensure_out(ctx, y, x->dims, 2));
    B200_TRY(launch_add_same(x->v, b->v, (*y)->v, ctx->stream));
  } else {
    B200_FAIL(B200_EUNSUPPORTED, "Add: input rank %d (the reference handles rank 4 and rank 2)", x->rank);
  }
  ctx->launches++;
  return 0;
}

// ------------------------------------------------------------------ MatMul (+ fused Add of [1,N])
int b200_matmul(b200_ctx* ctx, const b200_tensor* a, const b200_tensor* b, const b200_tensor* bias, b200_tensor** y) {
  if (!ctx || !a || !b || !y) B200_FAIL(B200_EINVAL, "NULL argument");
  if (a->rank != 2 || b->rank != 2) B200_FAIL(B200_EINVAL, "MatMul operands must be rank 2 (mul_op.rs:16-19)");
  if (a->dims[1] != b->dims[0]) B200_FAIL(B200_EINVAL, "MatMul: inner dims %lld vs %lld", (long long)a->dims[1], (long long)b->dims[0]);
  const int R = (int)a->dims[0], K = (int)a->dims[1], N = (int)b->dims[1];
  if (bias && bias->v.numel() != N) B200_FAIL(B200_EINVAL, "MatMul bias must hold N values");
  Guard gd(ctx);
  int64_t yd[2] = {R, N};
  B200_TRY(ensure_out(ctx, y, yd, 2));
  // B^T [N][K] so that both operands are K-contiguous; the GEMM is a pointwise "convolution" over R pixels on the
  // convolution's tcgen05 path (TMA-fed A tiles, one channel tile of N columns).  The split / permuted weight tiles are
  // cached on b (a rank-2 tensor is never a Conv weight) and rebuilt when b is uploaded again.
  ConvArgs c{};
  c.x = a->v.p; c.N = R; c.C = K; c.H = 1; c.W = 1; c.ldx = a->v.ld;
  c.w = nullptr; c.M = N; c.KH = 1; c.KW = 1; c.K = K; c.ldw = K; c.wc = K;
  c.bias = bias ? bias->v.p : nullptr; c.chan_add = nullptr;
  c.y = (*y)->v.p; c.Ho = 1; c.Wo = 1; c.ldy = (*y)->v.ld;
  c.sh = c.sw = 1; c.pt = c.pl = 0; c.relu = 0;
  static const int forced_path = [] { const char* e = getenv("B200_CONV_PATH"); return e ? atoi(e) : 0; }();
  const bool use_tc = forced_path != 1 && tc_supported(c) == 0;
  b200_tensor* bm = const_cast<b200_tensor*>(b);
  int rc = 0;
  if (!use_tc || !bm->tc) {
    float* bt = nullptr;
    B200_CUDA(cudaMallocAsync((void**)&bt, (size_t)N * K * sizeof(float) + 16, ctx->stream));
    rc = launch_transpose2d(b->v.p, K, N, bt, ctx->stream);
    ctx->launches++;
    c.w = bt;
    if (rc == 0 && use_tc) { rc = tc_prepare_weights(bt, N, K, ctx->stream, &bm->tc); ctx->launches++; }
    if (rc == 0 && !use_tc) { rc = launch_conv_simt(c, ctx->stream); ctx->launches++; }
    cudaFreeAsync(bt, ctx->stream);
  }
  if (rc == 0 && use_tc) { rc = launch_conv_tc(c, *bm->tc, ctx->stream); ctx->launches++; }
  return rc;
}

// ------------------------------------------------------------------ Reshape / Concat / Dropout
int b200_reshape(b200_ctx* ctx, const b200_tensor* x, const int64_t* shape, int n_shape, b200_tensor** y) {
  if (!ctx || !x || !shape || !y) B200_FAIL(B200_EINVAL, "NULL argument");
  if (n_shape < 2) B200_FAIL(B200_EINVAL, "Reshape: shape needs >= 2 entries (reshape_op.rs:87)");
  if (x->rank != 4 && x->rank != 2) B200_FAIL(B200_EINVAL, "Reshape data must be rank 4 (reshape_op.rs:27,30)");
  int64_t d[2];
  for (int i = 0; i < 2; ++i) {
    d[i] = shape[i];
    if (d[i] == 0) d[i] = x->dims[i];  // reshape_op.rs:69-83
    if (d[i] < 0) B200_FAIL(B200_EUNSUPPORTED, "Reshape: -1 is not supported by the reference (reshape_op.rs:87)");
  }
  if (d[0] * d[1] != numel_of(x)) B200_FAIL(B200_EINVAL, "Reshape: %lldx%lld != %lld elements (reshape_op.rs:90)", (long long)d[0], (long long)d[1], (long long)numel_of(x));
  Guard gd(ctx);
  const TView& v = x->v;
  const bool order_free = v.dense() && (v.H * v.W == 1 || v.C == 1);  // NCHW order == physical order
  if (*y == nullptr && order_free) {
    std::unique_ptr<b200_tensor> t(new b200_tensor());
    t->ctx = ctx; t->rank = 2; t->dims[0] = d[0]; t->dims[1] = d[1];
    t->v.p = v.p; t->v.N = (int)d[0]; t->v.C = (int)d[1]; t->v.H = t->v.W = 1; t->v.ld = (int)d[1];
    t->storage = x->storage;  // zero-copy alias
    ctx_retain(ctx);
    *y = t.release();
    return 0;
  }
  B200_TRY(ensure_out(ctx, y, d, 2));
  if (!(*y)->v.dense()) B200_FAIL(B200_EINVAL, "Reshape output must be dense");
  if (order_free) {
    B200_CUDA(cudaMemcpyAsync((*y)->v.p, v.p, (size_t)numel_of(x) * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  } else {
    B200_TRY(launch_rows_to_nchw(v, (*y)->v.p, ctx->stream));
    ctx->launches++;
  }
  return 0;
}

int b200_concat(b200_ctx* ctx, const b200_tensor* a, const b200_tensor* b, int64_t axis, b200_tensor** y) {
  if (!ctx || !a || !b || !y) B200_FAIL(B200_EINVAL, "NULL argument");
  if (a->rank != 4 || b->rank != 4) B200_FAIL(B200_EINVAL, "Concat inputs must be rank 4 (concatenate_op.rs:15-18)");
  if (axis < 0 || axis > 3) B200_FAIL(B200_EINVAL, "Concat: axis %lld out of range for rank 4 (concatenate_op.rs:31 Axis(axis as usize))", (long long)axis);
  for (int d = 0; d < 4; ++d)
    if (d != axis && a->dims[d] != b->dims[d]) B200_FAIL(B200_EINVAL, "Concat: non-axis dims differ (ndarray::concatenate returns Err, unwrapped at concatenate_op.rs:32)");
  Guard gd(ctx);
  int64_t yd[4] = {a->dims[0], a->dims[1], a->dims[2], a->dims[3]};
  yd[axis] += b->dims[axis];
  B200_TRY(ensure_out(ctx, y, yd, 4));
  if (axis == 1) {
    // channels: the inputs may already BE the matching channel views of *y (zero-copy Concat)
    TView ya = (*y)->v; ya.C = a->v.C;
    TView yb = (*y)->v; yb.C = b->v.C; yb.p += a->v.C;
    if (!(a->v.p == ya.p && a->v.ld == ya.ld)) { B200_TRY(launch_copy_rows(a->v, ya, ctx->stream)); ctx->launches++; }
    if (!(b->v.p == yb.p && b->v.ld == yb.ld)) { B200_TRY(launch_copy_rows(b->v, yb, ctx->stream)); ctx->launches++; }
  } else {
    // batch / rows / columns: two block copies in the channels-last layout
    B200_TRY(launch_copy_block(a->v, (*y)->v, 0, 0, 0, ctx->stream));
    B200_TRY(launch_copy_block(b->v, (*y)->v, axis == 0 ? a->v.N : 0, axis == 2 ? a->v.H : 0, axis == 3 ? a->v.W : 0, ctx->stream));
    ctx->launches += 2;
  }
  return 0;
}

int b200_dropout(b200_ctx* ctx, const b200_tensor* x, float ratio, b200_tensor** y) {
  (void)ratio;  // inference mode: identity (dropout_op.rs:66-71)
  if (!ctx || !x || !y) B200_FAIL(B200_EINVAL, "NULL argument");
  if (x->rank != 4) B200_FAIL(B200_EINVAL, "Dropout input must be rank 4 (dropout_op.rs:16)");
  Guard gd(ctx);
  if (*y == nullptr) {
    std::unique_ptr<b200_tensor> t(new b200_tensor(*x));  // alias: shares storage
    t->tc.reset();
    ctx_retain(ctx);
    *y = t.release();
    return 0;
  }
  B200_TRY(ensure_out(ctx, y, x->dims, 4));
  B200_TRY(launch_copy_rows(x->v, (*y)->v, ctx->stream));
  ctx->launches++;
  return 0;
}

// ------------------------------------------------------------------ GlobalAveragePool / Softmax
int b200_global_avgpool(b200_ctx* ctx, const b200_tensor* x, b200_tensor** y) {
  if (!ctx || !x || !y) B200_FAIL(B200_EINVAL, "NULL argument");
  if (x->rank != 4) B200_FAIL(B200_EINVAL, "GlobalAveragePool input must be rank 4");
  Guard gd(ctx);
  int64_t yd[4] = {x->dims[0], x->dims[1], 1, 1};
  B200_TRY(ensure_out(ctx, y, yd, 4));
  B200_TRY(launch_global_avgpool(x->v, (*y)->v, ctx->stream));
  ctx->launches++;
  return 0;
}

int b200_softmax(b200_ctx* ctx, const b200_tensor* x, b200_tensor** y) {
  if (!ctx || !x || !y) B200_FAIL(B200_EINVAL, "NULL argument");
  if (x->rank != 4) B200_FAIL(B200_EINVAL, "Softmax input must be rank 4 (softmax_op.rs:18)");
  Guard gd(ctx);
  int64_t yd[2] = {x->dims[0], x->dims[1] * x->dims[2] * x->dims[3]};
  B200_TRY(ensure_out(ctx, y, yd, 2));
  if (!(*y)->v.dense()) B200_FAIL(B200_EINVAL, "Softmax output must be dense");
  B200_TRY(launch_softmax(x->v, (*y)->v.p, ctx->stream));
  ctx->launches++;
  return 0;
}

int b200_tensorproto_read(const uint8_t* bytes, size_t len, float* out, size_t cap, int64_t* dims_out, int* rank_out,
                          size_t* n);

}  // extern "C"
