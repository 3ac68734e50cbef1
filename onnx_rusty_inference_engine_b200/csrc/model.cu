// Graph-level executor: the device-resident successor of inference() / node_inference()
// (src/inference_engine/model_inference.rs:29-162) and of the branch scheduler in
// src/inference_engine/multithreading/*.rs.
//
// What the reference does per inference          What this file does once per (model, batch size)
//   walk graph.node in file order                  same order, but into a launch plan
//   re-decode every initializer on every use       weights laid out + uploaded once (utils.rs:113-185)
//   clone tensors in and out of a HashMap          activations live in one HBM arena; names -> views
//   Conv, Add([C,1,1]), Relu as three passes       one kernel, Add/Relu in the epilogue
//   Concat copies both inputs                      producers write at a channel offset of the result
//   Dropout / Reshape copy                         aliases (Reshape's NCHW order folded into MatMul's weight)
//   2 worker threads per Fire module               one stream; the plan is replayed as a CUDA graph
// Results are independent of node scheduling in the reference, so a single in-order stream is faithful.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <set>
#include <sstream>
#include <thread>

#include "internal.h"
#include "onnx_wire.h"
#include "tc_common.h"

namespace b200 {
namespace {

struct DevBuf {
  float* p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { if (p) cudaFree(p); }
};

struct Val {  // one graph value for the planned batch
  int rank = 0;
  int64_t dims[4] = {0, 0, 0, 0};  // logical, dims[0] already scaled to the batch
  TView v;
  bool planned = false;      // has a device location
  int alloc = -1;            // arena allocation that backs it (liveness tracking); aliases and channel views share it
  bool pad_zeroed = false;   // lanes [C, ld) are zero
  bool s2d = false;          // graph input stored 2x2 space-to-depth: v is [N, H/2, W/2, 4*C], channel (dy*2+dx)*C + c
  bool is_init = false;      // initializer (constant)
  const WireTensor* init = nullptr;
  std::shared_ptr<std::vector<float>> host2d;  // constant rank-2 value produced by Reshape(initializer)
  // set when this value is a flatten of a channels-last activation whose NCHW order has been
  // folded into the consumer MatMul's weight: rows are in (h, w, c) order
  int perm_C = 0, perm_HW = 0;
  // set when this value is a MaxPool 3x3 / 2 output that is never materialised: its only consumer, a pointwise
  // convolution, pools on the fly (ConvArgs::pool).  pool_src is the pool's input; alloc is the input's allocation
  bool pooled = false;
  TView pool_src;
  int pool_pt = 0, pool_pl = 0;
};

struct Step {
  std::string name, kind;
  double flops = 0, bytes = 0;
  std::function<int(cudaStream_t)> run;
};

struct Plan {
  int64_t batch = 0;
  std::vector<Step> steps;
  std::unique_ptr<DevBuf> arena;
  size_t arena_used = 0;    // bytes of the coloured arena
  size_t arena_bump = 0;    // sum of all allocations (no reuse)
  uint64_t last_used = 0;   // LRU stamp (b200_model::use_counter)
  TView in_view;            // where the input transform writes
  bool in_zero_pad = false;
  bool in_direct = false;   // input needs no transform (C == 1 or H*W == 1): memcpy
  bool in_s2d = false;      // input transform writes the 2x2 space-to-depth layout (stride-2 stem convolution)
  int in_c0 = 0, in_h0 = 0, in_w0 = 0;   // logical C, H, W of the graph input (in_s2d)
  // Fused small-CNN path (mnist8_fused.cu): the first launch reads the caller's NCHW input directly, so it replaces the
  // input transform (its pointer is per call: outside the CUDA graph, like the transform)
  std::function<int(const float*, cudaStream_t, int*)> in_stem;   // (input, stream, non-finite flag or null)
  bool safe = false;        // the finite guard's fallback plan
  std::string fused_name;   // profile entry of in_stem
  double fused_flops = 0, fused_bytes = 0;
  float* out_ptr = nullptr; // dense [batch, out_per_image]
  int64_t out_per_image = 0;
  bool out_needs_nchw = false;
  TView out_view;
  cudaGraphExec_t graph_exec = nullptr;
  ~Plan() { if (graph_exec) cudaGraphExecDestroy(graph_exec); }
};

int64_t attr_i(const WireNode& n, const char* name, int64_t dflt) {
  for (auto& a : n.attr) if (a.name == name) return a.i;
  return dflt;
}

}  // namespace
}  // namespace b200

using namespace b200;

struct b200_model {
  b200_ctx* ctx = nullptr;
  WireModel wm;
  std::string input_name;
  int64_t in_dims[4] = {1, 0, 0, 0};  // static dims of the graph input (dims[0] = the model's batch, normally 1)
  int64_t out_per_image = 0;
  std::map<std::string, std::unique_ptr<DevBuf>> consts;              // device-resident constants by key
  std::map<std::string, std::shared_ptr<TcWeights>> tc_weights;       // tcgen05 weight preparations by key
  std::map<int64_t, std::unique_ptr<Plan>> plans;
  int opt_cuda_graph = 1;
  int opt_conv_path = 0;
  int opt_fire_fusion = 1;
  int opt_pdl = 1;           // tcgen05 launches as programmatic dependent launches (prologue under the predecessor's tail)
  int opt_pool_fusion = 1;   // MaxPool 3x3 / 2 -> pointwise Conv as one tcgen05 launch that never writes the pooled tensor
  int opt_alt_order = 1;
  int opt_s2d = 1;           // stride-2 stem convolution on a space-to-depth copy of the graph input     // alternate the tile walking direction from launch to launch (L2 reuse)   // expand1x1 + expand3x3 of a Fire module as one conv when both fit one channel tile
  int opt_fused_cnn = 2;     // the MNIST-8 graph fused (mnist8_fused.cu) when the graph matches: 2 = one launch, 1 = two launches, 0 = node by node
  int opt_verbose = 0;
  float* stage_in = nullptr;  size_t stage_in_bytes = 0;   // device staging for host-to-host runs
  float* stage_out = nullptr; size_t stage_out_bytes = 0;
  // pipelined host-to-host runs (b200_model_run_async): two slots, copy streams separate from the compute stream
  struct Slot {
    float* in = nullptr; size_t in_bytes = 0;
    float* out = nullptr; size_t out_bytes = 0;
    cudaEvent_t in_ready = nullptr, in_free = nullptr, done = nullptr, out_done = nullptr;
  } slots[2];
  cudaStream_t h2d = nullptr, d2h = nullptr;
  uint64_t seq = 0;
  uint64_t use_counter = 0;   // LRU stamps of `plans`
  // Finite guard.  The split-precision tensor-core path (x = hi + lo: Inf - Inf = NaN) and the zero-weight fusions (Fire
  // expand fusion, space-to-depth stem: 0 * Inf = NaN) are exact for FINITE activations only, where the reference
  // (convolution_op.rs:480 multiplies real taps only) keeps Inf / NaN local.  The input stage of every run raises this
  // device flag when the input holds an Inf / NaN; the synchronous host entry then reruns the batch on the fallback plan
  // (CUDA-core fp32 convolutions, no fusions: the reference's semantics), the asynchronous entries report it at sync.
  int* d_nonfinite = nullptr;
  int* h_nonfinite = nullptr;   // pinned
  int opt_finite_guard = 1;
  ~b200_model() {
    if (stage_in) cudaFree(stage_in);
    if (stage_out) cudaFree(stage_out);
    for (auto& sl : slots) {
      if (sl.in) cudaFree(sl.in);
      if (sl.out) cudaFree(sl.out);
      if (sl.in_ready) cudaEventDestroy(sl.in_ready);
      if (sl.in_free) cudaEventDestroy(sl.in_free);
      if (sl.done) cudaEventDestroy(sl.done);
      if (sl.out_done) cudaEventDestroy(sl.out_done);
    }
    if (h2d) cudaStreamDestroy(h2d);
    if (d2h) cudaStreamDestroy(d2h);
    if (d_nonfinite) cudaFree(d_nonfinite);
    if (h_nonfinite) cudaFreeHost(h_nonfinite);
  }
};

namespace b200 {
namespace {

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

// ---------------------------------------------------------------- constants
int upload_const(b200_model* m, const std::string& key, const std::vector<float>& host, float** out) {
  auto it = m->consts.find(key);
  if (it != m->consts.end()) { *out = it->second->p; return 0; }
  std::unique_ptr<DevBuf> b(new DevBuf());
  b->bytes = std::max<size_t>(host.size() * sizeof(float), 16);
  if (cudaMalloc((void**)&b->p, b->bytes) != cudaSuccess) B200_FAIL(B200_ENOMEM, "cudaMalloc(%zu) for constant %s", b->bytes, key.c_str());
  // on the context's stream (the consumers -- the weight-split kernel, the plan's launches -- run there; the legacy
  // stream a plain cudaMemcpy uses is not ordered against a non-blocking stream), then wait: `host` is pageable and
  // may be freed by the caller
  if (!host.empty()) {
    B200_CUDA(cudaMemcpyAsync(b->p, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, m->ctx->stream));
    B200_CUDA(cudaStreamSynchronize(m->ctx->stream));
  }
  *out = b->p;
  m->consts[key] = std::move(b);
  return 0;
}

// Shape of an initializer as the reference sees it: graph.input dims (utils.rs:122) else TensorProto.dims.
std::vector<int64_t> init_dims(const WireModel& wm, const WireTensor& t) {
  if (const WireValueInfo* vi = wm.find_input(t.name)) {
    bool ok = !vi->dims.empty();
    for (auto d : vi->dims) if (d < 0) ok = false;
    if (ok) return vi->dims;
  }
  return t.dims;
}

struct Planner {
  b200_model* m;
  Plan* plan;
  int64_t B;
  bool safe = false;   // the finite guard's fallback plan: CUDA-core fp32 convolutions, no zero-weight fusions
  int conv_path() const { return safe ? 1 : m->opt_conv_path; }
  std::map<std::string, Val> env;
  std::map<std::string, int> n_consumers;
  std::map<std::string, std::pair<std::string, int>> redirect;  // value -> (concat output, channel offset)
  std::set<size_t> consumed;                                    // node indices folded into an earlier step
  std::map<std::string, std::string> fused_pool_label;          // pooled value that is never materialised -> its MaxPool's name
  // ----- arena with liveness reuse.  The dry pass records every allocation with the launch that
  // first writes it (def) and the last launch that reads it (last); offsets then come from a first-fit interval
  // colouring: two allocations may share bytes iff their [def, last] launch ranges are disjoint (one in-order stream, so
  // a launch never sees a buffer whose last reader has not finished; an output never aliases an input of its own
  // launch).  The second pass hands out the same allocations in the same order at their coloured offsets.
  struct ArenaRec { size_t bytes; int def, last; size_t off; };
  std::vector<ArenaRec> recs;
  size_t alloc_seq = 0;
  size_t arena_cursor = 0;   // dry pass: sum of all allocations (what a bump allocator would need); then the coloured size
  bool dry = true;
  static constexpr int LIVE_FOREVER = 0x7fffffff;
  float* arena_alloc(size_t floats, int* id = nullptr, bool forever = false) {
    const size_t bytes = (floats * sizeof(float) + 255) & ~(size_t)255;
    const size_t i = alloc_seq++;
    if (id) *id = (int)i;
    if (dry) {
      recs.push_back(ArenaRec{bytes, (int)step_counter, forever ? LIVE_FOREVER : (int)step_counter, 0});
      arena_cursor += bytes;
      return nullptr;
    }
    return (float*)((char*)plan->arena->p + recs[i].off);
  }
  void touch(const Val* v) {   // v is read by the launch being planned
    if (dry && v && v->alloc >= 0 && recs[(size_t)v->alloc].last < (int)step_counter) recs[(size_t)v->alloc].last = (int)step_counter;
  }
  void keep_forever(int alloc) { if (dry && alloc >= 0) recs[(size_t)alloc].last = LIVE_FOREVER; }
  size_t colour() {
    size_t top = 0;
    std::vector<size_t> order(recs.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return recs[a].bytes > recs[b].bytes; });   // big first
    std::vector<size_t> placed;
    for (size_t i : order) {
      // candidate offsets: 0 and the end of every conflicting allocation already placed; take the lowest that fits
      std::vector<std::pair<size_t, size_t>> busy;   // [begin, end) of conflicting allocations
      for (size_t j : placed)
        if (!(recs[j].last < recs[i].def || recs[i].last < recs[j].def)) busy.push_back({recs[j].off, recs[j].off + recs[j].bytes});
      std::sort(busy.begin(), busy.end());
      size_t off = 0;
      for (auto& b : busy) {
        if (off + recs[i].bytes <= b.first) break;
        if (b.second > off) off = b.second;
      }
      recs[i].off = off;
      placed.push_back(i);
      top = std::max(top, off + recs[i].bytes);
    }
    return top;
  }

  const WireNode* sole_consumer(const std::string& value, size_t* idx) {
    if (n_consumers[value] != 1) return nullptr;
    if (!m->wm.outputs.empty() && m->wm.outputs[0].name == value) return nullptr;
    for (size_t i = 0; i < m->wm.nodes.size(); ++i)
      for (auto& in : m->wm.nodes[i].input)
        if (in == value) { *idx = i; return &m->wm.nodes[i]; }
    return nullptr;
  }

  Val* get(const std::string& name) {
    auto it = env.find(name);
    if (it != env.end()) { touch(&it->second); return &it->second; }
    const WireTensor* t = m->wm.find_initializer(name);
    if (!t) return nullptr;
    Val v;
    v.is_init = true; v.init = t;
    auto d = init_dims(m->wm, *t);
    v.rank = (int)d.size();
    for (int i = 0; i < v.rank && i < 4; ++i) v.dims[i] = d[i];
    env[name] = v;
    return &env[name];
  }

  // Device location for a rank-4 / rank-2 activation about to be produced.
  int place(const std::string& name, Val* v) {
    auto r = redirect.find(name);
    if (v->rank == 4) {
      v->v.N = (int)v->dims[0]; v->v.C = (int)v->dims[1]; v->v.H = (int)v->dims[2]; v->v.W = (int)v->dims[3];
      if (r != redirect.end()) {
        Val* parent = &env[r->second.first];
        if (!parent->planned) {
          parent->v.N = (int)parent->dims[0]; parent->v.C = (int)parent->dims[1];
          parent->v.H = (int)parent->dims[2]; parent->v.W = (int)parent->dims[3];
          parent->v.ld = parent->v.C;
          parent->v.p = arena_alloc((size_t)parent->v.pixels() * parent->v.ld, &parent->alloc);
          parent->planned = true;
        }
        v->alloc = parent->alloc;
        v->v.ld = parent->v.ld;
        v->v.p = dry ? nullptr : parent->v.p + r->second.second;
      } else {
        v->v.ld = v->v.C;
        v->v.p = arena_alloc((size_t)v->v.pixels() * v->v.ld, &v->alloc);
      }
    } else {
      v->v.N = (int)v->dims[0]; v->v.C = (int)v->dims[1]; v->v.H = v->v.W = 1; v->v.ld = v->v.C;
      v->v.p = arena_alloc((size_t)v->v.numel(), &v->alloc);
    }
    v->planned = true;
    return 0;
  }

  // Launches alternate their walking direction (ConvArgs::reverse): launch k starts on what launch k-1 wrote last,
  // which is still in L2.  The Fire pattern keeps the alternation meaningful: squeeze F, expand R (, expand F), next
  // squeeze the opposite of the last expand.
  size_t step_counter = 0;
  int next_reverse() const { return m->opt_alt_order ? (int)(step_counter & 1) : 0; }
  int use_pdl() const { return m->opt_pdl ? 1 : 0; }
  void add_step(const std::string& name, const char* kind, double flops, double bytes, std::function<int(cudaStream_t)> fn) {
    ++step_counter;
    if (dry) return;
    Step s; s.name = name; s.kind = kind; s.flops = flops; s.bytes = bytes; s.run = std::move(fn);
    plan->steps.push_back(std::move(s));
  }

  // ----- weights
  // Conv weights [M,C,kh,kw] -> device [M][kh][kw][Cp] (K-contiguous rows); Cp > C adds zero lanes.
  int conv_weights(const WireTensor& w, const std::vector<int64_t>& d, int Cp, float** out) {
    const int M = (int)d[0], C = (int)d[1], kh = (int)d[2], kw = (int)d[3];
    std::string key = "convw:" + w.name + ":" + std::to_string(Cp);
    if (m->consts.count(key)) { *out = m->consts[key]->p; return 0; }
    if ((int64_t)w.f32.size() != (int64_t)M * C * kh * kw) B200_FAIL(B200_EINVAL, "initializer %s: %zu floats, dims say %lld", w.name.c_str(), w.f32.size(), (long long)M * C * kh * kw);
    std::vector<float> h((size_t)M * kh * kw * Cp, 0.f);
    for (int mm = 0; mm < M; ++mm)
      for (int c = 0; c < C; ++c)
        for (int r = 0; r < kh; ++r)
          for (int s = 0; s < kw; ++s)
            h[(((size_t)mm * kh + r) * kw + s) * Cp + c] = w.f32[(((size_t)mm * C + c) * kh + r) * kw + s];
    return upload_const(m, key, h, out);
  }
  int vec_const(const WireTensor& t, size_t expect, float** out) {
    if (t.f32.size() != expect) B200_FAIL(B200_EINVAL, "initializer %s: %zu floats, expected %zu", t.name.c_str(), t.f32.size(), expect);
    return upload_const(m, "vec:" + t.name, t.f32, out);
  }

  // ----- shape pre-pass results are written into env as unplanned Vals
  int run();
  int do_conv(size_t i);
  int do_maxpool(size_t i);
  int do_relu(size_t i);
  int do_add(size_t i);
  int do_matmul(size_t i);
  int do_reshape(size_t i);
  int do_concat(size_t i);
  int do_dropout(size_t i);
  int do_gap(size_t i);
  int do_softmax(size_t i);
  int plan_concat_redirects();
  int matmul_weights(const Val& b, int perm_C, int perm_HW, int K, int N, const std::string& key, float** out);
  int try_plan_mnist8(bool* done);
};

int Planner::plan_concat_redirects() {
  // A Concat(axis=1) of two rank-4 activations becomes a zero-copy view: each input is produced
  // directly at its channel offset inside the Concat result.  Needs a shape-only pass first: outputs of
  // Conv/MaxPool/Relu/... are only known while walking, so the offsets are fixed lazily in do_concat's
  // counterpart below; here we only decide eligibility (input produced by a node, used by no other Concat).
  std::map<std::string, int> concat_uses;
  for (auto& n : m->wm.nodes)
    if (n.op_type == "Concat")
      for (auto& in : n.input) concat_uses[in]++;
  for (auto& n : m->wm.nodes) {
    if (n.op_type != "Concat" || n.input.size() != 2 || attr_i(n, "axis", 1) != 1) continue;
    bool ok = n.input[0] != n.input[1];
    for (auto& in : n.input) {
      if (concat_uses[in] != 1) ok = false;
      if (in == m->input_name || m->wm.find_initializer(in)) ok = false;
      // the producer must be an op that writes its own output buffer (aliases such as Dropout do not)
      bool producer_ok = false;
      for (auto& pn : m->wm.nodes)
        if (!pn.output.empty() && pn.output[0] == in)
          producer_ok = pn.op_type == "Conv" || pn.op_type == "Relu" || pn.op_type == "MaxPool" || pn.op_type == "Add";
      if (!producer_ok) ok = false;
    }
    if (!ok) continue;
    // channel offsets are patched once shapes are known (second of the pair set in run())
    redirect[n.input[0]] = {n.output[0], 0};
    redirect[n.input[1]] = {n.output[0], -1};
  }
  return 0;
}

static int attr_check(const WireNode& n, std::initializer_list<const char*> allowed, const char* what) {
  for (auto& a : n.attr) {
    bool ok = false;
    for (auto s : allowed) if (a.name == s) ok = true;
    if (!ok) B200_FAIL(B200_EUNSUPPORTED, "ATTRIBUTE NAME FOR %s NOT FOUND, %s", what, a.name.c_str());
  }
  return 0;
}

static const WireAttr* find_attr(const WireNode& n, const char* name) {
  for (auto& a : n.attr) if (a.name == name) return &a;
  return nullptr;
}

static int parse_conv_attrs(const WireNode& n, b200_conv_params* p) {
  // convolution(), convolution_op.rs:137-163
  B200_TRY(attr_check(n, {"auto_pad", "dilations", "group", "kernel_shape", "pads", "strides"}, "CONVOLUTION"));
  memset(p, 0, sizeof(*p));
  p->auto_pad = B200_PAD_VALID;  // default, :134
  if (auto a = find_attr(n, "auto_pad")) {
    if (a->s == "SAME_UPPER") p->auto_pad = B200_PAD_SAME_UPPER;
    else if (a->s == "SAME_LOWER") p->auto_pad = B200_PAD_SAME_LOWER;
    else if (a->s == "VALID") p->auto_pad = B200_PAD_VALID;
    else if (a->s == "NOT_SET") p->auto_pad = B200_PAD_NOTSET;  // sic, :143
    else B200_FAIL(B200_EUNSUPPORTED, "Convolution Auto Pad specified not found: %s", a->s.c_str());
  }
  if (auto a = find_attr(n, "dilations")) { if (a->ints.size() >= 2) { p->dilations[0] = a->ints[0]; p->dilations[1] = a->ints[1]; } }
  if (auto a = find_attr(n, "group")) p->group = a->i;
  if (auto a = find_attr(n, "pads")) { if (a->ints.size() < 4) B200_FAIL(B200_EINVAL, "Conv pads needs 4 ints"); for (int i = 0; i < 4; ++i) p->pads[i] = a->ints[i]; }
  if (auto a = find_attr(n, "strides")) { if (a->ints.size() >= 2) { p->strides[0] = a->ints[0]; p->strides[1] = a->ints[1]; } }
  return 0;
}

static int parse_pool_attrs(const WireNode& n, b200_pool_params* p) {
  // max_pool(), max_pool_op.rs:88-113
  B200_TRY(attr_check(n, {"auto_pad", "kernel_shape", "pads", "storage_order", "strides"}, "MAX POOL"));
  memset(p, 0, sizeof(*p));
  p->auto_pad = B200_PAD_VALID;
  if (auto a = find_attr(n, "auto_pad")) {
    if (a->s == "SAME_UPPER") p->auto_pad = B200_PAD_SAME_UPPER;
    else if (a->s == "SAME_LOWER") p->auto_pad = B200_PAD_SAME_LOWER;
    else if (a->s == "VALID") p->auto_pad = B200_PAD_VALID;
    else if (a->s == "NOTSET") p->auto_pad = B200_PAD_NOTSET;  // sic, :96
    else B200_FAIL(B200_EUNSUPPORTED, "MaxPool Auto Pad specified not found: %s", a->s.c_str());
  }
  if (auto a = find_attr(n, "kernel_shape")) { if (a->ints.size() >= 2) { p->kernel[0] = a->ints[0]; p->kernel[1] = a->ints[1]; } }
  if (auto a = find_attr(n, "pads")) { if (a->ints.size() >= 4) for (int i = 0; i < 4; ++i) p->pads[i] = a->ints[i]; }
  if (auto a = find_attr(n, "strides")) { if (a->ints.size() >= 2) { p->strides[0] = a->ints[0]; p->strides[1] = a->ints[1]; } }
  return 0;
}

int Planner::do_conv(size_t i) {
  const WireNode& n = m->wm.nodes[i];
  if (n.input.size() < 2 || n.output.empty()) B200_FAIL(B200_EINVAL, "Conv node %s: needs 2 inputs", n.name.c_str());
  Val* x = get(n.input[0]);
  Val* w = get(n.input[1]);
  if (!x || !w) B200_FAIL(B200_EINVAL, "Conv %s: input %s not available", n.name.c_str(), (!x ? n.input[0] : n.input[1]).c_str());
  if (x->is_init) B200_FAIL(B200_EUNSUPPORTED, "Conv %s: constant data input is not supported by the graph path", n.name.c_str());
  if (!w->is_init) B200_FAIL(B200_EUNSUPPORTED, "Conv %s: weights must be an initializer", n.name.c_str());
  if (x->rank != 4 || w->rank != 4) B200_FAIL(B200_EINVAL, "Conv %s: operands must be rank 4 (convolution_op.rs:101,111)", n.name.c_str());
  b200_conv_params p;
  B200_TRY(parse_conv_attrs(n, &p));
  int64_t yd[4];
  B200_TRY(b200_conv2d_out_dims(x->dims, w->dims, &p, yd));
  Geo g;
  {
    int ap = p.auto_pad;
    if (p.pads[0] > 0 || p.pads[1] > 0 || p.pads[2] > 0 || p.pads[3] > 0) ap = B200_PAD_NOTSET;
    B200_TRY(ref_geometry(ap, (int)x->dims[2], (int)x->dims[3], (int)w->dims[2], (int)w->dims[3], (int)p.strides[0], (int)p.strides[1], p.pads, &g));
  }
  const int M = (int)w->dims[0], C = (int)w->dims[1], KH = (int)w->dims[2], KW = (int)w->dims[3];
  const WireTensor* bias = nullptr;
  if (n.input.size() > 2) {
    bias = m->wm.find_initializer(n.input[2]);
    if (!bias) B200_FAIL(B200_EINVAL, "Conv %s: bias %s must be an initializer (convolution_op.rs:124)", n.name.c_str(), n.input[2].c_str());
  }
  // ---- epilogue fusion: Conv -> [Add(initializer [M,1,1])] -> [Relu], each link single-consumer
  std::string out_name = n.output[0];
  const WireTensor* chan_add = nullptr;
  int relu = 0;
  std::string label = n.name.empty() ? n.output[0] : n.name;
  size_t j;
  if (const WireNode* c = sole_consumer(out_name, &j)) {
    if (c->op_type == "Add" && c->input.size() == 2 && c->input[0] == out_name) {
      const WireTensor* t = m->wm.find_initializer(c->input[1]);
      if (t) {
        auto d = init_dims(m->wm, *t);
        if (d.size() == 3 && d[0] == M && d[1] == 1 && d[2] == 1) {
          chan_add = t; consumed.insert(j); out_name = c->output[0]; label += "+Add";
        }
      }
    }
  }
  if (const WireNode* c = sole_consumer(out_name, &j)) {
    if (c->op_type == "Relu" && !consumed.count(j)) { relu = 1; consumed.insert(j); out_name = c->output[0]; label += "+Relu"; }
  }
  int Ceff = C;
  if (C % 4 != 0 && C >= 3 && x->pad_zeroed && x->v.ld == round_up4(C)) Ceff = round_up4(C);
  // ---- Fire-module fusion: this 1x1 expand and its 3x3 (pad 1) sibling read the same squeeze output and write
  // adjacent channel slices of one Concat result.  When both fit one channel tile (M1 + M3 <= 128) they run as ONE
  // 3x3 convolution whose first M1 filters are the 1x1 weights at the centre tap and exact zeros elsewhere: the
  // tcgen05 kernel pays per MMA instruction, not per channel, below 128 channels, so the 1x1 branch rides along for
  // ~8 % of the 3x3 branch's time instead of a launch of its own.  Adding 0 * x terms is exact for finite x.
  if (m->opt_fire_fusion && conv_path() != 1 && !x->pooled && !chan_add && KH == 1 && KW == 1 && p.strides[0] == 1 && p.strides[1] == 1 &&
      g.pt == 0 && g.pl == 0 && g.Ho == (int)x->dims[2] && g.Wo == (int)x->dims[3] && Ceff % 4 == 0) {
    auto r1 = redirect.find(out_name);
    if (r1 != redirect.end() && r1->second.second == 0) {
      for (size_t j3 = i + 1; j3 < m->wm.nodes.size(); ++j3) {
        const WireNode& n3 = m->wm.nodes[j3];
        if (consumed.count(j3) || n3.op_type != "Conv" || n3.input.size() < 2 || n3.output.empty() || n3.input[0] != n.input[0]) continue;
        const WireTensor* w3 = m->wm.find_initializer(n3.input[1]);
        if (!w3) break;
        auto wd3 = init_dims(m->wm, *w3);
        if (wd3.size() != 4 || wd3[1] != C || wd3[2] != 3 || wd3[3] != 3 || M + wd3[0] > 128) break;
        const int M3 = (int)wd3[0];
        b200_conv_params p3;
        if (parse_conv_attrs(n3, &p3) || p3.strides[0] != 1 || p3.strides[1] != 1) break;
        Geo g3;
        {
          int ap = p3.auto_pad;
          if (p3.pads[0] > 0 || p3.pads[1] > 0 || p3.pads[2] > 0 || p3.pads[3] > 0) ap = B200_PAD_NOTSET;
          if (ref_geometry(ap, (int)x->dims[2], (int)x->dims[3], 3, 3, 1, 1, p3.pads, &g3)) break;
        }
        if (g3.pt != 1 || g3.pl != 1 || g3.Ho != g.Ho || g3.Wo != g.Wo) break;
        const WireTensor* bias3 = nullptr;
        if (n3.input.size() > 2 && !(bias3 = m->wm.find_initializer(n3.input[2]))) break;
        // the sibling's tail must match: [Relu] with the same flag, no folded Add, then the Concat's second slot
        std::string out3 = n3.output[0];
        size_t jr = 0;
        bool relu3 = false;
        if (const WireNode* c = sole_consumer(out3, &jr)) {
          if (c->op_type == "Add") break;
          if (c->op_type == "Relu" && !consumed.count(jr)) { relu3 = true; out3 = c->output[0]; }
        }
        if ((relu3 ? 1 : 0) != relu) break;
        auto r3 = redirect.find(out3);
        if (r3 == redirect.end() || r3->second.first != r1->second.first || r3->second.second != M) break;
        if ((int64_t)w->init->f32.size() != (int64_t)M * C || (int64_t)w3->f32.size() != (int64_t)M3 * C * 9) break;
        // ---- committed: one 3x3 convolution with M + M3 filters
        const int Mt = M + M3;
        // K order of the fused filter bank: the centre tap first (tap_perm), so that the 1x1 branch's only non-zero
        // weights sit in the first k-block(s) and the tcgen05 kernel can leave its columns out of all later k-blocks
        static const int kPos2Tap[9] = {4, 0, 1, 2, 3, 5, 6, 7, 8};
        unsigned long long tap_perm = 0;
        for (int i = 0; i < 9; ++i) tap_perm |= (unsigned long long)kPos2Tap[i] << (4 * i);
        std::string key = "convw:fire:ctr1st:" + w->init->name + "+" + w3->name + ":" + std::to_string(Ceff);
        float *dwf = nullptr, *dbf = nullptr;
        if (m->consts.count(key)) dwf = m->consts[key]->p;
        else {
          std::vector<float> h((size_t)Mt * 9 * Ceff, 0.f);
          for (int mm = 0; mm < M; ++mm)
            for (int c = 0; c < C; ++c) h[((size_t)mm * 9 + 0) * Ceff + c] = w->init->f32[(size_t)mm * C + c];   // K position 0 = centre tap
          for (int mm = 0; mm < M3; ++mm)
            for (int c = 0; c < C; ++c)
              for (int pos = 0; pos < 9; ++pos) {
                const int r = kPos2Tap[pos] / 3, sx = kPos2Tap[pos] % 3;
                h[((size_t)(M + mm) * 9 + pos) * Ceff + c] = w3->f32[(((size_t)mm * C + c) * 3 + r) * 3 + sx];
              }
          B200_TRY(upload_const(m, key, h, &dwf));
        }
        if (bias || bias3) {
          if ((bias && bias->f32.size() != (size_t)M) || (bias3 && bias3->f32.size() != (size_t)M3)) B200_FAIL(B200_EINVAL, "Conv %s: bias length", n.name.c_str());
          std::vector<float> hb((size_t)Mt, 0.f);
          if (bias) std::copy(bias->f32.begin(), bias->f32.end(), hb.begin());
          if (bias3) std::copy(bias3->f32.begin(), bias3->f32.end(), hb.begin() + M);
          B200_TRY(upload_const(m, "vec:fire:" + (bias ? bias->name : std::string("-")) + "+" + (bias3 ? bias3->name : std::string("-")), hb, &dbf));
        }
        ConvArgs a{};
        a.x = x->v.p; a.N = x->v.N; a.C = Ceff; a.H = x->v.H; a.W = x->v.W; a.ldx = x->v.ld;
        a.w = dwf; a.M = Mt; a.KH = 3; a.KW = 3; a.K = 9 * Ceff; a.wc = Ceff; a.ldw = a.K;
        a.bias = dbf; a.chan_add = nullptr;
        a.Ho = g.Ho; a.Wo = g.Wo; a.sh = 1; a.sw = 1; a.pt = 1; a.pl = 1; a.relu = relu;
        a.reverse = next_reverse();
        a.pdl = use_pdl();
        a.tap_perm = tap_perm;
        a.skip_m = M;                        // the 1x1 filters ...
        a.skip_kb = (Ceff + 31) / 32;        // ... are zero past the k-blocks that hold K position 0 (the centre tap)
        if (tc_supported(a) != 0) break;
        Val y1, y3;
        y1.rank = 4; memcpy(y1.dims, yd, sizeof(yd));
        y3 = y1; y3.dims[1] = M3;
        B200_TRY(place(out_name, &y1));
        B200_TRY(place(out3, &y3));
        a.y = y1.v.p; a.ldy = y1.v.ld;
        consumed.insert(j3);
        if (relu3) consumed.insert(jr);
        const double P = (double)y1.v.pixels();
        const double flops = 2.0 * P * C * ((double)M + 9.0 * M3);
        const double bytes = 4.0 * ((double)x->v.pixels() * C + P * Mt + (double)C * (M + 9.0 * M3));
        std::shared_ptr<TcWeights> tcw;
        if (!dry) {
          auto it = m->tc_weights.find("tc:" + key);
          if (it == m->tc_weights.end()) {
            B200_TRY(tc_prepare_weights(dwf, Mt, a.K, m->ctx->stream, &tcw));
            m->tc_weights["tc:" + key] = tcw;
          } else tcw = it->second;
        }
        std::string l3 = n3.name.empty() ? n3.output[0] : n3.name;
        add_step(label + " | " + l3 + (relu3 ? "+Relu" : ""), "conv_tc", flops, bytes, [a, tcw](cudaStream_t st) { return launch_conv_tc(a, *tcw, st); });
        env[out_name] = y1;
        env[out3] = y3;
        return 0;
      }
    }
  }
  Val y;
  y.rank = 4; memcpy(y.dims, yd, sizeof(yd));
  B200_TRY(place(out_name, &y));
  // ---- operand preparation
  float *dw = nullptr, *db = nullptr, *dadd = nullptr;
  if (bias) B200_TRY(vec_const(*bias, (size_t)M, &db));
  if (chan_add) B200_TRY(vec_const(*chan_add, (size_t)M, &dadd));
  ConvArgs a{};
  if (x->s2d) {
    // the stride-2 convolution over the space-to-depth input (Planner::run): kernel ceil(k/2)^2, 4*C channels, stride 1;
    // W'[m][r][s][(dy*2+dx)*C + c] = W[m][c][2r+dy][2s+dx], zero where 2r+dy >= KH or 2s+dx >= KW
    const int KH2 = (KH + 1) / 2, KW2 = (KW + 1) / 2, Cs = 4 * C;
    std::string key = "convw:s2d:" + w->init->name;
    if (m->consts.count(key)) dw = m->consts[key]->p;
    else {
      std::vector<float> h((size_t)M * KH2 * KW2 * Cs, 0.f);
      for (int mm = 0; mm < M; ++mm)
        for (int c = 0; c < C; ++c)
          for (int r = 0; r < KH; ++r)
            for (int sx = 0; sx < KW; ++sx)
              h[(((size_t)mm * KH2 + r / 2) * KW2 + sx / 2) * Cs + ((r & 1) * 2 + (sx & 1)) * C + c] =
                  w->init->f32[(((size_t)mm * C + c) * KH + r) * KW + sx];
      B200_TRY(upload_const(m, key, h, &dw));
    }
    a.x = x->v.p; a.N = x->v.N; a.C = Cs; a.H = x->v.H; a.W = x->v.W; a.ldx = x->v.ld;
    a.w = dw; a.M = M; a.KH = KH2; a.KW = KW2; a.K = KH2 * KW2 * Cs; a.wc = Cs; a.ldw = a.K;
    a.sh = 1; a.sw = 1; a.pt = 0; a.pl = 0;
    Ceff = Cs;   // keys the tcgen05 weight cache below
  } else {
    B200_TRY(conv_weights(*w->init, {M, C, KH, KW}, Ceff, &dw));
    a.x = x->v.p; a.N = x->v.N; a.C = Ceff; a.H = x->v.H; a.W = x->v.W; a.ldx = x->v.ld;
    a.w = dw; a.M = M; a.KH = KH; a.KW = KW; a.K = KH * KW * Ceff; a.wc = Ceff; a.ldw = a.K;
    a.sh = (int)p.strides[0]; a.sw = (int)p.strides[1]; a.pt = g.pt; a.pl = g.pl;
    if (x->pooled) {   // do_maxpool: the MaxPool in front of this pointwise convolution runs inside its launch
      a.x = x->pool_src.p; a.H = x->pool_src.H; a.W = x->pool_src.W; a.ldx = x->pool_src.ld;
      a.pool = 1; a.pool_pt = x->pool_pt; a.pool_pl = x->pool_pl;
      label = fused_pool_label[n.input[0]] + " | " + label;
    }
  }
  a.bias = db; a.chan_add = dadd;
  a.y = y.v.p; a.Ho = g.Ho; a.Wo = g.Wo; a.ldy = y.v.ld;
  a.relu = relu;
  a.reverse = next_reverse();
  a.pdl = use_pdl();
  const double P = (double)y.v.pixels();
  const double flops = 2.0 * P * M * C * KH * KW;
  const double bytes = 4.0 * ((double)(x->pooled ? x->pool_src.pixels() : x->v.pixels()) * C + P * M + (double)M * C * KH * KW);
  bool use_tc = false;
  std::shared_ptr<TcWeights> tcw;
  if (conv_path() != 1 && tc_supported(a) == 0) {
    use_tc = true;
    if (!dry) {
      std::string key = "tc:" + w->init->name + ":" + std::to_string(Ceff) + (x->s2d ? ":s2d" : "");
      auto it = m->tc_weights.find(key);
      if (it == m->tc_weights.end()) {
        B200_TRY(tc_prepare_weights(dw, M, a.K, m->ctx->stream, &tcw));
        m->tc_weights[key] = tcw;
      } else tcw = it->second;
    }
  } else if (conv_path() == 2) {
    B200_FAIL(B200_EUNSUPPORTED, "Conv %s: conv_path=2 (tcgen05) requested but the shape is not eligible", label.c_str());
  }
  if (a.pool && !use_tc) B200_FAIL(B200_EUNSUPPORTED, "Conv %s: planned with a fused MaxPool but not eligible for the tcgen05 path", label.c_str());
  // a pool-fused launch is bound by the pool's input traffic: it is listed with the bandwidth kernels, not with the conv family
  if (use_tc) add_step(label, a.pool ? "maxpool+conv_tc" : "conv_tc", flops, bytes, [a, tcw](cudaStream_t st) { return launch_conv_tc(a, *tcw, st); });
  else add_step(label, "conv_simt", flops, bytes, [a](cudaStream_t st) { return launch_conv_simt(a, st); });
  env[out_name] = y;
  return 0;
}

int Planner::do_maxpool(size_t i) {
  const WireNode& n = m->wm.nodes[i];
  Val* x = get(n.input[0]);
  if (!x || x->rank != 4 || x->is_init) B200_FAIL(B200_EINVAL, "MaxPool %s: input must be a rank-4 activation", n.name.c_str());
  b200_pool_params p;
  B200_TRY(parse_pool_attrs(n, &p));
  int64_t yd[4];
  B200_TRY(b200_maxpool2d_out_dims(x->dims, &p, yd));
  Geo g;
  B200_TRY(ref_geometry(p.auto_pad, (int)x->dims[2], (int)x->dims[3], (int)p.kernel[0], (int)p.kernel[1], (int)p.strides[0], (int)p.strides[1], p.pads, &g));
  Val y; y.rank = 4; memcpy(y.dims, yd, sizeof(yd));
  // ---- MaxPool 3x3 / 2 whose only consumer is a pointwise stride-1 convolution with <= 64 filters (SqueezeNet: pool1 ->
  // fire2 squeeze, pool after fire4 -> fire5 squeeze): the convolution's launch pools on the fly -- the window rows arrive
  // by TMA (out-of-bounds = the reference's zero padding) and the converter warps take the maximum on their way to tensor
  // memory -- so the pooled tensor is neither written nor read back.  A tile is as many pooled rows of one image as fit
  // the 128 MMA rows (54-pixel rows: 2, 27: 4, 13: 7 + 6).
  size_t jc = 0;
  if (m->opt_pool_fusion && conv_path() != 1 && p.kernel[0] == 3 && p.kernel[1] == 3 && p.strides[0] == 2 && p.strides[1] == 2 &&
      g.Wo >= 4 && x->v.ld == x->v.C && !x->s2d && !x->pooled) {   // (dense rows: arena allocations are 256-byte aligned)
    const WireNode* c = sole_consumer(n.output[0], &jc);
    const WireTensor* cw = (c && c->op_type == "Conv" && c->input.size() >= 2 && c->input[0] == n.output[0]) ? m->wm.find_initializer(c->input[1]) : nullptr;
    if (cw) {
      auto wd = init_dims(m->wm, *cw);
      b200_conv_params cp;
      bool ok = wd.size() == 4 && wd[1] == x->v.C && wd[2] == 1 && wd[3] == 1 && wd[0] <= 64 && parse_conv_attrs(*c, &cp) == 0 &&
                cp.strides[0] == 1 && cp.strides[1] == 1 && cp.pads[0] == 0 && cp.pads[1] == 0 && cp.pads[2] == 0 && cp.pads[3] == 0 &&
                (cp.auto_pad == B200_PAD_NOTSET || cp.auto_pad == B200_PAD_VALID);
      if (ok) {   // the same test launch_conv_tc applies (both planning passes must decide alike: no pointers involved)
        ConvArgs t{};
        t.N = x->v.N; t.C = x->v.C; t.H = x->v.H; t.W = x->v.W; t.ldx = x->v.ld; t.wc = t.C;
        t.M = (int)wd[0]; t.KH = t.KW = 1; t.K = t.C; t.sh = t.sw = 1; t.Ho = g.Ho; t.Wo = g.Wo;
        t.pool = 1; t.pool_pt = g.pt; t.pool_pl = g.pl;
        ok = tc_supported(t) == 0;
      }
      if (ok) {
        y.v.N = (int)yd[0]; y.v.C = (int)yd[1]; y.v.H = (int)yd[2]; y.v.W = (int)yd[3]; y.v.ld = y.v.C; y.v.p = nullptr;
        y.pooled = true; y.pool_src = x->v; y.pool_pt = g.pt; y.pool_pl = g.pl;
        y.alloc = x->alloc;   // reading the pooled value reads the pool's input: keeps it alive up to the convolution
        y.planned = true;
        env[n.output[0]] = y;
        fused_pool_label[n.output[0]] = n.name.empty() ? n.output[0] : n.name;
        return 0;
      }
    }
  }
  B200_TRY(place(n.output[0], &y));
  PoolArgs a{};
  a.x = x->v.p; a.N = x->v.N; a.C = x->v.C; a.H = x->v.H; a.W = x->v.W; a.ldx = x->v.ld;
  a.y = y.v.p; a.Ho = g.Ho; a.Wo = g.Wo; a.ldy = y.v.ld;
  a.kh = (int)p.kernel[0]; a.kw = (int)p.kernel[1]; a.sh = (int)p.strides[0]; a.sw = (int)p.strides[1]; a.pt = g.pt; a.pl = g.pl;
  a.reverse = next_reverse();
  const double bytes = 4.0 * ((double)x->v.numel() + (double)y.v.numel());
  add_step(n.name.empty() ? n.output[0] : n.name, "maxpool", 0, bytes, [a](cudaStream_t st) { return launch_maxpool(a, st); });
  env[n.output[0]] = y;
  return 0;
}

int Planner::do_relu(size_t i) {
  const WireNode& n = m->wm.nodes[i];
  Val* x = get(n.input[0]);
  if (!x || x->is_init || x->rank != 4) B200_FAIL(B200_EINVAL, "Relu %s: input must be a rank-4 activation (relu_op.rs:16)", n.name.c_str());
  Val y; y.rank = x->rank; memcpy(y.dims, x->dims, sizeof(y.dims));
  B200_TRY(place(n.output[0], &y));
  TView xv = x->v, yv = y.v;
  add_step(n.name.empty() ? n.output[0] : n.name, "relu", 0, 8.0 * (double)xv.numel(), [xv, yv](cudaStream_t st) { return launch_relu(xv, yv, st); });
  env[n.output[0]] = y;
  return 0;
}

int Planner::do_add(size_t i) {
  // add(), add_op.rs:16-107: input 2 must be an initializer (:54-68)
  const WireNode& n = m->wm.nodes[i];
  if (n.input.size() != 2) B200_FAIL(B200_EINVAL, "Add %s: needs 2 inputs", n.name.c_str());
  Val* x = get(n.input[0]);
  const WireTensor* bt = m->wm.find_initializer(n.input[1]);
  if (!bt) B200_FAIL(B200_EUNSUPPORTED, "Cannot retrieve input 2 for Add operation (add_op.rs:66): %s is not an initializer", n.input[1].c_str());
  if (!x || x->is_init) B200_FAIL(B200_EUNSUPPORTED, "Add %s: constant first operand is not supported by the graph path", n.name.c_str());
  auto bd = init_dims(m->wm, *bt);
  Val y; y.rank = x->rank; memcpy(y.dims, x->dims, sizeof(y.dims));
  y.perm_C = 0;
  if (x->rank == 4) {
    if (bd.size() != 3 || bd[0] != x->dims[1] || bd[1] != 1 || bd[2] != 1)
      B200_FAIL(B200_EUNSUPPORTED, "Add %s: rank-4 input needs a [C,1,1] initializer (add_op.rs:55-58,75)", n.name.c_str());
    float* db = nullptr;
    B200_TRY(vec_const(*bt, (size_t)bd[0], &db));
    B200_TRY(place(n.output[0], &y));
    TView xv = x->v, yv = y.v;
    add_step(n.name.empty() ? n.output[0] : n.name, "add_channel", (double)xv.numel(), 8.0 * (double)xv.numel(),
             [xv, db, yv](cudaStream_t st) { return launch_add_channel(xv, db, yv, st); });
  } else if (x->rank == 2) {
    // reference: same-shape [1,K] + [1,K]; batch-N extension broadcasts the initializer over rows
    if (bd.size() != 2 || bd[1] != x->dims[1] || (bd[0] != 1 && bd[0] != x->dims[0]))
      B200_FAIL(B200_EINVAL, "Add %s: rank-2 operands must have the same shape (add_op.rs:84)", n.name.c_str());
    float* db = nullptr;
    B200_TRY(vec_const(*bt, (size_t)(bd[0] * bd[1]), &db));
    B200_TRY(place(n.output[0], &y));
    TView xv = x->v, yv = y.v, bv;
    bv.p = db; bv.N = (int)bd[0]; bv.C = (int)bd[1]; bv.H = bv.W = 1; bv.ld = bv.C;
    add_step(n.name.empty() ? n.output[0] : n.name, "add_rows", (double)xv.numel(), 8.0 * (double)xv.numel(),
             [xv, bv, yv](cudaStream_t st) { return launch_add_same(xv, bv, yv, st); });
  } else {
    B200_FAIL(B200_EUNSUPPORTED, "Add %s: input rank %d", n.name.c_str(), x->rank);
  }
  env[n.output[0]] = y;
  return 0;
}

int Planner::do_reshape(size_t i) {
  // reshape(), reshape_op.rs:16-92
  const WireNode& n = m->wm.nodes[i];
  if (n.input.size() != 2) B200_FAIL(B200_EINVAL, "Reshape %s: needs 2 inputs", n.name.c_str());
  const WireTensor* st = m->wm.find_initializer(n.input[1]);
  if (!st) B200_FAIL(B200_EUNSUPPORTED, "Unable to retrieve Shape for Reshape operation (reshape_op.rs:42)");
  if (st->i64.size() < 2) B200_FAIL(B200_EINVAL, "Reshape %s: shape must hold >= 2 int64 values (reshape_op.rs:87)", n.name.c_str());
  Val* x = get(n.input[0]);
  if (!x || x->rank != 4) B200_FAIL(B200_EINVAL, "Reshape %s: data must be rank 4 (reshape_op.rs:27,30)", n.name.c_str());
  int64_t s0 = st->i64[0], s1 = st->i64[1];
  Val y; y.rank = 2;
  if (x->is_init) {
    // constant folding: Reshape(initializer) happens once, on the host, in the reference's memory order
    if (s0 == 0) s0 = x->dims[0];
    if (s1 == 0) s1 = x->dims[1];
    const int64_t total = x->dims[0] * x->dims[1] * x->dims[2] * x->dims[3];
    if (s0 < 0 || s1 < 0 || s0 * s1 != total || (int64_t)x->init->f32.size() != total)
      B200_FAIL(B200_EINVAL, "Reshape %s: %lldx%lld does not match %lld elements (reshape_op.rs:90)", n.name.c_str(), (long long)s0, (long long)s1, (long long)total);
    y.dims[0] = s0; y.dims[1] = s1;
    y.host2d = std::make_shared<std::vector<float>>(x->init->f32);
    y.is_init = false; y.planned = false;
    env[n.output[0]] = y;
    return 0;
  }
  // activation: per-image element count is fixed; the leading dim scales with the batch
  const int64_t static_n = m->in_dims[0] > 0 ? m->in_dims[0] : 1;
  const int64_t per_image = x->dims[1] * x->dims[2] * x->dims[3];
  if (s0 == 0) s0 = static_n;
  if (s1 == 0) s1 = x->dims[1];
  if (s0 < 0 || s1 <= 0) B200_FAIL(B200_EUNSUPPORTED, "Reshape %s: -1 is not supported (reshape_op.rs:87)", n.name.c_str());
  if (s0 * s1 != static_n * per_image)
    B200_FAIL(B200_EINVAL, "Reshape %s: %lldx%lld does not match %lld elements (reshape_op.rs:90)", n.name.c_str(), (long long)s0, (long long)s1, (long long)(static_n * per_image));
  const int64_t total = x->dims[0] * per_image;
  y.dims[1] = s1; y.dims[0] = total / s1;
  const bool order_free = x->v.dense() && (x->v.H * x->v.W == 1 || x->v.C == 1);
  // Fold the NCHW flatten order into the consumer MatMul's constant weight when possible.
  bool fold = false;
  size_t j;
  if (!order_free && x->v.dense() && s1 == per_image) {
    if (const WireNode* c = sole_consumer(n.output[0], &j)) {
      if (c->op_type == "MatMul" && c->input.size() == 2 && c->input[0] == n.output[0]) {
        Val* b = get(c->input[1]);
        if (b && b->host2d) fold = true;
      }
    }
  }
  if (order_free || fold) {
    y.v.p = x->v.p; y.v.N = (int)y.dims[0]; y.v.C = (int)y.dims[1]; y.v.H = y.v.W = 1; y.v.ld = y.v.C;
    y.alloc = x->alloc;
    y.planned = true;
    if (fold) { y.perm_C = x->v.C; y.perm_HW = x->v.H * x->v.W; }
  } else {
    B200_TRY(place(n.output[0], &y));
    TView xv = x->v; float* dst = y.v.p;
    add_step(n.name.empty() ? n.output[0] : n.name, "reshape_nchw", 0, 8.0 * (double)xv.numel(),
             [xv, dst](cudaStream_t st) { return launch_rows_to_nchw(xv, dst, st); });
  }
  env[n.output[0]] = y;
  return 0;
}

int Planner::do_matmul(size_t i) {
  // mul(), mul_op.rs:11-32: both operands come from the store's 2-D slot, i.e. from Reshape outputs
  const WireNode& n = m->wm.nodes[i];
  if (n.input.size() != 2) B200_FAIL(B200_EINVAL, "MatMul %s: needs 2 inputs", n.name.c_str());
  Val* a = get(n.input[0]);
  Val* b = get(n.input[1]);
  if (!a || !b || a->rank != 2 || b->rank != 2 || a->is_init || b->is_init)
    B200_FAIL(B200_EINVAL, "MatMul %s: operands must be rank-2 store values (mul_op.rs:16-19 unwrap the 2-D slot)", n.name.c_str());
  if (!b->host2d) B200_FAIL(B200_EUNSUPPORTED, "MatMul %s: the right operand must be a constant (Reshape of an initializer)", n.name.c_str());
  if (a->host2d) B200_FAIL(B200_EUNSUPPORTED, "MatMul %s: constant left operand is not supported", n.name.c_str());
  const int K = (int)a->dims[1], N = (int)b->dims[1];
  if (b->dims[0] != K) B200_FAIL(B200_EINVAL, "MatMul %s: inner dims %d vs %lld", n.name.c_str(), K, (long long)b->dims[0]);
  // weight as [N][K] rows (K-contiguous), with the activation's (h,w,c) flatten order folded in if needed
  std::string key = "matw:" + n.input[1] + ":" + std::to_string(a->perm_C) + "x" + std::to_string(a->perm_HW);
  float* dw = nullptr;
  B200_TRY(matmul_weights(*b, a->perm_C, a->perm_HW, K, N, key, &dw));
  // fuse the following Add of a [1,N] initializer (MNIST Plus214, add_op.rs:84)
  std::string out_name = n.output[0];
  std::string label = n.name.empty() ? n.output[0] : n.name;
  float* dbias = nullptr;
  size_t j;
  if (const WireNode* c = sole_consumer(out_name, &j)) {
    if (c->op_type == "Add" && c->input.size() == 2 && c->input[0] == out_name) {
      if (const WireTensor* t = m->wm.find_initializer(c->input[1])) {
        auto d = init_dims(m->wm, *t);
        if (d.size() == 2 && d[0] == 1 && d[1] == N) {
          B200_TRY(vec_const(*t, (size_t)N, &dbias));
          consumed.insert(j); out_name = c->output[0]; label += "+Add";
        }
      }
    }
  }
  Val y; y.rank = 2; y.dims[0] = a->dims[0]; y.dims[1] = N;
  B200_TRY(place(out_name, &y));
  ConvArgs c{};
  c.x = a->v.p; c.N = a->v.N; c.C = K; c.H = 1; c.W = 1; c.ldx = a->v.ld;
  c.w = dw; c.M = N; c.KH = 1; c.KW = 1; c.K = K; c.ldw = K; c.wc = K;
  c.bias = dbias; c.chan_add = nullptr;
  c.y = y.v.p; c.Ho = 1; c.Wo = 1; c.ldy = y.v.ld; c.sh = c.sw = 1; c.pt = c.pl = 0; c.relu = 0;
  const double R = (double)a->dims[0];
  c.reverse = next_reverse();
  c.pdl = use_pdl();
  // mul_op.rs:23 on the convolution's tcgen05 path: rows are the "pixels" of a pointwise layer (TMA-fed A tiles), the
  // N columns one channel tile (MNIST: 10 -> BN = 16)
  if (conv_path() != 1 && tc_supported(c) == 0) {
    std::shared_ptr<TcWeights> tcw;
    if (!dry) {
      auto it = m->tc_weights.find("tc:" + key);
      if (it == m->tc_weights.end()) {
        B200_TRY(tc_prepare_weights(dw, N, K, m->ctx->stream, &tcw));
        m->tc_weights["tc:" + key] = tcw;
      } else tcw = it->second;
    }
    add_step(label, "matmul_tc", 2.0 * R * K * N, 4.0 * (R * K + R * N + (double)K * N),
             [c, tcw](cudaStream_t st) { return launch_conv_tc(c, *tcw, st); });
  } else {
    if (conv_path() == 2) B200_FAIL(B200_EUNSUPPORTED, "MatMul %s: conv_path=2 (tcgen05) requested but the shape is not eligible", label.c_str());
    add_step(label, "matmul_simt", 2.0 * R * K * N, 4.0 * (R * K + R * N + (double)K * N),
             [c](cudaStream_t st) { return launch_conv_simt(c, st); });
  }
  env[out_name] = y;
  return 0;
}

int Planner::matmul_weights(const Val& b, int perm_C, int perm_HW, int K, int N, const std::string& key, float** out) {
  if (m->consts.count(key)) { *out = m->consts[key]->p; return 0; }
  std::vector<float> h((size_t)N * K);
  const std::vector<float>& src = *b.host2d;  // [K][N], k in the reference's (c,h,w) order
  if ((int64_t)src.size() != (int64_t)K * N) B200_FAIL(B200_EINVAL, "MatMul weight %s: %zu floats, expected %lld", key.c_str(), src.size(), (long long)K * N);
  for (int k = 0; k < K; ++k) {
    int kp = k;  // position of reference row k in the activation's physical row
    if (perm_C > 0) { const int c = k / perm_HW, hw = k % perm_HW; kp = hw * perm_C + c; }
    for (int nn = 0; nn < N; ++nn) h[(size_t)nn * K + kp] = src[(size_t)k * N + nn];
  }
  return upload_const(m, key, h, out);
}

// ---------------------------------------------------------------- the MNIST-8 graph as two fused launches
// Matches exactly: input [*,1,28,28] -> Conv(8x1x5x5, stride 1, pads 2) [-> Add [8,1,1]] -> Relu -> MaxPool 2x2/2
//   -> Conv(16x8x5x5, stride 1, pads 2) [-> Add [16,1,1]] -> Relu -> MaxPool 3x3/3 -> Reshape [*,256]
//   -> MatMul(Reshape(initializer) [256,10]) [-> Add [1,10]] = graph output, every link single-consumer.
// Everything else keeps the node-by-node plan.  (BASELINE.json config 5; kernels and layout in mnist8_fused.cu.)
int Planner::try_plan_mnist8(bool* done) {
  *done = false;
  if (!m->opt_fused_cnn || conv_path() == 1) return 0;
  if (m->in_dims[1] != 1 || m->in_dims[2] != 28 || m->in_dims[3] != 28 || (m->in_dims[0] != 1 && m->in_dims[0] > 0)) return 0;
  auto next = [&](const std::string& v, const char* op, size_t* idx) -> const WireNode* {
    const WireNode* c = sole_consumer(v, idx);
    return (c && c->op_type == op && !c->input.empty() && c->input[0] == v && !c->output.empty()) ? c : nullptr;
  };
  struct ConvStage { const WireNode* conv = nullptr; const WireTensor *w = nullptr, *bias = nullptr, *add = nullptr; std::string out; };
  std::vector<size_t> used;
  auto conv_stage = [&](const std::string& in, int C, int M, int HW, int pool_k, int pool_out, ConvStage* cs) -> bool {
    size_t i;
    const WireNode* c = next(in, "Conv", &i);
    if (!c || c->input.size() < 2) return false;
    cs->conv = c; used.push_back(i);
    cs->w = m->wm.find_initializer(c->input[1]);
    if (!cs->w) return false;
    auto wd = init_dims(m->wm, *cs->w);
    if (wd.size() != 4 || wd[0] != M || wd[1] != C || wd[2] != 5 || wd[3] != 5 || (int64_t)cs->w->f32.size() != (int64_t)M * C * 25) return false;
    if (c->input.size() > 2) {
      cs->bias = m->wm.find_initializer(c->input[2]);
      if (!cs->bias || cs->bias->f32.size() != (size_t)M) return false;
    }
    b200_conv_params p;
    if (parse_conv_attrs(*c, &p) || p.strides[0] != 1 || p.strides[1] != 1 || p.group > 1 || p.dilations[0] > 1 || p.dilations[1] > 1) return false;
    Geo g;
    int ap = p.auto_pad;
    if (p.pads[0] > 0 || p.pads[1] > 0 || p.pads[2] > 0 || p.pads[3] > 0) ap = B200_PAD_NOTSET;
    if (ref_geometry(ap, HW, HW, 5, 5, 1, 1, p.pads, &g) || g.Ho != HW || g.Wo != HW || g.pt != 2 || g.pl != 2 || g.pb != 2 || g.pr != 2) return false;
    std::string v = c->output[0];
    if (const WireNode* a = next(v, "Add", &i)) {
      if (a->input.size() != 2) return false;
      cs->add = m->wm.find_initializer(a->input[1]);
      if (!cs->add) return false;
      auto ad = init_dims(m->wm, *cs->add);
      if (ad.size() != 3 || ad[0] != M || ad[1] != 1 || ad[2] != 1 || cs->add->f32.size() != (size_t)M) return false;
      used.push_back(i); v = a->output[0];
    }
    const WireNode* r = next(v, "Relu", &i);
    if (!r) return false;
    used.push_back(i); v = r->output[0];
    const WireNode* mp = next(v, "MaxPool", &i);
    if (!mp) return false;
    b200_pool_params pp;
    if (parse_pool_attrs(*mp, &pp) || pp.kernel[0] != pool_k || pp.kernel[1] != pool_k || pp.strides[0] != pool_k || pp.strides[1] != pool_k) return false;
    Geo pg;
    if (ref_geometry(pp.auto_pad, HW, HW, pool_k, pool_k, pool_k, pool_k, pp.pads, &pg) || pg.Ho != pool_out || pg.Wo != pool_out || pg.pt || pg.pl || pg.pb || pg.pr) return false;
    used.push_back(i);
    cs->out = mp->output[0];
    return true;
  };
  ConvStage s1, s2;
  if (!conv_stage(m->input_name, 1, 8, 28, 2, 14, &s1)) return 0;
  if (!conv_stage(s1.out, 8, 16, 14, 3, 4, &s2)) return 0;
  size_t i;
  const WireNode* rs = next(s2.out, "Reshape", &i);
  if (!rs || rs->input.size() != 2) return 0;
  const WireTensor* shp = m->wm.find_initializer(rs->input[1]);
  if (!shp || shp->i64.size() < 2 || shp->i64[1] != 256 || (shp->i64[0] != 1 && shp->i64[0] != 0)) return 0;
  used.push_back(i);
  const WireNode* mm = next(rs->output[0], "MatMul", &i);
  if (!mm || mm->input.size() != 2) return 0;
  used.push_back(i);
  // the right operand: Reshape(initializer), constant-folded by do_reshape (file order puts it first in mnist-8)
  size_t ib = (size_t)-1;
  for (size_t k = 0; k < m->wm.nodes.size(); ++k)
    if (!m->wm.nodes[k].output.empty() && m->wm.nodes[k].output[0] == mm->input[1]) ib = k;
  if (ib == (size_t)-1 || m->wm.nodes[ib].op_type != "Reshape" || m->wm.nodes[ib].input.size() != 2 || !m->wm.find_initializer(m->wm.nodes[ib].input[0])) return 0;
  if (n_consumers[mm->input[1]] != 1) return 0;
  B200_TRY(do_reshape(ib));
  used.push_back(ib);
  Val* bval = get(mm->input[1]);
  if (!bval || !bval->host2d || bval->dims[0] != 256 || bval->dims[1] != 10) return 0;
  std::string out_name = mm->output[0];
  const WireTensor* bm = nullptr;
  if (const WireNode* a = next(out_name, "Add", &i)) {
    if (a->input.size() != 2) return 0;
    bm = m->wm.find_initializer(a->input[1]);
    if (!bm) return 0;
    auto bd = init_dims(m->wm, *bm);
    if (bd.size() != 2 || bd[0] != 1 || bd[1] != 10 || bm->f32.size() != 10) return 0;
    used.push_back(i); out_name = a->output[0];
  }
  if (m->wm.outputs.empty() || m->wm.outputs[0].name != out_name) return 0;
  if (used.size() != m->wm.nodes.size()) return 0;   // nothing else in the graph
  // ---- committed: constants (shared with the node-by-node plan through the keyed caches)
  float *dw2 = nullptr, *db2 = nullptr, *da2 = nullptr, *dwm = nullptr, *dbm = nullptr;
  Mnist8StemConsts k1;   // the stem's 200 weights, bias and folded Add travel as kernel parameters (constant bank)
  for (int mm = 0; mm < 8; ++mm) {
    for (int t = 0; t < 25; ++t) k1.w[t][mm] = s1.w->f32[(size_t)mm * 25 + t];
    k1.bias[mm] = s1.bias ? s1.bias->f32[(size_t)mm] : 0.f;
    k1.add[mm] = s1.add ? s1.add->f32[(size_t)mm] : 0.f;
  }
  B200_TRY(conv_weights(*s2.w, {16, 8, 5, 5}, 8, &dw2));
  if (s2.bias) B200_TRY(vec_const(*s2.bias, 16, &db2));
  if (s2.add) B200_TRY(vec_const(*s2.add, 16, &da2));
  B200_TRY(matmul_weights(*bval, 16, 16, 256, 10, "matw:" + mm->input[1] + ":16x16", &dwm));
  if (bm) B200_TRY(vec_const(*bm, 10, &dbm));
  const int N = (int)B;
  const bool onepass = m->opt_fused_cnn >= 2;
  // two-launch form only: the pooled stem output lives in HBM; its zero halo must survive, so it is never shared
  float* p1 = onepass ? nullptr : arena_alloc(mnist8_p1_floats(N), nullptr, /*forever=*/true);
  Val y; y.rank = 2; y.dims[0] = B; y.dims[1] = 10;
  B200_TRY(place(out_name, &y));
  for (size_t k : used) consumed.insert(k);
  env[out_name] = y;
  if (dry) { ++step_counter; return *done = true, 0; }
  std::shared_ptr<TcWeights> tcw;
  {
    const std::string key = "tc:" + s2.w->name + ":8:natural_k";
    auto it = m->tc_weights.find(key);
    if (it == m->tc_weights.end()) {
      B200_TRY(tc_prepare_weights(dw2, 16, 200, m->ctx->stream, &tcw, /*natural_k=*/true));
      m->tc_weights[key] = tcw;
    } else tcw = it->second;
  }
  float* dout = y.v.p;
  plan->fused_name = onepass ? "mnist8_onepass(all 12 nodes: stem on CUDA cores one group ahead of the tcgen05 head)" : "mnist8_stem(Convolution28+Plus30+ReLU32+Pooling66)";
  if (onepass) {
    // ONE launch: it reads the caller's input and writes the logits (the launch list of the plan is empty)
    plan->fused_flops = 2.0 * N * (8.0 * 25 * 784 + 16.0 * 200 * 196 + 256.0 * 10);   // the reference's 1.573 MFLOP per image
    plan->fused_bytes = 4.0 * N * (784.0 + 10.0);
    plan->in_stem = [=](const float* d_in, cudaStream_t st, int* flag) {
      return launch_mnist8_onepass(d_in, k1, *tcw, db2, da2, dwm, dbm, dout, N, st, flag);
    };
    *done = true;
    return 0;
  }
  // the halo (and the pad bytes per image) of the stem output are zero for the plan's lifetime: nothing else writes them
  B200_CUDA(cudaMemsetAsync(p1, 0, mnist8_p1_floats(N) * sizeof(float), m->ctx->stream));
  plan->fused_flops = 2.0 * N * 8.0 * 25 * 784;
  plan->fused_bytes = 4.0 * N * (784.0 + 14.0 * 14 * 8);
  plan->in_stem = [=](const float* d_in, cudaStream_t st, int* flag) { return launch_mnist8_stem(d_in, k1, p1, N, st, flag); };
  // the reference's FLOPs for these nodes (1.573 MFLOP per image with the stem's 0.314); the 52 conv2 outputs per image
  // that MaxPool 3x3/3 floors away are not computed here
  const double flops = 2.0 * N * (16.0 * 200 * 196 + 256.0 * 10);
  const double bytes = 4.0 * N * (14.0 * 14 * 8 + 10.0) + 4.0 * (3200 + 2560);
  add_step("mnist8_head(Convolution110+Plus112+ReLU114+Pooling160+Times212+Plus214)", "mnist8_fused", flops, bytes,
           [=](cudaStream_t st) { return launch_mnist8_head(p1, *tcw, db2, da2, dwm, dbm, dout, N, st); });
  *done = true;
  return 0;
}

int Planner::do_concat(size_t i) {
  const WireNode& n = m->wm.nodes[i];
  if (n.input.size() != 2) B200_FAIL(B200_EUNSUPPORTED, "Concat %s: exactly two inputs (concatenate_op.rs:15-18)", n.name.c_str());
  B200_TRY(attr_check(n, {"axis"}, "CONCATENATE"));
  const int64_t axis = attr_i(n, "axis", 1);
  if (axis < 0 || axis > 3) B200_FAIL(B200_EINVAL, "Concat %s: axis %lld out of range (concatenate_op.rs:31)", n.name.c_str(), (long long)axis);
  Val* a = get(n.input[0]);
  Val* b = get(n.input[1]);
  if (!a || !b || a->rank != 4 || b->rank != 4 || a->is_init || b->is_init)
    B200_FAIL(B200_EINVAL, "Concat %s: inputs must be rank-4 activations (concatenate_op.rs:16,18)", n.name.c_str());
  for (int d = 0; d < 4; ++d)
    if (d != axis && a->dims[d] != b->dims[d]) B200_FAIL(B200_EINVAL, "Concat %s: non-axis dims differ", n.name.c_str());
  auto it = env.find(n.output[0]);
  if (axis == 1 && it != env.end() && it->second.planned && redirect.count(n.input[0])) {
    return 0;  // both producers already wrote into the result: zero-copy
  }
  Val y; y.rank = 4;
  for (int d = 0; d < 4; ++d) y.dims[d] = a->dims[d];
  y.dims[axis] += b->dims[axis];
  B200_TRY(place(n.output[0], &y));
  std::string nm = n.name.empty() ? n.output[0] : n.name;
  TView av = a->v, bv = b->v;
  if (axis == 1) {
    TView ya = y.v, yb = y.v;
    ya.C = av.C; yb.C = bv.C; if (!dry) yb.p += av.C;
    add_step(nm + ":0", "concat_copy", 0, 8.0 * (double)av.numel(), [av, ya](cudaStream_t st) { return launch_copy_rows(av, ya, st); });
    add_step(nm + ":1", "concat_copy", 0, 8.0 * (double)bv.numel(), [bv, yb](cudaStream_t st) { return launch_copy_rows(bv, yb, st); });
  } else {
    // batch / rows / columns (never used by the two bundled models): block copies in the channels-last layout.  Along the
    // batch axis the result no longer scales with the planned batch like the inputs do; it is still what
    // ndarray::concatenate(Axis(0)) returns for the planned tensors.
    TView yv = y.v;
    const int n0 = axis == 0 ? av.N : 0, h0 = axis == 2 ? av.H : 0, w0 = axis == 3 ? av.W : 0;
    add_step(nm + ":0", "concat_copy", 0, 8.0 * (double)av.numel(), [av, yv](cudaStream_t st) { return launch_copy_block(av, yv, 0, 0, 0, st); });
    add_step(nm + ":1", "concat_copy", 0, 8.0 * (double)bv.numel(), [bv, yv, n0, h0, w0](cudaStream_t st) { return launch_copy_block(bv, yv, n0, h0, w0, st); });
  }
  env[n.output[0]] = y;
  return 0;
}

int Planner::do_dropout(size_t i) {
  const WireNode& n = m->wm.nodes[i];
  B200_TRY(attr_check(n, {"ratio"}, "DROP OUT"));  // dropout_op.rs:22-28
  Val* x = get(n.input[0]);
  if (!x || x->rank != 4 || x->is_init) B200_FAIL(B200_EINVAL, "Dropout %s: input must be a rank-4 activation", n.name.c_str());
  env[n.output[0]] = *x;  // identity at inference (dropout_op.rs:66-71): alias, no kernel
  return 0;
}

int Planner::do_gap(size_t i) {
  const WireNode& n = m->wm.nodes[i];
  Val* x = get(n.input[0]);
  if (!x || x->rank != 4 || x->is_init) B200_FAIL(B200_EINVAL, "GlobalAveragePool %s: input must be a rank-4 activation", n.name.c_str());
  size_t j;
  const WireNode* c = sole_consumer(n.output[0], &j);
  TView xv = x->v;
  if (c && c->op_type == "Softmax" && (size_t)xv.C * 5 * sizeof(float) <= 48 * 1024) {
    // SqueezeNet tail: GAP + Softmax in one kernel (global_average_pool_op.rs:33-52 + softmax_op.rs:45-57)
    consumed.insert(j);
    Val y; y.rank = 2; y.dims[0] = x->dims[0]; y.dims[1] = x->dims[1];
    B200_TRY(place(c->output[0], &y));
    float* dst = y.v.p;
    add_step((n.name.empty() ? n.output[0] : n.name) + "+Softmax", "gap_softmax", (double)xv.numel(), 4.0 * ((double)xv.numel() + (double)y.v.numel()),
             [xv, dst](cudaStream_t st) { return launch_gap_softmax(xv, dst, st); });
    env[c->output[0]] = y;
    return 0;
  }
  Val y; y.rank = 4; y.dims[0] = x->dims[0]; y.dims[1] = x->dims[1]; y.dims[2] = 1; y.dims[3] = 1;
  B200_TRY(place(n.output[0], &y));
  TView yv = y.v;
  add_step(n.name.empty() ? n.output[0] : n.name, "global_avgpool", (double)xv.numel(), 4.0 * ((double)xv.numel() + (double)yv.numel()),
           [xv, yv](cudaStream_t st) { return launch_global_avgpool(xv, yv, st); });
  env[n.output[0]] = y;
  return 0;
}

int Planner::do_softmax(size_t i) {
  const WireNode& n = m->wm.nodes[i];
  Val* x = get(n.input[0]);
  if (!x || x->rank != 4 || x->is_init) B200_FAIL(B200_EINVAL, "Softmax %s: input must be a rank-4 activation (softmax_op.rs:18)", n.name.c_str());
  Val y; y.rank = 2; y.dims[0] = x->dims[0]; y.dims[1] = x->dims[1] * x->dims[2] * x->dims[3];
  B200_TRY(place(n.output[0], &y));
  TView xv = x->v; float* dst = y.v.p;
  add_step(n.name.empty() ? n.output[0] : n.name, "softmax", 4.0 * (double)xv.numel(), 8.0 * (double)xv.numel(),
           [xv, dst](cudaStream_t st) { return launch_softmax(xv, dst, st); });
  env[n.output[0]] = y;
  return 0;
}

int Planner::run() {
  env.clear(); consumed.clear(); alloc_seq = 0; step_counter = 0;
  if (dry) { recs.clear(); arena_cursor = 0; }
  n_consumers.clear();
  for (auto& n : m->wm.nodes) for (auto& in : n.input) n_consumers[in]++;
  redirect.clear();
  B200_TRY(plan_concat_redirects());
  // graph input
  Val in; in.rank = 4; in.dims[0] = B * (m->in_dims[0] > 0 ? m->in_dims[0] : 1);
  in.dims[1] = m->in_dims[1]; in.dims[2] = m->in_dims[2]; in.dims[3] = m->in_dims[3];
  in.v.N = (int)in.dims[0]; in.v.C = (int)in.dims[1]; in.v.H = (int)in.dims[2]; in.v.W = (int)in.dims[3];
  in.v.ld = in.v.C >= 3 ? round_up4(in.v.C) : in.v.C;
  // Stride-2 stem (SqueezeNet conv1: 7x7 / 2 on 3 channels): store the input 2x2 space-to-depth, [N, H/2, W/2, 4*C]
  // with channel (dy*2+dx)*C + c.  The convolution becomes a stride-1 one with a ceil(k/2)^2 kernel over 4*C channels
  // (do_conv builds the zero-padded weights): K = 4*4*12 = 192 = six full k-blocks instead of 7*7*4 = 196 (the
  // channel-padded layout, seven k-blocks), 12 channels are three 16-byte chunks per tap, and the transform writes
  // 25 % fewer bytes.  Exact: the extra taps have zero weights.
  plan->in_s2d = false;
  if (m->opt_s2d && conv_path() != 1 && in.v.C >= 1 && (4 * in.v.C) % 4 == 0 && in.v.H % 2 == 0 && in.v.W % 2 == 0) {
    size_t jc = 0;
    const WireNode* c = sole_consumer(m->input_name, &jc);
    const WireTensor* w = (c && c->op_type == "Conv" && c->input.size() >= 2 && c->input[0] == m->input_name) ? m->wm.find_initializer(c->input[1]) : nullptr;
    b200_conv_params cp;
    if (w && parse_conv_attrs(*c, &cp) == 0 && cp.strides[0] == 2 && cp.strides[1] == 2 && cp.pads[0] == 0 && cp.pads[1] == 0 &&
        cp.pads[2] == 0 && cp.pads[3] == 0 && (cp.auto_pad == B200_PAD_VALID || cp.auto_pad == B200_PAD_NOTSET)) {
      auto wd = init_dims(m->wm, *w);
      if (wd.size() == 4 && wd[1] == in.v.C && wd[2] >= 2 && wd[3] >= 2 && wd[2] <= 16 && wd[3] <= 16 && wd[2] <= in.v.H && wd[3] <= in.v.W &&
          (int64_t)w->f32.size() == wd[0] * wd[1] * wd[2] * wd[3]) {
        plan->in_s2d = true;
        plan->in_c0 = in.v.C; plan->in_h0 = in.v.H; plan->in_w0 = in.v.W;
        in.s2d = true;
        in.v.C = 4 * in.v.C; in.v.H /= 2; in.v.W /= 2; in.v.ld = in.v.C;
      }
    }
  }
  in.v.p = arena_alloc((size_t)in.v.pixels() * in.v.ld, &in.alloc);
  in.planned = true; in.pad_zeroed = true;
  env[m->input_name] = in;
  plan->in_view = in.v;
  plan->in_zero_pad = !plan->in_s2d && in.v.ld != in.v.C;
  plan->in_direct = !plan->in_s2d && in.v.dense() && (in.v.C == 1 || in.v.H * in.v.W == 1);

  // Concat results need their dims before the producers run: shape-only evaluation happens naturally in
  // file order because each producer's `place` looks the parent up in env; so register Concat outputs
  // lazily: when the first redirected producer is placed, the parent's dims must exist.  Compute them here
  // with a light shape walk (Conv / MaxPool / Relu / Dropout / Concat are the only rank-4 -> rank-4 ops).
  {
    std::map<std::string, std::vector<int64_t>> shp;
    shp[m->input_name] = {in.dims[0], in.dims[1], in.dims[2], in.dims[3]};
    for (auto& n : m->wm.nodes) {
      auto have = [&](const std::string& s) { return shp.count(s) > 0; };
      if (n.input.empty() || n.output.empty()) continue;
      if (n.op_type == "Conv" && have(n.input[0])) {
        const WireTensor* w = n.input.size() > 1 ? m->wm.find_initializer(n.input[1]) : nullptr;
        if (!w) continue;
        auto wd = init_dims(m->wm, *w);
        if (wd.size() != 4) continue;
        b200_conv_params p;
        if (parse_conv_attrs(n, &p)) continue;
        int64_t xd[4], wdd[4], yd[4];
        for (int k = 0; k < 4; ++k) { xd[k] = shp[n.input[0]][k]; wdd[k] = wd[k]; }
        if (b200_conv2d_out_dims(xd, wdd, &p, yd)) continue;
        shp[n.output[0]] = {yd[0], yd[1], yd[2], yd[3]};
      } else if (n.op_type == "MaxPool" && have(n.input[0])) {
        b200_pool_params p;
        if (parse_pool_attrs(n, &p)) continue;
        int64_t xd[4], yd[4];
        for (int k = 0; k < 4; ++k) xd[k] = shp[n.input[0]][k];
        if (b200_maxpool2d_out_dims(xd, &p, yd)) continue;
        shp[n.output[0]] = {yd[0], yd[1], yd[2], yd[3]};
      } else if ((n.op_type == "Relu" || n.op_type == "Dropout" || n.op_type == "Add") && have(n.input[0])) {
        shp[n.output[0]] = shp[n.input[0]];
      } else if (n.op_type == "Concat" && n.input.size() == 2 && have(n.input[0]) && have(n.input[1])) {
        auto a = shp[n.input[0]], b = shp[n.input[1]];
        const int64_t ax = attr_i(n, "axis", 1);
        if (ax < 0 || ax > 3) continue;
        a[(size_t)ax] += b[(size_t)ax];
        shp[n.output[0]] = a;
      }
    }
    // fix channel offsets; drop redirects whose shapes are unknown or whose offsets break 16-byte alignment
    std::vector<std::string> drop;
    for (auto& n : m->wm.nodes) {
      if (n.op_type != "Concat" || n.input.size() != 2) continue;
      if (!redirect.count(n.input[0]) || redirect[n.input[0]].first != n.output[0]) continue;
      bool ok = shp.count(n.input[0]) && shp.count(n.input[1]) && shp.count(n.output[0]);
      if (ok) {
        auto a = shp[n.input[0]], b = shp[n.input[1]];
        // a slice that does not start on a 16-byte boundary would push its producer to scalar stores and make its
        // consumer ineligible for the TMA / 128-bit gather paths: such a Concat is copied instead
        ok = a[0] == b[0] && a[2] == b[2] && a[3] == b[3] && a[1] % 4 == 0;
        if (ok) {
          redirect[n.input[1]].second = (int)a[1];
          Val parent; parent.rank = 4;
          for (int k = 0; k < 4; ++k) parent.dims[k] = shp[n.output[0]][k];
          env[n.output[0]] = parent;
        }
      }
      if (!ok) { drop.push_back(n.input[0]); drop.push_back(n.input[1]); }
    }
    for (auto& d : drop) redirect.erase(d);
  }

  plan->in_stem = nullptr;
  bool fused_cnn = false;
  B200_TRY(try_plan_mnist8(&fused_cnn));
  for (size_t i = 0; i < m->wm.nodes.size(); ++i) {
    if (consumed.count(i)) continue;
    const WireNode& n = m->wm.nodes[i];
    if (n.input.empty() || n.output.empty()) B200_FAIL(B200_EINVAL, "node %zu (%s) has no inputs/outputs", i, n.op_type.c_str());
    const std::string& op = n.op_type;
    if (op == "Conv") B200_TRY(do_conv(i));
    else if (op == "Relu") B200_TRY(do_relu(i));
    else if (op == "MaxPool") B200_TRY(do_maxpool(i));
    else if (op == "Concat") B200_TRY(do_concat(i));
    else if (op == "Dropout") B200_TRY(do_dropout(i));
    else if (op == "GlobalAveragePool") B200_TRY(do_gap(i));
    else if (op == "Softmax") B200_TRY(do_softmax(i));
    else if (op == "Reshape") B200_TRY(do_reshape(i));
    else if (op == "Add") B200_TRY(do_add(i));
    else if (op == "MatMul") B200_TRY(do_matmul(i));
    else B200_FAIL(B200_EUNSUPPORTED, "INFERENCE OPERATION '%s' NOT FOUND FOR NODE %s", op.c_str(), n.name.c_str());  // model_inference.rs:158
  }
  // ---- result: the value the reference prints (Softmax result / final Add) == graph.output[0] for both models
  std::string out_name = m->wm.outputs.empty() ? m->wm.nodes.back().output[0] : m->wm.outputs[0].name;
  auto it = env.find(out_name);
  if (it == env.end() || !it->second.planned) {
    // fall back to the last produced activation (e.g. the output was folded into a fused step)
    for (auto r = m->wm.nodes.rbegin(); r != m->wm.nodes.rend() && (it == env.end() || !it->second.planned); ++r)
      it = env.find(r->output[0]);
    if (it == env.end() || !it->second.planned) B200_FAIL(B200_EINVAL, "graph output %s was not produced", out_name.c_str());
  }
  Val& o = it->second;
  ++step_counter;            // the result outlives every launch (read by the caller after the run)
  touch(&o);
  keep_forever(o.alloc);
  plan->out_per_image = o.v.numel() / std::max<int64_t>(1, B);
  plan->out_view = o.v;
  if (o.rank == 4 && o.v.H * o.v.W > 1 && o.v.C > 1) {
    plan->out_needs_nchw = true;
    plan->out_ptr = arena_alloc((size_t)o.v.numel(), nullptr, true);
  } else if (!o.v.dense()) {
    plan->out_needs_nchw = true;  // strided rows -> dense copy via the same kernel (HW == 1 or C == 1)
    plan->out_ptr = arena_alloc((size_t)o.v.numel(), nullptr, true);
  } else {
    plan->out_needs_nchw = false;
    plan->out_ptr = o.v.p;
  }
  if (plan->out_needs_nchw && !dry) {
    TView ov = o.v; float* dst = plan->out_ptr;
    add_step("output_to_nchw", "reshape_nchw", 0, 8.0 * (double)ov.numel(), [ov, dst](cudaStream_t st) { return launch_rows_to_nchw(ov, dst, st); });
  }
  return 0;
}

int build_plan(b200_model* m, int64_t batch, Plan** out, bool safe = false) {
  if (batch <= 0) B200_FAIL(B200_EINVAL, "batch must be positive");
  const int64_t key = safe ? -batch : batch;
  auto it = m->plans.find(key);
  if (it != m->plans.end()) { it->second->last_used = ++m->use_counter; *out = it->second.get(); return 0; }
  // one arena per planned batch size: keep the most recently used few (a server that sees many batch sizes must not
  // accumulate arenas); kernels of an evicted plan may still be in flight, so drain the streams first
  constexpr size_t MAX_PLANS = 6;
  if (m->plans.size() >= MAX_PLANS) {
    cudaStreamSynchronize(m->ctx->stream);
    if (m->h2d) cudaStreamSynchronize(m->h2d);
    if (m->d2h) cudaStreamSynchronize(m->d2h);
    while (m->plans.size() >= MAX_PLANS) {
      auto lru = m->plans.begin();
      for (auto p = m->plans.begin(); p != m->plans.end(); ++p) if (p->second->last_used < lru->second->last_used) lru = p;
      m->plans.erase(lru);
    }
  }
  std::unique_ptr<Plan> plan(new Plan());
  plan->batch = batch;
  plan->safe = safe;
  Planner pl;
  pl.m = m; pl.plan = plan.get(); pl.B = batch; pl.safe = safe;
  pl.dry = true;
  B200_TRY(pl.run());
  plan->arena_bump = pl.arena_cursor;         // what a bump allocator (round 1) would have taken
  const size_t coloured = pl.colour();
  plan->arena.reset(new DevBuf());
  plan->arena->bytes = coloured + 256;
  if (cudaMalloc((void**)&plan->arena->p, plan->arena->bytes) != cudaSuccess) {
    cudaGetLastError();
    B200_FAIL(B200_ENOMEM, "cudaMalloc(%zu) for the activation arena (batch %lld)", plan->arena->bytes, (long long)batch);
  }
  plan->arena_used = coloured;
  pl.dry = false;
  B200_TRY(pl.run());
  B200_CUDA(cudaStreamSynchronize(m->ctx->stream));  // constant uploads / weight preparation done
  plan->last_used = ++m->use_counter;
  *out = plan.get();
  m->plans[key] = std::move(plan);
  return 0;
}

int run_steps(b200_model* m, Plan* plan, cudaStream_t st) {
  for (auto& s : plan->steps) {
    int rc = s.run(st);
    if (rc) return rc;
  }
  (void)m;
  return 0;
}

// `input_consumed` (optional) is recorded right after the input transform: from then on d_in may be overwritten
int run_plan(b200_model* m, Plan* plan, const float* d_in, float* d_out, cudaEvent_t input_consumed = nullptr) {
  cudaStream_t st = m->ctx->stream;
  // 1. input: logical NCHW (what the reference's manage_input_data holds, utils.rs:29-45) -> channels-last rows
  int* flag = (m->opt_finite_guard && !plan->safe) ? m->d_nonfinite : nullptr;
  if (plan->in_stem) {
    B200_TRY(plan->in_stem(d_in, st, flag));
    m->ctx->launches++;
  } else if (plan->in_direct) {
    if (flag) { B200_TRY(launch_nonfinite_scan(d_in, (size_t)plan->in_view.numel(), flag, st)); m->ctx->launches++; }
    B200_CUDA(cudaMemcpyAsync(plan->in_view.p, d_in, (size_t)plan->in_view.numel() * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else if (plan->in_s2d) {
    B200_TRY(launch_nchw_to_s2d(d_in, plan->in_view.N, plan->in_c0, plan->in_h0, plan->in_w0, plan->in_view.p, st, flag));
    m->ctx->launches++;
  } else {
    if (flag) { B200_TRY(launch_nonfinite_scan(d_in, (size_t)plan->batch * m->in_dims[0] * m->in_dims[1] * m->in_dims[2] * m->in_dims[3], flag, st)); m->ctx->launches++; }
    B200_TRY(launch_nchw_to_rows(d_in, plan->in_view, plan->in_zero_pad, st));
    m->ctx->launches++;
  }
  if (input_consumed) B200_CUDA(cudaEventRecord(input_consumed, st));
  // 2. the node walk, replayed as one CUDA graph
  if (plan->steps.empty()) {
    // everything ran in the input stage (the one-launch MNIST form)
  } else if (m->opt_cuda_graph) {
    if (!plan->graph_exec) {
      cudaGraph_t g = nullptr;
      B200_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      int rc = run_steps(m, plan, st);
      cudaError_t e = cudaStreamEndCapture(st, &g);
      if (rc) { if (g) cudaGraphDestroy(g); return rc; }
      if (e != cudaSuccess) B200_FAIL(B200_ECUDA, "graph capture failed: %s", cudaGetErrorString(e));
      e = cudaGraphInstantiate(&plan->graph_exec, g, 0);
      cudaGraphDestroy(g);
      if (e != cudaSuccess) B200_FAIL(B200_ECUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
    }
    B200_CUDA(cudaGraphLaunch(plan->graph_exec, st));
  } else {
    B200_TRY(run_steps(m, plan, st));
  }
  m->ctx->launches += (int64_t)plan->steps.size();
  // 3. result
  if (d_out != plan->out_ptr)
    B200_CUDA(cudaMemcpyAsync(d_out, plan->out_ptr, (size_t)plan->batch * plan->out_per_image * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // namespace
}  // namespace b200

extern "C" {

int b200_model_load_onnx(b200_ctx* ctx, const uint8_t* bytes, size_t len, b200_model** out) {
  if (!ctx || !bytes || !out) B200_FAIL(B200_EINVAL, "NULL argument");
  std::unique_ptr<b200_model> m(new b200_model());
  m->ctx = ctx;
  std::string err;
  if (!parse_model(bytes, len, &m->wm, &err)) B200_FAIL(B200_EPARSE, "ONNX parse error: %s", err.c_str());
  if (m->wm.nodes.empty()) B200_FAIL(B200_EPARSE, "ONNX graph has no nodes");
  // the single data input: the graph.input entries that are not initializers (utils.rs:35)
  int found = 0;
  for (auto& vi : m->wm.inputs) {
    if (m->wm.find_initializer(vi.name)) continue;
    ++found;
    m->input_name = vi.name;
    if (vi.dims.size() != 4) B200_FAIL(B200_EUNSUPPORTED, "input %s: the reference needs 4 static dims (utils.rs:36-40)", vi.name.c_str());
    for (int i = 0; i < 4; ++i) {
      if (vi.dims[i] <= 0 && i > 0) B200_FAIL(B200_EUNSUPPORTED, "input %s: symbolic dim (utils.rs:67 panics on DimParam)", vi.name.c_str());
      m->in_dims[i] = vi.dims[i] > 0 ? vi.dims[i] : 1;
    }
  }
  if (found != 1) B200_FAIL(B200_EUNSUPPORTED, "expected exactly one non-initializer graph input, found %d", found);
  Guard g(ctx);
  B200_CUDA(cudaMalloc((void**)&m->d_nonfinite, sizeof(int)));
  B200_CUDA(cudaMemsetAsync(m->d_nonfinite, 0, sizeof(int), ctx->stream));
  B200_CUDA(cudaHostAlloc((void**)&m->h_nonfinite, sizeof(int), cudaHostAllocDefault));
  *m->h_nonfinite = 0;
  Plan* p = nullptr;
  B200_TRY(build_plan(m.get(), 1, &p));  // validates every node up front (unknown op / attribute errors surface here)
  m->out_per_image = p->out_per_image;
  ctx_retain(ctx);
  *out = m.release();
  return 0;
}

int b200_model_load_file(b200_ctx* ctx, const char* path, b200_model** out) {
  if (!path) B200_FAIL(B200_EINVAL, "path is NULL");
  std::ifstream f(path, std::ios::binary);
  if (!f) B200_FAIL(B200_EINVAL, "cannot open %s", path);
  std::vector<char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  return b200_model_load_onnx(ctx, (const uint8_t*)buf.data(), buf.size(), out);
}

int b200_model_free(b200_model* m) {
  if (!m) return 0;
  b200_ctx* ctx = m->ctx;
  {
    Guard g(ctx);
    cudaStreamSynchronize(ctx->stream);
    if (m->h2d) cudaStreamSynchronize(m->h2d);
    if (m->d2h) cudaStreamSynchronize(m->d2h);
    delete m;
  }
  ctx_release(ctx);
  return 0;
}

int b200_model_io(const b200_model* m, int64_t in_chw[3], int64_t* out_per_image) {
  if (!m) B200_FAIL(B200_EINVAL, "model is NULL");
  if (in_chw) { in_chw[0] = m->in_dims[1]; in_chw[1] = m->in_dims[2]; in_chw[2] = m->in_dims[3]; }
  if (out_per_image) *out_per_image = m->out_per_image;
  return 0;
}

int b200_model_set_option(b200_model* m, const char* key, int64_t value) {
  if (!m || !key) B200_FAIL(B200_EINVAL, "NULL argument");
  Guard g(m->ctx);
  std::string k(key);
  if (k == "cuda_graph") m->opt_cuda_graph = value ? 1 : 0;
  else if (k == "conv_path") {
    if (value < 0 || value > 2) B200_FAIL(B200_EINVAL, "conv_path must be 0, 1 or 2");
    if (m->opt_conv_path != (int)value) { cudaStreamSynchronize(m->ctx->stream); m->plans.clear(); }
    m->opt_conv_path = (int)value;
  } else if (k == "pdl") {
    if (m->opt_pdl != (value ? 1 : 0)) { cudaStreamSynchronize(m->ctx->stream); m->plans.clear(); }
    m->opt_pdl = value ? 1 : 0;
  } else if (k == "pool_fusion") {
    if (m->opt_pool_fusion != (value ? 1 : 0)) { cudaStreamSynchronize(m->ctx->stream); m->plans.clear(); }
    m->opt_pool_fusion = value ? 1 : 0;
  } else if (k == "fire_fusion") {
    if (m->opt_fire_fusion != (value ? 1 : 0)) { cudaStreamSynchronize(m->ctx->stream); m->plans.clear(); }
    m->opt_fire_fusion = value ? 1 : 0;
  } else if (k == "s2d") {
    if (m->opt_s2d != (value ? 1 : 0)) { cudaStreamSynchronize(m->ctx->stream); m->plans.clear(); }
    m->opt_s2d = value ? 1 : 0;
  } else if (k == "alt_order") {
    if (m->opt_alt_order != (value ? 1 : 0)) { cudaStreamSynchronize(m->ctx->stream); m->plans.clear(); }
    m->opt_alt_order = value ? 1 : 0;
  } else if (k == "fused_cnn") {
    if (value < 0 || value > 2) B200_FAIL(B200_EINVAL, "fused_cnn must be 0, 1 or 2");
    if (m->opt_fused_cnn != (int)value) { cudaStreamSynchronize(m->ctx->stream); m->plans.clear(); }
    m->opt_fused_cnn = (int)value;
  } else if (k == "finite_guard") m->opt_finite_guard = value ? 1 : 0;
  else if (k == "verbose") m->opt_verbose = value ? 1 : 0;
  else B200_FAIL(B200_EINVAL, "unknown option %s", key);
  return 0;
}

int b200_model_run_device(b200_model* m, const float* d_in, int64_t batch, float* d_out) {
  if (!m || !d_in || !d_out) B200_FAIL(B200_EINVAL, "NULL argument");
  Guard g(m->ctx);
  Plan* p = nullptr;
  B200_TRY(build_plan(m, batch, &p));
  return run_plan(m, p, d_in, d_out);
}

int b200_model_run(b200_model* m, const float* host_in, int64_t batch, float* host_out) {
  if (!m || !host_in || !host_out) B200_FAIL(B200_EINVAL, "NULL argument");
  Guard g(m->ctx);
  Plan* p = nullptr;
  B200_TRY(build_plan(m, batch, &p));
  const size_t in_bytes = (size_t)batch * m->in_dims[0] * m->in_dims[1] * m->in_dims[2] * m->in_dims[3] * sizeof(float);
  const size_t out_bytes = (size_t)batch * p->out_per_image * sizeof(float);
  if (m->stage_in_bytes < in_bytes) {
    if (m->stage_in) cudaFree(m->stage_in);
    m->stage_in = nullptr; m->stage_in_bytes = 0;
    if (cudaMalloc((void**)&m->stage_in, in_bytes) != cudaSuccess) { cudaGetLastError(); B200_FAIL(B200_ENOMEM, "cudaMalloc(%zu) for input staging", in_bytes); }
    m->stage_in_bytes = in_bytes;
  }
  if (m->stage_out_bytes < out_bytes) {
    if (m->stage_out) cudaFree(m->stage_out);
    m->stage_out = nullptr; m->stage_out_bytes = 0;
    if (cudaMalloc((void**)&m->stage_out, out_bytes) != cudaSuccess) { cudaGetLastError(); B200_FAIL(B200_ENOMEM, "cudaMalloc(%zu) for output staging", out_bytes); }
    m->stage_out_bytes = out_bytes;
  }
  cudaStream_t st = m->ctx->stream;
  B200_CUDA(cudaMemcpyAsync(m->stage_in, host_in, in_bytes, cudaMemcpyHostToDevice, st));
  B200_TRY(run_plan(m, p, m->stage_in, m->stage_out));
  B200_CUDA(cudaMemcpyAsync(host_out, m->stage_out, out_bytes, cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaMemcpyAsync(m->h_nonfinite, m->d_nonfinite, sizeof(int), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  if (*m->h_nonfinite) {
    // finite guard: the input holds an Inf / NaN -- rerun on the fallback plan, whose arithmetic keeps them where the
    // reference keeps them (CUDA-core fp32, real taps only)
    *m->h_nonfinite = 0;
    B200_CUDA(cudaMemsetAsync(m->d_nonfinite, 0, sizeof(int), st));
    Plan* ps = nullptr;
    B200_TRY(build_plan(m, batch, &ps, /*safe=*/true));
    B200_TRY(run_plan(m, ps, m->stage_in, m->stage_out));
    B200_CUDA(cudaMemcpyAsync(host_out, m->stage_out, out_bytes, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}

int b200_model_run_async(b200_model* m, const float* host_in, int64_t batch, float* host_out) {
  if (!m || !host_in || !host_out) B200_FAIL(B200_EINVAL, "NULL argument");
  Guard g(m->ctx);
  Plan* p = nullptr;
  B200_TRY(build_plan(m, batch, &p));
  const size_t in_bytes = (size_t)batch * m->in_dims[0] * m->in_dims[1] * m->in_dims[2] * m->in_dims[3] * sizeof(float);
  const size_t out_bytes = (size_t)batch * p->out_per_image * sizeof(float);
  if (!m->h2d) {
    B200_CUDA(cudaStreamCreateWithFlags(&m->h2d, cudaStreamNonBlocking));
    B200_CUDA(cudaStreamCreateWithFlags(&m->d2h, cudaStreamNonBlocking));
    for (auto& sl : m->slots) {
      B200_CUDA(cudaEventCreateWithFlags(&sl.in_ready, cudaEventDisableTiming));
      B200_CUDA(cudaEventCreateWithFlags(&sl.in_free, cudaEventDisableTiming));
      B200_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
      B200_CUDA(cudaEventCreateWithFlags(&sl.out_done, cudaEventDisableTiming));
    }
  }
  b200_model::Slot& sl = m->slots[m->seq & 1];
  const bool reused = m->seq >= 2;
  m->seq++;
  if (sl.in_bytes < in_bytes || sl.out_bytes < out_bytes) {
    if (reused) { B200_CUDA(cudaEventSynchronize(sl.out_done)); }   // growing a slot that may still be in flight
    if (sl.in_bytes < in_bytes) {
      if (sl.in) cudaFree(sl.in);
      sl.in = nullptr; sl.in_bytes = 0;
      if (cudaMalloc((void**)&sl.in, in_bytes) != cudaSuccess) { cudaGetLastError(); B200_FAIL(B200_ENOMEM, "cudaMalloc(%zu) for input staging", in_bytes); }
      sl.in_bytes = in_bytes;
    }
    if (sl.out_bytes < out_bytes) {
      if (sl.out) cudaFree(sl.out);
      sl.out = nullptr; sl.out_bytes = 0;
      if (cudaMalloc((void**)&sl.out, out_bytes) != cudaSuccess) { cudaGetLastError(); B200_FAIL(B200_ENOMEM, "cudaMalloc(%zu) for output staging", out_bytes); }
      sl.out_bytes = out_bytes;
    }
  }
  cudaStream_t st = m->ctx->stream;
  // H2D of this batch overlaps the compute of the previous one (and of the one before: the slot's input buffer is
  // free as soon as its previous occupant's input transform has run, so the copy engine never waits for a whole run)
  if (reused) { B200_CUDA(cudaStreamWaitEvent(m->h2d, sl.in_free, 0)); }
  B200_CUDA(cudaMemcpyAsync(sl.in, host_in, in_bytes, cudaMemcpyHostToDevice, m->h2d));
  B200_CUDA(cudaEventRecord(sl.in_ready, m->h2d));
  B200_CUDA(cudaStreamWaitEvent(st, sl.in_ready, 0));
  if (reused) { B200_CUDA(cudaStreamWaitEvent(st, sl.out_done, 0)); }
  B200_TRY(run_plan(m, p, sl.in, sl.out, sl.in_free));
  B200_CUDA(cudaEventRecord(sl.done, st));
  B200_CUDA(cudaStreamWaitEvent(m->d2h, sl.done, 0));
  B200_CUDA(cudaMemcpyAsync(host_out, sl.out, out_bytes, cudaMemcpyDeviceToHost, m->d2h));
  B200_CUDA(cudaEventRecord(sl.out_done, m->d2h));
  return 0;
}

int b200_model_sync(b200_model* m) {
  if (!m) B200_FAIL(B200_EINVAL, "model is NULL");
  Guard g(m->ctx);
  B200_CUDA(cudaStreamSynchronize(m->ctx->stream));
  if (m->h2d) B200_CUDA(cudaStreamSynchronize(m->h2d));
  if (m->d2h) B200_CUDA(cudaStreamSynchronize(m->d2h));
  if (m->opt_finite_guard) {
    B200_CUDA(cudaMemcpy(m->h_nonfinite, m->d_nonfinite, sizeof(int), cudaMemcpyDeviceToHost));
    if (*m->h_nonfinite) {
      *m->h_nonfinite = 0;
      B200_CUDA(cudaMemset(m->d_nonfinite, 0, sizeof(int)));
      B200_FAIL(B200_EUNSUPPORTED, "a batch run through b200_model_run_async / b200_model_run_device since the last sync held an Inf or NaN input: the "
                "split-precision tensor-core path is exact for finite data only -- rerun that batch through b200_model_run (it falls back to the "
                "CUDA-core fp32 plan by itself) or set the option conv_path=1");
    }
  }
  return 0;
}

// inference() over several GPUs of one box: images are independent units on this path, so the
// batch is split contiguously (the first batch % n shards get one extra image), every shard runs through the pipelined
// host-to-host entry of its own model / context / device on its own host thread, and the logits land in host_out in
// shard order.  No data-path collective and no dependency on torch or NCCL: the only exchange is each device's D2H copy
// of its logits rows into the caller's buffer.
int b200_model_run_sharded(b200_model* const* models, int n, const float* host_in, int64_t batch, float* host_out) {
  if (!models || n <= 0 || !host_in || !host_out) B200_FAIL(B200_EINVAL, "NULL / empty argument");
  if (batch < 0) B200_FAIL(B200_EINVAL, "batch must be non-negative");
  for (int i = 0; i < n; ++i) {
    if (!models[i]) B200_FAIL(B200_EINVAL, "models[%d] is NULL", i);
    for (int k = 1; k < 4; ++k)
      if (models[i]->in_dims[k] != models[0]->in_dims[k]) B200_FAIL(B200_EINVAL, "models[%d] has a different input shape", i);
    if (models[i]->out_per_image != models[0]->out_per_image) B200_FAIL(B200_EINVAL, "models[%d] has a different output size", i);
    for (int j = 0; j < i; ++j)
      if (models[j]->ctx == models[i]->ctx) B200_FAIL(B200_EINVAL, "models[%d] and models[%d] share a context: one context per shard", j, i);
  }
  const int64_t in_per = models[0]->in_dims[0] * models[0]->in_dims[1] * models[0]->in_dims[2] * models[0]->in_dims[3];
  const int64_t out_per = models[0]->out_per_image;
  const int64_t base = batch / n, rem = batch % n;
  std::vector<int> rcs((size_t)n, 0);
  std::vector<std::string> errs((size_t)n);
  auto shard = [&](int i) {
    const int64_t lo = i * base + std::min<int64_t>(i, rem), cnt = base + (i < rem ? 1 : 0);
    if (cnt == 0) return;
    int rc = b200_model_run_async(models[i], host_in + lo * in_per, cnt, host_out + lo * out_per);
    if (rc == 0) rc = b200_model_sync(models[i]);
    if (rc) { rcs[(size_t)i] = rc; errs[(size_t)i] = b200_last_error(); }   // the message is thread-local: carry it over
  };
  std::vector<std::thread> th;
  for (int i = 1; i < n; ++i) th.emplace_back(shard, i);
  shard(0);
  for (auto& t : th) t.join();
  for (int i = 0; i < n; ++i)
    if (rcs[(size_t)i]) B200_FAIL(rcs[(size_t)i], "shard %d of %d: %s", i, n, errs[(size_t)i].c_str());
  return 0;
}

int b200_model_arena_bytes(b200_model* m, int64_t batch, int64_t* arena_bytes, int64_t* no_reuse_bytes) {
  if (!m) B200_FAIL(B200_EINVAL, "model is NULL");
  Guard g(m->ctx);
  Plan* p = nullptr;
  B200_TRY(build_plan(m, batch, &p));
  if (arena_bytes) *arena_bytes = (int64_t)p->arena_used;
  if (no_reuse_bytes) *no_reuse_bytes = (int64_t)p->arena_bump;
  return 0;
}

int64_t b200_model_launches_per_run(b200_model* m, int64_t batch) {
  if (!m) return B200_EINVAL;
  Guard g(m->ctx);
  Plan* p = nullptr;
  if (build_plan(m, batch, &p)) return B200_EINVAL;
  return (int64_t)p->steps.size() + ((p->in_direct && !p->in_stem) ? 0 : 1);
}

int b200_model_profile(b200_model* m, int64_t batch, int iters, int flush_l2, char* buf, size_t cap) {
  if (!m || !buf || cap < 64) B200_FAIL(B200_EINVAL, "bad arguments");
  if (iters < 1) iters = 1;
  Guard g(m->ctx);
  Plan* p = nullptr;
  B200_TRY(build_plan(m, batch, &p));
  cudaStream_t st = m->ctx->stream;
  const size_t flush_bytes = 256u << 20;
  if (flush_l2 && !m->ctx->l2_flush) {
    if (cudaMalloc((void**)&m->ctx->l2_flush, flush_bytes) != cudaSuccess) { cudaGetLastError(); B200_FAIL(B200_ENOMEM, "cudaMalloc for the L2 flush buffer"); }
  }
  cudaEvent_t e0, e1;
  B200_CUDA(cudaEventCreate(&e0));
  B200_CUDA(cudaEventCreate(&e1));
  std::ostringstream js;
  js << "[";
  // The input stage (layout transform, or the fused stem that reads the caller's NCHW input) is part of every run but not
  // of the captured launch list: time it as entry 0, on a scratch input of the right size (contents do not matter).
  int rc = 0;
  {
    const size_t in_floats = (size_t)batch * m->in_dims[0] * m->in_dims[1] * m->in_dims[2] * m->in_dims[3];
    float* scratch = nullptr;
    const bool has_stage = p->in_stem || !p->in_direct;
    if (has_stage) {
      if (cudaMalloc((void**)&scratch, in_floats * sizeof(float)) != cudaSuccess) { cudaGetLastError(); B200_FAIL(B200_ENOMEM, "cudaMalloc for the profile's scratch input"); }
      cudaMemsetAsync(scratch, 0, in_floats * sizeof(float), st);
      auto stage = [&]() -> int {
        if (p->in_stem) return p->in_stem(scratch, st, nullptr);
        if (p->in_s2d) return launch_nchw_to_s2d(scratch, p->in_view.N, p->in_c0, p->in_h0, p->in_w0, p->in_view.p, st);
        return launch_nchw_to_rows(scratch, p->in_view, p->in_zero_pad, st);
      };
      double total_ms = 0;
      rc = stage();
      for (int it = 0; it < iters && rc == 0; ++it) {
        if (flush_l2 == 1) launch_fill_zero(m->ctx->l2_flush, flush_bytes / sizeof(float), st);
        cudaEventRecord(e0, st);
        rc = stage();
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess) { rc = B200_ECUDA; set_error("profile: input stage failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        total_ms += ms;
      }
      cudaFree(scratch);
      const double out_floats = (double)p->in_view.pixels() * p->in_view.ld;
      js << "{\"name\":\"" << (p->in_stem ? p->fused_name.c_str() : p->in_s2d ? "input_to_space_to_depth" : "input_to_channels_last")
         << "\",\"kind\":\"" << (p->in_stem ? "mnist8_fused" : "input_transform") << "\",\"ms\":" << (total_ms / iters)
         << ",\"flops\":" << (p->in_stem ? p->fused_flops : 0.0)
         << ",\"bytes\":" << (p->in_stem ? p->fused_bytes : 4.0 * ((double)in_floats + out_floats)) << "}";
      if (!p->steps.empty()) js << ",";
    }
  }
  if (flush_l2 == 2) {
    // in-order mode: the whole launch list runs `iters` times in model order and every launch is timed where it
    // stands, so it sees the cache state it sees in a real run (its input was just written by its predecessor)
    std::vector<double> acc(p->steps.size(), 0.0);
    for (int it = -1; it < iters && rc == 0; ++it) {   // it == -1: warm pass
      for (size_t i = 0; i < p->steps.size() && rc == 0; ++i) {
        cudaEventRecord(e0, st);
        rc = p->steps[i].run(st);
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess) { rc = B200_ECUDA; set_error("profile: step %s failed: %s", p->steps[i].name.c_str(), cudaGetErrorString(cudaGetLastError())); break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 0) acc[i] += ms;
      }
    }
    for (size_t i = 0; i < p->steps.size() && rc == 0; ++i) {
      Step& s = p->steps[i];
      if (i) js << ",";
      js << "{\"name\":\"" << s.name << "\",\"kind\":\"" << s.kind << "\",\"ms\":" << (acc[i] / iters) << ",\"flops\":" << s.flops
         << ",\"bytes\":" << s.bytes << "}";
    }
  } else
  for (size_t i = 0; i < p->steps.size() && rc == 0; ++i) {
    Step& s = p->steps[i];
    double total_ms = 0;
    rc = s.run(st);  // warm
    for (int it = 0; it < iters && rc == 0; ++it) {
      if (flush_l2) launch_fill_zero(m->ctx->l2_flush, flush_bytes / sizeof(float), st);
      cudaEventRecord(e0, st);
      rc = s.run(st);
      cudaEventRecord(e1, st);
      if (cudaEventSynchronize(e1) != cudaSuccess) { rc = B200_ECUDA; set_error("profile: step %s failed: %s", s.name.c_str(), cudaGetErrorString(cudaGetLastError())); break; }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      total_ms += ms;
    }
    if (i) js << ",";
    js << "{\"name\":\"" << s.name << "\",\"kind\":\"" << s.kind << "\",\"ms\":" << (total_ms / iters) << ",\"flops\":" << s.flops
       << ",\"bytes\":" << s.bytes << "}";
  }
  js << "]";
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc) return rc;
  std::string out = js.str();
  if (out.size() + 1 > cap) B200_FAIL(B200_EINVAL, "profile buffer too small: need %zu bytes", out.size() + 1);
  memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

int b200_tensorproto_read(const uint8_t* bytes, size_t len, float* out, size_t cap, int64_t* dims_out, int* rank_out, size_t* n) {
  if (!bytes || !n) B200_FAIL(B200_EINVAL, "NULL argument");
  WireTensor t;
  std::string err;
  if (!parse_tensor(bytes, len, &t, &err)) B200_FAIL(B200_EPARSE, "%s", err.c_str());
  if (t.f32.empty() && !t.i64.empty()) B200_FAIL(B200_EUNSUPPORTED, "TensorProto holds int64 data; read_input_data (main.rs:44-53) reads f32");
  *n = t.f32.size();
  if (rank_out) *rank_out = (int)t.dims.size();
  if (dims_out) for (size_t i = 0; i < t.dims.size() && i < 8; ++i) dims_out[i] = t.dims[i];
  if (out) memcpy(out, t.f32.data(), std::min(cap, t.f32.size()) * sizeof(float));
  return 0;
}

}  // extern "C"
