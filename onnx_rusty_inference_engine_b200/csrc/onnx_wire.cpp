// See onnx_wire.h.  A cursor-based proto2/proto3 wire decoder: varint (0), fixed64 (1),
// length-delimited (2), fixed32 (5).  Groups (3/4) are rejected.  Unknown fields are skipped.
#include "onnx_wire.h"

#include <cstring>

namespace b200 {
namespace {

struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  bool done() const { return p >= end || !ok; }
  uint64_t varint() {
    uint64_t v = 0;
    int shift = 0;
    while (p < end) {
      uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
      shift += 7;
      if (shift > 63) break;
    }
    ok = false;
    return 0;
  }
};

struct Field {
  uint32_t num;
  uint32_t wt;
  uint64_t v;           // wt 0: value; wt 1/5: raw little-endian bits
  const uint8_t* data;  // wt 2
  size_t len;
};

bool next(Cursor& c, Field& f) {
  if (c.done()) return false;
  uint64_t key = c.varint();
  if (!c.ok) return false;
  f.num = (uint32_t)(key >> 3);
  f.wt = (uint32_t)(key & 7);
  f.v = 0; f.data = nullptr; f.len = 0;
  switch (f.wt) {
    case 0: f.v = c.varint(); return c.ok;
    case 1:
      if (c.end - c.p < 8) { c.ok = false; return false; }
      memcpy(&f.v, c.p, 8); c.p += 8; return true;
    case 5: {
      if (c.end - c.p < 4) { c.ok = false; return false; }
      uint32_t t; memcpy(&t, c.p, 4); f.v = t; c.p += 4; return true;
    }
    case 2: {
      uint64_t n = c.varint();
      if (!c.ok || n > (uint64_t)(c.end - c.p)) { c.ok = false; return false; }
      f.data = c.p; f.len = (size_t)n; c.p += n; return true;
    }
    default: c.ok = false; return false;
  }
}

std::string str(const Field& f) { return std::string((const char*)f.data, f.len); }

void packed_i64(const Field& f, std::vector<int64_t>& out, bool& ok) {
  if (f.wt == 0) { out.push_back((int64_t)f.v); return; }
  if (f.wt != 2) { ok = false; return; }
  Cursor c{f.data, f.data + f.len};
  while (!c.done()) { uint64_t v = c.varint(); if (!c.ok) { ok = false; return; } out.push_back((int64_t)v); }
}

void packed_f32(const Field& f, std::vector<float>& out, bool& ok) {
  if (f.wt == 5) { uint32_t b = (uint32_t)f.v; float x; memcpy(&x, &b, 4); out.push_back(x); return; }
  if (f.wt != 2 || f.len % 4) { ok = false; return; }
  size_t n = f.len / 4, o = out.size();
  out.resize(o + n);
  memcpy(out.data() + o, f.data, f.len);  // little-endian host assumed (x86-64 / aarch64)
}

bool tensor(const uint8_t* d, size_t n, WireTensor& t) {
  Cursor c{d, d + n};
  Field f;
  bool ok = true;
  const uint8_t* raw = nullptr; size_t raw_len = 0;
  std::vector<float> fdata;
  while (next(c, f)) {
    switch (f.num) {
      case 1: packed_i64(f, t.dims, ok); break;
      case 2: t.data_type = (int32_t)f.v; break;
      case 4: packed_f32(f, fdata, ok); break;
      case 7: packed_i64(f, t.i64, ok); break;
      case 8: if (f.wt == 2) t.name = str(f); break;
      case 9: if (f.wt == 2) { raw = f.data; raw_len = f.len; } break;
      default: break;
    }
    if (!ok) return false;
  }
  if (!c.ok) return false;
  if (raw_len > 0) {  // raw_data wins, as in utils.rs:128
    if (t.data_type == 7) {
      if (raw_len % 8) return false;
      t.i64.resize(raw_len / 8);
      memcpy(t.i64.data(), raw, raw_len);
    } else {
      if (raw_len % 4) return false;
      t.f32.resize(raw_len / 4);
      memcpy(t.f32.data(), raw, raw_len);
    }
  } else if (!fdata.empty()) {
    t.f32.swap(fdata);
  }
  return true;
}

bool attribute(const uint8_t* d, size_t n, WireAttr& a) {
  Cursor c{d, d + n};
  Field f;
  bool ok = true;
  while (next(c, f)) {
    switch (f.num) {
      case 1: if (f.wt == 2) a.name = str(f); break;
      case 2: if (f.wt == 5) { uint32_t b = (uint32_t)f.v; memcpy(&a.f, &b, 4); } break;
      case 3: a.i = (int64_t)f.v; break;
      case 4: if (f.wt == 2) a.s = str(f); break;
      case 7: packed_f32(f, a.floats, ok); break;
      case 8: packed_i64(f, a.ints, ok); break;
      case 20: a.type = (int32_t)f.v; break;
      default: break;
    }
    if (!ok) return false;
  }
  return c.ok;
}

bool node(const uint8_t* d, size_t n, WireNode& nd) {
  Cursor c{d, d + n};
  Field f;
  while (next(c, f)) {
    if (f.wt != 2) continue;
    switch (f.num) {
      case 1: nd.input.push_back(str(f)); break;
      case 2: nd.output.push_back(str(f)); break;
      case 3: nd.name = str(f); break;
      case 4: nd.op_type = str(f); break;
      case 5: { WireAttr a; if (!attribute(f.data, f.len, a)) return false; nd.attr.push_back(std::move(a)); break; }
      default: break;
    }
  }
  return c.ok;
}

bool value_info(const uint8_t* d, size_t n, WireValueInfo& vi) {
  Cursor c{d, d + n};
  Field f;
  while (next(c, f)) {
    if (f.num == 1 && f.wt == 2) vi.name = str(f);
    if (f.num == 2 && f.wt == 2) {  // TypeProto
      Cursor c2{f.data, f.data + f.len}; Field f2;
      while (next(c2, f2)) {
        if (f2.num != 1 || f2.wt != 2) continue;  // tensor_type
        Cursor c3{f2.data, f2.data + f2.len}; Field f3;
        while (next(c3, f3)) {
          if (f3.num == 1) vi.elem_type = (int32_t)f3.v;
          if (f3.num == 2 && f3.wt == 2) {  // TensorShapeProto
            Cursor c4{f3.data, f3.data + f3.len}; Field f4;
            while (next(c4, f4)) {
              if (f4.num != 1 || f4.wt != 2) continue;  // Dimension
              int64_t dv = -1;
              Cursor c5{f4.data, f4.data + f4.len}; Field f5;
              while (next(c5, f5)) if (f5.num == 1 && f5.wt == 0) dv = (int64_t)f5.v;
              if (!c5.ok) return false;
              vi.dims.push_back(dv);
            }
            if (!c4.ok) return false;
          }
        }
        if (!c3.ok) return false;
      }
      if (!c2.ok) return false;
    }
  }
  return c.ok;
}

bool graph(const uint8_t* d, size_t n, WireModel& m) {
  Cursor c{d, d + n};
  Field f;
  while (next(c, f)) {
    if (f.wt != 2) continue;
    switch (f.num) {
      case 1: { WireNode nd; if (!node(f.data, f.len, nd)) return false; m.nodes.push_back(std::move(nd)); break; }
      case 2: m.graph_name = str(f); break;
      case 5: { WireTensor t; if (!tensor(f.data, f.len, t)) return false; m.initializers.push_back(std::move(t)); break; }
      case 11: { WireValueInfo v; if (!value_info(f.data, f.len, v)) return false; m.inputs.push_back(std::move(v)); break; }
      case 12: { WireValueInfo v; if (!value_info(f.data, f.len, v)) return false; m.outputs.push_back(std::move(v)); break; }
      default: break;
    }
  }
  return c.ok;
}

}  // namespace

bool parse_tensor(const uint8_t* data, size_t len, WireTensor* out, std::string* err) {
  if (!tensor(data, len, *out)) { if (err) *err = "malformed TensorProto"; return false; }
  return true;
}

bool parse_model(const uint8_t* data, size_t len, WireModel* out, std::string* err) {
  Cursor c{data, data + len};
  Field f;
  bool have_graph = false;
  while (next(c, f)) {
    if (f.num == 1 && f.wt == 0) out->ir_version = (int64_t)f.v;
    else if (f.num == 2 && f.wt == 2) out->producer = str(f);
    else if (f.num == 7 && f.wt == 2) {
      if (!graph(f.data, f.len, *out)) { if (err) *err = "malformed GraphProto"; return false; }
      have_graph = true;
    } else if (f.num == 8 && f.wt == 2) {
      Cursor c2{f.data, f.data + f.len}; Field f2;
      while (next(c2, f2)) if (f2.num == 2 && f2.wt == 0) out->opset = (int64_t)f2.v;
    }
  }
  if (!c.ok) { if (err) *err = "malformed ModelProto"; return false; }
  if (!have_graph) { if (err) *err = "ModelProto has no graph"; return false; }
  return true;
}

}  // namespace b200
