// Minimal protobuf wire reader for the subset of ONNX the reference consumes.
// Replaces the third-party crates `onnx-protobuf = "0.2.3"` / `protobuf = "=3.4.0"` used at
// main.rs:29-30 (ModelProto::parse_from_bytes) and main.rs:50 (TensorProto::parse_from_bytes).
// Field numbers are those of the schema the reference ships, models/onnx.proto (cited per struct).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace b200 {

struct WireTensor {  // TensorProto, onnx.proto:527-595
  std::string name;
  std::vector<int64_t> dims;      // field 1
  int32_t data_type = 1;          // field 2 (1 = FLOAT, 7 = INT64)
  std::vector<float> f32;         // raw_data (9) as LE f32, else float_data (4)   [utils.rs:128-137]
  std::vector<int64_t> i64;       // int64_data (7), or raw_data when data_type == INT64 [utils.rs:138-142]
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : dims) n *= d;
    return n;
  }
};

struct WireAttr {  // AttributeProto, onnx.proto:142-173
  std::string name;
  int32_t type = 0;               // field 20
  float f = 0.f;                  // field 2
  int64_t i = 0;                  // field 3
  std::string s;                  // field 4
  std::vector<int64_t> ints;      // field 8
  std::vector<float> floats;      // field 7
};

struct WireNode {  // NodeProto, onnx.proto:201-214
  std::vector<std::string> input, output;
  std::string name, op_type;
  std::vector<WireAttr> attr;
};

struct WireValueInfo {  // ValueInfoProto, onnx.proto:185-188 -> TypeProto.Tensor -> TensorShapeProto
  std::string name;
  int32_t elem_type = 0;
  std::vector<int64_t> dims;  // -1 for dim_param / unknown
};

struct WireModel {  // ModelProto onnx.proto:347-384, GraphProto :445-468
  int64_t ir_version = 0;
  int64_t opset = 0;
  std::string producer, graph_name;
  std::vector<WireNode> nodes;
  std::vector<WireTensor> initializers;
  std::vector<WireValueInfo> inputs, outputs;
  const WireTensor* find_initializer(const std::string& n) const {
    for (auto& t : initializers)
      if (t.name == n) return &t;
    return nullptr;
  }
  const WireValueInfo* find_input(const std::string& n) const {
    for (auto& v : inputs)
      if (v.name == n) return &v;
    return nullptr;
  }
};

// Both return false and fill `err` on malformed input; neither throws.
bool parse_model(const uint8_t* data, size_t len, WireModel* out, std::string* err);
bool parse_tensor(const uint8_t* data, size_t len, WireTensor* out, std::string* err);

}  // namespace b200
