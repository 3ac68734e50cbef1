// onnx_rusty_inference_engine_bin -- the reference's binary (Cargo.toml:11-13, src/main.rs:9-53) over the C ABI.
//
//   onnx_rusty_inference_engine_bin [model.onnx input.pb expected.pb input_name...]
//
// With no arguments it runs what main.rs hard-codes (:17-20): models/squeezenet1.0-8.onnx on squeezenet_data_0.pb,
// expected squeezenet_output_0.pb, input tensor "data_0".  Like read_and_make_inference (main.rs:27-42) it parses the
// model, reads both TensorProto files (read_input_data, :44-53), runs inference() and prints the reference's result line
// (softmax_op.rs:41 / add_op.rs:104: 1-based class) followed by "Expected Data: [...]" (main.rs:41).  Additionally it
// reports whether the output matches the expected data within the north_star tolerance, and exits 1 if not.
// Links only libb200rt.so (include/b200rt.h); plain C++ host code, no CUDA or torch types.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "../../include/b200rt.h"

static bool read_file(const std::string& path, std::vector<uint8_t>* out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  out->assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
  return true;
}

static bool read_tensor_pb(const std::string& path, std::vector<float>* out) {
  std::vector<uint8_t> bytes;
  if (!read_file(path, &bytes)) { fprintf(stderr, "Cannot open input file %s\n", path.c_str()); return false; }   // main.rs:45
  size_t n = 0;
  int rank = 0;
  int64_t dims[8];
  if (b200_tensorproto_read(bytes.data(), bytes.size(), nullptr, 0, dims, &rank, &n)) {
    fprintf(stderr, "Error while deserializing the message: %s\n", b200_last_error());                              // main.rs:50
    return false;
  }
  out->resize(n);
  return b200_tensorproto_read(bytes.data(), bytes.size(), out->data(), n, dims, &rank, &n) == 0;
}

static void print_vec(const char* label, const std::vector<float>& v) {
  printf("%s[", label);
  for (size_t i = 0; i < v.size(); ++i) printf("%s%.9g", i ? ", " : "", v[i]);
  printf("]\n");
}

int main(int argc, char** argv) {
  std::string onnx_file = "models/squeezenet1.0-8.onnx", input_path = "squeezenet_data_0.pb", output_path = "squeezenet_output_0.pb";
  std::vector<std::string> input_tensor_name = {"data_0"};
  if (argc >= 4) {
    onnx_file = argv[1]; input_path = argv[2]; output_path = argv[3];
    input_tensor_name.assign(argv + 4, argv + argc);
  } else if (argc != 1) {
    fprintf(stderr, "usage: %s [model.onnx input.pb expected.pb input_name...]\n", argv[0]);
    return 2;
  }
  std::vector<float> input_data, output_data;
  if (!read_tensor_pb(input_path, &input_data) || !read_tensor_pb(output_path, &output_data)) return 2;
  b200_ctx* ctx = nullptr;
  b200_model* model = nullptr;
  if (b200_ctx_create(0, nullptr, &ctx)) { fprintf(stderr, "%s\n", b200_last_error()); return 2; }   // no CPU fallback
  if (b200_model_load_file(ctx, onnx_file.c_str(), &model)) { fprintf(stderr, "Failed to convert the file: %s\n", b200_last_error()); return 2; }   // main.rs:30
  int64_t chw[3], out_per_image = 0;
  b200_model_io(model, chw, &out_per_image);
  // the names that are initializers are skipped upstream (utils.rs:35); what remains is the single data input the model
  // handle feeds.  manage_input_data's from_shape_vec (utils.rs:40) requires the static element count:
  if ((int64_t)input_data.size() != chw[0] * chw[1] * chw[2]) {
    fprintf(stderr, "input length %zu != static model shape %lldx%lldx%lld (utils.rs:40)\n", input_data.size(), (long long)chw[0], (long long)chw[1], (long long)chw[2]);
    return 2;
  }
  std::vector<float> out((size_t)out_per_image);
  if (b200_model_run(model, input_data.data(), 1, out.data())) { fprintf(stderr, "%s\n", b200_last_error()); return 2; }
  size_t best = 0;
  for (size_t i = 1; i < out.size(); ++i) if (out[i] > out[best]) best = i;
  printf("\n%s Inference results: Class %zu-nth predicted.\n", out.size() == 1000 ? "Squeezenet1.0-8" : "MNist-8", best + 1);
  print_vec("Actual Data: ", out);
  print_vec("Expected Data: ", output_data);
  int rc = 0;
  if (output_data.size() == out.size()) {
    double worst = 0;
    size_t eb = 0;
    for (size_t i = 0; i < out.size(); ++i) {
      worst = std::fmax(worst, std::fabs((double)out[i] - output_data[i]) / (1e-5 + 1e-4 * std::fabs((double)output_data[i])));
      if (output_data[i] > output_data[eb]) eb = i;
    }
    const bool ok = worst <= 1.0 && eb == best;
    printf("Match (1e-4 rel + 1e-5 abs, argmax): %s (max err/tol %.3f)\n", ok ? "yes" : "NO", worst);
    rc = ok ? 0 : 1;
  }
  b200_model_free(model);
  b200_ctx_destroy(ctx);
  return rc;
}
