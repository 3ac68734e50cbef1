// Host-side types shared by the tcgen05 translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace b200 {

struct TcWeights {
  float* buf = nullptr;     // [2*Mpad][Kpad]: hi rows then lo rows, zero padded
  CUtensorMap tmap;         // 2-D tiled map over buf, box = 32 floats x BN rows, SWIZZLE_128B
  int M = 0, K = 0, Mpad = 0, Kpad = 0, BN = 0;
  ~TcWeights() { if (buf) cudaFree(buf); }
};

}  // namespace b200
