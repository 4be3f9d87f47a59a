// TEST INFRASTRUCTURE (see ../../../cuda_emul.h): host stand-in for the CUB calls of pmp_cloud_build.cu and plainmp_b200.cu.
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

#include "cuda_emul.h"

namespace cub {
struct DeviceRadixSort {
  // stable ascending sort of (key, value) pairs on key bits [begin_bit, end_bit)
  template <class K, class V>
  static cudaError_t SortPairs(void* tmp, size_t& tmp_bytes, const K* keys_in, K* keys_out, const V* vals_in, V* vals_out,
                               int n, int begin_bit = 0, int end_bit = (int)sizeof(K) * 8, cudaStream_t = nullptr) {
    if (!tmp) {
      tmp_bytes = 16;
      return cudaSuccess;
    }
    const K mask = (end_bit - begin_bit >= (int)sizeof(K) * 8) ? (K)~K(0) : (K)(((K(1) << (end_bit - begin_bit)) - 1) << begin_bit);
    std::vector<int> order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return (keys_in[a] & mask) < (keys_in[b] & mask); });
    for (int i = 0; i < n; i++) {
      keys_out[i] = keys_in[order[(size_t)i]];
      vals_out[i] = vals_in[order[(size_t)i]];
    }
    return cudaSuccess;
  }
};
}  // namespace cub
