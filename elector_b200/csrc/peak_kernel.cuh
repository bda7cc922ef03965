// peak_kernel.cuh -- measures the INT32 issue peak used as the DP kernel's roofline denominator.
#pragma once
#include <cuda_runtime.h>
namespace elector {
// register-only dependent chains, 8 independent per thread
template <bool MIXED>
__global__ void __launch_bounds__(256) int32_peak_kernel(int *out, int iters, int a, int b) {
  int v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MIXED) {  // half fma-pipe (IMAD), half alu-pipe (IADD3 / VIMNMX / LOP3)
        v0 = v0 * a + b; v1 = max(v1 + a, b); v2 = v2 * a + b; v3 = (v3 ^ a) + b;
        v4 = v4 * a + b; v5 = max(v5 + a, b); v6 = v6 * a + b; v7 = (v7 ^ a) + b;
      } else {
        v0 = max(v0 + a, b); v1 = (v1 ^ a) + b; v2 = max(v2 + a, b); v3 = (v3 ^ a) + b;
        v4 = max(v4 + a, b); v5 = (v5 ^ a) + b; v6 = max(v6 + a, b); v7 = (v7 ^ a) + b;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
}

}  // namespace elector
