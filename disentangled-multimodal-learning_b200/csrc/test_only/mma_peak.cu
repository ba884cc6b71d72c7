// TEST-ONLY: issue-rate microbenchmark of the legacy warp-level tensor path (mma.sync.m16n8k16 bf16, fp32 accumulate) - the
// denominator the position-bias MLP kernels of csrc/deform2d_bias.cu are measured against.  Every warp keeps
// `kChains` independent accumulators; operands never change (no memory traffic).
#include "../common.cuh"

namespace {
template <int kChains>
__global__ void __launch_bounds__(256) mma_peak_kernel(int iters, float* out) {
  float acc[kChains][4];
#pragma unroll
  for (int c = 0; c < kChains; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
  uint32_t a[4] = {0x3f803f80u + threadIdx.x, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
  const uint32_t b0 = 0x3f803f80u, b1 = 0x3c003c00u + blockIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < kChains; ++c) dml::mma_bf16_16816(acc[c], a, b0, b1);
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kChains; ++c) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
  if (s == 123.456f) out[0] = s;
}
}  // namespace

extern "C" int dml_test_mma_sync_peak(int ctas, int chains, int iters, float* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (chains == 4) mma_peak_kernel<4><<<ctas, 256, 0, st>>>(iters, out);
  else if (chains == 8) mma_peak_kernel<8><<<ctas, 256, 0, st>>>(iters, out);
  else if (chains == 16) mma_peak_kernel<16><<<ctas, 256, 0, st>>>(iters, out);
  else return DML_EINVAL;
  DML_RETURN_LAUNCH();
}
