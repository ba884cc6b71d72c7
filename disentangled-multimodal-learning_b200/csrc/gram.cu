// Similarity matrices of the batch losses (utils/loss.py:25-64 PathBatchLoss, :90-143 OmicDomainScaleLoss, :220-253 BatchLoss):
//     sim[g] = A[g] B[g]^T,   A[g], B[g] = N rows (N = batch_size x world_size <= 64) of K floats,
// K = L1 L2 = 2 500 x 144 = 360 000 per head (2.9 M for the whole map): skinny products that stream the gathered attention
// maps once - HBM-bound (N / 2 FLOP per byte) - and their adjoint  dA[g][i, :] = sum_j W[g][i, j] X[g][j, :]  for the LOCAL
// rows only (utils/gather.py:16-20 keeps the local slice of the gradient).  Rows are addressed through a pointer table
// (one base pointer per row + a group stride), so the per-rank tensors of the all_gather are read where they are: no
// torch.cat copy, and `att.view(N, 8, -1).transpose(0, 1)` (loss.py:42-43) is a stride, not a copy.
//
//   gram_fwd_kernel  CTA = one K-range of one group; 128-float k-tiles staged in shared memory by a 3-deep cp.async ring,
//                    4 x 4 or 8 x 8 outputs per thread (rows interleaved so the float4 shared loads are conflict-free), fp32 FMAs;
//                    per-CTA partial matrices go to a workspace and are summed by the caller (deterministic, no atomics).
//   rows_mix_kernel  thread = 4 consecutive k of all (<= 16) output rows: reads each X row once, coalesced; W in shared memory.
#include "common.cuh"

namespace dml {
namespace gr {

constexpr int kThreads = 256;
constexpr int kKc = 128;           // floats of K per stage
constexpr int kStages = 3;         // cp.async ring depth
constexpr int kPitch = kKc + 4;    // shared-memory row pitch: 16-byte aligned, consecutive rows 4 banks apart

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;      // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

template <int NP, int TR>      // NP padded rows, TR x TR outputs per thread
__global__ void __launch_bounds__(kThreads)
gram_fwd_kernel(const float* const* __restrict__ a_rows, long long a_gs, const float* const* __restrict__ b_rows, long long b_gs,
                int same, int N, long long K, long long k_per_cta, float* __restrict__ part) {
  constexpr int TI = NP / TR;                      // TI x TI thread tiles
  constexpr int KG = kThreads / (TI * TI);         // k-groups sharing a tile position
  constexpr int kLd = NP * kKc / 4 / kThreads;     // float4 loads per thread, stage and matrix
  static_assert(KG >= 1 && KG <= kKc / 4 && kLd >= 1, "tile shape");
  constexpr int kStageFloats = 2 * NP * kPitch;    // A tile [NP][kPitch], then the B tile
  extern __shared__ __align__(16) float sm[];
  const int g = blockIdx.y, split = blockIdx.x, nsplit = gridDim.x;
  const int tid = threadIdx.x;
  const int tile = tid % (TI * TI), kg = tid / (TI * TI);
  const int ti = tile / TI, tj = tile % TI;
  const long long k_begin = (long long)split * k_per_cta, k_end = min(K, k_begin + k_per_cta);

  float acc[TR][TR];
#pragma unroll
  for (int r = 0; r < TR; ++r)
#pragma unroll
    for (int c = 0; c < TR; ++c) acc[r][c] = 0.f;

  // global -> shared memory with cp.async (16 bytes per request, zero fill outside the rows / the K range), kStages deep
  const int nst = (int)((k_end - k_begin + kKc - 1) / kKc);
  auto issue = [&](int it) {
    if (it < nst) {
      const long long k0 = k_begin + (long long)it * kKc;
      float* dA = sm + (size_t)(it % kStages) * kStageFloats;
      float* dB = dA + NP * kPitch;
#pragma unroll
      for (int l = 0; l < kLd; ++l) {
        const int idx = tid + l * kThreads, row = idx / (kKc / 4), kq = (idx % (kKc / 4)) * 4;
        const bool ok = row < N && k0 + kq < k_end;      // K is a multiple of 4: a float4 never straddles k_end
        const int rr = ok ? row : 0;
        const long long kk = ok ? k0 + kq : k_begin;
        cp_async16(smem_u32(dA + row * kPitch + kq), a_rows[rr] + (size_t)g * a_gs + kk, ok);
        if (!same) cp_async16(smem_u32(dB + row * kPitch + kq), b_rows[rr] + (size_t)g * b_gs + kk, ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int it = 0; it < kStages - 1; ++it) issue(it);
  for (int it = 0; it < nst; ++it) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 2) : "memory");
    __syncthreads();            // stage `it` has landed for every thread; the slot of stage it - 1 is free again
    issue(it + kStages - 1);
    const float* sA = sm + (size_t)(it % kStages) * kStageFloats;
    const float* sB = same ? sA : sA + NP * kPitch;
    // this thread's share of the stage: k = 4 (kg + KG q); rows ti + TI r / tj + TI c (interleaved: the lanes of a warp read
    // consecutive rows, 4 banks apart - conflict-free float4 loads); 2 TR float4 loads feed 4 TR^2 FMAs
#pragma unroll 2
    for (int q = kg; q < kKc / 4; q += KG) {
      float4 a[TR], b[TR];
#pragma unroll
      for (int r = 0; r < TR; ++r) {
        a[r] = *reinterpret_cast<const float4*>(sA + (ti + TI * r) * kPitch + 4 * q);
        b[r] = *reinterpret_cast<const float4*>(sB + (tj + TI * r) * kPitch + 4 * q);
      }
#pragma unroll
      for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int c = 0; c < TR; ++c) {
          acc[r][c] = fmaf(a[r].x, b[c].x, acc[r][c]);
          acc[r][c] = fmaf(a[r].y, b[c].y, acc[r][c]);
          acc[r][c] = fmaf(a[r].z, b[c].z, acc[r][c]);
          acc[r][c] = fmaf(a[r].w, b[c].w, acc[r][c]);
        }
    }
  }
  // k-groups -> one matrix: the groups add their tiles into shared memory one after the other, then the partial [g][split][N][N]
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  float* red = sm;               // [NP][NP + 1] <= NP kPitch
  for (int w = 0; w < KG; ++w) {
    if (kg == w) {
#pragma unroll
      for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int c = 0; c < TR; ++c) {
          float* d = red + (ti + TI * r) * (NP + 1) + tj + TI * c;
          *d = (w == 0 ? 0.f : *d) + acc[r][c];
        }
    }
    __syncthreads();
  }
  float* po = part + ((size_t)g * nsplit + split) * N * N;
  for (int idx = tid; idx < N * N; idx += kThreads) po[idx] = red[(idx / N) * (NP + 1) + idx % N];
}

// out[g][i][k] = sum_j W[g][i][j] X[g][j][k]   (i < RB local rows, j < N rows of X, k < K); out rows K floats apart, groups
// out_gs apart; W float [G][RB][N].  One thread = 4 consecutive k.
template <int RB>
__global__ void __launch_bounds__(kThreads)
rows_mix_kernel(const float* __restrict__ W, const float* const* __restrict__ x_rows, long long x_gs, int nrows, int N, long long K,
                float* __restrict__ out, long long out_gs, long long out_rs) {
  __shared__ float sW[RB * 64];
  const int g = blockIdx.y;
  for (int i = threadIdx.x; i < RB * N; i += kThreads) sW[i] = (i / N) < nrows ? W[((size_t)g * nrows) * N + i] : 0.f;
  __syncthreads();
  const long long k4 = ((long long)blockIdx.x * kThreads + threadIdx.x) * 4;
  if (k4 >= K) return;
  float4 acc[RB];
#pragma unroll
  for (int i = 0; i < RB; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int j = 0; j < N; ++j) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(x_rows[j] + (size_t)g * x_gs + k4));
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const float w = sW[i * N + j];
      acc[i].x = fmaf(w, x.x, acc[i].x); acc[i].y = fmaf(w, x.y, acc[i].y);
      acc[i].z = fmaf(w, x.z, acc[i].z); acc[i].w = fmaf(w, x.w, acc[i].w);
    }
  }
#pragma unroll
  for (int i = 0; i < RB; ++i)
    if (i < nrows) *reinterpret_cast<float4*>(out + (size_t)g * out_gs + (size_t)i * out_rs + k4) = acc[i];
}

template <typename Kern>
static int set_smem(Kern k, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;
  return (int)cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

static int sm_count() {
  int dev = 0, nsm = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0) nsm = 148;
  return nsm;
}

}  // namespace gr
}  // namespace dml

extern "C" {

/* K-splits per group: about two CTAs per SM over all groups, every split a multiple of the 64-float stage */
int dml_gram_splits(int G, long long K, int nsm) {
  if (G <= 0 || K <= 0) return 0;
  if (nsm <= 0) nsm = dml::gr::sm_count();
  long long want = (2LL * nsm + G - 1) / G;
  const long long stages = (K + dml::gr::kKc - 1) / dml::gr::kKc;
  if (want > stages) want = stages;
  if (want < 1) want = 1;
  const long long per = ((stages + want - 1) / want) * dml::gr::kKc;
  return (int)((K + per - 1) / per);
}

int dml_gram_fwd(const float* const* a_rows, long long a_gs, const float* const* b_rows, long long b_gs, int G, int N, long long K,
                 float* part, void* stream) {
  using namespace dml;
  using namespace dml::gr;
  DML_CHECK_ARG(a_rows && b_rows && part && G > 0 && N > 0 && K > 0);
  if (N > 64 || (K % 4) != 0 || (a_gs % 4) != 0 || (b_gs % 4) != 0) return DML_EUNSUPPORTED;
  const int nsplit = dml_gram_splits(G, K, 0);
  const long long stages = (K + kKc - 1) / kKc;
  const long long per = ((stages + nsplit - 1) / nsplit) * kKc;
  const int same = (a_rows == b_rows && a_gs == b_gs) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(nsplit, G);
  int e = 0;
  if (N <= 16) {
    const size_t smem = sizeof(float) * kStages * 2 * 16 * kPitch;
    e = set_smem(gram_fwd_kernel<16, 4>, smem);
    if (!e) gram_fwd_kernel<16, 4><<<grid, kThreads, smem, st>>>(a_rows, a_gs, b_rows, b_gs, same, N, K, per, part);
  } else if (N <= 32) {
    const size_t smem = sizeof(float) * kStages * 2 * 32 * kPitch;
    e = set_smem(gram_fwd_kernel<32, 4>, smem);
    if (!e) gram_fwd_kernel<32, 4><<<grid, kThreads, smem, st>>>(a_rows, a_gs, b_rows, b_gs, same, N, K, per, part);
  } else {
    const size_t smem = sizeof(float) * kStages * 2 * 64 * kPitch;
    e = set_smem(gram_fwd_kernel<64, 8>, smem);
    if (!e) gram_fwd_kernel<64, 8><<<grid, kThreads, smem, st>>>(a_rows, a_gs, b_rows, b_gs, same, N, K, per, part);
  }
  if (e) return e;
  DML_RETURN_LAUNCH();
}

int dml_rows_mix(const float* W, const float* const* x_rows, long long x_gs, int G, int nrows, int N, long long K, float* out,
                 long long out_gs, long long out_rs, void* stream) {
  using namespace dml;
  using namespace dml::gr;
  DML_CHECK_ARG(W && x_rows && out && G > 0 && nrows > 0 && N > 0 && K > 0);
  if (N > 64 || nrows > 16 || (K % 4) != 0 || (x_gs % 4) != 0 || (out_gs % 4) != 0 || (out_rs % 4) != 0) return DML_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const long long nthreads = K / 4;
  dim3 grid((unsigned)((nthreads + kThreads - 1) / kThreads), G);
  if (nrows <= 4) rows_mix_kernel<4><<<grid, kThreads, 0, st>>>(W, x_rows, x_gs, nrows, N, K, out, out_gs, out_rs);
  else if (nrows <= 8) rows_mix_kernel<8><<<grid, kThreads, 0, st>>>(W, x_rows, x_gs, nrows, N, K, out, out_gs, out_rs);
  else rows_mix_kernel<16><<<grid, kThreads, 0, st>>>(W, x_rows, x_gs, nrows, N, K, out, out_gs, out_rs);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
