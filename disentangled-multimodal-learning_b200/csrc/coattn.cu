// Genomic-guided co-attention (MCAT / CMTA: models/MultiheadAttention.py:7-321 called with ONE head, embed_dim 256 and a
// handful of genomic tokens on one side - model.py:1007,1047 and 1168-1170,1229-1238) as single streaming passes over the
// patch tokens (sm_100a).
//
// With few tokens on one side, the projections of the LONG side fold into the short side exactly:
//   few queries (q_l, l < F) over S keys x_s:   score[l, s] = q_l . (W_k x_s + b_k) = (W_k^T q_l) . x_s + q_l . b_k
//                                               out_l = sum_s P[l, s] (W_v x_s + b_v) = W_v (sum_s P[l, s] x_s) + b_v
//   S queries x_s over few keys (k_f, v_f):     score[s, f] = scale (W_q x_s + b_q) . k_f = (scale W_q^T k_f) . x_s + scale b_q . k_f
//                                               out_s = W_o (sum_f P[s, f] v_f) + b_o = sum_f P[s, f] (W_o v_f) + b_o
// so the kernels read every patch row x_s (E floats) exactly once per direction, the O(F E^2) algebra on the short side
// stays with the caller, and nothing of size S x E besides x (and dx / out) touches HBM.  The raw (pre-softmax) scores the
// reference returns (need_raw=True, MultiheadAttention.py:300-303) are written as a by-product.
//
// A CTA is 16 compute warps that also drive the copy engine: 32-row stages of the long side (1 KB rows, any row stride) stream
// into a shared-memory ring with cp.async.bulk (one 32 KB copy per stage when the rows are contiguous, mbarrier transaction counts), 96-160 KB
// in flight per SM independent of register pressure; a compute warp takes 2 rows of every stage (lane = 4 consecutive floats
// of each 128-float span, packed fp32 FMAs), reduces the F dot products of its rows with a reduce-scatter over the warp
// and keeps the short-side vectors in shared memory.  Per-CTA partial sums (online-softmax state, short-side gradients) go
// to a caller-owned workspace and are reduced in a second, tiny launch or by the caller: no atomics, deterministic results.
#include "tc_common.cuh"

namespace dml {
namespace ca {

using tc::mbar_arrive;
using tc::mbar_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::mbar_wait_relaxed;

constexpr int kStageRows = 32;                // rows per ring stage: 2 per warp (16-warp forward CTAs) or 4 (8-warp backward CTAs)
constexpr int kWarpsFwd = 16, kWarpsBwd = 8;  // the backward kernels carry twice the accumulators: fewer, fatter warps

template <int V>
struct Row {
  float4 v[V];
};

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

template <int V>
__device__ __forceinline__ Row<V> ldg_row(const float* p, int lane) {
  Row<V> r;
#pragma unroll
  for (int c = 0; c < V; ++c) r.v[c] = __ldg(reinterpret_cast<const float4*>(p) + c * 32 + lane);
  return r;
}
template <int V>
__device__ __forceinline__ Row<V> lds_row(const float* p, int lane) {
  Row<V> r;
#pragma unroll
  for (int c = 0; c < V; ++c) r.v[c] = reinterpret_cast<const float4*>(p)[c * 32 + lane];
  return r;
}
template <int V>
__device__ __forceinline__ Row<V> lds_row_if(const float* p, int lane, bool valid) {      // stale ring rows read as zeros
  Row<V> r = lds_row<V>(p, lane);
  if (!valid) {
#pragma unroll
    for (int c = 0; c < V; ++c) r.v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  return r;
}
template <int V>
__device__ __forceinline__ void store_row(float* p, int lane, const Row<V>& r) {
#pragma unroll
  for (int c = 0; c < V; ++c) reinterpret_cast<float4*>(p)[c * 32 + lane] = r.v[c];
}
__device__ __forceinline__ float2 lo2(const float4& a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ float2 hi2(const float4& a) { return make_float2(a.z, a.w); }
template <int V>
__device__ __forceinline__ float dot_part(const Row<V>& a, const Row<V>& b) {
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < V; ++c) {
    s = __ffma2_rn(lo2(a.v[c]), lo2(b.v[c]), s);
    s = __ffma2_rn(hi2(a.v[c]), hi2(b.v[c]), s);
  }
  return s.x + s.y;
}
template <int V>
__device__ __forceinline__ void axpy(Row<V>& y, float a, const Row<V>& x) {
  const float2 aa = make_float2(a, a);
#pragma unroll
  for (int c = 0; c < V; ++c) {
    const float2 l = __ffma2_rn(aa, lo2(x.v[c]), lo2(y.v[c]));
    const float2 h = __ffma2_rn(aa, hi2(x.v[c]), hi2(y.v[c]));
    y.v[c] = make_float4(l.x, l.y, h.x, h.y);
  }
}
template <int V>
__device__ __forceinline__ void scale_row(Row<V>& y, float a) {
  const float2 aa = make_float2(a, a);
#pragma unroll
  for (int c = 0; c < V; ++c) {
    const float2 l = __fmul2_rn(aa, lo2(y.v[c])), h = __fmul2_rn(aa, hi2(y.v[c]));
    y.v[c] = make_float4(l.x, l.y, h.x, h.y);
  }
}
template <int V>
__device__ __forceinline__ void zero_row(Row<V>& y) {
#pragma unroll
  for (int c = 0; c < V; ++c) y.v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__host__ __device__ constexpr int pow2_ceil(int n) { int p = 1; while (p < n) p *= 2; return p; }

// All-reduce of N per-lane partial sums (every lane ends with every total).  N a power of two uses a reduce-scatter (each
// exchange halves the number of values a lane carries) followed by broadcasts: 2 N + log2(32 / N) shuffles instead of 5 N.
// own / own_idx: the one total this lane holds natively (own_idx = -1 on the duplicate lanes) - the lane that writes it out.
template <int N>
__device__ __forceinline__ void allreduce(float (&v)[N], int lane, float& own, int& own_idx) {
  if constexpr (N == 32 || N == 16 || N == 8 || N == 4 || N == 2) {
    constexpr int kSteps = N == 32 ? 5 : N == 16 ? 4 : N == 8 ? 3 : N == 4 ? 2 : 1;
    float w[N];
#pragma unroll
    for (int i = 0; i < N; ++i) w[i] = v[i];
#pragma unroll
    for (int st = 0; st < kSteps; ++st) {
      const int bit = 16 >> st;
      const bool up = lane & bit;
      const int cnt = N >> (st + 1);
#pragma unroll
      for (int i = 0; i < N / 2; ++i) {
        if (i < cnt) {
          const float send = up ? w[i] : w[i + cnt];
          const float keep = up ? w[i + cnt] : w[i];
          w[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
      }
    }
    // w[0] of a lane = the total of value (lane >> (5 - kSteps)) over the lanes that share its top kSteps lane bits: finish
    // over the remaining low bits, then hand every total to every lane
    float t = w[0];
#pragma unroll
    for (int o = (16 >> kSteps); o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = __shfl_sync(0xffffffffu, t, i << (5 - kSteps));
    own = t;
    own_idx = (lane & ((32 >> kSteps) - 1)) == 0 ? lane >> (5 - kSteps) : -1;
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
    own = 0.f;
    own_idx = lane < N ? lane : -1;
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (i == lane) own = v[i];
  }
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- the ring: kStages stages of NIN x 32 rows x E floats, full / empty mbarriers -----------------------------------------
template <int V, int NIN, int kStages, int kWarps>
struct Ring {
  static constexpr int kR = kStageRows / kWarps;
  static constexpr int E = 128 * V;
  static constexpr int kStageFloats = NIN * kStageRows * E;
  static constexpr size_t kBytes = sizeof(float) * kStages * kStageFloats + 16 * kStages;
  float* data;          // generic pointer to stage 0
  uint32_t bars;        // shared address of full[kStages], empty[kStages]
  __device__ __forceinline__ uint32_t full(int st) const { return bars + 8u * st; }
  __device__ __forceinline__ uint32_t empty(int st) const { return bars + 8u * (kStages + st); }
  __device__ __forceinline__ const float* row(int st, int in, int r) const { return data + st * kStageFloats + (in * kStageRows + r) * E; }
  __device__ __forceinline__ void init(int tid) {
    if (tid == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), kWarps); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  // Start the copies of stage `it` - rows [s_begin + 32 it, ..) of up to two row-strided inputs - once every compute warp has
  // released the slot's previous contents.  An input whose rows are contiguous (row stride = E) goes out as ONE bulk copy
  // from warp 0; otherwise every warp issues the copies of its own R rows (cp.async.bulk is a warp-uniform instruction:
  // one warp issuing 32 different rows would serialise them).
  __device__ __forceinline__ void issue(int it, const float* x, long long xs_r, const float* g, long long gs_r, int s_begin, int s_end,
                                        int warp, int lane) const {
    const bool xc = xs_r == E, gc = NIN == 1 || gs_r == E;
    if (warp != 0 && xc && gc) return;
    const int st = it % kStages;
    if (it >= kStages) mbar_wait(empty(st), ((it / kStages) - 1) & 1);
    const int row0 = s_begin + it * kStageRows, nrows = min(kStageRows, s_end - row0);
    if (warp == 0 && lane == 0) {
      mbar_expect_tx(full(st), (uint32_t)(nrows * NIN * E * 4));
      if (xc) bulk_g2s(smem_u32(row(st, 0, 0)), x + (size_t)row0 * E, (uint32_t)(nrows * E * 4), full(st));
      if (NIN == 2 && gc) bulk_g2s(smem_u32(row(st, 1, 0)), g + (size_t)row0 * E, (uint32_t)(nrows * E * 4), full(st));
    }
    const int r = warp * kR + lane;
    if (lane < kR && r < nrows) {
      if (!xc) bulk_g2s(smem_u32(row(st, 0, r)), x + (size_t)(row0 + r) * xs_r, E * 4, full(st));
      if (NIN == 2 && !gc) bulk_g2s(smem_u32(row(st, 1, r)), g + (size_t)(row0 + r) * gs_r, E * 4, full(st));
    }
  }
  // prologue and the per-iteration top-up: stage it + kStages - 1 goes out while stage it is consumed
  __device__ __forceinline__ void prologue(int niter, const float* x, long long xs_r, const float* g, long long gs_r, int s_begin,
                                           int s_end, int warp, int lane) const {
    for (int it = 0; it < kStages - 1 && it < niter; ++it) issue(it, x, xs_r, g, gs_r, s_begin, s_end, warp, lane);
  }
  __device__ __forceinline__ void top_up(int it, int niter, const float* x, long long xs_r, const float* g, long long gs_r, int s_begin,
                                         int s_end, int warp, int lane) const {
    if (it + kStages - 1 < niter) issue(it + kStages - 1, x, xs_r, g, gs_r, s_begin, s_end, warp, lane);
  }
  __device__ __forceinline__ void wait_full(int it) const { mbar_wait(full(it % kStages), (it / kStages) & 1); }
  __device__ __forceinline__ void release(int it, int lane) const {
    __syncwarp();
    if (lane == 0) mbar_arrive(empty(it % kStages));
  }
};

// =================================================================================================================
// few queries over many keys, forward:   raw[b, f, s] = qt[b, f] . x[b, s] + c[b, f];   online softmax over s;
// partial (M, L, acc = sum_s exp(raw - M) x_s) per CTA -> part[b][chunk][f][E + 2] = {acc[E], M, L}
// =================================================================================================================
template <int F, int V, int kStages>
__global__ void __launch_bounds__(32 * kWarpsFwd, 1)
fq_fwd_kernel(const float* __restrict__ X, long long xs_b, long long xs_r, const float* __restrict__ qt, const float* __restrict__ cv,
              float* __restrict__ raw, float* __restrict__ part, int S, int rows_per_cta) {
  constexpr int kWarps = kWarpsFwd, kThreads = 32 * kWarps;
  using RingT = Ring<V, 1, kStages, kWarps>;
  constexpr int E = 128 * V, R = RingT::kR, NP = pow2_ceil(R * F);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  RingT ring;
  ring.data = reinterpret_cast<float*>(smem_raw);
  ring.bars = smem_u32(smem_raw) + sizeof(float) * kStages * RingT::kStageFloats;
  float* s_qt = reinterpret_cast<float*>(smem_raw + RingT::kBytes);      // [F][E]; reused as the CTA accumulator at the end
  float* s_ml = s_qt + F * E;                                           // [kWarps][2 F]
  const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  ring.init(tid);
  for (int i = tid; i < F * E / 4; i += kThreads) {      // scores are kept in the exp2 domain: qt and c times log2(e)
    const float4 t = __ldg(reinterpret_cast<const float4*>(qt + (size_t)b * F * E) + i);
    reinterpret_cast<float4*>(s_qt)[i] = make_float4(t.x * kLog2e, t.y * kLog2e, t.z * kLog2e, t.w * kLog2e);
  }
  __syncthreads();
  const float* xb = X + (size_t)b * xs_b;
  const int s_begin = chunk * rows_per_cta, s_end = min(S, s_begin + rows_per_cta);
  const int niter = cdiv(s_end - s_begin, kStageRows);
  float m[F], l[F];
  Row<V> acc[F];
#pragma unroll
  for (int f = 0; f < F; ++f) { m[f] = -INFINITY; l[f] = 0.f; zero_row(acc[f]); }
  {
    float cf[F];
#pragma unroll
    for (int f = 0; f < F; ++f) cf[f] = __ldg(cv + b * F + f) * kLog2e;
    ring.prologue(niter, xb, xs_r, nullptr, 0, s_begin, s_end, warp, lane);
    for (int it = 0; it < niter; ++it) {
      const int st = it % kStages, s0 = s_begin + it * kStageRows + warp * R;
      ring.top_up(it, niter, xb, xs_r, nullptr, 0, s_begin, s_end, warp, lane);
      ring.wait_full(it);
      if (s0 < s_end) {
        Row<V> x[R];
#pragma unroll
        for (int r = 0; r < R; ++r) x[r] = lds_row_if<V>(ring.row(st, 0, warp * R + r), lane, s0 + r < s_end);
        float d[NP];
#pragma unroll
        for (int i = R * F; i < NP; ++i) d[i] = 0.f;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const Row<V> q = lds_row<V>(s_qt + f * E, lane);
#pragma unroll
          for (int r = 0; r < R; ++r) d[r * F + f] = dot_part(x[r], q);
        }
        float own;
        int oi;
        allreduce<NP>(d, lane, own, oi);
        if (oi >= 0 && oi < R * F && s0 + oi / F < s_end)      // raw scores [b, f, s]: the lane that holds total (r, f) writes it
          raw[((size_t)b * F + oi % F) * S + s0 + oi / F] = own * kLn2 + __ldg(cv + b * F + oi % F);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int f = 0; f < F; ++f) d[r * F + f] = (s0 + r < s_end) ? d[r * F + f] + cf[f] : -INFINITY;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          float mx = d[f];
#pragma unroll
          for (int r = 1; r < R; ++r) mx = fmaxf(mx, d[r * F + f]);
          if (mx > m[f]) {               // warp-uniform: every lane holds the same totals
            const float sc = ex2f(m[f] - mx);      // m = -inf -> 0
            scale_row(acc[f], sc);
            l[f] *= sc;
            m[f] = mx;
          }
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float p = ex2f(d[r * F + f] - m[f]);      // masked row: exp2(-inf) = 0
            l[f] += p;
            axpy(acc[f], p, x[r]);
          }
        }
      }
      ring.release(it, lane);
    }
  }
  // CTA combine: common maximum, then the compute warps add their rescaled sums into shared memory one after the other
  if (lane == 0) {
#pragma unroll
    for (int f = 0; f < F; ++f) { s_ml[warp * 2 * F + f] = m[f]; s_ml[warp * 2 * F + F + f] = l[f]; }
  }
  __syncthreads();      // also: every warp is done reading s_qt
  float M[F], L[F], myscale[F];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    float mm = -INFINITY;
    for (int w = 0; w < kWarps; ++w) mm = fmaxf(mm, s_ml[w * 2 * F + f]);
    float ll = 0.f;
    for (int w = 0; w < kWarps; ++w) {
      const float mw = s_ml[w * 2 * F + f];
      ll += mw == -INFINITY ? 0.f : s_ml[w * 2 * F + F + f] * ex2f(mw - mm);
    }
    M[f] = mm; L[f] = ll;
    myscale[f] = m[f] == -INFINITY ? 0.f : ex2f(m[f] - mm);
  }
  for (int w = 0; w < kWarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int f = 0; f < F; ++f) {
        Row<V> t;
        if (w == 0) zero_row(t); else t = lds_row<V>(s_qt + f * E, lane);
        axpy(t, myscale[f], acc[f]);
        store_row<V>(s_qt + f * E, lane, t);
      }
    }
    __syncthreads();
  }
  float* po = part + ((size_t)b * nchunk + chunk) * F * (E + 2);
  for (int i = tid; i < F * E; i += kThreads) po[(i / E) * (E + 2) + (i % E)] = s_qt[i];
#pragma unroll
  for (int f = 0; f < F; ++f)
    if (tid == f) { po[f * (E + 2) + E] = M[f]; po[f * (E + 2) + E + 1] = L[f]; }
}

// partials of one (b, f) -> px[b, f, :] = sum_s P x_s, lse[b, f] = log sum_s exp(raw).  E / 4 column threads x 4 chunk groups.
template <int V>
__global__ void __launch_bounds__(128 * V)
fq_combine_kernel(const float* __restrict__ part, int nchunk, int F, float* __restrict__ px, float* __restrict__ lse) {
  constexpr int E = 128 * V, kCols = E / 4, kGroups = 4;
  __shared__ float s_l[kGroups];
  __shared__ float4 s_a[kGroups][kCols];
  const int f = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, col = tid % kCols, grp = tid / kCols;
  const float* pb = part + (size_t)b * nchunk * F * (E + 2) + (size_t)f * (E + 2);
  const size_t stride = (size_t)F * (E + 2);
  __shared__ float s_mx[128 * V / 32];
  float M = -INFINITY;
  for (int c = tid; c < nchunk; c += 128 * V) M = fmaxf(M, __ldg(pb + c * stride + E));      // one chunk per thread, then a block maximum
  M = warp_max(M);
  if ((tid & 31) == 0) s_mx[tid >> 5] = M;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 128 * V / 32; ++w) M = fmaxf(M, s_mx[w]);
  float L = 0.f;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int c = grp; c < nchunk; c += kGroups) {
    const float mc = __ldg(pb + c * stride + E);
    const float w = mc == -INFINITY ? 0.f : ex2f(mc - M);      // partial maxima are exp2-domain
    L = fmaf(__ldg(pb + c * stride + E + 1), w, L);
    const float* src = pb + c * stride + col * 4;      // (E + 2) floats per entry: rows are only 8-byte aligned
    const float2 s0 = __ldg(reinterpret_cast<const float2*>(src)), s1 = __ldg(reinterpret_cast<const float2*>(src) + 1);
    a.x = fmaf(w, s0.x, a.x); a.y = fmaf(w, s0.y, a.y); a.z = fmaf(w, s1.x, a.z); a.w = fmaf(w, s1.y, a.w);
  }
  s_a[grp][col] = a;
  if (col == 0) s_l[grp] = L;
  __syncthreads();
  if (grp == 0) {
    float Lt = 0.f;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
      Lt += s_l[g];
      const float4 v = s_a[g][col];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    const float inv = 1.f / Lt;
    reinterpret_cast<float4*>(px + ((size_t)b * F + f) * E)[col] = make_float4(t.x * inv, t.y * inv, t.z * inv, t.w * inv);
    if (col == 0) lse[b * F + f] = (M + log2f(Lt)) * kLn2;
  }
}

// =================================================================================================================
// few queries over many keys, backward.  P = exp(raw - lse), dP[f, s] = dpx[f] . x_s, ds = P (dP - D[f]) + draw,
//   dx_s = sum_f ds qt[f] + P dpx[f];   dqt[f] = sum_s ds x_s;   dc[f] = sum_s ds   (per-CTA partials [F][E + 1])
// =================================================================================================================
template <int F, int V, int kStages>
__global__ void __launch_bounds__(32 * kWarpsBwd, 1)
fq_bwd_kernel(const float* __restrict__ X, long long xs_b, long long xs_r, const float* __restrict__ qt, const float* __restrict__ raw,
              const float* __restrict__ lse, const float* __restrict__ dpx, const float* __restrict__ Dv, const float* __restrict__ draw,
              float* __restrict__ dX, float* __restrict__ part, int S, int rows_per_cta) {
  constexpr int kWarps = kWarpsBwd, kThreads = 32 * kWarps;
  using RingT = Ring<V, 1, kStages, kWarps>;
  constexpr int E = 128 * V, R = RingT::kR, NP = pow2_ceil(R * F);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  RingT ring;
  ring.data = reinterpret_cast<float*>(smem_raw);
  ring.bars = smem_u32(smem_raw) + sizeof(float) * kStages * RingT::kStageFloats;
  float* s_qt = reinterpret_cast<float*>(smem_raw + RingT::kBytes);      // [F][E]
  float* s_dp = s_qt + F * E;                                           // [F][E]; reused as the CTA accumulator of dqt
  float* s_dc = s_dp + F * E;                                           // [kWarps][F]
  const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  ring.init(tid);
  for (int i = tid; i < F * E / 4; i += kThreads) {
    reinterpret_cast<float4*>(s_qt)[i] = __ldg(reinterpret_cast<const float4*>(qt + (size_t)b * F * E) + i);
    reinterpret_cast<float4*>(s_dp)[i] = __ldg(reinterpret_cast<const float4*>(dpx + (size_t)b * F * E) + i);
  }
  __syncthreads();
  const float* xb = X + (size_t)b * xs_b;
  const int s_begin = chunk * rows_per_cta, s_end = min(S, s_begin + rows_per_cta);
  const int niter = cdiv(s_end - s_begin, kStageRows);
  float dc[F];
  Row<V> dq[F];
#pragma unroll
  for (int f = 0; f < F; ++f) { dc[f] = 0.f; zero_row(dq[f]); }
  {
    float ls[F], Df[F];
#pragma unroll
    for (int f = 0; f < F; ++f) { ls[f] = __ldg(lse + b * F + f) * kLog2e; Df[f] = __ldg(Dv + b * F + f); }
    ring.prologue(niter, xb, xs_r, nullptr, 0, s_begin, s_end, warp, lane);
    for (int it = 0; it < niter; ++it) {
      const int st = it % kStages, s0 = s_begin + it * kStageRows + warp * R;
      ring.top_up(it, niter, xb, xs_r, nullptr, 0, s_begin, s_end, warp, lane);
      // the raw scores (and draw) of the warp's rows: issued before the wait on the ring
      float p[R * F];
#pragma unroll
      for (int f = 0; f < F; ++f)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const bool ok = s0 + r < s_end;
          const size_t ri = ((size_t)b * F + f) * S + min(s0 + r, S - 1);
          p[r * F + f] = ok ? ex2f(fmaf(__ldg(raw + ri), kLog2e, -ls[f])) : 0.f;
        }
      ring.wait_full(it);
      if (s0 < s_end) {
        float d[NP];
#pragma unroll
        for (int i = R * F; i < NP; ++i) d[i] = 0.f;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const Row<V> g = lds_row<V>(s_dp + f * E, lane);
#pragma unroll
          for (int r = 0; r < R; ++r) d[r * F + f] = dot_part(lds_row_if<V>(ring.row(st, 0, warp * R + r), lane, s0 + r < s_end), g);
        }
        {
          float own;
          int oi;
          allreduce<NP>(d, lane, own, oi);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const bool ok = s0 + r < s_end;
          asm volatile("" ::: "memory");      // keep the short-side rows in shared memory, not hoisted into registers
          const Row<V> x = lds_row_if<V>(ring.row(st, 0, warp * R + r), lane, ok);
          Row<V> o;
          zero_row(o);
#pragma unroll
          for (int f = 0; f < F; ++f) {
            float ds = p[r * F + f] * (d[r * F + f] - Df[f]);               // masked row: p = 0
            if (draw != nullptr && ok) ds += __ldg(draw + ((size_t)b * F + f) * S + s0 + r);
            dc[f] += ds;
            axpy(dq[f], ds, x);
            axpy(o, ds, lds_row<V>(s_qt + f * E, lane));
            axpy(o, p[r * F + f], lds_row<V>(s_dp + f * E, lane));
          }
          if (ok) store_row<V>(dX + ((size_t)b * S + s0 + r) * E, lane, o);
        }
      }
      ring.release(it, lane);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int f = 0; f < F; ++f) s_dc[warp * F + f] = dc[f];
  }
  __syncthreads();      // all warps are done with s_dp
  for (int w = 0; w < kWarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int f = 0; f < F; ++f) {
        Row<V> t;
        if (w == 0) zero_row(t); else t = lds_row<V>(s_dp + f * E, lane);
        axpy(t, 1.f, dq[f]);
        store_row<V>(s_dp + f * E, lane, t);
      }
    }
    __syncthreads();
  }
  float* po = part + ((size_t)b * nchunk + chunk) * F * (E + 1);
  for (int i = tid; i < F * E; i += kThreads) po[(i / E) * (E + 1) + (i % E)] = s_dp[i];
  if (tid < F) {
    float t = 0.f;
    for (int w = 0; w < kWarps; ++w) t += s_dc[w * F + tid];
    po[tid * (E + 1) + E] = t;
  }
}

// =================================================================================================================
// many queries over few keys, forward:  raw[b, s, f] = kt[b, f] . x[b, s] + c[b, f];  P = softmax_f;  out = sum_f P vt[f] + bo
// =================================================================================================================
template <int F, int V, int kStages>
__global__ void __launch_bounds__(32 * kWarpsFwd, 1)
fk_fwd_kernel(const float* __restrict__ X, long long xs_b, long long xs_r, const float* __restrict__ kt, const float* __restrict__ cv,
              const float* __restrict__ vt, const float* __restrict__ bo, float* __restrict__ raw, float* __restrict__ out, int S,
              int rows_per_cta) {
  constexpr int kWarps = kWarpsFwd, kThreads = 32 * kWarps;
  using RingT = Ring<V, 1, kStages, kWarps>;
  constexpr int E = 128 * V, R = RingT::kR, NP = pow2_ceil(R * F);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  RingT ring;
  ring.data = reinterpret_cast<float*>(smem_raw);
  ring.bars = smem_u32(smem_raw) + sizeof(float) * kStages * RingT::kStageFloats;
  float* s_kt = reinterpret_cast<float*>(smem_raw + RingT::kBytes);      // [F][E]
  float* s_vt = s_kt + F * E;                                           // [F][E]
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  ring.init(tid);
  for (int i = tid; i < F * E / 4; i += kThreads) {      // scores are kept in the exp2 domain: kt and c times log2(e)
    const float4 t = __ldg(reinterpret_cast<const float4*>(kt + (size_t)b * F * E) + i);
    reinterpret_cast<float4*>(s_kt)[i] = make_float4(t.x * kLog2e, t.y * kLog2e, t.z * kLog2e, t.w * kLog2e);
    reinterpret_cast<float4*>(s_vt)[i] = __ldg(reinterpret_cast<const float4*>(vt + (size_t)b * F * E) + i);
  }
  __syncthreads();
  const float* xb = X + (size_t)b * xs_b;
  const int s_begin = chunk * rows_per_cta, s_end = min(S, s_begin + rows_per_cta);
  const int niter = cdiv(s_end - s_begin, kStageRows);
  float cf[F];
#pragma unroll
  for (int f = 0; f < F; ++f) cf[f] = __ldg(cv + b * F + f) * kLog2e;
  const Row<V> bias = ldg_row<V>(bo, lane);
  ring.prologue(niter, xb, xs_r, nullptr, 0, s_begin, s_end, warp, lane);
  for (int it = 0; it < niter; ++it) {
    const int st = it % kStages, s0 = s_begin + it * kStageRows + warp * R;
    ring.top_up(it, niter, xb, xs_r, nullptr, 0, s_begin, s_end, warp, lane);
    ring.wait_full(it);
    float d[NP];
#pragma unroll
    for (int i = R * F; i < NP; ++i) d[i] = 0.f;
    if (s0 < s_end) {
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const Row<V> k = lds_row<V>(s_kt + f * E, lane);
#pragma unroll
        for (int r = 0; r < R; ++r) d[r * F + f] = dot_part(lds_row_if<V>(ring.row(st, 0, warp * R + r), lane, s0 + r < s_end), k);
      }
    }
    ring.release(it, lane);      // the rows are consumed: the rest of the step works from registers
    if (s0 < s_end) {
      float own;
      int oi;
      allreduce<NP>(d, lane, own, oi);
      if (oi >= 0 && oi < R * F && s0 + oi / F < s_end)      // raw [b, s, f]: R * F consecutive floats, each from the lane that holds it
        raw[((size_t)b * S + s0) * F + oi] = own * kLn2 + __ldg(cv + b * F + oi % F);
#pragma unroll
      for (int i = 0; i < R * F; ++i) d[i] += cf[i % F];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float mx = d[r * F];
#pragma unroll
        for (int f = 1; f < F; ++f) mx = fmaxf(mx, d[r * F + f]);
        float sum = 0.f, p[F];
#pragma unroll
        for (int f = 0; f < F; ++f) { p[f] = ex2f(d[r * F + f] - mx); sum += p[f]; }
        const float inv = 1.f / sum;
        Row<V> o = bias;
#pragma unroll
        for (int f = 0; f < F; ++f) axpy(o, p[f] * inv, lds_row<V>(s_vt + f * E, lane));
        if (s0 + r < s_end) store_row<V>(out + ((size_t)b * S + s0 + r) * E, lane, o);
      }
    }
  }
}

// =================================================================================================================
// many queries over few keys, backward.  P = softmax_f(raw[s, :]), dP[s, f] = dout_s . vt[f], ds = P (dP - sum_f P dP) + draw,
//   dx_s = sum_f ds kt[f];   dkt[f] = sum_s ds x_s;   dvt[f] = sum_s P dout_s;   dc[f] = sum_s ds;   dbo = sum_s dout_s
//   per-CTA partials [2 F + 1][E] then [F] (dc)
// =================================================================================================================
template <int F, int V, int kStages>
__global__ void __launch_bounds__(32 * kWarpsBwd, 1)
fk_bwd_kernel(const float* __restrict__ X, long long xs_b, long long xs_r, const float* __restrict__ G, long long gs_b, long long gs_r,
              const float* __restrict__ kt, const float* __restrict__ vt, const float* __restrict__ raw, const float* __restrict__ draw,
              float* __restrict__ dX, float* __restrict__ part, int S, int rows_per_cta) {
  constexpr int kWarps = kWarpsBwd, kThreads = 32 * kWarps;
  using RingT = Ring<V, 2, kStages, kWarps>;
  constexpr int E = 128 * V, R = RingT::kR, NP = pow2_ceil(R * F);
  constexpr int kAcc = 2 * F + 1;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  RingT ring;
  ring.data = reinterpret_cast<float*>(smem_raw);
  ring.bars = smem_u32(smem_raw) + sizeof(float) * kStages * RingT::kStageFloats;
  float* s_kt = reinterpret_cast<float*>(smem_raw + RingT::kBytes);      // [F][E]; with s_vt reused as the CTA accumulators
  float* s_vt = s_kt + F * E;                                           // [F][E]
  float* s_db = s_vt + F * E;                                           // [E]
  float* s_dc = s_db + E;                                               // [kWarps][F]
  const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  ring.init(tid);
  for (int i = tid; i < F * E / 4; i += kThreads) {
    reinterpret_cast<float4*>(s_kt)[i] = __ldg(reinterpret_cast<const float4*>(kt + (size_t)b * F * E) + i);
    reinterpret_cast<float4*>(s_vt)[i] = __ldg(reinterpret_cast<const float4*>(vt + (size_t)b * F * E) + i);
  }
  __syncthreads();
  const float* xb = X + (size_t)b * xs_b;
  const float* gb = G + (size_t)b * gs_b;
  const int s_begin = chunk * rows_per_cta, s_end = min(S, s_begin + rows_per_cta);
  const int niter = cdiv(s_end - s_begin, kStageRows);
  Row<V> dk[F], dv[F], db;
  float dc[F];
#pragma unroll
  for (int f = 0; f < F; ++f) { zero_row(dk[f]); zero_row(dv[f]); dc[f] = 0.f; }
  zero_row(db);
  {
    ring.prologue(niter, xb, xs_r, gb, gs_r, s_begin, s_end, warp, lane);
    for (int it = 0; it < niter; ++it) {
      const int st = it % kStages, s0 = s_begin + it * kStageRows + warp * R;
      ring.top_up(it, niter, xb, xs_r, gb, gs_r, s_begin, s_end, warp, lane);
      float p[R * F];
#pragma unroll
      for (int r = 0; r < R; ++r) {      // softmax of the saved raw scores: issued before the wait on the ring
        const bool ok = s0 + r < s_end;
        const size_t ri = ((size_t)b * S + min(s0 + r, S - 1)) * F;
        float mx = -INFINITY;
#pragma unroll
        for (int f = 0; f < F; ++f) { p[r * F + f] = __ldg(raw + ri + f); mx = fmaxf(mx, p[r * F + f]); }
        float sum = 0.f;
        mx *= kLog2e;
#pragma unroll
        for (int f = 0; f < F; ++f) { p[r * F + f] = ex2f(fmaf(p[r * F + f], kLog2e, -mx)); sum += p[r * F + f]; }
        const float inv = ok ? 1.f / sum : 0.f;
#pragma unroll
        for (int f = 0; f < F; ++f) p[r * F + f] *= inv;
      }
      ring.wait_full(it);
      if (s0 < s_end) {
        float d[NP];
#pragma unroll
        for (int i = R * F; i < NP; ++i) d[i] = 0.f;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const Row<V> v = lds_row<V>(s_vt + f * E, lane);
#pragma unroll
          for (int r = 0; r < R; ++r) d[r * F + f] = dot_part(lds_row_if<V>(ring.row(st, 1, warp * R + r), lane, s0 + r < s_end), v);
        }
        {
          float own;
          int oi;
          allreduce<NP>(d, lane, own, oi);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const bool ok = s0 + r < s_end;
          asm volatile("" ::: "memory");      // keep the short-side rows in shared memory, not hoisted into registers
          const Row<V> x = lds_row_if<V>(ring.row(st, 0, warp * R + r), lane, ok);
          const Row<V> g = lds_row_if<V>(ring.row(st, 1, warp * R + r), lane, ok);
          float dsum = 0.f;
#pragma unroll
          for (int f = 0; f < F; ++f) dsum = fmaf(p[r * F + f], d[r * F + f], dsum);
          Row<V> o;
          zero_row(o);
#pragma unroll
          for (int f = 0; f < F; ++f) {
            float ds = p[r * F + f] * (d[r * F + f] - dsum);      // masked row: p = 0
            if (draw != nullptr && ok) ds += __ldg(draw + ((size_t)b * S + s0 + r) * F + f);
            dc[f] += ds;
            axpy(dk[f], ds, x);
            axpy(dv[f], p[r * F + f], g);
            axpy(o, ds, lds_row<V>(s_kt + f * E, lane));
          }
          axpy(db, 1.f, g);
          if (ok) store_row<V>(dX + ((size_t)b * S + s0 + r) * E, lane, o);
        }
      }
      ring.release(it, lane);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int f = 0; f < F; ++f) s_dc[warp * F + f] = dc[f];
  }
  __syncthreads();      // all warps are done with s_kt / s_vt: they become the accumulators of d kt / d vt
  for (int w = 0; w < kWarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int f = 0; f < F; ++f) {
        Row<V> t;
        if (w == 0) zero_row(t); else t = lds_row<V>(s_kt + f * E, lane);
        axpy(t, 1.f, dk[f]);
        store_row<V>(s_kt + f * E, lane, t);
        if (w == 0) zero_row(t); else t = lds_row<V>(s_vt + f * E, lane);
        axpy(t, 1.f, dv[f]);
        store_row<V>(s_vt + f * E, lane, t);
      }
      Row<V> t;
      if (w == 0) zero_row(t); else t = lds_row<V>(s_db, lane);
      axpy(t, 1.f, db);
      store_row<V>(s_db, lane, t);
    }
    __syncthreads();
  }
  float* po = part + ((size_t)b * nchunk + chunk) * (kAcc * E + F);
  for (int i = tid; i < kAcc * E; i += kThreads) po[i] = s_kt[i];      // s_kt, s_vt, s_db are contiguous
  if (tid < F) {
    float t = 0.f;
    for (int w = 0; w < kWarps; ++w) t += s_dc[w * F + tid];
    po[kAcc * E + tid] = t;
  }
}

template <typename Kern>
static int set_smem(Kern k, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;
  return (int)cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

#define DML_CA_DISPATCH_F(F_, ...) \
  switch (F_) {                     \
    case 1: { constexpr int kF = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int kF = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int kF = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int kF = 4; __VA_ARGS__; } break; \
    case 5: { constexpr int kF = 5; __VA_ARGS__; } break; \
    case 6: { constexpr int kF = 6; __VA_ARGS__; } break; \
    case 7: { constexpr int kF = 7; __VA_ARGS__; } break; \
    case 8: { constexpr int kF = 8; __VA_ARGS__; } break; \
    default: return DML_EUNSUPPORTED;               \
  }

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kStages1 = 5, kStages2 = 3;      // 160 KB / 192 KB of rows in flight per SM

// rows per CTA (a multiple of the 32-row stage): about one CTA per SM over all bags
static int rows_per_cta(int B, int S, int nsm) {
  if (nsm <= 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0) nsm = 148;
  }
  const int per_bag = nsm / B > 0 ? nsm / B : 1;
  const int rows = cdiv(cdiv(S, per_bag), kStageRows) * kStageRows;
  return rows;
}

}  // namespace ca
}  // namespace dml

extern "C" {

int dml_coattn_chunks(int B, int S, int nsm) {
  if (B <= 0 || S <= 0) return 0;
  return dml::cdiv(S, dml::ca::rows_per_cta(B, S, nsm));
}

size_t dml_coattn_fq_fwd_ws_floats(int B, int F, int S, int E) {
  if (B <= 0 || F <= 0 || S <= 0 || E <= 0) return 0;
  return (size_t)B * dml_coattn_chunks(B, S, 0) * F * (E + 2);
}
size_t dml_coattn_fq_bwd_ws_floats(int B, int F, int S, int E) {
  if (B <= 0 || F <= 0 || S <= 0 || E <= 0) return 0;
  return (size_t)B * dml_coattn_chunks(B, S, 0) * F * (E + 1);
}
size_t dml_coattn_fk_bwd_ws_floats(int B, int F, int S, int E) {
  if (B <= 0 || F <= 0 || S <= 0 || E <= 0) return 0;
  return (size_t)B * dml_coattn_chunks(B, S, 0) * ((2 * F + 1) * E + F);
}

int dml_coattn_fq_fwd(const float* x, long long xs_b, long long xs_r, const float* qt, const float* c, int B, int F, int S, int E,
                      float* raw, float* px, float* lse, float* ws, void* stream) {
  using namespace dml;
  using namespace dml::ca;
  DML_CHECK_ARG(x && qt && c && raw && px && lse && ws);
  DML_CHECK_ARG(B > 0 && S > 0 && F > 0 && xs_r >= E && (xs_r % 4) == 0 && (xs_b % 4) == 0 && aligned16(x) && aligned16(qt));
  if (E != 256 || F > 8) return DML_EUNSUPPORTED;
  constexpr int V = 2;
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = rows_per_cta(B, S, 0), nchunk = cdiv(S, rows);
  dim3 grid(nchunk, B);
  DML_CA_DISPATCH_F(F, {
    const size_t smem = Ring<V, 1, kStages1, kWarpsFwd>::kBytes + sizeof(float) * (kF * E + kWarpsFwd * 2 * kF);
    int e = set_smem(fq_fwd_kernel<kF, V, kStages1>, smem);
    if (e) return e;
    fq_fwd_kernel<kF, V, kStages1><<<grid, 32 * kWarpsFwd, smem, st>>>(x, xs_b, xs_r, qt, c, raw, ws, S, rows);
  });
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  fq_combine_kernel<V><<<dim3(F, B), 128 * V, 0, st>>>(ws, nchunk, F, px, lse);
  DML_RETURN_LAUNCH();
}

int dml_coattn_fq_bwd(const float* x, long long xs_b, long long xs_r, const float* qt, const float* raw, const float* lse,
                      const float* dpx, const float* dsum, const float* draw, int B, int F, int S, int E, float* dx, float* ws,
                      void* stream) {
  using namespace dml;
  using namespace dml::ca;
  DML_CHECK_ARG(x && qt && raw && lse && dpx && dsum && dx && ws);
  DML_CHECK_ARG(B > 0 && S > 0 && F > 0 && xs_r >= E && (xs_r % 4) == 0 && (xs_b % 4) == 0 && aligned16(x) && aligned16(qt) &&
                aligned16(dpx) && aligned16(dx));
  if (E != 256 || F > 8) return DML_EUNSUPPORTED;
  constexpr int V = 2;
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = rows_per_cta(B, S, 0);
  dim3 grid(cdiv(S, rows), B);
  DML_CA_DISPATCH_F(F, {
    const size_t smem = Ring<V, 1, kStages1, kWarpsBwd>::kBytes + sizeof(float) * (2 * kF * E + kWarpsBwd * kF);
    int e = set_smem(fq_bwd_kernel<kF, V, kStages1>, smem);
    if (e) return e;
    fq_bwd_kernel<kF, V, kStages1><<<grid, 32 * kWarpsBwd, smem, st>>>(x, xs_b, xs_r, qt, raw, lse, dpx, dsum, draw, dx, ws, S, rows);
  });
  DML_RETURN_LAUNCH();
}

int dml_coattn_fk_fwd(const float* x, long long xs_b, long long xs_r, const float* kt, const float* c, const float* vt, const float* bo,
                      int B, int F, int S, int E, float* raw, float* out, void* stream) {
  using namespace dml;
  using namespace dml::ca;
  DML_CHECK_ARG(x && kt && c && vt && bo && raw && out);
  DML_CHECK_ARG(B > 0 && S > 0 && F > 0 && xs_r >= E && (xs_r % 4) == 0 && (xs_b % 4) == 0 && aligned16(x) && aligned16(kt) &&
                aligned16(vt) && aligned16(bo) && aligned16(out));
  if (E != 256 || F > 8) return DML_EUNSUPPORTED;
  constexpr int V = 2;
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = rows_per_cta(B, S, 0);
  dim3 grid(cdiv(S, rows), B);
  DML_CA_DISPATCH_F(F, {
    const size_t smem = Ring<V, 1, kStages1, kWarpsFwd>::kBytes + sizeof(float) * (2 * kF * E);
    int e = set_smem(fk_fwd_kernel<kF, V, kStages1>, smem);
    if (e) return e;
    fk_fwd_kernel<kF, V, kStages1><<<grid, 32 * kWarpsFwd, smem, st>>>(x, xs_b, xs_r, kt, c, vt, bo, raw, out, S, rows);
  });
  DML_RETURN_LAUNCH();
}

int dml_coattn_fk_bwd(const float* x, long long xs_b, long long xs_r, const float* dout, long long gs_b, long long gs_r, const float* kt,
                      const float* vt, const float* raw, const float* draw, int B, int F, int S, int E, float* dx, float* ws,
                      void* stream) {
  using namespace dml;
  using namespace dml::ca;
  DML_CHECK_ARG(x && dout && kt && vt && raw && dx && ws);
  DML_CHECK_ARG(B > 0 && S > 0 && F > 0 && xs_r >= E && (xs_r % 4) == 0 && (xs_b % 4) == 0 && gs_r >= E && (gs_r % 4) == 0 &&
                (gs_b % 4) == 0 && aligned16(x) && aligned16(dout) && aligned16(kt) && aligned16(vt) && aligned16(dx));
  if (E != 256 || F > 8) return DML_EUNSUPPORTED;
  constexpr int V = 2;
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = rows_per_cta(B, S, 0);
  dim3 grid(cdiv(S, rows), B);
  DML_CA_DISPATCH_F(F, {
    const size_t smem = Ring<V, 2, kStages2, kWarpsBwd>::kBytes + sizeof(float) * ((2 * kF + 1) * E + kWarpsBwd * kF);
    int e = set_smem(fk_bwd_kernel<kF, V, kStages2>, smem);
    if (e) return e;
    fk_bwd_kernel<kF, V, kStages2><<<grid, 32 * kWarpsBwd, smem, st>>>(x, xs_b, xs_r, dout, gs_b, gs_r, kt, vt, raw, draw, dx, ws, S, rows);
  });
  DML_RETURN_LAUNCH();
}

}  // extern "C"
