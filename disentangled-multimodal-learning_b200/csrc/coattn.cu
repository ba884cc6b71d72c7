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
// One warp owns a row at a time (lane = 4 consecutive floats of each 128-float span, float4 loads), R rows per step for
// memory-level parallelism; the F dot products of a row are reduced with warp shuffles; the few-side vectors live in
// shared memory.  Per-CTA partial sums (online-softmax state, short-side gradients) go to a caller-owned workspace and are
// reduced in a second, tiny launch or by the caller: no atomics, results are deterministic.
#include "common.cuh"

namespace dml {
namespace ca {

constexpr int kWarps = 8;
constexpr int kThreads = 32 * kWarps;
constexpr int kRowsPerCta = 128;      // 16 rows per warp

template <int V>
struct Row {
  float4 v[V];
};

template <int V>
__device__ __forceinline__ Row<V> load_row(const float* p, int lane, bool valid) {
  Row<V> r;
#pragma unroll
  for (int c = 0; c < V; ++c)
    r.v[c] = valid ? __ldg(reinterpret_cast<const float4*>(p) + c * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
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
__device__ __forceinline__ void store_row(float* p, int lane, const Row<V>& r) {
#pragma unroll
  for (int c = 0; c < V; ++c) reinterpret_cast<float4*>(p)[c * 32 + lane] = r.v[c];
}
template <int V>
__device__ __forceinline__ float dot_part(const Row<V>& a, const Row<V>& b) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < V; ++c) {
    s = fmaf(a.v[c].x, b.v[c].x, s); s = fmaf(a.v[c].y, b.v[c].y, s);
    s = fmaf(a.v[c].z, b.v[c].z, s); s = fmaf(a.v[c].w, b.v[c].w, s);
  }
  return s;
}
template <int V>
__device__ __forceinline__ void axpy(Row<V>& y, float a, const Row<V>& x) {
#pragma unroll
  for (int c = 0; c < V; ++c) {
    y.v[c].x = fmaf(a, x.v[c].x, y.v[c].x); y.v[c].y = fmaf(a, x.v[c].y, y.v[c].y);
    y.v[c].z = fmaf(a, x.v[c].z, y.v[c].z); y.v[c].w = fmaf(a, x.v[c].w, y.v[c].w);
  }
}
template <int V>
__device__ __forceinline__ void scale_row(Row<V>& y, float a) {
#pragma unroll
  for (int c = 0; c < V; ++c) { y.v[c].x *= a; y.v[c].y *= a; y.v[c].z *= a; y.v[c].w *= a; }
}
template <int V>
__device__ __forceinline__ void zero_row(Row<V>& y) {
#pragma unroll
  for (int c = 0; c < V; ++c) y.v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// All-reduce of N per-lane partial sums (every lane ends with every total).  N a power of two up to 16 uses a
// reduce-scatter (each exchange halves the number of values a lane carries) followed by broadcasts: N + log2 N shuffles
// for the sums instead of 5 N.
template <int N>
__device__ __forceinline__ void allreduce(float (&v)[N], int lane) {
  if constexpr (N == 16 || N == 8 || N == 4) {
    constexpr int kSteps = N == 16 ? 4 : N == 8 ? 3 : 2;
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
    // w[0] of a lane = the total of value index(lane) over the lanes that share its top kSteps lane bits ... finish over
    // the remaining low bits, then hand every total to every lane
    float t = w[0];
#pragma unroll
    for (int o = (16 >> kSteps); o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      // value i sits on the lanes whose top bits spell i (bit 16 = most significant)
      int src = 0;
#pragma unroll
      for (int st = 0; st < kSteps; ++st)
        if (i & (N >> (st + 1))) src |= 16 >> st;
      v[i] = __shfl_sync(0xffffffffu, t, src);
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
  }
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Rows per warp step: enough loads in flight, register budget permitting.
template <int F> struct FwdRows { static constexpr int R = F <= 4 ? 4 : (F <= 8 ? 2 : 1); };
template <int F> struct BwdRows { static constexpr int R = F <= 2 ? 4 : (F <= 4 ? 2 : 1); };

// =================================================================================================================
// few queries over many keys, forward:   raw[b, f, s] = qt[b, f] . x[b, s] + c[b, f];   online softmax over s;
// partial (M, L, acc = sum_s exp(raw - M) x_s) per CTA -> part[b][chunk][f][E + 2] = {acc[E], M, L}
// =================================================================================================================
template <int F, int V>
__global__ void __launch_bounds__(kThreads)
fq_fwd_kernel(const float* __restrict__ X, long long xs_b, long long xs_r, const float* __restrict__ qt, const float* __restrict__ cv,
              float* __restrict__ raw, float* __restrict__ part, int S) {
  constexpr int E = 128 * V, R = FwdRows<F>::R;
  extern __shared__ __align__(16) float sm[];
  float* s_qt = sm;                 // [F][E]; reused as the CTA accumulator at the end
  float* s_ml = sm + F * E;         // [kWarps][2 F]
  const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < F * E / 4; i += kThreads)
    reinterpret_cast<float4*>(s_qt)[i] = __ldg(reinterpret_cast<const float4*>(qt + (size_t)b * F * E) + i);
  __syncthreads();
  float cf[F];
#pragma unroll
  for (int f = 0; f < F; ++f) cf[f] = __ldg(cv + b * F + f);
  const float* xb = X + (size_t)b * xs_b;
  float m[F], l[F];
  Row<V> acc[F];
#pragma unroll
  for (int f = 0; f < F; ++f) { m[f] = -INFINITY; l[f] = 0.f; zero_row(acc[f]); }

  const int s_begin = chunk * kRowsPerCta, s_end = min(S, s_begin + kRowsPerCta);
  for (int s0 = s_begin + warp * R; s0 < s_end; s0 += kWarps * R) {
    Row<V> x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = load_row<V>(xb + (size_t)(s0 + r) * xs_r, lane, s0 + r < s_end);
    float d[R * F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const Row<V> q = lds_row<V>(s_qt + f * E, lane);
#pragma unroll
      for (int r = 0; r < R; ++r) d[r * F + f] = dot_part(x[r], q);
    }
    allreduce<R * F>(d, lane);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int f = 0; f < F; ++f) d[r * F + f] = (s0 + r < s_end) ? d[r * F + f] + cf[f] : -INFINITY;
    // raw scores: lane (f, r) writes one (R consecutive s per f)
    if (lane < R * F) {
      const int f = lane / R, r = lane % R;
      float val = 0.f;
#pragma unroll
      for (int i = 0; i < R * F; ++i)
        if (i == r * F + f) val = d[i];
      if (s0 + r < s_end) raw[((size_t)b * F + f) * S + s0 + r] = val;
    }
#pragma unroll
    for (int f = 0; f < F; ++f) {
      float mx = d[f];
#pragma unroll
      for (int r = 1; r < R; ++r) mx = fmaxf(mx, d[r * F + f]);
      if (mx > m[f]) {               // warp-uniform: every lane holds the same totals
        const float sc = ex2f((m[f] - mx) * kLog2e);      // m = -inf -> 0
        scale_row(acc[f], sc);
        l[f] *= sc;
        m[f] = mx;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float p = ex2f((d[r * F + f] - m[f]) * kLog2e);      // masked row: exp(-inf) = 0
        l[f] += p;
        axpy(acc[f], p, x[r]);
      }
    }
  }
  // CTA combine: common maximum, then the warps add their rescaled sums into shared memory one after the other
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
      ll += mw == -INFINITY ? 0.f : s_ml[w * 2 * F + F + f] * ex2f((mw - mm) * kLog2e);
    }
    M[f] = mm; L[f] = ll;
    myscale[f] = m[f] == -INFINITY ? 0.f : ex2f((m[f] - mm) * kLog2e);
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

// partials of one (b, f) -> px[b, f, :] = sum_s P x_s, lse[b, f] = log sum_s exp(raw)
template <int V>
__global__ void __launch_bounds__(128 * V / 4)
fq_combine_kernel(const float* __restrict__ part, int nchunk, int F, float* __restrict__ px, float* __restrict__ lse) {
  constexpr int E = 128 * V;
  const int f = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;      // E / 4 threads: one float4 each
  const float* pb = part + (size_t)b * nchunk * F * (E + 2) + (size_t)f * (E + 2);
  const size_t stride = (size_t)F * (E + 2);
  float M = -INFINITY;
  for (int c = 0; c < nchunk; ++c) M = fmaxf(M, pb[c * stride + E]);
  float L = 0.f;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < nchunk; ++c) {
    const float mc = pb[c * stride + E];
    if (mc == -INFINITY) continue;
    const float w = ex2f((mc - M) * kLog2e);
    L += pb[c * stride + E + 1] * w;
    const float* src = pb + c * stride + tid * 4;      // (E + 2) floats per entry: rows are only 8-byte aligned
    a.x = fmaf(w, src[0], a.x); a.y = fmaf(w, src[1], a.y); a.z = fmaf(w, src[2], a.z); a.w = fmaf(w, src[3], a.w);
  }
  const float inv = 1.f / L;
  reinterpret_cast<float4*>(px + ((size_t)b * F + f) * E)[tid] = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
  if (tid == 0) lse[b * F + f] = M + logf(L);
}

// =================================================================================================================
// few queries over many keys, backward.  P = exp(raw - lse), dP[f, s] = dpx[f] . x_s, ds = P (dP - D[f]) + draw,
//   dx_s = sum_f ds qt[f] + P dpx[f];   dqt[f] = sum_s ds x_s;   dc[f] = sum_s ds   (per-CTA partials [F][E + 1])
// =================================================================================================================
template <int F, int V>
__global__ void __launch_bounds__(kThreads)
fq_bwd_kernel(const float* __restrict__ X, long long xs_b, long long xs_r, const float* __restrict__ qt, const float* __restrict__ raw,
              const float* __restrict__ lse, const float* __restrict__ dpx, const float* __restrict__ Dv, const float* __restrict__ draw,
              float* __restrict__ dX, float* __restrict__ part, int S) {
  constexpr int E = 128 * V, R = BwdRows<F>::R;
  extern __shared__ __align__(16) float sm[];
  float* s_qt = sm;                 // [F][E]
  float* s_dp = sm + F * E;         // [F][E]; reused as the CTA accumulator of dqt
  float* s_dc = sm + 2 * F * E;     // [kWarps][F]
  const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < F * E / 4; i += kThreads) {
    reinterpret_cast<float4*>(s_qt)[i] = __ldg(reinterpret_cast<const float4*>(qt + (size_t)b * F * E) + i);
    reinterpret_cast<float4*>(s_dp)[i] = __ldg(reinterpret_cast<const float4*>(dpx + (size_t)b * F * E) + i);
  }
  __syncthreads();
  float ls[F], Df[F], dc[F];
  Row<V> dq[F];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    ls[f] = __ldg(lse + b * F + f); Df[f] = __ldg(Dv + b * F + f); dc[f] = 0.f; zero_row(dq[f]);
  }
  const float* xb = X + (size_t)b * xs_b;
  const int s_begin = chunk * kRowsPerCta, s_end = min(S, s_begin + kRowsPerCta);
  for (int s0 = s_begin + warp * R; s0 < s_end; s0 += kWarps * R) {
    Row<V> x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = load_row<V>(xb + (size_t)(s0 + r) * xs_r, lane, s0 + r < s_end);
    float d[R * F], p[R * F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const Row<V> g = lds_row<V>(s_dp + f * E, lane);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        d[r * F + f] = dot_part(x[r], g);
        const bool ok = s0 + r < s_end;
        const size_t ri = ((size_t)b * F + f) * S + min(s0 + r, S - 1);
        p[r * F + f] = ok ? ex2f((__ldg(raw + ri) - ls[f]) * kLog2e) : 0.f;
      }
    }
    allreduce<R * F>(d, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool ok = s0 + r < s_end;
      Row<V> o;
      zero_row(o);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        float ds = p[r * F + f] * (d[r * F + f] - Df[f]);
        if (draw != nullptr && ok) ds += __ldg(draw + ((size_t)b * F + f) * S + s0 + r);
        if (!ok) ds = 0.f;
        dc[f] += ds;
        axpy(dq[f], ds, x[r]);
        axpy(o, ds, lds_row<V>(s_qt + f * E, lane));
        axpy(o, p[r * F + f], lds_row<V>(s_dp + f * E, lane));
      }
      if (ok) store_row<V>(dX + ((size_t)b * S + s0 + r) * E, lane, o);
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
template <int F, int V>
__global__ void __launch_bounds__(kThreads)
fk_fwd_kernel(const float* __restrict__ X, long long xs_b, long long xs_r, const float* __restrict__ kt, const float* __restrict__ cv,
              const float* __restrict__ vt, const float* __restrict__ bo, float* __restrict__ raw, float* __restrict__ out, int S) {
  constexpr int E = 128 * V, R = FwdRows<F>::R;
  extern __shared__ __align__(16) float sm[];
  float* s_kt = sm;                 // [F][E]
  float* s_vt = sm + F * E;         // [F][E]
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < F * E / 4; i += kThreads) {
    reinterpret_cast<float4*>(s_kt)[i] = __ldg(reinterpret_cast<const float4*>(kt + (size_t)b * F * E) + i);
    reinterpret_cast<float4*>(s_vt)[i] = __ldg(reinterpret_cast<const float4*>(vt + (size_t)b * F * E) + i);
  }
  __syncthreads();
  float cf[F];
#pragma unroll
  for (int f = 0; f < F; ++f) cf[f] = __ldg(cv + b * F + f);
  const Row<V> bias = load_row<V>(bo, lane, true);
  const float* xb = X + (size_t)b * xs_b;
  const int s_begin = chunk * kRowsPerCta, s_end = min(S, s_begin + kRowsPerCta);
  for (int s0 = s_begin + warp * R; s0 < s_end; s0 += kWarps * R) {
    Row<V> x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = load_row<V>(xb + (size_t)(s0 + r) * xs_r, lane, s0 + r < s_end);
    float d[R * F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const Row<V> k = lds_row<V>(s_kt + f * E, lane);
#pragma unroll
      for (int r = 0; r < R; ++r) d[r * F + f] = dot_part(x[r], k);
    }
    allreduce<R * F>(d, lane);
#pragma unroll
    for (int i = 0; i < R * F; ++i) d[i] += cf[i % F];
    if (lane < R * F) {              // raw [b, s, f]: R * F consecutive floats
      float val = 0.f;
#pragma unroll
      for (int i = 0; i < R * F; ++i)
        if (i == lane) val = d[i];
      if (s0 + lane / F < s_end) raw[((size_t)b * S + s0) * F + lane] = val;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float mx = d[r * F];
#pragma unroll
      for (int f = 1; f < F; ++f) mx = fmaxf(mx, d[r * F + f]);
      float sum = 0.f, p[F];
#pragma unroll
      for (int f = 0; f < F; ++f) { p[f] = ex2f((d[r * F + f] - mx) * kLog2e); sum += p[f]; }
      const float inv = 1.f / sum;
      Row<V> o = bias;
#pragma unroll
      for (int f = 0; f < F; ++f) axpy(o, p[f] * inv, lds_row<V>(s_vt + f * E, lane));
      if (s0 + r < s_end) store_row<V>(out + ((size_t)b * S + s0 + r) * E, lane, o);
    }
  }
}

// =================================================================================================================
// many queries over few keys, backward.  P = softmax_f(raw[s, :]), dP[s, f] = dout_s . vt[f], ds = P (dP - sum_f P dP) + draw,
//   dx_s = sum_f ds kt[f];   dkt[f] = sum_s ds x_s;   dvt[f] = sum_s P dout_s;   dc[f] = sum_s ds;   dbo = sum_s dout_s
//   per-CTA partials [2 F + 1][E] then [F] (dc)
// =================================================================================================================
template <int F, int V>
__global__ void __launch_bounds__(kThreads)
fk_bwd_kernel(const float* __restrict__ X, long long xs_b, long long xs_r, const float* __restrict__ G, long long gs_b, long long gs_r,
              const float* __restrict__ kt, const float* __restrict__ vt, const float* __restrict__ raw, const float* __restrict__ draw,
              float* __restrict__ dX, float* __restrict__ part, int S) {
  constexpr int E = 128 * V, R = BwdRows<F>::R;
  constexpr int kAcc = 2 * F + 1;
  extern __shared__ __align__(16) float sm[];
  float* s_kt = sm;                         // [F][E]
  float* s_vt = sm + F * E;                 // [F][E]
  float* s_acc = sm + 2 * F * E;            // [2 F + 1][E]
  float* s_dc = sm + (2 * F + kAcc) * E;    // [kWarps][F]
  const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < F * E / 4; i += kThreads) {
    reinterpret_cast<float4*>(s_kt)[i] = __ldg(reinterpret_cast<const float4*>(kt + (size_t)b * F * E) + i);
    reinterpret_cast<float4*>(s_vt)[i] = __ldg(reinterpret_cast<const float4*>(vt + (size_t)b * F * E) + i);
  }
  __syncthreads();
  Row<V> dk[F], dv[F], db;
  float dc[F];
#pragma unroll
  for (int f = 0; f < F; ++f) { zero_row(dk[f]); zero_row(dv[f]); dc[f] = 0.f; }
  zero_row(db);
  const float* xb = X + (size_t)b * xs_b;
  const float* gb = G + (size_t)b * gs_b;
  const int s_begin = chunk * kRowsPerCta, s_end = min(S, s_begin + kRowsPerCta);
  for (int s0 = s_begin + warp * R; s0 < s_end; s0 += kWarps * R) {
    Row<V> x[R], g[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      x[r] = load_row<V>(xb + (size_t)(s0 + r) * xs_r, lane, s0 + r < s_end);
      g[r] = load_row<V>(gb + (size_t)(s0 + r) * gs_r, lane, s0 + r < s_end);
    }
    float d[R * F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const Row<V> v = lds_row<V>(s_vt + f * E, lane);
#pragma unroll
      for (int r = 0; r < R; ++r) d[r * F + f] = dot_part(g[r], v);
    }
    allreduce<R * F>(d, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool ok = s0 + r < s_end;
      const size_t ri = ((size_t)b * S + min(s0 + r, S - 1)) * F;
      float sc[F], mx = -INFINITY;
#pragma unroll
      for (int f = 0; f < F; ++f) { sc[f] = __ldg(raw + ri + f); mx = fmaxf(mx, sc[f]); }
      float sum = 0.f, p[F];
#pragma unroll
      for (int f = 0; f < F; ++f) { p[f] = ex2f((sc[f] - mx) * kLog2e); sum += p[f]; }
      const float inv = ok ? 1.f / sum : 0.f;
      float dsum = 0.f;
#pragma unroll
      for (int f = 0; f < F; ++f) { p[f] *= inv; dsum = fmaf(p[f], d[r * F + f], dsum); }
      Row<V> o;
      zero_row(o);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        float ds = p[f] * (d[r * F + f] - dsum);
        if (draw != nullptr && ok) ds += __ldg(draw + ri + f);
        dc[f] += ds;
        axpy(dk[f], ds, x[r]);
        axpy(dv[f], p[f], g[r]);
        axpy(o, ds, lds_row<V>(s_kt + f * E, lane));
      }
      axpy(db, 1.f, g[r]);      // masked rows were loaded as zeros
      if (ok) store_row<V>(dX + ((size_t)b * S + s0 + r) * E, lane, o);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int f = 0; f < F; ++f) s_dc[warp * F + f] = dc[f];
  }
  for (int w = 0; w < kWarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int j = 0; j < kAcc; ++j) {
        Row<V> t;
        if (w == 0) zero_row(t); else t = lds_row<V>(s_acc + j * E, lane);
        const Row<V>& src = j < F ? dk[j < F ? j : 0] : (j < 2 * F ? dv[j < 2 * F && j >= F ? j - F : 0] : db);
        axpy(t, 1.f, src);
        store_row<V>(s_acc + j * E, lane, t);
      }
    }
    __syncthreads();
  }
  float* po = part + ((size_t)b * nchunk + chunk) * (kAcc * E + F);
  for (int i = tid; i < kAcc * E; i += kThreads) po[i] = s_acc[i];
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

}  // namespace ca
}  // namespace dml

extern "C" {

int dml_coattn_chunks(int S) { return S > 0 ? dml::cdiv(S, dml::ca::kRowsPerCta) : 0; }

size_t dml_coattn_fq_fwd_ws_floats(int B, int F, int S, int E) {
  if (B <= 0 || F <= 0 || S <= 0 || E <= 0) return 0;
  return (size_t)B * dml_coattn_chunks(S) * F * (E + 2);
}
size_t dml_coattn_fq_bwd_ws_floats(int B, int F, int S, int E) {
  if (B <= 0 || F <= 0 || S <= 0 || E <= 0) return 0;
  return (size_t)B * dml_coattn_chunks(S) * F * (E + 1);
}
size_t dml_coattn_fk_bwd_ws_floats(int B, int F, int S, int E) {
  if (B <= 0 || F <= 0 || S <= 0 || E <= 0) return 0;
  return (size_t)B * dml_coattn_chunks(S) * ((2 * F + 1) * E + F);
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
  const int nchunk = dml_coattn_chunks(S);
  dim3 grid(nchunk, B);
  DML_CA_DISPATCH_F(F, {
    const size_t smem = sizeof(float) * (kF * E + kWarps * 2 * kF);
    int e = set_smem(fq_fwd_kernel<kF, V>, smem);
    if (e) return e;
    fq_fwd_kernel<kF, V><<<grid, kThreads, smem, st>>>(x, xs_b, xs_r, qt, c, raw, ws, S);
  });
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  fq_combine_kernel<V><<<dim3(F, B), 128 * V / 4, 0, st>>>(ws, nchunk, F, px, lse);
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
  dim3 grid(dml_coattn_chunks(S), B);
  DML_CA_DISPATCH_F(F, {
    const size_t smem = sizeof(float) * (2 * kF * E + kWarps * kF);
    int e = set_smem(fq_bwd_kernel<kF, V>, smem);
    if (e) return e;
    fq_bwd_kernel<kF, V><<<grid, kThreads, smem, st>>>(x, xs_b, xs_r, qt, raw, lse, dpx, dsum, draw, dx, ws, S);
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
  dim3 grid(dml_coattn_chunks(S), B);
  DML_CA_DISPATCH_F(F, {
    const size_t smem = sizeof(float) * (2 * kF * E);
    int e = set_smem(fk_fwd_kernel<kF, V>, smem);
    if (e) return e;
    fk_fwd_kernel<kF, V><<<grid, kThreads, smem, st>>>(x, xs_b, xs_r, kt, c, vt, bo, raw, out, S);
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
  dim3 grid(dml_coattn_chunks(S), B);
  DML_CA_DISPATCH_F(F, {
    const size_t smem = sizeof(float) * ((4 * kF + 1) * E + kWarps * kF);
    int e = set_smem(fk_bwd_kernel<kF, V>, smem);
    if (e) return e;
    fk_bwd_kernel<kF, V><<<grid, kThreads, smem, st>>>(x, xs_b, xs_r, dout, gs_b, gs_r, kt, vt, raw, draw, dx, ws, S);
  });
  DML_RETURN_LAUNCH();
}

}  // extern "C"
