// cpb_table_build / cpb_param_grad: exact piecewise-linear form of the CPB bias MLP and its
// parameter gradients.  Reference maths: CPB.forward, DeformableAttention1D.py:84-102.
#include <math.h>

#include "../../include/dml_b200.h"
#include "cpb_table.cuh"

namespace dml {

constexpr int kSortN = 2048;  // >= 32 + 33*32 candidates

struct MlpSmem {
  float w1[32], b1[32], W2[32 * 32], b2[32], W3[2 * 32], b3[2];
};

__device__ void load_mlp(MlpSmem& m, const float* w1, const float* b1, const float* W2, const float* b2,
                         const float* W3, const float* b3, int hid, int nout, int tid, int nthreads) {
  for (int i = tid; i < 32; i += nthreads) {
    m.w1[i] = i < hid ? w1[i] : 0.f;
    m.b1[i] = i < hid ? b1[i] : 0.f;
    m.b2[i] = i < hid ? b2[i] : 0.f;
  }
  for (int i = tid; i < 1024; i += nthreads) {
    int k = i >> 5, mm = i & 31;
    m.W2[i] = (k < hid && mm < hid) ? W2[k * hid + mm] : 0.f;
  }
  for (int i = tid; i < 64; i += nthreads) {
    int o = i >> 5, k = i & 31;
    m.W3[i] = (o < nout && k < hid) ? W3[o * hid + k] : 0.f;
  }
  if (tid < 2) m.b3[tid] = tid < nout ? b3[tid] : 0.f;
}

__global__ void __launch_bounds__(1024, 1)
cpb_table_build_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ W2,
                       const float* __restrict__ b2, const float* __restrict__ W3, const float* __restrict__ b3,
                       int hid, int nout, float T, uint32_t* __restrict__ table) {
  __shared__ double cand[kSortN];
  __shared__ double l1[34];
  __shared__ MlpSmem m;
  __shared__ int s_n1, s_nbp, s_kmax;
  __shared__ unsigned s_amax[2];
  const int tid = threadIdx.x;
  const double dT = (double)T;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);

  load_mlp(m, w1, b1, W2, b2, W3, b3, hid, nout, tid, blockDim.x);
  for (int i = tid; i < kSortN; i += blockDim.x) cand[i] = INF;
  if (tid == 0) { s_kmax = 0; s_amax[0] = 0u; s_amax[1] = 0u; }
  __syncthreads();

  // ---- layer-1 breakpoints: w1[k] t + b1[k] = 0 ----
  if (tid < 32) {
    double t = INF;
    if (m.w1[tid] != 0.f) {
      t = -(double)m.b1[tid] / (double)m.w1[tid];
      if (!(t > -dT && t < dT)) t = INF;
    }
    cand[tid] = t;
  }
  __syncthreads();
  if (tid == 0) {  // 32 values: insertion sort is fine
    double a[32];
    int n1 = 0;
    for (int i = 0; i < 32; ++i)
      if (cand[i] < INF) {
        double v = cand[i];
        int j = n1++;
        while (j > 0 && a[j - 1] > v) { a[j] = a[j - 1]; --j; }
        a[j] = v;
      }
    for (int i = 0; i < n1; ++i) l1[i + 1] = a[i];
    l1[0] = -dT;
    l1[n1 + 1] = dT;
    s_n1 = n1;
  }
  __syncthreads();
  const int n1 = s_n1;

  // ---- layer-2 breakpoints inside each layer-1 interval ----
  for (int idx = tid; idx < 33 * 32; idx += blockDim.x) {
    int q = idx >> 5, k = idx & 31;
    if (q > n1 || k >= hid) continue;
    double lo = l1[q], hi = l1[q + 1];
    double tm = 0.5 * (lo + hi);
    double A = 0.0, B = (double)m.b2[k];
    for (int mm = 0; mm < 32; ++mm) {
      double pre = (double)m.w1[mm] * tm + (double)m.b1[mm];
      if (pre > 0.0) {
        A += (double)m.W2[k * 32 + mm] * (double)m.w1[mm];
        B += (double)m.W2[k * 32 + mm] * (double)m.b1[mm];
      }
    }
    if (A != 0.0) {
      double r = -B / A;
      if (r > lo && r < hi) cand[32 + idx] = r;
    }
  }
  __syncthreads();

  // ---- bitonic sort (ascending) of kSortN doubles ----
  for (int k = 2; k <= kSortN; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < kSortN; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          double a = cand[i], b = cand[ixj];
          bool up = ((i & k) == 0);
          if ((a > b) == up) { cand[i] = b; cand[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  int cnt = 0;
  for (int i = tid; i < kSortN; i += blockDim.x) cnt += (cand[i] < INF) ? 1 : 0;
  __shared__ int s_cnt;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  atomicAdd(&s_cnt, cnt);
  __syncthreads();
  const int nbp = min(s_cnt, kCpbSegMax - 1 - 15);
  const int nseg = nbp + 1;
  if (tid == 0) s_nbp = nbp;

  float4* cellcoef = reinterpret_cast<float4*>(table + kTabCellCoef);
  float* cellbp = reinterpret_cast<float*>(table + kTabCellBp);
  uint16_t* cellseg = reinterpret_cast<uint16_t*>(table + kTabCellSeg);
  float4* segcoef = reinterpret_cast<float4*>(table + kTabSegCoef);
  float* segbp = reinterpret_cast<float*>(table + kTabSegBp);
  uint32_t* mask1 = table + kTabMask1;
  uint32_t* mask2 = table + kTabMask2;
  const double LN2 = 0.693147180559945309417, LOG2E = 1.44269504088896340736;

  // ---- per-segment affine coefficients, evaluated from the active sets at the segment midpoint ----
  // stored as (a0, c0*log2e, a1, c1*log2e): in x = t/ln2 and in the softmax's log2 domain the slope is unchanged.
  for (int s = tid; s < kCpbSegMax + kCpbBpPad; s += blockDim.x) {
    if (s < nseg) {
      double lo = (s == 0) ? -dT : cand[s - 1];
      double hi = (s == nseg - 1) ? dT : cand[s];
      double tm = 0.5 * (lo + hi);
      uint32_t m1 = 0, m2 = 0;
      for (int mm = 0; mm < 32; ++mm)
        if ((double)m.w1[mm] * tm + (double)m.b1[mm] > 0.0) m1 |= 1u << mm;
      double a[2] = {0.0, 0.0}, c[2] = {(double)m.b3[0], (double)m.b3[1]};
      for (int k = 0; k < 32; ++k) {
        double P = 0.0, Q = (double)m.b2[k];
        for (int mm = 0; mm < 32; ++mm)
          if (m1 >> mm & 1u) {
            P += (double)m.W2[k * 32 + mm] * (double)m.w1[mm];
            Q += (double)m.W2[k * 32 + mm] * (double)m.b1[mm];
          }
        if (P * tm + Q > 0.0 && k < hid) {
          m2 |= 1u << k;
          a[0] += (double)m.W3[k] * P;      c[0] += (double)m.W3[k] * Q;
          a[1] += (double)m.W3[32 + k] * P; c[1] += (double)m.W3[32 + k] * Q;
        }
      }
      segcoef[s] = make_float4((float)a[0], (float)(c[0] * LOG2E), (float)a[1], (float)(c[1] * LOG2E));
      atomicMax(&s_amax[0], __float_as_uint(fabsf((float)a[0]) * 1.0000002f));   // non-negative floats order like uints
      atomicMax(&s_amax[1], __float_as_uint(fabsf((float)a[1]) * 1.0000002f));
      mask1[s] = m1;
      mask2[s] = m2;
    } else if (s < kCpbSegMax) {
      segcoef[s] = make_float4(0.f, 0.f, 0.f, 0.f);
      mask1[s] = 0;
      mask2[s] = 0;
    }
    segbp[s] = (s < nbp) ? (float)(cand[s] / LN2) : __int_as_float(0x7f800000);
  }
  __syncthreads();

  // ---- uniform cells over x in [-X, X]: affine pieces on both sides of the first breakpoint of the cell ----
  const double X = dT / LN2;
  const double cw = 2.0 * X / (double)kCpbCells;
  int my_dirty = 0;
  for (int c = tid; c < kCpbCells; c += blockDim.x) {
    const double x0 = (-X + c * cw) * LN2, x1 = (-X + (c + 1) * cw) * LN2;   // cell bounds in t units
    int lo = 0, hi = nbp;  // #{bp <= x0}
    while (lo < hi) { int mid = (lo + hi) >> 1; if (cand[mid] <= x0) lo = mid + 1; else hi = mid; }
    const int s0 = lo;
    lo = s0; hi = nbp;     // #{bp < x1}
    while (lo < hi) { int mid = (lo + hi) >> 1; if (cand[mid] < x1) lo = mid + 1; else hi = mid; }
    const int inside = lo - s0;
    const int s1 = inside >= 1 ? s0 + 1 : s0;
    float bpv = __int_as_float(0x7f800000);
    if (inside == 1) bpv = (float)(cand[s0] / LN2);
    else if (inside >= 2) { bpv = __int_as_float(0x7fc00000); ++my_dirty; }
    const float4 e0 = segcoef[s0], e1 = segcoef[s1];
    cellcoef[c] = make_float4(e0.x, e0.y, e1.x, e1.y);
    cellcoef[kCpbCells + c] = make_float4(e0.z, e0.w, e1.z, e1.w);
    cellbp[c] = bpv;
    cellseg[c] = (uint16_t)s0;
  }
  atomicAdd(&s_kmax, my_dirty);
  __syncthreads();
  if (tid == 0) {
    table[0] = (uint32_t)nseg;
    table[1] = (uint32_t)s_kmax;   // number of flagged (>= 2 breakpoints) cells
    table[2] = __float_as_uint((float)X);
    table[3] = __float_as_uint((float)((double)kCpbCells / (2.0 * X)));
    table[4] = (uint32_t)hid;
    table[5] = (uint32_t)nout;
    table[6] = s_amax[0];
    table[7] = s_amax[1];
    for (int i = 8; i < 16; ++i) table[i] = 0;
  }
}

// Parameter gradients from the per-segment sums  segsum[s] = (A0, Bx0, A1, Bx1),
// A_o = sum delta_o, Bx_o = sum delta_o * x (x = t / ln2) over all (i, j) pairs whose t fell in segment s.
// Inside segment s the MLP is  z2_k = P_k t + Q_k  (layer-1 active set m1), out_o = sum_k W3[o,k] act2_k z2_k + b3[o].
// grads layout (floats, must be zeroed by the caller): dw1[32] db1[32] dW2[32*32] db2[32] dW3[2*32] db3[2]
__global__ void __launch_bounds__(1024, 1)
cpb_param_grad_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ W2,
                      const float* __restrict__ b2, const float* __restrict__ W3, const float* __restrict__ b3,
                      int hid, int nout, const uint32_t* __restrict__ table, const float* __restrict__ segsum,
                      float* __restrict__ grads) {
  __shared__ MlpSmem m;
  __shared__ float sU[32], sV[32];
  const int tid = threadIdx.x;
  load_mlp(m, w1, b1, W2, b2, W3, b3, hid, nout, tid, blockDim.x);
  __syncthreads();
  const int nseg = (int)table[0];
  const uint32_t* mask1 = table + kTabMask1;
  const uint32_t* mask2 = table + kTabMask2;
  const int k = tid >> 5, mm = tid & 31;  // thread (k, mm) owns dW2[k][mm]
  float acc_W2 = 0.f, acc_w1 = 0.f, acc_b1 = 0.f, acc_b2 = 0.f, acc_W3[2] = {0.f, 0.f}, acc_b3[2] = {0.f, 0.f};

  for (int s = blockIdx.x; s < nseg; s += gridDim.x) {
    const float4 ss = reinterpret_cast<const float4*>(segsum)[s];
    const float A[2] = {ss.x, ss.z}, B[2] = {ss.y * kLn2, ss.w * kLn2};   // sums of delta*x arrive in x = t/ln2 units
    if (A[0] == 0.f && A[1] == 0.f && B[0] == 0.f && B[1] == 0.f) continue;  // uniform across the CTA
    const uint32_t m1 = mask1[s], m2 = mask2[s];
    if (tid < 32) {
      const int kk = tid;
      float P = 0.f, Q = m.b2[kk];
      for (int j = 0; j < 32; ++j)
        if (m1 >> j & 1u) { P += m.W2[kk * 32 + j] * m.w1[j]; Q += m.W2[kk * 32 + j] * m.b1[j]; }
      const bool on = (m2 >> kk) & 1u;
      float U = 0.f, V = 0.f;
      if (on) {
        U = m.W3[kk] * A[0] + m.W3[32 + kk] * A[1];
        V = m.W3[kk] * B[0] + m.W3[32 + kk] * B[1];
        acc_W3[0] += P * B[0] + Q * A[0];
        acc_W3[1] += P * B[1] + Q * A[1];
      }
      sU[kk] = U;
      sV[kk] = V;
      acc_b2 += U;
      if (kk < 2) acc_b3[0] += A[kk];  // thread kk accumulates db3[kk]
    }
    __syncthreads();
    if (m1 >> mm & 1u) acc_W2 += m.w1[mm] * sV[k] + m.b1[mm] * sU[k];
    if (tid < 32 && (m1 >> tid & 1u)) {
      float X = 0.f, Y = 0.f;
      for (int kk = 0; kk < 32; ++kk) { X += m.W2[kk * 32 + tid] * sU[kk]; Y += m.W2[kk * 32 + tid] * sV[kk]; }
      acc_b1 += X;
      acc_w1 += Y;
    }
    __syncthreads();
  }
  float* dw1 = grads, *db1 = grads + 32, *dW2 = grads + 64, *db2 = grads + 64 + 1024, *dW3 = grads + 96 + 1024,
        *db3 = grads + 160 + 1024;
  if (acc_W2 != 0.f) atomicAdd(dW2 + tid, acc_W2);
  if (tid < 32) {
    atomicAdd(dw1 + tid, acc_w1);
    atomicAdd(db1 + tid, acc_b1);
    atomicAdd(db2 + tid, acc_b2);
    atomicAdd(dW3 + tid, acc_W3[0]);
    atomicAdd(dW3 + 32 + tid, acc_W3[1]);
    if (tid < 2) atomicAdd(db3 + tid, acc_b3[0]);
  }
}

// Evaluate the table at arbitrary t (diagnostics / tests): out[i] = (bias0, bias1) in natural units, seg[i] = segment.
__global__ void cpb_eval_kernel(const uint32_t* __restrict__ table, const float* __restrict__ t, int count,
                                float* __restrict__ out, int* __restrict__ seg) {
  extern __shared__ __align__(16) uint8_t sm[];
  for (int o = 0; o < 2; ++o) {
    __syncthreads();
    const CpbView tb = cpb_stage(sm, table, o, true, threadIdx.x, blockDim.x);
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
      const float x = t[i] * kLog2e;
      float a, c;
      int s = 0;
      cpb_lookup<true>(tb, x, a, c, s);
      out[2 * i + o] = fmaf(a, x, c) * kLn2;
      if (seg && o == 0) seg[i] = s;
    }
  }
}

}  // namespace dml

extern "C" {

int dml_cpb_eval(const void* table, const float* t, int count, float* out, int* seg, void* stream) {
  DML_CHECK_ARG(table && t && out && count > 0);
  {   // per-device attribute: set on every call (cheap, no process-global flag)
    cudaError_t e = cudaFuncSetAttribute(dml::cpb_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dml::kCpbSmemBwdBytes);
    if (e != cudaSuccess) return (int)e;
  }
  dml::cpb_eval_kernel<<<dml::cdiv(count, 256) < 296 ? dml::cdiv(count, 256) : 296, 256, dml::kCpbSmemBwdBytes,
                         (cudaStream_t)stream>>>((const uint32_t*)table, t, count, out, seg);
  DML_RETURN_LAUNCH();
}

size_t dml_cpb_table_bytes(void) { return (size_t)dml::kTabWords * 4; }
int dml_cpb_seg_max(void) { return dml::kCpbSegMax; }

int dml_cpb_table_build(const float* w1, const float* b1, const float* W2, const float* b2, const float* W3,
                        const float* b3, int hid, int nout, float t_max, void* table, void* stream) {
  DML_CHECK_ARG(w1 && b1 && W2 && b2 && W3 && b3 && table);
  if (hid < 1 || hid > dml::kCpbHidMax || nout < 1 || nout > dml::kCpbOutMax) return DML_EUNSUPPORTED;
  DML_CHECK_ARG(t_max > 0.f);
  dml::cpb_table_build_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(w1, b1, W2, b2, W3, b3, hid, nout, t_max,
                                                                   (uint32_t*)table);
  DML_RETURN_LAUNCH();
}

int dml_cpb_param_grad(const float* w1, const float* b1, const float* W2, const float* b2, const float* W3,
                       const float* b3, int hid, int nout, const void* table, const float* segsum, float* grads,
                       void* stream) {
  DML_CHECK_ARG(w1 && b1 && W2 && b2 && W3 && b3 && table && segsum && grads);
  if (hid < 1 || hid > dml::kCpbHidMax || nout < 1 || nout > dml::kCpbOutMax) return DML_EUNSUPPORTED;
  cudaError_t e = cudaMemsetAsync(grads, 0, sizeof(float) * DML_CPB_GRAD_FLOATS, (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  dml::cpb_param_grad_kernel<<<32, 1024, 0, (cudaStream_t)stream>>>(w1, b1, W2, b2, W3, b3, hid, nout,
                                                                   (const uint32_t*)table, segsum, grads);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
