// Piecewise-linear table of the continuous-position-bias MLP (reference: CPB, DeformableAttention1D.py:60-102).
//
// The CPB network maps ONE scalar t = sign(p) log(|p|+1) through Linear(1,H)+ReLU, Linear(H,H)+ReLU,
// Linear(H,O) (H = dim/4 <= 32, O = heads/groups <= 2).  A scalar-input ReLU MLP is an exactly
// piecewise-linear function of t with at most H + H*(H+1) breakpoints, so per step we extract the
// breakpoints once (cpb_table_build) and every (query i, key j) pair evaluates
//     bias_o = a[s][o] * t + c[s][o],   s = segment containing t
// instead of 2*H*H dense MACs.  Same function, identical up to fp32 rounding; its backward is the
// per-segment sums  A[s][o] = sum delta_o,  B[s][o] = sum delta_o * t  (see cpb_param_grad).
#pragma once
#include "common.cuh"

namespace dml {

constexpr int kCpbHidMax = 32;
constexpr int kCpbOutMax = 2;
constexpr int kCpbCells = 4096;                                   // uniform cells over [-T, T]
constexpr int kCpbSegMax = kCpbHidMax + kCpbHidMax * (kCpbHidMax + 1) + 1 + 15;  // 1104 (padded)
constexpr int kCpbBpPad = 16;

// Table layout in 32-bit words (one device buffer, produced by cpb_table_build):
constexpr int kTabHdr = 0;                                        // int nseg, int kmax, float T, float inv_cell, int hid, int nout
constexpr int kTabCoef = 16;                                      // float4[kCpbSegMax]   (a0, c0, a1, c1)
constexpr int kTabBp = kTabCoef + 4 * kCpbSegMax;                 // float[kCpbSegMax+pad] upper boundary of segment s (+inf past the end)
constexpr int kTabCell = kTabBp + kCpbSegMax + kCpbBpPad;         // uint16[kCpbCells]     first candidate segment of each cell
constexpr int kTabSmemWords = kTabCell + kCpbCells / 2;           // everything above is what the attention kernels stage in smem
constexpr int kTabMask1 = kTabSmemWords;                          // uint32[kCpbSegMax]   active set of layer 1
constexpr int kTabMask2 = kTabMask1 + kCpbSegMax;                 // uint32[kCpbSegMax]   active set of layer 2
constexpr int kTabWords = kTabMask2 + kCpbSegMax;

struct CpbView {            // pointers into the (shared-memory) copy of the table
  const float4* coef;
  const float* bpf;
  const uint16_t* cellseg;
  float T, inv_cell;
  int kmax, nseg;
};

__device__ __forceinline__ CpbView cpb_view(const uint32_t* tab) {
  CpbView v;
  v.coef = reinterpret_cast<const float4*>(tab + kTabCoef);
  v.bpf = reinterpret_cast<const float*>(tab + kTabBp);
  v.cellseg = reinterpret_cast<const uint16_t*>(tab + kTabCell);
  v.nseg = (int)tab[0];
  v.kmax = (int)tab[1];
  v.T = __uint_as_float(tab[2]);
  v.inv_cell = __uint_as_float(tab[3]);
  return v;
}

// Cooperative copy of the smem part of the table (all threads of the CTA; caller syncs).
__device__ __forceinline__ void cpb_stage(uint32_t* dst, const uint32_t* __restrict__ src, int tid, int nthreads) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  for (int i = tid; i < kTabSmemWords / 4; i += nthreads) d4[i] = __ldg(s4 + i);
}

// t = sign(p) * log(|p| + 1)   (DeformableAttention1D.py:93)
__device__ __forceinline__ float cpb_t(float p) {
  float L = __log2f(fabsf(p) + 1.0f) * kLn2;
  return copysignf(L, p);
}

__device__ __forceinline__ int cpb_segment(const CpbView& tb, float t) {
  int cell = (int)((t + tb.T) * tb.inv_cell);
  cell = min(max(cell, 0), kCpbCells - 1);
  int s = tb.cellseg[cell];
  for (int k = 0; k < tb.kmax; ++k) s += (t >= tb.bpf[s]) ? 1 : 0;
  return s;
}

}  // namespace dml
