// Piecewise-linear table of the continuous-position-bias MLP (reference: CPB, DeformableAttention1D.py:60-102).
//
// The CPB network maps ONE scalar t = sign(p) log(|p|+1) through Linear(1,H)+ReLU, Linear(H,H)+ReLU,
// Linear(H,O) (H = dim/4 <= 32, O = heads/groups <= 2).  A scalar-input ReLU MLP is an exactly
// piecewise-linear function of t with at most H + H*(H+1) breakpoints, so per step we extract the
// breakpoints once (cpb_table_build) and every (query i, key j) pair evaluates
//     bias_o = a[s][o] * t + c[s][o],   s = segment containing t
// instead of 2*H*H dense MACs.  Same function, identical up to fp32 rounding; its backward is the
// per-segment sums  A[s][o] = sum delta_o,  B[s][o] = sum delta_o * t  (see cpb_param_grad).
//
// Lookup structure.  The kernels work in x = sign(p) log2(|p|+1) = t / ln2 (one MUFU.LG2) and in the
// log2 domain of the softmax, where  bias * log2(e) = a * x + c * log2(e)  with the SAME slope a.
// [-X, X] is cut into kCpbCells uniform cells; per head-output o and cell the table holds the affine
// pieces on both sides of the first breakpoint inside the cell:
//     cellcoef[o][cell] = (a_lo, c_lo*log2e, a_hi, c_hi*log2e),  cellbp[cell] = first breakpoint (x units, +inf if none)
// so the common case is two shared-memory loads and a select, with no loop and no divergence.  Cells that
// contain two or more breakpoints are flagged (cellbp = NaN) and resolved through the per-segment arrays.
#pragma once
#include "common.cuh"

namespace dml {

constexpr int kCpbHidMax = 32;
constexpr int kCpbOutMax = 2;
constexpr int kCpbCells = 1024;
constexpr int kCpbSegMax = kCpbHidMax + kCpbHidMax * (kCpbHidMax + 1) + 1 + 15;  // 1104 (padded)
constexpr int kCpbBpPad = 16;

// Table layout in 32-bit words (one device buffer, produced by cpb_table_build):
constexpr int kTabHdr = 0;                                   // int nseg, int ndirty, float X, float inv_cell, int hid, int nout, float amax0, float amax1 (max |slope| per output)
constexpr int kTabCellCoef = 16;                             // float4[2][kCpbCells]
constexpr int kTabCellBp = kTabCellCoef + 2 * 4 * kCpbCells; // float[kCpbCells]
constexpr int kTabCellSeg = kTabCellBp + kCpbCells;          // uint16[kCpbCells]   segment index at the start of the cell
constexpr int kTabSegCoef = kTabCellSeg + kCpbCells / 2;     // float4[kCpbSegMax]  (a0, c0*log2e, a1, c1*log2e)
constexpr int kTabSegBp = kTabSegCoef + 4 * kCpbSegMax;      // float[kCpbSegMax+pad] upper boundary of segment s in x units (+inf past the end)
constexpr int kTabMask1 = kTabSegBp + kCpbSegMax + kCpbBpPad;  // uint32[kCpbSegMax]  active set of layer 1
constexpr int kTabMask2 = kTabMask1 + kCpbSegMax;            // uint32[kCpbSegMax]  active set of layer 2
constexpr int kTabWords = kTabMask2 + kCpbSegMax;

// shared-memory image used by the attention kernels for ONE head output o:
//   float4 coef[kCpbCells] | float bp[kCpbCells] | uint16 seg[kCpbCells] (backward only)
constexpr int kCpbSmemFwdBytes = kCpbCells * 16 + kCpbCells * 4;
constexpr int kCpbSmemBwdBytes = kCpbSmemFwdBytes + kCpbCells * 2;

struct CpbView {
  const float4* coef;        // smem, this head's output
  const float* bp;           // smem
  const uint16_t* seg;       // smem (backward) or nullptr
  const uint32_t* gtab;      // global table (slow path)
  float c1, c2;              // cell = floor(x * c1 + c2)
  int oidx;
};

// Stage this head's cell table into shared memory (all threads; caller syncs) and build the view.
__device__ __forceinline__ CpbView cpb_stage(uint8_t* smem, const uint32_t* __restrict__ tab, int oidx, bool with_seg,
                                             int tid, int nthreads) {
  float4* coef = reinterpret_cast<float4*>(smem);
  float* bp = reinterpret_cast<float*>(smem + kCpbCells * 16);
  uint16_t* seg = reinterpret_cast<uint16_t*>(smem + kCpbSmemFwdBytes);
  const float4* gcoef = reinterpret_cast<const float4*>(tab + kTabCellCoef) + oidx * kCpbCells;
  for (int i = tid; i < kCpbCells; i += nthreads) coef[i] = __ldg(gcoef + i);
  const float4* gbp = reinterpret_cast<const float4*>(tab + kTabCellBp);
  for (int i = tid; i < kCpbCells / 4; i += nthreads) reinterpret_cast<float4*>(bp)[i] = __ldg(gbp + i);
  if (with_seg) {
    const uint4* gs = reinterpret_cast<const uint4*>(tab + kTabCellSeg);
    for (int i = tid; i < kCpbCells / 8; i += nthreads) reinterpret_cast<uint4*>(seg)[i] = __ldg(gs + i);
  }
  CpbView v;
  v.coef = coef;
  v.bp = bp;
  v.seg = with_seg ? seg : nullptr;
  v.gtab = tab;
  const float X = __uint_as_float(__ldg(tab + 2)), inv = __uint_as_float(__ldg(tab + 3));
  v.c1 = inv;
  v.c2 = X * inv;
  v.oidx = oidx;
  return v;
}

// x = sign(p) * log2(|p| + 1)     (t of DeformableAttention1D.py:93 divided by ln 2)
__device__ __forceinline__ float cpb_x(float p) { return copysignf(__log2f(fabsf(p) + 1.0f), p); }

// Slow path: resolve a flagged cell through the per-segment arrays in global memory.
static __device__ __noinline__ void cpb_slow(const CpbView& tb, int cell, float x, float& a, float& c, int& seg) {
  const uint16_t* cs = reinterpret_cast<const uint16_t*>(tb.gtab + kTabCellSeg);
  const float* sbp = reinterpret_cast<const float*>(tb.gtab + kTabSegBp);
  const float4* sc = reinterpret_cast<const float4*>(tb.gtab + kTabSegCoef);
  int s = cs[cell];
  while (s < kCpbSegMax - 1 && x >= __ldg(sbp + s)) ++s;
  const float4 e = __ldg(sc + s);
  a = tb.oidx ? e.z : e.x;
  c = tb.oidx ? e.w : e.y;
  seg = s;
}

// bias (log2 domain) = a * x + c; returns a, c (and the segment index when kSeg).
template <bool kSeg>
__device__ __forceinline__ void cpb_lookup(const CpbView& tb, float x, float& a, float& c, int& seg) {
  const int cell = min(max(__float2int_rd(fmaf(x, tb.c1, tb.c2)), 0), kCpbCells - 1);
  const float4 e = tb.coef[cell];
  const float b = tb.bp[cell];
  const bool hi = x >= b;
  a = hi ? e.z : e.x;
  c = hi ? e.w : e.y;
  if (kSeg) seg = (int)tb.seg[cell] + (hi ? 1 : 0);
  if (__builtin_expect(b != b, 0)) cpb_slow(tb, cell, x, a, c, seg);
}

}  // namespace dml
