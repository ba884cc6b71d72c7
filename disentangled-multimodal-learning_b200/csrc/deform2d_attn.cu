// DeformCrossAttention2D attention core on the tensor cores (models/DeformableAttention2D.py:290-321 and its adjoint):
//   forward   S = scale q k^T + bias  ->  attn = softmax(S) (an OUTPUT of the module)  ->  o = dropout(attn) v
//   backward  dP = dO v^T (+ dA),  D = sum P dP,  dS = P (dP - D),  dq = scale dS k,  dk = scale dS^T q,  dv = P^T dO
// Every product is mma.sync.m16n8k16 on THREE bf16 parts per operand (24 bits, 6 MMAs: q, k, v and the gradients have no known
// range, and dS feeds the cancellation-dominated sums of the position-bias backward, deform2d_bias.cu).  The products are a small
// part of the work (2 x 64 MACs per pair against 1 120 in the bias MLP); what the tensor cores remove is the shared-memory operand
// traffic that bound the fp32 CUDA-core version of these kernels (8 LDS per 16 FMAs).
// Row kernels: warp = 16 query rows whose A fragments (q or dO, split in the kernel) stay in registers; 48-key tiles of the
// pre-split k / v planes stream through a double-buffered, XOR-swizzled cp.async ring; B fragments by ldmatrix (k as [key][channel]
// directly, v through ldmatrix.trans); softmax statistics on the accumulator fragments (thread-local + two shuffles per row);
// probabilities / dS go from the accumulator layout straight into A fragments (two n-tiles = one k-step).
// The column pass (dk, dv) stays on the CUDA cores (see the note at the kernel).
#include "common.cuh"

namespace dml {
namespace {

constexpr int kC = 512, kHd = 64;
// tile of R rows x 64 channels in three bf16 parts, 128-byte rows, 16-byte chunks XOR-swizzled by the row
template <int R>
struct TileGeo {
  static constexpr int kRows = R, kNT = R / 8, kNTP = R / 16, kKS = R / 16;
  static constexpr uint32_t kPart = R * kHd * 2, kBytes = 3 * kPart;
};
using KeyTile = TileGeo<48>;      // row kernels: 48 keys per tile (the reference's 144 keys = 3 tiles exactly)
constexpr uint32_t kRowSmem = 2 * KeyTile::kBytes;       // double buffer

__device__ __forceinline__ void split3(float x0, float x1, uint32_t& h, uint32_t& m, uint32_t& l) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
  const float r0 = x0 - bf16_lo_f(h), r1 = x1 - bf16_hi_f(h);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(m) : "f"(r1), "f"(r0));
  const float s0 = r0 - bf16_lo_f(m), s1 = r1 - bf16_hi_f(m);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(s1), "f"(s0));
}

// x fp32 [rows, 512] -> planes bf16 [3][rows][512] (as 32-bit words [3][rows][256])
__global__ void __launch_bounds__(256) planes_kernel(const float* __restrict__ x, long long rows, float mult, uint32_t* __restrict__ planes) {
  const long long total = rows * 128;                       // float4 per thread
  const size_t plane = (size_t)rows * 256;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const float4 v = *reinterpret_cast<const float4*>(x + i * 4);
    uint32_t h0, m0, l0, h1, m1, l1;
    split3(v.x * mult, v.y * mult, h0, m0, l0);
    split3(v.z * mult, v.w * mult, h1, m1, l1);
    uint32_t* o = planes + i * 2;
    *reinterpret_cast<uint2*>(o) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(o + plane) = make_uint2(m0, m1);
    *reinterpret_cast<uint2*>(o + 2 * plane) = make_uint2(l0, l1);
  }
}

// R rows (row0 .. of `nrows`) x 64 channels (head h) of the three planes -> swizzled shared tile; rows beyond nrows = 0
template <class G>
__device__ __forceinline__ void load_tile_async(unsigned char* dst, const uint32_t* __restrict__ planes, size_t plane_w, size_t base_row,
                                                int row0, int nrows, int h) {
  for (int c = threadIdx.x; c < 3 * G::kRows * 8; c += blockDim.x) {
    const int part = c / (G::kRows * 8), rc = c - part * (G::kRows * 8), row = rc >> 3, ch = rc & 7;
    const bool ok = row0 + row < nrows;
    const uint32_t* src = planes + part * plane_w + (base_row + (ok ? row0 + row : 0)) * 256 + h * 32 + ch * 4;
    cp_async16(smem_u32(dst + part * G::kPart + row * 128 + ((ch ^ (row & 7)) << 4)), src, ok);
  }
}

// B fragments of one k-step (16 channels) for operands stored [n][k] (k = channel contiguous): all n-tiles (rows of the tile), 3 parts
template <class G>
__device__ __forceinline__ void bfrags_nk(uint32_t sbase, int ks, int lane, uint32_t (&bf)[3][G::kNT][2]) {
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int ntp = 0; ntp < G::kNTP; ++ntp) {
      const int mi = lane >> 3, row = 16 * ntp + (mi >> 1) * 8 + (lane & 7), ch = 2 * ks + (mi & 1);
      uint32_t r4[4];
      ldmatrix_x4(r4, sbase + p * G::kPart + row * 128 + ((ch ^ (row & 7)) << 4));
      bf[p][2 * ntp][0] = r4[0];
      bf[p][2 * ntp][1] = r4[1];
      bf[p][2 * ntp + 1][0] = r4[2];
      bf[p][2 * ntp + 1][1] = r4[3];
    }
}
// B fragments of one k-step for operands stored [k][n] (n = channel contiguous, k = row of the tile): rows 16 ks .. 16 ks + 15, the
// 8 n-tiles of the 64 channels
template <class G>
__device__ __forceinline__ void bfrags_kn(uint32_t sbase, int ks, int lane, uint32_t (&bf)[3][8][2]) {
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int ntp = 0; ntp < 4; ++ntp) {
      const int mi = lane >> 3, row = 16 * ks + (mi & 1) * 8 + (lane & 7), ch = 2 * ntp + (mi >> 1);
      uint32_t r4[4];
      ldmatrix_x4_trans(r4, sbase + p * G::kPart + row * 128 + ((ch ^ (row & 7)) << 4));
      bf[p][2 * ntp][0] = r4[0];
      bf[p][2 * ntp][1] = r4[1];
      bf[p][2 * ntp + 1][0] = r4[2];
      bf[p][2 * ntp + 1][1] = r4[3];
    }
}
// acc[nt] += A (three parts) x B (three parts): 6 MMAs per n-tile, smallest terms first, term-major
template <int NT>
__device__ __forceinline__ void mma6(float (&acc)[NT][4], const uint32_t (&a0)[4], const uint32_t (&a1)[4], const uint32_t (&a2)[4],
                                     const uint32_t (&bf)[3][NT][2]) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) mma_bf16_16816(acc[nt], a2, bf[0][nt][0], bf[0][nt][1]);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) mma_bf16_16816(acc[nt], a0, bf[2][nt][0], bf[2][nt][1]);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) mma_bf16_16816(acc[nt], a1, bf[1][nt][0], bf[1][nt][1]);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) mma_bf16_16816(acc[nt], a1, bf[0][nt][0], bf[0][nt][1]);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) mma_bf16_16816(acc[nt], a0, bf[1][nt][0], bf[1][nt][1]);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) mma_bf16_16816(acc[nt], a0, bf[0][nt][0], bf[0][nt][1]);
}

// A fragments (three parts, four k-steps over the 64 channels) of rows r0 (regs 0, 2) and r1 (regs 1, 3) of a fp32 [.., 512] tensor
__device__ __forceinline__ void afrags_rows(const float* __restrict__ x, size_t row0, size_t row1, bool ok0, bool ok1, int h, int t, float mult,
                                            uint32_t (&a)[3][4][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int col = h * 64 + 16 * ks + 2 * t + 8 * half;
      float2 v0 = make_float2(0.f, 0.f), v1 = v0;
      if (ok0) v0 = *reinterpret_cast<const float2*>(x + row0 * kC + col);
      if (ok1) v1 = *reinterpret_cast<const float2*>(x + row1 * kC + col);
      split3(v0.x * mult, v0.y * mult, a[0][ks][2 * half], a[1][ks][2 * half], a[2][ks][2 * half]);
      split3(v1.x * mult, v1.y * mult, a[0][ks][2 * half + 1], a[1][ks][2 * half + 1], a[2][ks][2 * half + 1]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The row kernels are bound by the latency of their global accesses to the map (two warps per scheduler), so every read of
// the map is issued one stage ahead of its use: the values of the NEXT tile are requested before the MMAs of the current one.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kNT = KeyTile::kNT;        // score n-tiles per key tile (6)

// the thread's 2 x (2 kNT) values of a [rows, m] fp32 map for the tile starting at key j0: v[nt][0..1] row r0, v[nt][2..3] row r1
__device__ __forceinline__ void load_map(const float* __restrict__ p0, const float* __restrict__ p1, bool ok0, bool ok1, int j0, int m, int t,
                                         float (&v)[kNT][4], float fill) {
#pragma unroll
  for (int nt = 0; nt < kNT; ++nt) {
    const int j = j0 + 8 * nt + 2 * t;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      v[nt][u] = (ok0 && j + u < m) ? p0[j + u] : fill;
      v[nt][2 + u] = (ok1 && j + u < m) ? p1[j + u] : fill;
    }
  }
}
__device__ __forceinline__ void load_keep(const unsigned char* __restrict__ p0, const unsigned char* __restrict__ p1, bool ok0, bool ok1, int j0,
                                          int m, int t, uint32_t& bits) {
  bits = 0u;                                                 // bit (4 nt + c) = keep flag of value v[nt][c]
#pragma unroll
  for (int nt = 0; nt < kNT; ++nt) {
    const int j = j0 + 8 * nt + 2 * t;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (ok0 && j + u < m && p0[j + u]) bits |= 1u << (4 * nt + u);
      if (ok1 && j + u < m && p1[j + u]) bits |= 1u << (4 * nt + 2 + u);
    }
  }
}

// forward rows: attn in = bias, out = probabilities; o [B, n, 512]
__global__ void __launch_bounds__(256, 1) attn_fwd_mma_kernel(const float* __restrict__ q, const uint32_t* __restrict__ kplanes,
                                                               const uint32_t* __restrict__ vplanes, float* __restrict__ attn,
                                                               const unsigned char* __restrict__ keep, float keep_scale, int n, int m,
                                                               float scale, float* __restrict__ o) {
  extern __shared__ __align__(128) unsigned char smem[];
  using G = KeyTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y >> 3, h = blockIdx.y & 7;
  const int i0 = blockIdx.x * 128 + warp * 16 + g, i1 = i0 + 8;
  const bool ok0 = i0 < n, ok1 = i1 < n;
  const size_t plane_w = (size_t)(gridDim.y >> 3) * m * 256;
  const size_t krow = (size_t)b * m;
  uint32_t a[3][4][4];
  afrags_rows(q, (size_t)b * n + i0, (size_t)b * n + i1, ok0, ok1, h, t, scale, a);
  float* arow0 = attn + (((size_t)(b * 8 + h) * n) + (ok0 ? i0 : 0)) * m;
  float* arow1 = attn + (((size_t)(b * 8 + h) * n) + (ok1 ? i1 : 0)) * m;
  const unsigned char* krow0 = keep ? keep + (((size_t)(b * 8 + h) * n) + (ok0 ? i0 : 0)) * m : nullptr;
  const unsigned char* krow1 = keep ? keep + (((size_t)(b * 8 + h) * n) + (ok1 ? i1 : 0)) * m : nullptr;
  const int ntiles = cdiv(m, G::kRows);
  float mx0 = -INFINITY, mx1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  // ---- pass 1: raw scores into the map, running maximum and sum (per thread over its columns; merged over the quad afterwards)
  load_tile_async<G>(smem, kplanes, plane_w, krow, 0, m, h);
  cp_async_commit();
  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1, j0 = tile * G::kRows;
    if (tile + 1 < ntiles) load_tile_async<G>(smem + (buf ^ 1) * G::kBytes, kplanes, plane_w, krow, j0 + G::kRows, m, h);
    cp_async_commit();
    float bv[kNT][4];
    load_map(arow0, arow1, ok0, ok1, j0, m, t, bv, 0.f);    // the bias of this tile: in flight during the MMAs
    cp_async_wait<1>();
    __syncthreads();
    const uint32_t sbase = smem_u32(smem + buf * G::kBytes);
    float acc[kNT][4];
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t bf[3][kNT][2];
      bfrags_nk<G>(sbase, ks, lane, bf);
      mma6<kNT>(acc, a[0][ks], a[1][ks], a[2][ks], bf);
    }
    float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt) {
      const int j = j0 + 8 * nt + 2 * t;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (j + u < m) {
          acc[nt][u] += bv[nt][u];
          acc[nt][2 + u] += bv[nt][2 + u];
          if (ok0) arow0[j + u] = acc[nt][u];
          if (ok1) arow1[j + u] = acc[nt][2 + u];
          tm0 = fmaxf(tm0, acc[nt][u]);
          tm1 = fmaxf(tm1, acc[nt][2 + u]);
        } else {
          acc[nt][u] = acc[nt][2 + u] = -INFINITY;
        }
      }
    }
    const float nm0 = fmaxf(mx0, tm0), nm1 = fmaxf(mx1, tm1);
    float s0 = 0.f, s1 = 0.f;
    if (nm0 > -INFINITY) {
#pragma unroll
      for (int nt = 0; nt < kNT; ++nt) s0 += __expf(acc[nt][0] - nm0) + __expf(acc[nt][1] - nm0);
      l0 = l0 * __expf(mx0 - nm0) + s0;
      mx0 = nm0;
    }
    if (nm1 > -INFINITY) {
#pragma unroll
      for (int nt = 0; nt < kNT; ++nt) s1 += __expf(acc[nt][2] - nm1) + __expf(acc[nt][3] - nm1);
      l1 = l1 * __expf(mx1 - nm1) + s1;
      mx1 = nm1;
    }
    __syncthreads();
  }
  // merge the statistics of the four threads of a row
#pragma unroll
  for (int off = 1; off <= 2; off <<= 1) {
    const float om0 = __shfl_xor_sync(0xffffffffu, mx0, off), ol0 = __shfl_xor_sync(0xffffffffu, l0, off);
    const float om1 = __shfl_xor_sync(0xffffffffu, mx1, off), ol1 = __shfl_xor_sync(0xffffffffu, l1, off);
    const float nm0 = fmaxf(mx0, om0), nm1 = fmaxf(mx1, om1);
    l0 = (mx0 > -INFINITY ? l0 * __expf(mx0 - nm0) : 0.f) + (om0 > -INFINITY ? ol0 * __expf(om0 - nm0) : 0.f);
    l1 = (mx1 > -INFINITY ? l1 * __expf(mx1 - nm1) : 0.f) + (om1 > -INFINITY ? ol1 * __expf(om1 - nm1) : 0.f);
    mx0 = nm0;
    mx1 = nm1;
  }
  const float inv0 = 1.f / fmaxf(l0, 1e-30f), inv1 = 1.f / fmaxf(l1, 1e-30f);
  // ---- pass 2: probabilities out, o += P v (the raw scores of tile + 1 are requested before the MMAs of tile)
  float oacc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) oacc[nt][0] = oacc[nt][1] = oacc[nt][2] = oacc[nt][3] = 0.f;
  load_tile_async<G>(smem, vplanes, plane_w, krow, 0, m, h);
  cp_async_commit();
  float raw[kNT][4];
  uint32_t kb = 0xffffffffu;
  load_map(arow0, arow1, ok0, ok1, 0, m, t, raw, -INFINITY);
  if (keep) load_keep(krow0, krow1, ok0, ok1, 0, m, t, kb);
  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1, j0 = tile * G::kRows;
    if (tile + 1 < ntiles) load_tile_async<G>(smem + (buf ^ 1) * G::kBytes, vplanes, plane_w, krow, j0 + G::kRows, m, h);
    cp_async_commit();
    float p[kNT][4];
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt) {
      const int j = j0 + 8 * nt + 2 * t;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float p0 = __expf(raw[nt][u] - mx0) * inv0, p1 = __expf(raw[nt][2 + u] - mx1) * inv1;      // exp(-inf) = 0 beyond m
        if (j + u < m) {
          if (ok0) arow0[j + u] = p0;
          if (ok1) arow1[j + u] = p1;
        }
        if (keep) {
          p0 = (kb >> (4 * nt + u)) & 1u ? p0 * keep_scale : 0.f;
          p1 = (kb >> (4 * nt + 2 + u)) & 1u ? p1 * keep_scale : 0.f;
        }
        p[nt][u] = ok0 ? p0 : 0.f;
        p[nt][2 + u] = ok1 ? p1 : 0.f;
      }
    }
    if (tile + 1 < ntiles) {
      load_map(arow0, arow1, ok0, ok1, j0 + G::kRows, m, t, raw, -INFINITY);
      if (keep) load_keep(krow0, krow1, ok0, ok1, j0 + G::kRows, m, t, kb);
    }
    cp_async_wait<1>();
    __syncthreads();
    const uint32_t sbase = smem_u32(smem + buf * G::kBytes);
#pragma unroll
    for (int ks = 0; ks < G::kKS; ++ks) {                  // 16 keys per k-step = n-tiles 2 ks, 2 ks + 1 of the score tile
      uint32_t pa[3][4];
      split3(p[2 * ks][0], p[2 * ks][1], pa[0][0], pa[1][0], pa[2][0]);
      split3(p[2 * ks][2], p[2 * ks][3], pa[0][1], pa[1][1], pa[2][1]);
      split3(p[2 * ks + 1][0], p[2 * ks + 1][1], pa[0][2], pa[1][2], pa[2][2]);
      split3(p[2 * ks + 1][2], p[2 * ks + 1][3], pa[0][3], pa[1][3], pa[2][3]);
      uint32_t bf[3][8][2];
      bfrags_kn<G>(sbase, ks, lane, bf);
      mma6<8>(oacc, pa[0], pa[1], pa[2], bf);
    }
    __syncthreads();
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = h * 64 + 8 * nt + 2 * t;
    if (ok0) *reinterpret_cast<float2*>(o + ((size_t)b * n + i0) * kC + col) = make_float2(oacc[nt][0], oacc[nt][1]);
    if (ok1) *reinterpret_cast<float2*>(o + ((size_t)b * n + i1) * kC + col) = make_float2(oacc[nt][2], oacc[nt][3]);
  }
}

// backward rows: ds [B, 8, n, m] out, dq [B, n, 512] out
__global__ void __launch_bounds__(256, 1) attn_bwd_rows_mma_kernel(const uint32_t* __restrict__ kplanes, const uint32_t* __restrict__ vplanes,
                                                                    const float* __restrict__ attn, const float* __restrict__ dO,
                                                                    const float* __restrict__ dA, const unsigned char* __restrict__ keep,
                                                                    float keep_scale, int n, int m, float scale, float* __restrict__ ds,
                                                                    float* __restrict__ dq) {
  extern __shared__ __align__(128) unsigned char smem[];
  using G = KeyTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y >> 3, h = blockIdx.y & 7;
  const int i0 = blockIdx.x * 128 + warp * 16 + g, i1 = i0 + 8;
  const bool ok0 = i0 < n, ok1 = i1 < n;
  const size_t plane_w = (size_t)(gridDim.y >> 3) * m * 256;
  const size_t krow = (size_t)b * m;
  uint32_t a[3][4][4];
  afrags_rows(dO, (size_t)b * n + i0, (size_t)b * n + i1, ok0, ok1, h, t, 1.f, a);
  const size_t r0 = (((size_t)(b * 8 + h) * n) + (ok0 ? i0 : 0)) * m, r1 = (((size_t)(b * 8 + h) * n) + (ok1 ? i1 : 0)) * m;
  const int ntiles = cdiv(m, G::kRows);
  float D0 = 0.f, D1 = 0.f;
  // ---- pass 1: dP = dO v^T (* keep) + dA -> ds (raw), D = sum P dP
  load_tile_async<G>(smem, vplanes, plane_w, krow, 0, m, h);
  cp_async_commit();
  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1, j0 = tile * G::kRows;
    if (tile + 1 < ntiles) load_tile_async<G>(smem + (buf ^ 1) * G::kBytes, vplanes, plane_w, krow, j0 + G::kRows, m, h);
    cp_async_commit();
    float pv[kNT][4], av[kNT][4];
    uint32_t kb = 0xffffffffu;
    load_map(attn + r0, attn + r1, ok0, ok1, j0, m, t, pv, 0.f);          // in flight during the MMAs
    if (dA) load_map(dA + r0, dA + r1, ok0, ok1, j0, m, t, av, 0.f);
    if (keep) load_keep(keep + r0, keep + r1, ok0, ok1, j0, m, t, kb);
    cp_async_wait<1>();
    __syncthreads();
    const uint32_t sbase = smem_u32(smem + buf * G::kBytes);
    float acc[kNT][4];
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t bf[3][kNT][2];
      bfrags_nk<G>(sbase, ks, lane, bf);
      mma6<kNT>(acc, a[0][ks], a[1][ks], a[2][ks], bf);
    }
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt) {
      const int j = j0 + 8 * nt + 2 * t;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int u = c & 1;
        const bool okr = (c < 2) ? ok0 : ok1;
        if (okr && j + u < m) {
          float d = acc[nt][c];
          if (keep) d = (kb >> (4 * nt + c)) & 1u ? d * keep_scale : 0.f;
          if (dA) d += av[nt][c];
          ds[((c < 2) ? r0 : r1) + j + u] = d;
          if (c < 2) D0 = fmaf(pv[nt][c], d, D0);
          else D1 = fmaf(pv[nt][c], d, D1);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int off = 1; off <= 2; off <<= 1) {
    D0 += __shfl_xor_sync(0xffffffffu, D0, off);
    D1 += __shfl_xor_sync(0xffffffffu, D1, off);
  }
  // ---- pass 2: dS = P (dP - D) -> ds, dq += dS k (P and dP of tile + 1 are requested before the MMAs of tile)
  float qacc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) qacc[nt][0] = qacc[nt][1] = qacc[nt][2] = qacc[nt][3] = 0.f;
  load_tile_async<G>(smem, kplanes, plane_w, krow, 0, m, h);
  cp_async_commit();
  float pn[kNT][4], dn[kNT][4];
  load_map(attn + r0, attn + r1, ok0, ok1, 0, m, t, pn, 0.f);
  load_map(ds + r0, ds + r1, ok0, ok1, 0, m, t, dn, 0.f);
  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1, j0 = tile * G::kRows;
    if (tile + 1 < ntiles) load_tile_async<G>(smem + (buf ^ 1) * G::kBytes, kplanes, plane_w, krow, j0 + G::kRows, m, h);
    cp_async_commit();
    float s[kNT][4];
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt) {
      const int j = j0 + 8 * nt + 2 * t;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int u = c & 1;
        const bool okr = (c < 2) ? ok0 : ok1;
        float v = 0.f;
        if (okr && j + u < m) {
          v = pn[nt][c] * (dn[nt][c] - ((c < 2) ? D0 : D1));
          ds[((c < 2) ? r0 : r1) + j + u] = v;
        }
        s[nt][c] = v;
      }
    }
    if (tile + 1 < ntiles) {
      load_map(attn + r0, attn + r1, ok0, ok1, j0 + G::kRows, m, t, pn, 0.f);
      load_map(ds + r0, ds + r1, ok0, ok1, j0 + G::kRows, m, t, dn, 0.f);
    }
    cp_async_wait<1>();
    __syncthreads();
    const uint32_t sbase = smem_u32(smem + buf * G::kBytes);
#pragma unroll
    for (int ks = 0; ks < G::kKS; ++ks) {
      uint32_t pa[3][4];
      split3(s[2 * ks][0], s[2 * ks][1], pa[0][0], pa[1][0], pa[2][0]);
      split3(s[2 * ks][2], s[2 * ks][3], pa[0][1], pa[1][1], pa[2][1]);
      split3(s[2 * ks + 1][0], s[2 * ks + 1][1], pa[0][2], pa[1][2], pa[2][2]);
      split3(s[2 * ks + 1][2], s[2 * ks + 1][3], pa[0][3], pa[1][3], pa[2][3]);
      uint32_t bf[3][8][2];
      bfrags_kn<G>(sbase, ks, lane, bf);
      mma6<8>(qacc, pa[0], pa[1], pa[2], bf);
    }
    __syncthreads();
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = h * 64 + 8 * nt + 2 * t;
    if (ok0) *reinterpret_cast<float2*>(dq + ((size_t)b * n + i0) * kC + col) = make_float2(qacc[nt][0] * scale, qacc[nt][1] * scale);
    if (ok1) *reinterpret_cast<float2*>(dq + ((size_t)b * n + i1) * kC + col) = make_float2(qacc[nt][2] * scale, qacc[nt][3] * scale);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward columns on the CUDA cores (fp32).  A tensor-core version of this pass (P^T / dS^T tiles split and transposed with
// movmatrix, q / dO planes as B operands) was built and measured at 181 us against 150 us for this kernel at B = 4 x 2 500 x 144
// (six warps per SM at 242 registers): the transposing operand path costs more than the shared-memory traffic it removes.
// ---------------------------------------------------------------------------------------------------------------------
// Backward over the columns: dv[j] = sum_i (P keep)_ij dO_i, dk[j] = scale sum_i dS_ij q_i.  CTA = 32 keys x one query chunk, a
// thread = 4 keys x 4 channels of both products for half of the rows of a 32-query tile (4 LDS.128 per 32 FMAs);
// parts[chunk][2][B, m, 512] holds dk and dv of the chunk (summed afterwards by reduce_parts).
__global__ void __launch_bounds__(256) attn_bwd_cols_kernel(const float* __restrict__ q, const float* __restrict__ attn,
                                                            const float* __restrict__ ds, const float* __restrict__ dO,
                                                            const unsigned char* __restrict__ keep, float keep_scale, int n, int m,
                                                            int B, float scale, int chunk_rows, float* __restrict__ parts) {
  __shared__ __align__(16) float qs[32][68], gs[32][68];
  __shared__ __align__(16) float ps[32][36], ss[32][36];              // [query][key]
  const int b = blockIdx.y >> 3, h = blockIdx.y & 7;
  const int j0 = blockIdx.x * 32;
  const int half = threadIdx.x >> 7, tl = threadIdx.x & 127;
  const int kq = (tl >> 4) * 4, cq = (tl & 15) * 4;                     // 4 keys x 4 channels
  const int i_begin = blockIdx.z * chunk_rows, i_end = min(n, i_begin + chunk_rows);
  const size_t arow = ((size_t)(b * 8 + h) * n) * m;
  float ak[4][4], av[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int w = 0; w < 4; ++w) ak[u][w] = av[u][w] = 0.f;
  for (int it = i_begin; it < i_end; it += 32) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 16; i += 256) {
      const int r = i >> 4, c4 = i & 15;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), g = a;
      if (it + r < i_end) {
        a = *reinterpret_cast<const float4*>(q + ((size_t)b * n + it + r) * kC + h * 64 + c4 * 4);
        g = *reinterpret_cast<const float4*>(dO + ((size_t)b * n + it + r) * kC + h * 64 + c4 * 4);
      }
      *reinterpret_cast<float4*>(&qs[r][c4 * 4]) = a;
      *reinterpret_cast<float4*>(&gs[r][c4 * 4]) = g;
    }
    for (int i = threadIdx.x; i < 32 * 32; i += 256) {
      const int r = i >> 5, c = i & 31;
      float p = 0.f, s = 0.f;
      if (it + r < i_end && j0 + c < m) {
        const size_t at = arow + (size_t)(it + r) * m + j0 + c;
        p = attn[at];
        if (keep) p = keep[at] ? p * keep_scale : 0.f;
        s = ds[at];
      }
      ps[r][c] = p;
      ss[r][c] = s;
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < 16; ++rr) {
      const int r = half * 16 + rr;
      const float4 p4 = *reinterpret_cast<const float4*>(&ps[r][kq]), s4 = *reinterpret_cast<const float4*>(&ss[r][kq]);
      const float4 g4 = *reinterpret_cast<const float4*>(&gs[r][cq]), q4 = *reinterpret_cast<const float4*>(&qs[r][cq]);
      const float pv[4] = {p4.x, p4.y, p4.z, p4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        av[u][0] += pv[u] * g4.x; av[u][1] += pv[u] * g4.y; av[u][2] += pv[u] * g4.z; av[u][3] += pv[u] * g4.w;
        ak[u][0] += sv[u] * q4.x; ak[u][1] += sv[u] * q4.y; ak[u][2] += sv[u] * q4.z; ak[u][3] += sv[u] * q4.w;
      }
    }
  }
  // combine the two row halves through shared memory (fixed order), then write the chunk's partial
  __syncthreads();
  float* red = &qs[0][0];                                             // 32 x 68 floats >= 128 threads x 16
  float* red2 = &gs[0][0];
  if (half == 1) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        red[(u * 4 + w) * 128 + tl] = ak[u][w];
        red2[(u * 4 + w) * 128 + tl] = av[u][w];
      }
  }
  __syncthreads();
  if (half == 0) {
    const size_t plane = (size_t)B * m * kC;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (j0 + kq + u < m) {
        float* o = parts + (size_t)blockIdx.z * 2 * plane + ((size_t)b * m + j0 + kq + u) * kC + h * 64 + cq;
        float4 vk, vv;
        vk.x = (ak[u][0] + red[(u * 4 + 0) * 128 + tl]) * scale;
        vk.y = (ak[u][1] + red[(u * 4 + 1) * 128 + tl]) * scale;
        vk.z = (ak[u][2] + red[(u * 4 + 2) * 128 + tl]) * scale;
        vk.w = (ak[u][3] + red[(u * 4 + 3) * 128 + tl]) * scale;
        vv.x = av[u][0] + red2[(u * 4 + 0) * 128 + tl];
        vv.y = av[u][1] + red2[(u * 4 + 1) * 128 + tl];
        vv.z = av[u][2] + red2[(u * 4 + 2) * 128 + tl];
        vv.w = av[u][3] + red2[(u * 4 + 3) * 128 + tl];
        *reinterpret_cast<float4*>(o) = vk;
        *reinterpret_cast<float4*>(o + plane) = vv;
      }
    }
  }
}

__global__ void reduce_chunks_kernel(const float* __restrict__ parts, int nparts, long long len, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += parts[(size_t)p * len + i];
  out[i] = s;
}

inline int plane_blocks(long long rows) {
  const long long b = (rows * 128 + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

}  // namespace
}  // namespace dml

using namespace dml;

extern "C" {

/* bytes of the scratch the attention entry points need: three bf16 planes of k and of v */
size_t dml_da2_attn_ws_bytes(int B, int n, int m, int backward) {
  (void)n;
  (void)backward;
  return (size_t)2 * 3 * B * m * kC * 2;
}

int dml_da2_attn_fwd(const float* q, const float* k, const float* v, float* attn, const unsigned char* keep, float keep_scale, int B, int n,
                     int m, float scale, void* ws, float* o, void* stream) {
  DML_CHECK_ARG(q && k && v && attn && o && ws && B > 0 && n > 0 && m > 0 && B * 8 <= 65535);
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* kp = reinterpret_cast<uint32_t*>(ws);
  uint32_t* vp = kp + (size_t)3 * B * m * 256;
  planes_kernel<<<plane_blocks((long long)B * m), 256, 0, st>>>(k, (long long)B * m, 1.f, kp);
  planes_kernel<<<plane_blocks((long long)B * m), 256, 0, st>>>(v, (long long)B * m, 1.f, vp);
  cudaError_t e = cudaFuncSetAttribute(attn_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowSmem);
  if (e != cudaSuccess) return (int)e;
  attn_fwd_mma_kernel<<<dim3(cdiv(n, 128), B * 8), 256, kRowSmem, st>>>(q, kp, vp, attn, keep, keep_scale, n, m, scale, o);
  DML_RETURN_LAUNCH();
}

int dml_da2_cols_chunks(int B, int n, int m) {
  const int base = B * 8 * cdiv(m, 32);
  int chunks = cdiv(148 * 8, base);
  const int max_chunks = cdiv(n, 64);
  if (chunks > max_chunks) chunks = max_chunks;
  return chunks < 1 ? 1 : chunks;
}

/* ds: float [B, 8, n, m] (out: dS); dq [B, n, 512]; dkv [2][B, m, 512] = dk, dv; parts: float [chunks][2][B, m, 512] */
int dml_da2_attn_bwd(const float* q, const float* k, const float* v, const float* attn, const float* dO, const float* dA,
                     const unsigned char* keep, float keep_scale, int B, int n, int m, float scale, void* ws, float* ds, float* dq,
                     float* parts, float* dkv, void* stream) {
  DML_CHECK_ARG(q && k && v && attn && dO && ws && ds && dq && parts && dkv && B > 0 && n > 0 && m > 0 && B * 8 <= 65535);
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* kp = reinterpret_cast<uint32_t*>(ws);
  uint32_t* vp = kp + (size_t)3 * B * m * 256;
  planes_kernel<<<plane_blocks((long long)B * m), 256, 0, st>>>(k, (long long)B * m, 1.f, kp);
  planes_kernel<<<plane_blocks((long long)B * m), 256, 0, st>>>(v, (long long)B * m, 1.f, vp);
  cudaError_t e = cudaFuncSetAttribute(attn_bwd_rows_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowSmem);
  if (e != cudaSuccess) return (int)e;
  attn_bwd_rows_mma_kernel<<<dim3(cdiv(n, 128), B * 8), 256, kRowSmem, st>>>(kp, vp, attn, dO, dA, keep, keep_scale, n, m, scale, ds, dq);
  const int chunks = dml_da2_cols_chunks(B, n, m);
  const int chunk_rows = cdiv(cdiv(n, chunks), 32) * 32;
  attn_bwd_cols_kernel<<<dim3(cdiv(m, 32), B * 8, chunks), 256, 0, st>>>(q, attn, ds, dO, keep, keep_scale, n, m, B, scale, chunk_rows, parts);
  const long long len = (long long)2 * B * m * kC;
  reduce_chunks_kernel<<<(unsigned)((len + 255) / 256), 256, 0, st>>>(parts, chunks, len, dkv);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
