// kernels_tc.cuh — K2x: the batched sweep on the 5th-generation tensor cores (tcgen05.mma kind::tf32, 3xTF32).
//
//   Y[t][k] = sum_p H[p][k] * X[t - p][k]          (FFTConvolver.cpp:176-187, Utilities.cpp:62-111)
//
// For a long launch group the sweep is, per frequency bin k, a 1-D convolution ALONG THE BLOCK INDEX t of the
// bin's time line x_k[t] with the bin's P partition values H[.][k].  Cut t into segments of R = 64 steps: the
// outputs of segment n are a Toeplitz matrix of H applied to a window of x,
//
//   D[i][n] = sum_j A[i][j] * B[j][n],   A[i][j] = H[i + Q - j],   B[j][n] = x[64 n - Q + j],   0 <= j < K = Q + 64
//
// (Q = P - 1 rounded up to 64) — a GEMM with M = 64 outputs per segment, N = segments, K = Q + 64, one per bin
// and channel.  The complex product becomes real GEMMs by stacking [Hr ; Hi] into M = 128 rows and running the
// same A against the real and the imaginary time line (two accumulators D, D2 in tensor memory):
//   y.re = D[0:64] - D2[64:128],  y.im = D[64:128] + D2[0:64]        (entry 0 = DC / Nyquist: y = (D[0:64], D2[64:128]))
// FP32 accuracy comes from the 3xTF32 split: a = a_hi + a_lo, b = b_hi + b_lo (each tf32-exact), and
// a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi accumulated in FP32 (dropped term ~2^-22 relative).
//
// What makes this cheap on sm_100a: the B operand of chunk c (32 values of j) is a ROW-SHIFTED WINDOW of one
// shared-memory strip.  The bin's time line is stored as rows of 64 samples; plane e in {0, 1} holds the 32-sample
// half rows (128 B, SWIZZLE_128B K-major).  B[32c + jj][n] = x[64 (n + c/2) + 32 (c%2) + jj - Q] is row n + c/2
// of plane c%2, i.e. the same strip with the descriptor start address advanced by (c/2) * 128 bytes — the
// hardware applies the 128-byte swizzle to absolute address bits, so a start address inside the 1024-byte swizzle
// atom is fine (tools/tc_probe.cu, profiles/r02_tc_probe.txt).  One 144-row strip per plane (8 planes: re/im x
// hi/lo x e, 147 KB, one 1-D bulk copy each: the global time lines are stored as pre-swizzled strip images) feeds all 32 K chunks of a 128-segment tile: every x sample enters
// shared memory once, and only the Toeplitz tiles of H (16 KB, pre-swizzled images, 1-D bulk copies through a 4-stage
// ring) stream during the tile.
//
// Kernels: k_tc_build_a (H -> tf32 hi/lo Toeplitz tile images, once per IR), k_tc_split_x (timeline rows -> per-bin
// hi/lo time lines), k_tc_sweep (TMA producer warp / MMA warp with one elected issuing lane / 8 epilogue warps reading TMEM),
// k_tc_merge_y (partial planes -> Y rows, combines the complex product).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

namespace pc {
namespace tc {

constexpr int kR = 64;                              // block steps per segment
constexpr int kN = 128;                             // segments per tile (MMA N)
constexpr int kMaxChunks = 32;                      // K chunks of 32 -> K <= 1024, row shifts 0..15
constexpr int kStripRows = kN + 16;
constexpr int kStripBytes = kStripRows * 128;       // 18432 (a multiple of the 1024-byte swizzle atom)
constexpr int kATileBytes = 128 * 128;              // one 128 x 32 tf32 Toeplitz tile image
constexpr int kAStages = 4;
constexpr int kSmemBytes = 8 * kStripBytes + kAStages * kATileBytes + 1024;
constexpr int kThreads = 320;
constexpr int kFlush = 4;                           // K chunks accumulated in tensor memory before the FP32 register add

struct Geom {
  int P, Q, nchunk, nb, nseg, ntile, rows;
  long long Lt, Lty;
};

inline __host__ __device__ Geom make_geom(int P, int nb) {
  Geom g;
  g.P = P;
  g.Q = ((P > 1 ? P - 1 : 0) + 63) / 64 * 64;
  g.nchunk = g.Q / 32 + 2;
  g.nb = nb;
  g.nseg = (nb + kR - 1) / kR;
  g.ntile = (g.nseg + kN - 1) / kN;
  g.rows = g.ntile * kN + 16;
  g.Lt = (long long)g.rows * 64;
  g.Lty = (long long)g.ntile * kN * 64;
  return g;
}
inline __host__ __device__ bool geom_ok(const Geom& g, int B) { return g.nchunk <= kMaxChunks && B % 32 == 0 && g.nb > 0; }

// byte offset of element (row r, float e < 32) in a SWIZZLE_128B K-major image with a 1024-byte aligned base
inline __host__ __device__ uint32_t sw128(uint32_t r, uint32_t e) { return r * 128u + ((((e >> 2) ^ (r & 7u)) & 7u) << 4) + (e & 3u) * 4u; }

// Xt layout: [line][re_hi, re_lo, im_hi, im_lo][e][row R][32 floats], the sample tau = 64 R + 32 e + jj stored at
// 16-byte chunk (jj / 4) ^ (R % 8) of its 128-byte row: a strip (kStripRows consecutive rows of one plane, first row a
// multiple of 8) is ONE contiguous 18 KB piece of global memory that is already the SWIZZLE_128B shared-memory image,
// so the sweep fetches it with a single 1-D bulk copy (no tensor map, no per-row TMA requests).
inline __host__ __device__ size_t xt_index(long long line, int pl, int e, long long R, int jj, int rows) {
  return ((((size_t)line * 4 + pl) * 2 + e) * (size_t)rows + (size_t)R) * 32 + (size_t)(((((jj >> 2) ^ (int)(R & 7)) & 7) << 2) | (jj & 3));
}

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// ---- H -> Toeplitz tile images ---------------------------------------------------------------------------------
struct BuildAParams {
  const float2* H;          // [C][Prows][B]
  long long h_cstride;
  int B, P, Q, nchunk;
  float* A;                 // [C*B lines][nchunk][hi, lo][4096]
};

// grid (nchunk, B, C), block 256
__global__ void __launch_bounds__(256) k_tc_build_a(BuildAParams p) {
  const int c = blockIdx.x, k = blockIdx.y, ch = blockIdx.z;
  const long long line = (long long)ch * p.B + k;
  float* hi = p.A + ((line * p.nchunk + c) * 2) * 4096;
  float* lo = hi + 4096;
  const float2* Hk = p.H + (long long)ch * p.h_cstride + k;
  for (int idx = threadIdx.x; idx < 4096; idx += 256) {
    const int m = idx >> 5, jj = idx & 31;
    const int part = m >> 6, i = m & 63;
    const int pp = i + p.Q - (32 * c + jj);
    float v = 0.0f;
    if (pp >= 0 && pp < p.P) { const float2 h = Hk[(long long)pp * p.B]; v = part ? h.y : h.x; }
    const float vh = tf32_rn(v), vl = tf32_rn(v - vh);
    const uint32_t off = sw128((uint32_t)m, (uint32_t)jj) >> 2;
    hi[off] = vh;
    lo[off] = vl;
  }
}

// ---- timeline rows -> per-bin hi / lo time lines ---------------------------------------------------------------
struct SplitXParams {
  const float2* X;          // [C][R][B]
  long long x_cstride;
  long long row_base;       // timeline row of tau = 0 (= row of output block 0 minus Q); may be negative
  long long row_lo, row_hi; // rows outside [row_lo, row_hi) read as zero
  int B;
  int rows;                 // 64-sample rows per plane
  float* Xt;
};

// grid (rows * 2, B / 32, C), block (32, 8)
__global__ void __launch_bounds__(256) k_tc_split_x(SplitXParams p) {
  __shared__ float2 tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const long long tau0 = (long long)blockIdx.x * 32;
  const int k0 = blockIdx.y * 32, ch = blockIdx.z;
  for (int r = ty; r < 32; r += 8) {
    const long long row = p.row_base + tau0 + r;
    float2 v = make_float2(0.0f, 0.0f);
    if (row >= p.row_lo && row < p.row_hi) v = p.X[(long long)ch * p.x_cstride + row * p.B + k0 + tx];
    tile[r][tx] = v;
  }
  __syncthreads();
  const long long R = tau0 >> 6;
  const int e = (int)((tau0 >> 5) & 1);
  for (int kk = ty; kk < 32; kk += 8) {
    const float2 v = tile[tx][kk];
    const long long line = (long long)ch * p.B + k0 + kk;
    const float rh = tf32_rn(v.x), ih = tf32_rn(v.y);
    p.Xt[xt_index(line, 0, e, R, tx, p.rows)] = rh;
    p.Xt[xt_index(line, 1, e, R, tx, p.rows)] = tf32_rn(v.x - rh);
    p.Xt[xt_index(line, 2, e, R, tx, p.rows)] = ih;
    p.Xt[xt_index(line, 3, e, R, tx, p.rows)] = tf32_rn(v.y - ih);
  }
}

// ---- partial planes -> Y rows ----------------------------------------------------------------------------------
struct MergeYParams {
  const float* Yt;          // [C*B lines][D part0, D part1, D2 part0, D2 part1][Lty]
  long long Lty;
  int B, nb;
  float2* Y;
  long long y_cstride, y_rstride, yrow0;
};

// grid (ceil(nb / 32), B / 32, C), block (32, 8)
__global__ void __launch_bounds__(256) k_tc_merge_y(MergeYParams p) {
  __shared__ float2 tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const long long t0 = (long long)blockIdx.x * 32;
  const int k0 = blockIdx.y * 32, ch = blockIdx.z;
  for (int kk = ty; kk < 32; kk += 8) {
    const long long line = (long long)ch * p.B + k0 + kk;
    const float* src = p.Yt + line * 4 * p.Lty + t0 + tx;
    float2 y = make_float2(0.0f, 0.0f);
    if (t0 + tx < p.nb) {
      const float d0 = src[0], d1 = src[p.Lty], e0 = src[2 * p.Lty], e1 = src[3 * p.Lty];
      y = (k0 + kk == 0) ? make_float2(d0, e1) : make_float2(d0 - e1, d1 + e0);
    }
    tile[tx][kk] = y;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const long long t = t0 + r;
    if (t < p.nb) p.Y[(long long)ch * p.y_cstride + (p.yrow0 + t) * p.y_rstride + k0 + tx] = tile[r][tx];
  }
}

// ---- the sweep -------------------------------------------------------------------------------------------------
struct SweepParams {
  const float* A;
  const float* Xt;
  float* Yt;
  int lines, ntile, nchunk, rows;
  long long Lty;
  int dbg;                  // timing experiments only (tools/tc_sweep_test.cu): 1 no Yt stores, 2 strips loaded once, 4 A ring loaded once
  int* err;                 // (mapped host word) set non-zero when a barrier wait gave up — a bug, not a data condition
};

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive1(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, unsigned parity) {
  const long long t0 = clock64();
  unsigned ok = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    if (ok) return true;
    if (clock64() - t0 > 4000000000LL) return false;
  }
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// SWIZZLE_128B K-major operand descriptor (rows of 128 bytes, 8-row groups 1024 bytes apart), split in two words:
// only the low word (start address in 16-byte units | leading-offset field) changes between MMAs
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
               :: "r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kDescHi), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}

// grid: any (persistent, tiles walked round-robin); block 320 = TMA producer warp, MMA warp, 8 epilogue warps.
// Tensor memory: two accumulator buffers of [D | D2] (2 x 256 columns).  The MMA warp fills a buffer with the products
// of kFlush chunks and hands it to the epilogue warps, which add it to FP32 registers (round-to-nearest) while the
// other buffer fills: the tensor core's accumulate truncates, so short accumulation chains keep the error at the
// level of the FFMA sweep, and the epilogue of a tile overlaps the MMAs of the next one.
__global__ void __launch_bounds__(kThreads, 1) k_tc_sweep(SweepParams P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* strips = base;                           // [comp][hi, lo][e] x kStripBytes
  unsigned char* ring = base + 8 * kStripBytes;
  __shared__ unsigned long long bar_strip_full[2], bar_strip_empty, bar_tmem_full[2], bar_tmem_empty[2], bar_a_full[kAStages], bar_a_empty[kAStages];
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) {
    mbar_init(&bar_strip_full[0], 1); mbar_init(&bar_strip_full[1], 1); mbar_init(&bar_strip_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_tmem_full[i], 1); mbar_init(&bar_tmem_empty[i], 8); }
    for (int i = 0; i < kAStages; ++i) { mbar_init(&bar_a_full[i], 1); mbar_init(&bar_a_empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_addr(&tmem_holder)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  const int total = P.lines * P.ntile;

  if (warp == 0) {
    if (lane == 0) {                                      // ---- producer
      unsigned it_a = 0;
      int n = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < total && ok; tile += gridDim.x, ++n) {
        const int line = tile / P.ntile, nt = tile - line * P.ntile;
        const bool load_strips = !(P.dbg & 2) || n == 0;
        if (load_strips && n > 0 && !mbar_wait(&bar_strip_empty, (unsigned)(n - 1) & 1u)) { *reinterpret_cast<volatile int*>(P.err) = 1; break; }
        for (int e = 0; e < 2 && load_strips; ++e) {      // plane e = 0 first: chunk 0 needs only that one
          mbar_expect(&bar_strip_full[e], 4 * kStripBytes);
          for (int pl = 0; pl < 4; ++pl)
            bulk_load(strips + (pl * 2 + e) * kStripBytes, P.Xt + ((((size_t)line * 4 + pl) * 2 + e) * (size_t)P.rows + (size_t)nt * kN) * 32, kStripBytes,
                      &bar_strip_full[e]);
        }
        const float* Aline = P.A + (size_t)line * P.nchunk * 2 * 4096;
        for (int s = 0; s < P.nchunk * 2; ++s, ++it_a) {
          const unsigned stage = it_a % kAStages, use = it_a / kAStages;
          if ((P.dbg & 4) && use > 0) continue;
          if (use > 0 && !mbar_wait(&bar_a_empty[stage], (use - 1) & 1u)) { *reinterpret_cast<volatile int*>(P.err) = 2; ok = false; break; }
          mbar_expect(&bar_a_full[stage], kATileBytes);
          bulk_load(ring + stage * kATileBytes, Aline + (size_t)s * 4096, kATileBytes, &bar_a_full[stage]);
        }
      }
    }
  } else if (warp == 1) {                                 // ---- MMA issuer: the warp stays converged, one elected lane issues
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t strip_lo = desc_lo(smem_addr(strips)), ring_lo = desc_lo(smem_addr(ring));
    unsigned it_a = 0, gcount = 0;
    int n = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < total && ok; tile += gridDim.x, ++n) {
      for (int c = 0; c < P.nchunk && ok; ++c) {
        const int g_in = c % kFlush;
        const uint32_t buf = gcount & 1u;
        if (g_in == 0 && gcount >= 2 && !mbar_wait(&bar_tmem_empty[buf], ((gcount >> 1) - 1) & 1u)) { if (lane == 0) *reinterpret_cast<volatile int*>(P.err) = 4; ok = false; break; }
        if (c < 2 && (!(P.dbg & 2) || n == 0) && !mbar_wait(&bar_strip_full[c], (unsigned)n & 1u)) { if (lane == 0) *reinterpret_cast<volatile int*>(P.err) = 3; ok = false; break; }
        const uint32_t e = (uint32_t)c & 1u, q = (uint32_t)c >> 1;
        const bool last_of_group = (g_in == kFlush - 1) || (c == P.nchunk - 1);
#pragma unroll
        for (int hl = 0; hl < 2; ++hl, ++it_a) {
          const unsigned stage = it_a % kAStages, use = it_a / kAStages;
          if ((!(P.dbg & 4) || use == 0) && !mbar_wait(&bar_a_full[stage], use & 1u)) { if (lane == 0) *reinterpret_cast<volatile int*>(P.err) = 5; ok = false; break; }
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = ring_lo + stage * (kATileBytes >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
              for (int comp = 0; comp < 2; ++comp) {
                const uint32_t d = tmem + buf * 256u + (uint32_t)comp * kN;
                const uint32_t bhi = strip_lo + (((comp * 2 + 0) * 2 + e) * kStripBytes >> 4) + q * 8 + kk * 2;
                const uint32_t blo = strip_lo + (((comp * 2 + 1) * 2 + e) * kStripBytes >> 4) + q * 8 + kk * 2;
                if (hl == 0) {
                  mma_tf32(d, a_lo + kk * 2, bhi, idesc, (g_in | kk) ? 1u : 0u);
                  mma_tf32(d, a_lo + kk * 2, blo, idesc, 1u);
                } else {
                  mma_tf32(d, a_lo + kk * 2, bhi, idesc, 1u);
                }
              }
            }
            mma_commit(&bar_a_empty[stage]);
            if (hl == 1 && last_of_group) mma_commit(&bar_tmem_full[buf]);
            if (hl == 1 && c == P.nchunk - 1) mma_commit(&bar_strip_empty);
          }
          __syncwarp();
        }
        if (last_of_group) ++gcount;
      }
    }
  } else {                                                // ---- epilogue: warp w reads TMEM lanes 32 (w % 4) .., 64 columns
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int m = quarter * 32 + lane, part = m >> 6, i = m & 63;
    const int ngroups = (P.nchunk + kFlush - 1) / kFlush;
    unsigned gcount = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < total && ok; tile += gridDim.x) {
      const int line = tile / P.ntile, nt = tile - line * P.ntile;
      float acc[2][64];
#pragma unroll
      for (int comp = 0; comp < 2; ++comp)
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[comp][j] = 0.0f;
      for (int g = 0; g < ngroups; ++g, ++gcount) {
        const uint32_t buf = gcount & 1u;
        if (!mbar_wait(&bar_tmem_full[buf], (gcount >> 1) & 1u)) { if (lane == 0) *reinterpret_cast<volatile int*>(P.err) = 6; ok = false; break; }
        tc_fence_after();
#pragma unroll
        for (int comp = 0; comp < 2; ++comp) {
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + buf * 256u + (uint32_t)(comp * kN + half * 64 + h2 * 32);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                           "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                           "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[comp][h2 * 32 + j] += __uint_as_float(v[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive1(&bar_tmem_empty[buf]);
      }
      if (!ok) break;
      if (P.dbg & 1) continue;
#pragma unroll
      for (int comp = 0; comp < 2; ++comp) {
        float* dst = P.Yt + ((size_t)line * 4 + comp * 2 + part) * P.Lty + ((size_t)nt * kN + half * 64) * 64 + i;
#pragma unroll
        for (int j = 0; j < 64; ++j) dst[(size_t)j * 64] = acc[comp][j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

#endif  // __CUDACC__

}  // namespace tc
}  // namespace pc
