// kernels_fft512.cuh — K1r / K3r: register-resident real FFTs for the block size of the headline shapes
// (B = 512: 1024-point real transform = 512-point complex transform + even/odd split).
//
// Replaces, for B = 512, the shared-memory Stockham kernels of kernels.cuh (CopyAndPad + AudioFFT::fft,
// FFTConvolver.cpp:172-173 / AudioFFT.cpp:114-137;  AudioFFT::ifft + Sum + overlap save, FFTConvolver.cpp:190-204 /
// AudioFFT.cpp:139-159).  Those spend ~2300 warp-instructions per transform, half of them index arithmetic
// and shared-memory traffic of three out-of-place radix-8 passes, at 33 % occupancy (two 4 KB ping-pong buffers per
// transform): the round-1 captures show them issue-bound at ~50 % of the issue slots.  Here one WARP owns one
// transform and keeps its 16 complex points per lane in registers through a three-step 8 x 8 x 8 decomposition
//
//   n = 64 n2 + 8 n1 + n0        k = k2 + 8 q0 + 64 q1                (all digits in [0, 8))
//   W512^{nk} = W8^{n2 k2} . W512^{(8 n1 + n0) k2} . W8^{n1 q0} . W64^{n0 q0} . W8^{n0 q1}
//
//   step 1  DFT8 over n2 for the two columns m = 8 n1 + n0 in {lane, lane + 32}   -> twiddle W512^{m k2}
//   step 2  DFT8 over n1 for the two pairs (k2, n0) = (pid / 8, pid % 8), pid in {lane, lane + 32} -> twiddle W64^{n0 q0}
//   step 3  DFT8 over n0 for the two residues l = k2 + 8 q0 in {lane, 64 - lane}  (lane 0: {0, 32})
//
// with two exchanges through ONE 4 KB shared buffer per warp (XOR-swizzled so that both the writing and the reading
// side of each exchange are bank-conflict free: every half-warp of a 64-bit access touches 16 distinct bank pairs)
// and __syncwarp() only.  Everything that touches global memory is coalesced: for a fixed
// register index consecutive lanes hold consecutive points (first step: consecutive columns; last step: consecutive /
// mirrored residues).  The lane that owns residue l also owns 64 - l, i.e. bins k and M - k of every mirror pair, so
// the even/odd split of the real transform (and its inverse, merged with the frequency-domain overlap-add) needs no
// data movement at all.  The forward transform exploits the zero half of [x ; 0] (n2 >= 4 is zero), the inverse
// computes only the first half of its output (n2 < 4) — the other half is what the overlap-add merge replaced.
// The inverse runs the same three steps backwards with conjugated twiddles.
//
// Tables (device array `tab512`, 1088 float2, built in double on the host, staged in shared memory once per CTA):
//   T1[k2 * 64 + m] = exp(-2 pi i m k2 / 512)     T2[a * 8 + b] = exp(-2 pi i a b / 64)     TS[k] = exp(-2 pi i k / 1024), k < 512
//
// The per-lane phase bodies are plain inline functions so that tests/emu runs the identical arithmetic on the CPU.
#pragma once

#include "kernels.cuh"

namespace pc {

constexpr int kF512_M = 512;
constexpr int kF512_T1 = 0, kF512_T2 = 512, kF512_TS = 576, kF512_TabLen = 1088;
constexpr int kF512_Xch = 512;          // float2 per warp exchange buffer

// complex add / sub / scale on float2: on the device these are single packed-FP32 instructions (FADD2 / FFMA2)
PC_HD float2 f2_add(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  return __fadd2_rn(a, b);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
PC_HD float2 f2_sub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a);
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}
PC_HD float2 f2_scale(float2 a, float s) {
#if defined(__CUDA_ARCH__)
  return __fmul2_rn(a, make_float2(s, s));
#else
  return make_float2(a.x * s, a.y * s);
#endif
}
// a * w (INV: a * conj(w))
template <bool INV>
PC_HD float2 f2_cmul(float2 a, float2 w) {
  if (INV) w.y = -w.y;
  // (a.x w.x - a.y w.y, a.x w.y + a.y w.x) = (a.x, a.x) * w + (-a.y w.y, a.y w.x)
  const float2 t = make_float2(-a.y * w.y, a.y * w.x);
#if defined(__CUDA_ARCH__)
  return __ffma2_rn(make_float2(a.x, a.x), w, t);
#else
  return make_float2(fmaf(a.x, w.x, t.x), fmaf(a.x, w.y, t.y));
#endif
}

// 8-point DFT, natural order in and out (forward: e^{-2 pi i / 8}; INV: conjugate), unscaled
template <bool INV>
PC_HD void f512_dft8(float2* a) {
  float2 s[8];
#pragma unroll
  for (int r = 0; r < 4; ++r) { s[r] = f2_add(a[r], a[r + 4]); s[r + 4] = f2_sub(a[r], a[r + 4]); }
  const float h = 0.70710678118654752440f;
  if (!INV) {   // W8^1, W8^2, W8^3 on the odd half
    s[5] = make_float2(h * (s[5].x + s[5].y), h * (s[5].y - s[5].x));
    s[6] = make_float2(s[6].y, -s[6].x);
    s[7] = make_float2(h * (s[7].y - s[7].x), -h * (s[7].x + s[7].y));
  } else {
    s[5] = make_float2(h * (s[5].x - s[5].y), h * (s[5].x + s[5].y));
    s[6] = make_float2(-s[6].y, s[6].x);
    s[7] = make_float2(-h * (s[7].x + s[7].y), h * (s[7].x - s[7].y));
  }
#pragma unroll
  for (int g = 0; g < 2; ++g) {     // 4-point DFTs: g = 0 -> even outputs 0,2,4,6 ; g = 1 -> odd outputs 1,3,5,7
    const float2 b0 = f2_add(s[4 * g], s[4 * g + 2]), b1 = f2_sub(s[4 * g], s[4 * g + 2]);
    const float2 b2 = f2_add(s[4 * g + 1], s[4 * g + 3]), d = f2_sub(s[4 * g + 1], s[4 * g + 3]);
    const float2 b3 = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
    a[g] = f2_add(b0, b2);
    a[g + 2] = f2_add(b1, b3);
    a[g + 4] = f2_sub(b0, b2);
    a[g + 6] = f2_sub(b1, b3);
  }
}

// residues owned by a lane in step 3 (and, mirrored, in the first inverse step)
PC_HD int f512_la(int lane) { return lane; }
PC_HD int f512_lb(int lane) { return lane == 0 ? 32 : 64 - lane; }
// exchange-buffer layouts (float2 index; 16 consecutive float2 = the 32 banks).
// s1 [k2][m = 8 n1 + n0]: written with consecutive m per k2, read with (k2, k2 + 1) x n0 = 0..7 per half-warp at a
//    fixed n1 -> bit 3 of m is flipped for odd k2 so that the two k2 land in different halves of the bank set.
// s2 [k2][q0][n0]: written with (k2, k2 + 1) x n0 = 0..7 at a fixed q0, read with k2 = 0..7 x (q0, q0 + 1) at a fixed
//    n0 -> n0 ^ k2 spreads the eight k2 over eight bank pairs, bit 0 of q0 ^ k2 picks the half.
PC_HD int f512_s1(int k2, int m) { return k2 * 64 + (m ^ ((k2 & 1) << 3)); }
PC_HD int f512_s2(int k2, int q0, int n0) { return k2 * 64 + ((q0 ^ (k2 & 1)) << 3) + (n0 ^ k2); }

// ---------------------------------------------------------------------------------------------------------
// forward: z[n] = (x[2n], x[2n+1]), n < 256 valid (the upper half of [x ; 0] is zero)
// ---------------------------------------------------------------------------------------------------------
// loads of step 1: the 8 non-zero points of the lane's two columns (z[64 n2 + m], n2 < 4), a[4 h + n2]
PC_HD void f512_fwd_load(int lane, const float* src, int nv, bool vec, float2* a) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int m = lane + 32 * h;
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
      const int n = 64 * n2 + m;
      if (vec) {
        a[4 * h + n2] = *reinterpret_cast<const float2*>(src + 2 * n);
      } else {
        const int i0 = 2 * n, i1 = i0 + 1;
        a[4 * h + n2] = make_float2(i0 < nv ? src[i0] : 0.0f, i1 < nv ? src[i1] : 0.0f);
      }
    }
  }
}

// step 1: DFT8 over n2 (upper half of the input is the zero padding) + twiddle, result into the exchange buffer (layout s1)
PC_HD void f512_fwd_p1(int lane, const float2* in, float2* S, const float2* tab) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int m = lane + 32 * h;
    float2 a[8];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) a[n2] = in[4 * h + n2];
#pragma unroll
    for (int n2 = 4; n2 < 8; ++n2) a[n2] = make_float2(0.0f, 0.0f);
    f512_dft8<false>(a);
    S[f512_s1(0, m)] = a[0];
#pragma unroll
    for (int k2 = 1; k2 < 8; ++k2) S[f512_s1(k2, m)] = f2_cmul<false>(a[k2], tab[kF512_T1 + k2 * 64 + m]);
  }
}

// step 2 (both directions): the two (k2, n0) pairs of the lane; forward reads layout s1 over n1, inverse layout s2 over q0
template <bool INV>
PC_HD void f512_mid_load(int lane, const float2* S, const float2* tab, float2* A, float2* B) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float2* a = h ? B : A;
    const int pid = lane + 32 * h, k2 = pid >> 3, n0 = pid & 7;
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = INV ? S[f512_s2(k2, j, n0)] : S[f512_s1(k2, 8 * j + n0)];
    f512_dft8<INV>(a);
    // forward: C[q0] *= W64^{n0 q0} ; inverse: E[n1] *= conj(W64^{n1 k2})
#pragma unroll
    // (T2 is symmetric, W64^{ab} = W64^{ba}: index it row-major in the loop variable so that the 8 distinct
    //  addresses of an instruction are consecutive words)
#pragma unroll
    for (int j = 1; j < 8; ++j) a[j] = f2_cmul<INV>(a[j], tab[kF512_T2 + j * 8 + (INV ? k2 : n0)]);
  }
}
template <bool INV>
PC_HD void f512_mid_store(int lane, float2* S, const float2* A, const float2* B) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float2* a = h ? B : A;
    const int pid = lane + 32 * h, k2 = pid >> 3, n0 = pid & 7;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (INV) S[f512_s1(k2, 8 * j + n0)] = a[j];     // E[n1 = j] at [k2][8 n1 + n0]
      else S[f512_s2(k2, j, n0)] = a[j];              // C[q0 = j] at [k2][q0][n0]
    }
  }
}

// even/odd split of one mirror pair: a = Z[k], bm = Z[M-k]  ->  X[k], X[M-k]   (fwd_split of kernels.cuh)
PC_HD void f512_split_pair(float2 a, float2 bm, float2 w, float2* xk, float2* xmk) {
  const float2 b = make_float2(bm.x, -bm.y);
  const float2 E = f2_scale(f2_add(a, b), 0.5f), D = f2_scale(f2_sub(a, b), 0.5f);
  const float2 O = make_float2(D.y, -D.x);              // -i * D
  const float2 wO = f2_cmul<false>(O, w);
  *xk = f2_add(E, wO);
  const float2 t = f2_sub(E, wO);
  *xmk = make_float2(t.x, -t.y);
}

// step 3: DFT8 over n0 for both residues, split, packed spectrum row to global memory
PC_HD void f512_fwd_p3(int lane, const float2* S, const float2* tab, float2* X) {
  const int la = f512_la(lane), lb = f512_lb(lane);
  float2 A[8], B[8];
#pragma unroll
  for (int n0 = 0; n0 < 8; ++n0) {
    A[n0] = S[f512_s2(la & 7, la >> 3, n0)];
    B[n0] = S[f512_s2(lb & 7, lb >> 3, n0)];
  }
  f512_dft8<false>(A);      // A[q1] = Z[la + 64 q1]
  f512_dft8<false>(B);      // B[q1] = Z[lb + 64 q1]
  float2 XA[8], XB[8];
  if (lane != 0) {          // mirror of la + 64 q1 is lb + 64 (7 - q1)
#pragma unroll
    for (int q1 = 0; q1 < 8; ++q1)
      f512_split_pair(A[q1], B[7 - q1], tab[kF512_TS + la + 64 * q1], &XA[q1], &XB[7 - q1]);
  } else {                  // lane 0 owns the two self-mirrored residues 0 and 32
    XA[0] = make_float2(A[0].x + A[0].y, A[0].x - A[0].y);          // (DC, Nyquist)
    XA[4] = make_float2(A[4].x, -A[4].y);                            // k = M/2
#pragma unroll
    for (int q1 = 1; q1 < 4; ++q1) f512_split_pair(A[q1], A[8 - q1], tab[kF512_TS + 64 * q1], &XA[q1], &XA[8 - q1]);
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) f512_split_pair(B[q1], B[7 - q1], tab[kF512_TS + 32 + 64 * q1], &XB[q1], &XB[7 - q1]);
  }
#pragma unroll
  for (int q1 = 0; q1 < 8; ++q1) {
    X[la + 64 * q1] = XA[q1];
    X[lb + 64 * q1] = XB[q1];
  }
}

// ---------------------------------------------------------------------------------------------------------
// inverse: W[k] = Yt[k] + (-1)^k Yp[k]  (frequency-domain overlap-add), un-split, inverse steps, first half of the output
// ---------------------------------------------------------------------------------------------------------
PC_HD void f512_unsplit_pair(float2 a, float2 bm, float2 w, float2* zk, float2* zmk) {
  const float2 b = make_float2(bm.x, -bm.y);
  const float2 E = f2_scale(f2_add(a, b), 0.5f), D = f2_scale(f2_sub(a, b), 0.5f);
  const float2 O = f2_cmul<true>(D, w);                 // conj(w) * D
  *zk = make_float2(E.x - O.y, E.y + O.x);
  *zmk = make_float2(E.x + O.y, O.x - E.y);
}

// first inverse step: loads + merge + un-split + inverse DFT8 over q1 + twiddle -> exchange buffer (layout s2)
PC_HD void f512_inv_p1(int lane, const float2* Yt, const float2* Yp, float2* S, const float2* tab) {
  const int la = f512_la(lane), lb = f512_lb(lane);
  float2 A[8], B[8], PA[8], PB[8];
#pragma unroll
  for (int q1 = 0; q1 < 8; ++q1) {       // all 32 loads first
    A[q1] = PC_LD(Yt + la + 64 * q1); B[q1] = PC_LD(Yt + lb + 64 * q1);
    PA[q1] = PC_LD(Yp + la + 64 * q1); PB[q1] = PC_LD(Yp + lb + 64 * q1);
  }
  const float sg = (lane & 1) ? -1.0f : 1.0f;          // (-1)^k: k has the parity of the lane for both residues
#if defined(__CUDA_ARCH__)
  const float2 sg2 = make_float2(sg, sg);
#pragma unroll
  for (int q1 = 0; q1 < 8; ++q1) { A[q1] = __ffma2_rn(PA[q1], sg2, A[q1]); B[q1] = __ffma2_rn(PB[q1], sg2, B[q1]); }
#else
#pragma unroll
  for (int q1 = 0; q1 < 8; ++q1) {
    A[q1] = make_float2(fmaf(PA[q1].x, sg, A[q1].x), fmaf(PA[q1].y, sg, A[q1].y));
    B[q1] = make_float2(fmaf(PB[q1].x, sg, B[q1].x), fmaf(PB[q1].y, sg, B[q1].y));
  }
#endif
  float2 ZA[8], ZB[8];
  if (lane != 0) {
#pragma unroll
    for (int q1 = 0; q1 < 8; ++q1)
      f512_unsplit_pair(A[q1], B[7 - q1], tab[kF512_TS + la + 64 * q1], &ZA[q1], &ZB[7 - q1]);
  } else {
    // entry 0 packs (DC, Nyquist); the Nyquist index M is even, so its overlap sign is + (ola_merge)
    ZA[0] = make_float2(0.5f * (A[0].x + A[0].y), 0.5f * (A[0].x - A[0].y));
    ZA[4] = make_float2(A[4].x, -A[4].y);
#pragma unroll
    for (int q1 = 1; q1 < 4; ++q1) f512_unsplit_pair(A[q1], A[8 - q1], tab[kF512_TS + 64 * q1], &ZA[q1], &ZA[8 - q1]);
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) f512_unsplit_pair(B[q1], B[7 - q1], tab[kF512_TS + 32 + 64 * q1], &ZB[q1], &ZB[7 - q1]);
  }
  f512_dft8<true>(ZA);      // ZA[n0] = sum_q1 Z[la + 64 q1] W8^{-n0 q1}
  f512_dft8<true>(ZB);
#pragma unroll
  for (int n0 = 0; n0 < 8; ++n0) {
    const float2 da = n0 ? f2_cmul<true>(ZA[n0], tab[kF512_T1 + n0 * 64 + la]) : ZA[0];
    const float2 db = n0 ? f2_cmul<true>(ZB[n0], tab[kF512_T1 + n0 * 64 + lb]) : ZB[0];
    S[f512_s2(la & 7, la >> 3, n0)] = da;
    S[f512_s2(lb & 7, lb >> 3, n0)] = db;
  }
}

// last inverse step: inverse DFT8 over k2 for the two columns; only n2 < 4 (the first B of the 2B output samples)
// FAST: the whole block is inside the destination, no look-ahead rings, linear, 8-byte aligned -> float2 stores
template <bool FAST>
PC_HD void f512_inv_p3(int lane, const float2* S, float scale, const OutSpec& o) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int m = lane + 32 * h;
    float2 e[8];
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) e[k2] = S[f512_s1(k2, m)];
    f512_dft8<true>(e);
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
      const int n = 64 * n2 + m;
      if (FAST) {
        *reinterpret_cast<float2*>(o.dst + o.index0 + 2 * n) = f2_scale(e[n2], scale);
      } else {
        inv_store_sample(e[n2].x, scale, o, 2 * n);
        inv_store_sample(e[n2].y, scale, o, 2 * n + 1);
      }
    }
  }
}

#if defined(__CUDACC__)
// grid (ceil(nblocks / 8) capped, C), block (32, 8): warp = one transform, looping over blocks with stride 8 * gridDim.x
// dynamic smem = (1088 + 8 * 512) float2 = 41472 bytes
__global__ void __launch_bounds__(256, 4) k_fwd_fft512(FwdParams P, const float2* __restrict__ tab512) {
  extern __shared__ float2 pc_smem512[];
  float2* tab = pc_smem512;
  float2* S = pc_smem512 + kF512_TabLen + threadIdx.y * kF512_Xch;
  const int lane = threadIdx.x, tid = threadIdx.y * 32 + lane;
  for (int j = tid; j < kF512_TabLen; j += 256) tab[j] = tab512[j];
  __syncthreads();
  const int c = blockIdx.y;
  const long long nv_total = P.nvalid_c ? (long long)P.nvalid_c[c] : P.nvalid;
  const float* src_c = P.src + (long long)(P.use_cmap ? P.cmap[c] : c) * P.src_cstride;
  // software pipeline: the loads of the warp's NEXT block are issued before the current one is transformed
  auto issue = [&](int blk, float2* a) {
    const long long rem = nv_total - (long long)blk * kF512_M;
    const int nv = rem <= 0 ? 0 : (rem > kF512_M ? kF512_M : (int)rem);
    const float* src = src_c + (long long)blk * kF512_M;
    const bool vec = nv == kF512_M && (reinterpret_cast<size_t>(src) & 7) == 0;
    f512_fwd_load(lane, src, nv, vec, a);
  };
  const int stride = gridDim.x * 8;
  int blk = blockIdx.x * 8 + threadIdx.y;
  float2 nxt[8];
  if (blk < P.nblocks) issue(blk, nxt);
  for (; blk < P.nblocks; blk += stride) {
    float2 cur[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
    if (blk + stride < P.nblocks) issue(blk + stride, nxt);
    f512_fwd_p1(lane, cur, S, tab);
    __syncwarp();
    float2 A[8], B[8];
    f512_mid_load<false>(lane, S, tab, A, B);
    __syncwarp();
    f512_mid_store<false>(lane, S, A, B);
    __syncwarp();
    f512_fwd_p3(lane, S, tab, P.dst + (long long)c * P.dst_cstride + (P.dst_row0 + blk) * (long long)kF512_M);
    __syncwarp();
  }
}

template <bool FAST>
__global__ void __launch_bounds__(256, 3) k_inv_fft512(InvParams P, const float2* __restrict__ tab512) {
  extern __shared__ float2 pc_smem512[];
  float2* tab = pc_smem512;
  float2* S = pc_smem512 + kF512_TabLen + threadIdx.y * kF512_Xch;
  const int lane = threadIdx.x, tid = threadIdx.y * 32 + lane;
  for (int j = tid; j < kF512_TabLen; j += 256) tab[j] = tab512[j];
  __syncthreads();
  const int c = blockIdx.y;
  OutSpec o;
  o.dst = P.dst + (long long)c * P.dst_cstride;
  o.lo = P.lo; o.hi = P.hi; o.mask = P.mask;
  o.n_add = P.n_add;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    o.add[a] = a < P.n_add ? P.add[a] + (long long)c * P.add_cstride[a] : nullptr;
    o.add_mask[a] = P.add_mask[a];
  }
  for (int blk = blockIdx.x * 8 + threadIdx.y; blk < P.nblocks; blk += gridDim.x * 8) {
    o.index0 = P.index0 + (long long)blk * kF512_M;
    o.abs0 = P.abs0 + (long long)blk * kF512_M;
    const float2* Yt = P.Y + (long long)c * P.y_cstride + (P.yrow0 + blk) * P.y_rstride;
    f512_inv_p1(lane, Yt, Yt - P.y_rstride, S, tab);
    __syncwarp();
    float2 A[8], B[8];
    f512_mid_load<true>(lane, S, tab, A, B);
    __syncwarp();
    f512_mid_store<true>(lane, S, A, B);
    __syncwarp();
    f512_inv_p3<FAST>(lane, S, P.scale, o);
    __syncwarp();
  }
}
#else
// CPU emulation (tests/emu): the lanes of a warp become loops, __syncwarp() the loop boundaries
inline void emu_fwd_fft512(int nblocks, int C, const FwdParams& P, const float2* tab) {
  float2 S[kF512_Xch];
  float2 A[32][8], B[32][8];
  for (int c = 0; c < C; ++c) {
    const long long nv_total = P.nvalid_c ? (long long)P.nvalid_c[c] : P.nvalid;
    const float* src_c = P.src + (long long)(P.use_cmap ? P.cmap[c] : c) * P.src_cstride;
    for (int blk = 0; blk < nblocks; ++blk) {
      const long long rem = nv_total - (long long)blk * kF512_M;
      const int nv = rem <= 0 ? 0 : (rem > kF512_M ? kF512_M : (int)rem);
      const float* src = src_c + (long long)blk * kF512_M;
      for (int l = 0; l < 32; ++l) {
        float2 a[8];
        f512_fwd_load(l, src, nv, false, a);
        f512_fwd_p1(l, a, S, tab);
      }
      for (int l = 0; l < 32; ++l) f512_mid_load<false>(l, S, tab, A[l], B[l]);
      for (int l = 0; l < 32; ++l) f512_mid_store<false>(l, S, A[l], B[l]);
      for (int l = 0; l < 32; ++l) f512_fwd_p3(l, S, tab, P.dst + (long long)c * P.dst_cstride + (P.dst_row0 + blk) * (long long)kF512_M);
    }
  }
}

inline void emu_inv_fft512(int nblocks, int C, const InvParams& P, const float2* tab, bool fast) {
  float2 S[kF512_Xch];
  float2 A[32][8], B[32][8];
  for (int c = 0; c < C; ++c) {
    OutSpec o;
    o.dst = P.dst + (long long)c * P.dst_cstride;
    o.lo = P.lo; o.hi = P.hi; o.mask = P.mask;
    o.n_add = P.n_add;
    for (int a = 0; a < 3; ++a) {
      o.add[a] = a < P.n_add ? P.add[a] + (long long)c * P.add_cstride[a] : nullptr;
      o.add_mask[a] = P.add_mask[a];
    }
    for (int blk = 0; blk < nblocks; ++blk) {
      o.index0 = P.index0 + (long long)blk * kF512_M;
      o.abs0 = P.abs0 + (long long)blk * kF512_M;
      const float2* Yt = P.Y + (long long)c * P.y_cstride + (P.yrow0 + blk) * P.y_rstride;
      for (int l = 0; l < 32; ++l) f512_inv_p1(l, Yt, Yt - P.y_rstride, S, tab);
      for (int l = 0; l < 32; ++l) f512_mid_load<true>(l, S, tab, A[l], B[l]);
      for (int l = 0; l < 32; ++l) f512_mid_store<true>(l, S, A[l], B[l]);
      for (int l = 0; l < 32; ++l) {
        if (fast) f512_inv_p3<true>(l, S, P.scale, o); else f512_inv_p3<false>(l, S, P.scale, o);
      }
    }
  }
}
#endif

}  // namespace pc
