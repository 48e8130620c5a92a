// kernels_rt.cuh — K0 k_rt_block: ONE launch per real-time call.
//
// A plugin calls process() with one host block (64..1024 samples) at a time (src/PluginProcessor.cpp:1793 ->
// StereoConvolver::process -> TwoStageFFTConvolver::process -> FFTConvolver::process, FFTConvolver.cpp:155-212).  On the
// multi-kernel path such a call is launch-bound: H2D copy, forward FFT, memset, streaming sweep, inverse FFT, mixdown,
// D2H copy, each a separate stream operation (56 us per call in round 1 for 7 us of work).  This kernel does the
// whole head-stage step of a call that stays inside the open block in one launch of one thread-block CLUSTER:
//
//   cluster = C convolvers x NC CTAs (<= 16 CTAs), 256 threads each
//   A  every CTA of convolver c assembles the open block [samples of earlier calls ; this call's samples ; 0] in shared
//      memory — the new samples are read straight from the caller's pinned host buffer (zero-copy) — and (q = 0)
//      appends them to the open blocks of all stages on the device
//   B  forward real FFT of the open block by the whole CTA (the NC CTAs of a convolver compute it redundantly: 1-2 us,
//      cheaper than a cluster exchange) -> spectrum row Xnew in shared memory; q = 0 also stores it as timeline row `head`
//   C  FDL sweep of the CTA's bin tile [q M/NC, (q+1) M/NC): partition 0 from Xnew in shared memory, partitions >= 1
//      from the timeline (FFTConvolver.cpp:176-187), partition groups reduced through shared memory
//   D  the tile of Y goes into the q = 0 CTA's shared memory through DSMEM (and, if the block completes, to the
//      overlap row of the next block) — cluster barrier
//   E  q = 0: frequency-domain overlap-add with the previous block's spectrum, inverse FFT, look-ahead rings of the
//      tail stages added, the call's samples written to the caller's pinned host buffer (or, with routing, to CTA 0's
//      mix buffer through DSMEM — cluster barrier — CTA 0 applies the mixdown matrix, src/PluginProcessor.cpp:1833-1838)
//
// The host keeps the bookkeeping (fill, head, buffer parity); tail stages whose block completes in the call are
// enqueued on a low-priority stream and consumed one tail period later (TwoStageFFTConvolver.cpp:213-222).
#pragma once

#include "kernels.cuh"

namespace pc {

struct RtParams {
  int M, C, NC, P;
  int fill, len, complete;
  // input: routed input i at in + i * in_stride (pinned host memory, device-accessible, or device memory)
  const float* in; long long in_stride; int in_map[8];
  // stage 0
  float* inbuf0; long long inbuf0_stride;
  const float2* H; long long h_cstride;
  float2* X; long long x_cstride; long long head;
  const float2* Yprev; float2* Ynext; long long y_cstride;
  const float2* tw;
  // open blocks of the later stages (the new samples are appended at later_fill[s])
  int n_later; float* later_inbuf[3]; long long later_stride[3]; int later_fill[3];
  // look-ahead rings added on top of the head output
  int n_add; const float* add[3]; long long add_cstride[3]; long long add_mask[3]; long long abs0;
  // output: n_out mixed channels (mix_on) or C channels at out + o * out_stride
  float* out; long long out_stride;
  int mix_on, n_out; float mix[64];
  // completion word in pinned host memory (nullptr: none): set to done_val once every output sample is written, so
  // that the caller can spin on it instead of paying the driver's stream-synchronise latency
  unsigned int* done_flag; unsigned int done_val;
  // mode 0: the whole step in this launch.  Head stages too large for one cluster (uniform long IRs) split it:
  // mode 1 = FRONT (phases A, B: assemble, forward FFT, timeline row) -> the all-SM TMA sweep (K2t) writes Yt ->
  // mode 2 = BACK (phase E from the global row Yt: overlap-add, inverse FFT, output, overlap row of the next block)
  int mode; const float2* Yt;
};

// shared-memory layout of one CTA (float2 units unless noted)
struct RtSmem {
  int tw, bufA, bufB, xnew, yfull;       // float2 offsets
  int xs, ys, red, mixbuf;               // byte offsets of float / float4 regions
  int bytes;
};
PC_HD RtSmem rt_smem_layout(int M, int C) {
  RtSmem L;
  const int MB = M < 16 ? 16 : M;
  int o = 0;
  L.tw = o; o += (tw_table_len(M) + 15) & ~15;
  L.bufA = o; o += MB;
  L.bufB = o; o += MB;
  L.xnew = o; o += MB;
  L.yfull = o; o += MB;
  int b = o * 8;
  L.xs = b; b += M * 4;
  L.ys = b; b += M * 4;
  L.red = b; b += 256 * 16;
  L.mixbuf = b; b += C * M * 4;
  L.bytes = (b + 15) & ~15;
  return L;
}

// sweep geometry of a CTA: tile of TB = M / NC bins = TB / 2 bin pairs, PG partition groups
PC_HD int rt_pairs(int M, int NC) { return M / NC / 2; }

// first forward pass reads the assembled time block from shared memory
struct RtSmemIn {
  const float* xs; int nv;
  PC_HD int prep(int base) const { return base; }
  PC_HD float2 at(int tok, int off) const {
    const int i0 = 2 * (tok + off), i1 = i0 + 1;
    return make_float2(i0 < nv ? xs[i0] : 0.0f, i1 < nv ? xs[i1] : 0.0f);
  }
};
// last inverse pass writes the scaled time samples of the block to shared memory
struct RtSmemOut {
  float* ys; float scale; int half;
  PC_HD int prep(int base) const { return base; }
  PC_HD void put(int tok, int off, float2 v) const {
    const int n = tok + off;
    if (n >= half) return;
    ys[2 * n] = v.x * scale;
    ys[2 * n + 1] = v.y * scale;
  }
};

// ---- per-thread phase bodies (shared with the CPU emulation) ---------------------------------------------
// A: sample i of the open block
PC_HD void rt_assemble(const RtParams& P, int c, int q, int i, float* xs) {
  float v = 0.0f;
  if (i < P.fill) {
    v = P.inbuf0[(long long)c * P.inbuf0_stride + i];
  } else if (i < P.fill + P.len) {
    v = P.in[(long long)P.in_map[c] * P.in_stride + (i - P.fill)];
    if (q == 0) {
      P.inbuf0[(long long)c * P.inbuf0_stride + i] = v;
      for (int s = 0; s < P.n_later; ++s)
        P.later_inbuf[s][(long long)c * P.later_stride[s] + P.later_fill[s] + (i - P.fill)] = v;
    }
  }
  xs[i] = v;
}

// C: partial sum of one thread: bin pair at k, partitions pg, pg + PG, ...
PC_HD float4c rt_sweep_thread(const RtParams& P, int c, int k, int pg, int PG, const float2* xnew) {
  const float2* Hk = P.H + (long long)c * P.h_cstride + k;
  const float2* Xk = P.X + (long long)c * P.x_cstride + P.head * (long long)P.M + k;
  const bool packed_first = (k == 0);
  const float m = packed_first ? 0.0f : 1.0f;
  float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
  // batches of 8 partitions: 16 independent 16-byte loads in flight per thread before any arithmetic (one SM only
  // reaches its share of the L2 bandwidth with deep memory-level parallelism); the ragged last batch is predicated,
  // partition 0 takes the spectrum of the open block from shared memory
  for (int p = pg; p < P.P; p += 8 * PG) {
    float4c h[8], x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int pp = p + u * PG;
      if (pp < P.P) {
        h[u] = ld_pair(Hk + (long long)pp * P.M);
        if (pp == 0) { x[u].a = xnew[k]; x[u].b = xnew[k + 1]; }
        else x[u] = ld_pair(Xk - (long long)pp * P.M);
      } else {
        h[u].a = h[u].b = x[u].a = x[u].b = make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float re = fmaf(h[u].a.x, x[u].a.x, a0.x);
      re = fmaf(-m * h[u].a.y, x[u].a.y, re);
      const float im = packed_first ? fmaf(h[u].a.y, x[u].a.y, a0.y) : fmaf(h[u].a.y, x[u].a.x, fmaf(h[u].a.x, x[u].a.y, a0.y));
      a0 = make_float2(re, im);
      a1.x = fmaf(-h[u].b.y, x[u].b.y, fmaf(h[u].b.x, x[u].b.x, a1.x));
      a1.y = fmaf(h[u].b.y, x[u].b.x, fmaf(h[u].b.x, x[u].b.y, a1.y));
    }
  }
  float4c r; r.a = a0; r.b = a1;
  return r;
}

// E: sample s of the block after the inverse transform: look-ahead rings added
PC_HD float rt_out_sample(const RtParams& P, int c, const float* ys, int s) {
  float r = ys[s];
  for (int a = 0; a < P.n_add; ++a)
    r += P.add[a][(long long)c * P.add_cstride[a] + ((P.abs0 + s) & P.add_mask[a])];
  return r;
}

#if defined(__CUDACC__)
PC_D void rt_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// A CTA may only be written through DSMEM once it is known to have STARTED (its shared memory exists): every CTA
// arrives on the cluster barrier first thing and waits on it right before its first remote store (split barrier, so
// the wait is free by then).
PC_D void rt_cluster_arrive() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
PC_D void rt_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// generic address of `p` (own shared memory) in CTA `rank` of the cluster
template <class T>
PC_D T* rt_map_rank(T* p, unsigned rank) {
  unsigned long long out;
  asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(reinterpret_cast<unsigned long long>(p)), "r"(rank));
  return reinterpret_cast<T*>(out);
}

// all passes of the forward transform by the whole CTA (256 threads), first pass from the assembled block
template <int M, int PASS>
__device__ __forceinline__ float2* rt_fwd_passes(float2* in, float2* out, const float2* tw, int tid) {
  if constexpr (PASS >= M) {
    return in;
  } else {
    constexpr int R = pass_radix(M, PASS);
    for (int i = tid; i < M / R; i += 256)
      stockham_butterfly<false>(SmemIn{in}, SmemOut{out}, tw + tw_pass_offset(M, PASS), M, PASS, R, i);
    __syncthreads();
    return rt_fwd_passes<M, PASS * R>(out, in, tw, tid);
  }
}
template <int M, int PASS>
__device__ __forceinline__ void rt_inv_passes(float2* in, float2* out, const float2* tw, int tid, float* ys, float scale) {
  constexpr int R = pass_radix(M, PASS);
  if constexpr (PASS * R == M) {       // last pass: scaled samples to shared memory (first half of the transform only)
    for (int i = tid; i < M / R; i += 256)
      stockham_butterfly<true>(SmemIn{in}, RtSmemOut{ys, scale, M / 2}, tw + tw_pass_offset(M, PASS), M, PASS, R, i);
    __syncthreads();
  } else {
    for (int i = tid; i < M / R; i += 256)
      stockham_butterfly<true>(SmemIn{in}, SmemOut{out}, tw + tw_pass_offset(M, PASS), M, PASS, R, i);
    __syncthreads();
    rt_inv_passes<M, PASS * R>(out, in, tw, tid, ys, scale);
  }
}

// launched with cluster dimension (C * NC, 1, 1) = the whole grid; block 256; dynamic smem rt_smem_layout(M, C).bytes
template <int M>
__global__ void __launch_bounds__(256) k_rt_block(RtParams P) {
  extern __shared__ __align__(16) unsigned char pc_rt_smem[];
  const RtSmem L = rt_smem_layout(M, P.C);
  float2* sm2 = reinterpret_cast<float2*>(pc_rt_smem);
  float2* tw = sm2 + L.tw;
  float2* bufA = sm2 + L.bufA;
  float2* bufB = sm2 + L.bufB;
  float2* xnew = sm2 + L.xnew;
  float2* yfull = sm2 + L.yfull;
  float* xs = reinterpret_cast<float*>(pc_rt_smem + L.xs);
  float* ys = reinterpret_cast<float*>(pc_rt_smem + L.ys);
  float4* red = reinterpret_cast<float4*>(pc_rt_smem + L.red);
  float* mixbuf = reinterpret_cast<float*>(pc_rt_smem + L.mixbuf);
  const int tid = threadIdx.x;
  const int rank = blockIdx.x, c = rank / P.NC, q = rank % P.NC;
  const bool remote_stores = (P.mode == 0) || (P.mode == 2 && P.mix_on);     // uniform over the cluster
  if (remote_stores) rt_cluster_arrive();

  // ---- A: twiddles + the open block
  for (int j = tid; j < tw_table_len(M); j += 256) tw[j] = P.tw[j];
  if (P.mode != 2) {
    for (int i = tid; i < M; i += 256) rt_assemble(P, c, q, i, xs);
    __syncthreads();

    // ---- B: forward real FFT -> xnew
    if constexpr (M == 1) {
      if (tid == 0) { bufA[0] = make_float2(xs[0], 0.0f); }
      __syncthreads();
      if (tid == 0) fwd_split(bufA, xnew, tw, M, 0);
    } else {
      constexpr int R0 = pass_radix(M, 1);
      for (int i = tid; i < M / R0; i += 256)
        stockham_butterfly<false>(RtSmemIn{xs, P.fill + P.len}, SmemOut{bufA}, tw + tw_pass_offset(M, 1), M, 1, R0, i);
      __syncthreads();
      float2* res = rt_fwd_passes<M, R0>(bufA, bufB, tw, tid);
      for (int k = tid; k <= M / 2; k += 256) fwd_split(res, xnew, tw, M, k);
    }
    __syncthreads();
    if (q == 0) {
      float2* row = P.X + (long long)c * P.x_cstride + P.head * (long long)M;
      for (int k = tid; k < M; k += 256) row[k] = xnew[k];
    }
    if (P.mode == 1) return;            // FRONT: the sweep is a separate all-SM launch (uniform across the cluster)
  }

  const float2* Yt_src = yfull;
  if (P.mode == 0) {
    // ---- C: sweep of this CTA's bin tile
    const int pairs = rt_pairs(M, P.NC);
    const int PG = 256 / pairs;                 // host guarantees 1 <= pairs <= 256
    const int pi = tid % pairs, pg = tid / pairs;
    const int k = q * (M / P.NC) + 2 * pi;
    if (pg < PG) {
      const float4c r = rt_sweep_thread(P, c, k, pg, PG, xnew);
      red[tid] = make_float4(r.a.x, r.a.y, r.b.x, r.b.y);
    }
    __syncthreads();
    // ---- D: reduce the partition groups, tile -> q = 0 CTA of the convolver (DSMEM), overlap row of the next block
    rt_cluster_wait();                          // every CTA of the cluster has started
    if (tid < pairs) {
      float4 v = red[tid];
      for (int g = 1; g < PG; ++g) {
        const float4 u = red[tid + g * pairs];
        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
      }
      float2* dst = rt_map_rank(yfull, (unsigned)(c * P.NC));
      dst[k] = make_float2(v.x, v.y);
      dst[k + 1] = make_float2(v.z, v.w);
      if (P.complete) {
        float2* yn = P.Ynext + (long long)c * P.y_cstride + k;
        yn[0] = make_float2(v.x, v.y);
        yn[1] = make_float2(v.z, v.w);
      }
    }
    rt_cluster_sync();
  } else {                                      // BACK: the sweep's row is in global memory
    Yt_src = P.Yt + (long long)c * P.y_cstride;
    if (P.complete && q == 0) {
      float2* yn = P.Ynext + (long long)c * P.y_cstride;
      for (int k = tid; k < M; k += 256) yn[k] = Yt_src[k];
    }
    __syncthreads();                            // twiddles staged
  }

  // ---- E: overlap-add in the frequency domain, inverse FFT, output
  if (q == 0) {
    const float2* Yp = P.Yprev + (long long)c * P.y_cstride;
    for (int kk = tid; kk <= M / 2; kk += 256) inv_pre(Yt_src, Yp, bufA, tw, M, kk, 1, 0);
    __syncthreads();
    const float scale = 1.0f / (float)M;
    if constexpr (M == 1) {
      if (tid == 0) { ys[0] = bufA[0].x * scale; }
      __syncthreads();
    } else {
      rt_inv_passes<M, 1>(bufA, bufB, tw, tid, ys, scale);
    }
    if (P.mode == 2 && P.mix_on) rt_cluster_wait();      // BACK mode: first remote store of this launch (all CTAs have q = 0)
    if (P.mix_on) {
      float* mb = rt_map_rank(mixbuf, 0u) + (long long)c * M;
      for (int i = tid; i < P.len; i += 256) mb[i] = rt_out_sample(P, c, ys, P.fill + i);
    } else {
      float* o = P.out + (long long)c * P.out_stride;
      for (int i = tid; i < P.len; i += 256) o[i] = rt_out_sample(P, c, ys, P.fill + i);
    }
  }
  if (P.mix_on) {
    rt_cluster_sync();
    if (rank == 0) {
      for (int j = tid; j < P.n_out * P.len; j += 256) {
        const int o = j / P.len, i = j % P.len;
        float acc = 0.0f;
        for (int cc = 0; cc < P.C; ++cc) {
          const float mm = P.mix[o * P.C + cc];
          if (mm != 0.0f) acc = fmaf(mm, mixbuf[(long long)cc * M + i], acc);
        }
        P.out[(long long)o * P.out_stride + i] = acc;
      }
    }
  }
  if (P.done_flag) {
    // every CTA that wrote output makes its stores visible system-wide, the cluster meets, CTA 0 raises the flag
    __threadfence_system();
    rt_cluster_sync();
    if (rank == 0 && tid == 0) {
      *reinterpret_cast<volatile unsigned int*>(P.done_flag) = P.done_val;
      __threadfence_system();
    }
  }
}
#else
// CPU emulation (tests/emu): the CTAs of the cluster run phase by phase; a DSMEM store is a store into the other
// CTA's arrays
inline void emu_rt_block(const RtParams& P) {
  const int M = P.M, n = P.C * P.NC;
  const int MB = M < 16 ? 16 : M;
  struct Cta { float2 *bufA, *bufB, *xnew, *yfull; float *xs, *ys, *mix; float4* red; };
  Cta* ct = new Cta[n];
  for (int r = 0; r < n; ++r) {
    ct[r].bufA = new float2[MB]; ct[r].bufB = new float2[MB]; ct[r].xnew = new float2[MB]; ct[r].yfull = new float2[MB];
    ct[r].xs = new float[M]; ct[r].ys = new float[M]; ct[r].mix = new float[(size_t)P.C * M]; ct[r].red = new float4[256];
  }
  for (int r = 0; r < n && P.mode != 2; ++r) {
    const int c = r / P.NC, q = r % P.NC;
    Cta& t = ct[r];
    for (int i = 0; i < M; ++i) rt_assemble(P, c, q, i, t.xs);
    float2* res = t.bufA;
    if (M == 1) {
      t.bufA[0] = make_float2(t.xs[0], 0.0f);
    } else {
      const int R0 = pass_radix(M, 1);
      for (int i = 0; i < M / R0; ++i)
        stockham_butterfly<false>(RtSmemIn{t.xs, P.fill + P.len}, SmemOut{t.bufA}, P.tw + tw_pass_offset(M, 1), M, 1, R0, i);
      float2* in = t.bufA; float2* out = t.bufB;
      for (int p = R0; p < M;) {
        const int R = pass_radix(M, p);
        for (int i = 0; i < M / R; ++i) stockham_butterfly<false>(SmemIn{in}, SmemOut{out}, P.tw + tw_pass_offset(M, p), M, p, R, i);
        float2* x = in; in = out; out = x;
        p *= R;
      }
      res = in;
    }
    for (int k = 0; k <= M / 2; ++k) fwd_split(res, t.xnew, P.tw, M, k);
    if (q == 0) {
      float2* row = P.X + (long long)c * P.x_cstride + P.head * (long long)M;
      for (int k = 0; k < M; ++k) row[k] = t.xnew[k];
    }
  }
  // C + D  (q = 0 wrote the timeline row before any CTA reads older rows: rows < head only)
  for (int r = 0; r < n && P.mode == 0; ++r) {
    const int c = r / P.NC, q = r % P.NC;
    Cta& t = ct[r];
    const int pairs = rt_pairs(M, P.NC), PG = 256 / pairs;
    for (int tid = 0; tid < 256; ++tid) {
      const int pi = tid % pairs, pg = tid / pairs, k = q * (M / P.NC) + 2 * pi;
      if (pg < PG) {
        const float4c v = rt_sweep_thread(P, c, k, pg, PG, t.xnew);
        t.red[tid].x = v.a.x; t.red[tid].y = v.a.y; t.red[tid].z = v.b.x; t.red[tid].w = v.b.y;
      }
    }
    for (int tid = 0; tid < pairs; ++tid) {
      float4 v = t.red[tid];
      for (int g = 1; g < PG; ++g) { const float4 u = t.red[tid + g * pairs]; v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
      const int k = q * (M / P.NC) + 2 * tid;
      float2* dst = ct[c * P.NC].yfull;
      dst[k] = make_float2(v.x, v.y); dst[k + 1] = make_float2(v.z, v.w);
      if (P.complete) {
        float2* yn = P.Ynext + (long long)c * P.y_cstride + k;
        yn[0] = make_float2(v.x, v.y); yn[1] = make_float2(v.z, v.w);
      }
    }
  }
  // E
  for (int c = 0; c < P.C && P.mode != 1; ++c) {
    Cta& t = ct[c * P.NC];
    const float2* Yp = P.Yprev + (long long)c * P.y_cstride;
    const float2* Yt_src = t.yfull;
    if (P.mode == 2) {
      Yt_src = P.Yt + (long long)c * P.y_cstride;
      if (P.complete) for (int k = 0; k < M; ++k) P.Ynext[(long long)c * P.y_cstride + k] = Yt_src[k];
    }
    for (int kk = 0; kk <= M / 2; ++kk) inv_pre(Yt_src, Yp, t.bufA, P.tw, M, kk, 1, 0);
    const float scale = 1.0f / (float)M;
    if (M == 1) {
      t.ys[0] = t.bufA[0].x * scale;
    } else {
      float2* in = t.bufA; float2* out = t.bufB;
      for (int p = 1; p < M;) {
        const int R = pass_radix(M, p);
        if (p * R == M) {
          for (int i = 0; i < M / R; ++i)
            stockham_butterfly<true>(SmemIn{in}, RtSmemOut{t.ys, scale, M / 2}, P.tw + tw_pass_offset(M, p), M, p, R, i);
        } else {
          for (int i = 0; i < M / R; ++i) stockham_butterfly<true>(SmemIn{in}, SmemOut{out}, P.tw + tw_pass_offset(M, p), M, p, R, i);
          float2* x = in; in = out; out = x;
        }
        p *= R;
      }
    }
    for (int i = 0; i < P.len; ++i) {
      const float v = rt_out_sample(P, c, t.ys, P.fill + i);
      if (P.mix_on) ct[0].mix[(size_t)c * M + i] = v; else P.out[(long long)c * P.out_stride + i] = v;
    }
  }
  if (P.mix_on && P.mode != 1)
    for (int o = 0; o < P.n_out; ++o)
      for (int i = 0; i < P.len; ++i) {
        float acc = 0.0f;
        for (int cc = 0; cc < P.C; ++cc) {
          const float mm = P.mix[o * P.C + cc];
          if (mm != 0.0f) acc = fmaf(mm, ct[0].mix[(size_t)cc * M + i], acc);
        }
        P.out[(long long)o * P.out_stride + i] = acc;
      }
  if (P.done_flag && P.mode != 1) *P.done_flag = P.done_val;
  for (int r = 0; r < n; ++r) {
    delete[] ct[r].bufA; delete[] ct[r].bufB; delete[] ct[r].xnew; delete[] ct[r].yfull;
    delete[] ct[r].xs; delete[] ct[r].ys; delete[] ct[r].mix; delete[] ct[r].red;
  }
  delete[] ct;
}
#endif

}  // namespace pc
