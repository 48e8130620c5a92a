// kernels_chain.cuh — SURVEY 8f-4 and the remainder of 8f-1: the per-sample work REEV-R does on the host around the
// convolver, moved to the device so that the device boundary sits at the plugin's dry input / final output:
//
//   k_chain_send   send = dry * ysend ; low cut / high cut state-variable filters ; predelay ring
//                  (src/PluginProcessor.cpp:1639-1653 + src/dsp/Filter.cpp:23-68 ; :1766-1790)
//   k_chain_wet    wet L = LL (+ RL), R = RR (+ LR) ; * yrev ; mid/side width ; out = drygain * dry + wetgain * wet
//                  (src/PluginProcessor.cpp:1832-1876)
//
// The filters are recursive (two TPT state-variable sections per 24 dB filter), i.e. sequential in time — on a GPU the
// block of n samples is cut into T chunks, one per thread:
//   pass 1  every thread runs the whole cascade (low cut, high cut: up to 8 state variables) over its chunk from a ZERO
//           state and keeps the final state — the zero-state response end point Z_t
//   powers  the cascade is linear and time-invariant: state' = A state + b x.  Column j of A^L (L = chunk length) is the
//           state after L steps with zero input from the unit state e_j — 8 threads run that once per launch
//   scan    S_{t+1} = A^L S_t + Z_t  gives every chunk's true initial state (one thread, T small steps of an 8x8
//           matrix-vector product); S_0 is the state carried over from the previous call
//   pass 2  every thread re-runs its chunk from its true initial state and writes the filtered samples
// The arithmetic inside a chunk is the reference's (same order, float32); only the split into chunks re-associates
// the recurrence (differences ~1e-7 of peak).  The predelay ring is written after the filters and read `predelay`
// samples back, in the same launch.
#pragma once

#include "kernels.cuh"

namespace pc {

constexpr int kChainStates = 8;      // low cut: 4 (two sections, or 1 for 6 dB) + high cut: 4

// coefficients of one reference Filter (src/dsp/Filter.h:53-70), computed on the host as Filter::init does
struct ChainFilter {
  int on, slope, mode;               // slope 0/1/2 = 6/12/24 dB ; mode 0 = LP, 2 = HP
  float g, k, k2, a1, a2, a3, a12, a22, a32;
};

struct ChainSendParams {
  const float* dry; long long dry_stride;       // 2 channels
  const float* ysend;                           // send envelope, n samples (nullptr: 1)
  float* conv_in; long long conv_stride;        // output: the convolver's input, 2 channels
  float* filt; long long filt_stride;           // scratch: filtered samples of this call, 2 channels
  float* state;                                 // [2][kChainStates] filter state carried between calls
  float* ring; long long ring_stride; long long ring_mask; long long ring_pos;   // predelay ring (power of two)
  int predelay;
  long long n;
  ChainFilter lc, hc;
};

// one sample through one Filter: Filter::eval (src/dsp/Filter.cpp:23-68); s = its 4 state variables (ic1..ic4, 6 dB: s[0])
PC_HD float chain_filter_eval(const ChainFilter& f, float* s, float sample) {
  if (f.slope == 0) {
    const float delta = f.g * (sample - s[0]);
    s[0] += delta;
    return f.mode == 0 ? s[0] : sample - s[0];
  }
  float v3 = sample - s[1];
  float v1 = f.a1 * s[0] + f.a2 * v3;
  float v2 = s[1] + f.a2 * s[0] + f.a3 * v3;
  s[0] = 2.0f * v1 - s[0];
  s[1] = 2.0f * v2 - s[1];
  float out = f.mode == 0 ? v2 : (f.mode == 1 ? v1 : sample - f.k * v1 - v2);
  if (f.slope == 1) return out;
  v3 = out - s[3];
  v1 = f.a12 * s[2] + f.a22 * v3;
  v2 = s[3] + f.a22 * s[2] + f.a32 * v3;
  s[2] = 2.0f * v1 - s[2];
  s[3] = 2.0f * v2 - s[3];
  return f.mode == 0 ? v2 : (f.mode == 1 ? v1 : out - f.k2 * v1 - v2);
}

// the cascade of the send path for one sample (s: 8 state variables)
PC_HD float chain_cascade(const ChainSendParams& P, float* s, float x) {
  if (P.lc.on) x = chain_filter_eval(P.lc, s, x);
  if (P.hc.on) x = chain_filter_eval(P.hc, s + 4, x);
  return x;
}

PC_HD float chain_send_in(const ChainSendParams& P, int ch, long long i) {
  const float v = P.dry[(long long)ch * P.dry_stride + i];
  return P.ysend ? v * P.ysend[i] : v;
}

// chunk [i0, i1) of channel ch from state s (updated); out != nullptr: write the filtered samples
PC_HD void chain_run_chunk(const ChainSendParams& P, int ch, long long i0, long long i1, float* s, float* out) {
  for (long long i = i0; i < i1; ++i) {
    const float y = chain_cascade(P, s, chain_send_in(P, ch, i));
    if (out) out[i] = y;
  }
}

// zero-input response: column j of A^L
PC_HD void chain_power_column(const ChainSendParams& P, long long L, int j, float* col) {
  for (int q = 0; q < kChainStates; ++q) col[q] = (q == j) ? 1.0f : 0.0f;
  for (long long i = 0; i < L; ++i) (void)chain_cascade(P, col, 0.0f);
}

// predelay: conv_in[i] = ring[(pos + i - predelay) & mask] after ring[(pos + i) & mask] = filtered[i]
PC_HD void chain_ring_write(const ChainSendParams& P, int ch, long long i) {
  P.ring[(long long)ch * P.ring_stride + ((P.ring_pos + i) & P.ring_mask)] = P.filt[(long long)ch * P.filt_stride + i];
}
PC_HD void chain_ring_read(const ChainSendParams& P, int ch, long long i) {
  P.conv_in[(long long)ch * P.conv_stride + i] =
      P.ring[(long long)ch * P.ring_stride + ((P.ring_pos + i - P.predelay) & P.ring_mask)];
}

struct ChainWetParams {
  const float* dry; long long dry_stride;       // 2 channels
  const float* conv; long long conv_stride;     // per-convolver outputs: LL, RR[, LR, RL]
  const float* yrev;                            // reverb envelope (nullptr: 1)
  float* out; long long out_stride;             // 2 channels
  long long n;
  int quad_ts;                                  // add RL / LR (quad IR and true stereo enabled)
  float width, drygain, wetgain;
};

PC_HD void chain_wet_sample(const ChainWetParams& P, long long i) {     // src/PluginProcessor.cpp:1832-1876
  float wl = P.conv[i], wr = P.conv[P.conv_stride + i];
  if (P.quad_ts) { wl += P.conv[3 * P.conv_stride + i]; wr += P.conv[2 * P.conv_stride + i]; }
  const float e = P.yrev ? P.yrev[i] : 1.0f;
  const float lin = wl * e, rin = wr * e;
  const float mid = (lin + rin) * 0.5f, side = (lin - rin) * 0.5f;
  const float norm = 1.0f / (1.0f + P.width);
  const float lout = (mid + side * P.width) * norm;
  const float rout = (mid - side * P.width) * norm;
  P.out[i] = P.dry[i] * P.drygain + lout * P.wetgain;
  P.out[P.out_stride + i] = P.dry[P.dry_stride + i] * P.drygain + rout * P.wetgain;
}

#if defined(__CUDACC__)
// grid (2 channels), block T threads (T = 64 for real-time calls, 1024 for batches); static smem
static __global__ void __launch_bounds__(1024) k_chain_send(ChainSendParams P) {
  __shared__ float AL[kChainStates][kChainStates];     // A^L, column j in AL[.][j]
  __shared__ float Z[1024][kChainStates];              // pass 1: zero-state end points ; after the scan: initial states
  const int ch = blockIdx.x, t = threadIdx.x, T = blockDim.x;
  const long long L = (P.n + T - 1) / T;
  const long long i0 = (long long)t * L < P.n ? (long long)t * L : P.n;
  const long long i1 = i0 + L < P.n ? i0 + L : P.n;
  const bool filtered = P.lc.on || P.hc.on;
  float* filt = P.filt + (long long)ch * P.filt_stride;
  if (filtered) {
    float s[kChainStates];
#pragma unroll
    for (int q = 0; q < kChainStates; ++q) s[q] = 0.0f;
    chain_run_chunk(P, ch, i0, i1, s, nullptr);
#pragma unroll
    for (int q = 0; q < kChainStates; ++q) Z[t][q] = s[q];
    if (t < kChainStates) {
      float col[kChainStates];
      chain_power_column(P, L, t, col);
#pragma unroll
      for (int q = 0; q < kChainStates; ++q) AL[q][t] = col[q];
    }
    __syncthreads();
    if (t == 0) {       // S_{t+1} = A^L S_t + Z_t ; Z[t] becomes the initial state of chunk t
      float S[kChainStates];
#pragma unroll
      for (int q = 0; q < kChainStates; ++q) S[q] = P.state[ch * kChainStates + q];
      for (int c = 0; c < T; ++c) {
        float nx[kChainStates];
        const long long c0 = (long long)c * L;
        const bool full = c0 + L <= P.n;               // a ragged / empty last chunk does not advance by A^L
#pragma unroll
        for (int q = 0; q < kChainStates; ++q) {
          float a = Z[c][q];
#pragma unroll
          for (int r = 0; r < kChainStates; ++r) a = fmaf(AL[q][r], S[r], a);
          nx[q] = a;
        }
#pragma unroll
        for (int q = 0; q < kChainStates; ++q) { Z[c][q] = S[q]; if (full) S[q] = nx[q]; }
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kChainStates; ++q) s[q] = Z[t][q];
    chain_run_chunk(P, ch, i0, i1, s, filt);
    // the state after the LAST sample of the call belongs to the thread whose chunk ends at n
    if (i1 == P.n && i0 < P.n) {
#pragma unroll
      for (int q = 0; q < kChainStates; ++q) P.state[ch * kChainStates + q] = s[q];
    }
  } else {
    for (long long i = i0; i < i1; ++i) filt[i] = chain_send_in(P, ch, i);
  }
  __syncthreads();
  for (long long i = t; i < P.n; i += T) chain_ring_write(P, ch, i);
  __syncthreads();
  for (long long i = t; i < P.n; i += T) chain_ring_read(P, ch, i);
}

static __global__ void k_chain_wet(ChainWetParams P) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P.n) chain_wet_sample(P, i);
}
#else
// CPU emulation (tests/emu): same chunking, same two passes
inline void emu_chain_send(const ChainSendParams& P, int T) {
  for (int ch = 0; ch < 2; ++ch) {
    const long long L = (P.n + T - 1) / T;
    float* filt = P.filt + (long long)ch * P.filt_stride;
    if (P.lc.on || P.hc.on) {
      float (*Z)[kChainStates] = new float[T][kChainStates];
      float AL[kChainStates][kChainStates];
      for (int t = 0; t < T; ++t) {
        const long long i0 = (long long)t * L < P.n ? (long long)t * L : P.n;
        const long long i1 = i0 + L < P.n ? i0 + L : P.n;
        float s[kChainStates] = {0, 0, 0, 0, 0, 0, 0, 0};
        chain_run_chunk(P, ch, i0, i1, s, nullptr);
        for (int q = 0; q < kChainStates; ++q) Z[t][q] = s[q];
      }
      for (int j = 0; j < kChainStates; ++j) {
        float col[kChainStates];
        chain_power_column(P, L, j, col);
        for (int q = 0; q < kChainStates; ++q) AL[q][j] = col[q];
      }
      float S[kChainStates];
      for (int q = 0; q < kChainStates; ++q) S[q] = P.state[ch * kChainStates + q];
      for (int c = 0; c < T; ++c) {
        float nx[kChainStates];
        const bool full = (long long)c * L + L <= P.n;
        for (int q = 0; q < kChainStates; ++q) {
          float a = Z[c][q];
          for (int r = 0; r < kChainStates; ++r) a = fmaf(AL[q][r], S[r], a);
          nx[q] = a;
        }
        for (int q = 0; q < kChainStates; ++q) { Z[c][q] = S[q]; if (full) S[q] = nx[q]; }
      }
      for (int t = 0; t < T; ++t) {
        const long long i0 = (long long)t * L < P.n ? (long long)t * L : P.n;
        const long long i1 = i0 + L < P.n ? i0 + L : P.n;
        float s[kChainStates];
        for (int q = 0; q < kChainStates; ++q) s[q] = Z[t][q];
        chain_run_chunk(P, ch, i0, i1, s, filt);
        if (i1 == P.n && i0 < P.n)
          for (int q = 0; q < kChainStates; ++q) P.state[ch * kChainStates + q] = s[q];
      }
      delete[] Z;
    } else {
      for (long long i = 0; i < P.n; ++i) filt[i] = chain_send_in(P, ch, i);
    }
    for (long long i = 0; i < P.n; ++i) chain_ring_write(P, ch, i);
    for (long long i = 0; i < P.n; ++i) chain_ring_read(P, ch, i);
  }
}
inline void emu_chain_wet(const ChainWetParams& P) {
  for (long long i = 0; i < P.n; ++i) chain_wet_sample(P, i);
}
#endif

}  // namespace pc
