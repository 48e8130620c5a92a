// irshape.cu — SURVEY 8f-3 ("next" row): the FFT-heavy core of REEV-R's IR shaping on the device.
//
// Impulse::applyDecay (src/dsp/Impulse.cpp:602-648): 4096-point STFT, hop 1024 (Impulse.h:21-22),
// analysis window of Impulse.cpp:65-69, per-bin decay that compounds once per block after the
// early-reflection blocks (:612, :626-633), inverse transform, overlap-add normalised by the summed
// window (:637-648).  On the device every STFT block is independent: the compounded decay of block b
// is lut[k]^(b - skip) in closed form, so all blocks are transformed, scaled and inverse-transformed
// in three batched launches and a fourth kernel gathers the (up to 4) overlapping blocks per sample.
// Reuses the Stockham passes / split functions of kernels.cuh (M = 2048 complex points per 4096 real).
//
// STATUS: written after the round-1 GPU budget was spent — verified on the CPU emulation against the C
// restatement (oracle/partconv_oracle.c::oc_apply_decay, itself unpinned); not yet run on a GPU.
#if defined(PC_EMULATE)
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/b200conv.h"
#include "kernels.cuh"

namespace {

constexpr int kN = 4096;          // STFT size (Impulse.h:21)
constexpr int kM = kN / 2;        // complex points of the real transform
constexpr int kHop = kN / 4;      // Impulse.h:22

// analysis-windowed block straight from global memory: z[n] = (x[2n] w[2n], x[2n+1] w[2n+1])
struct StftIn {
  const float* src; const float* win; int nv;
  PC_HD int prep(int base) const { return base; }
  PC_HD float2 at(int tok, int off) const {
    const int i0 = 2 * (tok + off), i1 = i0 + 1;
    return make_float2(i0 < nv ? src[i0] * win[i0] : 0.0f, i1 < nv ? src[i1] * win[i1] : 0.0f);
  }
};
// last inverse pass: all 2M samples of the block, scaled, to a dense scratch row
struct StftOut {
  float* dst; float scale;
  PC_HD int prep(int base) const { return base; }
  PC_HD void put(int tok, int off, float2 v) const {
    const int n = tok + off;
    dst[2 * n] = v.x * scale;
    dst[2 * n + 1] = v.y * scale;
  }
};

// spectrum row of block b scaled by lut^m (m = b - skip > 0); entry 0 packs (DC, Nyquist): DC is left alone
// (the reference loop starts at k = 1, Impulse.cpp:627) and the Nyquist bin uses lut[M]
PC_HD void decay_scale(float2* row, const double* lut, int m, int k) {
  if (m <= 0) return;
  if (k == 0) { row[0].y *= (float)std::pow(lut[kM], (double)m); return; }
  const float g = (float)std::pow(lut[k], (double)m);
  row[k].x *= g;
  row[k].y *= g;
}

// out[i] = sum over the blocks covering i of scratch[b][i - b*hop], divided by the summed window (:637-648)
PC_HD float stft_gather(const float* scratch, const float* win, long long n, long long nblocks, long long i) {
  float acc = 0.0f, norm = 0.0f;
  const long long b_hi = i / kHop;
  for (long long b = b_hi; b >= 0 && b > b_hi - kN / kHop; --b) {
    if (b >= nblocks) continue;
    const long long off = i - b * kHop;
    if (off < kN) { acc += scratch[b * kN + off]; norm += win[off]; }
  }
  (void)n;
  return norm > 0.0f ? acc / norm : 0.0f;
}

#if !defined(PC_EMULATE)
__global__ void __launch_bounds__(512) k_stft_fwd(const float* x, long long n, const float* win, const float2* tw, float2* spec, long long nblocks) {
  extern __shared__ float2 sm[];
  const long long b = blockIdx.x;
  if (b >= nblocks) return;
  constexpr int NT = pc::fft_threads(kM);
  const int tx = threadIdx.x;
  float2* bufA = sm; float2* bufB = sm + kM;
  const long long start = b * kHop;
  const long long rem = n - start;
  const int nv = rem > kN ? kN : (int)rem;
  constexpr int R0 = pc::pass_radix(kM, 1);
  for (int i = tx; i < kM / R0; i += NT)
    pc::stockham_butterfly<false>(StftIn{x + start, win, nv}, pc::SmemOut{bufA}, tw + pc::tw_pass_offset(kM, 1), kM, 1, R0, i);
  __syncthreads();
  float2* res = pc::fft_mid_passes<false, kM, R0>(bufA, bufB, tw, tx, true);
  float2* row = spec + b * kM;
  for (int k = tx; k <= kM / 2; k += NT) pc::fwd_split(res, row, tw, kM, k);
}

__global__ void k_stft_decay(float2* spec, const double* lut, long long nblocks, int skip) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const long long b = blockIdx.y;
  if (k < kM && b < nblocks) decay_scale(spec + b * kM, lut, (int)(b - skip), k);
}

__global__ void __launch_bounds__(512) k_stft_inv(const float2* spec, const float2* zero_row, const float2* tw, float* scratch, long long nblocks) {
  extern __shared__ float2 sm[];
  const long long b = blockIdx.x;
  if (b >= nblocks) return;
  constexpr int NT = pc::fft_threads(kM);
  const int tx = threadIdx.x;
  float2* bufA = sm; float2* bufB = sm + kM;
  const float2* row = spec + b * kM;
  for (int k = tx; k <= kM / 2; k += NT) pc::inv_pre(row, zero_row, bufA, tw, kM, k, 1, 0);
  __syncthreads();
  float2* in = pc::fft_mid_passes<true, kM, 1>(bufA, bufB, tw, tx, true);
  // last pass (the one that reaches length M) writes all 2M samples of the block
  int p = 1;
  while (p * pc::pass_radix(kM, p) != kM) p *= pc::pass_radix(kM, p);
  const int R = pc::pass_radix(kM, p);
  for (int i = tx; i < kM / R; i += NT)
    pc::stockham_butterfly<true>(pc::SmemIn{in}, StftOut{scratch + b * kN, 1.0f / (float)kM}, tw + pc::tw_pass_offset(kM, p), kM, p, R, i);
}

__global__ void k_stft_gather(const float* scratch, const float* win, float* out, long long n, long long nblocks) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = stft_gather(scratch, win, n, nblocks, i);
}
#endif

// host reference of the per-block pipeline for the emulation build (same phase functions, loops for threads)
#if defined(PC_EMULATE)
void emu_stft(const float* x, long long n, const float* win, const float2* tw, const double* lut, int skip, float* out) {
  const long long nblocks = (n + kHop - 1) / kHop;
  std::vector<float2> spec((size_t)nblocks * kM), bufA(kM), bufB(kM), zero(kM, make_float2(0.f, 0.f));
  std::vector<float> scratch((size_t)nblocks * kN, 0.0f);
  for (long long b = 0; b < nblocks; ++b) {
    const long long start = b * kHop, rem = n - start;
    const int nv = rem > kN ? kN : (int)rem;
    const int R0 = pc::pass_radix(kM, 1);
    for (int i = 0; i < kM / R0; ++i)
      pc::stockham_butterfly<false>(StftIn{x + start, win, nv}, pc::SmemOut{bufA.data()}, tw + pc::tw_pass_offset(kM, 1), kM, 1, R0, i);
    float2* in = bufA.data(); float2* o = bufB.data();
    for (int p = R0; p < kM;) {
      const int R = pc::pass_radix(kM, p);
      for (int i = 0; i < kM / R; ++i) pc::stockham_butterfly<false>(pc::SmemIn{in}, pc::SmemOut{o}, tw + pc::tw_pass_offset(kM, p), kM, p, R, i);
      std::swap(in, o);
      p *= R;
    }
    float2* row = spec.data() + b * kM;
    for (int k = 0; k <= kM / 2; ++k) pc::fwd_split(in, row, tw, kM, k);
    for (int k = 0; k < kM; ++k) decay_scale(row, lut, (int)(b - skip), k);
    for (int k = 0; k <= kM / 2; ++k) pc::inv_pre(row, zero.data(), bufA.data(), tw, kM, k, 1, 0);
    in = bufA.data(); o = bufB.data();
    for (int p = 1; p < kM;) {
      const int R = pc::pass_radix(kM, p);
      const bool last = p * R == kM;
      for (int i = 0; i < kM / R; ++i) {
        if (!last) pc::stockham_butterfly<true>(pc::SmemIn{in}, pc::SmemOut{o}, tw + pc::tw_pass_offset(kM, p), kM, p, R, i);
        else pc::stockham_butterfly<true>(pc::SmemIn{in}, StftOut{scratch.data() + b * kN, 1.0f / (float)kM}, tw + pc::tw_pass_offset(kM, p), kM, p, R, i);
      }
      std::swap(in, o);
      p *= R;
    }
  }
  for (long long i = 0; i < n; ++i) out[i] = stft_gather(scratch.data(), win, n, nblocks, i);
}
#endif

}  // namespace

// window (Impulse.cpp:65-69) and twiddles on the host, in the layout of kernels.cuh
static void stft_tables(std::vector<float>& win, std::vector<float2>& tw) {
  win.resize(kN);
  const float step = 2.0f * 3.14159265358979323846f / (float)kN;
  for (int i = 0; i < kN / 2; ++i) win[i] = 0.42f - 0.50f * std::cos((float)i * step) + 0.08f * std::cos(2.0f * (float)i * step);
  for (int i = kN / 2; i < kN; ++i) win[i] = win[kN - 1 - i];
  tw.resize(pc::tw_table_len(kM));
  for (int k = 0; k <= kM / 2; ++k) {
    const double a = -2.0 * M_PI * (double)k / (2.0 * (double)kM);
    tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
  }
  for (int p = 1; p < kM;) {
    const int R = pc::pass_radix(kM, p);
    const int off = pc::tw_pass_offset(kM, p);
    for (int r = 1; r < R; ++r)
      for (int k = 0; k < p; ++k) {
        const double a = -2.0 * M_PI * (double)r * (double)k / ((double)p * (double)R);
        tw[off + (r - 1) * p + k] = make_float2((float)std::cos(a), (float)std::sin(a));
      }
    p *= R;
  }
}

// The decay-EQ STFT on DEVICE-resident taps, in place, C channels of n taps at dx + c * stride: one set of tables and
// scratch buffers for all channels, everything on stream st (no host synchronisation inside).
struct DecayScratch {
  float* dwin = nullptr; float* dscratch = nullptr; float* dtmp = nullptr;
  float2* dtw = nullptr; float2* dspec = nullptr; float2* dzero = nullptr; double* dlut = nullptr;
  void release() {
    cudaFree(dwin); cudaFree(dscratch); cudaFree(dtmp); cudaFree(dtw); cudaFree(dspec); cudaFree(dzero); cudaFree(dlut);
    *this = DecayScratch();
  }
};

static bool decay_eq_device(float* dx, size_t stride, int C, size_t n, const double* lut, double srate, cudaStream_t st,
                            DecayScratch& sc, std::vector<float>& win, std::vector<float2>& tw) {
  if (n == 0) return true;
  stft_tables(win, tw);
  const int skip = (int)std::ceil(100.0 * srate / (1000.0 * (double)kN));       // EARLY_REFLECTIONS_MS = 100, :612
  const long long nblocks = ((long long)n + kHop - 1) / kHop;
#if defined(PC_EMULATE)
  (void)st; (void)sc;
  std::vector<float> out(n);
  for (int c = 0; c < C; ++c) {
    emu_stft(dx + (size_t)c * stride, (long long)n, win.data(), tw.data(), lut, skip, out.data());
    std::memcpy(dx + (size_t)c * stride, out.data(), n * sizeof(float));
  }
  return true;
#else
  const size_t smem = 2 * kM * sizeof(float2);
  if (!sc.dtmp) {        // scratch and tables once per shaping job, shared by all channels (same n)
    bool ok = cudaMalloc(&sc.dtmp, n * sizeof(float)) == cudaSuccess;
    ok = ok && cudaMalloc(&sc.dwin, kN * sizeof(float)) == cudaSuccess;
    ok = ok && cudaMalloc(&sc.dtw, tw.size() * sizeof(float2)) == cudaSuccess;
    ok = ok && cudaMalloc(&sc.dlut, (kM + 1) * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&sc.dspec, (size_t)nblocks * kM * sizeof(float2)) == cudaSuccess;
    ok = ok && cudaMalloc(&sc.dzero, kM * sizeof(float2)) == cudaSuccess;
    ok = ok && cudaMalloc(&sc.dscratch, (size_t)nblocks * kN * sizeof(float)) == cudaSuccess;
    if (!ok) return false;
    cudaFuncSetAttribute(k_stft_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_stft_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemcpyAsync(sc.dwin, win.data(), kN * sizeof(float), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(sc.dtw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(sc.dlut, lut, (kM + 1) * sizeof(double), cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(sc.dzero, 0, kM * sizeof(float2), st);
  }
  const int NT = pc::fft_threads(kM);
  for (int c = 0; c < C; ++c) {
    float* x = dx + (size_t)c * stride;
    k_stft_fwd<<<(unsigned)nblocks, NT, smem, st>>>(x, (long long)n, sc.dwin, sc.dtw, sc.dspec, nblocks);
    k_stft_decay<<<dim3((kM + 255) / 256, (unsigned)nblocks), 256, 0, st>>>(sc.dspec, sc.dlut, nblocks, skip);
    k_stft_inv<<<(unsigned)nblocks, NT, smem, st>>>(sc.dspec, sc.dzero, sc.dtw, sc.dscratch, nblocks);
    k_stft_gather<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sc.dscratch, sc.dwin, sc.dtmp, (long long)n, nblocks);
    cudaMemcpyAsync(x, sc.dtmp, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
  }
  return cudaGetLastError() == cudaSuccess;
#endif
}

extern "C" int b200conv_ir_decay_eq(int device, float* ir, size_t n, const double* lut, double srate) {
  if (!ir || !lut) return B200CONV_EINVAL;
  if (n == 0) return B200CONV_OK;
  std::vector<float> win; std::vector<float2> tw;
#if defined(PC_EMULATE)
  (void)device;
  DecayScratch sc;
  decay_eq_device(ir, n, 1, n, lut, srate, nullptr, sc, win, tw);
  return B200CONV_OK;
#else
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return B200CONV_ECUDA; }
  float* dx = nullptr;
  cudaStream_t st = nullptr;
  DecayScratch sc;
  bool ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMalloc(&dx, n * sizeof(float)) == cudaSuccess;
  if (ok) {
    cudaMemcpyAsync(dx, ir, n * sizeof(float), cudaMemcpyHostToDevice, st);
    ok = decay_eq_device(dx, n, 1, n, lut, srate, st, sc, win, tw);
    cudaMemcpyAsync(ir, dx, n * sizeof(float), cudaMemcpyDeviceToHost, st);
    ok = ok && cudaStreamSynchronize(st) == cudaSuccess && cudaGetLastError() == cudaSuccess;
  }
  sc.release();
  cudaFree(dx);
  if (st) cudaStreamDestroy(st);
  if (!ok) { cudaGetLastError(); return B200CONV_ECUDA; }
  return B200CONV_OK;
#endif
}

// ---------------------------------------------------------------------------------------------------------
// The device-resident subset of Impulse::recalcImpulse (src/dsp/Impulse.cpp:297-360), in the reference's order:
// auto gain (:313-320, calculateAutoGain :703-720) -> reverse (:322-330) -> trim (:437-470) -> gain (:472-486) ->
// decay EQ (:602-648) -> clip (:488-501) -> attack / decay envelope (:651-680).  Resampling, stretch and the
// parametric EQ (JUCE interpolators / SVF class) stay on the host.
// ---------------------------------------------------------------------------------------------------------
namespace {
// dst[i] = (raw[src index of i] * autogain) * gain, src index = trim start + i, mirrored when reversed
PC_HD float ir_pick(const float* raw, long long n_raw, long long start, int reverse, float ag, float g, long long i) {
  const long long j = start + i;
  const float v = raw[reverse ? n_raw - 1 - j : j];
  return (v * ag) * g;
}
PC_HD float ir_clip_env(float v, int clip, long long i, long long n, long long attack_n, long long decay_n) {
  if (clip) v = v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v);
  if (i < attack_n) v *= (float)i / (float)attack_n;
  if (i >= n - decay_n) {
    const float t = (float)(i - (n - decay_n)) / (float)decay_n;
    v *= 1.0f - (float)std::pow((double)t, 0.5);
  }
  return v;
}

#if !defined(PC_EMULATE)
__global__ void k_ir_energy(const float* l, const float* r, long long n, double* out) {
  __shared__ double red[256];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double a = (double)l[i], b = (double)r[i];
    acc += a * a + b * b;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
  if (threadIdx.x == 0) atomicAdd(out, red[0]);
}
__global__ void k_ir_pick(float* dst, const float* raw, long long n_raw, long long start, long long n, int reverse, float ag, float g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = ir_pick(raw, n_raw, start, reverse, ag, g, i);
}
__global__ void k_ir_clip_env(float* x, long long n, int clip, long long attack_n, long long decay_n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = ir_clip_env(x[i], clip, i, n, attack_n, decay_n);
}
// 1 + index of the last tap with |h| >= 1e-6 (the trim rule of FFTConvolver.cpp:103-106), 0 if none
__global__ void k_ir_last_significant(const float* x, long long n, unsigned long long* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(fabsf(x[i]) < 0.000001f)) atomicMax(out, (unsigned long long)(i + 1));
}
#endif
}  // namespace

// internal (not part of the C ABI): shapes C raw channels (host) into C device buffers dev_out[c] (cudaMalloc'ed here,
// *out_len taps each, caller frees) and reports the post-trim lengths trimmed[c] that FFTConvolver::init would use
extern "C" int pc_ir_shape_to_device(int device, const float* const* raw, int C, size_t n, const b200conv_ir_shape_params* sp,
                                     float** dev_out, size_t* out_len, size_t* trimmed) {
  if (!raw || !sp || !dev_out || !out_len || C < 2 || C > 8) return B200CONV_EINVAL;
  for (int c = 0; c < C; ++c) dev_out[c] = nullptr;
  // applyTrim (:437-445)
  const size_t start = (size_t)(sp->trim_left * (float)n);
  const size_t end = n - (size_t)(sp->trim_right * (float)n);
  if (n == 0 || start >= end || start >= n || end > n) { *out_len = 0; for (int c = 0; c < C; ++c) if (trimmed) trimmed[c] = 0; return B200CONV_OK; }
  const size_t m = end - start;
  *out_len = m;
  const long long attack_n = (long long)(int)(sp->attack * (float)(int)m), decay_n = (long long)(int)(sp->decay * (float)(int)m);
#if defined(PC_EMULATE)
  (void)device;
  double energy = 0.0;
  for (size_t i = 0; i < n; ++i) { const double a = raw[0][i], b = raw[1][i]; energy += a * a + b * b; }
  float ag = 1.0f;
  if (sp->autogain && energy > 0.0) ag = (float)std::min(1.0 / std::sqrt(energy), 1.0);
  std::vector<float> win; std::vector<float2> tw;
  DecayScratch sc;
  for (int c = 0; c < C; ++c) {
    dev_out[c] = (float*)std::malloc(m * sizeof(float));
    for (size_t i = 0; i < m; ++i) dev_out[c][i] = ir_pick(raw[c], (long long)n, (long long)start, sp->reverse, ag, sp->gain, (long long)i);
    if (sp->decay_lut) decay_eq_device(dev_out[c], m, 1, m, sp->decay_lut, sp->srate, nullptr, sc, win, tw);
    for (size_t i = 0; i < m; ++i) dev_out[c][i] = ir_clip_env(dev_out[c][i], sp->clip, (long long)i, (long long)m, attack_n, decay_n);
    if (trimmed) { size_t t = m; while (t > 0 && std::fabs(dev_out[c][t - 1]) < 0.000001f) --t; trimmed[c] = t; }
  }
  return B200CONV_OK;
#else
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return B200CONV_ECUDA; }
  cudaStream_t st = nullptr;
  float* draw = nullptr; double* denergy = nullptr; unsigned long long* dlast = nullptr;
  DecayScratch sc;
  std::vector<float> win; std::vector<float2> tw;
  bool ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMalloc(&draw, (size_t)C * n * sizeof(float)) == cudaSuccess;
  ok = ok && cudaMalloc(&denergy, sizeof(double)) == cudaSuccess;
  ok = ok && cudaMalloc(&dlast, C * sizeof(unsigned long long)) == cudaSuccess;
  for (int c = 0; c < C && ok; ++c) ok = cudaMalloc(&dev_out[c], m * sizeof(float)) == cudaSuccess;
  float ag = 1.0f;
  if (ok) {
    for (int c = 0; c < C; ++c) cudaMemcpyAsync(draw + (size_t)c * n, raw[c], n * sizeof(float), cudaMemcpyHostToDevice, st);   // the ONE upload
    if (sp->autogain) {
      double energy = 0.0;
      cudaMemsetAsync(denergy, 0, sizeof(double), st);
      k_ir_energy<<<296, 256, 0, st>>>(draw, draw + n, (long long)n, denergy);
      cudaMemcpyAsync(&energy, denergy, sizeof(double), cudaMemcpyDeviceToHost, st);
      ok = cudaStreamSynchronize(st) == cudaSuccess;
      if (ok && energy > 0.0) ag = (float)std::min(1.0 / std::sqrt(energy), 1.0);
    }
  }
  if (ok) {
    const unsigned gb = (unsigned)((m + 255) / 256);
    for (int c = 0; c < C; ++c)
      k_ir_pick<<<gb, 256, 0, st>>>(dev_out[c], draw + (size_t)c * n, (long long)n, (long long)start, (long long)m, sp->reverse, ag, sp->gain);
    if (sp->decay_lut)
      for (int c = 0; c < C && ok; ++c)
        ok = decay_eq_device(dev_out[c], m, 1, m, sp->decay_lut, sp->srate, st, sc, win, tw);   // one scratch set for all channels
    cudaMemsetAsync(dlast, 0, C * sizeof(unsigned long long), st);
    for (int c = 0; c < C; ++c) {
      k_ir_clip_env<<<gb, 256, 0, st>>>(dev_out[c], (long long)m, sp->clip, attack_n, decay_n);
      k_ir_last_significant<<<gb, 256, 0, st>>>(dev_out[c], (long long)m, dlast + c);
    }
    unsigned long long last[8] = {};
    cudaMemcpyAsync(last, dlast, C * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
    ok = ok && cudaStreamSynchronize(st) == cudaSuccess && cudaGetLastError() == cudaSuccess;
    if (trimmed) for (int c = 0; c < C; ++c) trimmed[c] = (size_t)last[c];
  }
  sc.release();
  cudaFree(draw); cudaFree(denergy); cudaFree(dlast);
  if (st) cudaStreamDestroy(st);
  if (!ok) {
    for (int c = 0; c < C; ++c) { cudaFree(dev_out[c]); dev_out[c] = nullptr; }
    cudaGetLastError();
    return B200CONV_ECUDA;
  }
  return B200CONV_OK;
#endif
}

extern "C" void pc_ir_shape_free(float** dev_out, int C) {
  for (int c = 0; c < C; ++c) {
#if defined(PC_EMULATE)
    std::free(dev_out[c]);
#else
    cudaFree(dev_out[c]);
#endif
    dev_out[c] = nullptr;
  }
}

// stand-alone: the shaped taps back on the host (tests, waveform display)
extern "C" int b200conv_ir_shape(int device, const float* const* raw, int n_channels, size_t n, const b200conv_ir_shape_params* sp,
                                 float* const* out, size_t* out_len) {
  if (!out || !out_len) return B200CONV_EINVAL;
  float* dev[8] = {};
  size_t m = 0;
  const int rc = pc_ir_shape_to_device(device, raw, n_channels, n, sp, dev, &m, nullptr);
  if (rc != B200CONV_OK) return rc;
  *out_len = m;
  bool ok = true;
  for (int c = 0; c < n_channels && m > 0; ++c) {
#if defined(PC_EMULATE)
    std::memcpy(out[c], dev[c], m * sizeof(float));
#else
    ok = ok && cudaMemcpy(out[c], dev[c], m * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess;
#endif
  }
  pc_ir_shape_free(dev, n_channels);
  return ok ? B200CONV_OK : B200CONV_ECUDA;
}
