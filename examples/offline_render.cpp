// offline_render.cpp — minimal host program on the drop-in surface: renders a stereo convolution reverb
// offline the way a REEV-R-style host would, once call-by-call through the reference's class surface
// (fftconvolver::TwoStageFFTConvolver, one object per channel, src/dsp/StereoConvolver.cpp:22-42) and once
// through ONE multi-channel C-ABI handle in a single batched call, and checks that both agree.
//
//   g++ -O2 -std=c++17 -I include examples/offline_render.cpp -L reevr_b200 -l:libb200conv.so \
//       -Wl,-rpath,$PWD/reevr_b200 -o offline_render && ./offline_render
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "TwoStageFFTConvolver.h"

static std::vector<float> decay_ir(size_t n, unsigned seed)
{
  std::vector<float> h(n);
  unsigned s = seed;
  for (size_t i = 0; i < n; ++i)
  {
    s = s * 1664525u + 1013904223u;
    const float g = (static_cast<float>(s >> 8) / 8388608.0f) - 1.0f;
    h[i] = g * std::exp(-6.9f * static_cast<float>(i) / static_cast<float>(n));
  }
  return h;
}

int main()
{
  const size_t sr = 48000, hostBlock = 480, seconds = 4;
  const size_t n = sr * seconds;
  const std::vector<float> irL = decay_ir(2 * sr, 1), irR = decay_ir(2 * sr, 2);
  std::vector<float> inL(n), inR(n);
  for (size_t i = 0; i < n; ++i)
  {
    inL[i] = 0.3f * std::sin(0.01f * i) * std::exp(-1e-4f * (i % 20000));
    inR[i] = 0.3f * std::cos(0.013f * i) * std::exp(-1e-4f * (i % 17000));
  }

  // (1) the reference's way: two convolver objects, one process() call per host block
  size_t head = 1;
  while (head < hostBlock) head *= 2;
  const size_t tail = head * 2 > 8192 ? head * 2 : 8192;          // StereoConvolver.cpp:11-15
  fftconvolver::TwoStageFFTConvolver cl, cr;
  if (!cl.init(head, tail, irL.data(), irL.size()) || !cr.init(head, tail, irR.data(), irR.size()))
  {
    std::printf("init failed: %s\n", cl.lastError());
    return 2;
  }
  std::vector<float> outL(n), outR(n);
  for (size_t pos = 0; pos < n; pos += hostBlock)
  {
    const size_t k = n - pos < hostBlock ? n - pos : hostBlock;
    cl.process(&inL[pos], &outL[pos], k);
    cr.process(&inR[pos], &outR[pos], k);
  }

  // (2) one stereo handle, the whole file in one batched call
  b200conv_config cfg = {};
  cfg.n_channels = 2;
  cfg.shard_count = 1;
  b200conv_t* h = b200conv_create(&cfg);
  const float* irs[2] = { irL.data(), irR.data() };
  const size_t lens[2] = { irL.size(), irR.size() };
  if (!h || b200conv_init_twostage(h, head, tail, irs, lens) != B200CONV_OK)
  {
    std::printf("b200conv init failed: %s\n", h ? b200conv_last_error(h) : "no handle");
    return 2;
  }
  std::vector<float> bL(n), bR(n);
  const float* in[2] = { inL.data(), inR.data() };
  float* out[2] = { bL.data(), bR.data() };
  if (b200conv_process(h, in, out, n) != B200CONV_OK)
  {
    std::printf("process failed: %s\n", b200conv_last_error(h));
    return 2;
  }
  double peak = 0, err = 0;
  for (size_t i = 0; i < n; ++i)
  {
    peak = std::fmax(peak, std::fmax(std::fabs(outL[i]), std::fabs(outR[i])));
    err = std::fmax(err, std::fmax(std::fabs(outL[i] - bL[i]), std::fabs(outR[i] - bR[i])));
  }
  std::printf("rendered %zu stereo frames; call-by-call vs batched: max diff %.3g of peak %.3g; %llu kernel launches in the batched handle\n",
              n, err / peak, peak, b200conv_launch_count(h));
  b200conv_destroy(h);
  return (err / peak < 1e-5) ? 0 : 1;
}
