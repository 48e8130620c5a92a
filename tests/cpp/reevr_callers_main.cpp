// TEST INFRASTRUCTURE — drives REEV-R's own (unmodified) StereoConvolver / Convolver sources, compiled against the
// drop-in headers of include/, the way PluginProcessor does (src/PluginProcessor.cpp:613-614 prepare, :638 loadImpulse,
// :1793 process, :1762 clear): usage  reevr_callers <in.f32> <ir.f32> <out.f32> <block> <taps> <quad 0|1>
// in.f32: L then R samples; ir.f32: LL, RR[, LR, RL] taps; out.f32: bufferLL, bufferRR[, bufferLR, bufferRL] per block.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "StereoConvolver.h"

static std::vector<float> slurp(const char* path)
{
  std::vector<float> v;
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::perror(path); std::exit(2); }
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f) / (long)sizeof(float);
  std::fseek(f, 0, SEEK_SET);
  v.resize((size_t)n);
  if (std::fread(v.data(), sizeof(float), (size_t)n, f) != (size_t)n) std::exit(2);
  std::fclose(f);
  return v;
}

int main(int argc, char** argv)
{
  if (argc != 7) return 2;
  const std::vector<float> in = slurp(argv[1]), ir = slurp(argv[2]);
  const int block = std::atoi(argv[4]);
  const size_t taps = (size_t)std::atol(argv[5]);
  const bool quad = std::atoi(argv[6]) != 0;
  const size_t n = in.size() / 2;
  Impulse imp;
  imp.isQuad = quad;
  imp.bufferLL.assign(ir.begin(), ir.begin() + taps);
  imp.bufferRR.assign(ir.begin() + taps, ir.begin() + 2 * taps);
  if (quad)
  {
    imp.bufferLR.assign(ir.begin() + 2 * taps, ir.begin() + 3 * taps);
    imp.bufferRL.assign(ir.begin() + 3 * taps, ir.begin() + 4 * taps);
  }
  StereoConvolver sc;
  sc.prepare(block);
  sc.loadImpulse(imp);
  FILE* out = std::fopen(argv[3], "wb");
  size_t pos = 0, calls = 0;
  while (pos < n)
  {
    const size_t k = (n - pos < (size_t)block) ? n - pos : (size_t)block;
    sc.process(in.data() + pos, in.data() + n + pos, k);
    std::fwrite(sc.bufferLL.data(), sizeof(float), k, out);
    std::fwrite(sc.bufferRR.data(), sizeof(float), k, out);
    if (quad)
    {
      std::fwrite(sc.bufferLR.data(), sizeof(float), k, out);
      std::fwrite(sc.bufferRL.data(), sizeof(float), k, out);
    }
    pos += k;
    if (++calls == 40) sc.clear();      // "clear tails" mid-stream (src/PluginProcessor.cpp:1762)
  }
  std::fclose(out);
  std::printf("finishedLoading=%d calls=%zu\n", sc.finishedLoading() ? 1 : 0, calls);
  sc.reset();
  return 0;
}
