// TEST INFRASTRUCTURE — the members of REEV-R's Impulse (src/dsp/Impulse.h:37-69) and SVF::EQBand (src/dsp/SVF.h:28)
// that src/dsp/StereoConvolver.{h,cpp} touch; the real class decodes audio files through JUCE.
#pragma once
#include <vector>

namespace SVF { struct EQBand { int mode = 0; float freq = 0, q = 0, gain = 0; }; }

class Impulse
{
public:
  std::vector<float> bufferLL = {};
  std::vector<float> bufferRR = {};
  std::vector<float> bufferLR = {};
  std::vector<float> bufferRL = {};
  bool isQuad = false;
};
