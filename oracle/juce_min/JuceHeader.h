// TEST INFRASTRUCTURE ONLY — the two JUCE names src/dsp/Filter.{h,cpp} of the reference use (juce::jlimit,
// juce::MathConstants<float>::pi), so that the UNMODIFIED Filter.cpp can be compiled into oracle/_ref without the
// JUCE tree.  Semantics from the JUCE documentation (jlimit(lo, hi, v) clamps v to [lo, hi]); no JUCE code.
#pragma once
#include <algorithm>
#include <cstddef>
#include <stdexcept>   // JUCE pulls the standard headers in; the reference relies on that (src/dsp/Utils.h:58)
#include <vector>
namespace juce
{
template <typename T>
struct MathConstants { static constexpr T pi = static_cast<T>(3.141592653589793238L); };
template <typename T>
inline T jlimit(T lo, T hi, T v) { return v < lo ? lo : (hi < v ? hi : v); }
} // namespace juce
using namespace juce;
