// TEST INFRASTRUCTURE — minimal stand-in for the parts of JUCE that REEV-R's src/dsp/Convolver.{h,cpp} and
// StereoConvolver.{h,cpp} use (juce::Thread, juce::WaitableEvent, uint32), so that the reference's UNMODIFIED caller
// sources can be compiled against include/TwoStageFFTConvolver.h without the JUCE tree (GUI dependencies).
// Semantics follow juce_Thread.h / juce_WaitableEvent.h: wait(-1) blocks until notify(); WaitableEvent(true) is a
// manual-reset event.  Written from the JUCE documentation, no JUCE code.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace juce
{
typedef std::uint32_t uint32;

class WaitableEvent
{
public:
  explicit WaitableEvent(bool manualReset = false) : _manual(manualReset), _set(false) {}
  bool wait(double timeoutMs = -1.0) const
  {
    std::unique_lock<std::mutex> l(_m);
    if (timeoutMs < 0) _cv.wait(l, [this] { return _set; });
    else if (!_cv.wait_for(l, std::chrono::duration<double, std::milli>(timeoutMs), [this] { return _set; })) return false;
    if (!_manual) _set = false;
    return true;
  }
  void signal() const { { std::lock_guard<std::mutex> l(_m); _set = true; } _cv.notify_all(); }
  void reset() const { std::lock_guard<std::mutex> l(_m); _set = false; }
private:
  bool _manual;
  mutable bool _set;
  mutable std::mutex _m;
  mutable std::condition_variable _cv;
};

class Thread
{
public:
  enum class Priority { highest = 2, high = 1, normal = 0, low = -1, background = -2 };
  explicit Thread(const std::string& name) : _name(name), _exit(false), _notified(false), _running(false) {}
  virtual ~Thread() { stopThread(-1); }
  virtual void run() = 0;
  bool startThread(Priority = Priority::normal)
  {
    if (_running) return true;
    _exit = false;
    _running = true;
    _t = std::thread([this] { run(); });
    return true;
  }
  void signalThreadShouldExit() { _exit = true; }
  bool threadShouldExit() const { return _exit.load(); }
  void notify() { { std::lock_guard<std::mutex> l(_m); _notified = true; } _cv.notify_all(); }
  bool wait(double timeoutMs) const
  {
    std::unique_lock<std::mutex> l(_m);
    bool ok = true;
    if (timeoutMs < 0) _cv.wait(l, [this] { return _notified; });
    else ok = _cv.wait_for(l, std::chrono::duration<double, std::milli>(timeoutMs), [this] { return _notified; });
    _notified = false;
    return ok;
  }
  bool stopThread(int /*timeoutMs*/)
  {
    if (_running)
    {
      signalThreadShouldExit();
      notify();
      if (_t.joinable()) _t.join();
      _running = false;
    }
    return true;
  }
private:
  std::string _name;
  std::atomic<bool> _exit;
  mutable bool _notified;
  bool _running;
  mutable std::mutex _m;
  mutable std::condition_variable _cv;
  std::thread _t;
};
} // namespace juce

using namespace juce;   // JuceHeader.h does this unless JUCE_DONT_DECLARE_PROJECTINFO-style opt-outs are set
