// Host-side test of the C++ drop-in classes (include/FFTConvolver.h, TwoStageFFTConvolver.h),
// written like the reference's own self-test reads (libs/FFTConvolver/test/Test.cpp): ramp
// signals, random call lengths, naive convolution as truth — plus a subclass that moves the
// background hooks onto a thread exactly the way REEV-R's Convolver does
// (src/dsp/Convolver.cpp:21-95) to prove that protocol still works.
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

#include "TwoStageFFTConvolver.h"

using fftconvolver::Sample;

static void naive(const std::vector<Sample>& x, const std::vector<Sample>& h, std::vector<double>& y)
{
  y.assign(x.size() + h.size() - 1, 0.0);
  for (size_t n = 0; n < x.size(); ++n)
    for (size_t m = 0; m < h.size(); ++m)
      y[n + m] += static_cast<double>(x[n]) * static_cast<double>(h[m]);
}

// REEV-R style: background thread + manual-reset event + atomic "finished" flag
class ThreadedConvolver : public fftconvolver::TwoStageFFTConvolver
{
public:
  ThreadedConvolver() : _finished(1), _go(false), _done(true), _exit(false), _starts(0), _thread(&ThreadedConvolver::run, this) {}
  ~ThreadedConvolver()
  {
    { std::lock_guard<std::mutex> l(_m); _exit = true; _go = true; }
    _cv.notify_all();
    _thread.join();
  }
  bool isFinished() const { return _finished.load() != 0; }
  int starts() const { return _starts; }
protected:
  void startBackgroundProcessing() override
  {
    _finished.store(0);
    { std::lock_guard<std::mutex> l(_m); _done = false; _go = true; }
    ++_starts;
    _cv.notify_all();
  }
  void waitForBackgroundProcessing() override
  {
    std::unique_lock<std::mutex> l(_m);
    _cv.wait(l, [this] { return _done; });
  }
private:
  void run()
  {
    for (;;)
    {
      std::unique_lock<std::mutex> l(_m);
      _cv.wait(l, [this] { return _go; });
      _go = false;
      if (_exit) return;
      l.unlock();
      doBackgroundProcessing();
      _finished.store(1);
      l.lock();
      _done = true;
      l.unlock();
      _cv.notify_all();
    }
  }
  std::atomic<unsigned> _finished;
  std::mutex _m;
  std::condition_variable _cv;
  bool _go, _done, _exit;
  int _starts;
  std::thread _thread;
};

template <class Conv>
static double drive(Conv& c, const std::vector<Sample>& x, size_t total, size_t bmin, size_t bmax, const std::vector<double>& truth)
{
  std::vector<Sample> in(total, 0.0f), out(total, 0.0f);
  for (size_t i = 0; i < x.size(); ++i) in[i] = x[i];
  size_t pos = 0;
  while (pos < total)
  {
    size_t n = bmin + static_cast<size_t>(rand()) % (1 + bmax - bmin);
    if (n > total - pos) n = total - pos;
    c.process(&in[pos], &out[pos], n);
    pos += n;
  }
  double peak = 0, err = 0;
  for (size_t i = 0; i < total; ++i)
  {
    peak = std::fmax(peak, std::fabs(truth[i]));
    err = std::fmax(err, std::fabs(truth[i] - out[i]));
  }
  return err / peak;
}

int main()
{
  int failures = 0;
  std::vector<Sample> x(20000), h(4321);
  for (size_t i = 0; i < x.size(); ++i) x[i] = 0.25f * std::sin(0.37f * i) + 0.1f * std::cos(1.3f * i * i * 1e-3f);
  for (size_t i = 0; i < h.size(); ++i) h[i] = std::exp(-0.002f * i) * std::cos(0.11f * i);
  std::vector<double> truth;
  naive(x, h, truth);
  const size_t total = truth.size();

  {
    fftconvolver::FFTConvolver c;
    if (c.init(0, &h[0], h.size())) { std::printf("init(0) must fail\n"); ++failures; }
    if (!c.init(100, &h[0], h.size())) { std::printf("init failed: %s\n", c.lastError()); ++failures; }
    const double e = drive(c, x, total, 1, 300, truth);
    std::printf("FFTConvolver           rel-to-peak error %.3g %s\n", e, e < 2e-5 ? "[OK]" : "[FAILED]");
    failures += !(e < 2e-5);
    c.clear();
    std::vector<Sample> z(1000, 0.0f), o(1000, 1.0f);
    c.process(&z[0], &o[0], z.size());
    bool silent = true;
    for (size_t i = 0; i < o.size(); ++i) silent = silent && o[i] == 0.0f;
    std::printf("FFTConvolver clear()   %s\n", silent ? "[OK]" : "[FAILED]");
    failures += !silent;
  }
  {
    fftconvolver::TwoStageFFTConvolver c;
    if (!c.init(64, 512, &h[0], h.size())) { std::printf("init failed: %s\n", c.lastError()); ++failures; }
    const double e = drive(c, x, total, 1, 300, truth);
    std::printf("TwoStageFFTConvolver   rel-to-peak error %.3g %s\n", e, e < 2e-5 ? "[OK]" : "[FAILED]");
    failures += !(e < 2e-5);
  }
  {
    ThreadedConvolver c;
    if (!c.init(64, 512, &h[0], h.size())) { std::printf("init failed: %s\n", c.lastError()); ++failures; }
    const double e = drive(c, x, total, 48, 48, truth);
    const bool hooks = c.starts() == static_cast<int>(total / 512);
    std::printf("threaded subclass      rel-to-peak error %.3g, %d tail hand-offs %s\n", e, c.starts(),
                (e < 2e-5 && hooks) ? "[OK]" : "[FAILED]");
    failures += !(e < 2e-5 && hooks);
    c.reset();
    std::vector<Sample> z(100, 1.0f), o(100, 1.0f);
    c.process(&z[0], &o[0], z.size());
    bool silent = true;
    for (size_t i = 0; i < o.size(); ++i) silent = silent && o[i] == 0.0f;
    std::printf("reset() -> zeros       %s\n", silent ? "[OK]" : "[FAILED]");
    failures += !silent;
  }
  std::printf(failures ? "FAILED (%d)\n" : "ALL OK\n", failures);
  return failures ? 1 : 0;
}
