// Drop-in replacements for four plugins of the reference (plus two siblings of the same family), same class names, same public shape, same
// XML keys and files, same stdout banners and imshow windows -- the arithmetic runs on a B200 through
// the C ABI of include/bgsb200.h (libbgsb200.so).  Header-only; link with -lbgsb200.
//
//   FrameDifferenceBGS           replaces package_bgs/FrameDifferenceBGS.{h,cpp}
//   WeightedMovingVarianceBGS    replaces package_bgs/WeightedMovingVarianceBGS.{h,cpp}
//   AdaptiveBackgroundLearning   replaces package_bgs/AdaptiveBackgroundLearning.{h,cpp}
//   MixtureOfGaussianV2BGS       replaces package_bgs/MixtureOfGaussianV2BGS.{h,cpp}
//
// FrameProcessor (FrameProcessor.cpp:40-59,163), USTC_BGS (ustc_src/ustc_bgs.cpp:8-14,94), Demo.cpp:179
// and Demo2.cpp:160 keep compiling and calling `new <Class>` / `bgs->process(in, fg, bg)` unchanged.
// Error convention of the reference is kept: an empty input returns silently with the outputs untouched;
// a failing GPU call throws cv::Exception through CV_Assert (Main.cpp:63-72 catches it).
#pragma once
#include <iostream>
#include <string>
#include <vector>

#include <opencv2/opencv.hpp>

#include "IBGS.h"
#include "bgsb200.h"

namespace bgsb_adapter {

class FanOut;

class PluginBase : public IBGS
{
  friend class FanOut;

protected:
  bgsb_ctx *ctx;
  bool firstTime;
  cv::Mat img_foreground, img_background;   // members of the reference classes; own the outputs
  // set by FanOut::process: the next process() call finds this frame's outputs in img_foreground / img_background
  // (one-shot: consumed by that call, dropped by the next FanOut::process -- capture loops reuse one buffer, so the
  // frame's address says nothing about its identity)
  bool pre_valid;
  bool pre_fg, pre_bg;
  int bg_type;                              // CV_8UC3; CV_8UC1 for the gray model of AdaptiveSelectiveBackgroundLearning

  explicit PluginBase(int algo) : ctx(0), firstTime(true), pre_valid(false), pre_fg(false), pre_bg(false), bg_type(CV_8UC3)
  {
    int rc = bgsb_create(&ctx, algo, 0);
    if (rc != BGSB_OK) std::cerr << "bgsb200: " << bgsb_last_error() << std::endl;
    CV_Assert(rc == BGSB_OK);
    // The arithmetic of the OpenCV this translation unit is compiled against: the reference only
    // builds with 2.4 (legacy blob tracking, CvFileStorage), whose BGR2GRAY constants and fp32 addWeighted differ from
    // 3.x / 4.x in the last bit.  BgsbBlobDetectorCC picks cvFindContours' border behaviour the same way.
#if defined(CV_MAJOR_VERSION) && CV_MAJOR_VERSION >= 3
    set("grayVariant", 0);
#else
    set("grayVariant", 1);
    if (algo == BGSB_ALGO_ADAPTIVE_BG_LEARNING) set("ablBlend", 1);
#endif
  }
  virtual ~PluginBase() { bgsb_destroy(ctx); }

  void set(const char *key, double v) { CV_Assert(bgsb_set_param(ctx, key, v) == BGSB_OK); }

public:
  virtual void configure() = 0;

  // Queued form of process() for a capture loop (VideoCapture.cpp:151-239): returns once the frame's upload, kernel and
  // downloads are enqueued (bgsb_submit), so the next frame's upload overlaps this frame's download.  img_output /
  // img_bgmodel are the caller's Mats (created here if empty; give preallocated page-locked ones -- bgsb_host_alloc --
  // for the copies to overlap); they and img_input must stay untouched until wait().  No imshow on this path.
  void submit(const cv::Mat &img_input, cv::Mat &img_output, cv::Mat &img_bgmodel, bool *fg_valid = 0, bool *bg_valid = 0)
  {
    if (img_input.empty()) return;
    CV_Assert(img_input.type() == CV_8UC3);
    configure();
    img_output.create(img_input.rows, img_input.cols, CV_8UC1);
    img_bgmodel.create(img_input.rows, img_input.cols, bg_type);
    int fv = 0, bv = 0;
    int rc = bgsb_submit(ctx, img_input.data, img_input.cols, img_input.rows, (size_t)img_input.step, img_output.data,
                         (size_t)img_output.step, img_bgmodel.data, (size_t)img_bgmodel.step, &fv, &bv);
    if (rc != BGSB_OK) std::cerr << "bgsb200: " << bgsb_last_error() << std::endl;
    CV_Assert(rc == BGSB_OK);
    if (fg_valid) *fg_valid = fv != 0;
    if (bg_valid) *bg_valid = bv != 0;
    firstTime = false;
  }
  void wait() { CV_Assert(bgsb_wait(ctx) == BGSB_OK); }

protected:

  // Runs one frame; returns which outputs were produced (reference early returns leave them untouched).
  void run(const cv::Mat &img_input, bool &fg_valid, bool &bg_valid, bool want_bg)
  {
    CV_Assert(img_input.type() == CV_8UC3);     // PreProcessor.cpp:56 hands BGR 8UC3 frames to every plugin
    if (pre_valid) {                            // FanOut already ran this frame through this plugin
      fg_valid = pre_fg;
      bg_valid = pre_bg && want_bg;
      pre_valid = false;
      return;
    }
    img_foreground.create(img_input.rows, img_input.cols, CV_8UC1);
    if (want_bg) img_background.create(img_input.rows, img_input.cols, bg_type);
    int fv = 0, bv = 0;
    int rc = bgsb_process(ctx, img_input.data, img_input.cols, img_input.rows, (size_t)img_input.step,
                          img_foreground.data, (size_t)img_foreground.step,
                          want_bg ? img_background.data : 0, want_bg ? (size_t)img_background.step : 0, &fv, &bv);
    if (rc != BGSB_OK) std::cerr << "bgsb200: " << bgsb_last_error() << std::endl;
    CV_Assert(rc == BGSB_OK);
    fg_valid = fv != 0;
    bg_valid = bv != 0;
  }
};

// FrameProcessor::process (FrameProcessor.cpp:169-215) hands the same img_prep to every enabled plugin, each of which
// would upload it again.  With one added line in front of those calls --
//     fanout.process(img_prep);          // FanOut fanout; fanout.add(frameDifference); fanout.add(mixtureOfGaussianV2BGS); ...
// -- the frame crosses PCIe once (bgsb_process_fanout) and the unchanged `plugin->process(img_prep, fg, bg)` calls that
// follow find their outputs ready (each plugin's next process() call consumes them; it must be given the same frame).
// Parameters re-read by a plugin's loadConfig() apply from the next frame on.
class FanOut
{
  std::vector<PluginBase *> plugins;

public:
  void add(IBGS *p)
  {
    PluginBase *b = dynamic_cast<PluginBase *>(p);
    CV_Assert(b != 0);
    plugins.push_back(b);
  }
  void process(const cv::Mat &img_input)
  {
    if (img_input.empty() || plugins.empty()) return;
    CV_Assert(img_input.type() == CV_8UC3);
    const size_t n = plugins.size();
    std::vector<bgsb_ctx *> ctxs(n);
    std::vector<uint8_t *> fg(n), bg(n);
    std::vector<size_t> fgs(n), bgs(n);
    std::vector<int> fv(n), bv(n);
    for (size_t k = 0; k < n; k++) {
      PluginBase *p = plugins[k];
      p->img_foreground.create(img_input.rows, img_input.cols, CV_8UC1);
      p->img_background.create(img_input.rows, img_input.cols, p->bg_type);
      ctxs[k] = p->ctx;
      fg[k] = p->img_foreground.data; fgs[k] = (size_t)p->img_foreground.step;
      bg[k] = p->img_background.data; bgs[k] = (size_t)p->img_background.step;
    }
    int rc = bgsb_process_fanout(&ctxs[0], (int)n, img_input.data, img_input.cols, img_input.rows, (size_t)img_input.step,
                                 &fg[0], &fgs[0], &bg[0], &bgs[0], &fv[0], &bv[0]);
    if (rc != BGSB_OK) std::cerr << "bgsb200: " << bgsb_last_error() << std::endl;
    CV_Assert(rc == BGSB_OK);
    for (size_t k = 0; k < n; k++) {
      plugins[k]->pre_valid = true;             // replaces whatever an earlier fan-out left unconsumed
      plugins[k]->pre_fg = fv[k] != 0;
      plugins[k]->pre_bg = bv[k] != 0;
    }
  }
};

}  // namespace bgsb_adapter

// ---------------------------------------------------------------------------------------------------------------
class FrameDifferenceBGS : public bgsb_adapter::PluginBase
{
private:
  bool enableThreshold;
  int threshold;
  bool showOutput;

public:
  FrameDifferenceBGS() : PluginBase(BGSB_ALGO_FRAME_DIFFERENCE), enableThreshold(true), threshold(15), showOutput(true)
  {
    std::cout << "FrameDifferenceBGS()" << std::endl;
  }
  ~FrameDifferenceBGS() { std::cout << "~FrameDifferenceBGS()" << std::endl; }

  // the configuration preamble of process() (also run by the queued submit())
  void configure()
  {
    loadConfig();
    if (firstTime) saveConfig();
  }

  void process(const cv::Mat &img_input, cv::Mat &img_output, cv::Mat &img_bgmodel)
  {
    (void)img_bgmodel;                       // never written (FrameDifferenceBGS.cpp:29-61)
    if (img_input.empty()) return;
    configure();
    bool fg, bg;
    run(img_input, fg, bg, false);
    if (!fg) return;                         // first frame: only remembered (.cpp:39-43)
    if (showOutput) cv::imshow("Frame Difference", img_foreground);
    img_foreground.copyTo(img_output);
    firstTime = false;
  }

private:
  void saveConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/FrameDifferenceBGS.xml", 0, CV_STORAGE_WRITE);
    cvWriteInt(fs, "enableThreshold", enableThreshold);
    cvWriteInt(fs, "threshold", threshold);
    cvWriteInt(fs, "showOutput", showOutput);
    cvReleaseFileStorage(&fs);
  }
  void loadConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/FrameDifferenceBGS.xml", 0, CV_STORAGE_READ);
    enableThreshold = cvReadIntByName(fs, 0, "enableThreshold", true);
    threshold = cvReadIntByName(fs, 0, "threshold", 15);
    showOutput = cvReadIntByName(fs, 0, "showOutput", true);
    cvReleaseFileStorage(&fs);
    set("enableThreshold", enableThreshold);
    set("threshold", threshold);
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Sibling plugins, same wrapper pattern.
class StaticFrameDifferenceBGS : public bgsb_adapter::PluginBase
{
private:
  bool enableThreshold;
  int threshold;
  bool showOutput;

public:
  StaticFrameDifferenceBGS() : PluginBase(BGSB_ALGO_STATIC_FRAME_DIFFERENCE), enableThreshold(true), threshold(15), showOutput(true)
  {
    std::cout << "StaticFrameDifferenceBGS()" << std::endl;
  }
  ~StaticFrameDifferenceBGS() { std::cout << "~StaticFrameDifferenceBGS()" << std::endl; }

  // the configuration preamble of process() (also run by the queued submit())
  void configure()
  {
    loadConfig();
    if (firstTime) saveConfig();
  }

  void process(const cv::Mat &img_input, cv::Mat &img_output, cv::Mat &img_bgmodel)
  {
    if (img_input.empty()) return;
    configure();
    bool fg, bg;
    run(img_input, fg, bg, true);
    if (showOutput) cv::imshow("Static Frame Difference", img_foreground);
    img_foreground.copyTo(img_output);       // StaticFrameDifferenceBGS.cpp:53-54
    img_background.copyTo(img_bgmodel);
    firstTime = false;
  }

private:
  void saveConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/StaticFrameDifferenceBGS.xml", 0, CV_STORAGE_WRITE);
    cvWriteInt(fs, "enableThreshold", enableThreshold);
    cvWriteInt(fs, "threshold", threshold);
    cvWriteInt(fs, "showOutput", showOutput);
    cvReleaseFileStorage(&fs);
  }
  void loadConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/StaticFrameDifferenceBGS.xml", 0, CV_STORAGE_READ);
    enableThreshold = cvReadIntByName(fs, 0, "enableThreshold", true);
    threshold = cvReadIntByName(fs, 0, "threshold", 15);
    showOutput = cvReadIntByName(fs, 0, "showOutput", true);
    cvReleaseFileStorage(&fs);
    set("enableThreshold", enableThreshold);
    set("threshold", threshold);
  }
};

class WeightedMovingMeanBGS : public bgsb_adapter::PluginBase
{
private:
  bool enableWeight;
  bool enableThreshold;
  int threshold;
  bool showOutput;
  bool showBackground;

public:
  WeightedMovingMeanBGS() : PluginBase(BGSB_ALGO_WEIGHTED_MOVING_MEAN), enableWeight(true), enableThreshold(true),
    threshold(15), showOutput(true), showBackground(false)
  {
    std::cout << "WeightedMovingMeanBGS()" << std::endl;
  }
  ~WeightedMovingMeanBGS() { std::cout << "~WeightedMovingMeanBGS()" << std::endl; }

  // the configuration preamble of process() (also run by the queued submit())
  void configure()
  {
    loadConfig();
    if (firstTime) saveConfig();
  }

  void process(const cv::Mat &img_input, cv::Mat &img_output, cv::Mat &img_bgmodel)
  {
    if (img_input.empty()) return;
    configure();
    bool fg, bg;
    run(img_input, fg, bg, true);
    if (!fg) return;                         // first two frames fill the history (WeightedMovingMeanBGS.cpp:40-51)
    if (showBackground) cv::imshow("W Moving Mean BG Model", img_background);
    if (showOutput) cv::imshow("W Moving Mean FG Mask", img_foreground);
    img_foreground.copyTo(img_output);       // :87-88
    img_background.copyTo(img_bgmodel);
    firstTime = false;
  }

private:
  void saveConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/WeightedMovingMeanBGS.xml", 0, CV_STORAGE_WRITE);
    cvWriteInt(fs, "enableWeight", enableWeight);
    cvWriteInt(fs, "enableThreshold", enableThreshold);
    cvWriteInt(fs, "threshold", threshold);
    cvWriteInt(fs, "showOutput", showOutput);
    cvWriteInt(fs, "showBackground", showBackground);
    cvReleaseFileStorage(&fs);
  }
  void loadConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/WeightedMovingMeanBGS.xml", 0, CV_STORAGE_READ);
    enableWeight = cvReadIntByName(fs, 0, "enableWeight", true);
    enableThreshold = cvReadIntByName(fs, 0, "enableThreshold", true);
    threshold = cvReadIntByName(fs, 0, "threshold", 15);
    showOutput = cvReadIntByName(fs, 0, "showOutput", true);
    showBackground = cvReadIntByName(fs, 0, "showBackground", false);
    cvReleaseFileStorage(&fs);
    set("enableWeight", enableWeight);
    set("enableThreshold", enableThreshold);
    set("threshold", threshold);
  }
};

// ---------------------------------------------------------------------------------------------------------------
class WeightedMovingVarianceBGS : public bgsb_adapter::PluginBase
{
private:
  bool enableWeight;
  bool enableThreshold;
  int threshold;
  bool showOutput;

public:
  WeightedMovingVarianceBGS() : PluginBase(BGSB_ALGO_WEIGHTED_MOVING_VARIANCE), enableWeight(true),
    enableThreshold(true), threshold(15), showOutput(true)
  {
    std::cout << "WeightedMovingVarianceBGS()" << std::endl;
  }
  ~WeightedMovingVarianceBGS() { std::cout << "~WeightedMovingVarianceBGS()" << std::endl; }

  // the configuration preamble of process() (also run by the queued submit())
  void configure()
  {
    loadConfig();
    if (firstTime) saveConfig();
  }

  void process(const cv::Mat &img_input, cv::Mat &img_output, cv::Mat &img_bgmodel)
  {
    (void)img_bgmodel;                       // never written (WeightedMovingVarianceBGS.cpp:30-117)
    if (img_input.empty()) return;
    configure();
    bool fg, bg;
    run(img_input, fg, bg, false);
    if (!fg) return;                         // first two frames fill the history (.cpp:40-51)
    if (showOutput) cv::imshow("W Moving Variance", img_foreground);
    img_foreground.copyTo(img_output);
    firstTime = false;
  }

private:
  void saveConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/WeightedMovingVarianceBGS.xml", 0, CV_STORAGE_WRITE);
    cvWriteInt(fs, "enableWeight", enableWeight);
    cvWriteInt(fs, "enableThreshold", enableThreshold);
    cvWriteInt(fs, "threshold", threshold);
    cvWriteInt(fs, "showOutput", showOutput);
    cvReleaseFileStorage(&fs);
  }
  void loadConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/WeightedMovingVarianceBGS.xml", 0, CV_STORAGE_READ);
    enableWeight = cvReadIntByName(fs, 0, "enableWeight", true);
    enableThreshold = cvReadIntByName(fs, 0, "enableThreshold", true);
    threshold = cvReadIntByName(fs, 0, "threshold", 15);
    showOutput = cvReadIntByName(fs, 0, "showOutput", true);
    cvReleaseFileStorage(&fs);
    set("enableWeight", enableWeight);
    set("enableThreshold", enableThreshold);
    set("threshold", threshold);
  }
};

// ---------------------------------------------------------------------------------------------------------------
class AdaptiveBackgroundLearning : public bgsb_adapter::PluginBase
{
private:
  double alpha;
  long limit;
  bool enableThreshold;
  int threshold;
  bool showForeground;
  bool showBackground;

public:
  AdaptiveBackgroundLearning() : PluginBase(BGSB_ALGO_ADAPTIVE_BG_LEARNING), alpha(0.05), limit(-1),
    enableThreshold(true), threshold(15), showForeground(true), showBackground(true)
  {
    std::cout << "AdaptiveBackgroundLearning()" << std::endl;
  }
  ~AdaptiveBackgroundLearning() { std::cout << "~AdaptiveBackgroundLearning()" << std::endl; }

  // the configuration preamble of process() (also run by the queued submit())
  void configure()
  {
    loadConfig();
    if (firstTime) saveConfig();
  }

  void process(const cv::Mat &img_input, cv::Mat &img_output, cv::Mat &img_bgmodel)
  {
    if (img_input.empty()) return;
    configure();
    bool fg, bg;
    run(img_input, fg, bg, true);
    if (showForeground) cv::imshow("A-Learning FG", img_foreground);
    if (showBackground) cv::imshow("A-Learning BG", img_background);
    img_foreground.copyTo(img_output);       // AdaptiveBackgroundLearning.cpp:79-80
    img_background.copyTo(img_bgmodel);
    firstTime = false;
  }

private:
  void saveConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/AdaptiveBackgroundLearning.xml", 0, CV_STORAGE_WRITE);
    cvWriteReal(fs, "alpha", alpha);
    cvWriteInt(fs, "limit", limit);
    cvWriteInt(fs, "enableThreshold", enableThreshold);
    cvWriteInt(fs, "threshold", threshold);
    cvWriteInt(fs, "showForeground", showForeground);
    cvWriteInt(fs, "showBackground", showBackground);
    cvReleaseFileStorage(&fs);
  }
  void loadConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/AdaptiveBackgroundLearning.xml", 0, CV_STORAGE_READ);
    alpha = cvReadRealByName(fs, 0, "alpha", 0.05);
    limit = cvReadIntByName(fs, 0, "limit", -1);
    enableThreshold = cvReadIntByName(fs, 0, "enableThreshold", true);
    threshold = cvReadIntByName(fs, 0, "threshold", 15);
    showForeground = cvReadIntByName(fs, 0, "showForeground", true);
    showBackground = cvReadIntByName(fs, 0, "showBackground", true);
    cvReleaseFileStorage(&fs);
    set("alpha", alpha);
    set("limit", (double)limit);
    set("enableThreshold", enableThreshold);
    set("threshold", threshold);
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Sibling plugin, USTC_BGS type 7 (ustc_src/ustc_bgs.cpp:15).
class AdaptiveSelectiveBackgroundLearning : public bgsb_adapter::PluginBase
{
private:
  double alphaLearn;
  double alphaDetection;
  long learningFrames;
  int threshold;
  bool showOutput;

public:
  AdaptiveSelectiveBackgroundLearning() : PluginBase(BGSB_ALGO_ADAPTIVE_SELECTIVE_BG_LEARNING), alphaLearn(0.05),
    alphaDetection(0.05), learningFrames(-1), threshold(15), showOutput(true)
  {
    bg_type = CV_8UC1;                       // the model is the gray image (.cpp:103)
    std::cout << "AdaptiveSelectiveBackgroundLearning()" << std::endl;
  }
  ~AdaptiveSelectiveBackgroundLearning() { std::cout << "~AdaptiveSelectiveBackgroundLearning()" << std::endl; }

  // the configuration preamble of process() (also run by the queued submit())
  void configure()
  {
    loadConfig();                            // before saveConfig, as in the reference (.cpp:41-46): the file's
    if (firstTime) saveConfig();             // defaults (90 frames, threshold 25) replace the constructor's
  }

  void process(const cv::Mat &img_input, cv::Mat &img_output, cv::Mat &img_bgmodel)
  {
    if (img_input.empty()) return;
    configure();
    bool fg, bg;
    run(img_input, fg, bg, true);
    if (showOutput) {
      cv::imshow("AS-Learning FG", img_foreground);
      cv::imshow("AS-Learning BG", img_background);
    }
    img_foreground.copyTo(img_output);       // .cpp:102-103
    img_background.copyTo(img_bgmodel);
    firstTime = false;
  }

private:
  void saveConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/AdaptiveSelectiveBackgroundLearning.xml", 0, CV_STORAGE_WRITE);
    cvWriteInt(fs, "learningFrames", learningFrames);
    cvWriteReal(fs, "alphaLearn", alphaLearn);
    cvWriteReal(fs, "alphaDetection", alphaDetection);
    cvWriteInt(fs, "threshold", threshold);
    cvWriteInt(fs, "showOutput", showOutput);
    cvReleaseFileStorage(&fs);
  }
  void loadConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/AdaptiveSelectiveBackgroundLearning.xml", 0, CV_STORAGE_READ);
    learningFrames = cvReadIntByName(fs, 0, "learningFrames", 90);
    alphaLearn = cvReadRealByName(fs, 0, "alphaLearn", 0.05);
    alphaDetection = cvReadRealByName(fs, 0, "alphaDetection", 0.05);
    threshold = cvReadIntByName(fs, 0, "threshold", 25);
    showOutput = cvReadIntByName(fs, 0, "showOutput", true);
    cvReleaseFileStorage(&fs);
    set("learningFrames", (double)learningFrames);
    set("alphaLearn", alphaLearn);
    set("alphaDetection", alphaDetection);
    set("threshold", threshold);
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Sibling plugin, USTC_BGS type 11 (ustc_src/ustc_bgs.cpp:21): package_bgs/dp/DPZivkovicAGMMBGS.{h,cpp}.
class DPZivkovicAGMMBGS : public bgsb_adapter::PluginBase
{
private:
  double threshold;
  double alpha;
  int gaussians;
  bool showOutput;

public:
  DPZivkovicAGMMBGS() : PluginBase(BGSB_ALGO_DP_ZIVKOVIC_AGMM), threshold(25.0f), alpha(0.001f), gaussians(3), showOutput(true)
  {
    std::cout << "DPZivkovicAGMMBGS()" << std::endl;
  }
  ~DPZivkovicAGMMBGS() { std::cout << "~DPZivkovicAGMMBGS()" << std::endl; }

  // the configuration preamble of process() (also run by the queued submit())
  void configure()
  {
    loadConfig();
    if (firstTime) {                         // the reference hands the parameters over once, on the first frame (:58-65)
      saveConfig();
      set("threshold", threshold);
      set("alpha", alpha);
      set("gaussians", gaussians);
    }
  }

  void process(const cv::Mat &img_input, cv::Mat &img_output, cv::Mat &img_bgmodel)
  {
    (void)img_bgmodel;                       // never written (DPZivkovicAGMMBGS.cpp:32-84)
    if (img_input.empty()) return;
    configure();
    bool fg, bg;
    run(img_input, fg, bg, false);
    if (showOutput) cv::imshow("Gaussian Mixture Model (Zivkovic)", img_foreground);
    img_foreground.copyTo(img_output);       // :78
    firstTime = false;
  }

private:
  void saveConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/DPZivkovicAGMMBGS.xml", 0, CV_STORAGE_WRITE);
    cvWriteReal(fs, "threshold", threshold);
    cvWriteReal(fs, "alpha", alpha);
    cvWriteInt(fs, "gaussians", gaussians);
    cvWriteInt(fs, "showOutput", showOutput);
    cvReleaseFileStorage(&fs);
  }
  void loadConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/DPZivkovicAGMMBGS.xml", 0, CV_STORAGE_READ);
    threshold = cvReadRealByName(fs, 0, "threshold", 25.0f);
    alpha = cvReadRealByName(fs, 0, "alpha", 0.001f);
    gaussians = cvReadIntByName(fs, 0, "gaussians", 3);
    showOutput = cvReadIntByName(fs, 0, "showOutput", true);
    cvReleaseFileStorage(&fs);
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Sibling plugins of the DP package, USTC_BGS types 9 / 12 / 13 (ustc_src/ustc_bgs.cpp:19,22,23):
// package_bgs/dp/{DPAdaptiveMedianBGS,DPMeanBGS,DPWrenGABGS}.{h,cpp}.  The three wrappers have one shape -- threshold, a
// second parameter, learningFrames, showOutput; handed to the model once, on the first frame; the high-threshold mask
// is the output, img_bgmodel is never written -- so they share one base; what differs is the XML file, the window title
// and which of the two parameters are integers.
namespace bgsb_adapter {
class DPSimplePlugin : public PluginBase
{
protected:
  const char *xml, *title, *key2;
  bool thr_int, p2_int;
  double threshold, p2, d_threshold, d_p2;
  int learningFrames;
  bool showOutput;

  DPSimplePlugin(int algo, const char *xml_, const char *title_, const char *key2_, bool thr_int_, bool p2_int_, double thr, double second)
    : PluginBase(algo), xml(xml_), title(title_), key2(key2_), thr_int(thr_int_), p2_int(p2_int_), threshold(thr), p2(second),
      d_threshold(thr), d_p2(second), learningFrames(30), showOutput(true) {}

public:
  void configure()
  {
    loadConfig();
    if (firstTime) {                         // e.g. DPMeanBGS.cpp:34-35, :56-61
      saveConfig();
      set("threshold", threshold);
      set(key2, p2);
      set("learningFrames", learningFrames);
    }
  }

  void process(const cv::Mat &img_input, cv::Mat &img_output, cv::Mat &img_bgmodel)
  {
    (void)img_bgmodel;                       // never written (e.g. DPMeanBGS.cpp:28-84)
    if (img_input.empty()) return;
    configure();
    bool fg, bg;
    run(img_input, fg, bg, false);
    if (showOutput) cv::imshow(title, img_foreground);
    img_foreground.copyTo(img_output);
    firstTime = false;
  }

private:
  void saveConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage(xml, 0, CV_STORAGE_WRITE);
    if (thr_int) cvWriteInt(fs, "threshold", (int)threshold); else cvWriteReal(fs, "threshold", threshold);
    if (p2_int) cvWriteInt(fs, key2, (int)p2); else cvWriteReal(fs, key2, p2);
    cvWriteInt(fs, "learningFrames", learningFrames);
    cvWriteInt(fs, "showOutput", showOutput);
    cvReleaseFileStorage(&fs);
  }
  void loadConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage(xml, 0, CV_STORAGE_READ);
    threshold = thr_int ? (double)cvReadIntByName(fs, 0, "threshold", (int)d_threshold) : cvReadRealByName(fs, 0, "threshold", d_threshold);
    p2 = p2_int ? (double)cvReadIntByName(fs, 0, key2, (int)d_p2) : cvReadRealByName(fs, 0, key2, d_p2);
    learningFrames = cvReadIntByName(fs, 0, "learningFrames", 30);
    showOutput = cvReadIntByName(fs, 0, "showOutput", true);
    cvReleaseFileStorage(&fs);
  }
};
}  // namespace bgsb_adapter

class DPAdaptiveMedianBGS : public bgsb_adapter::DPSimplePlugin
{
public:
  DPAdaptiveMedianBGS() : DPSimplePlugin(BGSB_ALGO_DP_ADAPTIVE_MEDIAN, "./config/DPAdaptiveMedianBGS.xml", "Adaptive Median (McFarlane&Schofield)",
                                         "samplingRate", true, true, 40, 7)
  {
    std::cout << "DPAdaptiveMedianBGS()" << std::endl;
  }
  ~DPAdaptiveMedianBGS() { std::cout << "~DPAdaptiveMedianBGS()" << std::endl; }
};

class DPMeanBGS : public bgsb_adapter::DPSimplePlugin
{
public:
  DPMeanBGS() : DPSimplePlugin(BGSB_ALGO_DP_MEAN, "./config/DPMeanBGS.xml", "Temporal Mean (Donovan Parks)", "alpha", true, false, 2700, 1e-6f)
  {
    std::cout << "DPMeanBGS()" << std::endl;
  }
  ~DPMeanBGS() { std::cout << "~DPMeanBGS()" << std::endl; }
};

class DPWrenGABGS : public bgsb_adapter::DPSimplePlugin
{
public:
  DPWrenGABGS() : DPSimplePlugin(BGSB_ALGO_DP_WREN_GA, "./config/DPWrenGABGS.xml", "Gaussian Average (Wren)", "alpha", false, false, 12.25f, 0.005f)
  {
    std::cout << "DPWrenGABGS()" << std::endl;
  }
  ~DPWrenGABGS() { std::cout << "~DPWrenGABGS()" << std::endl; }
};

// USTC_BGS type 14 (ustc_src/ustc_bgs.cpp:24): package_bgs/dp/DPPratiMediodBGS.{h,cpp}.
class DPPratiMediodBGS : public bgsb_adapter::PluginBase
{
private:
  int threshold;
  int samplingRate;
  int historySize;
  int weight;
  bool showOutput;

public:
  DPPratiMediodBGS() : PluginBase(BGSB_ALGO_DP_PRATI_MEDIOD), threshold(30), samplingRate(5), historySize(16), weight(5), showOutput(true)
  {
    std::cout << "DPPratiMediodBGS()" << std::endl;
  }
  ~DPPratiMediodBGS() { std::cout << "~DPPratiMediodBGS()" << std::endl; }

  void configure()
  {
    loadConfig();
    if (firstTime) {                         // handed to the model once, on the first frame (:57-62)
      saveConfig();
      set("threshold", threshold);
      set("samplingRate", samplingRate);
      set("historySize", historySize);
      set("weight", weight);
    }
  }

  void process(const cv::Mat &img_input, cv::Mat &img_output, cv::Mat &img_bgmodel)
  {
    (void)img_bgmodel;                       // never written (DPPratiMediodBGS.cpp:28-88)
    if (img_input.empty()) return;
    configure();
    bool fg, bg;
    run(img_input, fg, bg, false);
    if (showOutput) cv::imshow("Temporal Median (Cucchiara&Calderara)", img_foreground);
    img_foreground.copyTo(img_output);       // :75
    firstTime = false;
  }

private:
  void saveConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/DPPratiMediodBGS.xml", 0, CV_STORAGE_WRITE);
    cvWriteInt(fs, "threshold", threshold);
    cvWriteInt(fs, "samplingRate", samplingRate);
    cvWriteInt(fs, "historySize", historySize);
    cvWriteInt(fs, "weight", weight);
    cvWriteInt(fs, "showOutput", showOutput);
    cvReleaseFileStorage(&fs);
  }
  void loadConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/DPPratiMediodBGS.xml", 0, CV_STORAGE_READ);
    threshold = cvReadIntByName(fs, 0, "threshold", 30);
    samplingRate = cvReadIntByName(fs, 0, "samplingRate", 5);
    historySize = cvReadIntByName(fs, 0, "historySize", 16);
    weight = cvReadIntByName(fs, 0, "weight", 5);
    showOutput = cvReadIntByName(fs, 0, "showOutput", true);
    cvReleaseFileStorage(&fs);
  }
};

// USTC_BGS type 35 (ustc_src/ustc_bgs.cpp:62): package_bgs/bl/SigmaDeltaBGS.{h,cpp} (+ sdLaMa091.{h,cpp}).
class SigmaDeltaBGS : public bgsb_adapter::PluginBase
{
private:
  int ampFactor;
  int minVar;
  int maxVar;
  bool showOutput;

public:
  SigmaDeltaBGS() : PluginBase(BGSB_ALGO_SIGMA_DELTA), ampFactor(1), minVar(15), maxVar(255), showOutput(true)
  {
    std::cout << "SigmaDeltaBGS()" << std::endl;
  }
  ~SigmaDeltaBGS() { std::cout << "~SigmaDeltaBGS()" << std::endl; }

  void configure()
  {
    loadConfig();                            // applies the parameters, every frame (:26, :74)
    if (firstTime) saveConfig();
  }

  void process(const cv::Mat &img_input, cv::Mat &img_output, cv::Mat &img_bgmodel)
  {
    (void)img_bgmodel;                       // never written (SigmaDeltaBGS.cpp:16-50)
    if (img_input.empty()) return;
    configure();
    bool fg, bg;
    run(img_input, fg, bg, false);
    firstTime = false;
    if (!fg) return;                         // first frame: the model is initialised, img_output stays untouched (:28-33)
    if (showOutput) cv::imshow("Sigma-Delta", img_foreground);
    img_foreground.copyTo(img_output);
  }

private:
  void saveConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/SigmaDeltaBGS.xml", 0, CV_STORAGE_WRITE);
    cvWriteInt(fs, "ampFactor", ampFactor);
    cvWriteInt(fs, "minVar", minVar);
    cvWriteInt(fs, "maxVar", maxVar);
    cvWriteInt(fs, "showOutput", showOutput);
    cvReleaseFileStorage(&fs);
  }
  void loadConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/SigmaDeltaBGS.xml", 0, CV_STORAGE_READ);
    ampFactor = cvReadIntByName(fs, 0, "ampFactor", 1);
    minVar = cvReadIntByName(fs, 0, "minVar", 15);
    maxVar = cvReadIntByName(fs, 0, "maxVar", 255);
    showOutput = cvReadIntByName(fs, 0, "showOutput", true);
    set("ampFactor", ampFactor);
    set("minVar", minVar);
    set("maxVar", maxVar);
    cvReleaseFileStorage(&fs);
  }
};

// ---------------------------------------------------------------------------------------------------------------
class MixtureOfGaussianV2BGS : public bgsb_adapter::PluginBase
{
private:
  double alpha;
  bool enableThreshold;
  int threshold;
  bool showOutput;

public:
  MixtureOfGaussianV2BGS() : PluginBase(BGSB_ALGO_MOG2), alpha(0.05), enableThreshold(true), threshold(15), showOutput(true)
  {
    std::cout << "MixtureOfGaussianV2BGS()" << std::endl;
  }
  ~MixtureOfGaussianV2BGS() { std::cout << "~MixtureOfGaussianV2BGS()" << std::endl; }

  // the configuration preamble of process() (also run by the queued submit())
  void configure()
  {
    loadConfig();
    if (firstTime) saveConfig();
  }

  void process(const cv::Mat &img_input, cv::Mat &img_output, cv::Mat &img_bgmodel)
  {
    if (img_input.empty()) return;
    configure();
    bool fg, bg;
    run(img_input, fg, bg, true);            // mog(in, fg, alpha) + getBackgroundImage + threshold, one kernel
    if (showOutput)
    {
      cv::imshow("GMM (Zivkovic&Heijden)", img_foreground);
      cv::imshow("GMM BKG (Zivkovic&Heijden)", img_background);
    }
    img_foreground.copyTo(img_output);       // MixtureOfGaussianV2BGS.cpp:70-71
    img_background.copyTo(img_bgmodel);
    firstTime = false;
  }

private:
  void saveConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/MixtureOfGaussianV2BGS.xml", 0, CV_STORAGE_WRITE);
    cvWriteReal(fs, "alpha", alpha);
    cvWriteInt(fs, "enableThreshold", enableThreshold);
    cvWriteInt(fs, "threshold", threshold);
    cvWriteInt(fs, "showOutput", showOutput);
    cvReleaseFileStorage(&fs);
  }
  void loadConfig()
  {
    CvFileStorage *fs = cvOpenFileStorage("./config/MixtureOfGaussianV2BGS.xml", 0, CV_STORAGE_READ);
    alpha = cvReadRealByName(fs, 0, "alpha", 0.05);
    enableThreshold = cvReadIntByName(fs, 0, "enableThreshold", true);
    threshold = cvReadIntByName(fs, 0, "threshold", 15);
    showOutput = cvReadIntByName(fs, 0, "showOutput", true);
    cvReleaseFileStorage(&fs);
    set("alpha", alpha);
    set("enableThreshold", enableThreshold);
    set("threshold", threshold);
  }
};
