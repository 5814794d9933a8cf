// TEST INFRASTRUCTURE.  C entry points around the REFERENCE's own ZivkovicAGMM class, compiled together with
// /root/reference/package_bgs/dp/{ZivkovicAGMM,Image}.cpp into oracle/_ref/libdp_ref.so (oracle/Makefile: `make ref`).
// The call sequence is DPZivkovicAGMMBGS::process's (package_bgs/dp/DPZivkovicAGMMBGS.cpp:32-84): Initalize +
// InitModel on the first frame, then Subtract / Update per frame; the plugin's output is the HIGH-threshold mask.
#include <cstring>

#include "ZivkovicAGMM.h"
#include "AdaptiveMedianBGS.h"
#include "MeanBGS.h"
#include "WrenGA.h"
#include "PratiMediodBGS.h"

using namespace Algorithms::BackgroundSubtraction;

struct dpz_ref {
    ZivkovicParams params;
    ZivkovicAGMM bgs;
    RgbImage frame;
    BwImage low, high;
    int w, h, frame_number;
};

extern "C" {

__attribute__((visibility("default"))) dpz_ref *dpz_ref_create(int w, int h, double threshold, double alpha, int gaussians)
{
    dpz_ref *r = new dpz_ref;
    r->w = w; r->h = h; r->frame_number = 0;
    r->frame = cvCreateImage(cvSize(w, h), IPL_DEPTH_8U, 3);
    r->low = cvCreateImage(cvSize(w, h), IPL_DEPTH_8U, 1);
    r->high = cvCreateImage(cvSize(w, h), IPL_DEPTH_8U, 1);
    r->params.SetFrameSize(w, h);                         // DPZivkovicAGMMBGS.cpp:58-62
    r->params.LowThreshold() = threshold;
    r->params.HighThreshold() = 2 * r->params.LowThreshold();
    r->params.Alpha() = alpha;
    r->params.MaxModes() = gaussians;
    r->bgs.Initalize(r->params);                          // :64
    r->bgs.InitModel(r->frame);                           // :65
    return r;
}

// bgr: h rows of w*3 bytes; fg: h*w bytes (the high-threshold mask, img_output of the plugin)
__attribute__((visibility("default"))) void dpz_ref_process(dpz_ref *r, const unsigned char *bgr, unsigned char *fg)
{
    IplImage *f = r->frame.Ptr();
    for (int y = 0; y < r->h; y++) std::memcpy(f->imageData + (size_t)y * f->widthStep, bgr + (size_t)y * r->w * 3, (size_t)r->w * 3);
    r->bgs.Subtract(r->frame_number, r->frame, r->low, r->high);     // :68
    r->low.Clear();                                                  // :69
    r->bgs.Update(r->frame_number, r->frame, r->low);                // :70
    IplImage *m = r->high.Ptr();
    for (int y = 0; y < r->h; y++) std::memcpy(fg + (size_t)y * r->w, m->imageData + (size_t)y * m->widthStep, (size_t)r->w);
    r->frame_number++;                                               // :83
}

__attribute__((visibility("default"))) void dpz_ref_destroy(dpz_ref *r) { delete r; }


}   // extern "C"

// ---- DPAdaptiveMedianBGS / DPMeanBGS / DPWrenGABGS (USTC_BGS types 9, 12, 13): the reference's own classes, driven as
// their wrappers' process() does (package_bgs/dp/DPAdaptiveMedianBGS.cpp:28-82, DPMeanBGS.cpp:28-84, DPWrenGABGS.cpp:28-84):
// Initalize + InitModel(first frame) on the first call, then Subtract / Clear the low mask / Update; output = high mask.
template <class BGS, class PARAMS>
struct dp_simple_ref {
    PARAMS params;
    BGS bgs;
    RgbImage frame;
    BwImage low, high;
    int w, h, frame_number;
    bool first;
};

template <class R>
static R *simple_create(int w, int h)
{
    R *r = new R;
    r->w = w; r->h = h; r->frame_number = 0; r->first = true;
    r->frame = cvCreateImage(cvSize(w, h), IPL_DEPTH_8U, 3);
    r->low = cvCreateImage(cvSize(w, h), IPL_DEPTH_8U, 1);
    r->high = cvCreateImage(cvSize(w, h), IPL_DEPTH_8U, 1);
    r->params.SetFrameSize(w, h);
    return r;
}

template <class R>
static void simple_process(R *r, const unsigned char *bgr, unsigned char *fg)
{
    IplImage *f = r->frame.Ptr();
    for (int y = 0; y < r->h; y++) std::memcpy(f->imageData + (size_t)y * f->widthStep, bgr + (size_t)y * r->w * 3, (size_t)r->w * 3);
    if (r->first) { r->bgs.Initalize(r->params); r->bgs.InitModel(r->frame); r->first = false; }
    r->bgs.Subtract(r->frame_number, r->frame, r->low, r->high);
    r->low.Clear();
    r->bgs.Update(r->frame_number, r->frame, r->low);
    IplImage *m = r->high.Ptr();
    for (int y = 0; y < r->h; y++) std::memcpy(fg + (size_t)y * r->w, m->imageData + (size_t)y * m->widthStep, (size_t)r->w);
    r->frame_number++;
}

typedef dp_simple_ref<AdaptiveMedianBGS, AdaptiveMedianParams> dpmed_ref;
typedef dp_simple_ref<MeanBGS, MeanParams> dpmean_ref;
typedef dp_simple_ref<WrenGA, WrenParams> dpwren_ref;
typedef dp_simple_ref<PratiMediodBGS, PratiParams> dpprati_ref;
#define DP_EXPORT __attribute__((visibility("default")))

extern "C" {

DP_EXPORT dpmed_ref *dpmed_ref_create(int w, int h, int threshold, int samplingRate, int learningFrames)
{
    dpmed_ref *r = simple_create<dpmed_ref>(w, h);
    r->params.LowThreshold() = threshold;                             // DPAdaptiveMedianBGS.cpp:56-59
    r->params.HighThreshold() = 2 * r->params.LowThreshold();
    r->params.SamplingRate() = samplingRate;
    r->params.LearningFrames() = learningFrames;
    return r;
}
DP_EXPORT void dpmed_ref_process(dpmed_ref *r, const unsigned char *bgr, unsigned char *fg) { simple_process(r, bgr, fg); }
DP_EXPORT void dpmed_ref_destroy(dpmed_ref *r) { delete r; }

DP_EXPORT dpmean_ref *dpmean_ref_create(int w, int h, int threshold, double alpha, int learningFrames)
{
    dpmean_ref *r = simple_create<dpmean_ref>(w, h);
    r->params.LowThreshold() = threshold;                             // DPMeanBGS.cpp:56-61
    r->params.HighThreshold() = 2 * r->params.LowThreshold();
    r->params.Alpha() = alpha;
    r->params.LearningFrames() = learningFrames;
    return r;
}
DP_EXPORT void dpmean_ref_process(dpmean_ref *r, const unsigned char *bgr, unsigned char *fg) { simple_process(r, bgr, fg); }
DP_EXPORT void dpmean_ref_destroy(dpmean_ref *r) { delete r; }

DP_EXPORT dpwren_ref *dpwren_ref_create(int w, int h, double threshold, double alpha, int learningFrames)
{
    dpwren_ref *r = simple_create<dpwren_ref>(w, h);
    r->params.LowThreshold() = threshold;                             // DPWrenGABGS.cpp:56-60
    r->params.HighThreshold() = 2 * r->params.LowThreshold();
    r->params.Alpha() = alpha;
    r->params.LearningFrames() = learningFrames;
    return r;
}
DP_EXPORT void dpwren_ref_process(dpwren_ref *r, const unsigned char *bgr, unsigned char *fg) { simple_process(r, bgr, fg); }
DP_EXPORT void dpwren_ref_destroy(dpwren_ref *r) { delete r; }


DP_EXPORT dpprati_ref *dpprati_ref_create(int w, int h, int threshold, int samplingRate, int historySize, int weight)
{
    dpprati_ref *r = simple_create<dpprati_ref>(w, h);
    r->params.LowThreshold() = threshold;                             // DPPratiMediodBGS.cpp:57-62
    r->params.HighThreshold() = 2 * r->params.LowThreshold();
    r->params.SamplingRate() = samplingRate;
    r->params.HistorySize() = historySize;
    r->params.Weight() = weight;
    return r;
}
DP_EXPORT void dpprati_ref_process(dpprati_ref *r, const unsigned char *bgr, unsigned char *fg) { simple_process(r, bgr, fg); }
DP_EXPORT void dpprati_ref_destroy(dpprati_ref *r) { delete r; }

}
