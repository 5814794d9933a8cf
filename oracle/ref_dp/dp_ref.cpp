// TEST INFRASTRUCTURE.  C entry points around the REFERENCE's own ZivkovicAGMM class, compiled together with
// /root/reference/package_bgs/dp/{ZivkovicAGMM,Image}.cpp into oracle/_ref/libdp_ref.so (oracle/Makefile: `make ref`).
// The call sequence is DPZivkovicAGMMBGS::process's (package_bgs/dp/DPZivkovicAGMMBGS.cpp:32-84): Initalize +
// InitModel on the first frame, then Subtract / Update per frame; the plugin's output is the HIGH-threshold mask.
#include <cstring>

#include "ZivkovicAGMM.h"

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

}
