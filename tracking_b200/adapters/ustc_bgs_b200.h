// CvFGDetector / CvBlobDetector side of the boundary (OpenCV 2.4 legacy blob-tracking pipeline).
//
//   USTC_BGS              same class as ustc_src/ustc_bgs.h:58-76 / ustc_bgs.cpp:3-113, restricted to the
//                         four B200 plugins (ids 0 FD, 3 WMV, 5 MOG2, 6 ABL of the factory, .cpp:8-14)
//   BgsbBlobDetectorCC    CvBlobDetector with the behaviour of cvCreateBlobDetectorCC()
//                         (ustc_src/trackingMain.cpp:56 module table, :626 creation)
//   cvCreateBlobDetectorCC_B200()  factory with the signature the module table expects (:46-51)
//
// ustc_src/trackingMain.cpp needs two edits to use them: `type` (:34) set to 0/3/5/6, and the
// "BD_CC" row of the detector table (:56) pointing at cvCreateBlobDetectorCC_B200.
#pragma once
#include <vector>

#include "opencv2/legacy/blobtrack.hpp"

#include "bgsb_plugins.h"

class USTC_BGS : public CvFGDetector
{
public:
    int frameNum;
    IplImage *c_mask;
    IBGS *bgs;
    cv::Mat img_mask;
    cv::Mat img_bkgmodel;
    cv::Mat img_input;
    IplImage b;

    USTC_BGS(int type) : frameNum(0), c_mask(0), bgs(0)
    {
        CV_Assert(type >= 0 && type <= 37);                 // ustc_bgs.cpp:6
        switch (type) {
        case 0: bgs = new FrameDifferenceBGS; break;        // ustc_bgs.cpp:8
        case 1: bgs = new StaticFrameDifferenceBGS; break;  // :9
        case 2: bgs = new WeightedMovingMeanBGS; break;     // :10
        case 3: bgs = new WeightedMovingVarianceBGS; break; // :11
        case 5: bgs = new MixtureOfGaussianV2BGS; break;    // :13
        case 6: bgs = new AdaptiveBackgroundLearning; break;// :14
        case 7: bgs = new AdaptiveSelectiveBackgroundLearning; break;   // :15
        case 9: bgs = new DPAdaptiveMedianBGS; break;       // :19
        case 11: bgs = new DPZivkovicAGMMBGS; break;        // :21
        case 12: bgs = new DPMeanBGS; break;                // :22
        case 13: bgs = new DPWrenGABGS; break;              // :23
        case 14: bgs = new DPPratiMediodBGS; break;         // :24
        case 35: bgs = new SigmaDeltaBGS; break;            // :62
        default: CV_Assert(!"this plugin id is not on the B200 hot path (0 FD, 1 StaticFD, 2 WMM, 3 WMV, 5 MOG2, 6 ABL, 7 ASBL, 11 DPZivkovicAGMM)");
        }
    }
    ~USTC_BGS() {}

    void Release() { delete bgs; bgs = 0; }                 // ustc_bgs.cpp:75-77

    IplImage *GetMask()                                      // ustc_bgs.cpp:79-85
    {
        if (frameNum == 0) return NULL;
        return c_mask;     // NULL until the plugin produced its first mask (FD/WMV warm-up frames)
    }

    void Process(IplImage *pImg)                             // ustc_bgs.cpp:87-113
    {
        img_input = cv::Mat(pImg);
        bgs->process(img_input, img_mask, img_bkgmodel);
        if (!img_mask.empty()) {
            b = img_mask.operator IplImage();
            c_mask = &b;
        }
        frameNum++;
    }
};

class BgsbBlobDetectorCC : public CvBlobDetector
{
    bgsb_blobdetector *bd;

public:
    BgsbBlobDetectorCC() : bd(0)
    {
        CV_Assert(bgsb_blobdetector_create(&bd, 0) == BGSB_OK);
        // cvFindContours clears the mask's 1-px frame up to OpenCV 3.1 and keeps it afterwards
#if defined(CV_MAJOR_VERSION) && (CV_MAJOR_VERSION > 3 || (CV_MAJOR_VERSION == 3 && CV_MINOR_VERSION >= 2))
        bgsb_blobdetector_set_param(bd, "zeroBorder", 0);
#else
        bgsb_blobdetector_set_param(bd, "zeroBorder", 1);
#endif
    }
    ~BgsbBlobDetectorCC() { bgsb_blobdetector_destroy(bd); }
    void Release() { delete this; }

    // CvBlobDetectorCC::DetectNewBlob: returns 1 and appends to pNewBlobList when a new blob is confirmed
    int DetectNewBlob(IplImage * /*pImg*/, IplImage *pFGMask, CvBlobSeq *pNewBlobList, CvBlobSeq *pOldBlobList)
    {
        CV_Assert(pFGMask && pFGMask->nChannels == 1 && pFGMask->depth == IPL_DEPTH_8U);
        std::vector<bgsb_blob> old;
        if (pOldBlobList)
            for (int i = 0; i < pOldBlobList->GetBlobNum(); i++) {
                CvBlob *p = pOldBlobList->GetBlob(i);
                bgsb_blob o = {p->x, p->y, p->w, p->h, p->ID};
                old.push_back(o);
            }
        bgsb_blob nb[1];
        int n_new = 0, result = 0;
        int rc = bgsb_blobdetector_detect(bd, (const uint8_t *)pFGMask->imageData, pFGMask->width, pFGMask->height,
                                          (size_t)pFGMask->widthStep, old.empty() ? 0 : &old[0], (int)old.size(),
                                          nb, 1, &n_new, &result, 0, 0, 0);
        if (rc != BGSB_OK) {          // CvBlobDetectorCC has no failure path: report and return "no new blob"
            std::cerr << "bgsb200: DetectNewBlob: " << bgsb_last_error() << std::endl;
            return 0;
        }
        if (result && n_new == 1 && pNewBlobList) {
            CvBlob B = cvBlob(nb[0].x, nb[0].y, nb[0].w, nb[0].h);
            pNewBlobList->AddBlob(&B);
        }
        return result;
    }
};

inline CvBlobDetector *cvCreateBlobDetectorCC_B200() { return new BgsbBlobDetectorCC; }

// Tracker-side use of the same mask.  The tracker modules the reference lists at
// ustc_src/trackingMain.cpp:70-78 and runs per frame at :166 (pTracker->Process(pImg, pFG)) look at the foreground
// mask again: CvBlobTrackerCC / CCMSPF re-run cvFindContours(RETR_EXTERNAL) on it to get the blobs' contour
// rectangles, and CvBlobTrackerAuto1's blob deleter sums the mask under each tracked blob (cvSum of the ROI).  Both
// are answered here from ONE labelling of the mask on the GPU -- the component table and exact integer ROI sums --
// so the mask is not walked again on the CPU.
class BgsbTrackerFeed
{
    bgsb_ccl *ccl;
    int cw, ch;
    std::vector<bgsb_component> comps;
    std::vector<CvRect> rects;

public:
    BgsbTrackerFeed() : ccl(0), cw(0), ch(0) {}
    ~BgsbTrackerFeed() { if (ccl) bgsb_ccl_destroy(ccl); }

    // once per frame, with the mask CvFGDetector::GetMask() returned
    void Update(IplImage *pFGMask)
    {
        CV_Assert(pFGMask && pFGMask->nChannels == 1 && pFGMask->depth == IPL_DEPTH_8U);
        const int w = pFGMask->width, h = pFGMask->height;
        if (!ccl || w > cw || h > ch) {
            if (ccl) bgsb_ccl_destroy(ccl);
            ccl = 0;
            CV_Assert(bgsb_ccl_create(&ccl, 0, w, h) == BGSB_OK);
            cw = w; ch = h;
        }
        // cvFindContours clears the mask's 1-px frame up to OpenCV 3.1 and keeps it afterwards
#if defined(CV_MAJOR_VERSION) && (CV_MAJOR_VERSION > 3 || (CV_MAJOR_VERSION == 3 && CV_MINOR_VERSION >= 2))
        const int zero_border = 0;
#else
        const int zero_border = 1;
#endif
        comps.resize((size_t)((w + 1) / 2) * ((h + 1) / 2) + 1);
        int n = 0;
        int rc = bgsb_ccl_label(ccl, (const uint8_t *)pFGMask->imageData, w, h, (size_t)pFGMask->widthStep, zero_border, 0,
                                &comps[0], (int)comps.size(), &n);
        if (rc != BGSB_OK) std::cerr << "bgsb200: " << bgsb_last_error() << std::endl;
        CV_Assert(rc == BGSB_OK);
        comps.resize(n);
        rects.clear();
        // cvFindContours(RETR_EXTERNAL) lists the outer contours in reverse raster order of their first pixels
        for (int i = n - 1; i >= 0; i--)
            if (comps[i].external) rects.push_back(cvRect(comps[i].x, comps[i].y, comps[i].w, comps[i].h));
    }
    // the external contours of the mask: count, and ((CvContour*)cnt)->rect of the i-th one in cvFindContours order
    int GetContourNum() const { return (int)rects.size(); }
    CvRect GetContourRect(int i) const { return rects[i]; }
    // cvSum(mask(ROI)).val[0] for each rectangle (clipped to the mask), exact
    void SumROIs(const CvRect *rois, int n, double *sums)
    {
        if (n <= 0) return;
        std::vector<int32_t> flat((size_t)n * 4);
        for (int i = 0; i < n; i++) {
            int x0 = rois[i].x < 0 ? 0 : rois[i].x, y0 = rois[i].y < 0 ? 0 : rois[i].y;
            int x1 = rois[i].x + rois[i].width, y1 = rois[i].y + rois[i].height;
            if (x1 > cw) x1 = cw;
            if (y1 > ch) y1 = ch;
            flat[4 * i] = x0; flat[4 * i + 1] = y0;
            flat[4 * i + 2] = x1 > x0 ? x1 - x0 : 0; flat[4 * i + 3] = y1 > y0 ? y1 - y0 : 0;
        }
        std::vector<uint64_t> mom((size_t)n * 6);
        CV_Assert(bgsb_ccl_rect_moments(ccl, &flat[0], n, &mom[0]) == BGSB_OK);
        for (int i = 0; i < n; i++) sums[i] = (double)mom[6 * (size_t)i];
    }
    double SumROI(CvRect roi) { double s = 0; SumROIs(&roi, 1, &s); return s; }
};

// cvErode / cvDilate(mask, mask, NULL, n) on an 8-bit single-channel IplImage, on the GPU.
inline void bgsbMorph(IplImage *mask, int op /* BGSB_MORPH_ERODE | BGSB_MORPH_DILATE */, int iterations)
{
    int ops[2] = {op, iterations};
    CV_Assert(mask && mask->nChannels == 1);
    CV_Assert(bgsb_morph((const uint8_t *)mask->imageData, mask->width, mask->height, (size_t)mask->widthStep, ops, 1,
                         (uint8_t *)mask->imageData, (size_t)mask->widthStep) == BGSB_OK);
}
