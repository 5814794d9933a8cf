// Syntax check of the header-only adapters (tests/test_abi.py): the reference's call patterns must compile.
#include "ustc_bgs_b200.h"

int compile_check_calls()
{
    // FrameProcessor.cpp:40-59,157-167 pattern
    IBGS *plugins[13] = {new FrameDifferenceBGS, new WeightedMovingVarianceBGS, new MixtureOfGaussianV2BGS,
                         new AdaptiveBackgroundLearning, new StaticFrameDifferenceBGS, new WeightedMovingMeanBGS,
                         new AdaptiveSelectiveBackgroundLearning, new DPZivkovicAGMMBGS, new DPAdaptiveMedianBGS,
                         new DPMeanBGS, new DPWrenGABGS, new DPPratiMediodBGS, new SigmaDeltaBGS};
    cv::Mat img_input, img_bgs, img_bkgmodel;
    // FrameProcessor.cpp:169-215 with the one added line: a single upload feeds all enabled plugins
    bgsb_adapter::FanOut fanout;
    for (int i = 0; i < 13; i++) fanout.add(plugins[i]);
    fanout.process(img_input);
    for (int i = 0; i < 13; i++) {
        plugins[i]->process(img_input, img_bgs, img_bkgmodel);
        // capture-loop form (VideoCapture.cpp:151-239): queue the frame, collect later
        bgsb_adapter::PluginBase *q = dynamic_cast<bgsb_adapter::PluginBase *>(plugins[i]);
        bool fgv = false, bgv = false;
        q->submit(img_input, img_bgs, img_bkgmodel, &fgv, &bgv);
        q->wait();
        delete plugins[i];
    }
    // ustc_src/trackingMain.cpp:33-35,613-615 pattern
    CvFGDetector *fg = new USTC_BGS(5);
    IplImage *frame = 0;
    fg->Process(frame);
    IplImage *mask = fg->GetMask();
    // :626 pattern
    CvBlobDetector *bd = cvCreateBlobDetectorCC_B200();
    CvBlobSeq newb, oldb;
    int r = bd->DetectNewBlob(frame, mask, &newb, &oldb);
    bgsbMorph(mask, BGSB_MORPH_ERODE, 1);
    bd->Release();
    fg->Release();
    return r;
}
