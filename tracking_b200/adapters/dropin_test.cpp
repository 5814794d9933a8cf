// Drop-in test program: the reference's own call patterns, COMPILED, LINKED against libbgsb200.so and RUN
// (tests/test_gpu_cpp_dropin.py builds it against the functional OpenCV stand-in of adapters/stub_opencv and compares
// what it writes with the golden hashes / the oracle).
//
//   dropin_test <clip.raw> <outdir>
//     clip.raw : int32 n, h, w, then n frames of h*w*3 bytes (BGR)
//     outdir   : receives <Plugin>.fg / <Plugin>.bg (concatenated outputs of the frames that produced one), ustc.fg,
//                blobs.txt, feed.txt, and config/<Plugin>.xml written by the plugins' saveConfig
//
// Patterns exercised:
//   FrameProcessor.cpp:40-59,157-167,459-478   new <Plugin>; bgs->process(img_input, img_bgs, img_bkgmodel); delete
//   FrameProcessor.cpp:169-215 (+ FanOut)      every plugin on the same prepared frame, one upload
//   ustc_src/ustc_bgs.cpp:3-113                USTC_BGS(type): Process(IplImage*), GetMask(), Release()
//   ustc_src/trackingMain.cpp:56,626           cvCreateBlobDetectorCC -> DetectNewBlob(pImg, pFG, &NewBlobs, &OldBlobs)
//   ustc_src/trackingMain.cpp:70-78,166        the tracker's second look at the mask (BgsbTrackerFeed)
#include <stdio.h>
#include <string>
#include <vector>

#include "ustc_bgs_b200.h"

static void append(const std::string &path, const cv::Mat &m)
{
    FILE *f = fopen(path.c_str(), "ab");
    if (!f) { perror(path.c_str()); exit(2); }
    for (int y = 0; y < m.rows; y++) fwrite(m.data + (size_t)y * m.step, 1, (size_t)m.cols * m.channels(), f);
    fclose(f);
}

template <class Plugin>
static void run_plugin(const char *name, const std::vector<cv::Mat> &frames, const std::string &out)
{
    IBGS *bgs = new Plugin;                                   // FrameProcessor.cpp:40-59
    cv::Mat img_bgs, img_bkgmodel;
    int nfg = 0, nbg = 0;
    for (size_t t = 0; t < frames.size(); t++) {
        img_bgs = cv::Mat(); img_bkgmodel = cv::Mat();        // fresh outputs: "left untouched" shows as empty
        bgs->process(frames[t], img_bgs, img_bkgmodel);       // FrameProcessor.cpp:163
        if (!img_bgs.empty()) { append(out + "/" + name + ".fg", img_bgs); nfg++; }
        if (!img_bkgmodel.empty()) { append(out + "/" + name + ".bg", img_bkgmodel); nbg++; }
    }
    delete bgs;                                               // FrameProcessor.cpp:459-478
    printf("%s: %d masks, %d background images\n", name, nfg, nbg);
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: dropin_test <clip.raw> <outdir>\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int hdr[3];
    if (fread(hdr, 4, 3, f) != 3) return 2;
    const int n = hdr[0], h = hdr[1], w = hdr[2];
    std::vector<cv::Mat> frames;
    for (int t = 0; t < n; t++) {
        cv::Mat m(h, w, CV_8UC3);
        if (fread(m.data, 1, (size_t)h * w * 3, f) != (size_t)h * w * 3) return 2;
        frames.push_back(m);
    }
    fclose(f);
    const std::string out = argv[2];
    try {
        // an empty frame returns silently with the outputs untouched (every plugin, e.g. WeightedMovingVarianceBGS.cpp:32-33)
        {
            FrameDifferenceBGS fd;
            cv::Mat none, a, b;
            fd.process(none, a, b);
            if (!a.empty() || !b.empty()) { fprintf(stderr, "empty input wrote an output\n"); return 1; }
        }
        run_plugin<FrameDifferenceBGS>("FrameDifferenceBGS", frames, out);
        run_plugin<WeightedMovingVarianceBGS>("WeightedMovingVarianceBGS", frames, out);
        run_plugin<MixtureOfGaussianV2BGS>("MixtureOfGaussianV2BGS", frames, out);
        run_plugin<AdaptiveBackgroundLearning>("AdaptiveBackgroundLearning", frames, out);
        run_plugin<StaticFrameDifferenceBGS>("StaticFrameDifferenceBGS", frames, out);
        run_plugin<WeightedMovingMeanBGS>("WeightedMovingMeanBGS", frames, out);
        run_plugin<DPAdaptiveMedianBGS>("DPAdaptiveMedianBGS", frames, out);      // DP package, USTC_BGS types 9 / 12 / 13
        run_plugin<DPMeanBGS>("DPMeanBGS", frames, out);
        run_plugin<DPWrenGABGS>("DPWrenGABGS", frames, out);
        run_plugin<DPPratiMediodBGS>("DPPratiMediodBGS", frames, out);             // type 14
        run_plugin<SigmaDeltaBGS>("SigmaDeltaBGS", frames, out);                   // BL package, type 35

        // FrameProcessor::process with the one added line: a single upload feeds every enabled plugin
        {
            IBGS *p[4] = {new FrameDifferenceBGS, new WeightedMovingVarianceBGS, new MixtureOfGaussianV2BGS, new AdaptiveBackgroundLearning};
            const char *names[4] = {"fan_FrameDifferenceBGS", "fan_WeightedMovingVarianceBGS", "fan_MixtureOfGaussianV2BGS",
                                    "fan_AdaptiveBackgroundLearning"};
            bgsb_adapter::FanOut fanout;
            for (int i = 0; i < 4; i++) fanout.add(p[i]);
            for (size_t t = 0; t < frames.size(); t++) {
                fanout.process(frames[t]);
                for (int i = 0; i < 4; i++) {
                    cv::Mat fg, bg;
                    p[i]->process(frames[t], fg, bg);
                    if (!fg.empty()) append(out + "/" + names[i] + ".fg", fg);
                }
            }
            for (int i = 0; i < 4; i++) delete p[i];
        }

        // ustc_src/trackingMain.cpp: USTC_BGS as the CvFGDetector, the CC blob detector, the tracker's second look
        {
            CvFGDetector *fg = new USTC_BGS(0);                         // :33-35, type 0 = FrameDifference
            CvBlobDetector *bd = cvCreateBlobDetectorCC_B200();         // :626
            BgsbTrackerFeed feed;
            CvBlobSeq oldb;
            FILE *fb = fopen((out + "/blobs.txt").c_str(), "w"), *ff = fopen((out + "/feed.txt").c_str(), "w");
            for (size_t t = 0; t < frames.size(); t++) {
                IplImage img = frames[t];                               // cvQueryFrame, :161
                fg->Process(&img);                                      // CvBlobTrackerAuto1::Process -> FG detector
                IplImage *mask = fg->GetMask();
                if (!mask) { fprintf(fb, "%d nomask\n", (int)t); continue; }
                cv::Mat m(mask);
                append(out + "/ustc.fg", m);
                bgsbMorph(mask, BGSB_MORPH_ERODE, 1);                   // the clean-up stage (SURVEY 8a row aM): OPEN 3x3
                bgsbMorph(mask, BGSB_MORPH_DILATE, 1);
                append(out + "/ustc_open.fg", cv::Mat(mask));
                CvBlobSeq newb;
                int r = bd->DetectNewBlob(&img, mask, &newb, &oldb);
                fprintf(fb, "%d %d", (int)t, r);
                for (int i = 0; i < newb.GetBlobNum(); i++) {
                    CvBlob *b = newb.GetBlob(i);
                    fprintf(fb, " %.9g %.9g %.9g %.9g", b->x, b->y, b->w, b->h);
                }
                fprintf(fb, "\n");
                feed.Update(mask);
                fprintf(ff, "%d %d", (int)t, feed.GetContourNum());
                std::vector<CvRect> rois;
                for (int i = 0; i < feed.GetContourNum(); i++) {
                    CvRect rc = feed.GetContourRect(i);
                    fprintf(ff, " %d,%d,%d,%d", rc.x, rc.y, rc.width, rc.height);
                    rois.push_back(rc);
                }
                std::vector<double> sums(rois.size() + 1);
                feed.SumROIs(rois.empty() ? 0 : &rois[0], (int)rois.size(), &sums[0]);
                fprintf(ff, " |");
                for (size_t i = 0; i < rois.size(); i++) fprintf(ff, " %.0f", sums[i]);
                fprintf(ff, " | %.0f\n", feed.SumROI(cvRect(0, 0, mask->width, mask->height)));
            }
            fclose(fb); fclose(ff);
            bd->Release();
            fg->Release();                                              // :763 cvReleaseFGDetector
            delete fg;
        }
    } catch (const cv::Exception &e) {
        fprintf(stderr, "cv::Exception: %s (%s)\n", e.what(), bgsb_last_error());
        return 1;
    }
    printf("kernels launched: %llu\n", (unsigned long long)bgsb_kernel_launch_count());
    return 0;
}
