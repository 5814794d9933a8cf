// TEST INFRASTRUCTURE.  Minimal stand-in for <opencv2/opencv.hpp>, written for this repository, so that the
// reference's own DP background-subtraction sources (package_bgs/dp/{ZivkovicAGMM,Image}.cpp) compile from
// where they lie under /root/reference without OpenCV: those files use OpenCV only as an image CONTAINER
// (IplImage, cvCreateImage, cvReleaseImage, cvZero, cvSet, cvSize) -- every arithmetic operation of the algorithm is
// the reference's own C++.  Nothing here computes anything.
#pragma once
#include <cassert>      // the real header pulls these in; Image.h / PratiMediodBGS.cpp (INT_MAX) rely on it
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>

#define IPL_DEPTH_8U 8
#define IPL_DEPTH_32F 32
#define IPL_ORIGIN_TL 0
#define IPL_ORIGIN_BL 1

struct CvSize { int width, height; };
static inline CvSize cvSize(int w, int h) { CvSize s; s.width = w; s.height = h; return s; }

struct IplImage {
    int nChannels, depth, width, height, widthStep, origin, imageSize;
    char *imageData;
};

static inline IplImage *cvCreateImage(CvSize size, int depth, int channels)
{
    IplImage *img = (IplImage *)std::calloc(1, sizeof(IplImage));
    img->nChannels = channels; img->depth = depth; img->width = size.width; img->height = size.height;
    img->widthStep = (size.width * channels * (depth / 8) + 3) / 4 * 4;       // OpenCV aligns rows to 4 bytes
    img->imageSize = img->widthStep * size.height;
    img->imageData = (char *)std::calloc(1, img->imageSize ? img->imageSize : 1);
    return img;
}
static inline void cvReleaseImage(IplImage **img)
{
    if (img && *img) { std::free((*img)->imageData); std::free(*img); *img = 0; }
}
static inline void cvZero(IplImage *img) { std::memset(img->imageData, 0, img->imageSize); }

// AdaptiveMedianBGS::Initalize fills its model image with a constant before InitModel overwrites it
struct CvScalar { double val[4]; };
#define CV_RGB(r, g, b) cvScalarRGB((b), (g), (r))
static inline CvScalar cvScalarRGB(double v0, double v1, double v2) { CvScalar s; s.val[0] = v0; s.val[1] = v1; s.val[2] = v2; s.val[3] = 0; return s; }
static inline void cvSet(IplImage *img, CvScalar v)
{
    assert(img->depth == IPL_DEPTH_8U);
    for (int y = 0; y < img->height; y++)
        for (int x = 0; x < img->width; x++)
            for (int c = 0; c < img->nChannels; c++)
                ((unsigned char *)(img->imageData + (size_t)y * img->widthStep))[x * img->nChannels + c] = (unsigned char)v.val[c];
}
