// COMPILE-CHECK STUB ONLY.  The real adapters are built against OpenCV 2.4 headers (the version the
// reference needs: opencv2/legacy/blobtrack.hpp, CvFileStorage, cv::Mat(IplImage*)).  This image has
// no OpenCV C++ headers, so tests/test_abi.py compiles the adapters with -fsyntax-only against this
// minimal declaration set: exactly the OpenCV surface the adapters touch, nothing more.
#pragma once
#include <stddef.h>
#include <string>

#define CV_8UC1 0
#define CV_8UC3 16
#define CV_STORAGE_READ 0
#define CV_STORAGE_WRITE 1
#define IPL_DEPTH_8U 8
#define CV_Assert(expr) do { if (!(expr)) throw cv::Exception(); } while (0)

struct IplImage {
    int nChannels, depth, width, height, widthStep;
    char *imageData;
};
struct CvFileStorage;
struct CvFileNode;
CvFileStorage *cvOpenFileStorage(const char *filename, void *memstorage, int flags);
void cvReleaseFileStorage(CvFileStorage **fs);
void cvWriteInt(CvFileStorage *fs, const char *name, int value);
void cvWriteReal(CvFileStorage *fs, const char *name, double value);
int cvReadIntByName(const CvFileStorage *fs, const CvFileNode *map, const char *name, int default_value = 0);
double cvReadRealByName(const CvFileStorage *fs, const CvFileNode *map, const char *name, double default_value = 0.);

namespace cv {
struct Exception {};
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
class Mat {
public:
    int rows, cols;
    unsigned char *data;
    struct Step { size_t v; operator size_t() const { return v; } } step;
    Mat();
    Mat(int rows, int cols, int type);
    Mat(const IplImage *img, bool copyData = false);
    void create(int rows, int cols, int type);
    bool empty() const;
    int type() const;
    int channels() const;
    bool isContinuous() const;
    Size size() const;
    void copyTo(Mat &m) const;
    operator IplImage() const;
};
void imshow(const std::string &winname, const Mat &mat);
}  // namespace cv
