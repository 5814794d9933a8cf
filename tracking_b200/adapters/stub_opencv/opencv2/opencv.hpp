// MINIMAL FUNCTIONAL STAND-IN for the slice of the OpenCV C / C++ API the adapters touch -- containers only, no image
// arithmetic.  The real adapters are built against OpenCV 2.4 (the version the reference needs:
// opencv2/legacy/blobtrack.hpp, CvFileStorage, cv::Mat(IplImage*)).  This image has no OpenCV C++ headers, so
// tests/test_abi.py and tests/test_gpu_cpp_dropin.py build the adapters and the drop-in test program
// (adapters/dropin_test.cpp) against these declarations and RUN them against libbgsb200.so.
//   cv::Mat       reference-counted byte matrix (create / copyTo / header over an IplImage / operator IplImage)
//   IplImage      the header fields the adapters read
//   CvFileStorage flat key -> value XML files with the cvWriteInt / cvReadIntByName calls of saveConfig / loadConfig
//   cv::imshow    no-op
// -DBGSB_STUB_CV_MAJOR=<n> selects which OpenCV generation the stand-in claims to be (default 2, like the reference's
// build); the adapters pick their BGR2GRAY / addWeighted / cvFindContours variants from CV_MAJOR_VERSION.
#pragma once
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <memory>
#include <string>

#ifndef BGSB_STUB_CV_MAJOR
#define BGSB_STUB_CV_MAJOR 2
#endif
#define CV_MAJOR_VERSION BGSB_STUB_CV_MAJOR
#define CV_MINOR_VERSION 4

#define CV_8UC1 0
#define CV_8UC3 16
#define CV_STORAGE_READ 0
#define CV_STORAGE_WRITE 1
#define IPL_DEPTH_8U 8
#define CV_Assert(expr) do { if (!(expr)) throw cv::Exception(#expr, __FILE__, __LINE__); } while (0)

struct IplImage {
    int nChannels, depth, width, height, widthStep;
    char *imageData;
};

namespace cv {
struct Exception {
    std::string msg;
    Exception() {}
    Exception(const char *expr, const char *file, int line) { msg = std::string(file) + ":" + std::to_string(line) + ": " + expr; }
    const char *what() const { return msg.c_str(); }
};
}  // namespace cv

// ---- CvFileStorage: <opencv_storage><key>value</key>...</opencv_storage>, one level -------------------------------------
struct CvFileNode;
struct CvFileStorage {
    std::string path;
    int flags;
    std::map<std::string, std::string> kv;
    std::string order;            // keys in insertion order, '\n' separated (files are written in call order)
};
inline CvFileStorage *cvOpenFileStorage(const char *filename, void * /*memstorage*/, int flags)
{
    if (flags == CV_STORAGE_WRITE) {
        FILE *f = fopen(filename, "w");            // like OpenCV: NULL when the directory does not exist
        if (!f) return 0;
        fclose(f);
        CvFileStorage *fs = new CvFileStorage();
        fs->path = filename; fs->flags = flags;
        return fs;
    }
    FILE *f = fopen(filename, "r");
    if (!f) return 0;
    std::string text;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
    fclose(f);
    CvFileStorage *fs = new CvFileStorage();
    fs->path = filename; fs->flags = flags;
    size_t pos = text.find("<opencv_storage>");
    pos = pos == std::string::npos ? 0 : pos + 16;
    for (;;) {
        size_t a = text.find('<', pos);
        if (a == std::string::npos || text.compare(a, 2, "</") == 0) break;
        size_t b = text.find('>', a);
        if (b == std::string::npos) break;
        const std::string key = text.substr(a + 1, b - a - 1);
        size_t c = text.find("</" + key + ">", b);
        if (c == std::string::npos) break;
        fs->kv[key] = text.substr(b + 1, c - b - 1);
        pos = c + key.size() + 3;
    }
    return fs;
}
inline void cvReleaseFileStorage(CvFileStorage **fs)
{
    if (!fs || !*fs) return;
    if ((*fs)->flags == CV_STORAGE_WRITE) {
        FILE *f = fopen((*fs)->path.c_str(), "w");
        if (f) {
            fprintf(f, "<?xml version=\"1.0\"?>\n<opencv_storage>\n");
            size_t p = 0;
            while (p < (*fs)->order.size()) {
                size_t e = (*fs)->order.find('\n', p);
                const std::string k = (*fs)->order.substr(p, e - p);
                fprintf(f, "<%s>%s</%s>\n", k.c_str(), (*fs)->kv[k].c_str(), k.c_str());
                p = e + 1;
            }
            fprintf(f, "</opencv_storage>\n");
            fclose(f);
        }
    }
    delete *fs;
    *fs = 0;
}
inline void cvWriteInt(CvFileStorage *fs, const char *name, int value)
{
    if (!fs) return;
    if (!fs->kv.count(name)) fs->order += std::string(name) + "\n";
    fs->kv[name] = std::to_string(value);
}
inline void cvWriteReal(CvFileStorage *fs, const char *name, double value)
{
    if (!fs) return;
    char b[64];
    snprintf(b, sizeof(b), "%.17g", value);
    if (!fs->kv.count(name)) fs->order += std::string(name) + "\n";
    fs->kv[name] = b;
}
// a missing file (fs == NULL) or key yields the default, as in OpenCV (cvGetFileNodeByName(NULL, ...) == NULL)
inline int cvReadIntByName(const CvFileStorage *fs, const CvFileNode *, const char *name, int default_value = 0)
{
    if (!fs) return default_value;
    std::map<std::string, std::string>::const_iterator it = fs->kv.find(name);
    return it == fs->kv.end() ? default_value : (int)strtol(it->second.c_str(), 0, 10);
}
inline double cvReadRealByName(const CvFileStorage *fs, const CvFileNode *, const char *name, double default_value = 0.)
{
    if (!fs) return default_value;
    std::map<std::string, std::string>::const_iterator it = fs->kv.find(name);
    return it == fs->kv.end() ? default_value : strtod(it->second.c_str(), 0);
}

namespace cv {
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };

class Mat {
    std::shared_ptr<unsigned char> owner;         // empty for headers over foreign data
    int type_;

public:
    int rows, cols;
    unsigned char *data;
    struct Step { size_t v; Step() : v(0) {} operator size_t() const { return v; } } step;

    Mat() : type_(CV_8UC1), rows(0), cols(0), data(0) {}
    Mat(int r, int c, int type) : type_(CV_8UC1), rows(0), cols(0), data(0) { create(r, c, type); }
    // header over an IplImage (no copy), or a copy of it
    Mat(const IplImage *img, bool copyData = false) : type_(CV_8UC1), rows(0), cols(0), data(0)
    {
        if (!img) return;
        const int type = img->nChannels == 3 ? CV_8UC3 : CV_8UC1;
        if (copyData) {
            create(img->height, img->width, type);
            for (int y = 0; y < rows; y++) memcpy(data + (size_t)y * step.v, img->imageData + (size_t)y * img->widthStep, (size_t)cols * channels());
        } else {
            type_ = type; rows = img->height; cols = img->width;
            data = (unsigned char *)img->imageData; step.v = (size_t)img->widthStep;
        }
    }
    void create(int r, int c, int type)
    {
        if (data && r == rows && c == cols && type == type_) return;
        type_ = type; rows = r; cols = c;
        const int ch = type == CV_8UC3 ? 3 : 1;
        step.v = (size_t)c * ch;
        const size_t bytes = (size_t)r * step.v;
        owner.reset(bytes ? new unsigned char[bytes] : 0, std::default_delete<unsigned char[]>());
        data = owner.get();
    }
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    int type() const { return type_; }
    int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
    bool isContinuous() const { return step.v == (size_t)cols * channels(); }
    Size size() const { return Size(cols, rows); }
    void copyTo(Mat &m) const
    {
        if (empty()) { m = Mat(); return; }
        m.create(rows, cols, type_);
        for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step.v, data + (size_t)y * step.v, (size_t)cols * channels());
    }
    operator IplImage() const
    {
        IplImage h;
        h.nChannels = channels(); h.depth = IPL_DEPTH_8U; h.width = cols; h.height = rows;
        h.widthStep = (int)step.v; h.imageData = (char *)data;
        return h;
    }
};
inline void imshow(const std::string & /*winname*/, const Mat & /*mat*/) {}
}  // namespace cv
