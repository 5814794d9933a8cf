// MINIMAL FUNCTIONAL STAND-IN: the slice of OpenCV 2.4 legacy/blobtrack.hpp the adapters derive from and use
// (CvBlob, CvBlobSeq as a growable list, the CvVSModule / CvFGDetector / CvBlobDetector interfaces).
#pragma once
#include <vector>

#include "opencv2/opencv.hpp"

struct CvRect { int x, y, width, height; };
inline CvRect cvRect(int x, int y, int w, int h) { CvRect r = {x, y, w, h}; return r; }

struct CvBlob { float x, y, w, h; int ID; };
inline CvBlob cvBlob(float x, float y, float w, float h) { CvBlob b = {x, y, w, h, 0}; return b; }

class CvBlobSeq {
    std::vector<CvBlob> blobs;

public:
    virtual ~CvBlobSeq() {}
    virtual CvBlob *GetBlob(int BlobIndex) { return (BlobIndex >= 0 && BlobIndex < (int)blobs.size()) ? &blobs[BlobIndex] : 0; }
    virtual int GetBlobNum() { return (int)blobs.size(); }
    virtual void AddBlob(CvBlob *pB) { if (pB) blobs.push_back(*pB); }
    virtual void Clear() { blobs.clear(); }
};

class CvVSModule {
public:
    virtual ~CvVSModule() {}
    virtual void Release() = 0;
};

class CvFGDetector : public CvVSModule {
public:
    virtual IplImage *GetMask() = 0;
    virtual void Process(IplImage *pImg) = 0;
};

class CvBlobDetector : public CvVSModule {
public:
    virtual int DetectNewBlob(IplImage *pImg, IplImage *pImgFG, CvBlobSeq *pNewBlobList, CvBlobSeq *pOldBlobList) = 0;
};
