// COMPILE-CHECK STUB ONLY: the slice of OpenCV 2.4 legacy/blobtrack.hpp the adapters derive from.
#pragma once
#include "opencv2/opencv.hpp"

struct CvBlob { float x, y, w, h; int ID; };
inline CvBlob cvBlob(float x, float y, float w, float h) { CvBlob b = {x, y, w, h, 0}; return b; }

class CvBlobSeq {
public:
    virtual ~CvBlobSeq();
    virtual CvBlob *GetBlob(int BlobIndex);
    virtual int GetBlobNum();
    virtual void AddBlob(CvBlob *pB);
};

class CvVSModule {
public:
    virtual ~CvVSModule();
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
