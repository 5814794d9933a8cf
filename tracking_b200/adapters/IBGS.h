// The plugin interface of the reference, unchanged (package_bgs/IBGS.h:21-33).  When the adapters are
// dropped into the reference tree this file is NOT used -- the tree's own package_bgs/IBGS.h is; it is
// repeated here only so that the adapters can be compiled and checked on their own.
#pragma once
#include <opencv2/opencv.hpp>

class IBGS
{
public:
  virtual void process(const cv::Mat &img_input, cv::Mat &img_foreground, cv::Mat &img_background) = 0;
  virtual ~IBGS(){}

private:
  virtual void saveConfig() = 0;
  virtual void loadConfig() = 0;
};
