// stereo.h -- drop-in facade of the reference's stereo:: classes for the ADCensus path,
// backed by the B200-native C-ABI (include/tsm.h).  Class names, method signatures,
// default arguments and exception types are those of the reference:
//   stereo::ColorModel, CensusWin, ADCensusParams   include/stereo_utils.h:188-244
//   stereo::EpipolarRectifyMap                        include/stereo_utils.h:109-148
//   stereo::StereoMatching                            include/stereo.h:325-331
//   stereo::ADCensus                                  include/stereo.h:388-422
//   stereo::EpipolarRectify                           include/stereo.h:254-296
// Out of scope here ("out of scope"): YAML loaders, calibration, the
// file-glob rectify overload, DL inference wrappers, visualisation helpers.
#pragma once
#include <memory>
#include <string>
#include <vector>
#if __has_include(<opencv2/core/mat.hpp>)
#include <opencv2/core/mat.hpp>
#else
#include "cvmini.hpp"
#endif

namespace stereo
{

enum class ColorModel
{
	RGB = 0,
	HSI = 1,
};

enum class CensusWin
{
	CENSUSWIN_9x7 = 0,
	CENSUSWIN_7x5 = 1,
};

// Tunables of the matcher (reference values: source/stereo_utils.cpp:271-326).  The
// CUDA kernels implement the RGB set; the struct is kept for API parity.
class ADCensusParams
{
public:
	ADCensusParams() { setADCensusParams(ColorModel::RGB); }
	ADCensusParams(const ColorModel& colorModel) { setADCensusParams(colorModel); }
	~ADCensusParams() {}
	void setADCensusParams(const ColorModel& colorModel);
public:
	float lambdaAD; CensusWin censusWin; float lambdaCensus;
	float lambdaHue, lambdaSaturation, lambdaIntensity;
	int colorThresh1, colorThresh2, saturationThresh1, saturationThresh2, intensityThresh1, intensityThresh2;
	int maxLength1, maxLength2, iterations, colorDiff;
	float pi1, pi2;
	int dispTolerance, votingThresh; float votingRatioThresh;
	int maxSearchDepth, blurKernelSize, cannyThresh1, cannyThresh2, cannyKernelSize;
};

// include/stereo_utils.h:16-80 of the reference
template <typename T>
struct StereoPair
{
	T left, right;
	StereoPair() : left(), right() {}
	StereoPair(const T& left, const T& right) : left(left), right(right) {}
	void swap() { using std::swap; swap(left, right); }
};

class CameraIntrinsic
{
public:
	cv::Mat intrinsic_matrix;        // 3x3 CV_64FC1
	cv::Mat distortion_coefficients; // 1xN / Nx1 CV_64FC1, N = 4, 5, 8, 12 or 14
	CameraIntrinsic() = default;
	CameraIntrinsic(const cv::Mat& intrinsic_matrix, const cv::Mat& distortion_coefficients)
		: intrinsic_matrix(intrinsic_matrix), distortion_coefficients(distortion_coefficients) {}
	bool empty() const { return intrinsic_matrix.empty(); }
};

class StereoExtrinsic
{
public:
	cv::Mat E, F, R, T;
	bool empty() const { return R.empty() or T.empty(); }
};

class EpipolarRectifyMap
{
public:
	cv::Mat R1, R2, P1, P2;
	cv::Mat map00, map01, map10, map11;
	EpipolarRectifyMap() = default;
	// stereo_utils.cpp:157-169: the four CV_16SC2 / CV_16UC1 maps, computed on the device (row f2)
	void compute(const StereoPair<CameraIntrinsic>& intrinsic, const cv::Size& imgsz);
	// stereo_utils.cpp:134-155: map00 / map01 / map10 / map11 from an OpenCV FileStorage YAML file
	void loadRectifyMapsYMLFile(const std::string& ymlFilePath);
	EpipolarRectifyMap(const cv::Mat& R1, const cv::Mat& R2, const cv::Mat& P1, const cv::Mat& P2,
		const cv::Mat& map00, const cv::Mat& map01, const cv::Mat& map10, const cv::Mat& map11);
	bool empty() const;
};

class EpipolarRectify
{
public:
	EpipolarRectify();
	EpipolarRectify(const EpipolarRectifyMap& rectifyMap, const cv::Size& imgsz);
	~EpipolarRectify();
	void loadEpipolarRectifyMap(const EpipolarRectifyMap& rectifyMap, const cv::Size& imgsz);
	void rectify(const cv::Mat& stereoImage, cv::Mat& rectifiedStereoImage);
	void rectify(const cv::Mat& stereoImage, cv::Mat& rectifyLeftImage, cv::Mat& rectifiedRightImage);
	void rectify(const cv::Mat& leftImage, const cv::Mat& rightImage, cv::Mat& rectifyLeftImage, cv::Mat& rectifiedRightImage);
private:
	class EpipolarRectifyImpl;
	std::unique_ptr<EpipolarRectifyImpl> impl;
	friend class ADCensus;
};

class StereoMatching
{
public:
	virtual ~StereoMatching() = 0;
	virtual void compute(const cv::Mat& leftImage, const cv::Mat& rightImage, cv::Mat& disparity) = 0;
};

class ADCensus : public StereoMatching
{
public:
	ADCensus();
	~ADCensus();
	void setMinMaxDisparity(const int& minDisparity, const int& maxDisparity);
	void setMatchingStrategy(const ColorModel& colorModel = ColorModel::RGB, const bool& roiMatching = false, const bool& maskMatching = false);
	void setOffset(const int& offset);
	void compute(const cv::Mat& leftImage, const cv::Mat& rightImage, cv::Mat& disparity) override;
	// Batched form in the style of the reference's batched ONNX signature (include/stereo.h:381,
	// SURVEY 8(f) row f4): pair i runs on device i % N (one worker thread per device, four pairs in flight each);
	// N = 1 (the device of setDevice) or, after setDevice(-1), every visible device.
	void compute(const std::vector<cv::Mat>& leftImages, const std::vector<cv::Mat>& rightImages, std::vector<cv::Mat>& disparities);
	// Extension: fused rectify -> ADCensus on a side-by-side frame (BASELINE config C4).
	void compute(EpipolarRectify& rectify, const cv::Mat& stereoImage, cv::Mat& disparity);
	// Extension: CUDA device ordinal used by this object (default 0, or env TSM_DEVICE); -1 = the batched compute
	// shards over all visible devices (single-pair calls then use the default device).
	void setDevice(const int& device);
private:
	class ADCensusImpl;
	std::unique_ptr<ADCensusImpl> impl;
};

// include/stereo_utils.h:151-190, source/stereo_utils.cpp:176-238
class StereoParams
{
public:
	StereoPair<CameraIntrinsic> intrinsic;
	StereoExtrinsic extrinsic;
	EpipolarRectifyMap map;
	cv::Mat Q;
	float rectified_f = 0.f, rectified_cx = 0.f, rectified_cy = 0.f, baseline = 0.f;
	cv::Size imgsz;
	StereoParams() = default;
	StereoParams(const std::string& ymlFilePath) { loadYAMLFile(ymlFilePath); }
	void loadYAMLFile(const std::string& ymlFilePath);
	bool empty() const;
};

// ---- consumers of the disparity map: same free functions as the reference (include/stereo.h:194-235,
// source/stereo.cpp:75-202), computed on the device (csrc/k_consumers.cu).  disparity is CV_32FC1.
cv::Mat JETColorMap();
void applyColorMap(const cv::Mat& src, cv::Mat& dst, const cv::Mat& colorMap);
void applyColorMap(const cv::Mat& src, cv::Mat& dst, float minVal, float maxVal, const cv::Mat& colorMap);
void reprojectToDepth(const cv::Mat& disparity, float focalLength, float baseline, cv::Mat& depth);
void reprojectTo3D(const cv::Mat& disparity, float focalLength, float baseline, float cx, float cy, cv::Mat& XYZPoints);
void reprojectTo3D(const cv::Mat& disparity, const cv::Mat& Q, cv::Mat& XYZPoints);  // Q: 4x4 CV_64FC1 or CV_32FC1
// source/stereo.cpp:250-278, 328-356: ASCII point clouds of the finite points (RGBImage CV_8UC3, XYZPoints CV_32FC3)
void writePointCloudToPCD(const cv::Mat& RGBImage, const cv::Mat& XYZPoints, const std::string& pcdPath);
void writePointCloudToPLY(const cv::Mat& RGBImage, const cv::Mat& XYZPoints, const std::string& plyPath);

}
