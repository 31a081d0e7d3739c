// stereo.cpp -- implementation of the facade in stereo.h on top of the tsm_* C-ABI.
// Error behaviour follows the reference: setters and the image check throw std::string
// (source/ADCensus.cpp:310,326,333), internal failures std::runtime_error (:383-387);
// EpipolarRectify logs and returns on bad input (source/EpipolarRectify.cpp:48-57,70-79,89-98)
// but loadEpipolarRectifyMap throws std::runtime_error on empty maps (:35-40).
#include "stereo.h"
#include "../../include/tsm.h"
#include <algorithm>
#include <fstream>
#include <map>
#include <sstream>
#include <iterator>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <atomic>
#include <exception>
#include <mutex>
#include <thread>

namespace {

// Process-wide id of rectify-map CONTENT (tsm.h: map_generation): every loadEpipolarRectifyMap takes a new one.
std::atomic<unsigned long long> g_map_generation{1};

bool log_enabled() { static const bool on = std::getenv("TSM_LOG") != nullptr; return on; }
void log_info(const std::string& m) { if (log_enabled()) std::fprintf(stderr, "[INFO] %s\n", m.c_str()); }
void log_error(const std::string& m) { std::fprintf(stderr, "[ERROR] %s\n", m.c_str()); }

int default_device()
{
	const char* e = std::getenv("TSM_DEVICE");
	return e ? std::atoi(e) : 0;
}

struct Ctx {
	tsm_ctx* h = nullptr;
	int device = -1;
	Ctx() = default;
	Ctx(const Ctx&) = delete;
	Ctx& operator=(const Ctx&) = delete;
	Ctx(Ctx&& o) noexcept : h(o.h), device(o.device) { o.h = nullptr; o.device = -1; }
	void ensure(int dev)
	{
		if (h && device == dev) return;
		reset();
		if (tsm_create(dev, &h) != TSM_OK) throw std::runtime_error(tsm_last_error(nullptr));
		device = dev;
	}
	void reset() { if (h) tsm_destroy(h); h = nullptr; device = -1; }
	~Ctx() { reset(); }
};

int map_kind_of(const cv::Mat& m1, const cv::Mat& m2)
{
	if (m1.type() == CV_16SC2 && m2.type() == CV_16UC1) return TSM_MAP_FIXED_16SC2_16UC1;
	if (m1.type() == CV_32FC1 && m2.type() == CV_32FC1) return TSM_MAP_FLOAT_32FC1_X2;
	return -1;
}

}  // namespace

// ---------------------------------------------------------------- ADCensusParams
void stereo::ADCensusParams::setADCensusParams(const ColorModel& colorModel)
{
	lambdaAD = 10.f; censusWin = CensusWin::CENSUSWIN_9x7; lambdaCensus = 30.f;
	lambdaHue = 1.f; lambdaSaturation = 2.5f; lambdaIntensity = 2.5f;
	iterations = 4; pi1 = 1.f; pi2 = 3.f; dispTolerance = 0; votingThresh = 20; votingRatioThresh = 0.4f;
	maxSearchDepth = 20; blurKernelSize = 3; cannyThresh1 = 30; cannyThresh2 = 90; cannyKernelSize = 3;
	const bool rgb = colorModel == ColorModel::RGB;
	colorThresh1 = rgb ? 20 : 5; colorThresh2 = rgb ? 6 : 1;
	maxLength1 = rgb ? 34 : 17; maxLength2 = rgb ? 17 : 8; colorDiff = rgb ? 15 : 3;
	saturationThresh1 = rgb ? 0 : 10; saturationThresh2 = rgb ? 0 : 2;
	intensityThresh1 = rgb ? 0 : 12; intensityThresh2 = rgb ? 0 : 3;
}

// ------------------------------------------------------------ EpipolarRectifyMap
stereo::EpipolarRectifyMap::EpipolarRectifyMap(const cv::Mat& R1_, const cv::Mat& R2_, const cv::Mat& P1_, const cv::Mat& P2_,
	const cv::Mat& m00, const cv::Mat& m01, const cv::Mat& m10, const cv::Mat& m11)
	: R1(R1_), R2(R2_), P1(P1_), P2(P2_), map00(m00), map01(m01), map10(m10), map11(m11) {}

bool stereo::EpipolarRectifyMap::empty() const
{
	return map00.empty() || map01.empty() || map10.empty() || map11.empty();
}

// --------------------------------------------------------------- EpipolarRectify
class stereo::EpipolarRectify::EpipolarRectifyImpl
{
public:
	EpipolarRectifyMap m_rectifyMap;
	cv::Size m_imgsz;
	cv::Mat c00, c01, c10, c11;  // continuous copies handed to the C-ABI
	int kind = -1;
	unsigned long long gen = 0;  // identifies the content of c00..c11 to the device-side map cache
	Ctx ctx;
};

stereo::EpipolarRectify::EpipolarRectify() { impl = std::make_unique<EpipolarRectifyImpl>(); }
stereo::EpipolarRectify::EpipolarRectify(const EpipolarRectifyMap& rectifyMap, const cv::Size& imgsz)
{
	impl = std::make_unique<EpipolarRectifyImpl>();
	loadEpipolarRectifyMap(rectifyMap, imgsz);
}
stereo::EpipolarRectify::~EpipolarRectify() {}

void stereo::EpipolarRectify::loadEpipolarRectifyMap(const EpipolarRectifyMap& rectifyMap, const cv::Size& imgsz)
{
	log_info("Loading stereo epipolar rectify params...");
	if (rectifyMap.empty()) {
		std::string msg = "stereo params is empty, please load it first";
		log_error(msg);
		throw std::runtime_error(msg);
	}
	impl->m_rectifyMap = rectifyMap;
	impl->m_imgsz = imgsz;
	impl->c00 = rectifyMap.map00.clone(); impl->c01 = rectifyMap.map01.clone();
	impl->c10 = rectifyMap.map10.clone(); impl->c11 = rectifyMap.map11.clone();
	impl->kind = map_kind_of(impl->c00, impl->c01);
	if (impl->kind < 0 || map_kind_of(impl->c10, impl->c11) != impl->kind)
		throw std::runtime_error("[EpipolarRectify] unsupported map types (need CV_16SC2+CV_16UC1 or CV_32FC1 x2)");
	impl->gen = g_map_generation.fetch_add(1);
	log_info("Loaded stereo epipolar rectify params!");
}

void stereo::EpipolarRectify::rectify(const cv::Mat& stereoImage, cv::Mat& rectifiedStereoImage)
{
	if (impl->m_rectifyMap.empty()) { log_error("Stereo epipolar rectify params is empty, please load it first."); return; }
	if (stereoImage.empty()) { log_error("Stereo image is empty."); return; }
	cv::Mat l, r;
	rectify(stereoImage, l, r);
	cv::hconcat(l, r, rectifiedStereoImage);
}

void stereo::EpipolarRectify::rectify(const cv::Mat& stereoImage, cv::Mat& rectifyLeftImage, cv::Mat& rectifiedRightImage)
{
	if (impl->m_rectifyMap.empty()) { log_error("Stereo epipolar rectify params is empty, please load it first."); return; }
	if (stereoImage.empty()) { log_error("Left or Right image is empty."); return; }
	const int W = impl->m_imgsz.width, H = impl->m_imgsz.height;
	if (impl->c00.rows == H && impl->c00.cols == W && stereoImage.cols >= 2 * W && stereoImage.rows >= H && stereoImage.type() == CV_8UC3) {
		impl->ctx.ensure(default_device());
		cv::Mat l(H, W, CV_8UC3), r(H, W, CV_8UC3);
		int rc = tsm_rectify_stereo(impl->ctx.h, stereoImage.data, stereoImage.step, H, W, impl->c00.data, impl->c01.data,
			impl->c10.data, impl->c11.data, impl->kind, impl->gen, l.data, l.step, r.data, r.step);
		if (rc != TSM_OK) throw std::runtime_error(tsm_last_error(impl->ctx.h));
		rectifyLeftImage = l; rectifiedRightImage = r;
		return;
	}
	cv::Mat left = stereoImage(cv::Rect(0, 0, W, H)).clone();
	cv::Mat right = stereoImage(cv::Rect(W, 0, W, H)).clone();
	rectify(left, right, rectifyLeftImage, rectifiedRightImage);
}

void stereo::EpipolarRectify::rectify(const cv::Mat& leftImage, const cv::Mat& rightImage, cv::Mat& rectifyLeftImage, cv::Mat& rectifiedRightImage)
{
	if (impl->m_rectifyMap.empty()) { log_error("Stereo epipolar rectify params is empty, please load it first."); return; }
	if (leftImage.empty() || rightImage.empty()) { log_error("Left or Right image is empty."); return; }
	impl->ctx.ensure(default_device());
	const cv::Mat* src[2] = {&leftImage, &rightImage};
	const cv::Mat* m1[2] = {&impl->c00, &impl->c10};
	const cv::Mat* m2[2] = {&impl->c01, &impl->c11};
	cv::Mat out[2];
	for (int k = 0; k < 2; ++k) {
		out[k].create(m1[k]->rows, m1[k]->cols, CV_8UC3);
		int rc = tsm_remap(impl->ctx.h, src[k]->data, src[k]->step, src[k]->rows, src[k]->cols, m1[k]->data, m2[k]->data,
			impl->kind, impl->gen, m1[k]->rows, m1[k]->cols, out[k].data, out[k].step);
		if (rc != TSM_OK) throw std::runtime_error(tsm_last_error(impl->ctx.h));
	}
	rectifyLeftImage = out[0]; rectifiedRightImage = out[1];
}

// ---------------------------------------------------------------- StereoMatching
stereo::StereoMatching::~StereoMatching() {}

// ---------------------------------------------------------------------- ADCensus
class stereo::ADCensus::ADCensusImpl
{
public:
	// defaults of ADCensusImpl::ADCensusImpl, source/ADCensus.cpp:409-420
	int m_minDisparity = 0, m_maxDisparity = 64;
	ADCensusParams m_paMatching{ColorModel::HSI};
	ColorModel m_colorModel = ColorModel::HSI;
	bool m_roiMatching = false, m_maskMatching = false;
	int m_offset = 0;
	int device = default_device();  // -1: the batched compute shards over every visible device
	Ctx ctx[1];                     // single-pair compute / fused rectify
	static constexpr size_t kInFlight = 4;  // pairs in flight per device of the batched compute (measured: 2 -> 17.1, 3 -> 16.7, 4 -> 16.5, 6 -> 16.5 ms per 1080p pair)
	std::vector<Ctx> batch;         // [device][kInFlight] contexts of the batched compute
	tsm_adcensus_config config() const
	{
		tsm_adcensus_config c;
		c.min_disparity = m_minDisparity; c.max_disparity = m_maxDisparity;
		c.color_model = m_colorModel == ColorModel::RGB ? TSM_COLOR_RGB : TSM_COLOR_HSI;
		c.roi_matching = m_roiMatching; c.mask_matching = m_maskMatching; c.offset = m_offset;
		return c;
	}
};

stereo::ADCensus::ADCensus() { impl = std::make_unique<ADCensusImpl>(); }
stereo::ADCensus::~ADCensus() {}

void stereo::ADCensus::setMinMaxDisparity(const int& minDisparity, const int& maxDisparity)
{
	if (minDisparity * maxDisparity < 0 or minDisparity >= maxDisparity)
		throw(std::string("[ADCensus] Set MinMaxDisparity error."));
	impl->m_minDisparity = minDisparity;
	impl->m_maxDisparity = maxDisparity;
}

void stereo::ADCensus::setMatchingStrategy(const ColorModel& colorModel, const bool& roiMatching, const bool& maskMatching)
{
	impl->m_colorModel = colorModel;
	impl->m_paMatching = ADCensusParams(colorModel);
	impl->m_roiMatching = roiMatching;
	impl->m_maskMatching = maskMatching;
}

void stereo::ADCensus::setOffset(const int& offset)
{
	if (offset < 0) throw(std::string("[ADCensus] Offset must be positive."));
	impl->m_offset = offset;
}

void stereo::ADCensus::setDevice(const int& device)
{
	if (device < -1) throw(std::string("[ADCensus] Device error."));
	impl->device = device;
}

static void check_pair(const cv::Mat& l, const cv::Mat& r)
{
	if (l.empty() or r.empty() or l.size() != r.size()) throw(std::string("[ADCensus] Image error."));
	// stricter than the reference (which reads any Mat through at<Vec3b>, UB for other types)
	if (l.type() != CV_8UC3 or r.type() != CV_8UC3) throw(std::string("[ADCensus] Image error."));
}

void stereo::ADCensus::compute(const cv::Mat& leftImage, const cv::Mat& rightImage, cv::Mat& disparity)
{
	check_pair(leftImage, rightImage);
	log_info("Computing disparity...");
	auto start = std::chrono::steady_clock::now();
	impl->ctx[0].ensure(impl->device < 0 ? default_device() : impl->device);
	cv::Mat out(leftImage.rows, leftImage.cols, CV_32FC1);
	const tsm_adcensus_config cfg = impl->config();
	int rc = tsm_adcensus_compute(impl->ctx[0].h, &cfg, leftImage.data, leftImage.step, rightImage.data, rightImage.step,
		leftImage.rows, leftImage.cols, (float*)out.data, out.step);
	if (rc == TSM_E_ARG) throw(std::string(tsm_last_error(impl->ctx[0].h)));
	if (rc != TSM_OK) throw std::runtime_error(tsm_last_error(impl->ctx[0].h));
	disparity = out;
	auto tt = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - start);
	log_info("Disparity map computed. Timing: " + std::to_string(tt.count() / 1000.0) + " ms.");
}

// Batched form: stereo pairs are independent, so pair i goes to device i % N (the frame
// sharding bench.py measures), one worker thread per device, kInFlight contexts (= streams, arenas) per device so
// that the copies and the small refinement kernels of one pair overlap with the bandwidth kernels of another.
// setDevice(d >= 0): that one device; setDevice(-1): every visible device.  Results do not depend on the sharding:
// every pair runs the same kernels on its own context.
void stereo::ADCensus::compute(const std::vector<cv::Mat>& leftImages, const std::vector<cv::Mat>& rightImages, std::vector<cv::Mat>& disparities)
{
	if (leftImages.size() != rightImages.size()) throw(std::string("[ADCensus] Image error."));
	const size_t n = leftImages.size();
	for (size_t i = 0; i < n; ++i) check_pair(leftImages[i], rightImages[i]);
	disparities.assign(n, cv::Mat());
	if (n == 0) return;
	const tsm_adcensus_config cfg = impl->config();
	std::vector<int> devices;
	if (impl->device >= 0) devices.push_back(impl->device);
	else {
		int count = 0;
		if (tsm_device_count(&count) != TSM_OK or count < 1) throw std::runtime_error(tsm_last_error(nullptr));
		for (int d = 0; d < count and (size_t)d < n; ++d) devices.push_back(d);
	}
	const size_t nd = devices.size();
	if (impl->batch.size() != nd * ADCensusImpl::kInFlight) {
		impl->batch.clear();
		impl->batch.resize(nd * ADCensusImpl::kInFlight);
	}
	std::mutex err_mtx;
	std::exception_ptr err;
	std::string err_str;  // the reference throws std::string for argument errors
	auto worker = [&](size_t w) {
		Ctx* ctx = &impl->batch[w * ADCensusImpl::kInFlight];
		constexpr size_t K = ADCensusImpl::kInFlight;
		std::vector<size_t> mine;
		for (size_t i = w; i < n; i += nd) mine.push_back(i);
		bool pending[K] = {};
		size_t slot_pair[K] = {};
		auto drain = [&] {  // error path: leave no context with an un-waited pair
			for (size_t k = 0; k < K; ++k)
				if (pending[k]) {
					cv::Mat scratch(leftImages[slot_pair[k]].rows, leftImages[slot_pair[k]].cols, CV_32FC1);
					tsm_adcensus_wait(ctx[k].h, (float*)scratch.data, scratch.step);
					pending[k] = false;
				}
		};
		try {
			for (size_t k = 0; k < K; ++k) ctx[k].ensure(devices[w]);
			for (size_t j = 0; j < mine.size() + K; ++j) {
				const size_t k = j % K;
				if (pending[k]) {
					const size_t i = slot_pair[k];
					cv::Mat out(leftImages[i].rows, leftImages[i].cols, CV_32FC1);
					pending[k] = false;
					int rc = tsm_adcensus_wait(ctx[k].h, (float*)out.data, out.step);
					if (rc != TSM_OK) throw std::make_pair(rc, std::string(tsm_last_error(ctx[k].h)));
					disparities[i] = out;
				}
				if (j < mine.size()) {
					const size_t i = mine[j];
					int rc = tsm_adcensus_enqueue(ctx[k].h, &cfg, leftImages[i].data, leftImages[i].step, rightImages[i].data,
						rightImages[i].step, leftImages[i].rows, leftImages[i].cols);
					if (rc != TSM_OK) throw std::make_pair(rc, std::string(tsm_last_error(ctx[k].h)));
					pending[k] = true;
					slot_pair[k] = i;
				}
			}
		} catch (const std::pair<int, std::string>& e) {
			drain();
			std::lock_guard<std::mutex> g(err_mtx);
			if (!err and err_str.empty()) {
				if (e.first == TSM_E_ARG) err_str = e.second;
				else err = std::make_exception_ptr(std::runtime_error(e.second));
			}
		} catch (...) {
			drain();
			std::lock_guard<std::mutex> g(err_mtx);
			if (!err and err_str.empty()) err = std::current_exception();
		}
	};
	if (nd == 1) worker(0);
	else {
		std::vector<std::thread> pool;
		for (size_t w = 0; w < nd; ++w) pool.emplace_back(worker, w);
		for (auto& t : pool) t.join();
	}
	if (!err_str.empty()) throw(std::string(err_str));
	if (err) std::rethrow_exception(err);
}

void stereo::ADCensus::compute(EpipolarRectify& rectify, const cv::Mat& stereoImage, cv::Mat& disparity)
{
	auto& r = *rectify.impl;
	if (r.m_rectifyMap.empty()) throw std::runtime_error("stereo params is empty, please load it first");
	if (stereoImage.empty() or stereoImage.type() != CV_8UC3) throw(std::string("[ADCensus] Image error."));
	const int W = r.m_imgsz.width, H = r.m_imgsz.height;
	if (stereoImage.cols < 2 * W or stereoImage.rows < H or r.c00.rows != H or r.c00.cols != W)
		throw(std::string("[ADCensus] Image error."));
	impl->ctx[0].ensure(impl->device < 0 ? default_device() : impl->device);
	cv::Mat out(H, W, CV_32FC1);
	const tsm_adcensus_config cfg = impl->config();
	int rc = tsm_rectify_adcensus(impl->ctx[0].h, &cfg, stereoImage.data, stereoImage.step, H, W, r.c00.data, r.c01.data, r.c10.data,
		r.c11.data, r.kind, r.gen, (float*)out.data, out.step);
	if (rc == TSM_E_ARG) throw(std::string(tsm_last_error(impl->ctx[0].h)));
	if (rc != TSM_OK) throw std::runtime_error(tsm_last_error(impl->ctx[0].h));
	disparity = out;
}

// ---- consumers of the disparity map -------------------------------------------
namespace {
// One library context for the free functions (the reference's are stateless); device = env TSM_DEVICE or 0.
tsm_ctx* consumer_ctx()
{
	static tsm_ctx* ctx = [] {
		tsm_ctx* c = nullptr;
		const char* e = std::getenv("TSM_DEVICE");
		if (tsm_create(e ? std::atoi(e) : 0, &c) != TSM_OK) throw std::runtime_error(tsm_last_error(nullptr));
		return c;
	}();
	return ctx;
}
void check_disparity(const cv::Mat& m, const char* who)
{
	if (m.empty() or m.type() != CV_32FC1) throw std::runtime_error(std::string(who) + ": disparity must be a non-empty CV_32FC1 map");
}
void consumer_check(tsm_ctx* c, int rc)
{
	if (rc != TSM_OK) throw std::runtime_error(tsm_last_error(c));
}
}

cv::Mat stereo::JETColorMap()
{
	cv::Mat colorMap(1, 256, CV_8UC3);
	tsm_jet_colormap(colorMap.data);
	return colorMap;
}

static void apply_color_map(const cv::Mat& src, cv::Mat& dst, bool autoRange, float minVal, float maxVal, const cv::Mat& colorMap)
{
	check_disparity(src, "applyColorMap");
	if (colorMap.empty() or colorMap.type() != CV_8UC3 or colorMap.rows * colorMap.cols != 256 or !colorMap.isContinuous())
		throw std::runtime_error("applyColorMap: colorMap must be a continuous 1x256 CV_8UC3 table");
	cv::Mat out(src.rows, src.cols, CV_8UC3);
	tsm_ctx* c = consumer_ctx();
	consumer_check(c, tsm_apply_colormap(c, (const float*)src.data, src.step, src.rows, src.cols, autoRange ? 1 : 0, minVal, maxVal,
		colorMap.data, out.data, out.step));
	dst = out;
}

void stereo::applyColorMap(const cv::Mat& src, cv::Mat& dst, const cv::Mat& colorMap)
{
	apply_color_map(src, dst, true, 0.f, 0.f, colorMap);
}

void stereo::applyColorMap(const cv::Mat& src, cv::Mat& dst, float minVal, float maxVal, const cv::Mat& colorMap)
{
	apply_color_map(src, dst, false, minVal, maxVal, colorMap);
}

static void write_cloud(const cv::Mat& RGBImage, const cv::Mat& XYZPoints, const std::string& path, int format)
{
	if (RGBImage.empty() or XYZPoints.empty() or path.empty()) { log_error("Empty input."); return; }  // stereo.cpp:252-256
	if (RGBImage.type() != CV_8UC3 or XYZPoints.type() != CV_32FC3 or RGBImage.size() != XYZPoints.size())
		throw std::runtime_error("writePointCloud: RGBImage must be CV_8UC3 and XYZPoints CV_32FC3 of the same size");
	size_t n = 0;
	if (tsm_write_point_cloud(RGBImage.data, RGBImage.step, (const float*)XYZPoints.data, XYZPoints.step, RGBImage.rows,
		RGBImage.cols, path.c_str(), format, &n) != TSM_OK)
		throw std::runtime_error(tsm_last_error(nullptr));
	log_info("Write Done. Points: " + std::to_string(n) + ".");
}

void stereo::writePointCloudToPCD(const cv::Mat& RGBImage, const cv::Mat& XYZPoints, const std::string& pcdPath)
{
	write_cloud(RGBImage, XYZPoints, pcdPath, TSM_CLOUD_PCD);
}

void stereo::writePointCloudToPLY(const cv::Mat& RGBImage, const cv::Mat& XYZPoints, const std::string& plyPath)
{
	write_cloud(RGBImage, XYZPoints, plyPath, TSM_CLOUD_PLY);
}

void stereo::reprojectToDepth(const cv::Mat& disparity, float focalLength, float baseline, cv::Mat& depth)
{
	check_disparity(disparity, "reprojectToDepth");
	cv::Mat out(disparity.rows, disparity.cols, CV_32FC1);
	tsm_ctx* c = consumer_ctx();
	consumer_check(c, tsm_reproject_to_depth(c, (const float*)disparity.data, disparity.step, disparity.rows, disparity.cols,
		focalLength, baseline, (float*)out.data, out.step));
	depth = out;
}

void stereo::reprojectTo3D(const cv::Mat& disparity, float focalLength, float baseline, float cx, float cy, cv::Mat& XYZPoints)
{
	check_disparity(disparity, "reprojectTo3D");
	cv::Mat out(disparity.rows, disparity.cols, CV_32FC3);
	tsm_ctx* c = consumer_ctx();
	consumer_check(c, tsm_reproject_to_3d(c, (const float*)disparity.data, disparity.step, disparity.rows, disparity.cols,
		focalLength, baseline, cx, cy, (float*)out.data, out.step));
	XYZPoints = out;
}

void stereo::reprojectTo3D(const cv::Mat& disparity, const cv::Mat& Q, cv::Mat& XYZPoints)
{
	check_disparity(disparity, "reprojectTo3D");
	if (Q.rows != 4 or Q.cols != 4 or (Q.type() != CV_64FC1 and Q.type() != CV_32FC1))
		throw std::runtime_error("reprojectTo3D: Q must be a 4x4 CV_64FC1 or CV_32FC1 matrix");
	double q[16];
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j)
			q[4 * i + j] = Q.type() == CV_64FC1 ? ((const double*)(Q.data + (size_t)i * Q.step))[j]
				: (double)((const float*)(Q.data + (size_t)i * Q.step))[j];
	cv::Mat out(disparity.rows, disparity.cols, CV_32FC3);
	tsm_ctx* c = consumer_ctx();
	consumer_check(c, tsm_reproject_to_3d_q(c, (const float*)disparity.data, disparity.step, disparity.rows, disparity.cols, q,
		(float*)out.data, out.step));
	XYZPoints = out;
}

// ---- rectify-map generation and calibration files ---------------------------
namespace {
// The subset of OpenCV FileStorage YAML the reference reads: "key: !!opencv-matrix" blocks and "key: [ a, b ]".
struct YamlDoc {
	std::map<std::string, cv::Mat> mats;
	std::map<std::string, std::vector<double>> seqs;
};

bool parse_opencv_yaml(const std::string& path, YamlDoc& doc)
{
	std::ifstream f(path);
	if (!f) return false;
	std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	size_t pos = 0;
	auto number_list = [&](size_t open) {  // parses "[ a, b, ... ]" starting at text[open] == '['
		std::vector<double> v;
		size_t close = text.find(']', open);
		std::string body = text.substr(open + 1, close == std::string::npos ? std::string::npos : close - open - 1);
		for (char& ch : body) if (ch == ',' or ch == '\n' or ch == '\r') ch = ' ';
		std::istringstream is(body);
		double x;
		while (is >> x) v.push_back(x);
		pos = close == std::string::npos ? text.size() : close + 1;
		return v;
	};
	while (pos < text.size()) {
		size_t eol = text.find('\n', pos);
		if (eol == std::string::npos) eol = text.size();
		std::string line = text.substr(pos, eol - pos);
		size_t colon = line.find(':');
		if (line.empty() or line[0] == ' ' or line[0] == '%' or line[0] == '-' or colon == std::string::npos) { pos = eol + 1; continue; }
		const std::string key = line.substr(0, colon);
		if (line.find("!!opencv-matrix") != std::string::npos) {
			int rows = 0, cols = 0;
			std::string dt = "d";
			size_t p = eol + 1;
			for (int k = 0; k < 3; ++k) {  // rows / cols / dt lines
				size_t e = text.find('\n', p);
				std::string l = text.substr(p, e - p);
				if (l.find("rows:") != std::string::npos) rows = std::atoi(l.c_str() + l.find(':') + 1);
				else if (l.find("cols:") != std::string::npos) cols = std::atoi(l.c_str() + l.find(':') + 1);
				else if (l.find("dt:") != std::string::npos) { dt = l.substr(l.find(':') + 1); dt.erase(std::remove_if(dt.begin(), dt.end(), [](char c) { return c == ' ' or c == '"' or c == '\r'; }), dt.end()); }
				p = e + 1;
			}
			size_t open = text.find('[', p);
			if (open == std::string::npos) return false;
			std::vector<double> v = number_list(open);
			const char kind = dt.empty() ? 'd' : dt.back();
			const int ch = dt.size() > 1 ? std::atoi(dt.c_str()) : 1;
			int type = kind == 'd' ? CV_64FC1 : kind == 'f' ? CV_32FC1 : kind == 's' and ch == 2 ? CV_16SC2 : kind == 'w' ? CV_16UC1 : -1;
			if (type < 0 or (size_t)rows * cols * ch != v.size()) return false;
			cv::Mat m(rows, cols, type);
			for (size_t i = 0; i < v.size(); ++i) {
				if (type == CV_64FC1) ((double*)m.data)[i] = v[i];
				else if (type == CV_32FC1) ((float*)m.data)[i] = (float)v[i];
				else if (type == CV_16SC2) ((int16_t*)m.data)[i] = (int16_t)v[i];
				else ((uint16_t*)m.data)[i] = (uint16_t)v[i];
			}
			doc.mats[key] = m;
		} else if (line.find('[', colon) != std::string::npos) {
			doc.seqs[key] = number_list(pos + line.find('[', colon));
		} else {
			pos = eol + 1;
		}
	}
	return true;
}

cv::Mat yaml_mat(const YamlDoc& d, const char* key)
{
	auto it = d.mats.find(key);
	return it == d.mats.end() ? cv::Mat() : it->second;
}

void to_doubles(const cv::Mat& m, std::vector<double>& out)
{
	out.clear();
	for (int r = 0; r < m.rows; ++r)
		for (int c = 0; c < m.cols; ++c)
			out.push_back(m.type() == CV_64FC1 ? ((const double*)(m.data + (size_t)r * m.step))[c] : (double)((const float*)(m.data + (size_t)r * m.step))[c]);
}

void undistort_rectify_map(const cv::Mat& K, const cv::Mat& D, const cv::Mat& R, const cv::Mat& P, const cv::Size& sz, cv::Mat& m1, cv::Mat& m2)
{
	std::vector<double> k, d, r, p;
	to_doubles(K, k);
	to_doubles(D, d);
	to_doubles(R, r);
	to_doubles(P, p);
	if (k.size() != 9 or (!r.empty() and r.size() != 9) or (!p.empty() and p.size() != 9 and p.size() != 12))
		throw std::runtime_error("EpipolarRectifyMap::compute: K and R must be 3x3, P 3x3 or 3x4");
	cv::Mat a(sz.height, sz.width, CV_16SC2), b(sz.height, sz.width, CV_16UC1);
	tsm_ctx* c = consumer_ctx();
	consumer_check(c, tsm_init_undistort_rectify_map(c, k.data(), d.empty() ? nullptr : d.data(), (int)d.size(), r.empty() ? nullptr : r.data(),
		p.empty() ? nullptr : p.data(), p.empty() ? 0 : (int)p.size() / 3, sz.height, sz.width, (int16_t*)a.data, a.step, (uint16_t*)b.data, b.step));
	m1 = a;
	m2 = b;
}
}

void stereo::EpipolarRectifyMap::compute(const StereoPair<CameraIntrinsic>& intrinsic, const cv::Size& imgsz)
{
	if (intrinsic.left.empty() or intrinsic.right.empty()) return;  // stereo_utils.cpp:159-160
	undistort_rectify_map(intrinsic.left.intrinsic_matrix, intrinsic.left.distortion_coefficients, R1, P1, imgsz, map00, map01);
	undistort_rectify_map(intrinsic.right.intrinsic_matrix, intrinsic.right.distortion_coefficients, R2, P2, imgsz, map10, map11);
}

void stereo::EpipolarRectifyMap::loadRectifyMapsYMLFile(const std::string& ymlFilePath)
{
	YamlDoc doc;
	if (!parse_opencv_yaml(ymlFilePath, doc)) throw std::runtime_error("Cannot open rectify maps yml file.");  // :138-145
	map00 = yaml_mat(doc, "map00");
	map01 = yaml_mat(doc, "map01");
	map10 = yaml_mat(doc, "map10");
	map11 = yaml_mat(doc, "map11");
	if (empty()) std::fprintf(stderr, "[ERROR] Cannot load rectify maps from yml file.\n");
}

void stereo::StereoParams::loadYAMLFile(const std::string& ymlFilePath)
{
	if (ymlFilePath.empty()) throw std::invalid_argument("Stereo YAML file path is empty.");  // :187-192
	YamlDoc doc;
	if (!parse_opencv_yaml(ymlFilePath, doc)) throw std::runtime_error("Cannot open stereo yml file.");  // :197-203
	intrinsic.left = CameraIntrinsic(yaml_mat(doc, "leftK"), yaml_mat(doc, "leftD"));
	intrinsic.right = CameraIntrinsic(yaml_mat(doc, "rightK"), yaml_mat(doc, "rightD"));
	extrinsic.E = yaml_mat(doc, "E");
	extrinsic.F = yaml_mat(doc, "F");
	extrinsic.R = yaml_mat(doc, "R");
	extrinsic.T = yaml_mat(doc, "T");
	map.R1 = yaml_mat(doc, "R1");
	map.R2 = yaml_mat(doc, "R2");
	map.P1 = yaml_mat(doc, "P1");
	map.P2 = yaml_mat(doc, "P2");
	Q = yaml_mat(doc, "Q");
	auto sz = doc.seqs.find("imgsz");
	if (sz != doc.seqs.end() and sz->second.size() == 2) imgsz = cv::Size((int)sz->second[0], (int)sz->second[1]);
	if (Q.empty()) return;
	auto q = [&](int r, int c) { return ((const double*)(Q.data + (size_t)r * Q.step))[c]; };
	rectified_f = static_cast<float>(q(2, 3));   // :222-225
	rectified_cx = static_cast<float>(-q(0, 3));
	rectified_cy = static_cast<float>(-q(1, 3));
	baseline = 1.f / static_cast<float>(q(3, 2));
	map.compute(intrinsic, imgsz);
}

bool stereo::StereoParams::empty() const
{
	return intrinsic.left.empty() or intrinsic.right.empty() or extrinsic.empty() or map.empty() or Q.empty();
}
