// Exercises the C++ facade exactly like the reference's README demo (README.md:174-192):
//   stereo::ADCensus adcensus; setMatchingStrategy(RGB); setMinMaxDisparity(0, D); compute(l, r, disp)
// Modes:
//   api                     -- error behaviour only (no GPU needed)
//   run in.bin out.bin      -- in.bin = int32 H, W, D + left BGR + right BGR; out.bin = float32 H*W
//   batch in.bin out.bin n  -- the batched overload on n copies of the pair
//   params calib.yml out.bin -- stereo::StereoParams(calib.yml): out.bin = map00, map01, map10, map11, then f, cx, cy, B
//   consume in.bin out.bin  -- disparity, then the README demo's consumers (README.md demo 5): out.bin = disparity,
//                              reprojectToDepth (f=700,B=0.1), reprojectTo3D(f,B,cx,cy), reprojectTo3D(Q), applyColorMap(JET)
#include "../../tea_stereo_matching_b200/cpp/stereo.h"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>

static int api_checks()
{
    int fails = 0;
    stereo::ADCensus m;
    auto expect_string = [&](auto&& fn, const char* what) {
        try { fn(); std::printf("FAIL %s: no throw\n", what); ++fails; }
        catch (const std::string& e) { std::printf("ok   %s -> std::string(\"%s\")\n", what, e.c_str()); }
        catch (...) { std::printf("FAIL %s: wrong exception type\n", what); ++fails; }
    };
    expect_string([&] { m.setMinMaxDisparity(5, 5); }, "setMinMaxDisparity(5,5)");
    expect_string([&] { m.setMinMaxDisparity(-2, 7); }, "setMinMaxDisparity(-2,7)");
    expect_string([&] { m.setOffset(-1); }, "setOffset(-1)");
    cv::Mat a, b, d;
    expect_string([&] { m.compute(a, b, d); }, "compute(empty, empty)");
    cv::Mat l(8, 8, CV_8UC3), r(8, 9, CV_8UC3);
    expect_string([&] { m.compute(l, r, d); }, "compute(size mismatch)");
    stereo::StereoMatching* base = &m;  // ADCensus is-a StereoMatching
    (void)base;
    stereo::EpipolarRectify rect;
    cv::Mat out;
    rect.rectify(l, out);  // logs + returns, output untouched
    if (!out.empty()) { std::printf("FAIL rectify without maps touched the output\n"); ++fails; }
    try { rect.loadEpipolarRectifyMap(stereo::EpipolarRectifyMap(), cv::Size(8, 8)); std::printf("FAIL load empty maps\n"); ++fails; }
    catch (const std::runtime_error& e) { std::printf("ok   loadEpipolarRectifyMap(empty) -> runtime_error(\"%s\")\n", e.what()); }
    stereo::ADCensusParams p(stereo::ColorModel::RGB);
    if (p.maxLength1 != 34 || p.colorThresh1 != 20 || p.votingThresh != 20) { std::printf("FAIL params\n"); ++fails; }
    std::printf(fails ? "API CHECKS FAILED\n" : "API CHECKS PASSED\n");
    return fails;
}

int main(int argc, char** argv)
{
    if (argc >= 2 && !std::strcmp(argv[1], "api")) return api_checks();
    if (argc >= 4 && !std::strcmp(argv[1], "params")) {
        try {
            stereo::StereoParams sp(argv[2]);
            if (sp.empty()) { std::fprintf(stderr, "StereoParams is empty\n"); return 5; }
            std::ofstream o(argv[3], std::ios::binary);
            const int H = sp.imgsz.height, W = sp.imgsz.width;
            o.write((const char*)sp.map.map00.data, (std::streamsize)H * W * 4);
            o.write((const char*)sp.map.map01.data, (std::streamsize)H * W * 2);
            o.write((const char*)sp.map.map10.data, (std::streamsize)H * W * 4);
            o.write((const char*)sp.map.map11.data, (std::streamsize)H * W * 2);
            const float tail[4] = {sp.rectified_f, sp.rectified_cx, sp.rectified_cy, sp.baseline};
            o.write((const char*)tail, sizeof tail);
        } catch (const std::exception& e) { std::fprintf(stderr, "exception: %s\n", e.what()); return 4; }
        return 0;
    }
    if (argc < 4) { std::fprintf(stderr, "usage: %s api | run in out | batch in out n | batchall in out n\n", argv[0]); return 2; }
    std::ifstream f(argv[2], std::ios::binary);
    int32_t hdr[3];
    f.read((char*)hdr, sizeof hdr);
    const int H = hdr[0], W = hdr[1], D = hdr[2];
    cv::Mat left(H, W, CV_8UC3), right(H, W, CV_8UC3);
    f.read((char*)left.data, (std::streamsize)H * W * 3);
    f.read((char*)right.data, (std::streamsize)H * W * 3);
    try {
        stereo::ADCensus adcensus;
        adcensus.setMatchingStrategy(stereo::ColorModel::RGB, false, false);
        adcensus.setMinMaxDisparity(0, D);
        std::ofstream o(argv[3], std::ios::binary);
        if (!std::strcmp(argv[1], "batch") || !std::strcmp(argv[1], "batchall")) {
            const int n = argc > 4 ? std::atoi(argv[4]) : 3;
            if (!std::strcmp(argv[1], "batchall")) adcensus.setDevice(-1);  // shard the pairs over every visible device
            std::vector<cv::Mat> ls(n, left), rs(n, right), ds;
            adcensus.compute(ls, rs, ds);
            for (auto& d : ds) o.write((const char*)d.data, (std::streamsize)H * W * 4);
        } else if (!std::strcmp(argv[1], "consume")) {
            cv::Mat disparity, depth, xyz, xyzq, color;
            adcensus.compute(left, right, disparity);
            stereo::reprojectToDepth(disparity, 700.f, 0.1f, depth);
            stereo::reprojectTo3D(disparity, 700.f, 0.1f, W / 2.f, H / 2.f, xyz);
            cv::Mat Q(4, 4, CV_64FC1);
            const double q[16] = {1, 0, 0, -W / 2.0, 0, 1, 0, -H / 2.0, 0, 0, 0, 700.0, 0, 0, 10.0, 0.5};
            std::memcpy(Q.data, q, sizeof q);
            stereo::reprojectTo3D(disparity, Q, xyzq);
            stereo::applyColorMap(disparity, color, stereo::JETColorMap());
            o.write((const char*)disparity.data, (std::streamsize)H * W * 4);
            o.write((const char*)depth.data, (std::streamsize)H * W * 4);
            o.write((const char*)xyz.data, (std::streamsize)H * W * 12);
            o.write((const char*)xyzq.data, (std::streamsize)H * W * 12);
            o.write((const char*)color.data, (std::streamsize)H * W * 3);
        } else {
            cv::Mat disparity;
            adcensus.compute(left, right, disparity);
            o.write((const char*)disparity.data, (std::streamsize)H * W * 4);
        }
    } catch (const std::string& e) { std::fprintf(stderr, "std::string: %s\n", e.c_str()); return 3; }
    catch (const std::exception& e) { std::fprintf(stderr, "exception: %s\n", e.what()); return 4; }
    return 0;
}
