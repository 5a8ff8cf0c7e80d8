// facade_bench.cpp -- NvjpegCompressRunner::compress / decode / secondaryCompress at the headline size through the C++
// facade, on the PAGEABLE memory a cv::Mat really owns: wall-clock per call as the reference
// prints it (ImageCompressor.cpp:58,85).   usage: facade_bench [W H [ngpus]]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/ImageCompressor.h"

int main(int argc, char *argv[]) {
    const int W = argc > 2 ? atoi(argv[1]) : 8320, H = argc > 2 ? atoi(argv[2]) : 40000, ngpus = argc > 3 ? atoi(argv[3]) : 1;
    cv::Mat img(H, W, CV_8UC3);
    uint32_t st = 12345;
    for (int y = 0; y < H; y++) {          // smooth gradients + +-8 noise, in the spirit of the synthetic workload
        unsigned char *r = img.data + (size_t)y * img.step;
        for (int x = 0; x < W; x++) {
            st = st * 1664525u + 1013904223u;
            const int n = (int)((st >> 24) & 15) - 8;
            const int b = ((x * 255) / W + n), g = ((y * 255) / H + n), rr = (((x + y) & 511) / 2 + n);
            r[3 * x] = (unsigned char)(b < 0 ? 0 : b > 255 ? 255 : b);
            r[3 * x + 1] = (unsigned char)(g < 0 ? 0 : g > 255 ? 255 : g);
            r[3 * x + 2] = (unsigned char)(rr < 0 ? 0 : rr > 255 ? 255 : rr);
        }
    }
    NvjpegCompressRunner runner(W, H, 95, true, 422, ngpus);
    runner.buildCompressEnv();
    runner.buildDecodeEnv();
    int ok = 0;
    std::vector<unsigned char> jpg;
    double best_c = 1e30, best_d = 1e30, best_s = 1e30;
    for (int i = 0; i < 4; i++) {
        auto t0 = std::chrono::steady_clock::now();
        jpg = runner.compress(img, &ok);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (!ok) { fprintf(stderr, "compress failed\n"); return 1; }
        if (i) best_c = ms < best_c ? ms : best_c;
    }
    for (int i = 0; i < 3; i++) {
        auto t0 = std::chrono::steady_clock::now();
        cv::Mat rec = runner.reconstruct(jpg, &ok);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (!ok) { fprintf(stderr, "reconstruct failed\n"); return 1; }
        if (i) best_d = ms < best_d ? ms : best_d;
    }
    for (int i = 0; i < 3; i++) {
        std::vector<unsigned char> dj;
        double ps = 0;
        auto t0 = std::chrono::steady_clock::now();
        std::vector<unsigned char> j1 = runner.secondaryCompress(img, &dj, nullptr, &ps, true, &ok);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (!ok || j1 != jpg) { fprintf(stderr, "secondaryCompress failed\n"); return 1; }
        if (i) best_s = ms < best_s ? ms : best_s;
    }
    printf("{\"facade\": \"NvjpegCompressRunner (C++, pageable cv::Mat / std::vector)\", \"W\": %d, \"H\": %d, \"ngpus\": %d, \"jpeg_bytes\": %zu, "
           "\"compress_ms\": %.2f, \"reconstruct_ms\": %.2f, \"secondaryCompress_ms\": %.2f}\n",
           W, H, ngpus, jpg.size(), best_c, best_d, best_s);
    return 0;
}
