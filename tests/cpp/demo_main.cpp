// demo_main.cpp -- the reference demo's call order (src/ImageCompressor/main.cpp:16-82) against the drop-in facade:
// construct -> buildCompressEnv -> compress x2 -> deleteCompressEnv -> buildDecodeEnv -> save x2 -> decode x2 ->
// deleteDecodeEnv -> delete. Images come from raw BGR files written by the test (no imread in this image).
// usage: demo W H in1.bgr in2.bgr outdir [ngpus]     exit code 0 = every run_state was 1
// ngpus > 1: compress() runs as MCU-row strips over that many contexts of this one process (B2J_MULTI_DEVICES picks the GPUs)
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/ImageCompressor.h"

static void printState(int run_state) { std::cout << (run_state == 1 ? "[INFO] Successful." : "[INFO] Failed.") << std::endl; }

static cv::Mat load(const char *path, int W, int H) {
    cv::Mat m(H, W, CV_8UC3);
    FILE *f = fopen(path, "rb");
    if (!f || fread(m.data, 1, (size_t)W * H * 3, f) != (size_t)W * H * 3) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return m;
}

int main(int argc, char *argv[]) {
    if (argc < 6) return 2;
    const int W = atoi(argv[1]), H = atoi(argv[2]);
    const std::string outdir = argv[5];
    int bad = 0;
    const int ngpus = argc > 6 ? atoi(argv[6]) : 1;
    NvjpegCompressRunner *compressor = new NvjpegCompressRunner(W, H, 95, true, 422, ngpus);
    int compress_run_state, decode_run_state;
    compressor->buildCompressEnv();
    cv::Mat image1 = load(argv[3], W, H);
    std::vector<unsigned char> obuffer1 = compressor->compress(image1, &compress_run_state);
    printState(compress_run_state); bad += compress_run_state != 1;
    cv::Mat image2 = load(argv[4], W, H);
    std::vector<unsigned char> obuffer2 = compressor->compress(image2, &compress_run_state);
    printState(compress_run_state); bad += compress_run_state != 1;
    // README.md:8 extras while the encoder is alive
    std::vector<unsigned char> diffjpg; cv::Mat rec; double ps = 0; int st = 0;
    std::vector<unsigned char> again = compressor->secondaryCompress(image1, &diffjpg, &rec, &ps, true, &st);
    bad += st != 1 || again != obuffer1 || diffjpg.empty();
    std::cout << "[INFO] secondary: psnr " << ps << " dB, diff jpeg " << diffjpg.size() << " bytes" << std::endl;
    compressor->deleteCompressEnv();

    compressor->buildDecodeEnv();
    const std::string p1 = outdir + "/1.jpeg", p2 = outdir + "/2.jpeg";
    compressor->save(p1, obuffer1);
    compressor->save(p2, obuffer2);
    cv::Mat d1 = compressor->decode(p1, &decode_run_state);
    printState(decode_run_state); bad += decode_run_state != 1;
    cv::Mat d2 = compressor->decode(p2, &decode_run_state);
    printState(decode_run_state); bad += decode_run_state != 1;
    cv::Mat none = compressor->decode(outdir + "/missing.jpeg", &decode_run_state);
    bad += decode_run_state != 0 || !none.empty();
    FILE *f = fopen((outdir + "/1.dec.bgr").c_str(), "wb"); fwrite(d1.data, 1, d1.total() * 3, f); fclose(f);
    f = fopen((outdir + "/2.dec.bgr").c_str(), "wb"); fwrite(d2.data, 1, d2.total() * 3, f); fclose(f);
    double p = compressor->psnr(image1, d1, &st); bad += st != 1;
    std::cout << "[INFO] psnr(image1, decode1) = " << p << std::endl;
    bad += (p != ps);
    compressor->deleteDecodeEnv();
    delete compressor;
    return bad ? 1 : EXIT_SUCCESS;
}
