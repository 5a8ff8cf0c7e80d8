// ref_nvjpeg.cu -- BASELINE HARNESS, NOT PRODUCT CODE. Never linked into libb2jpeg.so.
//
// The reference (OroChippw/Nvjpeg-ImageCompressor) cannot be compiled in this image: it includes
// <opencv2/opencv.hpp> (ImageCompressor.h:12, ImageCompressorImpl.cu:14; no OpenCV C++ SDK here) and uses
// MSVC-only __declspec/fopen_s (ImageCompressor.h:15, ImageCompressor.cpp:66). Its hot path is nothing but a
// fixed sequence of nvJPEG calls, so this file issues the SAME library calls in the SAME order on raw buffers:
//   env setup      = initCompressEnv        ImageCompressorImpl.cu:19-45
//   encode         = CompressWorker         :269-294  (cv::split -> 3x cudaMemcpy -> nvjpegEncodeImage on the NULL
//                                                       stream -> two-call nvjpegEncodeRetrieveBitstream)
//   decode env     = initDecodeEnv          :67-95
//   decode         = DecodeWorker           :311-385 + getCVImageOnCPU :184-232
// The only additions are the switches the README sweep needed by editing the source (encoding progressive|baseline,
// sampling factor; ImageCompressorImpl.cu:28,31) and the timers. Built by baseline/Makefile into baseline/_ref/.
#include <cuda_runtime.h>
#include <nvjpeg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#if defined(__SSSE3__)
#include <tmmintrin.h>
#endif
#include <vector>

#define RCK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { fprintf(stderr, "cuda error %d at %s:%d\n", (int)_e, __FILE__, __LINE__); return -2; } } while (0)
#define RNJ(x) do { nvjpegStatus_t _s = (x); if (_s != NVJPEG_STATUS_SUCCESS) { fprintf(stderr, "nvjpeg error %d at %s:%d\n", (int)_s, __FILE__, __LINE__); return -3; } } while (0)

struct RefCtx {
    int W, H, quality, optimize, css, progressive;
    nvjpegHandle_t h_enc = nullptr, h_dec = nullptr;
    nvjpegEncoderState_t enc_state = nullptr;
    nvjpegEncoderParams_t enc_params = nullptr;
    nvjpegImage_t input{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool enc_init = false, dec_init = false;
    // decode
    nvjpegJpegState_t dec_state = nullptr, decoupled = nullptr;
    nvjpegJpegDecoder_t decoder = nullptr;
    nvjpegDecodeParams_t dec_params = nullptr;
    nvjpegBufferPinned_t pinned[2] = {nullptr, nullptr};
    nvjpegJpegStream_t streams[2] = {nullptr, nullptr};
    nvjpegBufferDevice_t dev_buf = nullptr;
    nvjpegImage_t out_img{};
    size_t out_size[3] = {0, 0, 0};
    std::vector<unsigned char> planes[3];
    float last_gpu_ms = 0, last_split_ms = 0, last_h2d_ms = 0, last_retrieve_ms = 0;
};

static nvjpegChromaSubsampling_t css_of(int c) {
    switch (c) {
    case 0: return NVJPEG_CSS_444;
    case 1: return NVJPEG_CSS_422;
    case 2: return NVJPEG_CSS_440;
    case 3: return NVJPEG_CSS_420;
    default: return NVJPEG_CSS_411;
    }
}

extern "C" {

__attribute__((visibility("default"))) int ref_create(int W, int H, int quality, int optimize, int css, int progressive, void **out) {
    RefCtx *c = new RefCtx();
    c->W = W; c->H = H; c->quality = quality; c->optimize = optimize; c->css = css; c->progressive = progressive;
    *out = c;
    return 0;
}

__attribute__((visibility("default"))) int ref_build_compress_env(void *p) {
    RefCtx *c = (RefCtx *)p;
    RNJ(nvjpegCreate(NVJPEG_BACKEND_GPU_HYBRID, nullptr, &c->h_enc));
    RNJ(nvjpegEncoderParamsCreate(c->h_enc, &c->enc_params, NULL));
    RNJ(nvjpegEncoderStateCreate(c->h_enc, &c->enc_state, NULL));
    RCK(cudaEventCreate(&c->ev0));
    RCK(cudaEventCreate(&c->ev1));
    RNJ(nvjpegEncoderParamsSetEncoding(c->enc_params, c->progressive ? NVJPEG_ENCODING_PROGRESSIVE_DCT_HUFFMAN : NVJPEG_ENCODING_BASELINE_DCT, NULL));
    RNJ(nvjpegEncoderParamsSetOptimizedHuffman(c->enc_params, c->optimize, NULL));
    RNJ(nvjpegEncoderParamsSetQuality(c->enc_params, c->quality, NULL));
    RNJ(nvjpegEncoderParamsSetSamplingFactors(c->enc_params, css_of(c->css), NULL));
    const size_t sz = (size_t)c->W * c->H;
    for (int i = 0; i < 3; i++) {
        c->input.pitch[i] = c->W;
        RCK(cudaMalloc((void **)&c->input.channel[i], sz));
        c->planes[i].resize(sz);
    }
    c->enc_init = true;
    return 0;
}

// CompressWorker: host split -> 3 pageable H2D copies -> nvjpegEncodeImage (event bracket) -> retrieve
__attribute__((visibility("default"))) int ref_compress(void *p, const unsigned char *bgr, size_t step, unsigned char *out, size_t cap, size_t *len) {
    RefCtx *c = (RefCtx *)p;
    const int W = c->W, H = c->H;
    auto t0 = std::chrono::steady_clock::now();
    for (int y = 0; y < H; y++) {  // cv::split (OpenCV's is SIMD, one thread: byte shuffles here, 16 pixels per step)
        const unsigned char *row = bgr + (size_t)y * step;
        unsigned char *b = c->planes[0].data() + (size_t)y * W, *g = c->planes[1].data() + (size_t)y * W, *r = c->planes[2].data() + (size_t)y * W;
        int x = 0;
#if defined(__SSSE3__)
        const __m128i m0 = _mm_setr_epi8(0, 3, 6, 9, 12, 15, 1, 4, 7, 10, 13, 2, 5, 8, 11, 14);
        const __m128i m1 = _mm_setr_epi8(2, 5, 8, 11, 14, 0, 3, 6, 9, 12, 15, 1, 4, 7, 10, 13);
        const __m128i m2 = _mm_setr_epi8(1, 4, 7, 10, 13, 2, 5, 8, 11, 14, 0, 3, 6, 9, 12, 15);
        for (; x + 16 <= W; x += 16) {
            // s0 = [b0..b5 | g0..g4 | r0..r4], s1 = [b6..b10 | g5..g10 | r5..r9], s2 = [b11..b15 | g11..g15 | r10..r15]
            const __m128i s0 = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(row + 3 * x)), m0);
            const __m128i s1 = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(row + 3 * x + 16)), m1);
            const __m128i s2 = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(row + 3 * x + 32)), m2);
            const __m128i k5 = _mm_setr_epi8(-1, -1, -1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
            const __m128i k6 = _mm_setr_epi8(-1, -1, -1, -1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
            const __m128i bb = _mm_or_si128(_mm_or_si128(_mm_and_si128(s0, k6), _mm_slli_si128(_mm_and_si128(s1, k5), 6)),
                                            _mm_slli_si128(_mm_and_si128(s2, k5), 11));
            const __m128i gg = _mm_or_si128(_mm_or_si128(_mm_and_si128(_mm_srli_si128(s0, 6), k5),
                                                         _mm_slli_si128(_mm_and_si128(_mm_srli_si128(s1, 5), k6), 5)),
                                            _mm_slli_si128(_mm_and_si128(_mm_srli_si128(s2, 5), k5), 11));
            const __m128i rr = _mm_or_si128(_mm_or_si128(_mm_srli_si128(s0, 11), _mm_slli_si128(_mm_srli_si128(s1, 11), 5)),
                                            _mm_slli_si128(_mm_srli_si128(s2, 10), 10));
            _mm_storeu_si128((__m128i *)(b + x), bb);
            _mm_storeu_si128((__m128i *)(g + x), gg);
            _mm_storeu_si128((__m128i *)(r + x), rr);
        }
#endif
        for (; x < W; x++) { b[x] = row[3 * x]; g[x] = row[3 * x + 1]; r[x] = row[3 * x + 2]; }
    }
    auto t1 = std::chrono::steady_clock::now();
    for (int i = 0; i < 3; i++) RCK(cudaMemcpy(c->input.channel[i], c->planes[i].data(), (size_t)W * H, cudaMemcpyHostToDevice));
    auto t2 = std::chrono::steady_clock::now();
    RCK(cudaEventRecord(c->ev0));
    RNJ(nvjpegEncodeImage(c->h_enc, c->enc_state, c->enc_params, &c->input, NVJPEG_INPUT_BGR, W, H, NULL));
    RCK(cudaEventRecord(c->ev1));
    size_t length = 0;
    RNJ(nvjpegEncodeRetrieveBitstream(c->h_enc, c->enc_state, NULL, &length, NULL));
    if (length > cap) return -4;
    RNJ(nvjpegEncodeRetrieveBitstream(c->h_enc, c->enc_state, out, &length, NULL));
    auto t3 = std::chrono::steady_clock::now();
    RCK(cudaEventSynchronize(c->ev1));
    RCK(cudaEventElapsedTime(&c->last_gpu_ms, c->ev0, c->ev1));
    c->last_split_ms = std::chrono::duration<float, std::milli>(t1 - t0).count();
    c->last_h2d_ms = std::chrono::duration<float, std::milli>(t2 - t1).count();
    c->last_retrieve_ms = std::chrono::duration<float, std::milli>(t3 - t2).count();
    *len = length;
    return 0;
}

// the reference's cudaEvent bracket alone (ImageCompressorImpl.cu:279-281): planes already on the device.
// sync=1 additionally waits for the bitstream to be complete on the device (size query), so that the whole encode is timed.
__attribute__((visibility("default"))) int ref_encode_resident(void *p, float *gpu_ms, size_t *len) {
    RefCtx *c = (RefCtx *)p;
    RCK(cudaEventRecord(c->ev0));
    RNJ(nvjpegEncodeImage(c->h_enc, c->enc_state, c->enc_params, &c->input, NVJPEG_INPUT_BGR, c->W, c->H, NULL));
    size_t length = 0;
    RNJ(nvjpegEncodeRetrieveBitstream(c->h_enc, c->enc_state, NULL, &length, NULL));
    RCK(cudaEventRecord(c->ev1));
    RCK(cudaEventSynchronize(c->ev1));
    RCK(cudaEventElapsedTime(gpu_ms, c->ev0, c->ev1));
    if (len) *len = length;
    return 0;
}

// strict = 1: the event bracket exactly as the reference places it (around nvjpegEncodeImage only, :279-281);
// strict = 0: the size query (which waits for the bitstream to be complete on the device) inside the bracket
__attribute__((visibility("default"))) int ref_encode_resident2(void *p, int strict, float *gpu_ms, size_t *len) {
    RefCtx *c = (RefCtx *)p;
    RCK(cudaEventRecord(c->ev0));
    RNJ(nvjpegEncodeImage(c->h_enc, c->enc_state, c->enc_params, &c->input, NVJPEG_INPUT_BGR, c->W, c->H, NULL));
    if (strict) RCK(cudaEventRecord(c->ev1));
    size_t length = 0;
    RNJ(nvjpegEncodeRetrieveBitstream(c->h_enc, c->enc_state, NULL, &length, NULL));
    if (!strict) RCK(cudaEventRecord(c->ev1));
    RCK(cudaEventSynchronize(c->ev1));
    RCK(cudaEventElapsedTime(gpu_ms, c->ev0, c->ev1));
    if (len) *len = length;
    return 0;
}

__attribute__((visibility("default"))) int ref_last_times(void *p, float *gpu_ms, float *split_ms, float *h2d_ms, float *retrieve_ms) {
    RefCtx *c = (RefCtx *)p;
    *gpu_ms = c->last_gpu_ms; *split_ms = c->last_split_ms; *h2d_ms = c->last_h2d_ms; *retrieve_ms = c->last_retrieve_ms;
    return 0;
}

__attribute__((visibility("default"))) int ref_build_decode_env(void *p) {
    RefCtx *c = (RefCtx *)p;
    const nvjpegBackend_t be = NVJPEG_BACKEND_GPU_HYBRID;
    RNJ(nvjpegCreate(be, nullptr, &c->h_dec));
    RNJ(nvjpegJpegStateCreate(c->h_dec, &c->dec_state));
    RNJ(nvjpegDecoderCreate(c->h_dec, be, &c->decoder));
    RNJ(nvjpegDecoderStateCreate(c->h_dec, c->decoder, &c->decoupled));
    RNJ(nvjpegDecodeParamsCreate(c->h_dec, &c->dec_params));
    RNJ(nvjpegBufferPinnedCreate(c->h_dec, nullptr, &c->pinned[0]));
    RNJ(nvjpegBufferPinnedCreate(c->h_dec, nullptr, &c->pinned[1]));
    RNJ(nvjpegJpegStreamCreate(c->h_dec, &c->streams[0]));
    RNJ(nvjpegJpegStreamCreate(c->h_dec, &c->streams[1]));
    RNJ(nvjpegBufferDeviceCreate(c->h_dec, nullptr, &c->dev_buf));
    RNJ(nvjpegDecodeParamsSetOutputFormat(c->dec_params, NVJPEG_OUTPUT_BGR));
    if (!c->ev0) { RCK(cudaEventCreate(&c->ev0)); RCK(cudaEventCreate(&c->ev1)); }
    c->dec_init = true;
    return 0;
}

// DecodeWorker + getCVImageOnCPU: parse, 3-phase decoupled decode to planar BGR, 3x D2H, scalar interleave
__attribute__((visibility("default"))) int ref_decode(void *p, const unsigned char *jpg, size_t len, unsigned char *bgr, size_t step, int *W, int *H, float *gpu_ms) {
    RefCtx *c = (RefCtx *)p;
    int ncomp = 0, widths[NVJPEG_MAX_COMPONENT], heights[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t ss;
    RNJ(nvjpegGetImageInfo(c->h_dec, jpg, len, &ncomp, &ss, widths, heights));
    widths[1] = widths[2] = widths[0]; heights[1] = heights[2] = heights[0];
    for (int i = 0; i < 3; i++) {
        const size_t sz = (size_t)widths[i] * heights[i];
        c->out_img.pitch[i] = widths[i];
        if (sz > c->out_size[i]) {
            if (c->out_img.channel[i]) RCK(cudaFree(c->out_img.channel[i]));
            RCK(cudaMalloc((void **)&c->out_img.channel[i], sz));
            c->out_size[i] = sz;
        }
    }
    cudaStream_t stream;
    RCK(cudaStreamCreate(&stream));
    RCK(cudaEventRecord(c->ev0, stream));
    RNJ(nvjpegStateAttachDeviceBuffer(c->decoupled, c->dev_buf));
    RNJ(nvjpegJpegStreamParse(c->h_dec, jpg, len, 0, 0, c->streams[0]));
    RNJ(nvjpegStateAttachPinnedBuffer(c->decoupled, c->pinned[0]));
    RNJ(nvjpegDecodeJpegHost(c->h_dec, c->decoder, c->decoupled, c->dec_params, c->streams[0]));
    RNJ(nvjpegDecodeJpegTransferToDevice(c->h_dec, c->decoder, c->decoupled, c->streams[0], stream));
    RNJ(nvjpegDecodeJpegDevice(c->h_dec, c->decoder, c->decoupled, &c->out_img, stream));
    RCK(cudaEventRecord(c->ev1, stream));
    RCK(cudaEventSynchronize(c->ev1));
    if (gpu_ms) RCK(cudaEventElapsedTime(gpu_ms, c->ev0, c->ev1));
    const int w = widths[0], h = heights[0];
    *W = w; *H = h;
    if (bgr) {
        for (int i = 0; i < 3; i++) {
            c->planes[i].resize((size_t)w * h);
            RCK(cudaMemcpy2D(c->planes[i].data(), (size_t)w, c->out_img.channel[i], (size_t)c->out_img.pitch[i], w, h, cudaMemcpyDeviceToHost));
        }
        for (int y = 0; y < h; y++) {
            unsigned char *row = bgr + (size_t)y * step;
            const unsigned char *b = c->planes[0].data() + (size_t)y * w, *g = c->planes[1].data() + (size_t)y * w, *r = c->planes[2].data() + (size_t)y * w;
            for (int x = 0; x < w; x++) { row[3 * x] = b[x]; row[3 * x + 1] = g[x]; row[3 * x + 2] = r[x]; }
        }
    }
    cudaStreamDestroy(stream);
    return 0;
}

__attribute__((visibility("default"))) void ref_destroy(void *p) {
    RefCtx *c = (RefCtx *)p;
    if (!c) return;
    if (c->enc_init) {
        nvjpegEncoderParamsDestroy(c->enc_params); nvjpegEncoderStateDestroy(c->enc_state); nvjpegDestroy(c->h_enc);
        for (int i = 0; i < 3; i++) cudaFree(c->input.channel[i]);
    }
    if (c->dec_init) {
        nvjpegJpegStateDestroy(c->decoupled); nvjpegJpegStateDestroy(c->dec_state); nvjpegDecoderDestroy(c->decoder);
        nvjpegDecodeParamsDestroy(c->dec_params);
        for (int i = 0; i < 2; i++) { nvjpegJpegStreamDestroy(c->streams[i]); nvjpegBufferPinnedDestroy(c->pinned[i]); }
        nvjpegBufferDeviceDestroy(c->dev_buf); nvjpegDestroy(c->h_dec);
        for (int i = 0; i < 3; i++) if (c->out_img.channel[i]) cudaFree(c->out_img.channel[i]);
    }
    if (c->ev0) { cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); }
    delete c;
}

}  // extern "C"
