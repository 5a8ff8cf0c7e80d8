// hostpipe.h -- pageable host memory <-> device, at PCIe speed.
// The reference hands the codec pageable buffers (cv::Mat pixels, std::vector bytes: ImageCompressorImpl.cu:273-287,
// :204-221) and pays cv::split + pageable cudaMemcpy for it. Here a small pool of copy threads moves row groups between
// the caller's pages and a ring of pinned staging buffers while the DMA engine moves the previous group, so a pageable
// image goes up (or comes down) at the speed of the slower of {host memcpy bandwidth, PCIe}.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace b2j {

// rows of `row_bytes` bytes from (src, sstep) to (dst, dstep), split over the pool's threads (the caller takes part)
class CopyPool {
public:
    explicit CopyPool(int nthreads);
    ~CopyPool();
    void copy2d(uint8_t *dst, size_t dstep, const uint8_t *src, size_t sstep, size_t row_bytes, size_t rows);
    int threads() const { return (int)workers_.size() + 1; }

private:
    struct Job { uint8_t *dst; const uint8_t *src; size_t dstep, sstep, row_bytes, rows, rows_per; };
    void worker();
    static void run_slices(const Job &j, std::atomic<size_t> &next);
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    Job job_{};
    std::atomic<size_t> next_{0};
    uint64_t gen_ = 0;
    int active_ = 0;
    bool stop_ = false;
};

bool is_pageable_host(const void *p);

// ring of pinned staging buffers shared by the up- and download helpers
struct StageRing {
    static constexpr int N = 3;
    uint8_t *buf[N] = {nullptr, nullptr, nullptr};
    size_t bytes = 0;
    cudaEvent_t ev[N] = {nullptr, nullptr, nullptr};
    bool busy[N] = {false, false, false};
    cudaError_t ensure(size_t need);
    void release();
};

}  // namespace b2j
