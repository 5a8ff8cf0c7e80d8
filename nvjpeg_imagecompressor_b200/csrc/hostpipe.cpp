// hostpipe.cpp -- see hostpipe.h
#include "hostpipe.h"

#include <string.h>

#include <algorithm>

#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define B2J_HAVE_NT 1
#endif

namespace b2j {

// Streaming copy: the destination is written with non-temporal stores (whole 64-byte lines), so the copy costs one read
// and one write of memory instead of read + read-for-ownership + write, and the staging ring / the caller's image do
// not evict each other from the caches. Rows of an image are far below the size at which memcpy() switches by itself.
static inline void copy_stream(uint8_t *dst, const uint8_t *src, size_t n) {
#ifdef B2J_HAVE_NT
    if (n >= 256) {
        const size_t head = (64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63;
        if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
        const size_t lines = n / 64;
        for (size_t i = 0; i < lines; i++) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src) + 0);
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src) + 1);
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src) + 2);
            const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src) + 3);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst) + 0, a);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst) + 1, b);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst) + 2, c);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst) + 3, d);
            src += 64; dst += 64;
        }
        n -= lines * 64;
    }
#endif
    if (n) memcpy(dst, src, n);
}

CopyPool::CopyPool(int nthreads) {
    const int n = std::max(0, nthreads - 1);
    for (int i = 0; i < n; i++) workers_.emplace_back([this] { worker(); });
}

CopyPool::~CopyPool() {
    {
        std::lock_guard<std::mutex> lk(m_);
        stop_ = true;
        gen_++;
    }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
}

void CopyPool::run_slices(const Job &j, std::atomic<size_t> &next) {
    for (;;) {
        const size_t r0 = next.fetch_add(j.rows_per);
        if (r0 >= j.rows) break;
        const size_t r1 = std::min(j.rows, r0 + j.rows_per);
        if (j.dstep == j.row_bytes && j.sstep == j.row_bytes) {
            copy_stream(j.dst + r0 * j.dstep, j.src + r0 * j.sstep, (r1 - r0) * j.row_bytes);
        } else {
            for (size_t r = r0; r < r1; r++) copy_stream(j.dst + r * j.dstep, j.src + r * j.sstep, j.row_bytes);
        }
    }
#ifdef B2J_HAVE_NT
    _mm_sfence();   // the streamed lines are globally visible before the DMA engine (or the caller) reads them
#endif
}

void CopyPool::worker() {
    uint64_t seen = 0;
    for (;;) {
        Job j;
        {
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return gen_ != seen; });
            seen = gen_;
            if (stop_) return;
            j = job_;
        }
        run_slices(j, next_);
        {
            std::lock_guard<std::mutex> lk(m_);
            if (--active_ == 0) done_.notify_one();
        }
    }
}

void CopyPool::copy2d(uint8_t *dst, size_t dstep, const uint8_t *src, size_t sstep, size_t row_bytes, size_t rows) {
    if (rows == 0 || row_bytes == 0) return;
    Job j{dst, src, dstep, sstep, row_bytes, rows, 1};
    // slices of about 1 MB keep the threads balanced without much traffic on the counter
    j.rows_per = std::max<size_t>(1, (1u << 20) / row_bytes);
    if (workers_.empty() || rows * row_bytes < (4u << 20)) {
        std::atomic<size_t> next{0};
        run_slices(j, next);
        return;
    }
    {
        std::lock_guard<std::mutex> lk(m_);
        job_ = j;
        next_.store(0);
        active_ = (int)workers_.size();
        gen_++;
    }
    cv_.notify_all();
    run_slices(j, next_);
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [&] { return active_ == 0; });
}

bool is_pageable_host(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

cudaError_t StageRing::ensure(size_t need) {
    if (need <= bytes) return cudaSuccess;
    release();
    for (int i = 0; i < N; i++) {
        cudaError_t e = cudaHostAlloc(&buf[i], need, cudaHostAllocDefault);
        if (e != cudaSuccess) { release(); return e; }
        e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
        if (e != cudaSuccess) { release(); return e; }
        busy[i] = false;
    }
    bytes = need;
    return cudaSuccess;
}

void StageRing::release() {
    for (int i = 0; i < N; i++) {
        if (buf[i]) cudaFreeHost(buf[i]);
        if (ev[i]) cudaEventDestroy(ev[i]);
        buf[i] = nullptr; ev[i] = nullptr; busy[i] = false;
    }
    bytes = 0;
}

}  // namespace b2j
