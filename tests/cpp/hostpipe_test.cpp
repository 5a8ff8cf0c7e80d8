// CopyPool::copy2d (csrc/hostpipe.cpp: copy threads + non-temporal line stores) against memcpy for every destination /
// source alignment, ragged row sizes and pitches, small and multi-threaded jobs. Host only: no CUDA call is made.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../nvjpeg_imagecompressor_b200/csrc/hostpipe.h"

static uint32_t rng_state = 12345;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

static int check(b2j::CopyPool &pool, size_t row_bytes, size_t rows, size_t dpad, size_t spad, size_t doff, size_t soff) {
    const size_t dstep = row_bytes + dpad, sstep = row_bytes + spad;
    std::vector<uint8_t> src(soff + sstep * rows + 64), dst(doff + dstep * rows + 64, 0xA5), ref;
    for (auto &b : src) b = (uint8_t)rnd();
    ref = dst;
    for (size_t r = 0; r < rows; r++) memcpy(ref.data() + doff + r * dstep, src.data() + soff + r * sstep, row_bytes);
    pool.copy2d(dst.data() + doff, dstep, src.data() + soff, sstep, row_bytes, rows);
    if (dst != ref) {
        printf("MISMATCH row_bytes=%zu rows=%zu dpad=%zu spad=%zu doff=%zu soff=%zu\n", row_bytes, rows, dpad, spad, doff, soff);
        return 1;
    }
    return 0;
}

int main() {
    int bad = 0;
    for (int threads : {1, 4}) {
        b2j::CopyPool pool(threads);
        for (size_t doff = 0; doff < 70; doff += 7)
            for (size_t soff = 0; soff < 70; soff += 13)
                for (size_t rb : {1u, 63u, 255u, 256u, 257u, 1000u, 4099u, 24960u})
                    bad += check(pool, rb, 5, (doff * 3) % 17, (soff * 5) % 11, doff, soff);
        bad += check(pool, 24960, 700, 0, 0, 3, 5);        // > 4 MB: goes to the worker threads, contiguous rows
        bad += check(pool, 24957, 700, 19, 3, 1, 2);       // ragged pitch, worker threads
        bad += check(pool, 1 << 20, 9, 0, 0, 0, 0);        // the linear helpers' 1 MB rows
    }
    printf(bad ? "FAILED %d\n" : "ok\n", bad);
    return bad ? 1 : 0;
}
