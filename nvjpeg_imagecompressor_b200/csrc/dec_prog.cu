// dec_prog.cu -- progressive (SOF2) decode: what the reference as shipped writes (NVJPEG_ENCODING_PROGRESSIVE_DCT_HUFFMAN,
// ImageCompressorImpl.cu:28) and nvJPEG's decoder accepts (:361-366). Every scan is absorbed into the coefficient
// array in file order (jdphuff.c decode_mcu_DC_first / DC_refine / AC_first / AC_refine; tests compare the pixels with
// the CPU checker and cv2.imdecode), then the baseline back end runs (k_idct on the
// coefficients' own DCs, k_upcolor). Parallelism: restart intervals of a scan are independent (one thread each); a
// scan without restart markers is ONE sequential chain -- refinement scans read the coefficients' history, so the
// self-synchronising scheme of the baseline decoder does not carry over. Correct for every stream the parser accepts,
// fast only for streams with restart markers; see DESIGN.md.
#include <stdio.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "dec.h"
#include "dec_kernels.h"

namespace b2j {

constexpr int PROG_LUT_BITS = 9;
struct ProgTable {                       // one Huffman table (jdhuff.c jpeg_make_d_derived_tbl)
    uint16_t lut[1 << PROG_LUT_BITS];    // (len << 8) | symbol for codes of length <= PROG_LUT_BITS, else 0
    int32_t maxcode[18];
    int32_t valoff[17];
    uint8_t vals[256];
};
struct ProgScanDev {
    ProgTable tb[2][3];                  // [0 DC | 1 AC][component of the scan]
    int ncomp, comp[3], Ss, Se, Ah, Al, ri;
    uint32_t nint;                       // restart intervals
    uint64_t units;                      // MCUs (interleaved) or blocks of the component
};

static int build_table(const uint8_t bits[17], const uint8_t vals[256], ProgTable *t) {
    memset(t, 0, sizeof(*t));
    int code = 0, p = 0;
    for (int l = 1; l <= 16; l++) {
        const int n = bits[l];
        if (code + n > (1 << l) || p + n > 256) return B2J_EFORMAT;
        t->valoff[l] = p - code;
        for (int i = 0; i < n; i++, p++, code++)
            if (l <= PROG_LUT_BITS) {
                const int lo = code << (PROG_LUT_BITS - l), cnt = 1 << (PROG_LUT_BITS - l);
                for (int j = 0; j < cnt; j++) t->lut[lo + j] = (uint16_t)((l << 8) | vals[p]);
            }
        t->maxcode[l] = n ? code - 1 : -1;
        code <<= 1;
    }
    t->maxcode[17] = 0x7fffffff;
    memcpy(t->vals, vals, 256);
    return B2J_OK;
}

// bit reader over STUFFED bytes (jdhuff.c jpeg_fill_bit_buffer): FF 00 -> FF; any other marker ends the data (zero bits
// follow, as libjpeg feeds them)
struct BitReader {
    const uint8_t *d;
    size_t pos, end;
    uint64_t acc;     // next bits left-aligned
    int nav;
    bool marker;
    __device__ __forceinline__ void init(const uint8_t *p, size_t a, size_t b) { d = p; pos = a; end = b; acc = 0; nav = 0; marker = false; }
    __device__ __forceinline__ void fill() {
        while (nav <= 56) {
            uint32_t b = 0;
            if (!marker && pos < end) {
                b = d[pos];
                if (b == 0xFF) {
                    const uint32_t b2 = pos + 1 < end ? d[pos + 1] : 0xD9;
                    if (b2 == 0) pos += 2; else { marker = true; b = 0; }
                } else pos++;
            }
            acc |= (uint64_t)b << (56 - nav);
            nav += 8;
        }
    }
    __device__ __forceinline__ uint32_t peek(int n) { return (uint32_t)(acc >> (64 - n)); }   // 1 <= n <= 32, after fill()
    __device__ __forceinline__ void skip(int n) { acc <<= n; nav -= n; }
    __device__ __forceinline__ uint32_t get(int n) { if (n == 0) return 0; fill(); const uint32_t v = peek(n); skip(n); return v; }
    __device__ __forceinline__ int sym(const ProgTable &t) {
        fill();
        const uint32_t top = peek(16);
        const uint32_t e = t.lut[top >> (16 - PROG_LUT_BITS)];
        if (e) { skip((int)(e >> 8)); return (int)(e & 0xFF); }
        for (int l = PROG_LUT_BITS + 1; l <= 16; l++) {
            const int code = (int)(top >> (16 - l));
            if (code <= t.maxcode[l]) { skip(l); return t.vals[(t.valoff[l] + code) & 255]; }
        }
        skip(16);
        return 0;   // invalid code: jdhuff.c warns and returns 0
    }
};

__device__ __forceinline__ int prog_extend(int v, int n) { return n == 0 ? 0 : (v < (1 << (n - 1)) ? v - (1 << n) + 1 : v); }

// scan-order index of block (bx, by) of component c (jdcoefct.c: non-interleaved scans walk the component's own blocks)
__device__ __forceinline__ size_t prog_block_index(const Geom &g, int c, int bx, int by) {
    const int h = c ? 1 : g.hs, v = c ? 1 : g.vs;
    const int mx = bx / h, my = by / v;
    const int blkn = c ? g.bpm - 3 + c : (by % v) * h + (bx % h);
    return ((size_t)my * g.mcux + mx) * g.bpm + blkn;
}

// one block of one scan (jdphuff.c); blk: 64 coefficients in zig-zag order
__device__ void prog_block(BitReader &r, const ProgScanDev &sc, int ci, int16_t *__restrict__ blk, int *pred, int &eobrun) {
    const int Al = sc.Al, p1 = 1 << Al, m1 = -(1 << Al);
    if (sc.Ss == 0) {
        if (sc.Ah == 0) {          // decode_mcu_DC_first
            const int s = r.sym(sc.tb[0][ci]);
            const int diff = s ? prog_extend((int)r.get(s), s) : 0;
            pred[ci] += diff;
            blk[0] = (int16_t)(pred[ci] * (1 << Al));
        } else if (r.get(1)) {     // decode_mcu_DC_refine
            blk[0] |= (int16_t)p1;
        }
        return;
    }
    const ProgTable &ac = sc.tb[1][ci];
    if (sc.Ah == 0) {              // decode_mcu_AC_first
        if (eobrun > 0) { eobrun--; return; }
        for (int k = sc.Ss; k <= sc.Se; k++) {
            const int rs = r.sym(ac), rr = rs >> 4, ss = rs & 15;
            if (ss) {
                k += rr;
                const int v = prog_extend((int)r.get(ss), ss);
                if (k <= 63) blk[k] = (int16_t)(v * (1 << Al));
            } else {
                if (rr == 15) k += 15;
                else { eobrun = 1 << rr; if (rr) eobrun += (int)r.get(rr); eobrun--; break; }
            }
        }
        return;
    }
    // decode_mcu_AC_refine
    int k = sc.Ss;
    if (eobrun == 0) {
        for (; k <= sc.Se; k++) {
            const int rs = r.sym(ac);
            int rr = rs >> 4;
            const int ss = rs & 15;
            int val = 0;
            if (ss) val = r.get(1) ? p1 : m1;      // the size of a new coefficient is always 1
            else if (rr != 15) { eobrun = 1 << rr; if (rr) eobrun += (int)r.get(rr); break; }
            // advance over already-nonzero coefficients and rr still-zero ones, refining the nonzero ones
            do {
                int16_t *cp = blk + k;
                if (*cp != 0) {
                    if (r.get(1) && (*cp & p1) == 0) *cp = (int16_t)(*cp + (*cp >= 0 ? p1 : m1));
                } else if (--rr < 0) break;
                k++;
            } while (k <= sc.Se);
            if (val && k <= 63) blk[k] = (int16_t)val;
        }
    }
    if (eobrun > 0) {
        for (; k <= sc.Se; k++) {
            int16_t *cp = blk + k;
            if (*cp != 0 && r.get(1) && (*cp & p1) == 0) *cp = (int16_t)(*cp + (*cp >= 0 ? p1 : m1));
        }
        eobrun--;
    }
}

// one thread per restart interval; iv[j] = file offset of interval j's first byte, iv[nint] = end of the segment
__global__ void __launch_bounds__(32)
k_prog_scan(const uint8_t *__restrict__ file, const uint64_t *__restrict__ iv, const ProgScanDev *__restrict__ scd, Geom g,
            int16_t *__restrict__ coef) {
    __shared__ ProgScanDev sc;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(scd);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&sc);
        for (int i = threadIdx.x; i < (int)(sizeof(ProgScanDev) / 4); i += 32) dst[i] = src[i];
    }
    __syncwarp();
    const uint32_t j = blockIdx.x * 32 + threadIdx.x;
    if (j >= sc.nint) return;
    BitReader r;
    r.init(file, (size_t)iv[j], (size_t)iv[j + 1]);
    int pred[3] = {0, 0, 0}, eobrun = 0;
    const uint64_t u0 = sc.ri ? (uint64_t)j * sc.ri : 0, u1 = sc.ri ? min(sc.units, u0 + sc.ri) : sc.units;
    if (sc.ncomp > 1) {            // interleaved: MCU order, every block of the MCU (padding blocks too)
        for (uint64_t mi = u0; mi < u1; mi++)
            for (int i = 0; i < sc.ncomp; i++) {
                const int c = sc.comp[i], nb = c ? 1 : g.hs * g.vs, b0 = c ? g.bpm - 3 + c : 0;
                for (int b = 0; b < nb; b++) prog_block(r, sc, i, coef + ((size_t)mi * g.bpm + b0 + b) * 64, pred, eobrun);
            }
    } else {                       // one component: raster order over its own blocks (real data only)
        const int c = sc.comp[0], wib = g.wib[c];
        for (uint64_t u = u0; u < u1; u++)
            prog_block(r, sc, 0, coef + prog_block_index(g, c, (int)(u % wib), (int)(u / wib)) * 64, pred, eobrun);
    }
}

#define PCK(call)                                                                                                  \
    do {                                                                                                           \
        cudaError_t _e = (call);                                                                                   \
        if (_e != cudaSuccess) {                                                                                   \
            snprintf(err, errlen, "%s at %s:%d", cudaGetErrorString(_e), __FILE__, __LINE__);                       \
            rc = B2J_ECUDA;                                                                                        \
            goto done;                                                                                             \
        }                                                                                                          \
    } while (0)

int dec_progressive(const uint8_t *jpg, size_t len, const ProgInfo &info, const Geom &g, int16_t *d_coef, void *d_tb,
                    uint8_t *d_planes, uint8_t *d_bgr, size_t step, cudaStream_t s, uint64_t *launches, char *err, size_t errlen) {
    int rc = B2J_OK;
    uint8_t *d_file = nullptr;
    ProgScanDev *d_sc = nullptr;
    uint64_t *d_iv = nullptr;
    std::vector<ProgScanDev> hsc(info.nscans);
    std::vector<uint64_t> hiv;
    std::vector<size_t> iv_base(info.nscans);
    DecTables *htb = nullptr;
    // per scan: tables, units, restart intervals (the markers are found on the host: FF Dn is unambiguous in stuffed data)
    for (int i = 0; i < info.nscans; i++) {
        const ProgScan &ps = info.scans[i];
        ProgScanDev &sd = hsc[i];
        memset(&sd, 0, sizeof(sd));
        sd.ncomp = ps.ncomp; sd.Ss = ps.Ss; sd.Se = ps.Se; sd.Ah = ps.Ah; sd.Al = ps.Al; sd.ri = ps.restart_interval;
        for (int c = 0; c < ps.ncomp; c++) {
            sd.comp[c] = ps.comp[c];
            if ((ps.Ss == 0 && ps.Ah == 0 && build_table(ps.bits[0][c], ps.vals[0][c], &sd.tb[0][c])) ||
                (ps.Ss > 0 && build_table(ps.bits[1][c], ps.vals[1][c], &sd.tb[1][c]))) {
                snprintf(err, errlen, "invalid Huffman table in scan %d", i);
                return B2J_EFORMAT;
            }
        }
        sd.units = ps.ncomp > 1 ? (uint64_t)g.mcux * g.mcuy : (uint64_t)g.wib[ps.comp[0]] * g.hib[ps.comp[0]];
        const uint64_t want = sd.ri ? (sd.units + sd.ri - 1) / sd.ri : 1;
        iv_base[i] = hiv.size();
        hiv.push_back(ps.seg_off);
        if (sd.ri) {
            const uint8_t *p = jpg + ps.seg_off, *e = p + ps.seg_len;
            while (hiv.size() - iv_base[i] < want) {
                const uint8_t *f = (const uint8_t *)memchr(p, 0xFF, (size_t)(e - p));
                if (!f || f + 1 >= e) break;
                if ((f[1] & 0xF8) == 0xD0) { hiv.push_back((uint64_t)(f + 2 - jpg)); p = f + 2; }
                else p = f + 1;
            }
        }
        sd.nint = (uint32_t)(hiv.size() - iv_base[i]);   // fewer than `want`: the missing intervals keep their zeros (truncated file)
        hiv.push_back(ps.seg_off + ps.seg_len);
    }
    PCK(cudaMalloc(&d_file, len + 16));
    PCK(cudaMalloc(&d_sc, sizeof(ProgScanDev) * info.nscans));
    PCK(cudaMalloc(&d_iv, sizeof(uint64_t) * hiv.size()));
    PCK(cudaMemcpyAsync(d_file, jpg, len, cudaMemcpyDefault, s));
    PCK(cudaMemcpyAsync(d_sc, hsc.data(), sizeof(ProgScanDev) * info.nscans, cudaMemcpyHostToDevice, s));
    PCK(cudaMemcpyAsync(d_iv, hiv.data(), sizeof(uint64_t) * hiv.size(), cudaMemcpyHostToDevice, s));
    PCK(cudaMemsetAsync(d_coef, 0, (size_t)g.nblocks * 128, s));
    for (int i = 0; i < info.nscans; i++) {
        k_prog_scan<<<(hsc[i].nint + 31) / 32, 32, 0, s>>>(d_file, d_iv + iv_base[i], d_sc + i, g, d_coef);
        PCK(cudaGetLastError());
        if (launches) (*launches)++;
    }
    {   // de-quantisation tables for the back end
        htb = (DecTables *)calloc(1, dec_tables_size());
        if (!htb) { rc = B2J_ENOMEM; goto done; }
        memcpy(htb->q, info.qt, sizeof(htb->q));
        PCK(cudaMemcpyAsync(d_tb, htb, dec_tables_size(), cudaMemcpyHostToDevice, s));
        uint8_t *py = d_planes;
        uint8_t *pcb = py + (((size_t)g.mcux * 8 * g.hs * g.mcuy * 8 * g.vs + 63) & ~(size_t)63);
        uint8_t *pcr = pcb + (((size_t)g.mcux * 8 * g.mcuy * 8 + 63) & ~(size_t)63);
        PCK(launch_idct(d_coef, nullptr, g, d_tb, py, pcb, pcr, 0, s));
        PCK(launch_upcolor(py, pcb, pcr, g, d_bgr, step, s));
        if (launches) (*launches) += 2;
    }
    PCK(cudaStreamSynchronize(s));   // the host vectors and the file copy go away below
done:
    if (rc != B2J_OK) cudaStreamSynchronize(s);
    cudaFree(d_file); cudaFree(d_sc); cudaFree(d_iv);
    free(htb);
    return rc;
}

}  // namespace b2j
