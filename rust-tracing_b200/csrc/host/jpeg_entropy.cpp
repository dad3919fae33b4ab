// Marker parsing and Huffman decoding of baseline (SOF0 / SOF1, 8-bit) JPEG streams: ITU-T T.81 Annex B (markers),
// F.2.2 (decoding procedures), Annex C (table generation). The output is what libjpeg's jdhuff.c hands to its inverse
// DCT: quantised coefficients in natural order. ImageTexture::new (texture.rs:76-80) decodes assets/earth-large.jpg,
// a baseline 4:2:0 file; progressive, arithmetic-coded, lossless and 12-bit streams are refused (RT_ERR_UNSUPPORTED).
#include "jpeg_entropy.h"

#include <cstring>

#include "../../../include/rt_b200.h"
#include "host_common.h"

namespace rt_host {
namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
    bool present = false;
    uint8_t bits[17] = {0};
    uint8_t vals[256] = {0};
    // T.81 F.2.2.3: codes of length l are the integers mincode[l] .. maxcode[l]; valptr[l] indexes vals
    int32_t mincode[17], maxcode[18], valptr[17];
    // 9-bit lookahead: (length << 8) | value, 0 = longer than 9 bits
    uint16_t look[512];

    int build(std::string* err) {
        int code = 0, k = 0;
        std::memset(look, 0, sizeof(look));
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k;
            mincode[l] = code;
            for (int i = 0; i < bits[l]; ++i, ++k, ++code) {
                if (l <= 9) {
                    const int first = code << (9 - l);
                    for (int j = 0; j < (1 << (9 - l)); ++j) look[first + j] = (uint16_t)((l << 8) | vals[k]);
                }
            }
            maxcode[l] = bits[l] ? code - 1 : -1;
            if (code > (1 << l)) { *err = "JPEG: over-subscribed Huffman table"; return RT_ERR_INVALID_ARGUMENT; }
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        return RT_OK;
    }
};

struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t acc = 0;     // bits are taken from the top
    int n = 0;            // valid bits in acc
    bool hit_marker = false;

    void fill() {
        // fast path: the next eight bytes hold no 0xff (no stuffing, no marker) - take as many whole bytes as fit
        if (!hit_marker && p + 8 <= end) {
            uint64_t v;
            std::memcpy(&v, p, 8);
            if (!(((v ^ 0xffffffffffffffffull) - 0x0101010101010101ull) & ~(v ^ 0xffffffffffffffffull) & 0x8080808080808080ull)) {
                v = __builtin_bswap64(v);
                const int nb = (64 - n) >> 3;
                if (nb > 0) {
                    const uint64_t chunk = nb == 8 ? v : (v >> (64 - 8 * nb));
                    acc |= nb == 8 ? chunk : (chunk << (64 - n - 8 * nb));
                    n += 8 * nb;
                    p += nb;
                }
                return;
            }
        }
        while (n <= 56) {
            int byte = 0;
            if (!hit_marker && p < end) {
                byte = *p;
                if (byte == 0xff) {
                    if (p + 1 < end && p[1] == 0x00) { p += 2; }             // stuffed zero
                    else { hit_marker = true; byte = 0; }                    // a marker: feed zeros, leave p on it
                } else {
                    ++p;
                }
            } else {
                hit_marker = true;
            }
            acc |= (uint64_t)byte << (56 - n);
            n += 8;
        }
    }
    inline int peek(int k) { if (n < k) fill(); return (int)(acc >> (64 - k)); }
    inline void skip(int k) { acc <<= k; n -= k; }
    inline int get(int k) { if (k == 0) return 0; const int v = peek(k); skip(k); return v; }
    void reset() { acc = 0; n = 0; hit_marker = false; }
};

inline int decode_symbol(BitReader& br, const HuffTable& t, bool* bad) {
    const int look = br.peek(9);
    const uint16_t e = t.look[look];
    if (e) { br.skip(e >> 8); return e & 0xff; }
    int code = br.peek(16);
    for (int l = 10; l <= 16; ++l) {
        const int c = code >> (16 - l);
        if (t.maxcode[l] >= 0 && c <= t.maxcode[l] && c >= t.mincode[l]) { br.skip(l); return t.vals[t.valptr[l] + c - t.mincode[l]]; }
    }
    *bad = true;
    return 0;
}

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }   // T.81 F.2.2.1 EXTEND

uint16_t be16(const uint8_t* p) { return (uint16_t)((p[0] << 8) | p[1]); }

}  // namespace

int jpeg_entropy_decode(const uint8_t* data, size_t n, JpegFrame* f, int16_t* coef, size_t capacity, std::string* err) {
    if (n < 4 || data[0] != 0xff || data[1] != 0xd8) { *err = "JPEG: no SOI marker"; return RT_ERR_INVALID_ARGUMENT; }
    *f = JpegFrame();
    std::memset(f->quant, 0, sizeof(f->quant));
    HuffTable dc[4], ac[4];
    bool have_quant[4] = {false, false, false, false};
    bool have_frame = false, saw_scan = false;
    int restart_interval = 0;
    bool comp_done[3] = {false, false, false};   // a component is coded by exactly one scan in a baseline stream
    size_t pos = 2;
    while (pos + 4 <= n) {
        if (data[pos] != 0xff) { *err = "JPEG: marker expected"; return RT_ERR_INVALID_ARGUMENT; }
        while (pos < n && data[pos] == 0xff) ++pos;          // fill bytes
        if (pos >= n) break;
        const int m = data[pos++];
        if (m == 0xd9) break;                                // EOI
        if (m == 0x01 || (m >= 0xd0 && m <= 0xd7)) continue; // TEM / stray RSTn: no length
        if (pos + 2 > n) { *err = "JPEG: truncated segment"; return RT_ERR_INVALID_ARGUMENT; }
        const size_t len = be16(data + pos);
        if (len < 2 || pos + len > n) { *err = "JPEG: truncated segment"; return RT_ERR_INVALID_ARGUMENT; }
        const uint8_t* s = data + pos + 2;
        const size_t sl = len - 2;
        if (m == 0xdb) {                                     // DQT
            size_t k = 0;
            while (k < sl) {
                const int pq = s[k] >> 4, tq = s[k] & 15;
                ++k;
                if (tq > 3 || k + (pq ? 128 : 64) > sl) { *err = "JPEG: bad DQT"; return RT_ERR_INVALID_ARGUMENT; }
                for (int i = 0; i < 64; ++i) {
                    f->quant[tq][kZigzag[i]] = pq ? be16(s + k + 2 * i) : s[k + i];
                }
                k += pq ? 128 : 64;
                have_quant[tq] = true;
            }
        } else if (m == 0xc4) {                              // DHT
            size_t k = 0;
            while (k < sl) {
                if (k + 17 > sl) { *err = "JPEG: bad DHT"; return RT_ERR_INVALID_ARGUMENT; }
                const int tc = s[k] >> 4, th = s[k] & 15;
                if (tc > 1 || th > 3) { *err = "JPEG: bad DHT"; return RT_ERR_INVALID_ARGUMENT; }
                HuffTable& t = tc ? ac[th] : dc[th];
                t = HuffTable();
                int total = 0;
                for (int l = 1; l <= 16; ++l) { t.bits[l] = s[k + l]; total += t.bits[l]; }
                k += 17;
                if (total > 256 || k + total > sl) { *err = "JPEG: bad DHT"; return RT_ERR_INVALID_ARGUMENT; }
                std::memcpy(t.vals, s + k, total);
                k += total;
                const int rc = t.build(err);
                if (rc < 0) return rc;
                t.present = true;
            }
        } else if (m == 0xc0 || m == 0xc1) {                 // SOF0 / SOF1 with 8-bit samples
            if (have_frame || sl < 6) { *err = "JPEG: bad SOF"; return RT_ERR_INVALID_ARGUMENT; }
            if (s[0] != 8) { *err = "JPEG: only 8-bit samples are supported"; return RT_ERR_UNSUPPORTED; }
            f->height = be16(s + 1);
            f->width = be16(s + 3);
            f->ncomp = s[5];
            if (f->width <= 0 || f->height <= 0) { *err = "JPEG: empty frame (DNL is not supported)"; return RT_ERR_UNSUPPORTED; }
            if (f->ncomp != 1 && f->ncomp != 3) { *err = "JPEG: 1 or 3 components are supported"; return RT_ERR_UNSUPPORTED; }
            if (sl < (size_t)(6 + 3 * f->ncomp)) { *err = "JPEG: bad SOF"; return RT_ERR_INVALID_ARGUMENT; }
            for (int c = 0; c < f->ncomp; ++c) {
                JpegComponent& C = f->comp[c];
                C.id = s[6 + 3 * c];
                C.h = s[7 + 3 * c] >> 4;
                C.v = s[7 + 3 * c] & 15;
                C.tq = s[8 + 3 * c];
                if (C.h < 1 || C.h > 4 || C.v < 1 || C.v > 4 || C.tq > 3) { *err = "JPEG: bad SOF component"; return RT_ERR_INVALID_ARGUMENT; }
                if (C.h > f->hmax) f->hmax = C.h;
                if (C.v > f->vmax) f->vmax = C.v;
            }
            if (f->ncomp == 1) { f->comp[0].h = f->comp[0].v = 1; f->hmax = f->vmax = 1; }   // a single component is never subsampled
            const int mcus_x = (f->width + 8 * f->hmax - 1) / (8 * f->hmax), mcus_y = (f->height + 8 * f->vmax - 1) / (8 * f->vmax);
            size_t off = 0;
            for (int c = 0; c < f->ncomp; ++c) {
                JpegComponent& C = f->comp[c];
                C.blocks_w = mcus_x * C.h;
                C.blocks_h = mcus_y * C.v;
                C.ds_w = (f->width * C.h + f->hmax - 1) / f->hmax;
                C.ds_h = (f->height * C.v + f->vmax - 1) / f->vmax;
                C.coef_offset = off;
                off += (size_t)C.blocks_w * C.blocks_h * 64;
            }
            f->coef_count = off;
            have_frame = true;
        } else if (m == 0xc2 || m == 0xc3 || (m >= 0xc5 && m <= 0xcf && m != 0xcc)) {
            *err = "JPEG: only baseline Huffman streams are supported (this one is progressive, lossless or arithmetic-coded)";
            return RT_ERR_UNSUPPORTED;
        } else if (m == 0xdd) {                              // DRI
            if (sl < 2) { *err = "JPEG: bad DRI"; return RT_ERR_INVALID_ARGUMENT; }
            restart_interval = be16(s);
        } else if (m == 0xee) {                              // APP14 "Adobe": transform 0 with three components = RGB
            if (sl >= 12 && std::memcmp(s, "Adobe", 5) == 0 && s[11] == 0) f->adobe_rgb = true;
        } else if (m == 0xda) {                              // SOS
            if (!have_frame) { *err = "JPEG: scan before frame header"; return RT_ERR_INVALID_ARGUMENT; }
            if (sl < 1) { *err = "JPEG: bad SOS"; return RT_ERR_INVALID_ARGUMENT; }
            const int ns = s[0];
            if (ns < 1 || ns > f->ncomp || sl < (size_t)(4 + 2 * ns)) { *err = "JPEG: bad SOS"; return RT_ERR_INVALID_ARGUMENT; }
            int ci[3], td[3], ta[3];
            for (int k = 0; k < ns; ++k) {
                ci[k] = -1;
                for (int c = 0; c < f->ncomp; ++c) if (f->comp[c].id == s[1 + 2 * k]) ci[k] = c;
                td[k] = s[2 + 2 * k] >> 4;
                ta[k] = s[2 + 2 * k] & 15;
                if (ci[k] < 0 || td[k] > 3 || ta[k] > 3 || comp_done[ci[k]]) { *err = "JPEG: bad SOS component"; return RT_ERR_INVALID_ARGUMENT; }
                if (!dc[td[k]].present || !ac[ta[k]].present || !have_quant[f->comp[ci[k]].tq]) { *err = "JPEG: scan uses a table that was never defined"; return RT_ERR_INVALID_ARGUMENT; }
            }
            if (s[1 + 2 * ns] != 0 || s[2 + 2 * ns] != 63 || s[3 + 2 * ns] != 0) { *err = "JPEG: spectral selection / successive approximation in a baseline scan"; return RT_ERR_UNSUPPORTED; }
            pos += len;
            if (!coef) {                                     // headers only: the frame is known, tables of the first scan checked
                return RT_OK;
            }
            if (capacity < f->coef_count) { *err = "rt_jpeg: coefficient buffer too small"; return RT_ERR_OUT_OF_RANGE; }
            if (!saw_scan) std::memset(coef, 0, f->coef_count * sizeof(int16_t));
            saw_scan = true;
            // T.81 A.2: an interleaved scan walks MCUs (h x v blocks of every component); a one-component scan walks the
            // component's own blocks, ceil(ds / 8) per row, with no padding to the MCU
            int units_x, units_y;
            if (ns > 1) {
                units_x = (f->width + 8 * f->hmax - 1) / (8 * f->hmax);
                units_y = (f->height + 8 * f->vmax - 1) / (8 * f->vmax);
            } else {
                units_x = (f->comp[ci[0]].ds_w + 7) / 8;
                units_y = (f->comp[ci[0]].ds_h + 7) / 8;
            }
            BitReader br{data + pos, data + n};
            int pred[3] = {0, 0, 0};
            int until_restart = restart_interval, next_rst = 0;
            bool bad = false;
            for (int uy = 0; uy < units_y; ++uy) {
                for (int ux = 0; ux < units_x; ++ux) {
                    if (restart_interval && until_restart == 0) {
                        // byte-align, expect RSTn
                        const uint8_t* q = br.p;
                        while (q + 1 < data + n && !(q[0] == 0xff && q[1] >= 0xd0 && q[1] <= 0xd7)) {
                            if (q[0] == 0xff && q[1] != 0x00 && q[1] != 0xff) break;
                            ++q;
                        }
                        if (q + 1 >= data + n || q[0] != 0xff || q[1] != 0xd0 + next_rst) { *err = "JPEG: restart marker missing"; return RT_ERR_INVALID_ARGUMENT; }
                        br.p = q + 2;
                        br.reset();
                        next_rst = (next_rst + 1) & 7;
                        pred[0] = pred[1] = pred[2] = 0;
                        until_restart = restart_interval;
                    }
                    for (int k = 0; k < ns; ++k) {
                        const JpegComponent& C = f->comp[ci[k]];
                        const int bh = ns > 1 ? C.h : 1, bv = ns > 1 ? C.v : 1;
                        const HuffTable& tdc = dc[td[k]];
                        const HuffTable& tac = ac[ta[k]];
                        for (int by = 0; by < bv; ++by) {
                            for (int bx = 0; bx < bh; ++bx) {
                                const int row = uy * bv + by, col = ux * bh + bx;
                                int16_t* blk = coef + C.coef_offset + ((size_t)row * C.blocks_w + col) * 64;
                                const int sdc = decode_symbol(br, tdc, &bad);
                                if (sdc > 11) bad = true;
                                if (bad) { *err = "JPEG: corrupt entropy-coded data"; return RT_ERR_INVALID_ARGUMENT; }
                                if (sdc) pred[k] += extend(br.get(sdc), sdc);
                                blk[0] = (int16_t)pred[k];
                                for (int z = 1; z < 64;) {
                                    const int rs = decode_symbol(br, tac, &bad);
                                    if (bad) { *err = "JPEG: corrupt entropy-coded data"; return RT_ERR_INVALID_ARGUMENT; }
                                    const int r = rs >> 4, sz = rs & 15;
                                    if (sz == 0) {
                                        if (r != 15) break;          // EOB
                                        z += 16;                     // ZRL
                                        continue;
                                    }
                                    z += r;
                                    if (z > 63) { *err = "JPEG: corrupt entropy-coded data"; return RT_ERR_INVALID_ARGUMENT; }
                                    blk[kZigzag[z]] = (int16_t)extend(br.get(sz), sz);
                                    ++z;
                                }
                            }
                        }
                    }
                    if (restart_interval) --until_restart;
                }
            }
            if (br.p + 1 >= data + n) { *err = "JPEG: the entropy-coded data is truncated (no marker follows the scan)"; return RT_ERR_INVALID_ARGUMENT; }
            for (int k = 0; k < ns; ++k) comp_done[ci[k]] = true;
            // continue after the scan's data: at the marker the reader stopped on, or search for it
            const uint8_t* q = br.p;
            while (q + 1 < data + n && !(q[0] == 0xff && q[1] != 0x00 && !(q[1] >= 0xd0 && q[1] <= 0xd7) && q[1] != 0xff)) ++q;
            pos = (size_t)(q - data);
            continue;
        }
        pos += len;
    }
    if (!have_frame) { *err = "JPEG: no frame header"; return RT_ERR_INVALID_ARGUMENT; }
    if (coef) {
        for (int c = 0; c < f->ncomp; ++c)
            if (!comp_done[c]) { *err = "JPEG: a component has no scan"; return RT_ERR_INVALID_ARGUMENT; }
    }
    return RT_OK;
}

}  // namespace rt_host

extern "C" int rt_jpeg_entropy_decode(const uint8_t* jpeg, size_t n_bytes, rt_jpeg_info* info, int16_t* coef, size_t capacity) {
    using namespace rt_host;
    if (!jpeg || !info) return fail(RT_ERR_INVALID_ARGUMENT, "rt_jpeg_entropy_decode: null argument");
    JpegFrame f;
    std::string err;
    const int rc = jpeg_entropy_decode(jpeg, n_bytes, &f, coef, capacity, &err);
    if (rc < 0) return fail(rc, "rt_jpeg_entropy_decode: " + err);
    std::memset(info, 0, sizeof(*info));
    info->width = f.width;
    info->height = f.height;
    info->components = f.ncomp;
    info->adobe_rgb = f.adobe_rgb ? 1 : 0;
    for (int c = 0; c < f.ncomp; ++c) {
        info->h_samp[c] = f.comp[c].h;
        info->v_samp[c] = f.comp[c].v;
        info->blocks_w[c] = f.comp[c].blocks_w;
        info->blocks_h[c] = f.comp[c].blocks_h;
        info->coef_offset[c] = (int64_t)f.comp[c].coef_offset;
        std::memcpy(info->quant[c], f.quant[f.comp[c].tq], sizeof(info->quant[c]));
    }
    info->coef_count = (int64_t)f.coef_count;
    return RT_OK;
}
