// Device half of rt_jpeg_decode (ImageTexture::new's decode, texture.rs:76-80): quantised coefficients -> RGB8.
//
// The target is byte equality with libjpeg-turbo's default decode - what PIL hands to every other test and to the
// oracle - so the arithmetic is libjpeg's, integer for integer:
//   jpeg_idct_kernel     dequantisation + the "slow but accurate" integer inverse DCT (jidctint.c jpeg_idct_islow:
//                        Loeffler-Ligtenberg-Moschytz, 13-bit constants, 2 extra bits kept between the passes), one
//                        thread per 8x8 block, both passes in registers;
//   jpeg_colour_kernel   "fancy" chroma upsampling (jdsample.c h2v1_fancy_upsample / h2v2_fancy_upsample: triangle
//                        filter, 3/4 nearer + 1/4 further sample, alternating rounding bias) fused with YCbCr -> RGB
//                        (jdcolor.c ycc_rgb_convert: 16-bit fixed point), one thread per output pixel pair.
// Both are pure streaming kernels, bound by HBM: per 8x8 block 128 B of coefficients in and 64 B of samples out; per
// pixel 1 + 2/(h v) sample bytes in (neighbours come from L1/L2) and 3 B out.
//
// Included by rt_cuda.cu.
#pragma once

namespace {

struct JpegPlane {
    const int16_t* coef;     // blocks_h x blocks_w x 64
    uint8_t* samples;        // (blocks_h * 8) x pitch
    int blocks_w, blocks_h, pitch;
    int ds_w, ds_h;          // real samples (the rest of the plane is block padding)
    int h, v;                // sampling factors
};

__constant__ uint16_t c_jpeg_quant[3][64];

#define JFIX_0_298631336 2446
#define JFIX_0_390180644 3196
#define JFIX_0_541196100 4433
#define JFIX_0_765366865 6270
#define JFIX_0_899976223 7373
#define JFIX_1_175875602 9633
#define JFIX_1_501321110 12299
#define JFIX_1_847759065 15137
#define JFIX_1_961570560 16069
#define JFIX_2_053119869 16819
#define JFIX_2_562915447 20995
#define JFIX_3_072711026 25172

__device__ __forceinline__ int jdescale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// one 8-point pass of jpeg_idct_islow; `shift` is the descale of the pass
__device__ __forceinline__ void idct8(const int in[8], int out[8], int shift) {
    // even part
    int z2 = in[2], z3 = in[6];
    int z1 = (z2 + z3) * JFIX_0_541196100;
    const int tmp2 = z1 + z3 * (-JFIX_1_847759065);
    const int tmp3 = z1 + z2 * JFIX_0_765366865;
    z2 = in[0];
    z3 = in[4];
    const int tmp0 = (z2 + z3) << 13;
    const int tmp1 = (z2 - z3) << 13;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    // odd part
    int t0 = in[7], t1 = in[5], t2 = in[3], t3 = in[1];
    z1 = t0 + t3;
    z2 = t1 + t2;
    z3 = t0 + t2;
    int z4 = t1 + t3;
    const int z5 = (z3 + z4) * JFIX_1_175875602;
    t0 *= JFIX_0_298631336;
    t1 *= JFIX_2_053119869;
    t2 *= JFIX_3_072711026;
    t3 *= JFIX_1_501321110;
    z1 *= -JFIX_0_899976223;
    z2 *= -JFIX_2_562915447;
    z3 *= -JFIX_1_961570560;
    z4 *= -JFIX_0_390180644;
    z3 += z5;
    z4 += z5;
    t0 += z1 + z3;
    t1 += z2 + z4;
    t2 += z2 + z3;
    t3 += z1 + z4;
    out[0] = jdescale(tmp10 + t3, shift);
    out[7] = jdescale(tmp10 - t3, shift);
    out[1] = jdescale(tmp11 + t2, shift);
    out[6] = jdescale(tmp11 - t2, shift);
    out[2] = jdescale(tmp12 + t1, shift);
    out[5] = jdescale(tmp12 - t1, shift);
    out[3] = jdescale(tmp13 + t0, shift);
    out[4] = jdescale(tmp13 - t0, shift);
}

struct JpegIdctArgs { JpegPlane plane[3]; int first_cta[4]; };   // CTAs [first_cta[c], first_cta[c + 1]) work on component c

__global__ void __launch_bounds__(128) jpeg_idct_kernel(JpegIdctArgs A) {
    const int comp = (int)blockIdx.x >= A.first_cta[2] ? 2 : ((int)blockIdx.x >= A.first_cta[1] ? 1 : 0);
    const JpegPlane& P = A.plane[comp];
    const int b = ((int)blockIdx.x - A.first_cta[comp]) * blockDim.x + threadIdx.x;
    if (b >= P.blocks_w * P.blocks_h) return;
    const uint4* src = reinterpret_cast<const uint4*>(P.coef + (size_t)b * 64);
    int ws[8][8];
    // pass 1: columns. Row r of the block arrives as one 16-byte load; dequantise (DEQUANTIZE = coefficient x step).
    int blk[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint4 q = __ldg(src + r);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            blk[r][2 * k] = (int)(int16_t)(w[k] & 0xffffu) * (int)c_jpeg_quant[comp][r * 8 + 2 * k];
            blk[r][2 * k + 1] = (int)(int16_t)(w[k] >> 16) * (int)c_jpeg_quant[comp][r * 8 + 2 * k + 1];
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        int in[8], out[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) in[r] = blk[r][c];
        idct8(in, out, 13 - 2);                  // CONST_BITS - PASS1_BITS
#pragma unroll
        for (int r = 0; r < 8; ++r) ws[r][c] = out[r];
    }
    // pass 2: rows; descale by CONST_BITS + PASS1_BITS + 3, add the level shift, clamp (the range_limit table)
    const int by = b / P.blocks_w, bx = b - by * P.blocks_w;
    uint8_t* dst = P.samples + (size_t)(by * 8) * P.pitch + bx * 8;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int out[8];
        idct8(ws[r], out, 13 + 2 + 3);
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            lo |= (uint32_t)min(255, max(0, out[c] + 128)) << (8 * c);
            hi |= (uint32_t)min(255, max(0, out[c + 4] + 128)) << (8 * c);
        }
        *reinterpret_cast<uint2*>(dst + (size_t)r * P.pitch) = make_uint2(lo, hi);
    }
}

// jdsample.c: the two output samples a chroma sample `c` of row sums expands to. Row sums are 3 * nearer + further row
// (h2v2, 4 bits of fraction) or the plain sample (h2v1, handled by the caller with 2 bits).
__device__ __forceinline__ int chroma_h2v2(const uint8_t* near_row, const uint8_t* far_row, int x_out, int ds_w) {
    const int c = x_out >> 1;
    const int this_sum = 3 * near_row[c] + far_row[c];
    if (x_out & 1) {
        if (c == ds_w - 1) return (this_sum * 4 + 7) >> 4;
        const int next_sum = 3 * near_row[c + 1] + far_row[c + 1];
        return (this_sum * 3 + next_sum + 7) >> 4;
    }
    if (c == 0) return (this_sum * 4 + 8) >> 4;
    const int last_sum = 3 * near_row[c - 1] + far_row[c - 1];
    return (this_sum * 3 + last_sum + 8) >> 4;
}
__device__ __forceinline__ int chroma_h2v1(const uint8_t* row, int x_out, int ds_w) {
    const int c = x_out >> 1;
    if (x_out & 1) return c == ds_w - 1 ? row[c] : (3 * row[c] + row[c + 1] + 2) >> 2;
    return c == 0 ? row[c] : (3 * row[c] + row[c - 1] + 1) >> 2;
}

// mode: 0 = grey, 1 = 4:4:4, 2 = 4:2:2 (h2v1), 3 = 4:2:0 (h2v2); rgb_passthrough: the three planes are R, G, B (Adobe)
struct JpegColourArgs {
    JpegPlane Y, Cb, Cr;
    int mode, rgb_passthrough, width, height;
};

__device__ __forceinline__ uint32_t jpeg_pixel(const JpegColourArgs& A, int x, int y) {
    const JpegPlane& Cb = A.Cb;
    const JpegPlane& Cr = A.Cr;
    const int yy = A.Y.samples[(size_t)y * A.Y.pitch + x];
    int cb = 128, cr = 128;
    if (A.mode == 1) {
        cb = Cb.samples[(size_t)y * Cb.pitch + x];
        cr = Cr.samples[(size_t)y * Cr.pitch + x];
    } else if (A.mode == 2) {
        const uint8_t* rb = Cb.samples + (size_t)y * Cb.pitch;
        const uint8_t* rr = Cr.samples + (size_t)y * Cr.pitch;
        if (Cb.ds_w > 2) { cb = chroma_h2v1(rb, x, Cb.ds_w); cr = chroma_h2v1(rr, x, Cr.ds_w); }   // jdsample.c: fancy only when
        else { cb = rb[x >> 1]; cr = rr[x >> 1]; }                                                 // downsampled_width > 2
    } else if (A.mode == 3) {
        const int cy = y >> 1;
        // the further row: above for even output rows, below for odd ones; at the image edge the edge row itself
        // (jdmainct.c duplicates the first / last real sample row as context)
        const int fy = (y & 1) ? min(cy + 1, Cb.ds_h - 1) : max(cy - 1, 0);
        const uint8_t* nb = Cb.samples + (size_t)cy * Cb.pitch;
        const uint8_t* fb = Cb.samples + (size_t)fy * Cb.pitch;
        const uint8_t* nr = Cr.samples + (size_t)cy * Cr.pitch;
        const uint8_t* fr = Cr.samples + (size_t)fy * Cr.pitch;
        if (Cb.ds_w > 2) { cb = chroma_h2v2(nb, fb, x, Cb.ds_w); cr = chroma_h2v2(nr, fr, x, Cr.ds_w); }
        else { cb = nb[x >> 1]; cr = nr[x >> 1]; }
    }
    int r, g, b;
    if (A.mode == 0) {
        r = g = b = yy;
    } else if (A.rgb_passthrough) {
        r = yy; g = cb; b = cr;
    } else {
        // jdcolor.c build_ycc_rgb_table / ycc_rgb_convert, SCALEBITS = 16
        const int u = cb - 128, w = cr - 128;
        r = yy + ((91881 * w + 32768) >> 16);
        g = yy + ((-22554 * u + 32768 - 46802 * w) >> 16);
        b = yy + ((116130 * u + 32768) >> 16);
        r = min(255, max(0, r)); g = min(255, max(0, g)); b = min(255, max(0, b));
    }
    return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16);
}

__device__ __forceinline__ uint32_t jpeg_ycc_rgb(int yy, int cb, int cr) {   // jdcolor.c ycc_rgb_convert, packed r | g << 8 | b << 16
    const int u = cb - 128, w = cr - 128;
    const int r = min(255, max(0, yy + ((91881 * w + 32768) >> 16)));
    const int g = min(255, max(0, yy + ((-22554 * u + 32768 - 46802 * w) >> 16)));
    const int b = min(255, max(0, yy + ((116130 * u + 32768) >> 16)));
    return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16);
}

// Four chroma samples [c0, c0 + 4) of a row as one 32-bit load plus the two neighbours the triangle filter needs, the
// neighbour index clamped at the row's ends: with prev = this (next = this) the filter's general form gives exactly
// jdsample.c's special first / last column ((4 this + 8) >> 4, (4 this + 7) >> 4; h2v1: this).
__device__ __forceinline__ void chroma_window(const uint8_t* row, int c0, int ds_w, int v[6]) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(row + c0);     // pitch and c0 are multiples of 4
    v[1] = (int)(w & 0xffu); v[2] = (int)((w >> 8) & 0xffu); v[3] = (int)((w >> 16) & 0xffu); v[4] = (int)(w >> 24);
    v[0] = c0 > 0 ? (int)row[c0 - 1] : v[1];
    v[5] = c0 + 4 < ds_w ? (int)row[c0 + 4] : v[4];
    // a row whose real samples end inside this window: the last real sample is its own neighbour
    if (c0 + 3 >= ds_w) {
#pragma unroll
        for (int k = 2; k <= 4; ++k) if (c0 + k - 1 >= ds_w) v[k] = v[k - 1];
    }
}

// Fast form for widths that are a multiple of 8 (the reference's 6400 x 3200 asset): one thread per eight pixels of a row -
// 8 luma bytes as one 64-bit load, the chroma of both planes through chroma_window, 24 output bytes as three 64-bit
// stores. Byte for byte the arithmetic of jpeg_pixel.
__global__ void __launch_bounds__(256) jpeg_colour8_kernel(JpegColourArgs A, uint8_t* __restrict__ rgb) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8, y = blockIdx.y;
    if (x0 >= A.width) return;
    const uint2 yw = *reinterpret_cast<const uint2*>(A.Y.samples + (size_t)y * A.Y.pitch + x0);
    int yy[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) { yy[k] = (int)((yw.x >> (8 * k)) & 0xffu); yy[4 + k] = (int)((yw.y >> (8 * k)) & 0xffu); }
    int cb[8], cr[8];
    if (A.mode == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { cb[k] = 128; cr[k] = 128; }
    } else if (A.mode == 1) {
        const uint2 bw = *reinterpret_cast<const uint2*>(A.Cb.samples + (size_t)y * A.Cb.pitch + x0);
        const uint2 rw = *reinterpret_cast<const uint2*>(A.Cr.samples + (size_t)y * A.Cr.pitch + x0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            cb[k] = (int)((bw.x >> (8 * k)) & 0xffu); cb[4 + k] = (int)((bw.y >> (8 * k)) & 0xffu);
            cr[k] = (int)((rw.x >> (8 * k)) & 0xffu); cr[4 + k] = (int)((rw.y >> (8 * k)) & 0xffu);
        }
    } else {
        const int c0 = x0 >> 1;
#pragma unroll
        for (int plane = 0; plane < 2; ++plane) {
            const JpegPlane& C = plane ? A.Cr : A.Cb;
            int* out = plane ? cr : cb;
            int s[6];
            if (A.mode == 2) {          // h2v1: the samples themselves, 2 fraction bits
                chroma_window(C.samples + (size_t)y * C.pitch, c0, C.ds_w, s);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    out[2 * k] = (3 * s[k + 1] + s[k] + 1) >> 2;
                    out[2 * k + 1] = (3 * s[k + 1] + s[k + 2] + 2) >> 2;
                }
            } else {                    // h2v2: column sums 3 * nearer row + further row, 4 fraction bits
                const int cy = y >> 1;
                const int fy = (y & 1) ? min(cy + 1, C.ds_h - 1) : max(cy - 1, 0);
                int f[6];
                chroma_window(C.samples + (size_t)cy * C.pitch, c0, C.ds_w, s);
                chroma_window(C.samples + (size_t)fy * C.pitch, c0, C.ds_w, f);
#pragma unroll
                for (int k = 0; k < 6; ++k) s[k] = 3 * s[k] + f[k];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    out[2 * k] = (3 * s[k + 1] + s[k] + 8) >> 4;
                    out[2 * k + 1] = (3 * s[k + 1] + s[k + 2] + 7) >> 4;
                }
            }
        }
    }
    uint32_t px[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (A.mode == 0) px[k] = (uint32_t)yy[k] * 0x010101u;
        else if (A.rgb_passthrough) px[k] = (uint32_t)yy[k] | ((uint32_t)cb[k] << 8) | ((uint32_t)cr[k] << 16);
        else px[k] = jpeg_ycc_rgb(yy[k], cb[k], cr[k]);
    }
    uint2* o = reinterpret_cast<uint2*>(rgb + ((size_t)y * A.width + x0) * 3);     // 24-byte groups: 8-byte aligned
    o[0] = make_uint2(px[0] | (px[1] << 24), (px[1] >> 8) | (px[2] << 16));
    o[1] = make_uint2((px[2] >> 16) | (px[3] << 8), px[4] | (px[5] << 24));
    o[2] = make_uint2((px[5] >> 8) | (px[6] << 16), (px[6] >> 16) | (px[7] << 8));
}

// General form (any width): one thread per four pixels of a row; 12 output bytes = three aligned 32-bit stores when the
// width is a multiple of 4.
__global__ void __launch_bounds__(256) jpeg_colour_kernel(JpegColourArgs A, uint8_t* __restrict__ rgb) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x0 >= A.width) return;
    uint8_t* o = rgb + ((size_t)y * A.width + x0) * 3;
    if ((A.width & 3) == 0) {
        const uint32_t p0 = jpeg_pixel(A, x0, y), p1 = jpeg_pixel(A, x0 + 1, y), p2 = jpeg_pixel(A, x0 + 2, y), p3 = jpeg_pixel(A, x0 + 3, y);
        uint32_t* w = reinterpret_cast<uint32_t*>(o);
        w[0] = p0 | (p1 << 24);
        w[1] = (p1 >> 8) | (p2 << 16);
        w[2] = (p2 >> 16) | (p3 << 8);
    } else {
        for (int k = 0; k < 4 && x0 + k < A.width; ++k) {
            const uint32_t p = jpeg_pixel(A, x0 + k, y);
            o[3 * k] = (uint8_t)p; o[3 * k + 1] = (uint8_t)(p >> 8); o[3 * k + 2] = (uint8_t)(p >> 16);
        }
    }
}

}  // namespace
