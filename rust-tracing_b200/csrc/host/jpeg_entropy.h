// Host half of rt_jpeg_decode: marker parsing and the entropy (Huffman) decode of a baseline JPEG into quantised DCT
// coefficients. Entropy decoding is one serial bit stream (a code's position depends on every code before it), so it
// stays on the host; everything after it - dequantisation, the inverse DCT, chroma upsampling, YCbCr -> RGB - is
// independent per block / per pixel and runs on the device (csrc/device/jpeg_kernels.cuh).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace rt_host {

struct JpegComponent {
    int id = 0, h = 1, v = 1, tq = 0;
    int blocks_w = 0, blocks_h = 0;      // blocks stored (padded to whole MCUs): the coefficient plane is blocks_h x blocks_w x 64
    int ds_w = 0, ds_h = 0;              // ceil(width * h / hmax), ceil(height * v / vmax): the real samples of the component
    size_t coef_offset = 0;              // first coefficient of the component, in int16 units
};

struct JpegFrame {
    int width = 0, height = 0, ncomp = 0, hmax = 1, vmax = 1;
    JpegComponent comp[3];
    uint16_t quant[4][64];               // natural (row-major) order
    bool adobe_rgb = false;              // Adobe APP14 with transform 0: the three components are R, G, B already
    size_t coef_count = 0;
};

// Parses headers into `f`; when `coef` is not null also decodes every scan into it (int16, natural order within a block,
// blocks row-major per component plane; capacity in int16 units). Returns 0 or a negative rt_status, message in *err.
int jpeg_entropy_decode(const uint8_t* data, size_t n, JpegFrame* f, int16_t* coef, size_t capacity, std::string* err);

}  // namespace rt_host
