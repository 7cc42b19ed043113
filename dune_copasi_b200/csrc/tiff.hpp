// Grayscale TIFF images as two-argument parser_context functions (`type = tiff`, src/dune/copasi/parser/context.cc:66-71,
// dune/copasi/common/tiff_grayscale.hh, src/dune/copasi/common/tiff_{file,grayscale}.cc).  The reference reads the files
// through libtiff (not in this image): this is an own reader for the baseline layouts its images use -- one sample per
// pixel, 8 / 16 / 32 / 64 bits, strips, uncompressed / PackBits / LZW / Deflate (zlib resolved at run time), optional
// horizontal predictor; anything else (tiles, JPEG, several samples) fails loudly.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace dcb {

struct TiffImage {
  uint32_t rows = 0, cols = 0;          // ImageLength, ImageWidth (the reference's row_size, col_size)
  int bits = 0;
  bool zero = true;                     // PhotometricInterpretation != 0 (MinIsBlack): value = raw / 2^bits
  float x_res = 0, y_res = 0, x_off = 0, y_off = 0;
  std::vector<double> values;           // [rows][cols], already scaled: (zero ? raw : 2^bits - raw) / 2^bits
  // TIFFGrayscale::operator()(x, y), tiff_grayscale.cc:91-105: pixel column from x, scanline from the top by y,
  // float arithmetic, truncation to uint32, clamped to the image.  (Negative offsets are undefined behaviour in the
  // reference's cast; here they clamp to pixel 0.)
  double operator()(double x, double y) const;
};

TiffImage read_tiff(const std::string& path);

}  // namespace dcb
