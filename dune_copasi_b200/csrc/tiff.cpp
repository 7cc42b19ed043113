#include "tiff.hpp"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iterator>

#include "util.hpp"

namespace dcb {

double TiffImage::operator()(double x, double y) const {
  auto pixel = [](float v) -> uint32_t { return v <= 0.0f ? 0u : (v >= 4294967040.0f ? 4294967295u : (uint32_t)v); };
  uint32_t px = pixel(x_res * ((float)x - x_off));
  uint32_t line = rows - pixel(y_res * ((float)y - y_off)) - 1u;   // wraps like the reference's uint32 arithmetic
  if (px > cols - 1) px = cols - 1;
  if (line > rows - 1) line = rows - 1;
  return values[(size_t)line * cols + px];
}

namespace {

struct Reader {
  const std::vector<unsigned char>& d;
  bool big;
  uint64_t get(size_t off, int n) const {
    if (off + n > d.size()) fail("TIFF file is truncated");
    uint64_t v = 0;
    for (int i = 0; i < n; ++i) v |= (uint64_t)d[off + (big ? n - 1 - i : i)] << (8 * i);
    return v;
  }
};

// TIFF LZW (compression 5): MSB-first codes of 9..12 bits, ClearCode 256, EndOfInformation 257, code width grows one
// code early ("early change", what every TIFF writer since 5.0 does)
void lzw_decode(const unsigned char* in, size_t n, std::vector<unsigned char>& out) {
  std::vector<std::vector<unsigned char>> table;
  auto reset = [&] {
    table.assign(258, {});
    for (int i = 0; i < 256; ++i) table[i] = {(unsigned char)i};
  };
  reset();
  int width = 9;
  long prev = -1;
  uint32_t acc = 0;
  int bits = 0;
  size_t pos = 0;
  for (;;) {
    while (bits < width && pos < n) { acc = (acc << 8) | in[pos++]; bits += 8; }
    if (bits < width) break;
    const uint32_t code = (acc >> (bits - width)) & ((1u << width) - 1);
    bits -= width;
    if (code == 257) break;
    if (code == 256) { reset(); width = 9; prev = -1; continue; }
    std::vector<unsigned char> entry;
    if (code < table.size()) entry = table[code];
    else if (prev >= 0 && code == table.size()) { entry = table[prev]; entry.push_back(table[prev][0]); }
    else fail("TIFF file: corrupt LZW stream");
    out.insert(out.end(), entry.begin(), entry.end());
    if (prev >= 0) {
      std::vector<unsigned char> add = table[prev];
      add.push_back(entry[0]);
      table.push_back(std::move(add));
    }
    prev = code;
    if (table.size() + 1 >= (1u << width) && width < 12) ++width;
  }
}

// Deflate (compression 8 / 32946) through the system's zlib, resolved at run time like NCCL: no link-time dependency
void inflate_strip(const unsigned char* in, size_t n, size_t expect, std::vector<unsigned char>& out) {
  typedef int (*uncompress_t)(unsigned char*, unsigned long*, const unsigned char*, unsigned long);
  static uncompress_t fn = [] {
    void* h = dlopen("libz.so.1", RTLD_NOW);
    if (!h) h = dlopen("libz.so", RTLD_NOW);
    return h ? (uncompress_t)dlsym(h, "uncompress") : (uncompress_t) nullptr;
  }();
  if (!fn) fail("TIFF file: Deflate-compressed strips need libz.so.1, which could not be loaded");
  std::vector<unsigned char> buf(expect);
  unsigned long len = (unsigned long)expect;
  if (fn(buf.data(), &len, in, (unsigned long)n) != 0) fail("TIFF file: corrupt Deflate stream");
  out.insert(out.end(), buf.begin(), buf.begin() + len);
}

}  // namespace

TiffImage read_tiff(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) fail("File '", path, "' does not exists.");   // tiff_file.cc:18
  std::vector<unsigned char> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (d.size() < 8 || !((d[0] == 'I' && d[1] == 'I') || (d[0] == 'M' && d[1] == 'M')))
    fail("Error opening TIFF file '", path, "'");
  Reader r{d, d[0] == 'M'};
  if (r.get(2, 2) != 42) fail("Error opening TIFF file '", path, "' (BigTIFF and other variants need libtiff)");
  size_t ifd = (size_t)r.get(4, 4);
  const int nent = (int)r.get(ifd, 2);
  TiffImage im;
  uint32_t compression = 1, photometric = 99, samples = 1, rows_per_strip = 0xffffffffu, planar = 1, predictor = 1;
  std::vector<uint64_t> offsets, counts;
  static const int tsize[13] = {0, 1, 1, 2, 4, 8, 1, 1, 2, 4, 8, 4, 8};
  for (int e = 0; e < nent; ++e) {
    const size_t p = ifd + 2 + (size_t)e * 12;
    const int tag = (int)r.get(p, 2), type = (int)r.get(p + 2, 2);
    const uint64_t count = r.get(p + 4, 4);
    if (type < 1 || type > 12) continue;
    const size_t at = (uint64_t)tsize[type] * count <= 4 ? p + 8 : (size_t)r.get(p + 8, 4);
    auto val = [&](uint64_t k) { return r.get(at + (size_t)k * tsize[type], tsize[type]); };
    auto rational = [&]() -> float {
      if (type == 5) { double n = (double)r.get(at, 4), q = (double)r.get(at + 4, 4); return q != 0 ? (float)(n / q) : 0.0f; }
      if (type == 11) { uint32_t b = (uint32_t)r.get(at, 4); float v; std::memcpy(&v, &b, 4); return v; }
      return (float)val(0);
    };
    switch (tag) {
      case 256: im.cols = (uint32_t)val(0); break;
      case 257: im.rows = (uint32_t)val(0); break;
      case 258: im.bits = (int)val(0); break;
      case 259: compression = (uint32_t)val(0); break;
      case 262: photometric = (uint32_t)val(0); break;
      case 273: for (uint64_t k = 0; k < count; ++k) offsets.push_back(val(k)); break;
      case 277: samples = (uint32_t)val(0); break;
      case 278: rows_per_strip = (uint32_t)val(0); break;
      case 279: for (uint64_t k = 0; k < count; ++k) counts.push_back(val(k)); break;
      case 282: im.x_res = rational(); break;
      case 283: im.y_res = rational(); break;
      case 284: planar = (uint32_t)val(0); break;
      case 286: im.x_off = rational(); break;
      case 317: predictor = (uint32_t)val(0); break;
      case 287: im.y_off = rational(); break;
      default: break;
    }
  }
  if (photometric != 0 && photometric != 1) fail("TIFF file '", path, "' must be in grayscale");   // tiff_file.cc:30
  im.zero = photometric != 0;
  if (!(im.x_res > 0.0f) || !(im.y_res > 0.0f)) fail("TIFF file '", path, "' has negative resolution");   // :39
  if (im.bits != 8 && im.bits != 16 && im.bits != 32 && im.bits != 64)
    fail("Encoding with ", im.bits, " bits not implemented");   // tiff_grayscale.cc:79
  if (samples != 1 || planar != 1) fail("TIFF file '", path, "': only one sample per pixel is read by this build");
  if (compression != 1 && compression != 32773 && compression != 5 && compression != 8 && compression != 32946)
    fail("TIFF file '", path, "': compression ", compression, " needs libtiff (this build reads uncompressed, PackBits, LZW and Deflate strips)");
  if (predictor != 1 && predictor != 2) fail("TIFF file '", path, "': predictor ", predictor, " needs libtiff");
  if (im.rows == 0 || im.cols == 0 || offsets.empty() || offsets.size() != counts.size()) fail("TIFF file '", path, "' has no image data");
  if (rows_per_strip > im.rows) rows_per_strip = im.rows;
  const size_t bpp = im.bits / 8, line = (size_t)im.cols * bpp;
  std::vector<unsigned char> raw;
  raw.reserve((size_t)im.rows * line);
  for (size_t s = 0; s < offsets.size(); ++s) {
    const size_t b = (size_t)offsets[s], n = (size_t)counts[s];
    if (b + n > d.size()) fail("TIFF file '", path, "' is truncated");
    const size_t strip_rows = std::min<size_t>(rows_per_strip, im.rows - s * (size_t)rows_per_strip);
    if (compression == 1) {
      raw.insert(raw.end(), d.begin() + b, d.begin() + b + n);
    } else if (compression == 5) {
      lzw_decode(d.data() + b, n, raw);
    } else if (compression == 8 || compression == 32946) {
      inflate_strip(d.data() + b, n, strip_rows * line, raw);
    } else {   // PackBits
      size_t i = b;
      while (i < b + n) {
        const int c = (signed char)d[i++];
        if (c >= 0) { for (int k = 0; k <= c && i < b + n; ++k) raw.push_back(d[i++]); }
        else if (c != -128) { if (i < b + n) { raw.insert(raw.end(), (size_t)(1 - c), d[i]); ++i; } }
      }
    }
  }
  if (raw.size() < (size_t)im.rows * line) fail("TIFF file '", path, "' holds fewer pixels than its header says");
  if (predictor == 2) {   // horizontal differencing: every sample is stored as the difference to its left neighbour
    Reader in{raw, r.big};
    for (size_t row = 0; row < im.rows; ++row)
      for (size_t col = 1; col < im.cols; ++col) {
        const size_t at = (row * im.cols + col) * bpp;
        const uint64_t v = in.get(at, (int)bpp) + in.get(at - bpp, (int)bpp);
        for (size_t k = 0; k < bpp; ++k) raw[at + (r.big ? bpp - 1 - k : k)] = (unsigned char)(v >> (8 * k));
      }
  }
  const double maxv = std::ldexp(1.0, im.bits);   // std::size_t{1} << bits in the reference (2^64 wraps there)
  im.values.resize((size_t)im.rows * im.cols);
  Reader px{raw, r.big};
  for (size_t k = 0; k < im.values.size(); ++k) {
    const double v = (double)px.get(k * bpp, (int)bpp);
    im.values[k] = (im.zero ? v : maxv - v) / maxv;
  }
  return im;
}

}  // namespace dcb
