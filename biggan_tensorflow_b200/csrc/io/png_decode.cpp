// PNG -> uint8 HWC (see include/biggan_io.h).  Chunk walk with CRC check, one zlib inflate over the concatenated
// IDAT stream, the five scanline filters of the PNG specification, palette expansion, channel conversion.
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>
#include <vector>
#include "../../../include/biggan_io.h"

static thread_local char g_err[256];
static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}
extern "C" const char* bgio_last_error(void) { return g_err; }

static inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// header dimensions come from the file: callers size buffers from them, so they are bounded where they are parsed
static const int kMaxSide = 1 << 16;
struct PngHeader { int w, h, depth, color, interlace, channels; };
struct PngParts { PngHeader hdr; std::vector<uint8_t> idat; uint8_t palette[256][3]; int n_palette = 0; };

static int walk(const uint8_t* d, size_t n, PngParts* parts, bool header_only) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  if (n < 8 + 25 || memcmp(d, sig, 8) != 0) return fail("not a PNG file");
  size_t pos = 8;
  bool have_ihdr = false, done = false;
  while (!done) {
    if (pos + 12 > n) return fail("truncated PNG (chunk header at %zu)", pos);
    const uint32_t len = be32(d + pos);
    const uint8_t* type = d + pos + 4;
    if ((size_t)len > n - pos - 12) return fail("truncated PNG (chunk of %u bytes at %zu)", len, pos);
    const uint8_t* body = d + pos + 8;
    const uint32_t crc = be32(body + len);
    if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), type, len + 4) != crc) return fail("PNG chunk %.4s: CRC mismatch", type);
    if (!memcmp(type, "IHDR", 4)) {
      if (len != 13) return fail("bad IHDR length");
      PngHeader& h = parts->hdr;
      h.w = (int)be32(body); h.h = (int)be32(body + 4); h.depth = body[8]; h.color = body[9]; h.interlace = body[12];
      if (h.w <= 0 || h.h <= 0 || body[10] != 0 || body[11] != 0) return fail("unsupported PNG header");
      if (h.w > kMaxSide || h.h > kMaxSide) return fail("PNG dimensions %d x %d are out of range (1..%d)", h.w, h.h, kMaxSide);
      switch (h.color) {
        case 0: h.channels = 1; break;
        case 2: h.channels = 3; break;
        case 3: h.channels = 3; break;
        case 4: h.channels = 2; break;
        case 6: h.channels = 4; break;
        default: return fail("bad PNG colour type %d", h.color);
      }
      have_ihdr = true;
      if (header_only) return 0;
    } else if (!have_ihdr) {
      return fail("PNG does not start with IHDR");
    } else if (!memcmp(type, "PLTE", 4)) {
      if (len % 3 || len > 768) return fail("bad PLTE length");
      parts->n_palette = (int)len / 3;
      memcpy(parts->palette, body, len);
    } else if (!memcmp(type, "IDAT", 4)) {
      parts->idat.insert(parts->idat.end(), body, body + len);
    } else if (!memcmp(type, "IEND", 4)) {
      done = true;
    }
    pos += 12 + (size_t)len;
  }
  return have_ihdr ? 0 : fail("PNG has no IHDR");
}

extern "C" int bgio_png_info(const void* data, size_t n, int* width, int* height, int* channels, int* bit_depth) {
  PngParts p;
  if (walk((const uint8_t*)data, n, &p, true)) return -1;
  *width = p.hdr.w; *height = p.hdr.h; *channels = p.hdr.channels; *bit_depth = p.hdr.depth;
  return 0;
}

static inline int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

extern "C" int bgio_png_decode(const void* data, size_t n, void* out_v, size_t out_bytes, int out_channels) {
  PngParts p;
  if (walk((const uint8_t*)data, n, &p, false)) return -1;
  const PngHeader& h = p.hdr;
  // header dimensions come from the file: bound them before any size arithmetic (a crafted IHDR near 2^31 would wrap the
  // products below, and zlib's 32-bit avail_in / avail_out would truncate a > 4 GiB scanline buffer)
  if (h.w <= 0 || h.h <= 0 || h.w > kMaxSide || h.h > kMaxSide) return fail("PNG dimensions %d x %d are out of range (1..%d)", h.w, h.h, kMaxSide);
  if (h.interlace) return fail("interlaced PNG is not supported");
  const bool pal = h.color == 3;
  if (pal ? (h.depth != 1 && h.depth != 2 && h.depth != 4 && h.depth != 8) : h.depth != 8)
    return fail("PNG bit depth %d is not supported (8-bit images only)", h.depth);
  if (pal && p.n_palette == 0) return fail("palette PNG without PLTE");
  const int stored = pal ? 1 : h.channels;                       // samples per pixel in the scanlines
  const int oc = out_channels ? out_channels : h.channels;
  if (oc != 1 && oc != 2 && oc != 3 && oc != 4) return fail("bad channel count %d", out_channels);
  if (oc <= 2 && h.channels >= 3) return fail("RGB -> gray conversion is not supported");
  if (out_bytes < (size_t)h.w * h.h * oc) return fail("output buffer too small");
  const size_t row_bytes = pal ? ((size_t)h.w * h.depth + 7) / 8 : (size_t)h.w * stored;
  const int bpp = pal ? 1 : stored;                              // filter distance in bytes (>= 1)
  // (row_bytes + 1) <= 4 * 65536 + 1 and h <= 65536: the product fits size_t; it must also fit zlib's 32-bit counters
  const size_t raw_bytes = (row_bytes + 1) * (size_t)h.h;
  if (raw_bytes > 0xFFFFFFFFull || p.idat.size() > 0xFFFFFFFFull) return fail("PNG image data exceeds 4 GiB");
  std::vector<uint8_t> raw(raw_bytes);
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if (inflateInit(&zs) != Z_OK) return fail("zlib init failed");
  zs.next_in = p.idat.data(); zs.avail_in = (uInt)p.idat.size();
  zs.next_out = raw.data(); zs.avail_out = (uInt)raw.size();
  const int zr = inflate(&zs, Z_FINISH);
  const size_t got = raw.size() - zs.avail_out;
  inflateEnd(&zs);
  if (zr != Z_STREAM_END || got != raw.size()) return fail("PNG image data: inflate failed, short or overlong (%zu of %zu bytes, zlib %d)", got, raw.size(), zr);
  // unfilter in place
  std::vector<uint8_t> zero(row_bytes, 0);
  for (int y = 0; y < h.h; ++y) {
    uint8_t* cur = raw.data() + (size_t)y * (row_bytes + 1) + 1;
    const uint8_t* up = y ? cur - (row_bytes + 1) : zero.data();
    const int ft = cur[-1];
    switch (ft) {
      case 0: break;
      case 1: for (size_t i = bpp; i < row_bytes; ++i) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]); break;
      case 2: for (size_t i = 0; i < row_bytes; ++i) cur[i] = (uint8_t)(cur[i] + up[i]); break;
      case 3: for (size_t i = 0; i < row_bytes; ++i) cur[i] = (uint8_t)(cur[i] + (((i >= (size_t)bpp ? cur[i - bpp] : 0) + up[i]) >> 1)); break;
      case 4: for (size_t i = 0; i < row_bytes; ++i) cur[i] = (uint8_t)(cur[i] + paeth(i >= (size_t)bpp ? cur[i - bpp] : 0, up[i], i >= (size_t)bpp ? up[i - bpp] : 0)); break;
      default: return fail("bad PNG filter type %d in row %d", ft, y);
    }
  }
  // to the requested channel count
  uint8_t* out = (uint8_t*)out_v;
  for (int y = 0; y < h.h; ++y) {
    const uint8_t* row = raw.data() + (size_t)y * (row_bytes + 1) + 1;
    uint8_t* o = out + (size_t)y * h.w * oc;
    for (int x = 0; x < h.w; ++x, o += oc) {
      uint8_t px[4] = {0, 0, 0, 255};
      int have;                                                  // channels available in px: 1 gray, 2 gray+alpha, 3, 4
      if (pal) {
        const int per = 8 / h.depth, idx = (row[x / per] >> ((per - 1 - x % per) * h.depth)) & ((1 << h.depth) - 1);
        if (idx >= p.n_palette) return fail("palette index %d out of range", idx);
        px[0] = p.palette[idx][0]; px[1] = p.palette[idx][1]; px[2] = p.palette[idx][2];
        have = 3;
      } else {
        for (int c = 0; c < stored; ++c) px[c] = row[(size_t)x * stored + c];
        have = stored;
      }
      const bool gray = have <= 2;
      const uint8_t alpha = have == 2 ? px[1] : (have == 4 ? px[3] : 255);
      switch (oc) {
        case 1: o[0] = px[0]; break;
        case 2: o[0] = px[0]; o[1] = alpha; break;
        case 3: if (gray) { o[0] = o[1] = o[2] = px[0]; } else { o[0] = px[0]; o[1] = px[1]; o[2] = px[2]; } break;
        case 4: if (gray) { o[0] = o[1] = o[2] = px[0]; } else { o[0] = px[0]; o[1] = px[1]; o[2] = px[2]; } o[3] = alpha; break;
      }
    }
  }
  return 0;
}
