/* oracle/tile_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's per-tile encode path (palettised tile -> BCn payload), one
 * function per reference function it follows.  The 4x4 block arithmetic is squish::Compress from
 * squish_oracle.cpp.  Exposed as a plain C ABI so tests/ (ctypes) and bench.py's cpu_baseline leg
 * can call it; the product never links this file.
 *
 * Follows (reference file:line):
 *   tco_required_space   ConverterDxt::getRequiredSpace / getPaddedValue  converter_dxt.cpp:41-62
 *   expand_pixel         Colors::palToARGB                               colors.cpp:42-59
 *   gather_block         ConverterDxt::encodeTile gather + PadBlock      converter_dxt.cpp:126-133, converter.cpp:62-97
 *                        + ReorderColors(ARGB->ABGR), intended semantics  converter_dxt.cpp:219-220, converter.cpp:127,160-168
 *   squish_flags         ConverterDxt::getFlags                          converter_dxt.cpp:179-212
 *   tco_encode_tile      ConverterDxt::convert (encode) + encodeTile     converter_dxt.cpp:65-77,109-144
 *   tco_decode_tile      ConverterDxt::decodeTile + decodeBlockDxt1/3/5  converter_dxt.cpp:147-176,251-430
 *   tco_encode_tiles     TileData::encode with -u semantics, N threads   tiledata.cpp:111-162 (CPU baseline "B")
 */
#include "squish.h"

#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace {

std::mutex g_totalsMutex;
SquishOracleCounters g_totals;

void addTotals(const SquishOracleCounters& c)
{
  std::lock_guard<std::mutex> lock(g_totalsMutex);
  g_totals.blocks += c.blocks; g_totals.single_blocks += c.single_blocks;
  g_totals.range_blocks += c.range_blocks; g_totals.cluster_blocks += c.cluster_blocks;
  g_totals.cand3 += c.cand3; g_totals.cand4 += c.cand4;
  g_totals.sweeps3 += c.sweeps3; g_totals.sweeps4 += c.sweeps4; g_totals.points += c.points;
}

inline int padded(int v) { return (v + 3) & ~3; }

/* colors.cpp:45-54 followed by the byte0<->byte2 swap of converter_dxt.cpp:220: writes R,G,B,A. */
inline void expand_pixel(const uint8_t* palette, uint8_t index, bool keyed, uint8_t* rgba)
{
  if (index != 0 || !keyed) {
    rgba[0] = palette[4 * index + 2];
    rgba[1] = palette[4 * index + 1];
    rgba[2] = palette[4 * index + 0];
    rgba[3] = 255;
  } else {
    rgba[0] = rgba[1] = rgba[2] = rgba[3] = 0;
  }
}

/* palette entry 0 == bytes 00 FF 00 00 (u32le 0x0000ff00) marks index 0 as transparent: colors.cpp:47 */
inline bool palette_is_keyed(const uint8_t* palette)
{
  return palette[0] == 0x00 && palette[1] == 0xff && palette[2] == 0x00 && palette[3] == 0x00;
}

/* One 4x4 block at pixel (x0,y0) of a w*h tile whose index rows have pitch w; pixels outside the
 * tile are {0,0,0,0} (PadBlock with useCopy=false). */
void gather_block(const uint8_t* palette, const uint8_t* indexed, int w, int h, int x0, int y0, uint8_t rgba[64])
{
  const bool keyed = palette_is_keyed(palette);
  for (int py = 0; py < 4; ++py)
    for (int px = 0; px < 4; ++px) {
      uint8_t* dst = rgba + 4 * (4 * py + px);
      int x = x0 + px, y = y0 + py;
      if (x < w && y < h) expand_pixel(palette, indexed[y * w + x], keyed, dst);
      else dst[0] = dst[1] = dst[2] = dst[3] = 0;
    }
}

int squish_flags(int type, int quality)
{
  int flags;
  switch (type) {
    case 1: flags = squish::kDxt1; break;
    case 2: flags = squish::kDxt3; break;
    case 3: flags = squish::kDxt5; break;
    default: return 0;
  }
  switch (quality) {
    case 0: case 1: case 2: flags |= squish::kColourRangeFit; break;
    case 3: case 4:         flags |= squish::kColourClusterFit; break;
    case 5: case 6:         flags |= squish::kColourClusterFit | squish::kWeightColourByAlpha; break;
    case 7: case 8: case 9: flags |= squish::kColourIterativeClusterFit | squish::kWeightColourByAlpha; break;
    default: break;
  }
  return flags;
}

/* converter_dxt.cpp:413-429 */
void unpack565(const uint8_t* src, uint8_t* dst)
{
  unsigned c1 = src[0] | (src[1] << 8), c2 = src[2] | (src[3] << 8);
  dst[0] = (uint8_t)(((c1 << 3) & 0xf8) | ((c1 >> 2) & 0x07));
  dst[1] = (uint8_t)(((c1 >> 3) & 0xfc) | ((c1 >> 9) & 0x03));
  dst[2] = (uint8_t)(((c1 >> 8) & 0xf8) | ((c1 >> 13) & 0x07));
  dst[4] = (uint8_t)(((c2 << 3) & 0xf8) | ((c2 >> 2) & 0x07));
  dst[5] = (uint8_t)(((c2 >> 3) & 0xfc) | ((c2 >> 9) & 0x03));
  dst[6] = (uint8_t)(((c2 >> 8) & 0xf8) | ((c2 >> 13) & 0x07));
  dst[3] = dst[7] = 255;
}

/* Decodes the colour half of a block into B,G,R,(A) pixels; `bc1` selects the c0<=c1 3-colour mode. */
void decode_colour(const uint8_t* src, bool bc1, uint8_t out[64])
{
  uint8_t e[8];
  unsigned c0 = src[0] | (src[1] << 8), c1 = src[2] | (src[3] << 8);
  unpack565(src, e);
  uint32_t code = src[4] | (src[5] << 8) | (src[6] << 16) | ((uint32_t)src[7] << 24);
  const bool four = !bc1 || c0 > c1;
  for (int i = 0; i < 16; ++i, code >>= 2) {
    uint8_t* d = out + 4 * i;
    switch (code & 3) {
      case 0: d[0] = e[0]; d[1] = e[1]; d[2] = e[2]; if (bc1) d[3] = 255; break;
      case 1: d[0] = e[4]; d[1] = e[5]; d[2] = e[6]; if (bc1) d[3] = 255; break;
      case 2:
        if (four) for (int c = 0; c < 3; ++c) d[c] = (uint8_t)(((e[c] << 1) + e[4 + c]) / 3);
        else      for (int c = 0; c < 3; ++c) d[c] = (uint8_t)((e[c] + e[4 + c]) >> 1);
        if (bc1) d[3] = 255;
        break;
      default:
        if (four) { for (int c = 0; c < 3; ++c) d[c] = (uint8_t)((e[c] + (e[4 + c] << 1)) / 3); if (bc1) d[3] = 255; }
        else      { d[0] = d[1] = d[2] = d[3] = 0; }
        break;
    }
  }
}

/* referenceQuirk: the reference's decodeBlockDxt5 takes its two alpha endpoints and the 48 index bits
 * from src[2..9] instead of src[0..7] (converter_dxt.cpp:351-359 loads the control word at &src[2] and
 * then peels alpha[0], alpha[1] off it).  That is a decode-side defect of the reference (out of scope
 * here); the oracle decodes BC3 alpha per the format unless asked to mimic it, which the tests do when
 * they compare with the reference's own decoder. */
void decode_block(const uint8_t* src, int type, uint8_t out[64], bool referenceQuirk)
{
  if (type == 1) { decode_colour(src, true, out); return; }
  decode_colour(src + 8, false, out);
  if (type == 2) {
    uint64_t alpha; std::memcpy(&alpha, src, 8);
    for (int i = 0; i < 16; ++i, alpha >>= 4)
      out[4 * i + 3] = (uint8_t)((alpha & 0x0f) | ((alpha & 0x0f) << 4));
  } else {
    uint8_t a[8];
    const uint8_t* as = referenceQuirk ? src + 2 : src;
    a[0] = as[0]; a[1] = as[1];
    if (a[0] > a[1]) {
      for (int k = 1; k < 7; ++k) a[1 + k] = (uint8_t)(((7 - k) * (unsigned)a[0] + k * (unsigned)a[1]) / 7);
    } else {
      for (int k = 1; k < 5; ++k) a[1 + k] = (uint8_t)(((5 - k) * (unsigned)a[0] + k * (unsigned)a[1]) / 5);
      a[6] = 0; a[7] = 255;
    }
    uint64_t ctrl = 0;
    for (int k = 0; k < 6; ++k) ctrl |= (uint64_t)as[2 + k] << (8 * k);
    for (int i = 0; i < 16; ++i, ctrl >>= 3) out[4 * i + 3] = a[ctrl & 7];
  }
}

}  // namespace

extern "C" {

int tco_required_space(int type, int w, int h)
{
  if (w <= 0 || h <= 0) return 0;
  int blocks = (padded(w) * padded(h)) >> 4;
  if (type == 1) return blocks << 3;
  if (type == 2 || type == 3) return blocks << 4;
  return 0;
}

int tco_squish_flags(int type, int quality) { return squish_flags(type, quality); }

/* 16 R,G,B,A pixels -> one block, flags as squish's.  Thin export for block-level tests. */
void tco_compress_block(const uint8_t* rgba, int flags, uint8_t* out)
{
  squish::Compress(rgba, out, flags);
}

/* The 64-byte R,G,B,A block squish sees for block (bx,by) of a tile (tests of expand/pad/swizzle). */
void tco_gather_block(const uint8_t* palette, const uint8_t* indexed, int w, int h, int bx, int by, uint8_t* rgba)
{
  gather_block(palette, indexed, w, h, 4 * bx, 4 * by, rgba);
}

/* Colors::palToARGB restated: `size` pixels, output memory order B,G,R,A (colors.cpp:42-59). */
int tco_pal_to_argb(const uint8_t* indexed, const uint8_t* palette, uint8_t* dst, uint32_t size)
{
  if (!indexed || !palette || !dst || size == 0) return 0;
  const bool keyed = palette_is_keyed(palette);
  for (uint32_t i = 0; i < size; ++i) {
    uint8_t px[4];
    expand_pixel(palette, indexed[i], keyed, px);
    dst[4 * i + 0] = px[2]; dst[4 * i + 1] = px[1]; dst[4 * i + 2] = px[0]; dst[4 * i + 3] = px[3];
  }
  return (int)size;
}

/* Returns bytes written (4-byte header + blocks), 0 on error -- ConverterDxt::convert's contract. */
int tco_encode_tile(const uint8_t* palette, const uint8_t* indexed, int w, int h, int type, int quality, uint8_t* out)
{
  if (!palette || !indexed || !out || w <= 0 || h <= 0) return 0;
  const int space = tco_required_space(type, w, h);
  if (space == 0) return 0;
  const int blockSize = (type == 1) ? 8 : 16;
  const int flags = squish_flags(type, quality);
  out[0] = (uint8_t)(w & 0xff); out[1] = (uint8_t)((w >> 8) & 0xff);
  out[2] = (uint8_t)(h & 0xff); out[3] = (uint8_t)((h >> 8) & 0xff);
  uint8_t* dst = out + 4;
  for (int y = 0; y < h; y += 4)
    for (int x = 0; x < w; x += 4) {
      uint8_t rgba[64];
      gather_block(palette, indexed, w, h, x, y, rgba);
      squish::Compress(rgba, dst, flags);
      dst += blockSize;
    }
  return space + 4;
}

/* N full 64x64 tiles laid out as in a headerless TIS (per tile: 1024 B palette, 4096 B indices;
 * graphics.cpp:101-119) -> N records of (4 + blocks) bytes each.  `threads` static partitions. */
void tco_encode_tiles(const uint8_t* tiles, int n, int type, int quality, uint8_t* out, int threads)
{
  const int record = tco_required_space(type, 64, 64) + 4;
  if (threads < 1) threads = 1;
  if (threads > n) threads = n > 0 ? n : 1;
  auto work = [&](int t) {
    squish_oracle_counters_reset();
    int lo = (int)((long long)n * t / threads), hi = (int)((long long)n * (t + 1) / threads);
    for (int i = lo; i < hi; ++i)
      tco_encode_tile(tiles + (size_t)i * 5120, tiles + (size_t)i * 5120 + 1024, 64, 64, type, quality,
                      out + (size_t)i * record);
    SquishOracleCounters c; squish_oracle_counters_get(&c);
    addTotals(c);
  };
  if (threads == 1) { work(0); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
  for (auto& th : pool) th.join();
}

void tco_totals_reset(void) { std::lock_guard<std::mutex> lock(g_totalsMutex); std::memset(&g_totals, 0, sizeof(g_totals)); }
void tco_totals_get(SquishOracleCounters* out) { std::lock_guard<std::mutex> lock(g_totalsMutex); *out = g_totals; }

/* Decodes one encoded tile record (u16 w, u16 h, blocks) into w*h pixels, memory order B,G,R,A
 * (ColorFormat::ARGB as a little-endian u32).  Returns w*h*4 or 0. */
int tco_decode_tile_ex(const uint8_t* encoded, int type, uint8_t* bgra, int referenceQuirk);
int tco_decode_tile(const uint8_t* encoded, int type, uint8_t* bgra) { return tco_decode_tile_ex(encoded, type, bgra, 0); }

int tco_decode_tile_ex(const uint8_t* encoded, int type, uint8_t* bgra, int referenceQuirk)
{
  if (!encoded || !bgra || type < 1 || type > 3) return 0;
  const int w = encoded[0] | (encoded[1] << 8), h = encoded[2] | (encoded[3] << 8);
  if (w <= 0 || h <= 0) return 0;
  const int blockSize = (type == 1) ? 8 : 16;
  const uint8_t* src = encoded + 4;
  for (int y = 0; y < h; y += 4)
    for (int x = 0; x < w; x += 4) {
      uint8_t px[64];
      decode_block(src, type, px, referenceQuirk != 0);
      src += blockSize;
      for (int py = 0; py < 4 && y + py < h; ++py)
        for (int qx = 0; qx < 4 && x + qx < w; ++qx)
          std::memcpy(bgra + 4 * ((size_t)(y + py) * w + (x + qx)), px + 4 * (4 * py + qx), 4);
    }
  return w * h * 4;
}

}  // extern "C"
