/* oracle/ref_shim/ref_wrap.cpp -- TEST INFRASTRUCTURE.
 * Plain-C entry points into the REFERENCE's own per-tile encode path, compiled from the sources
 * where they lie under /root/reference/tileconv (never copied), linked with the restated squish in
 * oracle/squish_oracle.cpp.  Output: oracle/_ref/libtileconv_ref.so.  Used to (1) validate
 * oracle/tile_oracle.cpp byte for byte and (2) time the reference's CPU path (`cpu_baseline.kind =
 * "reference"`): every tile goes through tc::TileData::operator() -> TileData::encode
 * (tiledata.cpp:58-66,111-162) -> ConverterFactory::GetConverter (converterfactory.cpp:30-51) ->
 * ConverterDxt::convert (converter_dxt.cpp:65-77), with -u semantics (no zlib), which is what the
 * reference's pool workers run per tile (tilethreadpool_posix.cpp:119-145). */
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include <squish.h>
#include "converter.h"
#include "converterfactory.h"
#include "options.h"
#include "tiledata.h"

extern thread_local std::vector<unsigned char> g_tcref_captured_pixels;   /* ref_stubs.cpp */
extern thread_local int g_tcref_captured_w, g_tcref_captured_h;

namespace {

std::mutex g_totalsMutex;
SquishOracleCounters g_totals;   /* candidates the restated squish evaluated, summed over workers */

void addTotals()
{
  SquishOracleCounters c;
  squish_oracle_counters_get(&c);
  std::lock_guard<std::mutex> lock(g_totalsMutex);
  g_totals.blocks += c.blocks; g_totals.single_blocks += c.single_blocks;
  g_totals.range_blocks += c.range_blocks; g_totals.cluster_blocks += c.cluster_blocks;
  g_totals.cand3 += c.cand3; g_totals.cand4 += c.cand4;
  g_totals.sweeps3 += c.sweeps3; g_totals.sweeps4 += c.sweeps4; g_totals.points += c.points;
}

tc::Encoding encodingOf(int type)
{
  switch (type) {
    case 0: return tc::Encoding::RAW;
    case 1: return tc::Encoding::BC1;
    case 2: return tc::Encoding::BC2;
    case 3: return tc::Encoding::BC3;
    default: return tc::Encoding::UNKNOWN;
  }
}

int encodeOne(const tc::Options& options, int index, const uint8_t* palette, const uint8_t* indexed,
              int w, int h, uint8_t* out, int outCapacity)
{
  tc::BytePtr pal(new uint8_t[1024], std::default_delete<uint8_t[]>());
  tc::BytePtr idx(new uint8_t[(size_t)w * h], std::default_delete<uint8_t[]>());
  tc::BytePtr enc(new uint8_t[64 * 64 * 4 * 2], std::default_delete<uint8_t[]>());
  std::memcpy(pal.get(), palette, 1024);
  std::memcpy(idx.get(), indexed, (size_t)w * h);
  tc::TileData tile(options);
  tile.setEncoding(true);
  tile.setIndex(index);
  tile.setPaletteData(pal);
  tile.setIndexedData(idx);
  tile.setDeflatedData(enc);
  tile.setWidth(w);
  tile.setHeight(h);
  tile();
  if (tile.isError() || tile.getSize() <= 0 || tile.getSize() > outCapacity) return 0;
  std::memcpy(out, enc.get(), (size_t)tile.getSize());
  return tile.getSize();
}

void configure(tc::Options& options, int type, int quality)
{
  options.setEncoding(encodingOf(type));
  options.setDeflate(false);               /* -u */
  options.setEncodingQuality(quality);     /* -q -<quality> */
  options.setVerbosity(0);
}

}  // namespace

extern "C" {

/* Known answer KA1: the channel swizzle the block path depends on.  Returns 1
 * when Converter::ReorderColors was compiled with its intended semantics (see Makefile, F3). */
int tcref_selftest(void)
{
  uint8_t px[4] = { 0x11, 0x22, 0x33, 0x44 };
  tc::Converter::ReorderColors(px, 1, tc::Converter::ColorFormat::ARGB, tc::Converter::ColorFormat::ABGR);
  return px[0] == 0x33 && px[1] == 0x22 && px[2] == 0x11 && px[3] == 0x44;
}

int tcref_encode_tile(const uint8_t* palette, const uint8_t* indexed, int w, int h, int type, int quality,
                      uint8_t* out, int outCapacity)
{
  tc::Options options;
  configure(options, type, quality);
  return encodeOne(options, 0, palette, indexed, w, h, out, outCapacity);
}

/* N full tiles in headerless-TIS layout -> N fixed-size records; static partition over `threads`
 * std::threads (CPU baseline "B" of BASELINE.md: the reference's per-tile work without its 50 ms
 * polling pool).  Returns the number of tiles that failed. */
int tcref_encode_tiles(const uint8_t* tiles, int n, int type, int quality, uint8_t* out, int record, int threads)
{
  tc::Options options;
  configure(options, type, quality);
  if (threads < 1) threads = 1;
  if (threads > n) threads = n > 0 ? n : 1;
  std::vector<int> failed((size_t)threads, 0);
  auto work = [&](int t) {
    squish_oracle_counters_reset();
    int lo = (int)((long long)n * t / threads), hi = (int)((long long)n * (t + 1) / threads);
    for (int i = lo; i < hi; ++i)
      if (encodeOne(options, i, tiles + (size_t)i * 5120, tiles + (size_t)i * 5120 + 1024, 64, 64,
                    out + (size_t)i * record, record) != record)
        ++failed[(size_t)t];
    addTotals();
  };
  if (threads == 1) { work(0); return failed[0]; }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
  for (auto& th : pool) th.join();
  int total = 0;
  for (int f : failed) total += f;
  return total;
}

/* The reference's block decoders on one encoded tile record (u16 w, u16 h, blocks): runs
 * ConverterDxt::convert in the decode direction (converter_dxt.cpp:78-89); the quantiser stub captures
 * the decoded pixels, returned here in memory order R,G,B,A.  Returns w*h*4 or 0. */
int tcref_decode_tile(const uint8_t* encoded, int encodedSize, int type, uint8_t* rgba, int capacity)
{
  tc::Options options;
  options.setVerbosity(0);
  tc::ConverterPtr converter = tc::ConverterFactory::GetConverter(options, (unsigned)type);
  if (converter == nullptr || !converter->canDecode() || encodedSize < 4) return 0;
  converter->setEncoding(false);
  converter->setColorFormat(tc::Converter::ColorFormat::ARGB);
  std::vector<uint8_t> enc(encoded, encoded + encodedSize + 16), palette(1024), indexed(64 * 64);
  g_tcref_captured_pixels.clear();
  converter->convert(palette.data(), indexed.data(), enc.data());      /* fails in the stubbed quantiser, by design */
  const int bytes = g_tcref_captured_w * g_tcref_captured_h * 4;
  if (bytes <= 0 || bytes > capacity) return 0;
  std::memcpy(rgba, g_tcref_captured_pixels.data(), (size_t)bytes);
  return bytes;
}

void tcref_totals_reset(void) { std::lock_guard<std::mutex> lock(g_totalsMutex); std::memset(&g_totals, 0, sizeof(g_totals)); }
void tcref_totals_get(SquishOracleCounters* out) { std::lock_guard<std::mutex> lock(g_totalsMutex); *out = g_totals; }

}  // extern "C"
