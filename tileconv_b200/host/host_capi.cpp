// Plain-C test and tooling entry points into the C++ host layer (used by tests/ through ctypes).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "converter_dxt_cuda.h"
#include "tile_batch_queue.h"
#include "tile_stream.h"

using namespace tc;

namespace {

Encoding encodingOf(int type) { return type == 0 ? Encoding::RAW : type == 1 ? Encoding::BC1 : type == 2 ? Encoding::BC2 : type == 3 ? Encoding::BC3 : Encoding::UNKNOWN; }
thread_local std::string t_error;

}  // namespace

extern "C" {

const char* tcbh_last_error(void) { return t_error.c_str(); }

// One tile through the per-tile plugin surface, the way TileData::encode drives it
// (tiledata.cpp:111-162 with -u): factory -> setEncoding/setColorFormat -> getRequiredSpace -> convert.
int tcbh_convert_tile(int type, int quality, const uint8_t* palette, const uint8_t* indexed, int w, int h, uint8_t* encoded, int capacity)
{
  Options options;
  options.setEncoding(encodingOf(type));
  options.setEncodingQuality(quality);
  options.setDeflate(false);
  ConverterPtr converter = ConverterFactory::GetConverter(options, Options::GetEncodingCode(options.getEncoding(), false));
  if (converter == nullptr || !converter->canEncode()) return 0;
  converter->setEncoding(true);
  converter->setColorFormat(Converter::ColorFormat::ARGB);
  const int space = converter->getRequiredSpace(w, h);
  if (space <= 0 || space + 4 > capacity) return 0;
  uint8_t pal[1024];
  std::memcpy(pal, palette, 1024);            // convert() may reorder the palette in place (converter_dxt.cpp:74)
  return converter->convert(pal, const_cast<uint8_t*>(indexed), encoded, w, h);
}

}  // extern "C"

namespace {

// tile i carries `options` unless altEvery > 0 and i % altEvery == altEvery - 1, which carries `alt`
int queueEncode(const Options& options, const Options& alt, int altEvery, const uint8_t* tiles, const uint16_t* wh, int n,
                uint8_t* out, int outStride, int* sizes, int slabTiles, int gpus, int threads)
{
  TileBatchQueue pool((unsigned)(threads > 0 ? threads : 1), 64, slabTiles > 0 ? slabTiles : 4096, gpus);
  int added = 0, drained = 0, problems = 0;
  typedef std::chrono::steady_clock Clock;
  double tMake = 0, tAdd = 0, tDrain = 0, tWait = 0;
  const bool profile = std::getenv("TCBH_PROFILE") != nullptr;
  while (added < n || !pool.finished()) {
    Clock::time_point t0 = Clock::now();
    if (added < n) {
      const int w = wh ? wh[2 * added] : 64, h = wh ? wh[2 * added + 1] : 64;
      BytePtr palette(new uint8_t[1024], std::default_delete<uint8_t[]>());
      BytePtr indexed(new uint8_t[4096], std::default_delete<uint8_t[]>());
      BytePtr deflated(new uint8_t[4096 * 4 * 2], std::default_delete<uint8_t[]>());
      std::memcpy(palette.get(), tiles + (size_t)added * 5120, 1024);
      std::memcpy(indexed.get(), tiles + (size_t)added * 5120 + 1024, (size_t)(w > 0 && h > 0 && w <= 64 && h <= 64 ? w * h : 0));
      const Options& mine = (altEvery > 0 && added % altEvery == altEvery - 1) ? alt : options;
      TileDataPtr tile(new TileData(mine));
      tile->setEncoding(true);
      tile->setIndex(added);
      tile->setType(Options::GetEncodingCode(mine.getEncoding(), mine.isDeflate()));
      tile->setPaletteData(palette); tile->setIndexedData(indexed); tile->setDeflatedData(deflated);
      tile->setWidth(w); tile->setHeight(h);
      Clock::time_point t1 = Clock::now();
      pool.addTileData(tile);
      ++added;
      Clock::time_point t2 = Clock::now();
      if (profile) { tMake += std::chrono::duration<double>(t1 - t0).count(); tAdd += std::chrono::duration<double>(t2 - t1).count(); }
    }
    Clock::time_point t3 = Clock::now();
    while (pool.hasResult() && pool.peekResult() != nullptr && pool.peekResult()->getIndex() == drained) {
      TileDataPtr done = pool.getResult();
      sizes[drained] = done->getSize();
      if (done->isError() || done->getSize() <= 0 || done->getSize() > outStride) ++problems;
      else std::memcpy(out + (size_t)drained * outStride, done->getDeflatedData().get(), (size_t)done->getSize());
      ++drained;
    }
    Clock::time_point t4 = Clock::now();
    if (added >= n) pool.waitForResult();
    if (profile) { tDrain += std::chrono::duration<double>(t4 - t3).count(); tWait += std::chrono::duration<double>(Clock::now() - t4).count(); }
  }
  if (profile) std::fprintf(stderr, "tcbh_queue_encode n=%d: make %.1f ms, addTileData %.1f ms, drain %.1f ms, wait %.1f ms\n", n, tMake * 1e3, tAdd * 1e3, tDrain * 1e3, tWait * 1e3);
  t_error = pool.lastError();
  return problems + (n - drained);
}

}  // namespace

extern "C" {

// n tiles through createThreadPool()'s queue with the reference's producer/consumer protocol
// (graphics.cpp:97-147).  sizes[i] receives TileData::getSize(); out record i starts at i*outStride.
// Returns 0 when every tile came back, in index order, without error.
int tcbh_queue_encode(int type, int quality, int deflate, const uint8_t* tiles, const uint16_t* wh, int n,
                      uint8_t* out, int outStride, int* sizes, int slabTiles, int gpus, int threads)
{
  Options options;
  options.setEncoding(encodingOf(type));
  options.setEncodingQuality(quality);
  options.setDeflate(deflate != 0);
  return queueEncode(options, options, 0, tiles, wh, n, out, outStride, sizes, slabTiles, gpus, threads);
}

// Same, but every altEvery-th tile asks for (type2, quality2): one pool fed by tiles with different options.
int tcbh_queue_encode_mixed(int type, int quality, int type2, int quality2, int altEvery, const uint8_t* tiles, int n,
                            uint8_t* out, int outStride, int* sizes, int slabTiles, int threads)
{
  Options options, alt;
  options.setEncoding(encodingOf(type)); options.setEncodingQuality(quality); options.setDeflate(false);
  alt.setEncoding(encodingOf(type2)); alt.setEncodingQuality(quality2); alt.setDeflate(false);
  return queueEncode(options, alt, altEvery, tiles, nullptr, n, out, outStride, sizes, slabTiles, 0, threads);
}

// tileconv -t <type> -q -<quality> [-u] [-T] in -> out, encode direction only.
int tcbh_encode_file(const char* inFile, const char* outFile, int type, int quality, int deflate, int assumeTis, int threads)
{
  Options options;
  options.setEncoding(encodingOf(type));
  options.setEncodingQuality(quality);
  options.setDeflate(deflate != 0);
  options.setAssumeTis(assumeTis != 0);
  options.setThreads(threads);
  options.setVerbosity(0);
  std::string error;
  bool ok = false;
  switch (sniffInput(inFile, assumeTis != 0)) {
    case InputKind::TIS: ok = encodeTisFile(options, inFile, outFile, error); break;
    case InputKind::MOS: ok = encodeMosFile(options, inFile, outFile, error); break;
    default: error = "Unsupported input file"; break;
  }
  t_error = error;
  return ok ? 0 : -1;
}

}  // extern "C"
