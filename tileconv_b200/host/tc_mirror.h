// tileconv_b200/host/tc_mirror.h -- stand-alone mirror of the slice of the reference's interface that
// the encode hot path touches, so that the plugin (converter_dxt_cuda.*) and the batch queue
// (tile_batch_queue.*) build and test without /root/reference.  Same names, argument meaning and
// error behaviour as the reference (file:line given per item); nothing here is executed when the
// plugin is compiled against the reference's own headers (-DTCB200_WITH_REFERENCE_HEADERS, see
// INTEGRATION.md), where these declarations come from tileconv/*.h instead.
#ifndef TCB200_TC_MIRROR_H
#define TCB200_TC_MIRROR_H

#include <cstdint>
#include <memory>
#include <queue>
#include <string>
#include <vector>

namespace tc {

// tileconv/types.h:31-35,55-70
constexpr unsigned ENCODE_RAW = 0, ENCODE_DXT1 = 1, ENCODE_DXT3 = 2, ENCODE_DXT5 = 3, ENCODE_Z = 255;
enum class Encoding { UNKNOWN, RAW, BC1, BC2, BC3, Z };
typedef std::shared_ptr<uint8_t> BytePtr;
constexpr unsigned HEADER_TBC_SIZE = 16, HEADER_MBC_SIZE = 20;
constexpr unsigned HEADER_TILE_ENCODED_SIZE = 4, HEADER_TILE_COMPRESSED_SIZE = 4;
constexpr unsigned PALETTE_SIZE = 1024, TILE_DIMENSION = 64;
constexpr unsigned MAX_TILE_SIZE_8 = TILE_DIMENSION * TILE_DIMENSION, MAX_TILE_SIZE_32 = MAX_TILE_SIZE_8 * 4;

// The subset of tileconv/options.h the path reads: -t (encoding), -q second digit (encode
// quality), -u (deflate off), -j (threads), -T (assume TIS), verbosity.  Defaults: options.cpp:33-45.
class Options
{
public:
  static Encoding GetEncodingType(int code) noexcept          // options.cpp:451-463
  {
    switch (code & 0xff) {
      case ENCODE_RAW: return Encoding::RAW;
      case ENCODE_DXT1: return Encoding::BC1;
      case ENCODE_DXT3: return Encoding::BC2;
      case ENCODE_DXT5: return Encoding::BC3;
      case ENCODE_Z: return Encoding::Z;
      default: return Encoding::UNKNOWN;
    }
  }
  static bool IsTileDeflated(int code) noexcept { return (code & 0x100) == 0; }   // options.cpp:465-468
  static unsigned GetEncodingCode(Encoding type, bool deflate) noexcept          // options.cpp:470-493
  {
    unsigned code = deflate ? 0u : 0x100u;
    switch (type) {
      case Encoding::RAW: return code | ENCODE_RAW;
      case Encoding::BC1: return code | ENCODE_DXT1;
      case Encoding::BC2: return code | ENCODE_DXT3;
      case Encoding::BC3: return code | ENCODE_DXT5;
      case Encoding::Z: return code | ENCODE_Z;
      default: return (unsigned)-1;
    }
  }

  static const char* GetEncodingName(int code) noexcept                          // options.cpp:495-522
  {
    switch (code) {
      case ENCODE_RAW: return "Not encoded (zlib-compressed)";
      case ENCODE_DXT1: return "BC1/DXT1 (zlib-compressed)";
      case ENCODE_DXT3: return "BC2/DXT3 (zlib-compressed)";
      case ENCODE_DXT5: return "BC3/DXT5 (zlib-compressed)";
      case ENCODE_Z: case ENCODE_Z | 256: return "JPEG compressed TIZ/MOZ";
      case ENCODE_RAW | 256: return "Not encoded (uncompressed)";
      case ENCODE_DXT1 | 256: return "BC1/DXT1 (uncompressed)";
      case ENCODE_DXT3 | 256: return "BC2/DXT3 (uncompressed)";
      case ENCODE_DXT5 | 256: return "BC3/DXT5 (uncompressed)";
      default: return "Unknown (unknown)";
    }
  }

  void setEncoding(Encoding type) noexcept { m_encoding = type; }
  Encoding getEncoding() const noexcept { return m_encoding; }
  void setEncodingQuality(int v) noexcept { m_qualityEncoding = v < 0 ? 0 : (v > 9 ? 9 : v); }
  int getEncodingQuality() const noexcept { return m_qualityEncoding; }
  void setDeflate(bool b) noexcept { m_deflate = b; }
  bool isDeflate() const noexcept { return m_deflate; }
  void setThreads(int v) noexcept { m_threads = v < 0 ? 0 : (v > 256 ? 256 : v); }
  int getThreads() const noexcept;                             // 0 = autodetect, options.cpp:345-355
  void setAssumeTis(bool b) noexcept { m_assumeTis = b; }
  bool assumeTis() const noexcept { return m_assumeTis; }
  void setVerbosity(int level) noexcept { m_verbosity = level < 0 ? 0 : (level > 2 ? 2 : level); }
  int getVerbosity() const noexcept { return m_verbosity; }
  bool isSilent() const noexcept { return m_verbosity < 1; }
  bool isVerbose() const noexcept { return m_verbosity > 1; }

private:
  bool m_deflate = true, m_assumeTis = false;
  int m_verbosity = 1, m_qualityEncoding = 9, m_threads = 0;
  Encoding m_encoding = Encoding::BC1;
};

// tileconv/converter.h:30-145 (the pure-virtual plugin surface; only what the encode path needs)
class Converter
{
public:
  enum class ColorFormat { ARGB, ABGR, BGRA, RGBA };
  virtual ~Converter() noexcept {}
  virtual bool canEncode() const noexcept = 0;
  virtual bool canDecode() const noexcept = 0;
  virtual bool deflateAllowed() const noexcept = 0;
  const Options& getOptions() const noexcept { return m_options; }
  void setEncoding(bool b) noexcept { m_encoding = b; }
  bool isEncoding() const noexcept { return m_encoding; }
  void setType(unsigned type) noexcept { m_type = type & 0xff; }
  unsigned getType() const noexcept { return (unsigned)m_type; }
  void setColorFormat(ColorFormat fmt) noexcept { m_colorFormat = fmt; }
  ColorFormat getColorFormat() const noexcept { return m_colorFormat; }
  int getWidth() const noexcept { return m_width; }
  int getHeight() const noexcept { return m_height; }
  virtual int getRequiredSpace(int width, int height) const noexcept = 0;
  virtual int getPaddedValue(int v) const noexcept { return v; }
  // returns the total size of the resulting data, 0 on error (converter.h:74-85)
  virtual int convert(uint8_t* palette, uint8_t* indexed, uint8_t* encoded, int width, int height) noexcept = 0;

protected:
  Converter(const Options& options, unsigned type) noexcept
  : m_options(options), m_encoding(true), m_colorFormat(ColorFormat::ARGB), m_type((int)type), m_width(0), m_height(0) {}
  virtual bool isTypeValid() const noexcept = 0;
  void setWidth(int w) noexcept { m_width = w < 0 ? 0 : w; }
  void setHeight(int h) noexcept { m_height = h < 0 ? 0 : h; }

private:
  const Options& m_options;
  bool m_encoding;
  ColorFormat m_colorFormat;
  int m_type, m_width, m_height;
};
typedef std::shared_ptr<Converter> ConverterPtr;

// tileconv/converterfactory.h:37 -- defined in converterfactory_cuda.cpp
class ConverterFactory
{
public:
  static ConverterPtr GetConverter(const Options& options, unsigned type) noexcept;
};

// tileconv/tiledata.h:30-116 -- one tile travelling through the pool
class TileData
{
public:
  explicit TileData(const Options& options) noexcept : m_options(options) {}
  TileData& operator()() noexcept;                     // tiledata.cpp:58-66 (single-tile path)
  const Options& getOptions() const noexcept { return m_options; }
  void setEncoding(bool b) noexcept { m_encoding = b; }
  bool isEncoding() const noexcept { return m_encoding; }
  void setIndex(int index) noexcept { m_index = index < 0 ? 0 : index; }
  int getIndex() const noexcept { return m_index; }
  void setType(unsigned type) noexcept { m_type = (int)type; }
  unsigned getType() const noexcept { return (unsigned)m_type; }
  void setPaletteData(BytePtr p) noexcept { m_palette = p; }
  BytePtr getPaletteData() const noexcept { return m_palette; }
  void setIndexedData(BytePtr p) noexcept { m_indexed = p; }
  BytePtr getIndexedData() const noexcept { return m_indexed; }
  void setDeflatedData(BytePtr p) noexcept { m_deflated = p; }
  BytePtr getDeflatedData() const noexcept { return m_deflated; }
  void setWidth(int w) noexcept { m_width = w < 0 ? 0 : w; }
  int getWidth() const noexcept { return m_width; }
  void setHeight(int h) noexcept { m_height = h < 0 ? 0 : h; }
  int getHeight() const noexcept { return m_height; }
  void setSize(int size) noexcept { m_size = size < 0 ? 0 : size; }
  int getSize() const noexcept { return m_size; }
  bool isError() const noexcept { return m_error; }
  const std::string& getErrorMsg() const noexcept { return m_errorMsg; }

private:
  const Options& m_options;
  bool m_encoding = false, m_error = false;
  BytePtr m_palette, m_indexed, m_deflated;
  int m_index = -1, m_width = 0, m_height = 0, m_type = 0, m_size = 0;
  std::string m_errorMsg;
};
typedef std::shared_ptr<TileData> TileDataPtr;

// tileconv/tilethreadpool_base.h:47-115 -- the fan-out / ordered fan-in seam.  As in the reference
// the base owns the input queue and the result min-heap (keyed by tile index, tiledata.h:119-129);
// hasResult()/canAddTileData() are NON-virtual and read them directly, so an implementation has to
// publish its results through getResultQueue().
class TileThreadPool;
typedef std::shared_ptr<TileThreadPool> ThreadPoolPtr;
unsigned getThreadPoolAutoThreads();                           // tilethreadpool_base.h:38
ThreadPoolPtr createThreadPool(int threadNum, int tileNum);    // tilethreadpool_base.h:44

struct TileIndexGreater
{
  bool operator()(const TileDataPtr& a, const TileDataPtr& b) const noexcept
  {
    return (a != nullptr && b != nullptr) ? a->getIndex() > b->getIndex() : false;
  }
};

class TileThreadPool
{
public:
  virtual ~TileThreadPool() noexcept {}
  unsigned getMaxTiles() const noexcept { return m_maxTiles; }
  void setMaxTiles(unsigned maxTiles) noexcept { m_maxTiles = maxTiles < 1 ? 1 : maxTiles; }
  virtual void addTileData(TileDataPtr tileData) noexcept = 0;   // blocks while the input side is full
  bool canAddTileData() const noexcept { return m_tiles.size() < m_maxTiles; }
  bool hasResult() const noexcept { return !m_results.empty(); }
  virtual TileDataPtr getResult() noexcept = 0;                  // next result or nullptr
  virtual const TileDataPtr peekResult() noexcept = 0;
  virtual void waitForResult() noexcept = 0;                     // until a result is ready or the pool is idle
  virtual bool finished() noexcept = 0;                          // input empty, nothing active, results drained

protected:
  typedef std::queue<TileDataPtr> TileQueue;
  typedef std::priority_queue<TileDataPtr, std::vector<TileDataPtr>, TileIndexGreater> ResultQueue;
  explicit TileThreadPool(unsigned tileNum) noexcept : m_maxTiles(tileNum < 1 ? 1 : tileNum) {}
  virtual void threadActivated() noexcept = 0;
  virtual void threadDeactivated() noexcept = 0;
  virtual int getActiveThreads() noexcept = 0;
  TileQueue& getTileQueue() noexcept { return m_tiles; }
  ResultQueue& getResultQueue() noexcept { return m_results; }

private:
  unsigned m_maxTiles;
  TileQueue m_tiles;
  ResultQueue m_results;
};

}  // namespace tc

#endif
