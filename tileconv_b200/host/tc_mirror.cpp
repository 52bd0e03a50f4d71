// Out-of-line pieces of the stand-alone interface mirror (not built when the plugin is compiled
// against the reference's headers).
#ifndef TCB200_WITH_REFERENCE_HEADERS
#include "tc_mirror.h"

#include <cstring>
#include <thread>

#include <zlib.h>

namespace tc {

int Options::getThreads() const noexcept
{
  // 0 = autodetect (options.cpp:345-355, tilethreadpool_posix.cpp:27-30)
  if (m_threads > 0) return m_threads;
  unsigned n = std::thread::hardware_concurrency();
  return (int)(n < 1 ? 1 : (n > 256 ? 256 : n));
}

// Single-tile path, the unit of work the reference's pool workers run (tiledata.cpp:58-66,
// 111-162): factory -> getRequiredSpace -> convert -> copy out (-u semantics; the deflate step of
// the batched path lives in TileBatchQueue::harvest).
TileData& TileData::operator()() noexcept
{
  m_error = true;
  m_size = 0;
  if (!m_encoding) { m_errorMsg = "Decoding is not part of this build\n"; return *this; }
  if (!m_palette || !m_indexed || !m_deflated || m_index < 0 || m_width <= 0 || m_height <= 0) {
    m_errorMsg = "Invalid tile data found\n";
    return *this;
  }
  ConverterPtr converter = ConverterFactory::GetConverter(m_options, Options::GetEncodingCode(m_options.getEncoding(), m_options.isDeflate()));
  if (converter == nullptr) { m_errorMsg = "Unsupported source format found\n"; return *this; }
  converter->setEncoding(true);
  converter->setColorFormat(Converter::ColorFormat::ARGB);
  const int space = converter->getRequiredSpace(m_width, m_height);
  if (space <= 0) { m_errorMsg = "Error while calculating space\n"; return *this; }
  std::unique_ptr<uint8_t[]> encoded(new uint8_t[(size_t)space + HEADER_TILE_ENCODED_SIZE]);
  const int written = converter->convert(m_palette.get(), m_indexed.get(), encoded.get(), m_width, m_height);
  if (written <= 0) { m_errorMsg = "Error while encoding tile data\n"; return *this; }
  if (m_options.isDeflate()) {
    // one-shot zlib stream at level 9 into a buffer of twice the size (tiledata.cpp:139-148, compress.cpp:36)
    uLongf dstLen = (uLongf)written * 2;
    if (compress2(m_deflated.get(), &dstLen, encoded.get(), (uLong)written, 9) != Z_OK) {
      m_errorMsg = "Error while compressing tile data\n";
      return *this;
    }
    m_size = (int)dstLen;
  } else {
    std::memcpy(m_deflated.get(), encoded.get(), (size_t)written);
    m_size = written;
  }
  m_error = false;
  return *this;
}

}  // namespace tc
#endif
