#include "tile_stream.h"
#include "slab_pipeline.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <vector>

#include <zlib.h>

namespace tc {

namespace {

uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint32_t rd16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
void wr32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

bool readAll(const std::string& path, std::vector<uint8_t>& data)
{
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) return false;
  const std::streamoff size = f.tellg();
  if (size < 0) return false;
  data.resize((size_t)size);
  f.seekg(0);
  return size == 0 || (bool)f.read(reinterpret_cast<char*>(data.data()), size);
}

struct TileSource
{
  const uint8_t* palette;
  const uint8_t* indices;
  int width, height;
};

// Pushes `count` tiles through a pool created by createThreadPool() and writes
// "u32 size + payload" per tile strictly in index order, exactly the protocol of
// graphics.cpp:94-147: add one, drain whatever is next in order, after the last tile wait.
template <class Provider>
bool pumpTiles(const Options& options, unsigned count, Provider provider, std::FILE* out, std::string& error)
{
  ThreadPoolPtr pool = createThreadPool(options.getThreads(), 64);
  const unsigned code = Options::GetEncodingCode(options.getEncoding(), options.isDeflate());
  unsigned added = 0, written = 0;
  while (added < count || !pool->finished()) {
    if (added < count) {
      const TileSource src = provider(added);
      BytePtr indexed(new uint8_t[MAX_TILE_SIZE_8], std::default_delete<uint8_t[]>());
      BytePtr palette(new uint8_t[PALETTE_SIZE], std::default_delete<uint8_t[]>());
      BytePtr deflated(new uint8_t[MAX_TILE_SIZE_32 * 2], std::default_delete<uint8_t[]>());
      std::memcpy(palette.get(), src.palette, PALETTE_SIZE);
      std::memcpy(indexed.get(), src.indices, (size_t)src.width * src.height);
      TileDataPtr tile(new TileData(options));
      tile->setEncoding(true);
      tile->setIndex((int)added);
      tile->setType(code);
      tile->setPaletteData(palette);
      tile->setIndexedData(indexed);
      tile->setDeflatedData(deflated);
      tile->setWidth(src.width);
      tile->setHeight(src.height);
      pool->addTileData(tile);
      ++added;
    }
    while (pool->hasResult() && pool->peekResult() != nullptr && (unsigned)pool->peekResult()->getIndex() == written) {
      TileDataPtr done = pool->getResult();
      if (done == nullptr || done->isError() || done->getSize() <= 0) {
        error = (done != nullptr && !done->getErrorMsg().empty()) ? done->getErrorMsg() : "Error while encoding tile data";
        return false;
      }
      uint8_t size[4];
      wr32(size, (uint32_t)done->getSize());
      if (std::fwrite(size, 1, 4, out) != 4 ||
          std::fwrite(done->getDeflatedData().get(), 1, (size_t)done->getSize(), out) != (size_t)done->getSize()) {
        error = "Error while writing tile data";
        return false;
      }
      ++written;
    }
    if (added >= count) pool->waitForResult();
  }
  if (written < count) { error = "Missing tiles"; return false; }
  return true;
}

// ---- fast path (slab_pipeline.h): file descriptors instead of per-tile objects ------------------------------
struct Fd
{
  int fd;
  explicit Fd(int f) : fd(f) {}
  ~Fd() { if (fd >= 0) ::close(fd); }
};

bool preadAll(int fd, uint8_t* p, size_t n, unsigned long long off)
{
  while (n > 0) {
    const ssize_t r = ::pread(fd, p, n, (off_t)off);
    if (r <= 0) return false;
    p += r; n -= (size_t)r; off += (unsigned long long)r;
  }
  return true;
}

// A TIS body is count x 5120 bytes, i.e. already an array of tile records.
class TisFeed : public TileFeed
{
public:
  TisFeed(int fd, unsigned long long bodyOffset, unsigned count) : m_fd(fd), m_body(bodyOffset), m_count(count) {}
  unsigned count() const noexcept override { return m_count; }
  bool fill(unsigned first, unsigned n, uint8_t* records, uint16_t* wh) noexcept override
  {
    for (unsigned i = 0; i < 2 * n; ++i) wh[i] = 64;
    return preadAll(m_fd, records, (size_t)n * 5120, m_body + (unsigned long long)first * 5120);
  }
private:
  int m_fd;
  unsigned long long m_body;
  unsigned m_count;
};

// MOS: palettes, offset table and pixel data are separate arrays; edge tiles are narrower than 64
// (tileconv/graphics.cpp:296-297, 322-342).
class MosFeed : public TileFeed
{
public:
  MosFeed(const std::vector<uint8_t>& mos, unsigned width, unsigned height, uint32_t palOfs)
  : m_mos(mos), m_width(width), m_height(height), m_cols((width + 63) >> 6), m_rows((height + 63) >> 6), m_palOfs(palOfs)
  {
    m_tableOfs = (size_t)palOfs + (size_t)count() * PALETTE_SIZE;
    m_dataOfs = m_tableOfs + (size_t)count() * 4;
  }
  unsigned count() const noexcept override { return m_cols * m_rows; }
  void dims(unsigned i, uint16_t& w, uint16_t& h) const noexcept override
  {
    const unsigned row = i / m_cols, col = i % m_cols;
    w = (uint16_t)std::min(64u, m_width - col * 64); h = (uint16_t)std::min(64u, m_height - row * 64);
  }
  bool fill(unsigned first, unsigned n, uint8_t* records, uint16_t* wh) noexcept override
  {
    for (unsigned k = 0; k < n; ++k) {
      const unsigned i = first + k, row = i / m_cols, col = i % m_cols;
      const unsigned w = std::min(64u, m_width - col * 64), h = std::min(64u, m_height - row * 64);
      const size_t ofs = m_dataOfs + rd32(m_mos.data() + m_tableOfs + (size_t)i * 4);
      if (ofs + (size_t)w * h > m_mos.size()) return false;
      uint8_t* rec = records + (size_t)k * 5120;
      std::memcpy(rec, m_mos.data() + m_palOfs + (size_t)i * PALETTE_SIZE, PALETTE_SIZE);
      std::memcpy(rec + PALETTE_SIZE, m_mos.data() + ofs, (size_t)w * h);
      wh[2 * k] = (uint16_t)w; wh[2 * k + 1] = (uint16_t)h;
    }
    return true;
  }
private:
  const std::vector<uint8_t>& m_mos;
  unsigned m_width, m_height, m_cols, m_rows;
  uint32_t m_palOfs;
  size_t m_tableOfs, m_dataOfs;
};

void reportTimes(const PipelineTimes& t)
{
  if (!std::getenv("TCB200_TIMING")) return;
  std::fprintf(stderr, "{\"tiles\": %llu, \"devices\": %d, \"slab_tiles\": %d, \"threads\": %d, \"setup_s\": %.4f, \"total_s\": %.4f, "
               "\"tiles_per_s\": %.1f, \"fill_s\": %.4f, \"submit_s\": %.4f, \"gpu_wait_s\": %.4f, \"frame_s\": %.4f, \"write_s\": %.4f, "
               "\"bytes_in\": %llu, \"bytes_out\": %llu}\n",
               t.tiles, t.devices, t.slabTiles, t.threads, t.setup, t.total, t.total > 0 ? (double)t.tiles / t.total : 0.0,
               t.fill, t.submit, t.gpuWait, t.frame, t.write, t.bytesIn, t.bytesOut);
}

// header + framed tiles through the slab pipeline; removes the partial output on failure (graphics.cpp:77,161)
bool encodeFast(const Options& options, TileFeed& feed, const std::string& outFile, const uint8_t* header, size_t headerSize, std::string& error)
{
  Fd out(::open(outFile.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644));
  if (out.fd < 0) { error = "Error opening output file"; return false; }
  bool ok = ::pwrite(out.fd, header, headerSize, 0) == (ssize_t)headerSize;
  if (!ok) error = "Error while writing header";
  PipelineTimes times;
  if (ok) ok = runSlabPipeline(options, feed, out.fd, headerSize, error, &times);
  if (!ok) { ::unlink(outFile.c_str()); return false; }
  reportTimes(times);
  return true;
}

struct OutputFile
{
  std::string path;
  std::FILE* f;
  bool keep = false;
  explicit OutputFile(const std::string& p) : path(p), f(std::fopen(p.c_str(), "wb")) {}
  ~OutputFile() { if (f) std::fclose(f); if (!keep) std::remove(path.c_str()); }
};

}  // namespace

InputKind sniffInput(const std::string& path, bool assumeTis) noexcept
{
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) return InputKind::UNKNOWN;
  const std::streamoff size = f.tellg();
  char sig[4] = { 0, 0, 0, 0 };
  f.seekg(0);
  f.read(sig, 4);
  if (std::memcmp(sig, "TIS ", 4) == 0) return InputKind::TIS;
  if (std::memcmp(sig, "MOS ", 4) == 0 || std::memcmp(sig, "MOSC", 4) == 0) return InputKind::MOS;
  if (assumeTis && size > 0 && size % 5120 == 0) return InputKind::TIS;      // options.cpp:389-395
  return InputKind::UNKNOWN;
}

bool showInfo(const std::string& fileName, bool assumeTis) noexcept
{
  if (fileName.empty()) return false;
  std::printf("Parsing \"%s\"...\n", fileName.c_str());
  std::FILE* f = std::fopen(fileName.c_str(), "rb");
  if (!f) return false;
  struct Close { std::FILE* f; ~Close() { std::fclose(f); } } closer{ f };
  std::fseek(f, 0, SEEK_END);
  const long fileSize = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  auto read = [&](void* dst, size_t n) { return std::fread(dst, 1, n, f) == n; };
  uint8_t sig[4], ver[4], b[16];
  if (!read(sig, 4)) return false;
  if (std::memcmp(sig, "TIS ", 4) == 0) {
    if (!read(ver, 4)) return false;
    if (std::memcmp(ver, "V1  ", 4) != 0 && std::memcmp(ver, "V2  ", 4) != 0) { std::printf("Invalid TIS version.\n"); return false; }
    if (!read(b, 8)) return false;
    const uint32_t tileNum = rd32(b), tileSize = rd32(b + 4);
    std::printf("File type:       TIS\n");
    std::printf("File subtype:    %s\n", tileSize == 0x1400 ? "Paletted" : tileSize == 0x000c ? "PVRZ-based" : "Unknown");
    std::printf("Number of tiles: %d\n", (int)tileNum);
  } else if (std::memcmp(sig, "MOS ", 4) == 0 || std::memcmp(sig, "MOSC", 4) == 0) {
    std::vector<uint8_t> mos;
    if (!read(ver, 4)) return false;
    if (std::memcmp(sig, "MOSC", 4) == 0) {
      if (fileSize <= 12) { std::printf("Invalid MOSC size\n"); return false; }
      if (std::memcmp(ver, "V1  ", 4) != 0) { std::printf("Invalid MOSC version.\n"); return false; }
      if (!read(b, 4)) return false;
      const uint32_t mosSize = rd32(b);
      if (mosSize < 24) { std::printf("MOS size too small.\n"); return false; }
      std::vector<uint8_t> packed((size_t)fileSize - 12);
      if (!read(packed.data(), packed.size())) { std::printf("Incomplete or corrupted MOSC file.\n"); return false; }
      mos.resize(mosSize);
      uLongf got = mosSize;
      if (uncompress(mos.data(), &got, packed.data(), (uLong)packed.size()) != Z_OK || got != mosSize) {
        std::printf("Error while decompressing MOSC input file.\n");
        return false;
      }
    } else {
      if (fileSize < 24) { std::printf("MOS size too small\n"); return false; }
      mos.resize(24);
      std::fseek(f, 0, SEEK_SET);
      if (!read(mos.data(), 24)) return false;
    }
    // the reference never clears its isMosc flag (tileconv.cpp:281,284), so plain MOS files are reported as
    // "compressed" too; kept, since -I output is compared byte for byte
    std::printf("File type:       MOS (%s)\n", "compressed");
    if (std::memcmp(mos.data() + 4, "V1  ", 4) == 0) std::printf("File subtype:    Paletted (V1)\n");
    else if (std::memcmp(mos.data() + 4, "V2  ", 4) == 0) std::printf("File subtype:    PVRZ-based (V2)\n");
    else std::printf("File version:    Unknown\n");
    std::printf("Width:           %d\n", (int)rd16(mos.data() + 8));
    std::printf("Height:          %d\n", (int)rd16(mos.data() + 10));
    std::printf("Columns:         %d\n", (int)rd16(mos.data() + 12));
    std::printf("Rows:            %d\n", (int)rd16(mos.data() + 14));
  } else if (std::memcmp(sig, "TBC ", 4) == 0) {
    if (!read(ver, 4)) return false;
    if (std::memcmp(ver, "V1.0", 4) != 0) { std::printf("Invalid or unsupported TBC version.\n"); return false; }
    if (!read(b, 8)) return false;
    const uint32_t compType = rd32(b), tileNum = rd32(b + 4);
    std::printf("File type:       TBC\n");
    std::printf("TBC version:     1.0\n");
    std::printf("Compression:     0x%04x - %s\n", compType, Options::GetEncodingName((int)compType));
    std::printf("Number of tiles: %d\n", (int)tileNum);
  } else if (std::memcmp(sig, "MBC ", 4) == 0) {
    if (!read(ver, 4)) return false;
    if (std::memcmp(ver, "V1.0", 4) != 0) { std::printf("Invalid or unsupported MBC version.\n"); return false; }
    if (!read(b, 12)) return false;
    const uint32_t compType = rd32(b), width = rd32(b + 4), height = rd32(b + 8);
    std::printf("File type:       MBC\n");
    std::printf("MBC version:     1.0\n");
    std::printf("Compression:     0x%04x - %s\n", compType, Options::GetEncodingName((int)compType));
    std::printf("Width:           %d\n", (int)width);
    std::printf("Height:          %d\n", (int)height);
    std::printf("Number of tiles: %d\n", (int)(((width + 63) >> 6) * ((height + 63) >> 6)));
  } else if (std::memcmp(sig, "TIZ0", 4) == 0) {
    if (!read(b, 2)) return false;
    std::printf("File type:       TIZ\n");
    std::printf("Format version:  0\n");
    std::printf("Number of tiles: %d\n", (int)((b[0] << 8) | b[1]));           // big endian
  } else if (std::memcmp(sig, "MOZ0", 4) == 0) {
    if (!read(b, 4)) return false;
    const int width = (b[0] << 8) | b[1], height = (b[2] << 8) | b[3];
    std::printf("File type:       MOZ\n");
    std::printf("Format version:  0\n");
    std::printf("Width:           %d\n", width);
    std::printf("Height:          %d\n", height);
    std::printf("Number of tiles: %d\n", ((width + 63) >> 6) * ((height + 63) >> 6));
  } else if (assumeTis && (fileSize % 5120) == 0) {
    std::printf("File type:       TIS (assumed)\n");
    std::printf("File subtype:    Paletted\n");
    std::printf("Number of tiles: %d\n", (int)(fileSize / 5120));
  } else {
    std::printf("File type:       Unknown\n");
    return false;
  }
  return true;
}

bool encodeTisFile(const Options& options, const std::string& inFile, const std::string& outFile, std::string& error) noexcept
{
  if (inFile.empty() || outFile.empty() || inFile == outFile) { error = "Error opening input file"; return false; }
  if (slabPipelineApplies(options)) {
    // same checks as below, on the header alone; the body is streamed straight into the pinned slabs
    Fd in(::open(inFile.c_str(), O_RDONLY));
    struct stat st;
    if (in.fd < 0 || ::fstat(in.fd, &st) != 0) { error = "Error opening input file"; return false; }
    const unsigned long long size = (unsigned long long)st.st_size;
    uint8_t h[24] = { 0 };
    if (size >= 4 && !preadAll(in.fd, h, (size_t)std::min<unsigned long long>(24, size), 0)) { error = "Error opening input file"; return false; }
    unsigned long long first = 0;
    uint32_t count = 0;
    if (size >= 4 && std::memcmp(h, "TIS ", 4) == 0) {
      if (size < 24) { error = "Invalid TIS header"; return false; }
      if (std::memcmp(h + 4, "V1  ", 4) != 0 && std::memcmp(h + 4, "V2  ", 4) != 0) { error = "Invalid TIS version"; return false; }
      count = rd32(h + 8);
      if (count == 0) { error = "No tiles found"; return false; }
      if (rd32(h + 12) != 0x1400) { error = "Invalid tile size"; return false; }
      if (rd32(h + 16) < 0x18) { error = "Invalid header size"; return false; }
      if (rd32(h + 20) != 0x40) { error = "Invalid tile dimensions"; return false; }
      first = 24;
    } else if (options.assumeTis()) {
      if (size == 0 || size % 5120 != 0) { error = "Headerless TIS has wrong file size"; return false; }
      count = (uint32_t)(size / 5120);
    } else {
      error = "Invalid TIS signature";
      return false;
    }
    if (size < first + (unsigned long long)count * 5120) { error = "Incomplete TIS file"; return false; }
    uint8_t header[HEADER_TBC_SIZE];
    std::memcpy(header, "TBC V1.0", 8);
    wr32(header + 8, Options::GetEncodingCode(options.getEncoding(), options.isDeflate()));
    wr32(header + 12, count);
    TisFeed feed(in.fd, first, count);
    return encodeFast(options, feed, outFile, header, sizeof(header), error);
  }
  std::vector<uint8_t> data;
  if (!readAll(inFile, data)) { error = "Error opening input file"; return false; }
  size_t first = 0;
  uint32_t count = 0;
  if (data.size() >= 4 && std::memcmp(data.data(), "TIS ", 4) == 0) {
    // 'TIS ' 'V1  ' u32 count, u32 tile size (0x1400), u32 header size (>= 0x18), u32 dimension (0x40)
    if (data.size() < 24) { error = "Invalid TIS header"; return false; }
    if (std::memcmp(data.data() + 4, "V1  ", 4) != 0 && std::memcmp(data.data() + 4, "V2  ", 4) != 0) { error = "Invalid TIS version"; return false; }
    count = rd32(data.data() + 8);
    if (count == 0) { error = "No tiles found"; return false; }
    if (rd32(data.data() + 12) != 0x1400) { error = "Invalid tile size"; return false; }
    if (rd32(data.data() + 16) < 0x18) { error = "Invalid header size"; return false; }
    if (rd32(data.data() + 20) != 0x40) { error = "Invalid tile dimensions"; return false; }
    first = 24;                                            // the reference keeps reading right after these fields
  } else if (options.assumeTis()) {
    if (data.empty() || data.size() % 5120 != 0) { error = "Headerless TIS has wrong file size"; return false; }
    count = (uint32_t)(data.size() / 5120);
  } else {
    error = "Invalid TIS signature";
    return false;
  }
  if (data.size() < first + (size_t)count * 5120) { error = "Incomplete TIS file"; return false; }

  OutputFile out(outFile);
  if (!out.f) { error = "Error opening output file"; return false; }
  uint8_t header[HEADER_TBC_SIZE];
  std::memcpy(header, "TBC V1.0", 8);
  wr32(header + 8, Options::GetEncodingCode(options.getEncoding(), options.isDeflate()));
  wr32(header + 12, count);
  if (std::fwrite(header, 1, sizeof(header), out.f) != sizeof(header)) { error = "Error while writing header"; return false; }
  const uint8_t* base = data.data() + first;
  auto provider = [&](unsigned i) { TileSource s = { base + (size_t)i * 5120, base + (size_t)i * 5120 + 1024, 64, 64 }; return s; };
  if (!pumpTiles(options, count, provider, out.f, error)) return false;
  out.keep = true;
  return true;
}

bool encodeMosFile(const Options& options, const std::string& inFile, const std::string& outFile, std::string& error) noexcept
{
  std::vector<uint8_t> file, mos;
  if (inFile.empty() || outFile.empty() || inFile == outFile || !readAll(inFile, file)) { error = "Error opening input file"; return false; }
  if (file.size() >= 12 && std::memcmp(file.data(), "MOSC", 4) == 0) {
    // 'MOSC' 'V1  ' u32 uncompressed size, zlib stream (graphics.cpp:861-891)
    if (std::memcmp(file.data() + 4, "V1  ", 4) != 0) { error = "Invalid MOSC version"; return false; }
    const uint32_t size = rd32(file.data() + 8);
    if (size < 24) { error = "MOS size too small"; return false; }
    mos.resize(size);
    uLongf got = size;
    if (uncompress(mos.data(), &got, file.data() + 12, (uLong)(file.size() - 12)) != Z_OK || got != size) {
      error = "Error while decompressing MOSC input file";
      return false;
    }
  } else {
    mos.swap(file);
  }
  // 'MOS ' 'V1  ' u16 width, u16 height, u16 columns, u16 rows, u32 block size (0x40), u32 palette offset
  if (mos.size() < 24 || std::memcmp(mos.data(), "MOS ", 4) != 0) { error = "Invalid MOS signature"; return false; }
  if (std::memcmp(mos.data() + 4, "V1  ", 4) != 0) { error = "Unsupported MOS version"; return false; }
  const unsigned width = rd16(mos.data() + 8), height = rd16(mos.data() + 10);
  if (width == 0 || height == 0) { error = "Invalid MOS dimensions"; return false; }
  if (rd16(mos.data() + 12) == 0 || rd16(mos.data() + 14) == 0) { error = "Invalid number of tiles"; return false; }
  if (rd32(mos.data() + 16) != 0x40) { error = "Invalid tile dimensions"; return false; }
  const uint32_t palOfs = rd32(mos.data() + 20);
  if (palOfs < 24) { error = "MOS header too small"; return false; }
  const unsigned cols = (width + 63) >> 6, rows = (height + 63) >> 6, count = cols * rows;
  const size_t tableOfs = (size_t)palOfs + (size_t)count * PALETTE_SIZE, dataOfs = tableOfs + (size_t)count * 4;
  if (mos.size() < dataOfs + (size_t)width * height) { error = "Incomplete or corrupted MOS file"; return false; }

  uint8_t header[HEADER_MBC_SIZE];
  std::memcpy(header, "MBC V1.0", 8);
  wr32(header + 8, Options::GetEncodingCode(options.getEncoding(), options.isDeflate()));
  wr32(header + 12, width);
  wr32(header + 16, height);
  if (slabPipelineApplies(options)) {
    MosFeed feed(mos, width, height, palOfs);
    if (!encodeFast(options, feed, outFile, header, sizeof(header), error)) {
      if (error == "Error while reading tile data") error = "Tile data outside of the MOS file";
      return false;
    }
    return true;
  }
  OutputFile out(outFile);
  if (!out.f) { error = "Error opening output file"; return false; }
  if (std::fwrite(header, 1, sizeof(header), out.f) != sizeof(header)) { error = "Error while writing header"; return false; }
  bool bad = false;
  auto provider = [&](unsigned i) {
    const unsigned row = i / cols, col = i % cols;
    TileSource s;
    s.width = (int)std::min(64u, width - col * 64);
    s.height = (int)std::min(64u, height - row * 64);
    s.palette = mos.data() + palOfs + (size_t)i * PALETTE_SIZE;
    const size_t ofs = dataOfs + rd32(mos.data() + tableOfs + (size_t)i * 4);
    if (ofs + (size_t)s.width * s.height > mos.size()) { bad = true; s.indices = mos.data() + dataOfs; }
    else s.indices = mos.data() + ofs;
    return s;
  };
  if (!pumpTiles(options, count, provider, out.f, error)) return false;
  if (bad) { error = "Tile data outside of the MOS file"; return false; }
  out.keep = true;
  return true;
}

}  // namespace tc
