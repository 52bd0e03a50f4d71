#include "converter_dxt_cuda.h"

#include <cstring>
#include <map>
#include <mutex>
#include <utility>

#include "tcb200.h"

namespace tc {

namespace {

// The reference builds a new converter for every tile (tiledata.cpp:114); creating a CUDA context
// per tile would dominate, so encoder contexts are cached per calling thread, keyed by
// (pixel type, encode quality).  Worker threads of the reference's pool each get their own.
struct ContextCache
{
  std::map<std::pair<int, int>, tcb200_ctx*> entries;
  ~ContextCache() { for (auto& e : entries) tcb200_destroy(e.second); }
  tcb200_ctx* get(int type, int quality)
  {
    auto key = std::make_pair(type, quality);
    auto it = entries.find(key);
    if (it != entries.end()) return it->second;
    tcb200_ctx* ctx = tcb200_create(0, type, quality, 16);
    if (ctx) entries[key] = ctx;
    return ctx;
  }
};

thread_local ContextCache t_cache;

}  // namespace

ConverterDxtCuda::ConverterDxtCuda(const Options& options, unsigned type) noexcept
: ConverterDxtCudaBase(options, type)
{
}

ConverterDxtCuda::~ConverterDxtCuda() noexcept {}

int ConverterDxtCuda::getPaddedValue(int v) const noexcept { return (v + 3) & ~3; }

int ConverterDxtCuda::getRequiredSpace(int width, int height) const noexcept
{
  // blocks x 8 (type 1) or 16 (types 2,3); 0 on error (converter_dxt.cpp:41-56)
  if (width <= 0 || height <= 0) return 0;
  const int blocks = (getPaddedValue(width) * getPaddedValue(height)) >> 4;
  switch (getType()) {
    case 1: return blocks << 3;
    case 2: case 3: return blocks << 4;
    default: return 0;
  }
}

bool ConverterDxtCuda::isTypeValid() const noexcept
{
  const unsigned t = getType();
  return t == 1 || t == 2 || t == 3;
}

int ConverterDxtCuda::convert(uint8_t* palette, uint8_t* indexed, uint8_t* encoded, int width, int height) noexcept
{
  if (palette == nullptr || indexed == nullptr || encoded == nullptr) return 0;
  if (!isEncoding()) {
#ifdef TCB200_WITH_REFERENCE_HEADERS
    return ConverterDxtCudaBase::convert(palette, indexed, encoded, width, height);   // decode: reference code
#else
    return 0;
#endif
  }
  if (width <= 0 || height <= 0 || width > 64 || height > 64 || !isTypeValid()) return 0;
  setWidth(width); setHeight(height);
  tcb200_ctx* ctx = t_cache.get((int)getType(), getOptions().getEncodingQuality());
  if (ctx == nullptr) return 0;

  uint8_t record[TCB200_TILE_RECORD];
  std::memcpy(record, palette, 1024);
  std::memcpy(record + 1024, indexed, (size_t)width * height);
  const uint16_t wh[2] = { (uint16_t)width, (uint16_t)height };
  uint8_t out[4 + 4096];
  if (tcb200_encode(ctx, record, wh, 1, out) != TCB200_OK) return 0;
  const int total = getRequiredSpace(width, height) + 4;
  std::memcpy(encoded, out, (size_t)total);
  return total;
}

}  // namespace tc
