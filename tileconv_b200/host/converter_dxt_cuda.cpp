#include "converter_dxt_cuda.h"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "tcb200.h"

namespace tc {

namespace {

// The reference builds a new converter for every tile (tiledata.cpp:114) and runs convert() on up to 256 pool
// threads; creating a CUDA context per tile -- or per thread -- would dominate.  Encoder contexts therefore live in one
// process-wide pool keyed by (pixel type, encode quality): convert() borrows an idle context and returns it, so there
// are as many contexts as there are tiles in flight at once, and new ones are spread round-robin over the visible
// devices.  The pool is never torn down (the CUDA runtime may already be gone when static destructors run).
class ContextPool
{
public:
  tcb200_ctx* acquire(int type, int quality)
  {
    {
      std::lock_guard<std::mutex> lock(m_mutex);
      std::vector<tcb200_ctx*>& idle = m_idle[std::make_pair(type, quality)];
      if (!idle.empty()) { tcb200_ctx* ctx = idle.back(); idle.pop_back(); return ctx; }
    }
    const int devices = tcb200_device_count();
    const int device = devices > 0 ? (int)(m_next.fetch_add(1) % (unsigned)devices) : 0;
    tcb200_ctx* ctx = tcb200_create(device, type, quality, 16);
    if (ctx == nullptr) {
      // Converter::convert can only return 0 (converter.h:83); say why once, on stderr
      std::call_once(m_reported, [] { std::fprintf(stderr, "tileconv_b200: %s\n", tcb200_last_error(nullptr)); });
    }
    return ctx;
  }
  void release(int type, int quality, tcb200_ctx* ctx)
  {
    std::lock_guard<std::mutex> lock(m_mutex);
    m_idle[std::make_pair(type, quality)].push_back(ctx);
  }
private:
  std::mutex m_mutex;
  std::map<std::pair<int, int>, std::vector<tcb200_ctx*>> m_idle;
  std::atomic<unsigned> m_next{0};
  std::once_flag m_reported;
};

ContextPool& pool()
{
  static ContextPool* p = new ContextPool();
  return *p;
}

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
  const int type = (int)getType(), quality = getOptions().getEncodingQuality();
  tcb200_ctx* ctx = pool().acquire(type, quality);
  if (ctx == nullptr) return 0;

  uint8_t record[TCB200_TILE_RECORD];
  std::memcpy(record, palette, 1024);
  std::memcpy(record + 1024, indexed, (size_t)width * height);
  const uint16_t wh[2] = { (uint16_t)width, (uint16_t)height };
  uint8_t out[4 + 4096];
  const int rc = tcb200_encode(ctx, record, wh, 1, out);
  if (rc != TCB200_OK) {
    static std::once_flag reported;
    std::call_once(reported, [&] { std::fprintf(stderr, "tileconv_b200: GPU encode failed: %s\n", tcb200_last_error(ctx)); });
  }
  pool().release(type, quality, ctx);
  if (rc != TCB200_OK) return 0;
  const int total = getRequiredSpace(width, height) + 4;
  std::memcpy(encoded, out, (size_t)total);
  return total;
}

}  // namespace tc
