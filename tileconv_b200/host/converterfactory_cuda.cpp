// Replacement for the reference's tileconv/converterfactory.cpp:30-51: pixel types 1,2,3 resolve to
// the CUDA-backed ConverterDxtCuda.  Built INTO tileconv (-DTCB200_WITH_REFERENCE_HEADERS) it keeps
// the reference's converters for the other types; stand-alone only the BCn encoder exists.
#include "converter_dxt_cuda.h"
#ifdef TCB200_WITH_REFERENCE_HEADERS
#include "converter_raw.h"
#include "converter_z.h"
#endif

#include <cstring>

namespace tc {

#ifndef TCB200_WITH_REFERENCE_HEADERS
namespace {
// Type 0, "no pixel encoding" (FORMATS:41): u16 w, u16 h, the 1024-byte palette, then w*h indices --
// what the reference's ConverterRaw::convert writes in the encode direction
// (tileconv/converter_raw.cpp:48-62; its palette reorder ARGB->ARGB is the identity).
class ConverterRawHost : public Converter
{
public:
  ConverterRawHost(const Options& options, unsigned type) noexcept : Converter(options, type) {}
  bool canEncode() const noexcept { return true; }
  bool canDecode() const noexcept { return false; }
  bool deflateAllowed() const noexcept { return true; }
  int getRequiredSpace(int width, int height) const noexcept { return (width > 0 && height > 0) ? 1024 + width * height : 0; }
  int convert(uint8_t* palette, uint8_t* indexed, uint8_t* encoded, int width, int height) noexcept
  {
    if (!palette || !indexed || !encoded || !isEncoding() || width <= 0 || height <= 0 || width > 0xffff || height > 0xffff) return 0;
    setWidth(width); setHeight(height);
    encoded[0] = (uint8_t)width; encoded[1] = (uint8_t)(width >> 8);
    encoded[2] = (uint8_t)height; encoded[3] = (uint8_t)(height >> 8);
    std::memcpy(encoded + 4, palette, 1024);
    std::memcpy(encoded + 4 + 1024, indexed, (size_t)width * height);
    return getRequiredSpace(width, height) + (int)HEADER_TILE_ENCODED_SIZE;
  }
protected:
  bool isTypeValid() const noexcept { return getType() == 0; }
};
}  // namespace
#endif

ConverterPtr ConverterFactory::GetConverter(const Options& options, unsigned type) noexcept
{
  switch (type & 0xff) {
    case ENCODE_DXT1: case ENCODE_DXT3: case ENCODE_DXT5:
      return ConverterPtr(new ConverterDxtCuda(options, type & 0xff));
#ifdef TCB200_WITH_REFERENCE_HEADERS
    case ENCODE_RAW: return ConverterPtr(new ConverterRaw(options, type & 0xff));
    case ENCODE_Z: return ConverterPtr(new ConverterZ(options, type & 0xff));
#else
    case ENCODE_RAW: return ConverterPtr(new ConverterRawHost(options, 0));
#endif
    default: return nullptr;
  }
}

}  // namespace tc
