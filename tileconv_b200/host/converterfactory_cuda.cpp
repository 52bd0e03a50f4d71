// Replacement for the reference's tileconv/converterfactory.cpp:30-51: pixel types 1,2,3 resolve to
// the CUDA-backed ConverterDxtCuda.  Built INTO tileconv (-DTCB200_WITH_REFERENCE_HEADERS) it keeps
// the reference's converters for the other types; stand-alone only the BCn encoder exists.
#include "converter_dxt_cuda.h"
#ifdef TCB200_WITH_REFERENCE_HEADERS
#include "converter_raw.h"
#include "converter_z.h"
#endif

namespace tc {

ConverterPtr ConverterFactory::GetConverter(const Options& options, unsigned type) noexcept
{
  switch (type & 0xff) {
    case ENCODE_DXT1: case ENCODE_DXT3: case ENCODE_DXT5:
      return ConverterPtr(new ConverterDxtCuda(options, type & 0xff));
#ifdef TCB200_WITH_REFERENCE_HEADERS
    case ENCODE_RAW: return ConverterPtr(new ConverterRaw(options, type & 0xff));
    case ENCODE_Z: return ConverterPtr(new ConverterZ(options, type & 0xff));
#endif
    default: return nullptr;
  }
}

}  // namespace tc
