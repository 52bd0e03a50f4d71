// tileconv_b200/host/converter_dxt_cuda.h -- the B200 stand-in for tc::ConverterDxt's ENCODE half
// (reference tileconv/converter_dxt.h:29-73, converter_dxt.cpp:65-77,109-144,179-225).
//
// Same plugin surface: created by ConverterFactory::GetConverter(options, type) for types 1,2,3,
// configured with setEncoding(true)/setColorFormat(ARGB) by TileData::encode (tiledata.cpp:114-121),
// then convert(palette, indexed, encoded, w, h) returns the total bytes written (4-byte tile header
// included) or 0 on error; never throws.  The arithmetic runs on the GPU through libtcb200.so
// (include/tcb200.h); there is no CPU fallback -- without a device convert() returns 0.
//
// convert() handles ONE tile, which is the reference's calling convention but a poor fit for a GPU
// (one launch + two copies per tile); it exists so that every existing caller keeps working.  The
// measured path is TileBatchQueue (tile_batch_queue.h), which feeds whole slabs of tiles.
#ifndef TCB200_CONVERTER_DXT_CUDA_H
#define TCB200_CONVERTER_DXT_CUDA_H

#include "tc_iface.h"

namespace tc {

#ifdef TCB200_WITH_REFERENCE_HEADERS
typedef ConverterDxt ConverterDxtCudaBase;     // decode keeps the reference's implementation
#else
typedef Converter ConverterDxtCudaBase;
#endif

class ConverterDxtCuda : public ConverterDxtCudaBase
{
public:
  ConverterDxtCuda(const Options& options, unsigned type) noexcept;
  ~ConverterDxtCuda() noexcept;

  bool canEncode() const noexcept { return true; }
#ifdef TCB200_WITH_REFERENCE_HEADERS
  bool canDecode() const noexcept { return true; }
#else
  bool canDecode() const noexcept { return false; }
#endif
  bool deflateAllowed() const noexcept { return true; }

  int getRequiredSpace(int width, int height) const noexcept;
  int getPaddedValue(int v) const noexcept;
  int convert(uint8_t* palette, uint8_t* indexed, uint8_t* encoded, int width, int height) noexcept;

protected:
  bool isTypeValid() const noexcept;
};

}  // namespace tc

#endif
