/* oracle/ref_shim/ref_stubs.cpp -- TEST INFRASTRUCTURE.
 * Link-time stand-ins for the two reference classes whose own translation units need libraries
 * this image does not have (tileconv/colorquant.cpp -> libimagequant, tileconv/converter_z.cpp and
 * tileconv/jpeg.cpp -> jpeglib.h).  Both are DECODE-only (converter_z.h:35 canEncode()=false;
 * colors.cpp:61-99), i.e. never reached by the encode path the oracle build exists to run.
 * Every member reports failure. */
#include <cstring>
#include <vector>

#include "colorquant.h"
#include "converter_z.h"

/* The decode direction ends in ColorQuant (libimagequant, absent).  The stub refuses to quantise, but
 * it keeps a copy of the 32-bit pixels it was handed, which are exactly the output of the reference's
 * BC1/BC2/BC3 block DECODERS (converter_dxt.cpp:147-176,251-430) after Colors::ARGBToPal's reorder to
 * R,G,B,A (colors.cpp:68).  ref_wrap.cpp's tcref_decode_tile() reads it back, which pins the oracle's
 * restated decoder (the decoder behind the RMSE metric) to the reference's own code. */
thread_local std::vector<unsigned char> g_tcref_captured_pixels;
thread_local int g_tcref_captured_w = 0, g_tcref_captured_h = 0;

namespace tc {

ColorQuant::ColorQuant() noexcept
: m_dithering(false), m_lastTransparent(false), m_maxColors(256), m_qualityMin(0), m_qualityMax(100),
  m_speed(3), m_minOpacity(255), m_posterization(0), m_source(nullptr), m_width(0), m_height(0),
  m_gamma(0.0), m_target(nullptr), m_targetSize(0), m_palette(nullptr), m_paletteSize(0),
  m_liqAttr(nullptr), m_liqImage(nullptr), m_liqResult(nullptr) {}
ColorQuant::~ColorQuant() noexcept {}
bool ColorQuant::setSource(void* bitmap, int width, int height, double) noexcept
{
  g_tcref_captured_pixels.clear();
  g_tcref_captured_w = g_tcref_captured_h = 0;
  if (bitmap != nullptr && width > 0 && height > 0) {
    g_tcref_captured_pixels.assign((unsigned char*)bitmap, (unsigned char*)bitmap + (size_t)width * height * 4);
    g_tcref_captured_w = width; g_tcref_captured_h = height;
  }
  return false;
}
bool ColorQuant::setTarget(void*, int) noexcept { return false; }
bool ColorQuant::setPalette(void*, int) noexcept { return false; }
bool ColorQuant::quantize() noexcept { return false; }
void ColorQuant::setSpeed(int) noexcept {}
double ColorQuant::getQuantizationError() noexcept { return -1.0; }

ConverterZ::ConverterZ(const Options& options, unsigned type) noexcept : Converter(options, type) {}
ConverterZ::~ConverterZ() noexcept {}
int ConverterZ::getRequiredSpace(int, int) const noexcept { return 0; }
int ConverterZ::convert(uint8_t*, uint8_t*, uint8_t*, int, int) noexcept { return 0; }
bool ConverterZ::isTypeValid() const noexcept { return false; }

}  // namespace tc
