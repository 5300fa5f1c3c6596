// ORACLE — test infrastructure only. extern "C" doors onto the reference's own
// PixelFormatConversion functions (N/Encoder/PixelFormatConversion.cpp:16-121), compiled from
// /root/reference in place by oracle/Makefile (target `ref`), so tests can call them via ctypes.
#include <cstddef>
#include <cstdint>
#include "Common.h"
#include "PixelFormatConversion.h"
extern "C" {
void ref_BgraToGray(const BitmapData* b, uint8_t* d) { PixelFormatConversion::BgraToGray(b, d); }
void ref_BgraToGrayAlpha(const BitmapData* b, uint8_t* d) { PixelFormatConversion::BgraToGrayAlpha(b, d); }
void ref_BgraToRgb(const BitmapData* b, uint8_t* d) { PixelFormatConversion::BgraToRgb(b, d); }
void ref_BgraToRgba(const BitmapData* b, uint8_t* d) { PixelFormatConversion::BgraToRgba(b, d); }
}
