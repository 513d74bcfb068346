// gpusim_tables.h - host-side builders for the constant tables of the path.
#ifndef GPUSIM_TABLES_H
#define GPUSIM_TABLES_H
#include <cstdint>

namespace gpusim {

// FNV-1a over the first quarter wave of the carrier table (128 values)
constexpr uint32_t kCarrierLutFnv1a = 0x0b87c727u;
// chips 0..1022 followed by chips 0..63 again (a 32-chip window may straddle the 1023-chip wrap),
// plus one spare word so that a window can always read word+1: 35 words
constexpr int kCaWordsPerPrn = 35;

void carrier_lut(int32_t *sin512, int32_t *cos512);
int ca_code(int prn, uint8_t *chips1023);
void ca_words(int prn, uint32_t *words35);

} // namespace gpusim
#endif
