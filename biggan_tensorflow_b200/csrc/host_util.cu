// Host-side utilities of the C ABI (no device code).  bg_crc32c: CRC-32C (Castagnoli, reflected polynomial 0x82F63B78),
// the checksum TensorFlow's checkpoint format (TensorBundle) stores per tensor and per index block -- used by
// checkpoint.py to read / write checkpoints in the reference's format at memory speed (slicing-by-8).
#include <stddef.h>
#include <stdint.h>
#include "../../include/biggan_b200.h"

static uint32_t crc_table[8][256];
static bool crc_ready = false;

static void crc_init() {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
    crc_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) crc_table[t][i] = (crc_table[t - 1][i] >> 8) ^ crc_table[0][crc_table[t - 1][i] & 0xff];
  crc_ready = true;
}

// crc of `n` bytes continuing from `init` (0 for a fresh checksum); the 32 result bits are returned in an int
extern "C" int bg_crc32c(const void* data, size_t n, int init) {
  if (!crc_ready) crc_init();
  const uint8_t* p = (const uint8_t*)data;
  uint32_t c = ~(uint32_t)init;
  while (n && ((uintptr_t)p & 7)) { c = crc_table[0][(c ^ *p++) & 0xff] ^ (c >> 8); --n; }
  while (n >= 8) {
    uint64_t w = *(const uint64_t*)p ^ c;
    c = crc_table[7][w & 0xff] ^ crc_table[6][(w >> 8) & 0xff] ^ crc_table[5][(w >> 16) & 0xff] ^ crc_table[4][(w >> 24) & 0xff] ^
        crc_table[3][(w >> 32) & 0xff] ^ crc_table[2][(w >> 40) & 0xff] ^ crc_table[1][(w >> 48) & 0xff] ^ crc_table[0][(w >> 56) & 0xff];
    p += 8; n -= 8;
  }
  while (n--) c = crc_table[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return (int)~c;
}
