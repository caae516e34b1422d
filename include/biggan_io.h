/* libbiggan_io.so -- host-side input decoding for the training step's feeder.
 *
 * Replaces, for the reference's input pipeline (utils.py:12-38 ImageData.image_processing, BigGAN.py:771-787), the
 * TensorFlow ops tf.read_file + tf.image.decode_png(channels=c): PNG bytes -> uint8 HWC pixels.  Pure host code
 * (C ABI, zlib for inflate and CRC-32), thread safe, no global state except the thread-local error text.  The
 * resize / flip / normalise half of image_processing runs on the device (bg_image_prepare in biggan_b200.h).
 * Every function returns 0 on success, < 0 on error; bgio_last_error() gives the text.
 */
#ifndef BIGGAN_IO_H
#define BIGGAN_IO_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* bgio_last_error(void);
/* width, height, channels of the stored image (1 gray, 2 gray+alpha, 3 RGB or palette, 4 RGBA) and its bit depth */
int bgio_png_info(const void* data, size_t n, int* width, int* height, int* channels, int* bit_depth);
/* decode into out[height][width][out_channels] (uint8).  out_channels as tf.image.decode_png(channels=...):
 * 0 = as stored, 1 / 3 / 4 = convert (gray -> RGB by replication, alpha dropped or added as 255; RGB -> gray is refused).
 * 8-bit images (palette images of depth 1-8), non-interlaced. */
int bgio_png_decode(const void* data, size_t n, void* out, size_t out_bytes, int out_channels);

#ifdef __cplusplus
}
#endif
#endif /* BIGGAN_IO_H */
