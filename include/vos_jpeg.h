/* vos_jpeg.h -- C ABI of the JPEG front end of libvosprop.so.
 *
 * Replaces, for the frames the reference's loader decodes with Pillow (`Image.open(BytesIO(bytes)).convert('RGB')`,
 * /root/reference/src/utils/datasets.py:141-143, i.e. libjpeg-turbo's default decompression), the decode step by
 *   host   : marker parsing + Huffman decoding into quantised DCT coefficients (any thread; no CUDA call), and
 *   device : de-quantisation + integer inverse DCT + chroma up-sampling + YCbCr->RGB on the GPU,
 * with pixels BIT-IDENTICAL to Pillow's for the supported flavour: baseline / extended-sequential Huffman JPEG, 8 bit,
 * one interleaved scan, 1 or 3 components (YCbCr), chroma sampled 1x1, 2x1 or 2x2 relative to luma.  Anything else returns
 * VOSJPEG_ERR_UNSUPPORTED and the caller keeps Pillow for that file (the mirror's InferenceDataset does).
 *
 * Plain pointers and sizes only; no torch types.  Errors: negative codes, message through vosjpeg_last_error(). */
#ifndef VOS_JPEG_H_
#define VOS_JPEG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VOSJPEG_OK 0
#define VOSJPEG_ERR_INVALID (-1)      /* null pointer, truncated or corrupt stream */
#define VOSJPEG_ERR_UNSUPPORTED (-2)  /* a JPEG flavour outside the bit-exact path: decode it with Pillow */
#define VOSJPEG_ERR_CUDA (-3)

typedef struct vosjpeg_info {
    int32_t width, height;         /* image size in pixels */
    int32_t n_comp;                /* 1 (grey: R = G = B = Y) or 3 (YCbCr) */
    int32_t h_samp[3], v_samp[3];  /* sampling factors per component */
    int32_t blocks_w[3], blocks_h[3]; /* 8x8 blocks per row / column of each component's plane, padded to whole MCUs */
    int64_t coef_offset[3];        /* first int16 of each component inside the coefficient buffer */
    int64_t coef_count;            /* int16 values in the coefficient buffer = 64 * sum(blocks_w * blocks_h) */
    int64_t scan_offset;           /* byte offset of the entropy-coded data */
    int32_t restart_interval;      /* MCUs between RSTn markers, 0 = none */
    int32_t dc_table[3], ac_table[3];
    uint16_t quant[3][64];         /* quantisation table of each component, natural (row-major) order */
} vosjpeg_info;

/* Markers up to the start of scan (host).  Replaces the header part of Image.open (datasets.py:141). */
int vosjpeg_parse(const uint8_t* data, int64_t size, vosjpeg_info* info);

/* Huffman decoding of the whole scan (host, thread-safe, no CUDA): coef[coef_offset[c] + (by * blocks_w[c] + bx) * 64 + i] =
 * quantised coefficient i (natural order) of block (bx, by) of component c.  The buffer (info->coef_count int16, e.g. pinned
 * host memory) is overwritten completely.  Replaces the entropy-decoding half of Image.convert('RGB') (datasets.py:142). */
int vosjpeg_entropy_decode(const uint8_t* data, int64_t size, const vosjpeg_info* info, int16_t* coef);

/* Host stage for a batch of files on `n_threads` threads of the library's own (what a DataLoader's worker processes do for the
 * reference, datasets.py:141-143, without the inter-process copies): file i -> items[i] = [vosjpeg_info, zero-padded to
 * header_values int16][coefficients], status[i] = VOSJPEG_OK or the error (VOSJPEG_ERR_UNSUPPORTED: decode that file with Pillow;
 * also returned when header_values + coef_count exceeds item_capacity values). */
int vosjpeg_decode_files_host(const uint8_t* const* datas, const int64_t* sizes, int32_t n_files, int16_t* const* items, int64_t item_capacity,
                              int32_t header_values, int32_t n_threads, int32_t* status);

/* Scratch the device stage needs for one frame (the three sample planes). */
int64_t vosjpeg_scratch_bytes(const vosjpeg_info* info);

/* Device stage for one frame, ordered on `stream`, never synchronises: coefficients (device) -> rgb (device, height x width x 3
 * uint8, the array np.asarray(img.convert('RGB')) holds).  Replaces the reconstruction half of Image.convert('RGB'). */
int vosjpeg_reconstruct(const vosjpeg_info* info, const int16_t* coef_dev, uint8_t* scratch_dev, uint8_t* rgb_dev, void* stream);

/* The device stage for a batch of frames of ONE geometry (what a video's frames are) in two launches: frame f takes its
 * coefficients at coef_dev + f * coef_stride (int16 elements) and its quantisation tables ([3][64] uint16, natural order) at
 * quant_dev + f * quant_stride (device memory, e.g. the `quant` field of each frame's vosjpeg_info copied along with its
 * coefficients), or info->quant for every frame when quant_dev is NULL.  scratch_dev: n_frames x (vosjpeg_scratch_bytes rounded
 * up to 8) bytes; rgb_dev: n_frames x height x width x 3. */
int vosjpeg_reconstruct_batch(const vosjpeg_info* info, int32_t n_frames, const int16_t* coef_dev, int64_t coef_stride,
                              const uint16_t* quant_dev, int64_t quant_stride, uint8_t* scratch_dev, uint8_t* rgb_dev, void* stream);

/* The same stage on the host (single thread), for files decoded where no GPU is wanted and for the CPU tests of the
 * host code; identical output. */
int vosjpeg_reconstruct_host(const vosjpeg_info* info, const int16_t* coef, uint8_t* rgb);

const char* vosjpeg_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
