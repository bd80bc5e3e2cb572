// JPEG front end (include/vos_jpeg.h): host-side marker parsing + Huffman decoding, device-side reconstruction, with pixels
// bit-identical to Pillow / libjpeg-turbo's default path (the reference's loader, src/utils/datasets.py:141-143).
// The arithmetic restates libjpeg-turbo's integer code paths from their published definitions:
//   inverse DCT      : jidctint.c jpeg_idct_islow (CONST_BITS 13, PASS1_BITS 2, Loeffler-Ligtenberg-Moschytz, 12 multiplies)
//   up-sampling      : jdsample.c h2v1_fancy_upsample / h2v2_fancy_upsample (triangle filter; plain replication when the
//                      component is at most two samples wide)
//   colour           : jdcolor.c build_ycc_rgb_table / ycc_rgb_convert (16-bit fixed point, ONE_HALF rounding)
// HBM-bound byte / integer work: 1.2 MB of coefficients in, 1.2 MB of RGB out per 480p frame; no tensor cores involved.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "../../include/vos_jpeg.h"

namespace vosj {

thread_local char g_err[256] = "";
int jfail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// zig-zag position k -> natural (row-major) position
const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// ------------------------------------------------------------------------------------------------
// Huffman tables (ITU T.81 Annex C / F.2.2.3): an 11-bit look-ahead table for the short codes, the
// canonical maxcode / valptr arrays for the rest.
// ------------------------------------------------------------------------------------------------
constexpr int kLook = 11;
struct HuffTable {
    bool present = false;
    uint16_t look[1 << kLook];      // (code length << 8) | symbol; 0: code longer than kLook bits
    int16_t fast_ac[1 << kLook];    // AC tables: (value << 8) | (run << 4) | (code length + value bits) when both fit the look-ahead
                                    // window and the value fits 8 bits; 0 otherwise.  Most AC coefficients of a photograph do.
    int32_t maxcode[18];            // largest code of each length (-1: none); [17] = sentinel
    int32_t valoff[17];             // symbol index of the first code of each length minus that code
    uint8_t symbols[256];
};

bool build_table(const uint8_t* counts, const uint8_t* symbols, int n_sym, HuffTable& t) {
    memset(&t, 0, sizeof(t));
    memcpy(t.symbols, symbols, n_sym);
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
        t.valoff[len] = k - code;
        if (counts[len - 1]) {
            for (int i = 0; i < counts[len - 1]; ++i, ++code, ++k) {
                if (len <= kLook) {
                    const int first = code << (kLook - len);
                    for (int j = 0; j < (1 << (kLook - len)); ++j) t.look[first + j] = static_cast<uint16_t>((len << 8) | symbols[k]);
                }
            }
            t.maxcode[len] = code - 1;
            if (code > (1 << len)) return false;
        } else {
            t.maxcode[len] = -1;
        }
        code <<= 1;
    }
    t.maxcode[17] = 0x7fffffff;
    t.present = true;
    for (int i = 0; i < (1 << kLook); ++i) {
        const uint32_t e = t.look[i];
        const int len = static_cast<int>(e >> 8), run = static_cast<int>((e >> 4) & 15u), mag = static_cast<int>(e & 15u);
        if (e == 0 || mag == 0 || len + mag > kLook) continue;
        int v = ((i << len) & ((1 << kLook) - 1)) >> (kLook - mag);          // the value bits that follow the code
        if (v < (1 << (mag - 1))) v += 1 - (1 << mag);                         // T.81 F.2.2.1 EXTEND
        if (v >= -128 && v <= 127) t.fast_ac[i] = static_cast<int16_t>(v * 256 + run * 16 + len + mag);
    }
    return k == n_sym;
}

// Entropy-coded segment reader: removes the stuffed zero after 0xFF, feeds zero bits once a marker is reached.
// ensure() keeps at least 32 bits in the accumulator: enough for one Huffman code (<= 16 bits) plus its value bits (<= 15).
struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t acc = 0;
    int n = 0;
    bool hit_marker = false;
    inline void slow_byte() {
        uint32_t b = 0;
        if (!hit_marker && p < end) {
            b = *p;
            if (b == 0xFF) {
                if (p + 1 < end && p[1] == 0) {
                    p += 2;
                } else {
                    hit_marker = true;
                    b = 0;
                }
            } else {
                ++p;
            }
        }
        acc = (acc << 8) | b;
        n += 8;
    }
    inline void ensure() {
        if (n >= 32) return;
        if (!hit_marker && end - p >= 4) {
            uint32_t le;
            memcpy(&le, p, 4);
            const uint32_t v = __builtin_bswap32(le);
            const uint32_t inv = ~v;
            if (((inv - 0x01010101u) & ~inv & 0x80808080u) == 0u) {      // no 0xFF among the four bytes
                acc = (acc << 32) | v;
                n += 32;
                p += 4;
                return;
            }
        }
        while (n <= 56 - 24) slow_byte();                               // up to four bytes, one at a time
    }
    inline uint32_t peek(int k) const { return static_cast<uint32_t>(acc >> (n - k)) & ((1u << k) - 1u); }
    inline void skip(int k) { n -= k; }
    inline uint32_t get(int k) {
        const uint32_t v = peek(k);
        n -= k;
        return v;
    }
    // to the byte after the next RSTn marker
    bool restart() {
        acc = 0;
        n = 0;
        hit_marker = false;
        while (p + 1 < end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) ++p;
        if (p + 1 >= end) return false;
        p += 2;
        return true;
    }
};

// needs >= 16 bits in the accumulator (ensure())
inline int decode_symbol(BitReader& br, const HuffTable& t) {
    const uint32_t e = t.look[br.peek(kLook)];
    if (e) {
        br.skip(static_cast<int>(e >> 8));
        return static_cast<int>(e & 255u);
    }
    int l = kLook + 1;
    int32_t code = static_cast<int32_t>(br.peek(l));
    while (l <= 16 && code > t.maxcode[l]) {
        ++l;
        code = static_cast<int32_t>(br.peek(l));
    }
    if (l > 16) return -1;
    br.skip(l);
    return t.symbols[(code + t.valoff[l]) & 255];
}

// T.81 F.2.2.1 EXTEND, branch-free: values below 2^(s-1) are negative: v - (2^s - 1)
inline int extend(uint32_t v, int s) {
    const int x = static_cast<int>(v);
    return x + (((x - (1 << (s - 1))) >> 31) & (1 - (1 << s)));
}

struct Parsed {
    vosjpeg_info info;
    HuffTable dc[4], ac[4];
};

int parse_impl(const uint8_t* d, int64_t size, Parsed& ps, bool want_tables) {
    vosjpeg_info& info = ps.info;
    memset(&info, 0, sizeof(info));
    if (size < 4 || d[0] != 0xFF || d[1] != 0xD8) return jfail(VOSJPEG_ERR_INVALID, "not a JPEG stream");
    uint16_t qt[4][64];
    bool qt_present[4] = {false, false, false, false};
    int comp_id[3] = {0, 0, 0}, comp_tq[3] = {0, 0, 0};
    int adobe_transform = -1;
    bool have_sof = false, saw_jfif = false;
    int64_t i = 2;
    for (;;) {
        if (i + 4 > size) return jfail(VOSJPEG_ERR_INVALID, "truncated before the scan");
        if (d[i] != 0xFF) return jfail(VOSJPEG_ERR_INVALID, "marker expected at byte %lld", static_cast<long long>(i));
        while (i < size && d[i] == 0xFF) ++i;
        if (i >= size) return jfail(VOSJPEG_ERR_INVALID, "truncated before the scan");
        const int m = d[i++];
        if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (m == 0xD9) return jfail(VOSJPEG_ERR_INVALID, "end of image before a scan");
        if (i + 2 > size) return jfail(VOSJPEG_ERR_INVALID, "truncated segment");
        const int len = (d[i] << 8) | d[i + 1];
        if (len < 2 || i + len > size) return jfail(VOSJPEG_ERR_INVALID, "bad segment length");
        const uint8_t* s = d + i + 2;
        const int n = len - 2;
        i += len;
        if (m == 0xDB) {
            int j = 0;
            while (j < n) {
                const int pq = s[j] >> 4, tq = s[j] & 15;
                ++j;
                if (tq > 3 || j + (pq ? 128 : 64) > n) return jfail(VOSJPEG_ERR_INVALID, "bad quantisation table");
                for (int k = 0; k < 64; ++k) {
                    qt[tq][kZigzag[k]] = pq ? static_cast<uint16_t>((s[j] << 8) | s[j + 1]) : s[j];
                    j += pq ? 2 : 1;
                }
                qt_present[tq] = true;
            }
        } else if (m == 0xC0 || m == 0xC1) {
            if (n < 6 || s[0] != 8) return jfail(VOSJPEG_ERR_UNSUPPORTED, "sample precision %d", n ? s[0] : 0);
            info.height = (s[1] << 8) | s[2];
            info.width = (s[3] << 8) | s[4];
            info.n_comp = s[5];
            if (info.n_comp != 1 && info.n_comp != 3) return jfail(VOSJPEG_ERR_UNSUPPORTED, "%d components", info.n_comp);
            if (n < 6 + 3 * info.n_comp || info.width <= 0 || info.height <= 0) return jfail(VOSJPEG_ERR_INVALID, "bad frame header");
            for (int c = 0; c < info.n_comp; ++c) {
                comp_id[c] = s[6 + 3 * c];
                info.h_samp[c] = s[7 + 3 * c] >> 4;
                info.v_samp[c] = s[7 + 3 * c] & 15;
                comp_tq[c] = s[8 + 3 * c];
                if (comp_tq[c] > 3) return jfail(VOSJPEG_ERR_INVALID, "bad table selector");
            }
            have_sof = true;
        } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return jfail(VOSJPEG_ERR_UNSUPPORTED, "SOF%d (progressive, lossless or arithmetic coding)", m - 0xC0);
        } else if (m == 0xC4) {
            int j = 0;
            while (j < n) {
                if (j + 17 > n) return jfail(VOSJPEG_ERR_INVALID, "bad Huffman table");
                const int tc = s[j] >> 4, th = s[j] & 15;
                int total = 0;
                for (int k = 0; k < 16; ++k) total += s[j + 1 + k];
                if (tc > 1 || th > 3 || total > 256 || j + 17 + total > n) return jfail(VOSJPEG_ERR_INVALID, "bad Huffman table");
                if (want_tables && !build_table(s + j + 1, s + j + 17, total, tc ? ps.ac[th] : ps.dc[th]))
                    return jfail(VOSJPEG_ERR_INVALID, "inconsistent Huffman table");
                j += 17 + total;
            }
        } else if (m == 0xDD) {
            if (n < 2) return jfail(VOSJPEG_ERR_INVALID, "bad restart interval");
            info.restart_interval = (s[0] << 8) | s[1];
        } else if (m == 0xE0 && n >= 5 && memcmp(s, "JFIF", 5) == 0) {
            saw_jfif = true;
        } else if (m == 0xEE && n >= 12 && memcmp(s, "Adobe", 5) == 0) {
            adobe_transform = s[11];
        } else if (m == 0xDA) {
            if (!have_sof) return jfail(VOSJPEG_ERR_INVALID, "scan before frame header");
            const int ns = n ? s[0] : 0;
            if (ns != info.n_comp) return jfail(VOSJPEG_ERR_UNSUPPORTED, "scan with %d of %d components", ns, info.n_comp);
            if (n < 4 + 2 * ns) return jfail(VOSJPEG_ERR_INVALID, "bad scan header");
            for (int c = 0; c < ns; ++c) {
                if (s[1 + 2 * c] != comp_id[c]) return jfail(VOSJPEG_ERR_UNSUPPORTED, "scan components out of frame order");
                info.dc_table[c] = s[2 + 2 * c] >> 4;
                info.ac_table[c] = s[2 + 2 * c] & 15;
                if (info.dc_table[c] > 3 || info.ac_table[c] > 3) return jfail(VOSJPEG_ERR_INVALID, "bad table selector");
            }
            if (s[1 + 2 * ns] != 0 || s[2 + 2 * ns] != 63 || s[3 + 2 * ns] != 0)
                return jfail(VOSJPEG_ERR_UNSUPPORTED, "spectral selection / successive approximation");
            info.scan_offset = i;
            break;
        }
    }
    if (info.n_comp == 3 && adobe_transform != -1 && adobe_transform != 1)
        return jfail(VOSJPEG_ERR_UNSUPPORTED, "Adobe colour transform %d (RGB / CMYK data)", adobe_transform);
    // libjpeg's colour-space guess (jdapimin.c default_decompress_parms): without a JFIF or Adobe marker, components named
    // 'R', 'G', 'B' are taken as RGB data and not converted
    if (info.n_comp == 3 && !saw_jfif && adobe_transform == -1 && comp_id[0] == 'R' && comp_id[1] == 'G' && comp_id[2] == 'B')
        return jfail(VOSJPEG_ERR_UNSUPPORTED, "components named R, G, B without a JFIF / Adobe marker (RGB data)");
    if (info.n_comp == 1) info.h_samp[0] = info.v_samp[0] = 1;    // a one-component scan is never interleaved: MCU = one block
    int hmax = 1, vmax = 1;
    for (int c = 0; c < info.n_comp; ++c) {
        if (info.h_samp[c] < 1 || info.v_samp[c] < 1 || info.h_samp[c] > 4 || info.v_samp[c] > 4)
            return jfail(VOSJPEG_ERR_INVALID, "bad sampling factors");
        hmax = info.h_samp[c] > hmax ? info.h_samp[c] : hmax;
        vmax = info.v_samp[c] > vmax ? info.v_samp[c] : vmax;
    }
    for (int c = 0; c < info.n_comp; ++c) {
        // bit-exact up-sampling exists for 1:1, 2:1 horizontal and 2:1 x 2:1 (what 4:4:4, 4:2:2 and 4:2:0 files use)
        const int fh = hmax / info.h_samp[c], fv = vmax / info.v_samp[c];
        const bool ok = hmax % info.h_samp[c] == 0 && vmax % info.v_samp[c] == 0 &&
                        ((fh == 1 && fv == 1) || (fh == 2 && fv == 1) || (fh == 2 && fv == 2));
        if (!ok) return jfail(VOSJPEG_ERR_UNSUPPORTED, "sampling factors %dx%d of %dx%d", info.h_samp[c], info.v_samp[c], hmax, vmax);
        if (c == 0 && (fh != 1 || fv != 1)) return jfail(VOSJPEG_ERR_UNSUPPORTED, "first component below full resolution");
        if (!qt_present[comp_tq[c]]) return jfail(VOSJPEG_ERR_INVALID, "missing quantisation table %d", comp_tq[c]);
        memcpy(info.quant[c], qt[comp_tq[c]], sizeof(qt[0]));
    }
    const int mcus_x = (info.width + 8 * hmax - 1) / (8 * hmax), mcus_y = (info.height + 8 * vmax - 1) / (8 * vmax);
    int64_t off = 0;
    for (int c = 0; c < info.n_comp; ++c) {
        info.blocks_w[c] = mcus_x * info.h_samp[c];
        info.blocks_h[c] = mcus_y * info.v_samp[c];
        info.coef_offset[c] = off;
        off += static_cast<int64_t>(info.blocks_w[c]) * info.blocks_h[c] * 64;
    }
    info.coef_count = off;
    if (want_tables)
        for (int c = 0; c < info.n_comp; ++c)
            if (!ps.dc[info.dc_table[c]].present || !ps.ac[info.ac_table[c]].present)
                return jfail(VOSJPEG_ERR_INVALID, "missing Huffman table");
    return VOSJPEG_OK;
}

// ------------------------------------------------------------------------------------------------
// Reconstruction arithmetic, shared by the kernels and the host path.
// ------------------------------------------------------------------------------------------------
#define VJ_HD __host__ __device__ __forceinline__

VJ_HD int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// one 8-point pass of jpeg_idct_islow; `up` = CONST_BITS shift of the even part's DC terms, `down` = descale of the outputs
VJ_HD void idct8(int v0, int v1, int v2, int v3, int v4, int v5, int v6, int v7, int down, int* o) {
    int z1 = (v2 + v6) * 4433;
    const int tmp2 = z1 + v6 * -15137;
    const int tmp3 = z1 + v2 * 6270;
    const int tmp0 = (v0 + v4) * 8192;
    const int tmp1 = (v0 - v4) * 8192;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    int t0 = v7, t1 = v5, t2 = v3, t3 = v1;
    z1 = t0 + t3;
    int z2 = t1 + t2, z3 = t0 + t2, z4 = t1 + t3;
    const int z5 = (z3 + z4) * 9633;
    t0 *= 2446;
    t1 *= 16819;
    t2 *= 25172;
    t3 *= 12299;
    z1 *= -7373;
    z2 *= -20995;
    z3 = z3 * -16069 + z5;
    z4 = z4 * -3196 + z5;
    t0 += z1 + z3;
    t1 += z2 + z4;
    t2 += z2 + z3;
    t3 += z1 + z4;
    o[0] = descale(tmp10 + t3, down);
    o[7] = descale(tmp10 - t3, down);
    o[1] = descale(tmp11 + t2, down);
    o[6] = descale(tmp11 - t2, down);
    o[2] = descale(tmp12 + t1, down);
    o[5] = descale(tmp12 - t1, down);
    o[3] = descale(tmp13 + t0, down);
    o[4] = descale(tmp13 - t0, down);
}

// the IDCT's range-limit table (jdmaster.c prepare_range_limit_table), indexed with (x & 1023), samples centred on 128
VJ_HD uint8_t idct_limit(int x) {
    x &= 1023;
    return static_cast<uint8_t>(x < 128 ? x + 128 : (x < 512 ? 255 : (x < 896 ? 0 : x - 896)));
}

// one 8x8 block: quantised coefficients (natural order) -> 64 samples, row-major with `stride`
VJ_HD void idct_block(const int16_t* coef, const uint16_t* q, uint8_t* out, int stride) {
    int ws[64];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        int o[8];
        idct8(coef[c] * q[c], coef[8 + c] * q[8 + c], coef[16 + c] * q[16 + c], coef[24 + c] * q[24 + c], coef[32 + c] * q[32 + c],
              coef[40 + c] * q[40 + c], coef[48 + c] * q[48 + c], coef[56 + c] * q[56 + c], 11, o);
#pragma unroll
        for (int r = 0; r < 8; ++r) ws[r * 8 + c] = o[r];
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int o[8];
        idct8(ws[r * 8], ws[r * 8 + 1], ws[r * 8 + 2], ws[r * 8 + 3], ws[r * 8 + 4], ws[r * 8 + 5], ws[r * 8 + 6], ws[r * 8 + 7], 18, o);
#pragma unroll
        for (int c = 0; c < 8; ++c) out[r * stride + c] = idct_limit(o[c]);
    }
}

struct PlaneGeom {
    int32_t stride;      // samples per row of the padded plane (blocks_w * 8)
    int32_t w, h;        // real samples: ceil(width * h_samp / hmax), ceil(height * v_samp / vmax)
    int32_t fh, fv;      // up-sampling factors to the image grid (1 or 2)
    int64_t offset;      // first byte of the plane inside the scratch buffer
};

// sample of an up-sampled component at image position (x, y)
VJ_HD int upsampled(const uint8_t* plane, const PlaneGeom& g, int x, int y) {
    if (g.fh == 1) return plane[static_cast<int64_t>(y) * g.stride + x];
    const int j = x >> 1;
    if (g.fv == 1) {
        const uint8_t* row = plane + static_cast<int64_t>(y) * g.stride;
        const int a = row[j];
        if (g.w <= 2) return a;                                       // h2v1_upsample: replication
        if (x & 1) return j == g.w - 1 ? a : (3 * a + row[j + 1] + 2) >> 2;
        return j == 0 ? a : (3 * a + row[j - 1] + 1) >> 2;
    }
    const int i = y >> 1;
    const uint8_t* row0 = plane + static_cast<int64_t>(i) * g.stride;
    if (g.w <= 2) return row0[j];                                     // h2v2_upsample: replication
    const int io = (y & 1) ? (i + 1 < g.h ? i + 1 : g.h - 1) : (i > 0 ? i - 1 : 0);      // the nearer neighbouring row, edges replicated
    const uint8_t* row1 = plane + static_cast<int64_t>(io) * g.stride;
    const int s = 3 * row0[j] + row1[j];
    if (x & 1) return j == g.w - 1 ? (4 * s + 7) >> 4 : (3 * s + 3 * row0[j + 1] + row1[j + 1] + 7) >> 4;
    return j == 0 ? (4 * s + 8) >> 4 : (3 * s + 3 * row0[j - 1] + row1[j - 1] + 8) >> 4;
}

VJ_HD uint8_t clamp255(int v) { return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// jdcolor.c: FIX(1.40200) = 91881, FIX(1.77200) = 116130, FIX(0.71414) = 46802, FIX(0.34414) = 22554, ONE_HALF = 32768
VJ_HD void ycc_rgb(int y, int cb, int cr, uint8_t* rgb) {
    cb -= 128;
    cr -= 128;
    rgb[0] = clamp255(y + ((91881 * cr + 32768) >> 16));
    rgb[1] = clamp255(y + ((-22554 * cb + 32768 - 46802 * cr) >> 16));
    rgb[2] = clamp255(y + ((116130 * cb + 32768) >> 16));
}

struct ReconParams {
    vosjpeg_info info;
    PlaneGeom plane[3];
};

ReconParams make_params(const vosjpeg_info& info) {
    ReconParams rp;
    rp.info = info;
    int hmax = 1, vmax = 1;
    for (int c = 0; c < info.n_comp; ++c) {
        hmax = info.h_samp[c] > hmax ? info.h_samp[c] : hmax;
        vmax = info.v_samp[c] > vmax ? info.v_samp[c] : vmax;
    }
    int64_t off = 0;
    for (int c = 0; c < 3; ++c) {
        PlaneGeom& g = rp.plane[c];
        if (c >= info.n_comp) {
            g = rp.plane[0];
            continue;
        }
        g.stride = info.blocks_w[c] * 8;
        g.w = (info.width * info.h_samp[c] + hmax - 1) / hmax;
        g.h = (info.height * info.v_samp[c] + vmax - 1) / vmax;
        g.fh = hmax / info.h_samp[c];
        g.fv = vmax / info.v_samp[c];
        g.offset = off;
        off += static_cast<int64_t>(g.stride) * info.blocks_h[c] * 8;
    }
    return rp;
}

bool valid_info(const vosjpeg_info* info) {
    if (!info || info->width <= 0 || info->height <= 0 || (info->n_comp != 1 && info->n_comp != 3)) return false;
    for (int c = 0; c < info->n_comp; ++c)
        if (info->blocks_w[c] <= 0 || info->blocks_h[c] <= 0 || info->h_samp[c] < 1 || info->v_samp[c] < 1) return false;
    return true;
}

// ---- kernels ----
// A launch handles a batch of frames of one geometry (grid y / z = frame): frame f reads its coefficients at coef + f * coef_stride
// and its quantisation tables at quant + f * quant_stride (device memory: the loader item's own header), or the geometry's tables
// when quant is null.
struct BatchParams {
    ReconParams rp;
    int64_t coef_stride;        // int16 elements between frames
    const uint16_t* quant;      // per-frame tables [3][64] in device memory, or null
    int64_t quant_stride;       // uint16 elements between frames
    int64_t scratch_stride;     // bytes
};

// Eight threads per 8x8 block (32 blocks per CTA): thread r loads and de-quantises row r (16 contiguous bytes: a block's 128
// bytes are one coalesced line), the eight threads run the column pass (thread = column) and the row pass (thread = row) through
// shared memory, and thread r stores the 8 samples of row r (the 4 blocks of a warp that are neighbours in x write 32 contiguous
// bytes per row).  The block's slot is padded to 72 words so that the 4 blocks of a warp fall on different banks.
constexpr int kIdctBlocksPerCta = 32;
constexpr int kIdctSlot = 72;
__global__ void __launch_bounds__(kIdctBlocksPerCta * 8) vosjpeg_idct(const __grid_constant__ BatchParams bp, const int16_t* __restrict__ coef,
                                                                      uint8_t* __restrict__ planes, int64_t n_blocks) {
    __shared__ int ws[kIdctBlocksPerCta * kIdctSlot];
    __shared__ uint16_t qs[3 * 64];
    const ReconParams& rp = bp.rp;
    const int f = blockIdx.y;
    if (threadIdx.x < 3 * 64)
        qs[threadIdx.x] = bp.quant != nullptr ? __ldg(bp.quant + f * bp.quant_stride + threadIdx.x) : rp.info.quant[threadIdx.x >> 6][threadIdx.x & 63];
    __syncthreads();
    const int lb = threadIdx.x >> 3, t = threadIdx.x & 7;
    const int64_t b = static_cast<int64_t>(blockIdx.x) * kIdctBlocksPerCta + lb;
    if (b >= n_blocks) return;                       // whole groups of 8 lanes leave together; __syncwarp below names the live lanes
    const unsigned live = __activemask();
    int c = 0;
    int64_t local = b;
    while (c + 1 < rp.info.n_comp && local >= static_cast<int64_t>(rp.info.blocks_w[c]) * rp.info.blocks_h[c]) {
        local -= static_cast<int64_t>(rp.info.blocks_w[c]) * rp.info.blocks_h[c];
        ++c;
    }
    const int by = static_cast<int>(local / rp.info.blocks_w[c]), bx = static_cast<int>(local % rp.info.blocks_w[c]);
    int* w = ws + lb * kIdctSlot;
    {   // row t of the coefficient block, de-quantised
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(coef + f * bp.coef_stride + rp.info.coef_offset[c] + local * 64) + t);
        const int16_t* v = reinterpret_cast<const int16_t*>(&raw);
        const uint16_t* q = qs + c * 64 + t * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) w[t * 8 + i] = static_cast<int>(v[i]) * static_cast<int>(q[i]);
    }
    __syncwarp(live);
    int o[8];
    idct8(w[t], w[8 + t], w[16 + t], w[24 + t], w[32 + t], w[40 + t], w[48 + t], w[56 + t], 11, o);      // column t
    __syncwarp(live);
#pragma unroll
    for (int r = 0; r < 8; ++r) w[t * 8 + r] = o[r];                                                     // stored transposed: [column][row]
    __syncwarp(live);
    idct8(w[t], w[8 + t], w[16 + t], w[24 + t], w[32 + t], w[40 + t], w[48 + t], w[56 + t], 18, o);      // row t
    __align__(8) uint8_t px[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) px[i] = idct_limit(o[i]);
    uint8_t* dst = planes + f * bp.scratch_stride + rp.plane[c].offset + (static_cast<int64_t>(by) * 8 + t) * rp.plane[c].stride + bx * 8;
    *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(px);
}

// Four horizontally adjacent pixels per thread (x0 a multiple of 4): one 32-bit load of luma, the chroma samples of the two source
// columns and their neighbours once for all four pixels, 12 output bytes (16-bit stores when the row pitch allows).
struct Chroma4 {
    int v[4];
};
__device__ __forceinline__ Chroma4 upsampled4(const uint8_t* plane, const PlaneGeom& g, int x0, int y, int width) {
    Chroma4 r;
    if (g.fh == 1) {
        const uint8_t* row = plane + static_cast<int64_t>(y) * g.stride + x0;
#pragma unroll
        for (int i = 0; i < 4; ++i) r.v[i] = row[i];          // (the plane is padded to whole MCUs: reads past `width` stay inside it)
        return r;
    }
    const int j0 = x0 >> 1;                                   // source columns j0, j0 + 1; neighbours j0 - 1, j0 + 2
    const int jm = j0 > 0 ? j0 - 1 : 0, j1 = j0 + 1 < g.w ? j0 + 1 : g.w - 1, j2 = j0 + 2 < g.w ? j0 + 2 : g.w - 1;
    int sm, s0, s1, s2;                                       // column sums (h2v2) or samples (h2v1), scaled alike below
    if (g.fv == 1) {
        const uint8_t* row = plane + static_cast<int64_t>(y) * g.stride;
        sm = row[jm]; s0 = row[j0]; s1 = row[j1]; s2 = row[j2];
        if (g.w <= 2) {
            r.v[0] = r.v[1] = s0; r.v[2] = r.v[3] = s1;
            return r;
        }
        r.v[0] = j0 == 0 ? s0 : (3 * s0 + sm + 1) >> 2;
        r.v[1] = j0 == g.w - 1 ? s0 : (3 * s0 + s1 + 2) >> 2;
        r.v[2] = (3 * s1 + s0 + 1) >> 2;
        r.v[3] = j0 + 1 >= g.w - 1 ? s1 : (3 * s1 + s2 + 2) >> 2;
        return r;
    }
    const int i = y >> 1;
    const uint8_t* row0 = plane + static_cast<int64_t>(i) * g.stride;
    if (g.w <= 2) {
        r.v[0] = r.v[1] = row0[j0]; r.v[2] = r.v[3] = row0[j1];
        return r;
    }
    const int io = (y & 1) ? (i + 1 < g.h ? i + 1 : g.h - 1) : (i > 0 ? i - 1 : 0);
    const uint8_t* row1 = plane + static_cast<int64_t>(io) * g.stride;
    sm = 3 * row0[jm] + row1[jm]; s0 = 3 * row0[j0] + row1[j0]; s1 = 3 * row0[j1] + row1[j1]; s2 = 3 * row0[j2] + row1[j2];
    r.v[0] = j0 == 0 ? (4 * s0 + 8) >> 4 : (3 * s0 + sm + 8) >> 4;
    r.v[1] = j0 == g.w - 1 ? (4 * s0 + 7) >> 4 : (3 * s0 + s1 + 7) >> 4;
    r.v[2] = (3 * s1 + s0 + 8) >> 4;
    r.v[3] = j0 + 1 >= g.w - 1 ? (4 * s1 + 7) >> 4 : (3 * s1 + s2 + 7) >> 4;
    (void)width;
    return r;
}

__global__ void __launch_bounds__(128) vosjpeg_colour(const __grid_constant__ BatchParams bp, const uint8_t* __restrict__ planes,
                                                      uint8_t* __restrict__ rgb) {
    const ReconParams& rp = bp.rp;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, f = blockIdx.z;
    const int width = rp.info.width;
    if (x0 >= width) return;
    planes += f * bp.scratch_stride;
    const uint32_t luma = __ldg(reinterpret_cast<const uint32_t*>(planes + rp.plane[0].offset + static_cast<int64_t>(y) * rp.plane[0].stride + x0));
    __align__(4) uint8_t px[12];
    if (rp.info.n_comp == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) px[3 * i] = px[3 * i + 1] = px[3 * i + 2] = static_cast<uint8_t>((luma >> (8 * i)) & 255u);
    } else {
        const Chroma4 cb = upsampled4(planes + rp.plane[1].offset, rp.plane[1], x0, y, width);
        const Chroma4 cr = upsampled4(planes + rp.plane[2].offset, rp.plane[2], x0, y, width);
#pragma unroll
        for (int i = 0; i < 4; ++i) ycc_rgb(static_cast<int>((luma >> (8 * i)) & 255u), cb.v[i], cr.v[i], px + 3 * i);
    }
    uint8_t* out = rgb + ((static_cast<int64_t>(f) * rp.info.height + y) * width + x0) * 3;
    const int n = width - x0 < 4 ? width - x0 : 4;
    if (n == 4 && (reinterpret_cast<uintptr_t>(out) & 1) == 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) reinterpret_cast<uint16_t*>(out)[i] = reinterpret_cast<const uint16_t*>(px)[i];
    } else {
        for (int i = 0; i < 3 * n; ++i) out[i] = px[i];
    }
}

}  // namespace vosj

using namespace vosj;

extern "C" {

const char* vosjpeg_last_error(void) { return g_err; }

int vosjpeg_parse(const uint8_t* data, int64_t size, vosjpeg_info* info) {
    if (!data || !info) return jfail(VOSJPEG_ERR_INVALID, "null pointer");
    Parsed* ps = new (std::nothrow) Parsed();
    if (!ps) return jfail(VOSJPEG_ERR_INVALID, "out of memory");
    const int rc = parse_impl(data, size, *ps, false);
    if (rc == VOSJPEG_OK) *info = ps->info;
    delete ps;
    return rc;
}

int vosjpeg_entropy_decode(const uint8_t* data, int64_t size, const vosjpeg_info* info, int16_t* coef) {
    if (!data || !info || !coef) return jfail(VOSJPEG_ERR_INVALID, "null pointer");
    Parsed* ps = new (std::nothrow) Parsed();
    if (!ps) return jfail(VOSJPEG_ERR_INVALID, "out of memory");
    int rc = parse_impl(data, size, *ps, true);
    if (rc == VOSJPEG_OK && memcmp(&ps->info, info, sizeof(*info)) != 0) rc = jfail(VOSJPEG_ERR_INVALID, "info does not belong to this stream");
    if (rc != VOSJPEG_OK) {
        delete ps;
        return rc;
    }
    memset(coef, 0, static_cast<size_t>(info->coef_count) * sizeof(int16_t));
    BitReader br;
    br.p = data + info->scan_offset;
    br.end = data + size;
    int hmax = 1, vmax = 1;
    for (int c = 0; c < info->n_comp; ++c) {
        hmax = info->h_samp[c] > hmax ? info->h_samp[c] : hmax;
        vmax = info->v_samp[c] > vmax ? info->v_samp[c] : vmax;
    }
    const int mcus_x = info->blocks_w[0] / info->h_samp[0], mcus_y = info->blocks_h[0] / info->v_samp[0];
    int pred[3] = {0, 0, 0};
    int64_t count = 0;
    for (int my = 0; my < mcus_y && rc == VOSJPEG_OK; ++my) {
        for (int mx = 0; mx < mcus_x && rc == VOSJPEG_OK; ++mx) {
            if (info->restart_interval && count && count % info->restart_interval == 0) {
                if (!br.restart()) {
                    rc = jfail(VOSJPEG_ERR_INVALID, "restart marker missing");
                    break;
                }
                pred[0] = pred[1] = pred[2] = 0;
            }
            ++count;
            for (int c = 0; c < info->n_comp; ++c) {
                const HuffTable& dct = ps->dc[info->dc_table[c]];
                const HuffTable& act = ps->ac[info->ac_table[c]];
                for (int by = 0; by < info->v_samp[c]; ++by) {
                    for (int bx = 0; bx < info->h_samp[c]; ++bx) {
                        int16_t* blk = coef + info->coef_offset[c] +
                                       (static_cast<int64_t>(my * info->v_samp[c] + by) * info->blocks_w[c] + mx * info->h_samp[c] + bx) * 64;
                        br.ensure();
                        int s = decode_symbol(br, dct);
                        if (s < 0 || s > 15) {
                            rc = jfail(VOSJPEG_ERR_INVALID, "corrupt entropy-coded data");
                            goto done;
                        }
                        if (s) pred[c] += extend(br.get(s), s);
                        blk[0] = static_cast<int16_t>(pred[c]);
                        for (int k = 1; k < 64;) {
                            br.ensure();
                            const int fast = act.fast_ac[br.peek(kLook)];
                            if (fast) {                                     // code and value in one look-up
                                k += (fast >> 4) & 15;
                                if (k > 63) {
                                    rc = jfail(VOSJPEG_ERR_INVALID, "corrupt entropy-coded data");
                                    goto done;
                                }
                                br.skip(fast & 15);
                                blk[kZigzag[k]] = static_cast<int16_t>(fast >> 8);
                                ++k;
                                continue;
                            }
                            const int rs = decode_symbol(br, act);
                            if (rs < 0) {
                                rc = jfail(VOSJPEG_ERR_INVALID, "corrupt entropy-coded data");
                                goto done;
                            }
                            const int r = rs >> 4;
                            s = rs & 15;
                            if (s == 0) {
                                if (r != 15) break;
                                k += 16;
                                continue;
                            }
                            k += r;
                            if (k > 63) {
                                rc = jfail(VOSJPEG_ERR_INVALID, "corrupt entropy-coded data");
                                goto done;
                            }
                            blk[kZigzag[k]] = static_cast<int16_t>(extend(br.get(s), s));
                            ++k;
                        }
                    }
                }
            }
        }
    }
done:
    delete ps;
    return rc;
}

int vosjpeg_decode_files_host(const uint8_t* const* datas, const int64_t* sizes, int32_t n_files, int16_t* const* items, int64_t item_capacity,
                              int32_t header_values, int32_t n_threads, int32_t* status) {
    if (!datas || !sizes || !items || !status || n_files < 0) return jfail(VOSJPEG_ERR_INVALID, "null pointer");
    if (header_values < 0 || static_cast<size_t>(header_values) * sizeof(int16_t) < sizeof(vosjpeg_info) || (header_values & 7))
        return jfail(VOSJPEG_ERR_INVALID, "header: at least sizeof(vosjpeg_info) bytes, a multiple of 8 values");
    std::atomic<int32_t> next(0);
    auto work = [&]() {
        for (;;) {
            const int32_t i = next.fetch_add(1);
            if (i >= n_files) return;
            vosjpeg_info info;
            int rc = vosjpeg_parse(datas[i], sizes[i], &info);
            if (rc == VOSJPEG_OK && header_values + info.coef_count > item_capacity) rc = VOSJPEG_ERR_UNSUPPORTED;    // larger than the caller planned for
            if (rc == VOSJPEG_OK) {
                memset(items[i], 0, static_cast<size_t>(header_values) * sizeof(int16_t));
                memcpy(items[i], &info, sizeof(info));
                rc = vosjpeg_entropy_decode(datas[i], sizes[i], &info, items[i] + header_values);
            }
            status[i] = rc;
        }
    };
    const int32_t n_thr = n_threads < 1 ? 1 : (n_threads > n_files ? n_files : n_threads);
    std::vector<std::thread> pool;
    try {
        for (int32_t t = 1; t < n_thr; ++t) pool.emplace_back(work);
    } catch (...) {
        // no more threads to be had: the ones that started and this one share the files
    }
    work();
    for (auto& th : pool) th.join();
    return VOSJPEG_OK;
}

int64_t vosjpeg_scratch_bytes(const vosjpeg_info* info) {
    if (!valid_info(info)) return jfail(VOSJPEG_ERR_INVALID, "bad info");
    int64_t total = 0;
    for (int c = 0; c < info->n_comp; ++c) total += static_cast<int64_t>(info->blocks_w[c]) * info->blocks_h[c] * 64;
    return total;
}

int vosjpeg_reconstruct_batch(const vosjpeg_info* info, int32_t n_frames, const int16_t* coef_dev, int64_t coef_stride,
                              const uint16_t* quant_dev, int64_t quant_stride, uint8_t* scratch_dev, uint8_t* rgb_dev, void* stream) {
    if (!valid_info(info) || !coef_dev || !scratch_dev || !rgb_dev) return jfail(VOSJPEG_ERR_INVALID, "null pointer or bad info");
    if (n_frames < 1 || n_frames > 65535 || info->height > 65535) return jfail(VOSJPEG_ERR_INVALID, "1 .. 65535 frames (and rows) per launch");
    if ((reinterpret_cast<uintptr_t>(coef_dev) & 15) || (coef_stride & 7) || (reinterpret_cast<uintptr_t>(scratch_dev) & 7) ||
        (reinterpret_cast<uintptr_t>(quant_dev) & 1))
        return jfail(VOSJPEG_ERR_INVALID, "coefficients must be 16-byte aligned (stride a multiple of 8 values), scratch 8-byte aligned");
    if (n_frames > 1 && coef_stride < info->coef_count) return jfail(VOSJPEG_ERR_INVALID, "coefficient stride shorter than a frame");
    BatchParams bp;
    bp.rp = make_params(*info);
    bp.coef_stride = coef_stride;
    bp.quant = quant_dev;
    bp.quant_stride = quant_stride;
    bp.scratch_stride = (vosjpeg_scratch_bytes(info) + 7) & ~static_cast<int64_t>(7);
    const int64_t n_blocks = info->coef_count / 64;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    vosjpeg_idct<<<dim3(static_cast<unsigned>((n_blocks + kIdctBlocksPerCta - 1) / kIdctBlocksPerCta), n_frames), kIdctBlocksPerCta * 8, 0, st>>>(
        bp, coef_dev, scratch_dev, n_blocks);
    vosjpeg_colour<<<dim3((info->width + 511) / 512, info->height, n_frames), 128, 0, st>>>(bp, scratch_dev, rgb_dev);
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return jfail(VOSJPEG_ERR_CUDA, "%s", cudaGetErrorString(err));
    return VOSJPEG_OK;
}

int vosjpeg_reconstruct(const vosjpeg_info* info, const int16_t* coef_dev, uint8_t* scratch_dev, uint8_t* rgb_dev, void* stream) {
    return vosjpeg_reconstruct_batch(info, 1, coef_dev, info ? info->coef_count : 0, nullptr, 0, scratch_dev, rgb_dev, stream);
}

int vosjpeg_reconstruct_host(const vosjpeg_info* info, const int16_t* coef, uint8_t* rgb) {
    if (!valid_info(info) || !coef || !rgb) return jfail(VOSJPEG_ERR_INVALID, "null pointer or bad info");
    const ReconParams rp = make_params(*info);
    const int64_t bytes = vosjpeg_scratch_bytes(info);
    uint8_t* planes = new (std::nothrow) uint8_t[bytes];
    if (!planes) return jfail(VOSJPEG_ERR_INVALID, "out of memory");
    for (int c = 0; c < info->n_comp; ++c)
        for (int by = 0; by < info->blocks_h[c]; ++by)
            for (int bx = 0; bx < info->blocks_w[c]; ++bx)
                idct_block(coef + info->coef_offset[c] + (static_cast<int64_t>(by) * info->blocks_w[c] + bx) * 64, info->quant[c],
                           planes + rp.plane[c].offset + (static_cast<int64_t>(by) * 8) * rp.plane[c].stride + bx * 8, rp.plane[c].stride);
    for (int y = 0; y < info->height; ++y) {
        for (int x = 0; x < info->width; ++x) {
            uint8_t* out = rgb + (static_cast<int64_t>(y) * info->width + x) * 3;
            const int yy = upsampled(planes + rp.plane[0].offset, rp.plane[0], x, y);
            if (info->n_comp == 1) {
                out[0] = out[1] = out[2] = static_cast<uint8_t>(yy);
            } else {
                ycc_rgb(yy, upsampled(planes + rp.plane[1].offset, rp.plane[1], x, y), upsampled(planes + rp.plane[2].offset, rp.plane[2], x, y), out);
            }
        }
    }
    delete[] planes;
    return VOSJPEG_OK;
}

}  // extern "C"
