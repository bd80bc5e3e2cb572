// Work decomposition of the affinity kernel, shared by host and device code.
//
// The affinity of one target frame against R reference frames is a grid of
//   m_tiles x n_tiles  (m_tiles = TPF = ceil(P/128) target tiles; n_tiles = NT = R * TPF reference tiles)
// 128x128 logit tiles.  480p has only 51 target tiles for 148 SMs, so the
// linearised tile space (m-major) is cut into G = min(#SM, total) contiguous, equal ranges
// ("stream-K"): CTA c owns [cta_begin(c), cta_begin(c+1)).  A range may straddle m-tiles; each
// (CTA, m-tile) piece is a *segment* that produces one online-softmax partial per row, merged
// later by vos_merge_writeback.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define VOS_HD __host__ __device__ __forceinline__
#else
#define VOS_HD inline
#endif

namespace vosd {

constexpr int kTile = 128;

struct Decomp {
    int32_t tpf;      // tiles per frame = ceil(P / 128)
    int32_t nt;       // reference tiles per target tile = R * tpf
    int32_t grid;     // CTAs
    int32_t max_segs; // upper bound on segments per CTA
    int64_t total;    // tpf * nt
};

VOS_HD Decomp make_decomp(int32_t n_pixels, int32_t n_refs, int32_t num_sms) {
    Decomp d;
    d.tpf = (n_pixels + kTile - 1) / kTile;
    d.nt = n_refs * d.tpf;
    d.total = static_cast<int64_t>(d.tpf) * d.nt;
    d.grid = static_cast<int32_t>(d.total < num_sms ? d.total : num_sms);
    // a CTA's range has ceil(total/grid) tiles at most -> spans at most that/nt + 2 m-tiles
    const int64_t per = (d.total + d.grid - 1) / d.grid;
    d.max_segs = static_cast<int32_t>((per + d.nt - 1) / d.nt) + 1;
    return d;
}

VOS_HD int64_t cta_begin(const Decomp& d, int32_t c) { return static_cast<int64_t>(c) * d.total / d.grid; }

// the CTA whose range contains linear tile index x
VOS_HD int32_t cta_of(const Decomp& d, int64_t x) {
    return static_cast<int32_t>(((x + 1) * d.grid + d.total - 1) / d.total - 1);
}

// Iterates the segments of one CTA.
struct SegIter {
    int64_t lin, lin_end;
    int32_t nt, mt, seg;
    VOS_HD SegIter(const Decomp& d, int32_t cta)
        : lin(cta_begin(d, cta)), lin_end(cta_begin(d, cta + 1)), nt(d.nt), mt(0), seg(-1) {
        mt = static_cast<int32_t>(lin / nt);
    }
    // next segment: target tile `m_tile`, reference tiles [n0, n1)
    VOS_HD bool next(int32_t& m_tile, int32_t& n0, int32_t& n1) {
        if (lin >= lin_end) return false;
        m_tile = mt;
        n0 = static_cast<int32_t>(lin - static_cast<int64_t>(mt) * nt);
        const int64_t room = lin_end - lin;
        n1 = (nt - n0 <= room) ? nt : static_cast<int32_t>(n0 + room);
        lin += n1 - n0;
        ++mt;
        ++seg;
        return true;
    }
};

}  // namespace vosd
