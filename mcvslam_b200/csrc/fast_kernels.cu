// FAST-9/16 detection per grid cell for a batch of pyramids (sm_100a).
//
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (ORBextractor.cc:604-633): for every ~35x35 cell the
// reference calls cv::FAST(cell image, iniThFAST, nonmax=true) and, if that returns nothing, again with minThFAST.
// Two kernels:
//
//  k_fast_score — threshold-free FAST score map S of every level (u8, 0 where S < minThFAST). S = max over the 16 arcs of 9
//    contiguous circle pixels of min(v - p) resp. min(p - v), minus 1 (OpenCV cornerScore<16>); a pixel is a corner at
//    threshold T iff S >= T. Register-marching stencil, no shared-memory tile: a warp owns a 128-px strip (lane = one aligned
//    4-px word) and marches down the rows keeping the last 7 row words in a register ring. Per row: one coalesced 32-bit
//    load, two shuffles, and a 4-pixels-at-once compass reject built on VABSDIFF4 (a 9-arc always contains two compass
//    pixels that are 90 degrees apart, so a corner needs |v - p| > T on two adjacent compass points). The few survivors are
//    pushed to a per-warp queue and scored 32 at a time (one lane per pixel) so the expensive arc min/max network never runs
//    divergent.
//
//  k_cell_nms — one warp per cell: stages the cell's detection region of S in shared memory, keeps strict 8-neighbour local
//    maxima (pixels outside the cell's own detection rim count as 0 — exactly what the per-cell cv::FAST sees), applies the
//    ini -> min threshold fallback ("any maximum with S >= ini ? S >= ini : S >= min" — equivalent to re-running FAST at
//    minThFAST), and ballot-compacts the survivors in row-major order into the cell's candidate slots.
//
// Candidate order over the level (cell-row-major, then row-major inside the cell) is rebuilt by the quadtree kernel from the
// per-cell counts, so it matches vToDistributeKeys of the reference.
#include "engine.h"

namespace mcv {

constexpr int FS_ROWS = 32;   // output rows per warp
constexpr int FS_WARPS = 4;
constexpr int FS_QCAP = 32 + 128;

__device__ __noinline__ int fast_score16(const uint8_t* c, int pitch) {
    // Bresenham circle, OpenCV order (fast_score.cpp makeOffsets). Both polarities at once: P[k] = (256 + v - p, 256 + p - v) as
    // packed u16x2 = C + p * 0xFFFF with C = (256 + v) | (256 - v) << 16 (one IMAD per circle pixel; the low half never
    // borrows because 256 + v - p >= 1).
    const unsigned v = c[0];
    const unsigned C = (256u + v) | ((256u - v) << 16);
    unsigned p[16];
    p[0] = c[3 * pitch] * 0xFFFFu + C;          p[1] = c[3 * pitch + 1] * 0xFFFFu + C;   p[2] = c[2 * pitch + 2] * 0xFFFFu + C;
    p[3] = c[pitch + 3] * 0xFFFFu + C;          p[4] = c[3] * 0xFFFFu + C;               p[5] = c[-pitch + 3] * 0xFFFFu + C;
    p[6] = c[-2 * pitch + 2] * 0xFFFFu + C;     p[7] = c[-3 * pitch + 1] * 0xFFFFu + C;  p[8] = c[-3 * pitch] * 0xFFFFu + C;
    p[9] = c[-3 * pitch - 1] * 0xFFFFu + C;     p[10] = c[-2 * pitch - 2] * 0xFFFFu + C; p[11] = c[-pitch - 3] * 0xFFFFu + C;
    p[12] = c[-3] * 0xFFFFu + C;                p[13] = c[pitch - 3] * 0xFFFFu + C;      p[14] = c[2 * pitch - 2] * 0xFFFFu + C;
    p[15] = c[3 * pitch - 1] * 0xFFFFu + C;
    // min over every window of 9 consecutive (circular) values by doubling (2, 4, then 4 + 4 + 1 with a 3-input min),
    // VIMNMX.S16x2 / VIMNMX3.S16x2; then max over the 16 windows
    unsigned m2[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m2[k] = __vmins2(p[k], p[(k + 1) & 15]);
    unsigned m4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m4[k] = __vmins2(m2[k], m2[(k + 2) & 15]);
    unsigned best = 0u;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const unsigned a = __vimin3_s16x2(m4[k], m4[(k + 4) & 15], p[(k + 8) & 15]);
        const unsigned b = __vimin3_s16x2(m4[k + 1], m4[(k + 5) & 15], p[(k + 9) & 15]);
        best = __vimax3_s16x2(best, a, b);
    }
    const int A = (int)(best & 0xffffu) - 256, B = (int)(best >> 16) - 256;
    return max(A, B) - 1;
}

// per-byte (a > T) for four packed bytes, result in bit 7 of each byte. add = 255 - T.
__device__ __forceinline__ unsigned gt4(unsigned a, unsigned add_lo7, bool add_hi) {
    const unsigned s = (a & 0x7f7f7f7fu) + add_lo7;     // no carry across bytes: 127 + 127 < 256
    return add_hi ? (a | s) : (a & s);                  // carry out of bit 7 of a + add, add's bit 7 being a constant
}

// Scores queued pixels 32 at a time (one lane each) while at least `keep_below` + 1 are queued; returns the new fill.
__device__ __noinline__ int drain_queue(const unsigned* q, int qn, int keep_below, const uint8_t* src, uint8_t* dst, int pitch, int Tm) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    while (qn > keep_below) {
        const int n = min(qn, 32);
        if (lane < n) {
            const unsigned e = q[qn - n + lane];
            const int x = (int)(e & 0xffffu), y = (int)(e >> 16);
            const int sc = fast_score16(src + y * pitch + x, pitch);
            if (sc >= Tm) dst[y * pitch + x] = (uint8_t)sc;
        }
        qn -= n;
    }
    __syncwarp();
    return qn;
}

__global__ void __launch_bounds__(32 * FS_WARPS) k_fast_score(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ score,
                                                               const __grid_constant__ Plan P, const __grid_constant__ StripTable T) {
    __shared__ unsigned s_q[FS_WARPS][FS_QCAP];
    const int img = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sid = blockIdx.x * FS_WARPS + warp;
    if (sid >= T.first[P.n_levels]) return;
    int level = 0;
    while (level + 1 < P.n_levels && sid >= T.first[level + 1]) ++level;
    const LevelGeom& g = P.lv[level];
    const int t = sid - T.first[level];
    const int x0 = BORDER + (t % T.strips_x[level]) * 128 + lane * 4;   // BORDER is a multiple of 4: aligned words
    const int y0 = EDGE_THRESHOLD + (t / T.strips_x[level]) * FS_ROWS;
    const int w = g.w, h = g.h, pitch = g.pitch;
    const uint8_t* src = pyr + (size_t)img * P.pyr_bytes + g.img_off;
    uint8_t* dst = score + (size_t)img * P.pyr_bytes + g.img_off;
    const int x_lo = EDGE_THRESHOLD, x_hi = w - EDGE_THRESHOLD, y_hi = h - EDGE_THRESHOLD;   // detection region [19, n-19)
    const bool in_row = x0 < pitch;                         // word exists in memory
    const int ex = lane == 0 ? x0 - 4 : x0 + 4;             // lanes 0 / 31 fetch the strip's outer neighbour words
    const bool edge = lane == 0 || (lane == 31 && x0 + 4 < pitch);
    unsigned colmask = 0;                                   // bit 7 of byte k set iff column x0 + k is inside the region
#pragma unroll
    for (int k = 0; k < 4; ++k) if (x0 + k >= x_lo && x0 + k < x_hi) colmask |= 0x80u << (8 * k);
    const int Tm = P.min_th;
    const unsigned add = 255u - (unsigned)Tm, add_lo7 = (add & 0x7fu) * 0x01010101u;
    const bool add_hi = (add & 0x80u) != 0;
    const unsigned lt = (1u << lane) - 1;
    unsigned* q = s_q[warp];
    int qn = 0;                                             // warp-uniform queue fill

    auto row_ptr = [&](int r) { return src + min(y0 - 3 + r, h - 1) * pitch; };   // rows past the image feed nothing
    unsigned pw[7], pe[7], ring[7], ering[7];
#pragma unroll
    for (int d = 0; d < 7; ++d) {
        const uint8_t* row = row_ptr(d);
        pw[d] = in_row ? __ldg(reinterpret_cast<const unsigned*>(row + x0)) : 0u;
        pe[d] = edge ? __ldg(reinterpret_cast<const unsigned*>(row + ex)) : 0u;
        ring[d] = 0u; ering[d] = 0u;
    }
#pragma unroll 1
    for (int rb = 0; rb < FS_ROWS + 6 + 1; rb += 7) {
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const int r = rb + j;                           // input row y0 - 3 + r; completes the window of output row y0 + r - 6
            ring[j] = pw[j]; ering[j] = pe[j];
            {
                const uint8_t* row = row_ptr(r + 7);
                pw[j] = in_row ? __ldg(reinterpret_cast<const unsigned*>(row + x0)) : 0u;
                pe[j] = edge ? __ldg(reinterpret_cast<const unsigned*>(row + ex)) : 0u;
            }
            const int y = y0 + r - 6;
            const unsigned v4 = ring[(j + 4) % 7], top = ring[(j + 1) % 7], bot = ring[j], we = ering[(j + 4) % 7];
            unsigned w0 = __shfl_up_sync(0xffffffffu, v4, 1), w2 = __shfl_down_sync(0xffffffffu, v4, 1);
            w0 = lane == 0 ? we : w0;
            w2 = lane == 31 ? we : w2;
            const unsigned lft = __funnelshift_r(w0, v4, 8), rgt = __funnelshift_r(v4, w2, 24);   // pixels x-3.., x+3..
            const unsigned mt = gt4(__vabsdiffu4(v4, top), add_lo7, add_hi), mb = gt4(__vabsdiffu4(v4, bot), add_lo7, add_hi);
            const unsigned ml = gt4(__vabsdiffu4(v4, lft), add_lo7, add_hi), mr = gt4(__vabsdiffu4(v4, rgt), add_lo7, add_hi);
            const bool row_ok = r >= 6 && y < y_hi;         // warp-uniform
            const unsigned pass = row_ok ? ((mt | mb) & (ml | mr) & colmask) : 0u;
            if (row_ok && colmask) *reinterpret_cast<unsigned*>(dst + y * pitch + x0) = 0u;   // zero first; scores land later
            if (__any_sync(0xffffffffu, pass != 0)) {
                // queue position of (lane, k): entries of byte k of all lanes, then byte k + 1, ...
                int base = qn;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool pk = (pass >> (8 * k + 7)) & 1u;
                    const unsigned m = __ballot_sync(0xffffffffu, pk);
                    if (pk) q[base + __popc(m & lt)] = (unsigned)(x0 + k) | ((unsigned)y << 16);
                    base += __popc(m);
                }
                qn = base;
                if (qn >= 32) qn = drain_queue(q, qn, 31, src, dst, pitch, Tm);
            }
        }
    }
    drain_queue(q, qn, 0, src, dst, pitch, Tm);
}

// ---------------------------------------------------------------------------------------------------------
// per-cell NMS + threshold fallback + ordered compaction: one warp per cell.
// The cell's detection region is staged as whole aligned 32-bit words (same alignment phase as global memory, so the copy
// is LDG.32 -> mask -> STS.32) inside a zero frame; columns outside the region are masked to 0 on the way in, which is
// exactly "pixels outside the cell's detection rim count as 0". The scan then walks the words in (row, word) order =
// row-major pixel order, skips all-zero words with one compare, and tests the rare non-zero bytes.
// ---------------------------------------------------------------------------------------------------------
constexpr int NMS_WARPS = 8;

__global__ void __launch_bounds__(32 * NMS_WARPS) k_cell_nms(const uint8_t* __restrict__ score, uint32_t* __restrict__ cell_pts,
                                                             int* __restrict__ cell_cnt, const __grid_constant__ Plan P,
                                                             int tile_bytes, int list_cap) {
    extern __shared__ __align__(16) uint8_t nms_smem[];
    const int img = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int cell = blockIdx.x * NMS_WARPS + warp;
    if (cell >= P.cells_per_image) return;
    int level = 0;
    while (level + 1 < P.n_levels && cell >= P.lv[level + 1].cell_base) ++level;
    const LevelGeom& g = P.lv[level];
    cell -= g.cell_base;
    const int ci = cell / g.n_cols, cj = cell - ci * g.n_cols;
    int* out_cnt = cell_cnt + (size_t)img * P.cells_per_image + g.cell_base + cell;
    uint32_t* out_pts = cell_pts + (size_t)img * P.cand_per_image + g.cand_off + (size_t)cell * g.cell_cap;
    // cell image bounds — ORBextractor.cc:588-615; detection region = rows/cols [3, n-3) of the cell image
    const int max_bx = g.w - BORDER, max_by = g.h - BORDER;
    const int ini_y = BORDER + ci * g.h_cell, ini_x = BORDER + cj * g.w_cell;
    const int max_y = min(ini_y + g.h_cell + 6, max_by), max_x = min(ini_x + g.w_cell + 6, max_bx);
    const int dx0 = ini_x + 3, dy0 = ini_y + 3;              // first detection column / row (level coordinates)
    const int dw = max_x - ini_x - 6, dh = max_y - ini_y - 6;
    if (ini_y >= max_by - 3 || ini_x >= max_bx - 6 || dw <= 0 || dh <= 0) {
        if (lane == 0) *out_cnt = 0;
        return;
    }
    const int ax0 = dx0 & ~3;                                // aligned start column
    const int wpr = ((dx0 + dw + 3) >> 2) - (ax0 >> 2);      // words per row
    const int tp = 4 * (wpr + 2);                            // tile pitch: one zero word left and right
    uint8_t* tile = nms_smem + (size_t)warp * (tile_bytes + 4 * list_cap);
    uint32_t* list = reinterpret_cast<uint32_t*>(tile + tile_bytes);
    uint32_t* tw = reinterpret_cast<uint32_t*>(tile);
    const int tpw = wpr + 2, n_words = dh * wpr;
    // zero frame: rows -1 and dh, and the side words of every row
    for (int i = lane; i < tpw; i += 32) { tw[i] = 0u; tw[(dh + 1) * tpw + i] = 0u; }
    for (int y = lane; y < dh; y += 32) { tw[(y + 1) * tpw] = 0u; tw[(y + 1) * tpw + wpr + 1] = 0u; }
    // byte masks of the first / last word of a row
    const unsigned m_first = 0xffffffffu << (8 * (dx0 - ax0));
    const int tail = (dx0 + dw) & 3;
    const unsigned m_last = tail ? (0xffffffffu >> (8 * (4 - tail))) : 0xffffffffu;
    const uint8_t* src = score + (size_t)img * P.pyr_bytes + g.img_off + dy0 * g.pitch + ax0;
    const float inv_wpr = 1.0f / (float)wpr;
    for (int i = lane; i < n_words; i += 32) {
        const int y = (int)(((float)i + 0.5f) * inv_wpr), wi = i - y * wpr;
        unsigned v = __ldg(reinterpret_cast<const unsigned*>(src + y * g.pitch) + wi);
        if (wi == 0) v &= m_first;
        if (wi == wpr - 1) v &= m_last;
        tw[(y + 1) * tpw + wi + 1] = v;
    }
    __syncwarp();
    // pass A: strict 8-neighbour local maxima, in row-major order -> list (x | y << 8 | s << 24), x relative to ax0
    const unsigned lt = (1u << lane) - 1;
    int n_keep = 0;
    bool any_ini = false;
    for (int base = 0; base < n_words; base += 32) {
        const int i = base + lane;
        unsigned word = 0u;
        int y = 0, wi = 0;
        if (i < n_words) {
            y = (int)(((float)i + 0.5f) * inv_wpr); wi = i - y * wpr;
            word = tw[(y + 1) * tpw + wi + 1];
        }
        unsigned keep = 0;
        if (word) {
            const uint8_t* c0 = tile + (y + 1) * tp + 4 * (wi + 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int s = (word >> (8 * k)) & 0xff;
                if (s) {
                    const uint8_t* c = c0 + k;
                    if (s > c[-1] && s > c[1] && s > c[-tp - 1] && s > c[-tp] && s > c[-tp + 1] && s > c[tp - 1] && s > c[tp] && s > c[tp + 1]) {
                        keep |= 1u << k;
                        any_ini |= s >= P.ini_th;
                    }
                }
            }
        }
        if (__any_sync(0xffffffffu, keep != 0)) {
            const int mine = __popc(keep);
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            int pos = n_keep + incl - mine;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if ((keep >> k) & 1u) list[pos++] = (unsigned)(4 * wi + k) | ((unsigned)y << 8) | (((word >> (8 * k)) & 0xffu) << 24);
            n_keep += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    const int th = __any_sync(0xffffffffu, any_ini) ? P.ini_th : P.min_th;
    __syncwarp();
    // pass B: threshold + ordered write; reference coordinates = level position - BORDER
    int total = 0;
    const int ox = ax0 - BORDER, oy = dy0 - BORDER;
    for (int b = 0; b < n_keep; b += 32) {
        const int i = b + lane;
        const unsigned e = i < n_keep ? list[i] : 0u;
        const bool k = i < n_keep && (int)(e >> 24) >= th;
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (k) out_pts[total + __popc(m & lt)] = pack_pt((int)(e & 0xffu) + ox, (int)((e >> 8) & 0xffu) + oy, (int)(e >> 24));
        total += __popc(m);
    }
    if (lane == 0) *out_cnt = total;
}

int launch_fast_cells(const Plan& P, const uint8_t* d_pyr, uint8_t* d_score, uint32_t* d_cell_pts, int* d_cell_cnt, int n_images,
                      cudaStream_t s) {
    StripTable T{};
    int n = 0;
    for (int l = 0; l < P.n_levels; ++l) {
        T.first[l] = n;
        const int rw = P.lv[l].w - EDGE_THRESHOLD - BORDER, rh = P.lv[l].h - 2 * EDGE_THRESHOLD;
        T.strips_x[l] = std::max(1, (rw + 127) / 128);
        n += T.strips_x[l] * std::max(0, (rh + FS_ROWS - 1) / FS_ROWS);
    }
    for (int l = P.n_levels; l <= MAX_LEVELS; ++l) T.first[l] = n;
    if (n > 0) k_fast_score<<<dim3((n + FS_WARPS - 1) / FS_WARPS, n_images), 32 * FS_WARPS, 0, s>>>(d_pyr, d_score, P, T);
    // NMS tile: (h_cell + 2) rows of (words per row + 2) words; list: cell_cap entries
    int tile_bytes = 0, list_cap = 0;
    for (int l = 0; l < P.n_levels; ++l) {
        const LevelGeom& g = P.lv[l];
        tile_bytes = std::max(tile_bytes, ((g.h_cell + 2) * 4 * ((g.w_cell + 3) / 4 + 3) + 15) & ~15);
        list_cap = std::max(list_cap, g.cell_cap);
    }
    const size_t smem = (size_t)NMS_WARPS * (tile_bytes + 4 * (size_t)list_cap);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_cell_nms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_cell_nms<<<dim3((P.cells_per_image + NMS_WARPS - 1) / NMS_WARPS, n_images), 32 * NMS_WARPS, smem, s>>>(d_score, d_cell_pts, d_cell_cnt, P,
                                                                                                          tile_bytes, list_cap);
    return 2;
}

}  // namespace mcv
