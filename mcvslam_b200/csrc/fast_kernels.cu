// FAST-9/16 detection per grid cell for a batch of pyramids (sm_100a).
//
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (ORBextractor.cc:604-633): for every ~35x35 cell the
// reference calls cv::FAST(cell image, iniThFAST, nonmax=true) and, if that returns nothing, again with minThFAST.
// Four kernels (k_fast_score, k_nms_sparse, k_cell_order; the quadtree kernel consumes the result):
//
//  k_fast_score — threshold-free FAST score S of every pixel of every level. S = max over the 16 arcs of 9
//    contiguous circle pixels of min(v - p) resp. min(p - v), minus 1 (OpenCV cornerScore<16>); a pixel is a corner at
//    threshold T iff S >= T. Register-marching stencil, no shared-memory tile: a warp owns a 128-px strip (lane = one aligned
//    4-px word) and marches down the rows keeping the last 7 row words in a register ring. Per row: one coalesced 32-bit
//    load, two shuffles, and a 4-pixels-at-once compass reject built on VABSDIFF4 (a 9-arc always contains two compass
//    pixels that are 90 degrees apart, so a corner needs |v - p| > T on two adjacent compass points). The survivors (about
//    twice the true corners) are pushed to a per-warp queue and scored 32 at a time (one lane per pixel) so the expensive arc
//    min/max network never runs divergent. Corners come in clusters: before the per-lane emission loop, row j's pass bits of a
//    7-row block are handed to the lane 5 j places down, so that no lane has a whole cluster to emit while the others wait.
//
//    There is NO dense score map (round 1 wrote one — 97 % zeros — and the NMS read it back: 0.8 GB of DRAM traffic per
//    128-frame step). Every pixel with S >= threshold (3-7 % of the pixels) is appended to the strip's own list segment (the
//    warp owns the strip, so the fill count is a register: no atomics); the scores on the strip's four BORDER lines (first /
//    last row, first / last column) are also written to a 320-byte edge record per strip, which is all a neighbouring
//    strip's NMS needs to know about this one.
//
//  k_nms_sparse — one warp per strip: rebuilds the strip's score tile (+1 px ring) in shared memory from the strip's own
//    list (scatter) and the facing edge lines of its eight neighbours' records, then runs one lane per listed pixel: strict
//    8-neighbour maximum test, where neighbours outside the pixel's own cell (detection rims of adjacent cells tile the level
//    without overlap) count as 0 — exactly what the per-cell cv::FAST sees. Survivors are appended (atomicAdd, unordered) to
//    their cell's slot array.
//
//  k_cell_order — one warp per cell: applies the ini -> min threshold fallback ("any maximum with S >= ini ? S >= ini :
//    S >= min" — equivalent to re-running FAST at minThFAST) and sorts the survivors by (y, x), i.e. the
//    row-major order cv::FAST emits them in; (y, x) is unique, so the atomics' arrival order never shows.
//
// Candidate order over the level (cell-row-major, then row-major inside the cell) is rebuilt by the quadtree kernel from the
// per-cell counts, so it matches vToDistributeKeys of the reference.
#include <string.h>
#include "engine.h"
#include "tma.cuh"

namespace mcv {

#ifndef MCV_FS_WARPS
#define MCV_FS_WARPS 1
#endif
constexpr int FS_WARPS = MCV_FS_WARPS;
// One warp per CTA and 32 resident CTAs per SM (64 registers, a few spills outside the marching loop): the kernel is bound by
// latency (the scorer's 16 dependent byte loads, the shuffles of the marching loop), and a CTA of several warps keeps its slot
// until its slowest strip is done. B200, 128-frame step, 36-row strips: 4 warps x 5 CTAs (96 registers) 0.717 ms, x 6 0.674,
// x 7 0.667, x 8 0.716; 2 warps x 14 0.670; 1 warp x 28 0.628; 29-row strips: 1 x 24 0.648, 1 x 28 0.626, 1 x 32 0.623.
#ifndef FS_MINB
#define FS_MINB 32
#endif
constexpr int FS_QCAP = 32 + 7 * 128;   // leftover (< 32) + every pixel of a 7-row block

__device__ __noinline__ int fast_score16(const uint8_t* c, int pitch) {
    // Bresenham circle, OpenCV order (fast_score.cpp makeOffsets). Both polarities at once: P[k] = (256 + v - p, 256 + p - v) as
    // packed u16x2 = C + p * 0xFFFF with C = (256 + v) | (256 - v) << 16 (one IMAD per circle pixel; the low half never
    // borrows because 256 + v - p >= 1).
    const unsigned v = c[0];
    const unsigned C = (256u + v) | ((256u - v) << 16);
    // the six other rows by a chain of 64-bit adds (two instructions each; indexing c[k * pitch + d] costs ptxas twice that)
    const ptrdiff_t sp = pitch;
    const uint8_t *d1 = c + sp, *d2 = d1 + sp, *d3 = d2 + sp, *u1 = c - sp, *u2 = u1 - sp, *u3 = u2 - sp;
    unsigned p[16];
    p[0] = d3[0] * 0xFFFFu + C;    p[1] = d3[1] * 0xFFFFu + C;    p[2] = d2[2] * 0xFFFFu + C;     p[3] = d1[3] * 0xFFFFu + C;
    p[4] = c[3] * 0xFFFFu + C;     p[5] = u1[3] * 0xFFFFu + C;    p[6] = u2[2] * 0xFFFFu + C;     p[7] = u3[1] * 0xFFFFu + C;
    p[8] = u3[0] * 0xFFFFu + C;    p[9] = u3[-1] * 0xFFFFu + C;   p[10] = u2[-2] * 0xFFFFu + C;   p[11] = u1[-3] * 0xFFFFu + C;
    p[12] = c[-3] * 0xFFFFu + C;   p[13] = d1[-3] * 0xFFFFu + C;  p[14] = d2[-2] * 0xFFFFu + C;   p[15] = d3[-1] * 0xFFFFu + C;
    // min over every window of 9 consecutive (circular) values as three windows of three (3-input VIMNMX3.S16x2: 16 + 16
    // instructions instead of the 48 of a doubling scheme); then max over the 16 windows
    unsigned m3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m3[k] = __vimin3_s16x2(p[k], p[(k + 1) & 15], p[(k + 2) & 15]);
    unsigned best = 0u;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const unsigned a = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
        const unsigned b = __vimin3_s16x2(m3[k + 1], m3[(k + 4) & 15], m3[(k + 7) & 15]);
        best = __vimax3_s16x2(best, a, b);
    }
    const int A = (int)(best & 0xffffu) - 256, B = (int)(best >> 16) - 256;
    return max(A, B) - 1;
}

// per-byte (a1 > T || a2 > T) for two words of four packed bytes, result in bit 7 of each byte. add = 255 - T; add_lo7 = its low
// seven bits in every byte, HI = its bit 7: a byte exceeds T iff a + add carries out of bit 7, i.e. HI ? (a | s) : (a & s) with
// s = (a & 0x7f) + (add & 0x7f) (no carry across bytes: 127 + 127 < 256). The ORs of the two operands fold into 3-input LOP3s.
template <bool HI>
__device__ __forceinline__ unsigned gt4_or(unsigned a1, unsigned a2, unsigned add_lo7) {
    const unsigned s1 = (a1 & 0x7f7f7f7fu) + add_lo7, s2 = (a2 & 0x7f7f7f7fu) + add_lo7;
    return HI ? (a1 | s1 | a2 | s2) : ((a1 & s1) | (a2 & s2));
}

// Edge record of a strip: the scores on its first / last row (128 bytes each, by strip column) and first / last column
// (FS_ROWS bytes each, by strip row); zero where nothing scored.
constexpr int FE_TOP = 0, FE_BOT = 128, FE_LEFT = 256, FE_RIGHT = 256 + FS_ROWS;
static_assert(FS_EDGE_BYTES >= 256 + 2 * FS_ROWS && FS_EDGE_BYTES % 16 == 0, "edge record layout");

// Queue entry of a candidate: b | lane << 5 | yb << 16 — bit b of lane `lane`'s candidate word of the 7-row block whose last
// row + 1 is yb. Bit b = 8 k + 7 - j stands for row j of the block and column k of a lane's word; row j's bits were handed
// (k_fast_score) to the lane 5 j places down, so the pixel's own lane is (lane + 5 j) & 31.
#ifndef MCV_FS_ROT
#define MCV_FS_ROT 5
#endif
constexpr int FS_ROT = MCV_FS_ROT;

// Scores queued pixels 32 at a time (one lane each) while at least `keep_below` + 1 are queued. Pixels with score >= Tm go,
// packed x | y << 12 | score << 24 (level coordinates), to the strip's list, and to the strip's edge record when they lie on
// one of its border lines (sxy = the strip's first column | first row << 16; its last row follows from the level's y_hi).
// Returns (list fill << 8) | queue fill.
__device__ __noinline__ int drain_queue(const unsigned* q, int qn, int keep_below, const uint8_t* src, uint8_t* __restrict__ edge,
                                        int pitch, int Tm, unsigned* __restrict__ list, int ln, unsigned sxy, int y_hi) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1;
    const int sx0 = (int)(sxy & 0xffffu), sy0 = (int)(sxy >> 16), y_last = min(sy0 + FS_ROWS, y_hi) - 1;
    __syncwarp();
    while (qn > keep_below) {
        const int n = min(qn, 32);
        bool hit = false;
        unsigned packed = 0;
        if (lane < n) {
            const unsigned e = q[qn - n + lane];
            const int jr = (int)(e & 7u);                                  // 7 - j
            const int cx = 4 * (int)(((e >> 5) + FS_ROT * (7 - jr)) & 31u) + (int)((e >> 3) & 3u);   // strip column
            const int x = sx0 + cx, y = (int)(e >> 16) - jr, cy = y - sy0;
            const int sc = fast_score16(src + y * pitch + x, pitch);
            if (sc >= Tm) {
                hit = true; packed = pack_pt(x, y, sc);
                uint8_t* ec = edge + cx;
                uint8_t* er = edge + FE_LEFT + cy;
                if (cy == 0) ec[FE_TOP] = (uint8_t)sc;
                if (y == y_last) ec[FE_BOT] = (uint8_t)sc;
                if (cx == 0) er[0] = (uint8_t)sc;
                if (cx == 127) er[FS_ROWS] = (uint8_t)sc;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) list[ln + __popc(m & lt)] = packed;
        ln += __popc(m);
        qn -= n;
    }
    __syncwarp();
    return (ln << 8) | qn;
}

template <bool HI>
__global__ void __launch_bounds__(32 * FS_WARPS, FS_MINB) k_fast_score(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ edges,
                                                               unsigned* __restrict__ nz_list, int* __restrict__ nz_cnt,
                                                               const __grid_constant__ Plan P, const __grid_constant__ StripTable T, int Tm) {
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
    uint8_t* edge = edges + ((size_t)img * P.n_fast_strips + sid) * FS_EDGE_BYTES;   // this strip's edge record
    if (lane < FS_EDGE_BYTES / 16) reinterpret_cast<uint4*>(edge)[lane] = make_uint4(0u, 0u, 0u, 0u);   // ordered before the scores by drain_queue's __syncwarp
    unsigned* list = nz_list + ((size_t)img * P.n_fast_strips + sid) * FS_SEG;   // this strip's own segment
    int ln = 0;                                             // its fill (warp-uniform)
    const int x_lo = EDGE_THRESHOLD, x_hi = w - EDGE_THRESHOLD, y_hi = h - EDGE_THRESHOLD;   // detection region [19, n-19)
    // Every lane loads a word of every row, with no predicate: lanes past the end of the row read the row's last word instead
    // (their own columns are masked out, and the columns of their left neighbour that would look at it are outside the region
    // too: x + 3 >= pitch > x_hi + 3). Lanes 0 / 31 also fetch the strip's outer neighbour words; lanes 1-30 ride along on lane
    // 0's address (a broadcast).
    const int lx = min(x0, pitch - 4);
    const int ex = lane == 31 ? min(x0 + 4, pitch - 4) : x0 - lane * 4 - 4;
    unsigned colmask = 0;                                   // bit 7 of byte k set iff column x0 + k is inside the region
#pragma unroll
    for (int k = 0; k < 4; ++k) if (x0 + k >= x_lo && x0 + k < x_hi) colmask |= 0x80u << (8 * k);
    unsigned add_lo7 = ((255u - (unsigned)Tm) & 0x7fu) * 0x01010101u;
    asm volatile("" : "+r"(add_lo7));                       // a register, not a multiply rematerialised at every use
    unsigned* q = s_q[warp];
    int qn = 0;                                             // warp-uniform queue fill

    // rows past the level's end feed no output row (row_ok is false for them); they are read unclamped — the next level / image
    // follows in the same buffer and the allocation ends with FS_ROWS + 8 spare rows (enqueue_extract).
    // pw / pe: words requested seven iterations ahead. Iteration r consumes input row y0 - 3 + r into the ring and has input row
    // r - 3 as its centre, so the outer neighbour word is requested for the row that will be the CENTRE seven iterations on.
    unsigned pw[7], pe[7], ring[7];
#pragma unroll
    for (int d = 0; d < 7; ++d) {
        pw[d] = __ldg(reinterpret_cast<const unsigned*>(src + (y0 - 3 + d) * pitch + lx));
        pe[d] = __ldg(reinterpret_cast<const unsigned*>(src + (y0 - 6 + d) * pitch + ex));
        ring[d] = 0u;
    }
    const uint8_t* lp = src + (size_t)(y0 + 4) * pitch + lx;              // input row y0 - 3 + (r + 7) for r = 0
    const uint8_t* lpe = src + (size_t)(y0 + 1) * pitch + ex;             // centre row of iteration r + 7 for r = 0
#pragma unroll 1
    for (int rb = 0; rb < FS_ROWS + 6; rb += 7) {
        // candidates of the block's 7 rows: row j's four pass bits (bit 7 of each byte) shifted right by j never collide.
        // Corners come in clusters, and a lane that kept its own seven rows would have many times the average to emit (the warp
        // waits for the slowest lane): row j's bits go to the lane FS_ROT * j places down instead, which spreads a cluster over
        // seven lanes.
        unsigned acc = 0u;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const int r = rb + j;                           // input row y0 - 3 + r; completes the window of output row y0 + r - 6
            ring[j] = pw[j];
            const unsigned we = pe[j];
            pw[j] = __ldg(reinterpret_cast<const unsigned*>(lp));
            pe[j] = __ldg(reinterpret_cast<const unsigned*>(lpe));
            lp += pitch; lpe += pitch;
            const int y = y0 + r - 6;
            const unsigned v4 = ring[(j + 4) % 7], top = ring[(j + 1) % 7], bot = ring[j];
            unsigned w0 = __shfl_up_sync(0xffffffffu, v4, 1), w2 = __shfl_down_sync(0xffffffffu, v4, 1);
            w0 = lane == 0 ? we : w0;
            w2 = lane == 31 ? we : w2;
            const unsigned lft = __funnelshift_r(w0, v4, 8), rgt = __funnelshift_r(v4, w2, 24);   // pixels x-3.., x+3..
            const bool row_ok = r >= 6 && r < FS_ROWS + 6 && y < y_hi;   // warp-uniform; every row belongs to exactly one strip
                                                                         // (a pixel listed twice would survive NMS twice)
            unsigned pass = gt4_or<HI>(__vabsdiffu4(v4, top), __vabsdiffu4(v4, bot), add_lo7) &
                            gt4_or<HI>(__vabsdiffu4(v4, lft), __vabsdiffu4(v4, rgt), add_lo7) & (row_ok ? colmask : 0u);
            if (j) pass = __shfl_sync(0xffffffffu, pass, lane + FS_ROT * j);   // source lane taken modulo 32
            acc |= pass >> j;
        }
        if (__any_sync(0xffffffffu, acc != 0u)) {
            // one compaction per block: exclusive scan of the lanes' counts, then every lane appends its own entries
            const int cnt = __popc(acc);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t2 = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t2; }
            unsigned* qp = q + qn + incl - cnt;
            qn += __shfl_sync(0xffffffffu, incl, 31);
            const unsigned cb = ((unsigned)(y0 + rb + 1) << 16) | ((unsigned)lane << 5);   // yb = last row of the block + 1
            while (acc) {
                unsigned b;
                asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(acc));   // highest set bit (one FLO; 31 - __clz costs three instructions)
                acc ^= 1u << b;
                *qp++ = cb + b;
            }
            if (qn >= 32) { const int r2 = drain_queue(q, qn, 31, src, edge, pitch, Tm, list, ln, (unsigned)(x0 - lane * 4) | ((unsigned)y0 << 16), y_hi); qn = r2 & 0xff; ln = r2 >> 8; }
        }
    }
    ln = drain_queue(q, qn, 0, src, edge, pitch, Tm, list, ln, (unsigned)(x0 - lane * 4) | ((unsigned)y0 << 16), y_hi) >> 8;
    if (lane == 0) nz_cnt[(size_t)img * P.n_fast_strips + sid] = ln;
}

// ---------------------------------------------------------------------------------------------------------
// sparse NMS: one warp per strip (same strip table as k_fast_score).
// Tried and dropped (bit-exact, B200): per-strip shared-memory tables of each column's / row's cell index and in-cell neighbour flags
// instead of the division and range tests per listed pixel (NMS + cells stage 0.333 -> 0.353 ms: two more dependent
// shared-memory loads per pixel cost more than the ~25 ALU instructions they replace); 1 / 2 / 8 warps per CTA: 0.329 / 0.332 / 0.355.
// ---------------------------------------------------------------------------------------------------------
#ifndef MCV_NMS_WARPS
#define MCV_NMS_WARPS 2
#endif
constexpr int NMS_WARPS = MCV_NMS_WARPS;
constexpr int NMS_TP = 144;                     // tile pitch in bytes: strip column k at byte 4 + k, ring columns at bytes 3 and 132
constexpr int NMS_TR = FS_ROWS + 2;             // tile rows: the strip's rows + one above and below
constexpr int NMS_TILE_BYTES = NMS_TR * NMS_TP;

__global__ void __launch_bounds__(32 * NMS_WARPS) k_nms_sparse(const uint8_t* __restrict__ edges, const unsigned* __restrict__ nz_list,
                                                               const int* __restrict__ nz_cnt, uint32_t* __restrict__ cell_raw,
                                                               int* __restrict__ cell_cnt, const __grid_constant__ Plan P,
                                                               const __grid_constant__ StripTable T) {
    __shared__ __align__(16) uint8_t s_tile[NMS_WARPS][NMS_TILE_BYTES];
    const int img = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sid = blockIdx.x * NMS_WARPS + warp;
    if (sid >= P.n_fast_strips) return;
    const int count = nz_cnt[(size_t)img * P.n_fast_strips + sid];
    if (count == 0) return;
    int level = 0;
    while (level + 1 < P.n_levels && sid >= T.first[level + 1]) ++level;
    const LevelGeom& g = P.lv[level];
    const int t = sid - T.first[level];
    const int nsx = T.strips_x[level], nsy = (T.first[level + 1] - T.first[level]) / nsx;
    const int sxi = t % nsx, syi = t / nsx;
    const int sx0 = BORDER + sxi * 128;                                   // level x of strip column 0 = tile byte 4
    const int ty0 = EDGE_THRESHOLD + syi * FS_ROWS - 1;                   // level y of tile row 0
    uint8_t* tile = s_tile[warp];
    const unsigned* list = nz_list + ((size_t)img * P.n_fast_strips + sid) * FS_SEG;
    // the first (up to) 128 list entries are fetched once and stay in registers for both passes
    unsigned ent[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) ent[u] = lane + 32 * u < count ? __ldg(list + lane + 32 * u) : 0u;
    for (int i = lane; i < NMS_TILE_BYTES / 16; i += 32) reinterpret_cast<uint4*>(tile)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    // ring: the facing border lines of the eight neighbouring strips (their records are zero where nothing scored)
    const uint8_t* E = edges + ((size_t)img * P.n_fast_strips + T.first[level]) * FS_EDGE_BYTES;
    const bool up_ok = syi > 0, dn_ok = syi + 1 < nsy, lf_ok = sxi > 0, rt_ok = sxi + 1 < nsx;
    if (up_ok) reinterpret_cast<unsigned*>(tile + 4)[lane] = __ldg(reinterpret_cast<const unsigned*>(E + (size_t)(t - nsx) * FS_EDGE_BYTES + FE_BOT) + lane);
    if (dn_ok) reinterpret_cast<unsigned*>(tile + (NMS_TR - 1) * NMS_TP + 4)[lane] = __ldg(reinterpret_cast<const unsigned*>(E + (size_t)(t + nsx) * FS_EDGE_BYTES + FE_TOP) + lane);
    for (int r = lane; r < FS_ROWS; r += 32) {
        if (lf_ok) tile[(1 + r) * NMS_TP + 3] = __ldg(E + (size_t)(t - 1) * FS_EDGE_BYTES + FE_RIGHT + r);
        if (rt_ok) tile[(1 + r) * NMS_TP + 132] = __ldg(E + (size_t)(t + 1) * FS_EDGE_BYTES + FE_LEFT + r);
    }
    if (lane == 0 && up_ok && lf_ok) tile[3] = __ldg(E + (size_t)(t - nsx - 1) * FS_EDGE_BYTES + FE_BOT + 127);
    if (lane == 1 && up_ok && rt_ok) tile[132] = __ldg(E + (size_t)(t - nsx + 1) * FS_EDGE_BYTES + FE_BOT);
    if (lane == 2 && dn_ok && lf_ok) tile[(NMS_TR - 1) * NMS_TP + 3] = __ldg(E + (size_t)(t + nsx - 1) * FS_EDGE_BYTES + FE_TOP + 127);
    if (lane == 3 && dn_ok && rt_ok) tile[(NMS_TR - 1) * NMS_TP + 132] = __ldg(E + (size_t)(t + nsx + 1) * FS_EDGE_BYTES + FE_TOP);
    // own pixels: scatter the list
    for (int i0 = lane; i0 < count; i0 += 128) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + 32 * u >= count) break;
            const unsigned e = i0 < 128 ? ent[u] : __ldg(list + i0 + 32 * u);
            tile[(pt_y(e) - ty0) * NMS_TP + (pt_x(e) - sx0 + 4)] = (uint8_t)pt_r(e);
        }
    }
    __syncwarp();
    const uint8_t* tb = tile;
    int* cnts = cell_cnt + (size_t)img * P.cells_per_image + g.cell_base;
    uint32_t* cells = cell_raw + (size_t)img * P.cand_per_image + g.cand_off;
    const float inv_wc = 1.0f / (float)g.w_cell, inv_hc = 1.0f / (float)g.h_cell;
    const int x_hi = g.w - EDGE_THRESHOLD, y_hi = g.h - EDGE_THRESHOLD;
    constexpr int TP = NMS_TP;                                            // tile pitch in bytes
    for (int i0 = lane; i0 < count; i0 += 128) {
        if (i0 >= 128) {
#pragma unroll
            for (int u = 0; u < 4; ++u) ent[u] = i0 + 32 * u < count ? __ldg(list + i0 + 32 * u) : 0u;   // four list loads in flight
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + 32 * u >= count) break;
            const unsigned e = ent[u];
            const int x = pt_x(e), y = pt_y(e), s = pt_r(e);
            // cell of the pixel: detection region of cell (cy, cx) = rows [19 + cy*h_cell, ...), cols [19 + cx*w_cell, ...)
            const int ux = x - EDGE_THRESHOLD, uy = y - EDGE_THRESHOLD;
            const int cx = (int)(((float)ux + 0.5f) * inv_wc), cy = (int)(((float)uy + 0.5f) * inv_hc);
            const int rx = ux - cx * g.w_cell, ry = uy - cy * g.h_cell;
            const bool lf = rx > 0, rt = rx < g.w_cell - 1 && x + 1 < x_hi, up = ry > 0, dn = ry < g.h_cell - 1 && y + 1 < y_hi;
            const uint8_t* c = tb + (y - ty0) * TP + (x - sx0 + 4);
            // neighbours outside the cell's detection region count as 0
            const int l0 = lf ? 1 : 0, r0 = rt ? 1 : 0;
            int m = max(lf ? (int)c[-1] : 0, rt ? (int)c[1] : 0);
            if (up) m = max(m, max((int)c[-TP], max((int)c[-TP - l0], (int)c[-TP + r0])));
            if (dn) m = max(m, max((int)c[TP], max((int)c[TP - l0], (int)c[TP + r0])));
            if (s > m) {
                const int cell = cy * g.n_cols + cx;
                const int slot = atomicAdd(&cnts[cell], 1);
                cells[(size_t)cell * g.cell_cap + slot] = pack_pt(x - BORDER, y - BORDER, s);   // reference coordinates: level - BORDER
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// per-cell threshold fallback + (y, x) ordering: one warp per cell. cell_raw -> cell_pts, cell_cnt updated in place.
// ---------------------------------------------------------------------------------------------------------
#ifndef MCV_ORD_WARPS
#define MCV_ORD_WARPS 4
#endif
constexpr int ORD_WARPS = MCV_ORD_WARPS;

__global__ void __launch_bounds__(32 * ORD_WARPS) k_cell_order(const uint32_t* __restrict__ cell_raw, uint32_t* __restrict__ cell_pts,
                                                               int* __restrict__ cell_cnt, const __grid_constant__ Plan P, int* __restrict__ fallback) {
    const int img = blockIdx.y, lane = threadIdx.x & 31;
    int cell = blockIdx.x * ORD_WARPS + (threadIdx.x >> 5);
    if (cell >= P.cells_per_image) return;
    int level = 0;
    while (level + 1 < P.n_levels && cell >= P.lv[level + 1].cell_base) ++level;
    const LevelGeom& g = P.lv[level];
    cell -= g.cell_base;
    int* cnt_p = cell_cnt + (size_t)img * P.cells_per_image + g.cell_base + cell;
    const int n = *cnt_p;
    if (n == 0) {
        // nothing at iniThFAST: the reference re-runs cv::FAST on this cell with minThFAST (ORBextractor.cc:619-623) -> queue the
        // cell for k_cell_fallback. fallback[0] = count, fallback[1 + i] = image * cells_per_image + cell index
        if (fallback && lane == 0) fallback[1 + atomicAdd(&fallback[0], 1)] = img * P.cells_per_image + g.cell_base + cell;
        return;
    }
    const size_t off = (size_t)img * P.cand_per_image + g.cand_off + (size_t)cell * g.cell_cap;
    const uint32_t* in = cell_raw + off;
    uint32_t* out = cell_pts + off;
    // sort key: y (12 bits) | x (12 bits); the packed point already has y above x, so the low 24 bits ARE the key
    if (n <= 32) {
        const uint32_t p = lane < n ? in[lane] : 0u;
        const bool ini = lane < n && pt_r(p) >= P.ini_th;
        const int th = __any_sync(0xffffffffu, ini) ? P.ini_th : P.min_th;
        const bool keep = lane < n && pt_r(p) >= th;
        // bitonic sort over the warp of the point rotated so that the key (y, x) leads and the response trails: keys are unique,
        // dropped entries (all ones) sink to the end; 15 compare-exchange steps instead of a 32-step rank count
        unsigned v = keep ? __funnelshift_l(p, p, 8) : 0xffffffffu;
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                const unsigned o = __shfl_xor_sync(0xffffffffu, v, j);
                const bool up = ((lane & k) == 0) == ((lane & j) == 0);   // this lane keeps the smaller of the pair
                v = up ? min(v, o) : max(v, o);
            }
        }
        const int kept = __popc(__ballot_sync(0xffffffffu, keep));
        if (lane < kept) out[lane] = __funnelshift_r(v, v, 8);
        if (lane == 0) *cnt_p = kept;
        return;
    }
    // crowded cell: rank by counting over the whole list
    bool any_ini = false;
    for (int i = lane; i < n; i += 32) any_ini |= pt_r(in[i]) >= P.ini_th;
    const int th = __any_sync(0xffffffffu, any_ini) ? P.ini_th : P.min_th;
    int kept = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const uint32_t p = i < n ? in[i] : 0u;
        const bool keep = i < n && pt_r(p) >= th;
        if (keep) {
            const unsigned key = p & 0xffffffu;
            int rank = 0;
            for (int j = 0; j < n; ++j) { const uint32_t o = in[j]; rank += pt_r(o) >= th && (o & 0xffffffu) < key; }
            out[rank] = p;
        }
        kept += __popc(__ballot_sync(0xffffffffu, keep));
    }
    if (lane == 0) *cnt_p = kept;
}

// ---------------------------------------------------------------------------------------------------------
// minThFAST fallback for the cells that came up empty at iniThFAST (ORBextractor.cc:619-623): one warp per queued cell
// recomputes the cell's score map at minThFAST into a shared-memory tile (ring of zeros = "neighbours outside the cell's own
// cv::FAST image do not exist"), runs the strict 8-neighbour NMS on it and writes the survivors in row-major order — the
// order cv::FAST emits them in — straight into cell_pts. The dense pass only ever scores at iniThFAST: on textured images
// (every cell has an iniThFAST corner) this queue is almost empty and four fifths of the exact scoring work disappears.
// ---------------------------------------------------------------------------------------------------------
constexpr int FB_WARPS = 4;
constexpr int FB_MAX_CELL = 78;                  // w_cell, h_cell <= 78: tile of 80 x 80 bytes per warp
constexpr int FB_TP = FB_MAX_CELL + 2;

__global__ void __launch_bounds__(32 * FB_WARPS) k_cell_fallback(const uint8_t* __restrict__ pyr, const int* __restrict__ fallback,
                                                                uint32_t* __restrict__ cell_pts, int* __restrict__ cell_cnt,
                                                                const __grid_constant__ Plan P) {
    __shared__ uint8_t s_tile[FB_WARPS][FB_TP * FB_TP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_items = fallback[0];
    const unsigned lt = (1u << lane) - 1;
    uint8_t* tile = s_tile[warp];
    const int Tm = P.min_th;
    for (int item = blockIdx.x * FB_WARPS + warp; item < n_items; item += gridDim.x * FB_WARPS) {
        const int code = fallback[1 + item];
        const int img = code / P.cells_per_image;
        int cell = code - img * P.cells_per_image;
        int level = 0;
        while (level + 1 < P.n_levels && cell >= P.lv[level + 1].cell_base) ++level;
        const LevelGeom& g = P.lv[level];
        cell -= g.cell_base;
        const int cy = cell / g.n_cols, cx = cell - cy * g.n_cols;
        // the cell's own detection region (level coordinates), clipped to the level's
        const int x0 = EDGE_THRESHOLD + cx * g.w_cell, y0 = EDGE_THRESHOLD + cy * g.h_cell;
        const int cw = min(g.w_cell, g.w - EDGE_THRESHOLD - x0), ch = min(g.h_cell, g.h - EDGE_THRESHOLD - y0);
        const int pitch = g.pitch;
        const uint8_t* src = pyr + (size_t)img * P.pyr_bytes + g.img_off;
        __syncwarp();
        for (int i = lane; i < FB_TP * (ch + 2); i += 32) tile[i] = 0;
        __syncwarp();
        if (cw > 0 && ch > 0) {
            const int n_px = cw * ch;
            // four pixels per lane and round: their 20 byte loads are in flight together (one warp walks a whole cell, so the
            // kernel lasts as long as one cell's chain of global round trips)
            constexpr int FB_U = 4;
            for (int i0 = 0; i0 < n_px; i0 += 32 * FB_U) {
                const uint8_t* cp[FB_U];
                int v[FB_U], up[FB_U], dn[FB_U], lf[FB_U], rt[FB_U], ryx[FB_U];
#pragma unroll
                for (int u = 0; u < FB_U; ++u) {
                    const int i = min(i0 + 32 * u + lane, n_px - 1);                     // clamped: the tail repeats the last pixel
                    const int ry = i / cw, rx = i - ry * cw;
                    const uint8_t* c = src + (size_t)(y0 + ry) * pitch + (x0 + rx);
                    cp[u] = c; ryx[u] = (ry + 1) * FB_TP + rx + 1;
                    v[u] = c[0]; up[u] = c[-3 * pitch]; dn[u] = c[3 * pitch]; lf[u] = c[-3]; rt[u] = c[3];
                }
#pragma unroll
                for (int u = 0; u < FB_U; ++u) {
                    if (i0 + 32 * u + lane < n_px) {
                        // same necessary condition as the dense pass: two compass points 90 degrees apart differ by more than T
                        const bool vert = abs(up[u] - v[u]) > Tm || abs(dn[u] - v[u]) > Tm;
                        const bool horz = abs(lf[u] - v[u]) > Tm || abs(rt[u] - v[u]) > Tm;
                        if (vert && horz) {
                            const int sc = fast_score16(cp[u], pitch);
                            if (sc >= Tm) tile[ryx[u]] = (uint8_t)sc;
                        }
                    }
                }
            }
            __syncwarp();
            uint32_t* out = cell_pts + (size_t)img * P.cand_per_image + g.cand_off + (size_t)cell * g.cell_cap;
            int kept = 0;
            for (int i0 = 0; i0 < n_px; i0 += 32) {
                const int i = i0 + lane;
                bool keep = false;
                int ry = 0, rx = 0, sc = 0;
                if (i < n_px) {
                    ry = i / cw; rx = i - ry * cw;
                    const uint8_t* t = tile + (ry + 1) * FB_TP + rx + 1;
                    sc = t[0];
                    if (sc) {
                        const int m = max(max(max((int)t[-1], (int)t[1]), max((int)t[-FB_TP], (int)t[FB_TP])),
                                          max(max((int)t[-FB_TP - 1], (int)t[-FB_TP + 1]), max((int)t[FB_TP - 1], (int)t[FB_TP + 1])));
                        keep = sc > m;
                    }
                }
                const unsigned mk = __ballot_sync(0xffffffffu, keep);
                if (keep) out[kept + __popc(mk & lt)] = pack_pt(x0 + rx - BORDER, y0 + ry - BORDER, sc);
                kept += __popc(mk);
            }
            if (lane == 0) cell_cnt[(size_t)img * P.cells_per_image + g.cell_base + cell] = kept;
        }
    }
}

// strip table shared by k_fast_score and k_nms_sparse; its total is Plan::n_fast_strips (fast_strip_table is also what
// build_plan uses to size the list segments)
int fast_strip_table(const Plan& P, StripTable& T) {
    int n = 0;
    for (int l = 0; l < P.n_levels; ++l) {
        T.first[l] = n;
        const int rw = P.lv[l].w - EDGE_THRESHOLD - BORDER, rh = P.lv[l].h - 2 * EDGE_THRESHOLD;
        T.strips_x[l] = std::max(1, (rw + 127) / 128);
        n += T.strips_x[l] * std::max(0, (rh + FS_ROWS - 1) / FS_ROWS);
    }
    for (int l = P.n_levels; l <= MAX_LEVELS; ++l) T.first[l] = n;
    return n;
}

int launch_fast_cells(const Plan& P, const uint8_t* d_pyr, uint8_t* d_edges, unsigned* d_nz_list, int* d_nz_cnt, uint32_t* d_cell_raw,
                      uint32_t* d_cell_pts, int* d_cell_cnt, int* d_fallback, int n_images, cudaStream_t s, cudaEvent_t after_score) {
    StripTable T{};
    const int n = fast_strip_table(P, T);
    // two-pass thresholds: dense scoring at iniThFAST, minThFAST only for the cells that stay empty (k_cell_fallback). Needs the
    // cell to fit the fallback tile; otherwise (and when both thresholds agree) the dense pass scores at minThFAST as before
    // and k_cell_order applies the "any corner >= ini ? ini : min" rule on the survivors.
    const bool two_pass = d_fallback && P.ini_th > P.min_th && P.max_cell_w <= FB_MAX_CELL && P.max_cell_h <= FB_MAX_CELL;
    cudaMemsetAsync(d_cell_cnt, 0, (size_t)n_images * P.cells_per_image * sizeof(int), s);
    if (two_pass) cudaMemsetAsync(d_fallback, 0, sizeof(int), s);
    if (n > 0) {
        const int Tm = two_pass ? P.ini_th : P.min_th;
        const dim3 grid((n + FS_WARPS - 1) / FS_WARPS, n_images);
        if ((255 - Tm) & 0x80) k_fast_score<true><<<grid, 32 * FS_WARPS, 0, s>>>(d_pyr, d_edges, d_nz_list, d_nz_cnt, P, T, Tm);
        else k_fast_score<false><<<grid, 32 * FS_WARPS, 0, s>>>(d_pyr, d_edges, d_nz_list, d_nz_cnt, P, T, Tm);
    }
    if (after_score) cudaEventRecord(after_score, s);   // stage boundary for mcv_rig_stage_ms
    if (n > 0)
        k_nms_sparse<<<dim3((n + NMS_WARPS - 1) / NMS_WARPS, n_images), 32 * NMS_WARPS, 0, s>>>(d_edges, d_nz_list, d_nz_cnt, d_cell_raw, d_cell_cnt, P, T);
    k_cell_order<<<dim3((P.cells_per_image + ORD_WARPS - 1) / ORD_WARPS, n_images), 32 * ORD_WARPS, 0, s>>>(d_cell_raw, d_cell_pts, d_cell_cnt, P,
                                                                                                          two_pass ? d_fallback : nullptr);
    if (!two_pass) return 3;
    const int max_items = P.cells_per_image * n_images;
    const int ctas = std::max(1, std::min((max_items + FB_WARPS - 1) / FB_WARPS, NUM_SMS * 4));
    k_cell_fallback<<<ctas, 32 * FB_WARPS, 0, s>>>(d_pyr, d_fallback, d_cell_pts, d_cell_cnt, P);
    return 4;
}

}  // namespace mcv
