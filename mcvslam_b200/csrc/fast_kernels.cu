// FAST-9/16 detection per grid cell for a batch of pyramids (sm_100a).
//
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (ORBextractor.cc:604-633): for every ~35x35 cell it calls
// cv::FAST(cell image, iniThFAST, nonmax=true) and, if that returns nothing, again with minThFAST. One CTA owns one cell of
// one level of one image:
//   1. stage the (wCell+6) x (hCell+6) cell image in shared memory;
//   2. threshold-free FAST score S = max over 9-arcs of min(v-p) / min(p-v), minus 1 (OpenCV cornerScore<16>), kept where
//      S >= minThFAST; a compass-point test rejects most pixels first and the survivors are scored from a compacted work list;
//   3. non-maximum suppression inside the cell only (pixels outside the cell's detection rim count as 0, exactly what the
//      per-cell cv::FAST sees); corners at threshold T are {S >= T}, so the ini->min fallback is "keep S >= ini if any
//      strict local maximum has S >= ini, else keep all";
//   4. warp-ballot compaction in row-major order into the cell's candidate slots.
// Candidate order over the level (cell-row-major, then row-major inside the cell) is rebuilt by the quadtree kernel from the
// per-cell counts, so it matches vToDistributeKeys of the reference.
#include "engine.h"

namespace mcv {

constexpr int FAST_THREADS = 128;

__device__ __forceinline__ int fast_score16(const uint8_t* c, int pitch) {
    // Bresenham circle, OpenCV order (fast_score.cpp makeOffsets)
    const int v = c[0];
    int d[16];
    d[0] = v - c[3 * pitch];          d[1] = v - c[3 * pitch + 1];   d[2] = v - c[2 * pitch + 2];   d[3] = v - c[pitch + 3];
    d[4] = v - c[3];                  d[5] = v - c[-pitch + 3];      d[6] = v - c[-2 * pitch + 2];  d[7] = v - c[-3 * pitch + 1];
    d[8] = v - c[-3 * pitch];         d[9] = v - c[-3 * pitch - 1];  d[10] = v - c[-2 * pitch - 2]; d[11] = v - c[-pitch - 3];
    d[12] = v - c[-3];                d[13] = v - c[pitch - 3];      d[14] = v - c[2 * pitch - 2];  d[15] = v - c[3 * pitch - 1];
    // min / max over every window of 9 consecutive (circular) values, by doubling: 2, 4, 8, then +1
    int mn2[16], mx2[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { mn2[k] = min(d[k], d[(k + 1) & 15]); mx2[k] = max(d[k], d[(k + 1) & 15]); }
    int mn4[16], mx4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { mn4[k] = min(mn2[k], mn2[(k + 2) & 15]); mx4[k] = max(mx2[k], mx2[(k + 2) & 15]); }
    int A = -256, Bm = 256;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int mn9 = min(min(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]);
        const int mx9 = max(max(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]);
        A = max(A, mn9);
        Bm = min(Bm, mx9);
    }
    return max(A, -Bm) - 1;
}

__global__ void __launch_bounds__(FAST_THREADS) k_fast_cells(const uint8_t* __restrict__ pyr, uint32_t* __restrict__ cell_pts,
                                                             int* __restrict__ cell_cnt, const __grid_constant__ Plan P) {
    extern __shared__ uint8_t smem[];
    const int img = blockIdx.y;
    int level = 0, cell = blockIdx.x;
    while (level + 1 < P.n_levels && cell >= P.lv[level + 1].cell_base) ++level;
    const LevelGeom& g = P.lv[level];
    cell -= g.cell_base;
    const int ci = cell / g.n_cols, cj = cell % g.n_cols;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* out_cnt = cell_cnt + (size_t)img * P.cells_per_image + g.cell_base + cell;
    uint32_t* out_pts = cell_pts + (size_t)img * P.cand_per_image + g.cand_off + (size_t)cell * g.cell_cap;

    // cell image bounds — ORBextractor.cc:588-615
    const int max_bx = g.w - BORDER, max_by = g.h - BORDER;
    const int ini_y = BORDER + ci * g.h_cell, ini_x = BORDER + cj * g.w_cell;
    const int max_y = min(ini_y + g.h_cell + 6, max_by), max_x = min(ini_x + g.w_cell + 6, max_bx);
    const int cw = max_x - ini_x, ch = max_y - ini_y;        // cell image size
    const int dw = cw - 6, dh = ch - 6;                      // detection region (rows/cols [3, n-3))
    if (ini_y >= max_by - 3 || ini_x >= max_bx - 6 || dw <= 0 || dh <= 0) {
        if (tid == 0) *out_cnt = 0;
        return;
    }
    // shared layout: tile [ch][tp] u8 | score [(dh+2)][sp] u8 (1-px zero rim) | work list u16[dw*dh] | counters
    const int tp = (P.max_cell_w + 6 + 3) & ~3;
    const int sp = P.max_cell_w + 2;
    uint8_t* s_tile = smem;
    uint8_t* s_score = s_tile + (P.max_cell_h + 6) * tp;
    uint16_t* s_work = reinterpret_cast<uint16_t*>(s_score + (((P.max_cell_h + 2) * sp + 3) & ~3));
    __shared__ int s_nwork;
    __shared__ int s_warp_cnt[FAST_THREADS / 32];

    const uint8_t* src = pyr + (size_t)img * P.pyr_bytes + g.img_off + (size_t)ini_y * g.pitch + ini_x;
    for (int i = tid; i < ch * cw; i += FAST_THREADS) {
        const int y = i / cw, x = i - y * cw;
        s_tile[y * tp + x] = src[(size_t)y * g.pitch + x];
    }
    for (int i = tid; i < (dh + 2) * sp; i += FAST_THREADS) s_score[i] = 0;
    if (tid == 0) s_nwork = 0;
    __syncthreads();

    // phase 1: compass-point reject at minTh. A 9-arc always contains two compass pixels that are 4 apart.
    const int T = P.min_th;
    const int npx = dw * dh;
    for (int base = 0; base < npx; base += FAST_THREADS) {
        const int i = base + tid;
        bool pass = false;
        if (i < npx) {
            const int y = i / dw, x = i - y * dw;
            const uint8_t* c = s_tile + (y + 3) * tp + (x + 3);
            const int v = c[0];
            const int d0 = v - c[3 * tp], d4 = v - c[3], d8 = v - c[-3 * tp], d12 = v - c[-3];
            const bool b0 = d0 > T, b4 = d4 > T, b8 = d8 > T, b12 = d12 > T;
            const bool k0 = d0 < -T, k4 = d4 < -T, k8 = d8 < -T, k12 = d12 < -T;
            pass = (b0 && b4) || (b4 && b8) || (b8 && b12) || (b12 && b0) || (k0 && k4) || (k4 && k8) || (k8 && k12) || (k12 && k0);
        }
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        int wbase = 0;
        if (lane == 0 && m) wbase = atomicAdd(&s_nwork, __popc(m));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (pass) s_work[wbase + __popc(m & ((1u << lane) - 1))] = (uint16_t)i;
    }
    __syncthreads();
    // phase 2: full score for the survivors
    const int nwork = s_nwork;
    for (int k = tid; k < nwork; k += FAST_THREADS) {
        const int i = s_work[k];
        const int y = i / dw, x = i - y * dw;
        const int sc = fast_score16(s_tile + (y + 3) * tp + (x + 3), tp);
        if (sc >= T) s_score[(y + 1) * sp + (x + 1)] = (uint8_t)sc;
    }
    __syncthreads();
    // phase 3: strict 8-neighbour local maxima; bit k of `keep` = pixel (k * FAST_THREADS + tid)
    const int nchunk = (npx + FAST_THREADS - 1) / FAST_THREADS;  // <= 69*69/128 = 38 -> two 32-bit masks
    uint32_t keep_lo = 0, keep_hi = 0;
    bool any_ini = false;
    for (int k = 0; k < nchunk; ++k) {
        const int i = k * FAST_THREADS + tid;
        if (i < npx) {
            const int y = i / dw, x = i - y * dw;
            const uint8_t* c = s_score + (y + 1) * sp + (x + 1);
            const int s = c[0];
            if (s > 0 && s > c[-1] && s > c[1] && s > c[-sp - 1] && s > c[-sp] && s > c[-sp + 1] && s > c[sp - 1] && s > c[sp] && s > c[sp + 1]) {
                if (k < 32) keep_lo |= 1u << k; else keep_hi |= 1u << (k - 32);
                any_ini |= s >= P.ini_th;
            }
        }
    }
    const int th = __syncthreads_or(any_ini) ? P.ini_th : P.min_th;
    // phase 4: ordered compaction
    int total = 0;
    for (int k = 0; k < nchunk; ++k) {
        const int i = k * FAST_THREADS + tid;
        bool keep = (k < 32 ? (keep_lo >> k) : (keep_hi >> (k - 32))) & 1u;
        int y = 0, x = 0, s = 0;
        if (keep) {
            y = i / dw; x = i - y * dw;
            s = s_score[(y + 1) * sp + (x + 1)];
            keep = s >= th;
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp_cnt[warp] = __popc(m);
        __syncthreads();
        int before = 0, chunk_total = 0;
#pragma unroll
        for (int w = 0; w < FAST_THREADS / 32; ++w) { const int c = s_warp_cnt[w]; if (w < warp) before += c; chunk_total += c; }
        if (keep) {
            // reference coordinates: cell-image position + (j*wCell, i*hCell) == level position - BORDER
            const int px = ini_x + 3 + x - BORDER, py = ini_y + 3 + y - BORDER;
            out_pts[total + before + __popc(m & ((1u << lane) - 1))] = pack_pt(px, py, s);
        }
        total += chunk_total;
        __syncthreads();
    }
    if (tid == 0) *out_cnt = total;
}

size_t fast_smem_bytes(const Plan& P) {
    const int tp = (P.max_cell_w + 6 + 3) & ~3, sp = P.max_cell_w + 2;
    size_t b = (size_t)(P.max_cell_h + 6) * tp + (((P.max_cell_h + 2) * sp + 3) & ~3) + 2 * (size_t)P.max_cell_w * P.max_cell_h;
    return (b + 15) & ~(size_t)15;
}

int launch_fast_cells(const Plan& P, const uint8_t* d_pyr, uint32_t* d_cell_pts, int* d_cell_cnt, int n_images, cudaStream_t s) {
    dim3 grid(P.cells_per_image, n_images);
    k_fast_cells<<<grid, FAST_THREADS, fast_smem_bytes(P), s>>>(d_pyr, d_cell_pts, d_cell_cnt, P);
    return 1;
}

}  // namespace mcv
