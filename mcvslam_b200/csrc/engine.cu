// C ABI of libmcv_b200.so: plan construction, device workspaces, stream plumbing. See include/mcv_b200.h for the
// reference interface each entry point replaces. There is no CPU fallback anywhere in this file: every compute entry
// point launches the sm_100a kernels or fails with an error code.
#include "engine.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <cfloat>

namespace mcv {

static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }

// ---------------------------------------------------------------------------------------------------------
// device buffer that grows on demand
// ---------------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr; size_t bytes = 0;
    mcv_status reserve(size_t n) {
        if (n <= bytes) return MCV_OK;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        MCV_CUDA(cudaMalloc(&p, n));
        bytes = n;
        return MCV_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct HostBuf {  // pinned staging
    void* p = nullptr; size_t bytes = 0;
    mcv_status reserve(size_t n) {
        if (n <= bytes) return MCV_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; bytes = 0;
        MCV_CUDA(cudaMallocHost(&p, n));
        bytes = n;
        return MCV_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

static inline int cv_round_host(float v) { return (int)lrintf(v); }
static inline int cv_round_host(double v) { return (int)lrint(v); }
static inline int cv_floor_host(double v) { int i = (int)v; return i - (i > v); }

}  // namespace mcv

using namespace mcv;

// =========================================================================================================
// extractor handle
// =========================================================================================================
struct mcv_orb {
    mcv_orb_params prm{};
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // ORBextractor public vectors (ORBextractor.h:96-99) and quotas
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> quota;
    // plan for the current image size
    Plan plan{};
    bool have_plan = false;
    int cap_images = 0;     // workspace capacity (images)
    unsigned ws_epoch = 0;  // bumped whenever the workspace buffers are (re)allocated: captured graphs of older epochs are stale
    int last_images = 0;    // images processed by the last extract
    DevBuf tabs, src, pyr, blur, score, nz_list, nz_cnt, cell_raw, cell_pts, cell_cnt, fallback, arena_a, arena_b, oct_idx, out_pts, out_cnt, kps, desc, counts, seeds, misc;
    HostBuf h_stage;
    int last_cap = 0;       // per-image keypoint slots of the last extract (layout of kps/desc)
    int channels = 1;       // input images: 1 = CV_8UC1 gray, 3 = CV_8UC3 BGR (cvtColor fused into the level-0 write)
    int last_launches = 0;
    // optional per-stage timing: CUDA events recorded between the stages of every call (no synchronisation added)
    bool profile = false;
    std::vector<cudaEvent_t> ev;   // PROF_RING calls x (N_STAGES + 1) events
    int prof_calls = 0;
    // optional: the latency-bound quadtree kernel runs on its own highest-priority stream, fenced by two events, so that its
    // few-warp CTAs are dispatched ahead of the stencil CTAs of whatever chunk is running beside it (rig slots only)
    cudaStream_t quad_stream = nullptr;
    cudaEvent_t quad_in = nullptr, quad_out = nullptr;
};

constexpr int N_STAGES = 8;   // pyramid, blur, fast_score, nms_cells, quadtree, orient_desc, stereo_match, stereo_median
constexpr int PROF_RING = 512;
static inline void prof_mark(mcv_orb* h, int stage_boundary) {
    if (h->profile && h->prof_calls < PROF_RING) cudaEventRecord(h->ev[(size_t)h->prof_calls * (N_STAGES + 1) + stage_boundary], h->stream);
}

// ORBextractor::init — ORBextractor.cc:407-436 (same float arithmetic, evaluated on the host once)
static void orb_init_scales(mcv_orb* h) {
    const int n = h->prm.nlevels;
    const float sf = h->prm.scale_factor;
    h->scale.assign(n, 1.f); h->sigma2.assign(n, 1.f); h->inv_scale.assign(n, 1.f); h->inv_sigma2.assign(n, 1.f);
    for (int i = 1; i < n; i++) {
        h->scale[i] = h->scale[i - 1] * sf;
        h->sigma2[i] = h->scale[i] * h->scale[i];
    }
    for (int i = 0; i < n; i++) {
        h->inv_scale[i] = 1.0f / h->scale[i];
        h->inv_sigma2[i] = 1.0f / h->sigma2[i];
    }
    h->quota.assign(n, 0);
    float factor = 1.0f / sf;
    float nDesired = h->prm.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)n));
    int sum = 0;
    for (int level = 0; level < n - 1; level++) {
        h->quota[level] = cv_round_host(nDesired);
        sum += h->quota[level];
        nDesired *= factor;
    }
    h->quota[n - 1] = std::max(h->prm.nfeatures - sum, 0);
}

// Level geometry for an image size + resize coefficient tables (cv::resize INTER_LINEAR model, see image_kernels.cu)
static mcv_status build_plan(mcv_orb* h, int w, int hgt, std::vector<int>& tabs) {
    Plan& P = h->plan;
    memset(&P, 0, sizeof(P));
    P.n_levels = h->prm.nlevels; P.w = w; P.h = hgt; P.ini_th = h->prm.ini_th_fast; P.min_th = h->prm.min_th_fast;
    tabs.clear();
    int img_off = 0, cell_base = 0, cand_off = 0, out_off = 0;
    for (int l = 0; l < P.n_levels; ++l) {
        LevelGeom& g = P.lv[l];
        g.scale = h->scale[l]; g.inv_scale = h->inv_scale[l];
        g.w = cv_round_host((float)w * g.inv_scale);   // ORBextractor.cc:904
        g.h = cv_round_host((float)hgt * g.inv_scale);
        if (g.w > MAX_DIM || g.h > MAX_DIM) { set_error("image larger than 4128 px is not supported"); return MCV_ERR_BAD_ARG; }
        g.pitch = (g.w + 15) & ~15;
        g.img_off = img_off;
        img_off += ((g.pitch * g.h) + 255) & ~255;
        // cell grid — ORBextractor.cc:588-602
        const int max_bx = g.w - BORDER, max_by = g.h - BORDER;
        const float width = (float)(max_bx - BORDER), height = (float)(max_by - BORDER);
        g.n_cols = (int)(width / 35.f); g.n_rows = (int)(height / 35.f);
        if (g.n_cols < 1 || g.n_rows < 1) { set_error("pyramid level smaller than one FAST cell"); return MCV_ERR_IMAGE_TOO_SMALL; }
        g.w_cell = (int)ceilf(width / g.n_cols); g.h_cell = (int)ceilf(height / g.n_rows);
        g.cell_base = cell_base;
        cell_base += g.n_cols * g.n_rows;
        g.cell_cap = ((g.w_cell + 1) / 2) * ((g.h_cell + 1) / 2);
        g.cand_off = cand_off;
        g.cand_cap = g.n_cols * g.n_rows * g.cell_cap;
        cand_off += g.cand_cap;
        g.quota = h->quota[l];
        // quadtree roots — ORBextractor.cc:527-529
        g.n_ini = (int)roundf(width / (float)(max_by - BORDER));
        if (g.n_ini < 1) { set_error("image aspect ratio below 0.5 is not supported (reference divides by zero)"); return MCV_ERR_BAD_ARG; }
        g.h_x = width / (float)g.n_ini;
        g.out_off = out_off;
        g.out_cap = g.quota + 3 + g.n_ini;
        out_off += g.out_cap;
        g.kp_size = (int)(31 * g.scale);   // ORBextractor.cc:640
        P.max_cell_w = std::max(P.max_cell_w, g.w_cell); P.max_cell_h = std::max(P.max_cell_h, g.h_cell);
        P.max_quota = std::max(P.max_quota, g.quota);
        // resize tables for level l <- l-1
        g.tab_off = (int)tabs.size();
        g.area_fast = 0;
        if (l > 0) {
            const LevelGeom& s = P.lv[l - 1];
            const double scale_x = 1. / ((double)g.w / s.w), scale_y = 1. / ((double)g.h / s.h);
            const int isx = cv_round_host(scale_x), isy = cv_round_host(scale_y);
            g.area_fast = (std::abs(scale_x - isx) < DBL_EPSILON && std::abs(scale_y - isy) < DBL_EPSILON && isx == 2 && isy == 2) ? 1 : 0;
            std::vector<int> xofs(g.w), xa0(g.w), xa1(g.w), yofs(g.h), yb0(g.h), yb1(g.h);
            for (int dx = 0; dx < g.w; ++dx) {
                float fx = (float)((dx + 0.5) * scale_x - 0.5);
                int sx = cv_floor_host(fx);
                fx -= sx;
                if (sx < 0) { fx = 0; sx = 0; }
                if (sx >= s.w - 1) { fx = 0; sx = s.w - 1; }
                xofs[dx] = sx;
                xa0[dx] = (short)cv_round_host((1.f - fx) * 2048.f);
                xa1[dx] = (short)cv_round_host(fx * 2048.f);
            }
            for (int dy = 0; dy < g.h; ++dy) {
                float fy = (float)((dy + 0.5) * scale_y - 0.5);
                int sy = cv_floor_host(fy);
                fy -= sy;
                yofs[dy] = sy;
                yb0[dy] = (short)cv_round_host((1.f - fy) * 2048.f);
                yb1[dy] = (short)cv_round_host(fy * 2048.f);
            }
            // k_resize_march preconditions: S[sx], S[sx+1] of 4 consecutive output px inside one 8-byte window, non-negative
            // coefficients, source rows strictly increasing, source wide enough for three aligned word loads per row
            bool ok = !g.area_fast && s.pitch >= 16;
            for (int dx = 0; dx < g.w && ok; dx += 4) {
                const int last = std::min(dx + 3, g.w - 1);
                ok = std::min(xofs[last] + 1, s.w - 1) - xofs[dx] <= 7 && xofs[last] >= xofs[dx];
                for (int k = dx; k <= last && ok; ++k) ok = xofs[k] >= xofs[dx] && xa0[k] >= 0 && xa1[k] >= 0;
            }
            for (int dy = 0; dy < g.h && ok; ++dy) ok = yb0[dy] >= 0 && yb1[dy] >= 0 && (dy == 0 || yofs[dy] > yofs[dy - 1]);
            g.march_ok = ok ? 1 : 0;
            for (auto* v : {&xofs, &xa0, &xa1, &yofs, &yb0, &yb1}) tabs.insert(tabs.end(), v->begin(), v->end());
        }
    }
    P.pyr_bytes = img_off;
    P.cells_per_image = cell_base;
    P.cand_per_image = cand_off;
    P.out_per_image = out_off;
    { StripTable T{}; P.n_fast_strips = fast_strip_table(P, T); }
    P.max_quad_kp = out_off;
    return MCV_OK;
}

static mcv_status ensure_workspace(mcv_orb* h, int w, int hgt, int n_images, int cap) {
    if (!h->have_plan || h->plan.w != w || h->plan.h != hgt) {
        std::vector<int> tabs;
        h->have_plan = false;
        mcv_status st = build_plan(h, w, hgt, tabs);
        if (st) return st;
        st = h->tabs.reserve(std::max<size_t>(16, tabs.size() * sizeof(int)));
        if (st) return st;
        if (!tabs.empty()) MCV_CUDA(cudaMemcpyAsync(h->tabs.p, tabs.data(), tabs.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        MCV_CUDA(cudaStreamSynchronize(h->stream));  // tabs is a local
        h->have_plan = true;
        h->cap_images = 0;
    }
    const Plan& P = h->plan;
    if (n_images > h->cap_images) {
        mcv_status st;
        // + spare bytes: k_resize_march reads whole words past a row end, and k_fast_score prefetches up to FS_ROWS + 6 rows past the
        // end of a level (rows that feed no output) without clamping
        if ((st = h->pyr.reserve((size_t)P.pyr_bytes * n_images + 256 + (size_t)(FS_ROWS + 8) * P.lv[0].pitch))) return st;
        if ((st = h->blur.reserve((size_t)P.pyr_bytes * n_images))) return st;
        if ((st = h->score.reserve(std::max<size_t>(16, (size_t)P.n_fast_strips * FS_EDGE_BYTES * n_images)))) return st;   // per-strip edge records (no dense score map)
        if ((st = h->cell_pts.reserve((size_t)P.cand_per_image * n_images * 4))) return st;
        if ((st = h->cell_raw.reserve((size_t)P.cand_per_image * n_images * 4))) return st;
        if ((st = h->nz_list.reserve((size_t)P.n_fast_strips * FS_SEG * n_images * 4))) return st;
        if ((st = h->nz_cnt.reserve((size_t)std::max(1, P.n_fast_strips) * n_images * 4))) return st;
        if ((st = h->arena_a.reserve((size_t)P.cand_per_image * n_images * 4))) return st;
        if ((st = h->arena_b.reserve((size_t)P.cand_per_image * n_images * 4))) return st;
        // sorted indices (u16 per candidate) | bucket prefix sums of every quadtree task (launch_octree)
        if ((st = h->oct_idx.reserve((size_t)P.cand_per_image * n_images * 2 + 512 + (size_t)P.n_levels * n_images * OCT_S_BYTES))) return st;
        if ((st = h->cell_cnt.reserve((size_t)P.cells_per_image * n_images * 4))) return st;
        if ((st = h->fallback.reserve(((size_t)P.cells_per_image * n_images + 4) * 4))) return st;
        if ((st = h->out_pts.reserve((size_t)P.out_per_image * n_images * 4))) return st;
        if ((st = h->out_cnt.reserve((size_t)P.n_levels * n_images * 4 * 2))) return st;   // counts | per-task overflow flags (launch_octree)
        h->cap_images = n_images;
        ++h->ws_epoch;
    }
    mcv_status st;
    if ((st = h->kps.reserve((size_t)cap * n_images * sizeof(mcv_keypoint)))) return st;
    if ((st = h->desc.reserve((size_t)cap * n_images * 32))) return st;
    if ((st = h->counts.reserve((size_t)n_images * 4))) return st;
    return MCV_OK;
}

// Enqueue the whole extraction of n_images device-resident images. d_kps/d_desc/d_counts: device outputs, `cap` slots/image.
// `wait_front` / `signal_front` (optional) stagger concurrent chunks: the issue-bound front half (pyramid .. per-cell lists)
// of a chunk starts when the previous chunk's front half is done, so that it runs beside that chunk's latency-bound
// quadtree kernel instead of beside its front half.
static mcv_status enqueue_extract(mcv_orb* h, const uint8_t* d_imgs, size_t src_pitch, size_t src_image_stride, int n_images,
                                  const SeedInfo* seeds, mcv_keypoint* d_kps, uint8_t* d_desc, int* d_counts, int cap,
                                  cudaEvent_t wait_front = nullptr, cudaEvent_t signal_front = nullptr) {
    const Plan& P = h->plan;
    int n = 0;
    if (wait_front) MCV_CUDA(cudaStreamWaitEvent(h->stream, wait_front, 0));
    prof_mark(h, 0);
    n += launch_pyramid(P, d_imgs, src_pitch, src_image_stride, h->channels, h->pyr.as<uint8_t>(), h->tabs.as<int>(), n_images, h->stream);
    prof_mark(h, 1);
    n += launch_blur(P, h->pyr.as<uint8_t>(), h->blur.as<uint8_t>(), n_images, h->stream);
    prof_mark(h, 2);
    const int rf = launch_fast_cells(P, h->pyr.as<uint8_t>(), h->score.as<uint8_t>(), h->nz_list.as<unsigned>(), h->nz_cnt.as<int>(), h->cell_raw.as<uint32_t>(),
                           h->cell_pts.as<uint32_t>(), h->cell_cnt.as<int>(), h->fallback.as<int>(), n_images, h->stream,
                           (h->profile && h->prof_calls < PROF_RING) ? h->ev[(size_t)h->prof_calls * (N_STAGES + 1) + 3] : nullptr);
    if (rf < 0) return MCV_ERR_CUDA;
    n += rf;
    if (signal_front) MCV_CUDA(cudaEventRecord(signal_front, h->stream));
    prof_mark(h, 4);
    cudaStream_t qs = h->stream;
    if (h->quad_stream && !h->profile) {
        MCV_CUDA(cudaEventRecord(h->quad_in, h->stream));
        MCV_CUDA(cudaStreamWaitEvent(h->quad_stream, h->quad_in, 0));
        qs = h->quad_stream;
    }
    const int r = launch_octree(P, h->cell_pts.as<uint32_t>(), h->cell_cnt.as<int>(), h->arena_a.as<uint32_t>(), h->arena_b.as<uint32_t>(),
                                h->oct_idx.as<uint16_t>(), h->out_pts.as<uint32_t>(), h->out_cnt.as<int>(), n_images, qs);
    if (qs != h->stream) {
        MCV_CUDA(cudaEventRecord(h->quad_out, qs));
        MCV_CUDA(cudaStreamWaitEvent(h->stream, h->quad_out, 0));
    }
    if (r < 0) { set_error("nfeatures too large for the quadtree kernel's shared-memory heap"); return MCV_ERR_CAPACITY; }
    n += r;
    prof_mark(h, 5);
    const int rd = launch_orient_desc(P, h->pyr.as<uint8_t>(), h->blur.as<uint8_t>(), h->out_pts.as<uint32_t>(), h->out_cnt.as<int>(), seeds, d_kps,
                                      d_desc, d_counts, cap, n_images, h->stream);
    if (rd < 0) return MCV_ERR_CUDA;
    n += rd;
    prof_mark(h, 6);
    MCV_CUDA(cudaGetLastError());
    h->last_images = n_images; h->last_cap = cap; h->last_launches = n;
    return MCV_OK;
}

extern "C" {

const char* mcv_last_error(void) { return g_err.c_str(); }
const char* mcv_version(void) { return "mcv_b200 0.1 (sm_100a)"; }
int mcv_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

mcv_status mcv_orb_create(const mcv_orb_params* p, int device, void* stream, mcv_orb** out) {
    if (!p || !out) return MCV_ERR_BAD_ARG;
    *out = nullptr;
    if (p->nlevels < 1 || p->nlevels > MAX_LEVELS || p->nfeatures < 1 || !(p->scale_factor > 1.0f) || p->min_th_fast < 1 ||
        p->ini_th_fast < p->min_th_fast || p->ini_th_fast > 254) {
        set_error("bad extractor parameters");
        return MCV_ERR_BAD_ARG;
    }
    if (mcv_device_count() <= device) { set_error("no CUDA device"); return MCV_ERR_NO_DEVICE; }
    MCV_CUDA(cudaSetDevice(device));
    mcv_orb* h = new mcv_orb();
    h->prm = *p; h->device = device;
    if (stream) h->stream = (cudaStream_t)stream;
    else { if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; set_error("cudaStreamCreate failed"); return MCV_ERR_CUDA; } h->own_stream = true; }
    orb_init_scales(h);
    *out = h;
    return MCV_OK;
}

void mcv_orb_destroy(mcv_orb* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (DevBuf* b : {&h->tabs, &h->src, &h->pyr, &h->blur, &h->score, &h->nz_list, &h->nz_cnt, &h->cell_raw, &h->cell_pts, &h->cell_cnt, &h->fallback, &h->arena_a, &h->arena_b, &h->oct_idx, &h->out_pts, &h->out_cnt,
                      &h->kps, &h->desc, &h->counts, &h->seeds, &h->misc})
        b->release();
    h->h_stage.release();
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    if (h->quad_stream) { cudaStreamDestroy(h->quad_stream); cudaEventDestroy(h->quad_in); cudaEventDestroy(h->quad_out); }
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

mcv_status mcv_orb_get_scales(const mcv_orb* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int32_t* fpl) {
    if (!h) return MCV_ERR_BAD_ARG;
    for (int i = 0; i < h->prm.nlevels; ++i) {
        if (scale) scale[i] = h->scale[i];
        if (inv_scale) inv_scale[i] = h->inv_scale[i];
        if (sigma2) sigma2[i] = h->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = h->inv_sigma2[i];
        if (fpl) fpl[i] = h->quota[i];
    }
    return MCV_OK;
}

// per level: quota + 3 (overshoot of the last split, ORBextractor.cc:557-565) + n_ini roots (build_plan: out_cap)
static int max_keypoints_roots(const mcv_orb* h, int roots_per_level) {
    int n = 0;
    for (int l = 0; l < h->prm.nlevels; ++l) n += h->quota[l] + 3 + roots_per_level;
    return n;
}
int mcv_orb_max_keypoints(const mcv_orb* h, int n_seeds) {
    if (!h) return 0;
    // size-independent bound: 15 roots per level is the most the quadtree kernels accept (any aspect ratio)
    return max_keypoints_roots(h, 15) + std::max(0, n_seeds);
}
int mcv_orb_max_keypoints_for(const mcv_orb* h, int w, int hgt, int n_seeds) {
    if (!h || w <= 0 || hgt <= 0) return 0;
    int n = 0;
    for (int l = 0; l < h->prm.nlevels; ++l) {   // n_ini exactly as build_plan / ORBextractor.cc:527-529
        const int lw = cv_round_host((float)w * h->inv_scale[l]), lh = cv_round_host((float)hgt * h->inv_scale[l]);
        const int den = lh - 2 * BORDER;
        const int n_ini = den > 0 ? (int)roundf((float)(lw - 2 * BORDER) / (float)den) : 15;
        n += h->quota[l] + 3 + std::max(1, std::min(15, n_ini));
    }
    return n + std::max(0, n_seeds);
}

mcv_status mcv_orb_extract(mcv_orb* h, const uint8_t* img, int w, int hgt, size_t stride, const mcv_keypoint* seeds, int n_seeds,
                           mcv_keypoint* kps_out, uint8_t* desc_out, int cap, int* n_out) {
    if (!h || !n_out) return MCV_ERR_BAD_ARG;
    *n_out = 0;
    if (!img || w <= 0 || hgt <= 0) return MCV_ERR_EMPTY_IMAGE;  // ORBextractor.cc:834
    const size_t row_bytes = (size_t)w * h->channels;
    if (stride < row_bytes || n_seeds < 0 || (n_seeds > 0 && !seeds) || !kps_out || !desc_out) return MCV_ERR_BAD_ARG;
    MCV_CUDA(cudaSetDevice(h->device));
    mcv_status st = ensure_workspace(h, w, hgt, 1, std::max(cap, 1));
    if (st) return st;
    const Plan& P = h->plan;
    if (cap < P.max_quad_kp + n_seeds) { set_error("cap smaller than mcv_orb_max_keypoints_for(w, h)"); return MCV_ERR_CAPACITY; }
    SeedInfo si{};
    if (n_seeds > 0) {
        for (int i = 0; i < n_seeds; ++i) {
            const mcv_keypoint& k = seeds[i];
            if (k.octave < 0 || k.octave >= P.n_levels) { set_error("seed octave out of range"); return MCV_ERR_SEED_RANGE; }
            const LevelGeom& g = P.lv[k.octave];
            const int cx = cv_round_host(k.x), cy = cv_round_host(k.y);
            if (cx < EDGE_THRESHOLD || cy < EDGE_THRESHOLD || cx >= g.w - EDGE_THRESHOLD || cy >= g.h - EDGE_THRESHOLD) {
                set_error("seed keypoint within 19 px of its level border (the reference would read outside the image)");
                return MCV_ERR_SEED_RANGE;
            }
        }
        if ((st = h->seeds.reserve((size_t)n_seeds * sizeof(mcv_keypoint)))) return st;
        MCV_CUDA(cudaMemcpyAsync(h->seeds.p, seeds, (size_t)n_seeds * sizeof(mcv_keypoint), cudaMemcpyHostToDevice, h->stream));
        si.d_seeds = h->seeds.as<mcv_keypoint>(); si.n_seeds = n_seeds;
    }
    if ((st = h->src.reserve(row_bytes * hgt))) return st;
    MCV_CUDA(cudaMemcpy2DAsync(h->src.p, row_bytes, img, stride, row_bytes, hgt, cudaMemcpyHostToDevice, h->stream));
    st = enqueue_extract(h, h->src.as<uint8_t>(), row_bytes, row_bytes * hgt, 1, n_seeds ? &si : nullptr, h->kps.as<mcv_keypoint>(),
                         h->desc.as<uint8_t>(), h->counts.as<int>(), cap);
    if (st) return st;
    int n = 0;
    MCV_CUDA(cudaMemcpyAsync(&n, h->counts.p, 4, cudaMemcpyDeviceToHost, h->stream));
    MCV_CUDA(cudaStreamSynchronize(h->stream));
    if (n > cap) { set_error("keypoint capacity exceeded"); return MCV_ERR_CAPACITY; }
    if (n > 0) {
        MCV_CUDA(cudaMemcpyAsync(kps_out, h->kps.p, (size_t)n * sizeof(mcv_keypoint), cudaMemcpyDeviceToHost, h->stream));
        MCV_CUDA(cudaMemcpyAsync(desc_out, h->desc.p, (size_t)n * 32, cudaMemcpyDeviceToHost, h->stream));
        MCV_CUDA(cudaStreamSynchronize(h->stream));
    }
    *n_out = n;
    return MCV_OK;
}

mcv_status mcv_orb_extract_batch(mcv_orb* h, const uint8_t* imgs, int n_images, int w, int hgt, int imgs_on_device, mcv_keypoint* kps_out,
                                 uint8_t* desc_out, int32_t* counts, int cap, int out_on_device) {
    if (!h || !imgs || n_images <= 0 || !kps_out || !desc_out || !counts) return MCV_ERR_BAD_ARG;
    if (w <= 0 || hgt <= 0) return MCV_ERR_EMPTY_IMAGE;
    MCV_CUDA(cudaSetDevice(h->device));
    mcv_status st = ensure_workspace(h, w, hgt, n_images, out_on_device ? 1 : cap);
    if (st) return st;
    if (cap < h->plan.max_quad_kp) { set_error("cap smaller than mcv_orb_max_keypoints_for(w, h)"); return MCV_ERR_CAPACITY; }
    const uint8_t* d_imgs = imgs;
    const size_t img_bytes = (size_t)w * hgt * h->channels;
    if (!imgs_on_device) {
        if ((st = h->src.reserve(img_bytes * n_images))) return st;
        MCV_CUDA(cudaMemcpyAsync(h->src.p, imgs, img_bytes * n_images, cudaMemcpyHostToDevice, h->stream));
        d_imgs = h->src.as<uint8_t>();
    }
    mcv_keypoint* d_kps = out_on_device ? kps_out : h->kps.as<mcv_keypoint>();
    uint8_t* d_desc = out_on_device ? desc_out : h->desc.as<uint8_t>();
    int* d_counts = out_on_device ? counts : h->counts.as<int>();
    st = enqueue_extract(h, d_imgs, (size_t)w * h->channels, img_bytes, n_images, nullptr, d_kps, d_desc, d_counts, cap);
    if (st) return st;
    if (!out_on_device) {
        MCV_CUDA(cudaMemcpyAsync(kps_out, d_kps, (size_t)cap * n_images * sizeof(mcv_keypoint), cudaMemcpyDeviceToHost, h->stream));
        MCV_CUDA(cudaMemcpyAsync(desc_out, d_desc, (size_t)cap * n_images * 32, cudaMemcpyDeviceToHost, h->stream));
        MCV_CUDA(cudaMemcpyAsync(counts, d_counts, (size_t)n_images * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    MCV_CUDA(cudaStreamSynchronize(h->stream));
    return MCV_OK;
}

// Device-resident, asynchronous form of mcv_orb_extract_batch: only enqueues on the handle's stream.
mcv_status mcv_orb_extract_batch_async(mcv_orb* h, const uint8_t* d_imgs, int n_images, int w, int hgt, mcv_keypoint* d_kps, uint8_t* d_desc,
                                       int32_t* d_counts, int cap) {
    if (!h || !d_imgs || n_images <= 0 || !d_kps || !d_desc || !d_counts) return MCV_ERR_BAD_ARG;
    if (w <= 0 || hgt <= 0) return MCV_ERR_EMPTY_IMAGE;
    MCV_CUDA(cudaSetDevice(h->device));
    mcv_status st = ensure_workspace(h, w, hgt, n_images, 1);
    if (st) return st;
    if (cap < h->plan.max_quad_kp) { set_error("cap smaller than mcv_orb_max_keypoints_for(w, h)"); return MCV_ERR_CAPACITY; }
    return enqueue_extract(h, d_imgs, (size_t)w * h->channels, (size_t)w * hgt * h->channels, n_images, nullptr, d_kps, d_desc, d_counts, cap);
}

mcv_status mcv_orb_level_device(mcv_orb* h, int image_index, int level, const uint8_t** dev_ptr, int* w, int* hgt, size_t* pitch) {
    if (!h || !h->have_plan || image_index < 0 || image_index >= h->last_images || level < 0 || level >= h->plan.n_levels) return MCV_ERR_BAD_ARG;
    const LevelGeom& g = h->plan.lv[level];
    if (dev_ptr) *dev_ptr = h->pyr.as<uint8_t>() + (size_t)image_index * h->plan.pyr_bytes + g.img_off;
    if (w) *w = g.w;
    if (hgt) *hgt = g.h;
    if (pitch) *pitch = g.pitch;
    return MCV_OK;
}

static mcv_status download_plane(mcv_orb* h, const DevBuf& buf, int image_index, int level, uint8_t* dst, size_t dst_stride, int* w, int* hgt) {
    if (!h || !h->have_plan || image_index < 0 || image_index >= h->last_images || level < 0 || level >= h->plan.n_levels) return MCV_ERR_BAD_ARG;
    const LevelGeom& g = h->plan.lv[level];
    if (w) *w = g.w;
    if (hgt) *hgt = g.h;
    if (!dst) return MCV_OK;
    if (dst_stride < (size_t)g.w) return MCV_ERR_BAD_ARG;
    MCV_CUDA(cudaSetDevice(h->device));
    MCV_CUDA(cudaMemcpy2DAsync(dst, dst_stride, buf.as<uint8_t>() + (size_t)image_index * h->plan.pyr_bytes + g.img_off, g.pitch, g.w, g.h,
                               cudaMemcpyDeviceToHost, h->stream));
    MCV_CUDA(cudaStreamSynchronize(h->stream));
    return MCV_OK;
}

mcv_status mcv_orb_download_level(mcv_orb* h, int image_index, int level, uint8_t* dst, size_t dst_stride, int* w, int* hgt) {
    return download_plane(h, h ? h->pyr : DevBuf(), image_index, level, dst, dst_stride, w, hgt);
}
mcv_status mcv_debug_download_blurred(mcv_orb* h, int image_index, int level, uint8_t* dst, size_t dst_stride) {
    return download_plane(h, h ? h->blur : DevBuf(), image_index, level, dst, dst_stride, nullptr, nullptr);
}

mcv_status mcv_debug_level_keypoints(mcv_orb* h, int image_index, int level, int which, mcv_keypoint* out, int cap, int* n_out) {
    if (!h || !h->have_plan || image_index < 0 || image_index >= h->last_images || level < 0 || level >= h->plan.n_levels || !n_out) return MCV_ERR_BAD_ARG;
    MCV_CUDA(cudaSetDevice(h->device));
    const Plan& P = h->plan;
    const LevelGeom& g = P.lv[level];
    std::vector<uint32_t> pts;
    if (which == 0) {
        const int nc = g.n_cols * g.n_rows;
        std::vector<int> cnt(nc);
        MCV_CUDA(cudaMemcpy(cnt.data(), h->cell_cnt.as<int>() + (size_t)image_index * P.cells_per_image + g.cell_base, nc * 4, cudaMemcpyDeviceToHost));
        std::vector<uint32_t> all(g.cand_cap);
        MCV_CUDA(cudaMemcpy(all.data(), h->cell_pts.as<uint32_t>() + (size_t)image_index * P.cand_per_image + g.cand_off, (size_t)g.cand_cap * 4, cudaMemcpyDeviceToHost));
        for (int c = 0; c < nc; ++c) pts.insert(pts.end(), all.begin() + (size_t)c * g.cell_cap, all.begin() + (size_t)c * g.cell_cap + cnt[c]);
    } else {
        int n = 0;
        MCV_CUDA(cudaMemcpy(&n, h->out_cnt.as<int>() + (size_t)image_index * P.n_levels + level, 4, cudaMemcpyDeviceToHost));
        pts.resize(n);
        if (n) MCV_CUDA(cudaMemcpy(pts.data(), h->out_pts.as<uint32_t>() + (size_t)image_index * P.out_per_image + g.out_off, (size_t)n * 4, cudaMemcpyDeviceToHost));
    }
    *n_out = (int)pts.size();
    for (int i = 0; i < (int)pts.size() && i < cap; ++i)
        out[i] = mcv_keypoint{(float)pt_x(pts[i]), (float)pt_y(pts[i]), 7.f, -1.f, (float)pt_r(pts[i]), 0, -1};
    return MCV_OK;
}

mcv_status mcv_orb_distribute_octree(mcv_orb* h, const mcv_keypoint* in, int n, int min_x, int max_x, int min_y, int max_y, int n_target,
                                     mcv_keypoint* out, int cap, int* n_out) {
    if (!h || !n_out || n < 0 || (n > 0 && !in) || max_x <= min_x || max_y <= min_y || n_target < 0) return MCV_ERR_BAD_ARG;
    *n_out = 0;
    if (n == 0) return MCV_OK;
    MCV_CUDA(cudaSetDevice(h->device));
    const int bw = max_x - min_x, bh = max_y - min_y;
    if (bw > 4095 || bh > 4095) return MCV_ERR_BAD_ARG;
    std::vector<uint32_t> pts(n);
    for (int i = 0; i < n; ++i) {
        const int x = (int)in[i].x, y = (int)in[i].y, r = (int)in[i].response;
        if (x < 0 || y < 0 || x >= bw || y >= bh || r < 0 || r > 255 || (float)x != in[i].x || (float)y != in[i].y || (float)r != in[i].response) {
            set_error("distribute_octree: keypoints must be integer-valued, inside the box, response in [0,255]");
            return MCV_ERR_BAD_ARG;
        }
        pts[i] = pack_pt(x, y, r);
    }
    const int out_cap = n_target + 8 + (int)roundf((float)bw / (float)bh);
    mcv_status st;
    if ((st = h->misc.reserve(((size_t)3 * n + out_cap + 4) * 4))) return st;
    uint32_t* d_in = h->misc.as<uint32_t>(); uint32_t* d_a = d_in + n; uint32_t* d_b = d_a + n; uint32_t* d_out = d_b + n;
    int* d_cnt = reinterpret_cast<int*>(d_out + out_cap);
    MCV_CUDA(cudaMemcpyAsync(d_in, pts.data(), (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    if (launch_octree_standalone(d_in, n, bw, bh, n_target, d_a, d_b, d_out, d_cnt, out_cap, h->stream) < 0) { set_error("distribute_octree: unsupported size"); return MCV_ERR_CAPACITY; }
    MCV_CUDA(cudaGetLastError());
    int cnt = 0;
    MCV_CUDA(cudaMemcpyAsync(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost, h->stream));
    MCV_CUDA(cudaStreamSynchronize(h->stream));
    std::vector<uint32_t> res(cnt);
    if (cnt) MCV_CUDA(cudaMemcpy(res.data(), d_out, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
    *n_out = cnt;
    if (cnt > cap) return MCV_ERR_CAPACITY;
    for (int i = 0; i < cnt; ++i) out[i] = mcv_keypoint{(float)pt_x(res[i]), (float)pt_y(res[i]), 7.f, -1.f, (float)pt_r(res[i]), 0, -1};
    return MCV_OK;
}

}  // extern "C"

// =========================================================================================================
// matcher (stateless in the reference: process-wide scratch on the current device, default stream 0 unless given)
// =========================================================================================================
namespace {
// The reference's Matcher is stateless and callable from any thread (include/Matcher.hpp:58-92). Every host
// thread therefore gets its OWN context per device — a non-blocking stream plus every scratch buffer the matcher-side entry
// points use — created on first use on the calling thread's current device (cudaGetDevice, like any CUDA library) and freed
// when the thread exits. Concurrent calls from different threads, or on different devices, share no state.
struct MatchCtx {
    int device = -1;
    cudaStream_t stream = nullptr;
    DevBuf q, t, idx, dist, off, cidx, part;
    DevBuf slot[12];          // per-entry-point buffers (project / fuse / wnd_track / distinctive / lk / bow / debug)
};
struct MatchCtxList {
    std::vector<MatchCtx*> v;
    ~MatchCtxList() {
        for (MatchCtx* c : v) {
            if (cudaSetDevice(c->device) == cudaSuccess) {   // fails harmlessly once the runtime is shutting down
                DevBuf* all[] = {&c->q, &c->t, &c->idx, &c->dist, &c->off, &c->cidx, &c->part};
                for (DevBuf* b : all) b->release();
                for (DevBuf& b : c->slot) b.release();
                if (c->stream) cudaStreamDestroy(c->stream);
            }
            delete c;
        }
    }
};
thread_local MatchCtxList tl_match;

mcv_status match_ctx(MatchCtx** out) {
    if (mcv_device_count() < 1) { set_error("no CUDA device"); return MCV_ERR_NO_DEVICE; }
    int dev = 0;
    MCV_CUDA(cudaGetDevice(&dev));
    for (MatchCtx* c : tl_match.v) if (c->device == dev) { *out = c; return MCV_OK; }
    MatchCtx* c = new MatchCtx;
    c->device = dev;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; set_error("cudaStreamCreate failed"); return MCV_ERR_CUDA; }
    tl_match.v.push_back(c);
    *out = c;
    return MCV_OK;
}

// runs brute-force 2-NN on host descriptors, returns idx/dist vectors (nq*2)
mcv_status bf_host(const uint8_t* q, int nq, const uint8_t* t, int nt, std::vector<int32_t>& idx, std::vector<int32_t>& dist) {
    MatchCtx* mc;
    mcv_status st = match_ctx(&mc);
    if (st) return st;
    cudaStream_t s = mc->stream;
    idx.assign((size_t)nq * 2, -1); dist.assign((size_t)nq * 2, 0x7fffffff);
    if (nq == 0) return MCV_OK;
    if ((st = mc->q.reserve((size_t)nq * 32))) return st;
    if ((st = mc->t.reserve(std::max<size_t>(32, (size_t)nt * 32)))) return st;
    if ((st = mc->idx.reserve((size_t)nq * 8))) return st;
    if ((st = mc->dist.reserve((size_t)nq * 8))) return st;
    MCV_CUDA(cudaMemcpyAsync(mc->q.p, q, (size_t)nq * 32, cudaMemcpyHostToDevice, s));
    if (nt) MCV_CUDA(cudaMemcpyAsync(mc->t.p, t, (size_t)nt * 32, cudaMemcpyHostToDevice, s));
    if ((st = mc->part.reserve(std::max<size_t>(8, knn2_bf_part_bytes(nq, nt))))) return st;
    if (launch_knn2_bf(mc->q.as<uint8_t>(), nq, mc->t.as<uint8_t>(), nt, 0, mc->idx.as<int32_t>(), mc->dist.as<int32_t>(), mc->part.as<unsigned>(), s) < 0) {
        set_error("knn2_bf: train set larger than 4M rows per call (tile it with train_offset) or out of memory");
        return MCV_ERR_CAPACITY;
    }
    MCV_CUDA(cudaGetLastError());
    MCV_CUDA(cudaMemcpyAsync(idx.data(), mc->idx.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaMemcpyAsync(dist.data(), mc->dist.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaStreamSynchronize(s));
    return MCV_OK;
}
}  // namespace

extern "C" {

mcv_status mcv_knn2_bf(const uint8_t* q, int nq, const uint8_t* t, int nt, mcv_dmatch* out, int* k_out) {
    if (nq < 0 || nt < 0 || (nq > 0 && (!q || !out)) || (nt > 0 && !t)) return MCV_ERR_BAD_ARG;
    std::vector<int32_t> idx, dist;
    mcv_status st = bf_host(q, nq, t, nt, idx, dist);
    if (st) return st;
    for (int i = 0; i < nq; ++i)
        for (int e = 0; e < 2; ++e) {
            const int j = idx[2 * i + e];
            out[2 * i + e] = j >= 0 ? mcv_dmatch{i, j, 0, (float)dist[2 * i + e]} : mcv_dmatch{i, -1, 0, 0.f};
        }
    if (k_out) *k_out = std::min(2, nt);
    return MCV_OK;
}

mcv_status mcv_bf_match(const uint8_t* q, int nq, const uint8_t* t, int nt, mcv_dmatch* out) {
    if (nq < 0 || nt < 1 || (nq > 0 && (!q || !out)) || !t) return MCV_ERR_BAD_ARG;
    std::vector<int32_t> idx, dist;
    mcv_status st = bf_host(q, nq, t, nt, idx, dist);
    if (st) return st;
    for (int i = 0; i < nq; ++i) out[i] = mcv_dmatch{i, idx[2 * i], 0, (float)dist[2 * i]};
    return MCV_OK;
}

mcv_status mcv_knn2_firstparty(const uint8_t* q, int nq, const uint8_t* t, int nt, mcv_dmatch* out) {
    if (nq < 0 || nt < 0 || (nq > 0 && (!q || !out)) || (nt > 0 && !t)) return MCV_ERR_BAD_ARG;
    std::vector<int32_t> idx, dist;
    mcv_status st = bf_host(q, nq, t, nt, idx, dist);
    if (st) return st;
    for (int i = 0; i < nq; ++i)
        for (int e = 0; e < 2; ++e) {
            const int j = idx[2 * i + e];  // padding: d = 999, idx = 0 (src/Matcher.cpp:258-259)
            out[2 * i + e] = j >= 0 ? mcv_dmatch{i, j, -1, (float)dist[2 * i + e]} : mcv_dmatch{i, 0, -1, 999.f};
        }
    return MCV_OK;
}

mcv_status mcv_knn2_candidates(const uint8_t* q, int nq, const uint8_t* t, int nt, const int32_t* cand_off, const int32_t* cand_idx, mcv_dmatch* out) {
    if (nq < 0 || nt < 0 || (nq > 0 && (!q || !out || !cand_off))) return MCV_ERR_BAD_ARG;
    if (nq == 0) return MCV_OK;
    const int total = cand_off[nq];
    if (total < 0 || (total > 0 && (!cand_idx || !t))) return MCV_ERR_BAD_ARG;
    for (int i = 0; i < nq; ++i) if (cand_off[i + 1] < cand_off[i] || cand_off[i + 1] - cand_off[i] >= (1 << 20)) return MCV_ERR_BAD_ARG;
    for (int i = 0; i < total; ++i) if (cand_idx[i] < 0 || cand_idx[i] >= nt) return MCV_ERR_BAD_ARG;
    MatchCtx* mc;
    mcv_status st = match_ctx(&mc);
    if (st) return st;
    cudaStream_t s = mc->stream;
    if ((st = mc->q.reserve((size_t)nq * 32))) return st;
    if ((st = mc->t.reserve(std::max<size_t>(32, (size_t)nt * 32)))) return st;
    if ((st = mc->idx.reserve((size_t)nq * 8))) return st;
    if ((st = mc->dist.reserve((size_t)nq * 8))) return st;
    if ((st = mc->off.reserve((size_t)(nq + 1) * 4))) return st;
    if ((st = mc->cidx.reserve(std::max<size_t>(4, (size_t)total * 4)))) return st;
    MCV_CUDA(cudaMemcpyAsync(mc->q.p, q, (size_t)nq * 32, cudaMemcpyHostToDevice, s));
    if (nt) MCV_CUDA(cudaMemcpyAsync(mc->t.p, t, (size_t)nt * 32, cudaMemcpyHostToDevice, s));
    MCV_CUDA(cudaMemcpyAsync(mc->off.p, cand_off, (size_t)(nq + 1) * 4, cudaMemcpyHostToDevice, s));
    if (total) MCV_CUDA(cudaMemcpyAsync(mc->cidx.p, cand_idx, (size_t)total * 4, cudaMemcpyHostToDevice, s));
    launch_knn2_candidates(mc->q.as<uint8_t>(), nq, mc->t.as<uint8_t>(), mc->off.as<int32_t>(), mc->cidx.as<int32_t>(), mc->idx.as<int32_t>(),
                           mc->dist.as<int32_t>(), s);
    MCV_CUDA(cudaGetLastError());
    std::vector<int32_t> idx((size_t)nq * 2), dist((size_t)nq * 2);
    MCV_CUDA(cudaMemcpyAsync(idx.data(), mc->idx.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaMemcpyAsync(dist.data(), mc->dist.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < nq; ++i)
        for (int e = 0; e < 2; ++e) out[2 * i + e] = mcv_dmatch{i, idx[2 * i + e], -1, (float)dist[2 * i + e]};
    return MCV_OK;
}

// The device entry points take their scratch from the device's default stream-ordered pool on the CALLER's stream. By default
// that pool hands freed memory back to the driver at the next synchronisation, which turns every call into a cudaMalloc
// (measured: ~80 us per call, 12 ms for the 288 MB of a 131072 x 2^20 shard); keep freed blocks cached instead.
static void keep_pool_cached() {
    static bool done[16] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || done[dev & 15]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[dev & 15] = true;
}

mcv_status mcv_knn2_bf_device(const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int train_offset, int32_t* d_idx, int32_t* d_dist, void* stream) {
    if (nq < 0 || nt < 0 || (nq > 0 && (!d_q || !d_idx || !d_dist)) || (nt > 0 && !d_t)) return MCV_ERR_BAD_ARG;
    if (nq == 0) return MCV_OK;
    // the partial keys live in a stream-ordered allocation on the CALLER's stream: safe with any stream from any thread
    unsigned* d_part = nullptr;
    const size_t part_bytes = knn2_bf_part_bytes(nq, nt);  // 0: the one-launch kernel takes this size, nothing to allocate
    if (part_bytes) {
        keep_pool_cached();
        MCV_CUDA(cudaMallocAsync((void**)&d_part, part_bytes, (cudaStream_t)stream));
    }
    const int rc = launch_knn2_bf(d_q, nq, d_t, nt, train_offset, d_idx, d_dist, d_part, (cudaStream_t)stream);
    const cudaError_t le = cudaGetLastError();
    if (d_part) cudaFreeAsync(d_part, (cudaStream_t)stream);
    if (rc < 0) { set_error("knn2_bf_device: train set larger than 4M rows per call"); return MCV_ERR_CAPACITY; }
    MCV_CUDA(le);
    return MCV_OK;
}

mcv_status mcv_knn2_pairs_device(const uint8_t* d_desc, const int32_t* d_counts, int n_images, int cap, const int32_t* d_pair_q, const int32_t* d_pair_t,
                                 int n_pairs, int32_t* d_idx, int32_t* d_dist, void* stream) {
    if (n_images < 1 || cap < 2 || cap > (1 << 22) || n_pairs < 0 || n_pairs > 65535 /* grid.z */ || (long long)n_images * cap > (1ll << 27) || !d_desc || !d_counts || (n_pairs > 0 && (!d_pair_q || !d_pair_t || !d_idx || !d_dist)))
        return MCV_ERR_BAD_ARG;
    if (n_pairs == 0) return MCV_OK;
    void* d_scratch = nullptr;   // stream-ordered on the CALLER's stream, like mcv_knn2_bf_device
    keep_pool_cached();
    MCV_CUDA(cudaMallocAsync(&d_scratch, knn2_tc_scratch_bytes(n_images * cap, 0, cap, cap, n_pairs), (cudaStream_t)stream));
    const int rc = launch_knn2_tc(d_desc, cap, d_desc, cap, 0, d_idx, d_dist, d_scratch, n_images, d_counts, d_pair_q, d_pair_t, n_pairs, (cudaStream_t)stream);
    const cudaError_t le = cudaGetLastError();
    cudaFreeAsync(d_scratch, (cudaStream_t)stream);
    if (rc < 0) { set_error("knn2_pairs_device: tensor map encoding failed"); return MCV_ERR_CUDA; }
    MCV_CUDA(le);
    return MCV_OK;
}

mcv_status mcv_project_match(const mcv_keypoint* kps, const uint8_t* desc, int n, int w, int hgt, const float* scale_factors, int nlevels,
                             const float* Rcw, const float* tcw, const float* intr, const float* mp_xyz, const uint8_t* mp_desc,
                             const int32_t* mp_level, int n_mp, float r_threshold, int32_t* out_idx, int32_t* out_dist, int* n_matched) {
    if (n < 0 || n_mp < 0 || w <= 0 || hgt <= 0 || nlevels < 1 || !scale_factors || !Rcw || !tcw || !intr || n >= (1 << 21)) return MCV_ERR_BAD_ARG;
    if (n_mp > 0 && (!mp_xyz || !mp_desc || !mp_level || !out_idx || !out_dist)) return MCV_ERR_BAD_ARG;
    if (n > 0 && (!kps || !desc)) return MCV_ERR_BAD_ARG;
    for (int m = 0; m < n_mp; ++m) if (mp_level[m] < 0 || mp_level[m] >= nlevels) return MCV_ERR_BAD_ARG;
    if (n_matched) *n_matched = 0;
    if (n_mp == 0) return MCV_OK;
    MatchCtx* mc;
    mcv_status st = match_ctx(&mc);
    if (st) return st;
    cudaStream_t s = mc->stream;
    DevBuf &b_kps = mc->slot[0], &b_desc = mc->slot[1], &b_f = mc->slot[2], &b_xyz = mc->slot[3], &b_mpd = mc->slot[4], &b_lvl = mc->slot[5], &b_oi = mc->slot[6], &b_od = mc->slot[7];
    DevBuf& b_grid = mc->slot[11];   // 30 x 30 cell table: 901 prefix entries | n keypoint indices
    if ((st = b_grid.reserve((size_t)(1024 + std::max(n, 1)) * 4))) return st;
    if ((st = b_kps.reserve(std::max<size_t>(28, (size_t)n * sizeof(mcv_keypoint))))) return st;
    if ((st = b_desc.reserve(std::max<size_t>(32, (size_t)n * 32)))) return st;
    if ((st = b_f.reserve((size_t)(nlevels + 16) * 4))) return st;
    if ((st = b_xyz.reserve((size_t)n_mp * 12))) return st;
    if ((st = b_mpd.reserve((size_t)n_mp * 32))) return st;
    if ((st = b_lvl.reserve((size_t)n_mp * 4))) return st;
    if ((st = b_oi.reserve((size_t)n_mp * 4))) return st;
    if ((st = b_od.reserve((size_t)n_mp * 4))) return st;
    std::vector<float> f(nlevels + 16);
    memcpy(f.data(), Rcw, 36); memcpy(f.data() + 9, tcw, 12); memcpy(f.data() + 12, intr, 16);
    memcpy(f.data() + 16, scale_factors, (size_t)nlevels * 4);
    if (n) {
        MCV_CUDA(cudaMemcpyAsync(b_kps.p, kps, (size_t)n * sizeof(mcv_keypoint), cudaMemcpyHostToDevice, s));
        MCV_CUDA(cudaMemcpyAsync(b_desc.p, desc, (size_t)n * 32, cudaMemcpyHostToDevice, s));
    }
    MCV_CUDA(cudaMemcpyAsync(b_f.p, f.data(), f.size() * 4, cudaMemcpyHostToDevice, s));
    MCV_CUDA(cudaMemcpyAsync(b_xyz.p, mp_xyz, (size_t)n_mp * 12, cudaMemcpyHostToDevice, s));
    MCV_CUDA(cudaMemcpyAsync(b_mpd.p, mp_desc, (size_t)n_mp * 32, cudaMemcpyHostToDevice, s));
    MCV_CUDA(cudaMemcpyAsync(b_lvl.p, mp_level, (size_t)n_mp * 4, cudaMemcpyHostToDevice, s));
    launch_project(b_kps.as<mcv_keypoint>(), b_desc.as<uint8_t>(), n, w, hgt, b_f.as<float>() + 16, b_f.as<float>(), b_xyz.as<float>(),
                   b_mpd.as<uint8_t>(), b_lvl.as<int32_t>(), n_mp, r_threshold, b_grid.as<int32_t>(), b_grid.as<int32_t>() + 1024, b_oi.as<int32_t>(),
                   b_od.as<int32_t>(), s);
    MCV_CUDA(cudaGetLastError());
    MCV_CUDA(cudaMemcpyAsync(out_idx, b_oi.p, (size_t)n_mp * 4, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaMemcpyAsync(out_dist, b_od.p, (size_t)n_mp * 4, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaStreamSynchronize(s));
    if (n_matched) { int c = 0; for (int m = 0; m < n_mp; ++m) c += out_idx[m] >= 0; *n_matched = c; }
    return MCV_OK;
}

mcv_status mcv_fuse_match(const mcv_keypoint* kps, const uint8_t* desc, int n, int w, int hgt, const float* level_sigma2,
                          const float* inv_level_sigma2, int nlevels, const float* Rcw, const float* tcw, const float* Ow, const float* intr,
                          const float* depth_left, float bf, const float* mp_xyz, const float* mp_normal, const uint8_t* mp_desc,
                          const int32_t* mp_level, int n_mp, int32_t* out_idx, int32_t* out_dist, int* n_matched) {
    if (n < 0 || n_mp < 0 || w <= 0 || hgt <= 0 || nlevels < 1 || !level_sigma2 || !inv_level_sigma2 || !Rcw || !tcw || !Ow || !intr || n >= (1 << 21))
        return MCV_ERR_BAD_ARG;
    if (n_mp > 0 && (!mp_xyz || !mp_normal || !mp_desc || !mp_level || !out_idx || !out_dist)) return MCV_ERR_BAD_ARG;
    if (n > 0 && (!kps || !desc || !depth_left)) return MCV_ERR_BAD_ARG;
    for (int i = 0; i < n; ++i) if (kps[i].octave < 0 || kps[i].octave >= nlevels) return MCV_ERR_BAD_ARG;
    if (n_matched) *n_matched = 0;
    if (n_mp == 0) return MCV_OK;
    MatchCtx* mc;
    mcv_status st = match_ctx(&mc);
    if (st) return st;
    cudaStream_t s = mc->stream;
    DevBuf &b_kps = mc->slot[0], &b_desc = mc->slot[1], &b_f = mc->slot[2], &b_dl = mc->slot[3], &b_xyz = mc->slot[4], &b_nrm = mc->slot[5], &b_mpd = mc->slot[6], &b_lvl = mc->slot[7], &b_oi = mc->slot[8], &b_od = mc->slot[9];
    DevBuf& b_grid = mc->slot[11];
    if ((st = b_grid.reserve((size_t)(1024 + std::max(n, 1)) * 4))) return st;
    if ((st = b_kps.reserve(std::max<size_t>(28, (size_t)n * sizeof(mcv_keypoint))))) return st;
    if ((st = b_desc.reserve(std::max<size_t>(32, (size_t)n * 32)))) return st;
    if ((st = b_dl.reserve(std::max<size_t>(4, (size_t)n * 4)))) return st;
    if ((st = b_f.reserve((size_t)(2 * nlevels + 20) * 4))) return st;
    if ((st = b_xyz.reserve((size_t)n_mp * 12))) return st;
    if ((st = b_nrm.reserve((size_t)n_mp * 12))) return st;
    if ((st = b_mpd.reserve((size_t)n_mp * 32))) return st;
    if ((st = b_lvl.reserve((size_t)n_mp * 4))) return st;
    if ((st = b_oi.reserve((size_t)n_mp * 4))) return st;
    if ((st = b_od.reserve((size_t)n_mp * 4))) return st;
    std::vector<float> f(2 * nlevels + 20);
    memcpy(f.data(), Rcw, 36); memcpy(f.data() + 9, tcw, 12); memcpy(f.data() + 12, intr, 16); memcpy(f.data() + 16, Ow, 12);
    f[19] = bf;
    memcpy(f.data() + 20, level_sigma2, (size_t)nlevels * 4); memcpy(f.data() + 20 + nlevels, inv_level_sigma2, (size_t)nlevels * 4);
    if (n) {
        MCV_CUDA(cudaMemcpyAsync(b_kps.p, kps, (size_t)n * sizeof(mcv_keypoint), cudaMemcpyHostToDevice, s));
        MCV_CUDA(cudaMemcpyAsync(b_desc.p, desc, (size_t)n * 32, cudaMemcpyHostToDevice, s));
        MCV_CUDA(cudaMemcpyAsync(b_dl.p, depth_left, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    }
    MCV_CUDA(cudaMemcpyAsync(b_f.p, f.data(), f.size() * 4, cudaMemcpyHostToDevice, s));
    MCV_CUDA(cudaMemcpyAsync(b_xyz.p, mp_xyz, (size_t)n_mp * 12, cudaMemcpyHostToDevice, s));
    MCV_CUDA(cudaMemcpyAsync(b_nrm.p, mp_normal, (size_t)n_mp * 12, cudaMemcpyHostToDevice, s));
    MCV_CUDA(cudaMemcpyAsync(b_mpd.p, mp_desc, (size_t)n_mp * 32, cudaMemcpyHostToDevice, s));
    MCV_CUDA(cudaMemcpyAsync(b_lvl.p, mp_level, (size_t)n_mp * 4, cudaMemcpyHostToDevice, s));
    launch_fuse_match(b_kps.as<mcv_keypoint>(), b_desc.as<uint8_t>(), n, w, hgt, b_f.as<float>(), nlevels, b_dl.as<float>(), b_xyz.as<float>(),
                      b_nrm.as<float>(), b_mpd.as<uint8_t>(), b_lvl.as<int32_t>(), n_mp, b_grid.as<int32_t>(), b_grid.as<int32_t>() + 1024, b_oi.as<int32_t>(),
                      b_od.as<int32_t>(), s);
    MCV_CUDA(cudaGetLastError());
    MCV_CUDA(cudaMemcpyAsync(out_idx, b_oi.p, (size_t)n_mp * 4, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaMemcpyAsync(out_dist, b_od.p, (size_t)n_mp * 4, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaStreamSynchronize(s));
    if (n_matched) { int c = 0; for (int m = 0; m < n_mp; ++m) c += out_idx[m] >= 0; *n_matched = c; }
    return MCV_OK;
}

mcv_status mcv_wnd_track(const mcv_keypoint* kps1, const uint8_t* desc1, int n1, const int32_t* q_idx, int n_q, const mcv_keypoint* kps2,
                         const uint8_t* desc2, int n2, int w, int hgt, int32_t* out_idx, int32_t* out_best, int32_t* out_dist, int* n_matched) {
    if (n1 < 0 || n2 < 0 || n_q < 0 || w <= 0 || hgt <= 0 || n2 >= (1 << 21)) return MCV_ERR_BAD_ARG;
    if (n_q > 0 && (!kps1 || !desc1 || !q_idx || !out_idx || !out_best || !out_dist)) return MCV_ERR_BAD_ARG;
    if (n2 > 0 && (!kps2 || !desc2)) return MCV_ERR_BAD_ARG;
    for (int q = 0; q < n_q; ++q) if (q_idx[q] < 0 || q_idx[q] >= n1) return MCV_ERR_BAD_ARG;
    if (n_matched) *n_matched = 0;
    if (n_q == 0) return MCV_OK;
    MatchCtx* mc;
    mcv_status st = match_ctx(&mc);
    if (st) return st;
    cudaStream_t s = mc->stream;
    DevBuf &b_k1 = mc->slot[0], &b_d1 = mc->slot[1], &b_q = mc->slot[2], &b_k2 = mc->slot[3], &b_d2 = mc->slot[4], &b_oi = mc->slot[5], &b_ob = mc->slot[6], &b_od = mc->slot[7];
    DevBuf& b_grid = mc->slot[11];
    if ((st = b_grid.reserve((size_t)(1024 + std::max(n2, 1)) * 4))) return st;
    if ((st = b_k1.reserve((size_t)n1 * sizeof(mcv_keypoint)))) return st;
    if ((st = b_d1.reserve((size_t)n1 * 32))) return st;
    if ((st = b_q.reserve((size_t)n_q * 4))) return st;
    if ((st = b_k2.reserve(std::max<size_t>(28, (size_t)n2 * sizeof(mcv_keypoint))))) return st;
    if ((st = b_d2.reserve(std::max<size_t>(32, (size_t)n2 * 32)))) return st;
    if ((st = b_oi.reserve((size_t)n_q * 4))) return st;
    if ((st = b_ob.reserve((size_t)n_q * 4))) return st;
    if ((st = b_od.reserve((size_t)n_q * 4))) return st;
    MCV_CUDA(cudaMemcpyAsync(b_k1.p, kps1, (size_t)n1 * sizeof(mcv_keypoint), cudaMemcpyHostToDevice, s));
    MCV_CUDA(cudaMemcpyAsync(b_d1.p, desc1, (size_t)n1 * 32, cudaMemcpyHostToDevice, s));
    MCV_CUDA(cudaMemcpyAsync(b_q.p, q_idx, (size_t)n_q * 4, cudaMemcpyHostToDevice, s));
    if (n2) {
        MCV_CUDA(cudaMemcpyAsync(b_k2.p, kps2, (size_t)n2 * sizeof(mcv_keypoint), cudaMemcpyHostToDevice, s));
        MCV_CUDA(cudaMemcpyAsync(b_d2.p, desc2, (size_t)n2 * 32, cudaMemcpyHostToDevice, s));
    }
    launch_wnd_track(b_k1.as<mcv_keypoint>(), b_d1.as<uint8_t>(), b_q.as<int32_t>(), n_q, b_k2.as<mcv_keypoint>(), b_d2.as<uint8_t>(), n2, w, hgt,
                     b_grid.as<int32_t>(), b_grid.as<int32_t>() + 1024, b_oi.as<int32_t>(), b_ob.as<int32_t>(), b_od.as<int32_t>(), s);
    MCV_CUDA(cudaGetLastError());
    MCV_CUDA(cudaMemcpyAsync(out_idx, b_oi.p, (size_t)n_q * 4, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaMemcpyAsync(out_best, b_ob.p, (size_t)n_q * 4, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaMemcpyAsync(out_dist, b_od.p, (size_t)n_q * 4, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaStreamSynchronize(s));
    if (n_matched) { int c = 0; for (int q = 0; q < n_q; ++q) c += out_idx[q] >= 0; *n_matched = c; }
    return MCV_OK;
}

// ---- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cpp:101-150), batched over MapPoints ----
mcv_status mcv_distinctive_descriptors(const uint8_t* desc, const int32_t* mp_off, int n_mp, int32_t* best_idx, int32_t* best_median,
                                       uint8_t* out_desc) {
    if (n_mp < 0 || (n_mp > 0 && (!mp_off || !best_idx))) return MCV_ERR_BAD_ARG;
    if (n_mp == 0) return MCV_OK;
    if (mp_off[0] != 0) return MCV_ERR_BAD_ARG;
    for (int m = 0; m < n_mp; ++m) if (mp_off[m + 1] < mp_off[m] || mp_off[m + 1] - mp_off[m] >= (1 << 22)) return MCV_ERR_BAD_ARG;
    const int total = mp_off[n_mp];
    if (total > 0 && !desc) return MCV_ERR_BAD_ARG;
    MatchCtx* mc;
    mcv_status st = match_ctx(&mc);
    if (st) return st;
    cudaStream_t s = mc->stream;
    DevBuf &b_d = mc->slot[0], &b_off = mc->slot[1], &b_bi = mc->slot[2], &b_bm = mc->slot[3];
    if ((st = b_d.reserve(std::max<size_t>(32, (size_t)total * 32)))) return st;
    if ((st = b_off.reserve((size_t)(n_mp + 1) * 4))) return st;
    if ((st = b_bi.reserve((size_t)n_mp * 4))) return st;
    if ((st = b_bm.reserve((size_t)n_mp * 4))) return st;
    if (total) MCV_CUDA(cudaMemcpyAsync(b_d.p, desc, (size_t)total * 32, cudaMemcpyHostToDevice, s));
    MCV_CUDA(cudaMemcpyAsync(b_off.p, mp_off, (size_t)(n_mp + 1) * 4, cudaMemcpyHostToDevice, s));
    launch_distinctive(b_d.as<uint8_t>(), b_off.as<int32_t>(), n_mp, b_bi.as<int32_t>(), b_bm.as<int32_t>(), s);
    MCV_CUDA(cudaGetLastError());
    std::vector<int32_t> med_tmp;
    if (!best_median) { med_tmp.resize(n_mp); best_median = med_tmp.data(); }
    MCV_CUDA(cudaMemcpyAsync(best_idx, b_bi.p, (size_t)n_mp * 4, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaMemcpyAsync(best_median, b_bm.p, (size_t)n_mp * 4, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaStreamSynchronize(s));
    if (out_desc)   // MapPoint::desp = all_ob_desps[BestIdx].clone(); points without observations keep their row untouched
        for (int m = 0; m < n_mp; ++m)
            if (best_idx[m] >= 0) memcpy(out_desc + (size_t)m * 32, desc + (size_t)(mp_off[m] + best_idx[m]) * 32, 32);
    return MCV_OK;
}

// ---- KL_Track (src/Frame.cpp:34-76): cv::calcOpticalFlowPyrLK with the reference's fixed arguments ----
mcv_status mcv_lk_track_batch(const uint8_t* prev, const uint8_t* next, int n_pairs, int w, int hgt, size_t stride, const float* pts,
                              const int32_t* pt_off, float* next_pts, uint8_t* status, float* err) {
    if (!prev || !next || n_pairs <= 0 || !pt_off) return MCV_ERR_BAD_ARG;
    if (w <= 0 || hgt <= 0) return MCV_ERR_EMPTY_IMAGE;
    if (stride < (size_t)w || pt_off[0] != 0) return MCV_ERR_BAD_ARG;
    for (int k = 0; k < n_pairs; ++k) if (pt_off[k + 1] < pt_off[k]) return MCV_ERR_BAD_ARG;
    const int n = pt_off[n_pairs];
    if (n > 0 && (!pts || !next_pts || !status || !err)) return MCV_ERR_BAD_ARG;
    MatchCtx* mc;
    mcv_status st = match_ctx(&mc);
    if (st) return st;
    cudaStream_t s = mc->stream;
    DevBuf &b_img = mc->slot[0], &b_ws = mc->slot[1], &b_pts = mc->slot[2], &b_pair = mc->slot[3], &b_out = mc->slot[4], &b_st = mc->slot[5], &b_err = mc->slot[6];
    const size_t img_bytes = (size_t)w * hgt;
    if ((st = b_img.reserve(2 * img_bytes * n_pairs))) return st;
    if ((st = b_ws.reserve(lk_workspace_bytes(w, hgt) * n_pairs))) return st;
    if ((st = b_pts.reserve(std::max<size_t>(8, (size_t)n * 8)))) return st;
    if ((st = b_pair.reserve(std::max<size_t>(4, (size_t)n * 4)))) return st;
    if ((st = b_out.reserve(std::max<size_t>(8, (size_t)n * 8)))) return st;
    if ((st = b_st.reserve(std::max<size_t>(4, (size_t)n)))) return st;
    if ((st = b_err.reserve(std::max<size_t>(4, (size_t)n * 4)))) return st;
    uint8_t* d_prev = b_img.as<uint8_t>();
    uint8_t* d_next = d_prev + img_bytes * n_pairs;
    MCV_CUDA(cudaMemcpy2DAsync(d_prev, w, prev, stride, w, (size_t)hgt * n_pairs, cudaMemcpyHostToDevice, s));   // images back to back: one 2-D copy
    MCV_CUDA(cudaMemcpy2DAsync(d_next, w, next, stride, w, (size_t)hgt * n_pairs, cudaMemcpyHostToDevice, s));
    std::vector<int32_t> pair_of;
    if (n) {
        MCV_CUDA(cudaMemcpyAsync(b_pts.p, pts, (size_t)n * 8, cudaMemcpyHostToDevice, s));
        if (n_pairs > 1) {
            pair_of.resize((size_t)n);
            for (int k = 0; k < n_pairs; ++k) std::fill(pair_of.begin() + pt_off[k], pair_of.begin() + pt_off[k + 1], k);
            MCV_CUDA(cudaMemcpyAsync(b_pair.p, pair_of.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
        }
    }
    launch_lk_track(d_prev, d_next, n_pairs, w, hgt, w, img_bytes, b_ws.p, b_pts.as<float>(), n_pairs > 1 ? b_pair.as<int>() : nullptr, n, b_out.as<float>(),
                    b_st.as<uint8_t>(), b_err.as<float>(), s);
    MCV_CUDA(cudaGetLastError());
    if (n) {
        MCV_CUDA(cudaMemcpyAsync(next_pts, b_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        MCV_CUDA(cudaMemcpyAsync(status, b_st.p, (size_t)n, cudaMemcpyDeviceToHost, s));
        MCV_CUDA(cudaMemcpyAsync(err, b_err.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    }
    MCV_CUDA(cudaStreamSynchronize(s));   // (also keeps pair_of alive until its copy is done)
    return MCV_OK;
}

mcv_status mcv_lk_track(const uint8_t* prev, const uint8_t* next, int w, int hgt, size_t stride, const float* pts, int n, float* next_pts,
                        uint8_t* status, float* err) {
    if (n < 0) return MCV_ERR_BAD_ARG;
    const int32_t off[2] = {0, n};
    return mcv_lk_track_batch(prev, next, 1, w, hgt, stride, pts, off, next_pts, status, err);
}

mcv_status mcv_kl_track(const uint8_t* prev, const uint8_t* next, int w, int hgt, size_t stride, const mcv_keypoint* kps, int n, mcv_keypoint* new_kps,
                        uint8_t* ok, int* n_ok) {
    if (n < 0 || (n > 0 && (!kps || !new_kps || !ok))) return MCV_ERR_BAD_ARG;
    if (n_ok) *n_ok = 0;
    for (int i = 0; i < n; ++i) ok[i] = 0;
    if (n < 10) return (!prev || !next) ? MCV_ERR_BAD_ARG : MCV_OK;     // src/Frame.cpp:41
    std::vector<float> pts((size_t)2 * n), nxt((size_t)2 * n), err((size_t)n);
    std::vector<uint8_t> status((size_t)n);
    for (int i = 0; i < n; ++i) { pts[2 * i] = kps[i].x; pts[2 * i + 1] = kps[i].y; }
    mcv_status st = mcv_lk_track(prev, next, w, hgt, stride, pts.data(), n, nxt.data(), status.data(), err.data());
    if (st) return st;
    int cnt = 0;
    for (int i = 0; i < n; ++i)
        if (status[i] > 0 && err[i] < 1) {                                 // src/Frame.cpp:57-58
            ok[i] = 1; ++cnt;
            new_kps[i] = kps[i]; new_kps[i].x = nxt[2 * i]; new_kps[i].y = nxt[2 * i + 1]; new_kps[i].octave = 0;   // :65-68
        }
    if (n_ok) *n_ok = cnt;
    return MCV_OK;
}

// ---- Object::ComputeBow: DBoW3 vocabulary on the device + Vocabulary::transform ----
struct mcv_voc {
    int device = 0, n_nodes = 0, L = 0, weighting = 0, norm = 1;
    DevBuf child_off, child_ids, node_desc;
    std::vector<int32_t> word_id, h_child_off;
    std::vector<double> weight;
};

mcv_status mcv_voc_create(int n_nodes, const int32_t* child_off, const uint32_t* child_ids, const uint8_t* node_desc, const int32_t* word_id,
                          const double* weight, int L, int weighting, int norm, int device, mcv_voc** out) {
    if (!out) return MCV_ERR_BAD_ARG;
    *out = nullptr;
    if (n_nodes < 1 || !child_off || !node_desc || !word_id || !weight || L < 0 || weighting < 0 || weighting > 3 || norm < 0 || norm > 2) return MCV_ERR_BAD_ARG;
    if (child_off[0] != 0) return MCV_ERR_BAD_ARG;
    for (int i = 0; i < n_nodes; ++i) if (child_off[i + 1] < child_off[i] || child_off[i + 1] - child_off[i] >= (1 << 20)) return MCV_ERR_BAD_ARG;
    const int n_edges = child_off[n_nodes];
    if (n_edges > 0 && !child_ids) return MCV_ERR_BAD_ARG;
    for (int e = 0; e < n_edges; ++e) if (child_ids[e] == 0 || child_ids[e] >= (uint32_t)n_nodes) return MCV_ERR_BAD_ARG;
    if (mcv_device_count() <= device || device < 0) { set_error("no such CUDA device"); return MCV_ERR_NO_DEVICE; }
    MCV_CUDA(cudaSetDevice(device));
    mcv_voc* v = new mcv_voc();
    v->device = device; v->n_nodes = n_nodes; v->L = L; v->weighting = weighting; v->norm = norm;
    v->word_id.assign(word_id, word_id + n_nodes); v->weight.assign(weight, weight + n_nodes); v->h_child_off.assign(child_off, child_off + n_nodes + 1);
    mcv_status st;
    if ((st = v->child_off.reserve((size_t)(n_nodes + 1) * 4)) || (st = v->child_ids.reserve(std::max<size_t>(4, (size_t)n_edges * 4))) ||
        (st = v->node_desc.reserve((size_t)n_nodes * 32))) { delete v; return st; }
    cudaMemcpy(v->child_off.p, child_off, (size_t)(n_nodes + 1) * 4, cudaMemcpyHostToDevice);
    if (n_edges) cudaMemcpy(v->child_ids.p, child_ids, (size_t)n_edges * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(v->node_desc.p, node_desc, (size_t)n_nodes * 32, cudaMemcpyHostToDevice);
    MCV_CUDA(cudaGetLastError());
    *out = v;
    return MCV_OK;
}

void mcv_voc_destroy(mcv_voc* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    for (DevBuf* b : {&v->child_off, &v->child_ids, &v->node_desc}) b->release();
    delete v;
}

mcv_status mcv_bow_transform(mcv_voc* v, const uint8_t* desc, int n, int levelsup, int32_t* out_word, double* out_weight, uint32_t* out_nid,
                             uint32_t* bow_ids, double* bow_vals, int* n_bow, uint32_t* fv_nodes, int32_t* fv_off, int32_t* fv_idx, int* n_fv) {
    if (!v || n < 0 || !n_bow || !n_fv || (n > 0 && (!desc || !bow_ids || !bow_vals || !fv_nodes || !fv_off || !fv_idx))) return MCV_ERR_BAD_ARG;
    *n_bow = 0; *n_fv = 0;
    if (fv_off) fv_off[0] = 0;
    if (n == 0) return MCV_OK;
    MCV_CUDA(cudaSetDevice(v->device));
    MatchCtx* mc;
    mcv_status st = match_ctx(&mc);
    if (st) return st;
    cudaStream_t s = mc->stream;
    DevBuf &b_feat = mc->slot[0], &b_leaf = mc->slot[1], &b_nid = mc->slot[2];   // per calling thread: two threads may share one vocabulary
    if ((st = b_feat.reserve((size_t)n * 32)) || (st = b_leaf.reserve((size_t)n * 4)) || (st = b_nid.reserve((size_t)n * 4))) return st;
    MCV_CUDA(cudaMemcpyAsync(b_feat.p, desc, (size_t)n * 32, cudaMemcpyHostToDevice, s));
    launch_bow_descend(b_feat.as<uint8_t>(), n, v->child_off.as<int32_t>(), v->child_ids.as<uint32_t>(), v->node_desc.as<uint8_t>(), v->L - levelsup,
                       std::max(v->L, 1) + 32, b_leaf.as<uint32_t>(), b_nid.as<uint32_t>(), s);
    MCV_CUDA(cudaGetLastError());
    std::vector<uint32_t> leaf(n), nid(n);
    MCV_CUDA(cudaMemcpyAsync(leaf.data(), b_leaf.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaMemcpyAsync(nid.data(), b_nid.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaStreamSynchronize(s));
    // BowVector / FeatureVector assembly (Vocabulary.cpp:586-631, BowVector.cpp, FeatureVector.cpp): std::map insertions in
    // feature order — the order the weights of one word are summed in is part of the (double) result
    std::map<uint32_t, double> bow;
    std::map<uint32_t, std::vector<int32_t>> fv;
    const bool tf = v->weighting == 0 || v->weighting == 1;
    for (int i = 0; i < n; ++i) {
        if (v->h_child_off[leaf[i]] != v->h_child_off[leaf[i] + 1]) { set_error("bow_transform: descent ended on an inner node (malformed vocabulary)"); return MCV_ERR_BAD_ARG; }
        const uint32_t wid = (uint32_t)v->word_id[leaf[i]];
        const double w = v->weight[leaf[i]];
        if (out_word) out_word[i] = (int32_t)wid;
        if (out_weight) out_weight[i] = w;
        if (out_nid) out_nid[i] = nid[i];
        if (!(w > 0)) continue;                        // stopped word
        auto it = bow.lower_bound(wid);
        if (it != bow.end() && !(wid < it->first)) { if (tf) it->second += w; }
        else bow.insert(it, std::make_pair(wid, w));
        fv[nid[i]].push_back(i);
    }
    if (tf && !bow.empty() && v->norm == 0) {
        const double nd = (double)bow.size();
        for (auto& e : bow) e.second /= nd;
    }
    if (v->norm != 0) {
        double norm = 0.0;
        if (v->norm == 1) { for (auto& e : bow) norm += fabs(e.second); }
        else { for (auto& e : bow) norm += e.second * e.second; norm = sqrt(norm); }
        if (norm > 0.0) for (auto& e : bow) e.second /= norm;
    }
    int k = 0;
    for (auto& e : bow) { bow_ids[k] = e.first; bow_vals[k] = e.second; ++k; }
    int m = 0, at = 0;
    for (auto& e : fv) { fv_nodes[m] = e.first; fv_off[m] = at; for (int32_t f : e.second) fv_idx[at++] = f; ++m; }
    fv_off[m] = at;
    *n_bow = k; *n_fv = m;
    return MCV_OK;
}

mcv_status mcv_debug_sincosf(const float* a, int n, float* so, float* co) {
    if (n <= 0 || !a || !so || !co) return MCV_ERR_BAD_ARG;
    MatchCtx* mc;
    mcv_status st = match_ctx(&mc);
    if (st) return st;
    cudaStream_t s = mc->stream;
    DevBuf &b = mc->slot[0];
    if ((st = b.reserve((size_t)n * 12))) return st;
    float* d = b.as<float>();
    MCV_CUDA(cudaMemcpyAsync(d, a, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    launch_debug_sincosf(d, n, d + n, d + 2 * (size_t)n, s);
    MCV_CUDA(cudaGetLastError());
    MCV_CUDA(cudaMemcpyAsync(so, d + n, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaMemcpyAsync(co, d + 2 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaStreamSynchronize(s));
    return MCV_OK;
}

mcv_status mcv_debug_fast_atan2(const float* y, const float* x, int n, float* out) {
    if (n <= 0 || !y || !x || !out) return MCV_ERR_BAD_ARG;
    MatchCtx* mc;
    mcv_status st = match_ctx(&mc);
    if (st) return st;
    cudaStream_t s = mc->stream;
    DevBuf &b = mc->slot[0];
    if ((st = b.reserve((size_t)n * 12))) return st;
    float* d = b.as<float>();
    MCV_CUDA(cudaMemcpyAsync(d, y, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    MCV_CUDA(cudaMemcpyAsync(d + n, x, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    launch_debug_atan2(d, d + n, n, d + 2 * (size_t)n, s);
    MCV_CUDA(cudaGetLastError());
    MCV_CUDA(cudaMemcpyAsync(out, d + 2 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaStreamSynchronize(s));
    return MCV_OK;
}

mcv_status mcv_debug_octree_clocks(long long* out8) {
    if (!out8) return MCV_ERR_BAD_ARG;
    if (octree_debug_clocks(out8)) { set_error("cudaMemcpyFromSymbol failed"); return MCV_ERR_CUDA; }
    return MCV_OK;
}

mcv_status mcv_debug_popc_peak(int iters, double* popc_per_s, double* ms_out) {
    if (iters <= 0) return MCV_ERR_BAD_ARG;
    MatchCtx* mc;
    mcv_status st = match_ctx(&mc);
    if (st) return st;
    cudaStream_t s = mc->stream;
    DevBuf &b = mc->slot[0];
    if ((st = b.reserve(64))) return st;
    const int blocks = NUM_SMS * 8, threads = 256;
    launch_popc_peak(iters, b.as<unsigned>(), blocks, threads, s);  // warm-up
    cudaEvent_t e0, e1;
    MCV_CUDA(cudaEventCreate(&e0)); MCV_CUDA(cudaEventCreate(&e1));
    MCV_CUDA(cudaEventRecord(e0, s));
    launch_popc_peak(iters, b.as<unsigned>(), blocks, threads, s);
    MCV_CUDA(cudaEventRecord(e1, s));
    MCV_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    MCV_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (ms_out) *ms_out = ms;
    if (popc_per_s) *popc_per_s = (double)blocks * threads * 8.0 * iters / (ms * 1e-3);
    return MCV_OK;
}

}  // extern "C"

// =========================================================================================================
// three-camera rig
//
// A batch is cut into chunks of `chunk_frames` triplets that run round-robin on RIG_SLOTS slots. Every slot owns a stream
// and a full extractor workspace, so within ONE synchronous call the host->device copy of chunk i+1, the kernels of chunk i
// and the device->host copy of chunk i-1 overlap, and the latency-bound quadtree kernel of one chunk shares the SMs with
// the throughput-bound stencils of another (CUDA streams replace the reference's ThreadPool(3), src/Frame.cpp:22).
// =========================================================================================================
#ifndef MCV_RIG_SLOTS_N
#define MCV_RIG_SLOTS_N 6
#endif
constexpr int RIG_SLOTS = MCV_RIG_SLOTS_N;
constexpr int RIG_TICKETS = 8;

// A chunk's whole kernel sequence (about 25 launches and memsets) captured once per shape and pointer set and replayed with one
// cudaGraphLaunch: small chunks are bound by the host's launch rate, not by the GPU (B200: 128 frames as 8 chunks of 16 through
// mcv_rig_process = 5.6 ms with direct launches).
struct ChunkKey {
    int n_frames, w, h, cap, channels; unsigned epoch;
    const void *imgs, *kps, *desc, *counts, *ur, *dp, *best, *scratch;
    bool operator==(const ChunkKey& o) const { return memcmp(this, &o, sizeof(ChunkKey)) == 0; }
};
struct ChunkGraph { ChunkKey key; cudaGraphExec_t exec = nullptr; int launches = 0; bool warm_only = true; };

struct RigSlot {
    mcv_orb* orb = nullptr;
    DevBuf imgs, kps, desc, counts, u_right, depth, best_dist, st_scratch;
    cudaEvent_t done = nullptr, front = nullptr;   // all work of the last chunk / its front half
    std::vector<ChunkGraph> graphs;                // most recently used last; at most RIG_GRAPHS_PER_SLOT
};
constexpr int RIG_GRAPHS_PER_SLOT = 8;

struct mcv_rig {
    mcv_rig_params prm{};
    int device = 0;
    cudaStream_t stream = nullptr;   // the caller-visible stream: async work is ordered after / joined back into it
    bool own_stream = false;
    RigSlot slot[RIG_SLOTS];
    cudaEvent_t fork = nullptr;
    int chunk_frames = 32;          // host path: chunks pipeline H2D / kernels / D2H (capped at ceil(n / slots) so that every slot gets work)
    int chunk_frames_dev = 128;     // device-resident path: consecutive calls overlap instead (see mcv_rig_process_async)
    int next_slot = 0;              // chunks rotate over the slots across calls
    int use_slots = RIG_SLOTS;      // synchronous host path: slots in rotation (env MCV_RIG_SLOTS). B200, 128 frames per mcv_rig_process call:
                                    // 3 slots 25.1 k frames/s, 4 slots 27.7 k, 6 slots 29.4 k (a chunk no longer waits for the D2H of the chunk
                                    // that used its slot before); mcv_rig_submit keeps whole steps in flight and rotates over three
    int use_slots_submit = 4;       // mcv_rig_submit: whole steps in flight rotate over this many slots (env MCV_RIG_SLOTS_SUBMIT). A step is
                                    // H2D -> kernels -> D2H on its slot's stream; with three steps in flight one of them is always in a
                                    // copy, so only two overlap their kernels. B200, 128-frame steps: 3 slots / 3 in flight 40.9 k
                                    // frames/s, 4 / 4: 42.7 k, 5 / 5: 42.7 k, 6 / 8: 42.8 k (device-resident 45.7 k)
    int use_slots_dev = 3;          // device-resident path (env MCV_RIG_SLOTS_DEV): consecutive calls rotate over three
                                    // streams, so the latency-bound quadtree of one batch runs beside the stencils of the next
                                    // ones. B200, 128-frame calls: BASELINE config one stream 36.4 k frames/s, two 41.1 k, three
                                    // 41.9 k, four 39.8 k; the reference's shipped single-level 512 x 512 config (the quadtree is
                                    // 1.7 of a call's 1.9 ms) two 67.5 k, three 81.1 k, four 81.9 k, six 53.0 k. A dedicated
                                    // high-priority quadtree stream (MCV_RIG_QUAD_PRIORITY=1) changes nothing.
    cudaEvent_t last_front = nullptr;   // front-half event of the most recently enqueued chunk
    bool no_stagger = true;             // the slots' streams run free; env MCV_RIG_STAGGER=1 makes a chunk's front half wait for
                                        // the previous chunk's (B200, current kernels: free 38.5k frames/s, staggered 37.3k)
    cudaEvent_t ticket[RIG_TICKETS] = {};   // completion of the last RIG_TICKETS mcv_rig_submit calls
    long long submitted = 0;                // number of mcv_rig_submit calls so far (ticket ids start at 1)
    int submit_chunk = 128;                 // frames per chunk of mcv_rig_submit (env MCV_RIG_SUBMIT_CHUNK); B200, 3 steps in flight: 128 -> 35.7k frames/s, 64 -> 34.7k, 32 -> 29.2k
    bool pending_join = false;
    int last_launches = 0;
    int graph_max_frames = 32;              // chunks of at most this many frames replay a captured CUDA graph (env MCV_RIG_GRAPH_MAX, 0 = never)
};

// ORBE + SMatch of n_frames device-resident triplets on one slot (its stream); all pointers are device pointers.
static mcv_status rig_chunk_enqueue(mcv_rig* r, RigSlot& sl, const uint8_t* d_imgs, int n_frames, int w, int hgt, mcv_keypoint* d_kps,
                                    uint8_t* d_desc, int32_t* d_counts, float* d_u_right, float* d_depth, int cap, int* launches, bool stagger, bool events) {
    mcv_orb* h = sl.orb;
    mcv_status st = enqueue_extract(h, d_imgs, (size_t)w * h->channels, (size_t)w * hgt * h->channels, 3 * n_frames, nullptr, d_kps, d_desc, d_counts, cap,
                                    events && stagger && !r->no_stagger && r->last_front != sl.front ? r->last_front : nullptr, events ? sl.front : nullptr);
    if (st) return st;
    if (events) r->last_front = sl.front;
    int n = h->last_launches;
    cudaEvent_t mid = (h->profile && h->prof_calls < PROF_RING) ? h->ev[(size_t)h->prof_calls * (N_STAGES + 1) + 7] : nullptr;
    n += launch_stereo(h->plan, h->pyr.as<uint8_t>(), d_kps, d_desc, d_counts, cap, n_frames, 0, 1, 3, r->prm.bf, r->prm.baseline, d_u_right,
                       d_depth, sl.best_dist.as<int>(), nullptr, sl.st_scratch.p, h->stream, mid);
    n += launch_fill_tails(d_kps, d_desc, d_counts, cap, 3 * n_frames, d_u_right, d_depth, 3, h->stream);   // slots behind the counts: defined values
    prof_mark(h, 8);
    if (h->profile && h->prof_calls < PROF_RING) ++h->prof_calls;
    MCV_CUDA(cudaGetLastError());
    *launches += n;
    return MCV_OK;
}

static mcv_status rig_chunk(mcv_rig* r, RigSlot& sl, const uint8_t* d_imgs, int n_frames, int w, int hgt, mcv_keypoint* d_kps,
                            uint8_t* d_desc, int32_t* d_counts, float* d_u_right, float* d_depth, int cap, int* launches,
                            bool stagger = true) {
    mcv_orb* h = sl.orb;
    mcv_status st = ensure_workspace(h, w, hgt, 3 * n_frames, 1);
    if (st) return st;
    if (cap < h->plan.max_quad_kp) { set_error("cap smaller than mcv_rig_max_keypoints_for(w, h)"); return MCV_ERR_CAPACITY; }
    if ((st = sl.best_dist.reserve((size_t)n_frames * cap * 4))) return st;
    if ((st = sl.st_scratch.reserve(stereo_scratch_bytes(h->plan, n_frames, cap)))) return st;
    const bool graph_ok = n_frames <= r->graph_max_frames && !h->profile && !h->quad_stream && r->no_stagger;
    if (!graph_ok) return rig_chunk_enqueue(r, sl, d_imgs, n_frames, w, hgt, d_kps, d_desc, d_counts, d_u_right, d_depth, cap, launches, stagger, true);
    const ChunkKey key{n_frames, w, hgt, cap, h->channels, h->ws_epoch, d_imgs, d_kps, d_desc, d_counts, d_u_right, d_depth, sl.best_dist.p, sl.st_scratch.p};
    for (size_t i = 0; i < sl.graphs.size(); ++i) {
        if (!(sl.graphs[i].key == key)) continue;
        ChunkGraph g = sl.graphs[i];
        sl.graphs.erase(sl.graphs.begin() + i);
        if (g.warm_only) {
            // second call with this shape and pointer set: every lazy allocation and function attribute is settled -> capture
            cudaGraph_t graph = nullptr;
            int n = 0;
            if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); r->graph_max_frames = 0; break; }
            st = rig_chunk_enqueue(r, sl, d_imgs, n_frames, w, hgt, d_kps, d_desc, d_counts, d_u_right, d_depth, cap, &n, false, false);
            const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
            if (st || ce != cudaSuccess || !graph || cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) {
                cudaGetLastError();
                if (graph) cudaGraphDestroy(graph);
                r->graph_max_frames = 0;         // capture is not available here: direct launches from now on
                if (st) return st;
                break;
            }
            cudaGraphDestroy(graph);
            g.launches = n; g.warm_only = false;
        }
        MCV_CUDA(cudaGraphLaunch(g.exec, h->stream));
        *launches += g.launches;
        h->last_images = 3 * n_frames; h->last_cap = cap; h->last_launches = g.launches;
        sl.graphs.push_back(g);
        return MCV_OK;
    }
    if (r->graph_max_frames > 0) {
        if ((int)sl.graphs.size() >= RIG_GRAPHS_PER_SLOT) { if (sl.graphs[0].exec) cudaGraphExecDestroy(sl.graphs[0].exec); sl.graphs.erase(sl.graphs.begin()); }
        ChunkGraph g; g.key = key;
        sl.graphs.push_back(g);
    }
    return rig_chunk_enqueue(r, sl, d_imgs, n_frames, w, hgt, d_kps, d_desc, d_counts, d_u_right, d_depth, cap, launches, stagger, true);
}

// Host-buffer path: per chunk H2D -> kernels -> D2H on the slot's stream (all asynchronous), chunks rotating over the slots.
static mcv_status rig_enqueue_host(mcv_rig* r, const uint8_t* imgs, int n_frames, int w, int hgt, int imgs_on_device, mcv_keypoint* kps_out,
                                   uint8_t* desc_out, int32_t* counts, float* u_right, float* depth_left, int cap, int out_on_device, int chunk, int n_slots,
                                   const std::vector<int>* plan = nullptr) {
    const size_t img3 = (size_t)3 * w * hgt * r->slot[0].orb->channels, kb = sizeof(mcv_keypoint);
    int launches = 0;
    size_t pi = 0;
#ifdef MCV_EXPERIMENTS
    // MCV_RIG_TRACE=1 (experiment builds): per chunk the times its H2D started / ended, its kernels ended and its D2H ended,
    // relative to the first chunk's start — printed by the next call
    static const bool trace = getenv("MCV_RIG_TRACE") != nullptr;
    static std::vector<cudaEvent_t> tev;
    static std::vector<int> tnf;
    if (trace && !tev.empty()) {
        cudaDeviceSynchronize();
        for (size_t c = 0; c < tnf.size(); ++c) {
            float a, b, k, d;
            cudaEventElapsedTime(&a, tev[0], tev[4 * c]); cudaEventElapsedTime(&b, tev[0], tev[4 * c + 1]);
            cudaEventElapsedTime(&k, tev[0], tev[4 * c + 2]); cudaEventElapsedTime(&d, tev[0], tev[4 * c + 3]);
            fprintf(stderr, "chunk %zu (%d frames): h2d %.3f-%.3f kernels -%.3f d2h -%.3f ms\n", c, tnf[c], a, b, k, d);
        }
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
        tev.clear(); tnf.clear();
    }
    auto mark = [&](cudaStream_t st_) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st_); tev.push_back(e); } };
#else
    auto mark = [&](cudaStream_t) {};
#endif
    for (int f0 = 0, nf = 0; f0 < n_frames; f0 += nf) {
        RigSlot& sl = r->slot[r->next_slot % n_slots];
        r->next_slot = (r->next_slot + 1) % n_slots;
        cudaStream_t s = sl.orb->stream;
        nf = std::min(plan && pi < plan->size() ? (*plan)[pi++] : chunk, n_frames - f0);
        const size_t n_img = (size_t)3 * nf;
        mcv_status st;
        const uint8_t* d_imgs = imgs + f0 * img3;
        mark(s);
#ifdef MCV_EXPERIMENTS
        if (trace) tnf.push_back(nf);
#endif
        if (!imgs_on_device) {
            if ((st = sl.imgs.reserve(img3 * nf))) return st;
            MCV_CUDA(cudaMemcpyAsync(sl.imgs.p, imgs + f0 * img3, img3 * nf, cudaMemcpyHostToDevice, s));
            d_imgs = sl.imgs.as<uint8_t>();
        }
        mark(s);
        mcv_keypoint* d_kps = kps_out + (size_t)3 * f0 * cap; uint8_t* d_desc = desc_out + (size_t)3 * f0 * cap * 32;
        int* d_counts = counts + 3 * f0; float* d_ur = u_right + (size_t)f0 * cap; float* d_dp = depth_left + (size_t)f0 * cap;
        if (!out_on_device) {
            if ((st = sl.kps.reserve(n_img * cap * kb))) return st;
            if ((st = sl.desc.reserve(n_img * cap * 32))) return st;
            if ((st = sl.counts.reserve(n_img * 4))) return st;
            if ((st = sl.u_right.reserve((size_t)nf * cap * 4))) return st;
            if ((st = sl.depth.reserve((size_t)nf * cap * 4))) return st;
            d_kps = sl.kps.as<mcv_keypoint>(); d_desc = sl.desc.as<uint8_t>(); d_counts = sl.counts.as<int>();
            d_ur = sl.u_right.as<float>(); d_dp = sl.depth.as<float>();
        }
        st = rig_chunk(r, sl, d_imgs, nf, w, hgt, d_kps, d_desc, d_counts, d_ur, d_dp, cap, &launches);
        if (st) return st;
        mark(s);
        if (!out_on_device) {
            MCV_CUDA(cudaMemcpyAsync(kps_out + (size_t)3 * f0 * cap, d_kps, n_img * cap * kb, cudaMemcpyDeviceToHost, s));
            MCV_CUDA(cudaMemcpyAsync(desc_out + (size_t)3 * f0 * cap * 32, d_desc, n_img * cap * 32, cudaMemcpyDeviceToHost, s));
            MCV_CUDA(cudaMemcpyAsync(counts + 3 * f0, d_counts, n_img * 4, cudaMemcpyDeviceToHost, s));
            MCV_CUDA(cudaMemcpyAsync(u_right + (size_t)f0 * cap, d_ur, (size_t)nf * cap * 4, cudaMemcpyDeviceToHost, s));
            MCV_CUDA(cudaMemcpyAsync(depth_left + (size_t)f0 * cap, d_dp, (size_t)nf * cap * 4, cudaMemcpyDeviceToHost, s));
        }
        mark(s);
    }
    r->last_launches = launches;
    return MCV_OK;
}

static inline int rig_chunk_size(const mcv_rig* r, int n_frames) {
    // profiling measures whole-batch kernels on one stream; otherwise chunk so that all slots get work
    if (r->slot[0].orb->profile || r->chunk_frames <= 0) return n_frames;
    return std::max(1, std::min(r->chunk_frames, (n_frames + r->use_slots - 1) / r->use_slots));
}

extern "C" {

mcv_status mcv_rig_create(const mcv_rig_params* p, int device, void* stream, mcv_rig** out) {
    if (!p || !out) return MCV_ERR_BAD_ARG;
    *out = nullptr;
    if (!(p->baseline > 0.f) || !(p->bf > 0.f)) { set_error("bad rig parameters"); return MCV_ERR_BAD_ARG; }
    mcv_rig* r = new mcv_rig();
    r->prm = *p; r->device = device;
    for (int i = 0; i < RIG_SLOTS; ++i) {
        mcv_status st = mcv_orb_create(&p->orb, device, nullptr, &r->slot[i].orb);
        if (st) { for (int k = 0; k < i; ++k) mcv_orb_destroy(r->slot[k].orb); delete r; return st; }
        cudaEventCreateWithFlags(&r->slot[i].done, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&r->slot[i].front, cudaEventDisableTiming);
        if (const char* e = getenv("MCV_RIG_QUAD_PRIORITY")) {
            if (atoi(e) > 0) {
                int lo = 0, hi = 0;
                cudaDeviceGetStreamPriorityRange(&lo, &hi);   // hi = numerically lowest = greatest priority
                mcv_orb* h = r->slot[i].orb;
                cudaStreamCreateWithPriority(&h->quad_stream, cudaStreamNonBlocking, hi);
                cudaEventCreateWithFlags(&h->quad_in, cudaEventDisableTiming);
                cudaEventCreateWithFlags(&h->quad_out, cudaEventDisableTiming);
            }
        }
    }
    cudaEventCreateWithFlags(&r->fork, cudaEventDisableTiming);
    if (stream) r->stream = (cudaStream_t)stream;
    else { cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking); r->own_stream = true; }
    if (const char* e = getenv("MCV_RIG_CHUNK")) r->chunk_frames = atoi(e);
    if (const char* e = getenv("MCV_RIG_CHUNK_DEV")) r->chunk_frames_dev = atoi(e);
    if (const char* e = getenv("MCV_RIG_SUBMIT_CHUNK")) r->submit_chunk = atoi(e);
    if (const char* e = getenv("MCV_RIG_SLOTS")) r->use_slots = std::max(1, std::min(RIG_SLOTS, atoi(e)));
    if (const char* e = getenv("MCV_RIG_STAGGER")) r->no_stagger = atoi(e) == 0;
    if (const char* e = getenv("MCV_RIG_GRAPH_MAX")) r->graph_max_frames = atoi(e);
    if (const char* e = getenv("MCV_RIG_SLOTS_SUBMIT")) r->use_slots_submit = std::max(1, std::min(RIG_SLOTS, atoi(e)));
    if (const char* e = getenv("MCV_RIG_SLOTS_DEV")) r->use_slots_dev = std::max(1, std::min(RIG_SLOTS, atoi(e)));
    *out = r;
    return MCV_OK;
}

void mcv_rig_destroy(mcv_rig* r) {
    if (!r) return;
    cudaSetDevice(r->device);
    cudaStreamSynchronize(r->stream);
    for (RigSlot& sl : r->slot) {
        cudaStreamSynchronize(sl.orb->stream);
        for (DevBuf* b : {&sl.imgs, &sl.kps, &sl.desc, &sl.counts, &sl.u_right, &sl.depth, &sl.best_dist, &sl.st_scratch}) b->release();
        if (sl.done) cudaEventDestroy(sl.done);
        if (sl.front) cudaEventDestroy(sl.front);
        for (ChunkGraph& g : sl.graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
        mcv_orb_destroy(sl.orb);
    }
    if (r->fork) cudaEventDestroy(r->fork);
    for (cudaEvent_t e : r->ticket) if (e) cudaEventDestroy(e);
    if (r->own_stream) cudaStreamDestroy(r->stream);
    delete r;
}

int mcv_rig_max_keypoints(const mcv_rig* r) { return r ? max_keypoints_roots(r->slot[0].orb, 4) : 0; }   // aspect ratio <= 4.5
int mcv_rig_max_keypoints_for(const mcv_rig* r, int w, int hgt) { return r ? mcv_orb_max_keypoints_for(r->slot[0].orb, w, hgt, 0) : 0; }
mcv_orb* mcv_rig_extractor(mcv_rig* r) { return r ? r->slot[0].orb : nullptr; }
int mcv_rig_last_launches(const mcv_rig* r) { return r ? r->last_launches : 0; }
mcv_status mcv_rig_set_chunk_frames(mcv_rig* r, int chunk_frames) {
    if (!r) return MCV_ERR_BAD_ARG;
    r->chunk_frames = chunk_frames;
    return MCV_OK;
}

mcv_status mcv_orb_set_input_channels(mcv_orb* h, int channels) {
    if (!h || (channels != 1 && channels != 3)) return MCV_ERR_BAD_ARG;
    h->channels = channels;
    return MCV_OK;
}

mcv_status mcv_rig_set_input_channels(mcv_rig* r, int channels) {
    if (!r || (channels != 1 && channels != 3)) return MCV_ERR_BAD_ARG;
    mcv_status st = mcv_rig_sync(r);
    if (st) return st;
    for (RigSlot& sl : r->slot) sl.orb->channels = channels;
    return MCV_OK;
}

mcv_status mcv_rig_process_async(mcv_rig* r, const uint8_t* d_imgs, int n_frames, int w, int hgt, mcv_keypoint* d_kps, uint8_t* d_desc,
                                 int32_t* d_counts, float* d_u_right, float* d_depth, int cap) {
    if (!r || !d_imgs || n_frames <= 0 || !d_kps || !d_desc || !d_counts || !d_u_right || !d_depth) return MCV_ERR_BAD_ARG;
    if (w <= 0 || hgt <= 0) return MCV_ERR_EMPTY_IMAGE;
    if (cap >= (1 << 20)) return MCV_ERR_BAD_ARG;
    MCV_CUDA(cudaSetDevice(r->device));
    // Chunks rotate over the slots' streams ACROSS calls and are staggered by their front-half events, so the latency-bound
    // quadtree kernel of one chunk (or call) runs beside the issue-bound stencils of the next. The work is ordered after
    // what is already on the rig's stream (fork); the rig's stream is ordered after the results only by mcv_rig_join.
    const bool profiling = r->slot[0].orb->profile;
    const int chunk = profiling || r->chunk_frames_dev <= 0 ? n_frames : std::min(n_frames, r->chunk_frames_dev);
    const size_t img3 = (size_t)3 * w * hgt * r->slot[0].orb->channels;
    int launches = 0;
    MCV_CUDA(cudaEventRecord(r->fork, r->stream));
    for (int f0 = 0; f0 < n_frames; f0 += chunk) {
        RigSlot& sl = r->slot[profiling ? 0 : r->next_slot % r->use_slots_dev];
        if (!profiling) r->next_slot = (r->next_slot + 1) % r->use_slots_dev;
        MCV_CUDA(cudaStreamWaitEvent(sl.orb->stream, r->fork, 0));
        const int nf = std::min(chunk, n_frames - f0);
        mcv_status st = rig_chunk(r, sl, d_imgs + f0 * img3, nf, w, hgt, d_kps + (size_t)3 * f0 * cap, d_desc + (size_t)3 * f0 * cap * 32,
                                  d_counts + 3 * f0, d_u_right + (size_t)f0 * cap, d_depth + (size_t)f0 * cap, cap, &launches, !profiling);
        if (st) return st;
    }
    r->pending_join = true;
    r->last_launches = launches;
    return MCV_OK;
}

mcv_status mcv_rig_join(mcv_rig* r) {
    if (!r) return MCV_ERR_BAD_ARG;
    MCV_CUDA(cudaSetDevice(r->device));
    for (RigSlot& sl : r->slot) {
        MCV_CUDA(cudaEventRecord(sl.done, sl.orb->stream));
        MCV_CUDA(cudaStreamWaitEvent(r->stream, sl.done, 0));
    }
    r->pending_join = false;
    return MCV_OK;
}

mcv_status mcv_rig_set_profiling(mcv_rig* r, int on) {
    if (!r) return MCV_ERR_BAD_ARG;
    mcv_orb* h = r->slot[0].orb;
    MCV_CUDA(cudaSetDevice(h->device));
    if (on && h->ev.empty()) {
        h->ev.resize((size_t)PROF_RING * (N_STAGES + 1));
        for (auto& e : h->ev) MCV_CUDA(cudaEventCreate(&e));
    }
    h->profile = on != 0;
    h->prof_calls = 0;
    return MCV_OK;
}

mcv_status mcv_rig_stage_ms(mcv_rig* r, float* total_ms, int n_stages, int* n_calls) {
    if (!r || !total_ms || n_stages < N_STAGES) return MCV_ERR_BAD_ARG;
    mcv_orb* h = r->slot[0].orb;
    MCV_CUDA(cudaSetDevice(h->device));
    MCV_CUDA(cudaStreamSynchronize(h->stream));
    for (int s = 0; s < n_stages; ++s) total_ms[s] = 0.f;
    for (int c = 0; c < h->prof_calls; ++c)
        for (int s = 0; s < N_STAGES; ++s) {
            float ms = 0.f;
            MCV_CUDA(cudaEventElapsedTime(&ms, h->ev[(size_t)c * (N_STAGES + 1) + s], h->ev[(size_t)c * (N_STAGES + 1) + s + 1]));
            total_ms[s] += ms;
        }
    if (n_calls) *n_calls = h->prof_calls;
    return MCV_OK;
}

mcv_status mcv_rig_sync(mcv_rig* r) {
    if (!r) return MCV_ERR_BAD_ARG;
    MCV_CUDA(cudaSetDevice(r->device));
    MCV_CUDA(cudaStreamSynchronize(r->stream));
    for (RigSlot& sl : r->slot) MCV_CUDA(cudaStreamSynchronize(sl.orb->stream));
    return MCV_OK;
}

mcv_status mcv_rig_process(mcv_rig* r, const uint8_t* imgs, int n_frames, int w, int hgt, int imgs_on_device, mcv_keypoint* kps_out,
                           uint8_t* desc_out, int32_t* counts, float* u_right, float* depth_left, int cap, int out_on_device) {
    if (!r || !imgs || n_frames <= 0 || !kps_out || !desc_out || !counts || !u_right || !depth_left) return MCV_ERR_BAD_ARG;
    if (w <= 0 || hgt <= 0) return MCV_ERR_EMPTY_IMAGE;
    if (cap >= (1 << 20)) return MCV_ERR_BAD_ARG;
    MCV_CUDA(cudaSetDevice(r->device));
    if (imgs_on_device && out_on_device) {
        mcv_status st = mcv_rig_process_async(r, imgs, n_frames, w, hgt, kps_out, desc_out, counts, u_right, depth_left, cap);
        if (st) return st;
        return mcv_rig_sync(r);
    }
    // host side involved: per-chunk H2D -> kernels -> D2H on the slot's stream, chunks overlapping across slots
    MCV_CUDA(cudaStreamSynchronize(r->stream));
    r->next_slot = 0;   // every slot is idle here (the call is synchronous): the same chunk of the same call shape always lands on the
                        // same slot, so its captured graph and its workspace size are reused call after call
    mcv_status st = rig_enqueue_host(r, imgs, n_frames, w, hgt, imgs_on_device, kps_out, desc_out, counts, u_right, depth_left, cap,
                                     out_on_device, rig_chunk_size(r, n_frames), r->use_slots);
    if (st) return st;
    for (RigSlot& sl : r->slot) MCV_CUDA(cudaStreamSynchronize(sl.orb->stream));
    return MCV_OK;
}

mcv_status mcv_rig_submit(mcv_rig* r, const uint8_t* imgs, int n_frames, int w, int hgt, mcv_keypoint* kps_out, uint8_t* desc_out,
                          int32_t* counts, float* u_right, float* depth_left, int cap, long long* ticket) {
    if (!r || !imgs || n_frames <= 0 || !kps_out || !desc_out || !counts || !u_right || !depth_left) return MCV_ERR_BAD_ARG;
    if (w <= 0 || hgt <= 0) return MCV_ERR_EMPTY_IMAGE;
    if (cap >= (1 << 20)) return MCV_ERR_BAD_ARG;
    MCV_CUDA(cudaSetDevice(r->device));
    const int chunk = r->slot[0].orb->profile || r->submit_chunk <= 0 ? n_frames : std::min(n_frames, r->submit_chunk);
    mcv_status st = rig_enqueue_host(r, imgs, n_frames, w, hgt, 0, kps_out, desc_out, counts, u_right, depth_left, cap, 0, chunk, r->use_slots_submit);
    if (st) return st;
    // completion = every slot's stream has drained what this call put on it
    const long long id = ++r->submitted;
    cudaEvent_t& ev = r->ticket[id % RIG_TICKETS];
    if (!ev) MCV_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (RigSlot& sl : r->slot) {
        MCV_CUDA(cudaEventRecord(sl.done, sl.orb->stream));
        MCV_CUDA(cudaStreamWaitEvent(r->stream, sl.done, 0));
    }
    MCV_CUDA(cudaEventRecord(ev, r->stream));
    if (ticket) *ticket = id;
    return MCV_OK;
}

mcv_status mcv_rig_wait(mcv_rig* r, long long ticket) {
    if (!r || ticket <= 0 || ticket > r->submitted) return MCV_ERR_BAD_ARG;
    MCV_CUDA(cudaSetDevice(r->device));
    // A ticket older than the ring shares its event slot with a LATER ticket; tickets complete in submission order (one ordered
    // gather stream), so waiting on the slot's current occupant is a correct — merely later — completion point.
    MCV_CUDA(cudaEventSynchronize(r->ticket[ticket % RIG_TICKETS]));
    return MCV_OK;
}

mcv_status mcv_stereo_match(mcv_orb* left, mcv_orb* right, const mcv_keypoint* kps_l, const uint8_t* desc_l, int n_l, const mcv_keypoint* kps_r,
                            const uint8_t* desc_r, int n_r, float bf, float baseline, float* u_right, float* depth_left, int32_t* best_dist,
                            int32_t* best_r) {
    if (!left || !right || n_l < 0 || n_r < 0 || !u_right || !depth_left || !(baseline > 0.f)) return MCV_ERR_BAD_ARG;
    if (!left->have_plan || !right->have_plan || left->last_images < 1 || right->last_images < 1) { set_error("stereo_match: extract on both handles first"); return MCV_ERR_BAD_ARG; }
    if (left->device != right->device || memcmp(&left->plan, &right->plan, sizeof(Plan)) != 0) { set_error("stereo_match: handles must share device, parameters and image size"); return MCV_ERR_BAD_ARG; }
    if (n_l == 0) return MCV_OK;
    if (n_r >= (1 << 20) || !kps_l || !desc_l || (n_r > 0 && (!kps_r || !desc_r))) return MCV_ERR_BAD_ARG;
    for (int i = 0; i < n_l; ++i) if (kps_l[i].octave < 0 || kps_l[i].octave >= left->plan.n_levels) return MCV_ERR_BAD_ARG;
    for (int i = 0; i < n_r; ++i) if (kps_r[i].octave < 0 || kps_r[i].octave >= left->plan.n_levels) return MCV_ERR_BAD_ARG;
    MCV_CUDA(cudaSetDevice(left->device));
    cudaStream_t s = left->stream;
    MCV_CUDA(cudaStreamSynchronize(right->stream));
    mcv_status st;
    const size_t kb = sizeof(mcv_keypoint);
    const size_t need = (size_t)n_l * (kb + 32 + 16) + (size_t)std::max(n_r, 1) * (kb + 32) + 256;
    if ((st = left->misc.reserve(need))) return st;
    if ((st = left->seeds.reserve(stereo_scratch_bytes(left->plan, 1, std::max(n_r, 1))))) return st;  // seeds buffer is idle between extracts
    uint8_t* base = left->misc.as<uint8_t>();
    float* d_ur = reinterpret_cast<float*>(base); float* d_dp = d_ur + n_l; int* d_bd = reinterpret_cast<int*>(d_dp + n_l); int* d_br = d_bd + n_l;
    uint8_t* d_dl = reinterpret_cast<uint8_t*>(d_br + n_l); uint8_t* d_dr = d_dl + (size_t)n_l * 32;
    mcv_keypoint* d_kl = reinterpret_cast<mcv_keypoint*>(d_dr + (size_t)std::max(n_r, 1) * 32); mcv_keypoint* d_kr = d_kl + n_l;
    MCV_CUDA(cudaMemcpyAsync(d_kl, kps_l, n_l * kb, cudaMemcpyHostToDevice, s));
    MCV_CUDA(cudaMemcpyAsync(d_dl, desc_l, (size_t)n_l * 32, cudaMemcpyHostToDevice, s));
    if (n_r) {
        MCV_CUDA(cudaMemcpyAsync(d_kr, kps_r, n_r * kb, cudaMemcpyHostToDevice, s));
        MCV_CUDA(cudaMemcpyAsync(d_dr, desc_r, (size_t)n_r * 32, cudaMemcpyHostToDevice, s));
    }
    launch_stereo_pair(left->plan, left->pyr.as<uint8_t>(), right->pyr.as<uint8_t>(), d_kl, d_dl, n_l, d_kr, d_dr, n_r, bf, baseline, d_ur, d_dp, d_bd,
                       d_br, left->seeds.p, s);
    MCV_CUDA(cudaGetLastError());
    MCV_CUDA(cudaMemcpyAsync(u_right, d_ur, (size_t)n_l * 4, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaMemcpyAsync(depth_left, d_dp, (size_t)n_l * 4, cudaMemcpyDeviceToHost, s));
    if (best_dist) MCV_CUDA(cudaMemcpyAsync(best_dist, d_bd, (size_t)n_l * 4, cudaMemcpyDeviceToHost, s));
    if (best_r) MCV_CUDA(cudaMemcpyAsync(best_r, d_br, (size_t)n_l * 4, cudaMemcpyDeviceToHost, s));
    MCV_CUDA(cudaStreamSynchronize(s));
    return MCV_OK;
}

}  // extern "C"
