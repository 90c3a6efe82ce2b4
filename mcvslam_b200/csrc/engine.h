// Internal declarations of libmcv_b200.so (host side + kernel launchers). sm_100a only; no CPU fallback.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/mcv_b200.h"

namespace mcv {

constexpr int MAX_LEVELS = 16;
constexpr int EDGE_THRESHOLD = 19;   // ORBextractor.cc:73
constexpr int BORDER = 16;           // EDGE_THRESHOLD - 3: origin of FAST candidate coordinates (ORBextractor.cc:588)
constexpr int MAX_DIM = 4096 + 2 * BORDER;  // candidate coordinates are packed in 12 bits
constexpr int NUM_SMS = 148;
#ifndef MCV_FS_ROWS
#define MCV_FS_ROWS 29
#endif
constexpr int FS_ROWS = MCV_FS_ROWS;            // rows per FAST strip; FS_ROWS + 6 is a multiple of the kernel's 7-row register ring
constexpr int OCT_S_BYTES = 2 * 2064;    // quadtree: bucket prefix sums per (image, level) task (<= 2048 buckets + 1, u16)
constexpr int FS_SEG = 128 * FS_ROWS;  // list entries a strip owns: worst case every pixel of the strip scores
constexpr int FS_EDGE_BYTES = (256 + 2 * FS_ROWS + 15) & ~15;   // per-strip edge record: scores on its first / last row and column (fast_kernels.cu)

// Candidate / quadtree point: x (12 bits) | y (12 bits) << 12 | response (8 bits) << 24, coordinates relative to BORDER.
__host__ __device__ inline uint32_t pack_pt(int x, int y, int r) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)r << 24); }
__host__ __device__ inline int pt_x(uint32_t p) { return (int)(p & 0xfffu); }
__host__ __device__ inline int pt_y(uint32_t p) { return (int)((p >> 12) & 0xfffu); }
__host__ __device__ inline int pt_r(uint32_t p) { return (int)(p >> 24); }

// Per-level geometry, computed once on the host from (params, image size) — ORBextractor.cc:407-436,588-602,903-904.
struct LevelGeom {
    int w, h, pitch;       // level image; pitch in bytes (multiple of 16)
    int img_off;           // byte offset of the level inside one image's pyramid block
    int n_cols, n_rows, w_cell, h_cell;
    int cell_base;         // index of this level's first cell within one image's cell array
    int cell_cap;          // candidate slots per cell (upper bound of NMS survivors)
    int cand_off;          // entry offset of this level inside one image's candidate block
    int cand_cap;          // n_cells * cell_cap
    int quota;             // mnFeaturesPerLevel[level]
    int n_ini;             // quadtree roots
    float h_x;             // root width
    int out_off, out_cap;  // quadtree output slots inside one image's block
    float scale, inv_scale;
    int kp_size;           // (int)(31 * scale)
    int tab_off;           // int offset into the resize tables (level >= 1)
    int area_fast;         // exact 2x decimation -> 2x2 box path
    int march_ok;          // geometry fits k_resize_march (8-byte source window per 4 output px, increasing source rows)
};

struct Plan {
    int n_levels, w, h;
    int ini_th, min_th;
    int pyr_bytes;         // per image, multiple of 256
    int cells_per_image, cand_per_image, out_per_image;
    int n_fast_strips;     // FAST strips (128 px x FS_ROWS rows) over all levels; each owns FS_SEG list entries
    int max_quad_kp;       // sum of out_cap
    int max_cell_w, max_cell_h, max_quota;
    LevelGeom lv[MAX_LEVELS];
};

struct StripTable {         // prefix of per-level strip counts: linear strip id -> (level, strip)
    int first[MAX_LEVELS + 1];
    int strips_x[MAX_LEVELS];
};

int fast_strip_table(const Plan& P, StripTable& T);

struct SeedInfo {          // single-image path only
    const mcv_keypoint* d_seeds;  // device copy, original order
    int n_seeds;
    int level_count[MAX_LEVELS];
};

void set_error(const std::string& s);
#define MCV_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            mcv::set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                         \
            return MCV_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

// ---- kernel launchers (each returns the number of kernels it enqueued) ----
int launch_pyramid(const Plan& P, const uint8_t* d_src, size_t src_pitch, size_t src_image_stride, int src_channels, uint8_t* d_pyr,
                   const int* d_tabs, int n_images, cudaStream_t s);
int launch_blur(const Plan& P, const uint8_t* d_pyr, uint8_t* d_blur, int n_images, cudaStream_t s);
int launch_fast_cells(const Plan& P, const uint8_t* d_pyr, uint8_t* d_edges, unsigned* d_nz_list, int* d_nz_cnt, uint32_t* d_cell_raw,
                      uint32_t* d_cell_pts, int* d_cell_cnt, int* d_fallback, int n_images, cudaStream_t s, cudaEvent_t after_score = nullptr);
int launch_octree(const Plan& P, const uint32_t* d_cell_pts, const int* d_cell_cnt, uint32_t* d_arena_a, uint32_t* d_arena_b,
                  uint16_t* d_oct_idx, uint32_t* d_out_pts, int* d_out_cnt, int n_images, cudaStream_t s);
int launch_orient_desc(const Plan& P, const uint8_t* d_pyr, const uint8_t* d_blur, const uint32_t* d_out_pts, const int* d_out_cnt,
                       const SeedInfo* seeds, mcv_keypoint* d_kps, uint8_t* d_desc, int* d_counts, int cap, int n_images,
                       cudaStream_t s);
int launch_stereo(const Plan& P, const uint8_t* d_pyr, const mcv_keypoint* d_kps, const uint8_t* d_desc, const int* d_counts, int cap,
                  int n_frames, int left_cam, int right_cam, int cams_per_frame, float bf, float baseline, float* d_u_right,
                  float* d_depth, int* d_best_dist, int* d_best_r, void* d_scratch, cudaStream_t s, cudaEvent_t mid = nullptr);
int launch_fill_tails(mcv_keypoint* d_kps, uint8_t* d_desc, const int* d_counts, int cap, int n_images, float* d_u_right, float* d_depth,
                      int cams_per_frame, cudaStream_t s);
// bytes of d_scratch (row tables + sorted right-keypoint records) for launch_stereo / launch_stereo_pair
size_t stereo_scratch_bytes(const Plan& P, int n_frames, int max_right);
// stereo across two separate pyramids (mcv_stereo_match on two handles): one "frame", explicit pointers
int launch_stereo_pair(const Plan& P, const uint8_t* d_pyr_l, const uint8_t* d_pyr_r, const mcv_keypoint* d_kl, const uint8_t* d_dl, int nl,
                       const mcv_keypoint* d_kr, const uint8_t* d_dr, int nr, float bf, float baseline, float* d_u_right, float* d_depth,
                       int* d_best_dist, int* d_best_r, void* d_scratch, cudaStream_t s);
int launch_octree_standalone(const uint32_t* d_pts, int n, int w_box, int h_box, int n_target, uint32_t* d_arena_a, uint32_t* d_arena_b,
                             uint32_t* d_out, int* d_out_cnt, int out_cap, cudaStream_t s);

int octree_debug_clocks(long long out[8]);

// matching
// tensor-core brute force (match_tc_kernels.cu): single problem (n_images == 0) or pairs of images of one descriptor array
bool knn2_tc_usable(int nq, int nt);
size_t knn2_tc_scratch_bytes(int nq_rows, int nt_rows, int nq, int nt, int n_pairs);
int launch_knn2_tc(const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int train_offset, int32_t* d_idx, int32_t* d_dist, void* d_scratch,
                   int n_images, const int32_t* d_counts, const int32_t* d_pair_q, const int32_t* d_pair_t, int n_pairs, cudaStream_t s);
size_t knn2_bf_part_bytes(int nq, int nt);   // partial-key scratch one call needs (owned by the caller: the launchers are stateless)
int launch_knn2_bf(const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int train_offset, int32_t* d_idx, int32_t* d_dist, unsigned* d_part,
                   cudaStream_t s);
int launch_knn2_candidates(const uint8_t* d_q, int nq, const uint8_t* d_t, const int32_t* d_off, const int32_t* d_cidx, int32_t* d_idx,
                           int32_t* d_dist, cudaStream_t s);
int launch_project(const mcv_keypoint* d_kps, const uint8_t* d_desc, int n, int w, int h, const float* d_scale, const float* d_pose,
                   const float* d_xyz, const uint8_t* d_mp_desc, const int32_t* d_level, int n_mp, float r_th, int32_t* d_cell_start,
                   int32_t* d_cell_idx, int32_t* d_idx, int32_t* d_dist, cudaStream_t s);   // d_cell_start: 901 ints, d_cell_idx: n ints (30 x 30 cell table, built here)
int launch_fuse_match(const mcv_keypoint* d_kps, const uint8_t* d_desc, int n, int w, int h, const float* d_par, int n_levels,
                      const float* d_depth_left, const float* d_xyz, const float* d_normal, const uint8_t* d_mp_desc, const int32_t* d_level,
                      int n_mp, int32_t* d_cell_start, int32_t* d_cell_idx, int32_t* d_idx, int32_t* d_dist, cudaStream_t s);
int launch_wnd_track(const mcv_keypoint* d_kps1, const uint8_t* d_desc1, const int32_t* d_qidx, int n_q, const mcv_keypoint* d_kps2,
                     const uint8_t* d_desc2, int n2, int w, int h, int32_t* d_cell_start, int32_t* d_cell_idx, int32_t* d_idx, int32_t* d_best,
                     int32_t* d_dist, cudaStream_t s);
int launch_bow_descend(const uint8_t* d_desc, int n, const int32_t* d_child_off, const uint32_t* d_child_ids, const uint8_t* d_node_desc,
                       int nid_level, int max_depth, uint32_t* d_leaf, uint32_t* d_nid, cudaStream_t s);
int launch_distinctive(const uint8_t* d_desc, const int32_t* d_off, int n_mp, int32_t* d_best_idx, int32_t* d_best_median, cudaStream_t s);
size_t lk_workspace_bytes(int w, int h);
int launch_lk_track(const uint8_t* d_prev, const uint8_t* d_next, int n_pairs, int w, int h, int src_pitch, size_t img_stride, void* d_ws,
                    const float* d_pts, const int* d_pair_of, int n, float* d_next_pts, uint8_t* d_status, float* d_err, cudaStream_t s);
int launch_debug_sincosf(const float* d_a, int n, float* d_s, float* d_c, cudaStream_t s);
int launch_debug_atan2(const float* d_y, const float* d_x, int n, float* d_o, cudaStream_t s);
int launch_popc_peak(int iters, unsigned* d_sink, int blocks, int threads, cudaStream_t s);

}  // namespace mcv
