/*
 * mcv_b200.h — C ABI of the B200-native ORB-extract + Hamming-match engine (libmcv_b200.so).
 *
 * This is the drop-in boundary for the tracking front-end hot path of Sologala/MCVSLAM. The reference has no FFI
 * of its own (single C++ process); the seam is its C++ virtual/static methods. Each entry point below names the
 * reference interface it replaces (paths relative to the reference root). The C++ host mirror that keeps the
 * reference's class surface (MCVSLAM::ORB, MCVSLAM::Matcher, ...) on top of this ABI is
 * mcvslam_b200/host/mcvslam_b200.hpp; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions: plain pointers and sizes only; caller owns every host buffer; the engine owns device workspaces;
 * no exception crosses the boundary; every call returns mcv_status (0 = ok, <0 = error) and is synchronous with
 * respect to its outputs unless it says "async". There is NO CPU fallback: without a CUDA device every compute
 * entry point returns MCV_ERR_NO_DEVICE / MCV_ERR_CUDA.
 */
#ifndef MCV_B200_H
#define MCV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int mcv_status;
enum {
    MCV_OK = 0,
    MCV_ERR_EMPTY_IMAGE = -1,   /* mirrors `return -1` on empty input, ORBextractor.cc:834 */
    MCV_ERR_BAD_ARG = -2,
    MCV_ERR_IMAGE_TOO_SMALL = -3, /* a pyramid level is narrower than one 35-px cell (reference divides by zero) */
    MCV_ERR_CAPACITY = -4,      /* output buffer / internal capacity too small */
    MCV_ERR_CUDA = -5,          /* CUDA runtime failure; see mcv_last_error() */
    MCV_ERR_NO_DEVICE = -6,
    MCV_ERR_SEED_RANGE = -7     /* a pre-seeded keypoint has a bad octave or lies within 19 px of the level border */
};

/* Layout-identical to cv::KeyPoint (28 bytes) — the Frame/Object `kps` element type (BaseExtractor.hpp:10). */
typedef struct mcv_keypoint {
    float x, y;      /* pt */
    float size;
    float angle;     /* degrees [0,360) */
    float response;  /* FAST corner score (ORBextractor.cc:619) */
    int32_t octave;
    int32_t class_id;
} mcv_keypoint;

/* Layout-identical to cv::DMatch (16 bytes) — element of MatchRes / MatchResKnn (include/Matcher.hpp:39-56). */
typedef struct mcv_dmatch {
    int32_t queryIdx, trainIdx, imgIdx;
    float distance;
} mcv_dmatch;

/* The five keys MCVSLAM::ORB::Parse reads from extractor.yaml (ORBExtractor.cpp:10-18). */
typedef struct mcv_orb_params {
    int32_t nfeatures;    /* nkeypoints */
    float scale_factor;   /* scale_factor */
    int32_t nlevels;      /* nlevels */
    int32_t ini_th_fast;  /* ORBextractor.iniThFAST */
    int32_t min_th_fast;  /* ORBextractor.minThFAST */
} mcv_orb_params;

typedef struct mcv_orb mcv_orb; /* one per extractor instance (Frame::extractor_left/right/wide, src/Frame.cpp:24-26) */

const char* mcv_last_error(void);      /* thread-local text of the last failure */
const char* mcv_version(void);
int mcv_device_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Extractor — replaces ORB_SLAM3::ORBextractor / MCVSLAM::ORB.
 * ---------------------------------------------------------------------------------------------------------- */

/* ORBextractor::ORBextractor + init (ORBextractor.cc:402-457); ORB::ORB(config) (ORBExtractor.cpp:20-23).
 * `stream` is a cudaStream_t to run on, or NULL for an engine-owned stream. */
mcv_status mcv_orb_create(const mcv_orb_params* params, int device, void* stream, mcv_orb** out);
void mcv_orb_destroy(mcv_orb* h);

/* Public per-level vectors (ORBextractor.h:60-70,96-99: mvScaleFactor, mvInvScaleFactor, mvLevelSigma2,
 * mvInvLevelSigma2) and mnFeaturesPerLevel. Each out array has nlevels entries; any may be NULL. Host only. */
mcv_status mcv_orb_get_scales(const mcv_orb* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                              int32_t* features_per_level);
/* Upper bound of keypoints one Extract call can return for `n_seeds` pre-seeded keypoints (quadtree overshoot
 * included, ORBextractor.cc:557-565): per level quota + 3 + the number of quadtree roots. mcv_orb_max_keypoints is valid for
 * every image size the engine accepts (15 roots per level); mcv_orb_max_keypoints_for is the exact bound for a w x h image
 * (roots = round(width / height) of the level's usable area, ORBextractor.cc:527). Size kps_out/desc_out with either. */
int mcv_orb_max_keypoints(const mcv_orb* h, int n_seeds);
int mcv_orb_max_keypoints_for(const mcv_orb* h, int w, int hgt, int n_seeds);

/* BaseExtractor::Extract(img, kps, desps) (BaseExtractor.hpp:17) == ORBextractor::operator() (ORBextractor.cc:831-899).
 * img: CV_8UC1 host image (w x h, row stride in bytes). seeds: caller's pre-seeded keypoints (may be NULL/0), appended
 * per octave after the quadtree keypoints exactly as the reference does. kps_out/desc_out: host buffers with room for
 * `cap` keypoints / cap*32 bytes (desc is the continuous N x 32 CV_8U matrix). *n_out = returned count.
 * Returns MCV_ERR_EMPTY_IMAGE for an empty image. */
mcv_status mcv_orb_extract(mcv_orb* h, const uint8_t* img, int w, int hgt, size_t stride, const mcv_keypoint* seeds,
                           int n_seeds, mcv_keypoint* kps_out, uint8_t* desc_out, int cap, int* n_out);

/* Batched Extract over n_images same-sized images (the ThreadPool(3) fan-out of src/Frame.cpp:118-126 generalised to
 * a frame batch). imgs: n_images contiguous w*h images, on the host (imgs_on_device=0) or already in HBM (=1).
 * Outputs: per image `cap` slots; counts[n_images]. Outputs on the host (out_on_device=0) or device (=1).
 * No seeds on this path. */
mcv_status mcv_orb_extract_batch(mcv_orb* h, const uint8_t* imgs, int n_images, int w, int hgt, int imgs_on_device,
                                 mcv_keypoint* kps_out, uint8_t* desc_out, int32_t* counts, int cap, int out_on_device);
/* The same with device-resident inputs and outputs, ASYNCHRONOUS on the handle's stream (the stream given to mcv_orb_create):
 * only enqueues; order consumers after it on that stream or synchronise it. Lets several handles / streams overlap — the
 * latency-bound quadtree of one batch beside the stencils of another. */
mcv_status mcv_orb_extract_batch_async(mcv_orb* h, const uint8_t* d_imgs, int n_images, int w, int hgt, mcv_keypoint* d_kps,
                                       uint8_t* d_desc, int32_t* d_counts, int cap);

/* mvImagePyramid[level] of image `image_index` of the last extract call (ORBextractor.h:72; read by the stereo SAD,
 * src/Frame.cpp:239-261). Copies the un-bordered level into dst (dst_stride bytes per row); w and hgt receive the level size.
 * Pass dst = NULL to query the size only. */
mcv_status mcv_orb_download_level(mcv_orb* h, int image_index, int level, uint8_t* dst, size_t dst_stride, int* w, int* hgt);
/* Device view of the same level (pitch in bytes); valid until the next extract on this handle. */
mcv_status mcv_orb_level_device(mcv_orb* h, int image_index, int level, const uint8_t** dev_ptr, int* w, int* hgt, size_t* pitch);

/* Static ORBextractor::DistributeOctTree (ORBextractor.h:73-74, ORBextractor.cc:524-580) on caller keypoints
 * (pt integer-valued, relative to (minX,minY)); output in the reference's heap-pop order. Host buffers. */
mcv_status mcv_orb_distribute_octree(mcv_orb* h, const mcv_keypoint* in, int n, int min_x, int max_x, int min_y, int max_y,
                                     int n_target, mcv_keypoint* out, int cap, int* n_out);

/* ------------------------------------------------------------------------------------------------------------
 * Matcher — replaces MCVSLAM::Matcher statics (include/Matcher.hpp:58-92, src/Matcher.cpp).
 * Descriptors are rows of 32 bytes. All buffers host unless the name ends in _device.
 *
 * Threading (the reference's Matcher is stateless, "may be called from any thread"): every entry point of this section and of
 * the caller-side sections below (projection / fuse / window tracking / distinctive descriptors / optical flow / BoW) is
 * re-entrant. Each host thread owns a private context per device — a non-blocking stream and all scratch buffers — created
 * on first use on the calling thread's CURRENT device (cudaGetDevice; the library never switches devices here) and released
 * when the thread exits. mcv_knn2_bf_device works on the caller's stream with a stream-ordered scratch allocation. Extractor
 * and rig handles are stateful like the reference's static ORB instances: one call at a time per handle, different handles
 * concurrently.
 * ---------------------------------------------------------------------------------------------------------- */

/* Matcher::KnnMatch(const cv::Mat&, const cv::Mat&, 2) / KnnMatch_cv (src/Matcher.cpp:134-138,304-308) ==
 * cv::BFMatcher(NORM_HAMMING).knnMatch: per query the first two train rows ordered by (distance, trainIdx).
 * out: nq*2 entries, imgIdx 0; if nt < 2 the unused entries have trainIdx -1. *k_out = min(2, nt). */
mcv_status mcv_knn2_bf(const uint8_t* q, int nq, const uint8_t* t, int nt, mcv_dmatch* out, int* k_out);
/* Matcher::BFMatch (src/Matcher.cpp:140-144): best train row per query (distance, trainIdx order); out: nq entries. */
mcv_status mcv_bf_match(const uint8_t* q, int nq, const uint8_t* t, int nt, mcv_dmatch* out);
/* Matcher::KnnMatch(vector<Mat>, vector<Mat>, 2) (src/Matcher.cpp:245-302): streaming top-2 with strict '<',
 * ALWAYS two entries per query, padded with (trainIdx 0, distance 999); imgIdx -1. */
mcv_status mcv_knn2_firstparty(const uint8_t* q, int nq, const uint8_t* t, int nt, mcv_dmatch* out);
/* Same 2-NN over a per-query candidate list (the call sites src/Frame.cpp:199-225, src/Object.cpp:217-226,
 * src/Map.cpp:495-527, src/Matcher.cpp:162-181): candidates of query i are t[cand_idx[cand_off[i] .. cand_off[i+1])],
 * trainIdx = position within that list. out: nq*2 entries with the (0, 999) padding. */
mcv_status mcv_knn2_candidates(const uint8_t* q, int nq, const uint8_t* t, int nt, const int32_t* cand_off,
                               const int32_t* cand_idx, mcv_dmatch* out);
/* MatchResKnn::FilterRatio (src/Matcher.cpp:100-111): knn = nq rows of `per` (1 or 2) entries; keeps row[0] when per == 1
 * or row[0].distance / row[1].distance <= ratio (float divide; 0/0 drops the row, the 999 sentinel passes). Host code.
 * out has room for nq entries; *n_out = kept. */
mcv_status mcv_filter_ratio(const mcv_dmatch* knn, int nq, int per, float ratio, mcv_dmatch* out, int* n_out);
/* MatchRes::FilterThreshold (src/Matcher.cpp:23-35): in-place swap-remove of entries with distance > thres_hold (this
 * reorders the survivors exactly as the reference does). *n_io = size in / size out. */
mcv_status mcv_filter_threshold(mcv_dmatch* m, int* n_io, int thres_hold);
/* MatchRes::FilterOrientation (src/Matcher.cpp:44-74): 40-bin rotation histogram, std::sort of the bins by size, keep the
 * three largest. In place; *n_io = size in / size out. */
mcv_status mcv_filter_orientation(mcv_dmatch* m, int* n_io, const mcv_keypoint* kps1, int n1, const mcv_keypoint* kps2, int n2);
/* MatchRes::FilterFMatrix + CheckDistEpipolarLine (src/Matcher.cpp:76-91,310-325): F12 3x3 row-major float,
 * level_sigma2[nlevels]; swap-remove in place. */
mcv_status mcv_filter_fmatrix(mcv_dmatch* m, int* n_io, const mcv_keypoint* kps1, int n1, const mcv_keypoint* kps2, int n2,
                              const float* F12, const float* level_sigma2, int nlevels);
/* Matcher::DBowMatch (src/Matcher.cpp:146-193): 2-NN restricted to features that share a vocabulary node. The two
 * DBoW3::FeatureVector maps are passed flattened and sorted by node id: node_ids[n_nodes], feat_off[n_nodes+1],
 * feat_idx[]. Emits a pair only when a second neighbour exists. out: up to n1*2 entries (pairs); *n_pairs = pairs. */
mcv_status mcv_dbow_match(const uint8_t* desc1, int n1, const uint32_t* node_ids1, const int32_t* feat_off1, const int32_t* feat_idx1,
                          int n_nodes1, const uint8_t* desc2, int n2, const uint32_t* node_ids2, const int32_t* feat_off2,
                          const int32_t* feat_idx2, int n_nodes2, mcv_dmatch* out, int* n_pairs);

/* Device-resident brute-force 2-NN for large sets (BASELINE config "LargeScaleMatching"): d_q/d_t device pointers,
 * d_idx (nq*2 int32) / d_dist (nq*2 int32) device outputs, (distance, trainIdx) order, -1 / INT_MAX padding.
 * Asynchronous on `stream` (cudaStream_t or NULL = default stream). train_offset is added to every trainIdx. */
mcv_status mcv_knn2_bf_device(const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int train_offset, int32_t* d_idx,
                              int32_t* d_dist, void* stream);
/* Many brute-force 2-NN problems in one launch: BF matching between images of one device-resident descriptor array
 * d_desc [n_images][cap][32] with d_counts [n_images] valid rows each (the layout mcv_orb_extract_batch / mcv_rig_process leave
 * on the device) — BASELINE config "batch of 4096 frames, extract + match": pair p matches the rows of image d_pair_q[p]
 * (queries) against image d_pair_t[p] (train), e.g. consecutive frames (i, i + 1). Outputs d_idx / d_dist [n_pairs][cap][2],
 * rows >= the query image's count are left untouched. Every pair == mcv_knn2_bf on the two descriptor sets. At most 65535
 * pairs and 2^27 descriptor rows per call. Asynchronous on
 * `stream`. (All brute-force entry points run on the tensor cores from 2^23 pairs per problem up — int8 GEMM of the +-1
 * expanded descriptors, exact — and on the integer pipe below; env MCV_KNN_POPC=1 forces the integer-pipe kernel.) */
mcv_status mcv_knn2_pairs_device(const uint8_t* d_desc, const int32_t* d_counts, int n_images, int cap, const int32_t* d_pair_q,
                                 const int32_t* d_pair_t, int n_pairs, int32_t* d_idx, int32_t* d_dist, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Frame — three-camera rig: extract x3 + left/right stereo (src/Frame.cpp:78-138,150-328).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct mcv_rig mcv_rig;

typedef struct mcv_rig_params {
    mcv_orb_params orb;   /* left/right/wide extractors share one config (config/frame.yaml:7-9 all -> extractor.yaml) */
    float bf;             /* config/frame.yaml:3 */
    float baseline;       /* config/frame.yaml:4 */
} mcv_rig_params;

mcv_status mcv_rig_create(const mcv_rig_params* params, int device, void* stream, mcv_rig** out);
/* Input pixel format of every later extract / process / submit call on the handle: 1 = CV_8UC1 gray (default, what
 * BaseExtractor::Extract asserts, ORBextractor.cc:837), 3 = CV_8UC3 interleaved BGR as the capture delivers it — then
 * System::Track's cv::cvtColor(COLOR_BGR2GRAY) (src/System.cpp:60-64) runs on the device, fused into the level-0 write of the
 * pyramid (bit-exact to OpenCV's fixed-point path), and image sizes / strides are in bytes of the 3-channel rows. */
mcv_status mcv_rig_set_input_channels(mcv_rig* r, int channels);
mcv_status mcv_orb_set_input_channels(mcv_orb* h, int channels);
void mcv_rig_destroy(mcv_rig* r);
/* Per-image keypoint slots (`cap`) the rig's outputs need: the bound for aspect ratios <= 4.5 (4 quadtree roots per level), and
 * the exact bound for a w x h image. A too small cap is MCV_ERR_CAPACITY, never an overrun. */
int mcv_rig_max_keypoints(const mcv_rig* r);
int mcv_rig_max_keypoints_for(const mcv_rig* r, int w, int hgt);
/* The extractor of slot 0 (image index = 3*frame + cam within the chunk it processed last; cam 0=L,1=R,2=W). The whole batch
 * is in it when it ran as one chunk: n_frames <= chunk_frames/... see mcv_rig_set_chunk_frames(r, 0). */
mcv_orb* mcv_rig_extractor(mcv_rig* r);
/* Batches are cut into chunks of at most `chunk_frames` triplets pipelined over 6 internal streams (H2D / kernels / D2H overlap
 * inside one synchronous call; every stream gets a chunk). Default 32 (env MCV_RIG_CHUNK); 0 = never chunk. Chunks of up to 32
 * frames replay a CUDA graph captured on the second call with the same shape (env MCV_RIG_GRAPH_MAX, 0 = direct launches). */
mcv_status mcv_rig_set_chunk_frames(mcv_rig* r, int chunk_frames);

/* Frame::Frame ORBE + SMatch stages for a batch of n_frames triplets. imgs: [n_frames][3][hgt][w] u8 (L, R, W), host
 * or device. Outputs (host or device): kps [n_frames*3][cap], desc [n_frames*3][cap][32], counts [n_frames*3],
 * u_right / depth_left [n_frames][cap] (Frame::u_right, Frame::depth_left, include/Frame.hpp:44-45; -1 = none). The slots behind
 * an image's count are defined: zero keypoints / descriptors, -1 in u_right / depth_left. */
mcv_status mcv_rig_process(mcv_rig* r, const uint8_t* imgs, int n_frames, int w, int hgt, int imgs_on_device,
                           mcv_keypoint* kps_out, uint8_t* desc_out, int32_t* counts, float* u_right, float* depth_left,
                           int cap, int out_on_device);
/* As mcv_rig_process with device-resident inputs and outputs, but only ENQUEUES the work (async): it runs on the rig's
 * internal streams, ordered after everything already on the rig's stream, and consecutive calls overlap (the latency-bound
 * quadtree stage of one call runs beside the stencils of the next ones): calls of up to the engine's device chunk (128 frames)
 * rotate over 3 internal streams (env MCV_RIG_SLOTS_DEV). The results are complete after mcv_rig_sync() (host wait) or, for
 * work queued on the rig's stream afterwards, after mcv_rig_join(). Do not reuse the output buffers of a call before one of
 * the two — a caller that keeps calls in flight gives each of them its own output set. */
mcv_status mcv_rig_process_async(mcv_rig* r, const uint8_t* d_imgs, int n_frames, int w, int hgt, mcv_keypoint* d_kps,
                                 uint8_t* d_desc, int32_t* d_counts, float* d_u_right, float* d_depth_left, int cap);
/* As mcv_rig_process with HOST inputs and outputs (pinned memory, or the copies degrade to synchronous ones), but only ENQUEUES
 * the step — host->device copy of the images, kernels, device->host copy of every result — on the rig's internal streams and
 * returns a ticket; consecutive submits overlap (copies of one step beside the kernels of another; this replaces the reference's
 * capture thread running ahead of Frame construction, src/System.cpp:60-66). The outputs of a submit are complete after
 * mcv_rig_wait(ticket) (or mcv_rig_sync). The caller keeps the buffers alive and untouched until then. Tickets complete in
 * submission order; the library tracks the last 8 individually — waiting on an older one waits for the later ticket that
 * took over its slot (never returns early). Chunk size: env MCV_RIG_SUBMIT_CHUNK (default 128 frames); whole steps rotate over
 * four internal slots (env MCV_RIG_SLOTS_SUBMIT), so four steps in flight keep three of them in their kernels. */
mcv_status mcv_rig_submit(mcv_rig* r, const uint8_t* imgs, int n_frames, int w, int hgt, mcv_keypoint* kps_out, uint8_t* desc_out,
                          int32_t* counts, float* u_right, float* depth_left, int cap, long long* ticket);
mcv_status mcv_rig_wait(mcv_rig* r, long long ticket);
/* Orders the rig's stream after all work enqueued so far by mcv_rig_process_async (no host wait). */
mcv_status mcv_rig_join(mcv_rig* r);
mcv_status mcv_rig_sync(mcv_rig* r);
/* Number of engine kernels launched by the last process call on this rig (for bench.py's gpu_launches). */
int mcv_rig_last_launches(const mcv_rig* r);
/* Per-stage device timing (the reference brackets the same stages with MyTimer "ORBE"/"SMatch", src/Frame.cpp:119,135).
 * When on, CUDA events are recorded between the stages of every process call on the rig's stream (no extra
 * synchronisation). mcv_rig_stage_ms synchronises and returns the summed milliseconds of the 8 stages — pyramid, blur,
 * fast_score, nms_cells, quadtree, orient_desc, stereo_match, stereo_median — over the calls since profiling was switched on. */
mcv_status mcv_rig_set_profiling(mcv_rig* r, int on);
mcv_status mcv_rig_stage_ms(mcv_rig* r, float* total_ms, int n_stages, int* n_calls);

/* Frame::ComputeStereoMatch (src/Frame.cpp:150-328) on caller keypoints, using the pyramids held by two extractor
 * handles (image 0 of each), as the reference reads extractor_left/right.mvImagePyramid. Host buffers.
 * best_dist (SAD of the chosen shift, -1 none) and best_r (matched right index, -1 none) may be NULL. */
mcv_status mcv_stereo_match(mcv_orb* left, mcv_orb* right, const mcv_keypoint* kps_l, const uint8_t* desc_l, int n_l,
                            const mcv_keypoint* kps_r, const uint8_t* desc_r, int n_r, float bf, float baseline,
                            float* u_right, float* depth_left, int32_t* best_dist, int32_t* best_r);

/* Object::AssignFeaturesToGrid + ProjectBunchMapPoints (src/Object.cpp:182-236,249-308) for an ORDERED array of
 * MapPoints. kps/desc: the camera's n keypoints (image w x hgt). scale_factors: mvScaleFactor[nlevels].
 * Rcw (3x3 row-major), tcw (3), intr (fx fy cx cy). mp_xyz [n_mp][3], mp_desc [n_mp][32], mp_level [n_mp].
 * out_idx[m] = matched keypoint index or -1; out_dist[m] = its Hamming distance or -1. *n_matched = `cnt`. */
mcv_status mcv_project_match(const mcv_keypoint* kps, const uint8_t* desc, int n, int w, int hgt, const float* scale_factors,
                             int nlevels, const float* Rcw, const float* tcw, const float* intr, const float* mp_xyz,
                             const uint8_t* mp_desc, const int32_t* mp_level, int n_mp, float r_threshold, int32_t* out_idx,
                             int32_t* out_dist, int* n_matched);

/* Map::Fuse matching front-end (src/Map.cpp:478-527) for an ORDERED array of MapPoints against one camera's keypoints: the
 * viewing-angle gate, Object::Map / Project, GetFeaturesInArea(uv, 10), the level and reprojection gates (as written in the
 * reference, incl. its use of kf->depth_left and mvLevelSigma2 in the stereo branch), KnnMatch + FilterRatio() +
 * FilterThreshold(). Ow = camera centre (3), depth_left [n] = KeyFrame::depth_left, bf = KeyFrame::bf, mp_normal [n_mp][3] =
 * MapPoint::GetNormalVector(). out_idx[m] = the keypoint index the reference passes to AddMapPoint / ReplaceMappoint, or -1.
 * The map surgery itself (src/Map.cpp:528-547) stays with the caller. */
mcv_status mcv_fuse_match(const mcv_keypoint* kps, const uint8_t* desc, int n, int w, int hgt, const float* level_sigma2,
                          const float* inv_level_sigma2, int nlevels, const float* Rcw, const float* tcw, const float* Ow, const float* intr,
                          const float* depth_left, float bf, const float* mp_xyz, const float* mp_normal, const uint8_t* mp_desc,
                          const int32_t* mp_level, int n_mp, int32_t* out_idx, int32_t* out_dist, int* n_matched);

/* Tracker::Wnd_Track (src/Tracker.cpp:341-360): every listed keypoint q_idx[q] of camera 1 (those owning a MapPoint, in the
 * caller's order) against the keypoints of camera 2 inside a +-20 px window, KnnMatch + FilterRatio() + FilterThreshold()
 * (+ FilterOrientation, the identity on one match). out_idx[q] = the camera-2 index the reference hands to AddMapPoint — as
 * written that is candi_idxs[queryIdx], the FIRST candidate of the window — out_best[q] = the candidate the match belongs to,
 * out_dist[q] = its Hamming distance; -1 where nothing passes. */
mcv_status mcv_wnd_track(const mcv_keypoint* kps1, const uint8_t* desc1, int n1, const int32_t* q_idx, int n_q, const mcv_keypoint* kps2,
                         const uint8_t* desc2, int n2, int w, int hgt, int32_t* out_idx, int32_t* out_best, int32_t* out_dist, int* n_matched);

/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cpp:101-150), batched over MapPoints (the reference runs it per
 * point on keyframe insertion and after Map::Fuse). desc = the observed descriptors of all points back to back (32 B rows, the
 * order all_ob_desps is filled in, :108-115), mp_off [n_mp + 1] = row range of each point. Per point: all-pairs
 * HammingDistance, per-row std::sort, median = row[0.5 * (N - 1)], first row with the least median wins. best_idx[m] = BestIdx
 * (row within the point's range; -1 for a point without observations, which the reference leaves untouched), best_median[m]
 * (may be NULL) = BestMedian, out_desc (may be NULL) [n_mp][32] = the descriptor the reference clones into MapPoint::desp. */
mcv_status mcv_distinctive_descriptors(const uint8_t* desc, const int32_t* mp_off, int n_mp, int32_t* best_idx, int32_t* best_median,
                                       uint8_t* out_desc);

/* KL_Track's optical flow (src/Frame.cpp:52-54): cv::calcOpticalFlowPyrLK(prev, next, pts, next_pts, res, err, Size(10, 10), 1,
 * TermCriteria(COUNT + EPS, 10, 0.01), 0, 0.001) on two 8-bit gray host images of the same size. pts / next_pts = n (x, y)
 * pairs, status = res, err = the L1 window residual / 32. Bit-exact with an SSE-baseline x86-64 OpenCV 4 (oracle pinned
 * against cv2 4.13). */
mcv_status mcv_lk_track(const uint8_t* prev, const uint8_t* next, int w, int hgt, size_t stride, const float* pts, int n, float* next_pts,
                        uint8_t* status, float* err);
/* The same for n_pairs image pairs in one call (the Frame constructor runs KL_Track once per earlier frame and camera,
 * src/Frame.cpp:89-103): prev / next hold n_pairs images back to back (rows `stride` bytes apart, image k starts at row k * hgt),
 * the points of pair k are pts[pt_off[k] .. pt_off[k + 1]). All pairs' points are tracked concurrently. */
mcv_status mcv_lk_track_batch(const uint8_t* prev, const uint8_t* next, int n_pairs, int w, int hgt, size_t stride, const float* pts,
                              const int32_t* pt_off, float* next_pts, uint8_t* status, float* err);
/* KL_Track (src/Frame.cpp:34-76) without the MapPoint bookkeeping: kps = obj1->kps[GetMapPointIdx(mp)] in the order of
 * GetMapPointsVector(). ok[i] = res[i] > 0 && err[i] < 1 (:57-58); new_kps[i] (valid where ok[i]) = the keypoint :65-69 pushes
 * onto obj2->kps: kps[i] with pt = next_pts[i] and octave = 0. With fewer than 10 points nothing is tracked (:41). The caller
 * keeps the mp2idx map (skip MapPoints already seen, :61-63). */
mcv_status mcv_kl_track(const uint8_t* prev, const uint8_t* next, int w, int hgt, size_t stride, const mcv_keypoint* kps, int n, mcv_keypoint* new_kps,
                        uint8_t* ok, int* n_ok);

/* Object::ComputeBow (src/Object.cpp:238-247): DBoW3::Vocabulary::transform(features, BowVector&, FeatureVector&, levelsup)
 * (modules/DBow3/src/Vocabulary.cpp:572-672). The vocabulary is handed over once as flat arrays — what a maintainer gets by
 * walking DBoW3's m_nodes after Vocabulary::load: child_off [n_nodes+1] / child_ids = m_nodes[i].children in stored order
 * (node 0 = root), node_desc [n_nodes][32] = m_nodes[i].descriptor, word_id / weight [n_nodes] = the leaf fields, L = m_L,
 * weighting = DBoW3::WeightingType (0 TF_IDF, 1 TF, 2 IDF, 3 BINARY), norm = what ScoringObject::mustNormalize reports
 * (0 none, 1 L1, 2 L2; orbvoc: TF_IDF + L1). The tree lives in device memory; the descent (k Hamming distances per level) runs
 * on the GPU, the std::map assembly of the two vectors — whose insertion order defines the double sums — on the host. */
typedef struct mcv_voc mcv_voc;
mcv_status mcv_voc_create(int n_nodes, const int32_t* child_off, const uint32_t* child_ids, const uint8_t* node_desc, const int32_t* word_id,
                          const double* weight, int L, int weighting, int norm, int device, mcv_voc** out);
void mcv_voc_destroy(mcv_voc* v);
/* Per feature: out_word / out_weight / out_nid (each may be NULL). BowVector: bow_ids ascending + bow_vals, *n_bow entries (<= n).
 * FeatureVector flattened as mcv_dbow_match takes it: fv_nodes ascending [*n_fv], fv_off [*n_fv + 1], fv_idx [<= n]. Buffers
 * must hold n (+1 for fv_off) entries. */
mcv_status mcv_bow_transform(mcv_voc* v, const uint8_t* desc, int n, int levelsup, int32_t* out_word, double* out_weight, uint32_t* out_nid,
                             uint32_t* bow_ids, double* bow_vals, int* n_bow, uint32_t* fv_nodes, int32_t* fv_off, int32_t* fv_idx, int* n_fv);

/* ------------------------------------------------------------------------------------------------------------
 * Test / measurement taps (used by tests/ and bench.py only).
 * ---------------------------------------------------------------------------------------------------------- */
/* Device ports of the float routines that must be bit-exact: glibc sincosf (ORBextractor.cc:102-103 resolve to the
 * float overloads) and cv::fastAtan2 (ORBextractor.cc:97). Host arrays of n floats. */
mcv_status mcv_debug_sincosf(const float* angles, int n, float* sin_out, float* cos_out);
mcv_status mcv_debug_fast_atan2(const float* y, const float* x, int n, float* out);
/* Intermediate stages of the last mcv_orb_extract* call, image `image_index`:
 *  which = 0: FAST candidates of `level` in reference order (x,y relative to the 16-px border, response);
 *  which = 1: quadtree output of `level` in heap-pop order (same coordinates). */
mcv_status mcv_debug_level_keypoints(mcv_orb* h, int image_index, int level, int which, mcv_keypoint* out, int cap, int* n_out);
/* Blurred level (GaussianBlur 7x7 sigma 2, ORBextractor.cc:874-875) of the last extract call. */
mcv_status mcv_debug_download_blurred(mcv_orb* h, int image_index, int level, uint8_t* dst, size_t dst_stride);
/* SM clock stamps of the quadtree kernel's phases for level 0 of image 0 of the last launch (start, candidates gathered,
 * roots, split loop done, heap drained, keypoints selected, 0, 0) — where the latency of that kernel goes. */
mcv_status mcv_debug_octree_clocks(long long* out8);
/* Integer-pipe microbenchmark: runs `iters` dependent-free rounds of (xor+popc) x8 per thread on the whole GPU and
 * returns achieved popc32/s and the kernel time; the matching roofline denominator. */
mcv_status mcv_debug_popc_peak(int iters, double* popc_per_s, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* MCV_B200_H */
