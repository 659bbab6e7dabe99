/* polychase_b200 -- C ABI of the B200-native Polychase hot path.
 *
 * Drop-in boundary for the reference's analyze / track / refine path.  The reference
 * has no C ABI of its own: its boundary is the pybind11 module `polychase_core`
 * (/root/reference/cpp/polychase_pybind.cc:29-348) over three blocking C++ entry points
 *   GenerateOpticalFlowDatabase   /root/reference/cpp/opticalflow.h:35-41
 *   TrackSequence                 /root/reference/cpp/tracker.h:27-33
 *   RefineTrajectory              /root/reference/cpp/refiner.h:22-27
 * This header is the thin layer *under* that surface: each entry point below names the
 * reference code it replaces.  Plain pointers and sizes only; no torch / Eigen / OpenCV
 * types.  All functions return PC_OK (0) or a negative pc_status; the message of the last
 * failure on a context is pc_last_error(ctx).  A context owns one CUDA device, its
 * streams and all device buffers; calls on one context must not race, different contexts
 * are independent.  Host pointers are caller-owned unless stated.
 *
 * There is no CPU fallback: pc_create fails if no sm_100 device is usable.
 */
#ifndef POLYCHASE_B200_H_
#define POLYCHASE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PC_API __attribute__((visibility("default")))

typedef enum pc_status {
    PC_OK = 0,
    PC_ERR_INVALID = -1,   /* bad argument (the reference's CHECK -> std::logic_error) */
    PC_ERR_CUDA = -2,      /* CUDA runtime failure */
    PC_ERR_NOMEM = -3,
    PC_ERR_NOT_FOUND = -4, /* unknown frame id etc. */
    PC_ERR_CAPACITY = -5,  /* caller buffer / context limit too small */
    PC_ERR_STATE = -6,     /* call sequence violated */
    PC_ERR_NOT_ENOUGH_FEATURES = -7, /* tracker.cc:162-166 */
    PC_ERR_IO = -8,        /* sqlite / file errors (database.cc:12-38) */
    PC_ERR_CANCELLED = -9
} pc_status;

typedef struct pc_ctx pc_ctx;

/* Context limits; 0 selects the default in brackets. */
typedef struct pc_limits {
    int device;          /* CUDA ordinal [0] */
    int max_width;       /* [3840] */
    int max_height;      /* [2160] */
    int max_features;    /* cap on keypoints per frame held on device; when a detector
                            runs with max_corners==0 (unlimited, gftt.h:10) results above
                            this cap are an error [16384] */
    int ring_frames;     /* frame slots kept resident (>= 9 for the +-8 window) [20] */
    int pipeline_depth;  /* frames in flight in the streaming analyzer [8] */
} pc_limits;

/* GFTTOptions, /root/reference/cpp/feature_detection/gftt.h:5-21 (same defaults). */
typedef struct pc_gftt_opts {
    double quality_level;   /* 0.01 */
    double min_distance;    /* 5.0 */
    int block_size;         /* 3 (only 3 supported) */
    int gradient_size;      /* 3 (only 3 supported) */
    int max_corners;        /* 0 = unlimited */
    int use_harris;         /* 0 (1 unsupported: never enabled by the addon) */
    double harris_k;        /* 0.04 */
    int grid_rows;          /* 4 */
    int grid_cols;          /* 4 */
} pc_gftt_opts;

/* OpticalFlowOptions, /root/reference/cpp/opticalflow.h:27-33 (same defaults). */
typedef struct pc_flow_opts {
    int window_size;             /* 10 */
    int max_level;               /* 3 */
    int term_max_iters;          /* 30 */
    double term_epsilon;         /* 0.01 */
    double min_eigen_threshold;  /* 1e-4 */
} pc_flow_opts;

/* VideoInfo, /root/reference/cpp/opticalflow.h:20-25. */
typedef struct pc_video_info {
    uint32_t width, height;
    int32_t first_frame;
    uint32_t num_frames;
} pc_video_info;

/* CameraIntrinsics + Pose = CameraState, /root/reference/cpp/pnp/types.h:18-198,
 * /root/reference/cpp/pose.h:9-12.  16 floats = 64 bytes; this is also the packed record
 * exchanged by the trajectory all-gather.  q is (w,x,y,z) as the Python surface shows it
 * (polychase_pybind.cc:224-232). */
typedef struct pc_camera_state {
    float fx, fy, cx, cy, aspect_ratio, width, height;
    float convention;   /* 0 = OpenGL (looks down -Z), 1 = OpenCV (types.h:13-16) */
    float q[4];         /* w, x, y, z */
    float t[3];
    float filled;       /* 1 = slot holds a state (CameraTrajectory's optional), 0 = empty */
} pc_camera_state;

/* BundleOptions, /root/reference/cpp/pnp/types.h:200-215 (same defaults). */
typedef struct pc_bundle_opts {
    uint64_t max_iterations;           /* 100 */
    uint64_t max_allowed_parallelism;  /* 8; ignored on the GPU path */
    int loss_type;                     /* 0 TRIVIAL, 1 HUBER (default), 2 CAUCHY */
    float loss_scale;                  /* 1.0 */
    float gradient_tol;                /* 1e-10 */
    float step_tol;                    /* 1e-8 */
    float initial_lambda;              /* 1e-5 */
    float min_lambda;                  /* 1e-10 */
    float max_lambda;                  /* 1e10 */
    int verbose;
} pc_bundle_opts;

/* BundleStats, /root/reference/cpp/pnp/types.h:217-225. */
typedef struct pc_bundle_stats {
    uint64_t iterations;
    float initial_cost, cost, lambda;
    uint64_t invalid_steps;
    float step_norm, grad_norm;
} pc_bundle_stats;

PC_API void pc_default_gftt_opts(pc_gftt_opts*);
PC_API void pc_default_flow_opts(pc_flow_opts*);
PC_API void pc_default_bundle_opts(pc_bundle_opts*);
PC_API const char* pc_version(void);

/* ---- context ----------------------------------------------------------------------- */
PC_API int pc_create(const pc_limits* limits, pc_ctx** out);
PC_API void pc_destroy(pc_ctx*);
PC_API const char* pc_last_error(pc_ctx*); /* ctx may be NULL: last pc_create failure */
PC_API int pc_synchronize(pc_ctx*);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
PC_API uint64_t pc_kernel_launches(pc_ctx*);

/* ---- analyze: frames, pyramid, detector, LK ---------------------------------------- */
/* Upload one RGB8 frame (H x W x 3, row stride in bytes) and build its gray pyramid:
 * cv::cvtColor(RGB2GRAY) + GeneratePyramid, opticalflow.cc:259,180-187.  `rgb` is a host
 * pointer (pc_frame_upload_rgb8) or a device pointer (pc_frame_from_device_rgb8).
 * The frame takes a ring slot; the least recently uploaded frame is evicted. */
PC_API int pc_frame_upload_rgb8(pc_ctx*, int32_t frame_id, const uint8_t* rgb, int w, int h,
                                size_t stride, const pc_flow_opts*);
PC_API int pc_frame_from_device_rgb8(pc_ctx*, int32_t frame_id, const uint8_t* rgb_dev, int w,
                                     int h, size_t stride, const pc_flow_opts*);
/* Same, from an already-gray image (tests; the reference's detector takes gray). */
PC_API int pc_frame_upload_gray8(pc_ctx*, int32_t frame_id, const uint8_t* gray, int w, int h,
                                 size_t stride, const pc_flow_opts*);
PC_API int pc_frame_release(pc_ctx*, int32_t frame_id);
/* Levels actually built (cv::buildOpticalFlowPyramid may stop early on small images). */
PC_API int pc_frame_num_levels(pc_ctx*, int32_t frame_id, int* levels_out);
/* Copy pyramid level `level` to host (tightly packed w*h bytes). */
PC_API int pc_frame_read_level(pc_ctx*, int32_t frame_id, int level, uint8_t* out, size_t cap,
                               int* w_out, int* h_out);

/* GoodFeaturesToTrack, gftt.cc:14-192, on the frame's level-0 gray image.  Writes up to
 * `cap` keypoints as (x,y) float pairs in the reference's order, keeps them on device as
 * the frame's keypoints (GenerateKeypoints, opticalflow.cc:154-166).  kps_out may be NULL. */
PC_API int pc_detect(pc_ctx*, int32_t frame_id, const pc_gftt_opts*, float* kps_out, int cap,
                     int* n_out);
/* Min-eigenvalue map of the frame (cv::cornerMinEigenVal, gftt.cc:35-36), w*h floats. */
PC_API int pc_min_eig_map(pc_ctx*, int32_t frame_id, float* eig_out, size_t cap_floats);
/* Resume path (ReadOrGenerateKeypoints, opticalflow.cc:168-178): keypoints read from a DB. */
PC_API int pc_set_keypoints(pc_ctx*, int32_t frame_id, const float* kps, int n);

/* GenerateOpticalFlowForAPair, opticalflow.cc:110-152: pyramidal LK of frame `from`'s
 * keypoints into frame `to`, then the status==1 filter.  Outputs (host, capacity `cap`
 * rows): src_idx (u32 index into `from`'s keypoints), tgt (x,y), err. */
PC_API int pc_lk_pair(pc_ctx*, int32_t from, int32_t to, const pc_flow_opts*, uint32_t* src_idx_out,
                      float* tgt_out, float* err_out, int cap, int* n_out);
/* Unfiltered cv::calcOpticalFlowPyrLK outputs for parity tests: next (n x 2), status, err. */
PC_API int pc_lk_raw(pc_ctx*, int32_t from, int32_t to, const pc_flow_opts*, float* next_out,
                     uint8_t* status_out, float* err_out, int cap, int* n_out);

/* ---- analyze: streaming pipeline (GenerateOpticalFlowDatabase, opticalflow.cc:209-321)
 * Frames are pushed in ascending id order; each push enqueues upload, gray+pyramid,
 * detection and every LK pair (f, f+d), d in {-8,-4,-2,-1,1,2,4,8}, that became
 * computable.  Results come back in push order, one record per pushed frame. */
typedef struct pc_pair_rows {
    int32_t image_id_from, image_id_to;
    int32_t rows;
    const uint32_t* src_kps_indices; /* rows            */
    const float* tgt_kps;            /* rows x 2        */
    const float* flow_errors;        /* rows            */
} pc_pair_rows;

typedef struct pc_frame_result {
    int32_t frame_id;
    int32_t num_keypoints;
    const float* keypoints;          /* num_keypoints x 2, valid until the next pop */
    int32_t num_pairs;
    pc_pair_rows pairs[8];
    /* fused Analyze -> Track (pc_analyze_track_begin): this frame's pose */
    int32_t tracked;                 /* 0 no pose, 1 solved on the device, 2 seeded (known) pose */
    int32_t num_matches;             /* rays that hit the mesh (tracker.cc:64-92) */
    float inlier_ratio;
    pc_camera_state camera;
    pc_bundle_stats stats;
} pc_frame_result;

#define PC_MEM_HOST 0
#define PC_MEM_DEVICE 1
#define PC_MEM_HOST_PINNED 2

PC_API int pc_analyze_begin(pc_ctx*, const pc_video_info*, const pc_gftt_opts*, const pc_flow_opts*);
PC_API int pc_analyze_push_frame(pc_ctx*, int32_t frame_id, const uint8_t* rgb, size_t stride,
                                 int mem_kind);
/* Keypoints already known for this frame (DB resume); call before push_frame. */
PC_API int pc_analyze_preset_keypoints(pc_ctx*, int32_t frame_id, const float* kps, int n);
/* Pops the oldest pushed frame's record (blocks until its work is done).  Returns
 * PC_ERR_NOT_FOUND when nothing is pending.  With `download`==0 only counts are
 * fetched (pointers NULL): results stay device resident. */
PC_API int pc_analyze_pop(pc_ctx*, pc_frame_result* out, int download);
PC_API int pc_analyze_pending(pc_ctx*);
/* Fused Analyze -> Track: a forward TrackSequence (tracker.cc:133-213) chained on the device
 * behind the analyzer.  Call after pc_analyze_begin and pc_mesh_set, before the first push.
 * Every pushed non-halo frame that has flows from already posed earlier frames is solved
 * (SolveFrame, tracker.cc:36-131: ray cast of the matched source keypoints + robust LM from the
 * previous frame's pose) right after its LK batch, from the flow rows still resident in HBM --
 * no host round trip; the result arrives with pc_analyze_pop.  Frames whose pose is known (the
 * sweep's first frame -- TrackSequence seeds it from the scene's view matrix, tracker.cc:205-207
 * -- or a shard's halo) are given with pc_analyze_track_seed before they are pushed. */
PC_API int pc_analyze_track_begin(pc_ctx*, const float model[16], const pc_bundle_opts*,
                                  int optimize_focal_length, int optimize_principal_point);
PC_API int pc_analyze_track_seed(pc_ctx*, int32_t frame_id, const pc_camera_state* cam);
/* Multi-GPU sharding: the first `halo_frames` frames pushed after
 * pc_analyze_begin are prepared and detected but emit no pair rows -- they are the previous
 * shard's last frames, needed only as partners of this shard's pairs. */
PC_API int pc_analyze_set_halo(pc_ctx*, int halo_frames);
PC_API int pc_analyze_end(pc_ctx*);

/* ---- synthetic frames (bench/test input generator; not on the reference path) --------
 * Warps a u8 texture (w x h, device resident after pc_synth_set_texture) by the inverse of
 * homography H (row-major 3x3, texture->image) into an RGB8 device frame (R=G=B). */
PC_API int pc_synth_set_texture(pc_ctx*, const uint8_t* tex, int w, int h);
PC_API int pc_synth_render_rgb8(pc_ctx*, const double H[9], uint8_t* rgb_dev, size_t stride);
PC_API int pc_device_alloc(pc_ctx*, size_t bytes, void** out);
PC_API int pc_device_free(pc_ctx*, void* p);
PC_API int pc_host_alloc_pinned(pc_ctx*, size_t bytes, void** out);
PC_API int pc_host_free_pinned(pc_ctx*, void* p);
PC_API int pc_memcpy_d2h(pc_ctx*, void* dst, const void* src, size_t bytes);
PC_API int pc_memcpy_h2d(pc_ctx*, void* dst, const void* src, size_t bytes);
/* Per-kernel-family device time accumulated since the last reset, in ms (CUDA events on
 * the launching stream; only collected after pc_timing_enable(ctx,1)). */
typedef struct pc_kernel_times {
    /* lk = the 8-pair LK batch launches only; lk_tmpl = the once-per-frame source template launches */
    double gray_pyr_ms, min_eig_ms, select_ms, lk_ms, compact_ms, raycast_ms, pnp_ms, ba_ms, lk_tmpl_ms;
    uint64_t gray_pyr_n, min_eig_n, select_n, lk_n, compact_n, raycast_n, pnp_n, ba_n, lk_tmpl_n;
} pc_kernel_times;
PC_API int pc_timing_enable(pc_ctx*, int on);
/* Whole-region device timing: pc_mark joins the context's three streams and records CUDA
 * event `slot` (0..7) after all work enqueued so far; pc_elapsed_ms waits for both events. */
PC_API int pc_mark(pc_ctx*, int slot);
PC_API int pc_elapsed_ms(pc_ctx*, int slot_begin, int slot_end, float* ms_out);
PC_API int pc_timing_read(pc_ctx*, pc_kernel_times* out, int reset);

/* ---- mesh + track (tracker.cc:36-213, ray_casting.cc:65-133, pnp/) ------------------ */
/* AcceleratedMesh(vertices, triangles, masked_triangles), ray_casting.cc:21-63,
 * geometry.h:52-95.  mask_bits may be NULL (nothing masked). */
PC_API int pc_mesh_set(pc_ctx*, const float* verts, int nv, const uint32_t* tris, int nt,
                       const uint32_t* mask_bits, int n_mask_words);
/* RayCast(accel_mesh, scene_transform, pos, check_mask), ray_casting.cc:128-133, for n
 * image positions under one camera.  Outputs per ray: hit flag, position (object space),
 * primitive id, barycentric (u,v), t. */
PC_API int pc_ray_cast(pc_ctx*, const float model[16], const pc_camera_state* cam, const float* pos,
                       int n, int check_mask, uint8_t* hit_out, float* pos_out, uint32_t* prim_out,
                       float* uv_out, float* t_out);
/* SolvePnPIterative, pnp/solvers.cc:11-78: robust LM pose (+intrinsics) refinement.
 * X: m x 3 world points, x: m x 2 image points, weights m or NULL.  cam is in/out. */
PC_API int pc_solve_pnp(pc_ctx*, const float* X, const float* x, const float* weights, int m,
                        const pc_bundle_opts*, float max_inlier_error, int optimize_focal_length,
                        int optimize_principal_point, pc_camera_state* cam, pc_bundle_stats* stats,
                        float* inlier_ratio);
/* One source frame's contribution to SolveFrame (tracker.cc:45-93). */
typedef struct pc_match_source {
    pc_camera_state camera;          /* pose/intrinsics of the source frame */
    const float* keypoints;          /* source keypoints, nk x 2 (host) */
    int32_t nk;
    const uint32_t* src_kps_indices; /* flow rows (host) */
    const float* tgt_kps;
    int32_t rows;
} pc_match_source;
/* SolveFrame, tracker.cc:36-131: ray-cast the sources' matched keypoints onto the mesh,
 * then SolvePnPIterative from `init`.  Returns PC_ERR_NOT_ENOUGH_FEATURES if < 3 hits. */
PC_API int pc_track_frame(pc_ctx*, const pc_match_source* srcs, int nsrc, const float model[16],
                          const pc_camera_state* init, const pc_bundle_opts*, int optimize_focal_length,
                          int optimize_principal_point, pc_camera_state* out, pc_bundle_stats* stats,
                          float* inlier_ratio, int* num_matches);

/* ---- refine (refiner.cc:71-725, pnp/lev_marq.h:391-871) ----------------------------- */
typedef struct pc_ba_edge {
    int32_t src_frame_idx, tgt_frame_idx; /* indices into the trajectory segment */
    int32_t first_row, rows;              /* rows of the shared match arrays */
} pc_ba_edge;
typedef struct pc_ba_problem {
    int32_t num_frames;                   /* trajectory segment length (>= 3) */
    const int32_t* kp_offsets;            /* num_frames+1 offsets into keypoints */
    const float* keypoints;               /* filtered keypoints of all frames, x,y */
    int32_t num_edges;
    const pc_ba_edge* edges;              /* CachedDatabase flows, refiner.cc:97-161 */
    const uint32_t* src_kps_indices;      /* per match: index into the src frame's kps */
    const float* tgt_kps;                 /* per match: x,y */
    float model[16];                      /* row-major model matrix */
    int optimize_focal_length, optimize_principal_point;
} pc_ba_problem;
PC_API int pc_ba_load(pc_ctx*, const pc_ba_problem*);
/* TotalCost, lev_marq.h:773-824 (updates the primitive-id cache like refiner.cc:323-350). */
PC_API int pc_ba_cost(pc_ctx*, const pc_camera_state* traj, const pc_bundle_opts*, float* cost_out);
/* BuildNormalEquations, lev_marq.h:653-771.  JtJ_blocks: for each frame i the lower
 * block row [i-8..i] as dense p x p blocks (block (i,j) at ((i*9)+(i-j))*p*p, row-major,
 * p = 6 or 9); Jtr: num_frames*p.  Either may be NULL. */
PC_API int pc_ba_normal_equations(pc_ctx*, const pc_camera_state* traj, const pc_bundle_opts*,
                                  float* JtJ_blocks, float* Jtr);
/* The per-(frame, keypoint) primitive-id cache of RefinementProblemBase (refiner.cc:235-242,
 * 547-559), in the order of pc_ba_problem.keypoints; 0xFFFFFFFF = no intersection cached. */
PC_API int pc_ba_read_cache(pc_ctx*, uint32_t* out, int cap);
/* ComputeStep (lev_marq.h:826-841) on the normal equations the last pc_ba_normal_equations left on the device:
 * step = -(JtJ with diag * (1 + lambda))^-1 Jtr through the block-banded Cholesky kernel (K14).  step_out:
 * num_frames * p floats (p = 6, or 9 with intrinsics); PC_ERR_STATE when a pivot is not positive. */
PC_API int pc_ba_solve_step(pc_ctx*, float lambda, float* step_out, float* step_norm_out);
typedef int (*pc_ba_iter_cb)(const pc_bundle_stats*, void* user);
/* LevMarqSparseSolve over the loaded problem, lev_marq.h:492-588; traj is in/out. */
PC_API int pc_ba_solve(pc_ctx*, const pc_bundle_opts*, pc_camera_state* traj, pc_bundle_stats* stats,
                       pc_ba_iter_cb cb, void* user);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink -------------------------
 * Rank 0 makes the id, the launcher carries its bytes to the other ranks, every rank joins with its context.
 * libnccl.so.2 is opened at run time; without it these calls fail with PC_ERR_STATE. */
#define PC_COMM_ID_BYTES 128
PC_API int pc_comm_unique_id(uint8_t id_out[PC_COMM_ID_BYTES]);
PC_API int pc_comm_init(pc_ctx*, int world, int rank, const uint8_t id[PC_COMM_ID_BYTES]);
PC_API int pc_comm_destroy(pc_ctx*);
/* The path's one collective: stitches per-GPU trajectory segments before the global refine (the addon's segment
 * notion, blender_addon/operators/tracking.py:103-109; TrackSequence per segment, tracker.cc:194-213).  `local`:
 * this rank's n_local records; counts[world]: every rank's segment length; `all`: sum(counts) records in rank order. */
PC_API int pc_traj_allgather(pc_ctx*, const pc_camera_state* local, int n_local, const int* counts, pc_camera_state* all);
/* Edge-sharded refine: after pc_comm_init + pc_ba_load (in that order) every rank evaluates a contiguous share of
 * the edges (BuildNormalEquations / TotalCost, lev_marq.h:653-824); the per-edge blocks and costs are all-gathered
 * once per LM iteration each and every rank then assembles, factors and decides identically.  The result equals the
 * single-GPU pc_ba_solve bit for bit. */
PC_API int pc_ba_set_edge_shard(pc_ctx*, int on);

#ifdef __cplusplus
}
#endif
#endif /* POLYCHASE_B200_H_ */
