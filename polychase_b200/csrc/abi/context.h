// pc_ctx: device, streams, resident frame ring and scratch buffers behind the C ABI.
#pragma once

#include <cuda_runtime.h>

#include <deque>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/polychase_b200.h"
#include "../kernels/kernels.h"
#include "../kernels/track_kernels.h"

namespace pc {

struct FrameSlot {
    bool used = false;
    int32_t frame_id = 0;
    uint64_t stamp = 0;
    int w = 0, h = 0, levels = 0;
    uint8_t* base = nullptr;      // the slot's one allocation
    Image8 level[kMaxLevels];     // allocated once at the context's max size (data = pixel (0,0) inside the apron)
    float* kps = nullptr;         // device, max_features x 2
    int* n_kps = nullptr;         // device count
    int* n_accepted = nullptr;    // device: kept corners before the max_corners cut
    int* greedy_remaining = nullptr;
    int n_kps_host = -1;          // -1 = not known on host yet
    bool has_kps = false;
    // streaming analyzer, 10x10 window: the LK source templates of this frame's keypoints (lk10.cu)
    uint4* tmpl = nullptr;
    float* tmpl_sums = nullptr;
    int tmpl_levels = 0;          // levels the allocation holds
    bool has_tmpl = false;        // computed for the frame and keypoints the slot holds now
    bool tmpl_queue_layout = false;   // which kernel wrote them (lk10q.cu's layout or lk10.cu's)
    int* order = nullptr;             // spatial order of the keypoints (launch_spatial_order), valid with has_tmpl
    bool has_order = false;
};

enum KernelFamily { KF_GRAY_PYR = 0, KF_MIN_EIG, KF_SELECT, KF_LK, KF_COMPACT, KF_RAYCAST, KF_PNP, KF_BA, KF_LK_TMPL, KF_COUNT };

struct TimedSpan {
    int family;
    cudaEvent_t start, stop;
};

struct PairOut {              // compacted LK rows of one pair (device) + pinned mirror
    uint32_t* idx = nullptr;
    float* tgt = nullptr;
    float* err = nullptr;
    int* count = nullptr;
};

struct Stage {                // one in-flight frame of the streaming analyzer
    uint8_t* rgb_dev = nullptr;       // staging for host-provided frames
    size_t rgb_bytes = 0;
    cudaStream_t det_stream = nullptr;   // the detector stream this frame's pyramid + detector were queued on
    cudaEvent_t templates = nullptr;     // this frame's LK source templates are written (recorded on its LK stream)
    bool templates_recorded = false;
    PairOut dev[8], host[8];          // views into the two slabs below
    // One device slab and one pinned mirror per stage: [8 row counts | pair 0: idx, tgt, err | pair 1 ...],
    // so a frame's result travels in one copy (capi.cu: alloc_stage_rows).
    uint8_t* rows_dev = nullptr;
    uint8_t* rows_host = nullptr;
    size_t rows_bytes = 0;
    bool rows_downloaded = false;     // the row slab + keypoints were queued with the counts (single-phase pop)
    int32_t from[8], to[8];
    int num_pairs = 0;
    float* kps_host = nullptr;        // pinned
    int* counts_host = nullptr;       // pinned: [0]=n_kps [1]=n_accepted [2]=greedy_remaining
    int32_t frame_id = 0;
    cudaEvent_t uploaded = nullptr, gray_done = nullptr, detected = nullptr, computed = nullptr, downloaded = nullptr;
    bool is_halo = false;
    bool busy = false;
    bool gray_pending = false;
    // fused analyze -> track chain (track.cu): this frame's pose solve
    int track_state = 0;              // 0 none, 1 solved on the device, 2 seeded (known pose)
    PnpResult* trk_result_dev = nullptr;
    PnpResult* trk_result_host = nullptr;        // pinned
    pc_camera_state* trk_cam_host = nullptr;     // pinned
    cudaEvent_t tracked = nullptr;
};

// Device-resident forward tracking sweep chained behind the analyzer (SolveFrame,
// /root/reference/cpp/tracker.cc:36-131, called per frame by TrackCameraTrajectory :133-192).
constexpr int kCamRing = 32;                     // >= 8 (largest skip) + pipeline depth
struct TrackChain {
    bool on = false;
    float model[16];
    pc_bundle_opts bo{};
    int opt_f = 0, opt_pp = 0;
    cudaStream_t stream = nullptr;               // high priority: the latency-bound chain goes first
    cudaEvent_t join = nullptr;                  // pc_mark joins this stream into the compute stream
    pc_camera_state* d_cams = nullptr;           // kCamRing slots, slot = frame_id mod kCamRing
    int32_t cam_frame[kCamRing];                 // host: which frame's pose the slot holds
    bool cam_known[kCamRing];
    std::unordered_map<int32_t, pc_camera_state> seeds;   // known poses, uploaded when their frame is pushed
    bool have_bounds = false;
    Bounds bounds{};                             // solvers.cc:19-21, from the first seed's intrinsics
    float* d_X = nullptr; float* d_x = nullptr; uint8_t* d_valid = nullptr;
    size_t cap_rows = 0;
};

// One set of detector scratch (a detector stream owns one).
struct DetScratch {
    float* eig = nullptr; int eig_pitch = 0;
    uint8_t* state = nullptr; int state_pitch = 0;
    int* cell_max = nullptr;
    unsigned long long* cand = nullptr; int cand_cap = 0; int* cand_count = nullptr;
    int* det_zero = nullptr; int det_zero_ints = 0;
    SelectWorkspace sel{};
};

struct MeshData;   // track.cu
struct BAData;     // ba.cu
struct CommData;   // comm.cu

}  // namespace pc

struct pc_ctx {
    pc_limits lim{};
    int device = 0;
    int sm_count = 0;
    cudaStream_t compute = nullptr, h2d = nullptr, d2h = nullptr;
    cudaStream_t d2h_rows = nullptr;     // pop-time row downloads: never queued behind a later frame's counts
    cudaStream_t h2d_extra[3] = {nullptr, nullptr, nullptr};   // PC_H2D_SPLIT: extra upload streams (frame slices)
    cudaEvent_t h2d_extra_ev[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t h2d_fork = nullptr;
    bool download_hint = false;          // the last pop asked for rows: queue the next frames' rows with their counts
    // streaming analyzer: LK batches run on this low-priority stream next to the following frame's
    // pyramid + detector on `compute` (nullptr = everything on `compute`)
    cudaStream_t side = nullptr;
    // extra detector streams (+ scratch sets, allocated by the first analyze pass): frames take `compute` and the
    // extra streams in turn, so that one frame's latency-bound selection chain (a 16-CTA cluster) overlaps the next
    // frames' pyramid / min-eig / NMS instead of serialising the detector.  PC_DET_STREAMS=n (1..4, default 3: resident throughput is the same from 2 up, e2e gains 1.8 % from the third, profiles/r2_y_det_streams_sweep.json).
    static constexpr int kMaxExtraDet = 3;
    int n_det_extra = 0;
    cudaStream_t det_stream[kMaxExtraDet] = {nullptr, nullptr, nullptr};
    cudaEvent_t det_join[kMaxExtraDet] = {nullptr, nullptr, nullptr};
    pc::DetScratch* det_set[kMaxExtraDet] = {nullptr, nullptr, nullptr};
    // second LK stream (+ dense scratch set): the batches of consecutive frames alternate between `side` and `side2`,
    // so one batch's tail wave, template launch and compaction overlap the next batch.  PC_LK_STREAMS=1 disables it.
    cudaStream_t side2 = nullptr;
    cudaEvent_t join_e = nullptr;
    float* lk_next2 = nullptr; uint8_t* lk_status2 = nullptr; float* lk_err2 = nullptr;
    std::string err;
    uint64_t launches = 0;
    uint64_t stamp = 0;

    std::vector<pc::FrameSlot> slots;

    // detector scratch (one compute stream -> one set)
    float* eig = nullptr; int eig_pitch = 0;
    uint8_t* state = nullptr; int state_pitch = 0;
    int* cell_max = nullptr; int cell_cap = 0;
    unsigned long long* cand = nullptr; int cand_cap = 0; int* cand_count = nullptr;
    int* det_zero = nullptr; int det_zero_ints = 0;   // counter block cleared at the start of every detector run
    pc::SelectWorkspace sel{};

    // LK dense scratch: [8][cap]
    float* lk_next = nullptr; uint8_t* lk_status = nullptr; float* lk_err = nullptr;
    // 10x10 LK as a work queue (lk10q.cu), opt-in with PC_LK_QUEUE=1: measured slower than the lock-step kernel on
    // the benchmark clip (profiles/r2_ij_lk_queue_ab.json).  lk_queue = the launch's item counter;
    // PC_LK_BUDGET=n makes a block leave after n items (0: stay until the queue is empty)
    int* lk_queue = nullptr;
    bool lk_queue_mode = false;
    int lk_queue_budget = 0;
    pc::PairOut sync_out;           // outputs of the synchronous pc_lk_pair
    uint8_t* rgb_scratch = nullptr; size_t rgb_scratch_bytes = 0;   // synchronous uploads

    // streaming analyzer
    bool analyzing = false;
    pc_video_info vinfo{};
    pc_gftt_opts gopts{};
    pc_flow_opts fopts{};
    std::vector<pc::Stage> stages;
    std::deque<int> inflight;       // stage indices in push order
    int next_stage = 0;
    int32_t last_pushed = 0; bool any_pushed = false;
    int halo_frames = 0, pushed_count = 0;
    cudaEvent_t marks[8] = {nullptr}; cudaEvent_t join_a = nullptr, join_b = nullptr, join_c = nullptr;
    std::unordered_map<int32_t, std::vector<float>> preset_kps;

    // synthetic texture
    uint8_t* tex = nullptr; int tex_w = 0, tex_h = 0, tex_pitch = 0;

    // timing
    bool timing = false;
    std::vector<pc::TimedSpan> spans;
    std::vector<cudaEvent_t> event_pool;
    double fam_ms[pc::KF_COUNT] = {0};
    uint64_t fam_n[pc::KF_COUNT] = {0};

    // track / refine state
    pc::MeshData* mesh = nullptr;
    pc::BAData* ba = nullptr;
    pc::TrackChain* track = nullptr;
    pc::CommData* comm = nullptr;

    ~pc_ctx();
};

namespace pc {

// Set ctx->err and return `code`.
int fail(pc_ctx* c, int code, const std::string& msg);
int cuda_fail(pc_ctx* c, cudaError_t e, const char* what, const char* file, int line);
void set_global_error(const std::string& msg);

#define PC_CUDA(ctx, expr)                                                         \
    do {                                                                           \
        cudaError_t e__ = (expr);                                                  \
        if (e__ != cudaSuccess) return ::pc::cuda_fail(ctx, e__, #expr, __FILE__, __LINE__); \
    } while (0)

#define PC_CHECK(ctx, cond, msg)                                                   \
    do {                                                                           \
        if (!(cond)) return ::pc::fail(ctx, PC_ERR_INVALID, std::string("check failed: ") + #cond + " -- " + (msg)); \
    } while (0)

// RAII-ish timing span helpers (no-ops unless ctx->timing)
void span_begin(pc_ctx* c, int family, cudaStream_t s);
void span_end(pc_ctx* c, cudaStream_t s);
int check_launch(pc_ctx* c, const char* what, int n_kernels);

FrameSlot* find_slot(pc_ctx* c, int32_t frame_id);
PyramidView view_of(const FrameSlot& f);

void free_mesh(MeshData*);
void free_track_chain(TrackChain*);
// fused chain hooks used by the streaming analyzer (capi.cu)
int track_chain_enqueue(pc_ctx* c, Stage& st, int32_t frame_id, bool is_halo, int cap);
int track_chain_collect(pc_ctx* c, Stage& st, pc_frame_result* out);
void free_ba(BAData*);
void free_comm(CommData*);

}  // namespace pc
