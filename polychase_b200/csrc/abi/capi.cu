// C ABI: context management and the analyze path (frames, detector, LK, streaming
// analyzer).  See include/polychase_b200.h for the reference interfaces each entry point
// replaces.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "context.h"

#include <ctype.h>
#include <stdio.h>
#ifdef __linux__
#include <sys/syscall.h>
#include <unistd.h>
#endif

namespace pc {

static std::mutex g_err_mtx;
static std::string g_err;

void set_global_error(const std::string& msg) {
    std::lock_guard<std::mutex> lk(g_err_mtx);
    g_err = msg;
}

int fail(pc_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    else set_global_error(msg);
    return code;
}

int cuda_fail(pc_ctx* c, cudaError_t e, const char* what, const char* file, int line) {
    std::string m = std::string("CUDA error ") + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ") at " + file +
                    ":" + std::to_string(line) + " in " + what;
    return fail(c, e == cudaErrorMemoryAllocation ? PC_ERR_NOMEM : PC_ERR_CUDA, m);
}

static cudaEvent_t get_event(pc_ctx* c) {
    if (!c->event_pool.empty()) {
        cudaEvent_t e = c->event_pool.back();
        c->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

void span_begin(pc_ctx* c, int family, cudaStream_t s) {
    if (!c->timing) return;
    TimedSpan sp;
    sp.family = family;
    sp.start = get_event(c);
    sp.stop = get_event(c);
    cudaEventRecord(sp.start, s);
    c->spans.push_back(sp);
}

void span_end(pc_ctx* c, cudaStream_t s) {
    if (!c->timing || c->spans.empty()) return;
    cudaEventRecord(c->spans.back().stop, s);
}

int check_launch(pc_ctx* c, const char* what, int n_kernels) {
    c->launches += (uint64_t)n_kernels;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(c, e, what, __FILE__, __LINE__);
    return PC_OK;
}

FrameSlot* find_slot(pc_ctx* c, int32_t frame_id) {
    for (auto& s : c->slots)
        if (s.used && s.frame_id == frame_id) return &s;
    return nullptr;
}

PyramidView view_of(const FrameSlot& f) {
    PyramidView v{};
    v.levels = f.levels;
    for (int i = 0; i < kMaxLevels; i++) {
        v.data[i] = f.level[i].data;
        v.w[i] = f.level[i].w;
        v.h[i] = f.level[i].h;
        v.pitch[i] = f.level[i].pitch;
    }
    return v;
}

static FrameSlot* acquire_slot(pc_ctx* c, int32_t frame_id) {
    FrameSlot* s = find_slot(c, frame_id);
    if (!s) {
        for (auto& t : c->slots)
            if (!t.used) { s = &t; break; }
    }
    if (!s) {  // evict the least recently uploaded frame
        s = &c->slots[0];
        for (auto& t : c->slots)
            if (t.stamp < s->stamp) s = &t;
    }
    s->used = true;
    s->frame_id = frame_id;
    s->stamp = ++c->stamp;
    s->has_kps = false;
    s->has_tmpl = false;
    s->has_order = false;
    s->n_kps_host = -1;
    return s;
}

// cv::buildOpticalFlowPyramid level rule: after producing level L the next size is
// ((w+1)/2, (h+1)/2); the pyramid stops at L if that size is <= winSize in either dimension.
static int plan_levels(FrameSlot* f, int w, int h, int win, int max_level) {
    int levels = 0;
    int lw = w, lh = h;
    for (int L = 0; L <= max_level && L < kMaxLevels; L++) {
        f->level[L].w = lw;
        f->level[L].h = lh;
        levels = L + 1;
        lw = (lw + 1) / 2;
        lh = (lh + 1) / 2;
        if (lw <= win || lh <= win) break;
    }
    return levels;
}

static int validate_flow_opts(pc_ctx* c, const pc_flow_opts* o) {
    extern bool lk_window_supported(int);
    PC_CHECK(c, o != nullptr, "flow options are required");
    PC_CHECK(c, lk_window_supported(o->window_size), "window_size must be in [3,16]");
    PC_CHECK(c, o->max_level >= 0 && o->max_level < kMaxLevels, "max_level must be in [0,5]");
    return PC_OK;
}

static int validate_gftt_opts(pc_ctx* c, const pc_gftt_opts* o) {
    PC_CHECK(c, o != nullptr, "detector options are required");
    // gftt.cc:19-20
    PC_CHECK(c, o->quality_level > 0 && o->min_distance >= 0 && o->max_corners >= 0,
             "options.quality_level > 0 && options.min_distance >= 0 && options.max_corners >= 0");
    PC_CHECK(c, o->block_size == 3 && o->gradient_size == 3, "only block_size 3 / gradient_size 3 are built");
    PC_CHECK(c, !o->use_harris, "use_harris is out of scope (never enabled by the addon)");
    const int gr = std::max(1, o->grid_rows), gc = std::max(1, o->grid_cols);
    PC_CHECK(c, gr * gc <= c->cell_cap, "grid_rows*grid_cols too large (limit 1024)");
    return PC_OK;
}

// Builds gray + pyramid of slot f from an RGB (channels==3) or gray (channels==1) device image.
static int build_pyramid(pc_ctx* c, FrameSlot* f, const uint8_t* img_dev, size_t stride, int channels, int w, int h,
                         const pc_flow_opts* fo, cudaStream_t s) {
    f->w = w;
    f->h = h;
    f->levels = plan_levels(f, w, h, fo->window_size, fo->max_level);
    span_begin(c, KF_GRAY_PYR, s);
    int done = 0, n_launch = 0;
    static const bool no_tma = getenv("PC_NO_TMA_PYRAMID") != nullptr;
    if (channels == 3 && !no_tma) done = launch_pyramid_tma(img_dev, stride, f->level, f->levels, s, &n_launch);
    if (done == 0) {
        if (channels == 3) launch_rgb_to_gray(img_dev, stride, f->level[0], s);
        else launch_copy_gray(img_dev, stride, f->level[0], s);
        done = 1;
        n_launch += 1 + (channels == 3 && (w % 16) ? 1 : 0);
    }
    for (int L = done; L < f->levels; L++, n_launch++) launch_pyr_down(f->level[L - 1], f->level[L], s);
    launch_pad_border(f->level, f->levels, s);
    span_end(c, s);
    return check_launch(c, "gray+pyramid", n_launch + 1);
}

// One set of detector scratch for frames up to W x H (a detector stream owns one set).
static int alloc_det_scratch(int W, int H, DetScratch& d) {
    d.eig_pitch = (W + 31) / 32 * 32;
    PC_CUDA(nullptr, cudaMalloc(&d.eig, sizeof(float) * (size_t)d.eig_pitch * H));
    d.state_pitch = (W + 127) / 128 * 128;
    PC_CUDA(nullptr, cudaMalloc(&d.state, (size_t)d.state_pitch * H));
    PC_CUDA(nullptr, cudaMalloc(&d.cell_max, sizeof(int) * 1024));
    // 3x3 NMS leaves at most one candidate per 2x2 block except on plateaus; w*h/4 is ample
    d.cand_cap = std::max(1024, (int)(((size_t)W * H) / 4));
    PC_CUDA(nullptr, cudaMalloc(&d.cand, sizeof(unsigned long long) * d.cand_cap));
    // one block holds every detector counter that must be zero at the start of a frame, so the init
    // launch of the min-eig stage clears them all: [0] candidate count, [8..15] select scratch,
    // then the greedy round counters, the two 4096-bin value histograms and the short list's bin cursors
    d.det_zero_ints = 16 + 3 * kMaxGreedyRounds + 3 * 4096;
    PC_CUDA(nullptr, cudaMalloc(&d.det_zero, sizeof(int) * d.det_zero_ints));
    PC_CUDA(nullptr, cudaMemset(d.det_zero, 0, sizeof(int) * d.det_zero_ints));
    d.cand_count = d.det_zero;
    d.sel.cap = d.cand_cap;
    PC_CUDA(nullptr, cudaMalloc(&d.sel.accepted, sizeof(unsigned long long) * d.cand_cap));
    d.sel.sorted_cap = 1;
    while (d.sel.sorted_cap < d.cand_cap) d.sel.sorted_cap <<= 1;
    PC_CUDA(nullptr, cudaMalloc(&d.sel.sorted, sizeof(unsigned long long) * d.sel.sorted_cap));
    d.sel.sel = d.det_zero + 8;
    d.sel.round_counters = d.det_zero + 16;
    d.sel.hist = d.sel.round_counters + 3 * kMaxGreedyRounds;
    d.sel.kept_hist = d.sel.hist + 4096;
    d.sel.bin_cursor = d.sel.kept_hist + 4096;
    PC_CUDA(nullptr, cudaMalloc(&d.sel.bin_start, sizeof(int) * 4096));
    PC_CUDA(nullptr, cudaMalloc(&d.sel.strong, sizeof(unsigned long long) * d.cand_cap));
    d.sel.cub_temp_bytes = select_cub_temp_bytes(d.cand_cap);
    PC_CUDA(nullptr, cudaMalloc(&d.sel.cub_temp, d.sel.cub_temp_bytes));
    return PC_OK;
}

static void free_det_scratch(DetScratch& d) {
    cudaFree(d.eig); cudaFree(d.state); cudaFree(d.cell_max); cudaFree(d.cand); cudaFree(d.det_zero);
    cudaFree(d.sel.accepted); cudaFree(d.sel.sorted); cudaFree(d.sel.cub_temp); cudaFree(d.sel.strong); cudaFree(d.sel.bin_start);
    d = DetScratch{};
}

static DetScratch primary_scratch(const pc_ctx* c) {
    DetScratch d;
    d.eig = c->eig; d.eig_pitch = c->eig_pitch; d.state = c->state; d.state_pitch = c->state_pitch;
    d.cell_max = c->cell_max; d.cand = c->cand; d.cand_cap = c->cand_cap; d.cand_count = c->cand_count;
    d.det_zero = c->det_zero; d.det_zero_ints = c->det_zero_ints; d.sel = c->sel;
    return d;
}

static int run_detector(pc_ctx* c, FrameSlot* f, const pc_gftt_opts* go, cudaStream_t s, const DetScratch* scratch = nullptr) {
    const DetScratch d = scratch ? *scratch : primary_scratch(c);
    DetectGrid g;
    g.grid_rows = std::max(1, go->grid_rows);
    g.grid_cols = std::max(1, go->grid_cols);
    g.block_h = (f->h + g.grid_rows - 1) / g.grid_rows;   // gftt.cc:42-43
    g.block_w = (f->w + g.grid_cols - 1) / g.grid_cols;
    span_begin(c, KF_MIN_EIG, s);
    launch_min_eig(f->level[0], d.eig, d.eig_pitch, g, d.cell_max, d.det_zero, d.det_zero_ints, f->n_kps, s);
    span_end(c, s);
    span_begin(c, KF_SELECT, s);
    launch_nms_candidates(d.eig, d.eig_pitch, f->w, f->h, g, d.cell_max, go->quality_level, d.state,
                          d.state_pitch, d.cand, d.cand_cap, d.cand_count, d.sel.hist, s);
    SelectWorkspace ws = d.sel;
    ws.accepted_count = f->n_accepted;
    ws.remaining = f->greedy_remaining;
    launch_select(d.cand, d.cand_count, d.cand_cap, d.eig, d.eig_pitch, d.state, d.state_pitch, f->w, f->h,
                  go->min_distance, go->max_corners, ws, f->kps, c->lim.max_features, f->n_kps, c->sm_count, s);
    span_end(c, s);
    // the candidate count lives in the shared detector scratch, which the next frame's detector clears: keep
    // this frame's value with its other counters so that an overflow is still seen when the frame is popped
    PC_CUDA(c, cudaMemcpyAsync(f->n_kps + 3, d.cand_count, sizeof(int), cudaMemcpyDeviceToDevice, s));
    f->has_kps = true;
    f->n_kps_host = -1;
    return check_launch(c, "detector", 6);
}

static LKParams make_lk_params(const pc_flow_opts* fo) {
    LKParams p;
    p.win = fo->window_size;
    p.max_level = fo->max_level;
    // cv::calcOpticalFlowPyrLK clamps the criteria: maxCount to [0,100], epsilon to [0,10]
    p.iters = std::min(std::max(fo->term_max_iters, 0), 100);
    p.eps = std::min(std::max(fo->term_epsilon, 0.0), 10.0);
    p.min_eig = fo->min_eigen_threshold;
    return p;
}

static void fill_pair(pc_ctx* c, LKPair& p, const FrameSlot& a, const FrameSlot& b, int k, const PairOut& out,
                      bool use_templates = false, bool second_set = false) {
    const int cap = c->lim.max_features;
    p.a = view_of(a);
    p.b = view_of(b);
    p.tmpl = LKTemplates{nullptr, nullptr, 0};
    if (use_templates && a.has_tmpl) p.tmpl = LKTemplates{a.tmpl, a.tmpl_queue_layout ? nullptr : a.tmpl_sums, cap};
    p.order = (use_templates && a.has_tmpl && a.has_order) ? a.order : nullptr;
    p.pts = a.kps;
    p.n_pts = a.n_kps;
    p.next = (second_set ? c->lk_next2 : c->lk_next) + (size_t)k * cap * 2;
    p.status = (second_set ? c->lk_status2 : c->lk_status) + (size_t)k * cap;
    p.err = (second_set ? c->lk_err2 : c->lk_err) + (size_t)k * cap;
    p.out_idx = out.idx;
    p.out_tgt = out.tgt;
    p.out_err = out.err;
    p.out_count = out.count;
}

static int alloc_pair_out(pc_ctx* c, PairOut& o, int cap, bool pinned) {
    if (pinned) {
        PC_CUDA(c, cudaMallocHost(&o.idx, sizeof(uint32_t) * cap));
        PC_CUDA(c, cudaMallocHost(&o.tgt, sizeof(float) * 2 * cap));
        PC_CUDA(c, cudaMallocHost(&o.err, sizeof(float) * cap));
        PC_CUDA(c, cudaMallocHost(&o.count, sizeof(int)));
    } else {
        PC_CUDA(c, cudaMalloc(&o.idx, sizeof(uint32_t) * cap));
        PC_CUDA(c, cudaMalloc(&o.tgt, sizeof(float) * 2 * cap));
        PC_CUDA(c, cudaMalloc(&o.err, sizeof(float) * cap));
        PC_CUDA(c, cudaMalloc(&o.count, sizeof(int)));
    }
    return PC_OK;
}

static void free_pair_out(PairOut& o, bool pinned) {
    if (pinned) {
        cudaFreeHost(o.idx); cudaFreeHost(o.tgt); cudaFreeHost(o.err); cudaFreeHost(o.count);
    } else {
        cudaFree(o.idx); cudaFree(o.tgt); cudaFree(o.err); cudaFree(o.count);
    }
    o = PairOut{};
}

// Row slab of one analyzer stage: [8 counts, padded to 128 B | pair k: idx (cap u32), tgt (2 cap f32),
// err (cap f32)] with cap rounded to 32 rows so every array starts on a 128-byte line.
static int alloc_stage_rows(pc_ctx* c, Stage& st, int cap) {
    const size_t capr = ((size_t)cap + 31) / 32 * 32;
    const size_t pair_bytes = capr * 16;
    st.rows_bytes = 128 + 8 * pair_bytes;
    PC_CUDA(c, cudaMalloc(&st.rows_dev, st.rows_bytes));
    PC_CUDA(c, cudaMallocHost(&st.rows_host, st.rows_bytes));
    PC_CUDA(c, cudaMemset(st.rows_dev, 0, 128));
    memset(st.rows_host, 0, 128);
    for (int side = 0; side < 2; side++) {
        uint8_t* base = side ? st.rows_host : st.rows_dev;
        PairOut* o = side ? st.host : st.dev;
        for (int k = 0; k < 8; k++) {
            uint8_t* p = base + 128 + k * pair_bytes;
            o[k].count = reinterpret_cast<int*>(base) + k;
            o[k].idx = reinterpret_cast<uint32_t*>(p);
            o[k].tgt = reinterpret_cast<float*>(p + capr * 4);
            o[k].err = reinterpret_cast<float*>(p + capr * 12);
        }
    }
    return PC_OK;
}

static void free_stage_rows(Stage& st) {
    cudaFree(st.rows_dev);
    cudaFreeHost(st.rows_host);
    st.rows_dev = st.rows_host = nullptr;
    st.rows_bytes = 0;
    for (int k = 0; k < 8; k++) st.dev[k] = st.host[k] = PairOut{};
}

}  // namespace pc

using namespace pc;

pc_ctx::~pc_ctx() {
    cudaSetDevice(device);
    if (compute) cudaStreamSynchronize(compute);
    if (h2d) cudaStreamSynchronize(h2d);
    if (d2h) cudaStreamSynchronize(d2h);
    if (d2h_rows) cudaStreamSynchronize(d2h_rows);
    if (side) cudaStreamSynchronize(side);
    for (int k = 0; k < n_det_extra; k++) cudaStreamSynchronize(det_stream[k]);
    if (side2) cudaStreamSynchronize(side2);
    if (track && track->stream) cudaStreamSynchronize(track->stream);
    cudaFree(lk_next2); cudaFree(lk_status2); cudaFree(lk_err2);
    for (auto& d : det_set)
        if (d) { free_det_scratch(*d); delete d; d = nullptr; }
    for (auto& s : slots) {
        cudaFree(s.tmpl);
        cudaFree(s.tmpl_sums);
        cudaFree(s.base);
        cudaFree(s.kps); cudaFree(s.n_kps); cudaFree(s.order);
    }
    cudaFree(eig); cudaFree(state); cudaFree(cell_max); cudaFree(cand); cudaFree(det_zero);
    cudaFree(sel.accepted); cudaFree(sel.sorted); cudaFree(sel.cub_temp); cudaFree(sel.strong); cudaFree(sel.bin_start);
    cudaFree(lk_next); cudaFree(lk_status); cudaFree(lk_err); cudaFree(lk_queue);
    free_pair_out(sync_out, false);
    cudaFree(rgb_scratch);
    for (auto& st : stages) {
        cudaFree(st.rgb_dev);
        free_stage_rows(st);
        cudaFreeHost(st.kps_host); cudaFreeHost(st.counts_host);
        cudaFree(st.trk_result_dev); cudaFreeHost(st.trk_result_host); cudaFreeHost(st.trk_cam_host);
        if (st.tracked) cudaEventDestroy(st.tracked);
        if (st.uploaded) cudaEventDestroy(st.uploaded);
        if (st.gray_done) cudaEventDestroy(st.gray_done);
        if (st.detected) cudaEventDestroy(st.detected);
        if (st.templates) cudaEventDestroy(st.templates);
        if (st.computed) cudaEventDestroy(st.computed);
        if (st.downloaded) cudaEventDestroy(st.downloaded);
    }
    cudaFree(tex);
    for (auto& sp : spans) { cudaEventDestroy(sp.start); cudaEventDestroy(sp.stop); }
    for (auto e : event_pool) cudaEventDestroy(e);
    for (auto e : marks) if (e) cudaEventDestroy(e);
    if (join_a) cudaEventDestroy(join_a);
    if (join_b) cudaEventDestroy(join_b);
    if (join_c) cudaEventDestroy(join_c);
    if (track) free_track_chain(track);
    if (mesh) free_mesh(mesh);
    if (ba) free_ba(ba);
    if (comm) free_comm(comm);
    if (side) cudaStreamDestroy(side);
    for (int k = 0; k < kMaxExtraDet; k++) {
        if (det_stream[k]) cudaStreamDestroy(det_stream[k]);
        if (det_join[k]) cudaEventDestroy(det_join[k]);
    }
    if (side2) cudaStreamDestroy(side2);
    if (join_e) cudaEventDestroy(join_e);
    if (compute) cudaStreamDestroy(compute);
    if (h2d) cudaStreamDestroy(h2d);
    for (int k = 0; k < 3; k++) {
        if (h2d_extra[k]) cudaStreamDestroy(h2d_extra[k]);
        if (h2d_extra_ev[k]) cudaEventDestroy(h2d_extra_ev[k]);
    }
    if (h2d_fork) cudaEventDestroy(h2d_fork);
    if (d2h) cudaStreamDestroy(d2h);
    if (d2h_rows) cudaStreamDestroy(d2h_rows);
}

extern "C" {

void pc_default_gftt_opts(pc_gftt_opts* o) {
    o->quality_level = 0.01; o->min_distance = 5.0; o->block_size = 3; o->gradient_size = 3; o->max_corners = 0;
    o->use_harris = 0; o->harris_k = 0.04; o->grid_rows = 4; o->grid_cols = 4;
}
void pc_default_flow_opts(pc_flow_opts* o) {
    o->window_size = 10; o->max_level = 3; o->term_max_iters = 30; o->term_epsilon = 0.01;
    o->min_eigen_threshold = 1e-4;
}
void pc_default_bundle_opts(pc_bundle_opts* o) {
    o->max_iterations = 100; o->max_allowed_parallelism = 8; o->loss_type = 1; o->loss_scale = 1.0f;
    o->gradient_tol = 1e-10f; o->step_tol = 1e-8f; o->initial_lambda = 1e-5f; o->min_lambda = 1e-10f;
    o->max_lambda = 1e10f; o->verbose = 0;
}
const char* pc_version(void) { return "polychase_b200 0.1.0 (sm_100a)"; }

const char* pc_last_error(pc_ctx* c) {
    if (c) return c->err.c_str();
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lk(g_err_mtx);
    copy = g_err;
    return copy.c_str();
}

int pc_create(const pc_limits* limits, pc_ctx** out) {
    if (!out) return fail(nullptr, PC_ERR_INVALID, "pc_create: out is NULL");
    *out = nullptr;
    pc_limits lim{};
    if (limits) lim = *limits;
    if (lim.max_width <= 0) lim.max_width = 3840;
    if (lim.max_height <= 0) lim.max_height = 2160;
    if (lim.max_features <= 0) lim.max_features = 16384;
    if (lim.ring_frames <= 0) lim.ring_frames = 20;
    if (lim.pipeline_depth <= 0) lim.pipeline_depth = 8;
    if (lim.ring_frames < 9 + lim.pipeline_depth) lim.ring_frames = 9 + lim.pipeline_depth;

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, PC_ERR_CUDA,
                    std::string("pc_create: no CUDA device available (") + cudaGetErrorString(e) +
                        "); this library has no CPU fallback");
    if (lim.device < 0 || lim.device >= ndev) return fail(nullptr, PC_ERR_INVALID, "pc_create: bad device ordinal");
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, lim.device);
    if (prop.major != 10)
        return fail(nullptr, PC_ERR_CUDA,
                    std::string("pc_create: device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                        std::to_string(prop.minor) + "; this library is built for sm_100a only");
    std::unique_ptr<pc_ctx> c(new pc_ctx());
    c->lim = lim;
    c->device = lim.device;
    c->sm_count = prop.multiProcessorCount;
    pc_ctx* cp = c.get();
    PC_CUDA(nullptr, cudaSetDevice(lim.device));
    {
        // priorities: track chain (highest, track.cu) > detector / everything else > streaming LK batches.
        // The analyzer's LK batches run on their own stream, so frame j's batch (ALU / gather bound)
        // shares the SMs with frame j+1's pyramid + detector (HBM streaming, then short latency-bound
        // selection launches) instead of running back to back with them.
        int lo = 0, hi = 0;                        // numerically lower = higher priority
        PC_CUDA(nullptr, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        PC_CUDA(nullptr, cudaStreamCreateWithPriority(&cp->compute, cudaStreamNonBlocking, hi < lo ? hi + 1 : lo));
        const char* e = getenv("PC_LK_STREAM");     // PC_LK_STREAM=0: everything on the compute stream
        if (!e || atoi(e) != 0) PC_CUDA(nullptr, cudaStreamCreateWithPriority(&cp->side, cudaStreamNonBlocking, lo));
        const char* l2 = getenv("PC_LK_STREAMS");   // PC_LK_STREAMS=1: one LK stream
        if (cp->side && (!l2 || atoi(l2) >= 2)) {
            PC_CUDA(nullptr, cudaStreamCreateWithPriority(&cp->side2, cudaStreamNonBlocking, lo));
            PC_CUDA(nullptr, cudaEventCreateWithFlags(&cp->join_e, cudaEventDisableTiming));
        }
        const char* d = getenv("PC_DET_STREAMS");   // detector streams (1..4, default 3); 1: one detector stream
        const int n_det = cp->side ? std::min(std::max(d ? atoi(d) : 3, 1), 1 + pc_ctx::kMaxExtraDet) : 1;
        for (int k = 0; k + 1 < n_det; k++) {
            PC_CUDA(nullptr, cudaStreamCreateWithPriority(&cp->det_stream[k], cudaStreamNonBlocking, hi < lo ? hi + 1 : lo));
            PC_CUDA(nullptr, cudaEventCreateWithFlags(&cp->det_join[k], cudaEventDisableTiming));
            cp->n_det_extra = k + 1;
        }
    }
    PC_CUDA(nullptr, cudaStreamCreateWithFlags(&cp->h2d, cudaStreamNonBlocking));
    PC_CUDA(nullptr, cudaEventCreateWithFlags(&cp->h2d_fork, cudaEventDisableTiming));
    PC_CUDA(nullptr, cudaStreamCreateWithFlags(&cp->d2h, cudaStreamNonBlocking));
    PC_CUDA(nullptr, cudaStreamCreateWithFlags(&cp->d2h_rows, cudaStreamNonBlocking));

    const int W = lim.max_width, H = lim.max_height, cap = lim.max_features;
    // frame ring: all levels of a slot in one allocation
    cp->slots.resize(lim.ring_frames);
    for (auto& s : cp->slots) {
        size_t total = 0;
        int lw = W, lh = H;
        size_t offs[kMaxLevels];
        int pitches[kMaxLevels];
        for (int L = 0; L < kMaxLevels; L++) {                  // image + REFLECT_101 apron (kernels.h)
            pitches[L] = (lw + 2 * kPadX + 127) / 128 * 128;
            offs[L] = total;
            total += (size_t)pitches[L] * (lh + 2 * kPadY);
            total = (total + 255) / 256 * 256;
            lw = (lw + 1) / 2; lh = (lh + 1) / 2;
        }
        uint8_t* base = nullptr;
        total += 256;   // lk10.cu reads aligned words that may run a few bytes past the apron
        PC_CUDA(nullptr, cudaMalloc(&base, total));
        s.base = base;
        for (int L = 0; L < kMaxLevels; L++) {
            s.level[L].data = base + offs[L] + (size_t)kPadY * pitches[L] + kPadX;
            s.level[L].pitch = pitches[L];
        }
        PC_CUDA(nullptr, cudaMalloc(&s.kps, sizeof(float) * 2 * cap));
        PC_CUDA(nullptr, cudaMalloc(&s.order, sizeof(int) * (size_t)cap));
        PC_CUDA(nullptr, cudaMalloc(&s.n_kps, sizeof(int) * 4));
        PC_CUDA(nullptr, cudaMemset(s.n_kps, 0, sizeof(int) * 4));
        s.n_accepted = s.n_kps + 1;
        s.greedy_remaining = s.n_kps + 2;
    }
    {
        DetScratch d;
        int rc = alloc_det_scratch(W, H, d);
        if (rc) return rc;
        cp->eig = d.eig; cp->eig_pitch = d.eig_pitch; cp->state = d.state; cp->state_pitch = d.state_pitch;
        cp->cell_cap = 1024;      // = NMS_MAX_CELLS (mineig.cu)
        cp->cell_max = d.cell_max; cp->cand = d.cand; cp->cand_cap = d.cand_cap; cp->cand_count = d.cand_count;
        cp->det_zero = d.det_zero; cp->det_zero_ints = d.det_zero_ints; cp->sel = d.sel;
    }
    PC_CUDA(nullptr, cudaMalloc(&cp->lk_next, sizeof(float) * 2 * (size_t)cap * 8));
    PC_CUDA(nullptr, cudaMalloc(&cp->lk_status, (size_t)cap * 8));
    PC_CUDA(nullptr, cudaMalloc(&cp->lk_err, sizeof(float) * (size_t)cap * 8));
    PC_CUDA(nullptr, cudaMalloc(&cp->lk_queue, sizeof(int) * 4));
    PC_CUDA(nullptr, cudaMemset(cp->lk_queue, 0, sizeof(int) * 4));
    if (const char* e = getenv("PC_LK_QUEUE")) cp->lk_queue_mode = atoi(e) != 0;
    if (const char* e = getenv("PC_LK_BUDGET")) cp->lk_queue_budget = std::max(0, atoi(e));
    int rc = alloc_pair_out(nullptr, cp->sync_out, cap, false);
    if (rc) return rc;
    *out = c.release();
    return PC_OK;
}

void pc_destroy(pc_ctx* c) { delete c; }

int pc_synchronize(pc_ctx* c) {
    PC_CUDA(c, cudaSetDevice(c->device));
    PC_CUDA(c, cudaStreamSynchronize(c->h2d));
    PC_CUDA(c, cudaStreamSynchronize(c->compute));
    PC_CUDA(c, cudaStreamSynchronize(c->d2h));
    PC_CUDA(c, cudaStreamSynchronize(c->d2h_rows));
    if (c->side) PC_CUDA(c, cudaStreamSynchronize(c->side));
    for (int k = 0; k < c->n_det_extra; k++) PC_CUDA(c, cudaStreamSynchronize(c->det_stream[k]));
    if (c->side2) PC_CUDA(c, cudaStreamSynchronize(c->side2));
    if (c->track && c->track->stream) PC_CUDA(c, cudaStreamSynchronize(c->track->stream));
    return PC_OK;
}

uint64_t pc_kernel_launches(pc_ctx* c) { return c->launches; }

// ---- frames ---------------------------------------------------------------------------
static int frame_common(pc_ctx* c, int32_t frame_id, const uint8_t* img, int w, int h, size_t stride, int channels,
                        bool on_device, const pc_flow_opts* fo) {
    PC_CUDA(c, cudaSetDevice(c->device));
    int rc = validate_flow_opts(c, fo);
    if (rc) return rc;
    PC_CHECK(c, img != nullptr, "image pointer is NULL");
    PC_CHECK(c, w >= 16 && h >= 16, "frames smaller than 16x16 are not supported");
    PC_CHECK(c, w <= c->lim.max_width && h <= c->lim.max_height, "frame larger than the context limits");
    PC_CHECK(c, stride >= (size_t)w * channels, "stride smaller than a row");
    FrameSlot* f = acquire_slot(c, frame_id);
    const uint8_t* dev = img;
    if (!on_device) {
        const size_t bytes = stride * (size_t)h;
        if (c->rgb_scratch_bytes < bytes) {
            cudaFree(c->rgb_scratch);
            c->rgb_scratch = nullptr;
            c->rgb_scratch_bytes = 0;
            PC_CUDA(c, cudaMalloc(&c->rgb_scratch, bytes));
            c->rgb_scratch_bytes = bytes;
        }
        PC_CUDA(c, cudaMemcpyAsync(c->rgb_scratch, img, bytes, cudaMemcpyHostToDevice, c->compute));
        dev = c->rgb_scratch;
    }
    rc = build_pyramid(c, f, dev, stride, channels, w, h, fo, c->compute);
    if (rc) return rc;
    PC_CUDA(c, cudaStreamSynchronize(c->compute));
    return PC_OK;
}

int pc_frame_upload_rgb8(pc_ctx* c, int32_t frame_id, const uint8_t* rgb, int w, int h, size_t stride,
                         const pc_flow_opts* fo) {
    return frame_common(c, frame_id, rgb, w, h, stride, 3, false, fo);
}
int pc_frame_from_device_rgb8(pc_ctx* c, int32_t frame_id, const uint8_t* rgb, int w, int h, size_t stride,
                              const pc_flow_opts* fo) {
    return frame_common(c, frame_id, rgb, w, h, stride, 3, true, fo);
}
int pc_frame_upload_gray8(pc_ctx* c, int32_t frame_id, const uint8_t* gray, int w, int h, size_t stride,
                          const pc_flow_opts* fo) {
    return frame_common(c, frame_id, gray, w, h, stride, 1, false, fo);
}

int pc_frame_release(pc_ctx* c, int32_t frame_id) {
    FrameSlot* f = find_slot(c, frame_id);
    if (!f) return fail(c, PC_ERR_NOT_FOUND, "frame " + std::to_string(frame_id) + " is not resident");
    f->used = false;
    return PC_OK;
}

int pc_frame_num_levels(pc_ctx* c, int32_t frame_id, int* levels_out) {
    FrameSlot* f = find_slot(c, frame_id);
    if (!f) return fail(c, PC_ERR_NOT_FOUND, "frame " + std::to_string(frame_id) + " is not resident");
    *levels_out = f->levels;
    return PC_OK;
}

int pc_frame_read_level(pc_ctx* c, int32_t frame_id, int level, uint8_t* out, size_t cap, int* w_out, int* h_out) {
    PC_CUDA(c, cudaSetDevice(c->device));
    FrameSlot* f = find_slot(c, frame_id);
    if (!f) return fail(c, PC_ERR_NOT_FOUND, "frame " + std::to_string(frame_id) + " is not resident");
    PC_CHECK(c, level >= 0 && level < f->levels, "no such pyramid level");
    const Image8& im = f->level[level];
    if (w_out) *w_out = im.w;
    if (h_out) *h_out = im.h;
    if (!out) return PC_OK;
    if (cap < (size_t)im.w * im.h) return fail(c, PC_ERR_CAPACITY, "output buffer too small for level");
    PC_CUDA(c, cudaMemcpy2DAsync(out, im.w, im.data, im.pitch, im.w, im.h, cudaMemcpyDeviceToHost, c->compute));
    PC_CUDA(c, cudaStreamSynchronize(c->compute));
    return PC_OK;
}

// ---- detector -------------------------------------------------------------------------
static int fetch_kps_count(pc_ctx* c, FrameSlot* f) {
    int host[3];
    PC_CUDA(c, cudaMemcpyAsync(host, f->n_kps, sizeof(int) * 3, cudaMemcpyDeviceToHost, c->compute));
    PC_CUDA(c, cudaStreamSynchronize(c->compute));
    f->n_kps_host = host[0];
    return PC_OK;
}

static int check_detector_result(pc_ctx* c, const pc_gftt_opts* go, int n_kps, int n_accepted, int remaining,
                                 int cand_count) {
    if (remaining != 0)
        return fail(c, PC_ERR_STATE, "min-distance suppression did not reach its fixed point in " +
                                         std::to_string(kMaxGreedyRounds) + " rounds");
    if (cand_count > c->cand_cap) return fail(c, PC_ERR_CAPACITY, "corner candidate buffer overflow");
    const int want = go->max_corners > 0 ? std::min(go->max_corners, n_accepted) : n_accepted;
    if (want > c->lim.max_features)
        return fail(c, PC_ERR_CAPACITY, "detector produced " + std::to_string(want) +
                                            " corners, above the context's max_features " +
                                            std::to_string(c->lim.max_features));
    (void)n_kps;
    return PC_OK;
}

int pc_detect(pc_ctx* c, int32_t frame_id, const pc_gftt_opts* go, float* kps_out, int cap, int* n_out) {
    PC_CUDA(c, cudaSetDevice(c->device));
    int rc = validate_gftt_opts(c, go);
    if (rc) return rc;
    FrameSlot* f = find_slot(c, frame_id);
    if (!f) return fail(c, PC_ERR_NOT_FOUND, "frame " + std::to_string(frame_id) + " is not resident");
    rc = run_detector(c, f, go, c->compute);
    if (rc) return rc;
    int host[3], ncand = 0;
    PC_CUDA(c, cudaMemcpyAsync(host, f->n_kps, sizeof(int) * 3, cudaMemcpyDeviceToHost, c->compute));
    PC_CUDA(c, cudaMemcpyAsync(&ncand, c->cand_count, sizeof(int), cudaMemcpyDeviceToHost, c->compute));
    PC_CUDA(c, cudaStreamSynchronize(c->compute));
    rc = check_detector_result(c, go, host[0], host[1], host[2], ncand);
    if (rc) return rc;
    f->n_kps_host = host[0];
    if (n_out) *n_out = host[0];
    if (kps_out) {
        if (cap < host[0]) return fail(c, PC_ERR_CAPACITY, "keypoint output buffer too small");
        PC_CUDA(c, cudaMemcpy(kps_out, f->kps, sizeof(float) * 2 * host[0], cudaMemcpyDeviceToHost));
    }
    return PC_OK;
}

int pc_min_eig_map(pc_ctx* c, int32_t frame_id, float* eig_out, size_t cap_floats) {
    PC_CUDA(c, cudaSetDevice(c->device));
    FrameSlot* f = find_slot(c, frame_id);
    if (!f) return fail(c, PC_ERR_NOT_FOUND, "frame " + std::to_string(frame_id) + " is not resident");
    if (cap_floats < (size_t)f->w * f->h) return fail(c, PC_ERR_CAPACITY, "eig output buffer too small");
    DetectGrid g{1, 1, f->w, f->h};
    span_begin(c, KF_MIN_EIG, c->compute);
    launch_min_eig(f->level[0], c->eig, c->eig_pitch, g, c->cell_max, c->det_zero, c->det_zero_ints, nullptr, c->compute);
    span_end(c, c->compute);
    int rc = check_launch(c, "min_eig", 2);
    if (rc) return rc;
    PC_CUDA(c, cudaMemcpy2DAsync(eig_out, sizeof(float) * f->w, c->eig, sizeof(float) * c->eig_pitch,
                                 sizeof(float) * f->w, f->h, cudaMemcpyDeviceToHost, c->compute));
    PC_CUDA(c, cudaStreamSynchronize(c->compute));
    return PC_OK;
}

int pc_set_keypoints(pc_ctx* c, int32_t frame_id, const float* kps, int n) {
    PC_CUDA(c, cudaSetDevice(c->device));
    FrameSlot* f = find_slot(c, frame_id);
    if (!f) return fail(c, PC_ERR_NOT_FOUND, "frame " + std::to_string(frame_id) + " is not resident");
    PC_CHECK(c, n >= 0 && (n == 0 || kps != nullptr), "bad keypoint array");
    if (n > c->lim.max_features) return fail(c, PC_ERR_CAPACITY, "more keypoints than the context's max_features");
    int hdr[4] = {n, n, 0, 0};
    if (n) PC_CUDA(c, cudaMemcpyAsync(f->kps, kps, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, c->compute));
    PC_CUDA(c, cudaMemcpyAsync(f->n_kps, hdr, sizeof(hdr), cudaMemcpyHostToDevice, c->compute));
    PC_CUDA(c, cudaStreamSynchronize(c->compute));
    f->has_kps = true;
    f->n_kps_host = n;
    return PC_OK;
}

// ---- LK ---------------------------------------------------------------------------------
static int lk_sync_common(pc_ctx* c, int32_t from, int32_t to, const pc_flow_opts* fo, FrameSlot** fa_out) {
    PC_CUDA(c, cudaSetDevice(c->device));
    int rc = validate_flow_opts(c, fo);
    if (rc) return rc;
    FrameSlot* a = find_slot(c, from);
    FrameSlot* b = find_slot(c, to);
    if (!a) return fail(c, PC_ERR_NOT_FOUND, "frame " + std::to_string(from) + " is not resident");
    if (!b) return fail(c, PC_ERR_NOT_FOUND, "frame " + std::to_string(to) + " is not resident");
    if (!a->has_kps) return fail(c, PC_ERR_STATE, "frame " + std::to_string(from) + " has no keypoints");
    PC_CHECK(c, a->w == b->w && a->h == b->h, "frame sizes differ");
    if (a->n_kps_host < 0) {
        rc = fetch_kps_count(c, a);
        if (rc) return rc;
    }
    LKBatch batch{};
    batch.num_pairs = 1;
    batch.cap = std::max(1, a->n_kps_host);
    fill_pair(c, batch.pair[0], *a, *b, 0, c->sync_out);
    const LKParams p = make_lk_params(fo);
    if (a->n_kps_host > 0) {
        span_begin(c, KF_LK, c->compute);
        launch_lk(batch, p, c->compute);
        span_end(c, c->compute);
    }
    span_begin(c, KF_COMPACT, c->compute);
    launch_lk_compact(batch, c->compute);
    span_end(c, c->compute);
    rc = check_launch(c, "lk", a->n_kps_host > 0 ? 2 : 1);
    if (rc) return rc;
    *fa_out = a;
    return PC_OK;
}

int pc_lk_pair(pc_ctx* c, int32_t from, int32_t to, const pc_flow_opts* fo, uint32_t* src_idx_out, float* tgt_out,
               float* err_out, int cap, int* n_out) {
    FrameSlot* a = nullptr;
    int rc = lk_sync_common(c, from, to, fo, &a);
    if (rc) return rc;
    int n = 0;
    PC_CUDA(c, cudaMemcpyAsync(&n, c->sync_out.count, sizeof(int), cudaMemcpyDeviceToHost, c->compute));
    PC_CUDA(c, cudaStreamSynchronize(c->compute));
    if (n_out) *n_out = n;
    if (n > cap) return fail(c, PC_ERR_CAPACITY, "flow output buffers too small");
    if (src_idx_out) PC_CUDA(c, cudaMemcpy(src_idx_out, c->sync_out.idx, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    if (tgt_out) PC_CUDA(c, cudaMemcpy(tgt_out, c->sync_out.tgt, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost));
    if (err_out) PC_CUDA(c, cudaMemcpy(err_out, c->sync_out.err, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return PC_OK;
}

int pc_lk_raw(pc_ctx* c, int32_t from, int32_t to, const pc_flow_opts* fo, float* next_out, uint8_t* status_out,
              float* err_out, int cap, int* n_out) {
    FrameSlot* a = nullptr;
    int rc = lk_sync_common(c, from, to, fo, &a);
    if (rc) return rc;
    PC_CUDA(c, cudaStreamSynchronize(c->compute));
    const int n = a->n_kps_host;
    if (n_out) *n_out = n;
    if (n > cap) return fail(c, PC_ERR_CAPACITY, "flow output buffers too small");
    if (next_out) PC_CUDA(c, cudaMemcpy(next_out, c->lk_next, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost));
    if (status_out) PC_CUDA(c, cudaMemcpy(status_out, c->lk_status, (size_t)n, cudaMemcpyDeviceToHost));
    if (err_out) PC_CUDA(c, cudaMemcpy(err_out, c->lk_err, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return PC_OK;
}

// ---- streaming analyzer ----------------------------------------------------------------
int pc_analyze_begin(pc_ctx* c, const pc_video_info* vi, const pc_gftt_opts* go, const pc_flow_opts* fo) {
    PC_CUDA(c, cudaSetDevice(c->device));
    PC_CHECK(c, vi != nullptr, "video info is required");
    if (c->analyzing) return fail(c, PC_ERR_STATE, "an analyze pass is already open");
    int rc = validate_gftt_opts(c, go);
    if (rc) return rc;
    rc = validate_flow_opts(c, fo);
    if (rc) return rc;
    PC_CHECK(c, (int)vi->width <= c->lim.max_width && (int)vi->height <= c->lim.max_height,
             "video larger than the context limits");
    PC_CHECK(c, vi->width >= 16 && vi->height >= 16, "frames smaller than 16x16 are not supported");
    c->vinfo = *vi;
    c->gopts = *go;
    c->fopts = *fo;
    const int cap = c->lim.max_features;
    if (c->stages.empty()) {
        c->stages.resize(c->lim.pipeline_depth);
        for (auto& st : c->stages) {
            rc = alloc_stage_rows(c, st, cap);
            if (rc) return rc;
            PC_CUDA(c, cudaMallocHost(&st.kps_host, sizeof(float) * 2 * cap));
            PC_CUDA(c, cudaMallocHost(&st.counts_host, sizeof(int) * 4));
            PC_CUDA(c, cudaEventCreateWithFlags(&st.uploaded, cudaEventDisableTiming));
            PC_CUDA(c, cudaEventCreateWithFlags(&st.gray_done, cudaEventDisableTiming));
            PC_CUDA(c, cudaEventCreateWithFlags(&st.detected, cudaEventDisableTiming));
            PC_CUDA(c, cudaEventCreateWithFlags(&st.templates, cudaEventDisableTiming));
            PC_CUDA(c, cudaEventCreateWithFlags(&st.computed, cudaEventDisableTiming));
            PC_CUDA(c, cudaEventCreateWithFlags(&st.downloaded, cudaEventDisableTiming));
        }
    }
    if (c->side) PC_CUDA(c, cudaStreamSynchronize(c->side));
    if (c->side2) {
        PC_CUDA(c, cudaStreamSynchronize(c->side2));
        if (!c->lk_next2) {                          // the second LK stream's dense scratch
            const size_t cap = (size_t)c->lim.max_features;
            if (cudaMalloc(&c->lk_next2, sizeof(float) * 2 * cap * 8) != cudaSuccess ||
                cudaMalloc(&c->lk_status2, cap * 8) != cudaSuccess ||
                cudaMalloc(&c->lk_err2, sizeof(float) * cap * 8) != cudaSuccess) {
                cudaGetLastError();
                cudaFree(c->lk_next2); cudaFree(c->lk_status2); cudaFree(c->lk_err2);
                c->lk_next2 = nullptr; c->lk_status2 = nullptr; c->lk_err2 = nullptr;
            }
        }
    }
    for (int k = 0; k < c->n_det_extra; k++) {
        PC_CUDA(c, cudaStreamSynchronize(c->det_stream[k]));
        if (!c->det_set[k]) {                        // the extra detector stream's scratch set
            c->det_set[k] = new DetScratch();
            int rc2 = alloc_det_scratch(c->lim.max_width, c->lim.max_height, *c->det_set[k]);
            if (rc2) {                               // not enough memory: fewer detector streams
                cudaGetLastError();
                free_det_scratch(*c->det_set[k]);
                delete c->det_set[k];
                c->det_set[k] = nullptr;
                c->n_det_extra = k;
                break;
            }
        }
    }
    for (auto& s : c->slots) { s.used = false; s.has_tmpl = false; s.has_order = false; }
    // template cache of the 10x10 LK kernel: (max_level + 1) x max_features x 640 B per slot, kept as
    // long as it stays under 4 GB for the whole ring (the kernel computes templates itself otherwise)
    {
        const int tl = std::min(std::max(fo->max_level, 0) + 1, kMaxLevels);
        const size_t per_slot = (size_t)tl * c->lim.max_features *
                                (c->lk_queue_mode ? kLkQueueTemplateBytesPerPoint : kLkTemplateBytesPerPoint);
        const bool want = fo->window_size == 10 && per_slot * c->slots.size() <= ((size_t)4 << 30) && !getenv("PC_NO_LK_TEMPLATES");
        for (auto& s : c->slots) {
            if (want && s.tmpl && s.tmpl_levels >= tl) continue;
            cudaFree(s.tmpl); cudaFree(s.tmpl_sums);
            s.tmpl = nullptr; s.tmpl_sums = nullptr; s.tmpl_levels = 0;
            if (!want) continue;
            if (cudaMalloc(&s.tmpl, per_slot) != cudaSuccess ||
                cudaMalloc(&s.tmpl_sums, (size_t)tl * c->lim.max_features * 4 * sizeof(float)) != cudaSuccess) {
                cudaGetLastError();
                cudaFree(s.tmpl); cudaFree(s.tmpl_sums);
                s.tmpl = nullptr; s.tmpl_sums = nullptr;
                continue;
            }
            s.tmpl_levels = tl;
        }
    }
    c->download_hint = false;
    c->inflight.clear();
    c->next_stage = 0;
    c->any_pushed = false;
    c->halo_frames = 0;
    c->pushed_count = 0;
    c->preset_kps.clear();
    for (auto& st : c->stages) { st.busy = false; st.gray_pending = false; st.track_state = 0; }
    if (c->track) c->track->on = false;
    c->analyzing = true;
    return PC_OK;
}

int pc_analyze_preset_keypoints(pc_ctx* c, int32_t frame_id, const float* kps, int n) {
    if (!c->analyzing) return fail(c, PC_ERR_STATE, "no analyze pass is open");
    PC_CHECK(c, n >= 0 && (n == 0 || kps != nullptr), "bad keypoint array");
    if (n > c->lim.max_features) return fail(c, PC_ERR_CAPACITY, "more keypoints than the context's max_features");
    c->preset_kps[frame_id].assign(kps, kps + 2 * (size_t)n);
    return PC_OK;
}

int pc_analyze_pending(pc_ctx* c) { return (int)c->inflight.size(); }

int pc_analyze_set_halo(pc_ctx* c, int halo_frames) {
    if (!c->analyzing) return fail(c, PC_ERR_STATE, "no analyze pass is open");
    PC_CHECK(c, halo_frames >= 0 && !c->any_pushed, "set the halo before the first push");
    c->halo_frames = halo_frames;
    return PC_OK;
}

int pc_mark(pc_ctx* c, int slot) {
    PC_CUDA(c, cudaSetDevice(c->device));
    PC_CHECK(c, slot >= 0 && slot < 8, "mark slot must be in [0,8)");
    if (!c->join_a) {
        PC_CUDA(c, cudaEventCreateWithFlags(&c->join_a, cudaEventDisableTiming));
        PC_CUDA(c, cudaEventCreateWithFlags(&c->join_b, cudaEventDisableTiming));
        PC_CUDA(c, cudaEventCreateWithFlags(&c->join_c, cudaEventDisableTiming));
    }
    if (!c->marks[slot]) PC_CUDA(c, cudaEventCreate(&c->marks[slot]));
    PC_CUDA(c, cudaEventRecord(c->join_a, c->h2d));
    PC_CUDA(c, cudaEventRecord(c->join_b, c->d2h));
    PC_CUDA(c, cudaStreamWaitEvent(c->compute, c->join_a, 0));
    PC_CUDA(c, cudaStreamWaitEvent(c->compute, c->join_b, 0));
    if (c->side) {
        PC_CUDA(c, cudaEventRecord(c->join_c, c->side));
        PC_CUDA(c, cudaStreamWaitEvent(c->compute, c->join_c, 0));
    }
    for (int k = 0; k < c->n_det_extra; k++) {
        PC_CUDA(c, cudaEventRecord(c->det_join[k], c->det_stream[k]));
        PC_CUDA(c, cudaStreamWaitEvent(c->compute, c->det_join[k], 0));
    }
    if (c->side2) {
        PC_CUDA(c, cudaEventRecord(c->join_e, c->side2));
        PC_CUDA(c, cudaStreamWaitEvent(c->compute, c->join_e, 0));
    }
    if (c->track && c->track->stream) {
        PC_CUDA(c, cudaEventRecord(c->track->join, c->track->stream));
        PC_CUDA(c, cudaStreamWaitEvent(c->compute, c->track->join, 0));
    }
    PC_CUDA(c, cudaEventRecord(c->marks[slot], c->compute));
    // work queued after the mark starts after it on every detector stream (a timed region opens with a mark)
    for (int k = 0; k < c->n_det_extra; k++) PC_CUDA(c, cudaStreamWaitEvent(c->det_stream[k], c->marks[slot], 0));
    return PC_OK;
}

int pc_elapsed_ms(pc_ctx* c, int a, int b, float* ms_out) {
    PC_CHECK(c, a >= 0 && a < 8 && b >= 0 && b < 8 && c->marks[a] && c->marks[b] && ms_out, "bad mark slots");
    PC_CUDA(c, cudaEventSynchronize(c->marks[a]));
    PC_CUDA(c, cudaEventSynchronize(c->marks[b]));
    PC_CUDA(c, cudaEventElapsedTime(ms_out, c->marks[a], c->marks[b]));
    return PC_OK;
}

// Largest row slab that travels whole (one copy) instead of as size-exact per-array copies.
static const size_t kSlabCopyMax = 8u << 20;

// Queues the LK batch of a pushed frame -- every pair whose later frame is this one: (j-d -> j) and
// (j -> j-d), d in {1,2,4,8} -- its compaction, the fused track step and the result downloads.
static int enqueue_lk(pc_ctx* c, Stage& st, FrameSlot* f) {
    const int32_t frame_id = st.frame_id;
    const int32_t first = c->vinfo.first_frame;
    cudaStream_t lks = c->compute;
    // batches alternate between the two LK streams (each has its dense scratch set)
    const bool second_lk = c->side2 && c->lk_next2 && (frame_id & 1);
    if (c->side) {
        lks = second_lk ? c->side2 : c->side;
        // this frame's pyramid and keypoints
        PC_CUDA(c, cudaEventRecord(st.detected, st.det_stream ? st.det_stream : c->compute));
        PC_CUDA(c, cudaStreamWaitEvent(lks, st.detected, 0));
        // ... and those of the earlier frames: frames j-2, j-4, j-8 by the order of this LK stream (their batches ran
        // on it and waited for their own events); frame j-1's pyramid, keypoints and templates were produced for / by
        // the batch on the other LK stream
        if (c->side2 && c->lk_next2) {
            for (auto& other : c->stages)
                if (&other != &st && other.busy && other.frame_id == frame_id - 1) {
                    PC_CUDA(c, cudaStreamWaitEvent(lks, other.detected, 0));
                    if (other.templates_recorded) PC_CUDA(c, cudaStreamWaitEvent(lks, other.templates, 0));
                }
        }
    }
    st.templates_recorded = false;
    const LKParams lkp = make_lk_params(&c->fopts);
    int rc;
    // source templates of this frame's keypoints (once per frame; its eight pairs load them)
    if (f->tmpl && lkp.win == 10 && lkp.max_level + 1 <= f->tmpl_levels) {
        span_begin(c, KF_LK_TMPL, lks);
        // PC_LK_SPATIAL=0: walk the keypoints in index (strength) order
        static const bool spatial = getenv("PC_LK_SPATIAL") == nullptr || atoi(getenv("PC_LK_SPATIAL")) != 0;
        f->has_order = false;
        if (spatial && !c->lk_queue_mode && f->order) {
            launch_spatial_order(f->kps, f->n_kps, c->lim.max_features, f->w, f->h, f->order, lks);
            f->has_order = true;
        }
        if (c->lk_queue_mode) launch_lk10q_templates(view_of(*f), f->kps, f->n_kps, c->lim.max_features, lkp, f->tmpl, lks);
        else launch_lk10_templates(view_of(*f), f->kps, f->n_kps, c->lim.max_features, lkp, f->tmpl, f->tmpl_sums, lks,
                                   f->has_order ? f->order : nullptr);
        f->tmpl_queue_layout = c->lk_queue_mode;
        span_end(c, lks);
        rc = check_launch(c, "lk templates", f->has_order ? 2 : 1);
        if (rc) return rc;
        f->has_tmpl = true;
        if (c->side) {
            PC_CUDA(c, cudaEventRecord(st.templates, lks));
            st.templates_recorded = true;
        }
    }
    LKBatch batch{};
    batch.cap = c->lim.max_features;
    if (c->gopts.max_corners > 0) batch.cap = std::min(batch.cap, c->gopts.max_corners);
    static const int kSkips[4] = {1, 2, 4, 8};   // |image_skips|, opticalflow.cc:76-77
    int np = 0;
    for (int k = 0; k < 4 && !st.is_halo; k++) {
        const int32_t other = frame_id - kSkips[k];
        if (other < first) continue;
        FrameSlot* o = find_slot(c, other);
        if (!o) return fail(c, PC_ERR_STATE, "frame " + std::to_string(other) + " fell out of the ring");
        st.from[np] = other; st.to[np] = frame_id;
        fill_pair(c, batch.pair[np], *o, *f, np, st.dev[np], true, second_lk);
        np++;
        st.from[np] = frame_id; st.to[np] = other;
        fill_pair(c, batch.pair[np], *f, *o, np, st.dev[np], true, second_lk);
        np++;
    }
    // presets may exceed max_corners
    for (int k = 0; k < np; k++) {
        const FrameSlot* src = find_slot(c, st.from[k]);
        if (src->n_kps_host > batch.cap) batch.cap = src->n_kps_host;
    }
    batch.num_pairs = np;
    st.num_pairs = np;
    // work-queue kernel: every pair of the batch brings queue-layout templates (indexed with stride max_features)
    if (c->lk_queue_mode && lkp.win == 10 && np > 0 && batch.cap <= c->lim.max_features) {
        bool all = true;
        for (int k = 0; k < np; k++) {
            const FrameSlot* src = find_slot(c, st.from[k]);
            all = all && src->has_tmpl && src->tmpl_queue_layout && batch.pair[k].tmpl.words != nullptr;
        }
        if (all) { batch.queue = c->lk_queue + (second_lk ? 1 : 0); batch.queue_budget = c->lk_queue_budget; }
    }
    if (np > 0) {
        const LKParams& p = lkp;
        span_begin(c, KF_LK, lks);
        launch_lk(batch, p, lks);
        span_end(c, lks);
        span_begin(c, KF_COMPACT, lks);
        launch_lk_compact(batch, lks);
        span_end(c, lks);
        rc = check_launch(c, "lk batch", batch.queue ? 3 : 2);
        if (rc) return rc;
    }
    PC_CUDA(c, cudaEventRecord(st.computed, lks));
    rc = track_chain_enqueue(c, st, frame_id, st.is_halo, batch.cap);     // fused Track (no-op unless enabled)
    if (rc) return rc;
    // results -> pinned host, on the download stream
    PC_CUDA(c, cudaStreamWaitEvent(c->d2h, st.computed, 0));
    PC_CUDA(c, cudaMemcpyAsync(st.counts_host, f->n_kps, sizeof(int) * 4, cudaMemcpyDeviceToHost, c->d2h));
    st.rows_downloaded = false;
    if (c->download_hint && st.rows_bytes <= kSlabCopyMax && batch.cap <= c->lim.max_features) {
        // the caller has been taking rows: send the whole slab (counts + rows) and the keypoints now,
        // so the pop of this frame is a single wait
        PC_CUDA(c, cudaMemcpyAsync(st.rows_host, st.rows_dev, np > 0 ? st.rows_bytes : 128, cudaMemcpyDeviceToHost, c->d2h));
        PC_CUDA(c, cudaMemcpyAsync(st.kps_host, f->kps, sizeof(float) * 2 * batch.cap, cudaMemcpyDeviceToHost, c->d2h));
        st.rows_downloaded = true;
    } else if (np > 0) {
        PC_CUDA(c, cudaMemcpyAsync(st.rows_host, st.rows_dev, sizeof(int) * 8, cudaMemcpyDeviceToHost, c->d2h));
    }
    PC_CUDA(c, cudaEventRecord(st.downloaded, c->d2h));
    return PC_OK;
}

int pc_analyze_push_frame(pc_ctx* c, int32_t frame_id, const uint8_t* rgb, size_t stride, int mem_kind) {
    PC_CUDA(c, cudaSetDevice(c->device));
    if (!c->analyzing) return fail(c, PC_ERR_STATE, "no analyze pass is open");
    PC_CHECK(c, rgb != nullptr, "frame pointer is NULL");
    const int w = (int)c->vinfo.width, h = (int)c->vinfo.height;
    const int32_t first = c->vinfo.first_frame, last = first + (int32_t)c->vinfo.num_frames;  // opticalflow.cc:220-221
    PC_CHECK(c, frame_id >= first && frame_id < last, "frame id outside the video range");
    PC_CHECK(c, !c->any_pushed || frame_id == c->last_pushed + 1, "frames must be pushed in ascending order");
    PC_CHECK(c, stride >= (size_t)w * 3, "stride smaller than a row");
    if ((int)c->inflight.size() >= (int)c->stages.size())
        return fail(c, PC_ERR_STATE, "pipeline full: pop a result before pushing another frame");
    const int si = c->next_stage;
    Stage& st = c->stages[si];
    st.frame_id = frame_id;
    st.num_pairs = 0;

    FrameSlot* f = acquire_slot(c, frame_id);
    // detector stream of this frame: frames take the streams in turn (each stream owns a scratch set)
    const int det_turn = c->pushed_count % (1 + c->n_det_extra);     // 0 = `compute` and the primary set
    const DetScratch* det_scratch = det_turn > 0 ? c->det_set[det_turn - 1] : nullptr;
    cudaStream_t ds = (det_turn > 0 && det_scratch) ? c->det_stream[det_turn - 1] : c->compute;
    st.det_stream = ds;
    const uint8_t* dev = rgb;
    if (mem_kind != PC_MEM_DEVICE) {
        const size_t bytes = stride * (size_t)h;
        if (st.rgb_bytes < bytes) {
            cudaFree(st.rgb_dev);
            st.rgb_dev = nullptr;
            st.rgb_bytes = 0;
            PC_CUDA(c, cudaMalloc(&st.rgb_dev, bytes));
            st.rgb_bytes = bytes;
        }
        // the staging buffer's previous contents must have been consumed by its gray kernel
        if (st.gray_pending) PC_CUDA(c, cudaStreamWaitEvent(c->h2d, st.gray_done, 0));
        // the frame travels as 4 slices on 4 streams, i.e. on several copy engines at once (PC_H2D_SPLIT=n, 1..4, overrides):
        // one 24.9 MB copy reaches 36 GB/s on a PCIe 5 x16 box, slices on two or more streams 52 GB/s
        // (profiles/r2_j_h2d_probe.json); e2e 14 273 -> 15 161 pairs/s (profiles/r2_k_streams_ab.json)
        static const int split = [] { const char* e = getenv("PC_H2D_SPLIT"); return e ? std::min(std::max(atoi(e), 1), 4) : 4; }();
        if (split > 1 && mem_kind == PC_MEM_HOST_PINNED) {
            if (!c->h2d_extra[0])
                for (int k = 0; k < 3; k++) {
                    PC_CUDA(c, cudaStreamCreateWithFlags(&c->h2d_extra[k], cudaStreamNonBlocking));
                    PC_CUDA(c, cudaEventCreateWithFlags(&c->h2d_extra_ev[k], cudaEventDisableTiming));
                }
            PC_CUDA(c, cudaEventRecord(c->h2d_fork, c->h2d));                 // the slices wait for what the h2d stream waited for
            const size_t rows_per = ((size_t)h + split - 1) / split;
            for (int k = 0; k < split; k++) {
                const size_t r0 = k * rows_per, r1 = std::min((size_t)h, r0 + rows_per);
                if (r0 >= r1) break;
                cudaStream_t sk = k == 0 ? c->h2d : c->h2d_extra[k - 1];
                if (k > 0) PC_CUDA(c, cudaStreamWaitEvent(sk, c->h2d_fork, 0));
                PC_CUDA(c, cudaMemcpyAsync(st.rgb_dev + r0 * stride, rgb + r0 * stride, (r1 - r0) * stride, cudaMemcpyHostToDevice, sk));
                if (k > 0) {
                    PC_CUDA(c, cudaEventRecord(c->h2d_extra_ev[k - 1], sk));
                    PC_CUDA(c, cudaStreamWaitEvent(c->h2d, c->h2d_extra_ev[k - 1], 0));
                }
            }
        } else {
            PC_CUDA(c, cudaMemcpyAsync(st.rgb_dev, rgb, bytes, cudaMemcpyHostToDevice, c->h2d));
        }
        PC_CUDA(c, cudaEventRecord(st.uploaded, c->h2d));
        PC_CUDA(c, cudaStreamWaitEvent(ds, st.uploaded, 0));
        dev = st.rgb_dev;
    }
    int rc = build_pyramid(c, f, dev, stride, 3, w, h, &c->fopts, ds);
    if (rc) return rc;
    if (mem_kind != PC_MEM_DEVICE) {
        PC_CUDA(c, cudaEventRecord(st.gray_done, ds));
        st.gray_pending = true;
    }
    auto preset = c->preset_kps.find(frame_id);
    if (preset != c->preset_kps.end()) {   // ReadOrGenerateKeypoints: a stored row is used as it is, even when empty
        const int n = (int)(preset->second.size() / 2);
        int hdr[4] = {n, n, 0, 0};
        if (n) memcpy(st.kps_host, preset->second.data(), sizeof(float) * 2 * n);
        memcpy(st.counts_host, hdr, sizeof(hdr));
        if (n) PC_CUDA(c, cudaMemcpyAsync(f->kps, st.kps_host, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, ds));
        PC_CUDA(c, cudaMemcpyAsync(f->n_kps, st.counts_host, sizeof(hdr), cudaMemcpyHostToDevice, ds));
        f->has_kps = true;
        f->n_kps_host = n;
        c->preset_kps.erase(preset);
    } else {
        rc = run_detector(c, f, &c->gopts, ds, ds == c->compute ? nullptr : det_scratch);
        if (rc) return rc;
    }
    st.is_halo = c->pushed_count < c->halo_frames;
    c->pushed_count++;
    st.busy = true;
    c->inflight.push_back(si);
    c->next_stage = (si + 1) % (int)c->stages.size();
    c->last_pushed = frame_id;
    c->any_pushed = true;
    return enqueue_lk(c, st, f);
}

int pc_analyze_pop(pc_ctx* c, pc_frame_result* out, int download) {
    PC_CUDA(c, cudaSetDevice(c->device));
    if (!c->analyzing) return fail(c, PC_ERR_STATE, "no analyze pass is open");
    if (c->inflight.empty()) return fail(c, PC_ERR_NOT_FOUND, "no frame is pending");
    PC_CHECK(c, out != nullptr, "result pointer is NULL");
    const int si = c->inflight.front();
    Stage& st = c->stages[si];
    PC_CUDA(c, cudaEventSynchronize(st.downloaded));
    memset(out, 0, sizeof(*out));
    out->frame_id = st.frame_id;
    const int n_kps = st.counts_host[0];
    int rc = check_detector_result(c, &c->gopts, n_kps, st.counts_host[1], st.counts_host[2], st.counts_host[3]);
    if (rc) { c->inflight.pop_front(); st.busy = false; return rc; }
    out->num_keypoints = n_kps;
    out->num_pairs = st.num_pairs;
    FrameSlot* f = find_slot(c, st.frame_id);
    if (f) f->n_kps_host = n_kps;
    c->download_hint = download != 0;
    if (download && !st.rows_downloaded) {
        // second phase (the counts are known now), on its own stream: c->d2h already holds the count
        // downloads of the frames pushed after this one, which wait for their LK batches
        cudaStream_t s = c->d2h_rows;
        if (f && n_kps > 0)
            PC_CUDA(c, cudaMemcpyAsync(st.kps_host, f->kps, sizeof(float) * 2 * n_kps, cudaMemcpyDeviceToHost, s));
        if (st.num_pairs > 0 && st.rows_bytes <= kSlabCopyMax) {
            PC_CUDA(c, cudaMemcpyAsync(st.rows_host + 128, st.rows_dev + 128, st.rows_bytes - 128, cudaMemcpyDeviceToHost, s));
        } else {
            for (int k = 0; k < st.num_pairs; k++) {
                const int n = *st.host[k].count;
                if (n <= 0) continue;
                PC_CUDA(c, cudaMemcpyAsync(st.host[k].idx, st.dev[k].idx, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, s));
                PC_CUDA(c, cudaMemcpyAsync(st.host[k].tgt, st.dev[k].tgt, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, s));
                PC_CUDA(c, cudaMemcpyAsync(st.host[k].err, st.dev[k].err, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
            }
        }
        PC_CUDA(c, cudaStreamSynchronize(s));
    }
    if (download) out->keypoints = st.kps_host;
    for (int k = 0; k < st.num_pairs; k++) {
        pc_pair_rows& r = out->pairs[k];
        r.image_id_from = st.from[k];
        r.image_id_to = st.to[k];
        r.rows = *st.host[k].count;
        if (download) {
            r.src_kps_indices = st.host[k].idx;
            r.tgt_kps = st.host[k].tgt;
            r.flow_errors = st.host[k].err;
        }
    }
    rc = track_chain_collect(c, st, out);
    c->inflight.pop_front();
    st.busy = false;
    return rc;
}

int pc_analyze_end(pc_ctx* c) {
    if (!c->analyzing) return fail(c, PC_ERR_STATE, "no analyze pass is open");
    int rc = pc_synchronize(c);
    c->inflight.clear();
    for (auto& st : c->stages) st.busy = false;
    c->analyzing = false;
    return rc;
}

// ---- synth / memory helpers --------------------------------------------------------------
int pc_synth_set_texture(pc_ctx* c, const uint8_t* tex, int w, int h) {
    PC_CUDA(c, cudaSetDevice(c->device));
    PC_CHECK(c, tex && w > 0 && h > 0, "bad texture");
    cudaFree(c->tex);
    c->tex = nullptr;
    c->tex_pitch = (w + 127) / 128 * 128;
    PC_CUDA(c, cudaMalloc(&c->tex, (size_t)c->tex_pitch * h));
    PC_CUDA(c, cudaMemcpy2D(c->tex, c->tex_pitch, tex, w, w, h, cudaMemcpyHostToDevice));
    c->tex_w = w;
    c->tex_h = h;
    return PC_OK;
}

static bool invert3x3(const double m[9], double o[9]) {
    const double a = m[0], b = m[1], cc = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + cc * (d * h - e * g);
    if (det == 0.0) return false;
    const double id = 1.0 / det;
    o[0] = (e * i - f * h) * id; o[1] = (cc * h - b * i) * id; o[2] = (b * f - cc * e) * id;
    o[3] = (f * g - d * i) * id; o[4] = (a * i - cc * g) * id; o[5] = (cc * d - a * f) * id;
    o[6] = (d * h - e * g) * id; o[7] = (b * g - a * h) * id; o[8] = (a * e - b * d) * id;
    return true;
}

int pc_synth_render_rgb8(pc_ctx* c, const double H[9], uint8_t* rgb_dev, size_t stride) {
    PC_CUDA(c, cudaSetDevice(c->device));
    if (!c->tex) return fail(c, PC_ERR_STATE, "no texture set");
    double Hi[9];
    if (!invert3x3(H, Hi)) return fail(c, PC_ERR_INVALID, "singular homography");
    launch_synth_warp(c->tex, c->tex_w, c->tex_h, c->tex_pitch, Hi, rgb_dev, stride, c->compute);
    return check_launch(c, "synth", 1);
}

int pc_device_alloc(pc_ctx* c, size_t bytes, void** out) {
    PC_CUDA(c, cudaSetDevice(c->device));
    PC_CUDA(c, cudaMalloc(out, bytes));
    return PC_OK;
}
int pc_device_free(pc_ctx* c, void* p) {
    PC_CUDA(c, cudaSetDevice(c->device));
    PC_CUDA(c, cudaFree(p));
    return PC_OK;
}
// NUMA node the device hangs off (/sys/bus/pci/devices/<bus id>/numa_node), -1 when unknown.
static int device_numa_node(int device) {
    char id[32] = {0};
    if (cudaDeviceGetPCIBusId(id, sizeof(id), device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char* p = id; *p; p++) *p = (char)tolower(*p);
    const std::string path = std::string("/sys/bus/pci/devices/") + id + "/numa_node";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

int pc_host_alloc_pinned(pc_ctx* c, size_t bytes, void** out) {
    PC_CUDA(c, cudaSetDevice(c->device));
    // The pages of an upload ring should live on the NUMA node the GPU hangs off: with eight ranks on one box every
    // upload otherwise crosses the socket interconnect from whichever node the processes happened to start on
    // (profiles/r2_n_h2d_numa_probe.json).  Preferred, not bound: the allocation still succeeds when that node is
    // full or outside the cpuset.  PC_PINNED_NUMA=0 leaves the process policy alone.
    static const bool numa = getenv("PC_PINNED_NUMA") == nullptr || atoi(getenv("PC_PINNED_NUMA")) != 0;
    const int node = numa ? device_numa_node(c->device) : -1;
    bool policy_set = false;
#ifdef __linux__
    int saved_mode = 0;                                   // the calling thread's own policy is put back afterwards
    unsigned long saved_mask[16] = {0};                   // up to 1024 nodes
    if (node >= 0 && node < 64 &&
        syscall(SYS_get_mempolicy, &saved_mode, saved_mask, sizeof(saved_mask) * 8ul, nullptr, 0ul) == 0) {
        unsigned long mask = 1ul << node;
        policy_set = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, &mask, 64ul) == 0;
    }
#endif
    // PC_PINNED_WC=1: write-combined page-locked memory (frames are written once by the host and only read by the
    // copy engine; write-combined pages are not snooped during the transfer)
    static const bool wc = getenv("PC_PINNED_WC") != nullptr && atoi(getenv("PC_PINNED_WC")) != 0;
    const cudaError_t e = cudaHostAlloc(out, bytes, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
#ifdef __linux__
    if (policy_set && syscall(SYS_set_mempolicy, saved_mode, saved_mask, sizeof(saved_mask) * 8ul) != 0)
        syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul);
#endif
    PC_CUDA(c, e);
    return PC_OK;
}
int pc_host_free_pinned(pc_ctx* c, void* p) {
    PC_CUDA(c, cudaFreeHost(p));
    return PC_OK;
}
int pc_memcpy_d2h(pc_ctx* c, void* dst, const void* src, size_t bytes) {
    PC_CUDA(c, cudaSetDevice(c->device));
    PC_CUDA(c, cudaStreamSynchronize(c->compute));
    PC_CUDA(c, cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return PC_OK;
}
int pc_memcpy_h2d(pc_ctx* c, void* dst, const void* src, size_t bytes) {
    PC_CUDA(c, cudaSetDevice(c->device));
    PC_CUDA(c, cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return PC_OK;
}

int pc_timing_enable(pc_ctx* c, int on) {
    c->timing = on != 0;
    return PC_OK;
}

int pc_timing_read(pc_ctx* c, pc_kernel_times* out, int reset) {
    int rc = pc_synchronize(c);
    if (rc) return rc;
    for (auto& sp : c->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.start, sp.stop) == cudaSuccess) {
            c->fam_ms[sp.family] += ms;
            c->fam_n[sp.family] += 1;
        }
        c->event_pool.push_back(sp.start);
        c->event_pool.push_back(sp.stop);
    }
    cudaGetLastError();
    c->spans.clear();
    if (out) {
        out->gray_pyr_ms = c->fam_ms[KF_GRAY_PYR]; out->gray_pyr_n = c->fam_n[KF_GRAY_PYR];
        out->min_eig_ms = c->fam_ms[KF_MIN_EIG]; out->min_eig_n = c->fam_n[KF_MIN_EIG];
        out->select_ms = c->fam_ms[KF_SELECT]; out->select_n = c->fam_n[KF_SELECT];
        out->lk_ms = c->fam_ms[KF_LK]; out->lk_n = c->fam_n[KF_LK];
        out->compact_ms = c->fam_ms[KF_COMPACT]; out->compact_n = c->fam_n[KF_COMPACT];
        out->raycast_ms = c->fam_ms[KF_RAYCAST]; out->raycast_n = c->fam_n[KF_RAYCAST];
        out->pnp_ms = c->fam_ms[KF_PNP]; out->pnp_n = c->fam_n[KF_PNP];
        out->ba_ms = c->fam_ms[KF_BA]; out->ba_n = c->fam_n[KF_BA];
        out->lk_tmpl_ms = c->fam_ms[KF_LK_TMPL]; out->lk_tmpl_n = c->fam_n[KF_LK_TMPL];
    }
    if (reset) {
        for (int i = 0; i < KF_COUNT; i++) { c->fam_ms[i] = 0; c->fam_n[i] = 0; }
    }
    return PC_OK;
}

}  // extern "C"
