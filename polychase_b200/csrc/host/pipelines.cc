// Host drivers of the three operations, calling the CUDA library only through the C ABI
// (include/polychase_b200.h).
#include "pipelines.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <exception>
#include <map>
#include <mutex>
#include <thread>

namespace pch {

namespace {

std::string Format(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
std::string Format(const char* fmt, ...) {
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return buf;
}

void Check(pc_ctx* ctx, int rc) {
    if (rc != PC_OK) ThrowPcError(ctx, rc);
}

// Write-behind for the analyze pass: finished frames are copied out of the page-locked result slab
// and written by this thread, one transaction per frame, while the caller already waits for the
// next frame.  The connection is used by this thread only between Start() and Finish().
class FlowWriter {
   public:
    struct Pair {
        int32_t from, to;
        std::vector<uint32_t> idx;
        std::vector<float> tgt, err;
    };
    struct Item {
        int32_t frame_id = 0;
        std::vector<float> kps;        // x, y
        std::vector<Pair> pairs;
    };
    explicit FlowWriter(Database& db) : db_(db) { thread_ = std::thread([this] { Run(); }); }
    ~FlowWriter() {
        try {
            Finish();
        } catch (...) {
        }
    }
    // Rethrows what a previous write threw.  Blocks while kMaxQueued frames wait (back-pressure).
    void Push(Item&& it) {
        std::unique_lock<std::mutex> lk(mtx_);
        space_.wait(lk, [&] { return queue_.size() < kMaxQueued || error_; });
        if (error_) std::rethrow_exception(error_);
        queue_.push_back(std::move(it));
        work_.notify_one();
    }
    // Drains the queue, stops the thread, rethrows a write error.
    void Finish() {
        {
            std::lock_guard<std::mutex> lk(mtx_);
            done_ = true;
        }
        work_.notify_one();
        if (thread_.joinable()) thread_.join();
        if (error_) {
            std::exception_ptr e = error_;
            error_ = nullptr;
            std::rethrow_exception(e);
        }
    }

   private:
    static constexpr size_t kMaxQueued = 8;
    void Run() {
        for (;;) {
            Item it;
            {
                std::unique_lock<std::mutex> lk(mtx_);
                work_.wait(lk, [&] { return !queue_.empty() || done_; });
                if (queue_.empty()) return;
                it = std::move(queue_.front());
                queue_.pop_front();
            }
            space_.notify_one();
            if (error_) continue;                          // after a failure: drop the rest
            try {
                db_.Begin();
                try {
                    if (!db_.KeypointsExist(it.frame_id))                          // ReadOrGenerateKeypoints :168-178
                        db_.WriteKeypoints(it.frame_id, it.kps.data(), it.kps.size() / 2);
                    for (const Pair& p : it.pairs) {
                        if (db_.ImagePairFlowExists(p.from, p.to)) continue;       // :286
                        // a pair row references the source frame's keypoints row (FOREIGN KEY): the source
                        // is either this frame or an earlier one, both already written
                        db_.WriteImagePairFlow(p.from, p.to, p.idx.data(), p.tgt.data(), p.err.data(), p.idx.size());
                    }
                } catch (...) {
                    db_.Commit();
                    throw;
                }
                db_.Commit();
            } catch (...) {
                std::lock_guard<std::mutex> lk(mtx_);
                error_ = std::current_exception();
                space_.notify_all();
            }
        }
    }
    Database& db_;
    std::thread thread_;
    std::mutex mtx_;
    std::condition_variable work_, space_;
    std::deque<Item> queue_;
    std::exception_ptr error_;
    bool done_ = false;
};

}  // namespace

// ---- Analyze -------------------------------------------------------------------------------
// GenerateOpticalFlowDatabase (opticalflow.cc:209-321).  The reference walks frame_id1 and, for
// each of the 8 offsets, re-reads / re-grays / re-pyramids frame_id2.  Here every frame is
// requested once, in ascending order, pushed through the streaming analyzer (gray + pyramid +
// detection once per frame, all pairs whose later frame just arrived in one LK launch) and the
// finished rows are written behind the GPU.  The database ends up with the same rows: one
// keypoints row per frame, one optical_flow row per directed pair inside [first, first+num).
void GenerateOpticalFlowDatabase(const VideoInfo& video_info, const FrameAccessorFunction& frame_accessor,
                                 const OpticalFlowProgressCallback& callback, const std::string& database_path,
                                 const GFTTOptions& detector_options, const OpticalFlowOptions& flow_options,
                                 bool write_images) {
    PCH_CHECK(frame_accessor);                                            // opticalflow.cc:216
    (void)write_images;   // debug PNG dumps (opticalflow.cc:80-96) need an image codec: not provided
    Database db{database_path};
    const int32_t from = video_info.first_frame;
    const int32_t to = video_info.first_frame + (int32_t)video_info.num_frames;

    const int w = (int)video_info.width, h = (int)video_info.height;
    // max_corners == 0 means unlimited (gftt.h:10): corners at least min_distance apart cannot be denser than
    // a hexagonal packing (1.155 w h / d^2), and 3x3 NMS leaves at most one candidate per 2x2 block
    int max_features = detector_options.max_corners;
    if (max_features <= 0) {
        const double d = std::max(1.0, detector_options.min_distance);
        const double bound = std::min((double)w * h / 4.0, 1.2 * (double)w * h / (d * d) + 1024.0);
        max_features = std::max(16384, (int)bound);
    }
    // Resume: keypoints already in the database are used as they are (they may be more numerous than
    // max_corners).  They are read now, because once the write-behind thread runs the connection is its.
    std::map<int32_t, Keypoints> stored_kps;
    for (int32_t f = from; f < to; f++) {
        if (!db.KeypointsExist(f)) continue;
        Keypoints kps = db.ReadKeypoints(f);
        max_features = std::max(max_features, (int)kps.size());
        // a stored row is authoritative even when it is empty (ReadOrGenerateKeypoints, opticalflow.cc:168-178:
        // the reference only detects when no row exists), so flows never index keypoints that are not in the database
        stored_kps.emplace(f, std::move(kps));
    }
    auto dc = AcquireDeviceContext(w, h, max_features);
    pc_ctx* ctx = dc->ctx;
    std::lock_guard<std::mutex> lk(dc->mtx);

    pc_video_info vi{video_info.width, video_info.height, video_info.first_frame, video_info.num_frames};
    const pc_gftt_opts go = ToAbi(detector_options);
    const pc_flow_opts fo = ToAbi(flow_options);
    Check(ctx, pc_analyze_begin(ctx, &vi, &go, &fo));
    struct EndGuard {
        pc_ctx* c;
        ~EndGuard() { pc_analyze_end(c); }
    } guard{ctx};

    std::deque<std::shared_ptr<void>> alive;     // frames stay alive until their upload was consumed
    FlowWriter writer(db);
    auto drain_one = [&]() {
        pc_frame_result r;
        Check(ctx, pc_analyze_pop(ctx, &r, 1));
        FlowWriter::Item it;
        it.frame_id = r.frame_id;
        it.kps.assign(r.keypoints, r.keypoints + 2 * (size_t)r.num_keypoints);
        it.pairs.resize((size_t)r.num_pairs);
        for (int k = 0; k < r.num_pairs; k++) {
            const pc_pair_rows& p = r.pairs[k];
            FlowWriter::Pair& q = it.pairs[(size_t)k];
            q.from = p.image_id_from;
            q.to = p.image_id_to;
            q.idx.assign(p.src_kps_indices, p.src_kps_indices + p.rows);
            q.tgt.assign(p.tgt_kps, p.tgt_kps + 2 * (size_t)p.rows);
            q.err.assign(p.flow_errors, p.flow_errors + p.rows);
        }
        writer.Push(std::move(it));
        if (!alive.empty()) alive.pop_front();
    };

    for (int32_t frame_id = from; frame_id < to; frame_id++) {
        if (callback) {                                                   // :238-247
            const double progress = static_cast<float>(frame_id - from) / video_info.num_frames;
            const bool ok = callback((float)progress, Format("Processing frame %d", frame_id));
            if (!ok) {
                callback(1.0, "Cancelled");
                return;
            }
        }
        const std::optional<Frame> frame = frame_accessor(frame_id);
        if (!frame) throw std::runtime_error(Format("Rquested frame #%d was not provided", frame_id));   // :251-254
        PCH_CHECK((uint32_t)frame->height == video_info.height);          // :196-198
        PCH_CHECK((uint32_t)frame->width == video_info.width);
        const auto stored = stored_kps.find(frame_id);
        if (stored != stored_kps.end())
            Check(ctx, pc_analyze_preset_keypoints(ctx, frame_id, stored->second.empty() ? nullptr : stored->second[0].data(),
                                                   (int)stored->second.size()));
        if (pc_analyze_pending(ctx) >= 3) drain_one();
        Check(ctx, pc_analyze_push_frame(ctx, frame_id, frame->data, frame->stride,
                                         frame->pinned ? PC_MEM_HOST_PINNED : PC_MEM_HOST));
        alive.push_back(frame->keep_alive);
    }
    while (pc_analyze_pending(ctx) > 0) drain_one();
    writer.Finish();                                                      // every row is in the database
    if (callback) callback(1.0, "Done");
}

// ---- Track -----------------------------------------------------------------------------------
namespace {

struct SolveFrameCache {
    std::vector<Keypoints> keypoints;
    std::vector<ImagePairFlow> flows;
    std::vector<pc_match_source> srcs;
};

// SolveFrame (tracker.cc:36-131)
std::optional<PnPResult> SolveFrame(DeviceContext& dc, const Database& database, const CameraTrajectory& camera_traj,
                                    const Mat4& model_matrix, int32_t frame_id, bool optimize_focal_length,
                                    bool optimize_principal_point, const BundleOptions& bundle_opts,
                                    SolveFrameCache& cache) {
    cache.keypoints.clear();
    cache.flows.clear();
    cache.srcs.clear();
    const std::vector<int32_t> ids = database.FindOpticalFlowsToImage(frame_id);
    cache.keypoints.reserve(ids.size());
    cache.flows.reserve(ids.size());
    for (int32_t flow_frame_id : ids) {
        PCH_CHECK(flow_frame_id != frame_id);                             // :46
        if (!camera_traj.IsFrameFilled(flow_frame_id)) continue;          // :48-50
        cache.keypoints.push_back(database.ReadKeypoints(flow_frame_id));
        cache.flows.push_back(database.ReadImagePairFlow(flow_frame_id, frame_id));
        const ImagePairFlow& flow = cache.flows.back();
        PCH_CHECK(flow.src_kps_indices.size() == flow.tgt_kps.size());    // :55
        const Keypoints& kps = cache.keypoints.back();
        pc_match_source s{};
        s.camera = ToAbi(*camera_traj.Get(flow_frame_id));
        s.keypoints = kps.empty() ? nullptr : kps[0].data();
        s.nk = (int32_t)kps.size();
        s.src_kps_indices = flow.src_kps_indices.data();
        s.tgt_kps = flow.tgt_kps.empty() ? nullptr : flow.tgt_kps[0].data();
        s.rows = (int32_t)flow.src_kps_indices.size();
        cache.srcs.push_back(s);
    }
    PnPResult result;
    // the solution should be very close to the previous / next pose (:111-119)
    if (camera_traj.IsFrameFilled(frame_id)) result.camera = *camera_traj.Get(frame_id);
    else if (camera_traj.IsFrameFilled(frame_id - 1)) result.camera = *camera_traj.Get(frame_id - 1);
    else if (camera_traj.IsFrameFilled(frame_id + 1)) result.camera = *camera_traj.Get(frame_id + 1);

    const pc_camera_state init = ToAbi(result.camera);
    const pc_bundle_opts bo = ToAbi(bundle_opts);
    pc_camera_state out{};
    pc_bundle_stats stats{};
    float inlier_ratio = 0;
    int matches = 0;
    const int rc = pc_track_frame(dc.ctx, cache.srcs.data(), (int)cache.srcs.size(), model_matrix.data(), &init, &bo,
                                  optimize_focal_length, optimize_principal_point, &out, &stats, &inlier_ratio,
                                  &matches);
    if (rc == PC_ERR_NOT_ENOUGH_FEATURES) return std::nullopt;            // :95-97
    Check(dc.ctx, rc);
    result.camera = FromAbi(out);
    result.bundle_stats = FromAbi(stats);
    result.inlier_ratio = inlier_ratio;
    return result;
}

}  // namespace

void TrackCameraTrajectory(const Database& database, CameraTrajectory& camera_traj, int32_t frame_from,
                           int32_t frame_to_inclusive, const Mat4& model_matrix, const AcceleratedMesh& accel_mesh,
                           const TrackingCallback& callback, bool optimize_focal_length, bool optimize_principal_point,
                           const BundleOptions& opts) {
    const int32_t first_frame = std::min(frame_from, frame_to_inclusive);
    const int32_t last_frame = std::max(frame_from, frame_to_inclusive);
    const int32_t dir = (frame_from < frame_to_inclusive) ? 1 : -1;
    PCH_CHECK(camera_traj.IsValidFrame(first_frame));                     // tracker.cc:147-149
    PCH_CHECK(camera_traj.IsValidFrame(last_frame));
    PCH_CHECK(camera_traj.IsFrameFilled(frame_from));

    auto dc = AcquireDeviceContext(0, 0, 0);
    // the context is locked per frame and released around the user callback: a callback that calls back
    // into this module (or a Python callback waiting for the GIL) must not hold up other users of the context
    std::unique_lock<std::mutex> lk(dc->mtx);
    SolveFrameCache cache;
    for (int32_t frame_id = frame_from + dir; frame_id != frame_to_inclusive + dir; frame_id += dir) {
        if (!lk.owns_lock()) lk.lock();
        accel_mesh.Bind(*dc);                                             // no-op while this mesh is the context's current one
        const std::optional<PnPResult> maybe = SolveFrame(*dc, database, camera_traj, model_matrix, frame_id,
                                                          optimize_focal_length, optimize_principal_point, opts, cache);
        lk.unlock();
        if (!maybe)                                                       // :162-166
            throw std::runtime_error(Format("Could not track to frame: %d. Not enough features.", frame_id));
        if (callback) {
            FrameTrackingResult r;
            r.frame = frame_id;
            r.pose = maybe->camera.pose;
            r.intrinsics = maybe->camera.intrinsics;
            r.bundle_stats = maybe->bundle_stats;
            r.inlier_ratio = maybe->inlier_ratio;
            if (!callback(r)) return;                                     // :178-183
        }
        camera_traj.Set(frame_id, maybe->camera);
    }
}

void TrackSequence(const std::string& database_path, int32_t frame_from, int32_t frame_to_inclusive,
                   const SceneTransformations& scene_transform, const AcceleratedMesh& accel_mesh,
                   const TrackingCallback& callback, bool optimize_focal_length, bool optimize_principal_point,
                   BundleOptions bundle_opts) {
    const Database database{database_path};
    const size_t num_frames = (size_t)std::abs(frame_to_inclusive - frame_from) + 1;
    CameraTrajectory camera_traj{std::min(frame_from, frame_to_inclusive), num_frames};
    camera_traj.Set(frame_from, CameraState{scene_transform.intrinsics, Pose::FromRt(scene_transform.view_matrix)});
    TrackCameraTrajectory(database, camera_traj, frame_from, frame_to_inclusive, scene_transform.model_matrix,
                          accel_mesh, callback, optimize_focal_length, optimize_principal_point, bundle_opts);
}

// ---- Refine ----------------------------------------------------------------------------------
namespace {

struct Bbox2 {
    Vec2 pmin, pmax;
    bool Contains(const Vec2& p) const {                                  // geometry.h:45-48 (strict)
        return p[0] > pmin[0] && p[1] > pmin[1] && p[0] < pmax[0] && p[1] < pmax[1];
    }
};

// TransformBbox + ComputeBbox (refiner.cc:18-69)
Bbox2 ComputeBbox(const CameraState& state, const Mesh& mesh, const Mat4& model_matrix) {
    const Mat4 mvp = MatMul(MatMul(state.intrinsics.To4x4ProjectionMatrix(), state.pose.Rt4x4()), model_matrix);
    Vec2 pmin{std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    Vec2 pmax{std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest()};
    for (int c = 0; c < 8; c++) {
        const float x = (c & 4) ? mesh.bbox.pmax[0] : mesh.bbox.pmin[0];
        const float y = (c & 2) ? mesh.bbox.pmax[1] : mesh.bbox.pmin[1];
        const float z = (c & 1) ? mesh.bbox.pmax[2] : mesh.bbox.pmin[2];
        const float tx = mvp[0] * x + mvp[1] * y + mvp[2] * z + mvp[3];
        const float ty = mvp[4] * x + mvp[5] * y + mvp[6] * z + mvp[7];
        const float tw = mvp[12] * x + mvp[13] * y + mvp[14] * z + mvp[15];
        const float px = tx / tw, py = ty / tw;                           // hnormalized
        pmin[0] = std::min(px, pmin[0]); pmax[0] = std::max(px, pmax[0]);
        pmin[1] = std::min(py, pmin[1]); pmax[1] = std::max(py, pmax[1]);
    }
    constexpr float kPadding = 20.0f;
    return Bbox2{{pmin[0] - kPadding, pmin[1] - kPadding}, {pmax[0] + kPadding, pmax[1] + kPadding}};
}

}  // namespace

void RefineTrajectory(const std::string& database_path, CameraTrajectory& traj, const Mat4& model_matrix,
                      const AcceleratedMesh& mesh, bool optimize_focal_length, bool optimize_principal_point,
                      const RefineTrajectoryCallback& callback, BundleOptions bundle_opts) {
    Database database{database_path};
    PCH_CHECK(traj.Count() > 2);                                          // refiner.cc:661
    for (int32_t frame = traj.FirstFrame(); frame <= traj.LastFrame(); frame++) PCH_CHECK(traj.IsFrameFilled(frame));

    // CachedDatabase (refiner.cc:71-197): keypoints filtered to the projected mesh bbox + 20 px,
    // flows remapped to the filtered indices, empty flows dropped.
    const int nf = (int)traj.Count();
    constexpr size_t kInvalidIdx = std::numeric_limits<size_t>::max();
    std::vector<int32_t> kp_offsets(nf + 1, 0);
    std::vector<float> all_kps;
    std::vector<pc_ba_edge> edges;
    std::vector<uint32_t> src_idx;
    std::vector<float> tgt_kps;
    for (int f = 0; f < nf; f++) {
        const int32_t frame_id = traj.FirstFrame() + f;
        Keypoints keypoints = database.ReadKeypoints(frame_id);
        const Bbox2 bbox = ComputeBbox(*traj.Get(frame_id), mesh.Inner(), model_matrix);
        std::vector<size_t> to_filtered(keypoints.size());
        size_t k = 0;
        for (size_t j = 0; j < keypoints.size(); j++) {
            if (bbox.Contains(keypoints[j])) {
                keypoints[k] = keypoints[j];
                to_filtered[j] = k++;
            } else {
                to_filtered[j] = kInvalidIdx;
            }
        }
        keypoints.resize(k);
        for (const Vec2& p : keypoints) { all_kps.push_back(p[0]); all_kps.push_back(p[1]); }
        kp_offsets[f + 1] = kp_offsets[f] + (int32_t)k;
        for (int32_t frame_id_to : database.FindOpticalFlowsFromImage(frame_id)) {
            if (!traj.IsValidFrame(frame_id_to)) continue;                // refiner.cc:133-135
            const ImagePairFlow flow = database.ReadImagePairFlow(frame_id, frame_id_to);
            pc_ba_edge e{};
            e.src_frame_idx = f;
            e.tgt_frame_idx = frame_id_to - traj.FirstFrame();
            e.first_row = (int32_t)src_idx.size();
            for (size_t j = 0; j < flow.tgt_kps.size(); j++) {
                PCH_CHECK(flow.src_kps_indices[j] < to_filtered.size());
                const size_t fi = to_filtered[flow.src_kps_indices[j]];
                if (fi == kInvalidIdx) continue;
                src_idx.push_back((uint32_t)fi);
                tgt_kps.push_back(flow.tgt_kps[j][0]);
                tgt_kps.push_back(flow.tgt_kps[j][1]);
            }
            e.rows = (int32_t)src_idx.size() - e.first_row;
            if (e.rows != 0) edges.push_back(e);                          // refiner.cc:154-160
        }
    }

    auto dc = AcquireDeviceContext(0, 0, 0);
    std::lock_guard<std::mutex> lk(dc->mtx);
    mesh.Bind(*dc);
    pc_ba_problem pr{};
    pr.num_frames = nf;
    pr.kp_offsets = kp_offsets.data();
    pr.keypoints = all_kps.data();
    pr.num_edges = (int32_t)edges.size();
    pr.edges = edges.data();
    pr.src_kps_indices = src_idx.data();
    pr.tgt_kps = tgt_kps.data();
    memcpy(pr.model, model_matrix.data(), sizeof(pr.model));
    pr.optimize_focal_length = optimize_focal_length;
    pr.optimize_principal_point = optimize_principal_point;
    Check(dc->ctx, pc_ba_load(dc->ctx, &pr));

    std::vector<pc_camera_state> states(nf);
    for (int f = 0; f < nf; f++) states[f] = ToAbi(*traj.Get(traj.FirstFrame() + f));
    struct CbData {
        const RefineTrajectoryCallback* cb;
        size_t max_iterations;
    } cbd{&callback, bundle_opts.max_iterations};
    auto trampoline = [](const pc_bundle_stats* s, void* user) -> int {   // refiner.cc:670-678
        CbData* d = static_cast<CbData*>(user);
        if (!*d->cb) return 1;
        RefineTrajectoryUpdate u;
        u.progress = static_cast<float>(s->iterations) / d->max_iterations;
        u.message = Format("Cost: %.02f (Initial: %.02f)", s->cost, s->initial_cost);
        u.stats = FromAbi(*s);
        // The context stays locked: the loaded refine problem lives in it and must not be replaced mid-solve.
        // Nothing that holds the GIL takes this mutex (interactive ray casts use their own context,
        // types.cc::AcquireInteractiveContext), so a Python callback cannot deadlock against it.
        return (*d->cb)(u) ? 1 : 0;
    };
    const pc_bundle_opts bo = ToAbi(bundle_opts);
    pc_bundle_stats stats{};
    Check(dc->ctx, pc_ba_solve(dc->ctx, &bo, states.data(), &stats, trampoline, &cbd));
    for (int f = 0; f < nf; f++) traj.Set(traj.FirstFrame() + f, FromAbi(states[f]));
}

}  // namespace pch
