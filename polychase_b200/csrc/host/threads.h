// Worker-thread wrappers with the reference's contract (cpp/opticalflow_thread.h,
// cpp/tracker_thread.h, cpp/refiner_thread.h): the constructor starts the worker, results flow
// through a queue polled with TryPop()/Empty(), the terminal message is always `true`, preceded by
// an exception object on failure, RequestStop() cancels at the next frame / iteration, Join()
// waits.  No Python callback is ever invoked from the worker thread.
//
// Deliberate differences: the exception message is preserved (the
// reference slices it to "std::exception", tracker_thread.h:76), and a frame hand-off that
// times out raises instead of dereferencing an empty optional (opticalflow_thread.h:154).
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <optional>
#include <string>
#include <thread>
#include <variant>
#include <vector>

#include "pipelines.h"

namespace pch {

struct CppException {
    std::string message;
    const char* what() const { return message.c_str(); }
};

template <typename Msg>
class ResultQueue {
   public:
    void Push(Msg m) {
        std::lock_guard<std::mutex> lk(mtx_);
        q_.push_back(std::move(m));
    }
    std::optional<Msg> TryPop() {
        std::lock_guard<std::mutex> lk(mtx_);
        if (q_.empty()) return std::nullopt;
        Msg m = std::move(q_.front());
        q_.pop_front();
        return m;
    }
    bool Empty() const {
        std::lock_guard<std::mutex> lk(mtx_);
        return q_.empty();
    }

   private:
    mutable std::mutex mtx_;
    std::deque<Msg> q_;
};

struct OpticalFlowProgress {
    float progress;
    std::string progress_message;
};
struct OpticalFlowRequest {
    int32_t frame_id;
};
using OpticalFlowThreadMessage = std::variant<OpticalFlowProgress, OpticalFlowRequest, bool, CppException>;

// Page-locked frame buffers for the OpticalFlowThread hand-off, allocated through the C ABI on the
// process-wide interactive context, under its mutex (page-locked memory is usable by every context
// of the device).
class PinnedFrameRing {
   public:
    static constexpr int kFrameRing = 6;
    ~PinnedFrameRing() { Release(); }
    // next slot of `bytes` bytes, or nullptr if page-locked memory cannot be had
    uint8_t* Slot(size_t bytes) {
        if (failed_) return nullptr;
        if (bytes != bytes_) {
            Release();
            try {
                dc_ = AcquireInteractiveContext();
            } catch (...) {
                failed_ = true;
                return nullptr;
            }
            for (int i = 0; i < kFrameRing; i++) {
                void* p = nullptr;
                int rc;
                {
                    std::lock_guard<std::mutex> lk(dc_->mtx);
                    rc = pc_host_alloc_pinned(dc_->ctx, bytes, &p);
                }
                if (rc != PC_OK) {
                    Release();
                    failed_ = true;
                    return nullptr;
                }
                slots_[i] = static_cast<uint8_t*>(p);
            }
            bytes_ = bytes;
        }
        uint8_t* s = slots_[next_];
        next_ = (next_ + 1) % kFrameRing;
        return s;
    }

   private:
    void Release() {
        for (auto& p : slots_) {
            if (p && dc_) {
                std::lock_guard<std::mutex> lk(dc_->mtx);
                pc_host_free_pinned(dc_->ctx, p);
            }
            p = nullptr;
        }
        bytes_ = 0;
        next_ = 0;
    }
    std::shared_ptr<DeviceContext> dc_;
    uint8_t* slots_[kFrameRing] = {nullptr};
    size_t bytes_ = 0;
    int next_ = 0;
    bool failed_ = false;
};

// Persistent helper threads for the frame hand-off copy: a 4K RGB frame is 25 MB, and one core copies about
// 10 GB/s.  (Spawning threads per frame, as the first version did, costs ~0.2 ms of the ~1.5 ms.)
class CopyPool {
   public:
    explicit CopyPool(int helpers) {
        for (int i = 0; i < helpers; i++) threads_.emplace_back([this, i] { Run(i); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mtx_);
            quit_ = true;
            generation_++;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int parts() const { return (int)threads_.size() + 1; }
    // fn(part) for part = 0 .. parts()-1, part 0 on the calling thread; returns when all are done
    void Run(const std::function<void(int)>& fn) {
        {
            std::lock_guard<std::mutex> lk(mtx_);
            fn_ = &fn;
            pending_ = (int)threads_.size();
            generation_++;
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> lk(mtx_);
        done_cv_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

   private:
    void Run(int index) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* fn;
            {
                std::unique_lock<std::mutex> lk(mtx_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                if (quit_) return;
                fn = fn_;
            }
            (*fn)(index + 1);
            {
                std::lock_guard<std::mutex> lk(mtx_);
                pending_--;
            }
            done_cv_.notify_one();
        }
    }
    std::vector<std::thread> threads_;
    std::mutex mtx_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)>* fn_ = nullptr;
    int pending_ = 0;
    uint64_t generation_ = 0;
    bool quit_ = false;
};

class OpticalFlowThread {
   public:
    OpticalFlowThread(VideoInfo video_info, std::string database_path, GFTTOptions detector_options = {},
                      OpticalFlowOptions flow_options = {}, bool write_images = false)
        : video_info_(video_info), database_path_(std::move(database_path)), detector_options_(detector_options),
          flow_options_(flow_options), write_images_(write_images) {
        worker_ = std::thread([this] { Work(); });
    }
    ~OpticalFlowThread() {
        RequestStop();
        Join();
    }
    void RequestStop() {
        {
            std::lock_guard<std::mutex> lk(frame_mtx_);
            stop_ = true;
        }
        frame_cv_.notify_all();
    }
    void Join() {
        if (worker_.joinable()) worker_.join();
    }
    std::optional<OpticalFlowThreadMessage> TryPop() { return queue_.TryPop(); }
    bool Empty() const { return queue_.Empty(); }
    // The frame is deep-copied on the calling thread (opticalflow_thread.h:120-132) -- here straight
    // into a ring of page-locked buffers, so the upload that follows is one asynchronous DMA with no
    // staging copy and no per-frame allocation.  A slot is reused kFrameRing frames later; the
    // analyzer keeps at most 4 frames in flight (pipelines.cc), whose uploads finished long before.
    void ProvideFrame(int32_t frame_id, const uint8_t* rgb, int width, int height, size_t stride) {
        const size_t row = (size_t)width * 3, bytes = row * (size_t)height;
        Frame f;
        f.width = width;
        f.height = height;
        f.stride = row;
        uint8_t* dst = ring_.Slot(bytes);
        if (dst) {
            f.pinned = true;
        } else {                                   // no page-locked memory: an ordinary copy
            auto buf = std::shared_ptr<uint8_t[]>(new uint8_t[bytes]);
            dst = buf.get();
            f.keep_alive = buf;
        }
        f.data = dst;
        // 4K RGB is 25 MB: split the copy over the pool (7 helpers + this thread)
        auto copy_rows = [&](int y0, int y1) {
            if (stride == row) memcpy(dst + (size_t)y0 * row, rgb + (size_t)y0 * row, (size_t)(y1 - y0) * row);
            else for (int y = y0; y < y1; y++) memcpy(dst + (size_t)y * row, rgb + (size_t)y * stride, row);
        };
        if (bytes < (8u << 20)) {
            copy_rows(0, height);
        } else {
            if (!pool_) pool_ = std::make_unique<CopyPool>(7);
            const int parts = pool_->parts();
            const std::function<void(int)> job = [&](int k) { copy_rows((int)((int64_t)height * k / parts), (int)((int64_t)height * (k + 1) / parts)); };
            pool_->Run(job);
        }
        {
            std::lock_guard<std::mutex> lk(frame_mtx_);
            provided_ = std::make_pair(frame_id, std::move(f));
        }
        frame_cv_.notify_all();
    }

   private:
    void Work() {
        auto accessor = [this](int32_t frame_id) -> std::optional<Frame> {
            queue_.Push(OpticalFlowRequest{frame_id});
            std::unique_lock<std::mutex> lk(frame_mtx_);
            frame_cv_.wait_for(lk, std::chrono::seconds(10), [&] { return provided_.has_value() || stop_; });
            if (stop_) return std::nullopt;
            if (!provided_) throw std::runtime_error("Requested frame " + std::to_string(frame_id) + " was not provided in 10 s");
            if (provided_->first != frame_id)
                throw std::runtime_error("Requested frame " + std::to_string(frame_id) + " but got " +
                                         std::to_string(provided_->first));
            Frame f = std::move(provided_->second);
            provided_.reset();
            return f;
        };
        auto progress = [this](float p, const std::string& msg) {
            queue_.Push(OpticalFlowProgress{p, msg});
            std::lock_guard<std::mutex> lk(frame_mtx_);
            return !stop_;
        };
        try {
            GenerateOpticalFlowDatabase(video_info_, accessor, progress, database_path_, detector_options_,
                                        flow_options_, write_images_);
        } catch (const std::exception& e) {
            bool stopping;
            {
                std::lock_guard<std::mutex> lk(frame_mtx_);
                stopping = stop_;
            }
            if (!stopping) queue_.Push(CppException{e.what()});
        } catch (...) {
            queue_.Push(CppException{"Unknown exception type. This should never happen!"});
        }
        queue_.Push(true);
    }

    const VideoInfo video_info_;
    const std::string database_path_;
    const GFTTOptions detector_options_;
    const OpticalFlowOptions flow_options_;
    const bool write_images_;
    ResultQueue<OpticalFlowThreadMessage> queue_;
    std::optional<std::pair<int32_t, Frame>> provided_;
    PinnedFrameRing ring_;       // touched by the providing thread only
    std::unique_ptr<CopyPool> pool_;   // likewise
    std::mutex frame_mtx_;
    std::condition_variable frame_cv_;
    bool stop_ = false;
    std::thread worker_;
};

using TrackerThreadMessage = std::variant<FrameTrackingResult, bool, CppException>;

class TrackerThread {
   public:
    TrackerThread(std::string database_path, int32_t frame_from, int32_t frame_to_inclusive,
                  SceneTransformations scene_transform, std::shared_ptr<const AcceleratedMesh> accel_mesh,
                  bool optimize_focal_length, bool optimize_principal_point, BundleOptions bundle_opts)
        : database_path_(std::move(database_path)), frame_from_(frame_from), frame_to_(frame_to_inclusive),
          scene_(scene_transform), mesh_(std::move(accel_mesh)), opt_f_(optimize_focal_length),
          opt_pp_(optimize_principal_point), opts_(bundle_opts) {
        worker_ = std::thread([this] { Work(); });
    }
    ~TrackerThread() { Join(); }
    void RequestStop() { stop_ = true; }
    void Join() {
        if (worker_.joinable()) worker_.join();
    }
    std::optional<TrackerThreadMessage> TryPop() { return queue_.TryPop(); }
    bool Empty() const { return queue_.Empty(); }

   private:
    void Work() {
        auto cb = [this](const FrameTrackingResult& r) {
            queue_.Push(r);
            return !stop_.load();
        };
        try {
            TrackSequence(database_path_, frame_from_, frame_to_, scene_, *mesh_, cb, opt_f_, opt_pp_, opts_);
        } catch (const std::exception& e) {
            queue_.Push(CppException{e.what()});
        } catch (...) {
            queue_.Push(CppException{"Unknown exception type. This should never happen!"});
        }
        queue_.Push(true);
    }
    const std::string database_path_;
    const int32_t frame_from_, frame_to_;
    const SceneTransformations scene_;
    const std::shared_ptr<const AcceleratedMesh> mesh_;
    const bool opt_f_, opt_pp_;
    const BundleOptions opts_;
    ResultQueue<TrackerThreadMessage> queue_;
    std::atomic<bool> stop_{false};
    std::thread worker_;
};

using RefinerThreadMessage = std::variant<RefineTrajectoryUpdate, bool, CppException>;

class RefinerThread {
   public:
    RefinerThread(std::string database_path, std::shared_ptr<CameraTrajectory> traj, Mat4 model_matrix,
                  std::shared_ptr<const AcceleratedMesh> mesh, bool optimize_focal_length,
                  bool optimize_principal_point, BundleOptions bundle_opts)
        : database_path_(std::move(database_path)), traj_(std::move(traj)), model_(model_matrix),
          mesh_(std::move(mesh)), opt_f_(optimize_focal_length), opt_pp_(optimize_principal_point),
          opts_(bundle_opts) {
        worker_ = std::thread([this] { Work(); });
    }
    ~RefinerThread() { Join(); }
    void RequestStop() { stop_ = true; }
    void Join() {
        if (worker_.joinable()) worker_.join();
    }
    std::optional<RefinerThreadMessage> TryPop() { return queue_.TryPop(); }
    bool Empty() const { return queue_.Empty(); }

   private:
    void Work() {
        auto cb = [this](RefineTrajectoryUpdate u) {
            queue_.Push(std::move(u));
            return !stop_.load();
        };
        try {
            // the trajectory is mutated in place; Python reads it after the terminal message
            RefineTrajectory(database_path_, *traj_, model_, *mesh_, opt_f_, opt_pp_, cb, opts_);
        } catch (const std::exception& e) {
            queue_.Push(CppException{e.what()});
        } catch (...) {
            queue_.Push(CppException{"Unknown exception type. This should never happen!"});
        }
        queue_.Push(true);
    }
    const std::string database_path_;
    const std::shared_ptr<CameraTrajectory> traj_;
    const Mat4 model_;
    const std::shared_ptr<const AcceleratedMesh> mesh_;
    const bool opt_f_, opt_pp_;
    const BundleOptions opts_;
    ResultQueue<RefinerThreadMessage> queue_;
    std::atomic<bool> stop_{false};
    std::thread worker_;
};

}  // namespace pch
