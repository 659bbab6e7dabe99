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
#include <memory>
#include <mutex>
#include <optional>
#include <string>
#include <thread>
#include <variant>

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
    // The frame is deep-copied on the calling thread (opticalflow_thread.h:120-132).
    void ProvideFrame(int32_t frame_id, const uint8_t* rgb, int width, int height, size_t stride) {
        auto buf = std::shared_ptr<uint8_t[]>(new uint8_t[(size_t)height * width * 3]);
        for (int y = 0; y < height; y++) memcpy(buf.get() + (size_t)y * width * 3, rgb + (size_t)y * stride, (size_t)width * 3);
        Frame f;
        f.data = buf.get();
        f.width = width;
        f.height = height;
        f.stride = (size_t)width * 3;
        f.keep_alive = buf;
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
