// The three blocking entry points of the reference, same names and argument meaning:
//   GenerateOpticalFlowDatabase   /root/reference/cpp/opticalflow.h:35-41
//   TrackSequence / TrackCameraTrajectory   /root/reference/cpp/tracker.h:27-39
//   RefineTrajectory              /root/reference/cpp/refiner.h:22-27
#pragma once

#include <functional>
#include <optional>
#include <string>

#include "database.h"
#include "types.h"

namespace pch {

// A frame handed over by the caller: H x W x 3 uint8 RGB, rows `stride` bytes apart.  `keep_alive`
// owns the memory.
struct Frame {
    const uint8_t* data = nullptr;
    int width = 0, height = 0;
    size_t stride = 0;
    std::shared_ptr<void> keep_alive;
    bool pinned = false;      // `data` is page-locked (pc_host_alloc_pinned): uploaded without a staging copy
};
using FrameAccessorFunction = std::function<std::optional<Frame>(int32_t frame_id)>;
using OpticalFlowProgressCallback = std::function<bool(float progress, const std::string& progress_message)>;

// Callbacks are taken by const reference and never copied: a std::function that wraps a Python callable
// must not be copied or destroyed on a thread that does not hold the GIL (module.cc).
void GenerateOpticalFlowDatabase(const VideoInfo& video_info, const FrameAccessorFunction& frame_accessor,
                                 const OpticalFlowProgressCallback& callback, const std::string& database_path,
                                 const GFTTOptions& detector_options = {}, const OpticalFlowOptions& flow_options = {},
                                 bool write_images = false);

struct FrameTrackingResult {
    int32_t frame = 0;
    Pose pose;
    CameraIntrinsics intrinsics;
    BundleStats bundle_stats;
    float inlier_ratio = 0;
};
using TrackingCallback = std::function<bool(const FrameTrackingResult&)>;

void TrackSequence(const std::string& database_path, int32_t frame_from, int32_t frame_to_inclusive,
                   const SceneTransformations& scene_transform, const AcceleratedMesh& accel_mesh,
                   const TrackingCallback& callback, bool optimize_focal_length, bool optimize_principal_point,
                   BundleOptions opts);

void TrackCameraTrajectory(const Database& database, CameraTrajectory& camera_traj, int32_t frame_from,
                           int32_t frame_to_inclusive, const Mat4& model_matrix, const AcceleratedMesh& accel_mesh,
                           const TrackingCallback& callback, bool optimize_focal_length, bool optimize_principal_point,
                           const BundleOptions& opts);

struct RefineTrajectoryUpdate {
    float progress = 0;
    std::string message;
    BundleStats stats;
};
using RefineTrajectoryCallback = std::function<bool(RefineTrajectoryUpdate)>;

void RefineTrajectory(const std::string& database_path, CameraTrajectory& traj, const Mat4& model_matrix,
                      const AcceleratedMesh& mesh, bool optimize_focal_length, bool optimize_principal_point,
                      const RefineTrajectoryCallback& callback, BundleOptions bundle_opts);

}  // namespace pch
