// Pin mode: the interactive "drag a pinned vertex, the object (or the camera) follows" solve.
//   PinUpdate, FindTransformation     /root/reference/cpp/pin_mode.h:11-24, pin_mode.cc:16-246
// One and two pins are closed-form; three and more go through SolvePnPIterative (K11 on the GPU, through
// pc_solve_pnp) from the current transform, so that dragging stays smooth.
#pragma once

#include <vector>

#include "types.h"

namespace pch {

struct PinUpdate {
    uint32_t pin_idx = 0;
    Vec2 pos{0, 0};
};

// object_points: n x 3, row-major
SceneTransformations FindTransformation(const std::vector<float>& object_points,
                                        const SceneTransformations& initial_scene_transform,
                                        const SceneTransformations& current_scene_transform, const PinUpdate& update,
                                        TransformationType trans_type, bool optimize_focal_length,
                                        bool optimize_principal_point);

}  // namespace pch
