// pybind11 module `polychase_core`: the Python surface of /root/reference/cpp/polychase_pybind.cc:29-348
// (same class, method, argument and enum names) over the B200 host pipelines, including the interactive
// pin mode (PinUpdate, find_transformation; pin_mode.cc) whose three-and-more-pin solve runs on K11.
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdio>
#include <cstring>

#include "database.h"
#include "pin_mode.h"
#include "pipelines.h"
#include "threads.h"
#include "types.h"

namespace py = pybind11;
using namespace pch;

using FArr = py::array_t<float, py::array::c_style | py::array::forcecast>;
using UArr = py::array_t<uint32_t, py::array::c_style | py::array::forcecast>;
using U8Arr = py::array_t<uint8_t, py::array::c_style>;

static FArr KpsToNumpy(const Keypoints& k) {
    FArr a({(py::ssize_t)k.size(), (py::ssize_t)2});
    if (!k.empty()) memcpy(a.mutable_data(), k[0].data(), k.size() * 2 * sizeof(float));
    return a;
}
static Keypoints KpsFromNumpy(const FArr& a) {
    if (a.ndim() != 2 || a.shape(1) != 2) throw std::invalid_argument("expected an (N, 2) float32 array");
    Keypoints k((size_t)a.shape(0));
    if (!k.empty()) memcpy(k[0].data(), a.data(), k.size() * 2 * sizeof(float));
    return k;
}
template <typename T, typename A>
static std::vector<T> VecFromNumpy(const A& a) {
    std::vector<T> v((size_t)a.size());
    if (!v.empty()) memcpy(v.data(), a.data(), v.size() * sizeof(T));
    return v;
}
static Mat4 Mat4FromNumpy(const FArr& a) {
    if (a.size() != 16) throw std::invalid_argument("expected a 4x4 matrix");
    Mat4 m;
    memcpy(m.data(), a.data(), sizeof(float) * 16);
    return m;
}
static FArr Mat4ToNumpy(const Mat4& m) {
    FArr a({4, 4});
    memcpy(a.mutable_data(), m.data(), sizeof(float) * 16);
    return a;
}
template <size_t N>
static FArr ArrToNumpy(const std::array<float, N>& v) {
    FArr a((py::ssize_t)N);
    memcpy(a.mutable_data(), v.data(), sizeof(float) * N);
    return a;
}
template <size_t N>
static std::array<float, N> ArrFromNumpy(const FArr& a) {
    if ((size_t)a.size() != N) throw std::invalid_argument("wrong vector length");
    std::array<float, N> v;
    memcpy(v.data(), a.data(), sizeof(float) * N);
    return v;
}

static Frame FrameFromNumpy(const U8Arr& a) {
    if (a.ndim() != 3 || a.shape(2) != 3) throw std::invalid_argument("expected an H x W x 3 uint8 array");
    Frame f;
    f.height = (int)a.shape(0);
    f.width = (int)a.shape(1);
    f.stride = (size_t)a.strides(0);
    f.data = a.data();
    return f;
}

// Holder for a Python callable that is shared with code running without the GIL (the pipelines release
// it): the last reference may die on any thread, so the deleter takes the GIL before the decref.
static std::shared_ptr<py::function> HoldPy(py::function fn) {
    return std::shared_ptr<py::function>(new py::function(std::move(fn)), [](py::function* p) {
        py::gil_scoped_acquire gil;
        delete p;
    });
}

template <typename Variant>
static py::object VariantToPy(std::optional<Variant> m) {
    if (!m) return py::none();
    return std::visit([](auto&& v) -> py::object { return py::cast(std::move(v)); }, std::move(*m));
}

PYBIND11_MODULE(polychase_core, m) {
    m.doc() = "polychase_core on B200: OpticalFlow / Tracker / Refiner over hand-written sm_100a CUDA";

    py::class_<Mesh>(m, "Mesh")
        .def_property_readonly("vertices", [](const Mesh& s) {
            FArr a({(py::ssize_t)s.NumVertices(), (py::ssize_t)3});
            memcpy(a.mutable_data(), s.vertices.data(), s.vertices.size() * sizeof(float));
            return a; })
        .def_property_readonly("triangles", [](const Mesh& s) {
            UArr a({(py::ssize_t)s.NumTriangles(), (py::ssize_t)3});
            memcpy(a.mutable_data(), s.triangles.data(), s.triangles.size() * sizeof(uint32_t));
            return a; })
        .def_property("masked_triangles",
                      [](const Mesh& s) {
                          UArr a((py::ssize_t)s.masked_triangles.size());
                          memcpy(a.mutable_data(), s.masked_triangles.data(), s.masked_triangles.size() * 4);
                          return a; },
                      [](Mesh& s, const UArr& a) { s.masked_triangles = VecFromNumpy<uint32_t>(a); })
        .def("is_triangle_masked", &Mesh::IsTriangleMasked)
        .def("mask_triangle", &Mesh::MaskTriangle)
        .def("unmask_triangle", &Mesh::UnmaskTriangle)
        .def("toggle_mask_triangle", &Mesh::ToggleMaskTriangle);

    py::class_<AcceleratedMesh, std::shared_ptr<AcceleratedMesh>>(m, "AcceleratedMesh")
        .def(py::init([](const FArr& v, const UArr& t, const UArr& mask) {
                 if (v.ndim() != 2 || v.shape(1) != 3) throw std::invalid_argument("vertices must be (N, 3) float32");
                 if (t.ndim() != 2 || t.shape(1) != 3) throw std::invalid_argument("triangles must be (M, 3) uint32");
                 return std::make_shared<AcceleratedMesh>(VecFromNumpy<float>(v), VecFromNumpy<uint32_t>(t),
                                                          VecFromNumpy<uint32_t>(mask));
             }),
             py::arg("vertices"), py::arg("triangles"), py::arg("masked_triangles") = UArr(0))
        .def("inner", &AcceleratedMesh::Inner, py::return_value_policy::reference_internal)
        .def("inner_mut", &AcceleratedMesh::InnerMut, py::return_value_policy::reference_internal);

    py::enum_<TransformationType>(m, "TransformationType")
        .value("Camera", TransformationType::Camera)
        .value("Model", TransformationType::Model);
    py::enum_<CameraConvention>(m, "CameraConvention")
        .value("OpenGL", CameraConvention::OpenGL)
        .value("OpenCV", CameraConvention::OpenCV);

    py::class_<CameraIntrinsics>(m, "CameraIntrinsics")
        .def(py::init([](float fx, float fy, float cx, float cy, float aspect, float w, float h, CameraConvention c) {
                 return CameraIntrinsics{fx, fy, cx, cy, aspect, w, h, c};
             }),
             py::arg("fx"), py::arg("fy"), py::arg("cx"), py::arg("cy"), py::arg("aspect_ratio"), py::arg("width"),
             py::arg("height"), py::arg("convention") = CameraConvention::OpenGL)
        .def_readwrite("fx", &CameraIntrinsics::fx)
        .def_readwrite("fy", &CameraIntrinsics::fy)
        .def_readwrite("cx", &CameraIntrinsics::cx)
        .def_readwrite("cy", &CameraIntrinsics::cy)
        .def_readwrite("aspect_ratio", &CameraIntrinsics::aspect_ratio)
        .def_readwrite("width", &CameraIntrinsics::width)
        .def_readwrite("height", &CameraIntrinsics::height)
        .def_readwrite("convention", &CameraIntrinsics::convention);

    py::class_<SceneTransformations>(m, "SceneTransformations")
        .def(py::init([](const FArr& model, const FArr& view, CameraIntrinsics intr) {
                 return SceneTransformations{Mat4FromNumpy(model), Mat4FromNumpy(view), intr};
             }),
             py::arg("model_matrix"), py::arg("view_matrix"), py::arg("intrinsics"))
        .def_property("model_matrix", [](const SceneTransformations& s) { return Mat4ToNumpy(s.model_matrix); },
                      [](SceneTransformations& s, const FArr& a) { s.model_matrix = Mat4FromNumpy(a); })
        .def_property("view_matrix", [](const SceneTransformations& s) { return Mat4ToNumpy(s.view_matrix); },
                      [](SceneTransformations& s, const FArr& a) { s.view_matrix = Mat4FromNumpy(a); })
        .def_readwrite("intrinsics", &SceneTransformations::intrinsics);

    py::class_<RayHit>(m, "RayHit")
        .def_property_readonly("pos", [](const RayHit& h) { return ArrToNumpy(h.pos); })
        .def_property_readonly("normal", [](const RayHit& h) { return ArrToNumpy(h.normal); })
        .def_property_readonly("barycentric_coordinate", [](const RayHit& h) { return ArrToNumpy(h.barycentric_coordinate); })
        .def_readwrite("t", &RayHit::t)
        .def_readwrite("primitive_id", &RayHit::primitive_id);

    py::class_<PinUpdate>(m, "PinUpdate")                                   // polychase_pybind.cc:65-69
        .def(py::init([](uint32_t pin_idx, const FArr& pos) { return PinUpdate{pin_idx, ArrFromNumpy<2>(pos)}; }),
             py::arg("pin_idx"), py::arg("pin_pos"))
        .def_readwrite("pin_idx", &PinUpdate::pin_idx)
        .def_property("pos", [](const PinUpdate& u) { return ArrToNumpy(u.pos); },
                      [](PinUpdate& u, const FArr& a) { u.pos = ArrFromNumpy<2>(a); });

    py::class_<ImagePairFlow>(m, "ImagePairFlow")
        .def(py::init<>())
        .def_readwrite("image_id_from", &ImagePairFlow::image_id_from)
        .def_readwrite("image_id_to", &ImagePairFlow::image_id_to)
        .def_property("src_kps_indices",
                      [](const ImagePairFlow& f) {
                          UArr a((py::ssize_t)f.src_kps_indices.size());
                          memcpy(a.mutable_data(), f.src_kps_indices.data(), f.src_kps_indices.size() * 4);
                          return a; },
                      [](ImagePairFlow& f, const UArr& a) { f.src_kps_indices = VecFromNumpy<uint32_t>(a); })
        .def_property("tgt_kps", [](const ImagePairFlow& f) { return KpsToNumpy(f.tgt_kps); },
                      [](ImagePairFlow& f, const FArr& a) { f.tgt_kps = KpsFromNumpy(a); })
        .def_property("flow_errors",
                      [](const ImagePairFlow& f) {
                          FArr a((py::ssize_t)f.flow_errors.size());
                          memcpy(a.mutable_data(), f.flow_errors.data(), f.flow_errors.size() * 4);
                          return a; },
                      [](ImagePairFlow& f, const FArr& a) { f.flow_errors = VecFromNumpy<float>(a); });

    py::class_<Database>(m, "Database")
        .def(py::init<const std::string&>(), py::arg("path"))
        .def("open", &Database::Open, py::arg("path"))
        .def("close", &Database::Close)
        .def("read_keypoints", [](const Database& d, int32_t id) { return KpsToNumpy(d.ReadKeypoints(id)); },
             py::arg("image_id"))
        .def("write_keypoints", [](Database& d, int32_t id, const FArr& k) { d.WriteKeypoints(id, KpsFromNumpy(k)); },
             py::arg("image_id"), py::arg("keypoints"))
        .def("read_image_pair_flow",
             [](const Database& d, int32_t a, int32_t b) { return d.ReadImagePairFlow(a, b); },
             py::arg("image_id_from"), py::arg("image_id_to"))
        .def("write_image_pair_flow",
             [](Database& d, int32_t a, int32_t b, const UArr& idx, const FArr& tgt, const FArr& err) {
                 d.WriteImagePairFlow(a, b, VecFromNumpy<uint32_t>(idx), KpsFromNumpy(tgt), VecFromNumpy<float>(err));
             },
             py::arg("image_id_from"), py::arg("image_id_to"), py::arg("src_kps_indices"), py::arg("tgt_kps"),
             py::arg("flow_errors"))
        .def("write_image_pair_flow", [](Database& d, const ImagePairFlow& f) { d.WriteImagePairFlow(f); },
             py::arg("image_pair_flow"))
        .def("find_optical_flows_from_image", &Database::FindOpticalFlowsFromImage, py::arg("image_id_from"))
        .def("find_optical_flows_to_image", &Database::FindOpticalFlowsToImage, py::arg("image_id_to"))
        .def("keypoints_exist", &Database::KeypointsExist, py::arg("image_id"))
        .def("image_pair_flow_exists", &Database::ImagePairFlowExists, py::arg("image_id_from"), py::arg("image_id_to"))
        .def("get_min_image_id_with_keypoints", &Database::GetMinImageIdWithKeypoints)
        .def("get_max_image_id_with_keypoints", &Database::GetMaxImageIdWithKeypoints);

    py::class_<VideoInfo>(m, "VideoInfo")
        .def(py::init([](uint32_t w, uint32_t h, int32_t first, uint32_t n) { return VideoInfo{w, h, first, n}; }),
             py::arg("width"), py::arg("height"), py::arg("first_frame"), py::arg("num_frames"))
        .def_readwrite("width", &VideoInfo::width)
        .def_readwrite("height", &VideoInfo::height)
        .def_readwrite("first_frame", &VideoInfo::first_frame)
        .def_readwrite("num_frames", &VideoInfo::num_frames);

    py::class_<GFTTOptions>(m, "GFTTOptions")     // grid_rows/grid_cols are not exposed (polychase_pybind.cc:128-136)
        .def(py::init<>())
        .def_readwrite("quality_level", &GFTTOptions::quality_level)
        .def_readwrite("min_distance", &GFTTOptions::min_distance)
        .def_readwrite("block_size", &GFTTOptions::block_size)
        .def_readwrite("gradient_size", &GFTTOptions::gradient_size)
        .def_readwrite("max_corners", &GFTTOptions::max_corners)
        .def_readwrite("use_harris", &GFTTOptions::use_harris)
        .def_readwrite("harris_k", &GFTTOptions::harris_k);

    py::class_<OpticalFlowOptions>(m, "OpticalFlowOptions")
        .def(py::init<>())
        .def_readwrite("window_size", &OpticalFlowOptions::window_size)
        .def_readwrite("max_level", &OpticalFlowOptions::max_level)
        .def_readwrite("term_max_iters", &OpticalFlowOptions::term_max_iters)
        .def_readwrite("term_epsilon", &OpticalFlowOptions::term_epsilon)
        .def_readwrite("min_eigen_threshold", &OpticalFlowOptions::min_eigen_threshold);

    py::class_<Pose>(m, "Pose")
        .def(py::init<>())
        // q is exposed as WXYZ (polychase_pybind.cc:224-232); read or assign it as a whole
        .def_property("q", [](const Pose& p) { return ArrToNumpy(p.q); },
                      [](Pose& p, const FArr& q) { p.q = ArrFromNumpy<4>(q); })
        .def_property("t", [](const Pose& p) { return ArrToNumpy(p.t); },
                      [](Pose& p, const FArr& t) { p.t = ArrFromNumpy<3>(t); });

    py::class_<CameraState>(m, "CameraState")
        .def(py::init<>())
        .def(py::init([](CameraIntrinsics i, Pose p) { return CameraState{i, p}; }), py::arg("intrinsics"),
             py::arg("pose"))
        .def_readwrite("intrinsics", &CameraState::intrinsics)
        .def_readwrite("pose", &CameraState::pose);

    py::enum_<BundleOptions::LossType>(m, "LossType")
        .value("Trivial", BundleOptions::LossType::TRIVIAL)
        .value("Huber", BundleOptions::LossType::HUBER)
        .value("Cauchy", BundleOptions::LossType::CAUCHY);

    py::class_<BundleOptions>(m, "BundleOptions")
        .def(py::init<>())
        .def_readwrite("max_iterations", &BundleOptions::max_iterations)
        .def_readwrite("max_allowed_parallelism", &BundleOptions::max_allowed_parallelism)
        .def_readwrite("loss_type", &BundleOptions::loss_type)
        .def_readwrite("loss_scale", &BundleOptions::loss_scale)
        .def_readwrite("gradient_tol", &BundleOptions::gradient_tol)
        .def_readwrite("step_tol", &BundleOptions::step_tol)
        .def_readwrite("initial_lambda", &BundleOptions::initial_lambda)
        .def_readwrite("min_lambda", &BundleOptions::min_lambda)
        .def_readwrite("max_lambda", &BundleOptions::max_lambda)
        .def_readwrite("verbose", &BundleOptions::verbose);

    py::class_<BundleStats>(m, "BundleStats")
        .def(py::init<>())
        .def_readwrite("iterations", &BundleStats::iterations)
        .def_readwrite("initial_cost", &BundleStats::initial_cost)
        .def_readwrite("cost", &BundleStats::cost)
        .def_readwrite("lambda", &BundleStats::lambda)
        .def_readwrite("invalid_steps", &BundleStats::invalid_steps)
        .def_readwrite("step_norm", &BundleStats::step_norm)
        .def_readwrite("grad_norm", &BundleStats::grad_norm)
        .def("__repr__", [](const BundleStats& s) {
            char buf[256];
            snprintf(buf, sizeof(buf),
                     "BundleStats(iterations=%zu, initial_cost=%g, cost=%g, lambda=%g, invalid_steps=%zu, "
                     "step_norm=%g, grad_norm=%g)",
                     s.iterations, s.initial_cost, s.cost, s.lambda, s.invalid_steps, s.step_norm, s.grad_norm);
            return std::string(buf);
        });

    py::class_<PnPResult>(m, "PnPResult")
        .def_readwrite("camera", &PnPResult::camera)
        .def_readwrite("bundle_stats", &PnPResult::bundle_stats);

    py::class_<FrameTrackingResult>(m, "FrameTrackingResult")
        .def_readwrite("frame", &FrameTrackingResult::frame)
        .def_readwrite("pose", &FrameTrackingResult::pose)
        .def_readwrite("intrinsics", &FrameTrackingResult::intrinsics)
        .def_readwrite("bundle_stats", &FrameTrackingResult::bundle_stats)
        .def_readwrite("inlier_ratio", &FrameTrackingResult::inlier_ratio);

    py::class_<CameraTrajectory, std::shared_ptr<CameraTrajectory>>(m, "CameraTrajectory")
        .def(py::init<int32_t, size_t>(), py::arg("first_frame_id"), py::arg("count"))
        .def("is_valid_frame", &CameraTrajectory::IsValidFrame, py::arg("frame_id"))
        .def("is_frame_filled", &CameraTrajectory::IsFrameFilled, py::arg("frame_id"))
        .def("get", &CameraTrajectory::Get, py::arg("frame_id"))
        .def("set", &CameraTrajectory::Set, py::arg("frame_id"), py::arg("state"))
        .def("count", &CameraTrajectory::Count)
        .def("first_frame", &CameraTrajectory::FirstFrame)
        .def("last_frame", &CameraTrajectory::LastFrame);

    py::class_<RefineTrajectoryUpdate>(m, "RefineTrajectoryUpdate")
        .def_readwrite("progress", &RefineTrajectoryUpdate::progress)
        .def_readwrite("message", &RefineTrajectoryUpdate::message)
        .def_readwrite("stats", &RefineTrajectoryUpdate::stats);

    py::class_<CppException>(m, "CppException").def("what", [](const CppException& e) { return e.message; });

    py::class_<OpticalFlowProgress>(m, "OpticalFlowProgress")
        .def_readonly("progress", &OpticalFlowProgress::progress)
        .def_readonly("progress_message", &OpticalFlowProgress::progress_message);
    py::class_<OpticalFlowRequest>(m, "OpticalFlowRequest").def_readonly("frame_id", &OpticalFlowRequest::frame_id);

    py::class_<OpticalFlowThread>(m, "OpticalFlowThread")
        .def(py::init<VideoInfo, std::string, GFTTOptions, OpticalFlowOptions, bool>(), py::arg("video_info"),
             py::arg("database_path"), py::arg("detector_options") = GFTTOptions{},
             py::arg("OpticalFlowOptions") = OpticalFlowOptions{}, py::arg("write_images") = false)
        .def("request_stop", &OpticalFlowThread::RequestStop)
        .def("join", &OpticalFlowThread::Join, py::call_guard<py::gil_scoped_release>())
        .def("try_pop", [](OpticalFlowThread& t) { return VariantToPy(t.TryPop()); })
        .def("empty", &OpticalFlowThread::Empty)
        .def("provide_frame", [](OpticalFlowThread& t, int32_t frame_id, const U8Arr& frame) {
            const Frame f = FrameFromNumpy(frame);
            t.ProvideFrame(frame_id, f.data, f.width, f.height, f.stride);
        });

    py::class_<TrackerThread>(m, "TrackerThread")
        .def(py::init<std::string, int32_t, int32_t, SceneTransformations, std::shared_ptr<const AcceleratedMesh>, bool,
                      bool, BundleOptions>(),
             py::arg("database_path"), py::arg("frame_from"), py::arg("frame_to_inclusive"), py::arg("scene_transform"),
             py::arg("accel_mesh"), py::arg("optimize_focal_length"), py::arg("optimize_principal_point"),
             py::arg("bundle_opts"))
        .def("request_stop", &TrackerThread::RequestStop)
        .def("join", &TrackerThread::Join, py::call_guard<py::gil_scoped_release>())
        .def("try_pop", [](TrackerThread& t) { return VariantToPy(t.TryPop()); })
        .def("empty", &TrackerThread::Empty);

    py::class_<RefinerThread>(m, "RefinerThread")
        .def(py::init([](std::string path, std::shared_ptr<CameraTrajectory> traj, const FArr& model,
                         std::shared_ptr<const AcceleratedMesh> mesh, bool of, bool opp, BundleOptions bo) {
                 return std::make_unique<RefinerThread>(std::move(path), std::move(traj), Mat4FromNumpy(model),
                                                        std::move(mesh), of, opp, bo);
             }),
             py::arg("database_path"), py::arg("camera_trajectory"), py::arg("model_matrix"), py::arg("mesh"),
             py::arg("optimize_focal_length"), py::arg("optimize_principal_point"), py::arg("bundle_opts"))
        .def("request_stop", &RefinerThread::RequestStop)
        .def("join", &RefinerThread::Join, py::call_guard<py::gil_scoped_release>())
        .def("try_pop", [](RefinerThread& t) { return VariantToPy(t.TryPop()); })
        .def("empty", &RefinerThread::Empty);

    m.def("ray_cast",
          [](const AcceleratedMesh& mesh, const SceneTransformations& scene, const FArr& pos, bool check_mask) {
              const auto p = ArrFromNumpy<2>(pos);
              py::gil_scoped_release release;      // the device round trip runs without the GIL
              return mesh.RayCast(scene, p, check_mask);
          },
          py::arg("accel_mesh"), py::arg("scene_transform"), py::arg("pos"), py::arg("check_mask"));

    m.def("find_transformation",                                            // polychase_pybind.cc:319-325
          [](const FArr& object_points, const SceneTransformations& initial, const SceneTransformations& current,
             const PinUpdate& update, TransformationType trans_type, bool of, bool opp) {
              if (object_points.ndim() != 2 || object_points.shape(1) != 3)
                  throw std::invalid_argument("object_points: expected an (N, 3) float32 array");
              std::vector<float> pts((size_t)object_points.size());
              if (!pts.empty()) memcpy(pts.data(), object_points.data(), pts.size() * sizeof(float));
              py::gil_scoped_release release;
              return FindTransformation(pts, initial, current, update, trans_type, of, opp);
          },
          py::arg("object_points"), py::arg("initial_scene_transform"), py::arg("current_scene_transform"),
          py::arg("update"), py::arg("trans_type"), py::arg("optimize_focal_length") = false,
          py::arg("optimize_principal_point") = false);

    m.def("generate_optical_flow_database",
          [](const VideoInfo& vi, py::function accessor, py::object callback, const std::string& path,
             const GFTTOptions& go, const OpticalFlowOptions& fo, bool write_images) {
              // Python callables are invoked with the GIL re-acquired; the pipeline itself runs without it
              // Python callables live behind a shared_ptr whose deleter takes the GIL: copying or destroying the
              // std::function that captures it never touches a Python reference count without the GIL
              auto accessor_h = HoldPy(std::move(accessor));
              FrameAccessorFunction acc = [accessor_h](int32_t frame_id) -> std::optional<Frame> {
                  py::gil_scoped_acquire gil;
                  py::object r = (*accessor_h)(frame_id);
                  if (r.is_none()) return std::nullopt;
                  auto arr = std::make_shared<U8Arr>(r.cast<U8Arr>());
                  Frame f = FrameFromNumpy(*arr);
                  // copy: the array must not be destroyed without the GIL
                  auto buf = std::shared_ptr<uint8_t[]>(new uint8_t[(size_t)f.height * f.width * 3]);
                  for (int y = 0; y < f.height; y++)
                      memcpy(buf.get() + (size_t)y * f.width * 3, f.data + (size_t)y * f.stride, (size_t)f.width * 3);
                  arr.reset();
                  f.data = buf.get();
                  f.stride = (size_t)f.width * 3;
                  f.keep_alive = buf;
                  return f;
              };
              OpticalFlowProgressCallback cb;
              if (!callback.is_none()) {
                  auto fn = HoldPy(callback.cast<py::function>());
                  cb = [fn](float p, const std::string& msg) {
                      py::gil_scoped_acquire gil;
                      return (*fn)(p, msg).cast<bool>();
                  };
              }
              py::gil_scoped_release release;
              GenerateOpticalFlowDatabase(vi, acc, cb, path, go, fo, write_images);
          },
          py::arg("video_info"), py::arg("frame_accessor_function"), py::arg("callback"), py::arg("database_path"),
          py::arg("detector_options") = GFTTOptions{}, py::arg("flow_options") = OpticalFlowOptions{},
          py::arg("write_images") = false);

    m.def("track_sequence",
          [](const std::string& path, int32_t from, int32_t to, const SceneTransformations& scene,
             const AcceleratedMesh& mesh, py::object callback, bool of, bool opp, BundleOptions bo) {
              TrackingCallback cb;
              if (!callback.is_none()) {
                  auto fn = HoldPy(callback.cast<py::function>());
                  cb = [fn](const FrameTrackingResult& r) {
                      py::gil_scoped_acquire gil;
                      return (*fn)(r).cast<bool>();
                  };
              }
              py::gil_scoped_release release;
              TrackSequence(path, from, to, scene, mesh, cb, of, opp, bo);
          },
          py::arg("database_path"), py::arg("frame_from"), py::arg("frame_to_inclusive"), py::arg("scene_transform"),
          py::arg("accel_mesh"), py::arg("callback"), py::arg("optimize_focal_length") = false,
          py::arg("optimize_principal_point") = false, py::arg("bundle_opts") = BundleOptions());

    m.def("refine_trajectory",
          [](const std::string& path, CameraTrajectory& traj, const FArr& model, const AcceleratedMesh& mesh, bool of,
             bool opp, py::object callback, BundleOptions bo) {
              RefineTrajectoryCallback cb;
              if (!callback.is_none()) {
                  auto fn = HoldPy(callback.cast<py::function>());
                  cb = [fn](RefineTrajectoryUpdate u) {
                      py::gil_scoped_acquire gil;
                      return (*fn)(u).cast<bool>();
                  };
              }
              const Mat4 mm = Mat4FromNumpy(model);
              py::gil_scoped_release release;
              RefineTrajectory(path, traj, mm, mesh, of, opp, cb, bo);
          },
          py::arg("database_path"), py::arg("camera_trajectory"), py::arg("model_matrix"), py::arg("mesh"),
          py::arg("optimize_focal_length"), py::arg("optimize_principal_point"), py::arg("callback"),
          py::arg("bundle_opts") = BundleOptions());
}
