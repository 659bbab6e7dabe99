// SQLite store of keypoints and optical flow: same tables, pragmas and little-endian blobs as
// /root/reference/cpp/database.{h,cc} (the on-disk contract between Analyze and Track/Refine:
// keypoints(image_id, rows, keypoints BLOB rows x 2 f32) and optical_flow(image_id_from,
// image_id_to, rows, src_keypoints_indices BLOB u32, tgt_keypoints BLOB rows x 2 f32,
// flow_errors BLOB f32), database.cc:108-135).
#pragma once

#include <string>
#include <vector>

#include "sqlite_shim.h"
#include "types.h"

namespace pch {

struct ImagePairFlow {
    int32_t image_id_from = 0;
    int32_t image_id_to = 0;
    KeypointsIndices src_kps_indices;
    Keypoints tgt_kps;
    FlowErrors flow_errors;
    void Clear() { src_kps_indices.clear(); tgt_kps.clear(); flow_errors.clear(); }
};

class Database {
   public:
    explicit Database(const std::string& path);
    Database(const Database&) = delete;
    Database& operator=(const Database&) = delete;
    ~Database();

    void Open(const std::string& path);
    void Close();

    Keypoints ReadKeypoints(int32_t image_id) const;
    void ReadKeypoints(int32_t image_id, Keypoints& keypoints) const;
    void WriteKeypoints(int32_t image_id, const Keypoints& keypoints);
    void WriteKeypoints(int32_t image_id, const float* xy, size_t rows);
    ImagePairFlow ReadImagePairFlow(int32_t image_id_from, int32_t image_id_to) const;
    void ReadImagePairFlow(int32_t image_id_from, int32_t image_id_to, ImagePairFlow& flow) const;
    void WriteImagePairFlow(int32_t image_id_from, int32_t image_id_to, const KeypointsIndices& src_kps_indices,
                            const Keypoints& tgt_kps, const FlowErrors& flow_errors);
    void WriteImagePairFlow(int32_t image_id_from, int32_t image_id_to, const uint32_t* idx, const float* tgt_xy,
                            const float* err, size_t rows);
    void WriteImagePairFlow(const ImagePairFlow& flow);
    std::vector<int32_t> FindOpticalFlowsFromImage(int32_t image_id_from) const;
    std::vector<int32_t> FindOpticalFlowsToImage(int32_t image_id_to) const;
    bool KeypointsExist(int32_t image_id) const;
    bool ImagePairFlowExists(int32_t image_id_from, int32_t image_id_to) const;
    int32_t GetMinImageIdWithKeypoints() const;
    int32_t GetMaxImageIdWithKeypoints() const;
    // One transaction around a batch of inserts (the reference autocommits every row).
    void Begin();
    void Commit();

   private:
    void Exec(const char* sql) const;
    int Call(int rc, int line) const;
    sqlite3_stmt* Prepare(const char* sql);

    sqlite3* db_ = nullptr;
    sqlite3_stmt* read_kps_ = nullptr;
    sqlite3_stmt* write_kps_ = nullptr;
    sqlite3_stmt* read_flow_ = nullptr;
    sqlite3_stmt* write_flow_ = nullptr;
    sqlite3_stmt* flows_from_ = nullptr;
    sqlite3_stmt* flows_to_ = nullptr;
    sqlite3_stmt* kps_exist_ = nullptr;
    sqlite3_stmt* flow_exist_ = nullptr;
    sqlite3_stmt* min_id_ = nullptr;
    sqlite3_stmt* max_id_ = nullptr;
};

}  // namespace pch
