// Internal launcher interface between the C ABI (csrc/abi) and the sm_100a kernels.
// Everything here takes raw device pointers and a stream; no ownership.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace pc {

constexpr int kMaxLevels = 6;

// Pyramid levels of the frame ring carry a REFLECT_101 apron of kPadX columns / kPadY rows around
// the image (what cv::buildOpticalFlowPyramid's winSize padding gives OpenCV's LK): the 10x10 LK
// kernel reads windows that hang over the image edge without any border arithmetic.
constexpr int kPadX = 16, kPadY = 12;

struct Image8 {          // one pitched u8 plane in HBM
    uint8_t* data = nullptr;   // pixel (0,0)
    int w = 0, h = 0;
    int pitch = 0;       // bytes per row, multiple of 128
};

struct PyramidView {     // what the LK kernel sees of one frame
    const uint8_t* data[kMaxLevels];
    int w[kMaxLevels], h[kMaxLevels], pitch[kMaxLevels];
    int levels;          // number of levels present (maxLevel + 1)
};

// ---- K1/K2: RGB->gray and pyrDown (gray_pyr.cu) --------------------------------------
void launch_rgb_to_gray(const uint8_t* rgb, size_t stride, Image8 gray, cudaStream_t s);
void launch_copy_gray(const uint8_t* src, size_t stride, Image8 gray, cudaStream_t s);
void launch_pyr_down(Image8 src, Image8 dst, cudaStream_t s);
// Fused TMA path (pyramid_tma.cu): gray + levels 1..3 from an RGB8 device image in two launches.  Returns how many
// levels of `lv` it wrote (0 = not applicable: nothing launched, use the kernels above); *launches = kernels launched.
int launch_pyramid_tma(const uint8_t* rgb, size_t stride, const Image8* lv, int levels, cudaStream_t s, int* launches);
// Fills the apron (kPadX / kPadY, BORDER_REFLECT_101) of `levels` planes in one launch.
void launch_pad_border(const Image8* planes, int levels, cudaStream_t s);

// ---- K4/K5: min-eigenvalue map, per-cell max, threshold + NMS (mineig.cu) -------------
struct DetectGrid {
    int grid_rows, grid_cols, block_w, block_h;   // gftt.cc:39-43
};
// eig: w*h floats (pitch in floats = eig_pitch); cell_max: grid_rows*grid_cols ordered ints
// Also clears the detector's counter block (`zero_block`, `zero_ints` ints: candidate count, value
// histograms, greedy round counters, select scratch) and the frame's three counters
// (`frame_counters`: n_kps, n_accepted, greedy_remaining) in its init launch: K5..K7 rely on that.
void launch_min_eig(Image8 gray, float* eig, int eig_pitch, DetectGrid g, int* cell_max, int* zero_block,
                    int zero_ints, int* frame_counters, cudaStream_t s);
// candidates: 64-bit keys (ordered value << 32 | address); state: u8 map (pitch = gray.pitch)
// value_hist: 4096 bins over the top 12 bits of the ordered candidate value; it and cand_count are
// zero on entry (launch_min_eig)
void launch_nms_candidates(const float* eig, int eig_pitch, int w, int h, DetectGrid g,
                           const int* cell_max, double quality_level, uint8_t* state, int state_pitch,
                           unsigned long long* cand, int cand_cap, int* cand_count, int* value_hist,
                           cudaStream_t s);

// ---- K6/K7: greedy min-distance suppression + ordering (select.cu) --------------------
struct SelectWorkspace {
    unsigned long long* accepted;      // cap entries
    unsigned long long* sorted;        // sorted_cap >= cap entries (CUB sort output, unlimited path)
    unsigned long long* strong;        // cap entries: the strongest candidates (max_corners > 0 path)
    int* accepted_count;               // device int
    int* round_counters;               // 3 * kMaxGreedyRounds ints (one set per fixed-point run of a launch)
    int* remaining;                    // device int: undecided candidates left (0 = converged)
    int* hist;                         // 4096 ints: 12-bit value histogram of all candidates (written by K5)
    int* kept_hist;                    // 4096 ints: 12-bit value histogram of kept keys
    int* sel;                          // 8 ints: [1] strong count, [2] threshold bin, [3] short-list size
    int* bin_cursor;                   // 4096 ints, zero on entry: fill cursors of the short list's bin groups
    int* bin_start;                    // 4096 ints: first slot of each bin group in the short list
    void* cub_temp; size_t cub_temp_bytes;
    int cap;
    int sorted_cap;
};
constexpr int kMaxGreedyRounds = 1024;   // per fixed-point run; a round resolves ~100 dependency levels
size_t select_cub_temp_bytes(int cap);
// Runs suppression over `cand` (count on device), sorts accepted keys descending and writes
// min(accepted, max_corners) keypoints (x,y floats) + their count.
void launch_select(const unsigned long long* cand, const int* cand_count, int cand_cap,
                   const float* eig, int eig_pitch, uint8_t* state, int state_pitch, int w, int h,
                   double min_distance, int max_corners, SelectWorkspace ws, float* kps_out,
                   int kps_cap, int* kps_count, int sm_count, cudaStream_t s);

// ---- K8/K9: pyramidal LK + status compaction (lk.cu) ---------------------------------
struct LKParams {
    int win, iters, max_level;
    double eps;          // already clamped; squared inside
    double min_eig;
};
// Source templates of one frame's keypoints, precomputed once per frame for the 10x10 kernel
// (lk10.cu): per (level, keypoint) the five lanes' 20 pixels of Ival / Ix / Iy packed in 32 words per
// lane, and A11, A12, A22.  words == nullptr: the LK kernel computes the templates itself.
struct LKTemplates {
    const uint4* words;          // [level][cap][5 lanes][8 uint4]; queue layout: [level][cap][16 quads][5 lanes]
    const float* sums;           // [level][cap][4] (nullptr in the queue layout: the sums live in quad 15)
    int cap;
};
constexpr size_t kLkTemplateBytesPerPoint = 5 * 8 * 16;     // per level
// queue layout (lk10q.cu): Ival / Ix / Iy unpacked (15 quads per lane) + one quad of sums
constexpr int kLkQueueQuadsPerLane = 16;
constexpr size_t kLkQueueTemplateBytesPerPoint = (size_t)kLkQueueQuadsPerLane * 5 * 16;
struct LKPair {
    PyramidView a, b;            // source / target frame
    LKTemplates tmpl;
    const float* pts;            // source keypoints (x,y)
    const int* n_pts;            // device count
    const int* order;            // optional: the order in which lk10_kernel walks the source keypoints (a permutation
                                 // of 0..n-1, spatially sorted: launch_spatial_order); nullptr = index order
    float* next;                 // dense outputs, capacity `cap`
    uint8_t* status;
    float* err;
    // compacted outputs
    uint32_t* out_idx; float* out_tgt; float* out_err; int* out_count;
};
constexpr int kMaxPairsPerLaunch = 8;
struct LKBatch {
    LKPair pair[kMaxPairsPerLaunch];
    int num_pairs;
    int cap;                     // max points per pair
    // lk10q.cu: non-null = every pair brings queue-layout templates and the launch runs as a work queue
    // over this device counter (the launcher zeroes it); queue_budget > 0 = items a block takes before it
    // leaves (0: blocks stay until the queue is empty)
    int* queue;
    int queue_budget;
};
void launch_lk(const LKBatch& batch, const LKParams& p, cudaStream_t s);
void launch_lk_compact(const LKBatch& batch, cudaStream_t s);
// Templates of frame `a`'s keypoints for every level the 10x10 kernel will visit (win must be 10).
void launch_lk10_templates(const PyramidView& a, const float* pts, const int* n_pts, int cap, const LKParams& p,
                           uint4* words, float* sums, cudaStream_t s, const int* order = nullptr);
// order[0..n) = the keypoints' indices sorted by 64 x 64-pixel cell (row-major cells; the order inside a cell is
// arbitrary).  The LK kernels use it only to decide which keypoints share a warp / a block -- results are written by
// keypoint index, so it cannot change them: neighbours in the image read the same cache lines and, for the +-8 pairs,
// tend to need similar iteration counts.  One block; scratch = 2 * kSpatialCellsMax ints are taken from shared memory.
void launch_spatial_order(const float* pts, const int* n_pts, int cap, int w, int h, int* order, cudaStream_t s);
// The same in the queue layout (kLkQueueTemplateBytesPerPoint per level and point).
void launch_lk10q_templates(const PyramidView& a, const float* pts, const int* n_pts, int cap, const LKParams& p,
                            uint4* words, cudaStream_t s);

// ---- synthetic frame warp (synth.cu) ------------------------------------------------
void launch_synth_warp(const uint8_t* tex, int w, int h, int tex_pitch, const double Hinv[9],
                       uint8_t* rgb, size_t stride, cudaStream_t s);

}  // namespace pc
