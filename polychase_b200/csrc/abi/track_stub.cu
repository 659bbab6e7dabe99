// TEMPORARY: track / refine entry points land in track.cu / ba.cu.
#include "context.h"
namespace pc {
struct MeshData {};
struct BAData {};
void free_mesh(MeshData* m) { delete m; }
void free_ba(BAData* b) { delete b; }
}
using namespace pc;
extern "C" {
int pc_mesh_set(pc_ctx* c, const float*, int, const uint32_t*, int, const uint32_t*, int) { return fail(c, PC_ERR_STATE, "not built yet"); }
int pc_ray_cast(pc_ctx* c, const float*, const pc_camera_state*, const float*, int, int, uint8_t*, float*, uint32_t*, float*, float*) { return fail(c, PC_ERR_STATE, "not built yet"); }
int pc_solve_pnp(pc_ctx* c, const float*, const float*, const float*, int, const pc_bundle_opts*, float, int, int, pc_camera_state*, pc_bundle_stats*, float*) { return fail(c, PC_ERR_STATE, "not built yet"); }
int pc_track_frame(pc_ctx* c, const pc_match_source*, int, const float*, const pc_camera_state*, const pc_bundle_opts*, int, int, pc_camera_state*, pc_bundle_stats*, float*, int*) { return fail(c, PC_ERR_STATE, "not built yet"); }
int pc_ba_load(pc_ctx* c, const pc_ba_problem*) { return fail(c, PC_ERR_STATE, "not built yet"); }
int pc_ba_cost(pc_ctx* c, const pc_camera_state*, const pc_bundle_opts*, float*) { return fail(c, PC_ERR_STATE, "not built yet"); }
int pc_ba_normal_equations(pc_ctx* c, const pc_camera_state*, const pc_bundle_opts*, float*, float*) { return fail(c, PC_ERR_STATE, "not built yet"); }
int pc_ba_solve(pc_ctx* c, const pc_bundle_opts*, pc_camera_state*, pc_bundle_stats*, pc_ba_iter_cb, void*) { return fail(c, PC_ERR_STATE, "not built yet"); }
}
