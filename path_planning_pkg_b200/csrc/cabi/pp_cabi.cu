// C ABI of libpp_b200.so (declared in include/pp_b200.h): context management, host prologues
// (csrc/host/pp_host.h) and kernel launches (csrc/kernels/pp_kernels.cuh).  No CPU compute path:
// every entry point that computes launches a kernel on the context's stream.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "../kernels/pp_kernels.cuh"
#include "../host/pp_host.h"
#include "../host/pp_footprint_host.h"
#include "../kernels/pp_fields.h"
#include "../../../include/pp_b200.h"

static thread_local std::string g_last_error;

struct PPNcclIdByValue { char internal[128]; };     // ncclUniqueId (nccl.h), passed by value to ncclCommInitRank

static int pp_fail(int code, const std::string& msg)
{
    g_last_error = msg;
    return code;
}

#define PP_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return pp_fail(PP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));           \
    } while (0)

template <class T>
struct DevBuf
{
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Per-slot scratch of the EXACT search kernel: small FIXED pools every resident query starts on, plus one arena
// (csrc/core/pp_arena.h) the containers of the queries that need more grow into.
struct WorkPools
{
    int alloc_slots = 0, chash_cap = 0, open3_cap = 0, closed_cap = 0, open2_cap = 0;
    DevBuf<PPNode3> open3; DevBuf<PPClosed3> closed; DevBuf<PPHashSlot> chash;
    DevBuf<unsigned> cell_state; DevBuf<float> nm_g, nm_f, cl_g; DevBuf<int> cl_prev; DevBuf<PPNode2> open2;
    DevBuf<unsigned char> arena_mem; DevBuf<PPArena> arena_ctl; size_t arena_bytes = 0;
    bool clamped = false;      // the slot count was limited by the memory budget, not by the request
    void release()
    {
        open3.release(); closed.release(); chash.release(); cell_state.release();
        nm_g.release(); nm_f.release(); cl_g.release(); cl_prev.release(); open2.release();
        arena_mem.release(); arena_ctl.release(); arena_bytes = 0;
        alloc_slots = 0;
    }
};

// first-pass capacities of the fixed per-slot pools (containers grow from here, x2 per step)
#define PP_INIT_CLOSED 8192
#define PP_INIT_OPEN3  4096
#define PP_INIT_OPEN2D 2048

struct pp_context
{
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    pp_params params;
    PPHostModel model;
    int num_groups = 0;
    int sm_count = 0;
    std::vector<PPHostFrame> frames;
    std::vector<PPGroup> groups;          // host mirror of d_groups
    std::vector<int> apf_cap;             // allocated obstacles per group
    std::vector<float*> d_apf;            // per group
    std::vector<DevBuf<int>> d_bin_off, d_bin_idx;   // per group: spatial index of the APF list
    float* d_maps = nullptr;              // num_groups x N*N
    float* d_map_tmp = nullptr;           // N*N (relocation double buffer)
    int*   d_cell_scratch = nullptr;      // N*N ints (relocation indices / line sample counts), kept zeroed
    float* d_off_xy = nullptr;
    PPGroup* d_groups = nullptr;
    bool groups_dirty = true;
    // generic scratch for stateless calls and map descriptors
    DevBuf<unsigned char> s0, s1, s2, s3, s4;
    // box descriptors of the map update: pinned host ring + device ring, so that an update is H2D + one launch with no
    // synchronisation (a slot is reused only after the launch that read it has finished)
    struct BoxStage { PPBoxDescDev* h = nullptr; PPBoxDescDev* d = nullptr; int cap = 0; unsigned next = 0;
                      cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr}; bool used[4] = {false, false, false, false}; } box_stage;
    // batch state
    int n_queries = 0;
    pp_search_opts opts;
    std::vector<PPQuery> h_queries;
    DevBuf<PPQuery> d_queries;
    DevBuf<PPResult> d_results;
    DevBuf<PPPathPt> d_paths;
    DevBuf<PPPop> d_trace;
    int* d_counter = nullptr;
    // per-slot scratch pools
    int n_slots = 0;
    WorkPools wp;          // EXACT mode: small fixed pools per resident query + the arena they grow into
    WorkPools wp_lazy;     // persistent cache of the stand-alone lazy 2D A* (AStar<T> handle), fixed pools
    size_t mem_budget = 0; // bytes the search scratch (slots + arena) may take; 0 = 75 % of the free device memory
    pp_context* parent = nullptr;   // lane context (pp_create_lane): shares the parent's maps / groups / tables, read-only
    cudaEvent_t ev_run0 = nullptr, ev_run1 = nullptr;   // bracket of the asynchronous first pass of a batch
    cudaEvent_t ev_re0 = nullptr, ev_re1 = nullptr;     // bracket of a re-run pass (ev0 / ev1 belong to pp_timer_begin / _end)
    bool run_pending = false;
    unsigned long long map_epoch = 1;     // bumped by every change of a map, goal frame or APF list (what lanes mirror)
    unsigned long long seen_epoch = 0;    // lane: the parent's epoch at the last refresh
    void* nccl_comm = nullptr;      // ncclComm_t of pp_comm_init (map replication, pp_broadcast_maps)
    int comm_rank = -1, comm_size = 0;
    unsigned* d_lazy_sid = nullptr;
    int lazy_group = -1;
    // planner-object history (pp_set_history): per group the carried 2D cache (`_visted` + `_node_map` costs) of one
    // reference planner object, plus the snapshot a capacity retry restores
    struct HistBuf { bool on = false; DevBuf<unsigned> cell_state, cell_state_bak, sid; DevBuf<float> nm_g, nm_f, nm_g_bak, nm_f_bak;
                     void release() { cell_state.release(); cell_state_bak.release(); sid.release(); nm_g.release(); nm_f.release();
                                      nm_g_bak.release(); nm_f_bak.release(); on = false; } };
    std::vector<HistBuf> hist;
    int batch_hist_group = -1;    // group whose history the uploaded batch runs on (-1 = fresh cache per query)
    // generic footprint collision check (pp_set_footprint): per-bin offset table on the host and the device
    std::vector<PPFootBin> foot_bins; std::vector<PPCellOff> foot_offs; int foot_win = 0;
    DevBuf<PPFootBin> d_foot_bins; DevBuf<PPCellOff> d_foot_offs; DevBuf<int> d_foot_lin;
    DevBuf<float> d_foot_xyh; DevBuf<int> d_foot_out;
    // velocity profile / trajectory (pp_velocity_profile_batch, pp_trajectory_batch)
    std::vector<float> h_vel;             // vel_init of the uploaded queries
    bool batch_done = false;              // pp_batch_run finished on the uploaded batch (results / path records valid)
    DevBuf<float> d_traj, d_traj_tmp, d_vel_in; DevBuf<int> d_traj_int; DevBuf<PPWorldFrame> d_wframes;
    // K-POP mode pools (per slot): node log, hash table, LSM queue arena + merge scratch
    struct KPools { int alloc_slots = 0, nodes_cap = 0, table_cap = 0, levels = 0; size_t arena_cap = 0, tmp_cap = 0; bool clamped = false;
                    DevBuf<PPKNode> nodes; DevBuf<PPKSlot> table; DevBuf<PPKEntry> arena, tmp_a, tmp_b;
                    void release() { nodes.release(); table.release(); arena.release(); tmp_a.release(); tmp_b.release(); alloc_slots = 0; } };
    KPools kp, kp_retry;
    // heuristic fields (throughput modes / C3)
    DevBuf<float> d_field2d;              // num_groups x N*N, filled per group on demand
    std::vector<char> field2d_valid;
    // K-POP scheduling hint: mean iterations per query the group needed in its last batch (0 = unknown; reset by any map /
    // goal change).  Only the order in which resident slots fetch the work items depends on it, never a result.
    std::vector<float> kpop_group_cost;
    DevBuf<float> d_group_cost;
    bool group_cost_dirty = true;
    DevBuf<unsigned> d_f2d_work; DevBuf<unsigned char> d_f2d_flags; DevBuf<float> d_dubins_field; DevBuf<int> d_f2d_ctl;
    DevBuf<int> d_qmap, d_order;
    DevBuf<int> d_exact_order; bool exact_order_valid = false;    // fetch order of the uploaded EXACT batch (longest expected first)
    int retried = 0;       // queries re-run in the last pp_batch_run
    unsigned long long launches = 0;      // kernels launched by this context
};

static size_t nn_of(const pp_context* c) { return (size_t)c->model.C.N * c->model.C.N; }

static int sync_groups(pp_context* c)
{
    if (!c->groups_dirty) return PP_SUCCESS;
    PP_CUDA(cudaMemcpyAsync(c->d_groups, c->groups.data(), sizeof(PPGroup) * c->num_groups, cudaMemcpyHostToDevice, c->stream));
    c->groups_dirty = false;
    return PP_SUCCESS;
}

static void map_changed(pp_context* c, int g)
{
    c->map_epoch++;
    if (g >= 0 && g < (int)c->field2d_valid.size()) c->field2d_valid[g] = 0;
    if (g >= 0 && g < (int)c->kpop_group_cost.size() && c->kpop_group_cost[g] != 0.0f) { c->kpop_group_cost[g] = 0.0f; c->group_cost_dirty = true; }
}

static int check_group(pp_context* c, int g)
{
    if (!c) return pp_fail(PP_ERR_INVALID, "null context");
    if (g < 0 || g >= c->num_groups) return pp_fail(PP_ERR_INVALID, "group index out of range");
    return PP_SUCCESS;
}

// a lane context reads its parent's map state: take over the host mirrors and make sure the parent's device copies are current
static int lane_refresh(pp_context* c)
{
    pp_context* p = c->parent;
    if (!p) return PP_SUCCESS;
    // nothing changed since the last look: do not wait for whatever the parent's stream is busy with (its own batch)
    if (c->seen_epoch == p->map_epoch && !p->groups_dirty) return PP_SUCCESS;
    c->seen_epoch = p->map_epoch;
    c->field2d_valid.assign(c->field2d_valid.size(), 0);      // fields derived from the parent's old maps
    int rc = sync_groups(p); if (rc) return rc;
    PP_CUDA(cudaStreamSynchronize(p->stream));
    c->frames = p->frames; c->groups = p->groups;
    c->d_groups = p->d_groups; c->d_maps = p->d_maps; c->d_off_xy = p->d_off_xy;
    c->groups_dirty = false;
    return PP_SUCCESS;
}

// ---- NCCL, resolved at run time (the library itself links only the CUDA runtime).  dlopen by soname returns the copy the
// process already holds (e.g. the one torch.distributed loaded); PP_B200_NCCL_LIB names another file.
struct PPNccl
{
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, PPNcclIdByValue, int) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

static PPNccl* nccl_api(std::string& err)
{
    static PPNccl api;
    if (api.handle) return &api;
    const char* names[] = {std::getenv("PP_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) { if (nm && *nm) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; } }
    if (!h) { err = "NCCL library not found (libnccl.so.2; set PP_B200_NCCL_LIB)"; return nullptr; }
    api.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(void**, int, PPNcclIdByValue, int))dlsym(h, "ncclCommInitRank");
    api.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclBroadcast");
    api.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.Broadcast || !api.CommDestroy) { err = "NCCL symbols missing"; return nullptr; }
    api.handle = h;
    return &api;
}

static int lane_guard(pp_context* c, const char* what)
{
    if (c && c->parent) return pp_fail(PP_ERR_INVALID, std::string(what) + ": lanes are read-only views of their parent's maps");
    return PP_SUCCESS;
}

extern "C"
{

const char* pp_last_error(void) { return g_last_error.c_str(); }

int pp_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

static int create_impl(pp_context* c, int num_groups)
{
    PP_CUDA(cudaSetDevice(c->device));
    cudaDeviceProp prop;
    PP_CUDA(cudaGetDeviceProperties(&prop, c->device));
    c->sm_count = prop.multiProcessorCount;
    PP_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    PP_CUDA(cudaEventCreate(&c->ev0));
    PP_CUDA(cudaEventCreate(&c->ev1));
    PP_CUDA(cudaEventCreate(&c->ev_run0));
    PP_CUDA(cudaEventCreate(&c->ev_run1));
    PP_CUDA(cudaEventCreate(&c->ev_re0));
    PP_CUDA(cudaEventCreate(&c->ev_re1));
    c->num_groups = num_groups;
    PP_CUDA(cudaMalloc(&c->d_counter, sizeof(int)));
    PP_CUDA(cudaMalloc(&c->d_lazy_sid, sizeof(unsigned)));
    PP_CUDA(cudaMemsetAsync(c->d_lazy_sid, 0, sizeof(unsigned), c->stream));
    if (c->parent) return lane_refresh(c);
    size_t nn = nn_of(c);
    PP_CUDA(cudaMalloc(&c->d_maps, sizeof(float) * nn * num_groups));
    PP_CUDA(cudaMemsetAsync(c->d_maps, 0, sizeof(float) * nn * num_groups, c->stream));
    PP_CUDA(cudaMalloc(&c->d_map_tmp, sizeof(float) * nn));
    PP_CUDA(cudaMalloc(&c->d_cell_scratch, sizeof(int) * nn));
    PP_CUDA(cudaMemsetAsync(c->d_cell_scratch, 0, sizeof(int) * nn, c->stream));
    PP_CUDA(cudaMalloc(&c->d_off_xy, sizeof(float) * c->model.off_xy.size()));
    PP_CUDA(cudaMemcpyAsync(c->d_off_xy, c->model.off_xy.data(), sizeof(float) * c->model.off_xy.size(), cudaMemcpyHostToDevice, c->stream));
    PP_CUDA(cudaMalloc(&c->d_groups, sizeof(PPGroup) * num_groups));
    c->frames.resize(num_groups);
    c->groups.resize(num_groups);
    c->apf_cap.assign(num_groups, 0);
    c->d_apf.assign(num_groups, nullptr);
    c->d_bin_off.resize(num_groups); c->d_bin_idx.resize(num_groups);
    float z[3] = {0, 0, 0};
    for (int g = 0; g < num_groups; g++)
    {
        // Grid3D ctor: goal (0,0,0), start (0,0,0) -> grid heading atan2(0,0) = 0
        pp_host_update_goal(c->model.C, z, z, c->frames[g]);
        c->groups[g].map = c->d_maps + nn * g;
        c->groups[g].apf = nullptr;
        c->groups[g].K = 0;
        c->groups[g].bin_shift = 0; c->groups[g].bin_n = 0; c->groups[g].bin_off = nullptr; c->groups[g].bin_idx = nullptr;
        c->groups[g].pad = 0;
        c->groups[g].frame = c->frames[g].F;
    }
    c->groups_dirty = true;
    PP_CUDA(cudaStreamSynchronize(c->stream));
    return PP_SUCCESS;
}

int pp_create(const pp_params* params, int device, int num_groups, pp_context** out)
{
    if (!params || !out || num_groups < 1) return pp_fail(PP_ERR_INVALID, "pp_create: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return pp_fail(PP_ERR_NO_DEVICE, "pp_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return pp_fail(PP_ERR_INVALID, "pp_create: bad device index");
    pp_context* c = new pp_context();
    c->device = device;
    c->params = *params;
    std::string err;
    if (!pp_host_build_model(*params, c->model, err)) { delete c; return pp_fail(PP_ERR_INVALID, "pp_create: " + err); }
    int rc = create_impl(c, num_groups);
    if (rc) { std::string keep = g_last_error; pp_destroy(c); g_last_error = keep; return rc; }    // nothing leaks on a CUDA failure
    *out = c;
    return PP_SUCCESS;
}

// A lane: a context that shares `parent`'s parameters, maps, goal frames and APF lists (read-only) and owns its stream,
// its search scratch and its batch buffers.  Batches uploaded to different lanes run concurrently on the device, so the
// drain of one batch (its last long queries) overlaps the bulk of the next -- continuous batching for the EXACT mode.
// Map updates go through the parent; a lane picks them up at its next pp_batch_upload.  Destroy lanes before the parent.
int pp_create_lane(pp_context* parent, pp_context** out)
{
    if (!parent || !out) return pp_fail(PP_ERR_INVALID, "pp_create_lane: bad arguments");
    if (parent->parent) return pp_fail(PP_ERR_INVALID, "pp_create_lane: a lane cannot have lanes");
    pp_context* c = new pp_context();
    c->device = parent->device;
    c->params = parent->params;
    c->model = parent->model;
    c->parent = parent;
    int rc = create_impl(c, parent->num_groups);
    if (rc) { std::string keep = g_last_error; pp_destroy(c); g_last_error = keep; return rc; }
    *out = c;
    return PP_SUCCESS;
}

int pp_set_memory_budget(pp_context* c, unsigned long long bytes)
{
    if (!c) return pp_fail(PP_ERR_INVALID, "null context");
    c->mem_budget = (size_t)bytes;
    return PP_SUCCESS;
}

void pp_destroy(pp_context* c)
{
    if (!c) return;
    pp_comm_destroy(c);
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (!c->parent)
    {
        for (float* p : c->d_apf) if (p) cudaFree(p);
        for (auto& b : c->d_bin_off) b.release();
        for (auto& b : c->d_bin_idx) b.release();
        cudaFree(c->d_maps); cudaFree(c->d_map_tmp); cudaFree(c->d_cell_scratch); cudaFree(c->d_off_xy);
        cudaFree(c->d_groups);
    }
    cudaFree(c->d_counter);
    if (c->box_stage.h) cudaFreeHost(c->box_stage.h);
    if (c->box_stage.d) cudaFree(c->box_stage.d);
    for (int q = 0; q < 4; q++) if (c->box_stage.ev[q]) cudaEventDestroy(c->box_stage.ev[q]);
    c->s0.release(); c->s1.release(); c->s2.release(); c->s3.release(); c->s4.release();
    c->d_queries.release(); c->d_results.release(); c->d_paths.release(); c->d_trace.release();
    c->wp.release(); c->wp_lazy.release(); c->d_qmap.release(); c->d_order.release(); c->d_exact_order.release(); c->d_group_cost.release();
    cudaFree(c->d_lazy_sid);
    for (auto& h : c->hist) h.release();
    c->d_foot_bins.release(); c->d_foot_offs.release(); c->d_foot_lin.release(); c->d_foot_xyh.release(); c->d_foot_out.release();
    c->d_traj.release(); c->d_traj_tmp.release(); c->d_vel_in.release(); c->d_traj_int.release(); c->d_wframes.release();
    c->kp.release(); c->kp_retry.release();
    c->d_field2d.release(); c->d_f2d_work.release(); c->d_f2d_flags.release(); c->d_dubins_field.release(); c->d_f2d_ctl.release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_run0) cudaEventDestroy(c->ev_run0);
    if (c->ev_run1) cudaEventDestroy(c->ev_run1);
    if (c->ev_re0) cudaEventDestroy(c->ev_re0);
    if (c->ev_re1) cudaEventDestroy(c->ev_re1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int pp_get_consts(pp_context* c, pp_consts_info* o)
{
    if (!c || !o) return pp_fail(PP_ERR_INVALID, "pp_get_consts: null");
    const PPConsts& C = c->model.C;
    o->log_threshold = C.log_thr; o->log_min = C.log_min; o->log_max = C.log_max; o->log_free = C.log_free;
    o->precision = C.precision; o->r_min = C.r_min; o->ang_step = C.ang_step; o->n2 = C.n2; o->n45 = C.n45;
    return PP_SUCCESS;
}

int pp_get_frame(pp_context* c, int g, pp_frame_info* o)
{
    int rc = check_group(c, g); if (rc) return rc;
    const PPHostFrame& f = c->frames[g];
    o->grid_heading = f.grid_heading;
    for (int q = 0; q < 3; q++) o->goal_world[q] = f.goal_world[q];
    o->goal_grid[0] = f.F.goal_x; o->goal_grid[1] = f.F.goal_y; o->goal_grid[2] = f.F.goal_h;
    o->goal_bin = f.F.goal_bin; o->goal_ci = f.F.goal_ci; o->goal_cj = f.F.goal_cj;
    o->num_apf = c->groups[g].K;
    return PP_SUCCESS;
}

int pp_get_tables(pp_context* c, float* offset_xy, float* offset_heading, float* actions_cost, float* abs_curv)
{
    if (!c) return pp_fail(PP_ERR_INVALID, "null context");
    const PPConsts& C = c->model.C;
    for (int i = 0; i < C.S; i++)
    {
        offset_heading[i] = C.off_heading[i]; actions_cost[i] = C.act_cost3d[i]; abs_curv[i] = C.abs_curv[i];
        for (int j = 0; j < C.bins; j++)
        {
            offset_xy[((size_t)i * C.bins + j) * 2] = c->model.off_xy[((size_t)i * (C.bins + 1) + j) * 2];
            offset_xy[((size_t)i * C.bins + j) * 2 + 1] = c->model.off_xy[((size_t)i * (C.bins + 1) + j) * 2 + 1];
        }
    }
    return PP_SUCCESS;
}

int pp_sync(pp_context* c)
{
    if (!c) return pp_fail(PP_ERR_INVALID, "null context");
    PP_CUDA(cudaSetDevice(c->device));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    return PP_SUCCESS;
}

// ---- map -------------------------------------------------------------------------------------------
int pp_update_goal(pp_context* c, int g, const float* goal3, const float* start3)
{
    int rc = check_group(c, g); if (rc) return rc;
    rc = lane_guard(c, "pp_update_goal"); if (rc) return rc;
    map_changed(c, g);
    PP_CUDA(cudaSetDevice(c->device));
    const PPConsts& C = c->model.C;
    PPHostFrame prev = c->frames[g];
    pp_host_update_goal(C, goal3, start3, c->frames[g]);
    c->groups[g].frame = c->frames[g].F;
    c->groups_dirty = true;
    // relocate_obstacles (Grid3D.cpp:169-203): forward scatter, last writer in raster order wins
    PPRelocDesc d;
    pp_host_reloc_desc(C, c->frames[g].grid_heading, prev.grid_heading, c->frames[g].goal_world, prev.goal_world, d);
    size_t nn = nn_of(c);
    float* map = c->d_maps + nn * g;
    int blocks = c->sm_count * 8;
    pp_map_reloc_fill_kernel<<<blocks, 256, 0, c->stream>>>(c->d_cell_scratch, nn);
    pp_map_reloc_scatter_kernel<<<blocks, 256, 0, c->stream>>>(c->d_cell_scratch, C.N, d);
    pp_map_reloc_gather_kernel<<<blocks, 256, 0, c->stream>>>(map, c->d_map_tmp, c->d_cell_scratch, nn);
    c->launches += 3;
    PP_CUDA(cudaMemcpyAsync(map, c->d_map_tmp, sizeof(float) * nn, cudaMemcpyDeviceToDevice, c->stream));
    PP_CUDA(cudaGetLastError());
    return PP_SUCCESS;
}

// Grid2D::update_goal_heading (lib/Grid2D.cpp:260-266): goal location and grid heading only -- the map is NOT relocated (that is
// Grid3D's override, pp_update_goal)
int pp_update_goal_frame(pp_context* c, int g, const float* goal3, const float* start3)
{
    int rc = check_group(c, g); if (rc) return rc;
    rc = lane_guard(c, "pp_update_goal_frame"); if (rc) return rc;
    if (!goal3 || !start3) return pp_fail(PP_ERR_INVALID, "pp_update_goal_frame: null");
    map_changed(c, g);
    pp_host_update_goal(c->model.C, goal3, start3, c->frames[g]);
    c->groups[g].frame = c->frames[g].F;
    c->groups_dirty = true;
    return PP_SUCCESS;
}

int pp_reset(pp_context* c, int g)
{
    int rc = check_group(c, g); if (rc) return rc;
    if ((int)c->hist.size() > g && c->hist[g].on)
    {
        // AStar::reset() on the carried cache: the visited flags go, the node costs stay
        PP_CUDA(cudaSetDevice(c->device));
        int nn = (int)nn_of(c);
        pp_hist_reset_kernel<<<std::min((nn + 255) / 256, 4 * c->sm_count), 256, 0, c->stream>>>(c->hist[g].cell_state.p, nn);
        c->launches += 1;
        PP_CUDA(cudaGetLastError());
    }
    return PP_SUCCESS;
}

int pp_set_history(pp_context* c, int g, int enable)
{
    int rc = check_group(c, g); if (rc) return rc;
    PP_CUDA(cudaSetDevice(c->device));
    if ((int)c->hist.size() != c->num_groups) c->hist.resize(c->num_groups);
    pp_context::HistBuf& h = c->hist[g];
    if (!enable) { PP_CUDA(cudaStreamSynchronize(c->stream)); h.release(); return PP_SUCCESS; }
    size_t nn = nn_of(c);
    PP_CUDA(h.cell_state.ensure(nn)); PP_CUDA(h.cell_state_bak.ensure(nn + 1)); PP_CUDA(h.sid.ensure(1));
    PP_CUDA(h.nm_g.ensure(nn)); PP_CUDA(h.nm_f.ensure(nn)); PP_CUDA(h.nm_g_bak.ensure(nn)); PP_CUDA(h.nm_f_bak.ensure(nn));
    // the freshly constructed planner: nothing visited, no node touched (g = 0, f = Euclidean h), search id 0
    PP_CUDA(cudaMemsetAsync(h.cell_state.p, 0, nn * sizeof(unsigned), c->stream));
    PP_CUDA(cudaMemsetAsync(h.sid.p, 0, sizeof(unsigned), c->stream));
    h.on = true;
    return PP_SUCCESS;
}

int pp_update_obstacles_decay(pp_context* c, int g)
{
    int rc = check_group(c, g); if (rc) return rc;
    rc = lane_guard(c, "pp_update_obstacles_decay"); if (rc) return rc;
    map_changed(c, g);
    PP_CUDA(cudaSetDevice(c->device));
    const PPConsts& C = c->model.C;
    size_t nn = nn_of(c);
    size_t want = (nn / 4 + 255) / 256;
    int blocks = (int)std::min<size_t>(std::max<size_t>(want, 1), (size_t)c->sm_count * 8);
    pp_map_decay_kernel<<<blocks, 256, 0, c->stream>>>(c->d_maps + nn * g, nn, C.log_free, C.log_min, C.log_max);
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    return PP_SUCCESS;
}

static int box_stage_slot(pp_context* c, int n, int* slot_out)
{
    pp_context::BoxStage& st = c->box_stage;
    if (n > st.cap)
    {
        PP_CUDA(cudaStreamSynchronize(c->stream));
        if (st.h) cudaFreeHost(st.h);
        if (st.d) cudaFree(st.d);
        st.h = nullptr; st.d = nullptr;
        int cap = 256; while (cap < n) cap <<= 1;
        PP_CUDA(cudaMallocHost(&st.h, sizeof(PPBoxDescDev) * (size_t)cap * 4));
        PP_CUDA(cudaMalloc(&st.d, sizeof(PPBoxDescDev) * (size_t)cap * 4));
        for (int q = 0; q < 4; q++) { if (!st.ev[q]) PP_CUDA(cudaEventCreateWithFlags(&st.ev[q], cudaEventDisableTiming)); st.used[q] = false; }
        st.cap = cap;
    }
    const int slot = (int)(st.next++ & 3u);
    if (st.used[slot]) PP_CUDA(cudaEventSynchronize(st.ev[slot]));
    *slot_out = slot;
    return PP_SUCCESS;
}

// Grid2D::update_obstacles(boxes, conf) [+ Grid2D::update_obstacles()] as one asynchronous launch
static int map_update_launch(pp_context* c, int g, const float* boxes, const float* conf, int n, int do_decay)
{
    const PPConsts& C = c->model.C;
    float ch = 1.0f, sh = 0.0f;
    int slot = 0;
    const PPBoxDescDev* d_desc = nullptr;
    if (n > 0)
    {
        std::vector<PPBoxDesc> d;
        pp_host_box_descs(C, c->frames[g], boxes, conf, n, ch, sh, d);
        int rc = box_stage_slot(c, n, &slot); if (rc) return rc;
        pp_context::BoxStage& st = c->box_stage;
        PPBoxDescDev* h = st.h + (size_t)slot * st.cap;
        for (int k = 0; k < n; k++)
        {
            PPBoxDescDev& o = h[k];
            o.start_i = d[k].start_i; o.start_j = d[k].start_j; o.ni = d[k].ni; o.nj = d[k].nj; o.delta = d[k].delta; o.pad = 0;
            const bool empty = d[k].lo_i > d[k].hi_i || d[k].lo_j > d[k].hi_j || d[k].ni <= 0 || d[k].nj <= 0;
            o.lo_i = (short)(empty ? 1 : d[k].lo_i); o.hi_i = (short)(empty ? 0 : d[k].hi_i);
            o.lo_j = (short)(empty ? 1 : d[k].lo_j); o.hi_j = (short)(empty ? 0 : d[k].hi_j);
        }
        PP_CUDA(cudaMemcpyAsync(st.d + (size_t)slot * st.cap, h, sizeof(PPBoxDescDev) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
        d_desc = st.d + (size_t)slot * st.cap;
    }
    else if (!do_decay) return PP_SUCCESS;
    const int T = (C.N + PP_TILE - 1) / PP_TILE;
    pp_map_update_kernel<<<dim3(T, T), 256, 0, c->stream>>>(c->d_maps + nn_of(c) * g, C.N, d_desc, n, ch, sh, C.log_min, C.log_max,
                                                            do_decay, C.log_free);
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    if (n > 0) { PP_CUDA(cudaEventRecord(c->box_stage.ev[slot], c->stream)); c->box_stage.used[slot] = true; }
    return PP_SUCCESS;
}

int pp_update_obstacles_boxes_2d(pp_context* c, int g, const float* boxes, const float* conf, int n)
{
    int rc = check_group(c, g); if (rc) return rc;
    rc = lane_guard(c, "pp_update_obstacles_boxes_2d"); if (rc) return rc;
    map_changed(c, g);
    if (n < 0 || (n > 0 && (!boxes || !conf))) return pp_fail(PP_ERR_INVALID, "boxes: bad arguments");
    if (c->model.C.N > 32767) return pp_fail(PP_ERR_INVALID, "boxes: at most 32767 cells per side");
    if (n == 0) return PP_SUCCESS;
    PP_CUDA(cudaSetDevice(c->device));
    return map_update_launch(c, g, boxes, conf, n, 0);
}

// One round of the map update -- Grid2D::update_obstacles(boxes, conf) followed by Grid2D::update_obstacles() -- as a single pass
// over the map (BASELINE configs[1]: every cell is read and written once).
int pp_update_obstacles_boxes_2d_decay(pp_context* c, int g, const float* boxes, const float* conf, int n)
{
    int rc = check_group(c, g); if (rc) return rc;
    rc = lane_guard(c, "pp_update_obstacles_boxes_2d_decay"); if (rc) return rc;
    map_changed(c, g);
    if (n < 0 || (n > 0 && (!boxes || !conf))) return pp_fail(PP_ERR_INVALID, "boxes: bad arguments");
    if (c->model.C.N > 32767) return pp_fail(PP_ERR_INVALID, "boxes: at most 32767 cells per side");
    PP_CUDA(cudaSetDevice(c->device));
    return map_update_launch(c, g, boxes, conf, n, 1);
}

// Grid3D::update_obstacles' own part (Grid3D.cpp:22-44): the APF obstacle list of the group, rebuilt from the boxes; the map is
// not touched.  A rank that receives its maps by pp_broadcast_maps calls this instead of pp_update_obstacles_boxes.
int pp_update_obstacles_apf(pp_context* c, int g, const float* boxes, int n, float apf_added_radius)
{
    int rc = check_group(c, g); if (rc) return rc;
    rc = lane_guard(c, "pp_update_obstacles_apf"); if (rc) return rc;
    if (n < 0 || (n > 0 && !boxes)) return pp_fail(PP_ERR_INVALID, "boxes: bad arguments");
    PP_CUDA(cudaSetDevice(c->device));
    // APF list rebuild, Grid3D.cpp:26-40
    std::vector<float> apf;
    pp_host_apf_list(c->model.C, c->frames[g], boxes, n, apf_added_radius, apf);
    if (n > c->apf_cap[g])
    {
        PP_CUDA(cudaStreamSynchronize(c->stream));
        if (c->d_apf[g]) PP_CUDA(cudaFree(c->d_apf[g]));
        c->d_apf[g] = nullptr;
        PP_CUDA(cudaMalloc(&c->d_apf[g], sizeof(float) * 3 * n));
        c->apf_cap[g] = n;
    }
    if (n > 0) PP_CUDA(cudaMemcpyAsync(c->d_apf[g], apf.data(), sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    c->groups[g].apf = c->d_apf[g];
    c->groups[g].K = n;
    {
        std::vector<int> off, idx;
        int bin_n = 0;
        const int shift = 4;
        pp_host_apf_bins(c->model.C, apf, n, shift, bin_n, off, idx);
        PP_CUDA(c->d_bin_off[g].ensure(off.size()));
        PP_CUDA(c->d_bin_idx[g].ensure(std::max<size_t>(idx.size(), 1)));
        PP_CUDA(cudaMemcpyAsync(c->d_bin_off[g].p, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice, c->stream));
        if (!idx.empty()) PP_CUDA(cudaMemcpyAsync(c->d_bin_idx[g].p, idx.data(), sizeof(int) * idx.size(), cudaMemcpyHostToDevice, c->stream));
        PP_CUDA(cudaStreamSynchronize(c->stream));      // the pageable host vectors above die at return
        c->groups[g].bin_shift = shift; c->groups[g].bin_n = bin_n;
        c->groups[g].bin_off = c->d_bin_off[g].p; c->groups[g].bin_idx = c->d_bin_idx[g].p;
    }
    c->groups_dirty = true;
    map_changed(c, g);
    return PP_SUCCESS;
}

int pp_update_obstacles_boxes(pp_context* c, int g, const float* boxes, const float* conf, int n, float apf_added_radius)
{
    int rc = pp_update_obstacles_apf(c, g, boxes, n, apf_added_radius); if (rc) return rc;
    if (n > 0 && !conf) return pp_fail(PP_ERR_INVALID, "boxes: bad arguments");
    return pp_update_obstacles_boxes_2d(c, g, boxes, conf, n);
}

int pp_update_obstacles_lines(pp_context* c, int g, const float* lines, const float* conf, int n, float width)
{
    int rc = check_group(c, g); if (rc) return rc;
    rc = lane_guard(c, "pp_update_obstacles_lines"); if (rc) return rc;
    map_changed(c, g);
    if (n < 0 || (n > 0 && (!lines || !conf))) return pp_fail(PP_ERR_INVALID, "lines: bad arguments");
    if (n == 0) return PP_SUCCESS;
    PP_CUDA(cudaSetDevice(c->device));
    const PPConsts& C = c->model.C;
    std::vector<PPLineDesc> d;
    pp_host_line_descs(C, c->frames[g], lines, conf, n, d);
    PP_CUDA(c->s4.ensure(sizeof(PPLineDesc) * n));
    PP_CUDA(cudaMemcpyAsync(c->s4.p, d.data(), sizeof(PPLineDesc) * n, cudaMemcpyHostToDevice, c->stream));
    const int T = (C.N + PP_TILE - 1) / PP_TILE;
    pp_map_lines_kernel<<<dim3(T, T), 256, 0, c->stream>>>(c->d_maps + nn_of(c) * g, C.N, C.n45, C.n2, C.res,
                                                          (const PPLineDesc*)c->s4.p, n, width, C.log_min, C.log_max);
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    PP_CUDA(cudaStreamSynchronize(c->stream));
    return PP_SUCCESS;
}

int pp_map_download(pp_context* c, int g, float* out)
{
    int rc = check_group(c, g); if (rc) return rc;
    PP_CUDA(cudaSetDevice(c->device));
    PP_CUDA(cudaMemcpyAsync(out, c->d_maps + nn_of(c) * g, sizeof(float) * nn_of(c), cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    return PP_SUCCESS;
}

int pp_map_upload(pp_context* c, int g, const float* in)
{
    int rc = check_group(c, g); if (rc) return rc;
    rc = lane_guard(c, "pp_map_upload"); if (rc) return rc;
    map_changed(c, g);
    PP_CUDA(cudaSetDevice(c->device));
    PP_CUDA(cudaMemcpyAsync(c->d_maps + nn_of(c) * g, in, sizeof(float) * nn_of(c), cudaMemcpyHostToDevice, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    return PP_SUCCESS;
}

// The caller may write through this pointer (e.g. its own NCCL receive): everything derived from the group's map -- the exact 2D
// field, K-POP scheduling hints -- is dropped here, and once more by pp_map_mark_dirty after the caller's write has completed.
void* pp_map_device_ptr(pp_context* c, int g)
{
    if (check_group(c, g)) return nullptr;
    map_changed(c, g);
    return c->d_maps + nn_of(c) * g;
}

int pp_map_mark_dirty(pp_context* c, int g)
{
    int rc = check_group(c, g); if (rc) return rc;
    map_changed(c, g);
    return PP_SUCCESS;
}

// ---- host-side start nodes ---------------------------------------------------------------------------
int pp_set_start_batch(pp_context* c, const pp_query* q, int n, pp_state* out)
{
    if (!c || !q || !out) return pp_fail(PP_ERR_INVALID, "pp_set_start_batch: null");
    for (int k = 0; k < n; k++)
    {
        if (q[k].group < 0 || q[k].group >= c->num_groups) return pp_fail(PP_ERR_INVALID, "query group out of range");
        PPState s = pp_host_set_start(c->model.C, c->frames[q[k].group], q[k].x, q[k].y, q[k].heading, q[k].vel);
        out[k].x = s.x; out[k].y = s.y; out[k].heading = s.heading; out[k].g = s.g; out[k].f = s.f;
        out[k].vmin_sqr = s.vmin_sqr; out[k].curvature_index = s.curv; out[k].angle_bin = s.bin; out[k].ci = s.ci; out[k].cj = s.cj;
    }
    return PP_SUCCESS;
}

// ---- stateless batches -------------------------------------------------------------------------------
static_assert(sizeof(pp_state) == sizeof(PPState), "pp_state layout");
static_assert(sizeof(pp_pop) == sizeof(PPPop), "pp_pop layout");

static int successors(pp_context* c, int g, const pp_state* in, int n, pp_state* out, int* n_out, int* flags, int expand)
{
    int rc = check_group(c, g); if (rc) return rc;
    if (n <= 0) return PP_SUCCESS;
    PP_CUDA(cudaSetDevice(c->device));
    const PPConsts& C = c->model.C;
    int stride = 2 * C.A + 1;
    PP_CUDA(c->s0.ensure(sizeof(PPState) * n));
    PP_CUDA(c->s1.ensure(sizeof(PPState) * (size_t)n * stride));
    PP_CUDA(c->s2.ensure(sizeof(int) * n));
    PP_CUDA(c->s3.ensure(sizeof(int) * n));
    PP_CUDA(cudaMemcpyAsync(c->s0.p, in, sizeof(PPState) * n, cudaMemcpyHostToDevice, c->stream));
    PP_CUDA(cudaMemsetAsync(c->s1.p, 0, sizeof(PPState) * (size_t)n * stride, c->stream));
    PPSuccArgs a;
    a.C = C; a.off_xy = c->d_off_xy; a.G = c->groups[g]; a.in = (const PPState*)c->s0.p; a.n = n; a.expand = expand;
    a.out = (PPState*)c->s1.p; a.n_out = (int*)c->s2.p; a.flags = (int*)c->s3.p;
    pp_successor_kernel<<<(n + 3) / 4, 128, 0, c->stream>>>(a);
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    PP_CUDA(cudaMemcpyAsync(out, c->s1.p, sizeof(PPState) * (size_t)n * stride, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaMemcpyAsync(n_out, c->s2.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaMemcpyAsync(flags, c->s3.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    return PP_SUCCESS;
}

int pp_rollout_batch(pp_context* c, const pp_state* in, int n, pp_state* out, int* n_out, int* flags)
{ return successors(c, 0, in, n, out, n_out, flags, 0); }

int pp_expand_batch(pp_context* c, int g, const pp_state* in, int n, pp_state* out, int* n_out, int* flags)
{ return successors(c, g, in, n, out, n_out, flags, 1); }

int pp_collision_batch(pp_context* c, int g, const float* xy, int n, int* free_out, int* cells_ij)
{
    int rc = check_group(c, g); if (rc) return rc;
    if (n <= 0) return PP_SUCCESS;
    PP_CUDA(cudaSetDevice(c->device));
    PP_CUDA(c->s0.ensure(sizeof(float) * 2 * n));
    PP_CUDA(c->s1.ensure(sizeof(int) * n));
    PP_CUDA(c->s2.ensure(sizeof(int) * 2 * n));
    PP_CUDA(cudaMemcpyAsync(c->s0.p, xy, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, c->stream));
    pp_collision_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->model.C, c->groups[g].map, (const float*)c->s0.p, 2, n, 0,
                                                              (int*)c->s1.p, (int*)c->s2.p);
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    PP_CUDA(cudaMemcpyAsync(free_out, c->s1.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    if (cells_ij) PP_CUDA(cudaMemcpyAsync(cells_ij, c->s2.p, sizeof(int) * 2 * n, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    return PP_SUCCESS;
}

// ---- generic vehicle-footprint collision check (north_star (c); core/pp_footprint.h) ----
int pp_set_footprint(pp_context* c, float length, float width, float rear_overhang)
{
    if (!c || !(length >= 0.0f) || !(width >= 0.0f) || !(rear_overhang >= 0.0f) || rear_overhang > length)
        return pp_fail(PP_ERR_INVALID, "pp_set_footprint: need 0 <= rear_overhang <= length, width >= 0");
    PP_CUDA(cudaSetDevice(c->device));
    int win = pp_footprint_build(c->model.C, length, width, rear_overhang, c->foot_bins, c->foot_offs);
    if (win > PP_FOOT_MAX_WIN) { c->foot_bins.clear(); return pp_fail(PP_ERR_CAPACITY, "pp_set_footprint: footprint spans more than 96 cells"); }
    c->foot_win = win;
    PP_CUDA(c->d_foot_bins.ensure(c->foot_bins.size()));
    PP_CUDA(c->d_foot_offs.ensure(c->foot_offs.size()));
    PP_CUDA(cudaMemcpyAsync(c->d_foot_bins.p, c->foot_bins.data(), sizeof(PPFootBin) * c->foot_bins.size(), cudaMemcpyHostToDevice, c->stream));
    PP_CUDA(cudaMemcpyAsync(c->d_foot_offs.p, c->foot_offs.data(), sizeof(PPCellOff) * c->foot_offs.size(), cudaMemcpyHostToDevice, c->stream));
    std::vector<int> lin(c->foot_offs.size());
    for (size_t k = 0; k < lin.size(); k++) lin[k] = (int)c->foot_offs[k].di * c->model.C.N + (int)c->foot_offs[k].dj;
    PP_CUDA(c->d_foot_lin.ensure(lin.size()));
    PP_CUDA(cudaMemcpyAsync(c->d_foot_lin.p, lin.data(), sizeof(int) * lin.size(), cudaMemcpyHostToDevice, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    return PP_SUCCESS;
}

int pp_get_footprint(pp_context* c, int bin, int* count, short* offs_ij, int cap)
{
    if (!c || c->foot_bins.empty()) return pp_fail(PP_ERR_INVALID, "pp_get_footprint: call pp_set_footprint first");
    if (bin < 0 || bin >= (int)c->foot_bins.size()) return pp_fail(PP_ERR_INVALID, "pp_get_footprint: bin out of range");
    const PPFootBin& b = c->foot_bins[bin];
    if (count) *count = b.count;
    for (int k = 0; offs_ij && k < b.count && k < cap; k++) { offs_ij[2 * k] = c->foot_offs[b.first + k].di; offs_ij[2 * k + 1] = c->foot_offs[b.first + k].dj; }
    return PP_SUCCESS;
}

int pp_footprint_batch(pp_context* c, int g, const float* xyh, int n, int* free_out, int* cells_ij, int* hits_out, float* kernel_ms)
{
    int rc = check_group(c, g); if (rc) return rc;
    if (c->foot_bins.empty()) return pp_fail(PP_ERR_INVALID, "pp_footprint_batch: call pp_set_footprint first");
    if (n <= 0) return PP_SUCCESS;
    if (!xyh || !free_out) return pp_fail(PP_ERR_INVALID, "pp_footprint_batch: null buffer");
    PP_CUDA(cudaSetDevice(c->device));
    PP_CUDA(c->d_foot_xyh.ensure((size_t)3 * n));
    PP_CUDA(c->d_foot_out.ensure((size_t)4 * n));
    PP_CUDA(cudaMemcpyAsync(c->d_foot_xyh.p, xyh, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    PPFootArgs a;
    a.C = c->model.C; a.map = c->d_maps + (size_t)g * nn_of(c); a.xyh = c->d_foot_xyh.p; a.n = n;
    a.bins = c->d_foot_bins.p; a.offs = c->d_foot_offs.p; a.lin = c->d_foot_lin.p; a.win = c->foot_win;
    a.free_out = c->d_foot_out.p; a.cells = cells_ij ? c->d_foot_out.p + n : nullptr; a.hits = hits_out ? c->d_foot_out.p + 3 * (size_t)n : nullptr;
    // default: direct gather (L1 is the staging buffer); PP_B200_FOOT_STAGED=1 selects the shared-memory-staged variant
    static const bool staged = [] { const char* e = std::getenv("PP_B200_FOOT_STAGED"); return e && e[0] == '1'; }();
    const size_t smem = staged ? (size_t)PP_FOOT_WARPS * a.win * a.win * sizeof(float) : 0;
    int occ = 0;
    if (staged)
    {
        PP_CUDA(cudaFuncSetAttribute(pp_footprint_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
        PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pp_footprint_kernel<true>, PP_FOOT_WARPS * 32, smem));
    }
    else PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pp_footprint_kernel<false>, PP_FOOT_WARPS * 32, 0));
    int blocks = std::min((n + 32 * PP_FOOT_WARPS - 1) / (32 * PP_FOOT_WARPS), std::max(1, occ) * c->sm_count);     // 32 poses per warp step, at most one full wave
    PP_CUDA(cudaEventRecord(c->ev0, c->stream));
    if (staged) pp_footprint_kernel<true><<<blocks, PP_FOOT_WARPS * 32, smem, c->stream>>>(a);
    else pp_footprint_kernel<false><<<blocks, PP_FOOT_WARPS * 32, 0, c->stream>>>(a);
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    PP_CUDA(cudaEventRecord(c->ev1, c->stream));
    PP_CUDA(cudaMemcpyAsync(free_out, c->d_foot_out.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    if (cells_ij) PP_CUDA(cudaMemcpyAsync(cells_ij, c->d_foot_out.p + n, sizeof(int) * 2 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    if (hits_out) PP_CUDA(cudaMemcpyAsync(hits_out, c->d_foot_out.p + 3 * (size_t)n, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    if (kernel_ms) PP_CUDA(cudaEventElapsedTime(kernel_ms, c->ev0, c->ev1));
    return PP_SUCCESS;
}

int pp_check_path(pp_context* c, int g, const float* xyh, int n, int* free_out)
{
    int rc = check_group(c, g); if (rc) return rc;
    *free_out = 1;
    if (n <= 0) return PP_SUCCESS;
    PP_CUDA(cudaSetDevice(c->device));
    PP_CUDA(c->s0.ensure(sizeof(float) * 3 * n));
    PP_CUDA(c->s1.ensure(sizeof(int) * n));
    PP_CUDA(cudaMemcpyAsync(c->s0.p, xyh, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    pp_collision_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->model.C, c->groups[g].map, (const float*)c->s0.p, 3, n, 1,
                                                              (int*)c->s1.p, nullptr);
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    std::vector<int> fr(n);
    PP_CUDA(cudaMemcpyAsync(fr.data(), c->s1.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < n; k++) if (!fr[k]) { *free_out = 0; break; }
    return PP_SUCCESS;
}

int pp_apf_batch(pp_context* c, int g, const float* xyh, int n, float* out)
{
    int rc = check_group(c, g); if (rc) return rc;
    if (n <= 0) return PP_SUCCESS;
    PP_CUDA(cudaSetDevice(c->device));
    PP_CUDA(c->s0.ensure(sizeof(float) * 3 * n));
    PP_CUDA(c->s1.ensure(sizeof(float) * n));
    PP_CUDA(cudaMemcpyAsync(c->s0.p, xyh, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    pp_apf_kernel<<<(n + 3) / 4, 128, 0, c->stream>>>(c->model.C, c->groups[g], (const float*)c->s0.p, n, (float*)c->s1.p);
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    PP_CUDA(cudaMemcpyAsync(out, c->s1.p, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    return PP_SUCCESS;
}

int pp_dubins_length_batch(pp_context* c, const float* starts, int n, const float* goal3, float* length, int* type, float* params4)
{
    if (!c) return pp_fail(PP_ERR_INVALID, "null context");
    if (n <= 0) return PP_SUCCESS;
    PP_CUDA(cudaSetDevice(c->device));
    PP_CUDA(c->s0.ensure(sizeof(float) * 3 * n));
    PP_CUDA(c->s1.ensure(sizeof(float) * n));
    PP_CUDA(c->s2.ensure(sizeof(int) * n));
    PP_CUDA(c->s3.ensure(sizeof(float) * 4 * n));
    PP_CUDA(cudaMemcpyAsync(c->s0.p, starts, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    size_t threads = (size_t)n * 4;
    pp_dubins_length_kernel<<<(int)((threads + 127) / 128), 128, 0, c->stream>>>(c->model.C.r_min, (const float*)c->s0.p, n,
        goal3[0], goal3[1], goal3[2], (float*)c->s1.p, (int*)c->s2.p, (float*)c->s3.p);
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    PP_CUDA(cudaMemcpyAsync(length, c->s1.p, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    if (type) PP_CUDA(cudaMemcpyAsync(type, c->s2.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    if (params4) PP_CUDA(cudaMemcpyAsync(params4, c->s3.p, sizeof(float) * 4 * n, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    return PP_SUCCESS;
}

int pp_dubins_length_fp32_batch(pp_context* c, const float* starts, int n, const float* goal3, float* length)
{
    if (!c) return pp_fail(PP_ERR_INVALID, "null context");
    if (n <= 0) return PP_SUCCESS;
    if (!starts || !goal3 || !length) return pp_fail(PP_ERR_INVALID, "pp_dubins_length_fp32_batch: bad arguments");
    PP_CUDA(cudaSetDevice(c->device));
    PP_CUDA(c->s0.ensure(sizeof(float) * 3 * (size_t)n));
    PP_CUDA(c->s1.ensure(sizeof(float) * (size_t)n));
    PP_CUDA(cudaMemcpyAsync(c->s0.p, starts, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    int blocks = std::min((n + 255) / 256, c->sm_count * 8);
    pp_dubins_length_fp32_kernel<<<blocks, 256, 0, c->stream>>>(c->model.C, (const float*)c->s0.p, n, goal3[0], goal3[1], goal3[2],
                                                                (float*)c->s1.p);
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    PP_CUDA(cudaMemcpyAsync(length, c->s1.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    return PP_SUCCESS;
}

int pp_dubins_path(pp_context* c, const float* s, const float* g, float* xyh, float* curvature, int cap, int* n_out,
                   float* length, int* long_turn_flag)
{
    if (!c) return pp_fail(PP_ERR_INVALID, "null context");
    PP_CUDA(cudaSetDevice(c->device));
    PP_CUDA(c->s0.ensure(sizeof(PPPathPt) * (size_t)cap));
    PP_CUDA(c->s1.ensure(16));
    int* d_n = (int*)c->s1.p; float* d_len = (float*)(c->s1.p + 4); int* d_flag = (int*)(c->s1.p + 8);
    pp_dubins_path_kernel<<<1, 32, 0, c->stream>>>(c->model.C, s[0], s[1], s[2], g[0], g[1], g[2], (PPPathPt*)c->s0.p, cap,
                                                   d_n, d_len, d_flag);
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    unsigned char hdr[16];
    PP_CUDA(cudaMemcpyAsync(hdr, c->s1.p, 16, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    int n; std::memcpy(&n, hdr, 4); std::memcpy(length, hdr + 4, 4); std::memcpy(long_turn_flag, hdr + 8, 4);
    *n_out = n;
    int m = std::min(n, cap);
    std::vector<PPPathPt> pts(std::max(m, 1));
    if (m > 0) PP_CUDA(cudaMemcpy(pts.data(), c->s0.p, sizeof(PPPathPt) * m, cudaMemcpyDeviceToHost));
    for (int k = 0; k < m; k++)
    {
        xyh[3 * k] = pts[k].x; xyh[3 * k + 1] = pts[k].y; xyh[3 * k + 2] = pts[k].heading;
        curvature[k] = pts[k].curvature;
    }
    return PP_SUCCESS;
}

// ---- batch search ------------------------------------------------------------------------------------
static void fill_args(pp_context* c, const WorkPools& w, const pp_search_opts& o, int n_slots, const int* qmap, int n_work, PPBatchArgs& a)
{
    a.C = c->model.C; a.off_xy = c->d_off_xy; a.groups = c->d_groups; a.queries = c->d_queries.p; a.n_queries = n_work;
    a.qmap = qmap; a.order = nullptr;
    a.n_slots = n_slots; a.counter = c->d_counter; a.results = c->d_results.p;
    a.paths = c->d_paths.p; a.path_cap = o.path_cap;
    a.trace = o.trace_cap > 0 ? c->d_trace.p : nullptr; a.trace_cap = o.trace_cap;
    a.open3 = w.open3.p; a.open3_cap = w.open3_cap; a.closed = w.closed.p; a.closed_cap = w.closed_cap;
    a.chash = w.chash.p; a.chash_cap = w.chash_cap; a.cell_state = w.cell_state.p; a.nm_g = w.nm_g.p; a.nm_f = w.nm_f.p;
    a.cl_g = w.cl_g.p; a.cl_prev = w.cl_prev.p; a.open2 = w.open2.p; a.open2_cap = w.open2_cap;
    a.arena = w.arena_bytes ? w.arena_ctl.p : nullptr;
    a.closed_max = std::max(o.max_expansions, w.closed_cap); a.open3_max = std::max(o.max_open, w.open3_cap);
    a.open2_max = std::max(o.max_open2d, w.open2_cap);
    a.hist_cell_state = nullptr; a.hist_nm_g = nullptr; a.hist_nm_f = nullptr; a.hist_sid = nullptr;
    if (c->batch_hist_group >= 0 && n_slots == 1 && n_work == 1)
    {
        pp_context::HistBuf& h = c->hist[c->batch_hist_group];
        a.hist_cell_state = h.cell_state.p; a.hist_nm_g = h.nm_g.p; a.hist_nm_f = h.nm_f.p; a.hist_sid = h.sid.p;
    }
}

// snapshot / restore of a carried cache around one query, so that a capacity retry re-runs it from the same history
static int hist_copy(pp_context* c, int g, bool restore)
{
    pp_context::HistBuf& h = c->hist[g];
    size_t nn = nn_of(c);
    unsigned *cs_a = restore ? h.cell_state.p : h.cell_state_bak.p, *cs_b = restore ? h.cell_state_bak.p : h.cell_state.p;
    PP_CUDA(cudaMemcpyAsync(cs_a, cs_b, nn * sizeof(unsigned), cudaMemcpyDeviceToDevice, c->stream));
    PP_CUDA(cudaMemcpyAsync(restore ? h.nm_g.p : h.nm_g_bak.p, restore ? h.nm_g_bak.p : h.nm_g.p, nn * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    PP_CUDA(cudaMemcpyAsync(restore ? h.nm_f.p : h.nm_f_bak.p, restore ? h.nm_f_bak.p : h.nm_f.p, nn * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    PP_CUDA(cudaMemcpyAsync(restore ? h.sid.p : h.cell_state_bak.p + nn, restore ? h.cell_state_bak.p + nn : h.sid.p, sizeof(unsigned),
                            cudaMemcpyDeviceToDevice, c->stream));
    return PP_SUCCESS;
}

// bytes the search scratch of this context may take (slots + arena)
static int work_budget(pp_context* c, size_t already, size_t* out)
{
    if (c->mem_budget) { *out = c->mem_budget; return PP_SUCCESS; }
    size_t free_b = 0, total_b = 0;
    PP_CUDA(cudaMemGetInfo(&free_b, &total_b));
    *out = (size_t)((free_b + already) * 0.75);
    return PP_SUCCESS;
}

// (Re)allocate `w` for up to `want_slots` resident queries.  with_arena: the per-slot pools get the small first-pass
// capacities (never more than the caller's caps) and whatever the budget leaves goes to the arena the containers grow
// into; otherwise the pools are fixed at the given capacities (stand-alone lazy A* handle).
static int ensure_work(pp_context* c, WorkPools& w, int want_slots, int max_exp, int max_open, int max_open2d, bool with_arena)
{
    size_t nn = nn_of(c);
    const int cap_c = with_arena ? std::min(max_exp, PP_INIT_CLOSED) : max_exp;
    const int cap_o = with_arena ? std::min(max_open, PP_INIT_OPEN3) : max_open;
    const int cap_2 = with_arena ? std::min(max_open2d, PP_INIT_OPEN2D) : max_open2d;
    int hc = 1; while (hc < 2 * cap_c) hc <<= 1;
    const size_t per_slot = sizeof(PPNode3) * (size_t)cap_o + sizeof(PPClosed3) * (size_t)cap_c + sizeof(PPHashSlot) * (size_t)hc +
                            nn * 20 + sizeof(PPNode2) * (size_t)cap_2;
    if (w.alloc_slots >= 1 && w.open3_cap == cap_o && w.closed_cap == cap_c && w.open2_cap == cap_2 &&
        (w.arena_bytes != 0) == with_arena && (w.alloc_slots >= want_slots || w.clamped))
        return PP_SUCCESS;      // big enough, or already as large as this context's budget allows
    PP_CUDA(cudaStreamSynchronize(c->stream));
    w.release();
    size_t budget = 0;
    int rc = work_budget(c, 0, &budget); if (rc) return rc;
    int slots = want_slots;
    size_t arena_bytes = 0;
    if (with_arena)
    {
        // the fixed parts may take at most half of the budget; the arena gets the rest (capped: what the caps could ever need)
        if ((size_t)slots * per_slot > budget / 2) slots = (int)((budget / 2) / per_slot);
        if (slots < 1) return pp_fail(PP_ERR_CAPACITY, "not enough device memory for one query slot");
        arena_bytes = budget - (size_t)slots * per_slot;
        // An explicit budget (pp_set_memory_budget) is taken in full.  Otherwise: what the caps could ever need, but no more than
        // 1 GB + 64 MB per resident query (the C4 batch's longest query, 1.3 M expansions, needs about 190 MB); a query that finds
        // the arena empty is re-run by pp_batch_wait on a larger one (arena_grow).
        if (!c->mem_budget)
        {
            const double per_query_max = 2.0 * ((double)sizeof(PPNode3) * max_open + (double)sizeof(PPClosed3) * max_exp +
                                                (double)sizeof(PPHashSlot) * 4 * max_exp + (double)sizeof(PPNode2) * max_open2d);
            double need = std::min(per_query_max * slots, (double)((size_t)1 << 30) + (double)((size_t)64 << 20) * slots);
            need = std::max(need, (double)((size_t)16 << 20));
            if (need < (double)arena_bytes) arena_bytes = (size_t)need;
        }
        arena_bytes = std::max<size_t>(arena_bytes & ~(size_t)4095, 1 << 16);
    }
    else
    {
        if ((size_t)slots * per_slot > budget) slots = (int)(budget / per_slot);
        if (slots < 1) return pp_fail(PP_ERR_CAPACITY, "not enough device memory for one query slot");
    }
    PP_CUDA(w.open3.ensure((size_t)slots * cap_o));
    PP_CUDA(w.closed.ensure((size_t)slots * cap_c));
    PP_CUDA(w.chash.ensure((size_t)slots * hc));
    PP_CUDA(w.cell_state.ensure((size_t)slots * nn));
    PP_CUDA(w.nm_g.ensure((size_t)slots * nn));
    PP_CUDA(w.nm_f.ensure((size_t)slots * nn));
    PP_CUDA(w.cl_g.ensure((size_t)slots * nn));
    PP_CUDA(w.cl_prev.ensure((size_t)slots * nn));
    PP_CUDA(w.open2.ensure((size_t)slots * cap_2));
    if (with_arena)
    {
        PP_CUDA(w.arena_mem.ensure(arena_bytes));
        PP_CUDA(w.arena_ctl.ensure(1));
        w.arena_bytes = arena_bytes;
    }
    w.open3_cap = cap_o; w.closed_cap = cap_c; w.open2_cap = cap_2; w.chash_cap = hc;
    w.alloc_slots = slots; w.clamped = slots < want_slots;
    return PP_SUCCESS;
}

// every launch starts on an empty arena (all blocks of the previous launch went back to their class lists; starting over
// also undoes the fragmentation across size classes)
static int arena_reset(pp_context* c, WorkPools& w)
{
    if (!w.arena_bytes) return PP_SUCCESS;
    PPArena a;
    std::memset(&a, 0, sizeof(a));
    a.base = (unsigned long long)w.arena_mem.p; a.size = w.arena_bytes;
    PP_CUDA(cudaMemcpyAsync(w.arena_ctl.p, &a, sizeof(a), cudaMemcpyHostToDevice, c->stream));
    return PP_SUCCESS;
}

// a query found the arena empty: give the context a larger one (x4, within the budget) before the re-run
static int arena_grow(pp_context* c, WorkPools& w)
{
    if (!w.arena_bytes) return PP_SUCCESS;
    const size_t nn = nn_of(c);
    int hc = w.chash_cap;
    const size_t per_slot = sizeof(PPNode3) * (size_t)w.open3_cap + sizeof(PPClosed3) * (size_t)w.closed_cap + sizeof(PPHashSlot) * (size_t)hc +
                            nn * 20 + sizeof(PPNode2) * (size_t)w.open2_cap;
    const size_t fixed = per_slot * (size_t)w.alloc_slots;
    PP_CUDA(cudaStreamSynchronize(c->stream));
    const size_t old = w.arena_bytes;
    w.arena_mem.release();
    size_t budget = 0;
    int rc = work_budget(c, fixed, &budget); if (rc) return rc;
    size_t want = std::min(old * 4, budget > fixed ? budget - fixed : old);
    want = std::max(want, old) & ~(size_t)4095;
    cudaError_t e = w.arena_mem.ensure(want);
    if (e != cudaSuccess) { (void)cudaGetLastError(); want = old; PP_CUDA(w.arena_mem.ensure(want)); }
    w.arena_bytes = want;
    return PP_SUCCESS;
}

static pp_search_opts default_opts(const pp_search_opts* in)
{
    pp_search_opts o;
    if (in) o = *in; else std::memset(&o, 0, sizeof(o));
    // the reference's containers are unbounded: the defaults are caps a query only meets when the whole state space of a
    // 512^2 x 72 grid is in play; containers grow towards them on demand (pp_arena.h)
    if (o.mode != PP_MODE_KPOP)
    {
        if (o.max_expansions <= 0) o.max_expansions = 1 << 24;
        if (o.max_open <= 0) o.max_open = 1 << 23;
        if (o.max_open2d <= 0) o.max_open2d = 1 << 20;
    }
    else if (o.max_expansions <= 0) o.max_expansions = 1 << 17;      // K-POP pools are fixed per slot (2 nodes per expansion)
    if (o.path_cap <= 0) o.path_cap = 2048;
    if (o.trace_cap < 0) o.trace_cap = 0;
    if (o.mode != PP_MODE_KPOP) o.mode = PP_MODE_EXACT;
    if (o.kpop <= 0 || o.kpop > PP_K_MAXPOP) o.kpop = PP_K_MAXPOP;
    return o;
}

static int hw_slots_of(pp_context* c, int* out)
{
    int occ = 0;
    PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pp_search_kernel, PP_SEARCH_WARPS * 32, 0));
    *out = std::max(1, occ) * c->sm_count * PP_SEARCH_WARPS;
    return PP_SUCCESS;
}

static int field2d_run(pp_context* c, int g, int* sweeps, float* ms);

// ---- K-POP pools / launch ----------------------------------------------------------------------------
static int kpop_levels_for(int nodes_cap)
{
    // the top level alone must hold every live entry (see pp_klsm_insert)
    int levels = 1;
    while (((size_t)PP_K_RUN0 << (levels - 1)) < (size_t)nodes_cap + PP_K_RUN0 && levels < PP_K_LEVELS) levels++;
    return levels;
}

static int ensure_kpop(pp_context* c, pp_context::KPools& k, int want_slots, int nodes_cap, double mem_frac)
{
    int tc = 1; while (tc < 2 * nodes_cap) tc <<= 1;
    int levels = kpop_levels_for(nodes_cap);
    size_t arena_cap = (size_t)PP_K_RUN0 * (((size_t)1 << levels) - 1), tmp_cap = (size_t)nodes_cap + 2 * PP_K_RUN0;
    size_t per_slot = sizeof(PPKNode) * (size_t)nodes_cap + sizeof(PPKSlot) * (size_t)tc + sizeof(PPKEntry) * (arena_cap + 2 * tmp_cap);
    int slots = want_slots;
    if (!((k.alloc_slots >= slots || (k.clamped && k.alloc_slots >= 1)) && k.nodes_cap == nodes_cap))
    {
        PP_CUDA(cudaStreamSynchronize(c->stream));
        k.release();
        size_t free_b = 0, total_b = 0;
        PP_CUDA(cudaMemGetInfo(&free_b, &total_b));
        size_t budget = (size_t)(free_b * mem_frac);
        if ((size_t)slots * per_slot > budget) slots = (int)(budget / per_slot);
        if (slots < 1) return pp_fail(PP_ERR_CAPACITY, "not enough device memory for one K-POP query slot");
        PP_CUDA(k.nodes.ensure((size_t)slots * nodes_cap));
        PP_CUDA(k.table.ensure((size_t)slots * tc));
        PP_CUDA(cudaMemsetAsync(k.table.p, 0xFF, sizeof(PPKSlot) * (size_t)slots * tc, c->stream));   // all-ones = empty (pp_kpop.h)
        PP_CUDA(k.arena.ensure((size_t)slots * arena_cap));
        PP_CUDA(k.tmp_a.ensure((size_t)slots * tmp_cap));
        PP_CUDA(k.tmp_b.ensure((size_t)slots * tmp_cap));
        k.nodes_cap = nodes_cap; k.table_cap = tc; k.levels = levels; k.arena_cap = arena_cap; k.tmp_cap = tmp_cap; k.alloc_slots = slots;
        k.clamped = slots < want_slots;
    }
    return PP_SUCCESS;
}

// warps cooperating on one K-POP query (CTA size / 32): 4 unless PP_B200_KPOP_WARPS says 1, 2 or 8 (tuning knob)
static int kpop_warps()
{
    static int nw = 0;
    if (nw == 0)
    {
        nw = 4;
        if (const char* e = std::getenv("PP_B200_KPOP_WARPS")) { int v = std::atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) nw = v; }
    }
    return nw;
}

static int kpop_hw_slots(pp_context* c, int* out)
{
    int occ = 0;
    switch (kpop_warps())
    {
        case 1:  PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pp_kpop_kernel<1>, 32, 0)); break;
        case 2:  PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pp_kpop_kernel<2>, 64, 0)); break;
        case 8:  PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pp_kpop_kernel<8>, 256, 0)); break;
        default: PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pp_kpop_kernel<4>, 128, 0)); break;
    }
    *out = std::max(1, occ) * c->sm_count;
    return PP_SUCCESS;
}

static int launch_kpop(pp_context* c, const pp_context::KPools& k, int n_slots, const int* qmap, int n_work)
{
    PPKpopArgs a;
    a.C = c->model.C; a.off_xy = c->d_off_xy; a.groups = c->d_groups; a.field2d = c->d_field2d.p; a.queries = c->d_queries.p;
    a.qmap = qmap; a.n_queries = n_work; a.n_slots = n_slots; a.kpop = c->opts.kpop; a.counter = c->d_counter; a.results = c->d_results.p;
    a.paths = c->d_paths.p; a.path_cap = c->opts.path_cap;
    a.trace = c->opts.trace_cap > 0 ? c->d_trace.p : nullptr; a.trace_cap = c->opts.trace_cap;
    a.nodes = k.nodes.p; a.nodes_cap = k.nodes_cap; a.table = k.table.p; a.table_cap = k.table_cap;
    a.arena = k.arena.p; a.arena_cap = k.arena_cap; a.tmp_a = k.tmp_a.p; a.tmp_b = k.tmp_b.p; a.tmp_cap = k.tmp_cap; a.lsm_levels = k.levels;
    // longest-expected-first fetch order when the batch is larger than the resident slots
    a.order = nullptr;
    if (n_work > n_slots && n_work <= (1 << 16))
    {
        PP_CUDA(c->d_order.ensure((size_t)n_work));
        if ((int)c->kpop_group_cost.size() != c->num_groups) { c->kpop_group_cost.assign(c->num_groups, 0.0f); c->group_cost_dirty = true; }
        float known = 0.0f; int n_known = 0;
        for (float v : c->kpop_group_cost) if (v > 0.0f) { known += v; n_known++; }
        const float unknown_cost = n_known ? known / n_known : 1.0f;
        if (c->group_cost_dirty)
        {
            PP_CUDA(c->d_group_cost.ensure((size_t)c->num_groups));
            PP_CUDA(cudaMemcpyAsync(c->d_group_cost.p, c->kpop_group_cost.data(), sizeof(float) * c->num_groups, cudaMemcpyHostToDevice, c->stream));
            PP_CUDA(cudaStreamSynchronize(c->stream));
            c->group_cost_dirty = false;
        }
        pp_kpop_order_kernel<<<(n_work + 255) / 256, 256, 0, c->stream>>>(a.queries, qmap, n_work, a.field2d, a.C.N, c->d_group_cost.p,
                                                                            unknown_cost, c->d_order.p);
        c->launches += 1;
        a.order = c->d_order.p;
    }
    PP_CUDA(cudaMemsetAsync(c->d_counter, 0, sizeof(int), c->stream));
    switch (kpop_warps())
    {
        case 1:  pp_kpop_kernel<1><<<n_slots, 32, 0, c->stream>>>(a); break;
        case 2:  pp_kpop_kernel<2><<<n_slots, 64, 0, c->stream>>>(a); break;
        case 8:  pp_kpop_kernel<8><<<n_slots, 256, 0, c->stream>>>(a); break;
        default: pp_kpop_kernel<4><<<n_slots, 128, 0, c->stream>>>(a); break;
    }
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    return PP_SUCCESS;
}

int pp_batch_upload(pp_context* c, const pp_query* q, int n, const pp_search_opts* opts)
{
    if (!c || !q || n <= 0) return pp_fail(PP_ERR_INVALID, "pp_batch_upload: bad arguments");
    PP_CUDA(cudaSetDevice(c->device));
    if (c->run_pending) return pp_fail(PP_ERR_INVALID, "pp_batch_upload: the previous batch is still running (pp_batch_wait first)");
    int rc = lane_refresh(c); if (rc) return rc;
    pp_search_opts o = default_opts(opts);
    c->opts = o;
    c->h_queries.resize(n); c->h_vel.resize(n);
    for (int k = 0; k < n; k++)
    {
        c->h_vel[k] = q[k].vel;
        if (q[k].group < 0 || q[k].group >= c->num_groups) return pp_fail(PP_ERR_INVALID, "query group out of range");
        c->h_queries[k].start = pp_host_set_start(c->model.C, c->frames[q[k].group], q[k].x, q[k].y, q[k].heading, q[k].vel);
        c->h_queries[k].group = q[k].group;
        c->h_queries[k].pad = 0;
    }
    c->n_queries = n; c->batch_done = false;
    // one EXACT-mode query on a group with history enabled continues on that planner's carried 2D cache
    c->batch_hist_group = (n == 1 && o.mode == PP_MODE_EXACT && (int)c->hist.size() > q[0].group && c->hist[q[0].group].on) ? q[0].group : -1;
    int hw_slots = 0;
    rc = hw_slots_of(c, &hw_slots); if (rc) return rc;
    int want = std::min(n, hw_slots);
    if (o.max_slots > 0) want = std::min(want, o.max_slots);
    PP_CUDA(c->d_queries.ensure(n));
    PP_CUDA(c->d_results.ensure(n));
    PP_CUDA(c->d_paths.ensure((size_t)n * o.path_cap));
    if (o.trace_cap > 0) PP_CUDA(c->d_trace.ensure((size_t)n * o.trace_cap));
    if (o.mode == PP_MODE_KPOP)
    {
        if (2 * c->model.C.A + 1 > PP_K_MAXSUCC) return pp_fail(PP_ERR_INVALID, "K-POP mode supports at most 8 successors per node (num_actions <= 3)");
        if (c->parent) return pp_fail(PP_ERR_INVALID, "K-POP batches run on the parent context (lanes serve the EXACT mode)");
        if (c->wp.alloc_slots) { PP_CUDA(cudaStreamSynchronize(c->stream)); c->wp.release(); }     // the two modes do not share scratch
        // the exact 2D distance field of every group the batch touches (recomputed only after a map / goal change)
        size_t nn_ = nn_of(c);
        PP_CUDA(c->d_field2d.ensure(nn_ * c->num_groups));
        if ((int)c->field2d_valid.size() != c->num_groups) c->field2d_valid.assign(c->num_groups, 0);
        for (int k = 0; k < n; k++)
            if (!c->field2d_valid[q[k].group]) { rc = field2d_run(c, q[k].group, nullptr, nullptr); if (rc) return rc; }
        rc = kpop_hw_slots(c, &hw_slots); if (rc) return rc;
        want = std::min(n, hw_slots);
        if (o.max_slots > 0) want = std::min(want, o.max_slots);
        // generated nodes, not expansions, fill the K-POP pools: about 1.5-4 per expansion
        int nodes_cap = (int)std::min<long long>(std::max(2ll * o.max_expansions, 4096ll), 1ll << 26);
        rc = ensure_kpop(c, c->kp, want, nodes_cap, 0.6); if (rc) return rc;
        c->n_slots = std::max(std::min(c->kp.alloc_slots, want), 1);
    }
    else
    {
        if (c->kp.alloc_slots || c->kp_retry.alloc_slots) { PP_CUDA(cudaStreamSynchronize(c->stream)); c->kp.release(); c->kp_retry.release(); }
        rc = ensure_work(c, c->wp, want, o.max_expansions, o.max_open, o.max_open2d, true); if (rc) return rc;
        c->n_slots = std::max(std::min(c->wp.alloc_slots, want), 1);
    }
    PP_CUDA(cudaMemcpyAsync(c->d_queries.p, c->h_queries.data(), sizeof(PPQuery) * n, cudaMemcpyHostToDevice, c->stream));
    rc = sync_groups(c); if (rc) return rc;
    c->exact_order_valid = false;
    if (o.mode == PP_MODE_EXACT && n > c->n_slots && n <= (1 << 16))
    {
        // More queries than resident slots: fetch the ones expected to run longest first (unreachable goal in the exact 2D field,
        // then by distance), so that they run beside the bulk instead of after it.  Scheduling only -- no result depends on it.
        size_t nn_ = nn_of(c);
        PP_CUDA(c->d_field2d.ensure(nn_ * c->num_groups));
        if ((int)c->field2d_valid.size() != c->num_groups) c->field2d_valid.assign(c->num_groups, 0);
        for (int k = 0; k < n; k++)
            if (!c->field2d_valid[q[k].group]) { rc = field2d_run(c, q[k].group, nullptr, nullptr); if (rc) return rc; }
        PP_CUDA(c->d_exact_order.ensure((size_t)n));
        // one warp per CTA: fits into a single freed warp slot while other lanes' search kernels occupy the SMs
        pp_kpop_order_kernel<<<(n + 31) / 32, 32, 0, c->stream>>>(c->d_queries.p, nullptr, n, c->d_field2d.p, c->model.C.N, nullptr, 1.0f,
                                                                   c->d_exact_order.p);
        c->launches += 1;
        PP_CUDA(cudaGetLastError());
        c->exact_order_valid = true;
    }
    PP_CUDA(cudaStreamSynchronize(c->stream));
    return PP_SUCCESS;
}

static int launch_search(pp_context* c, const WorkPools& w, const pp_search_opts& o, int n_slots, const int* qmap, int n_work)
{
    PPBatchArgs a;
    fill_args(c, w, o, n_slots, qmap, n_work, a);
    if (!qmap && c->exact_order_valid && n_work == c->n_queries) a.order = c->d_exact_order.p;
    PP_CUDA(cudaMemsetAsync(c->d_counter, 0, sizeof(int), c->stream));
    int rc = arena_reset(c, c->wp); if (rc) return rc;
    int blocks = (n_slots + PP_SEARCH_WARPS - 1) / PP_SEARCH_WARPS;
    pp_search_kernel<<<blocks, PP_SEARCH_WARPS * 32, 0, c->stream>>>(a);
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    return PP_SUCCESS;
}

// First pass of the uploaded batch, enqueued on the context's stream; returns at once.  pp_batch_wait completes it.
int pp_batch_run_async(pp_context* c)
{
    if (!c || c->n_queries <= 0) return pp_fail(PP_ERR_INVALID, "pp_batch_run: nothing uploaded");
    if (c->run_pending) return pp_fail(PP_ERR_INVALID, "pp_batch_run_async: a batch is already running");
    PP_CUDA(cudaSetDevice(c->device));
    const int n = c->n_queries;
    const bool kmode = (c->opts.mode == PP_MODE_KPOP);
    const int hist_g = kmode ? -1 : c->batch_hist_group;
    int rc = 0;
    PP_CUDA(cudaEventRecord(c->ev_run0, c->stream));
    if (hist_g >= 0) { rc = hist_copy(c, hist_g, false); if (rc) return rc; }
    rc = kmode ? launch_kpop(c, c->kp, c->n_slots, nullptr, n) : launch_search(c, c->wp, c->opts, c->n_slots, nullptr, n);
    if (rc) return rc;
    PP_CUDA(cudaEventRecord(c->ev_run1, c->stream));
    c->run_pending = true; c->batch_done = false;
    return PP_SUCCESS;
}

// Waits for the first pass and, because the reference's containers are unbounded, re-runs what still hit a cap: queries
// that met the caller's max_expansions / max_open / max_open2d get caps 8x larger (up to 3 escalations); queries that found
// the arena empty are re-run with fewer resident neighbours.  What still overflows stays flagged in pp_result.status.
// kernel_ms = device time of all passes (CUDA events on the context's stream).
int pp_batch_wait(pp_context* c, float* kernel_ms)
{
    if (!c || !c->run_pending) return pp_fail(PP_ERR_INVALID, "pp_batch_wait: nothing is running");
    PP_CUDA(cudaSetDevice(c->device));
    c->run_pending = false;
    const int n = c->n_queries;
    const bool kmode = (c->opts.mode == PP_MODE_KPOP);
    const int hist_g = kmode ? -1 : c->batch_hist_group;
    int rc = 0;
    PP_CUDA(cudaStreamSynchronize(c->stream));
    float total_ms = 0.0f;
    PP_CUDA(cudaEventElapsedTime(&total_ms, c->ev_run0, c->ev_run1));
    c->retried = 0;
    pp_search_opts o = c->opts;
    long long kp_nodes = c->kp.nodes_cap;
    std::vector<PPResult> r(n);
    const int overflow = PP_STATUS_OPEN_OVERFLOW | PP_STATUS_CLOSED_OVERFLOW | PP_STATUS_OPEN2D_OVERFLOW | PP_STATUS_ARENA_EXHAUSTED;
    int slots_now = c->n_slots;
    for (int level = 0; level < 3; level++)
    {
        PP_CUDA(cudaMemcpyAsync(r.data(), c->d_results.p, sizeof(PPResult) * n, cudaMemcpyDeviceToHost, c->stream));
        PP_CUDA(cudaStreamSynchronize(c->stream));
        if (kmode && level == 0)
        {
            // remember how hard every group was (mean iterations per query) for the next batch's launch order
            if ((int)c->kpop_group_cost.size() != c->num_groups) c->kpop_group_cost.assign(c->num_groups, 0.0f);
            std::vector<double> sum(c->num_groups, 0.0); std::vector<int> cnt(c->num_groups, 0);
            for (int k = 0; k < n; k++) { int g = c->h_queries[k].group; sum[g] += (double)r[k].n_lazy_searches; cnt[g]++; }
            for (int g = 0; g < c->num_groups; g++)
                if (cnt[g] > 0)
                {
                    float now = (float)(sum[g] / cnt[g]) + 1.0f, old = c->kpop_group_cost[g];
                    float upd = old > 0.0f ? 0.5f * old + 0.5f * now : now;
                    if (upd != old) { c->kpop_group_cost[g] = upd; c->group_cost_dirty = true; }
                }
        }
        std::vector<int> redo;
        bool arena_short = false;
        for (int k = 0; k < n; k++)
            if (r[k].status & overflow) { redo.push_back(k); if (r[k].status & PP_STATUS_ARENA_EXHAUSTED) arena_short = true; }
        if (redo.empty()) break;
        if (level == 0) c->retried = (int)redo.size();
        PP_CUDA(c->d_qmap.ensure(redo.size()));
        PP_CUDA(cudaMemcpyAsync(c->d_qmap.p, redo.data(), sizeof(int) * redo.size(), cudaMemcpyHostToDevice, c->stream));
        PP_CUDA(cudaEventRecord(c->ev_re0, c->stream));
        if (kmode)
        {
            int hw_slots = 0;
            rc = kpop_hw_slots(c, &hw_slots); if (rc) return rc;
            int want = std::min((int)redo.size(), hw_slots);
            kp_nodes = std::min<long long>(kp_nodes * 8, 1ll << 24);
            rc = ensure_kpop(c, c->kp_retry, want, (int)kp_nodes, 0.85); if (rc) return rc;
            int slots = std::max(std::min(c->kp_retry.alloc_slots, want), 1);
            rc = launch_kpop(c, c->kp_retry, slots, c->d_qmap.p, (int)redo.size()); if (rc) return rc;
        }
        else
        {
            o.max_expansions = (int)std::min<long long>((long long)o.max_expansions * 8, 1 << 26);
            o.max_open = (int)std::min<long long>((long long)o.max_open * 8, 1 << 25);
            o.max_open2d = (int)std::min<long long>((long long)o.max_open2d * 4, 1 << 22);
            if (arena_short) { slots_now = std::max(1, slots_now / 8); rc = arena_grow(c, c->wp); if (rc) return rc; }
            int slots = std::max(1, std::min((int)redo.size(), slots_now));
            if (hist_g >= 0) { rc = hist_copy(c, hist_g, true); if (rc) return rc; }     // the aborted attempt never happened
            rc = launch_search(c, c->wp, o, slots, c->d_qmap.p, (int)redo.size()); if (rc) return rc;
        }
        PP_CUDA(cudaEventRecord(c->ev_re1, c->stream));
        PP_CUDA(cudaStreamSynchronize(c->stream));
        float ms = 0.0f;
        PP_CUDA(cudaEventElapsedTime(&ms, c->ev_re0, c->ev_re1));
        total_ms += ms;
        if (level == 2)
        {
            // out of escalations: if the last attempt still overflowed on a carried cache, that attempt never happened either
            PP_CUDA(cudaMemcpyAsync(r.data(), c->d_results.p, sizeof(PPResult) * n, cudaMemcpyDeviceToHost, c->stream));
            PP_CUDA(cudaStreamSynchronize(c->stream));
            bool still = false;
            for (int k : redo) if (r[k].status & overflow) still = true;
            if (still && hist_g >= 0) { rc = hist_copy(c, hist_g, true); if (rc) return rc; PP_CUDA(cudaStreamSynchronize(c->stream)); }
        }
    }
    if (kernel_ms) *kernel_ms = total_ms;
    c->batch_done = true;
    return PP_SUCCESS;
}

int pp_batch_run(pp_context* c, float* kernel_ms)
{
    int rc = pp_batch_run_async(c); if (rc) return rc;
    return pp_batch_wait(c, kernel_ms);
}

int pp_batch_fetch(pp_context* c, pp_result* results, float* paths_xyh, float* curvature, pp_pop* trace)
{
    if (!c || c->n_queries <= 0 || !results) return pp_fail(PP_ERR_INVALID, "pp_batch_fetch: bad arguments");
    PP_CUDA(cudaSetDevice(c->device));
    int n = c->n_queries, pc = c->opts.path_cap;
    std::vector<PPResult> r(n);
    PP_CUDA(cudaMemcpyAsync(r.data(), c->d_results.p, sizeof(PPResult) * n, cudaMemcpyDeviceToHost, c->stream));
    std::vector<PPPathPt> pts;
    if (paths_xyh)
    {
        pts.resize((size_t)n * pc);
        PP_CUDA(cudaMemcpyAsync(pts.data(), c->d_paths.p, sizeof(PPPathPt) * (size_t)n * pc, cudaMemcpyDeviceToHost, c->stream));
    }
    if (trace && c->opts.trace_cap > 0)
        PP_CUDA(cudaMemcpyAsync(trace, c->d_trace.p, sizeof(PPPop) * (size_t)n * c->opts.trace_cap, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < n; k++)
    {
        pp_result& o = results[k];
        o.success = r[k].success; o.status = r[k].status; o.cost = r[k].cost; o.n_pops = r[k].n_pops;
        o.n_pops_bin_oob = r[k].n_pops_bin_oob; o.n_chain = r[k].n_chain; o.n_dubins = r[k].n_dubins;
        o.n_lazy_searches = r[k].n_lazy_searches; o.n_lazy_pops = r[k].n_lazy_pops; o.max_open = r[k].max_open;
        o.n_closed = r[k].n_closed; o.n_path = 0;
        if (!o.success || !paths_xyh) continue;
        // HybridAStar::reconstruct_path (HybridAStar.cpp:208-262): reversed Dubins samples, then the parent chain,
        // rotated back to the world frame; curvature shifted by one point
        const PPHostFrame& fr = c->frames[c->h_queries[k].group];
        const PPPathPt* src = pts.data() + (size_t)k * pc;
        int nd = std::min(r[k].n_dubins, pc), nc = std::min(r[k].n_chain, std::max(pc - nd, 0));
        int total = nd + nc;
        float* P = paths_xyh + (size_t)k * pc * 3;
        float* K = curvature ? curvature + (size_t)k * pc : nullptr;
        float prev_curv = 0.0f;
        for (int t = 0; t < total; t++)
        {
            const PPPathPt& s = (t < nd) ? src[nd - 1 - t] : src[t];
            pp_host_to_world(fr, s.x, s.y, s.heading, P[3 * t], P[3 * t + 1], P[3 * t + 2]);
            if (K) K[t] = prev_curv;
            prev_curv = s.curvature;
        }
        o.n_path = total;
    }
    return PP_SUCCESS;
}

int pp_find_path_batch(pp_context* c, const pp_query* q, int n, const pp_search_opts* opts, pp_result* results,
                       float* paths_xyh, float* curvature, pp_pop* trace)
{
    int rc = pp_batch_upload(c, q, n, opts); if (rc) return rc;
    rc = pp_batch_run(c, nullptr); if (rc) return rc;
    return pp_batch_fetch(c, results, paths_xyh, curvature, trace);
}

// ---- velocity profile / trajectory ----
static PPVelLimits limits_of(const pp_velocity_limits* l)
{
    PPVelLimits L; L.max_velocity = l->max_velocity; L.coast_velocity = l->coast_velocity; L.max_lat_acc = l->max_lat_acc;
    L.max_long_acc = l->max_long_acc; L.max_long_dec = l->max_long_dec;
    return L;
}

int pp_velocity_profile_batch(pp_context* c, const pp_velocity_limits* lim, const float* paths_xy, const float* curvature, const int* counts,
                              int n, int cap, const float* vel_init, const float* max_velocity_curr, const int* flags, float* velocity,
                              int* feasible)
{
    if (!c || !lim || !paths_xy || !curvature || !counts || !vel_init || !velocity || !feasible || cap < 1)
        return pp_fail(PP_ERR_INVALID, "pp_velocity_profile_batch: bad arguments");
    if (n <= 0) return PP_SUCCESS;
    PP_CUDA(cudaSetDevice(c->device));
    const size_t nc = (size_t)n * cap;
    // one float arena: xy | curvature | velocity | v2 | vel_init | vcap ; one int arena: counts | flags | feasible
    PP_CUDA(c->d_traj.ensure(nc * 5 + 2 * (size_t)n));
    PP_CUDA(c->d_traj_int.ensure(3 * (size_t)n));
    float* d_xy = c->d_traj.p; float* d_curv = d_xy + 2 * nc; float* d_vel = d_curv + nc; float* d_v2 = d_vel + nc;
    float* d_vi = d_v2 + nc; float* d_vc = d_vi + n;
    int* d_cnt = c->d_traj_int.p; int* d_fl = d_cnt + n; int* d_ok = d_fl + n;
    PP_CUDA(cudaMemcpyAsync(d_xy, paths_xy, sizeof(float) * 2 * nc, cudaMemcpyHostToDevice, c->stream));
    PP_CUDA(cudaMemcpyAsync(d_curv, curvature, sizeof(float) * nc, cudaMemcpyHostToDevice, c->stream));
    PP_CUDA(cudaMemcpyAsync(d_cnt, counts, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    PP_CUDA(cudaMemcpyAsync(d_vi, vel_init, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    if (max_velocity_curr) PP_CUDA(cudaMemcpyAsync(d_vc, max_velocity_curr, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    if (flags) PP_CUDA(cudaMemcpyAsync(d_fl, flags, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    PP_CUDA(cudaMemsetAsync(d_vel, 0, sizeof(float) * nc, c->stream));
    pp_velocity_profile_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(limits_of(lim), d_xy, d_curv, d_cnt, n, cap, d_vi,
                                                                       max_velocity_curr ? d_vc : nullptr, flags ? d_fl : nullptr, d_vel, d_v2, d_ok);
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    PP_CUDA(cudaMemcpyAsync(velocity, d_vel, sizeof(float) * nc, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaMemcpyAsync(feasible, d_ok, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    return PP_SUCCESS;
}

int pp_trajectory_batch(pp_context* c, const pp_velocity_limits* lim, const float* max_velocity_curr, const int* stop_at_goal, float* traj,
                        int* n_samples, int* feasible, float* kernel_ms)
{
    if (!c || !lim || !traj || !n_samples || !feasible) return pp_fail(PP_ERR_INVALID, "pp_trajectory_batch: bad arguments");
    if (c->n_queries <= 0 || (int)c->h_vel.size() != c->n_queries || !c->batch_done)
        return pp_fail(PP_ERR_INVALID, "pp_trajectory_batch: run a batch first (pp_find_path_batch or pp_batch_upload + pp_batch_run)");
    PP_CUDA(cudaSetDevice(c->device));
    const int n = c->n_queries, pc = c->opts.path_cap;
    // world frames of the groups: cos / sin of -grid_heading with the host libm, once per frame (pp_host_to_world does it per point)
    std::vector<PPWorldFrame> wf(c->num_groups);
    for (int g = 0; g < c->num_groups; g++)
    {
        const PPHostFrame& fr = c->frames[g];
        PPWorldFrame& F = wf[g];
        F.goal_gx = fr.F.goal_x; F.goal_gy = fr.F.goal_y; F.goal_wx = fr.goal_world[0]; F.goal_wy = fr.goal_world[1];
        F.angle = -fr.grid_heading; F.c = std::cos(F.angle); F.s = std::sin(F.angle);
    }
    PP_CUDA(c->d_wframes.ensure(wf.size()));
    PP_CUDA(cudaMemcpyAsync(c->d_wframes.p, wf.data(), sizeof(PPWorldFrame) * wf.size(), cudaMemcpyHostToDevice, c->stream));
    PP_CUDA(c->d_traj.ensure((size_t)n * 4 * pc));
    PP_CUDA(c->d_traj_tmp.ensure((size_t)n * 2 * pc));
    PP_CUDA(c->d_vel_in.ensure(2 * (size_t)n));
    PP_CUDA(c->d_traj_int.ensure(3 * (size_t)n));
    PP_CUDA(cudaMemcpyAsync(c->d_vel_in.p, c->h_vel.data(), sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    if (max_velocity_curr) PP_CUDA(cudaMemcpyAsync(c->d_vel_in.p + n, max_velocity_curr, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    if (stop_at_goal) PP_CUDA(cudaMemcpyAsync(c->d_traj_int.p, stop_at_goal, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    PPTrajArgs a;
    a.L = limits_of(lim); a.results = c->d_results.p; a.paths = c->d_paths.p; a.path_cap = pc; a.queries = c->d_queries.p;
    a.frames = c->d_wframes.p; a.vel_init = c->d_vel_in.p; a.vcap = max_velocity_curr ? c->d_vel_in.p + n : nullptr;
    a.stop = stop_at_goal ? c->d_traj_int.p : nullptr; a.n = n; a.traj = c->d_traj.p; a.tmp = c->d_traj_tmp.p;
    a.n_samples = c->d_traj_int.p + n; a.feasible = c->d_traj_int.p + 2 * (size_t)n;
    int blocks = std::min((n + 3) / 4, 16 * c->sm_count);
    PP_CUDA(cudaEventRecord(c->ev0, c->stream));
    pp_trajectory_kernel<<<blocks, 128, 0, c->stream>>>(a);
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    PP_CUDA(cudaEventRecord(c->ev1, c->stream));
    PP_CUDA(cudaMemcpyAsync(traj, c->d_traj.p, sizeof(float) * (size_t)n * 4 * pc, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaMemcpyAsync(n_samples, a.n_samples, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaMemcpyAsync(feasible, a.feasible, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    if (kernel_ms) PP_CUDA(cudaEventElapsedTime(kernel_ms, c->ev0, c->ev1));
    return PP_SUCCESS;
}

static int lazy_run(pp_context* c, int g, const int* ij, int n, float* out, int restart)
{
    int rc = check_group(c, g); if (rc) return rc;
    if (n <= 0) return PP_SUCCESS;
    PP_CUDA(cudaSetDevice(c->device));
    pp_search_opts o = default_opts(nullptr);      // local: the uploaded batch keeps its own options (and buffers sized for them)
    o.max_expansions = 16; o.max_open = 16; o.max_open2d = 1 << 18;
    if (c->wp_lazy.alloc_slots < 1 || c->lazy_group != g) restart = 1;      // no cache yet (or another planner's)
    rc = ensure_work(c, c->wp_lazy, 1, o.max_expansions, o.max_open, o.max_open2d, false); if (rc) return rc;
    c->lazy_group = g;
    rc = sync_groups(c); if (rc) return rc;
    PPBatchArgs a;
    const int keep_hist = c->batch_hist_group;
    c->batch_hist_group = -1;
    fill_args(c, c->wp_lazy, o, 1, nullptr, 0, a);
    c->batch_hist_group = keep_hist;
    PP_CUDA(c->s0.ensure(sizeof(int) * 2 * n));
    PP_CUDA(c->s1.ensure(sizeof(float) * n));
    PP_CUDA(c->s2.ensure(sizeof(int)));
    PP_CUDA(cudaMemcpyAsync(c->s0.p, ij, sizeof(int) * 2 * n, cudaMemcpyHostToDevice, c->stream));
    pp_lazy_astar_kernel<<<1, 32, 0, c->stream>>>(a, c->groups[g], (const int*)c->s0.p, n, (float*)c->s1.p, (int*)c->s2.p,
                                                 restart, c->d_lazy_sid);
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    int status = 0;
    PP_CUDA(cudaMemcpyAsync(out, c->s1.p, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaMemcpyAsync(&status, c->s2.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    if (status) return pp_fail(PP_ERR_CAPACITY, "lazy A*: 2D open-list pool exhausted");
    return PP_SUCCESS;
}

// AStar::reset() (lib/AStar.cpp:56-60) on the stand-alone cache of pp_astar_lazy_*: only the visited flags go, the node costs of
// earlier searches stay
int pp_astar_lazy_reset(pp_context* c, int g)
{
    int rc = check_group(c, g); if (rc) return rc;
    if (c->wp_lazy.alloc_slots < 1 || c->lazy_group != g) return PP_SUCCESS;      // no cache yet: nothing visited
    PP_CUDA(cudaSetDevice(c->device));
    int nn = (int)nn_of(c);
    pp_hist_reset_kernel<<<std::min((nn + 255) / 256, 4 * c->sm_count), 256, 0, c->stream>>>(c->wp_lazy.cell_state.p, nn);
    c->launches += 1;
    PP_CUDA(cudaGetLastError());
    return PP_SUCCESS;
}

int pp_astar_lazy_batch(pp_context* c, int g, const int* ij, int n, float* out) { return lazy_run(c, g, ij, n, out, 1); }
int pp_astar_lazy_continue(pp_context* c, int g, const int* ij, int n, float* out) { return lazy_run(c, g, ij, n, out, 0); }

int pp_override_dubins(pp_context* c, float r_min, float step_size)
{
    if (!c || !(r_min > 0.0f) || !(step_size > 0.0f)) return pp_fail(PP_ERR_INVALID, "pp_override_dubins: bad arguments");
    c->model.C.r_min = r_min;                     // Dubins::Dubins(r_min, step_size), Dubins.cpp:7-16
    c->model.C.step = step_size;
    c->model.C.ang_step = step_size / r_min;
    return PP_SUCCESS;
}

int pp_clear_obstacles(pp_context* c, int g)
{
    int rc = check_group(c, g); if (rc) return rc;
    rc = lane_guard(c, "pp_clear_obstacles"); if (rc) return rc;
    map_changed(c, g);
    PP_CUDA(cudaSetDevice(c->device));
    PP_CUDA(cudaMemsetAsync(c->d_maps + nn_of(c) * g, 0, sizeof(float) * nn_of(c), c->stream));
    return PP_SUCCESS;
}

// ---- heuristic fields ------------------------------------------------------------------------------------
static int field2d_run(pp_context* c, int g, int* sweeps, float* ms)
{
    const PPConsts& C = c->model.C;
    size_t nn = nn_of(c);
    int T = (C.N + 31) / 32;
    PP_CUDA(c->d_field2d.ensure(nn * c->num_groups));
    if ((int)c->field2d_valid.size() != c->num_groups) c->field2d_valid.assign(c->num_groups, 0);
    PP_CUDA(c->d_f2d_work.ensure(nn));
    PP_CUDA(c->d_f2d_flags.ensure((size_t)2 * T * T));
    PP_CUDA(c->d_f2d_ctl.ensure(4));
    float diag = C.n_act2d == 8 ? C.act_cost[1] : 0.0f;
    PP_CUDA(cudaEventRecord(c->ev0, c->stream));
    int sw = 0;
    int e = pp_launch_field2d(c->stream, c->groups[g].map, C.N, C.log_thr, C.act_cost[0], diag, C.n_act2d == 8 ? 1 : 0,
                              c->frames[g].F.goal_ci, c->frames[g].F.goal_cj, c->d_f2d_work.p, c->d_f2d_flags.p, c->d_f2d_ctl.p,
                              c->d_field2d.p + nn * g, c->sm_count, &sw, &c->launches);
    if (e != 0) return pp_fail(PP_ERR_CUDA, std::string("field2d: ") + cudaGetErrorString((cudaError_t)e));
    PP_CUDA(cudaEventRecord(c->ev1, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    if (ms) PP_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    if (sweeps) *sweeps = sw;
    c->field2d_valid[g] = 1;
    return PP_SUCCESS;
}

int pp_heuristic_field_2d(pp_context* c, int g, float* out_nn, int* sweeps, float* ms)
{
    int rc = check_group(c, g); if (rc) return rc;
    PP_CUDA(cudaSetDevice(c->device));
    rc = field2d_run(c, g, sweeps, ms); if (rc) return rc;
    if (out_nn)
    {
        PP_CUDA(cudaMemcpyAsync(out_nn, c->d_field2d.p + nn_of(c) * g, sizeof(float) * nn_of(c), cudaMemcpyDeviceToHost, c->stream));
        PP_CUDA(cudaStreamSynchronize(c->stream));
    }
    return PP_SUCCESS;
}

int pp_heuristic_field_3d(pp_context* c, int g, int use_h2d, float* out_nnb, float* ms)
{
    int rc = check_group(c, g); if (rc) return rc;
    PP_CUDA(cudaSetDevice(c->device));
    const PPConsts& C = c->model.C;
    size_t nn = nn_of(c), total = nn * C.bins;
    if (use_h2d && ((int)c->field2d_valid.size() != c->num_groups || !c->field2d_valid[g]))
    { rc = field2d_run(c, g, nullptr, nullptr); if (rc) return rc; }
    PP_CUDA(c->d_dubins_field.ensure(total));
    const PPFrame& F = c->frames[g].F;
    PP_CUDA(cudaEventRecord(c->ev0, c->stream));
    int e = pp_launch_dubins_field(c->stream, use_h2d ? c->d_field2d.p + nn * g : nullptr, c->d_dubins_field.p, C.N, C.bins, C.res,
                                   C.precision, C.r_min, F.goal_x, F.goal_y, F.goal_h, c->sm_count, &c->launches);
    if (e != 0) return pp_fail(PP_ERR_CUDA, std::string("dubins field: ") + cudaGetErrorString((cudaError_t)e));
    PP_CUDA(cudaEventRecord(c->ev1, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    if (ms) PP_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    if (out_nnb)
    {
        PP_CUDA(cudaMemcpyAsync(out_nnb, c->d_dubins_field.p, sizeof(float) * total, cudaMemcpyDeviceToHost, c->stream));
        PP_CUDA(cudaStreamSynchronize(c->stream));
    }
    return PP_SUCCESS;
}

// ---- map replication (north_star: "the map replicated by an NCCL broadcast over NVLink after each update") ----
int pp_comm_unique_id(void* id128)
{
    if (!id128) return pp_fail(PP_ERR_INVALID, "pp_comm_unique_id: null");
    std::string err;
    PPNccl* api = nccl_api(err);
    if (!api) return pp_fail(PP_ERR_INVALID, "pp_comm_unique_id: " + err);
    int e = api->GetUniqueId(id128);
    if (e) return pp_fail(PP_ERR_CUDA, std::string("ncclGetUniqueId: ") + (api->GetErrorString ? api->GetErrorString(e) : "error"));
    return PP_SUCCESS;
}

int pp_comm_init(pp_context* c, int nranks, int rank, const void* id128)
{
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return pp_fail(PP_ERR_INVALID, "pp_comm_init: bad arguments");
    if (c->nccl_comm) return pp_fail(PP_ERR_INVALID, "pp_comm_init: communicator already initialised");
    std::string err;
    PPNccl* api = nccl_api(err);
    if (!api) return pp_fail(PP_ERR_INVALID, "pp_comm_init: " + err);
    PP_CUDA(cudaSetDevice(c->device));
    PPNcclIdByValue id;
    std::memcpy(&id, id128, sizeof(id));
    void* comm = nullptr;
    int e = api->CommInitRank(&comm, nranks, id, rank);
    if (e) return pp_fail(PP_ERR_CUDA, std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(e) : "error"));
    c->nccl_comm = comm; c->comm_rank = rank; c->comm_size = nranks;
    return PP_SUCCESS;
}

// Replaces, on every rank but `root`, the maps of groups [first_group, first_group + n_groups) by root's, straight into the
// context's map storage on the context's stream; the exact 2D field and the scheduling hints of those groups are dropped.
int pp_broadcast_maps(pp_context* c, int first_group, int n_groups, int root)
{
    if (!c || !c->nccl_comm) return pp_fail(PP_ERR_INVALID, "pp_broadcast_maps: pp_comm_init first");
    int rc = lane_guard(c, "pp_broadcast_maps"); if (rc) return rc;
    if (n_groups < 1 || first_group < 0 || first_group + n_groups > c->num_groups || root < 0 || root >= c->comm_size)
        return pp_fail(PP_ERR_INVALID, "pp_broadcast_maps: bad group range or root");
    std::string err;
    PPNccl* api = nccl_api(err);
    if (!api) return pp_fail(PP_ERR_INVALID, "pp_broadcast_maps: " + err);
    PP_CUDA(cudaSetDevice(c->device));
    float* p = c->d_maps + nn_of(c) * first_group;
    int e = api->Broadcast(p, p, nn_of(c) * (size_t)n_groups, 7 /* ncclFloat32 */, root, c->nccl_comm, c->stream);
    if (e) return pp_fail(PP_ERR_CUDA, std::string("ncclBroadcast: ") + (api->GetErrorString ? api->GetErrorString(e) : "error"));
    for (int g = first_group; g < first_group + n_groups; g++) map_changed(c, g);
    return PP_SUCCESS;
}

int pp_comm_destroy(pp_context* c)
{
    if (!c || !c->nccl_comm) return PP_SUCCESS;
    std::string err;
    PPNccl* api = nccl_api(err);
    if (api) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); api->CommDestroy(c->nccl_comm); }
    c->nccl_comm = nullptr; c->comm_rank = -1; c->comm_size = 0;
    return PP_SUCCESS;
}

unsigned long long pp_kernel_launches(pp_context* c) { return c ? c->launches : 0ull; }
int pp_batch_retried(pp_context* c) { return c ? c->retried : 0; }

// CUDA-event bracket on the context's stream (the stream every kernel of this context is launched on)
int pp_timer_begin(pp_context* c)
{
    if (!c) return pp_fail(PP_ERR_INVALID, "null context");
    PP_CUDA(cudaSetDevice(c->device));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    PP_CUDA(cudaEventRecord(c->ev0, c->stream));
    return PP_SUCCESS;
}

int pp_timer_end(pp_context* c, float* ms)
{
    if (!c || !ms) return pp_fail(PP_ERR_INVALID, "null argument");
    PP_CUDA(cudaSetDevice(c->device));
    PP_CUDA(cudaEventRecord(c->ev1, c->stream));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    PP_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return PP_SUCCESS;
}

#ifdef PP_PROFILE
// development variant only (lib/libpp_b200_prof.so): per-phase SM cycles summed over all queries; reset on read
int pp_profile_read(pp_context* c, unsigned long long* out16)
{
    if (!c) return pp_fail(PP_ERR_INVALID, "null context");
    PP_CUDA(cudaSetDevice(c->device));
    PP_CUDA(cudaStreamSynchronize(c->stream));
    PP_CUDA(cudaMemcpyFromSymbol(out16, pp_prof_acc, sizeof(unsigned long long) * 16));
    unsigned long long z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    PP_CUDA(cudaMemcpyToSymbol(pp_prof_acc, z, sizeof(z)));
    return PP_SUCCESS;
}
#endif

} // extern "C"
