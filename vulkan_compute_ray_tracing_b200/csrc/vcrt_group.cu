// vcrt_group.cu -- multi-GPU groups of the C ABI (include/vcrt.h: vcrt_group_*): the scene is replicated, every GPU renders
// its share of one frame (interleaved 32x32 tiles, or a slice of the samples), and ONE collective at the end combines the
// shares -- NCCL all-gather of packed rgba8 tiles, or NCCL sum-reduce of the f32 accumulation buffers -- enqueued on the
// contexts' own render streams, so that nothing but the final read-back synchronises with the host.
//
// The reference has no multi-device path (one VkDevice, one queue: VulkanApplicationContext.cpp:95-119) and its shader
// invocations never communicate (ray-trace-compute.comp has no shared memory, atomics or barriers), so pixels and samples
// partition freely.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: inside a PyTorch process that is the library torch already loaded),
// so libvcrt.so itself links against nothing but the CUDA runtime and a host without NCCL can still use single-GPU contexts.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "vcrt_ctx.h"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("failed to load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return; }
        bool ok = true;
        auto sym = [&](const char* n) { void* p = dlsym(api.handle, n); if (!p) { ok = false; api.error = std::string("failed to resolve ") + n; } return p; };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.Reduce = (decltype(api.Reduce))sym("ncclReduce");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        if (!ok) { dlclose(api.handle); api.handle = nullptr; }
    });
    return &api;
}

thread_local std::string g_group_create_error;

}  // namespace

struct vcrt_group {
    int world = 1;                       // ranks in the group
    int first_rank = 0;                  // rank of members[0]
    bool owns_ctx = false;               // local groups create (and destroy) their contexts
    std::vector<vcrt_ctx*> members;      // the contexts this process drives (local group: all of them; rank group: one)
    std::vector<ncclComm_t> comms;       // one communicator per member (empty when world == 1)
    std::vector<DevBuf> mine, everyone;  // per member: packed tiles of this rank / of all ranks (pre-allocated, reused)
    std::string error;
};

namespace {

int gfail(vcrt_group* g, int code, const std::string& msg) {
    if (g) g->error = msg; else g_group_create_error = msg;
    return code;
}

int nccl_fail(vcrt_group* g, ncclResult_t r, const char* what) {
    NcclApi* api = nccl_api();
    return gfail(g, VCRT_ERR_CUDA, std::string("failed to ") + what + ": NCCL " + (api->GetErrorString ? api->GetErrorString(r) : "error"));
}

// every member's error text is the group's as well
int member_fail(vcrt_group* g, vcrt_ctx* c, int rc) {
    g->error = vcrt_last_error(c);
    return rc;
}

uint32_t max_owned_tiles(uint32_t W, uint32_t H, uint32_t world) {
    const uint32_t tiles = ((W + 31) / 32) * ((H + 31) / 32);
    return (tiles + world - 1) / world;
}

}  // namespace

namespace vcrt {
cudaError_t launch_unpack_all_tiles(void* image, const void* gathered, int elem_bytes, uint32_t W, uint32_t H, uint32_t world, uint32_t tiles_per_rank, cudaStream_t stream);
}

extern "C" {

const char* vcrt_group_last_error(const vcrt_group* g) { return g ? g->error.c_str() : g_group_create_error.c_str(); }

int vcrt_group_unique_id(void* id) {
    if (!id) return gfail(nullptr, VCRT_ERR_INVALID, "vcrt_group_unique_id: NULL argument");
    static_assert(sizeof(ncclUniqueId) == VCRT_GROUP_ID_BYTES, "ncclUniqueId size");
    NcclApi* api = nccl_api();
    if (!api->handle) return gfail(nullptr, VCRT_ERR_STATE, api->error);
    ncclUniqueId u;
    ncclResult_t r = api->GetUniqueId(&u);
    if (r != ncclSuccess) return nccl_fail(nullptr, r, "create a NCCL unique id");
    std::memcpy(id, &u, sizeof u);
    return VCRT_OK;
}

int vcrt_group_create_local(int n_devices, const int* devices, vcrt_group** out) {
    if (!out) return gfail(nullptr, VCRT_ERR_INVALID, "vcrt_group_create_local: out is NULL");
    *out = nullptr;
    if (n_devices < 1 || n_devices > 64) return gfail(nullptr, VCRT_ERR_INVALID, "vcrt_group_create_local: n_devices must be 1..64");
    vcrt_group* g = new (std::nothrow) vcrt_group();
    if (!g) return gfail(nullptr, VCRT_ERR_NOMEM, "vcrt_group_create_local: out of host memory");
    g->world = n_devices; g->first_rank = 0; g->owns_ctx = true;
    std::vector<int> devs(n_devices);
    for (int i = 0; i < n_devices; ++i) devs[i] = devices ? devices[i] : i;
    for (int i = 0; i < n_devices; ++i) {
        vcrt_ctx* c = nullptr;
        int rc = vcrt_create(devs[i], &c);
        if (rc) { g_group_create_error = vcrt_last_error(nullptr); vcrt_group_destroy(g); return rc; }
        g->members.push_back(c);
    }
    if (n_devices > 1) {
        NcclApi* api = nccl_api();
        if (!api->handle) { std::string e = api->error; vcrt_group_destroy(g); return gfail(nullptr, VCRT_ERR_STATE, e); }
        g->comms.resize(n_devices);
        ncclResult_t r = api->CommInitAll(g->comms.data(), n_devices, devs.data());
        if (r != ncclSuccess) { g->comms.clear(); vcrt_group_destroy(g); return nccl_fail(nullptr, r, "create NCCL communicators (ncclCommInitAll)"); }
    }
    g->mine.resize(n_devices); g->everyone.resize(n_devices);
    *out = g;
    return VCRT_OK;
}

int vcrt_group_create_rank(vcrt_ctx* ctx, const void* id, int rank, int world, vcrt_group** out) {
    if (!out) return gfail(nullptr, VCRT_ERR_INVALID, "vcrt_group_create_rank: out is NULL");
    *out = nullptr;
    if (!ctx || world < 1 || rank < 0 || rank >= world || (world > 1 && !id)) return gfail(nullptr, VCRT_ERR_INVALID, "vcrt_group_create_rank: bad argument");
    vcrt_group* g = new (std::nothrow) vcrt_group();
    if (!g) return gfail(nullptr, VCRT_ERR_NOMEM, "vcrt_group_create_rank: out of host memory");
    g->world = world; g->first_rank = rank; g->owns_ctx = false;
    g->members.push_back(ctx);
    if (world > 1) {
        NcclApi* api = nccl_api();
        if (!api->handle) { std::string e = api->error; delete g; return gfail(nullptr, VCRT_ERR_STATE, e); }
        cudaError_t ce = cudaSetDevice(ctx->device);
        if (ce != cudaSuccess) { delete g; return gfail(nullptr, VCRT_ERR_CUDA, std::string("failed to set device: ") + cudaGetErrorString(ce)); }
        ncclUniqueId u;
        std::memcpy(&u, id, sizeof u);
        g->comms.resize(1);
        ncclResult_t r = api->CommInitRank(&g->comms[0], world, u, rank);
        if (r != ncclSuccess) { g->comms.clear(); delete g; return nccl_fail(nullptr, r, "create the NCCL communicator (ncclCommInitRank)"); }
    }
    g->mine.resize(1); g->everyone.resize(1);
    *out = g;
    return VCRT_OK;
}

int vcrt_group_destroy(vcrt_group* g) {
    if (!g) return VCRT_OK;
    NcclApi* api = nccl_api();
    for (size_t i = 0; i < g->members.size(); ++i) {
        vcrt_ctx* c = g->members[i];
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        if (i < g->comms.size() && g->comms[i] && api->CommDestroy) api->CommDestroy(g->comms[i]);
        if (i < g->mine.size() && g->mine[i].ptr) cudaFree(g->mine[i].ptr);
        if (i < g->everyone.size() && g->everyone[i].ptr) cudaFree(g->everyone[i].ptr);
        if (g->owns_ctx) vcrt_destroy(c);
    }
    delete g;
    return VCRT_OK;
}

int vcrt_group_size(const vcrt_group* g) { return g ? g->world : 0; }
int vcrt_group_local_count(const vcrt_group* g) { return g ? (int)g->members.size() : 0; }
int vcrt_group_rank(const vcrt_group* g, int local_index) { return (g && local_index >= 0 && local_index < (int)g->members.size()) ? g->first_rank + local_index : -1; }
vcrt_ctx* vcrt_group_ctx(vcrt_group* g, int local_index) { return (g && local_index >= 0 && local_index < (int)g->members.size()) ? g->members[local_index] : nullptr; }

#define FOR_MEMBERS(call) do { if (!g) return VCRT_ERR_INVALID; for (vcrt_ctx* c : g->members) { int rc_ = (call); if (rc_) return member_fail(g, c, rc_); } return VCRT_OK; } while (0)

int vcrt_group_set_shader(vcrt_group* g, const char* shader_path) { FOR_MEMBERS(vcrt_set_shader(c, shader_path)); }
int vcrt_group_set_buffer(vcrt_group* g, int binding, const void* host, size_t bytes) { FOR_MEMBERS(vcrt_set_buffer(c, binding, host, bytes)); }
int vcrt_group_set_image_size(vcrt_group* g, uint32_t width, uint32_t height) { FOR_MEMBERS(vcrt_set_image_size(c, width, height)); }
int vcrt_group_set_ubo(vcrt_group* g, const vcrt_ubo* ubo) { FOR_MEMBERS(vcrt_set_ubo(c, ubo)); }
int vcrt_group_set_option(vcrt_group* g, const char* key, const char* value) { FOR_MEMBERS(vcrt_set_option(c, key, value)); }
int vcrt_group_synchronize(vcrt_group* g) { FOR_MEMBERS(vcrt_synchronize(c)); }

int vcrt_group_render(vcrt_group* g, const vcrt_render_params* params, int mode, float gamma) {
    if (!g || !params) return gfail(g, VCRT_ERR_INVALID, "vcrt_group_render: NULL argument");
    if (params->struct_size != sizeof(vcrt_render_params)) return gfail(g, VCRT_ERR_INVALID, "vcrt_group_render: struct_size mismatch");
    if (mode != VCRT_SHARD_TILES && mode != VCRT_SHARD_SAMPLES) return gfail(g, VCRT_ERR_INVALID, "vcrt_group_render: mode must be VCRT_SHARD_TILES or VCRT_SHARD_SAMPLES");
    if (params->accum_mode != VCRT_ACCUM_F32) return gfail(g, VCRT_ERR_INVALID, "vcrt_group_render: shares are combined through the f32 accumulation (accum_mode = VCRT_ACCUM_F32)");
    if (params->tile_count > 1) return gfail(g, VCRT_ERR_INVALID, "vcrt_group_render: tile_rank/tile_count are set by the group");
    NcclApi* api = nccl_api();
    const uint32_t world = (uint32_t)g->world;
    const uint32_t total = params->sample_count ? params->sample_count : 1u;
    const size_t nm = g->members.size();
    const uint32_t W = g->members[0]->W, H = g->members[0]->H;
    if (W == 0) return gfail(g, VCRT_ERR_STATE, "vcrt_group_render: no storage images bound (vcrt_group_set_image_size)");

    // ---- every member renders its share (asynchronous on its own stream)
    for (size_t i = 0; i < nm; ++i) {
        vcrt_ctx* c = g->members[i];
        if (c->W != W || c->H != H) return gfail(g, VCRT_ERR_STATE, "vcrt_group_render: members differ in image size");
        const uint32_t rank = (uint32_t)g->first_rank + (uint32_t)i;
        int rc = vcrt_clear_accum(c);
        if (rc) return member_fail(g, c, rc);
        vcrt_render_params p = *params;
        bool active = true;
        if (world > 1 && mode == VCRT_SHARD_TILES) { p.tile_rank = rank; p.tile_count = world; }
        else if (world > 1) {   // contiguous sample slices; the first total % world ranks take one extra sample
            const uint32_t base = total / world, extra = total % world;
            const uint32_t n = base + (rank < extra ? 1u : 0u);
            p.sample_begin = params->sample_begin + rank * base + (rank < extra ? rank : extra);
            p.sample_count = n;
            active = n > 0;     // sample_count 0 means "1" to vcrt_render: an empty slice is not launched
        }
        if (active && (rc = vcrt_render(c, &p))) return member_fail(g, c, rc);
    }
    // ---- combine
    if (mode == VCRT_SHARD_SAMPLES) {
        if (world > 1) {
            ncclResult_t r = api->GroupStart();
            if (r != ncclSuccess) return nccl_fail(g, r, "start the NCCL group");
            for (size_t i = 0; i < nm; ++i) {
                vcrt_ctx* c = g->members[i];
                r = api->Reduce(c->accumf.ptr, c->accumf.ptr, (size_t)W * H * 4, ncclFloat, ncclSum, 0, g->comms[i], c->stream);
                if (r != ncclSuccess) { api->GroupEnd(); return nccl_fail(g, r, "reduce the f32 accumulation buffers (ncclReduce)"); }
            }
            if ((r = api->GroupEnd()) != ncclSuccess) return nccl_fail(g, r, "reduce the f32 accumulation buffers (ncclGroupEnd)");
        }
        if (g->first_rank == 0) {
            int rc = vcrt_resolve(g->members[0], total, gamma);
            if (rc) return member_fail(g, g->members[0], rc);
        }
        return VCRT_OK;
    }
    // tiles: every rank resolves its own tiles to rgba8, packs them, all-gathers the packed buffers and scatters every peer's
    // tiles into its image with ONE launch
    const uint32_t tiles_per_rank = max_owned_tiles(W, H, world);
    const size_t n = (size_t)tiles_per_rank * 1024u * 4u;
    for (size_t i = 0; i < nm; ++i) {
        vcrt_ctx* c = g->members[i];
        int rc = vcrt_resolve(c, total, gamma);
        if (rc) return member_fail(g, c, rc);
        if (world == 1) continue;
        cudaSetDevice(c->device);
        if ((rc = vcrt_ensure(c, g->mine[i], n, "allocate packed tiles")) || (rc = vcrt_ensure(c, g->everyone[i], n * world, "allocate gathered tiles"))) return member_fail(g, c, rc);
        if ((rc = vcrt_pack_tiles(c, 0, (uint32_t)g->first_rank + (uint32_t)i, world, g->mine[i].ptr, n))) return member_fail(g, c, rc);
    }
    if (world == 1) return VCRT_OK;
    ncclResult_t r = api->GroupStart();
    if (r != ncclSuccess) return nccl_fail(g, r, "start the NCCL group");
    for (size_t i = 0; i < nm; ++i) {
        vcrt_ctx* c = g->members[i];
        r = api->AllGather(g->mine[i].ptr, g->everyone[i].ptr, n, ncclUint8, g->comms[i], c->stream);
        if (r != ncclSuccess) { api->GroupEnd(); return nccl_fail(g, r, "gather the packed tiles (ncclAllGather)"); }
    }
    if ((r = api->GroupEnd()) != ncclSuccess) return nccl_fail(g, r, "gather the packed tiles (ncclGroupEnd)");
    for (size_t i = 0; i < nm; ++i) {
        vcrt_ctx* c = g->members[i];
        cudaSetDevice(c->device);
        cudaError_t e = vcrt::launch_unpack_all_tiles(c->target.ptr, g->everyone[i].ptr, 4, W, H, world, tiles_per_rank, c->stream);
        if (e != cudaSuccess) { vcrt_cuda_fail(c, e, "launch tile kernel"); return member_fail(g, c, VCRT_ERR_CUDA); }
        c->launches += 1;
    }
    return VCRT_OK;
}

int vcrt_group_read_target_rgba8(vcrt_group* g, int local_index, void* dst, size_t bytes) {
    vcrt_ctx* c = vcrt_group_ctx(g, local_index);
    if (!c) return gfail(g, VCRT_ERR_INVALID, "vcrt_group_read_target_rgba8: no such member");
    int rc = vcrt_read_target_rgba8(c, dst, bytes);
    return rc ? member_fail(g, c, rc) : VCRT_OK;
}

}  // extern "C"
