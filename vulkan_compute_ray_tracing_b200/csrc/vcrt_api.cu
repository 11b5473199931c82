// vcrt_api.cu -- the C ABI of include/vcrt.h: context, buffers, dispatch.  No CPU rendering path exists in
// this library: every render call launches CUDA kernels or fails with an error.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vcrt.h"
#include "vcrt_ctx.h"
#include "vcrt_devbuild.h"
#include "vcrt_host_setup.h"
#include "vcrt_launch.h"
#include "vcrt_repack.h"

using namespace vcrt;

static thread_local std::string g_create_error;

static const size_t kStride[8] = {0, 0, 0, sizeof(vcrt_triangle), sizeof(vcrt_material), sizeof(vcrt_bvh_node), sizeof(vcrt_light), sizeof(vcrt_sphere)};

static_assert(sizeof(vcrt_material) == 32 && sizeof(vcrt_triangle) == 48 && sizeof(vcrt_sphere) == 32 && sizeof(vcrt_bvh_node) == 48 &&
              sizeof(vcrt_light) == 8 && sizeof(vcrt_ubo) == 32 && sizeof(vcrt_aov) == 16, "data ABI");
static_assert(offsetof(vcrt_material, albedo) == 16 && offsetof(vcrt_triangle, v1) == 16 && offsetof(vcrt_triangle, v2) == 32 &&
              offsetof(vcrt_triangle, materialIndex) == 44 && offsetof(vcrt_sphere, materialIndex) == 16 && offsetof(vcrt_bvh_node, max) == 16 &&
              offsetof(vcrt_bvh_node, leftNodeIndex) == 28 && offsetof(vcrt_bvh_node, rightNodeIndex) == 32 && offsetof(vcrt_bvh_node, objectIndex) == 36,
              "data ABI offsets (GpuModels.h:26-63)");

static int fail(vcrt_ctx* c, int code, const std::string& msg) {
    if (c) c->error = msg; else g_create_error = msg;
    return code;
}
static int cuda_fail(vcrt_ctx* c, cudaError_t e, const char* what) {
    return fail(c, VCRT_ERR_CUDA, std::string("failed to ") + what + ": " + cudaGetErrorString(e));
}
#define CU(c, call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail((c), e_, (what)); } while (0)

// calls that touch the context's images, buffers or render stream are not allowed between vcrt_frames_begin and vcrt_frames_end
#define NOFRAMES(c, name) do { if ((c)->frames_n) return fail((c), VCRT_ERR_STATE, std::string(name) + ": frames are in flight (call vcrt_frames_end first)"); } while (0)

int vcrt_fail(vcrt_ctx* c, int code, const std::string& msg) { return fail(c, code, msg); }
int vcrt_cuda_fail(vcrt_ctx* c, cudaError_t e, const char* what) { return cuda_fail(c, e, what); }
static int ensure(vcrt_ctx* c, DevBuf& b, size_t bytes, const char* what);
int vcrt_ensure(vcrt_ctx* c, DevBuf& b, size_t bytes, const char* what) { return ensure(c, b, bytes, what); }

static int ensure(vcrt_ctx* c, DevBuf& b, size_t bytes, const char* what) {
    if (bytes > b.capacity) {
        if (b.ptr) CU(c, cudaFree(b.ptr), "free device buffer");
        b.ptr = nullptr; b.capacity = 0;
        CU(c, cudaMalloc(&b.ptr, bytes ? bytes : 16), what);
        b.capacity = bytes ? bytes : 16;
    }
    b.bytes = bytes;
    return VCRT_OK;
}

extern "C" {

const char* vcrt_version(void) { return "vcrt 0.1 (sm_100a)"; }

const char* vcrt_last_error(const vcrt_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int vcrt_create(int device, vcrt_ctx** out) {
    if (!out) return fail(nullptr, VCRT_ERR_INVALID, "vcrt_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "enumerate CUDA devices");
    if (device < 0 || device >= n) return fail(nullptr, VCRT_ERR_INVALID, "vcrt_create: no such CUDA device");
    vcrt_ctx* c = new (std::nothrow) vcrt_ctx();
    if (!c) return fail(nullptr, VCRT_ERR_NOMEM, "vcrt_create: out of host memory");
    c->device = device;
    std::memset(&c->ubo, 0, sizeof c->ubo);
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaMalloc((void**)&c->d_counters, 8 * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMemsetAsync(c->d_counters, 0, 8 * sizeof(unsigned long long), c->stream)) != cudaSuccess) {
        int rc = cuda_fail(nullptr, e, "create context");
        delete c;
        return rc;
    }
    c->own_stream = c->stream;
    *out = c;
    return VCRT_OK;
}

int vcrt_set_option(vcrt_ctx* c, const char* key, const char* value) {
    if (!c || !key || !value) return fail(c, VCRT_ERR_INVALID, "vcrt_set_option: NULL argument");
    const std::string k(key), v(value);
    if (k == "fast_bvh") {
        if (v != "sah" && v != "sah_plain" && v != "topology") return fail(c, VCRT_ERR_INVALID, "vcrt_set_option: fast_bvh must be 'sah', 'sah_plain' or 'topology'");
        const bool sah = v != "topology", opt = v == "sah";
        if (sah != c->fast_sah || opt != c->fast_reinsert) { c->fast_sah = sah; c->fast_reinsert = opt; c->fast_dirty = true; }
        return VCRT_OK;
    }
    if (k == "fast_nodes") {
        const int m = v == "auto" ? 0 : v == "q15" ? 1 : v == "f32" ? 2 : v == "q15x4" ? 3 : -1;
        if (m < 0) return fail(c, VCRT_ERR_INVALID, "vcrt_set_option: fast_nodes must be 'auto', 'q15x4', 'q15' or 'f32'");
        if (m != c->fast_nodes) { c->fast_nodes = m; c->fast_dirty = true; }
        return VCRT_OK;
    }
    if (k == "fast_build") {
        const int m = v == "auto" ? 0 : v == "host" ? 1 : v == "device" ? 2 : -1;
        if (m < 0) return fail(c, VCRT_ERR_INVALID, "vcrt_set_option: fast_build must be 'auto', 'host' or 'device'");
        if (m != c->fast_build) { c->fast_build = m; c->fast_dirty = true; }
        return VCRT_OK;
    }
    if (k == "dispatch_traversal") {
        const int m = v == "auto" ? 0 : v == "reference" ? 1 : v == "fast" ? 2 : -1;
        if (m < 0) return fail(c, VCRT_ERR_INVALID, "vcrt_set_option: dispatch_traversal must be 'auto', 'reference' or 'fast'");
        c->dispatch_trav = m;
        return VCRT_OK;
    }
    if (k == "host_threads") {   // OpenMP threads of the host-side record build (launchers such as torchrun export OMP_NUM_THREADS=1)
        const int n = atoi(value);
        if (n < 1 || n > 1024) return fail(c, VCRT_ERR_INVALID, "vcrt_set_option: host_threads must be 1..1024");
        set_repack_threads(n);
        return VCRT_OK;
    }
    if (k == "wf_batch_paths") {
        const long long n = atoll(value);
        if (n < 1024 || n > (1ll << 30)) return fail(c, VCRT_ERR_INVALID, "vcrt_set_option: wf_batch_paths must be 1024..2^30");
        c->wf_batch = (uint32_t)n;
        return VCRT_OK;
    }
    if (k == "wf_streams") {
        const int n = v == "auto" ? 0 : atoi(value);
        if (n < 0 || n > VCRT_MAX_PIPES || (n == 0 && v != "auto")) return fail(c, VCRT_ERR_INVALID, "vcrt_set_option: wf_streams must be 'auto' or 1..4");
        c->wf_streams = n;
        return VCRT_OK;
    }
    if (k == "trace_timing") {   // "auto": around the trace launches of multi-sample renders only (a 1-spp frame has 16 such events: 2 % of its time)
        if (v != "on" && v != "off" && v != "auto") return fail(c, VCRT_ERR_INVALID, "vcrt_set_option: trace_timing must be 'auto', 'on' or 'off'");
        c->trace_timing = v == "auto" ? 0 : v == "on" ? 1 : 2;
        return VCRT_OK;
    }
    if (k == "leaf_threshold" || k == "shade_threshold" || k == "continue_threshold") {
        const int n = atoi(value);
        if (n < 1 || n > 32) return fail(c, VCRT_ERR_INVALID, "vcrt_set_option: threshold must be 1..32 lanes");
        (k == "leaf_threshold" ? c->leaf_threshold : k == "shade_threshold" ? c->shade_threshold : c->continue_threshold) = (uint32_t)n;
        return VCRT_OK;
    }
    return fail(c, VCRT_ERR_INVALID, "vcrt_set_option: unknown option '" + k + "'");
}

static int prepare_fast(vcrt_ctx* c);

int vcrt_get_info(vcrt_ctx* c, const char* key, char* value, size_t capacity) {
    if (!c || !key || !value || capacity == 0) return fail(c, VCRT_ERR_INVALID, "vcrt_get_info: NULL argument");
    const std::string k(key);
    std::string v;
    if (k == "fast_nodes" || k == "fast_node_count" || k == "fast_depth") {
        CU(c, cudaSetDevice(c->device), "set device");
        const bool ok = prepare_fast(c) == VCRT_OK;
        if (k == "fast_nodes") v = !ok ? "none" : c->wide ? "q15x4" : c->quantized ? "q15" : "f32";
        else if (k == "fast_node_count") v = std::to_string(ok ? (c->wide ? c->nf4nodes : c->nfnodes) : 0u);
        else v = std::to_string(ok ? c->fast_depth : 0u);
    } else if (k == "fast_build" || k == "fast_build_ms" || k == "fast_build_stages") {
        CU(c, cudaSetDevice(c->device), "set device");
        const bool ok = prepare_fast(c) == VCRT_OK;
        if (k == "fast_build") v = !ok ? "none" : c->built_on_device ? "device" : "host";
        else if (k == "fast_build_stages") v = ok && c->built_on_device ? c->fast_build_stages : "";
        else { char b[32]; snprintf(b, sizeof b, "%.3f", c->fast_build_ms); v = b; }
    } else if (k == "wf_batch_paths") v = std::to_string(c->wf_batch);
    else if (k == "wf_pipelines") v = std::to_string(c->wf_pipes_used);   // pipelines of the last wavefront render (option "wf_streams")
    else if (k == "dispatch_kernel") v = c->dispatch_fast ? "fast" : "reference";
    else if (k == "device") v = std::to_string(c->device);
    else return fail(c, VCRT_ERR_INVALID, "vcrt_get_info: unknown key '" + k + "'");
    if (v.size() + 1 > capacity) return fail(c, VCRT_ERR_INVALID, "vcrt_get_info: buffer too small");
    std::memcpy(value, v.c_str(), v.size() + 1);
    return VCRT_OK;
}

int vcrt_set_stream(vcrt_ctx* c, void* cuda_stream) {
    if (!c) return VCRT_ERR_INVALID;
    NOFRAMES(c, "vcrt_set_stream");
    CU(c, cudaSetDevice(c->device), "set device");
    CU(c, cudaStreamSynchronize(c->stream), "synchronize");
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return VCRT_OK;
}

int vcrt_destroy(vcrt_ctx* c) {
    if (!c) return VCRT_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& ev : c->events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    for (cudaEvent_t ev : c->ev_pool) cudaEventDestroy(ev);
    for (int i = 1; i < VCRT_MAX_PIPES; ++i) {
        if (c->pipe_stream[i]) { cudaStreamSynchronize(c->pipe_stream[i]); cudaStreamDestroy(c->pipe_stream[i]); }
        if (c->join_ev[i]) cudaEventDestroy(c->join_ev[i]);
    }
    if (c->fork_ev) cudaEventDestroy(c->fork_ev);
    for (FrameSlot& fs : c->frame_slot) {
        if (fs.stream) { cudaStreamSynchronize(fs.stream); cudaStreamDestroy(fs.stream); }
        if (fs.folded) cudaEventDestroy(fs.folded);
        if (fs.done) cudaEventDestroy(fs.done);
        if (fs.image.ptr) cudaFree(fs.image.ptr);
        if (fs.sample.ptr) cudaFree(fs.sample.ptr);
        if (fs.counter) cudaFree(fs.counter);
    }
    c->trace_timer.destroy();
    for (auto& b : c->ssbo) if (b.ptr) cudaFree(b.ptr);
    for (DevBuf* b : {&c->fnodes, &c->ftris, &c->target, &c->accum8, &c->accumf, &c->aov, &c->wf_q0, &c->wf_q1, &c->wf_hit, &c->wf_color, &c->wf_counts, &c->qnodes, &c->q4nodes, &c->present}) if (b->ptr) cudaFree(b->ptr);
    if (c->d_counters) cudaFree(c->d_counters);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return VCRT_OK;
}

int vcrt_set_shader(vcrt_ctx* c, const char* path) {
    if (!c || !path) return fail(c, VCRT_ERR_INVALID, "vcrt_set_shader: NULL argument");
    std::string s(path);
    size_t slash = s.find_last_of("/\\");
    if (slash != std::string::npos) s = s.substr(slash + 1);
    size_t dot = s.find_last_of('.');
    if (dot != std::string::npos) s = s.substr(0, dot);
    if (s == "ray-trace-compute") c->shader = VCRT_SHADER_FULL;
    else if (s == "ray-trace-compute-simple") c->shader = VCRT_SHADER_SIMPLE;
    else return fail(c, VCRT_ERR_INVALID, "failed to open file: no kernel for shader '" + std::string(path) + "'");
    return VCRT_OK;
}

static int set_buffer_common(vcrt_ctx* c, int binding, const void* src, size_t bytes, cudaMemcpyKind kind) {
    if (!c) return VCRT_ERR_INVALID;
    if (binding < VCRT_BINDING_TRIANGLES || binding > VCRT_BINDING_SPHERES) return fail(c, VCRT_ERR_INVALID, "vcrt_set_buffer: binding must be 3..7");
    if (bytes % kStride[binding] != 0) return fail(c, VCRT_ERR_INVALID, "vcrt_set_buffer: size is not a multiple of the record size");
    if (bytes / kStride[binding] > 0x7fffffffu) return fail(c, VCRT_ERR_INVALID, "vcrt_set_buffer: too many records");
    if (bytes && !src) return fail(c, VCRT_ERR_INVALID, "vcrt_set_buffer: NULL source");
    NOFRAMES(c, "vcrt_set_buffer");
    CU(c, cudaSetDevice(c->device), "set device");
    int rc = ensure(c, c->ssbo[binding], bytes, "allocate storage buffer");
    if (rc) return rc;
    if (bytes) CU(c, cudaMemcpyAsync(c->ssbo[binding].ptr, src, bytes, kind, c->stream), "copy storage buffer");
    if (binding == VCRT_BINDING_TRIANGLES || binding == VCRT_BINDING_BVH) {
        // the host-side record build reads these two buffers back from the device if and when it runs (fetch_host_copies)
        (binding == VCRT_BINDING_TRIANGLES ? c->host_tris_valid : c->host_bvh_valid) = false;
        c->fast_dirty = true;
    }
    if (kind == cudaMemcpyHostToDevice) CU(c, cudaStreamSynchronize(c->stream), "synchronize");  // host pointer is borrowed for the call only
    return VCRT_OK;
}

int vcrt_set_buffer(vcrt_ctx* c, int binding, const void* host, size_t bytes) { return set_buffer_common(c, binding, host, bytes, cudaMemcpyHostToDevice); }
int vcrt_set_buffer_device(vcrt_ctx* c, int binding, const void* dev, size_t bytes) { return set_buffer_common(c, binding, dev, bytes, cudaMemcpyDeviceToDevice); }

int vcrt_clear_accum(vcrt_ctx* c) {
    if (!c) return VCRT_ERR_INVALID;
    NOFRAMES(c, "vcrt_clear_accum");
    CU(c, cudaSetDevice(c->device), "set device");
    for (DevBuf* b : {&c->target, &c->accum8, &c->accumf, &c->aov, &c->present})
        if (b->ptr && b->bytes) CU(c, cudaMemsetAsync(b->ptr, 0, b->bytes, c->stream), "clear image");
    return VCRT_OK;
}

int vcrt_set_image_size(vcrt_ctx* c, uint32_t w, uint32_t h) {
    if (!c) return VCRT_ERR_INVALID;
    if (w == 0 || h == 0 || w > 65536u || h > 65536u) return fail(c, VCRT_ERR_INVALID, "vcrt_set_image_size: bad extent");
    NOFRAMES(c, "vcrt_set_image_size");
    CU(c, cudaSetDevice(c->device), "set device");
    const size_t npix = (size_t)w * h;
    int rc;
    if ((rc = ensure(c, c->target, npix * 4, "allocate target image")) || (rc = ensure(c, c->accum8, npix * 4, "allocate accumulation image")) ||
        (rc = ensure(c, c->accumf, npix * 16, "allocate f32 accumulation")) || (rc = ensure(c, c->aov, npix * sizeof(vcrt_aov), "allocate AOV buffer")) ||
        (rc = ensure(c, c->present, npix * 4, "allocate present image")))
        return rc;
    c->W = w; c->H = h;
    return vcrt_clear_accum(c);
}

int vcrt_set_ubo(vcrt_ctx* c, const vcrt_ubo* ubo) {
    if (!c || !ubo) return fail(c, VCRT_ERR_INVALID, "vcrt_set_ubo: NULL argument");
    c->ubo = *ubo;
    return VCRT_OK;
}

// Folds the timing events of renders that have already finished into the counters, without waiting for anything: keeps
// the event lists short for callers that render frame after frame and never read the counters.
static void harvest_events(vcrt_ctx* c) {
    size_t done = 0;
    while (done < c->events.size() && cudaEventQuery(c->events[done].second) == cudaSuccess) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, c->events[done].first, c->events[done].second) == cudaSuccess) c->kernel_ms += ms;
        c->ev_pool.push_back(c->events[done].first);
        c->ev_pool.push_back(c->events[done].second);
        ++done;
    }
    c->events.erase(c->events.begin(), c->events.begin() + (long)done);
    c->trace_timer.harvest(&c->trace_ms, &c->trace_launches);
}

static int fetch_host_copies(vcrt_ctx* c) {
    for (int b : {VCRT_BINDING_TRIANGLES, VCRT_BINDING_BVH}) {
        bool& valid = b == VCRT_BINDING_TRIANGLES ? c->host_tris_valid : c->host_bvh_valid;
        if (valid) continue;
        std::vector<uint8_t>& shadow = b == VCRT_BINDING_TRIANGLES ? c->host_tris : c->host_bvh;
        shadow.resize(c->ssbo[b].bytes);
        if (c->ssbo[b].bytes) CU(c, cudaMemcpyAsync(shadow.data(), c->ssbo[b].ptr, c->ssbo[b].bytes, cudaMemcpyDeviceToHost, c->stream), "read back storage buffer");
        CU(c, cudaStreamSynchronize(c->stream), "synchronize");
        valid = true;
    }
    return VCRT_OK;
}

// Traversal records of the fast path: by the host builder (vcrt_repack.cpp: binned SAH, every node format, the precise error
// messages) or, for the default tree (a rebuilt topology walked as 4-wide quantised nodes), by the device builder (vcrt_devbuild.cu).
static int prepare_fast(vcrt_ctx* c) {
    if (c->fast_dirty) {
        // The context counts as prepared only once every record is on the device: any early return below leaves it dirty, so the
        // next render retries (or fails again) instead of launching with missing or stale node / triangle buffers.
        const auto t0 = std::chrono::steady_clock::now();
        c->fast_ok = false;
        c->fast_dirty = true;
        c->fast_err.clear();
        const uint32_t nbvh = (uint32_t)(c->ssbo[VCRT_BINDING_BVH].bytes / sizeof(vcrt_bvh_node)), ntris = (uint32_t)(c->ssbo[VCRT_BINDING_TRIANGLES].bytes / sizeof(vcrt_triangle));
        const bool device_tree = c->fast_sah && (c->fast_nodes == 0 || c->fast_nodes == 3);
        if (c->fast_build == 2 && !device_tree) return fail(c, VCRT_ERR_INVALID, "fast_build=device builds the rebuilt 4-wide quantised tree only (fast_bvh=sah, fast_nodes=auto|q15x4)");
        if (c->fast_build == 2 || (c->fast_build == 0 && device_tree && c->auto_device)) {
            devbuild::Alloc alloc;
            alloc.ftris = [c](size_t bytes) { return ensure(c, c->ftris, bytes, "allocate repacked triangles") ? nullptr : c->ftris.ptr; };
            alloc.q4nodes = [c](size_t bytes) { return ensure(c, c->q4nodes, bytes, "allocate 4-wide nodes") ? nullptr : c->q4nodes.ptr; };
            devbuild::Result r;
            std::string why;
            const int rc = devbuild::run(c->ssbo[VCRT_BINDING_BVH].ptr, nbvh, c->ssbo[VCRT_BINDING_TRIANGLES].ptr, ntris, c->fast_nodes == 3 ? 3.0e38f : 2.5e-4f,
                                         VCRT_FAST_STACK, c->stream, alloc, r, why);
            if (rc < 0) return fail(c, VCRT_ERR_CUDA, "failed to build traversal records on the device: " + why);
            if (rc == 0) {
                c->quantized = true; c->wide = true; c->have_binary = false; c->built_on_device = true;
                c->froot4 = r.root4; c->nf4nodes = r.nwide;
                std::memcpy(c->qorg, r.qorg, sizeof c->qorg); std::memcpy(c->qext, r.qext, sizeof c->qext);
                c->froot = (int32_t)0x80000000; c->nfnodes = 0;
                c->fast_depth = r.depth; c->bound_depth = r.bound_depth;
                c->fast_ok = true; c->fast_dirty = false;
                c->fast_build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                char stages[160];
                snprintf(stages, sizeof stages, "ranks %.2f ms, records+sort %.2f, PLOC %.2f (%u rounds), collapse %.2f (%u levels)", r.ms_ranks, r.ms_sort - r.ms_ranks,
                         r.ms_ploc - r.ms_sort, r.ploc_rounds, r.ms_total - r.ms_ploc, r.wide_levels);
                c->fast_build_stages = stages;
                return VCRT_OK;
            }
            if (c->fast_build == 2) { c->fast_dirty = false; c->fast_err = why; return fail(c, VCRT_ERR_INVALID, "fast traversal unavailable (fast_build=device): " + why); }
        }
        int rc = fetch_host_copies(c);
        if (rc) return rc;
        FastBvh fb;
        bool ok = build_fast_bvh((const vcrt_bvh_node*)c->host_bvh.data(), (uint32_t)(c->host_bvh.size() / sizeof(vcrt_bvh_node)),
                                 (const vcrt_triangle*)c->host_tris.data(), (uint32_t)(c->host_tris.size() / sizeof(vcrt_triangle)), fb, c->fast_err);
        if (ok && c->fast_sah) ok = rebuild_fast_bvh_sah(fb, c->fast_err);
        // insertion-based optimisation of the rebuilt tree (3 passes over the 10 % of the inner nodes with the largest area): 5 % fewer
        // node visits per ray for 3 s of host time per million triangles; scenes beyond 2 Mi triangles keep the tree as built
        std::vector<float> plain_nodes;      // the tree as built, in case the optimised one cannot be walked as a 4-wide tree (its stack bound)
        uint32_t plain_depth = 0;
        const bool reinsert = ok && c->fast_sah && c->fast_reinsert && fb.num_slots() <= (2u << 20);
        if (reinsert) { plain_nodes = fb.nodes; plain_depth = fb.depth; ok = optimize_fast_bvh_reinsert(fb, 3, 0.10f, c->fast_err); }
        if (ok) ok = check_fast_depth(fb, c->fast_err);   // the tree that will be walked: a deep bound tree is fine once rebuilt
        if (!ok) { c->fast_dirty = false; return fail(c, VCRT_ERR_INVALID, "fast traversal unavailable: " + c->fast_err); }   // a property of the bound tree: no retry
        precompute_triangles(fb);
        // 32-byte quantised nodes: "auto" accepts quanta up to 2.5e-4 (the reference's own leaf padding is 1e-4), "q15" any
        const float max_quantum = (c->fast_nodes == 1 || c->fast_nodes == 3) ? 3.0e38f : 2.5e-4f;
        bool quantized = c->fast_nodes != 2 && quantize_fast_bvh(fb, max_quantum);
        // 4-wide form of the quantised tree for the wavefront trace kernel ("auto" and "q15x4"; "q15" keeps the binary tree)
        bool wide = quantized && c->fast_nodes != 1 && build_wide_bvh(fb, VCRT_FAST_STACK);
        if (!wide && quantized && c->fast_nodes != 1 && reinsert) {   // the optimised tree asks for a deeper stack than the kernel has: walk the tree as built
            fb.nodes.swap(plain_nodes); fb.depth = plain_depth;
            quantized = quantize_fast_bvh(fb, max_quantum);
            wide = quantized && build_wide_bvh(fb, VCRT_FAST_STACK);
        }
        if ((rc = ensure(c, c->fnodes, fb.nodes.size() * 4, "allocate repacked nodes")) || (rc = ensure(c, c->ftris, fb.tris64.size() * 4, "allocate repacked triangles"))) return rc;
        if (!fb.nodes.empty()) CU(c, cudaMemcpyAsync(c->fnodes.ptr, fb.nodes.data(), fb.nodes.size() * 4, cudaMemcpyHostToDevice, c->stream), "upload repacked nodes");
        if (!fb.tris64.empty()) CU(c, cudaMemcpyAsync(c->ftris.ptr, fb.tris64.data(), fb.tris64.size() * 4, cudaMemcpyHostToDevice, c->stream), "upload repacked triangles");
        if (wide) {
            if ((rc = ensure(c, c->q4nodes, fb.q4nodes.size() * 4, "allocate 4-wide nodes"))) return rc;
            CU(c, cudaMemcpyAsync(c->q4nodes.ptr, fb.q4nodes.data(), fb.q4nodes.size() * 4, cudaMemcpyHostToDevice, c->stream), "upload 4-wide nodes");
        }
        if (quantized) {
            if ((rc = ensure(c, c->qnodes, fb.qnodes.size() * 4, "allocate quantised nodes"))) return rc;
            CU(c, cudaMemcpyAsync(c->qnodes.ptr, fb.qnodes.data(), fb.qnodes.size() * 4, cudaMemcpyHostToDevice, c->stream), "upload quantised nodes");
        }
        CU(c, cudaStreamSynchronize(c->stream), "synchronize");
        c->quantized = quantized; c->wide = wide; c->have_binary = true; c->built_on_device = false;
        if (wide) { c->froot4 = fb.root4; c->nf4nodes = (uint32_t)(fb.q4nodes.size() / 16); }
        if (quantized) { std::memcpy(c->qorg, fb.qorg, sizeof c->qorg); std::memcpy(c->qext, fb.qext, sizeof c->qext); }
        c->froot = fb.root;
        c->nfnodes = fb.num_nodes();
        c->fast_depth = fb.depth;
        c->bound_depth = fb.bound_depth;
        c->fast_ok = true;
        c->fast_dirty = false;
        c->fast_build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    if (!c->fast_ok) return fail(c, VCRT_ERR_INVALID, "fast traversal unavailable: " + c->fast_err);
    return VCRT_OK;
}

// fs != nullptr: a frame in flight (vcrt_frame_submit) -- everything goes onto the slot's stream, the one-launch kernels leave the
// frame's sample in the slot's buffer instead of folding it (a.sample_out), the wavefront pipeline uses the slot's queue set and
// orders its accumulate kernel behind the previous frame's fold; *out_args receives the kernel arguments for the fold kernel.
static int render_common(vcrt_ctx* c, const vcrt_render_params& p, uint32_t covW, uint32_t covH, FrameSlot* fs = nullptr, int fs_index = 0, bool restart = false,
                         KernelArgs* out_args = nullptr) {
    if (c->W == 0) return fail(c, VCRT_ERR_STATE, "render: no storage images bound (vcrt_set_image_size)");
    if (!fs) NOFRAMES(c, "render");
    const cudaStream_t stream = fs ? fs->stream : c->stream;
    if (p.shader > VCRT_SHADER_SIMPLE || p.traversal > VCRT_TRAVERSAL_BRUTE_FORCE || p.rng_mode > VCRT_RNG_PHILOX || p.accum_mode > VCRT_ACCUM_F32 ||
        p.trig_mode > VCRT_TRIG_PORTABLE)
        return fail(c, VCRT_ERR_INVALID, "render: enum field out of range");
    if (p.stack_depth > VCRT_MAX_STACK) return fail(c, VCRT_ERR_INVALID, "render: stack_depth > 64");
    if (p.tile_count > 1 && p.tile_rank >= p.tile_count) return fail(c, VCRT_ERR_INVALID, "render: tile_rank >= tile_count");
    CU(c, cudaSetDevice(c->device), "set device");

    KernelArgs a;
    std::memset(&a, 0, sizeof a);
    SceneView& s = a.scene;
    s.tris = (const float4*)c->ssbo[VCRT_BINDING_TRIANGLES].ptr;   s.ntris = (uint32_t)(c->ssbo[VCRT_BINDING_TRIANGLES].bytes / sizeof(vcrt_triangle));
    s.mats = (const float4*)c->ssbo[VCRT_BINDING_MATERIALS].ptr;   s.nmats = (uint32_t)(c->ssbo[VCRT_BINDING_MATERIALS].bytes / sizeof(vcrt_material));
    s.bvh = (const float4*)c->ssbo[VCRT_BINDING_BVH].ptr;          s.nbvh = (uint32_t)(c->ssbo[VCRT_BINDING_BVH].bytes / sizeof(vcrt_bvh_node));
    s.lights = (const vcrt_light*)c->ssbo[VCRT_BINDING_LIGHTS].ptr; s.nlights = (uint32_t)(c->ssbo[VCRT_BINDING_LIGHTS].bytes / sizeof(vcrt_light));
    s.spheres = (const float4*)c->ssbo[VCRT_BINDING_SPHERES].ptr;  s.nspheres = (uint32_t)(c->ssbo[VCRT_BINDING_SPHERES].bytes / sizeof(vcrt_sphere));
    if (p.traversal == VCRT_TRAVERSAL_FAST) {
        int rc = prepare_fast(c);
        if (rc) return rc;
        s.fnodes = c->have_binary ? (const float4*)c->fnodes.ptr : nullptr; s.ftris = (const float4*)c->ftris.ptr; s.nfnodes = c->nfnodes; s.froot = c->froot;
        if (c->quantized) {
            s.qnodes = c->have_binary ? (const Words8*)c->qnodes.ptr : nullptr;
            s.qorg = make_float3(c->qorg[0], c->qorg[1], c->qorg[2]);
            s.qext = make_float3(c->qext[0], c->qext[1], c->qext[2]);
        }
        if (c->wide) { s.q4nodes = (const Words8*)c->q4nodes.ptr; s.froot4 = c->froot4; }
    } else {
        s.froot = (int32_t)0x80000000;
    }

    setup_args(a, c->ubo, p, c->W, c->H, covW, covH, s.nlights);
    a.flags &= ~VCRT_FLAG_INTERNAL_RESTART;
    if (restart) a.flags |= VCRT_FLAG_INTERNAL_RESTART;
    a.target = (uchar4*)c->target.ptr; a.accum8 = (uchar4*)c->accum8.ptr; a.accumf = (float4*)c->accumf.ptr; a.aov = (vcrt_aov*)c->aov.ptr;
    a.counters = c->d_counters;
    a.work_counter = fs ? (unsigned int*)fs->counter : (unsigned int*)(c->d_counters + 3);
    a.leaf_threshold = c->leaf_threshold; a.shade_threshold = c->shade_threshold; a.continue_threshold = c->continue_threshold;

    // timing events come from a pool (a frame loop renders thousands of frames: no event creation in steady state)
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    for (cudaEvent_t* ev : {&e0, &e1}) {
        if (!c->ev_pool.empty()) { *ev = c->ev_pool.back(); c->ev_pool.pop_back(); }
        else CU(c, cudaEventCreate(ev), "create event");
    }
    auto give_back = [&]() { c->ev_pool.push_back(e0); c->ev_pool.push_back(e1); };
    cudaError_t e;
    if ((e = cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), stream)) != cudaSuccess || (e = cudaEventRecord(e0, stream)) != cudaSuccess) {
        give_back();
        return cuda_fail(c, e, "start render");
    }
    const bool count = (p.flags & VCRT_FLAG_COUNT_TRAVERSAL) != 0;
    uint32_t nlaunch = 1;
    // A 1-spp frame of a shallow shader (the reference's own frame: NUM_BOUNCES 2 or 4) is one launch of the one-thread-per-pixel
    // kernel instead of three launches per bounce of the wavefront pipeline: 0.09 instead of 0.59 ms at 800x600 on the bundled
    // scene, identical results (r01 A/B; with more samples or deeper paths the wavefront pipeline wins at every size).
    // A deeper 1-spp frame goes to the megakernel: every trace launch of the wavefront pipeline lasts as long as its slowest ray
    // (~0.27 ms per bounce on C3 whatever the ray count), and eight such tails in a row cost more than the megakernel's lower lane
    // utilisation: 1.82 vs 2.09 ms per 1080p frame on C3, 0.84 vs 1.10 ms on the bundled scene (profiles/r02_v6_latency_*.log).
    bool one_launch = (p.flags & (VCRT_FLAG_STATIC_KERNEL | VCRT_FLAG_MEGAKERNEL)) != 0;
    // With two or more frames in flight (vcrt_frame_submit) the tails overlap the next frames' work anyway, and what counts is the
    // work per frame, where the wavefront pipeline is ahead: C3, depth 8, 1080p: 1.26 / 1.13 / 1.07 ms per frame with 2 / 3 / 4 frames in
    // flight against the megakernel's 1.57 / 1.47 / 1.47; bundled scene 0.71 / 0.61 / 0.57 against 0.72 / 0.60 / 0.60
    // (profiles/r02_v27_latency_*.log).
    if (p.traversal == VCRT_TRAVERSAL_FAST && !one_launch && !(p.flags & VCRT_FLAG_WAVEFRONT) && a.sample_count == 1u) {
        const bool deep = a.env.max_bounces > 4u;
        // one deep frame at a time: the megakernel (no barrier between bounces) on very small scenes, the wavefront pipeline otherwise --
        // the tail loop of its trace kernel has closed the gap (1080p, depth 8, ms per frame, wavefront vs megakernel: bundled scene of 1.4 K
        // triangles 0.97 vs 0.84; lit box of 20 K 1.22 vs 1.29, 100 K 1.38 vs 1.56, 300 K 1.50 vs 1.61, 1 M 1.70 vs 1.80)
        const bool wavefront = deep && ((fs && c->frames_n >= 2) || s.ntris >= 8192u);
        if (!wavefront) {
            a.flags |= deep ? VCRT_FLAG_MEGAKERNEL : VCRT_FLAG_STATIC_KERNEL;
            one_launch = true;
        }
    }
    if (p.traversal == VCRT_TRAVERSAL_FAST && !one_launch) {
        // wavefront pipeline: a batch = a range of pixels x all samples of the call; queues sized for what the call needs, at most
        // wf_batch paths per batch.  Option "wf_streams" = n cuts the call into n batches that run as parallel pipelines on their
        // own streams (each with its own queue set), so that the tail of one trace launch overlaps the others' work.
        // Measured on C3 (r02_v1): a 1-spp 1080p frame takes 2.19 ms as one pipeline and 2.41 ms as four -- every trace launch of a
        // small frame lasts as long as its longest ray (~0.2 ms) whatever its size, and four times as many launches cost more
        // than the overlap returns.  A large render is another matter (r02_v55/56, C3): two pipelines overlap one's shade launches
        // (ALU and streaming) with the other's trace launches (L2 gathers) -- 64 spp 40.0 -> 38.0 ms, 16 spp +1 %, 8 spp +-0, 4 spp
        // -1.5 %; three or four are slower again.  "auto" = 2 from VCRT_WF_AUTO2_PATHS paths per call, else 1.
        const uint64_t need = (uint64_t)a.owned_tiles * 1024u * a.sample_count;
        int sets = fs ? c->frames_n : c->wf_streams ? c->wf_streams : need >= VCRT_WF_AUTO2_PATHS ? 2 : 1;   // frames in flight: one queue set per slot
        uint64_t per_set = fs ? need : (need + (uint64_t)sets - 1) / (uint64_t)sets;   // a frame in flight is one pipeline
        per_set = (per_set + a.sample_count - 1) / a.sample_count * a.sample_count;   // whole pixels
        const uint64_t cap = fs ? c->wf_batch : c->wf_batch / (uint64_t)sets;          // the pipelines share the queue memory of one batch
        if (per_set > cap) per_set = cap / a.sample_count * a.sample_count;
        uint32_t want = (uint32_t)per_set;
        if (want < a.sample_count) want = a.sample_count;
        if (want < 1024u) want = 1024u;
        int rc = VCRT_OK;
        if (c->wf_capacity != want || c->wf_sets != sets) {
            for (int i = 0; i < c->frames_n; ++i)   // the queues are about to move: no frame in flight may still be using them
                if ((e = cudaStreamSynchronize(c->frame_slot[i].stream)) != cudaSuccess) { give_back(); return cuda_fail(c, e, "synchronize"); }
            const size_t n = (size_t)want * (size_t)sets;
            if ((rc = ensure(c, c->wf_q0, n * 48, "allocate ray queue")) || (rc = ensure(c, c->wf_q1, n * 48, "allocate ray queue")) ||
                (rc = ensure(c, c->wf_hit, n * 8, "allocate hit buffer")) || (rc = ensure(c, c->wf_color, n * 16, "allocate sample buffer")) ||
                (rc = ensure(c, c->wf_counts, 64 * VCRT_MAX_PIPES, "allocate queue counters"))) { give_back(); return rc; }
            c->wf_capacity = want; c->wf_sets = sets;
        }
        WfPipes pipes;
        std::memset(&pipes, 0, sizeof pipes);
        auto queue_set = [&](int i) {
            WfQueues q;
            q.q[0] = (float4*)c->wf_q0.ptr + (size_t)i * want * 3; q.q[1] = (float4*)c->wf_q1.ptr + (size_t)i * want * 3;
            q.hit = (uint2*)c->wf_hit.ptr + (size_t)i * want; q.sample_color = (float4*)c->wf_color.ptr + (size_t)i * want;
            q.counts = (unsigned int*)c->wf_counts.ptr + 16 * i;
            q.capacity = want;
            return q;
        };
        pipes.n = fs ? 1 : sets;
        c->wf_pipes_used = pipes.n;
        pipes.stream[0] = stream;
        if (fs) { pipes.q[0] = queue_set(fs_index); pipes.before_accumulate = c->last_folded; }
        for (int i = 0; i < (fs ? 0 : sets); ++i) {
            if (i > 0) {
                if (!c->pipe_stream[i] && (e = cudaStreamCreateWithFlags(&c->pipe_stream[i], cudaStreamNonBlocking)) != cudaSuccess) { give_back(); return cuda_fail(c, e, "create pipeline stream"); }
                if (!c->join_ev[i] && (e = cudaEventCreateWithFlags(&c->join_ev[i], cudaEventDisableTiming)) != cudaSuccess) { give_back(); return cuda_fail(c, e, "create event"); }
                pipes.stream[i] = c->pipe_stream[i];
                pipes.join[i] = c->join_ev[i];
            }
            pipes.q[i] = queue_set(i);
        }
        if (!fs && sets > 1 && !c->fork_ev && (e = cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming)) != cudaSuccess) { give_back(); return cuda_fail(c, e, "create event"); }
        pipes.fork = c->fork_ev;
        nlaunch = 0;
        c->trace_timer.enabled = c->trace_timing == 1 || (c->trace_timing == 0 && a.sample_count > 1u);
        e = launch_render_wavefront(a, (int)p.shader, (int)p.rng_mode, (int)p.trig_mode, count, pipes, &nlaunch, &c->trace_timer);
    } else {
        if (fs) a.sample_out = (float4*)fs->sample.ptr;   // one launch: the sample is handed to the fold kernel
        if (p.traversal == VCRT_TRAVERSAL_FAST) e = launch_render_fast(a, (int)p.shader, (int)p.rng_mode, (int)p.trig_mode, count, stream);
        else if (p.traversal == VCRT_TRAVERSAL_BRUTE_FORCE) e = launch_render_brute(a, (int)p.shader, (int)p.rng_mode, (int)p.trig_mode, count, stream);
        else e = launch_render_reference(a, (int)p.shader, (int)p.rng_mode, (int)p.trig_mode, count, stream);
    }
    if (e != cudaSuccess) { give_back(); return cuda_fail(c, e, "launch render kernel"); }
    if ((e = cudaEventRecord(e1, stream)) != cudaSuccess) { give_back(); return cuda_fail(c, e, "record event"); }
    if (out_args) *out_args = a;
    c->events.emplace_back(e0, e1);
    c->launches += nlaunch;
    if (c->events.size() >= 256 || c->trace_timer.pending.size() >= 2048) harvest_events(c);   // a frame loop that never asks for counters
    return VCRT_OK;
}

int vcrt_render(vcrt_ctx* c, const vcrt_render_params* p) {
    if (!c || !p) return fail(c, VCRT_ERR_INVALID, "vcrt_render: NULL argument");
    if (p->struct_size != sizeof(vcrt_render_params)) return fail(c, VCRT_ERR_INVALID, "vcrt_render: struct_size mismatch");
    const bool ref_cov = (p->flags & VCRT_FLAG_REF_DISPATCH_COVERAGE) != 0;
    return render_common(c, *p, ref_cov ? (c->W / 32) * 32 : c->W, ref_cov ? (c->H / 32) * 32 : c->H);
}

// What ComputeModel::computeCommand renders, as render parameters (shared by vcrt_dispatch and vcrt_frame_dispatch).
static int dispatch_params(vcrt_ctx* c, uint32_t gx, uint32_t gy, vcrt_render_params& p, uint32_t& covW, uint32_t& covH) {
    std::memset(&p, 0, sizeof p);
    p.struct_size = sizeof p;
    p.shader = (uint32_t)c->shader;
    p.traversal = VCRT_TRAVERSAL_REFERENCE;
    // The literal hit_bvh (16-entry stack, unordered, no culling) or -- same hit records by construction (vcrt_fast.cuh),
    // 5-6x less time per frame on the bundled scene -- the fast traversal in the one-launch kernel.  "auto" takes the fast one
    // when the reference's stack cannot overflow on the bound tree (deepest leaf <= 13 <=> at most 8192 triangles with the
    // reference's builder): past that depth the shader silently drops part of the tree, which only the
    // literal traversal reproduces.
    c->dispatch_fast = false;
    if (c->dispatch_trav != 1) {
        const size_t ntris = c->ssbo[VCRT_BINDING_TRIANGLES].bytes / sizeof(vcrt_triangle);
        if (c->dispatch_trav == 2 || ntris <= 8192) {
            CU(c, cudaSetDevice(c->device), "set device");
            const bool ok = prepare_fast(c) == VCRT_OK;
            if (ok && (c->dispatch_trav == 2 || c->bound_depth <= 13)) c->dispatch_fast = true;
            else if (c->dispatch_trav == 2) return fail(c, VCRT_ERR_INVALID, "vcrt_dispatch: dispatch_traversal=fast but the fast traversal is unavailable: " + c->fast_err);
            else c->error.clear();
        }
    }
    if (c->dispatch_fast) { p.traversal = VCRT_TRAVERSAL_FAST; p.flags |= VCRT_FLAG_STATIC_KERNEL; }
    p.rng_mode = VCRT_RNG_PCG_REF;
    p.accum_mode = VCRT_ACCUM_RGBA8_REF;
    p.trig_mode = VCRT_TRIG_LIBM;
    p.sample_begin = c->ubo.currentSample;
    p.sample_count = 1;
    const uint64_t cw = (uint64_t)gx * 32u, ch = (uint64_t)gy * 32u;   // invocations beyond the image neither load nor store
    covW = (uint32_t)(cw < c->W ? cw : c->W);
    covH = (uint32_t)(ch < c->H ? ch : c->H);
    return VCRT_OK;
}

int vcrt_dispatch(vcrt_ctx* c, uint32_t gx, uint32_t gy, uint32_t gz) {
    if (!c) return VCRT_ERR_INVALID;
    if (gz == 0 || gx == 0 || gy == 0) return VCRT_OK;   // vkCmdDispatch with a zero dimension does nothing
    NOFRAMES(c, "vcrt_dispatch");
    vcrt_render_params p;
    uint32_t covW, covH;
    int rc = dispatch_params(c, gx, gy, p, covW, covH);
    if (rc) return rc;
    return render_common(c, p, covW, covH);
}

int vcrt_resolve(vcrt_ctx* c, uint32_t total_samples, float gamma) {
    if (!c) return VCRT_ERR_INVALID;
    if (c->W == 0) return fail(c, VCRT_ERR_STATE, "vcrt_resolve: no storage images bound");
    if (total_samples == 0) return fail(c, VCRT_ERR_INVALID, "vcrt_resolve: total_samples is 0");
    NOFRAMES(c, "vcrt_resolve");
    CU(c, cudaSetDevice(c->device), "set device");
    CU(c, launch_resolve((const float4*)c->accumf.ptr, (uchar4*)c->target.ptr, c->W * c->H, 1.0f / (float)total_samples, gamma > 0.0f ? 1.0f / gamma : 0.0f, c->stream),
       "launch resolve kernel");
    c->launches += 1;
    return VCRT_OK;
}

int vcrt_post_process(vcrt_ctx* c, float mix, float sigma, float k_sigma, float threshold, float gamma) {
    if (!c) return VCRT_ERR_INVALID;
    if (c->W == 0) return fail(c, VCRT_ERR_STATE, "vcrt_post_process: no storage images bound");
    if (mix != 0.0f && (!(sigma > 0.0f) || !(k_sigma >= 0.0f) || !(threshold > 0.0f) || k_sigma * sigma > 64.0f))
        return fail(c, VCRT_ERR_INVALID, "vcrt_post_process: need sigma > 0, kSigma >= 0, threshold > 0, kSigma * sigma <= 64");
    NOFRAMES(c, "vcrt_post_process");
    CU(c, cudaSetDevice(c->device), "set device");
    CU(c, launch_post_process((const uchar4*)c->target.ptr, (uchar4*)c->present.ptr, c->W, c->H, mix, sigma, k_sigma, threshold, gamma > 0.0f ? 1.0f / gamma : 0.0f, c->stream),
       "launch post-process kernel");
    c->launches += 1;
    return VCRT_OK;
}

static int tiles_common(vcrt_ctx* c, int what, uint32_t rank, uint32_t count, void* packed, size_t bytes, bool pack) {
    if (!c) return VCRT_ERR_INVALID;
    if (c->W == 0) return fail(c, VCRT_ERR_STATE, "tiles: no storage images bound");
    if (what != 0 && what != 2) return fail(c, VCRT_ERR_INVALID, "tiles: only the rgba8 target (0) and the f32 accumulation (2) are tile-packed");
    if (count == 0) count = 1;
    if (rank >= count) return fail(c, VCRT_ERR_INVALID, "tiles: tile_rank >= tile_count");
    const uint32_t tiles = ((c->W + 31) / 32) * ((c->H + 31) / 32);
    const uint32_t owned = tiles > rank ? (tiles - rank + count - 1) / count : 0u;
    const size_t elem = what == 0 ? 4 : 16;
    if (!packed || bytes < (size_t)owned * 1024u * elem) return fail(c, VCRT_ERR_INVALID, "tiles: packed buffer smaller than owned_tiles * 1024 * element size");
    NOFRAMES(c, "tiles");
    CU(c, cudaSetDevice(c->device), "set device");
    CU(c, launch_tiles(what == 0 ? c->target.ptr : c->accumf.ptr, packed, (int)elem, pack, c->W, c->H, rank, count, owned, c->stream), "launch tile kernel");
    c->launches += owned ? 1 : 0;
    return VCRT_OK;
}

int vcrt_pack_tiles(vcrt_ctx* c, int what, uint32_t tile_rank, uint32_t tile_count, void* packed_dev, size_t bytes) {
    return tiles_common(c, what, tile_rank, tile_count, packed_dev, bytes, true);
}
int vcrt_unpack_tiles(vcrt_ctx* c, int what, uint32_t tile_rank, uint32_t tile_count, const void* packed_dev, size_t bytes) {
    return tiles_common(c, what, tile_rank, tile_count, const_cast<void*>(packed_dev), bytes, false);
}

static int read_common(vcrt_ctx* c, const DevBuf& b, void* dst, size_t bytes, const char* what) {
    if (!c) return VCRT_ERR_INVALID;
    if (c->W == 0) return fail(c, VCRT_ERR_STATE, std::string(what) + ": no storage images bound");
    if (!dst || bytes != b.bytes) return fail(c, VCRT_ERR_INVALID, std::string(what) + ": size mismatch");
    NOFRAMES(c, what);
    CU(c, cudaSetDevice(c->device), "set device");
    CU(c, cudaMemcpyAsync(dst, b.ptr, bytes, cudaMemcpyDeviceToHost, c->stream), what);
    CU(c, cudaStreamSynchronize(c->stream), what);
    return VCRT_OK;
}

int vcrt_read_target_rgba8(vcrt_ctx* c, void* dst, size_t bytes) { return c ? read_common(c, c->target, dst, bytes, "read target image") : VCRT_ERR_INVALID; }
int vcrt_read_accum_rgba8(vcrt_ctx* c, void* dst, size_t bytes) { return c ? read_common(c, c->accum8, dst, bytes, "read accumulation image") : VCRT_ERR_INVALID; }
int vcrt_read_accum_f32(vcrt_ctx* c, void* dst, size_t bytes) { return c ? read_common(c, c->accumf, dst, bytes, "read f32 accumulation") : VCRT_ERR_INVALID; }
int vcrt_read_present_rgba8(vcrt_ctx* c, void* dst, size_t bytes) { return c ? read_common(c, c->present, dst, bytes, "read present image") : VCRT_ERR_INVALID; }
int vcrt_read_aov(vcrt_ctx* c, void* dst, size_t bytes) { return c ? read_common(c, c->aov, dst, bytes, "read AOV buffer") : VCRT_ERR_INVALID; }

int vcrt_write_accum_f32(vcrt_ctx* c, const void* src, size_t bytes) {
    if (!c) return VCRT_ERR_INVALID;
    if (c->W == 0) return fail(c, VCRT_ERR_STATE, "vcrt_write_accum_f32: no storage images bound");
    if (!src || bytes != c->accumf.bytes) return fail(c, VCRT_ERR_INVALID, "vcrt_write_accum_f32: size mismatch");
    NOFRAMES(c, "vcrt_write_accum_f32");
    CU(c, cudaSetDevice(c->device), "set device");
    CU(c, cudaMemcpyAsync(c->accumf.ptr, src, bytes, cudaMemcpyHostToDevice, c->stream), "write f32 accumulation");
    CU(c, cudaStreamSynchronize(c->stream), "synchronize");
    return VCRT_OK;
}

int vcrt_device_ptr(vcrt_ctx* c, int what, void** out, size_t* bytes) {
    if (!c || !out) return fail(c, VCRT_ERR_INVALID, "vcrt_device_ptr: NULL argument");
    const DevBuf* b = what == 0 ? &c->target : what == 1 ? &c->accum8 : what == 2 ? &c->accumf : what == 3 ? &c->aov : nullptr;
    if (!b) return fail(c, VCRT_ERR_INVALID, "vcrt_device_ptr: unknown buffer");
    if (c->W == 0) return fail(c, VCRT_ERR_STATE, "vcrt_device_ptr: no storage images bound");
    *out = b->ptr;
    if (bytes) *bytes = b->bytes;
    return VCRT_OK;
}

// ---------------------------------------------------------------------------------------------- frames in flight (include/vcrt.h)
int vcrt_frames_begin(vcrt_ctx* c, uint32_t n) {
    if (!c) return VCRT_ERR_INVALID;
    if (n < 1 || n > VCRT_MAX_FRAMES_IN_FLIGHT) return fail(c, VCRT_ERR_INVALID, "vcrt_frames_begin: frames_in_flight must be 1..4");
    if (c->W == 0) return fail(c, VCRT_ERR_STATE, "vcrt_frames_begin: no storage images bound (vcrt_set_image_size)");
    NOFRAMES(c, "vcrt_frames_begin");
    CU(c, cudaSetDevice(c->device), "set device");
    CU(c, cudaStreamSynchronize(c->stream), "synchronize");   // the slots' streams start from a quiet context
    const size_t npix = (size_t)c->W * c->H;
    for (uint32_t i = 0; i < n; ++i) {
        FrameSlot& fs = c->frame_slot[i];
        if (!fs.stream) CU(c, cudaStreamCreateWithFlags(&fs.stream, cudaStreamNonBlocking), "create frame stream");
        if (!fs.folded) CU(c, cudaEventCreateWithFlags(&fs.folded, cudaEventDisableTiming), "create event");
        if (!fs.done) CU(c, cudaEventCreateWithFlags(&fs.done, cudaEventDisableTiming), "create event");
        if (!fs.counter) CU(c, cudaMalloc((void**)&fs.counter, 2 * sizeof(unsigned long long)), "allocate work counter");
        int rc;
        if ((rc = ensure(c, fs.image, npix * 4, "allocate frame image")) || (rc = ensure(c, fs.sample, npix * 16, "allocate frame sample buffer"))) return rc;
        fs.busy = false;
    }
    c->frames_n = (int)n;
    c->frame_next = 0;
    c->last_folded = nullptr;
    return VCRT_OK;
}

int vcrt_frame_wait(vcrt_ctx* c, uint32_t slot) {
    if (!c) return VCRT_ERR_INVALID;
    if (!c->frames_n || slot >= (uint32_t)c->frames_n) return fail(c, VCRT_ERR_INVALID, "vcrt_frame_wait: no such frame slot");
    FrameSlot& fs = c->frame_slot[slot];
    if (fs.busy) {
        CU(c, cudaSetDevice(c->device), "set device");
        CU(c, cudaEventSynchronize(fs.done), "wait for frame");
        fs.busy = false;
    }
    return VCRT_OK;
}

static int frame_submit_common(vcrt_ctx* c, const vcrt_render_params& p, uint32_t covW, uint32_t covH, uint32_t total_samples, float gamma, void* host_dst, size_t bytes,
                               uint32_t* slot_out) {
    if (host_dst && bytes != (size_t)c->W * c->H * 4) return fail(c, VCRT_ERR_INVALID, "frame submit: size mismatch");
    const uint32_t slot = c->frame_next;
    FrameSlot& fs = c->frame_slot[slot];
    int rc = vcrt_frame_wait(c, slot);   // main.cpp:325: wait for the fence of the slot about to be reused
    if (rc) return rc;
    CU(c, cudaSetDevice(c->device), "set device");
    const bool restart = p.accum_mode == VCRT_ACCUM_F32 && p.sample_begin == 0;
    KernelArgs a;
    if ((rc = render_common(c, p, covW, covH, &fs, (int)slot, restart, &a))) return rc;
    // fold in frame order.  One-launch kernels: the sample waits in the slot's buffer, the fold kernel applies it behind the previous
    // frame's fold.  Wavefront pipeline: its accumulate kernel already waited for that event; what is left is the resolve.
    if (a.sample_out && c->last_folded) CU(c, cudaStreamWaitEvent(fs.stream, c->last_folded, 0), "order frame folds");
    const float inv_total = 1.0f / (float)(total_samples ? total_samples : p.sample_begin + 1u);
    CU(c, launch_frame_fold(a, a.sample_out, (uchar4*)fs.image.ptr, inv_total, gamma > 0.0f ? 1.0f / gamma : 0.0f, fs.stream), "launch fold kernel");
    CU(c, cudaEventRecord(fs.folded, fs.stream), "record event");
    c->last_folded = fs.folded;
    if (host_dst) CU(c, cudaMemcpyAsync(host_dst, fs.image.ptr, bytes, cudaMemcpyDeviceToHost, fs.stream), "read frame back");
    CU(c, cudaEventRecord(fs.done, fs.stream), "record event");
    fs.busy = true;
    c->launches += 1;
    c->frame_next = (slot + 1u) % (uint32_t)c->frames_n;
    if (slot_out) *slot_out = slot;
    return VCRT_OK;
}

int vcrt_frame_submit(vcrt_ctx* c, const vcrt_render_params* p, uint32_t total_samples, float gamma, void* host_dst, size_t bytes, uint32_t* slot_out) {
    if (!c || !p) return fail(c, VCRT_ERR_INVALID, "vcrt_frame_submit: NULL argument");
    if (p->struct_size != sizeof(vcrt_render_params)) return fail(c, VCRT_ERR_INVALID, "vcrt_frame_submit: struct_size mismatch");
    if (!c->frames_n) return fail(c, VCRT_ERR_STATE, "vcrt_frame_submit: call vcrt_frames_begin first");
    if (p->sample_count > 1) return fail(c, VCRT_ERR_INVALID, "vcrt_frame_submit: a frame in flight renders one sample per pixel (sample_count 0 or 1)");
    const bool ref_cov = (p->flags & VCRT_FLAG_REF_DISPATCH_COVERAGE) != 0;
    return frame_submit_common(c, *p, ref_cov ? (c->W / 32) * 32 : c->W, ref_cov ? (c->H / 32) * 32 : c->H, total_samples, gamma, host_dst, bytes, slot_out);
}

int vcrt_frame_dispatch(vcrt_ctx* c, uint32_t gx, uint32_t gy, uint32_t gz, void* host_dst, size_t bytes, uint32_t* slot_out) {
    if (!c) return VCRT_ERR_INVALID;
    if (!c->frames_n) return fail(c, VCRT_ERR_STATE, "vcrt_frame_dispatch: call vcrt_frames_begin first");
    vcrt_render_params p;
    uint32_t covW, covH;
    int rc = dispatch_params(c, gx, gy, p, covW, covH);
    if (rc) return rc;
    if (gz == 0 || gx == 0 || gy == 0) covW = covH = 0;   // vkCmdDispatch with a zero dimension renders nothing; the frame is still presented
    return frame_submit_common(c, p, covW, covH, 0, 0.0f, host_dst, bytes, slot_out);
}

int vcrt_frames_end(vcrt_ctx* c) {
    if (!c) return VCRT_ERR_INVALID;
    if (!c->frames_n) return VCRT_OK;
    CU(c, cudaSetDevice(c->device), "set device");
    for (int i = 0; i < c->frames_n; ++i) {
        CU(c, cudaStreamSynchronize(c->frame_slot[i].stream), "synchronize");
        c->frame_slot[i].busy = false;
    }
    c->frames_n = 0;
    c->last_folded = nullptr;
    return VCRT_OK;
}

int vcrt_alloc_host(size_t bytes, void** out) {
    if (!out) return VCRT_ERR_INVALID;
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "allocate page-locked host memory");
    return VCRT_OK;
}

int vcrt_free_host(void* ptr) {
    if (!ptr) return VCRT_OK;
    cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "free page-locked host memory");
    return VCRT_OK;
}

int vcrt_synchronize(vcrt_ctx* c) {
    if (!c) return VCRT_ERR_INVALID;
    CU(c, cudaSetDevice(c->device), "set device");
    CU(c, cudaStreamSynchronize(c->stream), "synchronize");
    for (int i = 0; i < c->frames_n; ++i) CU(c, cudaStreamSynchronize(c->frame_slot[i].stream), "synchronize");
    return VCRT_OK;
}

static int drain_events(vcrt_ctx* c) {
    CU(c, cudaStreamSynchronize(c->stream), "synchronize");
    for (int i = 0; i < c->frames_n; ++i) CU(c, cudaStreamSynchronize(c->frame_slot[i].stream), "synchronize");   // frames in flight record their events on the slots' streams
    for (auto& ev : c->events) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) c->kernel_ms += ms;
        c->ev_pool.push_back(ev.first);
        c->ev_pool.push_back(ev.second);
    }
    c->events.clear();
    c->trace_timer.drain(&c->trace_ms, &c->trace_launches);
    return VCRT_OK;
}

int vcrt_get_counters(vcrt_ctx* c, vcrt_counters* out) {
    if (!c || !out) return fail(c, VCRT_ERR_INVALID, "vcrt_get_counters: NULL argument");
    CU(c, cudaSetDevice(c->device), "set device");
    int rc = drain_events(c);
    if (rc) return rc;
    unsigned long long h[6];
    CU(c, cudaMemcpy(h, c->d_counters, sizeof h, cudaMemcpyDeviceToHost), "read counters");
    out->rays = h[0]; out->nodes = h[1]; out->triangles = h[2];
    out->kernel_ms = c->kernel_ms;
    out->launches = c->launches;
    out->trace_ms = c->trace_ms;
    out->trace_launches = c->trace_launches;
    out->traversals = h[4];
    out->primary_rays = h[5];
    out->primary_trace_ms = c->trace_timer.primary_ms;
    return VCRT_OK;
}

int vcrt_reset_counters(vcrt_ctx* c) {
    if (!c) return VCRT_ERR_INVALID;
    CU(c, cudaSetDevice(c->device), "set device");
    int rc = drain_events(c);
    if (rc) return rc;
    CU(c, cudaMemset(c->d_counters, 0, 8 * sizeof(unsigned long long)), "reset counters");
    c->trace_timer.primary_ms = 0.0;
    c->kernel_ms = 0.0;
    c->launches = 0;
    c->trace_ms = 0.0;
    c->trace_launches = 0;
    return VCRT_OK;
}

}  // extern "C"
