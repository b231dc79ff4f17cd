// Context / buffers / command buffer: the CUDA side of backends/common/{context,buffer/*,command_buffer}.rs.
#include <cuda.h>
#include <cuda_profiler_api.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace uzu {

// The driver API (VMM for SparseBuffer) is resolved at run time through the runtime's
// cudaGetDriverEntryPoint, so the library has no link-time dependency on libcuda.so.1 and still loads
// (and exports its symbols) on a machine without a GPU driver.
struct DriverApi {
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    bool ok = false;
};
static DriverApi g_drv;

static bool load_driver_api() {
    if (g_drv.ok) return true;
    auto get = [](const char* name, void** fn) {
        cudaDriverEntryPointQueryResult q;
        return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
    };
    bool ok = get("cuGetErrorString", (void**)&g_drv.GetErrorString) && get("cuDeviceGet", (void**)&g_drv.DeviceGet) &&
              get("cuDeviceGetAttribute", (void**)&g_drv.DeviceGetAttribute) &&
              get("cuMemGetAllocationGranularity", (void**)&g_drv.MemGetAllocationGranularity) &&
              get("cuMemAddressReserve", (void**)&g_drv.MemAddressReserve) && get("cuMemAddressFree", (void**)&g_drv.MemAddressFree) &&
              get("cuMemCreate", (void**)&g_drv.MemCreate) && get("cuMemRelease", (void**)&g_drv.MemRelease) &&
              get("cuMemMap", (void**)&g_drv.MemMap) && get("cuMemUnmap", (void**)&g_drv.MemUnmap) &&
              get("cuMemSetAccess", (void**)&g_drv.MemSetAccess);
    g_drv.ok = ok;
    cudaGetLastError();
    return ok;
}

static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }
uzu_status fail(uzu_status st, const std::string& msg) {
    g_last_error = msg;
    return st;
}

static void track_alloc(uzu_context* ctx, size_t bytes) {
    ctx->live_bytes += bytes;
    if (ctx->live_bytes > ctx->peak_bytes) ctx->peak_bytes = ctx->live_bytes;
}

}  // namespace uzu

using namespace uzu;

struct uzu_buffer {
    uzu_context* ctx;
    void* dev;
    void* host;
    size_t size;
    uzu_buffer_kind kind;
};

struct uzu_sparse_buffer {
    uzu_context* ctx;
    CUdeviceptr base;
    size_t capacity;      // reserved VA, multiple of page
    size_t page;
    std::vector<CUmemGenericAllocationHandle> handles;  // 0 = unmapped
};

extern "C" {

const char* uzu_last_error(void) { return g_last_error.c_str(); }
const char* uzu_version(void) { return "uzu_b200 0.1 (sm_100a)"; }

size_t uzu_abi_struct_size(const char* name) {
#define UZU_SZ(T) if (!strcmp(name, #T)) return sizeof(T);
    UZU_SZ(uzu_fused_linear_args) UZU_SZ(uzu_matmul_args) UZU_SZ(uzu_normalization_args) UZU_SZ(uzu_qkv_norm_args) UZU_SZ(uzu_attention_prepare_args)
    UZU_SZ(uzu_attention_args) UZU_SZ(uzu_attention_two_pass2_args) UZU_SZ(uzu_kv_cache_update_args) UZU_SZ(uzu_gated_act_mul_args)
    UZU_SZ(uzu_quantized_embedding_lookup_args) UZU_SZ(uzu_unified_sampling_args) UZU_SZ(uzu_delta_net_conv_update_args)
    UZU_SZ(uzu_delta_net_update_args) UZU_SZ(uzu_engine_options) UZU_SZ(uzu_sampling_method) UZU_SZ(uzu_model_info)
    UZU_SZ(uzu_ring_params) UZU_SZ(uzu_trie_node) UZU_SZ(uzu_kv_copy) UZU_SZ(uzu_delta_net_fused_update_args) UZU_SZ(uzu_qk_norm_config) UZU_SZ(uzu_attention_prepare_norm_args) UZU_SZ(uzu_tp_all_gather_args) UZU_SZ(uzu_activation_transform_args)
#undef UZU_SZ
    return 0;
}

uzu_status uzu_context_create(int device_ordinal, uzu_context** out) {
    if (!out) return fail(UZU_ERROR_INVALID_ARGUMENT, "uzu_context_create: out is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(UZU_ERROR_NO_DEVICE, std::string("uzu_context_create: no CUDA device (") +
                                             (e == cudaSuccess ? "count = 0" : cudaGetErrorString(e)) +
                                             "); this backend has no CPU fallback");
    if (device_ordinal < 0) {
        const char* env = std::getenv("UZU_DEVICE");
        if (env) device_ordinal = std::atoi(env);
        else UZU_CUDA_TRY(cudaGetDevice(&device_ordinal));
    }
    if (device_ordinal >= count) return fail(UZU_ERROR_INVALID_ARGUMENT, "uzu_context_create: device ordinal out of range");
    UZU_CUDA_TRY(cudaSetDevice(device_ordinal));
    cudaDeviceProp prop;
    UZU_CUDA_TRY(cudaGetDeviceProperties(&prop, device_ordinal));
    if (prop.major != 10)
        return fail(UZU_ERROR_UNSUPPORTED, std::string("uzu_context_create: device '") + prop.name +
                                               "' is not sm_100 (kernels are built for sm_100a only)");
    auto* ctx = new uzu_context();
    ctx->device = device_ordinal;
    ctx->sm_count = prop.multiProcessorCount;
    UZU_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    // split-K workspace: 16 MiB of f32 partials + 64 Ki tile counters (zeroed once; kernels reset what they use)
    ctx->splitk_ws_bytes = 16u << 20;
    ctx->splitk_counter_count = 1u << 16;
    UZU_CUDA_TRY(cudaMalloc(&ctx->splitk_ws, ctx->splitk_ws_bytes));
    UZU_CUDA_TRY(cudaMalloc(&ctx->splitk_counters, ctx->splitk_counter_count * sizeof(unsigned int)));
    UZU_CUDA_TRY(cudaMemset(ctx->splitk_counters, 0, ctx->splitk_counter_count * sizeof(unsigned int)));
    ctx->attn_ws_bytes = 8u << 20;
    UZU_CUDA_TRY(cudaMalloc(&ctx->attn_ws, ctx->attn_ws_bytes));
    UZU_CUDA_TRY(cudaMalloc(&ctx->attn_counters, 65536 * sizeof(unsigned int)));
    UZU_CUDA_TRY(cudaMemset(ctx->attn_counters, 0, 65536 * sizeof(unsigned int)));
    UZU_CUDA_TRY(cudaMalloc(&ctx->sampling_ws, 64 * 1024));
    UZU_CUDA_TRY(cudaMemset(ctx->sampling_ws, 0, 64 * 1024));
    // VMM support (SparseBuffer)
    int vmm = 0;
    CUdevice cu_dev;
    if (load_driver_api() && g_drv.DeviceGet(&cu_dev, device_ordinal) == CUDA_SUCCESS &&
        g_drv.DeviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, cu_dev) == CUDA_SUCCESS && vmm) {
        CUmemAllocationProp p = {};
        p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        p.location.id = device_ordinal;
        size_t gran = 0;
        if (g_drv.MemGetAllocationGranularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_MINIMUM) == CUDA_SUCCESS && gran) {
            ctx->vmm_supported = true;
            ctx->vmm_granularity = gran;
        }
    }
    UZU_CUDA_TRY(cudaDeviceSynchronize());
    *out = ctx;
    return UZU_OK;
}

void uzu_context_destroy(uzu_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    uzu_context_tp_destroy(ctx);
    cudaFree(ctx->splitk_ws);
    cudaFree(ctx->splitk_counters);
    cudaFree(ctx->sampling_ws);
    cudaFree(ctx->attn_ws);
    cudaFree(ctx->attn_counters);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

uzu_status uzu_context_synchronize(uzu_context* ctx) {
    UZU_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return UZU_OK;
}

uzu_status uzu_context_peak_memory_usage(uzu_context* ctx, size_t* out_bytes) {
    if (!ctx || !out_bytes) return fail(UZU_ERROR_INVALID_ARGUMENT, "peak_memory_usage: null argument");
    *out_bytes = ctx->peak_bytes;
    return UZU_OK;
}

uint32_t uzu_context_device_capabilities(uzu_context* ctx) { return ctx && ctx->vmm_supported ? UZU_CAP_SPARSE_BUFFERS : 0u; }

uzu_status uzu_context_start_capture(uzu_context*, const char*) {
    UZU_CUDA_TRY(cudaProfilerStart());
    return UZU_OK;
}
uzu_status uzu_context_stop_capture(uzu_context*) {
    UZU_CUDA_TRY(cudaProfilerStop());
    return UZU_OK;
}
int uzu_context_device(uzu_context* ctx) { return ctx->device; }
int uzu_context_sm_count(uzu_context* ctx) { return ctx->sm_count; }
void* uzu_context_stream(uzu_context* ctx) { return (void*)ctx->stream; }

// ---- dense buffers ----------------------------------------------------------------------------
uzu_status uzu_buffer_create(uzu_context* ctx, size_t size, uzu_buffer_kind kind, uzu_buffer** out) {
    if (!ctx || !out) return fail(UZU_ERROR_INVALID_ARGUMENT, "uzu_buffer_create: null argument");
    if (size == 0) size = 16;
    UZU_CUDA_TRY(cudaSetDevice(ctx->device));
    auto* b = new uzu_buffer{ctx, nullptr, nullptr, size, kind};
    cudaError_t e = cudaSuccess;
    switch (kind) {
        case UZU_BUFFER_MANAGED:
            e = cudaMallocManaged(&b->dev, size, cudaMemAttachGlobal);
            if (e == cudaSuccess) {
                b->host = b->dev;
                cudaMemAdvise(b->dev, size, cudaMemAdviseSetPreferredLocation, ctx->device);
            }
            break;
        case UZU_BUFFER_PINNED_HOST:
            e = cudaHostAlloc(&b->host, size, cudaHostAllocMapped | cudaHostAllocPortable);
            if (e == cudaSuccess) e = cudaHostGetDevicePointer(&b->dev, b->host, 0);
            break;
        case UZU_BUFFER_DEVICE:
            e = cudaMalloc(&b->dev, size);
            break;
        default:
            delete b;
            return fail(UZU_ERROR_INVALID_ARGUMENT, "uzu_buffer_create: unknown buffer kind");
    }
    if (e != cudaSuccess) {
        delete b;
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? UZU_ERROR_OUT_OF_MEMORY : UZU_ERROR_CUDA,
                    std::string("uzu_buffer_create: ") + cudaGetErrorString(e));
    }
    if (kind != UZU_BUFFER_PINNED_HOST) track_alloc(ctx, size);
    *out = b;
    return UZU_OK;
}

void uzu_buffer_destroy(uzu_buffer* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    if (b->kind == UZU_BUFFER_PINNED_HOST) cudaFreeHost(b->host);
    else {
        cudaFree(b->dev);
        b->ctx->live_bytes -= b->size;
    }
    delete b;
}

uint64_t uzu_buffer_gpu_ptr(const uzu_buffer* b) { return (uint64_t)b->dev; }
void* uzu_buffer_cpu_ptr(const uzu_buffer* b) { return b->host; }
size_t uzu_buffer_size(const uzu_buffer* b) { return b->size; }

uzu_status uzu_buffer_make_resident(uzu_context* ctx, uzu_buffer* b) {
    if (!b || b->kind != UZU_BUFFER_MANAGED) return UZU_OK;
    UZU_CUDA_TRY(cudaMemPrefetchAsync(b->dev, b->size, ctx->device, ctx->stream));
    return UZU_OK;
}

// ---- sparse buffers (CUDA virtual memory management) -----------------------------------------
#define UZU_CU_TRY(expr)                                                                 \
    do {                                                                                 \
        CUresult _r = (expr);                                                            \
        if (_r != CUDA_SUCCESS) {                                                        \
            const char* _s = nullptr;                                                    \
            g_drv.GetErrorString(_r, &_s);                                                   \
            return fail(_r == CUDA_ERROR_OUT_OF_MEMORY ? UZU_ERROR_OUT_OF_MEMORY : UZU_ERROR_CUDA, \
                        std::string(#expr) + ": " + (_s ? _s : "?"));                    \
        }                                                                                \
    } while (0)

uzu_status uzu_sparse_buffer_create(uzu_context* ctx, size_t capacity, uzu_sparse_buffer** out) {
    if (!ctx || !out) return fail(UZU_ERROR_INVALID_ARGUMENT, "uzu_sparse_buffer_create: null argument");
    if (!ctx->vmm_supported) return fail(UZU_ERROR_UNSUPPORTED, "sparse buffers: CUDA VMM is not supported on this device");
    UZU_CUDA_TRY(cudaSetDevice(ctx->device));
    size_t page = ctx->vmm_granularity;
    size_t cap = (capacity + page - 1) / page * page;
    auto* sb = new uzu_sparse_buffer{ctx, 0, cap, page, {}};
    CUresult r = g_drv.MemAddressReserve(&sb->base, cap, page, 0, 0);
    if (r != CUDA_SUCCESS) {
        delete sb;
        return fail(UZU_ERROR_CUDA, "cuMemAddressReserve failed");
    }
    sb->handles.assign(cap / page, 0);
    *out = sb;
    return UZU_OK;
}

uzu_status uzu_sparse_buffer_unmap(uzu_sparse_buffer* sb, const uint32_t* pages, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        uint32_t p = pages[i];
        if (p >= sb->handles.size()) return fail(UZU_ERROR_INVALID_ARGUMENT, "sparse unmap: page out of range");
        if (!sb->handles[p]) continue;
        UZU_CU_TRY(g_drv.MemUnmap(sb->base + (size_t)p * sb->page, sb->page));
        UZU_CU_TRY(g_drv.MemRelease(sb->handles[p]));
        sb->handles[p] = 0;
        sb->ctx->live_bytes -= sb->page;
    }
    return UZU_OK;
}

uzu_status uzu_sparse_buffer_map(uzu_sparse_buffer* sb, const uint32_t* pages, size_t n) {
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = sb->ctx->device;
    CUmemAccessDesc access = {};
    access.location = prop.location;
    access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    for (size_t i = 0; i < n; ++i) {
        uint32_t p = pages[i];
        if (p >= sb->handles.size()) return fail(UZU_ERROR_INVALID_ARGUMENT, "sparse map: page out of range");
        if (sb->handles[p]) continue;
        CUmemGenericAllocationHandle h;
        UZU_CU_TRY(g_drv.MemCreate(&h, sb->page, &prop, 0));
        CUdeviceptr va = sb->base + (size_t)p * sb->page;
        CUresult r = g_drv.MemMap(va, sb->page, 0, h, 0);
        if (r == CUDA_SUCCESS) r = g_drv.MemSetAccess(va, sb->page, &access, 1);
        if (r != CUDA_SUCCESS) {
            g_drv.MemRelease(h);
            return fail(UZU_ERROR_CUDA, "cuMemMap/cuMemSetAccess failed");
        }
        sb->handles[p] = h;
        track_alloc(sb->ctx, sb->page);
    }
    return UZU_OK;
}

void uzu_sparse_buffer_destroy(uzu_sparse_buffer* sb) {
    if (!sb) return;
    cudaSetDevice(sb->ctx->device);
    cudaStreamSynchronize(sb->ctx->stream);
    for (size_t p = 0; p < sb->handles.size(); ++p)
        if (sb->handles[p]) {
            g_drv.MemUnmap(sb->base + p * sb->page, sb->page);
            g_drv.MemRelease(sb->handles[p]);
            sb->ctx->live_bytes -= sb->page;
        }
    g_drv.MemAddressFree(sb->base, sb->capacity);
    delete sb;
}

uint64_t uzu_sparse_buffer_gpu_ptr(const uzu_sparse_buffer* sb) { return (uint64_t)sb->base; }
size_t uzu_sparse_buffer_size(const uzu_sparse_buffer* sb) { return sb->capacity; }
size_t uzu_sparse_buffer_page_size_bytes(const uzu_sparse_buffer* sb) { return sb->page; }

// ---- command buffers --------------------------------------------------------------------------
uzu_status uzu_command_buffer_create(uzu_context* ctx, const char* name, uzu_command_buffer** out) {
    if (!ctx || !out) return fail(UZU_ERROR_INVALID_ARGUMENT, "uzu_command_buffer_create: null argument");
    UZU_CUDA_TRY(cudaSetDevice(ctx->device));
    auto* c = new uzu_command_buffer();
    c->ctx = ctx;
    c->name = name ? name : "";
    UZU_CUDA_TRY(cudaEventCreate(&c->ev_begin));
    UZU_CUDA_TRY(cudaEventCreate(&c->ev_end));
    *out = c;
    return UZU_OK;
}

void uzu_command_buffer_destroy(uzu_command_buffer* c) {
    if (!c) return;
    if (c->state == uzu_command_buffer::Pending) cudaEventSynchronize(c->ev_end);
    cudaEventDestroy(c->ev_begin);
    cudaEventDestroy(c->ev_end);
    delete c;
}

uzu_status uzu_command_buffer_start_encoding(uzu_command_buffer* c) {
    if (!c || c->state != uzu_command_buffer::Initial) return fail(UZU_ERROR_INVALID_ARGUMENT, "start_encoding: not Initial");
    UZU_CUDA_TRY(cudaSetDevice(c->ctx->device));
    UZU_CUDA_TRY(cudaEventRecord(c->ev_begin, c->ctx->stream));
    c->state = uzu_command_buffer::Encoding;
    return UZU_OK;
}

void uzu_command_buffer_encode_copy(uzu_command_buffer* c, uint64_t src, uint64_t dst, size_t bytes) {
    if (!encodable(c, "encode_copy") || bytes == 0) return;
    cudaError_t e = cudaMemcpyAsync((void*)dst, (const void*)src, bytes, cudaMemcpyDefault, c->ctx->stream);
    if (e != cudaSuccess) c->record_error(UZU_ERROR_CUDA, std::string("encode_copy: ") + cudaGetErrorString(e));
}

void uzu_command_buffer_encode_fill(uzu_command_buffer* c, uint64_t dst, size_t bytes, uint8_t value) {
    if (!encodable(c, "encode_fill") || bytes == 0) return;
    cudaError_t e = cudaMemsetAsync((void*)dst, value, bytes, c->ctx->stream);
    if (e != cudaSuccess) c->record_error(UZU_ERROR_CUDA, std::string("encode_fill: ") + cudaGetErrorString(e));
}

// One in-order stream: every hazard the reference's HazardTracker reports is already ordered
// (same as backends/cpu/command_buffer.rs:106-111).
void uzu_command_buffer_encode_barrier(uzu_command_buffer*, uint32_t, uint32_t) {}
void uzu_command_buffer_push_debug_group(uzu_command_buffer*, const char*) {}
void uzu_command_buffer_pop_debug_group(uzu_command_buffer*) {}

uzu_status uzu_command_buffer_end_encoding(uzu_command_buffer* c) {
    if (!c || c->state != uzu_command_buffer::Encoding) return fail(UZU_ERROR_INVALID_ARGUMENT, "end_encoding: not Encoding");
    UZU_CUDA_TRY(cudaEventRecord(c->ev_end, c->ctx->stream));
    c->state = uzu_command_buffer::Executable;
    return UZU_OK;
}

// Work was enqueued on the stream while encoding; submit only marks the transition (the reference's
// Metal commit). Kept separate so a Rust `Executable::submit` maps 1:1.
uzu_status uzu_command_buffer_submit(uzu_command_buffer* c) {
    if (!c || c->state != uzu_command_buffer::Executable) return fail(UZU_ERROR_INVALID_ARGUMENT, "submit: not Executable");
    c->state = uzu_command_buffer::Pending;
    return UZU_OK;
}

uzu_status uzu_command_buffer_wait_until_completed(uzu_command_buffer* c) {
    if (!c || c->state != uzu_command_buffer::Pending) return fail(UZU_ERROR_INVALID_ARGUMENT, "wait_until_completed: not Pending");
    cudaError_t e = cudaEventSynchronize(c->ev_end);
    c->state = uzu_command_buffer::Completed;
    if (e != cudaSuccess) return fail(UZU_ERROR_CUDA, std::string("wait_until_completed: ") + cudaGetErrorString(e));
    if (c->sticky != UZU_OK) return fail(c->sticky, c->sticky_msg);
    return UZU_OK;
}

uzu_status uzu_command_buffer_gpu_execution_time(uzu_command_buffer* c, double* out_seconds) {
    if (!c || c->state != uzu_command_buffer::Completed) return fail(UZU_ERROR_INVALID_ARGUMENT, "gpu_execution_time: not Completed");
    float ms = 0.0f;
    UZU_CUDA_TRY(cudaEventElapsedTime(&ms, c->ev_begin, c->ev_end));
    *out_seconds = (double)ms * 1e-3;
    return UZU_OK;
}

uint64_t uzu_command_buffer_launch_count(const uzu_command_buffer* c) { return c ? c->launches : 0; }

}  // extern "C"
