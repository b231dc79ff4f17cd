// Tensor-parallel exchange steps (extension -- the reference is single-device): one process per GPU, NCCL over NVLink.
//
//   uzu_tp_all_reduce_encode : f32 partial sums [count] of a row-parallel projection (out / down) are summed over the ranks with
//                              ncclAllReduce on the context's stream, then rounded to bf16 ONCE (the unsharded kernel's single
//                              rounding point) by a small conversion kernel -> the next norm / GEMV consumes the usual bf16 row.
//   uzu_tp_all_gather_encode : vocab-parallel readout: [rows, cols_local] bf16 logits of every rank -> [rows, ranks * cols_local].
//
// NCCL is resolved at run time (dlopen "libnccl.so.2", or the path in UZU_NCCL_LIB): the library keeps loading on hosts without
// NCCL / without a GPU, and a process that already imported torch reuses torch's bundled NCCL. The unique id is created by rank 0
// (uzu_tp_get_unique_id) and distributed by the host program (bench.py: torch.distributed broadcast).
// Both collectives are stream-ordered and capturable, so a decode step with its 2 all-reduces per layer stays ONE CUDA graph.
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace uzu {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
};
constexpr int NCCL_UINT8 = 1, NCCL_FLOAT32 = 7, NCCL_SUM = 0;   // nccl.h: ncclDataType_t / ncclRedOp_t

static NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("UZU_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.error = std::string("cannot load NCCL (set UZU_NCCL_LIB to libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
            return;
        }
        auto sym = [&](const char* name) {
            void* p = dlsym(api.handle, name);
            if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + name;
            return p;
        };
        api.GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
        api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))sym("ncclCommInitRank");
        api.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
        api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclAllReduce");
        api.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))sym("ncclAllGather");
        api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    });
    return api;
}

static std::string nccl_err(int rc) {
    NcclApi& n = nccl();
    return n.GetErrorString ? n.GetErrorString(rc) : ("NCCL error " + std::to_string(rc));
}

// f32 -> bf16 (RNE), 4 elements per thread where aligned
__global__ void __launch_bounds__(256) tp_round_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, uint32_t count) {
    pdl_wait();
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    if (i + 3 < count) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&a);
        o.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(dst + i) = o;
    } else {
        for (uint32_t j = i; j < count; ++j) dst[j] = f2bf(src[j]);
    }
}

// gathered [ranks][rows][cols_local] -> [rows][ranks * cols_local] (2-byte elements)
__global__ void __launch_bounds__(256) tp_interleave_rows_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, uint32_t ranks,
                                                                 uint32_t rows, uint32_t cols_local) {
    const size_t total = (size_t)ranks * rows * cols_local;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t)(i % cols_local);
        const uint32_t r = (uint32_t)((i / cols_local) % rows);
        const uint32_t k = (uint32_t)(i / ((size_t)cols_local * rows));
        dst[(size_t)r * ranks * cols_local + (size_t)k * cols_local + c] = src[i];
    }
}

}  // namespace uzu

using namespace uzu;

extern "C" {

uzu_status uzu_tp_get_unique_id(uint8_t* out128) {
    if (!out128) return fail(UZU_ERROR_INVALID_ARGUMENT, "tp_get_unique_id: null output");
    NcclApi& n = nccl();
    if (!n.error.empty()) return fail(UZU_ERROR_UNSUPPORTED, n.error);
    NcclUniqueId id;
    const int rc = n.GetUniqueId(&id);
    if (rc != 0) return fail(UZU_ERROR_CUDA, "ncclGetUniqueId: " + nccl_err(rc));
    memcpy(out128, id.internal, 128);
    return UZU_OK;
}

uzu_status uzu_context_tp_init(uzu_context* ctx, uint32_t rank, uint32_t size, const uint8_t* unique_id128) {
    if (!ctx || !unique_id128 || size == 0 || rank >= size) return fail(UZU_ERROR_INVALID_ARGUMENT, "tp_init: bad arguments");
    if (ctx->nccl_comm) return fail(UZU_ERROR_INVALID_ARGUMENT, "tp_init: the context already has a communicator");
    NcclApi& n = nccl();
    if (!n.error.empty()) return fail(UZU_ERROR_UNSUPPORTED, n.error);
    UZU_CUDA_TRY(cudaSetDevice(ctx->device));
    NcclUniqueId id;
    memcpy(id.internal, unique_id128, 128);
    NcclComm comm = nullptr;
    const int rc = n.CommInitRank(&comm, (int)size, id, (int)rank);
    if (rc != 0) return fail(UZU_ERROR_CUDA, "ncclCommInitRank: " + nccl_err(rc));
    ctx->nccl_comm = comm;
    ctx->tp_rank = rank;
    ctx->tp_size = size;
    return UZU_OK;
}

void uzu_context_tp_destroy(uzu_context* ctx) {
    if (!ctx || !ctx->nccl_comm) return;
    NcclApi& n = nccl();
    if (n.CommDestroy) n.CommDestroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->tp_rank = 0;
    ctx->tp_size = 1;
}

uint32_t uzu_context_tp_size(const uzu_context* ctx) { return ctx ? ctx->tp_size : 0; }
uint32_t uzu_context_tp_rank(const uzu_context* ctx) { return ctx ? ctx->tp_rank : 0; }

void uzu_tp_all_reduce_encode(uzu_command_buffer* cmd, uint64_t partial_f32, uint32_t count, uint64_t out_bf16) {
    if (!encodable(cmd, "tp_all_reduce")) return;
    uzu_context* ctx = cmd->ctx;
    if (!partial_f32 || !out_bf16 || count == 0) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "tp_all_reduce: null operand / empty");
        return;
    }
    if (ctx->tp_size > 1) {
        if (!ctx->nccl_comm) {
            cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "tp_all_reduce: the context has no communicator (uzu_context_tp_init)");
            return;
        }
        const int rc = nccl().AllReduce((const void*)partial_f32, (void*)partial_f32, count, NCCL_FLOAT32, NCCL_SUM, ctx->nccl_comm, ctx->stream);
        if (rc != 0) cmd->record_error(UZU_ERROR_CUDA, "ncclAllReduce: " + nccl_err(rc));
        cmd->launches++;
    }
    launch(cmd, "tp_round_bf16_kernel", tp_round_bf16_kernel, dim3((count / 4 + 256) / 256), dim3(256), 0, (const float*)partial_f32,
           (__nv_bfloat16*)out_bf16, count);
}

void uzu_tp_all_gather_encode(uzu_command_buffer* cmd, const uzu_tp_all_gather_args* a) {
    if (!encodable(cmd, "tp_all_gather")) return;
    uzu_context* ctx = cmd->ctx;
    if (!a || !a->src || !a->dst || a->rows == 0 || a->cols_local == 0 || (a->rows > 1 && ctx->tp_size > 1 && !a->scratch)) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "tp_all_gather: null operand / empty (rows > 1 needs scratch)");
        return;
    }
    const size_t bytes_local = (size_t)a->rows * a->cols_local * 2;
    if (ctx->tp_size <= 1) {
        cudaError_t e = cudaMemcpyAsync((void*)a->dst, (const void*)a->src, bytes_local, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) cmd->record_error(UZU_ERROR_CUDA, std::string("tp_all_gather copy: ") + cudaGetErrorString(e));
        return;
    }
    if (!ctx->nccl_comm) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, "tp_all_gather: the context has no communicator (uzu_context_tp_init)");
        return;
    }
    // one row: the gathered layout [rank][cols_local] already is the logits row; more rows: gather into scratch, interleave
    void* recv = a->rows == 1 ? (void*)a->dst : (void*)a->scratch;
    const int rc = nccl().AllGather((const void*)a->src, recv, bytes_local, NCCL_UINT8, ctx->nccl_comm, ctx->stream);
    if (rc != 0) cmd->record_error(UZU_ERROR_CUDA, "ncclAllGather: " + nccl_err(rc));
    cmd->launches++;
    if (a->rows > 1) {
        const size_t total = (size_t)ctx->tp_size * a->rows * a->cols_local;
        const uint32_t blocks = (uint32_t)std::min<size_t>((total + 255) / 256, 4096);
        tp_interleave_rows_kernel<<<blocks, 256, 0, ctx->stream>>>((const uint16_t*)a->scratch, (uint16_t*)a->dst, ctx->tp_size, a->rows, a->cols_local);
        after_launch(cmd, "tp_interleave_rows_kernel");
    }
}

}  // extern "C"
