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


// ---- one-shot all-reduce over peer memory (NVLink / NVSwitch P2P), for the 8-32 KB decode messages ----------------------------------
// NOT yet run on hardware (round 1 had no multi-GPU time): enabled only after uzu_tp_p2p_export / uzu_tp_p2p_import succeeded, and
// the 2-rank test that covers it is opt-in. NCCL stays the default exchange and the fallback for large (prefill) messages.
//
// Every rank owns an exchange buffer (cudaMalloc, opened by the peers through CUDA IPC):
//     data [2 slot sets][P ranks][capacity] f32      partials the peers push into this rank
//     flags[2 slot sets][P ranks][CTAS]     u32      "chunk c of rank r, call number e, has landed"
//     epoch[CTAS]                           u32      this rank's call counter per CTA (device resident: the call is captured in a CUDA
//                                                    graph, so nothing that changes per call may be a kernel argument)
// CTA c of every rank owns the same slice of the vector. Per call e = epoch[c] + 1, slot set = e & 1:
//   1. push my slice of my partial into data[set][me][slice] of every peer (16-byte P2P stores), __threadfence_system, then publish
//      flags[set][me][c] = e on every peer (release, system scope);
//   2. spin (acquire, system scope) until flags[set][r][c] == e for every peer r in my own buffer;
//   3. out[i] = bf16( sum over ranks IN RANK ORDER of the f32 partials ) -- the same value on every rank, rounded once;
//   4. epoch[c] = e.
// Two slot sets are enough: a peer publishes call e only after it finished its own call e - 1, so when some rank overwrites set
// (e + 1) & 1 nobody can still be reading call e - 1 from it. No CTA waits for another CTA of its own rank: no grid barrier.
constexpr int TP_P2P_CTAS = 8;
constexpr int TP_P2P_MAX_RANKS = 8;

struct TpP2pView {
    float* data[TP_P2P_MAX_RANKS];            // data region of rank r's exchange buffer (peer-mapped; [me] = local)
    unsigned int* flags[TP_P2P_MAX_RANKS];    // flags region of rank r's exchange buffer
    unsigned int* epoch;                      // local
    uint32_t rank, size, capacity;            // capacity = f32 elements per (slot set, rank)
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) tp_p2p_all_reduce_kernel(const TpP2pView v, const float* __restrict__ partial, __nv_bfloat16* __restrict__ out,
                                                                uint32_t count) {
    pdl_wait();
    const uint32_t c = blockIdx.x, me = v.rank, P = v.size;
    __shared__ unsigned int s_e;
    if (threadIdx.x == 0) s_e = v.epoch[c] + 1u;
    __syncthreads();
    const unsigned int e = s_e;
    const uint32_t set = e & 1u;
    // slice of this CTA in units of float4 (count is a multiple of 4: model_dim rows)
    const uint32_t vec = count / 4u, per = (vec + TP_P2P_CTAS - 1) / TP_P2P_CTAS;
    const uint32_t v0 = min(vec, c * per), v1 = min(vec, v0 + per);
    const size_t slot_me = ((size_t)set * P + me) * v.capacity;
    // 1. push
    for (uint32_t r = 0; r < P; ++r) {
        if (r == me) continue;
        float4* dst = reinterpret_cast<float4*>(v.data[r] + slot_me);
        for (uint32_t i = v0 + threadIdx.x; i < v1; i += blockDim.x) dst[i] = reinterpret_cast<const float4*>(partial)[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < P && threadIdx.x != me) st_release_sys(v.flags[threadIdx.x] + ((size_t)set * P + me) * TP_P2P_CTAS + c, e);
    // 2. wait for the peers' pushes into MY buffer
    if (threadIdx.x < P && threadIdx.x != me) {
        const unsigned int* f = v.flags[me] + ((size_t)set * P + threadIdx.x) * TP_P2P_CTAS + c;
        unsigned long long t0 = 0;
        unsigned int spins = 0;
        while ((int)(ld_acquire_sys(f) - e) < 0) {
            if ((++spins & 4095u) == 0) {     // a peer that never arrives (crashed rank) must not hang this GPU forever
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t0 == 0) t0 = t1;
                else if (t1 - t0 > 20000000000ull) __trap();
            }
        }
    }
    __syncthreads();
    // 3. reduce in rank order (peer slots are read with L1 bypass: the same addresses are reused every second call)
    for (uint32_t i = v0 + threadIdx.x; i < v1; i += blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t r = 0; r < P; ++r) {
            const float4 x = r == me ? reinterpret_cast<const float4*>(partial)[i]
                                     : __ldcg(reinterpret_cast<const float4*>(v.data[me] + ((size_t)set * P + r) * v.capacity) + i);
            acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
        }
        __nv_bfloat162 a = __floats2bfloat162_rn(acc.x, acc.y), b = __floats2bfloat162_rn(acc.z, acc.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&a);
        o.y = *reinterpret_cast<uint32_t*>(&b);
        reinterpret_cast<uint2*>(out)[i] = o;
    }
    // 4.
    if (threadIdx.x == 0) v.epoch[c] = e;
}

struct TpP2pState {
    void* local = nullptr;                      // this rank's exchange buffer
    void* peer[TP_P2P_MAX_RANKS] = {};          // opened peer buffers ([me] = local)
    uint32_t capacity = 0;                      // f32 elements per (slot set, rank)
    bool ready = false;
};
static size_t tp_p2p_bytes(uint32_t size, uint32_t capacity) {
    return (size_t)2 * size * capacity * 4 + (size_t)2 * size * TP_P2P_CTAS * 4 + TP_P2P_CTAS * 4;
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
    if (ctx && ctx->tp_p2p) {
        TpP2pState* st = static_cast<TpP2pState*>(ctx->tp_p2p);
        for (uint32_t r = 0; r < ctx->tp_size && r < (uint32_t)TP_P2P_MAX_RANKS; ++r)
            if (st->peer[r] && st->peer[r] != st->local) cudaIpcCloseMemHandle(st->peer[r]);
        cudaFree(st->local);
        delete st;
        ctx->tp_p2p = nullptr;
    }
    if (!ctx || !ctx->nccl_comm) return;
    NcclApi& n = nccl();
    if (n.CommDestroy) n.CommDestroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->tp_rank = 0;
    ctx->tp_size = 1;
}


uzu_status uzu_tp_p2p_export(uzu_context* ctx, uint32_t capacity_f32, uint8_t* handle_out64) {
    if (!ctx || !handle_out64 || capacity_f32 == 0 || (capacity_f32 & 3u)) return fail(UZU_ERROR_INVALID_ARGUMENT, "tp_p2p_export: bad arguments");
    if (ctx->tp_size < 2 || ctx->tp_size > (uint32_t)TP_P2P_MAX_RANKS) return fail(UZU_ERROR_INVALID_ARGUMENT, "tp_p2p_export: needs a 2..8 rank context");
    if (ctx->tp_p2p) return fail(UZU_ERROR_INVALID_ARGUMENT, "tp_p2p_export: already exported");
    UZU_CUDA_TRY(cudaSetDevice(ctx->device));
    TpP2pState* st = new TpP2pState();
    const size_t bytes = tp_p2p_bytes(ctx->tp_size, capacity_f32);
    cudaError_t e = cudaMalloc(&st->local, bytes);
    if (e != cudaSuccess) { delete st; return fail(UZU_ERROR_OUT_OF_MEMORY, std::string("tp_p2p_export: ") + cudaGetErrorString(e)); }
    cudaMemset(st->local, 0, bytes);
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, st->local);
    if (e != cudaSuccess) { cudaFree(st->local); delete st; return fail(UZU_ERROR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    memcpy(handle_out64, &h, 64);
    st->capacity = capacity_f32;
    ctx->tp_p2p = st;
    return UZU_OK;
}

uzu_status uzu_tp_p2p_import(uzu_context* ctx, const uint8_t* handles_size_x_64) {
    if (!ctx || !handles_size_x_64 || !ctx->tp_p2p) return fail(UZU_ERROR_INVALID_ARGUMENT, "tp_p2p_import: export first");
    TpP2pState* st = static_cast<TpP2pState*>(ctx->tp_p2p);
    UZU_CUDA_TRY(cudaSetDevice(ctx->device));
    for (uint32_t r = 0; r < ctx->tp_size; ++r) {
        if (r == ctx->tp_rank) { st->peer[r] = st->local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles_size_x_64 + (size_t)r * 64, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&st->peer[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(UZU_ERROR_CUDA, "cudaIpcOpenMemHandle(rank " + std::to_string(r) + "): " + cudaGetErrorString(e));
    }
    UZU_CUDA_TRY(cudaDeviceSynchronize());
    st->ready = true;       // the host program must barrier all ranks after this call and before the first exchange
    return UZU_OK;
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
    if (ctx->tp_size > 1 && ctx->tp_p2p && static_cast<TpP2pState*>(ctx->tp_p2p)->ready && (count & 3u) == 0 &&
        count <= static_cast<TpP2pState*>(ctx->tp_p2p)->capacity && !((partial_f32 | out_bf16) & 15u)) {
        // one kernel: push to the peers over NVLink, wait, reduce in rank order, round (see tp_p2p_all_reduce_kernel)
        TpP2pState* st = static_cast<TpP2pState*>(ctx->tp_p2p);
        TpP2pView v{};
        const size_t data_bytes = (size_t)2 * ctx->tp_size * st->capacity * 4, flag_bytes = (size_t)2 * ctx->tp_size * TP_P2P_CTAS * 4;
        for (uint32_t r = 0; r < ctx->tp_size; ++r) {
            v.data[r] = reinterpret_cast<float*>(st->peer[r]);
            v.flags[r] = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(st->peer[r]) + data_bytes);
        }
        v.epoch = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(st->local) + data_bytes + flag_bytes);
        v.rank = ctx->tp_rank; v.size = ctx->tp_size; v.capacity = st->capacity;
        launch(cmd, "tp_p2p_all_reduce_kernel", tp_p2p_all_reduce_kernel, dim3(TP_P2P_CTAS), dim3(256), 0, v, (const float*)partial_f32,
               (__nv_bfloat16*)out_bf16, count);
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
