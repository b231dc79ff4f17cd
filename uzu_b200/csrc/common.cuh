// Shared device/host helpers for libuzu_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/uzu_b200.h"

namespace uzu {

// ---- error plumbing -------------------------------------------------------------------------
void set_last_error(const std::string& msg);
uzu_status fail(uzu_status st, const std::string& msg);

#define UZU_CUDA_TRY(expr)                                                                         \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return ::uzu::fail(UZU_ERROR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// ---- runtime objects ------------------------------------------------------------------------
struct Context;

}  // namespace uzu

struct uzu_context {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    size_t peak_bytes = 0, live_bytes = 0;
    // split-K workspace shared by the matmul kernels (stream-ordered, so one copy is enough)
    float* splitk_ws = nullptr;
    unsigned int* splitk_counters = nullptr;
    size_t splitk_ws_bytes = 0;
    size_t splitk_counter_count = 0;
    // sampling workspace
    unsigned long long* sampling_ws = nullptr;
    // split-KV attention workspace (partials + tickets), stream-ordered
    float* attn_ws = nullptr;
    size_t attn_ws_bytes = 0;
    unsigned int* attn_counters = nullptr;
    bool vmm_supported = false;
    size_t vmm_granularity = 0;
    // tensor parallelism (tp.cu): NCCL communicator of this process's rank, created by uzu_context_tp_init
    void* nccl_comm = nullptr;
    uint32_t tp_rank = 0, tp_size = 1;
    void* tp_p2p = nullptr;   // peer-memory exchange state (tp.cu, opt-in)
};

struct uzu_command_buffer {
    uzu_context* ctx = nullptr;
    std::string name;
    enum State { Initial, Encoding, Executable, Pending, Completed } state = Initial;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    uzu_status sticky = UZU_OK;
    std::string sticky_msg;
    uint64_t launches = 0;
    // Programmatic dependent launch for kernels that opt in (set by the engine)
    bool use_pdl = false;

    void record_error(uzu_status st, const std::string& msg) {
        if (sticky == UZU_OK) {
            sticky = st;
            sticky_msg = msg;
        }
    }
};

namespace uzu {

// Checks the launch and counts it. Must follow every <<<>>> in an encode function.
inline void after_launch(uzu_command_buffer* cmd, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) cmd->record_error(UZU_ERROR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    cmd->launches++;
}

#ifdef __CUDACC__
// Launch through cudaLaunchKernelEx so the command buffer can opt kernels into programmatic dependent launch.
template <typename... KArgs, typename... Args>
inline void launch(uzu_command_buffer* cmd, const char* what, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = cmd->ctx->stream;
    cudaLaunchAttribute attr[1];
    if (cmd->use_pdl) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
    if (e != cudaSuccess) cmd->record_error(UZU_ERROR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    cmd->launches++;
}
#endif

// tcgen05 prefill GEMM (prefill_gemm.cu), dispatched from encode_matmul (matmul.cu)
bool prefill_gemm_applicable(const uzu_matmul_args& a);
void encode_prefill_gemm(uzu_command_buffer* cmd, const uzu_matmul_args& a);

// tensor-core prefill attention (attention_prefill.cu, opt-in), tried first by uzu_attention_single_pass_encode
bool encode_attention_prefill(uzu_command_buffer* cmd, const uzu_attention_args& a);

// multi-token DeltaNet recurrence in one launch (deltanet_prefill.cu, opt-in), tried by the engine's batched hybrid prefill
bool encode_delta_net_prefill(uzu_command_buffer* cmd, const uzu_delta_net_fused_update_args& f, uint32_t rows, uint32_t in_stride, uint32_t out_stride);

// The launch stream belongs to the context's device: make that device current on the calling thread (another context, or the
// host program, may have switched it since start_encoding).
inline void make_current(const uzu_context* ctx) {
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess || d != ctx->device) cudaSetDevice(ctx->device);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (function, device), not of the process: opt in once per device.
// `done` is a per-call-site (per template instance) bit mask of device ordinals; setting the attribute twice is harmless, so a
// race between two threads only repeats the call.
template <typename K>
inline void opt_in_dynamic_smem(uzu_command_buffer* cmd, K kernel, int bytes, std::atomic<uint64_t>& done) {
    const uint64_t bit = 1ull << (cmd->ctx->device & 63);
    if (done.load(std::memory_order_acquire) & bit) return;
    make_current(cmd->ctx);
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        cmd->record_error(UZU_ERROR_CUDA, std::string("cudaFuncSetAttribute(MaxDynamicSharedMemorySize): ") + cudaGetErrorString(e));
        return;
    }
    done.fetch_or(bit, std::memory_order_release);
}

inline bool encodable(uzu_command_buffer* cmd, const char* what) {
    if (!cmd) return false;
    if (cmd->state != uzu_command_buffer::Encoding) {
        cmd->record_error(UZU_ERROR_INVALID_ARGUMENT, std::string(what) + ": command buffer is not in the Encoding state");
        return false;
    }
    make_current(cmd->ctx);
    return true;
}

// ---- device helpers -------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ float bf16_bits_to_f32(uint16_t h) { return __uint_as_float(((uint32_t)h) << 16); }
__device__ __forceinline__ float bf2f(__nv_bfloat16 h) { return __bfloat162float(h); }
// round-to-nearest-even, identical to half::bf16::from_f32
__device__ __forceinline__ __nv_bfloat16 f2bf(float f) { return __float2bfloat16_rn(f); }
__device__ __forceinline__ float round_bf16(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum for blockDim.x <= 1024; `red` is >= 32 floats of shared memory. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (lane < nw) ? red[lane] : 0.0f;
    t = warp_sum(t);
    return t;
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may start
// while its predecessor is still running; it must not touch the predecessor's outputs (or write anything the predecessor
// reads) before pdl_wait(). Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// activation math: backends/common/gpu_types/activation_type.rs:16-65 (f32 in, f32 out)
__device__ __forceinline__ float act_f32(uint32_t act, float x) {
    switch (act) {
        case UZU_ACT_SILU: return x / (1.0f + expf(-x));
        case UZU_ACT_GELU_APPROX: {
            float t = 0.7978846f * (x + 0.044715f * x * x * x);
            return 0.5f * x * (1.0f + tanhf(t));
        }
        case UZU_ACT_GELU_EXACT: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
        case UZU_ACT_SOFTPLUS: return x > 20.0f ? x : logf(1.0f + expf(x));
        default: return x;
    }
}

// the same function for translation units compiled WITH fma contraction (matmul.cu): explicit roundings so that fused
// prologues / epilogues match the standalone kernels (built with -fmad=false) bit for bit
__device__ __forceinline__ float act_f32_nofma(uint32_t act, float x) {
    switch (act) {
        case UZU_ACT_SILU: return __fdiv_rn(x, __fadd_rn(1.0f, expf(-x)));
        case UZU_ACT_GELU_APPROX: {
            const float x3 = __fmul_rn(__fmul_rn(__fmul_rn(0.044715f, x), x), x);
            const float t = __fmul_rn(0.7978846f, __fadd_rn(x, x3));
            return __fmul_rn(__fmul_rn(0.5f, x), __fadd_rn(1.0f, tanhf(t)));
        }
        case UZU_ACT_GELU_EXACT: return __fmul_rn(__fmul_rn(0.5f, x), __fadd_rn(1.0f, erff(__fmul_rn(x, 0.70710678118654752440f))));
        case UZU_ACT_SOFTPLUS: return x > 20.0f ? x : logf(__fadd_rn(1.0f, expf(x)));
        default: return x;
    }
}

#endif  // __CUDACC__

}  // namespace uzu
