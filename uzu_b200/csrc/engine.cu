// Host engine: C++ mirror of the reference's backend-agnostic Rust host code, driving the C-ABI kernels of
// this library in the reference's op order. Mirrors (paths relative to crates/backend-uzu/src/):
//   Engine::load_language_model          engine/language_model/mod.rs:58-116 (config.json + model.safetensors)
//   ParameterLoader / leaf.read_allocation parameters/loader.rs:64-250 (pread into DenseBuffer::cpu_ptr)
//   WeightMatrix::load / matmul_b         encodable_block/weight_matrix.rs:101-215
//   Decoder::encode                       encodable_block/decoder.rs:138-203
//   Transformer::encode                   encodable_block/transformer.rs:226-329
//   TransformerLayer::encode              encodable_block/transformer_layer.rs:194-238
//   Attention::attend                     encodable_block/mixer/attention/mode.rs:45-144
//   AttentionState                        encodable_block/mixer/attention/state.rs:62-237
//   DeltaNet::encode (m == 1 branch)      encodable_block/mixer/delta_net.rs:473-535
//   DenseMlp::encode                      encodable_block/mlp/dense.rs:32-48
//   Embedding lookup / readout            encodable_block/embedding.rs:345-456
//   Sampling::encode                      encodable_block/sampling/mod.rs:83-195
//   LanguageModelStream::{new,generate}   engine/language_model/stream/stream.rs:131-782
// Differences that are B200 design, not semantics: scratch is a fixed arena (so a decode step can be a
// CUDA graph), RoPE tables are the reference's host f32 table evaluated once for every position and kept on
// the device, per-step scalars (prefix length, seeds) live in a device-side DecodeState so a captured
// step replays unchanged.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <string>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <vector>

#include "common.cuh"
#include "decode_mega.cuh"

namespace uzu {

// =====================================================================================================
// minimal JSON
// =====================================================================================================
struct Json {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;

    const Json* get(const std::string& k) const {
        for (auto& kv : obj)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
    const Json& at(const std::string& k) const {
        const Json* j = get(k);
        if (!j) throw std::runtime_error("config: missing field '" + k + "'");
        return *j;
    }
    bool is_null() const { return type == Null; }
    std::string type_tag() const { return at("type").str; }
    double number() const {
        if (type != Num) throw std::runtime_error("config: expected a number");
        return num;
    }
    uint32_t u32() const { return (uint32_t)number(); }
    bool boolean() const {
        if (type != Bool) throw std::runtime_error("config: expected a bool");
        return b;
    }
};

struct JsonParser {
    const char* p;
    const char* end;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    Json parse() {
        ws();
        if (p >= end) throw std::runtime_error("json: unexpected end");
        Json j;
        if (*p == '{') {
            j.type = Json::Obj;
            ++p; ws();
            if (*p == '}') { ++p; return j; }
            while (true) {
                ws();
                Json k = parse_string();
                ws();
                if (*p != ':') throw std::runtime_error("json: expected ':'");
                ++p;
                j.obj.emplace_back(k.str, parse());
                ws();
                if (*p == ',') { ++p; continue; }
                if (*p == '}') { ++p; break; }
                throw std::runtime_error("json: expected ',' or '}'");
            }
        } else if (*p == '[') {
            j.type = Json::Arr;
            ++p; ws();
            if (*p == ']') { ++p; return j; }
            while (true) {
                j.arr.push_back(parse());
                ws();
                if (*p == ',') { ++p; continue; }
                if (*p == ']') { ++p; break; }
                throw std::runtime_error("json: expected ',' or ']'");
            }
        } else if (*p == '"') {
            j = parse_string();
        } else if (!strncmp(p, "true", 4)) { j.type = Json::Bool; j.b = true; p += 4; }
        else if (!strncmp(p, "false", 5)) { j.type = Json::Bool; j.b = false; p += 5; }
        else if (!strncmp(p, "null", 4)) { j.type = Json::Null; p += 4; }
        else {
            char* e = nullptr;
            j.type = Json::Num;
            j.num = strtod(p, &e);
            if (e == p) throw std::runtime_error("json: bad token");
            p = e;
        }
        return j;
    }
    Json parse_string() {
        if (*p != '"') throw std::runtime_error("json: expected string");
        ++p;
        Json j;
        j.type = Json::Str;
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                ++p;
                switch (*p) {
                    case 'n': j.str += '\n'; break;
                    case 't': j.str += '\t'; break;
                    case 'u': j.str += '?'; p += 4; break;
                    default: j.str += *p;
                }
                ++p;
            } else {
                j.str += *p++;
            }
        }
        ++p;
        return j;
    }
};

static Json parse_json(const std::string& text) {
    JsonParser jp{text.data(), text.data() + text.size()};
    return jp.parse();
}

static std::string read_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::string s;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    fclose(f);
    return s;
}

// =====================================================================================================
// parameters (safetensors)
// =====================================================================================================
struct TensorInfo {
    std::string dtype;
    std::vector<uint64_t> shape;
    uint64_t begin = 0, end = 0;
};

struct ParameterLoader {
    int fd = -1;
    uint64_t data_offset = 0;
    std::map<std::string, TensorInfo> tensors;
    std::map<std::string, std::string> metadata;
    std::set<std::string> validated;
    uint64_t file_size = 0;

    static uint64_t dtype_size(const std::string& dt) {   // parameters/safetensors_metadata.rs:93-108
        if (dt == "F32" || dt == "I32" || dt == "U32") return 4;
        if (dt == "F16" || dt == "BF16") return 2;
        if (dt == "I8" || dt == "U8") return 1;
        if (dt == "I64" || dt == "U64") return 8;
        return 0;
    }

    explicit ParameterLoader(const std::string& path) {
        fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) throw std::runtime_error("cannot open " + path);
        struct stat st{};
        if (fstat(fd, &st) != 0) throw std::runtime_error("cannot stat " + path);
        file_size = (uint64_t)st.st_size;
        uint64_t hlen = 0;
        if (pread(fd, &hlen, 8, 0) != 8) throw std::runtime_error("safetensors: short header");
        if (hlen > (100ull << 20) || 8 + hlen > file_size) throw std::runtime_error("safetensors: implausible header length");
        std::string header(hlen, '\0');
        if ((uint64_t)pread(fd, &header[0], hlen, 8) != hlen) throw std::runtime_error("safetensors: short header");
        data_offset = 8 + hlen;
        Json j = parse_json(header);
        for (auto& kv : j.obj) {
            if (kv.first == "__metadata__") {
                for (auto& m : kv.second.obj) metadata[m.first] = m.second.str;
                continue;
            }
            TensorInfo t;
            t.dtype = kv.second.at("dtype").str;
            for (auto& d : kv.second.at("shape").arr) t.shape.push_back((uint64_t)d.num);
            t.begin = (uint64_t)kv.second.at("data_offsets").arr[0].num;
            t.end = (uint64_t)kv.second.at("data_offsets").arr[1].num;
            tensors[kv.first] = t;
        }
    }
    ~ParameterLoader() { if (fd >= 0) close(fd); }

    // leaf(name).validate(shape, dtype) (parameters/loader.rs:140-160)
    const TensorInfo& validate(const std::string& name, const std::vector<uint64_t>& shape, const std::string& dtype) {
        auto it = tensors.find(name);
        if (it == tensors.end()) throw std::runtime_error("missing tensor '" + name + "'");
        if (it->second.shape != shape) {
            std::ostringstream os;
            os << "tensor '" << name << "' has shape [";
            for (auto d : it->second.shape) os << d << ",";
            os << "], expected [";
            for (auto d : shape) os << d << ",";
            os << "]";
            throw std::runtime_error(os.str());
        }
        if (it->second.dtype != dtype) throw std::runtime_error("tensor '" + name + "' has dtype " + it->second.dtype + ", expected " + dtype);
        // the buffer is sized from the offsets while kernels index by shape: a truncated / corrupt file must be a load error
        uint64_t numel = 1;
        for (auto d : shape) numel *= d;
        const TensorInfo& t = it->second;
        if (t.end < t.begin || data_offset + t.end > file_size || t.end - t.begin != numel * dtype_size(dtype))
            throw std::runtime_error("tensor '" + name + "': data_offsets do not match shape x dtype or exceed the file");
        validated.insert(name);
        return it->second;
    }
    // read_allocation: pread straight into the buffer's CPU pointer (loader.rs:162-179)
    void read_into(const TensorInfo& t, void* dst) {
        uint64_t done = 0, total = t.end - t.begin;
        while (done < total) {
            ssize_t n = pread(fd, (char*)dst + done, std::min<uint64_t>(total - done, 1u << 30), data_offset + t.begin + done);
            if (n <= 0) throw std::runtime_error("safetensors: short read");
            done += (uint64_t)n;
        }
    }
    void assert_all_validated() {  // loader.rs:230-250
        for (auto& kv : tensors)
            if (!validated.count(kv.first)) throw std::runtime_error("tensor '" + kv.first + "' was not consumed by the model");
    }
};

// =====================================================================================================
// model
// =====================================================================================================
struct Buf {
    uzu_buffer* b = nullptr;
    uint64_t ptr() const { return b ? uzu_buffer_gpu_ptr(b) : 0; }
};

struct WeightMatrix {
    Buf values, scales, zero_points, biases;
    uint32_t prologue = UZU_B_FULL_PRECISION, mode = UZU_QMODE_U4, group_size = 0, bits = 16;
    uint32_t rows = 0, cols = 0;
    uint64_t bytes = 0;  // algorithmic bytes streamed by one full pass over the matrix
};

struct Linear {
    WeightMatrix w;
    uint32_t in_dim = 0, out_dim = 0;
    // RHTLinearWrapper (linear/rht_wrapper.rs): i32 +-1 signs of the 32-wide input / output randomized Hadamard; scratch = the engine's
    // buffer for the transformed activation (the reference transforms its consumed input allocation in place, rht_wrapper.rs:294)
    Buf in_signs, out_signs;
    uint64_t rht_scratch = 0;
    bool rht() const { return in_signs.b != nullptr; }
};

struct NormCfg {
    float epsilon = 1e-5f, scale_offset = 0.0f;
    bool full_layer = false, subtract_mean = false, has_scale = true, has_biases = false;
};
struct Norm {
    NormCfg cfg;
    Buf scales;
    uint32_t n = 0;
    bool present = false;
};

struct RopeCfg {
    int kind = 0;  // 0 unscaled, 1 linear, 2 llama
    float base = 10000.0f, scaling_factor = 1.0f, low = 1.0f, high = 1.0f;
    uint32_t head_dim = 0, original_context_length = 0, max_sequence_length = 0;
    bool operator==(const RopeCfg& o) const {
        return kind == o.kind && base == o.base && scaling_factor == o.scaling_factor && low == o.low && high == o.high &&
               head_dim == o.head_dim && original_context_length == o.original_context_length;
    }
};

struct AttentionLayer {
    Linear qkv, out, gate;
    bool has_gate = false;
    Norm qnorm, knorm;
    uint32_t num_heads = 0, num_groups = 0, head_dim = 0;
    bool is_causal = true;
    bool has_scale = false;
    float scale = 0.0f;
    int rope_index = -1;
};

struct DeltaNetLayer {
    Linear in_proj, out_proj;
    Buf conv_weight, conv_bias, a_log, dt_bias, norm_weight;
    bool conv_has_bias = false;
    uint32_t num_heads = 0, num_groups = 0, head_dim = 0, value_head_dim = 0, kernel_size = 0;
    uint32_t key_dim = 0, value_dim = 0, conv_dim = 0, total_proj_dim = 0;
    float norm_epsilon = 1e-6f;
};

struct Layer {
    bool is_attention = true;
    Norm pre_mixer, pre_mlp;
    AttentionLayer attn;
    DeltaNetLayer dn;
    Linear up, down;
    uint32_t hidden_dim = 0;
    uint32_t act = UZU_ACT_SILU;
};

struct LayerState {
    // attention
    uzu_sparse_buffer* k_sparse = nullptr;
    uzu_sparse_buffer* v_sparse = nullptr;
    Buf k_dense, v_dense;
    uint64_t keys = 0, values = 0;
    uint32_t length = 0;
    uint32_t mapped_pages = 0;
    size_t row_bytes = 0;
    // delta net
    Buf conv_state, ssm_state, conv_snapshot, ssm_snapshot;
    size_t conv_bytes = 0, ssm_bytes = 0;
};

struct DecodeState {      // device-resident per-stream scalars (one u32-aligned block)
    uint32_t position;    // prefix length == absolute position of the token being fed
    uint32_t step;        // decode steps issued so far
    uint64_t base_seed;
};

constexpr uint32_t MAX_ROWS = 1024;              // ATTENTION_SUFFIX_CAPACITY (mixer/attention/state.rs:14)
constexpr uint32_t TOKEN_RING = 256;
constexpr uint32_t MAX_TRIE = 16;                // nodes per speculation pass: the stream's full_batch_size (stream.rs:550) == logits_rows

// RoPE rows of a speculation pass: node i sits at position context + height(i) (transformer.rs:248), so its cos/sin row is a gather
__global__ void trie_rope_gather_kernel(const uzu_trie_node* nodes, uint32_t context_length, uint32_t dim, const float* cos_table,
                                        const float* sin_table, float* cos_out, float* sin_out) {
    const size_t row = (size_t)context_length + nodes[blockIdx.x].height;
    for (uint32_t d = threadIdx.x; d < dim; d += blockDim.x) {
        cos_out[(size_t)blockIdx.x * dim + d] = cos_table[row * dim + d];
        sin_out[(size_t)blockIdx.x * dim + d] = sin_table[row * dim + d];
    }
}

__global__ void decode_step_begin_kernel(const DecodeState* st, unsigned long long* seeds) {
    // PRng::derive(position) (encodable_block/sampling/prng.rs:12-23)
    unsigned long long h = st->base_seed + (unsigned long long)st->position;
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    seeds[0] = h;
}

__global__ void decode_step_end_kernel(DecodeState* st, const uint32_t* sampled, uint32_t* token_ids, volatile uint32_t* host_ring,
                                       uint32_t* dev_out, uint32_t dev_out_base_step) {
    const uint32_t tok = sampled[0];
    token_ids[0] = tok;                        // device-side chaining of the next input (stream.rs:611-615)
    host_ring[st->step % TOKEN_RING] = tok;    // pinned host ring read by next()/flush()
    if (dev_out) dev_out[st->step - dev_out_base_step] = tok;
    st->position += 1;
    st->step += 1;
}

}  // namespace uzu

using namespace uzu;

struct uzu_engine {
    uzu_context* ctx = nullptr;
    uzu_engine_options opts{};
    // model
    uint32_t model_dim = 0, hidden_dim = 0, vocab = 0;
    bool tied = false;
    float input_scale = 1.0f;
    bool has_logit_scale = false, has_logit_soft_cap = false;
    float logit_scale = 1.0f, logit_soft_cap = 0.0f;
    Linear in_emb, out_emb;
    std::vector<Layer> layers;
    Norm out_norm;
    std::vector<RopeCfg> ropes;
    std::vector<Buf> rope_cos, rope_sin;  // [positions, rope_dim] f32
    uint32_t rope_positions = 0;
    std::vector<uzu_buffer*> owned;
    // state
    std::vector<LayerState> state;
    uint32_t context_length = 0, snapshot_context = 0;
    uint32_t max_context = 0;
    // scratch arena
    Buf shortcut2;   // fused decode path: the residual ping-pongs between `shortcut` and `shortcut2`
    bool fused_ok = false;
    Buf token_ids, hidden_a, hidden_b, shortcut, qkv, queries, attn_out, gate, fused_up, gated, mixer_out, in_proj, delta_out, normed_out,
        logits, sampled, seeds, tp_parts, tp_sums, tp_maxs;
    Buf host_ring;        // pinned u32[TOKEN_RING]
    Buf host_tokens;      // pinned u32[MAX_ROWS] upload staging
    Buf decode_state;     // device DecodeState
    Buf snapshot_token;   // next-input token at snapshot time
    uint32_t logits_rows = 16;
    // speculation pass (trie): device nodes, per-rope gathered cos/sin rows, pinned staging; pending = nodes of an unaccepted pass
    Buf rht_scratch;
    Buf trie_nodes, host_trie;
    std::vector<Buf> trie_cos, trie_sin;
    uint32_t trie_pending = 0;
    // tensor parallelism (config.json "tensor_parallel" block written by uzu_b200/tp.py; tp.cu holds the collectives)
    uint32_t tp_rank = 0, tp_size = 1, vocab_local = 0;
    bool tp_sharded = false;       // the checkpoint is a shard: row-parallel partials in f32 + exchange, vocab-parallel readout
    Buf tp_partial, logits_local, tp_gather;
    // multi-sequence batched decode (extension, BASELINE config 4 "batch=8"): independent sequence states sharing one weight pass per step
    struct Sequence { std::vector<LayerState> layers; uint32_t context_length = 0; };
    std::vector<Sequence> seqs;
    Buf batch_pos;                 // device u32[16]: per-sequence prefix lengths (the captured batched step reads positions from here)
    bool batch_pos_valid = false;
    cudaGraphExec_t batch_graph = nullptr;
    uint32_t batch_graph_bucket = 0, batch_graph_B = 0;
    uint64_t batch_graph_launches = 0;
    // streaming
    uzu_sampling_method sampling{};
    uint32_t steps_issued = 0, steps_returned = 0;
    cudaEvent_t step_events[2] = {nullptr, nullptr};
    // CUDA graph of one decode step
    cudaGraphExec_t graph_exec = nullptr;
    uint32_t graph_bucket = 0;
    bool graph_stochastic = false;
    uzu_sampling_method graph_sampling{};   // the method baked into the captured sampling launch (seed excluded: it lives in DecodeState)
    uint64_t graph_launches_per_step = 0;
    bool use_pdl = true;   // programmatic dependent launch between the kernels of a decode step (UZU_NO_PDL=1 disables)
    uint64_t launches = 0;
    uzu_model_info info{};
    // persistent whole-token decode kernel (decode_mega.cu): program + buffers, built at load when the model is covered
    struct Mega {
        bool ok = false;
        std::string why;                 // why the model is not covered (diagnostics)
        uzu::MegaConfig cfg{};
        std::vector<uzu::MkOp> ops;
        Buf ops_dev, streams_dev, barrier, error_flag, argmax_keys, attn_part, attn_tickets, dn_raw;
        uint32_t nstreams = 0;
        uint64_t stream_bytes = 0;
        uint64_t barrier_base = 0;       // arrivals the monotonic grid-barrier counter has seen before the next launch
        std::map<uint64_t, uint64_t> stream_of;   // weight values pointer -> decode-stream copy (also feeds the per-kernel path's TMA GEMV)
    } mega;
};

namespace uzu {

static void check(uzu_status st) {
    if (st != UZU_OK) throw std::runtime_error(uzu_last_error());
}

static Buf make_buf(uzu_engine* e, size_t bytes, uzu_buffer_kind kind) {
    Buf b;
    check(uzu_buffer_create(e->ctx, bytes, kind, &b.b));
    e->owned.push_back(b.b);
    return b;
}

static Buf load_tensor(uzu_engine* e, ParameterLoader& pl, const std::string& name, const std::vector<uint64_t>& shape, const std::string& dtype) {
    const TensorInfo& t = pl.validate(name, shape, dtype);
    Buf b = make_buf(e, t.end - t.begin, UZU_BUFFER_MANAGED);
    pl.read_into(t, uzu_buffer_cpu_ptr(b.b));
    check(uzu_buffer_make_resident(e->ctx, b.b));
    return b;
}

// WeightMatrix::load (weight_matrix.rs:101-162). `layout_required`: "output_input" for linears, "input_output" for embeddings.
static WeightMatrix load_weight_matrix(uzu_engine* e, ParameterLoader& pl, const std::string& prefix, const char* layout_required,
                                       uint32_t rows, uint32_t cols) {
    auto it = pl.metadata.find(prefix + ".spec");
    if (it == pl.metadata.end()) throw std::runtime_error("missing weight spec '" + prefix + ".spec' in safetensors metadata");
    Json spec = parse_json(it->second);
    if (spec.at("layout").str != layout_required) throw std::runtime_error(prefix + ": expected " + layout_required + " layout");
    WeightMatrix w;
    w.rows = rows;
    w.cols = cols;
    const std::string ty = spec.type_tag();
    if (ty == "FullPrecisionSpec") {
        w.values = load_tensor(e, pl, prefix + ".weights", {rows, cols}, "BF16");
        w.bytes = (uint64_t)rows * cols * 2;
        return w;
    }
    if (ty != "MLXSpec" && ty != "IntSpec") throw std::runtime_error(prefix + ": unsupported weight spec " + ty);
    const uint32_t bits = spec.at("bits").u32();
    w.group_size = spec.at("group_size").u32();
    if (bits != 4 && bits != 8) throw std::runtime_error(prefix + ": unsupported bits");
    if (w.group_size == 0) throw std::runtime_error(prefix + ": group size must be non-zero");
    w.bits = bits;
    w.mode = bits == 4 ? UZU_QMODE_U4 : UZU_QMODE_U8;
    const uint32_t packing = bits == 4 ? 2 : 1;
    if (cols % packing) throw std::runtime_error(prefix + ": stored columns not divisible by the packing divisor");
    const uint32_t groups = (cols + w.group_size - 1) / w.group_size;
    w.values = load_tensor(e, pl, prefix + ".weights", {rows, cols / packing}, "U8");
    w.scales = load_tensor(e, pl, prefix + ".scales", {rows, groups}, "BF16");
    w.bytes = (uint64_t)rows * (cols / packing) + (uint64_t)rows * groups * 2;
    if (ty == "MLXSpec") {
        w.prologue = UZU_B_SCALE_BIAS_DEQUANT;
        w.biases = load_tensor(e, pl, prefix + ".biases", {rows, groups}, "BF16");
        w.bytes += (uint64_t)rows * groups * 2;
    } else if (spec.at("is_symmetric").boolean()) {
        w.prologue = UZU_B_SCALE_SYMMETRIC_DEQUANT;
    } else {
        w.prologue = UZU_B_SCALE_ZERO_POINT_DEQUANT;
        const uint32_t zcols = (groups + packing - 1) / packing;
        w.zero_points = load_tensor(e, pl, prefix + ".zero_points", {rows, zcols}, "U8");
        w.bytes += (uint64_t)rows * zcols;
    }
    return w;
}

static Linear load_linear(uzu_engine* e, ParameterLoader& pl, const std::string& prefix, uint32_t in_dim, uint32_t out_dim) {
    Linear l;
    l.in_dim = in_dim;
    l.out_dim = out_dim;
    const std::string wp = prefix + ".weights";
    auto it = pl.metadata.find(wp + ".spec");
    if (it != pl.metadata.end()) {
        Json spec = parse_json(it->second);
        if (spec.type_tag() == "HybridSpec") {
            // Linear::new_mixed_precision (linear/mod.rs:128-143) -> RHTLinearWrapper::load_inner_with_output_rht (rht_wrapper.rs:141-176)
            const Json* adapter = spec.get("adapter_spec");
            const Json* block = spec.get("incoherence_block_size");
            const Json* mode = spec.get("incoherence_processing_mode");
            if ((adapter && !adapter->is_null()) || !block || block->is_null() || block->u32() != 32 || !mode || mode->str != "input_output")
                throw std::runtime_error(wp + ": unsupported HybridSpec (only 32-wide input_output RHT without an adapter)");
            if (in_dim % 32 || out_dim % 32) throw std::runtime_error(wp + ": RHT dimensions must be multiples of 32");
            l.in_signs = load_tensor(e, pl, wp + ".incoherence_signs.input_signs", {in_dim}, "I32");
            l.out_signs = load_tensor(e, pl, wp + ".incoherence_signs.output_signs", {out_dim}, "I32");
            l.w = load_weight_matrix(e, pl, wp + ".quantized", "output_input", out_dim, in_dim);
            if (l.w.prologue == UZU_B_FULL_PRECISION) throw std::runtime_error(wp + ": fused output-hadamard factors require quantized weights");
            return l;
        }
    }
    l.w = load_weight_matrix(e, pl, wp, "output_input", out_dim, in_dim);
    return l;
}

static NormCfg parse_norm_cfg(const Json& j) {
    NormCfg c;
    c.epsilon = (float)j.at("epsilon").number();
    c.scale_offset = j.at("scale_offset").is_null() ? 0.0f : (float)j.at("scale_offset").number();
    c.full_layer = j.at("upcast_mode").str == "full_layer";
    c.subtract_mean = j.at("subtract_mean").boolean();
    c.has_scale = j.at("has_scale").boolean();
    c.has_biases = j.at("has_biases").boolean();
    if (c.has_biases) throw std::runtime_error("normalization biases are not supported by this engine mirror");
    return c;
}

static Norm load_norm(uzu_engine* e, ParameterLoader& pl, const std::string& prefix, const Json& cfg, uint32_t n) {
    Norm nm;
    nm.present = true;
    nm.cfg = parse_norm_cfg(cfg);
    nm.n = n;
    if (nm.cfg.has_scale) nm.scales = load_tensor(e, pl, prefix + ".scales", {n}, "F32");
    return nm;
}

static RopeCfg parse_rope(const Json& j) {
    RopeCfg r;
    const std::string ty = j.type_tag();
    r.base = (float)j.at("base").number();
    r.head_dim = j.at("head_dim").u32();
    r.max_sequence_length = j.at("max_sequence_length").u32();
    if (ty == "UnscaledRoPEConfig") r.kind = 0;
    else if (ty == "LinearScalingRoPEConfig") { r.kind = 1; r.scaling_factor = (float)j.at("scaling_factor").number(); }
    else if (ty == "LlamaRoPEConfig") {
        r.kind = 2;
        r.scaling_factor = (float)j.at("scaling_factor").number();
        r.original_context_length = j.at("original_context_length").u32();
        r.low = (float)j.at("low_frequency_factor").number();
        r.high = (float)j.at("high_frequency_factor").number();
    } else throw std::runtime_error("unsupported RoPE config " + ty + " (YaRN / LongRoPE are outside the BASELINE configs)");
    return r;
}

// PrecalculatedRoPE::precalculate (mixer/attention/rope.rs:13-114), evaluated for positions [0, count)
static void rope_tables(const RopeCfg& c, uint32_t count, float* cosines, float* sines) {
    const uint32_t head_dim = c.head_dim, half = head_dim / 2;
    for (uint32_t pair = 0; pair < half; ++pair) {
        const uint32_t channel = pair * 2;
        float inv = 1.0f / powf(c.base, (float)channel / (float)head_dim);
        if (c.kind == 1) inv = inv / c.scaling_factor;
        else if (c.kind == 2) {
            const float low_wl = (float)c.original_context_length / c.low;
            const float high_wl = (float)c.original_context_length / c.high;
            const float wavelength = 2.0f * 3.14159265358979323846f / inv;
            const float scaled = inv / c.scaling_factor;
            if (wavelength < high_wl) {
            } else if (wavelength > low_wl) inv = scaled;
            else {
                float smooth = (float)c.original_context_length / wavelength - c.low;
                smooth = smooth / (c.high - c.low);
                inv = smooth * inv + (1.0f - smooth) * scaled;
            }
        }
        for (uint32_t t = 0; t < count; ++t) {
            const float em = (float)t * inv;
            const float s = sinf(em) * 1.0f, co = cosf(em) * 1.0f;
            const size_t po = (size_t)t * head_dim + pair;
            sines[po] = s; sines[po + half] = s;
            cosines[po] = co; cosines[po + half] = co;
        }
    }
}

static void load_model(uzu_engine* e, const std::string& dir) {
    Json cfg = parse_json(read_file(dir + "/config.json"));
    if (cfg.type_tag() != "LanguageModelConfig") throw std::runtime_error("config.json: not a LanguageModelConfig");
    const Json& dec = cfg.at("decoder_config");
    const Json& tr = dec.at("transformer_config");
    e->model_dim = tr.at("model_dim").u32();
    e->hidden_dim = tr.at("hidden_dim").u32();
    e->vocab = dec.at("vocab_size").u32();
    if (!dec.at("ple_model_config").is_null() || !dec.at("embedding_norm_config").is_null())
        throw std::runtime_error("PLE / embedding norm are outside this backend's scope (SURVEY 2.1)");
    ParameterLoader pl(dir + "/model.safetensors");
    const uint32_t H = e->model_dim, V = e->vocab;
    // tensor-parallel shard (uzu_b200/tp.py): narrower heads / hidden_dim come from the shard's own config; the block says which rank
    // this is and how many vocabulary rows the local readout holds. The context must carry the matching communicator.
    const uint32_t want_tp = e->opts.tp_size ? e->opts.tp_size : 1;
    uint32_t vocab_out = V;
    if (const Json* tpj = cfg.get("tensor_parallel")) {
        if (!tpj->is_null()) {
            e->tp_sharded = true;
            e->tp_rank = tpj->at("rank").u32();
            e->tp_size = tpj->at("size").u32();
            e->vocab_local = tpj->at("vocab_size_local").u32();
            if (e->tp_size == 0 || e->tp_rank >= e->tp_size || (uint64_t)e->vocab_local * e->tp_size != V)
                throw std::runtime_error("config.json: inconsistent tensor_parallel block");
            if (e->ctx->tp_size != e->tp_size || e->ctx->tp_rank != e->tp_rank)
                throw std::runtime_error("tensor-parallel shard " + std::to_string(e->tp_rank) + "/" + std::to_string(e->tp_size) +
                                         " needs a context initialised with the same rank / size (uzu_context_tp_init)");
            vocab_out = e->vocab_local;
        }
    }
    if (want_tp != e->tp_size || (e->opts.tp_size && e->opts.tp_rank != e->tp_rank))
        throw std::runtime_error("uzu_engine_options tp_rank/tp_size do not match the checkpoint (shard it with uzu_b200/tp.py)");

    const Json& emb = dec.at("embedding_config");
    e->tied = emb.type_tag() == "TiedEmbeddingConfig";
    if (!emb.at("input_scale").is_null()) e->input_scale = (float)emb.at("input_scale").number();
    if (!emb.at("logit_scale").is_null()) { e->has_logit_scale = true; e->logit_scale = (float)emb.at("logit_scale").number(); }
    if (!emb.at("logit_soft_cap").is_null()) { e->has_logit_soft_cap = true; e->logit_soft_cap = (float)emb.at("logit_soft_cap").number(); }
    if (e->tied && e->tp_sharded) throw std::runtime_error("a tensor-parallel shard carries untied embeddings (uzu_b200/tp.py)");
    if (e->tied) {
        e->in_emb.in_dim = H; e->in_emb.out_dim = V;
        e->in_emb.w = load_weight_matrix(e, pl, "decoder.embedding.embedding", "input_output", V, H);
        e->out_emb = e->in_emb;
    } else {
        e->in_emb.in_dim = H; e->in_emb.out_dim = V;
        e->in_emb.w = load_weight_matrix(e, pl, "decoder.embedding.input_embedding", "input_output", V, H);
        e->out_emb.in_dim = H; e->out_emb.out_dim = vocab_out;
        e->out_emb.w = load_weight_matrix(e, pl, "decoder.embedding.output_embedding", "input_output", vocab_out, H);
    }

    uint64_t wbytes = e->out_emb.w.bytes, kvb = 0, stb = 0;
    uint32_t n_attn = 0, n_dn = 0;
    const auto& lcs = tr.at("layer_configs").arr;
    e->layers.resize(lcs.size());
    for (size_t i = 0; i < lcs.size(); ++i) {
        const Json& lc = lcs[i];
        Layer& L = e->layers[i];
        const std::string p = "decoder.transformer.layers." + std::to_string(i);
        if (lc.at("pre_mixer_norm_config").is_null()) throw std::runtime_error("layers without pre_mixer_norm are not supported");
        if (!lc.at("post_mixer_norm_config").is_null() || !lc.at("post_mlp_norm_config").is_null() || !lc.at("ple_config").is_null() ||
            lc.at("has_post_layer_scalar").boolean() || !lc.at("kv_source_layer_index").is_null())
            throw std::runtime_error("post norms / PLE / post-layer scalar / KV sharing are outside this backend's scope");
        L.pre_mixer = load_norm(e, pl, p + ".pre_mixer_norm", lc.at("pre_mixer_norm_config"), H);
        L.pre_mlp = load_norm(e, pl, p + ".pre_mlp_norm", lc.at("pre_mlp_norm_config"), H);
        const Json& mc = lc.at("mixer_config");
        const std::string mty = mc.type_tag();
        if (mty == "AttentionConfig") {
            L.is_attention = true;
            AttentionLayer& A = L.attn;
            A.num_heads = mc.at("num_heads").u32();
            A.num_groups = mc.at("num_groups").u32();
            A.head_dim = mc.at("head_dim").u32();
            A.is_causal = mc.at("is_causal").boolean();
            if (!mc.at("scale").is_null()) { A.has_scale = true; A.scale = (float)mc.at("scale").number(); }
            if (!mc.at("sliding_window_size").is_null() || mc.at("has_sinks").boolean() || mc.at("has_qkv_biases").boolean() ||
                mc.at("has_out_biases").boolean() || mc.at("normalize_values").boolean() || mc.at("is_kv_sharing").boolean() ||
                !mc.at("logit_soft_cap").is_null())
                throw std::runtime_error("attention variant (sliding window / sinks / biases / value norm / KV sharing) not supported by the engine mirror");
            const uint32_t qd = A.num_heads * A.head_dim, kvd = A.num_groups * A.head_dim;
            A.qkv = load_linear(e, pl, p + ".mixer.qkv_projection", H, qd + 2 * kvd);
            A.out = load_linear(e, pl, p + ".mixer.out_projection", qd, H);
            if (!mc.at("gate_projection_config").is_null()) {
                A.has_gate = true;
                A.gate = load_linear(e, pl, p + ".mixer.gate_projection", H, qd);
                wbytes += A.gate.w.bytes;
            }
            if (!mc.at("query_norm_config").is_null()) A.qnorm = load_norm(e, pl, p + ".mixer.query_norm", mc.at("query_norm_config"), A.head_dim);
            if (!mc.at("key_norm_config").is_null()) A.knorm = load_norm(e, pl, p + ".mixer.key_norm", mc.at("key_norm_config"), A.head_dim);
            if (!lc.at("rope_config").is_null()) {
                RopeCfg rc = parse_rope(lc.at("rope_config"));
                int idx = -1;
                for (size_t r = 0; r < e->ropes.size(); ++r)
                    if (e->ropes[r] == rc) idx = (int)r;
                if (idx < 0) { e->ropes.push_back(rc); idx = (int)e->ropes.size() - 1; }
                A.rope_index = idx;
            }
            wbytes += A.qkv.w.bytes + A.out.w.bytes;
            kvb += 2ull * kvd * 2;
            n_attn++;
        } else if (mty == "DeltaNetConfig") {
            if (e->tp_sharded) throw std::runtime_error("tensor parallelism covers attention mixers only");
            L.is_attention = false;
            DeltaNetLayer& D = L.dn;
            D.num_heads = mc.at("num_heads").u32();
            D.num_groups = mc.at("num_groups").u32();
            D.head_dim = mc.at("head_dim").u32();
            D.value_head_dim = mc.at("value_head_dim").u32();
            D.kernel_size = mc.at("kernel_size").u32();
            if (D.head_dim != 128 || D.value_head_dim != 128 || D.kernel_size < 2)
                throw std::runtime_error("DeltaNet: head_dim and value_head_dim must be 128, kernel_size >= 2 (delta_net.rs:168-186)");
            D.key_dim = D.num_groups * D.head_dim;
            D.value_dim = D.num_heads * D.value_head_dim;
            D.conv_dim = 2 * D.key_dim + D.value_dim;
            D.total_proj_dim = D.conv_dim + D.value_dim + 2 * D.num_heads;
            D.norm_epsilon = (float)mc.at("norm_config").at("epsilon").number();
            D.conv_has_bias = mc.at("conv_config").at("has_biases").boolean();
            D.in_proj = load_linear(e, pl, p + ".mixer.in_proj", H, D.total_proj_dim);
            D.conv_weight = load_tensor(e, pl, p + ".mixer.conv.weights", {D.conv_dim, D.kernel_size}, "F32");
            if (D.conv_has_bias) D.conv_bias = load_tensor(e, pl, p + ".mixer.conv.biases", {D.conv_dim}, "F32");
            D.a_log = load_tensor(e, pl, p + ".mixer.a_log", {D.num_heads}, "F32");
            D.dt_bias = load_tensor(e, pl, p + ".mixer.dt_bias", {D.num_heads}, "F32");
            D.norm_weight = load_tensor(e, pl, p + ".mixer.norm.scales", {D.value_head_dim}, "F32");
            D.out_proj = load_linear(e, pl, p + ".mixer.out_proj", D.value_dim, H);
            wbytes += D.in_proj.w.bytes + D.out_proj.w.bytes;
            stb += 2ull * D.num_heads * D.value_head_dim * D.head_dim * 4;
            n_dn++;
        } else {
            throw std::runtime_error("mixer " + mty + " is outside this backend's scope (SURVEY 2.1)");
        }
        const Json& mlp = lc.at("mlp_config");
        if (mlp.type_tag() != "DenseMLPConfig") throw std::runtime_error("only DenseMLPConfig is supported");
        if (mlp.at("has_up_biases").boolean() || mlp.at("has_down_biases").boolean() || !mlp.at("gate_clipping").is_null() ||
            !mlp.at("up_clipping").is_null())
            throw std::runtime_error("MLP biases / clipping are not supported by the engine mirror");
        const std::string act = mlp.at("activation").type_tag();
        if (act == "SiLU") L.act = UZU_ACT_SILU;
        else if (act == "GELU") L.act = mlp.at("activation").at("approximate").boolean() ? UZU_ACT_GELU_APPROX : UZU_ACT_GELU_EXACT;
        else throw std::runtime_error("Identity activation is not supported for kernel (mlp/gate_act_mul.rs:60-62)");
        L.hidden_dim = lc.at("hidden_dim").is_null() ? e->hidden_dim : lc.at("hidden_dim").u32();
        L.up = load_linear(e, pl, p + ".mlp.up_projection", H, 2 * L.hidden_dim);
        L.down = load_linear(e, pl, p + ".mlp.down_projection", L.hidden_dim, H);
        wbytes += L.up.w.bytes + L.down.w.bytes;
    }
    e->out_norm = load_norm(e, pl, "decoder.transformer.output_norm", tr.at("output_norm_config"), H);
    pl.assert_all_validated();

    e->info.model_dim = H;
    e->info.hidden_dim = e->hidden_dim;
    e->info.vocab_size = V;
    e->info.num_layers = (uint32_t)e->layers.size();
    e->info.num_attention_layers = n_attn;
    e->info.num_delta_net_layers = n_dn;
    e->info.weight_bytes_per_token = wbytes;
    e->info.kv_bytes_per_token_per_ctx = kvb;
    e->info.state_bytes_per_token = stb;
}

static void create_state_and_scratch(uzu_engine* e) {
    const uint32_t H = e->model_dim, V = e->vocab;
    uint32_t max_ctx = e->opts.max_context_length ? e->opts.max_context_length : 8192;
    for (auto& r : e->ropes) max_ctx = std::min(max_ctx, r.max_sequence_length);  // state.rs:80-85
    e->max_context = max_ctx;
    const uint32_t rows_total = max_ctx + MAX_ROWS;
    const bool sparse = (uzu_context_device_capabilities(e->ctx) & UZU_CAP_SPARSE_BUFFERS) != 0;
    e->state.resize(e->layers.size());
    uint32_t max_qkv = 0, max_qd = 0, max_f = 0, max_proj = 0, max_vd = 0, max_heads = 0, max_hd = 0;
    for (size_t i = 0; i < e->layers.size(); ++i) {
        Layer& L = e->layers[i];
        LayerState& S = e->state[i];
        max_f = std::max(max_f, L.hidden_dim);
        if (L.is_attention) {
            const AttentionLayer& A = L.attn;
            S.row_bytes = (size_t)A.num_groups * A.head_dim * 2;
            const size_t bytes = (size_t)rows_total * S.row_bytes;
            if (sparse) {
                check(uzu_sparse_buffer_create(e->ctx, bytes, &S.k_sparse));
                check(uzu_sparse_buffer_create(e->ctx, bytes, &S.v_sparse));
                S.keys = uzu_sparse_buffer_gpu_ptr(S.k_sparse);
                S.values = uzu_sparse_buffer_gpu_ptr(S.v_sparse);
            } else {
                S.k_dense = make_buf(e, bytes, UZU_BUFFER_DEVICE);
                S.v_dense = make_buf(e, bytes, UZU_BUFFER_DEVICE);
                S.keys = S.k_dense.ptr();
                S.values = S.v_dense.ptr();
            }
            max_qkv = std::max(max_qkv, (A.num_heads + 2 * A.num_groups) * A.head_dim);
            max_qd = std::max(max_qd, A.num_heads * A.head_dim);
            max_heads = std::max(max_heads, A.num_heads);
            max_hd = std::max(max_hd, A.head_dim);
        } else {
            const DeltaNetLayer& D = L.dn;
            S.conv_bytes = (size_t)D.conv_dim * (D.kernel_size - 1) * 4;
            S.ssm_bytes = (size_t)D.num_heads * D.value_head_dim * D.head_dim * 4;
            S.conv_state = make_buf(e, S.conv_bytes, UZU_BUFFER_DEVICE);
            S.ssm_state = make_buf(e, S.ssm_bytes, UZU_BUFFER_DEVICE);
            S.conv_snapshot = make_buf(e, S.conv_bytes, UZU_BUFFER_DEVICE);
            S.ssm_snapshot = make_buf(e, S.ssm_bytes, UZU_BUFFER_DEVICE);
            max_proj = std::max(max_proj, D.total_proj_dim);
            max_vd = std::max(max_vd, D.value_dim);
        }
    }
    auto dev = [&](size_t bytes) { return make_buf(e, std::max<size_t>(bytes, 256), UZU_BUFFER_DEVICE); };
    e->token_ids = dev(MAX_ROWS * 4);
    e->hidden_a = dev((size_t)MAX_ROWS * H * 2);
    e->hidden_b = dev((size_t)MAX_ROWS * H * 2);
    e->shortcut = dev((size_t)MAX_ROWS * H * 2);
    e->shortcut2 = dev((size_t)H * 2);
    e->mixer_out = dev((size_t)MAX_ROWS * H * 2);
    e->normed_out = dev((size_t)e->logits_rows * H * 2);
    e->qkv = dev((size_t)MAX_ROWS * max_qkv * 2);
    e->queries = dev((size_t)MAX_ROWS * max_qd * 2);
    e->attn_out = dev((size_t)MAX_ROWS * max_qd * 2);
    e->gate = dev((size_t)MAX_ROWS * max_qd * 2);
    e->fused_up = dev((size_t)MAX_ROWS * 2 * max_f * 2);
    e->gated = dev((size_t)MAX_ROWS * max_f * 2);
    e->in_proj = dev((size_t)MAX_ROWS * max_proj * 2);
    e->delta_out = dev((size_t)MAX_ROWS * max_vd * 2);
    e->logits = dev((size_t)e->logits_rows * V * 2);
    if (e->tp_sharded) {
        e->tp_partial = dev((size_t)MAX_ROWS * H * 4);
        e->logits_local = dev((size_t)e->logits_rows * e->vocab_local * 2);
        e->tp_gather = dev((size_t)e->logits_rows * V * 2);
    }
    e->sampled = dev(MAX_ROWS * 4);
    e->seeds = dev(MAX_ROWS * 8);
    // two-pass attention scratch (only used for suffix <= 8 and context > 1024, like the reference dispatch)
    e->tp_parts = dev((size_t)8 * max_heads * 32 * max_hd * 4);
    e->tp_sums = dev((size_t)8 * max_heads * 32 * 4);
    e->tp_maxs = dev((size_t)8 * max_heads * 32 * 4);
    e->host_ring = make_buf(e, TOKEN_RING * 4, UZU_BUFFER_PINNED_HOST);
    e->host_tokens = make_buf(e, MAX_ROWS * 4, UZU_BUFFER_PINNED_HOST);
    e->decode_state = dev(sizeof(DecodeState));
    e->batch_pos = dev(16 * 4);
    e->snapshot_token = dev(4);
    // RoPE tables for every position the state can hold
    e->rope_positions = rows_total;
    for (auto& rc : e->ropes) {
        std::vector<float> cosv((size_t)rows_total * rc.head_dim), sinv((size_t)rows_total * rc.head_dim);
        rope_tables(rc, rows_total, cosv.data(), sinv.data());
        Buf c = dev(cosv.size() * 4), s = dev(sinv.size() * 4);
        cudaMemcpy((void*)c.ptr(), cosv.data(), cosv.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy((void*)s.ptr(), sinv.data(), sinv.size() * 4, cudaMemcpyHostToDevice);
        e->rope_cos.push_back(c);
        e->rope_sin.push_back(s);
    }
    {   // one scratch row block for the input-transformed activation of RHT linears
        uint32_t max_in = 0;
        auto visit = [&](Linear& l) { if (l.rht()) max_in = std::max(max_in, l.in_dim); };
        auto each = [&](auto&& f) {
            for (auto& L : e->layers) {
                f(L.up); f(L.down);
                if (L.is_attention) { f(L.attn.qkv); f(L.attn.out); if (L.attn.has_gate) f(L.attn.gate); }
                else { f(L.dn.in_proj); f(L.dn.out_proj); }
            }
        };
        each(visit);
        if (max_in) {
            if (e->tp_sharded) throw std::runtime_error("tensor-parallel shards of RHT (HybridSpec) linears are not supported");
            e->rht_scratch = dev((size_t)MAX_ROWS * max_in * 2);
            each([&](Linear& l) { if (l.rht()) l.rht_scratch = e->rht_scratch.ptr(); });
        }
    }
    e->trie_nodes = dev(MAX_TRIE * sizeof(uzu_trie_node));
    e->host_trie = make_buf(e, MAX_TRIE * (sizeof(uzu_trie_node) + 8 + 4 + 4), UZU_BUFFER_PINNED_HOST);
    for (auto& rc : e->ropes) {
        e->trie_cos.push_back(dev((size_t)MAX_TRIE * rc.head_dim * 4));
        e->trie_sin.push_back(dev((size_t)MAX_TRIE * rc.head_dim * 4));
    }
    cudaEventCreateWithFlags(&e->step_events[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&e->step_events[1], cudaEventDisableTiming);
}

// One more independent sequence state (KV caches, DeltaNet states) with the same geometry as e->state: multi-sequence batched decode.
static void alloc_sequence_state(uzu_engine* e, std::vector<LayerState>& st) {
    const uint32_t rows_total = e->max_context + MAX_ROWS;
    const bool sparse = (uzu_context_device_capabilities(e->ctx) & UZU_CAP_SPARSE_BUFFERS) != 0;
    st.assign(e->layers.size(), LayerState{});
    for (size_t i = 0; i < e->layers.size(); ++i) {
        const Layer& L = e->layers[i];
        LayerState& S = st[i];
        if (L.is_attention) {
            S.row_bytes = (size_t)L.attn.num_groups * L.attn.head_dim * 2;
            const size_t bytes = (size_t)rows_total * S.row_bytes;
            if (sparse) {
                check(uzu_sparse_buffer_create(e->ctx, bytes, &S.k_sparse));
                check(uzu_sparse_buffer_create(e->ctx, bytes, &S.v_sparse));
                S.keys = uzu_sparse_buffer_gpu_ptr(S.k_sparse);
                S.values = uzu_sparse_buffer_gpu_ptr(S.v_sparse);
            } else {
                S.k_dense = make_buf(e, bytes, UZU_BUFFER_DEVICE);
                S.v_dense = make_buf(e, bytes, UZU_BUFFER_DEVICE);
                S.keys = S.k_dense.ptr();
                S.values = S.v_dense.ptr();
            }
        } else {
            const DeltaNetLayer& D = L.dn;
            S.conv_bytes = (size_t)D.conv_dim * (D.kernel_size - 1) * 4;
            S.ssm_bytes = (size_t)D.num_heads * D.value_head_dim * D.head_dim * 4;
            S.conv_state = make_buf(e, S.conv_bytes, UZU_BUFFER_DEVICE);
            S.ssm_state = make_buf(e, S.ssm_bytes, UZU_BUFFER_DEVICE);
            S.conv_snapshot = make_buf(e, S.conv_bytes, UZU_BUFFER_DEVICE);
            S.ssm_snapshot = make_buf(e, S.ssm_bytes, UZU_BUFFER_DEVICE);
            cudaMemsetAsync((void*)S.conv_state.ptr(), 0, S.conv_bytes, e->ctx->stream);
            cudaMemsetAsync((void*)S.ssm_state.ptr(), 0, S.ssm_bytes, e->ctx->stream);
        }
    }
}

// TransformerState::prepare (state.rs:141-172): make sure KV pages for rows [0, rows) are mapped
static void state_prepare(uzu_engine* e, uint32_t rows_needed) {
    for (auto& S : e->state) {
        if (!S.k_sparse) continue;
        const size_t page = uzu_sparse_buffer_page_size_bytes(S.k_sparse);
        const uint32_t need = (uint32_t)(((size_t)rows_needed * S.row_bytes + page - 1) / page);
        if (need <= S.mapped_pages) continue;
        std::vector<uint32_t> pages;
        for (uint32_t p = S.mapped_pages; p < need; ++p) pages.push_back(p);
        check(uzu_sparse_buffer_map(S.k_sparse, pages.data(), pages.size()));
        check(uzu_sparse_buffer_map(S.v_sparse, pages.data(), pages.size()));
        S.mapped_pages = need;
    }
}

static void upload_decode_state(uzu_engine* e);

static void reset_state(uzu_engine* e) {
    cudaStreamSynchronize(e->ctx->stream);
    e->context_length = 0;
    e->snapshot_context = 0;
    e->trie_pending = 0;
    for (auto& S : e->state) {
        S.length = 0;
        if (S.conv_state.b) {
            cudaMemsetAsync((void*)S.conv_state.ptr(), 0, S.conv_bytes, e->ctx->stream);
            cudaMemsetAsync((void*)S.ssm_state.ptr(), 0, S.ssm_bytes, e->ctx->stream);
        }
    }
    e->steps_issued = e->steps_returned = 0;
    cudaMemsetAsync((void*)e->token_ids.ptr(), 0, 4, e->ctx->stream);
    cudaStreamSynchronize(e->ctx->stream);
    upload_decode_state(e);
}

// -----------------------------------------------------------------------------------------------------
// encoding one forward pass (Decoder::encode)
// -----------------------------------------------------------------------------------------------------
static void encode_linear(uzu_command_buffer* cmd, const Linear& l, uint64_t a, uint32_t m, uint64_t d) {
    uzu_matmul_args ma{};
    ma.a = a;
    ma.b = l.w.values.ptr();
    ma.b_scales = l.w.scales.ptr();
    ma.b_zero_points = l.w.zero_points.ptr();
    ma.b_biases = l.w.biases.ptr();
    ma.d = d;
    ma.b_prologue = l.w.prologue;
    ma.b_mode = l.w.mode;
    ma.b_group_size = l.w.group_size;
    ma.b_transpose = 1;                 // linear/matmul.rs:136
    ma.ab_scale = 1.0f;
    ma.m = m; ma.n = l.out_dim; ma.k = l.in_dim;
    ma.weights_dt = ma.input_dt = ma.output_dt = UZU_DT_BF16;   // language_model/mod.rs:74
    if (l.rht()) {   // RHTLinearWrapper::encode_input (rht_wrapper.rs:286-296): InputRht, then the inner matmul with the output factors
        uzu_activation_transform_args t{};
        t.input = a; t.fp_out = l.rht_scratch; t.rht_factors = l.in_signs.ptr();
        t.batch_size = m; t.element_count = l.in_dim;
        t.ops = UZU_ACTIVATION_TRANSFORM_INPUT_RHT; t.in_place = 0; t.data_type = UZU_DT_BF16;
        uzu_activation_transform_encode(cmd, &t);
        ma.a = l.rht_scratch;
        ma.rht_factors = l.out_signs.ptr();
        ma.d_transform |= UZU_D_RHT;
    }
    uzu_matmul_encode(cmd, &ma);
}

// Row-parallel projection (attention out / MLP down). Unsharded: the plain linear. Tensor-parallel shard: this rank's K slice gives an
// f32 partial [m, H]; uzu_tp_all_reduce sums it over the ranks and rounds to bf16 once (the unsharded kernel's rounding point).
static void encode_row_parallel(uzu_engine* e, uzu_command_buffer* cmd, const Linear& l, uint64_t a, uint32_t m, uint64_t d) {
    if (!e->tp_sharded) {
        encode_linear(cmd, l, a, m, d);
        return;
    }
    uzu_matmul_args ma{};
    ma.a = a;
    ma.b = l.w.values.ptr(); ma.b_scales = l.w.scales.ptr(); ma.b_zero_points = l.w.zero_points.ptr(); ma.b_biases = l.w.biases.ptr();
    ma.d = e->tp_partial.ptr();
    ma.b_prologue = l.w.prologue; ma.b_mode = l.w.mode; ma.b_group_size = l.w.group_size; ma.b_transpose = 1; ma.ab_scale = 1.0f;
    ma.m = m; ma.n = l.out_dim; ma.k = l.in_dim;
    ma.weights_dt = ma.input_dt = UZU_DT_BF16;
    ma.output_dt = UZU_DT_F32;
    uzu_matmul_encode(cmd, &ma);
    uzu_tp_all_reduce_encode(cmd, e->tp_partial.ptr(), m * l.out_dim, d);
}

// Readout (embedding.rs:374-456). Tensor-parallel shard: local vocabulary rows, then all-gather into the full logits rows.
static void encode_readout(uzu_engine* e, uzu_command_buffer* cmd, uint64_t normed, uint32_t rows) {
    if (!e->tp_sharded) {
        encode_linear(cmd, e->out_emb, normed, rows, e->logits.ptr());
        return;
    }
    encode_linear(cmd, e->out_emb, normed, rows, e->logits_local.ptr());
    uzu_tp_all_gather_args ga{e->logits_local.ptr(), e->logits.ptr(), e->tp_gather.ptr(), rows, e->vocab_local};
    uzu_tp_all_gather_encode(cmd, &ga);
}

enum ShortcutMode { ShortcutNone, ShortcutCopy, ShortcutAdd };

static void encode_norm(uzu_command_buffer* cmd, const Norm& n, uint64_t input, uint64_t output, uint64_t shortcut, ShortcutMode mode, uint32_t rows) {
    uzu_normalization_args a{};
    a.input = input;
    a.scales = n.scales.ptr();
    a.output = output;
    a.shortcut = mode == ShortcutNone ? 0 : shortcut;
    a.batch_size = rows;
    a.element_count = n.n;
    a.epsilon = n.cfg.epsilon;
    a.scale_offset = n.cfg.scale_offset;
    a.post_layer_scalar = 1.0f;
    a.subtract_mean = n.cfg.subtract_mean;
    a.full_layer = n.cfg.full_layer;
    a.copy_to_shortcut = mode != ShortcutNone;
    a.residual_add = mode == ShortcutAdd;
    a.has_scales = n.cfg.has_scale;
    uzu_normalization_encode(cmd, &a);
}

struct PassCtx {
    uint32_t m = 1;
    bool dynamic = false;   // decode-graph mode: positions come from the device DecodeState
    uint64_t trie = 0;      // speculation pass: device uzu_trie_node[m] (BatchTopology not flat); RoPE rows come from e->trie_cos/sin
};

// q/k norms, RoPE + KV append, attention core, optional sigmoid gate: everything between the qkv and the out projections
static void encode_attention_mix(uzu_engine* e, uzu_command_buffer* cmd, const Layer& L, LayerState& S, const PassCtx& pc, bool with_gate);

static uint64_t encode_attention(uzu_engine* e, uzu_command_buffer* cmd, const Layer& L, LayerState& S, uint64_t hidden, const PassCtx& pc) {
    const AttentionLayer& A = L.attn;
    // gate projection first (mode.rs:54-61); `hidden` is not modified by our linears, so no copy is needed
    if (A.has_gate) encode_linear(cmd, A.gate, hidden, pc.m, e->gate.ptr());
    encode_linear(cmd, A.qkv, hidden, pc.m, e->qkv.ptr());
    encode_attention_mix(e, cmd, L, S, pc, true);
    encode_row_parallel(e, cmd, A.out, e->attn_out.ptr(), pc.m, e->mixer_out.ptr());
    return e->mixer_out.ptr();
}

static void encode_attention_mix(uzu_engine* e, uzu_command_buffer* cmd, const Layer& L, LayerState& S, const PassCtx& pc, bool with_gate) {
    const AttentionLayer& A = L.attn;
    const uint32_t m = pc.m, D = A.head_dim, Hq = A.num_heads, Hkv = A.num_groups;
    const uint64_t dyn = pc.dynamic ? e->decode_state.ptr() : 0;  // &DecodeState::position (first member)
    const uint32_t total_heads = Hq + 2 * Hkv;
    auto qkn = [&](const Norm& n, uint32_t off, uint32_t cnt) {
        if (!n.present || cnt == 0) return;
        uzu_qkv_norm_args qa{};
        qa.scales = n.scales.ptr();
        qa.qkv_output = e->qkv.ptr();
        qa.batch_size = m; qa.total_heads = total_heads; qa.head_dim = D;
        qa.epsilon = n.cfg.epsilon; qa.scale_offset = n.cfg.scale_offset;
        qa.head_offset = off; qa.head_count = cnt; qa.full_layer = n.cfg.full_layer;
        qa.in_place = 1; qa.has_scales = n.cfg.has_scale;
        uzu_qkv_norm_encode(cmd, &qa);
    };
    qkn(A.qnorm, 0, Hq);
    qkn(A.knorm, Hq, Hkv);
    const uint32_t prefix = S.length;
    uzu_attention_prepare_args pa{};
    pa.qkv = e->qkv.ptr(); pa.queries = e->queries.ptr();
    pa.keys = S.keys; pa.values = S.values;
    pa.num_q_heads = Hq; pa.num_kv_heads = Hkv; pa.head_dim = D;
    pa.kv_token_offset = prefix; pa.batch_dim = m; pa.has_kv = 1;
    if (A.rope_index >= 0) {
        const RopeCfg& rc = e->ropes[A.rope_index];
        pa.has_rope = 1; pa.rope_dim = rc.head_dim;
        const size_t row0 = pc.dynamic ? 0 : (size_t)e->context_length;   // token positions = context_length + i (transformer.rs:246-247)
        pa.cosines = e->rope_cos[A.rope_index].ptr() + row0 * rc.head_dim * 4;
        pa.sines = e->rope_sin[A.rope_index].ptr() + row0 * rc.head_dim * 4;
        if (pc.trie) {   // positions = context_length + height (transformer.rs:248): rows gathered by trie_rope_gather_kernel
            pa.cosines = e->trie_cos[A.rope_index].ptr();
            pa.sines = e->trie_sin[A.rope_index].ptr();
        }
    }
    pa.dynamic_position = dyn;
    uzu_attention_prepare_encode(cmd, &pa);

    uzu_attention_args aa{};
    aa.queries = e->queries.ptr(); aa.keys = S.keys; aa.values = S.values;
    aa.gqa_factor = Hq / Hkv;
    aa.sequence_length = prefix + m;
    aa.k_head_stride = D; aa.k_seq_stride = Hkv * D; aa.v_head_stride = D; aa.v_seq_stride = Hkv * D;
    aa.scale = A.has_scale ? A.scale : 1.0f / sqrtf((float)D);
    aa.num_heads = Hq; aa.suffix_length = m; aa.head_dim = D; aa.is_causal = A.is_causal;
    aa.dynamic_position = dyn;
    if (pc.trie) { aa.is_trie = 1; aa.trie = pc.trie; }   // run_core: trie_core + the nodes (mode.rs:178-184), mask.rs:21-29
    // AttentionCores::encode (core/mod.rs:81-93). This backend has no "gemm" core yet, so like the CPU backend long
    // suffixes go through the single/two-pass kernels. For suffix <= 8 with context > 1024 the reference uses the
    // two-pass core; the fused split-KV single-pass entry point computes the same function in one launch, so the
    // engine uses it for every decode step unless UZU_TWO_PASS=1 asks for the literal dispatch.
    static const bool literal_two_pass = getenv("UZU_TWO_PASS") != nullptr;
    if (literal_two_pass && !pc.dynamic && !pc.trie && m <= 8 && prefix + m > 1024) {
        aa.out = e->tp_parts.ptr(); aa.sums = e->tp_sums.ptr(); aa.maxs = e->tp_maxs.ptr();
        uzu_attention_two_pass1_encode(cmd, &aa);
        uzu_attention_two_pass2_args a2{e->tp_parts.ptr(), e->tp_sums.ptr(), e->tp_maxs.ptr(), e->attn_out.ptr(), Hq, m, D};
        uzu_attention_two_pass2_encode(cmd, &a2);
    } else {
        aa.out = e->attn_out.ptr();
        uzu_attention_single_pass_encode(cmd, &aa);
    }
    if (A.has_gate && with_gate) uzu_sigmoid_gate_encode(cmd, e->gate.ptr(), e->attn_out.ptr(), m * Hq * D);
}

static uint64_t encode_delta_net(uzu_engine* e, uzu_command_buffer* cmd, const Layer& L, LayerState& S, uint64_t hidden, const PassCtx& pc) {
    const DeltaNetLayer& D = L.dn;
    const uint32_t m = pc.m;
    // m > 1 (prefill): the two projections run batched over the m rows (tensor-core GEMM for m >= 64); the recurrence itself -- the
    // rolling conv state and the delta-rule state update -- is the decode branch (delta_net.rs flat m = 1 path) applied row by row in
    // token order, so a batched prefill computes exactly what m single-token passes compute. (The reference's chunked-prefill core,
    // Kernels::DeltaNetChunkedPrefill, evaluates the same recurrence in a different summation order; not restated here.)
    encode_linear(cmd, D.in_proj, hidden, m, e->in_proj.ptr());
    bool whole_pass = false;
    if (m > 1) {   // opt-in: the m-token recurrence of this layer in ONE launch (deltanet_prefill.cu), state kept on chip across tokens
        uzu_delta_net_fused_update_args f0{};
        f0.conv.conv_weight = D.conv_weight.ptr(); f0.conv.bias = D.conv_bias.ptr(); f0.conv.in_out = e->in_proj.ptr(); f0.conv.state = S.conv_state.ptr();
        f0.conv.kernel_size = D.kernel_size; f0.conv.conv_dim = D.conv_dim; f0.conv.state_stride = D.kernel_size - 1; f0.conv.has_bias = D.conv_has_bias;
        f0.update.in_proj = e->in_proj.ptr(); f0.update.a_log = D.a_log.ptr(); f0.update.dt_bias = D.dt_bias.ptr(); f0.update.norm_weight = D.norm_weight.ptr();
        f0.update.state = S.ssm_state.ptr(); f0.update.out = e->delta_out.ptr();
        f0.update.num_v_heads = D.num_heads; f0.update.num_k_heads = D.num_groups; f0.update.head_v_dim = D.value_head_dim; f0.update.key_dim = D.key_dim;
        f0.update.value_dim = D.value_dim; f0.update.norm_epsilon = D.norm_epsilon; f0.update.head_k_dim = D.head_dim;
        whole_pass = encode_delta_net_prefill(cmd, f0, m, D.total_proj_dim, D.value_dim);
    }
    for (uint32_t t = 0; t < m && !whole_pass; ++t) {
        const uint64_t row = e->in_proj.ptr() + (size_t)t * D.total_proj_dim * 2;
        uzu_delta_net_fused_update_args fa{};
        uzu_delta_net_conv_update_args& ca = fa.conv;
        ca.conv_weight = D.conv_weight.ptr(); ca.bias = D.conv_bias.ptr(); ca.in_out = row; ca.state = S.conv_state.ptr();
        ca.kernel_size = D.kernel_size; ca.conv_dim = D.conv_dim; ca.state_stride = D.kernel_size - 1; ca.has_bias = D.conv_has_bias;
        uzu_delta_net_update_args& ua = fa.update;
        ua.in_proj = row; ua.a_log = D.a_log.ptr(); ua.dt_bias = D.dt_bias.ptr(); ua.norm_weight = D.norm_weight.ptr();
        ua.state = S.ssm_state.ptr(); ua.out = e->delta_out.ptr() + (size_t)t * D.value_dim * 2;
        ua.num_v_heads = D.num_heads; ua.num_k_heads = D.num_groups; ua.head_v_dim = D.value_head_dim; ua.key_dim = D.key_dim;
        ua.value_dim = D.value_dim; ua.norm_epsilon = D.norm_epsilon; ua.head_k_dim = D.head_dim;
        if (m > 1 && uzu_delta_net_fused_update_supported(&fa)) {
            uzu_delta_net_fused_update_encode(cmd, &fa);          // conv + update in one launch per token
        } else {
            uzu_delta_net_conv_update_encode(cmd, &ca);
            uzu_delta_net_update_encode(cmd, &ua);
        }
    }
    encode_linear(cmd, D.out_proj, e->delta_out.ptr(), m, e->mixer_out.ptr());
    return e->mixer_out.ptr();
}

// Decoder::encode for `m` tokens already in e->token_ids; logits for rows [row_begin, row_end) land in e->logits.
static void encode_decoder(uzu_engine* e, uzu_command_buffer* cmd, const PassCtx& pc, uint32_t row_begin, uint32_t row_end) {
    const uint32_t m = pc.m, H = e->model_dim;
    // embedding lookup (embedding.rs:345-372)
    if (e->in_emb.w.prologue == UZU_B_FULL_PRECISION) {
        uzu_full_precision_embedding_lookup_encode(cmd, e->token_ids.ptr(), e->in_emb.w.values.ptr(), e->hidden_a.ptr(), m, e->vocab, H, e->input_scale);
    } else {
        uzu_quantized_embedding_lookup_args la{};
        la.token_ids = e->token_ids.ptr(); la.weights = e->in_emb.w.values.ptr(); la.scales = e->in_emb.w.scales.ptr();
        la.zero_points = e->in_emb.w.zero_points.ptr(); la.biases = e->in_emb.w.biases.ptr(); la.output = e->hidden_a.ptr();
        la.batch_size = m; la.vocab_size = e->vocab; la.model_dim = H; la.input_scale = e->input_scale;
        la.group_size = e->in_emb.w.group_size; la.quantization_mode = e->in_emb.w.mode;
        la.quantization_method = e->in_emb.w.prologue == UZU_B_SCALE_BIAS_DEQUANT ? UZU_QMETHOD_SCALE_BIAS
                                 : e->in_emb.w.prologue == UZU_B_SCALE_ZERO_POINT_DEQUANT ? UZU_QMETHOD_SCALE_ZERO_POINT : UZU_QMETHOD_SCALE_SYMMETRIC;
        uzu_quantized_embedding_lookup_encode(cmd, &la);
    }
    uint64_t hidden = e->hidden_a.ptr();
    for (size_t i = 0; i < e->layers.size(); ++i) {
        Layer& L = e->layers[i];
        LayerState& S = e->state[i];
        // pre_mixer_norm: layer 0 copies the input into the shortcut, later layers add (transformer_layer.rs:95-109)
        encode_norm(cmd, L.pre_mixer, hidden, e->hidden_b.ptr(), e->shortcut.ptr(), i == 0 ? ShortcutCopy : ShortcutAdd, m);
        uint64_t mixed = L.is_attention ? encode_attention(e, cmd, L, S, e->hidden_b.ptr(), pc) : encode_delta_net(e, cmd, L, S, e->hidden_b.ptr(), pc);
        encode_norm(cmd, L.pre_mlp, mixed, e->hidden_b.ptr(), e->shortcut.ptr(), ShortcutAdd, m);
        // DenseMlp::encode (mlp/dense.rs:32-48)
        encode_linear(cmd, L.up, e->hidden_b.ptr(), m, e->fused_up.ptr());
        uzu_gated_act_mul_args ga{};
        ga.act_operand = e->fused_up.ptr(); ga.fp_out = e->gated.ptr(); ga.gated_dim = L.hidden_dim; ga.batch_dim = m;
        ga.act_type = L.act; ga.interleaved = 1;
        uzu_gated_act_mul_encode(cmd, &ga);
        encode_row_parallel(e, cmd, L.down, e->gated.ptr(), m, e->hidden_a.ptr());
        hidden = e->hidden_a.ptr();
    }
    if (row_end <= row_begin) return;
    const uint32_t rows = row_end - row_begin;
    // output_norm over the requested rows, residual add into the same rows of the shortcut (transformer.rs:317-323)
    encode_norm(cmd, e->out_norm, hidden + (size_t)row_begin * H * 2, e->normed_out.ptr(), e->shortcut.ptr() + (size_t)row_begin * H * 2, ShortcutAdd, rows);
    encode_readout(e, cmd, e->normed_out.ptr(), rows);
    if (e->has_logit_scale || e->has_logit_soft_cap)
        uzu_logit_transform_encode(cmd, e->logits.ptr(), rows * e->vocab, e->has_logit_scale ? e->logit_scale : 1.0f, e->logit_soft_cap, e->has_logit_soft_cap);
}

// ---- fused decode step (m = 1): norm / gated-act / sigmoid-gate launches folded into the consuming GEMV ---------------------
static uzu_matmul_args linear_args(const Linear& l, uint64_t a, uint32_t m, uint64_t d) {
    uzu_matmul_args ma{};
    ma.a = a;
    ma.b = l.w.values.ptr();
    ma.b_scales = l.w.scales.ptr();
    ma.b_zero_points = l.w.zero_points.ptr();
    ma.b_biases = l.w.biases.ptr();
    ma.d = d;
    ma.b_prologue = l.w.prologue;
    ma.b_mode = l.w.mode;
    ma.b_group_size = l.w.group_size;
    ma.b_transpose = 1;
    ma.ab_scale = 1.0f;
    ma.m = m; ma.n = l.out_dim; ma.k = l.in_dim;
    ma.weights_dt = ma.input_dt = ma.output_dt = UZU_DT_BF16;
    return ma;
}

// decode-stream copy of a linear's weights (built with the persistent kernel's program) or 0
static uint64_t g_stream_lookup(const std::map<uint64_t, uint64_t>& m, const Linear& l) {
    auto it = m.find(l.w.values.ptr());
    return it == m.end() ? 0 : it->second;
}

static uzu_fused_linear_args fused_norm_args(const Linear& l, const Norm& n, uint64_t input, uint64_t sc_in, uint64_t sc_out, bool add, uint64_t d) {
    uzu_fused_linear_args f{};
    f.matmul = linear_args(l, 0, 1, d);
    f.prologue = 1;
    f.norm_input = input; f.norm_shortcut_in = add ? sc_in : 0; f.norm_scales = n.scales.ptr(); f.shortcut_out = sc_out;
    f.norm_epsilon = n.cfg.epsilon; f.norm_scale_offset = n.cfg.scale_offset;
    f.norm_residual_add = add; f.norm_full_layer = n.cfg.full_layer;
    return f;
}

// Which launches the fused decode path folds away (UZU_FUSE_MASK, default all):
//   bit 0: pre-mixer / pre-MLP RMSNorm (+ residual add) -> prologue of the consuming GEMV(s)
//   bit 1: GatedActMul -> epilogue of the up GEMV (paired up / gate tiles)
//   bit 2: SigmoidGate -> prologue of the attention out projection
//   bit 3: QKVNorm(q) + QKVNorm(k) + AttentionPrepare -> one launch
//   bit 4: DeltaNetConvUpdate -> inside DeltaNetUpdate (one k head per v head only)
static uint32_t fuse_mask() {
    static const uint32_t m = [] { const char* v = getenv("UZU_FUSE_MASK"); return v ? (uint32_t)atoi(v) : 31u; }();
    return m;
}

static uzu_fused_linear_args fused_up_args(uzu_engine* e, const Layer& L, bool norm_fused, uint64_t sc_in, uint64_t sc_out) {
    const bool gated = (fuse_mask() & 2u) != 0;
    uzu_fused_linear_args f{};
    if (norm_fused) f = fused_norm_args(L.up, L.pre_mlp, e->mixer_out.ptr(), sc_in, sc_out, true, gated ? e->gated.ptr() : e->fused_up.ptr());
    else f.matmul = linear_args(L.up, e->hidden_b.ptr(), 1, gated ? e->gated.ptr() : e->fused_up.ptr());
    if (gated) { f.epilogue = 1; f.act_type = L.act; }
    return f;
}

static bool any_rht(const uzu_engine* e) {
    for (auto& L : e->layers) {
        if (L.up.rht() || L.down.rht()) return true;
        if (L.is_attention ? (L.attn.qkv.rht() || L.attn.out.rht() || L.attn.gate.rht()) : (L.dn.in_proj.rht() || L.dn.out_proj.rht())) return true;
    }
    return e->out_emb.rht();
}

static bool fused_decode_supported(uzu_engine* e) {
    const uint32_t mask = fuse_mask();
    if (any_rht(e)) return false;   // RHT linears take the unfused sequence (input transform, GEMV, output transform)
    for (auto& L : e->layers) {
        if (L.pre_mixer.cfg.subtract_mean || L.pre_mlp.cfg.subtract_mean || !L.pre_mixer.cfg.has_scale || !L.pre_mlp.cfg.has_scale) return false;
        std::vector<uzu_fused_linear_args> fs;
        const uint64_t x = e->hidden_a.ptr(), s0 = e->shortcut.ptr(), s1 = e->shortcut2.ptr();
        if (mask & 1u) {
            if (L.is_attention) {
                if (L.attn.has_gate) fs.push_back(fused_norm_args(L.attn.gate, L.pre_mixer, x, s0, s1, true, e->gate.ptr()));
                fs.push_back(fused_norm_args(L.attn.qkv, L.pre_mixer, x, s0, s1, true, e->qkv.ptr()));
            } else {
                fs.push_back(fused_norm_args(L.dn.in_proj, L.pre_mixer, x, s0, s1, true, e->in_proj.ptr()));
            }
        }
        if (L.is_attention && L.attn.has_gate && (mask & 4u)) {
            uzu_fused_linear_args g{};
            g.matmul = linear_args(L.attn.out, 0, 1, e->mixer_out.ptr());
            g.prologue = 3; g.sg_attn = e->attn_out.ptr(); g.sg_gate = e->gate.ptr();
            fs.push_back(g);
        }
        if (mask & 3u) fs.push_back(fused_up_args(e, L, (mask & 1u) != 0, s0, s1));
        for (auto& f : fs)
            if (!uzu_fused_linear_supported(e->ctx, &f)) return false;
    }
    return true;
}

// a plain m = 1 GEMV of the decode step: through the TMA-fed stream variant when the linear has a decode-stream copy
static void encode_decode_gemv(uzu_engine* e, uzu_command_buffer* cmd, const Linear& lin, uint64_t x, uint64_t d) {
    const uint64_t stream = e->tp_sharded ? 0 : g_stream_lookup(e->mega.stream_of, lin);
    if (stream) {
        uzu_fused_linear_args f{};
        f.matmul = linear_args(lin, x, 1, d);
        f.decode_stream = stream;
        if (uzu_fused_linear_supported(e->ctx, &f)) {
            uzu_fused_linear_encode(cmd, &f);
            return;
        }
    }
    encode_row_parallel(e, cmd, lin, x, 1, d);
}

static void encode_decoder_fused(uzu_engine* e, uzu_command_buffer* cmd, const PassCtx& pc) {
    const uint32_t H = e->model_dim;
    const uint32_t mask = fuse_mask();
    const bool fuse_norm = (mask & 1u) != 0, fuse_gated = (mask & 2u) != 0, fuse_sigmoid = (mask & 4u) != 0, fuse_qknorm = (mask & 8u) != 0, fuse_conv = (mask & 16u) != 0;
    // embedding lookup (embedding.rs:345-372)
    if (e->in_emb.w.prologue == UZU_B_FULL_PRECISION) {
        uzu_full_precision_embedding_lookup_encode(cmd, e->token_ids.ptr(), e->in_emb.w.values.ptr(), e->hidden_a.ptr(), 1, e->vocab, H, e->input_scale);
    } else {
        uzu_quantized_embedding_lookup_args la{};
        la.token_ids = e->token_ids.ptr(); la.weights = e->in_emb.w.values.ptr(); la.scales = e->in_emb.w.scales.ptr();
        la.zero_points = e->in_emb.w.zero_points.ptr(); la.biases = e->in_emb.w.biases.ptr(); la.output = e->hidden_a.ptr();
        la.batch_size = 1; la.vocab_size = e->vocab; la.model_dim = H; la.input_scale = e->input_scale;
        la.group_size = e->in_emb.w.group_size; la.quantization_mode = e->in_emb.w.mode;
        la.quantization_method = e->in_emb.w.prologue == UZU_B_SCALE_BIAS_DEQUANT ? UZU_QMETHOD_SCALE_BIAS
                                 : e->in_emb.w.prologue == UZU_B_SCALE_ZERO_POINT_DEQUANT ? UZU_QMETHOD_SCALE_ZERO_POINT : UZU_QMETHOD_SCALE_SYMMETRIC;
        uzu_quantized_embedding_lookup_encode(cmd, &la);
    }
    // With the norm folded into its consumers the residual ping-pongs between two buffers (every consumer CTA re-reads the old
    // residual while CTA 0 of the first consumer writes the new one); with a standalone norm it is updated in place in S[cur].
    uint64_t S[2] = {e->shortcut.ptr(), e->shortcut2.ptr()};
    int cur = 0;
    const uint64_t dyn = e->decode_state.ptr();
    for (size_t i = 0; i < e->layers.size(); ++i) {
        Layer& L = e->layers[i];
        LayerState& St = e->state[i];
        const bool add = i > 0;   // layer 0: Copy mode (transformer_layer.rs:95-109)
        bool wrote = false;
        // `lin` consumes pre_mixer_norm(hidden_a): fused prologue or the standalone norm's output in hidden_b
        auto mixer_in_linear = [&](const Linear& lin, uint64_t d) {
            if (fuse_norm) {
                auto f = fused_norm_args(lin, L.pre_mixer, e->hidden_a.ptr(), S[cur], wrote ? 0 : S[cur ^ 1], add, d);
                f.decode_stream = g_stream_lookup(e->mega.stream_of, lin);
                uzu_fused_linear_encode(cmd, &f);
                wrote = true;
            } else {
                encode_linear(cmd, lin, e->hidden_b.ptr(), 1, d);
            }
        };
        if (!fuse_norm) encode_norm(cmd, L.pre_mixer, e->hidden_a.ptr(), e->hidden_b.ptr(), S[cur], add ? ShortcutAdd : ShortcutCopy, 1);
        if (L.is_attention) {
            const AttentionLayer& A = L.attn;
            const uint32_t D = A.head_dim, Hq = A.num_heads, Hkv = A.num_groups;
            if (A.has_gate) mixer_in_linear(A.gate, e->gate.ptr());
            mixer_in_linear(A.qkv, e->qkv.ptr());
            const uint32_t total_heads = Hq + 2 * Hkv;
            uzu_attention_prepare_args pa{};
            pa.qkv = e->qkv.ptr(); pa.queries = e->queries.ptr(); pa.keys = St.keys; pa.values = St.values;
            pa.num_q_heads = Hq; pa.num_kv_heads = Hkv; pa.head_dim = D; pa.kv_token_offset = St.length; pa.batch_dim = 1; pa.has_kv = 1;
            if (A.rope_index >= 0) {
                const RopeCfg& rc = e->ropes[A.rope_index];
                pa.has_rope = 1; pa.rope_dim = rc.head_dim;
                pa.cosines = e->rope_cos[A.rope_index].ptr(); pa.sines = e->rope_sin[A.rope_index].ptr();
            }
            pa.dynamic_position = dyn;
            if (fuse_qknorm && (A.qnorm.present || A.knorm.present) && D <= 256) {
                uzu_attention_prepare_norm_args pn{};
                pn.prepare = pa;
                auto cfg = [](const Norm& nm) {
                    uzu_qk_norm_config c{};
                    c.present = nm.present; c.scales = nm.scales.ptr(); c.epsilon = nm.cfg.epsilon; c.scale_offset = nm.cfg.scale_offset;
                    c.full_layer = nm.cfg.full_layer; c.has_scales = nm.cfg.has_scale;
                    return c;
                };
                pn.q_norm = cfg(A.qnorm); pn.k_norm = cfg(A.knorm);
                uzu_attention_prepare_norm_encode(cmd, &pn);
            } else {
                auto qkn = [&](const Norm& n, uint32_t off, uint32_t cnt) {
                    if (!n.present || cnt == 0) return;
                    uzu_qkv_norm_args qa{};
                    qa.scales = n.scales.ptr(); qa.qkv_output = e->qkv.ptr();
                    qa.batch_size = 1; qa.total_heads = total_heads; qa.head_dim = D;
                    qa.epsilon = n.cfg.epsilon; qa.scale_offset = n.cfg.scale_offset;
                    qa.head_offset = off; qa.head_count = cnt; qa.full_layer = n.cfg.full_layer;
                    qa.in_place = 1; qa.has_scales = n.cfg.has_scale;
                    uzu_qkv_norm_encode(cmd, &qa);
                };
                qkn(A.qnorm, 0, Hq);
                qkn(A.knorm, Hq, Hkv);
                uzu_attention_prepare_encode(cmd, &pa);
            }
            uzu_attention_args aa{};
            aa.queries = e->queries.ptr(); aa.keys = St.keys; aa.values = St.values; aa.out = e->attn_out.ptr();
            aa.gqa_factor = Hq / Hkv; aa.sequence_length = St.length + 1;
            aa.k_head_stride = D; aa.k_seq_stride = Hkv * D; aa.v_head_stride = D; aa.v_seq_stride = Hkv * D;
            aa.scale = A.has_scale ? A.scale : 1.0f / sqrtf((float)D);
            aa.num_heads = Hq; aa.suffix_length = 1; aa.head_dim = D; aa.is_causal = A.is_causal; aa.dynamic_position = dyn;
            uzu_attention_single_pass_encode(cmd, &aa);
            if (A.has_gate && fuse_sigmoid && !e->tp_sharded) {
                uzu_fused_linear_args g{};
                g.matmul = linear_args(A.out, 0, 1, e->mixer_out.ptr());
                g.prologue = 3; g.sg_attn = e->attn_out.ptr(); g.sg_gate = e->gate.ptr();
                g.decode_stream = g_stream_lookup(e->mega.stream_of, A.out);
                uzu_fused_linear_encode(cmd, &g);
            } else {
                if (A.has_gate) uzu_sigmoid_gate_encode(cmd, e->gate.ptr(), e->attn_out.ptr(), Hq * D);
                encode_decode_gemv(e, cmd, A.out, e->attn_out.ptr(), e->mixer_out.ptr());
            }
        } else {
            const DeltaNetLayer& Dn = L.dn;
            mixer_in_linear(Dn.in_proj, e->in_proj.ptr());
            uzu_delta_net_fused_update_args fa{};
            uzu_delta_net_conv_update_args& ca = fa.conv;
            ca.conv_weight = Dn.conv_weight.ptr(); ca.bias = Dn.conv_bias.ptr(); ca.in_out = e->in_proj.ptr(); ca.state = St.conv_state.ptr();
            ca.kernel_size = Dn.kernel_size; ca.conv_dim = Dn.conv_dim; ca.state_stride = Dn.kernel_size - 1; ca.has_bias = Dn.conv_has_bias;
            uzu_delta_net_update_args& ua = fa.update;
            ua.in_proj = e->in_proj.ptr(); ua.a_log = Dn.a_log.ptr(); ua.dt_bias = Dn.dt_bias.ptr(); ua.norm_weight = Dn.norm_weight.ptr();
            ua.state = St.ssm_state.ptr(); ua.out = e->delta_out.ptr();
            ua.num_v_heads = Dn.num_heads; ua.num_k_heads = Dn.num_groups; ua.head_v_dim = Dn.value_head_dim; ua.key_dim = Dn.key_dim;
            ua.value_dim = Dn.value_dim; ua.norm_epsilon = Dn.norm_epsilon; ua.head_k_dim = Dn.head_dim;
            if (fuse_conv && uzu_delta_net_fused_update_supported(&fa)) {
                uzu_delta_net_fused_update_encode(cmd, &fa);
            } else {
                uzu_delta_net_conv_update_encode(cmd, &ca);
                uzu_delta_net_update_encode(cmd, &ua);
            }
            encode_decode_gemv(e, cmd, Dn.out_proj, e->delta_out.ptr(), e->mixer_out.ptr());
        }
        if (fuse_norm) cur ^= 1;
        // pre_mlp_norm + DenseMlp (mlp/dense.rs:32-48)
        if (!fuse_norm) encode_norm(cmd, L.pre_mlp, e->mixer_out.ptr(), e->hidden_b.ptr(), S[cur], ShortcutAdd, 1);
        if (fuse_norm || fuse_gated) {
            auto fu = fused_up_args(e, L, fuse_norm, S[cur], S[cur ^ 1]);
            fu.decode_stream = g_stream_lookup(e->mega.stream_of, L.up);
            uzu_fused_linear_encode(cmd, &fu);
        } else {
            encode_linear(cmd, L.up, e->hidden_b.ptr(), 1, e->fused_up.ptr());
        }
        if (fuse_norm) cur ^= 1;
        if (!fuse_gated) {
            uzu_gated_act_mul_args ga{};
            ga.act_operand = e->fused_up.ptr(); ga.fp_out = e->gated.ptr(); ga.gated_dim = L.hidden_dim; ga.batch_dim = 1;
            ga.act_type = L.act; ga.interleaved = 1;
            uzu_gated_act_mul_encode(cmd, &ga);
        }
        encode_decode_gemv(e, cmd, L.down, e->gated.ptr(), e->hidden_a.ptr());
    }
    // output norm with the residual add in place on the current residual buffer (transformer.rs:317-323), readout, logit transform
    encode_norm(cmd, e->out_norm, e->hidden_a.ptr(), e->normed_out.ptr(), S[cur], ShortcutAdd, 1);
    if (!e->tp_sharded && g_stream_lookup(e->mega.stream_of, e->out_emb)) encode_decode_gemv(e, cmd, e->out_emb, e->normed_out.ptr(), e->logits.ptr());
    else encode_readout(e, cmd, e->normed_out.ptr(), 1);
    if (e->has_logit_scale || e->has_logit_soft_cap)
        uzu_logit_transform_encode(cmd, e->logits.ptr(), e->vocab, e->has_logit_scale ? e->logit_scale : 1.0f, e->logit_soft_cap, e->has_logit_soft_cap);
    (void)pc;
}


// =====================================================================================================
// persistent decode kernel: compile the model into a phase program (decode_mega.cu runs it, one launch per token)
// =====================================================================================================
namespace {

struct MegaBuilder {
    uzu_engine* e;
    uzu_engine::Mega& mg;
    uint32_t W;              // consumer warps of the whole grid
    uint32_t grid = 0;
    uint32_t bits = 0, group_size = 0;
    size_t scratch = 0;
    std::map<uint64_t, Buf> streams;   // original values pointer -> decode-stream copy

    explicit MegaBuilder(uzu_engine* e_) : e(e_), mg(e_->mega), W(0) {}

    void need(bool cond, const char* why) {
        if (!cond) throw std::runtime_error(why);
    }
    Buf dev_zero(size_t bytes) {
        Buf b = make_buf(e, std::max<size_t>(bytes, 256), UZU_BUFFER_DEVICE);
        cudaMemsetAsync((void*)b.ptr(), 0, std::max<size_t>(bytes, 256), e->ctx->stream);
        return b;
    }
    void check_linear(const Linear& l) {
        const WeightMatrix& w = l.w;
        need(!l.rht(), "RHT (HybridSpec) linear");
        need(w.prologue != UZU_B_FULL_PRECISION, "full-precision linear");
        need(w.mode == UZU_QMODE_U4 || w.mode == UZU_QMODE_U8, "signed codes");
        if (!bits) { bits = w.bits; group_size = w.group_size; }
        need(w.bits == bits && w.group_size == group_size, "mixed quantisation geometry");
        need(l.in_dim <= 16384 && (l.in_dim % 64) == 0 && (l.in_dim * bits / 8) % 16 == 0, "input dimension");
        need((l.out_dim % 4) == 0, "output dimension");
    }
    const uint8_t* stream_of(const Linear& l) {
        const uint64_t key = l.w.values.ptr();
        auto it = streams.find(key);
        if (it == streams.end()) {
            const size_t bytes = mega_stream_bytes(l.out_dim, l.in_dim, bits);
            Buf b = make_buf(e, bytes, UZU_BUFFER_DEVICE);
            const uint32_t method = l.w.prologue == UZU_B_SCALE_BIAS_DEQUANT ? UZU_QMETHOD_SCALE_BIAS
                                    : l.w.prologue == UZU_B_SCALE_ZERO_POINT_DEQUANT ? UZU_QMETHOD_SCALE_ZERO_POINT : UZU_QMETHOD_SCALE_SYMMETRIC;
            mega_repack(e->ctx, (const uint8_t*)l.w.values.ptr(), (const __nv_bfloat16*)l.w.scales.ptr(), (const uint8_t*)l.w.zero_points.ptr(),
                        (const __nv_bfloat16*)l.w.biases.ptr(), l.out_dim, l.in_dim, bits, group_size, method, (uint8_t*)b.ptr());
            mg.stream_bytes += bytes;
            mg.stream_of[key] = b.ptr();
            it = streams.emplace(key, b).first;
        }
        return (const uint8_t*)it->second.ptr();
    }
    static uint32_t range_of(uint64_t u, uint64_t U, uint64_t Weff) { return (uint32_t)(((u + 1) * Weff - 1) / U); }

    // one GEMV phase over 1..2 matrices sharing the input row; returns the op index
    size_t gemv(std::initializer_list<const Linear*> mats) {
        MkOp op{};
        op.kind = MK_GEMV;
        uint32_t unit0 = 0, nm = 0;
        for (const Linear* l : mats) {
            MkMat& M = op.mat[nm++];
            M.n = l->out_dim; M.k = l->in_dim;
            M.tiles = (l->out_dim + 15) / 16;
            M.C = (l->in_dim * bits / 4 + 511) / 512;
            M.bias_form = l->w.prologue == UZU_B_SCALE_BIAS_DEQUANT;
            M.unit0 = unit0;
            M.stream = stream_of(*l);
            unit0 += M.tiles * M.C;
            op.k = l->in_dim;
        }
        op.nmat = nm;
        op.units = unit0;
        need((uint64_t)unit0 * (grid + 1) < (1ull << 32), "phase too large for the 32-bit partition arithmetic");
        const uint64_t U = unit0, Weff = std::min<uint64_t>(grid, U);      // pieces per tile = CTAs whose range touches it (decode_mega.cu mk_my_range)
        for (uint32_t i = 0; i < nm; ++i) {
            MkMat& M = op.mat[i];
            need(M.k == op.k, "matrices of a phase must share the input row");
            uint32_t P = 1;
            for (uint32_t t = 0; t < M.tiles; ++t) {
                const uint32_t a = range_of((uint64_t)M.unit0 + (uint64_t)t * M.C, U, Weff), b = range_of((uint64_t)M.unit0 + (uint64_t)(t + 1) * M.C - 1, U, Weff);
                P = std::max(P, b - a + 1);
            }
            need(P <= 64, "too many pieces per tile");
            M.P = P;
            // unused slots of a tile are never written: they keep these zeros, so consumers sum all P slots without a per-tile count
            Buf pieces = dev_zero((size_t)M.tiles * P * 16 * 4);
            M.pieces = (float*)pieces.ptr();
        }
        const uint32_t C = op.mat[0].C, gps = 512 / (group_size * bits / 4);
        // xs + zero block + sx, plus (RMSNorm inputs) a bf16 copy of the row; the norm is only ever applied to model_dim rows
        scratch = std::max(scratch, (size_t)(C * 64 + 4) * 16 + (size_t)C * gps * 4 + 64);
        mg.ops.push_back(op);
        return mg.ops.size() - 1;
    }
    MkPieces pieces_of(size_t op_index, int mat) {
        const MkMat& M = mg.ops[op_index].mat[mat];
        MkPieces pc{};
        pc.pieces = M.pieces;
        pc.P = M.P;
        return pc;
    }
    void norm_input(size_t op_index, const Norm& n, bool add, uint64_t sc_in, uint64_t sc_out) {
        MkOp& op = mg.ops[op_index];
        need(n.present || true, "");
        need(!n.cfg.subtract_mean && n.cfg.has_scale, "norm variant");
        need(op.k <= 8192, "normalised rows longer than 8192");
        {   // the RMSNorm staging keeps a bf16 copy of the row behind xs / sx
            const uint32_t C = op.mat[0].C, gps = 512 / (group_size * bits / 4);
            scratch = std::max(scratch, (size_t)(C * 64 + 4) * 16 + (size_t)C * gps * 4 + 64 + (size_t)op.k * 2);
        }
        op.in_kind = MK_IN_NORM;
        op.shortcut_in = (const __nv_bfloat16*)sc_in;
        op.shortcut_out = (__nv_bfloat16*)sc_out;
        op.norm_scales = (const float*)n.scales.ptr();
        op.norm_eps = n.cfg.epsilon; op.norm_scale_offset = n.cfg.scale_offset;
        op.norm_residual_add = add; op.norm_full_layer = n.cfg.full_layer;
    }

    void build() {
        const char* env = getenv("UZU_MEGA");
        need(!env || atoi(env) != 0, "disabled by UZU_MEGA=0");
        need(!e->tp_sharded, "tensor-parallel shard");
        need(!e->has_logit_scale && !e->has_logit_soft_cap, "logit transform");
        need((e->vocab % 4) == 0, "vocabulary size");
        need((e->model_dim % 64) == 0, "model dimension");
        for (auto& L : e->layers) {
            if (L.is_attention) {
                check_linear(L.attn.qkv); check_linear(L.attn.out);
                if (L.attn.has_gate) check_linear(L.attn.gate);
                const uint32_t D = L.attn.head_dim, G = L.attn.num_heads / L.attn.num_groups;
                need(D == 64 || D == 128 || D == 256, "head dimension");
                need(G >= 1 && G <= 4 && L.attn.num_heads % L.attn.num_groups == 0, "query heads per kv head");
                need(L.attn.is_causal, "non-causal attention");
                need((uint32_t)e->ctx->sm_count >= L.attn.num_groups, "more kv heads than SMs");
            } else {
                check_linear(L.dn.in_proj); check_linear(L.dn.out_proj);
                need(L.dn.head_dim == 128, "DeltaNet key head dimension");
                need(L.dn.value_head_dim == 64 || L.dn.value_head_dim == 128 || L.dn.value_head_dim == 256, "DeltaNet value head dimension");
                need((L.dn.value_dim % 256) == 0 && (L.dn.key_dim % 128) == 0 && (L.dn.value_dim % 128) == 0, "DeltaNet dimensions");
                need(L.dn.kernel_size >= 2 && L.dn.kernel_size - 1 <= 7, "DeltaNet conv taps");
                need(L.dn.num_heads % L.dn.num_groups == 0, "DeltaNet heads");
            }
            check_linear(L.up); check_linear(L.down);
            need((L.hidden_dim % 16) == 0 && L.up.out_dim == 2 * L.hidden_dim, "MLP dimensions");
        }
        check_linear(e->out_emb);
        need(bits != 0, "no quantised linear");
        const uint32_t npg = group_size * bits / 4;
        // partition parameters first (the piece tables depend on the number of consumer warps)
        MegaConfig probe{};
        need(mega_config(e->ctx, npg, bits, 0, &probe), "quantisation geometry not covered by the persistent kernel");
        W = probe.grid * probe.ncw;
        grid = probe.grid;

        const uint32_t H = e->model_dim;
        uint64_t S[2] = {e->shortcut.ptr(), e->shortcut2.ptr()};
        int cur = 0;
        uint32_t max_attn_scratch = 0, max_parts = 0, max_hk = 1, max_vd = 1, max_kvh = 1;
        for (auto& L : e->layers) {
            if (L.is_attention) {
                const uint32_t G = L.attn.num_heads / L.attn.num_groups, D = L.attn.head_dim;
                // sq [G][D] + sc [G][kp] + stats + so [ceil(ncw / 2)][G][D]  (decode_mega.cu MK_ATTN), kp at the longest context
                const uint32_t cph = probe.grid / L.attn.num_groups, kpw = 32 / (D / 8), step = 16 * kpw;    // 16 >= any warp count of the table
                const uint32_t seq_max = e->max_context + MAX_ROWS;
                const uint32_t kp_max = std::max(((seq_max + cph - 1) / cph + step - 1) / step, 4u) * step;
                max_attn_scratch = std::max<uint32_t>(max_attn_scratch, (G * D + G * kp_max + 16 + 8 * G * D) * 4);
                max_parts = std::max<uint32_t>(max_parts, probe.grid * G * (L.attn.head_dim + 2));
                max_kvh = std::max(max_kvh, L.attn.num_groups);
            } else {
                max_hk = std::max(max_hk, L.dn.num_groups);
                max_vd = std::max(max_vd, L.dn.value_dim);
            }
        }
        scratch = std::max<size_t>(scratch, std::max<size_t>(max_attn_scratch + 64, (256 + 256 + 32) * 4));
        mg.attn_part = dev_zero((size_t)std::max(max_parts, 64u) * 4);
        mg.attn_tickets = dev_zero(max_kvh * 4);
        mg.dn_raw = dev_zero((size_t)max_vd * 4);
        (void)max_hk;
        mg.argmax_keys = dev_zero((size_t)probe.grid * 8);

        size_t prev_down = (size_t)-1;
        for (size_t i = 0; i < e->layers.size(); ++i) {
            Layer& L = e->layers[i];
            LayerState& St = e->state[i];
            auto mixer_source = [&](size_t oi) {
                MkOp& op = mg.ops[oi];
                if (i == 0) {
                    op.src_kind = MK_SRC_EMBED;
                    MkEmbed& E = op.embed;
                    const WeightMatrix& w = e->in_emb.w;
                    E.weights = (const uint8_t*)w.values.ptr(); E.scales = (const __nv_bfloat16*)w.scales.ptr();
                    E.zero_points = (const uint8_t*)w.zero_points.ptr(); E.biases = (const __nv_bfloat16*)w.biases.ptr();
                    E.full_precision = w.prologue == UZU_B_FULL_PRECISION; E.mode = w.mode; E.group_size = w.group_size ? w.group_size : 64;
                    E.method = w.prologue == UZU_B_SCALE_BIAS_DEQUANT ? UZU_QMETHOD_SCALE_BIAS
                               : w.prologue == UZU_B_SCALE_ZERO_POINT_DEQUANT ? UZU_QMETHOD_SCALE_ZERO_POINT : UZU_QMETHOD_SCALE_SYMMETRIC;
                    E.vocab = e->vocab; E.input_scale = e->input_scale;
                    need(E.full_precision || (E.group_size % 8) == 0, "embedding group size");
                } else {
                    op.src_kind = MK_SRC_PIECES;
                    op.src_pc = pieces_of(prev_down, 0);
                }
            };
            size_t mix_out;
            if (L.is_attention) {
                const AttentionLayer& A = L.attn;
                const uint32_t D = A.head_dim, Hq = A.num_heads, Hkv = A.num_groups;
                const size_t o_in = A.has_gate ? gemv({&A.gate, &A.qkv}) : gemv({&A.qkv});
                const int qkv_mat = A.has_gate ? 1 : 0;
                norm_input(o_in, L.pre_mixer, i > 0, S[cur], S[cur ^ 1]);
                mixer_source(o_in);
                cur ^= 1;
                MkOp prep{};
                prep.kind = MK_ATTN;
                prep.qkv_pc = pieces_of(o_in, qkv_mat);
                prep.queries = (__nv_bfloat16*)e->queries.ptr();
                prep.keys = (__nv_bfloat16*)St.keys; prep.values = (__nv_bfloat16*)St.values;
                prep.attn_out = (__nv_bfloat16*)e->attn_out.ptr();
                prep.num_q_heads = Hq; prep.num_kv_heads = Hkv; prep.head_dim = D;
                if (A.rope_index >= 0) {
                    prep.rope_dim = e->ropes[A.rope_index].head_dim;
                    need(prep.rope_dim <= D && (prep.rope_dim % 2) == 0, "rope dimension");
                    prep.rope_cos = (const float*)e->rope_cos[A.rope_index].ptr();
                    prep.rope_sin = (const float*)e->rope_sin[A.rope_index].ptr();
                }
                auto nq = [&](const Norm& n, const float*& sc, float& eps, float& off, uint32_t& full, uint32_t& has, uint32_t& present) {
                    present = n.present;
                    if (!n.present) return;
                    need(!n.cfg.subtract_mean, "q/k norm variant");
                    sc = (const float*)n.scales.ptr(); eps = n.cfg.epsilon; off = n.cfg.scale_offset; full = n.cfg.full_layer; has = n.cfg.has_scale;
                };
                nq(A.qnorm, prep.qnorm_scales, prep.qnorm_eps, prep.qnorm_offset, prep.qnorm_full_layer, prep.qnorm_has_scales, prep.qnorm_present);
                nq(A.knorm, prep.knorm_scales, prep.knorm_eps, prep.knorm_offset, prep.knorm_full_layer, prep.knorm_has_scales, prep.knorm_present);
                prep.attn_scale = A.has_scale ? A.scale : 1.0f / sqrtf((float)D);
                prep.attn_part = (float*)mg.attn_part.ptr();
                prep.attn_tickets = (unsigned int*)mg.attn_tickets.ptr();
                prep.kind = MK_ATTN;              // q/k norm + RoPE + KV append are folded into the attention phase
                mg.ops.push_back(prep);
                mix_out = gemv({&A.out});
                MkOp& oo = mg.ops[mix_out];
                oo.src_kind = MK_SRC_BF16;
                oo.src_vec = (const __nv_bfloat16*)e->attn_out.ptr();
                if (A.has_gate) { oo.in_kind = MK_IN_SIGMOID; oo.gate_pc = pieces_of(o_in, 0); }
                else oo.in_kind = MK_IN_PLAIN;
            } else {
                const DeltaNetLayer& Dn = L.dn;
                const size_t o_in = gemv({&Dn.in_proj});
                norm_input(o_in, L.pre_mixer, i > 0, S[cur], S[cur ^ 1]);
                mixer_source(o_in);
                cur ^= 1;
                MkOp dc{};
                dc.kind = MK_DN_UPDATE;
                dc.dn_in_pc = pieces_of(o_in, 0);
                dc.dn_conv_weight = (const float*)Dn.conv_weight.ptr();
                dc.dn_conv_bias = Dn.conv_has_bias ? (const float*)Dn.conv_bias.ptr() : nullptr;
                dc.dn_conv_state = (float*)St.conv_state.ptr();
                dc.dn_a_log = (const float*)Dn.a_log.ptr(); dc.dn_dt_bias = (const float*)Dn.dt_bias.ptr();
                dc.dn_state = (float*)St.ssm_state.ptr();
                dc.dn_out_raw = (float*)mg.dn_raw.ptr();
                dc.dn_kernel_size = Dn.kernel_size; dc.dn_key_dim = Dn.key_dim; dc.dn_value_dim = Dn.value_dim;
                dc.dn_num_k_heads = Dn.num_groups; dc.dn_num_v_heads = Dn.num_heads; dc.dn_hv_dim = Dn.value_head_dim;
                need(Dn.key_dim == Dn.num_groups * 128 && Dn.value_dim == Dn.num_heads * Dn.value_head_dim && Dn.conv_dim == 2 * Dn.key_dim + Dn.value_dim,
                     "DeltaNet geometry");
                need(Dn.total_proj_dim == Dn.conv_dim + Dn.value_dim + 2 * Dn.num_heads, "DeltaNet projection layout");
                {
                    const uint32_t cph = probe.grid / Dn.num_heads;
                    need(cph >= 1, "more DeltaNet heads than SMs");
                    const uint32_t rows_per = (Dn.value_head_dim + cph - 1) / cph;
                    need(probe.ncw >= 8 && probe.ncw * 32 >= 256 + rows_per && rows_per <= 256, "DeltaNet rows per CTA");
                }
                mg.ops.push_back(dc);
                mix_out = gemv({&Dn.out_proj});
                MkOp& oo = mg.ops[mix_out];
                // the staging of this phase also commits the rolling conv state (all readers of the old state are behind the barrier)
                oo.dn_commit = 1;
                oo.dn_in_pc = dc.dn_in_pc; oo.dn_conv_state = dc.dn_conv_state;
                oo.dn_kernel_size = dc.dn_kernel_size; oo.dn_key_dim = dc.dn_key_dim; oo.dn_value_dim = dc.dn_value_dim;
                oo.in_kind = MK_IN_DELTA;
                oo.src_kind = MK_SRC_BF16;       // unused by MK_IN_DELTA (the row comes from dn_raw)
                oo.src_vec = (const __nv_bfloat16*)e->delta_out.ptr();
                oo.dn_raw = (const float*)mg.dn_raw.ptr();
                oo.dn_norm_weight = (const float*)Dn.norm_weight.ptr();
                oo.dn_z_pc = pieces_of(o_in, 0);
                oo.dn_z_row0 = Dn.conv_dim;
                oo.dn_heads = Dn.num_heads; oo.dn_head_v_dim = Dn.value_head_dim; oo.dn_eps = Dn.norm_epsilon;
                need(Dn.out_proj.in_dim == Dn.value_dim, "DeltaNet out projection");
            }
            // pre_mlp_norm + DenseMlp (mlp/dense.rs:32-48)
            const size_t o_up = gemv({&L.up});
            norm_input(o_up, L.pre_mlp, true, S[cur], S[cur ^ 1]);
            mg.ops[o_up].src_kind = MK_SRC_PIECES;
            mg.ops[o_up].src_pc = pieces_of(mix_out, 0);
            cur ^= 1;
            // GatedActMul: small hidden dims fold it into the down projection's staging (every CTA recomputes the F activations, one
            // barrier less); large ones keep a separate elementwise phase spread over all SMs
            static const uint32_t fold_max = [] { const char* v = getenv("UZU_MEGA_ACT_FOLD_MAX"); return v ? (uint32_t)atoi(v) : 6144u; }();
            const bool fold_act = L.hidden_dim <= fold_max;
            if (!fold_act) {
                MkOp act{};
                act.kind = MK_ACT;
                act.up_pc = pieces_of(o_up, 0);
                act.hidden = (__nv_bfloat16*)e->gated.ptr();
                act.act_dim = L.hidden_dim; act.act_type = L.act;
                mg.ops.push_back(act);
            }
            const size_t o_down = gemv({&L.down});
            if (fold_act) {
                mg.ops[o_down].in_kind = MK_IN_GATED;
                mg.ops[o_down].gated_pc = pieces_of(o_up, 0);
                mg.ops[o_down].gated_act = L.act;
                need(L.down.in_dim == L.hidden_dim, "down projection input");
            } else {
                mg.ops[o_down].in_kind = MK_IN_PLAIN;
            }
            mg.ops[o_down].src_kind = MK_SRC_BF16;
            mg.ops[o_down].src_vec = (const __nv_bfloat16*)e->gated.ptr();
            prev_down = o_down;
            (void)H;
        }
        // output norm (residual add, transformer.rs:317-323) + readout + greedy argmax
        const size_t o_out = gemv({&e->out_emb});
        norm_input(o_out, e->out_norm, true, S[cur], S[cur ^ 1]);
        mg.ops[o_out].src_kind = MK_SRC_PIECES;
        mg.ops[o_out].src_pc = pieces_of(prev_down, 0);
        MkOp lg{};
        lg.kind = MK_LOGITS;
        lg.logits_pc = pieces_of(o_out, 0);
        lg.logits = (__nv_bfloat16*)e->logits.ptr();
        lg.vocab = e->vocab;
        lg.argmax_keys = (unsigned long long*)mg.argmax_keys.ptr();
        mg.ops.push_back(lg);
        MkOp fin = lg;
        fin.kind = MK_FINISH;
        mg.ops.push_back(fin);

        need(mega_config(e->ctx, npg, bits, (uint32_t)scratch, &mg.cfg), "shared memory budget");
        {
            std::vector<MkStream> st;
            uint32_t prev_kind = 0;
            for (auto& op : mg.ops) {
                const uint32_t before = prev_kind;
                prev_kind = op.kind;
                if (op.kind != MK_GEMV) continue;
                MkStream d{};
                // parking the prefetch cursor in front of latency-critical phases was measured (round 2): the KV loads did not get faster,
                // the phase after them lost its prefetch -> off unless UZU_MEGA_HOLD=1
                static const bool hold_on = [] { const char* v = getenv("UZU_MEGA_HOLD"); return v && atoi(v) != 0; }();
                d.hold = (hold_on && (before == MK_ATTN || before == MK_DN_UPDATE)) ? 1u : 0u;
                d.stream0 = op.mat[0].stream;
                d.stream1 = op.nmat > 1 ? op.mat[1].stream : nullptr;
                d.units = op.units;
                d.split = op.nmat > 1 ? op.mat[1].unit0 : op.units;
                st.push_back(d);
            }
            mg.nstreams = (uint32_t)st.size();
            mg.streams_dev = make_buf(e, st.size() * sizeof(MkStream), UZU_BUFFER_DEVICE);
            cudaMemcpy((void*)mg.streams_dev.ptr(), st.data(), st.size() * sizeof(MkStream), cudaMemcpyHostToDevice);
        }
        for (auto& L : e->layers)
            if (!L.is_attention) {
                const uint32_t cph = mg.cfg.grid / L.dn.num_heads, rows_per = (L.dn.value_head_dim + cph - 1) / cph;
                need(mg.cfg.ncw >= 8 && mg.cfg.ncw * 32 >= 256 + rows_per, "DeltaNet rows per CTA (final configuration)");
            }
        mg.ops_dev = make_buf(e, mg.ops.size() * sizeof(MkOp), UZU_BUFFER_DEVICE);
        cudaMemcpyAsync((void*)mg.ops_dev.ptr(), mg.ops.data(), mg.ops.size() * sizeof(MkOp), cudaMemcpyHostToDevice, e->ctx->stream);
        mg.barrier = dev_zero(256);
        mg.error_flag = make_buf(e, 256, UZU_BUFFER_PINNED_HOST);
        *(volatile unsigned int*)uzu_buffer_cpu_ptr(mg.error_flag.b) = 0u;
        cudaError_t err = cudaStreamSynchronize(e->ctx->stream);
        if (err != cudaSuccess) throw std::runtime_error(std::string("decode stream repack: ") + cudaGetErrorString(err));
        mg.ok = true;
    }
};

}  // namespace

static void build_mega(uzu_engine* e) {
    MegaBuilder b(e);
    try {
        b.build();
    } catch (const std::exception& ex) {
        e->mega.ok = false;
        e->mega.why = ex.what();
    }
}

static bool mega_usable(const uzu_engine* e) {
    return e->mega.ok && e->sampling.kind == UZU_SAMPLING_GREEDY;
}

static void issue_decode_step(uzu_engine* e, uint64_t dev_out, uint32_t dev_out_base_step);
static void upload_decode_state(uzu_engine* e);

// Both decode paths are parity-tested; which one is faster depends on the model (phase count vs bytes per phase). Unless UZU_DECODE_PATH
// forces one ("persistent" | "kernels"), time a few steps of each at a representative context right after load and keep the faster.
static void calibrate_decode_path(uzu_engine* e) {
    if (!e->mega.ok) return;
    const char* env = getenv("UZU_DECODE_PATH");
    if (env && !strcmp(env, "persistent")) { e->mega.why = "forced by UZU_DECODE_PATH=persistent"; return; }
    if (env && !strcmp(env, "kernels")) { e->mega.ok = false; e->mega.why = "forced by UZU_DECODE_PATH=kernels"; return; }
    const uint32_t P = std::min<uint32_t>(e->max_context / 2, 2048);
    cudaStream_t s = e->ctx->stream;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float ms[2] = {0.0f, 0.0f};
    for (int mode = 0; mode < 2; ++mode) {
        e->mega.ok = mode == 0;
        reset_state(e);
        for (auto& S : e->state) S.length = P;       // pseudo-context: the KV / state contents do not matter for the timing
        e->context_length = P;
        upload_decode_state(e);
        for (int i = 0; i < 3; ++i) issue_decode_step(e, 0, 0);
        cudaEventRecord(a, s);
        for (int i = 0; i < 8; ++i) issue_decode_step(e, 0, 0);
        cudaEventRecord(b, s);
        if (cudaEventSynchronize(b) != cudaSuccess) { cudaGetLastError(); ms[mode] = 1e30f; continue; }
        cudaEventElapsedTime(&ms[mode], a, b);
        e->steps_returned = e->steps_issued;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    const unsigned int code = *(volatile unsigned int*)uzu_buffer_cpu_ptr(e->mega.error_flag.b);
    e->mega.ok = code == 0 && ms[0] <= ms[1];
    char buf[160];
    snprintf(buf, sizeof buf, "auto-selected at load: persistent kernel %.3f ms/step vs per-kernel path %.3f ms/step at context %u", ms[0] / 8, ms[1] / 8, P);
    e->mega.why = buf;
    reset_state(e);
    e->launches = 0;
}

static void mega_check_error(uzu_engine* e) {
    if (!e->mega.ok) return;
    const unsigned int code = *(volatile unsigned int*)uzu_buffer_cpu_ptr(e->mega.error_flag.b);
    if (code) {
        e->mega.ok = false;
        char buf[96];
        snprintf(buf, sizeof buf, "persistent decode kernel timed out (code 0x%x); falling back to the per-kernel path", code);
        throw std::runtime_error(buf);
    }
}

static void launch_mega_step(uzu_engine* e, uint64_t dev_out, uint32_t dev_out_base_step, unsigned long long* trace = nullptr, uint32_t trace_cta = 0) {
    MkParams p{};
    p.trace = trace;
    p.trace_cta = trace_cta;
    p.ops = (const MkOp*)e->mega.ops_dev.ptr();
    p.streams = (const MkStream*)e->mega.streams_dev.ptr();
    p.nstreams = e->mega.nstreams;
    p.nops = (uint32_t)e->mega.ops.size();
    p.ncw = e->mega.cfg.ncw;
    p.state = (MkStepState*)e->decode_state.ptr();
    p.barrier = (unsigned long long*)e->mega.barrier.ptr();
    p.barrier_base = e->mega.barrier_base;
    e->mega.barrier_base += (uint64_t)(e->mega.ops.size() - 1) * e->mega.cfg.grid;     // every phase but the last ends with a grid barrier
    p.error_flag = (unsigned int*)e->mega.error_flag.ptr();
    p.token_ids = (const uint32_t*)e->token_ids.ptr();
    p.token_out = (uint32_t*)e->token_ids.ptr();
    p.sampled = (uint32_t*)e->sampled.ptr();
    p.host_ring = (volatile uint32_t*)e->host_ring.ptr();
    p.dev_out = (uint32_t*)dev_out;
    p.dev_out_base_step = dev_out_base_step;
    p.token_ring = TOKEN_RING;
    p.scratch_bytes = e->mega.cfg.scratch_bytes;
    p.stages = e->mega.cfg.stages;
    if (const char* err = mega_launch(e->ctx, e->mega.cfg, p)) throw std::runtime_error(std::string("decode_mega launch: ") + err);
    e->launches += 1;
}

static void encode_sampling(uzu_engine* e, uzu_command_buffer* cmd, uint32_t rows) {
    const uzu_sampling_method& s = e->sampling;
    uzu_unified_sampling_args sa{};
    sa.logits = e->logits.ptr();
    sa.output = e->sampled.ptr();
    sa.vocab_size = e->vocab;
    sa.batch_size = rows;
    if (s.kind == UZU_SAMPLING_STOCHASTIC) {
        sa.is_stochastic = 1;
        sa.seeds = e->seeds.ptr();
        sa.has_temperature = s.has_temperature; sa.temperature = s.temperature;
        sa.has_top_k = s.has_top_k; sa.top_k = s.top_k;
        sa.has_top_p = s.has_top_p; sa.top_p = s.top_p;
        sa.has_min_p = s.has_min_p; sa.min_p = s.min_p;
    }
    uzu_unified_sampling_encode(cmd, &sa);
}

static bool has_delta(const uzu_engine* e) {
    for (auto& L : e->layers)
        if (!L.is_attention) return true;
    return false;
}

// advance host-side bookkeeping after a pass over m tokens (TransformerState::encode_accept: flat, no copies)
static void accept(uzu_engine* e, uint32_t m) {
    for (size_t i = 0; i < e->layers.size(); ++i)
        if (e->layers[i].is_attention) e->state[i].length += m;
    e->context_length += m;
}

struct CmdGuard {
    uzu_command_buffer* c = nullptr;
    explicit CmdGuard(uzu_context* ctx, const char* name) { check(uzu_command_buffer_create(ctx, name, &c)); }
    ~CmdGuard() { uzu_command_buffer_destroy(c); }
};

static void run_cmd_to_completion(uzu_engine* e, uzu_command_buffer* c) {
    check(uzu_command_buffer_end_encoding(c));
    check(uzu_command_buffer_submit(c));
    check(uzu_command_buffer_wait_until_completed(c));
    e->launches += uzu_command_buffer_launch_count(c);
}

static uint64_t prng_derive(uint64_t seed, uint64_t index) {
    uint64_t h = seed + index;
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    return h;
}

// one non-graph pass over `count` host tokens; optionally samples the last row
static void run_pass(uzu_engine* e, const uint32_t* tokens, uint32_t count, uint32_t row_begin, uint32_t row_end, bool sample_last) {
    if (count == 0 || count > MAX_ROWS) throw std::runtime_error("pass size must be in 1..1024");
    if (e->trie_pending) throw std::runtime_error("a speculation pass is pending: uzu_engine_trie_accept first");
    if (e->context_length + count > e->max_context + MAX_ROWS || e->context_length + count > e->rope_positions)
        throw std::runtime_error("context overflow: raise max_context_length");
    state_prepare(e, e->context_length + count);
    memcpy(uzu_buffer_cpu_ptr(e->host_tokens.b), tokens, count * 4);
    CmdGuard g(e->ctx, "pass");
    check(uzu_command_buffer_start_encoding(g.c));
    uzu_command_buffer_encode_copy(g.c, e->host_tokens.ptr(), e->token_ids.ptr(), count * 4);
    PassCtx pc;
    pc.m = count;
    encode_decoder(e, g.c, pc, row_begin, row_end);
    if (sample_last) {
        if (e->sampling.kind == UZU_SAMPLING_STOCHASTIC) {
            // seed of the sampled row = PRng::derive(absolute index of its input token) (stream.rs:252-254)
            uint64_t seed = prng_derive(e->sampling.seed, (uint64_t)e->context_length + count - 1);
            cudaMemcpyAsync((void*)e->seeds.ptr(), &seed, 8, cudaMemcpyHostToDevice, e->ctx->stream);
            cudaStreamSynchronize(e->ctx->stream);
        }
        encode_sampling(e, g.c, 1);
    }
    run_cmd_to_completion(e, g.c);
    accept(e, count);
}

static uint32_t attention_bucket(uint32_t seq) {
    uint32_t b = 64;
    while (b < seq) b <<= 1;
    return b;
}

// Capture one decode step (m = 1, positions from DecodeState) into a CUDA graph.
static void capture_decode_graph(uzu_engine* e) {
    if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
    cudaStream_t s = e->ctx->stream;
    CmdGuard g(e->ctx, "decode-graph");
    cudaStreamSynchronize(s);
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) throw std::runtime_error("cudaStreamBeginCapture failed");
    g.c->state = uzu_command_buffer::Encoding;   // events are not recorded inside a capture
    g.c->use_pdl = e->use_pdl;
    if (e->sampling.kind == UZU_SAMPLING_STOCHASTIC) {
        decode_step_begin_kernel<<<1, 1, 0, s>>>((const DecodeState*)e->decode_state.ptr(), (unsigned long long*)e->seeds.ptr());
        g.c->launches++;
    }
    PassCtx pc;
    pc.m = 1;
    pc.dynamic = true;
    if (e->fused_ok) encode_decoder_fused(e, g.c, pc);
    else encode_decoder(e, g.c, pc, 0, 1);
    encode_sampling(e, g.c, 1);
    decode_step_end_kernel<<<1, 1, 0, s>>>((DecodeState*)e->decode_state.ptr(), (const uint32_t*)e->sampled.ptr(), (uint32_t*)e->token_ids.ptr(),
                                           (volatile uint32_t*)e->host_ring.ptr(), nullptr, 0);
    g.c->launches++;
    cudaGraph_t graph = nullptr;
    cudaError_t err = cudaStreamEndCapture(s, &graph);
    if (err != cudaSuccess || g.c->sticky != UZU_OK) {
        if (graph) cudaGraphDestroy(graph);
        throw std::runtime_error(std::string("decode graph capture failed: ") + (g.c->sticky != UZU_OK ? g.c->sticky_msg : cudaGetErrorString(err)));
    }
    err = cudaGraphInstantiate(&e->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (err != cudaSuccess) throw std::runtime_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(err));
    e->graph_launches_per_step = g.c->launches;
    e->graph_bucket = attention_bucket(e->context_length + 1);
    e->graph_stochastic = e->sampling.kind == UZU_SAMPLING_STOCHASTIC;
    e->graph_sampling = e->sampling;
}

// The captured graph passes temperature / top-k / top-p / min-p by value: it is only valid for the method it was captured with.
static bool graph_sampling_matches(const uzu_engine* e) {
    const uzu_sampling_method &a = e->graph_sampling, &b = e->sampling;
    return a.kind == b.kind && a.has_temperature == b.has_temperature && (!a.has_temperature || a.temperature == b.temperature) &&
           a.has_top_k == b.has_top_k && (!a.has_top_k || a.top_k == b.top_k) && a.has_top_p == b.has_top_p && (!a.has_top_p || a.top_p == b.top_p) &&
           a.has_min_p == b.has_min_p && (!a.has_min_p || a.min_p == b.min_p);
}

// enqueue one decode step (no host wait)
static void issue_decode_step(uzu_engine* e, uint64_t dev_out, uint32_t dev_out_base_step) {
    if (e->trie_pending) throw std::runtime_error("a speculation pass is pending: uzu_engine_trie_accept first");
    if (e->context_length + 1 > e->max_context + MAX_ROWS || e->context_length + 1 > e->rope_positions)
        throw std::runtime_error("context overflow: raise max_context_length");
    state_prepare(e, e->context_length + 1);
    cudaStream_t s = e->ctx->stream;
    if (mega_usable(e)) {
        launch_mega_step(e, dev_out, dev_out_base_step);
    } else if (e->opts.use_cuda_graph && !dev_out) {
        if (!e->graph_exec || e->graph_bucket != attention_bucket(e->context_length + 1) || !graph_sampling_matches(e))
            capture_decode_graph(e);
        cudaError_t err = cudaGraphLaunch(e->graph_exec, s);
        if (err != cudaSuccess) throw std::runtime_error(std::string("cudaGraphLaunch: ") + cudaGetErrorString(err));
        e->launches += e->graph_launches_per_step;
    } else {
        CmdGuard g(e->ctx, "decode");
        g.c->state = uzu_command_buffer::Encoding;
        g.c->use_pdl = e->use_pdl;
        if (e->sampling.kind == UZU_SAMPLING_STOCHASTIC) {
            decode_step_begin_kernel<<<1, 1, 0, s>>>((const DecodeState*)e->decode_state.ptr(), (unsigned long long*)e->seeds.ptr());
            g.c->launches++;
        }
        PassCtx pc;
        pc.m = 1;
        pc.dynamic = true;
        if (e->fused_ok) encode_decoder_fused(e, g.c, pc);
        else encode_decoder(e, g.c, pc, 0, 1);
        encode_sampling(e, g.c, 1);
        decode_step_end_kernel<<<1, 1, 0, s>>>((DecodeState*)e->decode_state.ptr(), (const uint32_t*)e->sampled.ptr(), (uint32_t*)e->token_ids.ptr(),
                                               (volatile uint32_t*)e->host_ring.ptr(), (uint32_t*)dev_out, dev_out_base_step);
        g.c->launches++;
        if (g.c->sticky != UZU_OK) throw std::runtime_error(g.c->sticky_msg);
        e->launches += g.c->launches;
    }
    cudaEventRecord(e->step_events[e->steps_issued & 1], s);
    e->steps_issued++;
    accept(e, 1);
}

static void upload_decode_state(uzu_engine* e) {
    DecodeState st{};
    st.position = e->context_length;
    st.step = e->steps_issued;
    st.base_seed = e->sampling.seed;
    cudaMemcpyAsync((void*)e->decode_state.ptr(), &st, sizeof st, cudaMemcpyHostToDevice, e->ctx->stream);
    cudaStreamSynchronize(e->ctx->stream);
}

}  // namespace uzu

#define UZU_ENGINE_TRY(...)                                            \
    if (!e) return uzu::fail(UZU_ERROR_INVALID_ARGUMENT, "null engine"); \
    try {                                                              \
        cudaSetDevice(e->ctx->device);                                 \
        __VA_ARGS__;                                                   \
        return UZU_OK;                                                 \
    } catch (const std::exception& ex) {                               \
        return uzu::fail(UZU_ERROR_INVALID_ARGUMENT, ex.what());       \
    }

extern "C" {

uzu_status uzu_engine_create(uzu_context* ctx, const char* model_dir, const uzu_engine_options* opts, uzu_engine** out) {
    if (!ctx || !model_dir || !out) return fail(UZU_ERROR_INVALID_ARGUMENT, "uzu_engine_create: null argument");
    auto* e = new uzu_engine();
    e->ctx = ctx;
    if (opts) e->opts = *opts;
    if (e->opts.tp_size == 0) e->opts.tp_size = 1;
    e->use_pdl = getenv("UZU_NO_PDL") == nullptr;
    try {
        cudaSetDevice(ctx->device);
        load_model(e, model_dir);
        create_state_and_scratch(e);
        reset_state(e);
        e->fused_ok = e->opts.fused_decode && getenv("UZU_NO_FUSED") == nullptr && fused_decode_supported(e);
        if (e->opts.fused_decode) {                   // the persistent decode kernel is the fused path's next step: one launch per token
            build_mega(e);
            calibrate_decode_path(e);
        }
        check(uzu_context_synchronize(ctx));
    } catch (const std::exception& ex) {
        std::string msg = ex.what();
        uzu_engine_destroy(e);
        return fail(UZU_ERROR_IO, std::string("uzu_engine_create: ") + msg);
    }
    *out = e;
    return UZU_OK;
}

void uzu_engine_destroy(uzu_engine* e) {
    if (!e) return;
    cudaSetDevice(e->ctx->device);
    cudaStreamSynchronize(e->ctx->stream);
    if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
    if (e->batch_graph) cudaGraphExecDestroy(e->batch_graph);
    for (auto& S : e->state) {
        if (S.k_sparse) uzu_sparse_buffer_destroy(S.k_sparse);
        if (S.v_sparse) uzu_sparse_buffer_destroy(S.v_sparse);
    }
    for (auto& q : e->seqs)
        for (auto& S : q.layers) {
            if (S.k_sparse) uzu_sparse_buffer_destroy(S.k_sparse);
            if (S.v_sparse) uzu_sparse_buffer_destroy(S.v_sparse);
        }
    for (auto* b : e->owned) uzu_buffer_destroy(b);
    if (e->step_events[0]) cudaEventDestroy(e->step_events[0]);
    if (e->step_events[1]) cudaEventDestroy(e->step_events[1]);
    delete e;
}

uzu_status uzu_engine_info(const uzu_engine* e, uzu_model_info* out) {
    if (!e || !out) return fail(UZU_ERROR_INVALID_ARGUMENT, "uzu_engine_info: null argument");
    *out = e->info;
    return UZU_OK;
}

uzu_status uzu_engine_reset(uzu_engine* e) { UZU_ENGINE_TRY(reset_state(e)); }

uint32_t uzu_engine_context_length(const uzu_engine* e) { return e ? e->context_length : 0; }

uzu_status uzu_engine_snapshot(uzu_engine* e) {
    UZU_ENGINE_TRY({
        cudaStream_t s = e->ctx->stream;
        for (auto& S : e->state)
            if (S.conv_state.b) {
                cudaMemcpyAsync((void*)S.conv_snapshot.ptr(), (void*)S.conv_state.ptr(), S.conv_bytes, cudaMemcpyDeviceToDevice, s);
                cudaMemcpyAsync((void*)S.ssm_snapshot.ptr(), (void*)S.ssm_state.ptr(), S.ssm_bytes, cudaMemcpyDeviceToDevice, s);
            }
        cudaMemcpyAsync((void*)e->snapshot_token.ptr(), (void*)e->token_ids.ptr(), 4, cudaMemcpyDeviceToDevice, s);
        cudaStreamSynchronize(s);
        e->snapshot_context = e->context_length;
    });
}

uzu_status uzu_engine_restore(uzu_engine* e) {
    UZU_ENGINE_TRY({
        cudaStream_t s = e->ctx->stream;
        cudaStreamSynchronize(s);
        for (size_t i = 0; i < e->state.size(); ++i) {
            auto& S = e->state[i];
            if (S.conv_state.b) {
                cudaMemcpyAsync((void*)S.conv_state.ptr(), (void*)S.conv_snapshot.ptr(), S.conv_bytes, cudaMemcpyDeviceToDevice, s);
                cudaMemcpyAsync((void*)S.ssm_state.ptr(), (void*)S.ssm_snapshot.ptr(), S.ssm_bytes, cudaMemcpyDeviceToDevice, s);
            } else {
                S.length = e->snapshot_context;   // KV rows beyond the snapshot are simply overwritten
            }
        }
        cudaMemcpyAsync((void*)e->token_ids.ptr(), (void*)e->snapshot_token.ptr(), 4, cudaMemcpyDeviceToDevice, s);
        e->context_length = e->snapshot_context;
        e->steps_returned = e->steps_issued;
        e->trie_pending = 0;
        cudaStreamSynchronize(s);
        upload_decode_state(e);
    });
}

uzu_status uzu_engine_prefill(uzu_engine* e, const uint32_t* tokens, uint32_t count, const uzu_sampling_method* sampling, uint32_t* out_token) {
    UZU_ENGINE_TRY({
        if (!tokens || count == 0) throw std::runtime_error("prefill: empty prompt");
        if (sampling) e->sampling = *sampling;
        // chunks of <= 1024 (stream.rs:194-195); only the last chunk's last row is sampled. Hybrid (DeltaNet) models take the same
        // chunks: projections and MLPs batched, the DeltaNet recurrence row by row inside the pass (encode_delta_net).
        // UZU_DELTA_PREFILL_STEPPED=1 restores one pass per token (A/B runs).
        static const bool stepped = getenv("UZU_DELTA_PREFILL_STEPPED") != nullptr;
        const uint32_t step = (stepped && has_delta(e)) ? 1 : MAX_ROWS;
        for (uint32_t s0 = 0; s0 < count; s0 += step) {
            const uint32_t n = std::min(step, count - s0);
            const bool last = s0 + n == count;
            run_pass(e, tokens + s0, n, last ? n - 1 : 0, last ? n : 0, last);
        }
        uint32_t tok = 0;
        cudaMemcpy(&tok, (void*)e->sampled.ptr(), 4, cudaMemcpyDeviceToHost);
        // the sampled token becomes the next input, chained on the device
        cudaMemcpy((void*)e->token_ids.ptr(), (void*)e->sampled.ptr(), 4, cudaMemcpyDeviceToDevice);
        e->steps_issued = e->steps_returned = 0;
        upload_decode_state(e);
        if (out_token) *out_token = tok;
    });
}

uzu_status uzu_engine_next(uzu_engine* e, uint32_t* out_token) {
    UZU_ENGINE_TRY({
        // keep one pass in flight: issue step N+1, then wait for step N (ForwardPassChaining::InFlight)
        issue_decode_step(e, 0, 0);
        if (e->steps_issued - e->steps_returned >= 2) {
            const uint32_t idx = e->steps_returned;
            cudaError_t err = cudaEventSynchronize(e->step_events[idx & 1]);
            if (err != cudaSuccess) throw std::runtime_error(std::string("decode step failed: ") + cudaGetErrorString(err));
            mega_check_error(e);
            if (out_token) *out_token = ((volatile uint32_t*)uzu_buffer_cpu_ptr(e->host_ring.b))[idx % TOKEN_RING];
            e->steps_returned++;
        } else if (out_token) {
            *out_token = 0xFFFFFFFFu;   // nothing resolved yet (first call after prefill)
        }
    });
}

uzu_status uzu_engine_flush(uzu_engine* e, uint32_t* out_token) {
    UZU_ENGINE_TRY({
        if (e->steps_returned < e->steps_issued) {
            const uint32_t idx = e->steps_returned;
            cudaError_t err = cudaEventSynchronize(e->step_events[idx & 1]);
            if (err != cudaSuccess) throw std::runtime_error(std::string("decode step failed: ") + cudaGetErrorString(err));
            mega_check_error(e);
            if (out_token) *out_token = ((volatile uint32_t*)uzu_buffer_cpu_ptr(e->host_ring.b))[idx % TOKEN_RING];
            e->steps_returned++;
        } else if (out_token) {
            *out_token = 0xFFFFFFFFu;
        }
    });
}

uzu_status uzu_engine_decode_device(uzu_engine* e, uint32_t steps, uint64_t out_tokens_dev) {
    UZU_ENGINE_TRY({
        const uint32_t base = e->steps_issued;
        for (uint32_t i = 0; i < steps; ++i) issue_decode_step(e, out_tokens_dev, base);
        e->steps_returned = e->steps_issued;   // device-resident run: tokens are not resolved on the host
    });
}

uzu_status uzu_engine_forward(uzu_engine* e, const uint32_t* tokens, uint32_t count, uint32_t row_begin, uint32_t row_end, uint16_t* out_logits) {
    UZU_ENGINE_TRY({
        if (row_end > count || row_begin > row_end || row_end - row_begin > e->logits_rows) throw std::runtime_error("forward: bad row range (<= 16 rows)");
        run_pass(e, tokens, count, row_begin, row_end, false);
        if (out_logits && row_end > row_begin)
            cudaMemcpy(out_logits, (void*)e->logits.ptr(), (size_t)(row_end - row_begin) * e->vocab * 2, cudaMemcpyDeviceToHost);
    });
}


// ---- speculation pass over a trie: the verify half of the stream's speculative decode (stream.rs:550-657) ----
// The proposer (a draft model / n-gram lookup) and the trie bookkeeping (trie.rs: linearize, accept) stay on the host side of the boundary;
// the backend runs Decoder::encode over the linearized nodes (positions = context + height, attention masked by subtrie range, one sampled
// token per node) and later compacts the KV rows of the accepted path (TransformerState::encode_accept).
int uzu_engine_speculation_supported(const uzu_engine* e) {
    // Mixer::speculation_supported: attention always, DeltaNet only with a tree-verify core (delta_net.rs:442-444) -- not built here
    return e && !has_delta(e) ? 1 : 0;
}

uzu_status uzu_engine_trie_pass(uzu_engine* e, const uint32_t* tokens, const uzu_trie_node* nodes, const uint64_t* seeds, uint32_t count,
                                const uzu_sampling_method* sampling, uint32_t* out_tokens, uint16_t* out_logits) {
    UZU_ENGINE_TRY({
        if (!tokens || !nodes || count == 0 || count > MAX_TRIE) throw std::runtime_error("trie_pass: 1..16 nodes");
        if (has_delta(e)) throw std::runtime_error("trie_pass: DeltaNet tree verification is not supported (speculation_supported == 0)");
        if (e->trie_pending) throw std::runtime_error("trie_pass: the previous speculation pass was not accepted");
        if (e->steps_returned != e->steps_issued) throw std::runtime_error("trie_pass: flush() the in-flight decode steps first");
        // a linearized trie: depth-first order, node i's subtree = [trie_start, trie_end] = [i, >= i], height = depth (trie.rs:158-178)
        for (uint32_t i = 0; i < count; ++i) {
            const uzu_trie_node& n = nodes[i];
            const bool ok = n.trie_start == i && n.trie_end >= i && n.trie_end < count && n.height <= i && (i == 0 ? n.height == 0 : n.height >= 1 && n.height <= nodes[i - 1].height + 1);
            if (!ok) throw std::runtime_error("trie_pass: nodes are not a linearized trie");
        }
        if (e->context_length + count > e->max_context + MAX_ROWS || e->context_length + count > e->rope_positions)
            throw std::runtime_error("context overflow: raise max_context_length");
        if (sampling) e->sampling = *sampling;
        state_prepare(e, e->context_length + count);
        cudaStream_t s = e->ctx->stream;
        // pinned staging: tokens | nodes | seeds | sampled
        uint8_t* host = (uint8_t*)uzu_buffer_cpu_ptr(e->host_trie.b);
        uint64_t* h_seeds = (uint64_t*)host;
        uzu_trie_node* h_nodes = (uzu_trie_node*)(host + MAX_TRIE * 8);
        uint32_t* h_tokens = (uint32_t*)(host + MAX_TRIE * (8 + sizeof(uzu_trie_node)));
        uint32_t* h_sampled = h_tokens + MAX_TRIE;
        memcpy(h_tokens, tokens, count * 4);
        memcpy(h_nodes, nodes, count * sizeof(uzu_trie_node));
        for (uint32_t i = 0; i < count; ++i)   // speculator seeds = derive(root position + depth) (dflash_tfm.rs:267,304)
            h_seeds[i] = seeds ? seeds[i] : prng_derive(e->sampling.seed, (uint64_t)e->context_length + nodes[i].height);
        CmdGuard g(e->ctx, "trie pass");
        check(uzu_command_buffer_start_encoding(g.c));
        uzu_command_buffer_encode_copy(g.c, e->host_trie.ptr() + MAX_TRIE * (8 + sizeof(uzu_trie_node)), e->token_ids.ptr(), count * 4);
        uzu_command_buffer_encode_copy(g.c, e->host_trie.ptr() + MAX_TRIE * 8, e->trie_nodes.ptr(), count * sizeof(uzu_trie_node));
        uzu_command_buffer_encode_copy(g.c, e->host_trie.ptr(), e->seeds.ptr(), count * 8);
        for (size_t r = 0; r < e->ropes.size(); ++r) {
            trie_rope_gather_kernel<<<count, 64, 0, s>>>((const uzu_trie_node*)e->trie_nodes.ptr(), e->context_length, e->ropes[r].head_dim,
                                                         (const float*)e->rope_cos[r].ptr(), (const float*)e->rope_sin[r].ptr(),
                                                         (float*)e->trie_cos[r].ptr(), (float*)e->trie_sin[r].ptr());
            g.c->launches++;
        }
        PassCtx pc;
        pc.m = count;
        pc.trie = e->trie_nodes.ptr();
        encode_decoder(e, g.c, pc, 0, count);        // logits for every node (stream.rs:640: Some(0..batch_dim.size()))
        encode_sampling(e, g.c, count);
        uzu_command_buffer_encode_copy(g.c, e->sampled.ptr(), e->host_trie.ptr() + MAX_TRIE * (8 + sizeof(uzu_trie_node)) + MAX_TRIE * 4, count * 4);
        run_cmd_to_completion(e, g.c);
        if (out_tokens) memcpy(out_tokens, h_sampled, count * 4);
        if (out_logits) cudaMemcpy(out_logits, (void*)e->logits.ptr(), (size_t)count * e->vocab * 2, cudaMemcpyDeviceToHost);
        e->trie_pending = count;
    });
}

uzu_status uzu_engine_trie_accept(uzu_engine* e, const uint32_t* accepted_indices, uint32_t count, uint32_t next_token) {
    UZU_ENGINE_TRY({
        if (!e->trie_pending) throw std::runtime_error("trie_accept: no speculation pass to accept");
        if (!accepted_indices || count == 0 || count > e->trie_pending) throw std::runtime_error("trie_accept: 1..pending indices");
        for (uint32_t i = 0; i < count; ++i)       // mixer/attention/state.rs:179: strictly increasing, inside the pass
            if (accepted_indices[i] >= e->trie_pending || (i > 0 && accepted_indices[i] <= accepted_indices[i - 1]))
                throw std::runtime_error("trie_accept: invalid accepted indices");
        CmdGuard g(e->ctx, "trie accept");
        check(uzu_command_buffer_start_encoding(g.c));
        for (size_t l = 0; l < e->layers.size(); ++l) {
            LayerState& S = e->state[l];
            // AttentionStateType::Full (state.rs:182-199): accepted row `length + index` moves to `length + i`; already-compact rows stay
            std::vector<uzu_kv_copy> copies;
            for (uint32_t i = 0; i < count; ++i)
                if (accepted_indices[i] != i) copies.push_back(uzu_kv_copy{S.length + accepted_indices[i], S.length + i});
            if (!copies.empty()) {
                uzu_kv_cache_update_args ka{};
                ka.in_place_keys = S.keys; ka.in_place_values = S.values;
                ka.copies = copies.data(); ka.copy_count = (uint32_t)copies.size();
                ka.element_dim = e->layers[l].attn.num_groups * e->layers[l].attn.head_dim;
                uzu_kv_cache_update_encode(g.c, &ka);
            }
        }
        // the last verified node's sampled token is the next pass's root (ForwardPassChaining::Constant, stream.rs:508-513)
        uint32_t* staging = (uint32_t*)uzu_buffer_cpu_ptr(e->host_tokens.b);
        staging[0] = next_token;
        uzu_command_buffer_encode_copy(g.c, e->host_tokens.ptr(), e->token_ids.ptr(), 4);
        run_cmd_to_completion(e, g.c);
        e->trie_pending = 0;
        accept(e, count);
        upload_decode_state(e);
    });
}

uint64_t uzu_engine_launch_count(const uzu_engine* e) { return e ? e->launches : 0; }

// ---- persistent decode kernel controls (extension) ----
int uzu_engine_decode_mode(const uzu_engine* e) { return e && e->mega.ok ? 1 : 0; }
const char* uzu_engine_decode_mode_reason(const uzu_engine* e) { return e ? e->mega.why.c_str() : ""; }
uzu_status uzu_engine_set_decode_mode(uzu_engine* e, int persistent) {
    UZU_ENGINE_TRY({
        cudaStreamSynchronize(e->ctx->stream);
        if (!persistent) { e->mega.ok = false; if (e->mega.why.empty()) e->mega.why = "disabled by uzu_engine_set_decode_mode"; }
        else if (!e->mega.ok) {
            if (e->mega.ops.empty() || !e->mega.ops_dev.b) throw std::runtime_error("persistent decode kernel not available: " + e->mega.why);
            *(volatile unsigned int*)uzu_buffer_cpu_ptr(e->mega.error_flag.b) = 0u;
            cudaMemset((void*)e->mega.barrier.ptr(), 0, 256);
            e->mega.barrier_base = 0;
            e->mega.ok = true;
        }
    });
}
uzu_status uzu_engine_debug_decode_trace(uzu_engine* e, uint32_t cta, uint32_t capacity, uint32_t* out_kinds, uint64_t* out_cycles, uint32_t* out_nops) {
    UZU_ENGINE_TRY({
        if (!mega_usable(e)) throw std::runtime_error("decode trace: the persistent decode kernel is not active");
        const uint32_t n = (uint32_t)e->mega.ops.size();
        if (out_nops) *out_nops = n;
        if (capacity < n || !out_kinds || !out_cycles) throw std::runtime_error("decode trace: capacity too small");
        if (e->context_length + 1 > e->max_context + MAX_ROWS) throw std::runtime_error("context overflow");
        state_prepare(e, e->context_length + 1);
        Buf tr = make_buf(e, (size_t)n * 8 * 8, UZU_BUFFER_DEVICE);
        cudaMemsetAsync((void*)tr.ptr(), 0, (size_t)n * 8 * 8, e->ctx->stream);
        launch_mega_step(e, 0, 0, (unsigned long long*)tr.ptr(), cta);
        cudaEventRecord(e->step_events[e->steps_issued & 1], e->ctx->stream);
        e->steps_issued++;
        accept(e, 1);
        e->steps_returned = e->steps_issued;
        cudaError_t err = cudaStreamSynchronize(e->ctx->stream);
        if (err != cudaSuccess) throw std::runtime_error(std::string("decode trace: ") + cudaGetErrorString(err));
        mega_check_error(e);
        cudaMemcpy(out_cycles, (void*)tr.ptr(), (size_t)n * 8 * 8, cudaMemcpyDeviceToHost);
        for (uint32_t i = 0; i < n; ++i) out_kinds[i] = e->mega.ops[i].kind;
    });
}
uzu_status uzu_engine_last_logits(uzu_engine* e, uint16_t* out_logits) {
    UZU_ENGINE_TRY({
        cudaError_t err = cudaStreamSynchronize(e->ctx->stream);
        if (err != cudaSuccess) throw std::runtime_error(std::string("last_logits: ") + cudaGetErrorString(err));
        mega_check_error(e);
        cudaMemcpy(out_logits, (void*)e->logits.ptr(), (size_t)e->vocab * 2, cudaMemcpyDeviceToHost);
    });
}

uzu_status uzu_engine_decode_timed(uzu_engine* e, uint32_t steps, double* out_seconds) {
    UZU_ENGINE_TRY({
        cudaStream_t s = e->ctx->stream;
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaStreamSynchronize(s);
        cudaEventRecord(a, s);
        for (uint32_t i = 0; i < steps; ++i) issue_decode_step(e, 0, 0);
        cudaEventRecord(b, s);
        cudaError_t err = cudaEventSynchronize(b);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, a, b);
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        e->steps_returned = e->steps_issued;
        if (err != cudaSuccess) throw std::runtime_error(std::string("decode_timed: ") + cudaGetErrorString(err));
        mega_check_error(e);
        if (out_seconds) *out_seconds = (double)ms * 1e-3;
    });
}

uzu_status uzu_engine_step_host(uzu_engine* e, uint32_t token_in, uint32_t* token_out) {
    UZU_ENGINE_TRY({
        cudaStream_t s = e->ctx->stream;
        uint32_t* staging = (uint32_t*)uzu_buffer_cpu_ptr(e->host_tokens.b);
        staging[0] = token_in;
        cudaMemcpyAsync((void*)e->token_ids.ptr(), staging, 4, cudaMemcpyHostToDevice, s);   // H2D: this step's input
        issue_decode_step(e, 0, 0);
        uint32_t tok = 0;
        cudaMemcpyAsync(staging + 1, (void*)e->sampled.ptr(), 4, cudaMemcpyDeviceToHost, s); // D2H: this step's result
        cudaError_t err = cudaStreamSynchronize(s);
        if (err != cudaSuccess) throw std::runtime_error(std::string("step_host: ") + cudaGetErrorString(err));
        mega_check_error(e);
        tok = staging[1];
        e->steps_returned = e->steps_issued;
        if (token_out) *token_out = tok;
    });
}

// ---- multi-sequence batched decode (extension; the reference is single-sequence: batch_dim = tokens of ONE sequence, SURVEY 7 hard part 5) ----
// B independent sequences (own KV caches / DeltaNet states / positions) advance by one token per step and share ONE pass over the weights:
// every linear runs with m = B rows (the m <= 16 GEMV streams the matrix once for all rows), norms / activations / embedding / sampling take
// B rows, only attention and the DeltaNet recurrence run per sequence (their state is per sequence). Plain stream-ordered launches (no CUDA
// graph yet). NOT yet run on hardware (written after round 1's GPU budget was spent); orchestration of parity-tested kernels only.
namespace uzu {

struct SeqSwap {   // RAII: make sequence b the engine's current state for code written against e->state / e->context_length
    uzu_engine* e; uint32_t b;
    SeqSwap(uzu_engine* e_, uint32_t b_) : e(e_), b(b_) { std::swap(e->state, e->seqs[b].layers); std::swap(e->context_length, e->seqs[b].context_length); }
    ~SeqSwap() { std::swap(e->state, e->seqs[b].layers); std::swap(e->context_length, e->seqs[b].context_length); }
};

__global__ void batch_step_end_kernel(uint32_t* pos, uint32_t B, const uint32_t* sampled, uint32_t* token_ids) {
    const uint32_t b = threadIdx.x;
    if (b < B) {
        pos[b] += 1;                 // every sequence advanced by one token
        token_ids[b] = sampled[b];   // device-side chaining of the next inputs (a host-fed step overwrites them before the next pass)
    }
}

// `dyn`: positions come from e->batch_pos on the device (CUDA-graph replay); otherwise from the host-side sequence state
static void encode_batch_step(uzu_engine* e, uzu_command_buffer* cmd, bool dyn) {
    const uint32_t B = (uint32_t)e->seqs.size(), H = e->model_dim;
    if (e->in_emb.w.prologue == UZU_B_FULL_PRECISION) {
        uzu_full_precision_embedding_lookup_encode(cmd, e->token_ids.ptr(), e->in_emb.w.values.ptr(), e->hidden_a.ptr(), B, e->vocab, H, e->input_scale);
    } else {
        uzu_quantized_embedding_lookup_args la{};
        la.token_ids = e->token_ids.ptr(); la.weights = e->in_emb.w.values.ptr(); la.scales = e->in_emb.w.scales.ptr();
        la.zero_points = e->in_emb.w.zero_points.ptr(); la.biases = e->in_emb.w.biases.ptr(); la.output = e->hidden_a.ptr();
        la.batch_size = B; la.vocab_size = e->vocab; la.model_dim = H; la.input_scale = e->input_scale;
        la.group_size = e->in_emb.w.group_size; la.quantization_mode = e->in_emb.w.mode;
        la.quantization_method = e->in_emb.w.prologue == UZU_B_SCALE_BIAS_DEQUANT ? UZU_QMETHOD_SCALE_BIAS
                                 : e->in_emb.w.prologue == UZU_B_SCALE_ZERO_POINT_DEQUANT ? UZU_QMETHOD_SCALE_ZERO_POINT : UZU_QMETHOD_SCALE_SYMMETRIC;
        uzu_quantized_embedding_lookup_encode(cmd, &la);
    }
    uint64_t hidden = e->hidden_a.ptr();
    for (size_t i = 0; i < e->layers.size(); ++i) {
        Layer& L = e->layers[i];
        encode_norm(cmd, L.pre_mixer, hidden, e->hidden_b.ptr(), e->shortcut.ptr(), i == 0 ? ShortcutCopy : ShortcutAdd, B);
        if (L.is_attention) {
            const AttentionLayer& A = L.attn;
            const uint32_t D = A.head_dim, Hq = A.num_heads, Hkv = A.num_groups, total_heads = Hq + 2 * Hkv;
            const size_t qkv_row = (size_t)total_heads * D * 2, q_row = (size_t)Hq * D * 2;
            if (A.has_gate) encode_linear(cmd, A.gate, e->hidden_b.ptr(), B, e->gate.ptr());
            encode_linear(cmd, A.qkv, e->hidden_b.ptr(), B, e->qkv.ptr());
            for (uint32_t b = 0; b < B; ++b) {     // per sequence: q/k norm, RoPE at its own position + KV append, attention over its own cache
                LayerState& S = e->seqs[b].layers[i];
                const uint32_t pos = e->seqs[b].context_length;
                const uint64_t qkv_b = e->qkv.ptr() + b * qkv_row;
                auto qkn = [&](const Norm& n, uint32_t off, uint32_t cnt) {
                    if (!n.present || cnt == 0) return;
                    uzu_qkv_norm_args qa{};
                    qa.scales = n.scales.ptr(); qa.qkv_output = qkv_b;
                    qa.batch_size = 1; qa.total_heads = total_heads; qa.head_dim = D;
                    qa.epsilon = n.cfg.epsilon; qa.scale_offset = n.cfg.scale_offset;
                    qa.head_offset = off; qa.head_count = cnt; qa.full_layer = n.cfg.full_layer;
                    qa.in_place = 1; qa.has_scales = n.cfg.has_scale;
                    uzu_qkv_norm_encode(cmd, &qa);
                };
                qkn(A.qnorm, 0, Hq);
                qkn(A.knorm, Hq, Hkv);
                uzu_attention_prepare_args pa{};
                pa.qkv = qkv_b; pa.queries = e->queries.ptr() + b * q_row;
                pa.keys = S.keys; pa.values = S.values;
                pa.num_q_heads = Hq; pa.num_kv_heads = Hkv; pa.head_dim = D;
                pa.kv_token_offset = S.length; pa.batch_dim = 1; pa.has_kv = 1;
                if (A.rope_index >= 0) {
                    const RopeCfg& rc = e->ropes[A.rope_index];
                    pa.has_rope = 1; pa.rope_dim = rc.head_dim;
                    // with a dynamic position the kernel offsets the tables by the device-resident position itself
                    pa.cosines = e->rope_cos[A.rope_index].ptr() + (dyn ? 0 : (size_t)pos * rc.head_dim * 4);
                    pa.sines = e->rope_sin[A.rope_index].ptr() + (dyn ? 0 : (size_t)pos * rc.head_dim * 4);
                }
                if (dyn) pa.dynamic_position = e->batch_pos.ptr() + (uint64_t)b * 4;
                uzu_attention_prepare_encode(cmd, &pa);
                uzu_attention_args aa{};
                aa.queries = pa.queries; aa.keys = S.keys; aa.values = S.values; aa.out = e->attn_out.ptr() + b * q_row;
                aa.gqa_factor = Hq / Hkv; aa.sequence_length = S.length + 1;
                aa.k_head_stride = D; aa.k_seq_stride = Hkv * D; aa.v_head_stride = D; aa.v_seq_stride = Hkv * D;
                aa.scale = A.has_scale ? A.scale : 1.0f / sqrtf((float)D);
                aa.num_heads = Hq; aa.suffix_length = 1; aa.head_dim = D; aa.is_causal = A.is_causal;
                if (dyn) aa.dynamic_position = e->batch_pos.ptr() + (uint64_t)b * 4;
                uzu_attention_single_pass_encode(cmd, &aa);
            }
            if (A.has_gate) uzu_sigmoid_gate_encode(cmd, e->gate.ptr(), e->attn_out.ptr(), B * Hq * D);
            encode_linear(cmd, A.out, e->attn_out.ptr(), B, e->mixer_out.ptr());
        } else {
            const DeltaNetLayer& D = L.dn;
            encode_linear(cmd, D.in_proj, e->hidden_b.ptr(), B, e->in_proj.ptr());
            for (uint32_t b = 0; b < B; ++b) {     // per sequence: its own rolling conv state and recurrent state
                LayerState& S = e->seqs[b].layers[i];
                const uint64_t row = e->in_proj.ptr() + (size_t)b * D.total_proj_dim * 2;
                uzu_delta_net_conv_update_args ca{};
                ca.conv_weight = D.conv_weight.ptr(); ca.bias = D.conv_bias.ptr(); ca.in_out = row; ca.state = S.conv_state.ptr();
                ca.kernel_size = D.kernel_size; ca.conv_dim = D.conv_dim; ca.state_stride = D.kernel_size - 1; ca.has_bias = D.conv_has_bias;
                uzu_delta_net_conv_update_encode(cmd, &ca);
                uzu_delta_net_update_args ua{};
                ua.in_proj = row; ua.a_log = D.a_log.ptr(); ua.dt_bias = D.dt_bias.ptr(); ua.norm_weight = D.norm_weight.ptr();
                ua.state = S.ssm_state.ptr(); ua.out = e->delta_out.ptr() + (size_t)b * D.value_dim * 2;
                ua.num_v_heads = D.num_heads; ua.num_k_heads = D.num_groups; ua.head_v_dim = D.value_head_dim; ua.key_dim = D.key_dim;
                ua.value_dim = D.value_dim; ua.norm_epsilon = D.norm_epsilon; ua.head_k_dim = D.head_dim;
                uzu_delta_net_update_encode(cmd, &ua);
            }
            encode_linear(cmd, D.out_proj, e->delta_out.ptr(), B, e->mixer_out.ptr());
        }
        encode_norm(cmd, L.pre_mlp, e->mixer_out.ptr(), e->hidden_b.ptr(), e->shortcut.ptr(), ShortcutAdd, B);
        encode_linear(cmd, L.up, e->hidden_b.ptr(), B, e->fused_up.ptr());
        uzu_gated_act_mul_args ga{};
        ga.act_operand = e->fused_up.ptr(); ga.fp_out = e->gated.ptr(); ga.gated_dim = L.hidden_dim; ga.batch_dim = B;
        ga.act_type = L.act; ga.interleaved = 1;
        uzu_gated_act_mul_encode(cmd, &ga);
        encode_linear(cmd, L.down, e->gated.ptr(), B, e->hidden_a.ptr());
        hidden = e->hidden_a.ptr();
    }
    encode_norm(cmd, e->out_norm, hidden, e->normed_out.ptr(), e->shortcut.ptr(), ShortcutAdd, B);
    encode_linear(cmd, e->out_emb, e->normed_out.ptr(), B, e->logits.ptr());
    if (e->has_logit_scale || e->has_logit_soft_cap)
        uzu_logit_transform_encode(cmd, e->logits.ptr(), B * e->vocab, e->has_logit_scale ? e->logit_scale : 1.0f, e->logit_soft_cap, e->has_logit_soft_cap);
    encode_sampling(e, cmd, B);
}

// Capture one batched step (positions from e->batch_pos) into a CUDA graph: ~740 launches per step for Llama-3-8B x 8 sequences, which
// the host cannot issue eagerly faster than ~22 ms; replayed as a graph the step is bound by the GPU again.
static void capture_batch_graph(uzu_engine* e, uint32_t bucket) {
    if (e->batch_graph) { cudaGraphExecDestroy(e->batch_graph); e->batch_graph = nullptr; }
    const uint32_t B = (uint32_t)e->seqs.size();
    cudaStream_t s = e->ctx->stream;
    CmdGuard g(e->ctx, "batch-graph");
    cudaStreamSynchronize(s);
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) throw std::runtime_error("cudaStreamBeginCapture failed");
    g.c->state = uzu_command_buffer::Encoding;
    g.c->use_pdl = e->use_pdl;
    encode_batch_step(e, g.c, true);
    batch_step_end_kernel<<<1, 32, 0, s>>>((uint32_t*)e->batch_pos.ptr(), B, (const uint32_t*)e->sampled.ptr(), (uint32_t*)e->token_ids.ptr());
    g.c->launches++;
    cudaGraph_t graph = nullptr;
    cudaError_t err = cudaStreamEndCapture(s, &graph);
    if (err != cudaSuccess || g.c->sticky != UZU_OK) {
        if (graph) cudaGraphDestroy(graph);
        throw std::runtime_error(std::string("batch graph capture failed: ") + (g.c->sticky != UZU_OK ? g.c->sticky_msg : cudaGetErrorString(err)));
    }
    err = cudaGraphInstantiate(&e->batch_graph, graph, 0);
    cudaGraphDestroy(graph);
    if (err != cudaSuccess) throw std::runtime_error(std::string("cudaGraphInstantiate (batch): ") + cudaGetErrorString(err));
    e->batch_graph_launches = g.c->launches;
    e->batch_graph_bucket = bucket;
    e->batch_graph_B = B;
}

// one step for all sequences; the sampled tokens are chained into token_ids on the device (next step's inputs)
static void run_batch_step(uzu_engine* e, bool chain) {
    const uint32_t B = (uint32_t)e->seqs.size();
    uint32_t longest = 0;
    for (uint32_t b = 0; b < B; ++b) {
        if (e->seqs[b].context_length + 1 > e->max_context + MAX_ROWS || e->seqs[b].context_length + 1 > e->rope_positions)
            throw std::runtime_error("batch step: context overflow in sequence " + std::to_string(b));
        SeqSwap sw(e, b);
        state_prepare(e, e->context_length + 1);
        longest = std::max(longest, e->context_length + 1);
    }
    if (e->opts.use_cuda_graph) {
        if (!e->batch_pos_valid) {
            uint32_t host_pos[16] = {0};
            for (uint32_t b = 0; b < B; ++b) host_pos[b] = e->seqs[b].context_length;
            cudaMemcpyAsync((void*)e->batch_pos.ptr(), host_pos, sizeof host_pos, cudaMemcpyHostToDevice, e->ctx->stream);
            cudaStreamSynchronize(e->ctx->stream);      // host_pos is a local
            e->batch_pos_valid = true;
        }
        const uint32_t bucket = attention_bucket(longest);
        if (!e->batch_graph || e->batch_graph_bucket != bucket || e->batch_graph_B != B) {
            // the capture encodes with the host-side lengths of THIS step (they pick the KV-split geometry of the bucket)
            capture_batch_graph(e, bucket);
        }
        cudaError_t err = cudaGraphLaunch(e->batch_graph, e->ctx->stream);
        if (err != cudaSuccess) throw std::runtime_error(std::string("cudaGraphLaunch (batch): ") + cudaGetErrorString(err));
        e->launches += e->batch_graph_launches;
    } else {
        CmdGuard g(e->ctx, "batch step");
        g.c->state = uzu_command_buffer::Encoding;
        encode_batch_step(e, g.c, false);
        if (g.c->sticky != UZU_OK) throw std::runtime_error(g.c->sticky_msg);
        e->launches += g.c->launches;
        if (chain) cudaMemcpyAsync((void*)e->token_ids.ptr(), (void*)e->sampled.ptr(), B * 4, cudaMemcpyDeviceToDevice, e->ctx->stream);
        e->batch_pos_valid = false;
    }
    for (uint32_t b = 0; b < B; ++b) {
        for (size_t i = 0; i < e->layers.size(); ++i)
            if (e->layers[i].is_attention) e->seqs[b].layers[i].length += 1;
        e->seqs[b].context_length += 1;
    }
}

}  // namespace uzu

uzu_status uzu_engine_batch_begin(uzu_engine* e, uint32_t sequences) {
    UZU_ENGINE_TRY({
        if (sequences == 0 || sequences > 16) throw std::runtime_error("batch_begin: 1..16 sequences (the m <= 16 GEMV rows)");
        if (e->tp_sharded) throw std::runtime_error("batch_begin: not combined with tensor parallelism yet");
        if (e->sampling.kind == UZU_SAMPLING_STOCHASTIC) throw std::runtime_error("batch decode samples greedily (per-row seeds are not wired yet)");
        if (e->logits_rows < sequences) throw std::runtime_error("batch_begin: logits scratch too small");
        cudaStreamSynchronize(e->ctx->stream);
        if (e->seqs.size() != sequences && e->batch_graph) {
            // the captured batched step holds the addresses of every sequence's state: a different set of sequences needs a new capture
            cudaGraphExecDestroy(e->batch_graph);
            e->batch_graph = nullptr;
            e->batch_graph_B = 0;
        }
        while (e->seqs.size() < sequences) {
            e->seqs.emplace_back();
            alloc_sequence_state(e, e->seqs.back().layers);
        }
        auto release = [&](Buf& b) {          // buffers of a dropped sequence leave the engine's ownership list with it
            if (!b.b) return;
            auto it = std::find(e->owned.begin(), e->owned.end(), b.b);
            if (it != e->owned.end()) e->owned.erase(it);
            uzu_buffer_destroy(b.b);
            b = Buf{};
        };
        while (e->seqs.size() > sequences) {
            for (auto& S : e->seqs.back().layers) {
                if (S.k_sparse) uzu_sparse_buffer_destroy(S.k_sparse);
                if (S.v_sparse) uzu_sparse_buffer_destroy(S.v_sparse);
                release(S.k_dense); release(S.v_dense);
                release(S.conv_state); release(S.ssm_state); release(S.conv_snapshot); release(S.ssm_snapshot);
            }
            e->seqs.pop_back();
        }
        e->batch_pos_valid = false;
        for (auto& q : e->seqs) {       // reset every sequence
            q.context_length = 0;
            for (auto& S : q.layers) {
                S.length = 0;
                if (S.conv_state.b) {
                    cudaMemsetAsync((void*)S.conv_state.ptr(), 0, S.conv_bytes, e->ctx->stream);
                    cudaMemsetAsync((void*)S.ssm_state.ptr(), 0, S.ssm_bytes, e->ctx->stream);
                }
            }
        }
        cudaStreamSynchronize(e->ctx->stream);
    });
}

uzu_status uzu_engine_batch_prefill(uzu_engine* e, uint32_t sequence, const uint32_t* tokens, uint32_t count, uint32_t* out_token) {
    UZU_ENGINE_TRY({
        if (sequence >= e->seqs.size()) throw std::runtime_error("batch_prefill: no such sequence (batch_begin first)");
        e->batch_pos_valid = false;
        if (!tokens || count == 0) throw std::runtime_error("batch_prefill: empty prompt");
        uint32_t tok = 0;
        {
            SeqSwap sw(e, sequence);    // the ordinary chunked prefill, against this sequence's state
            for (uint32_t s0 = 0; s0 < count; s0 += MAX_ROWS) {
                const uint32_t n = std::min(MAX_ROWS, count - s0);
                const bool last = s0 + n == count;
                run_pass(e, tokens + s0, n, last ? n - 1 : 0, last ? n : 0, last);
            }
            cudaMemcpy(&tok, (void*)e->sampled.ptr(), 4, cudaMemcpyDeviceToHost);
        }
        if (out_token) *out_token = tok;
    });
}

uzu_status uzu_engine_batch_step(uzu_engine* e, const uint32_t* tokens_in, uint32_t* tokens_out) {
    UZU_ENGINE_TRY({
        const uint32_t B = (uint32_t)e->seqs.size();
        if (B == 0 || !tokens_in) throw std::runtime_error("batch_step: batch_begin first / null tokens");
        uint32_t* staging = (uint32_t*)uzu_buffer_cpu_ptr(e->host_tokens.b);
        memcpy(staging, tokens_in, B * 4);
        cudaMemcpyAsync((void*)e->token_ids.ptr(), staging, B * 4, cudaMemcpyHostToDevice, e->ctx->stream);
        run_batch_step(e, false);
        cudaMemcpyAsync(staging + 16, (void*)e->sampled.ptr(), B * 4, cudaMemcpyDeviceToHost, e->ctx->stream);
        cudaError_t err = cudaStreamSynchronize(e->ctx->stream);
        if (err != cudaSuccess) throw std::runtime_error(std::string("batch_step: ") + cudaGetErrorString(err));
        if (tokens_out) memcpy(tokens_out, staging + 16, B * 4);
    });
}

uzu_status uzu_engine_batch_decode_timed(uzu_engine* e, const uint32_t* first_tokens, uint32_t steps, double* out_seconds) {
    UZU_ENGINE_TRY({
        const uint32_t B = (uint32_t)e->seqs.size();
        if (B == 0 || !first_tokens || steps == 0) throw std::runtime_error("batch_decode_timed: batch_begin first / null tokens / zero steps");
        cudaStream_t s = e->ctx->stream;
        uint32_t* staging = (uint32_t*)uzu_buffer_cpu_ptr(e->host_tokens.b);
        memcpy(staging, first_tokens, B * 4);
        cudaMemcpyAsync((void*)e->token_ids.ptr(), staging, B * 4, cudaMemcpyHostToDevice, s);
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, s);
        for (uint32_t i = 0; i < steps; ++i) run_batch_step(e, true);     // sampled tokens chained on the device, no host round trip
        cudaEventRecord(b, s);
        cudaError_t err = cudaEventSynchronize(b);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, a, b);
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        if (err != cudaSuccess) throw std::runtime_error(std::string("batch_decode_timed: ") + cudaGetErrorString(err));
        if (out_seconds) *out_seconds = (double)ms * 1e-3;
    });
}

uzu_status uzu_engine_batch_logits(uzu_engine* e, uint16_t* out_logits) {
    UZU_ENGINE_TRY({
        if (e->seqs.empty() || !out_logits) throw std::runtime_error("batch_logits: batch_begin first / null output");
        cudaStreamSynchronize(e->ctx->stream);
        cudaMemcpy(out_logits, (void*)e->logits.ptr(), e->seqs.size() * (size_t)e->vocab * 2, cudaMemcpyDeviceToHost);
    });
}

uint32_t uzu_engine_batch_context_length(const uzu_engine* e, uint32_t sequence) {
    return e && sequence < e->seqs.size() ? e->seqs[sequence].context_length : 0;
}

uzu_status uzu_engine_time_linears_select(uzu_engine* e, uint32_t iters, uint32_t select, double* out_seconds, uint64_t* out_launches) {
    UZU_ENGINE_TRY({
        cudaStream_t s = e->ctx->stream;
        CmdGuard g(e->ctx, "linears");
        g.c->state = uzu_command_buffer::Encoding;
        g.c->use_pdl = e->use_pdl && !(select & 0x80000000u);   // bit 31: plain stream-ordered launches
        if (select & 64u) state_prepare(e, e->context_length + 1);
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        auto once = [&]() {
            for (auto& L : e->layers) {
                if (select & 1u) {
                    if (L.is_attention) {
                        if (L.attn.has_gate) encode_decode_gemv(e, g.c, L.attn.gate, e->hidden_b.ptr(), e->gate.ptr());
                        encode_decode_gemv(e, g.c, L.attn.qkv, e->hidden_b.ptr(), e->qkv.ptr());
                    } else {
                        encode_decode_gemv(e, g.c, L.dn.in_proj, e->hidden_b.ptr(), e->in_proj.ptr());
                    }
                }
                if (select & 2u) {
                    if (L.is_attention) encode_decode_gemv(e, g.c, L.attn.out, e->attn_out.ptr(), e->mixer_out.ptr());
                    else encode_decode_gemv(e, g.c, L.dn.out_proj, e->delta_out.ptr(), e->mixer_out.ptr());
                }
                if (select & 4u) encode_decode_gemv(e, g.c, L.up, e->hidden_b.ptr(), e->fused_up.ptr());
                if (select & 32u) {
                    uzu_fused_linear_args f{};
                    f.matmul = linear_args(L.up, e->hidden_b.ptr(), 1, e->gated.ptr());
                    f.epilogue = 1; f.act_type = L.act;
                    f.decode_stream = g_stream_lookup(e->mega.stream_of, L.up);
                    uzu_fused_linear_encode(g.c, &f);
                }
                if (select & 8u) encode_decode_gemv(e, g.c, L.down, e->gated.ptr(), e->hidden_a.ptr());
                if ((select & 64u) && L.is_attention) {     // attention mix at the current context (rewrites the same KV row each time)
                    PassCtx pc{};
                    pc.m = 1;
                    encode_attention_mix(e, g.c, L, e->state[&L - &e->layers[0]], pc, true);
                }
                if ((select & 128u) && !L.is_attention) {   // DeltaNet conv + state update (advances the recurrent state)
                    const DeltaNetLayer& D = L.dn;
                    LayerState& S = e->state[&L - &e->layers[0]];
                    uzu_delta_net_conv_update_args ca{};
                    ca.conv_weight = D.conv_weight.ptr(); ca.bias = D.conv_bias.ptr(); ca.in_out = e->in_proj.ptr(); ca.state = S.conv_state.ptr();
                    ca.kernel_size = D.kernel_size; ca.conv_dim = D.conv_dim; ca.state_stride = D.kernel_size - 1; ca.has_bias = D.conv_has_bias;
                    uzu_delta_net_conv_update_encode(g.c, &ca);
                    uzu_delta_net_update_args ua{};
                    ua.in_proj = e->in_proj.ptr(); ua.a_log = D.a_log.ptr(); ua.dt_bias = D.dt_bias.ptr(); ua.norm_weight = D.norm_weight.ptr();
                    ua.state = S.ssm_state.ptr(); ua.out = e->delta_out.ptr();
                    ua.num_v_heads = D.num_heads; ua.num_k_heads = D.num_groups; ua.head_v_dim = D.value_head_dim; ua.key_dim = D.key_dim;
                    ua.value_dim = D.value_dim; ua.norm_epsilon = D.norm_epsilon; ua.head_k_dim = D.head_dim;
                    uzu_delta_net_update_encode(g.c, &ua);
                }
            }
            if (select & 16u) encode_decode_gemv(e, g.c, e->out_emb, e->normed_out.ptr(), e->logits.ptr());
        };
        once();   // warm-up
        cudaStreamSynchronize(s);
        const uint64_t before = g.c->launches;
        cudaEventRecord(a, s);
        for (uint32_t i = 0; i < iters; ++i) once();
        cudaEventRecord(b, s);
        cudaError_t err = cudaEventSynchronize(b);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, a, b);
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        if (err != cudaSuccess || g.c->sticky != UZU_OK) throw std::runtime_error("time_linears failed: " + g.c->sticky_msg);
        if (out_seconds) *out_seconds = (double)ms * 1e-3;
        if (out_launches) *out_launches = g.c->launches - before;
    });
}

uzu_status uzu_engine_time_prefill_linears(uzu_engine* e, uint32_t m, uint32_t iters, double* out_seconds, double* out_flops) {
    UZU_ENGINE_TRY({
        if (m == 0 || m > MAX_ROWS || iters == 0) throw std::runtime_error("time_prefill_linears: m must be in 1..1024, iters > 0");
        cudaStream_t s = e->ctx->stream;
        CmdGuard g(e->ctx, "prefill linears");
        g.c->state = uzu_command_buffer::Encoding;
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        double flops = 0.0;
        auto lin = [&](const Linear& l, uint64_t in, uint64_t out, bool count) {
            encode_linear(g.c, l, in, m, out);
            if (count) flops += 2.0 * m * (double)l.out_dim * (double)l.in_dim;
        };
        auto once = [&](bool count) {
            for (auto& L : e->layers) {
                if (L.is_attention) {
                    if (L.attn.has_gate) lin(L.attn.gate, e->hidden_b.ptr(), e->gate.ptr(), count);
                    lin(L.attn.qkv, e->hidden_b.ptr(), e->qkv.ptr(), count);
                    lin(L.attn.out, e->attn_out.ptr(), e->mixer_out.ptr(), count);
                } else {
                    lin(L.dn.in_proj, e->hidden_b.ptr(), e->in_proj.ptr(), count);
                    lin(L.dn.out_proj, e->delta_out.ptr(), e->mixer_out.ptr(), count);
                }
                lin(L.up, e->hidden_b.ptr(), e->fused_up.ptr(), count);
                lin(L.down, e->gated.ptr(), e->hidden_a.ptr(), count);
            }
        };
        once(false);   // warm-up
        cudaStreamSynchronize(s);
        cudaEventRecord(a, s);
        for (uint32_t i = 0; i < iters; ++i) once(i == 0);
        cudaEventRecord(b, s);
        cudaError_t err = cudaEventSynchronize(b);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, a, b);
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        if (err != cudaSuccess || g.c->sticky != UZU_OK) throw std::runtime_error("time_prefill_linears failed: " + g.c->sticky_msg);
        if (out_seconds) *out_seconds = (double)ms * 1e-3 / iters;
        if (out_flops) *out_flops = flops;
    });
}

uzu_status uzu_engine_time_linears(uzu_engine* e, uint32_t iters, double* out_seconds, uint64_t* out_launches) {
    return uzu_engine_time_linears_select(e, iters, 31u, out_seconds, out_launches);
}

}  // extern "C"
