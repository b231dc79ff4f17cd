/*
 * uzu_oracle.c -- CPU restatement of the uzu reference CPU backend for the transformer
 * decode hot path.
 *
 * THIS IS TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it. The product (libuzu_b200.so) never
 * links, loads or calls anything in oracle/.
 *
 * The reference (trymirai/uzu @ 9670da1) is Rust; no Rust toolchain exists in this image, so
 * the reference itself cannot be compiled or run here (DESIGN.md "Oracle"). Every function
 * below restates one reference CPU kernel loop-for-loop, with the same accumulation order and
 * the same rounding points, and cites the file:line it follows (paths relative to
 * /root/reference/crates/backend-uzu/src/).
 *
 * Pinning status: the reference tests hold NO stored outputs for this path except the
 * `unit_interval` endpoints (tests/unit/encodable_block/sampling/gumbel_test.rs:7-12), which
 * tests/test_oracle_pins.py checks, together with the published Philox4x32-10 known-answer
 * vectors (Random123 kat_vectors), the reference tests' own equivalence properties
 * (gather == dense readout, attention vs independent softmax reference, two-pass == single
 * pass) evaluated on the reference tests' closed-form inputs. Everything else is "parity
 * unpinned": it matches the reference by construction (restatement), not by known answers.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile). -ffp-contract=off keeps
 * every a*b+c as two IEEE roundings exactly like rustc emits for f32 `a * b + c`.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * bf16 <-> f32, `half` crate 2.7.1 semantics (bf16::from_f32 = round-to-nearest-even,
 * NaN quieted; bf16::to_f32 = shift). Call sites: backends/cpu/kernel/matmul/reference.rs:113-143
 * and every T::from(..)/to_f32() in the kernels.
 * ---------------------------------------------------------------------------------------- */
typedef uint16_t bf16_t;

static inline float bf2f(bf16_t h) {
    uint32_t u = ((uint32_t)h) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

static inline bf16_t f2bf(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    if ((x & 0x7FFFFFFFu) > 0x7F800000u) return (bf16_t)((x >> 16) | 0x0040u);
    uint32_t round_bit = 0x00008000u;
    if ((x & round_bit) != 0 && (x & (3u * round_bit - 1u)) != 0) return (bf16_t)((x >> 16) + 1u);
    return (bf16_t)(x >> 16);
}

ORACLE_API float oracle_bf16_to_f32(uint16_t h) { return bf2f(h); }
ORACLE_API uint16_t oracle_f32_to_bf16(float f) { return f2bf(f); }

/* bf16 (op) bf16 -> bf16: `half` implements arithmetic as from_f32(a.to_f32() op b.to_f32()). */
static inline bf16_t bf_add(bf16_t a, bf16_t b) { return f2bf(bf2f(a) + bf2f(b)); }
static inline bf16_t bf_mul(bf16_t a, bf16_t b) { return f2bf(bf2f(a) * bf2f(b)); }

/* Element type tags used for the few buffers whose dtype varies (matmul D, weights). */
enum { ORACLE_DT_BF16 = 0, ORACLE_DT_F32 = 1 };

static inline float read_f32(const void* base, int dt, size_t i) {
    return dt == ORACLE_DT_F32 ? ((const float*)base)[i] : bf2f(((const bf16_t*)base)[i]);
}
static inline void write_f32(void* base, int dt, size_t i, float v) {
    if (dt == ORACLE_DT_F32) ((float*)base)[i] = v;
    else ((bf16_t*)base)[i] = f2bf(v);
}

/* ------------------------------------------------------------------------------------------
 * Activation math: backends/common/gpu_types/activation_type.rs:16-65.
 * `activate<T>` computes in f32 and returns T (one rounding for T = bf16).
 * ---------------------------------------------------------------------------------------- */
enum { ACT_SILU = 0, ACT_GELU_APPROX = 1, ACT_GELU_EXACT = 2, ACT_IDENTITY = 3, ACT_SOFTPLUS = 4 };

static inline float act_f32(int act, float x, int* passthrough) {
    *passthrough = 0;
    switch (act) {
        case ACT_SILU: return x / (1.0f + expf(-1.0f * x)); /* activation_silu_alpha(x, 1.0), :31-38 */
        case ACT_GELU_APPROX: {
            const float k0 = 0.044715f, k1 = 0.7978846f;
            float t = k1 * (x + k0 * x * x * x);
            return 0.5f * x * (1.0f + tanhf(t));
        }
        case ACT_GELU_EXACT: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
        case ACT_IDENTITY: *passthrough = 1; return x;
        case ACT_SOFTPLUS:
            if (x > 20.0f) { *passthrough = 1; return x; }
            return logf(1.0f + expf(x));
    }
    return x;
}

static inline bf16_t act_bf16(int act, bf16_t x) {
    int pass;
    float y = act_f32(act, bf2f(x), &pass);
    return pass ? x : f2bf(y);
}

/* ------------------------------------------------------------------------------------------
 * Matmul: backends/cpu/kernel/matmul/kernel.rs:164-295 (+ reference.rs:12-143).
 *   D[r,c] = softcap(ab_scale * sum_k A[r,k]*W[b_col,k] (+D) (+bias[c]))
 * Quantized W: u32 little-endian words, element (b_col*K + k) at bit ((idx % pf) * bits)
 * (:236-243); signed_codes XORs the top bit (:243-245); correction: zero-point (:254-266),
 * MLX bias, or symmetric midpoint (:267-273). f32 accumulate in k order; weight value is
 * `scale * code + corr` rounded in f32 before the multiply (:274-277).
 * ---------------------------------------------------------------------------------------- */
enum { QM_NONE = 0, QM_SCALE_BIAS = 1, QM_ZERO_POINT = 2, QM_SYMMETRIC = 3 };

typedef struct {
    const void* a;          /* [m,k] bf16 (a_dt) row-major */
    int a_dt;
    const void* w;          /* quantized: packed codes; full precision: [n,ld] (w_dt) */
    const void* scales;     /* [n, ceil(k/gs)] w_dt */
    const uint8_t* zero_points; /* 4-bit: [n, ceil(groups/2)] nibble-packed; 8-bit: [n, groups] */
    const void* biases;     /* MLX: [n, groups] w_dt */
    int w_dt;               /* dtype of fp weights / scales / biases / epilogue bias */
    int method;             /* QM_* */
    int bits;               /* 4 or 8 */
    int group_size;
    int signed_codes;
    int b_transpose;        /* fp only; quantized is always [n,k] */
    int ld;                 /* fp leading dimension, 0 = default */
    void* d;                /* [m,n] d_dt */
    int d_dt;
    const uint32_t* gather; /* optional [m,n] row indices into W */
    float ab_scale;
    int accumulate;
    const void* bias;       /* optional [n] w_dt */
    int has_soft_cap;
    float soft_cap;
    int m, n, k;
} oracle_matmul_args;

static void matmul_rows(const oracle_matmul_args* p, int row, int col_begin, int col_end) {
    const int k_u = p->k, n_u = p->n;
    const int quant = p->method != QM_NONE;
    int num_groups_k = 0, zp_stride = 0, pack_factor = 0;
    if (quant) {
        num_groups_k = (k_u + p->group_size - 1) / p->group_size;
        zp_stride = p->bits == 4 ? (num_groups_k + 1) / 2 : num_groups_k;
        pack_factor = p->bits == 4 ? 8 : 4;
    }
    for (int col = col_begin; col < col_end; ++col) {
        size_t b_col = p->gather ? p->gather[(size_t)row * n_u + col] : (size_t)col;
        float acc = 0.0f;
        for (int inner = 0; inner < k_u; ++inner) {
            float a_value = read_f32(p->a, p->a_dt, (size_t)row * k_u + inner);
            float b_value;
            if (!quant) {
                size_t ldim = p->ld ? (size_t)p->ld : (size_t)(p->b_transpose ? k_u : n_u);
                size_t index = p->b_transpose ? b_col * ldim + inner : (size_t)inner * ldim + b_col;
                b_value = read_f32(p->w, p->w_dt, index);
            } else {
                size_t lin = b_col * (size_t)k_u + inner;
                size_t word_index = lin / pack_factor;
                unsigned bit_offset = (unsigned)(lin % pack_factor) * (unsigned)p->bits;
                uint32_t word;
                memcpy(&word, (const uint8_t*)p->w + word_index * 4, 4);
                uint32_t mask = (1u << p->bits) - 1u;
                uint8_t code = (uint8_t)((word >> bit_offset) & mask);
                if (p->signed_codes) code ^= (uint8_t)(1u << (p->bits - 1));
                float q = (float)code;
                int g = inner / p->group_size;
                float scale = read_f32(p->scales, p->w_dt, b_col * num_groups_k + g);
                float midpoint = (float)(1u << (p->bits - 1));
                float bias_term;
                if (p->method == QM_ZERO_POINT) {
                    float zp;
                    if (p->bits == 4) {
                        uint8_t byte = p->zero_points[b_col * zp_stride + (g >> 1)];
                        zp = (g & 1) == 0 ? (float)(byte & 0x0F) : (float)((byte >> 4) & 0x0F);
                    } else {
                        zp = (float)p->zero_points[b_col * zp_stride + g];
                    }
                    bias_term = -scale * zp;
                } else if (p->method == QM_SCALE_BIAS) {
                    bias_term = read_f32(p->biases, p->w_dt, b_col * num_groups_k + g);
                } else {
                    bias_term = -scale * midpoint;
                }
                b_value = scale * q + bias_term;
            }
            acc += a_value * b_value;
        }
        size_t out = (size_t)row * n_u + col;
        float value = p->ab_scale * acc;
        if (p->accumulate) value += read_f32(p->d, p->d_dt, out);
        if (p->bias) value += read_f32(p->bias, p->w_dt, col);
        if (p->has_soft_cap) value = p->soft_cap * tanhf(value / p->soft_cap);
        write_f32(p->d, p->d_dt, out, value);
    }
}

/* threads <= 1: the reference's execution model (one worker thread, cpu/context.rs:21-27).
 * threads  > 1: same arithmetic per output element, OpenMP over output columns ("courtesy"
 * baseline; not the reference's execution model). */
ORACLE_API void oracle_matmul(const oracle_matmul_args* p, int threads) {
    for (int row = 0; row < p->m; ++row) {
#ifdef _OPENMP
        if (threads > 1) {
            int chunk = 16;
            int nchunks = (p->n + chunk - 1) / chunk;
#pragma omp parallel for num_threads(threads) schedule(static)
            for (int c = 0; c < nchunks; ++c) {
                int b = c * chunk, e = b + chunk > p->n ? p->n : b + chunk;
                matmul_rows(p, row, b, e);
            }
            continue;
        }
#endif
        (void)threads;
        matmul_rows(p, row, 0, p->n);
    }
}

/* ------------------------------------------------------------------------------------------
 * Normalization: backends/cpu/kernel/normalization/normalization.rs:50-125, instantiated as
 * InputT = OutputT = bf16, AffineT = AccumT = f32 (encodable_block/normalization.rs:84-100).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const bf16_t* input;   /* NULL when in_place */
    const float* scales;   /* optional */
    const float* biases;   /* optional */
    bf16_t* output;
    bf16_t* shortcut;      /* optional (copy_to_shortcut) */
    int batch_size, element_count;
    float epsilon, scale_offset, post_layer_scalar;
    int in_place, subtract_mean, full_layer, copy_to_shortcut, residual_add;
    int scale_residual_sum, scale_output;
} oracle_norm_args;

ORACLE_API void oracle_normalization(const oracle_norm_args* p) {
    const bf16_t* input = p->in_place ? (const bf16_t*)p->output : p->input;
    const int n = p->element_count;
    const float nf = (float)n;
    for (int b = 0; b < p->batch_size; ++b) {
        size_t off = (size_t)b * n;
        float sum = 0.0f, sum_sq = 0.0f;
        for (int i = 0; i < n; ++i) {
            bf16_t val = input[off + i];
            if (p->copy_to_shortcut) {
                bf16_t* skip = p->shortcut + off + i;
                if (p->residual_add) {
                    val = bf_add(val, *skip);                                     /* :72-73 */
                    if (p->scale_residual_sum) val = f2bf(bf2f(val) * p->post_layer_scalar);
                }
                *skip = val;                                                      /* :78 */
            }
            float av = bf2f(val);
            if (p->subtract_mean) sum = sum + av;
            sum_sq = sum_sq + av * av;
        }
        float mean = p->subtract_mean ? sum / nf : 0.0f;
        float variance = sum_sq / nf - mean * mean;
        float rms_inv = 1.0f / sqrtf(variance + p->epsilon);                     /* .sqrt().recip() */
        for (int i = 0; i < n; ++i) {
            float iv = p->residual_add ? bf2f(p->shortcut[off + i]) : bf2f(input[off + i]);
            float normalized = (iv - mean) * rms_inv;
            bf16_t result;
            if (p->scales) {
                float sv = p->scales[i];
                if (p->full_layer) {
                    result = f2bf(normalized * (sv + p->scale_offset));          /* :104-106 */
                } else {
                    result = bf_mul(f2bf(normalized), f2bf(sv + p->scale_offset)); /* :108-111 */
                }
            } else {
                result = f2bf(normalized);
            }
            if (p->biases) result = f2bf(bf2f(result) + p->biases[i]);
            if (p->scale_output) result = bf_mul(result, f2bf(p->post_layer_scalar));
            p->output[off + i] = result;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * QKVNorm: backends/cpu/kernel/attention/qkv_norm.rs:36-76 (in place, AccumT = f32,
 * ScaleT = f32; encodable_block/mixer/attention/qkv_norm.rs:97-111).
 * ---------------------------------------------------------------------------------------- */
ORACLE_API void oracle_qkv_norm(bf16_t* qkv, const float* scales, int batch_size, int total_heads,
                                int head_dim, float epsilon, float scale_offset, int head_offset,
                                int head_count, int full_layer) {
    size_t stride = (size_t)total_heads * head_dim;
    float hd = (float)head_dim;
    for (int b = 0; b < batch_size; ++b)
        for (int h = 0; h < head_count; ++h) {
            size_t off = (size_t)b * stride + (size_t)(head_offset + h) * head_dim;
            float total = 0.0f;
            for (int i = 0; i < head_dim; ++i) {
                float v = bf2f(qkv[off + i]);
                total = total + v * v;
            }
            float mean_square = total / hd;
            float rms = 1.0f / sqrtf(mean_square + epsilon);
            for (int i = 0; i < head_dim; ++i) {
                float normalized = bf2f(qkv[off + i]) * rms;
                bf16_t r;
                if (!scales) r = f2bf(normalized);
                else if (full_layer) r = f2bf(normalized * (scales[i] + scale_offset));
                else r = bf_mul(f2bf(normalized), f2bf(scales[i] + scale_offset));
                qkv[off + i] = r;
            }
        }
}

/* ------------------------------------------------------------------------------------------
 * RoPE tables: encodable_block/mixer/attention/rope.rs:13-114 (host f32 math).
 * kind: 0 unscaled, 1 linear, 2 llama3. (YaRN / LongRoPE: not in any BASELINE config.)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int kind;
    float base;
    int head_dim;
    float scaling_factor;
    int original_context_length;
    float low_frequency_factor, high_frequency_factor;
} oracle_rope_config;

ORACLE_API void oracle_rope_tables(const oracle_rope_config* c, const uint32_t* positions, int count,
                                   float* cosines, float* sines) {
    int head_dim = c->head_dim, half = head_dim / 2;
    const float attention_scaling_factor = 1.0f;
    for (int pair = 0; pair < half; ++pair) {
        int channel = pair * 2;
        float inv = 1.0f / powf(c->base, (float)channel / (float)head_dim);
        if (c->kind == 1) {
            inv = inv / c->scaling_factor;
        } else if (c->kind == 2) {
            float low_wl = (float)c->original_context_length / c->low_frequency_factor;
            float high_wl = (float)c->original_context_length / c->high_frequency_factor;
            float wavelength = 2.0f * 3.14159265358979323846f / inv;
            float scaled = inv / c->scaling_factor;
            if (wavelength < high_wl) {
                /* keep */
            } else if (wavelength > low_wl) {
                inv = scaled;
            } else {
                float smooth = (float)c->original_context_length / wavelength - c->low_frequency_factor;
                smooth = smooth / (c->high_frequency_factor - c->low_frequency_factor);
                inv = smooth * inv + (1.0f - smooth) * scaled;
            }
        }
        for (int t = 0; t < count; ++t) {
            float e = (float)positions[t] * inv;
            float s = sinf(e) * attention_scaling_factor;
            float co = cosf(e) * attention_scaling_factor;
            size_t po = (size_t)t * head_dim + pair;
            sines[po] = s; sines[po + half] = s;
            cosines[po] = co; cosines[po + half] = co;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * AttentionPrepare: backends/cpu/kernel/attention/attention_prepare.rs:7-126.
 * ---------------------------------------------------------------------------------------- */
static inline bf16_t apply_rope(const bf16_t* head, const float* cosines, const float* sines,
                                int batch_idx, int d, int rope_dim) {
    int half = rope_dim / 2;
    int paired = d < half ? d + half : d - half;
    float input = bf2f(head[d]);
    float p = bf2f(head[paired]);
    float signed_p = d < half ? -p : p;
    float c = cosines[(size_t)batch_idx * rope_dim + d];
    float s = sines[(size_t)batch_idx * rope_dim + d];
    return f2bf(input * c + signed_p * s);
}

ORACLE_API void oracle_attention_prepare(const bf16_t* qkv, bf16_t* queries, bf16_t* keys, bf16_t* values,
                                         const float* cosines, const float* sines, int num_q_heads,
                                         int num_kv_heads, int head_dim, int rope_dim,
                                         int kv_token_offset, int batch_dim, int has_kv, int has_rope) {
    int total_heads = has_kv ? num_q_heads + 2 * num_kv_heads : num_q_heads;
    for (int b = 0; b < batch_dim; ++b)
        for (int h = 0; h < total_heads; ++h) {
            const bf16_t* head = qkv + ((size_t)b * total_heads + h) * head_dim;
            int is_query = !has_kv || h < num_q_heads;
            int is_key = has_kv && h >= num_q_heads && h < num_q_heads + num_kv_heads;
            for (int d = 0; d < head_dim; ++d) {
                bf16_t e = head[d];
                if (has_rope && d < rope_dim && (is_query || is_key))
                    e = apply_rope(head, cosines, sines, b, d, rope_dim);
                if (is_query) {
                    queries[((size_t)h * batch_dim + b) * head_dim + d] = e;
                } else if (is_key) {
                    keys[((size_t)(kv_token_offset + b) * num_kv_heads + (h - num_q_heads)) * head_dim + d] = e;
                } else {
                    values[((size_t)(kv_token_offset + b) * num_kv_heads + (h - num_q_heads - num_kv_heads)) * head_dim + d] = e;
                }
            }
        }
}

/* ------------------------------------------------------------------------------------------
 * Attention mask: backends/cpu/kernel/attention/mask.rs:3-62.
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint32_t trie_start, trie_end, height; } oracle_trie_node;

typedef struct {
    int has_ring; uint32_t ring_offset, ring_length;
    const oracle_trie_node* trie;            /* NULL unless is_trie */
    int has_sliding_window; uint32_t sliding_window_size;
    int is_causal;
} oracle_mask;

static int should_use_key(const oracle_mask* m, uint32_t q_seq_idx, uint32_t prefix_length,
                          uint32_t suffix_position, uint32_t query_position, uint32_t i) {
    int use_key = 1;
    uint32_t key_position;
    if (i >= prefix_length) {
        uint32_t kis = i - prefix_length;
        if (m->trie) {
            const oracle_trie_node* node = m->trie + kis;
            key_position = suffix_position + node->height;
            if (m->is_causal) use_key &= (q_seq_idx >= node->trie_start && q_seq_idx <= node->trie_end);
        } else {
            key_position = suffix_position + kis;
            if (m->is_causal) use_key &= (kis <= q_seq_idx);
        }
    } else {
        if (m->has_ring) {
            key_position = (prefix_length + i - m->ring_offset) % prefix_length;
            use_key &= key_position < m->ring_length;
        } else {
            key_position = i;
        }
    }
    if (m->has_sliding_window) {
        uint32_t w = m->sliding_window_size;
        if (m->is_causal) use_key &= (key_position <= query_position && (query_position - key_position) < w);
        else if (key_position <= query_position) use_key &= ((query_position - key_position) <= w / 2);
        else use_key &= ((key_position - query_position) <= w / 2);
    }
    return use_key;
}

typedef struct {
    const bf16_t* queries;  /* [num_heads, suffix, D] */
    const bf16_t* keys;
    const bf16_t* values;
    int head_dim, gqa_factor, sequence_length;
    int k_head_stride, k_seq_stride, v_head_stride, v_seq_stride;
    float scale;
    const bf16_t* sinks;    /* optional [num_heads] */
    int num_heads, suffix_length;
    oracle_mask mask;
} oracle_attn_args;

/* AttentionSinglePass: backends/cpu/kernel/attention/attention_single_pass.rs:49-126. */
ORACLE_API void oracle_attention_single_pass(const oracle_attn_args* p, bf16_t* out) {
    const int D = p->head_dim;
    float* q = (float*)malloc(sizeof(float) * D);
    float* o = (float*)malloc(sizeof(float) * D);
    uint32_t prefix_length = (uint32_t)(p->sequence_length - p->suffix_length);
    uint32_t suffix_position = p->mask.has_ring ? p->mask.ring_length : prefix_length;
    for (int h = 0; h < p->num_heads; ++h)
        for (int qs = 0; qs < p->suffix_length; ++qs) {
            int kvh = h / p->gqa_factor;
            size_t o_off = (size_t)qs * p->num_heads + h;
            size_t q_off = (size_t)h * p->suffix_length + qs;
            uint32_t query_position = p->mask.trie ? suffix_position + p->mask.trie[qs].height
                                                   : suffix_position + (uint32_t)qs;
            const bf16_t* qp = p->queries + q_off * D;
            const bf16_t* kb = p->keys + (size_t)kvh * p->k_head_stride;
            const bf16_t* vb = p->values + (size_t)kvh * p->v_head_stride;
            for (int j = 0; j < D; ++j) { q[j] = p->scale * bf2f(qp[j]); o[j] = 0.0f; }
            float max_score = -INFINITY, sum_exp = 0.0f;
            if (p->sinks) { max_score = bf2f(p->sinks[h % p->num_heads]); sum_exp = 1.0f; }
            for (int i = 0; i < p->sequence_length; ++i) {
                if (!should_use_key(&p->mask, (uint32_t)qs, prefix_length, suffix_position, query_position, (uint32_t)i)) continue;
                const bf16_t* kp = kb + (size_t)i * p->k_seq_stride;
                float score = 0.0f;
                for (int j = 0; j < D; ++j) score += q[j] * bf2f(kp[j]);
                float new_max = fmaxf(max_score, score);
                float factor = expf(max_score - new_max);
                float es = expf(score - new_max);
                max_score = new_max;
                sum_exp = sum_exp * factor + es;
                const bf16_t* vp = vb + (size_t)i * p->v_seq_stride;
                for (int j = 0; j < D; ++j) o[j] = o[j] * factor + es * bf2f(vp[j]);
            }
            bf16_t* op = out + o_off * D;
            for (int j = 0; j < D; ++j) op[j] = f2bf(o[j] / sum_exp);
        }
    free(q); free(o);
}

/* AttentionTwoPass1/2: backends/cpu/kernel/attention/attention_two_pass.rs:55-189. */
#define TWO_PASS_BLOCKS 32

ORACLE_API void oracle_attention_two_pass1(const oracle_attn_args* p, float* out, float* sums, float* maxs) {
    const int D = p->head_dim;
    float* q = (float*)malloc(sizeof(float) * D);
    float* o = (float*)malloc(sizeof(float) * D);
    uint32_t prefix_length = (uint32_t)(p->sequence_length - p->suffix_length);
    uint32_t suffix_position = p->mask.has_ring ? p->mask.ring_length : prefix_length;
    for (int h = 0; h < p->num_heads; ++h)
        for (int qs = 0; qs < p->suffix_length; ++qs) {
            uint32_t query_position = p->mask.trie ? suffix_position + p->mask.trie[qs].height
                                                   : suffix_position + (uint32_t)qs;
            for (int blk = 0; blk < TWO_PASS_BLOCKS; ++blk) {
                size_t o_off = (size_t)qs * p->num_heads + h;
                size_t q_off = (size_t)h * p->suffix_length + qs;
                int kvh = h / p->gqa_factor;
                const bf16_t* qp = p->queries + q_off * D;
                const bf16_t* kb = p->keys + (size_t)kvh * p->k_head_stride;
                const bf16_t* vb = p->values + (size_t)kvh * p->v_head_stride;
                float* ob = out + (o_off * TWO_PASS_BLOCKS + blk) * D;
                for (int j = 0; j < D; ++j) { q[j] = p->scale * bf2f(qp[j]); o[j] = 0.0f; }
                float max_score = -1e9f, sum_exp = 0.0f;
                if (p->sinks && blk == 0) { max_score = bf2f(p->sinks[h]); sum_exp = 1.0f; }
                for (int i = blk; i < p->sequence_length; i += TWO_PASS_BLOCKS) {
                    if (!should_use_key(&p->mask, (uint32_t)qs, prefix_length, suffix_position, query_position, (uint32_t)i)) continue;
                    const bf16_t* kp = kb + (size_t)i * p->k_seq_stride;
                    float score = 0.0f;
                    for (int j = 0; j < D; ++j) score += q[j] * bf2f(kp[j]);
                    float new_max = fmaxf(max_score, score);
                    float factor = expf(max_score - new_max);
                    float es = expf(score - new_max);
                    max_score = new_max;
                    sum_exp = sum_exp * factor + es;
                    const bf16_t* vp = vb + (size_t)i * p->v_seq_stride;
                    for (int j = 0; j < D; ++j) o[j] = o[j] * factor + es * bf2f(vp[j]);
                }
                for (int j = 0; j < D; ++j) ob[j] = o[j];
                sums[o_off * TWO_PASS_BLOCKS + blk] = sum_exp;
                maxs[o_off * TWO_PASS_BLOCKS + blk] = max_score;
            }
        }
    free(q); free(o);
}

ORACLE_API void oracle_attention_two_pass2(const float* partials, const float* sums, const float* maxs,
                                           bf16_t* out, int head_dim, int num_heads, int suffix_length) {
    const int D = head_dim;
    for (int h = 0; h < num_heads; ++h)
        for (int qs = 0; qs < suffix_length; ++qs) {
            size_t o_off = (size_t)qs * num_heads + h;
            float gmax = -INFINITY;
            for (int b = 0; b < TWO_PASS_BLOCKS; ++b) gmax = fmaxf(gmax, maxs[o_off * TWO_PASS_BLOCKS + b]);
            float gsum = 0.0f;
            for (int b = 0; b < TWO_PASS_BLOCKS; ++b)
                gsum += sums[o_off * TWO_PASS_BLOCKS + b] * expf(maxs[o_off * TWO_PASS_BLOCKS + b] - gmax);
            for (int j = 0; j < D; ++j) {
                float val = 0.0f;
                for (int b = 0; b < TWO_PASS_BLOCKS; ++b)
                    val += partials[(o_off * TWO_PASS_BLOCKS + b) * D + j] * expf(maxs[o_off * TWO_PASS_BLOCKS + b] - gmax);
                out[o_off * D + j] = f2bf(val / gsum);
            }
        }
}

/* KVCacheUpdate: backends/cpu/kernel/attention/kv_cache_update.rs:7-28. */
ORACLE_API void oracle_kv_cache_update(bf16_t* keys, bf16_t* values, const uint32_t* copies /* (src,dst) pairs */,
                                       int copy_count, int element_dim) {
    for (int e = 0; e < element_dim; ++e)
        for (int i = 0; i < copy_count; ++i) {
            size_t s = (size_t)copies[2 * i] * element_dim + e, d = (size_t)copies[2 * i + 1] * element_dim + e;
            keys[d] = keys[s];
            values[d] = values[s];
        }
}

/* SigmoidGate: backends/cpu/kernel/attention/sigmoid_gate.rs:7-22. */
ORACLE_API void oracle_sigmoid_gate(const bf16_t* gate, bf16_t* output, int total) {
    for (int i = 0; i < total; ++i) {
        float g = bf2f(gate[i]);
        float sg = 1.0f / (1.0f + expf(-g));
        output[i] = f2bf(bf2f(output[i]) * sg);
    }
}

/* ------------------------------------------------------------------------------------------
 * GatedActMul (FullPrecision op, interleaved, no Hadamard):
 * backends/cpu/kernel/gated_act_mul/{mod.rs:5-12, gated_act_mul.rs:43-66}. Two bf16 roundings:
 * activate() -> bf16, then bf16*bf16 -> bf16; result re-rounded (exact) by T::from(f32).
 * ---------------------------------------------------------------------------------------- */
ORACLE_API void oracle_gated_act_mul(const bf16_t* fused_up, bf16_t* out, int gated_dim, int batch_dim, int act) {
    for (int b = 0; b < batch_dim; ++b)
        for (int g = 0; g < gated_dim; ++g) {
            size_t base = (size_t)b * 2 * gated_dim;
            bf16_t value = fused_up[base + g];
            bf16_t gate = fused_up[base + gated_dim + g];
            out[(size_t)b * gated_dim + g] = bf_mul(value, act_bf16(act, gate));
        }
}

/* ------------------------------------------------------------------------------------------
 * Embedding lookups: backends/cpu/kernel/embedding/quant_embedding.rs:11-117 (byte-wise nibble,
 * even dim -> low nibble), full_precision_embedding.rs:7-32.
 * mode: 0 = U4, 1 = I8, 2 = U8 (gpu_types/quantization.rs:9).
 * ---------------------------------------------------------------------------------------- */
ORACLE_API void oracle_quant_embedding_lookup(const uint32_t* token_ids, const uint8_t* weights, const bf16_t* scales,
                                              const uint8_t* zero_points, const bf16_t* biases, bf16_t* output,
                                              int batch_size, uint32_t vocab_size, int model_dim, float input_scale,
                                              int group_size, int mode, int method) {
    int packing = mode == 0 ? 2 : 1;
    size_t wstride = (size_t)model_dim / packing;
    int num_groups = (model_dim + group_size - 1) / group_size;
    size_t zstride = mode == 0 ? (size_t)(num_groups + 1) / 2 : (size_t)num_groups;
    for (int b = 0; b < batch_size; ++b) {
        uint32_t tok = token_ids[b];
        for (int d = 0; d < model_dim; ++d) {
            size_t oi = (size_t)b * model_dim + d;
            if (tok >= vocab_size) { output[oi] = 0; continue; }
            int g = d / group_size;
            float scale = bf2f(scales[(size_t)tok * num_groups + g]);
            int32_t qv;
            if (mode == 0) {
                uint8_t packed = weights[(size_t)tok * wstride + d / 2];
                qv = (d & 1) == 0 ? (packed & 0x0F) : ((packed >> 4) & 0x0F);
            } else if (mode == 1) {
                qv = ((const int8_t*)weights)[(size_t)tok * wstride + d];
            } else {
                qv = weights[(size_t)tok * wstride + d];
            }
            float bias;
            if (method == QM_SCALE_BIAS) {
                bias = bf2f(biases[(size_t)tok * num_groups + g]);
            } else if (method == QM_ZERO_POINT) {
                uint8_t zp;
                if (mode == 0) {
                    uint8_t packed = zero_points[(size_t)tok * zstride + g / 2];
                    zp = (g & 1) == 0 ? (packed & 0x0F) : ((packed >> 4) & 0x0F);
                } else {
                    zp = zero_points[(size_t)tok * zstride + g];
                }
                bias = -scale * (float)zp;
            } else {
                int midpoint = 1 << ((mode == 0 ? 4 : 8) - 1);
                bias = -scale * (float)midpoint;
            }
            float of = scale * (float)qv + bias;
            of = of * input_scale;
            output[oi] = f2bf(of);
        }
    }
}

ORACLE_API void oracle_fp_embedding_lookup(const uint32_t* token_ids, const bf16_t* weights, bf16_t* output,
                                           int batch_size, uint32_t vocab_size, int model_dim, float input_scale) {
    for (int b = 0; b < batch_size; ++b) {
        uint32_t tok = token_ids[b];
        for (int d = 0; d < model_dim; ++d) {
            size_t oi = (size_t)b * model_dim + d;
            if (tok >= vocab_size) output[oi] = 0;
            else output[oi] = bf_mul(weights[(size_t)tok * model_dim + d], f2bf(input_scale));
        }
    }
}

/* LogitTransform: backends/cpu/kernel/logit_transform/logit_transform.rs:7-26. */
ORACLE_API void oracle_logit_transform(bf16_t* logits, int length, float scale, float soft_cap, int has_soft_cap) {
    for (int i = 0; i < length; ++i) {
        float v = bf2f(logits[i]) * scale;
        if (has_soft_cap) v = tanhf(v / soft_cap) * soft_cap;
        logits[i] = f2bf(v);
    }
}

/* Elementwise glue: backends/cpu/kernel/tensor_{add_scale,copy,add_bias,add_swap}/ *.rs. */
ORACLE_API void oracle_tensor_add_scale(const bf16_t* input /* NULL = in place */, const bf16_t* bias, bf16_t* output,
                                        int num_cols, int length, float scale) {
    for (int i = 0; i < length; ++i) {
        float iv = bf2f(input ? input[i] : output[i]);
        output[i] = f2bf((iv + bf2f(bias[i % num_cols])) * scale);
    }
}
ORACLE_API void oracle_tensor_copy(const bf16_t* src, bf16_t* dst, int length) {
    for (int i = 0; i < length; ++i) dst[i] = src[i];
}
ORACLE_API void oracle_tensor_add_bias(const bf16_t* input /* NULL = in place */, const bf16_t* bias, bf16_t* output,
                                       int num_cols, int length) {
    for (int i = 0; i < length; ++i) {
        float v = bf2f(input ? input[i] : output[i]);
        output[i] = f2bf(v + bf2f(bias[i % num_cols]));
    }
}
ORACLE_API void oracle_tensor_add_swap(bf16_t* skip, bf16_t* main_, int length) {
    for (int i = 0; i < length; ++i) {
        bf16_t r = bf_add(skip[i], main_[i]);
        skip[i] = r; main_[i] = r;
    }
}

/* ------------------------------------------------------------------------------------------
 * ActivationTransform (Mirai RHT): backends/cpu/kernel/activation_transform/
 * {activation_transform.rs:44-136, mod.rs:31-44}. 32-point Walsh-Hadamard per stripe in f32
 * (butterflies stride 1,2,4,8,16: lower = a+b, upper = a-b; then * 1/sqrt(32)); InputRht multiplies
 * by the +-1 factors before the transform, OutputRht after. Quantize ops: symmetric int8 per
 * activation group on the InputRht-transformed row (mod.rs:9-29, activation_transform.rs:11-42).
 * op: 0 InputRht, 1 OutputRht, 2 Quantize, 3 QuantizeWithGroupSums (gpu_types ActivationTransformOp).
 * `input` may alias `fp_out` (in_place): the row is transformed into a scratch row first, like the
 * reference's `transformed` vector.
 * ---------------------------------------------------------------------------------------- */
static void hadamard32(float* v) {
    for (int stride = 1; stride < 32; stride <<= 1)
        for (int lane = 0; lane < 32; ++lane)
            if ((lane & stride) == 0) {
                float a = v[lane], b = v[lane | stride];
                v[lane] = a + b;
                v[lane | stride] = a - b;
            }
    const float scale = 1.0f / sqrtf(32.0f);
    for (int i = 0; i < 32; ++i) v[i] *= scale;
}

static float min_max_symmetric_divisor(const float* v, int n) {
    float mn = INFINITY, mx = -INFINITY;
    for (int i = 0; i < n; ++i) { mn = fminf(mn, v[i]); mx = fmaxf(mx, v[i]); }
    float mag = fmaxf(fabsf(mn), fabsf(mx));
    return (isfinite(mag) && mag > 0.0f) ? mag / 127.0f : 1.0f;
}

ORACLE_API void oracle_activation_transform(const void* input, int is_f32, void* fp_out, int8_t* q_out, float* scales_out,
                                            int32_t* group_sums_out, const int32_t* factors, int rows, int cols, int op,
                                            int activation_group_size, int sum_group_size) {
    const int input_rht = op != 1;
    float* t = (float*)malloc((size_t)cols * sizeof(float));
    for (int r = 0; r < rows; ++r) {
        const size_t off = (size_t)r * cols;
        for (int s0 = 0; s0 < cols; s0 += 32) {
            float stripe[32];
            for (int l = 0; l < 32; ++l) {
                float v = is_f32 ? ((const float*)input)[off + s0 + l] : bf2f(((const bf16_t*)input)[off + s0 + l]);
                stripe[l] = input_rht ? v * (float)factors[s0 + l] : v;
            }
            hadamard32(stripe);
            for (int l = 0; l < 32; ++l) t[s0 + l] = input_rht ? stripe[l] : stripe[l] * (float)factors[s0 + l];
        }
        if (op >= 2) {
            const int groups = cols / activation_group_size;
            if (op == 3) for (int g = 0; g < cols / sum_group_size; ++g) group_sums_out[(size_t)r * (cols / sum_group_size) + g] = 0;
            for (int g = 0; g < groups; ++g) {
                const float* src = t + (size_t)g * activation_group_size;
                const float scale = min_max_symmetric_divisor(src, activation_group_size);
                scales_out[(size_t)r * groups + g] = scale;
                for (int i = 0; i < activation_group_size; ++i) {
                    float q = roundf(src[i] / scale);           /* f32::round: half away from zero */
                    q = fminf(fmaxf(q, -127.0f), 127.0f);
                    const int idx = g * activation_group_size + i;
                    q_out[off + idx] = (int8_t)q;
                    if (op == 3) group_sums_out[(size_t)r * (cols / sum_group_size) + idx / sum_group_size] += (int)q;
                }
            }
        } else {
            for (int i = 0; i < cols; ++i) {
                if (is_f32) ((float*)fp_out)[off + i] = t[i];
                else ((bf16_t*)fp_out)[off + i] = f2bf(t[i]);
            }
        }
    }
    free(t);
}

/* ------------------------------------------------------------------------------------------
 * Sampling RNG: encodable_block/sampling/gumbel.rs:1-81 (Philox4x32-10, key = 64-bit seed,
 * counter = [offset,0,0,0]); prng.rs:12-23.
 * ---------------------------------------------------------------------------------------- */
static inline void philox_round(uint32_t ctr[4], const uint32_t key[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * ctr[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * ctr[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ ctr[1] ^ key[0], n1 = lo1, n2 = hi0 ^ ctr[3] ^ key[1], n3 = lo0;
    ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
}

ORACLE_API void oracle_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t ctr[4] = {ctr_in[0], ctr_in[1], ctr_in[2], ctr_in[3]};
    uint32_t key[2] = {key_in[0], key_in[1]};
    philox_round(ctr, key);
    for (int i = 0; i < 9; ++i) {
        key[0] += 0x9E3779B9u; key[1] += 0xBB67AE85u;
        philox_round(ctr, key);
    }
    memcpy(out, ctr, 16);
}

ORACLE_API float oracle_unit_interval(uint32_t word) {
    uint32_t w = word >> 8;
    if (w < 1) w = 1;
    return (float)w * (1.0f / 16777216.0f);
}

static inline float uniform_float(uint64_t key64, uint32_t offset, uint32_t word) {
    uint32_t ctr[4] = {offset, 0, 0, 0}, key[2] = {(uint32_t)key64, (uint32_t)(key64 >> 32)}, out[4];
    oracle_philox4x32_10(ctr, key, out);
    return oracle_unit_interval(out[word]);
}

ORACLE_API void oracle_revidx(uint32_t logit_idx, uint32_t vocab_size, uint32_t* offset, uint32_t* word) {
    const uint32_t TG = 1024, WPO = 4;
    uint32_t thread_idx = logit_idx % TG;
    uint32_t thread_offset = ((vocab_size + TG * WPO - 1) / (TG * WPO)) * thread_idx;
    uint32_t block_idx = logit_idx / TG;
    *offset = thread_offset + block_idx / WPO;
    *word = block_idx % WPO;
}

ORACLE_API float oracle_gumbel_float(uint64_t key, uint32_t offset, uint32_t word) {
    return -logf(-logf(uniform_float(key, offset, word)));
}

ORACLE_API uint64_t oracle_prng_derive(uint64_t seed, uint64_t index) {
    uint64_t h = seed + index;
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    return h;
}

/* ------------------------------------------------------------------------------------------
 * UnifiedSampling: backends/cpu/kernel/sampling/unified_sampling.rs:22-98.
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint32_t idx; float v; } idxval;
static int cmp_desc(const void* a, const void* b) {
    const idxval* x = (const idxval*)a; const idxval* y = (const idxval*)b;
    /* b.1.partial_cmp(a.1).unwrap_or(Equal).then(a.0.cmp(b.0)) */
    if (y->v < x->v) return -1;
    if (y->v > x->v) return 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

typedef struct {
    const bf16_t* logits;
    uint32_t* output;
    const uint64_t* seeds;     /* NULL = greedy */
    const uint32_t* bitmask;   /* optional */
    int has_temperature; float temperature;
    int has_top_k; uint32_t top_k;
    int has_top_p; float top_p;
    int has_min_p; float min_p;
    uint32_t vocab_size, batch_size;
} oracle_sampling_args;

ORACLE_API void oracle_unified_sampling(const oracle_sampling_args* p) {
    uint32_t V = p->vocab_size;
    float* l = (float*)malloc(sizeof(float) * V);
    idxval* sorted = (p->has_top_k || p->has_top_p || p->has_min_p) ? (idxval*)malloc(sizeof(idxval) * V) : NULL;
    uint32_t words = (V + 31) / 32;
    for (uint32_t b = 0; b < p->batch_size; ++b) {
        for (uint32_t i = 0; i < V; ++i) l[i] = bf2f(p->logits[(size_t)V * b + i]);
        if (p->bitmask) {
            const uint32_t* bm = p->bitmask + (size_t)words * b;
            for (uint32_t i = 0; i < V; ++i)
                if ((bm[i / 32] & (1u << (i % 32))) == 0) l[i] = -INFINITY;
        }
        if (p->has_temperature) {
            float r = 1.0f / p->temperature;
            for (uint32_t i = 0; i < V; ++i) l[i] *= r;
        }
        if (sorted) {
            for (uint32_t i = 0; i < V; ++i) { sorted[i].idx = i; sorted[i].v = l[i]; }
            qsort(sorted, V, sizeof(idxval), cmp_desc); /* total order (ties by index) => same result as a stable sort */
            float lmax = sorted[0].v, norm = 0.0f;
            for (uint32_t i = 0; i < V; ++i) norm += expf(sorted[i].v - lmax);
            for (uint32_t i = 0; i < V; ++i) l[i] = -INFINITY;
            float mass = 0.0f;
            for (uint32_t r = 0; r < V; ++r) {
                float lv = sorted[r].v;
                if ((p->has_top_k && r >= p->top_k) || (p->has_top_p && mass >= p->top_p) ||
                    (p->has_min_p && lv < lmax + logf(p->min_p)))
                    break;
                l[sorted[r].idx] = lv;
                mass += expf(lv - lmax) / norm;
            }
        }
        if (p->seeds) {
            uint64_t seed = p->seeds[b];
            for (uint32_t i = 0; i < V; ++i) {
                uint32_t off, w;
                oracle_revidx(i, V, &off, &w);
                l[i] += oracle_gumbel_float(seed, off, w);
            }
        }
        /* max_by(a.1.partial_cmp(b.1).unwrap_or(Equal).then(b.0.cmp(a.0))): max value, ties -> lowest
         * index; NaN compares Equal and then falls to the index rule (:90-95). */
        uint32_t best = 0;
        for (uint32_t i = 1; i < V; ++i) {
            float a = l[best], c = l[i];
            /* Iterator::max_by keeps the later element when compare(best, cur) != Greater. */
            int ord; /* compare(best, cur) */
            if (a < c) ord = -1; else if (a > c) ord = 1; else ord = 0;
            if (ord == 0) ord = (i < best) ? -1 : 1; /* b.0.cmp(a.0) with a=best, b=cur: cur.idx cmp best.idx reversed */
            if (ord <= 0) best = i;
        }
        p->output[b] = best;
    }
    free(l); free(sorted);
}

/* ------------------------------------------------------------------------------------------
 * DeltaNet decode (Qwen3.5 hybrid layers): backends/cpu/kernel/gdn/
 * conv_update.rs:8-55 and update.rs:13-144, with the parameter/state element types the engine
 * actually allocates and Metal declares (f32 a_log/dt_bias/norm_weight/state; SURVEY.md row
 * a11 documents the reference CPU kernel's mistyped pointers).
 * ---------------------------------------------------------------------------------------- */
ORACLE_API void oracle_delta_net_conv_update(const float* conv_weight, const float* bias, bf16_t* in_out, float* state,
                                             int kernel_size, int conv_dim, int state_stride) {
    int taps = kernel_size - 1;
    for (int c = 0; c < conv_dim; ++c) {
        size_t so = (size_t)c * state_stride, wo = (size_t)c * kernel_size;
        float x = bf2f(in_out[c]);
        float acc = bias ? bias[c] : 0.0f;
        for (int t = 0; t < taps; ++t) acc += state[so + t] * conv_weight[wo + t];
        acc += x * conv_weight[wo + taps];
        int pass;
        in_out[c] = f2bf(act_f32(ACT_SILU, acc, &pass));
        for (int t = 1; t < taps; ++t) state[so + t - 1] = state[so + t];
        state[so + taps - 1] = x;
    }
}

ORACLE_API void oracle_delta_net_update(const bf16_t* in_proj, const float* a_log, const float* dt_bias,
                                        const float* norm_weight, float* state, bf16_t* out, int num_v_heads,
                                        int num_k_heads, int head_k_dim, int head_v_dim, int key_dim, int value_dim,
                                        float norm_epsilon) {
    int conv_dim = 2 * key_dim + value_dim;
    float* q = (float*)malloc(sizeof(float) * head_k_dim);
    float* k = (float*)malloc(sizeof(float) * head_k_dim);
    float* o = (float*)malloc(sizeof(float) * head_v_dim);
    for (int hv = 0; hv < num_v_heads; ++hv) {
        int hk = hv / (num_v_heads / num_k_heads);
        int q_off = hk * head_k_dim, k_off = key_dim + hk * head_k_dim;
        for (int j = 0; j < head_k_dim; ++j) { q[j] = bf2f(in_proj[q_off + j]); k[j] = bf2f(in_proj[k_off + j]); }
        float qn = 0.0f, kn = 0.0f;
        for (int j = 0; j < head_k_dim; ++j) qn += q[j] * q[j];
        for (int j = 0; j < head_k_dim; ++j) kn += k[j] * k[j];
        float qi = 1.0f / sqrtf(qn + 1e-6f), ki = 1.0f / sqrtf(kn + 1e-6f);
        for (int j = 0; j < head_k_dim; ++j) { q[j] *= qi; k[j] *= ki; }
        float qs = 1.0f / sqrtf((float)head_k_dim);
        for (int j = 0; j < head_k_dim; ++j) q[j] *= qs;
        float beta_raw = bf2f(in_proj[conv_dim + value_dim + hv]);
        float beta = 1.0f / (1.0f + expf(-beta_raw));
        float a_raw = bf2f(in_proj[conv_dim + value_dim + num_v_heads + hv]);
        float sp_in = a_raw + dt_bias[hv];
        float sp = sp_in > 20.0f ? sp_in : logf(1.0f + expf(sp_in));
        float g = -expf(a_log[hv]) * sp;
        float decay = expf(g);
        float kq = 0.0f;
        for (int j = 0; j < head_k_dim; ++j) kq += k[j] * q[j];
        for (int i = 0; i < head_v_dim; ++i) {
            float v_i = bf2f(in_proj[2 * key_dim + hv * head_v_dim + i]);
            size_t row = ((size_t)hv * head_v_dim + i) * head_k_dim;
            float sq = 0.0f, sk = 0.0f;
            for (int j = 0; j < head_k_dim; ++j) { float s = state[row + j]; sq += s * q[j]; sk += s * k[j]; }
            float retrieved = decay * sk;
            float delta = beta * (v_i - retrieved);
            o[i] = decay * sq + delta * kq;
            for (int j = 0; j < head_k_dim; ++j) state[row + j] = decay * state[row + j] + k[j] * delta;
        }
        float ss = 0.0f;
        for (int i = 0; i < head_v_dim; ++i) ss += o[i] * o[i];
        float inv_rms = 1.0f / sqrtf(ss / (float)head_v_dim + norm_epsilon);
        for (int i = 0; i < head_v_dim; ++i) {
            float z = bf2f(in_proj[conv_dim + hv * head_v_dim + i]);
            /* activate<f32> on an f32 z_i: z_i is read via to_f32() so no bf16 rounding of silu */
            int pass;
            float zs = act_f32(ACT_SILU, z, &pass);
            out[hv * head_v_dim + i] = f2bf(o[i] * inv_rms * norm_weight[i] * zs);
        }
    }
    free(q); free(k); free(o);
}

ORACLE_API int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
