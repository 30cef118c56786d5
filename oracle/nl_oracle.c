/*
 * nl_oracle.c — CPU restatement of the nanollama Go engine's quantized forward path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (nanollama_b200/, csrc/, host/)
 * may link, import or execute this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, and only as the checker / CPU arm.
 *
 * The Go toolchain is absent from the build image, so the reference engine itself cannot be
 * compiled here; this file restates it operation for operation:
 *   fp16 -> fp32           go/gguf.go:601-636   (half2floatLUT construction)
 *   Q4_0 dequant / matmul  go/quant.go:18-94
 *   Q8_0 dequant / matmul  go/quant.go:100-165
 *   Q5_0 / Q4_K / Q6_K     go/quant.go:171-484
 *   F32 / F16 matmul       go/quant.go:490-563, go/model.go:371-378
 *   RMSNorm / Softmax/SiLU go/quant.go:570-631
 *   RoPE tables, RoPE      go/model.go:346-358, :449-477
 *   embedding lookup       go/model.go:389-446
 *   Forward / Reset        go/model.go:490-631
 *   argmax                 go/main.go:400-408
 * Same fp32 sequential accumulation order, float64 sum of squares, float64 exp.
 * Build with -O2 -ffp-contract=off (Go/amd64 does not fuse multiply-adds).
 * Row-chunk parallelism mirrors go/quant.go:49-71 (serial when rows < 4*workers).
 *
 * Parity pin: the reference holds no golden vectors for this path.  This
 * oracle is pinned instead by tests/test_oracle.py against (1) bytes produced by the reference's
 * own quantizers decoded by gguf-py, (2) logits of nanollama.llama.Llama (fp32) on the same
 * exported GGUF — both generated in the build container by tests/golden/make_golden.py.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

enum { T_F32 = 0, T_F16 = 1, T_Q4_0 = 2, T_Q5_0 = 6, T_Q8_0 = 8, T_Q4_K = 12, T_Q6_K = 14 };

static float h2f_lut[65536];
static int lut_ready = 0;
static int num_workers = 0;

static float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* go/gguf.go:605-631 */
static void build_lut(void) {
    if (lut_ready) return;
    for (int h = 0; h < 65536; h++) {
        uint32_t sign = ((uint32_t)h >> 15) & 1, exp = ((uint32_t)h >> 10) & 0x1F, mant = (uint32_t)h & 0x3FF, f;
        if (exp == 0) {
            if (mant == 0) f = sign << 31;
            else {
                uint32_t e = 1;
                while ((mant & 0x400) == 0) { mant <<= 1; e--; }
                mant &= 0x3FF;
                f = (sign << 31) | ((e + 127 - 15) << 23) | (mant << 13);
            }
        } else if (exp == 0x1F) f = (sign << 31) | 0x7F800000u | (mant << 13);
        else f = (sign << 31) | ((exp - 15 + 127) << 23) | (mant << 13);
        h2f_lut[h] = bits2f(f);
    }
    lut_ready = 1;
    if (num_workers == 0) {
        long n = sysconf(_SC_NPROCESSORS_ONLN);
        num_workers = n > 0 ? (int)n : 1;
    }
}
static inline float h2f(const uint8_t *p) { return h2f_lut[(uint16_t)(p[0] | (p[1] << 8))]; }
static inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

void nlo_init(void) { build_lut(); }
void nlo_set_workers(int n) { build_lut(); num_workers = n > 0 ? n : 1; }
int nlo_get_workers(void) { build_lut(); return num_workers; }
float nlo_half2float(uint16_t h) { build_lut(); return h2f_lut[h]; }

/* ---------------- block dequant ---------------- */
/* go/quant.go:22-31 */
static void dq_q4_0_block(const uint8_t *b, float *out) {
    float d = h2f(b);
    for (int j = 0; j < 16; j++) {
        uint8_t v = b[2 + j];
        out[j] = (float)((int)(v & 0x0F) - 8) * d;
        out[j + 16] = (float)((int)(v >> 4) - 8) * d;
    }
}
/* go/quant.go:103-108 */
static void dq_q8_0_block(const uint8_t *b, float *out) {
    float d = h2f(b);
    for (int j = 0; j < 32; j++) out[j] = (float)(int8_t)b[2 + j] * d;
}
/* go/quant.go:405-420 */
static void dq_q5_0_block(const uint8_t *b, float *out) {
    float d = h2f(b);
    uint32_t qh = le32(b + 2);
    const uint8_t *qs = b + 6;
    for (int j = 0; j < 16; j++) {
        int lo = qs[j] & 0x0F, hi = qs[j] >> 4;
        int q0 = lo | (int)(((qh >> j) & 1) << 4), q1 = hi | (int)(((qh >> (j + 16)) & 1) << 4);
        out[j] = (float)(q0 - 16) * d;
        out[j + 16] = (float)(q1 - 16) * d;
    }
}
/* go/quant.go:285-294 */
static void scale_min_k4(int j, const uint8_t *s, uint8_t *sc, uint8_t *m) {
    if (j < 4) { *sc = s[j] & 63; *m = s[j + 4] & 63; }
    else { *sc = (uint8_t)((s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4)); *m = (uint8_t)((s[j + 4] >> 4) | ((s[j] >> 6) << 4)); }
}
/* go/quant.go:296-323 */
static void dq_q4_k_block(const uint8_t *b, float *out) {
    float d = h2f(b), dmin = h2f(b + 2);
    const uint8_t *scales = b + 4, *qs = b + 16;
    int is = 0, oi = 0, qi = 0;
    for (int j = 0; j < 256; j += 64) {
        uint8_t sc0, m0, sc1, m1v;
        scale_min_k4(is, scales, &sc0, &m0);
        float d1 = d * (float)sc0, m1 = dmin * (float)m0;
        scale_min_k4(is + 1, scales, &sc1, &m1v);
        float d2 = d * (float)sc1, m2 = dmin * (float)m1v;
        for (int l = 0; l < 32; l++) out[oi + l] = d1 * (float)(qs[qi + l] & 0x0F) - m1;
        for (int l = 0; l < 32; l++) out[oi + 32 + l] = d2 * (float)(qs[qi + l] >> 4) - m2;
        qi += 32; oi += 64; is += 2;
    }
}
/* go/quant.go:174-208 (one 256-element super-block) */
static void dq_q6_k_block(const uint8_t *b, float *out) {
    const uint8_t *ql = b, *qh = b + 128, *scales = b + 192;
    float d = h2f(b + 208);
    for (int n128 = 0; n128 < 2; n128++) {
        const uint8_t *qlP = ql + n128 * 64, *qhP = qh + n128 * 32, *scP = scales + n128 * 8;
        float *y = out + n128 * 128;
        for (int l = 0; l < 32; l++) {
            int is = l / 16;
            int q1 = (int)(qlP[l] & 0x0F) | (((int)(qhP[l] >> 0) & 3) << 4);
            int q2 = (int)(qlP[l + 32] & 0x0F) | (((int)(qhP[l] >> 2) & 3) << 4);
            int q3 = (int)(qlP[l] >> 4) | (((int)(qhP[l] >> 4) & 3) << 4);
            int q4 = (int)(qlP[l + 32] >> 4) | (((int)(qhP[l] >> 6) & 3) << 4);
            y[l + 0] = d * (float)(int8_t)scP[is + 0] * (float)(q1 - 32);
            y[l + 32] = d * (float)(int8_t)scP[is + 2] * (float)(q2 - 32);
            y[l + 64] = d * (float)(int8_t)scP[is + 4] * (float)(q3 - 32);
            y[l + 96] = d * (float)(int8_t)scP[is + 6] * (float)(q4 - 32);
        }
    }
}

static int blk_elems(int t) { return (t == T_F32 || t == T_F16) ? 1 : (t == T_Q4_K || t == T_Q6_K) ? 256 : 32; }
static int blk_bytes(int t) {
    switch (t) { case T_F32: return 4; case T_F16: return 2; case T_Q4_0: return 18; case T_Q5_0: return 22;
                 case T_Q8_0: return 34; case T_Q4_K: return 144; case T_Q6_K: return 210; default: return 0; }
}
int64_t nlo_tensor_bytes(int type, int64_t n) { int be = blk_elems(type), bb = blk_bytes(type); return bb ? (n / be) * bb : -1; }

/* Dequantise n elements (getF32Tensor / Dequant*: go/model.go:268-303).  returns 0 or -1 */
int nlo_dequant(int type, const uint8_t *src, int64_t n, float *dst) {
    build_lut();
    switch (type) {
    case T_F32: for (int64_t i = 0; i < n; i++) dst[i] = bits2f(le32(src + 4 * i)); return 0;
    case T_F16: for (int64_t i = 0; i < n; i++) dst[i] = h2f(src + 2 * i); return 0;
    case T_Q4_0: for (int64_t b = 0; b < n / 32; b++) dq_q4_0_block(src + 18 * b, dst + 32 * b); return 0;
    case T_Q5_0: for (int64_t b = 0; b < n / 32; b++) dq_q5_0_block(src + 22 * b, dst + 32 * b); return 0;
    case T_Q8_0: for (int64_t b = 0; b < n / 32; b++) dq_q8_0_block(src + 34 * b, dst + 32 * b); return 0;
    case T_Q4_K: for (int64_t b = 0; b < n / 256; b++) dq_q4_k_block(src + 144 * b, dst + 256 * b); return 0;
    case T_Q6_K: for (int64_t b = 0; b < n / 256; b++) dq_q6_k_block(src + 210 * b, dst + 256 * b); return 0;
    default: return -1;
    }
}

/* ---------------- matmul ranges ---------------- */
typedef struct { float *out; const uint8_t *w; const float *x; int start, end, rows, cols, type; } mm_job;

/* go/quant.go:74-94 */
static void mm_q4_0(const mm_job *j) {
    int bpr = j->cols / 32; int64_t rb = (int64_t)bpr * 18;
    for (int i = j->start; i < j->end; i++) {
        const uint8_t *row = j->w + (int64_t)i * rb; float sum = 0.0f;
        for (int b = 0; b < bpr; b++) {
            const uint8_t *blk = row + b * 18; float d = h2f(blk); const float *x = j->x + b * 32; float dot = 0.0f;
            for (int k = 0; k < 16; k++) {
                uint8_t bv = blk[2 + k];
                float v0 = (float)((int)(bv & 0x0F) - 8), v1 = (float)((int)(bv >> 4) - 8);
                dot += v0 * x[k] + v1 * x[k + 16];
            }
            sum += dot * d;
        }
        j->out[i] = sum;
    }
}
/* go/quant.go:149-165 */
static void mm_q8_0(const mm_job *j) {
    int bpr = j->cols / 32; int64_t rb = (int64_t)bpr * 34;
    for (int i = j->start; i < j->end; i++) {
        const uint8_t *row = j->w + (int64_t)i * rb; float sum = 0.0f;
        for (int b = 0; b < bpr; b++) {
            const uint8_t *blk = row + b * 34; float d = h2f(blk); const float *x = j->x + b * 32; float dot = 0.0f;
            for (int k = 0; k < 32; k++) dot += (float)(int8_t)blk[2 + k] * x[k];
            sum += dot * d;
        }
        j->out[i] = sum;
    }
}
/* go/quant.go:461-484 */
static void mm_q5_0(const mm_job *j) {
    int bpr = j->cols / 32; int64_t rb = (int64_t)bpr * 22;
    for (int r = j->start; r < j->end; r++) {
        const uint8_t *row = j->w + (int64_t)r * rb; float sum = 0.0f;
        for (int b = 0; b < bpr; b++) {
            const uint8_t *blk = row + b * 22; float d = h2f(blk); uint32_t qh = le32(blk + 2); const uint8_t *qs = blk + 6;
            const float *x = j->x + b * 32;
            for (int k = 0; k < 16; k++) {
                int lo = qs[k] & 0x0F, hi = qs[k] >> 4;
                int q0 = lo | (int)(((qh >> k) & 1) << 4), q1 = hi | (int)(((qh >> (k + 16)) & 1) << 4);
                sum += (float)(q0 - 16) * d * x[k];
                sum += (float)(q1 - 16) * d * x[k + 16];
            }
        }
        j->out[r] = sum;
    }
}
/* go/quant.go:364-396 */
static void mm_q4_k(const mm_job *j) {
    int bpr = j->cols / 256; int64_t rb = (int64_t)bpr * 144;
    for (int r = j->start; r < j->end; r++) {
        const uint8_t *row = j->w + (int64_t)r * rb; float sum = 0.0f;
        for (int b = 0; b < bpr; b++) {
            const uint8_t *blk = row + b * 144; float d = h2f(blk), dmin = h2f(blk + 2);
            const uint8_t *scales = blk + 4, *qs = blk + 16; const float *x = j->x + b * 256;
            int is = 0, qi = 0;
            for (int jj = 0; jj < 256; jj += 64) {
                uint8_t sc0, m0, sc1, m1v;
                scale_min_k4(is, scales, &sc0, &m0);
                float d1 = d * (float)sc0, m1 = dmin * (float)m0;
                scale_min_k4(is + 1, scales, &sc1, &m1v);
                float d2 = d * (float)sc1, m2 = dmin * (float)m1v;
                for (int l = 0; l < 32; l++) sum += (d1 * (float)(qs[qi + l] & 0x0F) - m1) * x[jj + l];
                for (int l = 0; l < 32; l++) sum += (d2 * (float)(qs[qi + l] >> 4) - m2) * x[jj + 32 + l];
                qi += 32; is += 2;
            }
        }
        j->out[r] = sum;
    }
}
/* go/quant.go:239-276 */
static void mm_q6_k(const mm_job *j) {
    int bpr = j->cols / 256; int64_t rb = (int64_t)bpr * 210;
    for (int r = j->start; r < j->end; r++) {
        const uint8_t *row = j->w + (int64_t)r * rb; float sum = 0.0f;
        for (int b = 0; b < bpr; b++) {
            const uint8_t *blk = row + b * 210, *ql = blk, *qh = blk + 128, *scales = blk + 192;
            float d = h2f(blk + 208); const float *xb = j->x + b * 256;
            for (int n128 = 0; n128 < 2; n128++) {
                const uint8_t *qlP = ql + n128 * 64, *qhP = qh + n128 * 32, *scP = scales + n128 * 8;
                const float *x = xb + n128 * 128;
                for (int l = 0; l < 32; l++) {
                    int is = l / 16;
                    int q1 = (int)(qlP[l] & 0x0F) | (((int)(qhP[l] >> 0) & 3) << 4);
                    int q2 = (int)(qlP[l + 32] & 0x0F) | (((int)(qhP[l] >> 2) & 3) << 4);
                    int q3 = (int)(qlP[l] >> 4) | (((int)(qhP[l] >> 4) & 3) << 4);
                    int q4 = (int)(qlP[l + 32] >> 4) | (((int)(qhP[l] >> 6) & 3) << 4);
                    float s0 = d * (float)(int8_t)scP[is + 0], s2 = d * (float)(int8_t)scP[is + 2];
                    float s4 = d * (float)(int8_t)scP[is + 4], s6 = d * (float)(int8_t)scP[is + 6];
                    sum += s0 * (float)(q1 - 32) * x[l + 0];
                    sum += s2 * (float)(q2 - 32) * x[l + 32];
                    sum += s4 * (float)(q3 - 32) * x[l + 64];
                    sum += s6 * (float)(q4 - 32) * x[l + 96];
                }
            }
        }
        j->out[r] = sum;
    }
}
/* go/quant.go:553-563 */
static void mm_f16(const mm_job *j) {
    for (int i = j->start; i < j->end; i++) {
        const uint8_t *row = j->w + (int64_t)i * j->cols * 2; float sum = 0.0f;
        for (int k = 0; k < j->cols; k++) sum += h2f(row + 2 * k) * j->x[k];
        j->out[i] = sum;
    }
}
/* go/quant.go:516-525 (+ byte decode of go/model.go:371-378) */
static void mm_f32(const mm_job *j) {
    for (int i = j->start; i < j->end; i++) {
        const uint8_t *row = j->w + (int64_t)i * j->cols * 4; float sum = 0.0f;
        for (int k = 0; k < j->cols; k++) sum += bits2f(le32(row + 4 * k)) * j->x[k];
        j->out[i] = sum;
    }
}
static void *mm_thread(void *p) {
    const mm_job *j = (const mm_job *)p;
    switch (j->type) {
    case T_Q4_0: mm_q4_0(j); break; case T_Q8_0: mm_q8_0(j); break; case T_Q5_0: mm_q5_0(j); break;
    case T_Q4_K: mm_q4_k(j); break; case T_Q6_K: mm_q6_k(j); break; case T_F16: mm_f16(j); break;
    case T_F32: mm_f32(j); break; default: break;
    }
    return NULL;
}

/* matmulDispatch (go/model.go:361-386) with the fork-join of go/quant.go:49-71.
 * Row results do not depend on the chunking (each row is a sequential sum). returns 0 or -1 */
int nlo_matmul(float *out, const uint8_t *w, int type, const float *x, int rows, int cols) {
    build_lut();
    if (blk_bytes(type) == 0) return -1;  /* reference prints a WARNING and leaves out stale */
    mm_job base = { out, w, x, 0, rows, rows, cols, type };
    if (rows < num_workers * 4 || num_workers == 1) { mm_thread(&base); return 0; }
    pthread_t th[256]; mm_job jobs[256];
    int workers = num_workers > 256 ? 256 : num_workers, nt = 0;
    int chunk = (rows + workers - 1) / workers;
    for (int wk = 0; wk < workers; wk++) {
        int s = wk * chunk, e = s + chunk; if (e > rows) e = rows; if (s >= e) break;
        jobs[nt] = base; jobs[nt].start = s; jobs[nt].end = e;
        pthread_create(&th[nt], NULL, mm_thread, &jobs[nt]); nt++;
    }
    for (int i = 0; i < nt; i++) pthread_join(th[i], NULL);
    return 0;
}

/* ---------------- math utilities ---------------- */
/* go/quant.go:597-607 (out may alias x => RMSNorm :570-580) */
void nlo_rmsnorm(float *out, const float *x, const float *w, int n, float eps) {
    double ss = 0.0;
    for (int i = 0; i < n; i++) ss += (double)x[i] * (double)x[i];
    float inv = (float)(1.0 / sqrt(ss / (double)n + (double)eps));
    for (int i = 0; i < n; i++) out[i] = x[i] * inv * w[i];
}
/* go/quant.go:584-594 */
void nlo_rmsnorm_bare(float *x, int n, float eps) {
    double ss = 0.0;
    for (int i = 0; i < n; i++) ss += (double)x[i] * (double)x[i];
    float inv = (float)(1.0 / sqrt(ss / (double)n + (double)eps));
    for (int i = 0; i < n; i++) x[i] *= inv;
}
/* go/quant.go:610-626 */
void nlo_softmax(float *x, int n) {
    float mx = x[0];
    for (int i = 1; i < n; i++) if (x[i] > mx) mx = x[i];
    float sum = 0.0f;
    for (int i = 0; i < n; i++) { x[i] = (float)exp((double)(x[i] - mx)); sum += x[i]; }
    float inv = 1.0f / sum;
    for (int i = 0; i < n; i++) x[i] *= inv;
}
/* go/quant.go:629-631 */
float nlo_silu(float x) { return x / (1.0f + (float)exp((double)(-x))); }
/* go/main.go:400-408 */
int nlo_argmax(const float *l, int n) { int best = 0; for (int i = 1; i < n; i++) if (l[i] > l[best]) best = i; return best; }

/* ---------------- model ---------------- */
typedef struct {
    int32_t n_layers, embed_dim, n_heads, n_kv_heads, head_dim, vocab_size, seq_len, interm_size;
    float rms_norm_eps, rope_theta;
    int32_t qk_norm, rope_conjugate;
} nlo_config;

typedef struct { const uint8_t *p; int type; } wt;
typedef struct { float *attn_norm, *ffn_norm; wt wq, wk, wv, wo, wgate, wup, wdown; float *bq, *bk, *bv, *bo; } layer_w;

typedef struct {
    nlo_config c;
    wt tok_embd, output; float *output_norm; layer_w *L;
    float *gamma; const int32_t *gamma_map; /* optional dense [n_gamma, dim] + token->row (-1 none) */
    float *x, *xb, *xb2, *hb, *hb2, *q, *k, *v, *att, *logits, *kc, *vc, *cosc, *sinc, *emb;
} nlo_model;

/* allocState + precomputeRoPE: go/model.go:324-358 */
nlo_model *nlo_model_new(const nlo_config *c) {
    build_lut();
    nlo_model *m = (nlo_model *)calloc(1, sizeof(nlo_model));
    m->c = *c;
    if (m->c.head_dim == 0 && m->c.n_heads > 0) m->c.head_dim = m->c.embed_dim / m->c.n_heads; /* model.go:140 */
    if (m->c.seq_len > 2048) m->c.seq_len = 2048;                                              /* model.go:145 */
    int dim = m->c.embed_dim, hd = m->c.head_dim, kvd = m->c.n_kv_heads * hd, S = m->c.seq_len;
    m->L = (layer_w *)calloc((size_t)m->c.n_layers, sizeof(layer_w));
    m->x = calloc(dim, 4); m->xb = calloc(dim, 4); m->xb2 = calloc(dim, 4); m->emb = calloc(dim, 4);
    m->hb = calloc(m->c.interm_size, 4); m->hb2 = calloc(m->c.interm_size, 4);
    m->q = calloc((size_t)m->c.n_heads * hd, 4); m->k = calloc(kvd, 4); m->v = calloc(kvd, 4);
    m->att = calloc((size_t)m->c.n_heads * S, 4); m->logits = calloc(m->c.vocab_size > 0 ? m->c.vocab_size : 1, 4);
    m->kc = calloc((size_t)m->c.n_layers * S * kvd, 4); m->vc = calloc((size_t)m->c.n_layers * S * kvd, 4);
    int half = hd / 2;
    m->cosc = calloc((size_t)S * half, 4); m->sinc = calloc((size_t)S * half, 4);
    double theta = (double)m->c.rope_theta;
    for (int pos = 0; pos < S; pos++)
        for (int i = 0; i < half; i++) {
            double freq = 1.0 / pow(theta, (double)(2 * i) / (double)hd);
            double angle = (double)pos * freq;
            m->cosc[pos * half + i] = (float)cos(angle);
            m->sinc[pos * half + i] = (float)sin(angle);
        }
    return m;
}
void nlo_model_free(nlo_model *m) {
    if (!m) return;
    for (int i = 0; i < m->c.n_layers; i++) { free(m->L[i].attn_norm); free(m->L[i].ffn_norm); free(m->L[i].bq); free(m->L[i].bk); free(m->L[i].bv); free(m->L[i].bo); }
    free(m->L); free(m->output_norm); free(m->gamma);
    free(m->x); free(m->xb); free(m->xb2); free(m->emb); free(m->hb); free(m->hb2); free(m->q); free(m->k); free(m->v);
    free(m->att); free(m->logits); free(m->kc); free(m->vc); free(m->cosc); free(m->sinc); free(m);
}

/* Tensor slots, same order as nanollama_cuda.h.  Matrices alias caller memory (like the Go slices
 * into GGUFFile.TensorData, gguf.go:573); vectors are dequantised to fp32 copies (getF32Tensor). */
enum { S_TOK_EMBD = 0, S_OUTPUT_NORM = 1, S_OUTPUT = 2, S_ATTN_NORM = 3, S_FFN_NORM = 4, S_WQ = 5, S_WK = 6, S_WV = 7, S_WO = 8,
       S_WGATE = 9, S_WUP = 10, S_WDOWN = 11, S_BQ = 12, S_BK = 13, S_BV = 14, S_BO = 15 };

int nlo_model_set_tensor(nlo_model *m, int slot, int layer, int type, const uint8_t *data, int64_t n_elems) {
    if (blk_bytes(type) == 0) return -1;
    float **vec = NULL; wt *mat = NULL;
    layer_w *l = (layer >= 0 && layer < m->c.n_layers) ? &m->L[layer] : NULL;
    switch (slot) {
    case S_TOK_EMBD: mat = &m->tok_embd; break; case S_OUTPUT: mat = &m->output; break;
    case S_OUTPUT_NORM: vec = &m->output_norm; break;
    default:
        if (!l) return -1;
        switch (slot) {
        case S_ATTN_NORM: vec = &l->attn_norm; break; case S_FFN_NORM: vec = &l->ffn_norm; break;
        case S_WQ: mat = &l->wq; break; case S_WK: mat = &l->wk; break; case S_WV: mat = &l->wv; break; case S_WO: mat = &l->wo; break;
        case S_WGATE: mat = &l->wgate; break; case S_WUP: mat = &l->wup; break; case S_WDOWN: mat = &l->wdown; break;
        case S_BQ: vec = &l->bq; break; case S_BK: vec = &l->bk; break; case S_BV: vec = &l->bv; break; case S_BO: vec = &l->bo; break;
        default: return -1;
        }
    }
    if (mat) { mat->p = data; mat->type = type; return 0; }
    free(*vec); *vec = (float *)malloc((size_t)n_elems * 4);
    return nlo_dequant(type, data, n_elems, *vec);
}
/* Gamma: dense rows [n_rows, dim] fp32 + token->row map (go/gamma.go:272-290 adds row to the embedding) */
void nlo_model_set_gamma(nlo_model *m, const float *rows, int n_rows, const int32_t *token_to_row) {
    free(m->gamma); m->gamma = NULL; m->gamma_map = NULL;
    if (!rows) return;
    m->gamma = (float *)malloc((size_t)n_rows * m->c.embed_dim * 4);
    memcpy(m->gamma, rows, (size_t)n_rows * m->c.embed_dim * 4);
    m->gamma_map = token_to_row;
}

/* go/model.go:389-446 */
static void embed_lookup(float *out, const wt *e, int token, int dim) {
    int be = blk_elems(e->type), bb = blk_bytes(e->type);
    if (bb == 0) { memset(out, 0, (size_t)dim * 4); return; }
    int64_t row_bytes = (int64_t)(dim / be) * bb;
    nlo_dequant(e->type, e->p + (int64_t)token * row_bytes, dim, out);
}
/* go/model.go:449-477 */
static void rope(float *v, int pos, const nlo_model *m, int conj) {
    int half = m->c.head_dim / 2; const float *cc = m->cosc + pos * half, *ss = m->sinc + pos * half;
    for (int i = 0; i < half; i++) {
        float x0 = v[i], x1 = v[i + half], c = cc[i], s = ss[i];
        if (!conj) { v[i] = x0 * c - x1 * s; v[i + half] = x0 * s + x1 * c; }
        else { v[i] = x0 * c + x1 * s; v[i + half] = -x0 * s + x1 * c; }
    }
}
static void add_bias(float *o, const float *b, int n) { if (b) for (int i = 0; i < n; i++) o[i] += b[i]; }

/* go/model.go:490-620 */
int nlo_forward(nlo_model *m, int token, int pos) {
    const nlo_config *c = &m->c;
    int dim = c->embed_dim, hd = c->head_dim, kvd = c->n_kv_heads * hd, S = c->seq_len, group = c->n_heads / c->n_kv_heads;
    if (token < 0 || token >= c->vocab_size || pos < 0 || pos >= S) return -1; /* Go would panic on the slice index */
    embed_lookup(m->emb, &m->tok_embd, token, dim);
    if (m->gamma && m->gamma_map && m->gamma_map[token] >= 0) {
        const float *g = m->gamma + (int64_t)m->gamma_map[token] * dim;
        for (int i = 0; i < dim; i++) m->emb[i] += g[i];
    }
    memcpy(m->x, m->emb, (size_t)dim * 4);
    float attn_scale = (float)(1.0 / sqrt((double)hd));
    for (int layer = 0; layer < c->n_layers; layer++) {
        layer_w *l = &m->L[layer];
        nlo_rmsnorm(m->xb, m->x, l->attn_norm, dim, c->rms_norm_eps);
        nlo_matmul(m->q, l->wq.p, l->wq.type, m->xb, c->n_heads * hd, dim);
        nlo_matmul(m->k, l->wk.p, l->wk.type, m->xb, kvd, dim);
        nlo_matmul(m->v, l->wv.p, l->wv.type, m->xb, kvd, dim);
        add_bias(m->q, l->bq, c->n_heads * hd); add_bias(m->k, l->bk, kvd); add_bias(m->v, l->bv, kvd);
        for (int h = 0; h < c->n_heads; h++) rope(m->q + h * hd, pos, m, c->rope_conjugate);
        for (int h = 0; h < c->n_kv_heads; h++) rope(m->k + h * hd, pos, m, c->rope_conjugate);
        if (c->qk_norm) {
            for (int h = 0; h < c->n_heads; h++) nlo_rmsnorm_bare(m->q + h * hd, hd, c->rms_norm_eps);
            for (int h = 0; h < c->n_kv_heads; h++) nlo_rmsnorm_bare(m->k + h * hd, hd, c->rms_norm_eps);
        }
        int64_t loff = (int64_t)layer * S * kvd;
        memcpy(m->kc + loff + (int64_t)pos * kvd, m->k, (size_t)kvd * 4);
        memcpy(m->vc + loff + (int64_t)pos * kvd, m->v, (size_t)kvd * 4);
        for (int h = 0; h < c->n_heads; h++) {
            int kvh = h / group; const float *qh = m->q + h * hd; float *att = m->att + (int64_t)h * S;
            for (int t = 0; t <= pos; t++) {
                const float *kk = m->kc + loff + (int64_t)t * kvd + kvh * hd; float dot = 0.0f;
                for (int d = 0; d < hd; d++) dot += qh[d] * kk[d];
                att[t] = dot * attn_scale;
            }
            nlo_softmax(att, pos + 1);
            float *o = m->xb2 + h * hd;
            for (int d = 0; d < hd; d++) o[d] = 0.0f;
            for (int t = 0; t <= pos; t++) {
                float a = att[t]; const float *vv = m->vc + loff + (int64_t)t * kvd + kvh * hd;
                for (int d = 0; d < hd; d++) o[d] += a * vv[d];
            }
        }
        nlo_matmul(m->xb, l->wo.p, l->wo.type, m->xb2, dim, dim);
        add_bias(m->xb, l->bo, dim);
        for (int i = 0; i < dim; i++) m->x[i] += m->xb[i];
        nlo_rmsnorm(m->xb, m->x, l->ffn_norm, dim, c->rms_norm_eps);
        nlo_matmul(m->hb, l->wgate.p, l->wgate.type, m->xb, c->interm_size, dim);
        nlo_matmul(m->hb2, l->wup.p, l->wup.type, m->xb, c->interm_size, dim);
        for (int i = 0; i < c->interm_size; i++) m->hb[i] = nlo_silu(m->hb[i]) * m->hb2[i];
        nlo_matmul(m->xb, l->wdown.p, l->wdown.type, m->hb, dim, c->interm_size);
        for (int i = 0; i < dim; i++) m->x[i] += m->xb[i];
    }
    nlo_rmsnorm(m->x, m->x, m->output_norm, dim, c->rms_norm_eps);
    nlo_matmul(m->logits, m->output.p, m->output.type, m->x, c->vocab_size, dim);
    return 0;
}
/* go/model.go:623-631 */
void nlo_reset(nlo_model *m) {
    size_t n = (size_t)m->c.n_layers * m->c.seq_len * m->c.n_kv_heads * m->c.head_dim;
    memset(m->kc, 0, n * 4); memset(m->vc, 0, n * 4);
}
float *nlo_logits(nlo_model *m) { return m->logits; }
float *nlo_state_x(nlo_model *m) { return m->x; }
float *nlo_key_cache(nlo_model *m) { return m->kc; }
float *nlo_value_cache(nlo_model *m) { return m->vc; }
float *nlo_rope_cos(nlo_model *m) { return m->cosc; }
float *nlo_rope_sin(nlo_model *m) { return m->sinc; }
/* ---- sampling step of Engine.Generate (go/main.go:177-197, :294-398).  r01 = the value rng.Float32() delivers (the Go generator itself
 * is not restated: the callers pass the same number to both sides).  Plain restatements: the O(vocab * k) insertion of sampleTopK, a
 * full sort in sampleTopP (the reference's sort.Slice is unstable; ties are broken by index here). ---- */
void nlo_rep_penalty(float *logits, int vocab, const int32_t *recent, int n_recent, float penalty) {   /* go/main.go:177-187 */
    if (!(penalty > 1.0f)) return;
    for (int i = 0; i < n_recent; i++) {
        int tok = recent[i];
        if (tok >= 0 && tok < vocab) {
            if (logits[tok] > 0) logits[tok] /= penalty; else logits[tok] *= penalty;
        }
    }
}
int nlo_sample_top_k(const float *logits, int vocab, float temp, int top_k, float r01) {   /* go/main.go:294-343 */
    if (temp <= 0 || top_k < 1) return nlo_argmax(logits, vocab);   /* (top_k < 1 panics in the reference) */
    if (top_k > vocab) top_k = vocab;
    int *idx = malloc(sizeof(int) * (size_t)top_k);
    float *val = malloc(sizeof(float) * (size_t)top_k), *probs = malloc(sizeof(float) * (size_t)top_k);
    for (int i = 0; i < top_k; i++) { idx[i] = -1; val[i] = -1e30f; }
    for (int i = 0; i < vocab; i++) {
        if (logits[i] > val[top_k - 1]) {
            idx[top_k - 1] = i; val[top_k - 1] = logits[i];
            for (int j = top_k - 1; j > 0 && val[j] > val[j - 1]; j--) {
                int ti = idx[j]; idx[j] = idx[j - 1]; idx[j - 1] = ti;
                float tv = val[j]; val[j] = val[j - 1]; val[j - 1] = tv;
            }
        }
    }
    float maxv = val[0], sum = 0.f;
    int n = 0;
    for (int i = 0; i < top_k; i++) {
        if (idx[i] < 0) break;
        probs[i] = (float)exp((double)((val[i] - maxv) / temp));
        sum += probs[i];
        n = i + 1;
    }
    for (int i = n; i < top_k; i++) probs[i] = 0.f;
    float r = r01 * sum, cdf = 0.f;
    int res = n > 0 ? idx[0] : 0;
    for (int i = 0; i < top_k; i++) {
        cdf += probs[i];
        if (r <= cdf) { res = idx[i]; break; }
    }
    free(idx); free(val); free(probs);
    return res;
}
typedef struct { int idx; float prob; } idx_prob;
static int cmp_prob_desc(const void *a, const void *b) {
    const idx_prob *x = a, *y = b;
    if (x->prob > y->prob) return -1;
    if (x->prob < y->prob) return 1;
    return x->idx - y->idx;
}
int nlo_sample_top_p(const float *logits, int vocab, float temp, float top_p, float r01) {   /* go/main.go:346-398 */
    if (temp <= 0) return nlo_argmax(logits, vocab);
    float maxv = logits[0];
    for (int i = 1; i < vocab; i++) if (logits[i] > maxv) maxv = logits[i];
    idx_prob *c = malloc(sizeof(idx_prob) * (size_t)vocab);
    float sum = 0.f;
    for (int i = 0; i < vocab; i++) {
        float p = (float)exp((double)((logits[i] - maxv) / temp));
        c[i].idx = i; c[i].prob = p;
        sum += p;
    }
    float inv = 1.0f / sum;
    for (int i = 0; i < vocab; i++) c[i].prob *= inv;
    qsort(c, (size_t)vocab, sizeof(idx_prob), cmp_prob_desc);
    float cum = 0.f;
    int res = c[0].idx;
    for (int i = 0; i < vocab; i++) {
        cum += c[i].prob;
        if (cum >= top_p) {
            float r = r01 * cum, cdf = 0.f;
            for (int j = 0; j <= i; j++) {
                cdf += c[j].prob;
                if (r <= cdf) { res = c[j].idx; break; }
            }
            break;
        }
    }
    free(c);
    return res;
}


/* Greedy loop of Engine.Generate with temp<=0, rep-penalty 1.0 (go/main.go:152-230): prefill token by token
 * (stops at seq_len-1), then argmax / forward until n_new tokens, EOS, or pos reaches seq_len.
 * margins (optional) receives top1-top2 logit gap per generated token.  returns number generated. */
int nlo_generate_greedy(nlo_model *m, const int32_t *prompt, int n_prompt, int n_new, int eos_id, int32_t *out, float *margins) {
    nlo_reset(m);
    int pos = 0;
    for (int i = 0; i < n_prompt; i++) { if (nlo_forward(m, prompt[i], pos)) return -1; pos++; if (pos >= m->c.seq_len - 1) break; }
    int n = 0;
    for (int i = 0; i < n_new; i++) {
        int next = nlo_argmax(m->logits, m->c.vocab_size);
        if (margins) {
            float top = m->logits[next], second = -INFINITY;
            for (int j = 0; j < m->c.vocab_size; j++) if (j != next && m->logits[j] > second) second = m->logits[j];
            margins[n] = top - second;
        }
        out[n++] = next;
        if (next == eos_id) break;
        if (nlo_forward(m, next, pos)) return -1;
        pos++;
        if (pos >= m->c.seq_len) break;
    }
    return n;
}
