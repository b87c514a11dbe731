/*
 * Plain-C CPU restatement of the fp8-mps-metal hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (fp8-mps-metal_b200/) never does.
 *
 * Each function restates one piece of the reference (audiohacking/fp8-mps-metal),
 * loop for loop, with the same fp32 arithmetic and summation order as the shader:
 *
 *   fp8o_decode        fp8_matmul.metal:19-40
 *   fp8o_decode_e5m2   (no reference counterpart: the e5m2 format's definition, see below)
 *   fp8o_encode        fp8_matmul.metal:44-92   (floor(log2 v) taken exactly, like the
 *                                                reference's test_fp8_correctness.py:84)
 *   fp8o_to_half       fp8_matmul.metal:215-223 + fp8_mps_native.py:121-122
 *   fp8o_encode_*      fp8_matmul.metal:228-236 + fp8_mps_native.py:142
 *   fp8o_vecmat        fp8_matmul.metal:155-210 (32 lanes x 4 bytes, then the lane sum)
 *   fp8o_matmul        fp8_matmul.metal:99-147  (4-way unrolled K loop)
 *
 * Parity pin: tests/test_oracle_golden.py checks it against tests/golden/ (fixtures
 * generated from the reference's own Python codec) and against oracle/fp8_oracle.py.
 *
 * Threads: pthreads (this image has no libgomp); fp8o_set_threads(n), default = online CPUs.
 * Build: make -C oracle     (gcc -O2 -pthread -ffp-contract=off)
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ---- minimal pthread parallel-for: body(ctx, begin, end) over [0, n) in equal slabs */
typedef void (*fp8o_body_t)(void* ctx, ptrdiff_t begin, ptrdiff_t end);
typedef struct { fp8o_body_t body; void* ctx; ptrdiff_t begin, end; } fp8o_job_t;
static int g_threads = 0;

static void* fp8o_thread_main(void* p)
{
    fp8o_job_t* j = (fp8o_job_t*)p;
    j->body(j->ctx, j->begin, j->end);
    return NULL;
}

int fp8o_num_threads(void)
{
    if (g_threads <= 0) {
        long n = sysconf(_SC_NPROCESSORS_ONLN);
        g_threads = n > 0 ? (int)n : 1;
        if (g_threads > 256) g_threads = 256;
    }
    return g_threads;
}

void fp8o_set_threads(int n) { g_threads = n > 0 ? (n > 256 ? 256 : n) : 0; }

static void parallel_for(ptrdiff_t n, ptrdiff_t min_serial, fp8o_body_t body, void* ctx)
{
    int nt = fp8o_num_threads();
    if (n < min_serial || nt == 1) { body(ctx, 0, n); return; }
    if ((ptrdiff_t)nt > n) nt = (int)n;
    pthread_t th[256];
    fp8o_job_t jobs[256];
    ptrdiff_t chunk = (n + nt - 1) / nt;
    int started = 0;
    for (int t = 0; t < nt; ++t) {
        ptrdiff_t b = (ptrdiff_t)t * chunk, e = b + chunk > n ? n : b + chunk;
        if (b >= e) break;
        jobs[t].body = body; jobs[t].ctx = ctx; jobs[t].begin = b; jobs[t].end = e;
        if (t == nt - 1 || e == n) { body(ctx, b, e); break; }  /* last slab on the caller */
        if (pthread_create(&th[t], NULL, fp8o_thread_main, &jobs[t]) != 0) { body(ctx, b, e); th[t] = 0; }
        started = t + 1;
    }
    for (int t = 0; t < started; ++t) if (th[t]) pthread_join(th[t], NULL);
}

static float g_lut[256];
static int g_lut_ready = 0;

/* fp8_matmul.metal:19-40 */
float fp8o_decode(uint8_t bits)
{
    if ((bits & 0x7F) == 0x7F) return 0.0f;                 /* :21 NaN -> 0 */
    unsigned sign = (bits >> 7) & 1;                        /* :23 */
    unsigned exp_bits = (bits >> 3) & 0xF;                  /* :24 */
    unsigned mant_bits = bits & 0x7;                        /* :25 */
    float value;
    if (exp_bits == 0) {
        value = (float)mant_bits / 8.0f * (1.0f / 64.0f);   /* :31 */
    } else {
        float mantissa = 1.0f + (float)mant_bits / 8.0f;    /* :34 */
        int exponent = (int)exp_bits - 7;                   /* :35 */
        value = mantissa * ldexpf(1.0f, exponent);          /* :36 exp2 */
    }
    return sign ? -value : value;                           /* :39 */
}

/* float8_e5m2: not a reference codec (the reference mis-routes e5m2 tensors into fp8o_decode's format).
 * The format's own definition, arithmetic form: sign(1) exponent(5, bias 15) mantissa(2), IEEE inf / NaN at
 * exponent 31.  Pinned against PyTorch's CPU cast by tests/test_oracle_golden.py. */
float fp8o_decode_e5m2(uint8_t bits)
{
    unsigned sign = (bits >> 7) & 1, e = (bits >> 2) & 0x1F, m = bits & 0x3;
    float value;
    if (e == 31) value = m ? NAN : INFINITY;
    else if (e == 0) value = (float)m / 4.0f * ldexpf(1.0f, -14);
    else value = (1.0f + (float)m / 4.0f) * ldexpf(1.0f, (int)e - 15);
    return sign ? -value : value;
}

static void ensure_lut(void)
{
    if (!g_lut_ready) {
        for (int i = 0; i < 256; ++i) g_lut[i] = fp8o_decode((uint8_t)i);
        g_lut_ready = 1;
    }
}

/* fp8_matmul.metal:44-92 */
uint8_t fp8o_encode(float val)
{
    if (val != val) return 0x7F;           /* NaN: undefined in the reference; build-defined */
    unsigned sign = 0;
    if (val < 0.0f) { sign = 1; val = -val; }               /* :46-49 */
    if (val >= 448.0f) return (uint8_t)((sign << 7) | 0x7E);/* :53-55 */
    if (val < (1.0f / 512.0f)) return (uint8_t)(sign << 7); /* :58-60 */
    if (val < (1.0f / 64.0f)) {                             /* :64-70 */
        float mant_f = val * 512.0f;
        unsigned mant = (unsigned)rintf(mant_f);
        if (mant > 7u) mant = 7u;
        return (uint8_t)((sign << 7) | mant);
    }
    int exp_val = ilogbf(val);                              /* :73 floor(log2), exact */
    if (exp_val < -6) exp_val = -6;                         /* :75 */
    if (exp_val > 8) exp_val = 8;
    float mantissa = val / ldexpf(1.0f, exp_val);           /* :77 */
    float mant_f = (mantissa - 1.0f) * 8.0f;                /* :79 */
    unsigned mant = (unsigned)rintf(mant_f);                /* :80 */
    if (mant > 7u) mant = 7u;                               /* :81 */
    unsigned exp_bits = (unsigned)(exp_val + 7);            /* :83 */
    if (exp_bits < 1u) exp_bits = 1u;                       /* :84 */
    if (exp_bits > 15u) exp_bits = 15u;
    if (exp_bits == 15u && mant == 7u) mant = 6u;           /* :87-89 */
    return (uint8_t)((sign << 7) | (exp_bits << 3) | mant); /* :91 */
}

static inline float bf16_to_f32(uint16_t h)
{
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

static inline float f16_to_f32(uint16_t h)
{
    _Float16 x;
    memcpy(&x, &h, 2);
    return (float)x;
}

/* fp8_matmul.metal:228-236, input widened exactly to fp32 first (fp8_mps_native.py:142) */
typedef struct { const void* in; uint8_t* out; int kind; } enc_ctx_t;

static void enc_body(void* p, ptrdiff_t b, ptrdiff_t e)
{
    enc_ctx_t* c = (enc_ctx_t*)p;
    if (c->kind == 0) { const float* in = (const float*)c->in;
        for (ptrdiff_t i = b; i < e; ++i) c->out[i] = fp8o_encode(in[i]); }
    else if (c->kind == 1) { const uint16_t* in = (const uint16_t*)c->in;
        for (ptrdiff_t i = b; i < e; ++i) c->out[i] = fp8o_encode(bf16_to_f32(in[i])); }
    else { const uint16_t* in = (const uint16_t*)c->in;
        for (ptrdiff_t i = b; i < e; ++i) c->out[i] = fp8o_encode(f16_to_f32(in[i])); }
}

void fp8o_encode_f32(const float* in, uint8_t* out, size_t n)
{ enc_ctx_t c = { in, out, 0 }; parallel_for((ptrdiff_t)n, 4096, enc_body, &c); }

void fp8o_encode_bf16(const uint16_t* in, uint8_t* out, size_t n)
{ enc_ctx_t c = { in, out, 1 }; parallel_for((ptrdiff_t)n, 4096, enc_body, &c); }

void fp8o_encode_f16(const uint16_t* in, uint8_t* out, size_t n)
{ enc_ctx_t c = { in, out, 2 }; parallel_for((ptrdiff_t)n, 4096, enc_body, &c); }

/* fp8_matmul.metal:215-223, then the host's fp16 multiply by half(scale)
 * (fp8_mps_native.py:121-122).  has_scale == 0 is the bare kernel. */
typedef struct { const uint8_t* in; uint16_t* out; float scale; int has_scale; } deq_ctx_t;

static void deq_body(void* p, ptrdiff_t b, ptrdiff_t e)
{
    deq_ctx_t* c = (deq_ctx_t*)p;
    _Float16 s = (_Float16)c->scale;                        /* native.py:121 */
    for (ptrdiff_t i = b; i < e; ++i) {
        _Float16 h = (_Float16)g_lut[c->in[i]];             /* metal:222 */
        if (c->has_scale) h = (_Float16)((float)h * (float)s); /* exact product, one rounding */
        memcpy(&c->out[i], &h, 2);
    }
}

void fp8o_to_half(const uint8_t* in, uint16_t* out, size_t n, float scale, int has_scale)
{
    ensure_lut();
    deq_ctx_t c = { in, out, scale, has_scale };
    parallel_for((ptrdiff_t)n, 4096, deq_body, &c);
}

/* fp8_matmul.metal:155-210.  One "simdgroup" of 32 lanes per row; lane l walks
 * k = 4l, 4l+128, ... (:177), adds x0*w0 + x1*w1 + x2*w2 + x3*w3 per step (:192),
 * the 32 partial sums are then added (simd_sum, :202).  sw_len is 1 or N. */
typedef struct { const uint8_t* x; const uint8_t* W; float* out; const float* sx; const float* sw;
                 int sw_len; uint32_t N, K; } vm_ctx_t;

static void vm_body(void* p, ptrdiff_t rb, ptrdiff_t re)
{
    vm_ctx_t* c = (vm_ctx_t*)p;
    const uint8_t* x = c->x;
    uint32_t K = c->K;
    for (ptrdiff_t row = rb; row < re; ++row) {
        const uint8_t* w = c->W + (size_t)row * K;
        float lane_sum[32];
        for (uint32_t lane = 0; lane < 32; ++lane) {
            float sum = 0.0f;
            for (uint32_t k = lane * 4; k < K; k += 32 * 4) {
                if (k + 3 < K) {
                    sum += g_lut[x[k]] * g_lut[w[k]] + g_lut[x[k + 1]] * g_lut[w[k + 1]] +
                           g_lut[x[k + 2]] * g_lut[w[k + 2]] + g_lut[x[k + 3]] * g_lut[w[k + 3]];
                } else {
                    uint32_t end = k + 4 < K ? k + 4 : K;
                    for (uint32_t kk = k; kk < end; ++kk) sum += g_lut[x[kk]] * g_lut[w[kk]];
                }
            }
            lane_sum[lane] = sum;
        }
        for (int off = 16; off > 0; off >>= 1)              /* simd_sum: butterfly order */
            for (int l = 0; l < off; ++l) lane_sum[l] += lane_sum[l + off];
        float sx = c->sx[0];                                      /* :206 */
        float sw = (c->sw_len == 1) ? c->sw[0] : c->sw[row];      /* :207 */
        c->out[row] = lane_sum[0] * sx * sw;                      /* :208 */
    }
}

void fp8o_vecmat(const uint8_t* x, const uint8_t* W, float* out,
                 const float* scale_x, const float* scale_w, int sw_len,
                 uint32_t N, uint32_t K)
{
    ensure_lut();
    vm_ctx_t c = { x, W, out, scale_x, scale_w, sw_len, N, K };
    parallel_for((ptrdiff_t)N, (size_t)N * K < (1u << 18) ? (ptrdiff_t)N + 1 : 2, vm_body, &c);
}

/* fp8_matmul.metal:99-147.  sa_len is 1 or M, sb_len is 1 or N (independent:
 * the reference's single scale_mode flag reads out of bounds when they differ).
 * Threads split the flattened (row, col) output index. */
typedef struct { const uint8_t* A; const uint8_t* B; float* C; const float* sa; int sa_len;
                 const float* sb; int sb_len; uint32_t M, N, K; } mm_ctx_t;

static void mm_body(void* p, ptrdiff_t ib, ptrdiff_t ie)
{
    mm_ctx_t* c = (mm_ctx_t*)p;
    uint32_t K = c->K, N = c->N;
    uint32_t K4 = (K / 4) * 4;
    for (ptrdiff_t idx = ib; idx < ie; ++idx) {
        size_t row = (size_t)idx / N, col = (size_t)idx % N;
        const uint8_t* a = c->A + row * K;
        const uint8_t* b = c->B + col * K;
        float sum = 0.0f;
        uint32_t k = 0;
        for (; k < K4; k += 4) {                                  /* :121-137 */
            sum += g_lut[a[k]] * g_lut[b[k]] + g_lut[a[k + 1]] * g_lut[b[k + 1]] +
                   g_lut[a[k + 2]] * g_lut[b[k + 2]] + g_lut[a[k + 3]] * g_lut[b[k + 3]];
        }
        for (; k < K; ++k) sum += g_lut[a[k]] * g_lut[b[k]];      /* :140-142 */
        float sa = (c->sa_len == 1) ? c->sa[0] : c->sa[row];      /* :144 */
        float sb = (c->sb_len == 1) ? c->sb[0] : c->sb[col];      /* :145 */
        c->C[row * N + col] = sum * sa * sb;                      /* :146 */
    }
}

void fp8o_matmul(const uint8_t* A, const uint8_t* B, float* C,
                 const float* scale_a, int sa_len, const float* scale_b, int sb_len,
                 uint32_t M, uint32_t N, uint32_t K)
{
    ensure_lut();
    mm_ctx_t c = { A, B, C, scale_a, sa_len, scale_b, sb_len, M, N, K };
    ptrdiff_t n = (ptrdiff_t)M * N;
    parallel_for(n, (size_t)n * K < (1u << 18) ? n + 1 : 2, mm_body, &c);
}
