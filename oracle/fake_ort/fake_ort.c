/*
 * TEST INFRASTRUCTURE — CPU fp32 restatement of what ONNXRuntime 1.13.1 (CPU EP) does for the
 * reference at its three call sites /root/reference/src/april_session.c:145,160,176
 * (`g_ort->Run` on the encoder / decoder / joiner graphs) and at model load
 * (/root/reference/src/april_model.c:45-82, src/ort_util.h:127-134, src/ort_util.c:20-50).
 *
 * ORT itself is absent from /root/reference (un-vendored dependency, pinned by URL to v1.13.1 in
 * download_onnx_linux_x64.sh:8) and from this image, so its *published* semantics are restated:
 * an ONNX ModelProto (opset 11) is parsed from memory and each node is evaluated in graph order
 * with the operator definitions of the ONNX spec (Conv, Gemm, MatMul, Add, Sub, Mul, Div, Pow,
 * Sigmoid, Tanh, Relu, Exp, Sqrt, Neg, ReduceMean, Transpose, Reshape, Flatten, Squeeze,
 * Unsqueeze, Slice, Split, Concat, Gather, Identity, Constant, Cast). Arithmetic is IEEE fp32 with
 * sequential-k accumulation for MatMul and 8-lane strided partial sums for Gemm(transB) —
 * MLAS's exact summation order is not reproducible, which is why logits are compared with a
 * tolerance while token sequences are compared exactly (DESIGN.md "Parity").
 *
 * PARITY STATUS: unpinned by the reference (it ships no golden vectors). This
 * interpreter is cross-checked against an independent NumPy restatement (tools/np_ref.py) in
 * tests/test_oracle.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may load it.
 */
#include "onnxruntime_c_api.h"

#include <math.h>
#include <stdio.h>
#include <stdatomic.h>

/* ------------------------------------------------------------------------------------------ */
/* tensors                                                                                      */
/* ------------------------------------------------------------------------------------------ */
#define MAXD 8
enum { DT_F32 = 1, DT_I32 = 6, DT_I64 = 7 };

typedef struct {
    int dtype, ndim;
    int64_t dims[MAXD];
    int64_t numel;
    void* data;
    int owned;
} T;

static int64_t numel_of(int ndim, const int64_t* d) { int64_t n = 1; for (int i = 0; i < ndim; i++) n *= d[i]; return n; }
static size_t dsize(int dt) { return dt == DT_I64 ? 8 : 4; }

static T t_new(int dtype, int ndim, const int64_t* dims) {
    T t; memset(&t, 0, sizeof t);
    t.dtype = dtype; t.ndim = ndim;
    for (int i = 0; i < ndim; i++) t.dims[i] = dims[i];
    t.numel = numel_of(ndim, dims);
    t.data = malloc((size_t)(t.numel > 0 ? t.numel : 1) * dsize(dtype));
    t.owned = 1;
    return t;
}
static void t_free(T* t) { if (t->owned && t->data) free(t->data); t->data = NULL; t->owned = 0; }
static T t_view(const T* s) { T t = *s; t.owned = 0; return t; }   /* shares storage */
static T t_copy(const T* s) { T t = t_new(s->dtype, s->ndim, s->dims); memcpy(t.data, s->data, (size_t)s->numel * dsize(s->dtype)); return t; }

/* ------------------------------------------------------------------------------------------ */
/* protobuf cursor                                                                              */
/* ------------------------------------------------------------------------------------------ */
typedef struct { const uint8_t* p; const uint8_t* end; } Cur;

static uint64_t rd_varint(Cur* c) {
    uint64_t v = 0; int sh = 0;
    while (c->p < c->end) { uint8_t b = *c->p++; v |= (uint64_t)(b & 0x7F) << sh; if (!(b & 0x80)) break; sh += 7; }
    return v;
}
/* returns 0 at end; else field number, sets wt and (val | sub) */
static int rd_field(Cur* c, int* wt, uint64_t* val, Cur* sub) {
    if (c->p >= c->end) return 0;
    uint64_t key = rd_varint(c);
    *wt = (int)(key & 7);
    int field = (int)(key >> 3);
    switch (*wt) {
        case 0: *val = rd_varint(c); break;
        case 1: memcpy(val, c->p, 8); c->p += 8; break;
        case 5: { uint32_t v; memcpy(&v, c->p, 4); *val = v; c->p += 4; break; }
        case 2: { uint64_t n = rd_varint(c); sub->p = c->p; sub->end = c->p + n; c->p += n; break; }
        default: fprintf(stderr, "fake_ort: bad wire type %d\n", *wt); abort();
    }
    return field;
}
static char* cur_strdup(Cur s) { size_t n = (size_t)(s.end - s.p); char* r = (char*)malloc(n + 1); memcpy(r, s.p, n); r[n] = 0; return r; }

/* ------------------------------------------------------------------------------------------ */
/* graph                                                                                        */
/* ------------------------------------------------------------------------------------------ */
enum Op { OP_CONV, OP_GEMM, OP_MATMUL, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_SIGMOID, OP_TANH, OP_RELU,
          OP_EXP, OP_SQRT, OP_NEG, OP_REDUCEMEAN, OP_TRANSPOSE, OP_RESHAPE, OP_FLATTEN, OP_SQUEEZE, OP_UNSQUEEZE,
          OP_SLICE, OP_SPLIT, OP_CONCAT, OP_GATHER, OP_IDENTITY, OP_CONSTANT, OP_CAST, OP_UNKNOWN };
static const char* OPN[] = { "Conv", "Gemm", "MatMul", "Add", "Sub", "Mul", "Div", "Pow", "Sigmoid", "Tanh", "Relu",
          "Exp", "Sqrt", "Neg", "ReduceMean", "Transpose", "Reshape", "Flatten", "Squeeze", "Unsqueeze",
          "Slice", "Split", "Concat", "Gather", "Identity", "Constant", "Cast" };

#define MAXIO 16
typedef struct {
    int op;
    char* op_name;
    int nin, nout;
    int in[MAXIO], out[MAXIO];
    /* attributes */
    int64_t ints[MAXD * 2]; int n_ints;       /* axes / perm / split */
    int64_t axis; int has_axis;
    int64_t strides[2], pads[4], group, keepdims, transA, transB, to;
    float alpha, beta;
    T value; int has_value;
} Node;

struct OrtSession {
    int net_index;
    int nvals; char** val_names; T* consts; uint8_t* is_const;
    int nnodes; Node* nodes;
    int n_in, n_out; int in_ids[8], out_ids[8];
    int in_ndim[8], out_ndim[8]; int64_t in_dims[8][MAXD], out_dims[8][MAXD];
};
struct OrtEnv { int n_sessions; };
struct OrtStatus { char msg[256]; };
struct OrtSessionOptions { int intra, inter; };
struct OrtMemoryInfo { int dummy; };
struct OrtValue { T t; };
struct OrtTypeInfo { int ndim; int64_t dims[MAXD]; };

static OrtStatus* mk_status(const char* fmt, const char* a) {
    OrtStatus* s = (OrtStatus*)calloc(1, sizeof *s); snprintf(s->msg, sizeof s->msg, fmt, a); return s;
}

static int val_id(OrtSession* s, const char* name, int create) {
    for (int i = 0; i < s->nvals; i++) if (strcmp(s->val_names[i], name) == 0) return i;
    if (!create) return -1;
    if ((s->nvals & 255) == 0) {
        s->val_names = (char**)realloc(s->val_names, sizeof(char*) * (size_t)(s->nvals + 256));
        s->consts = (T*)realloc(s->consts, sizeof(T) * (size_t)(s->nvals + 256));
        s->is_const = (uint8_t*)realloc(s->is_const, (size_t)(s->nvals + 256));
    }
    s->val_names[s->nvals] = strdup(name);
    memset(&s->consts[s->nvals], 0, sizeof(T));
    s->is_const[s->nvals] = 0;
    return s->nvals++;
}

static T parse_tensor(Cur c, char** name_out) {
    int64_t dims[MAXD]; int ndim = 0, dtype = DT_F32; Cur raw = { 0, 0 };
    float* fl = NULL; size_t nfl = 0; int64_t* il = NULL; size_t nil = 0;
    int wt; uint64_t v; Cur sub; int f;
    while ((f = rd_field(&c, &wt, &v, &sub))) {
        if (f == 1) { if (wt == 0) dims[ndim++] = (int64_t)v; else { while (sub.p < sub.end) dims[ndim++] = (int64_t)rd_varint(&sub); } }
        else if (f == 2) dtype = (int)v;
        else if (f == 8 && name_out) *name_out = cur_strdup(sub);
        else if (f == 9) raw = sub;
        else if (f == 4) { /* float_data */
            if (wt == 5) { fl = (float*)realloc(fl, 4 * (nfl + 1)); uint32_t u = (uint32_t)v; memcpy(&fl[nfl++], &u, 4); }
            else { size_t n = (size_t)(sub.end - sub.p) / 4; fl = (float*)realloc(fl, 4 * (nfl + n)); memcpy(fl + nfl, sub.p, 4 * n); nfl += n; }
        } else if (f == 7) { /* int64_data */
            if (wt == 0) { il = (int64_t*)realloc(il, 8 * (nil + 1)); il[nil++] = (int64_t)v; }
            else while (sub.p < sub.end) { il = (int64_t*)realloc(il, 8 * (nil + 1)); il[nil++] = (int64_t)rd_varint(&sub); }
        }
    }
    if (dtype != DT_F32 && dtype != DT_I64 && dtype != DT_I32) { fprintf(stderr, "fake_ort: unsupported tensor dtype %d\n", dtype); abort(); }
    T t = t_new(dtype, ndim, dims);
    if (raw.p) memcpy(t.data, raw.p, (size_t)t.numel * dsize(dtype));
    else if (dtype == DT_F32 && fl) memcpy(t.data, fl, 4 * nfl);
    else if (dtype == DT_I64 && il) memcpy(t.data, il, 8 * nil);
    free(fl); free(il);
    return t;
}

static void parse_value_info(Cur c, char** name, int* ndim, int64_t* dims) {
    int wt; uint64_t v; Cur sub; int f; *ndim = 0;
    while ((f = rd_field(&c, &wt, &v, &sub))) {
        if (f == 1) *name = cur_strdup(sub);
        else if (f == 2) { /* TypeProto */
            Cur c2 = sub, s2; int f2;
            while ((f2 = rd_field(&c2, &wt, &v, &s2))) if (f2 == 1) { /* tensor_type */
                Cur c3 = s2, s3; int f3;
                while ((f3 = rd_field(&c3, &wt, &v, &s3))) if (f3 == 2) { /* shape */
                    Cur c4 = s3, s4; int f4;
                    while ((f4 = rd_field(&c4, &wt, &v, &s4))) if (f4 == 1) { /* dim */
                        Cur c5 = s4, s5; int f5; int64_t dv = -1;
                        while ((f5 = rd_field(&c5, &wt, &v, &s5))) if (f5 == 1) dv = (int64_t)v;
                        dims[(*ndim)++] = dv;
                    }
                }
            }
        }
    }
}

static void parse_attr(Cur c, Node* n) {
    char* name = NULL; int wt; uint64_t v; Cur sub; int f;
    int64_t ints[MAXD * 2]; int n_ints = 0; int64_t iv = 0; float fv = 0; int has_t = 0; T tv; memset(&tv, 0, sizeof tv);
    while ((f = rd_field(&c, &wt, &v, &sub))) {
        if (f == 1) name = cur_strdup(sub);
        else if (f == 2) { uint32_t u = (uint32_t)v; memcpy(&fv, &u, 4); }
        else if (f == 3) iv = (int64_t)v;
        else if (f == 5) { tv = parse_tensor(sub, NULL); has_t = 1; }
        else if (f == 8) { if (wt == 0) ints[n_ints++] = (int64_t)v; else while (sub.p < sub.end) ints[n_ints++] = (int64_t)rd_varint(&sub); }
    }
    if (!name) return;
    if (!strcmp(name, "axes") || !strcmp(name, "perm") || !strcmp(name, "split")) { memcpy(n->ints, ints, sizeof(int64_t) * (size_t)n_ints); n->n_ints = n_ints; }
    else if (!strcmp(name, "axis")) { n->axis = iv; n->has_axis = 1; }
    else if (!strcmp(name, "strides")) { n->strides[0] = ints[0]; n->strides[1] = n_ints > 1 ? ints[1] : 1; }
    else if (!strcmp(name, "pads")) { for (int i = 0; i < n_ints && i < 4; i++) n->pads[i] = ints[i]; }
    else if (!strcmp(name, "group")) n->group = iv;
    else if (!strcmp(name, "keepdims")) n->keepdims = iv;
    else if (!strcmp(name, "transA")) n->transA = iv;
    else if (!strcmp(name, "transB")) n->transB = iv;
    else if (!strcmp(name, "alpha")) n->alpha = fv;
    else if (!strcmp(name, "beta")) n->beta = fv;
    else if (!strcmp(name, "to")) n->to = iv;
    else if (!strcmp(name, "value")) { n->value = tv; n->has_value = 1; has_t = 0; }
    else if (!strcmp(name, "dilations")) { for (int i = 0; i < n_ints; i++) if (ints[i] != 1) { fprintf(stderr, "fake_ort: dilation != 1\n"); abort(); } }
    /* kernel_shape is implied by the weight */
    if (has_t) t_free(&tv);
    free(name);
}

static void parse_node(OrtSession* s, Cur c) {
    s->nodes = (Node*)realloc(s->nodes, sizeof(Node) * (size_t)(s->nnodes + 1));
    Node* n = &s->nodes[s->nnodes++]; memset(n, 0, sizeof *n);
    n->strides[0] = n->strides[1] = 1; n->group = 1; n->keepdims = 1; n->alpha = 1.f; n->beta = 1.f; n->op = OP_UNKNOWN;
    int wt; uint64_t v; Cur sub; int f;
    while ((f = rd_field(&c, &wt, &v, &sub))) {
        if (f == 1) { char* nm = cur_strdup(sub); n->in[n->nin++] = nm[0] ? val_id(s, nm, 1) : -1; free(nm); }
        else if (f == 2) { char* nm = cur_strdup(sub); n->out[n->nout++] = val_id(s, nm, 1); free(nm); }
        else if (f == 4) { n->op_name = cur_strdup(sub); for (int i = 0; i < OP_UNKNOWN; i++) if (!strcmp(n->op_name, OPN[i])) n->op = i; }
        else if (f == 5) parse_attr(sub, n);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* kernels                                                                                      */
/* ------------------------------------------------------------------------------------------ */
static inline float dot8(const float* a, const float* b, int64_t k) {
    float acc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    int64_t i = 0;
    for (; i + 8 <= k; i += 8) for (int j = 0; j < 8; j++) acc[j] += a[i + j] * b[i + j];
    float s = ((acc[0] + acc[4]) + (acc[1] + acc[5])) + ((acc[2] + acc[6]) + (acc[3] + acc[7]));
    for (; i < k; i++) s += a[i] * b[i];
    return s;
}

static void bcast_strides(const T* t, int nd, const int64_t* od, int64_t* st) {
    int off = nd - t->ndim; int64_t s = 1;
    for (int i = nd - 1; i >= 0; i--) {
        int j = i - off;
        if (j < 0) { st[i] = 0; continue; }
        st[i] = (t->dims[j] == 1 && od[i] != 1) ? 0 : s;
        s *= t->dims[j];
    }
}

static T binary(int op, const T* a, const T* b) {
    int nd = a->ndim > b->ndim ? a->ndim : b->ndim; int64_t od[MAXD];
    for (int i = 0; i < nd; i++) {
        int ja = i - (nd - a->ndim), jb = i - (nd - b->ndim);
        int64_t da = ja >= 0 ? a->dims[ja] : 1, db = jb >= 0 ? b->dims[jb] : 1;
        od[i] = da > db ? da : db;
    }
    int dt = a->dtype;
    T o = t_new(dt, nd, od);
    int64_t sa[MAXD], sb[MAXD]; bcast_strides(a, nd, od, sa); bcast_strides(b, nd, od, sb);
    int64_t idx[MAXD] = { 0 };
    for (int64_t n = 0; n < o.numel; n++) {
        int64_t ia = 0, ib = 0; for (int i = 0; i < nd; i++) { ia += idx[i] * sa[i]; ib += idx[i] * sb[i]; }
        if (dt == DT_F32) {
            float x = ((float*)a->data)[ia], y = ((float*)b->data)[ib], r;
            switch (op) { case OP_ADD: r = x + y; break; case OP_SUB: r = x - y; break; case OP_MUL: r = x * y; break;
                          case OP_DIV: r = x / y; break; default: r = (y == 2.0f) ? x * x : powf(x, y); }
            ((float*)o.data)[n] = r;
        } else {
            int64_t x = ((int64_t*)a->data)[ia], y = ((int64_t*)b->data)[ib], r;
            switch (op) { case OP_ADD: r = x + y; break; case OP_SUB: r = x - y; break; case OP_MUL: r = x * y; break; default: r = y ? x / y : 0; }
            ((int64_t*)o.data)[n] = r;
        }
        for (int i = nd - 1; i >= 0; i--) { if (++idx[i] < od[i]) break; idx[i] = 0; }
    }
    return o;
}

static T unary(int op, const T* a) {
    T o = t_new(DT_F32, a->ndim, a->dims); const float* x = (const float*)a->data; float* y = (float*)o.data;
    for (int64_t i = 0; i < a->numel; i++) {
        float v = x[i];
        switch (op) {
            case OP_SIGMOID: v = 1.0f / (1.0f + expf(-v)); break;
            case OP_TANH: v = tanhf(v); break;
            case OP_RELU: v = v > 0 ? v : 0; break;
            case OP_EXP: v = expf(v); break;
            case OP_SQRT: v = sqrtf(v); break;
            case OP_NEG: v = -v; break;
        }
        y[i] = v;
    }
    return o;
}

static T matmul(const T* a, const T* b) {            /* a [..., M, K] x b [K, N] */
    if (b->ndim != 2) { fprintf(stderr, "fake_ort: MatMul rhs must be 2-D\n"); abort(); }
    int64_t K = b->dims[0], N = b->dims[1], M = a->numel / K;
    int64_t od[MAXD]; for (int i = 0; i < a->ndim - 1; i++) od[i] = a->dims[i]; od[a->ndim - 1] = N;
    T o = t_new(DT_F32, a->ndim, od);
    const float* A = (const float*)a->data; const float* B = (const float*)b->data; float* C = (float*)o.data;
    for (int64_t m = 0; m < M; m++) {
        float* c = C + m * N; for (int64_t n = 0; n < N; n++) c[n] = 0.f;
        for (int64_t k = 0; k < K; k++) { float av = A[m * K + k]; const float* br = B + k * N; for (int64_t n = 0; n < N; n++) c[n] += av * br[n]; }
    }
    return o;
}

static T gemm(const Node* nd, const T* a, const T* b, const T* c) {
    if (nd->transA) { fprintf(stderr, "fake_ort: Gemm transA unsupported\n"); abort(); }
    int64_t M = a->dims[0], K = a->dims[1], N = nd->transB ? b->dims[0] : b->dims[1];
    int64_t od[2] = { M, N }; T o = t_new(DT_F32, 2, od);
    const float* A = (const float*)a->data; const float* B = (const float*)b->data; float* C = (float*)o.data;
    for (int64_t m = 0; m < M; m++) for (int64_t n = 0; n < N; n++) {
        float s;
        if (nd->transB) s = dot8(A + m * K, B + n * K, K);
        else { s = 0; for (int64_t k = 0; k < K; k++) s += A[m * K + k] * B[k * N + n]; }
        s *= nd->alpha;
        if (c) { const float* cb = (const float*)c->data; s += nd->beta * (c->numel == N ? cb[n] : (c->numel == 1 ? cb[0] : cb[m * N + n])); }
        C[m * N + n] = s;
    }
    return o;
}

static T conv(const Node* nd, const T* x, const T* w, const T* b) {
    /* x [N,C,H,W] or [N,C,W]; w [O,C/g,kh,kw] or [O,C/g,kw] */
    int is1d = x->ndim == 3;
    int64_t N = x->dims[0], C = x->dims[1], H = is1d ? 1 : x->dims[2], W = is1d ? x->dims[2] : x->dims[3];
    int64_t O = w->dims[0], Cg = w->dims[1], kh = is1d ? 1 : w->dims[2], kw = is1d ? w->dims[2] : w->dims[3];
    int64_t sh = is1d ? 1 : nd->strides[0], sw = is1d ? nd->strides[0] : nd->strides[1];
    int64_t ph0 = is1d ? 0 : nd->pads[0], pw0 = is1d ? nd->pads[0] : nd->pads[1];
    int64_t ph1 = is1d ? 0 : nd->pads[2], pw1 = is1d ? nd->pads[1] : nd->pads[3];
    int64_t Ho = (H + ph0 + ph1 - kh) / sh + 1, Wo = (W + pw0 + pw1 - kw) / sw + 1;
    int64_t G = nd->group, Og = O / G;
    if (Cg * G != C) { fprintf(stderr, "fake_ort: Conv group mismatch\n"); abort(); }
    int64_t od4[4] = { N, O, Ho, Wo }, od3[3] = { N, O, Wo };
    T o = is1d ? t_new(DT_F32, 3, od3) : t_new(DT_F32, 4, od4);
    const float* X = (const float*)x->data; const float* Wt = (const float*)w->data; float* Y = (float*)o.data;
    for (int64_t n = 0; n < N; n++) for (int64_t oc = 0; oc < O; oc++) {
        int64_t g = oc / Og; float bias = b ? ((const float*)b->data)[oc] : 0.f;
        for (int64_t oy = 0; oy < Ho; oy++) for (int64_t ox = 0; ox < Wo; ox++) {
            float s = 0.f;
            for (int64_t ci = 0; ci < Cg; ci++) {
                const float* xp = X + ((n * C + g * Cg + ci) * H) * W; const float* wp = Wt + ((oc * Cg + ci) * kh) * kw;
                for (int64_t ky = 0; ky < kh; ky++) {
                    int64_t iy = oy * sh + ky - ph0; if (iy < 0 || iy >= H) continue;
                    for (int64_t kx = 0; kx < kw; kx++) {
                        int64_t ix = ox * sw + kx - pw0; if (ix < 0 || ix >= W) continue;
                        s += xp[iy * W + ix] * wp[ky * kw + kx];
                    }
                }
            }
            Y[((n * O + oc) * Ho + oy) * Wo + ox] = s + bias;
        }
    }
    return o;
}

static T transpose(const Node* nd, const T* a) {
    int n = a->ndim; int64_t perm[MAXD], od[MAXD], st[MAXD];
    for (int i = 0; i < n; i++) perm[i] = nd->n_ints ? nd->ints[i] : n - 1 - i;
    int64_t s = 1; for (int i = n - 1; i >= 0; i--) { st[i] = s; s *= a->dims[i]; }
    for (int i = 0; i < n; i++) od[i] = a->dims[perm[i]];
    T o = t_new(a->dtype, n, od); size_t es = dsize(a->dtype);
    int64_t idx[MAXD] = { 0 };
    for (int64_t k = 0; k < o.numel; k++) {
        int64_t src = 0; for (int i = 0; i < n; i++) src += idx[i] * st[perm[i]];
        memcpy((char*)o.data + (size_t)k * es, (const char*)a->data + (size_t)src * es, es);
        for (int i = n - 1; i >= 0; i--) { if (++idx[i] < od[i]) break; idx[i] = 0; }
    }
    return o;
}

static T reshape_to(const T* a, int nd, const int64_t* shape) {
    int64_t od[MAXD]; int64_t known = 1; int neg = -1;
    for (int i = 0; i < nd; i++) { od[i] = shape[i] == 0 ? a->dims[i] : shape[i]; if (od[i] < 0) neg = i; else known *= od[i]; }
    if (neg >= 0) od[neg] = a->numel / known;
    T o = t_copy(a); o.ndim = nd; for (int i = 0; i < nd; i++) o.dims[i] = od[i];
    return o;
}

static T reduce_mean(const Node* nd, const T* a) {
    int ax = (int)(nd->n_ints ? nd->ints[0] : -1); if (ax < 0) ax += a->ndim;
    if (nd->n_ints > 1 || ax != a->ndim - 1) { fprintf(stderr, "fake_ort: ReduceMean only over the last axis\n"); abort(); }
    int64_t K = a->dims[ax], M = a->numel / K; int64_t od[MAXD]; int ondim = 0;
    for (int i = 0; i < a->ndim; i++) { if (i == ax) { if (nd->keepdims) od[ondim++] = 1; } else od[ondim++] = a->dims[i]; }
    T o = t_new(DT_F32, ondim, od); const float* x = (const float*)a->data; float* y = (float*)o.data;
    for (int64_t m = 0; m < M; m++) { float s = 0.f; for (int64_t k = 0; k < K; k++) s += x[m * K + k]; y[m] = s / (float)K; }
    return o;
}

static const int64_t* i64_data(const T* t) { return (const int64_t*)t->data; }

static T slice(const T* a, const T* starts, const T* ends, const T* axes, const T* steps) {
    int n = a->ndim; int64_t st[MAXD], en[MAXD], sp[MAXD];
    for (int i = 0; i < n; i++) { st[i] = 0; en[i] = a->dims[i]; sp[i] = 1; }
    for (int64_t j = 0; j < starts->numel; j++) {
        int ax = (int)(axes ? i64_data(axes)[j] : j); if (ax < 0) ax += n;
        int64_t d = a->dims[ax], s0 = i64_data(starts)[j], e0 = i64_data(ends)[j];
        if (steps && i64_data(steps)[j] != 1) { fprintf(stderr, "fake_ort: Slice step != 1\n"); abort(); }
        if (s0 < 0) s0 += d; if (e0 < 0) e0 += d; if (s0 > d) s0 = d; if (e0 > d) e0 = d; if (s0 < 0) s0 = 0;
        st[ax] = s0; en[ax] = e0 > s0 ? e0 : s0;
    }
    int64_t od[MAXD], as[MAXD]; int64_t s = 1;
    for (int i = n - 1; i >= 0; i--) { as[i] = s; s *= a->dims[i]; od[i] = en[i] - st[i]; }
    T o = t_new(a->dtype, n, od); size_t es = dsize(a->dtype); int64_t idx[MAXD] = { 0 };
    for (int64_t k = 0; k < o.numel; k++) {
        int64_t src = 0; for (int i = 0; i < n; i++) src += (idx[i] + st[i]) * as[i];
        memcpy((char*)o.data + (size_t)k * es, (const char*)a->data + (size_t)src * es, es);
        for (int i = n - 1; i >= 0; i--) { if (++idx[i] < od[i]) break; idx[i] = 0; }
    }
    (void)sp;
    return o;
}

static T concat(const Node* nd, T** ins, int nin) {
    const T* a = ins[0]; int ax = (int)nd->axis; if (ax < 0) ax += a->ndim;
    int64_t od[MAXD]; for (int i = 0; i < a->ndim; i++) od[i] = a->dims[i];
    od[ax] = 0; for (int j = 0; j < nin; j++) od[ax] += ins[j]->dims[ax];
    T o = t_new(a->dtype, a->ndim, od); size_t es = dsize(a->dtype);
    int64_t outer = 1, inner = 1; for (int i = 0; i < ax; i++) outer *= od[i]; for (int i = ax + 1; i < a->ndim; i++) inner *= od[i];
    int64_t off = 0;
    for (int j = 0; j < nin; j++) {
        int64_t chunk = ins[j]->dims[ax] * inner;
        for (int64_t q = 0; q < outer; q++)
            memcpy((char*)o.data + (size_t)((q * od[ax] * inner) + off) * es, (const char*)ins[j]->data + (size_t)(q * chunk) * es, (size_t)chunk * es);
        off += chunk;
    }
    return o;
}

static T gather(const Node* nd, const T* data, const T* idx) {
    int ax = (int)nd->axis; if (ax < 0) ax += data->ndim;
    if (ax != 0) { fprintf(stderr, "fake_ort: Gather axis != 0\n"); abort(); }
    int64_t inner = data->numel / data->dims[0]; int64_t od[MAXD]; int ondim = 0;
    for (int i = 0; i < idx->ndim; i++) od[ondim++] = idx->dims[i];
    for (int i = 1; i < data->ndim; i++) od[ondim++] = data->dims[i];
    T o = t_new(data->dtype, ondim, od); size_t es = dsize(data->dtype);
    for (int64_t j = 0; j < idx->numel; j++) {
        int64_t k = idx->dtype == DT_I64 ? i64_data(idx)[j] : ((const int32_t*)idx->data)[j]; if (k < 0) k += data->dims[0];
        memcpy((char*)o.data + (size_t)(j * inner) * es, (const char*)data->data + (size_t)(k * inner) * es, (size_t)inner * es);
    }
    return o;
}

/* evaluate one node; vals[] holds views/owned tensors */
static void eval_node(const Node* nd, T* vals) {
    T* in[MAXIO]; for (int i = 0; i < nd->nin; i++) in[i] = nd->in[i] >= 0 ? &vals[nd->in[i]] : NULL;
    T out; memset(&out, 0, sizeof out);
    switch (nd->op) {
        case OP_CONV: out = conv(nd, in[0], in[1], nd->nin > 2 ? in[2] : NULL); break;
        case OP_GEMM: out = gemm(nd, in[0], in[1], nd->nin > 2 ? in[2] : NULL); break;
        case OP_MATMUL: out = matmul(in[0], in[1]); break;
        case OP_ADD: case OP_SUB: case OP_MUL: case OP_DIV: case OP_POW: out = binary(nd->op, in[0], in[1]); break;
        case OP_SIGMOID: case OP_TANH: case OP_RELU: case OP_EXP: case OP_SQRT: case OP_NEG: out = unary(nd->op, in[0]); break;
        case OP_REDUCEMEAN: out = reduce_mean(nd, in[0]); break;
        case OP_TRANSPOSE: out = transpose(nd, in[0]); break;
        case OP_RESHAPE: out = reshape_to(in[0], (int)in[1]->numel, i64_data(in[1])); break;
        case OP_FLATTEN: { int ax = nd->has_axis ? (int)nd->axis : 1; int64_t a = 1, b = 1; for (int i = 0; i < in[0]->ndim; i++) (i < ax ? (a *= in[0]->dims[i]) : (b *= in[0]->dims[i])); int64_t sh[2] = { a, b }; out = reshape_to(in[0], 2, sh); break; }
        case OP_SQUEEZE: { int64_t sh[MAXD]; int k = 0; for (int i = 0; i < in[0]->ndim; i++) { int drop = 0; if (nd->n_ints == 0) drop = in[0]->dims[i] == 1; else for (int j = 0; j < nd->n_ints; j++) { int ax = (int)nd->ints[j]; if (ax < 0) ax += in[0]->ndim; if (ax == i) drop = 1; } if (!drop) sh[k++] = in[0]->dims[i]; } out = reshape_to(in[0], k, sh); break; }
        case OP_UNSQUEEZE: { int ond = in[0]->ndim + nd->n_ints; int64_t sh[MAXD]; int src = 0; for (int i = 0; i < ond; i++) { int ins = 0; for (int j = 0; j < nd->n_ints; j++) { int ax = (int)nd->ints[j]; if (ax < 0) ax += ond; if (ax == i) ins = 1; } sh[i] = ins ? 1 : in[0]->dims[src++]; } out = reshape_to(in[0], ond, sh); break; }
        case OP_SLICE: out = slice(in[0], in[1], in[2], nd->nin > 3 ? in[3] : NULL, nd->nin > 4 ? in[4] : NULL); break;
        case OP_SPLIT: {
            int ax = (int)nd->axis; if (ax < 0) ax += in[0]->ndim; int64_t off = 0;
            for (int j = 0; j < nd->nout; j++) {
                int64_t len = nd->n_ints ? nd->ints[j] : in[0]->dims[ax] / nd->nout;
                int64_t s0[1] = { off }, e0[1] = { off + len }, a0[1] = { ax }; int64_t one[1] = { 1 };
                T ts = { DT_I64, 1, { 1 }, 1, s0, 0 }, te = { DT_I64, 1, { 1 }, 1, e0, 0 }, ta = { DT_I64, 1, { 1 }, 1, a0, 0 }; (void)one;
                vals[nd->out[j]] = slice(in[0], &ts, &te, &ta, NULL); off += len;
            }
            return;
        }
        case OP_CONCAT: out = concat(nd, in, nd->nin); break;
        case OP_GATHER: out = gather(nd, in[0], in[1]); break;
        case OP_IDENTITY: out = t_copy(in[0]); break;
        case OP_CONSTANT: out = t_view(&nd->value); break;
        case OP_CAST: {
            if (nd->to == in[0]->dtype) { out = t_copy(in[0]); break; }
            out = t_new((int)nd->to, in[0]->ndim, in[0]->dims);
            for (int64_t i = 0; i < out.numel; i++) {
                double v = in[0]->dtype == DT_F32 ? ((float*)in[0]->data)[i] : in[0]->dtype == DT_I64 ? (double)((int64_t*)in[0]->data)[i] : ((int32_t*)in[0]->data)[i];
                if (nd->to == DT_F32) ((float*)out.data)[i] = (float)v; else if (nd->to == DT_I64) ((int64_t*)out.data)[i] = (int64_t)v; else ((int32_t*)out.data)[i] = (int32_t)v;
            }
            break;
        }
        default: fprintf(stderr, "fake_ort: unsupported op %s\n", nd->op_name ? nd->op_name : "?"); abort();
    }
    vals[nd->out[0]] = out;
}

/* ------------------------------------------------------------------------------------------ */
/* OrtApi                                                                                       */
/* ------------------------------------------------------------------------------------------ */
static __thread FakeOrtTap g_tap = NULL; static __thread void* g_tap_ud = NULL;   /* per calling thread: synchronous sessions Run on their caller's thread, so concurrent test threads can each tap their own session */
static _Atomic uint64_t g_run_counts[3];
void fake_ort_set_tap(FakeOrtTap tap, void* ud) { g_tap = tap; g_tap_ud = ud; }
void fake_ort_get_run_counts(uint64_t c[3]) { for (int i = 0; i < 3; i++) c[i] = atomic_load(&g_run_counts[i]); }

static OrtStatus* api_CreateEnv(OrtLoggingLevel l, const char* id, OrtEnv** out) { (void)l; (void)id; *out = (OrtEnv*)calloc(1, sizeof(OrtEnv)); return NULL; }
static OrtStatus* api_CreateSessionOptions(OrtSessionOptions** out) { *out = (OrtSessionOptions*)calloc(1, sizeof(OrtSessionOptions)); return NULL; }
static OrtStatus* api_SetIntra(OrtSessionOptions* o, int n) { o->intra = n; return NULL; }
static OrtStatus* api_SetInter(OrtSessionOptions* o, int n) { o->inter = n; return NULL; }

static OrtStatus* api_CreateSessionFromArray(const OrtEnv* env, const void* data, size_t len, const OrtSessionOptions* o, OrtSession** out) {
    (void)o;
    OrtSession* s = (OrtSession*)calloc(1, sizeof *s);
    s->net_index = ((OrtEnv*)env)->n_sessions++;
    Cur c = { (const uint8_t*)data, (const uint8_t*)data + len }, graph = { 0, 0 }, sub; int wt; uint64_t v; int f;
    while ((f = rd_field(&c, &wt, &v, &sub))) if (f == 7) graph = sub;
    if (!graph.p) { free(s); return mk_status("fake_ort: %s", "no graph in model"); }
    c = graph;
    while ((f = rd_field(&c, &wt, &v, &sub))) {
        if (f == 1) parse_node(s, sub);
        else if (f == 5) { char* name = NULL; T t = parse_tensor(sub, &name); int id = val_id(s, name ? name : "", 1); s->consts[id] = t; s->is_const[id] = 1; free(name); }
        else if (f == 11 || f == 12) {
            char* name = NULL; int nd; int64_t dims[MAXD]; parse_value_info(sub, &name, &nd, dims);
            int id = val_id(s, name, 1);
            if (f == 11) { if (!s->is_const[id]) { int k = s->n_in++; s->in_ids[k] = id; s->in_ndim[k] = nd; memcpy(s->in_dims[k], dims, sizeof dims); } }
            else { int k = s->n_out++; s->out_ids[k] = id; s->out_ndim[k] = nd; memcpy(s->out_dims[k], dims, sizeof dims); }
            free(name);
        }
    }
    for (int i = 0; i < s->nnodes; i++) if (s->nodes[i].op == OP_UNKNOWN) {
        OrtStatus* st = mk_status("fake_ort: unsupported op %s", s->nodes[i].op_name ? s->nodes[i].op_name : "?");
        return st;
    }
    *out = s; return NULL;
}
static OrtStatus* api_InCount(const OrtSession* s, size_t* n) { *n = (size_t)s->n_in; return NULL; }
static OrtStatus* api_OutCount(const OrtSession* s, size_t* n) { *n = (size_t)s->n_out; return NULL; }
static OrtStatus* api_InInfo(const OrtSession* s, size_t i, OrtTypeInfo** out) { OrtTypeInfo* t = (OrtTypeInfo*)calloc(1, sizeof *t); t->ndim = s->in_ndim[i]; memcpy(t->dims, s->in_dims[i], sizeof t->dims); *out = t; return NULL; }
static OrtStatus* api_OutInfo(const OrtSession* s, size_t i, OrtTypeInfo** out) { OrtTypeInfo* t = (OrtTypeInfo*)calloc(1, sizeof *t); t->ndim = s->out_ndim[i]; memcpy(t->dims, s->out_dims[i], sizeof t->dims); *out = t; return NULL; }
static OrtStatus* api_Cast(const OrtTypeInfo* ti, const OrtTensorTypeAndShapeInfo** out) { *out = (const OrtTensorTypeAndShapeInfo*)ti; return NULL; }
static OrtStatus* api_DimCount(const OrtTensorTypeAndShapeInfo* i, size_t* n) { *n = (size_t)((const OrtTypeInfo*)i)->ndim; return NULL; }
static OrtStatus* api_Dims(const OrtTensorTypeAndShapeInfo* i, int64_t* d, size_t n) { for (size_t k = 0; k < n; k++) d[k] = ((const OrtTypeInfo*)i)->dims[k]; return NULL; }
static void api_RelTypeInfo(OrtTypeInfo* t) { free(t); }
static OrtStatus* api_CpuMemInfo(OrtAllocatorType a, OrtMemType m, OrtMemoryInfo** out) { (void)a; (void)m; *out = (OrtMemoryInfo*)calloc(1, sizeof(OrtMemoryInfo)); return NULL; }
static OrtStatus* api_CreateTensor(const OrtMemoryInfo* mi, void* data, size_t bytes, const int64_t* shape, size_t ndim, ONNXTensorElementDataType type, OrtValue** out) {
    (void)mi;
    OrtValue* v = (OrtValue*)calloc(1, sizeof *v);
    v->t.dtype = (int)type; v->t.ndim = (int)ndim; for (size_t i = 0; i < ndim; i++) v->t.dims[i] = shape[i];
    v->t.numel = numel_of((int)ndim, shape); v->t.data = data; v->t.owned = 0;
    if ((size_t)v->t.numel * dsize(v->t.dtype) != bytes) { free(v); return mk_status("fake_ort: %s", "tensor byte size mismatch"); }
    *out = v; return NULL;
}

static OrtStatus* api_Run(OrtSession* s, const OrtRunOptions* ro, const char* const* in_names, const OrtValue* const* in, size_t n_in,
                          const char* const* out_names, size_t n_out, OrtValue** out) {
    (void)ro;
    T* vals = (T*)calloc((size_t)s->nvals, sizeof(T));
    for (int i = 0; i < s->nvals; i++) if (s->is_const[i]) vals[i] = t_view(&s->consts[i]);
    for (size_t i = 0; i < n_in; i++) {
        int id = val_id(s, in_names[i], 0);
        if (id < 0) { free(vals); return mk_status("fake_ort: unknown input %s", in_names[i]); }
        vals[id] = t_view(&in[i]->t);
    }
    for (int i = 0; i < s->nnodes; i++) eval_node(&s->nodes[i], vals);
    OrtStatus* st = NULL;
    for (size_t i = 0; i < n_out && !st; i++) {
        int id = val_id(s, out_names[i], 0);
        if (id < 0 || !vals[id].data) { st = mk_status("fake_ort: unknown output %s", out_names[i]); break; }
        if (vals[id].numel != out[i]->t.numel || vals[id].dtype != out[i]->t.dtype) { st = mk_status("fake_ort: output shape mismatch for %s", out_names[i]); break; }
        memcpy(out[i]->t.data, vals[id].data, (size_t)vals[id].numel * dsize(vals[id].dtype));
    }
    if (!st && s->net_index < 3) {
        atomic_fetch_add(&g_run_counts[s->net_index], 1);
        if (g_tap && out[0]->t.dtype == DT_F32) g_tap(g_tap_ud, s->net_index, (const float*)out[0]->t.data, (size_t)out[0]->t.numel);
    }
    for (int i = 0; i < s->nvals; i++) t_free(&vals[i]);
    free(vals);
    return st;
}
static const char* api_ErrMsg(const OrtStatus* st) { return st->msg; }
static void api_RelStatus(OrtStatus* st) { free(st); }
static void api_RelValue(OrtValue* v) { free(v); }
static void api_RelMemInfo(OrtMemoryInfo* m) { free(m); }
static void api_RelSession(OrtSession* s) {
    if (!s) return;
    for (int i = 0; i < s->nvals; i++) { free(s->val_names[i]); if (s->is_const[i]) t_free(&s->consts[i]); }
    for (int i = 0; i < s->nnodes; i++) { free(s->nodes[i].op_name); if (s->nodes[i].has_value) t_free(&s->nodes[i].value); }
    free(s->val_names); free(s->consts); free(s->is_const); free(s->nodes); free(s);
}
static void api_RelOpts(OrtSessionOptions* o) { free(o); }
static void api_RelEnv(OrtEnv* e) { free(e); }

static const OrtApi g_api = {
    api_CreateEnv, api_CreateSessionOptions, api_SetIntra, api_SetInter, api_CreateSessionFromArray,
    api_InCount, api_OutCount, api_InInfo, api_OutInfo, api_Cast, api_DimCount, api_Dims, api_RelTypeInfo,
    api_CpuMemInfo, api_CreateTensor, api_Run, api_ErrMsg, api_RelStatus, api_RelValue, api_RelMemInfo,
    api_RelSession, api_RelOpts, api_RelEnv
};
static const OrtApi* base_GetApi(uint32_t version) { (void)version; return &g_api; }
static const char* base_Version(void) { return "fake-ort (CPU fp32 restatement of ORT 1.13.1 semantics for april-asr)"; }
static const OrtApiBase g_base = { base_GetApi, base_Version };
const OrtApiBase* OrtGetApiBase(void) { return &g_base; }
