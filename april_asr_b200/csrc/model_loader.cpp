// See model_loader.h. Container layout: /root/reference/extra/file-format.md:11-83 and
// src/file/model_file.c:58-129; params: src/params.c:46-112 (same validation ranges :71-82).
#include "model_loader.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <memory>
#include <set>
#include <sstream>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace april {
namespace {

// ------------------------------------------------------------------------------------------------
// little-endian cursor over a byte buffer
// ------------------------------------------------------------------------------------------------
struct Bytes {
    const uint8_t* p = nullptr;
    const uint8_t* end = nullptr;
    bool ok = true;
    size_t left() const { return (size_t)(end - p); }
    template <class T> T rd() {
        T v{};
        if (left() < sizeof(T)) { ok = false; p = end; return v; }
        std::memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    std::string str(size_t n) {
        if (left() < n) { ok = false; p = end; return {}; }
        std::string s((const char*)p, n);
        p += n;
        return s;
    }
};

// ------------------------------------------------------------------------------------------------
// protobuf wire reader (only what onnx.proto3 needs)
// ------------------------------------------------------------------------------------------------
struct Pb {
    const uint8_t* p; const uint8_t* end; bool ok = true;
    Pb(const uint8_t* b, const uint8_t* e) : p(b), end(e) {}
    uint64_t varint() {
        uint64_t v = 0; int sh = 0;
        while (p < end) { uint8_t b = *p++; v |= (uint64_t)(b & 0x7F) << sh; if (!(b & 0x80)) return v; sh += 7; if (sh > 63) break; }
        ok = false; return v;
    }
    // returns field number or 0 at end / error
    int next(int& wt, uint64_t& val, Pb& sub) {
        if (p >= end || !ok) return 0;
        uint64_t key = varint();
        wt = (int)(key & 7);
        int field = (int)(key >> 3);
        switch (wt) {
            case 0: val = varint(); break;
            case 1: if (end - p < 8) { ok = false; return 0; } std::memcpy(&val, p, 8); p += 8; break;
            case 5: { if (end - p < 4) { ok = false; return 0; } uint32_t v; std::memcpy(&v, p, 4); val = v; p += 4; break; }
            case 2: { uint64_t n = varint(); if ((uint64_t)(end - p) < n) { ok = false; return 0; } sub = Pb(p, p + n); p += n; break; }
            default: ok = false; return 0;
        }
        return ok ? field : 0;
    }
    std::string as_string() const { return std::string((const char*)p, (size_t)(end - p)); }
};

enum { ONNX_FLOAT = 1, ONNX_INT32 = 6, ONNX_INT64 = 7 };

struct OTensor {
    std::string name;
    int dtype = ONNX_FLOAT;
    std::vector<int64_t> dims;
    const uint8_t* raw = nullptr;   // view into the (mmap'ed) file: raw_data, little-endian, possibly unaligned
    std::vector<float> fl;          // float_data encoding (rare)
    std::vector<int64_t> i;         // INT64 / INT32 values (small tensors only)
    int64_t numel() const { int64_t n = 1; for (auto d : dims) n *= d; return n; }
    size_t fsize() const { return dtype == ONNX_FLOAT ? (size_t)numel() : 0; }
    float at(size_t k) const { if (raw) { float x; std::memcpy(&x, raw + 4 * k, 4); return x; } return fl[k]; }
    std::vector<float> copy_f() const {
        std::vector<float> v(fsize());
        if (raw) std::memcpy(v.data(), raw, v.size() * 4); else v = fl;
        return v;
    }
    // [rows, cols] row-major -> [cols, rows]
    std::vector<float> copy_f_transposed(int rows, int cols) const {
        std::vector<float> v((size_t)rows * cols);
        for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) v[(size_t)c * rows + r] = at((size_t)r * cols + c);
        return v;
    }
};

struct ONode {
    std::string op, name;
    std::vector<std::string> in, out;
    std::map<std::string, std::vector<int64_t>> ints;   // INT / INTS attributes
    std::map<std::string, float> floats;
    std::shared_ptr<OTensor> value;                      // Constant
};

struct OGraph {
    std::vector<ONode> nodes;
    std::map<std::string, OTensor> inits;
    std::vector<std::pair<std::string, std::vector<int64_t>>> inputs, outputs;   // non-initializer inputs
};

bool parse_tensor(Pb c, OTensor& t) {
    int wt; uint64_t v; Pb sub(nullptr, nullptr); int f;
    const uint8_t* raw = nullptr; size_t raw_len = 0; bool has_raw = false;
    while ((f = c.next(wt, v, sub))) {
        if (f == 1) { if (wt == 0) t.dims.push_back((int64_t)v); else while (sub.p < sub.end && sub.ok) t.dims.push_back((int64_t)sub.varint()); }
        else if (f == 2) t.dtype = (int)v;
        else if (f == 8) t.name = sub.as_string();
        else if (f == 9) { raw = sub.p; raw_len = (size_t)(sub.end - sub.p); has_raw = true; }
        else if (f == 4) { if (wt == 5) { uint32_t u = (uint32_t)v; float x; std::memcpy(&x, &u, 4); t.fl.push_back(x); } else { size_t n = (size_t)(sub.end - sub.p) / 4; size_t o = t.fl.size(); t.fl.resize(o + n); std::memcpy(t.fl.data() + o, sub.p, n * 4); } }
        else if (f == 7) { if (wt == 0) t.i.push_back((int64_t)v); else while (sub.p < sub.end && sub.ok) t.i.push_back((int64_t)sub.varint()); }
        else if (f == 5) { if (wt == 0) t.i.push_back((int64_t)(int32_t)v); else while (sub.p < sub.end && sub.ok) t.i.push_back((int64_t)(int32_t)sub.varint()); }
    }
    if (!c.ok) return false;
    int64_t n = 1;
    for (auto d : t.dims) { if (d < 0 || d > ((int64_t)1 << 31)) return false; n *= d; if (n > ((int64_t)1 << 40)) return false; }   // no negative / overflowing shapes
    if (has_raw) {
        if (t.dtype == ONNX_FLOAT) { if ((int64_t)raw_len != n * 4) return false; t.raw = raw; }
        else if (t.dtype == ONNX_INT64) { if ((int64_t)raw_len != n * 8) return false; t.i.resize((size_t)n); std::memcpy(t.i.data(), raw, raw_len); }
        else if (t.dtype == ONNX_INT32) { if ((int64_t)raw_len != n * 4) return false; t.i.resize((size_t)n); for (int64_t k = 0; k < n; k++) { int32_t x; std::memcpy(&x, raw + 4 * k, 4); t.i[(size_t)k] = x; } }
        else return false;
    } else if (t.dtype == ONNX_FLOAT && (int64_t)t.fl.size() != n) return false;
    return true;
}

bool parse_value_info(Pb c, std::string& name, std::vector<int64_t>& dims) {
    int wt; uint64_t v; Pb sub(nullptr, nullptr); int f;
    while ((f = c.next(wt, v, sub))) {
        if (f == 1) name = sub.as_string();
        else if (f == 2) {
            Pb c2 = sub, s2(nullptr, nullptr); int f2;
            while ((f2 = c2.next(wt, v, s2))) if (f2 == 1) {
                Pb c3 = s2, s3(nullptr, nullptr); int f3;
                while ((f3 = c3.next(wt, v, s3))) if (f3 == 2) {
                    Pb c4 = s3, s4(nullptr, nullptr); int f4;
                    while ((f4 = c4.next(wt, v, s4))) if (f4 == 1) {
                        Pb c5 = s4, s5(nullptr, nullptr); int f5; int64_t dv = -1;
                        while ((f5 = c5.next(wt, v, s5))) if (f5 == 1) dv = (int64_t)v;
                        dims.push_back(dv);
                    }
                }
            }
        }
    }
    return c.ok;
}

bool parse_node(Pb c, ONode& n) {
    int wt; uint64_t v; Pb sub(nullptr, nullptr); int f;
    while ((f = c.next(wt, v, sub))) {
        if (f == 1) n.in.push_back(sub.as_string());
        else if (f == 2) n.out.push_back(sub.as_string());
        else if (f == 3) n.name = sub.as_string();
        else if (f == 4) n.op = sub.as_string();
        else if (f == 5) {
            Pb a = sub, s2(nullptr, nullptr); int f2; std::string an; std::vector<int64_t> iv; bool has_f = false; float fv = 0;
            std::shared_ptr<OTensor> tv;
            while ((f2 = a.next(wt, v, s2))) {
                if (f2 == 1) an = s2.as_string();
                else if (f2 == 2) { uint32_t u = (uint32_t)v; std::memcpy(&fv, &u, 4); has_f = true; }
                else if (f2 == 3) iv.push_back((int64_t)v);
                else if (f2 == 5) { tv = std::make_shared<OTensor>(); if (!parse_tensor(s2, *tv)) return false; }
                else if (f2 == 8) { if (wt == 0) iv.push_back((int64_t)v); else while (s2.p < s2.end && s2.ok) iv.push_back((int64_t)s2.varint()); }
            }
            if (!a.ok) return false;
            if (tv) n.value = tv;
            else if (has_f) n.floats[an] = fv;
            else n.ints[an] = iv;
        }
    }
    return c.ok;
}

bool parse_model(const uint8_t* data, size_t len, OGraph& g, std::string& err) {
    Pb c(data, data + len), graph(nullptr, nullptr), sub(nullptr, nullptr); int wt; uint64_t v; int f; bool has_graph = false;
    while ((f = c.next(wt, v, sub))) if (f == 7) { graph = sub; has_graph = true; }
    if (!c.ok || !has_graph) { err = "ONNX: malformed ModelProto"; return false; }
    c = graph;
    std::vector<std::pair<std::string, std::vector<int64_t>>> ins;
    while ((f = c.next(wt, v, sub))) {
        if (f == 1) { ONode n; if (!parse_node(sub, n)) { err = "ONNX: malformed NodeProto"; return false; } g.nodes.push_back(std::move(n)); }
        else if (f == 5) { OTensor t; if (!parse_tensor(sub, t)) { err = "ONNX: malformed/unsupported initializer"; return false; } std::string nm = t.name; g.inits[nm] = std::move(t); }
        else if (f == 11 || f == 12) { std::string nm; std::vector<int64_t> d; if (!parse_value_info(sub, nm, d)) { err = "ONNX: malformed ValueInfo"; return false; } (f == 11 ? ins : g.outputs).push_back({ nm, d }); }
    }
    if (!c.ok) { err = "ONNX: malformed GraphProto"; return false; }
    for (auto& i : ins) if (!g.inits.count(i.first)) g.inputs.push_back(i);
    return true;
}

// ------------------------------------------------------------------------------------------------
// topology-driven weight extraction
// ------------------------------------------------------------------------------------------------
struct Linear { std::vector<float> w; std::vector<float> b; int in = 0, out = 0; bool from_h = false; std::string x_name, y_name; };

struct GraphView {
    const OGraph& g;
    std::map<std::string, const ONode*> producer;
    std::map<std::string, std::vector<const ONode*>> consumers;
    explicit GraphView(const OGraph& gr) : g(gr) {
        for (auto& n : g.nodes) { for (auto& o : n.out) producer[o] = &n; for (auto& i : n.in) consumers[i].push_back(&n); }
    }
    // constant tensor behind a value name (initializer or Constant node), or null
    const OTensor* constant(const std::string& name) const {
        auto it = g.inits.find(name); if (it != g.inits.end()) return &it->second;
        auto p = producer.find(name); if (p != producer.end() && p->second->op == "Constant" && p->second->value) return p->second->value.get();
        return nullptr;
    }
    // true if `name` is graph input `root` seen only through shape-only ops
    bool derives_from(const std::string& name, const std::string& root) const {
        static const std::set<std::string> pass = { "Slice", "Squeeze", "Unsqueeze", "Gather", "Reshape", "Transpose", "Identity", "Flatten" };
        std::string cur = name;
        for (int hop = 0; hop < 16; hop++) {
            if (cur == root) return true;
            auto p = producer.find(cur); if (p == producer.end()) return false;
            if (!pass.count(p->second->op)) return false;
            cur = p->second->in.empty() ? "" : p->second->in[0];
        }
        return false;
    }
};

// Collect Linear layers in graph order: Gemm(x, W[out,in] (transB=1), b) or MatMul(x, Wt[in,out]) (+ Add(bias)).
bool collect_linears(const GraphView& gv, const std::string& h_input, std::vector<Linear>& out, std::string& err) {
    for (auto& n : gv.g.nodes) {
        if ((n.op == "Gemm" || n.op == "MatMul") && (n.in.size() < 2 || n.out.empty())) { err = "ONNX: " + n.op + " node with missing inputs / outputs"; return false; }
        if (n.op == "Gemm") {
            const OTensor* w = n.in.size() > 1 ? gv.constant(n.in[1]) : nullptr;
            if (!w || w->dims.size() != 2 || w->dtype != ONNX_FLOAT) continue;
            if (w->dims[0] <= 0 || w->dims[1] <= 0 || w->dims[0] > (1 << 20) || w->dims[1] > (1 << 20)) { err = "ONNX: Gemm weight dims out of range"; return false; }
            auto ta = n.ints.find("transA"), tb = n.ints.find("transB");
            bool transA = ta != n.ints.end() && !ta->second.empty() && ta->second[0];
            bool transB = tb != n.ints.end() && !tb->second.empty() && tb->second[0];
            float alpha = n.floats.count("alpha") ? n.floats.at("alpha") : 1.f, beta = n.floats.count("beta") ? n.floats.at("beta") : 1.f;
            if (transA || alpha != 1.f || beta != 1.f) { err = "ONNX: unsupported Gemm attributes"; return false; }
            Linear l;
            if (transB) { l.out = (int)w->dims[0]; l.in = (int)w->dims[1]; l.w = w->copy_f(); }
            else { l.in = (int)w->dims[0]; l.out = (int)w->dims[1]; l.w = w->copy_f_transposed(l.in, l.out); }
            if (n.in.size() > 2 && !n.in[2].empty()) { const OTensor* b = gv.constant(n.in[2]); if (!b || (int)b->fsize() != l.out) { err = "ONNX: Gemm bias not a constant vector"; return false; } l.b = b->copy_f(); }
            l.from_h = !h_input.empty() && gv.derives_from(n.in[0], h_input);
            l.x_name = n.in[0]; l.y_name = n.out[0];
            out.push_back(std::move(l));
        } else if (n.op == "MatMul") {
            const OTensor* w = n.in.size() > 1 ? gv.constant(n.in[1]) : nullptr;
            if (!w || w->dims.size() != 2 || w->dtype != ONNX_FLOAT) continue;
            if (w->dims[0] <= 0 || w->dims[1] <= 0 || w->dims[0] > (1 << 20) || w->dims[1] > (1 << 20)) { err = "ONNX: MatMul weight dims out of range"; return false; }
            Linear l; l.in = (int)w->dims[0]; l.out = (int)w->dims[1];
            l.w = w->copy_f_transposed(l.in, l.out);
            l.from_h = !h_input.empty() && gv.derives_from(n.in[0], h_input);
            l.x_name = n.in[0]; l.y_name = n.out[0];
            // bias: a single Add consumer with a constant 1-D operand of size out
            auto cs = gv.consumers.find(n.out[0]);
            if (cs != gv.consumers.end() && cs->second.size() == 1 && cs->second[0]->op == "Add") {
                const ONode* a = cs->second[0];
                for (auto& nm : a->in) { if (nm == n.out[0]) continue; const OTensor* b = gv.constant(nm); if (b && (int)b->fsize() == l.out) { l.b = b->copy_f(); l.y_name = a->out[0]; } }
            }
            out.push_back(std::move(l));
        }
    }
    return true;
}

// BasicNorm eps constants in graph order: Add(ReduceMean(.), const scalar)
std::vector<float> collect_norm_eps(const GraphView& gv) {
    std::vector<float> eps;
    for (auto& n : gv.g.nodes) {
        if (n.op != "Add" || n.in.size() != 2) continue;
        for (int k = 0; k < 2; k++) {
            auto p = gv.producer.find(n.in[k]);
            if (p == gv.producer.end() || p->second->op != "ReduceMean") continue;
            const OTensor* c = gv.constant(n.in[1 - k]);
            if (c && c->fsize() == 1) eps.push_back(c->at(0));
        }
    }
    return eps;
}

#define FAIL(msg) do { err = (msg); return false; } while (0)

bool extract_encoder(const OGraph& g, ModelWeights& W, std::string& err) {
    GraphView gv(g);
    if (g.inputs.size() != 3 || g.outputs.size() != 3) FAIL("encoder: expected 3 inputs / 3 outputs (april_model.c:63-64)");
    // inputs by name as the reference binds them (april_session.c:121-122)
    const std::vector<int64_t>* xd = nullptr; const std::vector<int64_t>* hd = nullptr; const std::vector<int64_t>* cd = nullptr;
    for (auto& i : g.inputs) { if (i.first == "x") xd = &i.second; else if (i.first == "h") hd = &i.second; else if (i.first == "c") cd = &i.second; }
    if (!xd || !hd || !cd || xd->size() != 3 || hd->size() != 3 || cd->size() != 3) FAIL("encoder: inputs x,h,c with static 3-D shapes required");
    ModelDims& d = W.d;
    if ((*xd)[0] != 1 || (*hd)[1] != 1 || (*cd)[1] != 1) FAIL("encoder: graphs must be exported with batch 1 (params.c:71)");
    for (auto* v : { xd, hd, cd }) for (auto e : *v) if (e <= 0 || e > (1 << 20)) FAIL("encoder: input dims must be positive static sizes");
    d.SEG = (int)(*xd)[1]; d.MEL = (int)(*xd)[2]; d.L = (int)(*hd)[0]; d.D = (int)(*hd)[2]; d.H = (int)(*cd)[2];
    if ((*cd)[0] != d.L) FAIL("encoder: h/c layer count mismatch");
    if (d.L > 64 || d.SEG < 7) FAIL("encoder: layer count / segment size out of range");

    // convolutions, in order
    std::vector<const ONode*> convs;
    for (auto& n : g.nodes) if (n.op == "Conv") convs.push_back(&n);
    if (convs.size() != 3) FAIL("encoder: expected 3 Conv nodes (Conv2dSubsampling)");
    const int want_stride[3] = { 1, 2, 2 };
    std::vector<float>* cw[3] = { &W.conv1_w, &W.conv2_w, &W.conv3_w }; std::vector<float>* cb[3] = { &W.conv1_b, &W.conv2_b, &W.conv3_b };
    int cout[3], cin[3];
    for (int k = 0; k < 3; k++) {
        if (convs[k]->in.size() < 2 || convs[k]->out.empty()) FAIL("encoder: Conv node with missing inputs / outputs");
        const OTensor* w = gv.constant(convs[k]->in[1]);
        const OTensor* b = convs[k]->in.size() > 2 ? gv.constant(convs[k]->in[2]) : nullptr;
        if (!w || w->dims.size() != 4 || w->dims[2] != 3 || w->dims[3] != 3 || !b) FAIL("encoder: Conv must be 3x3 with constant weight and bias");
        if (w->dtype != ONNX_FLOAT || b->dtype != ONNX_FLOAT || w->dims[0] <= 0 || w->dims[1] <= 0 || w->dims[0] > 4096 || w->dims[1] > 4096) FAIL("encoder: Conv weights must be float with sane channel counts");
        auto st = convs[k]->ints.find("strides"); int s = (st != convs[k]->ints.end() && !st->second.empty()) ? (int)st->second[0] : 1;
        auto pd = convs[k]->ints.find("pads"); if (pd != convs[k]->ints.end()) for (auto p : pd->second) if (p != 0) FAIL("encoder: Conv padding unsupported");
        auto gr = convs[k]->ints.find("group"); if (gr != convs[k]->ints.end() && !gr->second.empty() && gr->second[0] != 1) FAIL("encoder: grouped Conv unsupported");
        if (s != want_stride[k]) FAIL("encoder: unexpected Conv strides (want 1,2,2)");
        cout[k] = (int)w->dims[0]; cin[k] = (int)w->dims[1];
        *cw[k] = w->copy_f(); *cb[k] = b->copy_f();
        if ((int)b->fsize() != cout[k]) FAIL("encoder: Conv bias size");
    }
    if (cin[0] != 1 || cin[1] != cout[0] || cin[2] != cout[1]) FAIL("encoder: Conv channel chain mismatch");
    d.C1 = cout[0]; d.C2 = cout[1]; d.C3 = cout[2];
    int t1 = d.SEG - 2, f1 = d.MEL - 2, t2 = (t1 - 3) / 2 + 1, f2 = (f1 - 3) / 2 + 1, t3 = (t2 - 3) / 2 + 1, f3 = (f2 - 3) / 2 + 1;
    if (t3 != 1 || f3 <= 0) FAIL("encoder: segment size must reduce to exactly one output frame (and at least one frequency bin)");
    d.F3 = f3;

    std::vector<Linear> lin;
    if (!collect_linears(gv, "h", lin, err)) return false;
    if ((int)lin.size() != 2 + 5 * d.L) { std::ostringstream os; os << "encoder: found " << lin.size() << " linear layers, expected " << 2 + 5 * d.L; FAIL(os.str()); }
    if (lin[0].in != d.C3 * d.F3 || lin[0].out != d.D || lin[0].b.empty()) FAIL("encoder: subsampling output Linear shape");
    W.out_w = std::move(lin[0].w); W.out_b = std::move(lin[0].b);
    W.layers.resize((size_t)d.L);
    for (int l = 0; l < d.L; l++) {
        Linear* q = &lin[(size_t)(1 + 5 * l)];
        // the two gate projections may appear in either order; W_hh is the one fed (directly) by `h`
        Linear *ih = &q[0], *hh = &q[1];
        if (q[0].from_h && !q[1].from_h) std::swap(ih, hh);
        else if (!(q[1].from_h && !q[0].from_h)) FAIL("encoder: cannot tell W_ih from W_hh (neither/both fed by input h)");
        LayerWeights& lw = W.layers[(size_t)l];
        if (ih->in != d.D || ih->out != 4 * d.H || hh->in != d.D || hh->out != 4 * d.H) FAIL("encoder: LSTM gate projection shapes");
        if (q[2].in != d.H || q[2].out != d.D || !q[2].b.empty()) FAIL("encoder: LSTM output projection W_hr shape (no bias expected)");
        if (l == 0) d.F = q[3].out;
        if (q[3].in != d.D || q[3].out != d.F || q[4].in != d.F || q[4].out != d.D || q[3].b.empty() || q[4].b.empty()) FAIL("encoder: feed-forward shapes");
        lw.w_ih = std::move(ih->w); lw.b_ih = ih->b.empty() ? std::vector<float>((size_t)4 * d.H, 0.f) : ih->b;
        lw.w_hh = std::move(hh->w); lw.b_hh = hh->b.empty() ? std::vector<float>((size_t)4 * d.H, 0.f) : hh->b;
        lw.w_hr = std::move(q[2].w); lw.ff1_w = std::move(q[3].w); lw.ff1_b = q[3].b; lw.ff2_w = std::move(q[4].w); lw.ff2_b = q[4].b;
    }
    Linear& proj = lin.back();
    if (proj.in != d.D || proj.b.empty()) FAIL("encoder: encoder_proj shape");
    d.J = proj.out; W.enc_proj_w = std::move(proj.w); W.enc_proj_b = proj.b;

    // gate order check: Split output #2 must feed Tanh (PyTorch order i,f,g,o)
    int splits = 0;
    for (auto& n : g.nodes) if (n.op == "Split" && n.out.size() == 4) {
        splits++;
        auto cs = gv.consumers.find(n.out[2]);
        if (cs == gv.consumers.end() || cs->second.empty() || cs->second[0]->op != "Tanh") FAIL("encoder: LSTM gate order is not i,f,g,o");
    }
    if (splits != d.L) FAIL("encoder: expected one 4-way gate Split per layer");
    // activation census: DoubleSwish = x * sigmoid(x - 1): 3 (conv) + L (ff) of them, plus 3 L gate sigmoids
    int n_sig = 0, n_tanh = 0; for (auto& n : g.nodes) { n_sig += n.op == "Sigmoid"; n_tanh += n.op == "Tanh"; }
    if (n_sig != 3 + d.L + 3 * d.L || n_tanh != 2 * d.L) FAIL("encoder: activation census does not match lstm_transducer_stateless2");

    std::vector<float> eps = collect_norm_eps(gv);
    if ((int)eps.size() != 1 + d.L) FAIL("encoder: expected 1 + L BasicNorm eps constants");
    W.out_norm_eps = eps[0];
    for (int l = 0; l < d.L; l++) W.layers[(size_t)l].norm_eps = eps[(size_t)l + 1];
    return true;
}

bool extract_decoder(const OGraph& g, ModelWeights& W, std::string& err) {
    GraphView gv(g);
    if (g.inputs.size() != 1 || g.outputs.size() != 1 || g.inputs[0].second.size() != 2) FAIL("decoder: expected input context (N, ctx)");
    if (g.inputs[0].second[0] != 1) FAIL("decoder: only batch size 1 graphs are supported (april_session.c:51-55)");
    ModelDims& d = W.d;
    d.CTX = (int)g.inputs[0].second[1];
    const ONode* gather = nullptr; const ONode* conv = nullptr;
    for (auto& n : g.nodes) { if (n.op == "Gather" && !gather) gather = &n; if (n.op == "Conv" && !conv) conv = &n; }
    if (!gather) FAIL("decoder: no embedding Gather");
    if (gather->in.size() < 2) FAIL("decoder: Gather node with missing inputs");
    const OTensor* emb = gv.constant(gather->in[0]);
    if (!emb || emb->dims.size() != 2) FAIL("decoder: embedding table must be a 2-D constant");
    if (emb->dtype != ONNX_FLOAT) FAIL("decoder: embedding must be float");
    if (emb->dims[0] <= 0 || emb->dims[1] <= 0 || emb->dims[0] > (1 << 20) || emb->dims[1] > (1 << 16)) FAIL("decoder: embedding dims out of range");
    if (d.CTX <= 0 || d.CTX > 16) FAIL("decoder: context size out of range");
    d.V = (int)emb->dims[0]; d.E = (int)emb->dims[1]; W.emb = emb->copy_f();
    if (d.CTX > 1) {
        if (!conv) FAIL("decoder: context_size > 1 needs the Conv1d");
        if (conv->in.size() < 2) FAIL("decoder: Conv node with missing inputs");
        const OTensor* w = gv.constant(conv->in[1]);
        auto gr = conv->ints.find("group"); int group = (gr != conv->ints.end() && !gr->second.empty()) ? (int)gr->second[0] : 1;
        if (w && w->dtype != ONNX_FLOAT) FAIL("decoder: Conv weight must be float");
        if (!w || w->dims.size() != 3 || w->dims[0] != d.E || w->dims[1] != 1 || w->dims[2] != d.CTX || group != d.E || conv->in.size() > 2)
            FAIL("decoder: only the depthwise, bias-free Conv1d(kernel = context_size) is supported");
        W.dec_conv_w = w->copy_f();
    } else {
        W.dec_conv_w.assign((size_t)d.E, 1.f);
    }
    std::vector<Linear> lin;
    if (!collect_linears(gv, "", lin, err)) return false;
    if (lin.size() != 1 || lin[0].in != d.E || lin[0].b.empty()) FAIL("decoder: expected one decoder_proj Linear");
    if (lin[0].out != d.J) FAIL("decoder: decoder_out dim != encoder_out dim");
    W.dec_proj_w = lin[0].w; W.dec_proj_b = lin[0].b;
    int n_relu = 0; for (auto& n : g.nodes) n_relu += n.op == "Relu";
    if (n_relu != 1) FAIL("decoder: expected exactly one Relu");
    return true;
}

bool extract_joiner(const OGraph& g, ModelWeights& W, std::string& err) {
    GraphView gv(g);
    if (g.inputs.size() != 2 || g.outputs.size() != 1) FAIL("joiner: expected 2 inputs / 1 output (april_model.c:69-70)");
    std::vector<Linear> lin;
    if (!collect_linears(gv, "", lin, err)) return false;
    if (lin.size() != 1 || lin[0].in != W.d.J || lin[0].out != W.d.V || lin[0].b.empty()) FAIL("joiner: expected one output Linear (J -> V)");
    int n_tanh = 0, n_soft = 0; for (auto& n : g.nodes) { n_tanh += n.op == "Tanh"; n_soft += (n.op == "Softmax" || n.op == "LogSoftmax"); }
    if (n_tanh != 1 || n_soft != 0) FAIL("joiner: expected tanh(enc + dec) -> Linear, no softmax");
    W.join_w = lin[0].w; W.join_b = lin[0].b;
    return true;
}

bool read_params(Bytes& b, Params& p, std::string& err) {
    if (b.left() < 8 || std::memcmp(b.p, "PARAMS\0\0", 8) != 0) FAIL("params: bad magic");
    b.p += 8;
    int32_t* dst[] = { &p.batch_size, &p.segment_size, &p.segment_step, &p.mel_features, &p.sample_rate, &p.frame_shift_ms,
                       &p.frame_length_ms, &p.round_pow2, &p.mel_low, &p.mel_high, &p.snip_edges, &p.token_count, &p.blank_id };
    for (auto* q : dst) *q = b.rd<int32_t>();
    if (!b.ok) FAIL("params: truncated");
    p.round_pow2 = p.round_pow2 != 0; p.snip_edges = p.snip_edges != 0;
    // same acceptance ranges as src/params.c:71-82
    bool ok = p.batch_size == 1 && p.segment_size > 0 && p.segment_size < 100 && p.segment_step > 0 && p.segment_step < 100 &&
              p.segment_step <= p.segment_size && p.mel_features > 0 && p.mel_features < 256 && p.sample_rate > 0 && p.sample_rate < 144000 &&
              p.token_count > 0 && p.token_count < 16384 && p.blank_id >= 0 && p.blank_id < p.token_count && p.frame_shift_ms > 0 &&
              p.frame_shift_ms <= p.frame_length_ms && p.frame_length_ms > 0 && p.frame_length_ms <= 5000 && p.mel_low > 0 &&
              p.mel_low < p.sample_rate && (p.mel_high == 0 || p.mel_high > p.mel_low);
    if (!ok) FAIL("params: value out of the accepted range (src/params.c:71-82)");
    if (p.frame_shift_ms * p.sample_rate / 1000 <= 0 || p.frame_length_ms * p.sample_rate / 1000 < 2) FAIL("params: frame shift / length of less than one sample at this sample rate");
    p.tokens.resize((size_t)p.token_count);
    for (int i = 0; i < p.token_count; i++) {
        int32_t n = b.rd<int32_t>();
        if (!b.ok || n < 0 || (size_t)n > b.left()) FAIL("params: truncated token table");
        p.tokens[(size_t)i] = b.str((size_t)n);
    }
    return true;
}

}  // namespace

bool load_april_file(const std::string& path, LoadedModel& out, std::string& err) {
    struct Map {
        const uint8_t* p = nullptr; size_t n = 0; int fd = -1;
        ~Map() { if (p) munmap((void*)p, n); if (fd >= 0) close(fd); }
        const uint8_t* data() const { return p; }
        size_t size() const { return n; }
    } buf;
    buf.fd = open(path.c_str(), O_RDONLY);
    if (buf.fd < 0) FAIL("cannot open " + path);
    struct stat st;
    if (fstat(buf.fd, &st) != 0 || st.st_size < 20) FAIL("file too small");
    buf.n = (size_t)st.st_size;
    void* mp = mmap(nullptr, buf.n, PROT_READ, MAP_PRIVATE, buf.fd, 0);
    if (mp == MAP_FAILED) FAIL("mmap failed");
    buf.p = (const uint8_t*)mp;
    Bytes b{ buf.data(), buf.data() + buf.size() };
    if (std::memcmp(b.p, "APRILMDL", 8) != 0) FAIL("magic check failed");
    b.p += 8;
    uint32_t version = b.rd<uint32_t>();
    if (version != 1) FAIL("unsupported model version");
    uint64_t header_size = b.rd<uint64_t>(); (void)header_size;
    out.language = b.str(8); out.language = out.language.c_str();   // cut at first NUL, like language[8]='\0' + strcpy
    uint64_t n = b.rd<uint64_t>(); if (!b.ok || n > b.left()) FAIL("header: bad name length"); out.name = b.str((size_t)n);
    n = b.rd<uint64_t>(); if (!b.ok || n > b.left()) FAIL("header: bad description length"); out.description = b.str((size_t)n);
    uint32_t type = b.rd<uint32_t>();
    if (type != 1) FAIL("unknown model type (only LSTM_TRANSDUCER_STATELESS = 1)");
    uint64_t params_off = b.rd<uint64_t>(), params_size = b.rd<uint64_t>();
    uint64_t nnet = b.rd<uint64_t>();
    if (!b.ok || nnet != 3) FAIL("wrong number of networks (april_model.c:36)");
    uint64_t off[3], len[3];
    for (int i = 0; i < 3; i++) { off[i] = b.rd<uint64_t>(); len[i] = b.rd<uint64_t>(); if (!b.ok || off[i] + len[i] > buf.size() || off[i] + len[i] < off[i]) FAIL("network out of bounds of file"); }
    if (params_off + params_size > buf.size() || params_off + params_size < params_off) FAIL("params out of bounds of file");
    // The reference reads params from wherever the fd sits after reading network[2] (april_model.c:61);
    // every file the exporter writes has params_offset == that position, so honour the header field.
    Bytes pb{ buf.data() + params_off, buf.data() + params_off + params_size };
    if (!read_params(pb, out.params, err)) return false;

    OGraph ge, gd, gj;
    if (!parse_model(buf.data() + off[0], (size_t)len[0], ge, err)) return false;
    if (!parse_model(buf.data() + off[1], (size_t)len[1], gd, err)) return false;
    if (!parse_model(buf.data() + off[2], (size_t)len[2], gj, err)) return false;
    if (!extract_encoder(ge, out.w, err)) return false;
    if (!extract_decoder(gd, out.w, err)) return false;
    if (!extract_joiner(gj, out.w, err)) return false;
    ModelDims& d = out.w.d;
    d.STEP = out.params.segment_step;
    // same cross-checks as april_model.c:99-102
    if (d.SEG != out.params.segment_size || d.MEL != out.params.mel_features) FAIL("x dims do not match params (april_model.c:99-101)");
    if (d.V != out.params.token_count) FAIL("logits dim does not match token_count (april_model.c:102)");
    if (d.CTX != 2) FAIL("only context_size 2 is supported (the reference's greedy search hard-codes it: april_session.c:322,326)");
    return true;
}

void model_digest(const LoadedModel& m, uint64_t out[16]) {
    const ModelDims& d = m.w.d;
    const int dims[15] = { d.L, d.D, d.H, d.F, d.J, d.E, d.V, d.CTX, d.MEL, d.SEG, d.STEP, d.C1, d.C2, d.C3, d.F3 };
    for (int i = 0; i < 15; i++) out[i] = (uint64_t)dims[i];
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } };
    auto vec = [&](const std::vector<float>& v) { const uint64_t n = v.size(); mix(&n, 8); if (n) mix(v.data(), n * 4); };
    const ModelWeights& w = m.w;
    for (auto* v : { &w.conv1_w, &w.conv1_b, &w.conv2_w, &w.conv2_b, &w.conv3_w, &w.conv3_b, &w.out_w, &w.out_b }) vec(*v);
    mix(&w.out_norm_eps, 4);
    for (auto& l : w.layers) { for (auto* v : { &l.w_ih, &l.b_ih, &l.w_hh, &l.b_hh, &l.w_hr, &l.ff1_w, &l.ff1_b, &l.ff2_w, &l.ff2_b }) vec(*v); mix(&l.norm_eps, 4); }
    for (auto* v : { &w.enc_proj_w, &w.enc_proj_b, &w.emb, &w.dec_conv_w, &w.dec_proj_w, &w.dec_proj_b, &w.join_w, &w.join_b }) vec(*v);
    out[15] = h;
}

}  // namespace april
