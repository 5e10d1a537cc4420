// Hand-written sm_100a kernels, exact-arithmetic tier (fp64 front-end, fp32 SIMT nets).
// These define the device-side ground truth that the tensor-core tier (kernels_tc.cuh) is checked
// against, and they carry the non-GEMM parts of the hot path in every configuration:
//   fbank_kernel          F1/F2  (/root/reference/src/fbank.c:174-306, pocketfft rfft_forward)
//   conv_embed_kernel     E1     (icefall Conv2dSubsampling, extra/export-april.py:183-203 [external])
//   gemm_simt_kernel      E1-E4  dense projections, fp32 FMA
//   lstm_gates_kernel     E2     gate pre-activations + cell update
//   lstm_proj_kernel      E2     h' = W_hr (o * tanh(c'))
//   add_rows / basic_norm E2/E3  residual, BasicNorm
//   decoder_kernel        D1     (april_session.c:151-163,181-196)
//   greedy_kernel         J1+G1  (april_session.c:166-179,306-429,431-476)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace april {
namespace k {

// ------------------------------------------------------------------------------------------------
// vocabulary flags used by the greedy rules (april_session.c:338-351,369-374)
// ------------------------------------------------------------------------------------------------
enum : uint8_t { VF_SPACE = 1, VF_EOS = 2, VF_COMMA = 4, VF_DIGIT = 8, VF_DOT = 16 };

#define APRIL_MAX_ACTIVE_TOKENS 72          // src/april_session.h:30

// Per-session greedy-search state that feeds back into the model.
struct GreedyState {
    int32_t ctx0, ctx1;                     // decoder context (context_size == 2)
    int32_t head;                           // active_token_head
    int32_t pad_;
    long long current_ms, last_emission_ms; // size_t in the reference
    int16_t ids[APRIL_MAX_ACTIVE_TOKENS];   // token ids of active_tokens[0..head)
};

struct DecisionRecord {                     // one joiner pass
    int32_t max_idx;
    int32_t is_blank;
    float max_val;
    float blank_val;
};

// ------------------------------------------------------------------------------------------------
// F1/F2: log-mel filterbank, bit-faithful to fbank.c
// ------------------------------------------------------------------------------------------------
struct FbankParams {
    int N;              // padded window (power of two, 512)
    int log2N;
    int shift;          // 160
    int mel;            // 80
    int nbins;          // N/2
    float pad_value;    // (float)log((double)kEps)
    float inv_scale_is_div; // unused
    const float* window;        // [N]
    const double2* twiddle;     // [N/2]  (cos, -sin)(2 pi k / N)
    const int* mel_start;       // [mel] first fft bin with non-zero weight
    const int* mel_count;       // [mel]
    const int* mel_off;         // [mel] offset into mel_w
    const float* mel_w;         // packed non-zero weights
};

// grid (ceil(max_new_rows / 4), B), block 128: one warp per output row.
// New rows of session b are, in order: pad_before[b] pad rows, n_frames[b] frames computed from this
// pass's PCM (frame i = samples [i*shift, i*shift + N)), then pad rows up to n_new[b].
__global__ void __launch_bounds__(128)
fbank_kernel(FbankParams P, const int16_t* __restrict__ pcm, const long long* __restrict__ pcm_off,
             const int* __restrict__ pad_before, const int* __restrict__ n_frames, const int* __restrict__ n_new,
             const int* __restrict__ carry, float* __restrict__ FB, int row_pitch) {
    extern __shared__ double smem_d[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int r = blockIdx.x * 4 + warp;
    if (r >= n_new[b]) return;
    float* out = FB + ((size_t)b * row_pitch + carry[b] + r) * P.mel;
    const int pb = pad_before[b];
    const int src = (r >= pb && r < pb + n_frames[b]) ? r - pb : -1;
    if (src < 0) {                                   // fbank_flush pad rows, fbank.c:312-316
        for (int m = lane; m < P.mel; m += 32) out[m] = P.pad_value;
        return;
    }
    const int N = P.N;
    double* dat = smem_d + (size_t)warp * 3 * N;     // natural order
    double* re = dat + N;
    double* im = re + N;
    const int16_t* x = pcm + pcm_off[b] + (long long)src * P.shift;

    // (float)pcm16 / 32768.0f, april_session.c:520-522; widened to double, fbank.c:228-236
    for (int j = lane; j < N; j += 32) dat[j] = (double)((float)x[j] / 32768.0f);
    __syncwarp();
    // DC removal with a FLOAT accumulator over double data, fbank.c:241-246
    float mean = 0.f;
    if (lane == 0) {
        float sum = 0.f;
        for (int j = 0; j < N; j++) sum = (float)((double)sum + dat[j]);
        mean = sum / (float)N;
    }
    mean = __shfl_sync(0xffffffffu, mean, 0);
    const double dmean = (double)mean;
    const double coeff = (double)0.97f;              // preemph_coeff is a float, fbank.c:188, april_model.c:97
    // pre-emphasis reads un-emphasised neighbours (the reference walks j downwards), fbank.c:249-253
    for (int base = 0; base < N; base += 32 * 8) {
        double cur[8], prev[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            int j = base + u * 32 + lane;
            if (j >= N) { cur[u] = prev[u] = 0.0; continue; }            // windows shorter than 256 samples (N = 64, 128)
            cur[u] = __dsub_rn(dat[j], dmean);
            prev[u] = __dsub_rn(dat[j > 0 ? j - 1 : 0], dmean);
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 8; u++) {
            int j = base + u * 32 + lane;
            if (j >= N) continue;
            double v = __dsub_rn(cur[u], __dmul_rn(coeff, prev[u]));
            v = __dmul_rn(v, (double)P.window[j]);   // fbank.c:256-257
            int rj = __brev((unsigned)j) >> (32 - P.log2N);
            re[rj] = v;
            im[rj] = 0.0;
        }
        __syncwarp();
    }
    // radix-2 DIT FFT in double (pocketfft computes the same DFT in double; fbank.c:259-270)
    for (int s = 1; s <= P.log2N; s++) {
        const int half = 1 << (s - 1);
        const int tw_stride = N >> s;
        for (int j = lane; j < N / 2; j += 32) {
            int pos = j & (half - 1);
            int i0 = ((j >> (s - 1)) << s) + pos;
            int i1 = i0 + half;
            double2 w = P.twiddle[pos * tw_stride];
            double xr = re[i1], xi = im[i1];
            double tr = w.x * xr - w.y * xi;
            double ti = w.x * xi + w.y * xr;
            double ar = re[i0], ai = im[i0];
            re[i0] = ar + tr; im[i0] = ai + ti;
            re[i1] = ar - tr; im[i1] = ai - ti;
        }
        __syncwarp();
    }
    // power in FLOAT, Nyquist dropped, im[0] forced to 0, fbank.c:269-280
    float* pw = (float*)dat;
    for (int kbin = lane; kbin < P.nbins; kbin += 32) {
        float fr = (float)re[kbin];
        float fi = kbin == 0 ? 0.f : (float)im[kbin];
        pw[kbin] = __fadd_rn(__fmul_rn(fr, fr), __fmul_rn(fi, fi));
    }
    __syncwarp();
    // mel: sequential float sum over fft bins (zero weights contribute exact +0), fbank.c:283-296
    for (int m = lane; m < P.mel; m += 32) {
        const int k0 = P.mel_start[m], cnt = P.mel_count[m];
        const float* w = P.mel_w + P.mel_off[m];
        float val = 0.f;
        for (int i = 0; i < cnt; i++) val = __fadd_rn(val, __fmul_rn(pw[k0 + i], w[i]));
        const float kEps = 1.1920928955078125e-07f;
        out[m] = (float)log((double)fmaxf(kEps, val));
    }
}

// ------------------------------------------------------------------------------------------------
// F1/F2 for the 512-sample window every shipped model uses (25 ms at 16 kHz, rounded to a power of two): same arithmetic
// as fbank_kernel wherever the reference's rounding is visible (DC mean, pre-emphasis, window, float power, sequential
// float mel sums, double log), restructured where it is not:
//  * the DC sum. fbank.c:241-246 accumulates 512 doubles into a FLOAT. Every sample is int16 / 32768 (a multiple of 2^-15,
//    april_session.c:520-522) and every partial sum is below 512 in magnitude, i.e. k * 2^-15 with |k| <= 2^24: exactly
//    representable in a float. The 512-step serial chain therefore never rounds, and equals the INTEGER sum of the samples,
//    which a warp reduces in five shuffles. (fbank_kernel runs the chain on one lane: 1 500 of its 6 150 instructions per frame.)
//  * the FFT: 512 = 8 x 8 x 8, three decimation-in-frequency radix-8 passes in fp64 with the butterflies in registers
//    (three shared-memory round trips instead of nine, rows padded by two doubles per eight so that the stride-8 and the
//    16-byte accesses are conflict-free), real input in the first pass, and only bins 0..255 (half of the last pass' outputs)
//    are formed. pocketfft (fbank.c:259-270) computes the same DFT in double; any two correct fp64 FFTs agree to ~1e-16
//    relative, far below the float rounding of the power that follows.
// One warp per frame, 4 frames per CTA. grid (ceil(max_new_rows / 4), B), block 128.
// ------------------------------------------------------------------------------------------------
constexpr int FB3_N = 512;
constexpr int FB3_PITCH = FB3_N + 2 * (FB3_N / 8);                         // doubles per padded array
__device__ __forceinline__ int fb3_phys(int i) { return i + 2 * (i >> 3); }

struct cplx { double x, y; };
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return { a.x + b.x, a.y + b.y }; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return { a.x - b.x, a.y - b.y }; }
__device__ __forceinline__ cplx cmul(cplx a, cplx w) { return { a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x }; }
__device__ __forceinline__ cplx cmul_mi(cplx a) { return { a.y, -a.x }; }   // a * (-i)

// forward 8-point DFT in place: y[m] = sum_k x[k] exp(-2 pi i m k / 8)
__device__ __forceinline__ void dft8(cplx (&v)[8]) {
    const double r = 0.70710678118654752440;
    cplx a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
    cplx b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
    b1 = { (b1.x + b1.y) * r, (b1.y - b1.x) * r };                          // * W8^1 = (1 - i) / sqrt 2
    b2 = cmul_mi(b2);                                                       // * W8^2 = -i
    b3 = { (b3.y - b3.x) * r, -(b3.x + b3.y) * r };                         // * W8^3 = (-1 - i) / sqrt 2
    // even outputs: DFT4 of a;  odd outputs: DFT4 of b
    cplx t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = cmul_mi(csub(a1, a3));
    v[0] = cadd(t0, t2); v[4] = csub(t0, t2); v[2] = cadd(t1, t3); v[6] = csub(t1, t3);
    cplx u0 = cadd(b0, b2), u1 = csub(b0, b2), u2 = cadd(b1, b3), u3 = cmul_mi(csub(b1, b3));
    v[1] = cadd(u0, u2); v[5] = csub(u0, u2); v[3] = cadd(u1, u3); v[7] = csub(u1, u3);
}
// W_512^idx from the half table (cos, -sin)(2 pi k / 512), k < 256
__device__ __forceinline__ cplx fb3_tw(const double2* __restrict__ tw, int idx) {
    idx &= FB3_N - 1;
    const double2 w = __ldg(tw + (idx & (FB3_N / 2 - 1)));
    return idx >= FB3_N / 2 ? cplx{ -w.x, -w.y } : cplx{ w.x, w.y };
}

__global__ void __launch_bounds__(128)
fbank512_kernel(FbankParams P, const int16_t* __restrict__ pcm, const long long* __restrict__ pcm_off,
                const int* __restrict__ pad_before, const int* __restrict__ n_frames, const int* __restrict__ n_new,
                const int* __restrict__ carry, float* __restrict__ FB, int row_pitch) {
    extern __shared__ double smem_d[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int r = blockIdx.x * 4 + warp;
    if (r >= n_new[b]) return;
    float* out = FB + ((size_t)b * row_pitch + carry[b] + r) * P.mel;
    const int pb = pad_before[b];
    const int src = (r >= pb && r < pb + n_frames[b]) ? r - pb : -1;
    if (src < 0) {                                   // fbank_flush pad rows, fbank.c:312-316
        for (int m = lane; m < P.mel; m += 32) out[m] = P.pad_value;
        return;
    }
    double* re = smem_d + (size_t)warp * 2 * FB3_PITCH;
    double* im = re + FB3_PITCH;
    const int16_t* x = pcm + pcm_off[b] + (long long)src * P.shift;

    // samples j = 16 lane + u: pcm_off is a multiple of 8 samples, shift * 2 bytes a multiple of 16 whenever shift is a multiple of 8
    short xs[16];
    if ((P.shift & 7) == 0) {
        const uint4 lo = __ldg(reinterpret_cast<const uint4*>(x + 16 * lane)), hi = __ldg(reinterpret_cast<const uint4*>(x + 16 * lane + 8));
        const uint32_t wds[8] = { lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w };
#pragma unroll
        for (int u = 0; u < 8; u++) { xs[2 * u] = (short)(wds[u] & 0xFFFFu); xs[2 * u + 1] = (short)(wds[u] >> 16); }
    } else {
#pragma unroll
        for (int u = 0; u < 16; u++) xs[u] = x[16 * lane + u];
    }
    int isum = 0;
#pragma unroll
    for (int u = 0; u < 16; u++) isum += (int)xs[u];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) isum += __shfl_xor_sync(0xffffffffu, isum, o);
    const float mean = ((float)isum / 32768.0f) / (float)FB3_N;            // == the reference's float chain (exact, see above), then its float division
    const double dmean = (double)mean;
    const double coeff = (double)0.97f;                                     // preemph_coeff is a float, fbank.c:188
    // data[j] = ((x[j] - mean) - 0.97 (x[j-1] - mean)) * window[j]; j = 0 uses itself as predecessor (fbank.c:249-257)
    const short x_prev_lane = (short)__shfl_up_sync(0xffffffffu, (int)xs[15], 1);
#pragma unroll
    for (int u = 0; u < 16; u++) {
        const int j = 16 * lane + u;
        const short xp = u > 0 ? xs[u - 1] : (lane > 0 ? x_prev_lane : xs[0]);
        const double cur = __dsub_rn((double)((float)xs[u] / 32768.0f), dmean);
        const double prev = __dsub_rn((double)((float)xp / 32768.0f), dmean);
        double v = __dsub_rn(cur, __dmul_rn(coeff, prev));
        v = __dmul_rn(v, (double)__ldg(P.window + j));
        re[fb3_phys(j)] = v;
    }
    __syncwarp();
    // ---- pass 1: span 64, real input. butterfly j: x_k = data[j + 64 k]; y_m *= W512^(j m); -> position j + 64 m
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
        const int j = lane + 32 * it;
        cplx v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = { re[fb3_phys(j + 64 * k)], 0.0 };
        dft8(v);
#pragma unroll
        for (int m = 1; m < 8; m++) v[m] = cmul(v[m], fb3_tw(P.twiddle, j * m));
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 8; m++) { re[fb3_phys(j + 64 * m)] = v[m].x; im[fb3_phys(j + 64 * m)] = v[m].y; }
    }
    __syncwarp();
    // ---- pass 2: span 8 inside blocks of 64. butterfly (blk, j): x_k = a[64 blk + j + 8 k]; y_m *= W64^(j m) = W512^(8 j m)
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
        const int id = lane + 32 * it, blk = id >> 3, j = id & 7;
        cplx v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { const int p = fb3_phys(64 * blk + j + 8 * k); v[k] = { re[p], im[p] }; }
        dft8(v);
#pragma unroll
        for (int m = 1; m < 8; m++) v[m] = cmul(v[m], fb3_tw(P.twiddle, 8 * j * m));
#pragma unroll
        for (int m = 0; m < 8; m++) { const int p = fb3_phys(64 * blk + j + 8 * m); re[p] = v[m].x; im[p] = v[m].y; }   // a butterfly only rewrites what it read
    }
    __syncwarp();
    // ---- pass 3: span 1. block q = 8 m1 + m2 holds, after its DFT8, X[m1 + 8 m2 + 64 m3] in element m3: bins < 256 are m3 < 4.
    // power in FLOAT, im[0] forced to 0 (fbank.c:269-280); written over the (no longer needed) start of `re` region of ANOTHER array: pw
    float* pw = reinterpret_cast<float*>(im);                               // im rows are consumed below before pw overwrites them: see the sync
    float pwr[2][4];
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const int q = lane + 32 * it;
        cplx v[8];
        const double2* rp = reinterpret_cast<const double2*>(re + fb3_phys(8 * q));
        const double2* ip = reinterpret_cast<const double2*>(im + fb3_phys(8 * q));
#pragma unroll
        for (int k = 0; k < 4; k++) { const double2 a = rp[k], c = ip[k]; v[2 * k] = { a.x, c.x }; v[2 * k + 1] = { a.y, c.y }; }
        dft8(v);
#pragma unroll
        for (int m3 = 0; m3 < 4; m3++) {
            const int kbin = (q >> 3) + 8 * (q & 7) + 64 * m3;
            const float fr = (float)v[m3].x, fi = kbin == 0 ? 0.f : (float)v[m3].y;
            pwr[it][m3] = __fadd_rn(__fmul_rn(fr, fr), __fmul_rn(fi, fi));
        }
    }
    __syncwarp();                                                           // every lane has read its rows of re / im
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const int q = lane + 32 * it;
#pragma unroll
        for (int m3 = 0; m3 < 4; m3++) pw[(q >> 3) + 8 * (q & 7) + 64 * m3] = pwr[it][m3];
    }
    __syncwarp();
    // mel: sequential float sum over fft bins (zero weights contribute exact +0), fbank.c:283-296
    for (int m = lane; m < P.mel; m += 32) {
        const int k0 = P.mel_start[m], cnt = P.mel_count[m];
        const float* w = P.mel_w + P.mel_off[m];
        float val = 0.f;
        for (int i = 0; i < cnt; i++) val = __fadd_rn(val, __fmul_rn(pw[k0 + i], __ldg(w + i)));
        const float kEps = 1.1920928955078125e-07f;
        out[m] = (float)log((double)fmaxf(kEps, val));
    }
}

// copy carried-over feature rows (<= carry_cap per session) into the pass buffer and back
__global__ void carry_in_kernel(const float* __restrict__ carry_store, const int* __restrict__ slot, const int* __restrict__ carry,
                                float* __restrict__ FB, int row_pitch, int mel, int carry_cap) {
    int b = blockIdx.x;
    const float* src = carry_store + (size_t)slot[b] * carry_cap * mel;
    float* dst = FB + (size_t)b * row_pitch * mel;
    for (int i = threadIdx.x; i < carry[b] * mel; i += blockDim.x) dst[i] = src[i];
}
__global__ void carry_out_kernel(float* __restrict__ carry_store, const int* __restrict__ slot, const int* __restrict__ first_row,
                                 const int* __restrict__ n_keep, const float* __restrict__ FB, int row_pitch, int mel, int carry_cap) {
    int b = blockIdx.x;
    float* dst = carry_store + (size_t)slot[b] * carry_cap * mel;
    const float* src = FB + ((size_t)b * row_pitch + first_row[b]) * mel;
    for (int i = threadIdx.x; i < n_keep[b] * mel; i += blockDim.x) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float double_swish(float x) { return x * sigmoidf_(x - 1.0f); }   // icefall scaling.py DoubleSwish

// ------------------------------------------------------------------------------------------------
// E1: Conv2dSubsampling up to the flattened (c * F3 + f) vector. One CTA per (t, b) window.
// Output row n = t * B + b of `out` [M, C3 * F3].
// ------------------------------------------------------------------------------------------------
struct ConvParams {
    int SEG, MEL, STEP, C1, C2, C3;
    int T1, F1, T2, F2, F3;
    const float *w1, *b1, *w2, *b2, *w3, *b3;
};

__global__ void __launch_bounds__(256)
conv_embed_kernel(ConvParams P, const float* __restrict__ FB, int row_pitch, const int* __restrict__ nsteps, int t0, int B,
                  float* __restrict__ out) {
    extern __shared__ float smem_f[];
    const int b = blockIdx.x, t = blockIdx.y;            // t relative to the chunk
    float* dst = out + ((size_t)t * B + b) * (P.C3 * P.F3);
    if (t0 + t >= nsteps[b]) {                           // ragged tail: keep the buffer defined
        for (int i = threadIdx.x; i < P.C3 * P.F3; i += blockDim.x) dst[i] = 0.f;
        return;
    }
    float* in = smem_f;                                   // [SEG][MEL]
    float* c1 = in + P.SEG * P.MEL;                       // [C1][T1][F1]
    float* c2 = c1 + P.C1 * P.T1 * P.F1;                  // [C2][T2][F2]
    float* w2s = c2 + P.C2 * P.T2 * P.F2;                 // [C2][C1][9]
    const float* src = FB + ((size_t)b * row_pitch + (size_t)(t0 + t) * P.STEP) * P.MEL;
    for (int i = threadIdx.x; i < P.SEG * P.MEL; i += blockDim.x) in[i] = src[i];
    for (int i = threadIdx.x; i < P.C2 * P.C1 * 9; i += blockDim.x) w2s[i] = P.w2[i];
    __syncthreads();
    const int n1 = P.C1 * P.T1 * P.F1;
    for (int i = threadIdx.x; i < n1; i += blockDim.x) {
        int f = i % P.F1, y = (i / P.F1) % P.T1, c = i / (P.F1 * P.T1);
        const float* w = P.w1 + c * 9;
        float s = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ky++)
#pragma unroll
            for (int kx = 0; kx < 3; kx++) s = fmaf(in[(y + ky) * P.MEL + f + kx], w[ky * 3 + kx], s);
        c1[i] = double_swish(s + P.b1[c]);
    }
    __syncthreads();
    const int n2 = P.C2 * P.T2 * P.F2;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        int f = i % P.F2, y = (i / P.F2) % P.T2, c = i / (P.F2 * P.T2);
        const float* w = w2s + c * P.C1 * 9;
        float s = 0.f;
        for (int ic = 0; ic < P.C1; ic++) {
            const float* ip = c1 + (ic * P.T1 + 2 * y) * P.F1 + 2 * f;
#pragma unroll
            for (int ky = 0; ky < 3; ky++)
#pragma unroll
                for (int kx = 0; kx < 3; kx++) s = fmaf(ip[ky * P.F1 + kx], w[ic * 9 + ky * 3 + kx], s);
        }
        c2[i] = double_swish(s + P.b2[c]);
    }
    __syncthreads();
    // conv3 (T3 == 1): each thread owns output channel oc and a strip of frequencies
    const int n3 = P.C3 * P.F3;
    for (int i = threadIdx.x; i < n3; i += blockDim.x) {
        int f = i % P.F3, c = i / P.F3;
        const float* w = P.w3 + (size_t)c * P.C2 * 9;
        float s = 0.f;
        for (int ic = 0; ic < P.C2; ic++) {
            const float* ip = c2 + (ic * P.T2) * P.F2 + 2 * f;
#pragma unroll
            for (int ky = 0; ky < 3; ky++)
#pragma unroll
                for (int kx = 0; kx < 3; kx++) s = fmaf(ip[ky * P.F2 + kx], __ldg(w + ic * 9 + ky * 3 + kx), s);
        }
        dst[i] = double_swish(s + P.b3[c]);              // transpose(1,2).view(b, t, c*f): index c * F3 + f
    }
}

// ------------------------------------------------------------------------------------------------
// fp32 SIMT GEMM: C[M,N] = A[M,K] * W[N,K]^T (+bias) with fused epilogue. K % 16 == 0.
// ------------------------------------------------------------------------------------------------
enum Epilogue { EPI_BIAS = 0, EPI_BIAS_DSWISH = 1, EPI_BIAS_RESIDUAL = 2 };

template <int EPI>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                 const float* __restrict__ R, float* __restrict__ C, int M, int N, int K) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int lr = tid >> 2, lc = (tid & 3) * 4;          // loader: row 0..63, k offset 0,4,8,12
    const int ty = tid >> 4, tx = tid & 15;               // compute: 4x4 micro-tile
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
        float4 av = make_float4(0, 0, 0, 0), wv = make_float4(0, 0, 0, 0);
        if (m0 + lr < M) av = *reinterpret_cast<const float4*>(A + (size_t)(m0 + lr) * K + k0 + lc);
        if (n0 + lr < N) wv = *reinterpret_cast<const float4*>(W + (size_t)(n0 + lr) * K + k0 + lc);
        As[lc + 0][lr] = av.x; As[lc + 1][lr] = av.y; As[lc + 2][lr] = av.z; As[lc + 3][lr] = av.w;
        Ws[lc + 0][lr] = wv.x; Ws[lc + 1][lr] = wv.y; Ws[lc + 2][lr] = wv.z; Ws[lc + 3][lr] = wv.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { a[i] = As[kk][ty * 4 + i]; w[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            if (EPI == EPI_BIAS_DSWISH) v = double_swish(v);
            if (EPI == EPI_BIAS_RESIDUAL) v += R[(size_t)m * N + n];
            C[(size_t)m * N + n] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// E2 recurrent step, exact tier. gates = G[t] + h W_hh^T ; c' = s(f) c + s(i) tanh(g) ; m = s(o) tanh(c')
// CTA: 16 hidden units x 64 batch rows. grid (H/16, ceil(B/64)).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lstm_gates_kernel(const float* __restrict__ G /*[B,4H] rows of step t*/, const float* __restrict__ Hst /*[B,D]*/,
                  const float* __restrict__ Whh /*[4H,D]*/, float* __restrict__ Cst /*[B,H]*/, float* __restrict__ Mout /*[B,H]*/,
                  const int* __restrict__ nsteps, int t, int B, int D, int H) {
    constexpr int BU = 16, BB = 64, BK = 16;
    __shared__ float Hs[BK][BB + 4];
    __shared__ float Ws[BK][4 * BU + 4];                  // [k][gate * 16 + unit]
    const int tid = threadIdx.x;
    const int u0 = blockIdx.x * BU, b0 = blockIdx.y * BB;
    const int lr = tid >> 2, lc = (tid & 3) * 4;
    const int tb = tid >> 4, tu = tid & 15;               // thread: batch rows tb*4..+3, unit tu, all 4 gates
    float acc[4][4] = {};                                  // [batch i][gate]
    for (int k0 = 0; k0 < D; k0 += BK) {
        float4 hv = make_float4(0, 0, 0, 0);
        if (b0 + lr < B) hv = *reinterpret_cast<const float4*>(Hst + (size_t)(b0 + lr) * D + k0 + lc);
        int gate = lr >> 4, unit = lr & 15;
        float4 wv = *reinterpret_cast<const float4*>(Whh + ((size_t)gate * H + u0 + unit) * D + k0 + lc);
        Hs[lc + 0][lr] = hv.x; Hs[lc + 1][lr] = hv.y; Hs[lc + 2][lr] = hv.z; Hs[lc + 3][lr] = hv.w;
        Ws[lc + 0][lr] = wv.x; Ws[lc + 1][lr] = wv.y; Ws[lc + 2][lr] = wv.z; Ws[lc + 3][lr] = wv.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { a[i] = Hs[kk][tb * 4 + i]; w[i] = Ws[kk][i * 16 + tu]; }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int g = 0; g < 4; g++) acc[i][g] = fmaf(a[i], w[g], acc[i][g]);
        }
        __syncthreads();
    }
    const int u = u0 + tu;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int b = b0 + tb * 4 + i;
        if (b >= B || t >= nsteps[b]) continue;           // ragged: finished sessions keep their state
        const float* g = G + (size_t)b * 4 * H;
        float gi = acc[i][0] + g[u], gf = acc[i][1] + g[H + u], gg = acc[i][2] + g[2 * H + u], go = acc[i][3] + g[3 * H + u];
        float c = Cst[(size_t)b * H + u];
        float cn = sigmoidf_(gf) * c + sigmoidf_(gi) * tanhf(gg);
        Cst[(size_t)b * H + u] = cn;
        Mout[(size_t)b * H + u] = sigmoidf_(go) * tanhf(cn);
    }
}

// h' = m W_hr^T ; writes the recurrent state and the layer output row. grid (D/64, ceil(B/64)).
__global__ void __launch_bounds__(256)
lstm_proj_kernel(const float* __restrict__ Min /*[B,H]*/, const float* __restrict__ Whr /*[D,H]*/, float* __restrict__ Hst /*[B,D]*/,
                 float* __restrict__ Hout /*[B,D] rows of step t*/, const int* __restrict__ nsteps, int t, int B, int D, int H) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int lr = tid >> 2, lc = (tid & 3) * 4;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < H; k0 += BK) {
        float4 av = make_float4(0, 0, 0, 0), wv = make_float4(0, 0, 0, 0);
        if (m0 + lr < B) av = *reinterpret_cast<const float4*>(Min + (size_t)(m0 + lr) * H + k0 + lc);
        if (n0 + lr < D) wv = *reinterpret_cast<const float4*>(Whr + (size_t)(n0 + lr) * H + k0 + lc);
        As[lc + 0][lr] = av.x; As[lc + 1][lr] = av.y; As[lc + 2][lr] = av.z; As[lc + 3][lr] = av.w;
        Ws[lc + 0][lr] = wv.x; Ws[lc + 1][lr] = wv.y; Ws[lc + 2][lr] = wv.z; Ws[lc + 3][lr] = wv.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { a[i] = As[kk][ty * 4 + i]; w[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int b = m0 + ty * 4 + i;
        if (b >= B) continue;
        bool live = t < nsteps[b];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int n = n0 + tx * 4 + j;
            if (n >= D) continue;
            if (live) { Hst[(size_t)b * D + n] = acc[i][j]; Hout[(size_t)b * D + n] = acc[i][j]; }
            else Hout[(size_t)b * D + n] = 0.f;
        }
    }
}

// y = a + b (row-major, n4 float4 elements)
__global__ void add_rows_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ y, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) { float4 x = a[i], z = b[i]; y[i] = make_float4(x.x + z.x, x.y + z.y, x.z + z.z, x.w + z.w); }
}

// BasicNorm: x * (mean(x^2) + eps)^-0.5, one warp per row (icefall scaling.py BasicNorm [external])
__global__ void basic_norm_kernel(const float* __restrict__ X, float* __restrict__ Y, int M, int D, float eps) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* x = X + (size_t)row * D;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) s = fmaf(x[i], x[i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float scale = 1.0f / sqrtf(s / (float)D + eps);
    float* y = Y + (size_t)row * D;
    for (int i = lane; i < D; i += 32) y[i] = x[i] * scale;
}

// gather / scatter of recurrent state between per-slot storage [slot][L][dim] and pass layout [L][B][dim]
__global__ void state_gather_kernel(const float* __restrict__ store, const int* __restrict__ slot, float* __restrict__ work, int L, int B, int dim) {
    int b = blockIdx.x, l = blockIdx.y;
    const float* s = store + ((size_t)slot[b] * L + l) * dim;
    float* d = work + ((size_t)l * B + b) * dim;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) d[i] = s[i];
}
__global__ void state_scatter_kernel(float* __restrict__ store, const int* __restrict__ slot, const float* __restrict__ work, int L, int B, int dim) {
    int b = blockIdx.x, l = blockIdx.y;
    float* s = store + ((size_t)slot[b] * L + l) * dim;
    const float* d = work + ((size_t)l * B + b) * dim;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) s[i] = d[i];
}

// ------------------------------------------------------------------------------------------------
// D1 + J1 + G1
// ------------------------------------------------------------------------------------------------
struct JoinerParams {
    int J, E, V, blank;
    const float *emb, *dec_conv /*[E][2]*/, *dec_w /*[J][E]*/, *dec_b, *join_w /*[V][J]*/, *join_b;
    const uint8_t* vflags;
};

// block-wide: dout[J] = W_p relu(emb[c0] k0 + emb[c1] k1) + b_p. scratch: E floats. blockDim multiple of 32.
__device__ void decoder_eval(const JoinerParams& P, int c0, int c1, float* scratch, float* dout) {
    for (int e = threadIdx.x; e < P.E; e += blockDim.x) {
        float s = 0.f;
        s = fmaf(P.emb[(size_t)c0 * P.E + e], P.dec_conv[e * 2 + 0], s);
        s = fmaf(P.emb[(size_t)c1 * P.E + e], P.dec_conv[e * 2 + 1], s);
        scratch[e] = fmaxf(s, 0.f);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = warp; j < P.J; j += nw) {
        const float* w = P.dec_w + (size_t)j * P.E;
        float s = 0.f;
        for (int e = lane; e < P.E; e += 32) s = fmaf(scratch[e], w[e], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) dout[j] = s + P.dec_b[j];
    }
    __syncthreads();
}

__global__ void decoder_kernel(JoinerParams P, int c0, int c1, float* __restrict__ dout) {
    extern __shared__ float smem_f[];
    decoder_eval(P, c0, c1, smem_f, smem_f + P.E);
    for (int j = threadIdx.x; j < P.J; j += blockDim.x) dout[j] = smem_f[P.E + j];
}

// The part of aas_process_logits (april_session.c:306-429) that feeds back into the model: decides
// is_blank and evolves {context, active-token count, last emission time}. Callback shaping is
// replayed on the host from the DecisionRecords (session.cpp), which re-derives is_blank
// independently and cross-checks it.
__device__ bool greedy_decide(GreedyState& st, const uint8_t* vf, int blank, float early_emit, int max_idx, float max_val,
                              float blank_val, bool& ctx_changed) {
    ctx_changed = false;
    const bool was_context_cleared = st.ctx1 == blank;                       // :322
    const bool is_equal_to_previous = st.ctx1 == max_idx;                    // :326
    if (is_equal_to_previous) early_emit = 0.f;
    bool is_blank = (blank_val - early_emit) > max_val;                      // :330
    const uint8_t f = max_idx >= 0 ? vf[max_idx] : 0;
    const bool wb = f & VF_SPACE;
    bool is_eos = f & VF_EOS;
    bool is_punct = is_eos || (f & VF_COMMA);
    if (is_punct && st.head > 0) {                                           // :345-351
        uint8_t lf = vf[st.ids[st.head - 1]];
        if ((lf & VF_DIGIT) && (f & VF_DOT)) { is_eos = false; is_punct = false; }   // "10.5" is not a sentence end
    }
    if (!was_context_cleared && is_punct && !is_equal_to_previous && max_val > (blank_val - 3.5f)) is_blank = false;   // :356-358
    if (max_idx < 0) is_blank = true;                                        // all non-blank logits non-finite: emit nothing (mirrored in session.cpp)
    if (!is_blank) {
        st.last_emission_ms = st.current_ms;                                 // :362
        st.ctx0 = st.ctx1; st.ctx1 = max_idx; ctx_changed = true;            // :364 aas_update_context
        bool is_final = st.head >= (APRIL_MAX_ACTIVE_TOKENS - 1);            // :366
        if (st.head > 0 && wb) {                                             // :369-388
            uint8_t lf = vf[st.ids[st.head - 1]];
            if (lf & VF_EOS) is_final = true;
        }
        if (is_final && st.head > 0) {                                       // aas_finalize_previous_words :213-255
            if (wb) st.head = 0;
            else {
                int start = APRIL_MAX_ACTIVE_TOKENS;
                for (int i = st.head - 1; i > 2; i--) if (vf[st.ids[i]] & VF_SPACE) { start = i; break; }
                if (start == APRIL_MAX_ACTIVE_TOKENS) st.head = 0;
                else {
                    for (int i = 0; i < st.head - start; i++) st.ids[i] = st.ids[start + i];
                    st.head -= start;
                }
            }
        }
        if (st.head >= (APRIL_MAX_ACTIVE_TOKENS - 1)) st.head = 0;           // :392-396
        st.ids[st.head++] = (int16_t)max_idx;                                // :398 aas_emit_token(force)
    } else {
        long long since = st.current_ms - st.last_emission_ms;              // :402
        if (since >= 2200) {                                                 // :411-416
            st.head = 0;                                                     // aas_finalize_tokens
            if (st.ctx0 != blank) { st.ctx0 = blank; st.ctx1 = blank; ctx_changed = true; }   // aas_clear_context :296-301
        }
        // tentative partials (:417-422) push then pop: no net state change
    }
    return is_blank;
}

// One CTA per session; walks the session's encoder steps of this pass sequentially (S1, april_session.c:431-476).
__global__ void __launch_bounds__(256)
greedy_kernel(JoinerParams P, const float* __restrict__ enc_out /*[T*B, J] time-major*/, const int* __restrict__ nsteps,
              const int* __restrict__ slot, const int* __restrict__ flush_flag, int B, int Tcap, int t_base, int stride_ms,
              GreedyState* __restrict__ gstate, float* __restrict__ dout_store /*[slot][J]*/, const float* __restrict__ dout_blank,
              DecisionRecord* __restrict__ rec /*[B][Tcap][3]*/, int* __restrict__ npass /*[B][Tcap]*/,
              float* __restrict__ tap_logits /*[B][Tcap][3][V] or null*/, int last_chunk, int chunk_T) {
    extern __shared__ float smem_f[];
    float* e_s = smem_f;                 // [J]
    float* d_s = e_s + P.J;              // [J]
    float* z_s = d_s + P.J;              // [J]
    float* scratch = z_s + P.J;          // [E]
    __shared__ float best_val[8];
    __shared__ int best_idx[8];
    __shared__ GreedyState st;
    __shared__ float blank_logit;
    __shared__ int s_flag;               // bit0: is_blank, bit1: ctx changed
    const int b = blockIdx.x;
    const int sl = slot[b];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) st = gstate[sl];
    for (int j = threadIdx.x; j < P.J; j += blockDim.x) d_s[j] = dout_store[(size_t)sl * P.J + j];
    __syncthreads();
    const int n = nsteps[b];
    const int t_end = min(n, t_base + chunk_T);
    for (int t = t_base; t < t_end; t++) {
        if (threadIdx.x == 0) st.current_ms += stride_ms;                    // :442-443
        const float* e = enc_out + ((size_t)(t - t_base) * B + b) * P.J;
        for (int j = threadIdx.x; j < P.J; j += blockDim.x) e_s[j] = e[j];
        __syncthreads();
        int pass = 0;
        for (; pass < 3; pass++) {                                           // :449-454
            const float early = pass == 0 ? 1.0f : 0.0f;
            for (int j = threadIdx.x; j < P.J; j += blockDim.x) z_s[j] = tanhf(e_s[j] + d_s[j]);
            __syncthreads();
            float bv = (float)(-9999999999.0); int bi = -1;                  // :311-312
            float* tl = tap_logits ? tap_logits + (((size_t)b * Tcap + t) * 3 + pass) * P.V : nullptr;
            for (int v = warp; v < P.V; v += nw) {
                const float* w = P.join_w + (size_t)v * P.J;
                float s = 0.f;
                for (int j = lane; j < P.J; j += 32) s = fmaf(z_s[j], w[j], s);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                s += P.join_b[v];
                if (lane == 0) {
                    if (tl) tl[v] = s;
                    if (v == P.blank) blank_logit = s;
                    else if (s > bv) { bv = s; bi = v; }                     // ascending v inside a warp: first max wins
                }
            }
            if (lane == 0) { best_val[warp] = bv; best_idx[warp] = bi; }
            __syncthreads();
            if (threadIdx.x == 0) {
                float mv = (float)(-9999999999.0); int mi = -1;
                for (int w = 0; w < nw; w++) {                               // strict '>' over ascending index == lowest index on ties
                    if (best_idx[w] < 0) continue;
                    if (best_val[w] > mv || (best_val[w] == mv && mi >= 0 && best_idx[w] < mi)) { mv = best_val[w]; mi = best_idx[w]; }
                }
                bool ctx_changed;
                bool is_blank = greedy_decide(st, P.vflags, P.blank, early, mi, mv, blank_logit, ctx_changed);
                DecisionRecord r; r.max_idx = mi; r.is_blank = is_blank ? 1 : 0; r.max_val = mv; r.blank_val = blank_logit;
                rec[((size_t)b * Tcap + t) * 3 + pass] = r;
                s_flag = (is_blank ? 1 : 0) | (ctx_changed ? 2 : 0);
            }
            __syncthreads();
            const int fl = s_flag;
            if (fl & 2) {                                                    // aas_run_decoder on the new context
                if (st.ctx0 == P.blank && st.ctx1 == P.blank) { for (int j = threadIdx.x; j < P.J; j += blockDim.x) d_s[j] = dout_blank[j]; __syncthreads(); }
                else decoder_eval(P, st.ctx0, st.ctx1, scratch, d_s);
            }
            if (fl & 1) { pass++; break; }
        }
        if (threadIdx.x == 0) npass[(size_t)b * Tcap + t] = pass;
        __syncthreads();
    }
    if (last_chunk && flush_flag[b]) {                                       // _aas_flush tail :561-563
        if (threadIdx.x == 0) {
            st.head = 0;
            if (st.ctx0 != P.blank) { st.ctx0 = P.blank; st.ctx1 = P.blank; s_flag = 2; } else s_flag = 0;
        }
        __syncthreads();
        if (s_flag & 2) for (int j = threadIdx.x; j < P.J; j += blockDim.x) d_s[j] = dout_blank[j];
        __syncthreads();
    }
    if (threadIdx.x == 0) gstate[sl] = st;
    for (int j = threadIdx.x; j < P.J; j += blockDim.x) dout_store[(size_t)sl * P.J + j] = d_s[j];
}

// ------------------------------------------------------------------------------------------------
// J1 + G1, cluster version: a thread-block cluster of 8 CTAs keeps the joiner output matrix
// resident in (distributed) shared memory — CTA r owns vocabulary rows [r*R, r*R+R) in fp32 — and
// walks S sessions step-synchronously. Per joiner pass every CTA forms z = tanh(e + d) for its
// sessions, multiplies its rows, publishes (best non-blank value, index, blank logit) to all 8
// CTAs over DSMEM, and after one cluster barrier every CTA evaluates the same greedy rule on the
// same 8 partials (replicated, deterministic state), so no second exchange is needed.
// The stateless decoder (D1) is a table lookup: dtab[(c0 * V + c1) * J + j], built once per model
// from the same arithmetic as decoder_eval (the decoder output is a pure function of the two
// context tokens: SURVEY.md §8a D1).
// ------------------------------------------------------------------------------------------------
constexpr int GC_CLUSTER = 8;
constexpr int GC_MAX_S = 8;          // sessions per cluster
constexpr int GC_MAX_JPL = 16;       // J / 32

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void dsmem_store_f32(const void* local_smem_ptr, uint32_t rank, float v) {
    uint32_t laddr = (uint32_t)__cvta_generic_to_shared(local_smem_ptr), raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}

struct GreedyPartial { float best_val; float best_idx; float blank_val; float has_blank; };

constexpr int GC_RPW = 8;            // vocabulary rows per warp (8 warps -> up to 64 rows per CTA, 512 per cluster)
constexpr int GC_SG = 4;             // sessions processed per register round

// sum over the 32 lanes of 32 per-lane values; lane l ends up with the total of value index l
__device__ __forceinline__ float warp_reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (i < n / 2) {
                const float send = up ? v[i] : v[i + n / 2];
                const float keep = up ? v[i + n / 2] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
    }
    return v[0];
}

__global__ void __cluster_dims__(GC_CLUSTER, 1, 1) __launch_bounds__(256, 1)
greedy_cluster_kernel(JoinerParams P, const float* __restrict__ dtab, const float* __restrict__ enc_out /*[T*B, J] chunk, time-major*/,
                      const int* __restrict__ nsteps, const int* __restrict__ slot, const int* __restrict__ flush_flag, int B, int Tcap, int t_base,
                      int stride_ms, GreedyState* __restrict__ gstate, DecisionRecord* __restrict__ rec, int* __restrict__ npass,
                      float* __restrict__ tap_logits, int last_chunk, int chunk_T, int S, int R) {
    extern __shared__ float smem_f[];
    const int J = P.J, V = P.V;
    const uint32_t rank = cluster_rank();
    const int cid = blockIdx.x / GC_CLUSTER;
    const int s0 = cid * S;
    const int ns = min(S, B - s0);                         // sessions of this cluster (cluster-uniform)
    const int Rp = (R + 3) & ~3;
    float* es = smem_f;                                    // [2][S][J] encoder rows: current step / prefetched next step
    float* ds = es + (size_t)2 * S * J;                    // [S][J] decoder rows of the current contexts
    float* zs = ds + (size_t)S * J;                        // [S][J] tanh(e + d)
    float* logit_s = zs + (size_t)S * J;                   // [S][Rp]
    GreedyPartial* inbox = reinterpret_cast<GreedyPartial*>(logit_s + (size_t)S * Rp);   // [2][8][S]
    uint8_t* vf_s = reinterpret_cast<uint8_t*>(inbox + 2 * GC_CLUSTER * S);               // [V]
    __shared__ GreedyState st[GC_MAX_S];
    __shared__ int s_done[GC_MAX_S], s_np[GC_MAX_S], s_newctx[GC_MAX_S], s_n[GC_MAX_S];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = (int)rank * R;
    const int jpl = J / 32;

    // this warp's rows of the joiner matrix live in registers for the whole launch: w[r][i] = W[row0 + warp*8 + r][lane + 32 i]
    float w[GC_RPW][GC_MAX_JPL];
    float bias_r = 0.f;                                    // bias of the row this lane reports after the reduce-scatter (row lane/4)
#pragma unroll
    for (int r = 0; r < GC_RPW; r++) {
        const int lr = warp * GC_RPW + r, v = row0 + lr;
#pragma unroll
        for (int i = 0; i < GC_MAX_JPL; i++) w[r][i] = (lr < R && v < V && i < jpl) ? P.join_w[(size_t)v * J + lane + 32 * i] : 0.f;
    }
    {
        const int lr = warp * GC_RPW + lane / GC_SG, v = row0 + lr;
        bias_r = (lr < R && v < V) ? P.join_b[v] : 0.f;
    }
    for (int i = threadIdx.x; i < V; i += blockDim.x) vf_s[i] = P.vflags[i];
    if (threadIdx.x < ns) { st[threadIdx.x] = gstate[slot[s0 + threadIdx.x]]; s_newctx[threadIdx.x] = 1; s_n[threadIdx.x] = nsteps[s0 + threadIdx.x]; }
    __syncthreads();
    if (ns <= 0) { cluster_sync_all(); return; }            // (never: the host launches only populated clusters)

    int t_end = 0;
    for (int s = 0; s < ns; s++) t_end = max(t_end, min(s_n[s], t_base + chunk_T));
    // encoder rows of the first step
    for (int s = 0; s < ns; s++)
        if (t_base < s_n[s]) for (int j = threadIdx.x; j < J; j += blockDim.x) es[(size_t)s * J + j] = enc_out[((size_t)0 * B + s0 + s) * J + j];
    int parity = 0;
    for (int t = t_base; t < t_end; t++) {
        const float* e_cur = es + (size_t)((t - t_base) & 1) * S * J;
        float* e_nxt = es + (size_t)((t - t_base + 1) & 1) * S * J;
        if (threadIdx.x < ns) {
            const int s = threadIdx.x;
            const bool act = t < s_n[s];
            s_done[s] = act ? 0 : 1; s_np[s] = 0;
            if (act) st[s].current_ms += stride_ms;                           // april_session.c:442-443
        }
        __syncthreads();
        for (int pass = 0; pass < 3; pass++) {
            // --- A: decoder rows for changed contexts (table lookup = aas_run_decoder), z = tanh(e + d); prefetch next step's e
            float e_pf[GC_MAX_S][2];                                          // registers: fully unrolled, statically indexed
            const bool do_pf = pass == 0 && t + 1 < t_end;
            if (do_pf) {
#pragma unroll
                for (int s = 0; s < GC_MAX_S; s++)
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int j = threadIdx.x + u * 256;
                        e_pf[s][u] = (s < ns && j < J && t + 1 < s_n[s]) ? enc_out[((size_t)(t + 1 - t_base) * B + s0 + s) * J + j] : 0.f;
                    }
            }
            // all sessions' table rows are requested before any is used: one DRAM round trip per pass, not one per session
            float d_new[GC_MAX_S][2];
#pragma unroll
            for (int s = 0; s < GC_MAX_S; s++) {
                const bool nc = s < ns && !s_done[s] && s_newctx[s] != 0;
                const float* drow = nc ? dtab + ((size_t)st[s].ctx0 * V + st[s].ctx1) * J : dtab;
#pragma unroll
                for (int u = 0; u < 2; u++) { const int j = threadIdx.x + u * 256; d_new[s][u] = (nc && j < J) ? __ldg(drow + j) : 0.f; }
            }
#pragma unroll
            for (int s = 0; s < GC_MAX_S; s++) {
                if (s >= ns || s_done[s]) continue;
                const bool nc = s_newctx[s] != 0;
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int j = threadIdx.x + u * 256;
                    if (j < J) {
                        float dv;
                        if (nc) { dv = d_new[s][u]; ds[(size_t)s * J + j] = dv; } else dv = ds[(size_t)s * J + j];
                        zs[(size_t)s * J + j] = tanhf(e_cur[(size_t)s * J + j] + dv);
                    }
                }
            }
            __syncthreads();
            // only a session that took part in this pass has consumed its flag: one that finished the step on a blank which ALSO
            // reset its context (2.2 s of silence, april_session.c:411-416) keeps the flag for its next step while its cluster mates run on
            if (threadIdx.x < ns && !s_done[threadIdx.x]) s_newctx[threadIdx.x] = 0;
            // --- B: logits of my rows: registers x shared z, 4 sessions per round
            for (int sg = 0; sg < ns; sg += GC_SG) {
                float acc[32];
#pragma unroll
                for (int p = 0; p < 32; p++) acc[p] = 0.f;
#pragma unroll
                for (int i = 0; i < GC_MAX_JPL; i++) {
                    if (i < jpl) {
                        float z[GC_SG];
#pragma unroll
                        for (int q = 0; q < GC_SG; q++) z[q] = (sg + q < ns) ? zs[(sg + q) * J + lane + 32 * i] : 0.f;
#pragma unroll
                        for (int r = 0; r < GC_RPW; r++)
#pragma unroll
                            for (int q = 0; q < GC_SG; q++) acc[r * GC_SG + q] = fmaf(w[r][i], z[q], acc[r * GC_SG + q]);
                    }
                }
                const float tot = warp_reduce_scatter32(acc, lane);          // lane -> (row lane/4, session sg + lane%4)
                const int lr = warp * GC_RPW + lane / GC_SG, sq = sg + lane % GC_SG;
                if (lr < R && sq < ns) logit_s[sq * Rp + lr] = tot + bias_r;
            }
            if (do_pf) {
#pragma unroll
                for (int s = 0; s < GC_MAX_S; s++)
#pragma unroll
                    for (int u = 0; u < 2; u++) { const int j = threadIdx.x + u * 256; if (s < ns && j < J && t + 1 < s_n[s]) e_nxt[(size_t)s * J + j] = e_pf[s][u]; }
            }
            __syncthreads();
            // --- C: warp s = local arg-max of session s over my rows (ties -> lowest index), published to all 8 CTAs over DSMEM
            if (warp < ns && !s_done[warp]) {
                const int s = warp;
                float bv = (float)(-9999999999.0); int bi = -1; float blank_v = 0.f; int has_blank = 0;
                float* tl = tap_logits ? tap_logits + (((size_t)(s0 + s) * Tcap + t) * 3 + pass) * V : nullptr;
                const float* lg = logit_s + s * Rp;
                for (int r = lane; r < R; r += 32) {
                    const int v = row0 + r;
                    if (v >= V) break;
                    const float x = lg[r];
                    if (tl) tl[v] = x;
                    if (v == P.blank) { blank_v = x; has_blank = 1; }
                    else if (x > bv) { bv = x; bi = v; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    const float ob = __shfl_xor_sync(0xffffffffu, blank_v, o);
                    const int oh = __shfl_xor_sync(0xffffffffu, has_blank, o);
                    if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
                    if (oh) { blank_v = ob; has_blank = 1; }
                }
                if (lane < GC_CLUSTER) {
                    GreedyPartial* slot_local = &inbox[(parity * GC_CLUSTER + rank) * S + s];
                    dsmem_store_f32(&slot_local->best_val, lane, bv);
                    dsmem_store_f32(&slot_local->best_idx, lane, (float)bi);
                    dsmem_store_f32(&slot_local->blank_val, lane, blank_v);
                    dsmem_store_f32(&slot_local->has_blank, lane, has_blank ? 1.f : 0.f);
                }
            }
            cluster_sync_all();
            // --- D: every CTA evaluates the same rule on the same 8 partials (replicated state)
            if (threadIdx.x < ns && !s_done[threadIdx.x]) {
                const int s = threadIdx.x;
                float mv = (float)(-9999999999.0); int mi = -1; float blank_v = 0.f;
                for (int c = 0; c < GC_CLUSTER; c++) {                        // rank order == ascending vocabulary order
                    const GreedyPartial& p = inbox[(parity * GC_CLUSTER + c) * S + s];
                    if (p.has_blank != 0.f) blank_v = p.blank_val;
                    if ((int)p.best_idx >= 0 && p.best_val > mv) { mv = p.best_val; mi = (int)p.best_idx; }
                }
                bool ctx_changed;
                const bool is_blank = greedy_decide(st[s], vf_s, P.blank, pass == 0 ? 1.0f : 0.0f, mi, mv, blank_v, ctx_changed);
                if (rank == 0) {
                    DecisionRecord r; r.max_idx = mi; r.is_blank = is_blank ? 1 : 0; r.max_val = mv; r.blank_val = blank_v;
                    rec[((size_t)(s0 + s) * Tcap + t) * 3 + pass] = r;
                }
                s_np[s] = pass + 1;
                if (ctx_changed) s_newctx[s] = 1;
                if (is_blank) s_done[s] = 1;
            }
            parity ^= 1;
            __syncthreads();
            bool all_done = true;
            for (int s = 0; s < ns; s++) all_done &= (s_done[s] != 0);
            if (all_done) break;
        }
        if (rank == 0 && threadIdx.x < ns && t < s_n[threadIdx.x]) npass[(size_t)(s0 + threadIdx.x) * Tcap + t] = s_np[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x < ns) {
        const int s = threadIdx.x;
        if (last_chunk && flush_flag[s0 + s]) {                               // _aas_flush tail, april_session.c:561-563
            st[s].head = 0;
            if (st[s].ctx0 != P.blank) { st[s].ctx0 = P.blank; st[s].ctx1 = P.blank; }
        }
        if (rank == 0) gstate[slot[s0 + s]] = st[s];
    }
    cluster_sync_all();                                                       // no CTA may exit while peers can still write its inbox
}

// decoder table rows for contexts (c0_base + blockIdx.y, *): Y[(y * V + c1)][e] = relu(emb[c0][e] k[e][0] + emb[c1][e] k[e][1])
__global__ void decoder_table_y_kernel(JoinerParams P, int c0_base, float* __restrict__ Y) {
    const int c1 = blockIdx.x, c0 = c0_base + blockIdx.y;
    for (int e = threadIdx.x; e < P.E; e += blockDim.x) {
        float s = 0.f;
        s = fmaf(P.emb[(size_t)c0 * P.E + e], P.dec_conv[e * 2 + 0], s);
        s = fmaf(P.emb[(size_t)c1 * P.E + e], P.dec_conv[e * 2 + 1], s);
        Y[((size_t)blockIdx.y * P.V + c1) * P.E + e] = fmaxf(s, 0.f);
    }
}

}  // namespace k
}  // namespace april
