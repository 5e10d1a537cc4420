"""NumPy restatement of the april-asr hot path (test/fixture infrastructure, NOT product).

Front-end follows /root/reference/src/fbank.c exactly (SURVEY.md Appendix A recipe):
  frame i = samples [160 i, 160 i + 512)            fbank.c:195-236
  float-accumulated mean over double data           fbank.c:241-246
  double pre-emphasis with (double)0.97f            fbank.c:249-253
  float Povey window over N=512, n/N                fbank.c:49-55,256-257
  double real FFT                                   fbank.c:259-270 (pocketfft)
  float power re^2+im^2, bins 0..255, im[0]=0       fbank.c:275-280
  sequential float mel dot (80x256)                 fbank.c:283-291, banks :65-95
  (float)log((double)max(eps, x))                   fbank.c:294-296
Net arithmetic follows extra/export-april.py:183-223,321-331 + icefall
lstm_transducer_stateless2 [external].
"""
import math

import numpy as np

F32 = np.float32
K_EPS = np.float32(1.1920928955078125e-07)


# ----------------------------------------------------------------------------- front-end
def povey_window(n):
    out = np.empty(n, dtype=np.float32)
    for i in range(n):
        out[i] = np.float32(math.pow(0.5 - 0.5 * math.cos(float(i) / float(n) * 6.283185307), 0.85))
    return out


def mel_banks(num_bins=80, num_fft_bins=256, padded=512, sr=16000, low=20, high=0):
    if high == 0:
        high = sr // 2
    ms = lambda f: 1127.0 * math.log(1.0 + f / 700.0)
    bw = F32(sr) / F32(padded)
    mel_low, mel_high = F32(ms(float(low))), F32(ms(float(high)))
    delta = F32(F32(mel_high - mel_low) / F32(F32(num_bins) + F32(1.0)))
    mat = np.zeros((num_bins, num_fft_bins), dtype=np.float32)
    for i in range(num_bins):
        left = F32(mel_low + F32(F32(i) * delta))
        center = F32(left + delta)
        right = F32(center + delta)
        for j in range(num_fft_bins):
            freq = F32(bw * F32(j))
            mel = F32(ms(float(freq)))
            w = F32(0.0)
            if mel > left and mel < right:
                if mel <= center:
                    w = F32(F32(mel - left) / F32(center - left))
                else:
                    w = F32(F32(right - mel) / F32(right - center))
            mat[i, j] = w
    return mat


_WIN = None
_MEL = None


def fbank_rows(pcm16, n_pad_zero=0):
    """pcm16: int16 [n]. Returns float32 [frames, 80] exactly as fbank.c produces for one stream
    (features are a pure function of absolute sample position)."""
    global _WIN, _MEL
    if _WIN is None:
        _WIN, _MEL = povey_window(512), mel_banks()
    wave = pcm16.astype(np.float32) / np.float32(32768.0)
    if n_pad_zero:
        wave = np.concatenate([wave, np.zeros(n_pad_zero, np.float32)])
    n = len(wave)
    if n < 512:
        return np.zeros((0, 80), np.float32)
    nfr = (n - 512) // 160 + 1
    rows = np.empty((nfr, 80), dtype=np.float32)
    pre = float(np.float32(0.97))
    win = _WIN.astype(np.float64)
    for i in range(nfr):
        d = wave[160 * i:160 * i + 512].astype(np.float64)
        s = np.float32(0.0)
        for v in d:                                   # float accumulator over double data
            s = np.float32(float(s) + v)
        mean = np.float32(s / np.float32(512))
        d = d - float(mean)
        d0 = d[0] - pre * d[0]
        d[1:] = d[1:] - pre * d[:-1]
        d[0] = d0
        d = d * win
        spec = np.fft.rfft(d)
        re = spec.real[:256].astype(np.float32)
        im = spec.imag[:256].astype(np.float32)
        im[0] = 0.0
        power = (re * re + im * im).astype(np.float32)
        for m in range(80):                           # sequential float sum over fft bins
            acc = np.float32(0.0)
            nz = np.nonzero(_MEL[m])[0]
            for k in nz:
                acc = np.float32(acc + np.float32(power[k] * _MEL[m, k]))
            rows[i, m] = np.float32(math.log(float(max(K_EPS, acc))))
    return rows


def fbank_rows_fast(pcm16):
    """Vectorised float64 version (not bit-exact; calibration only)."""
    global _WIN, _MEL
    if _WIN is None:
        _WIN, _MEL = povey_window(512), mel_banks()
    wave = pcm16.astype(np.float64) / 32768.0
    nfr = (len(wave) - 512) // 160 + 1
    idx = np.arange(512)[None, :] + 160 * np.arange(nfr)[:, None]
    d = wave[idx]
    d = d - d.mean(1, keepdims=True)
    d2 = d.copy()
    d2[:, 1:] -= 0.97 * d[:, :-1]
    d2[:, 0] -= 0.97 * d[:, 0]
    spec = np.fft.rfft(d2 * _WIN[None].astype(np.float64), axis=1)[:, :256]
    power = spec.real ** 2 + spec.imag ** 2
    power[:, 0] = spec.real[:, 0] ** 2
    mel = power @ _MEL.T.astype(np.float64)
    return np.log(np.maximum(float(K_EPS), mel)).astype(np.float32)


def segments(rows, size=9, step=4):
    n = (rows.shape[0] - size) // step + 1 if rows.shape[0] >= size else 0
    return np.stack([rows[step * s:step * s + size] for s in range(n)]) if n > 0 else np.zeros((0, size, rows.shape[1]), np.float32)


# ----------------------------------------------------------------------------- nets (fp32)
def sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(np.float32)


def double_swish(x):
    return x * sigmoid(x - np.float32(1.0))


def conv2d(x, w, b, stride):
    """x [N,C,H,W], w [O,C,3,3] -> valid conv."""
    N, C, H, W = x.shape
    O = w.shape[0]
    Ho, Wo = (H - 3) // stride + 1, (W - 3) // stride + 1
    cols = np.empty((N, C, 3, 3, Ho, Wo), dtype=np.float32)
    for i in range(3):
        for j in range(3):
            cols[:, :, i, j] = x[:, :, i:i + stride * Ho:stride, j:j + stride * Wo:stride]
    out = np.einsum("ncijhw,ocij->nohw", cols, w, optimize=True).astype(np.float32)
    return out + b[None, :, None, None]


def basic_norm(x, eps):
    ms = np.mean(x * x, axis=-1, keepdims=True, dtype=np.float32)
    return x * ((ms + np.float32(eps)) ** np.float32(-0.5)).astype(np.float32)


def encoder_embed(W, seg):
    """seg [N,9,80] -> [N, D]"""
    x = seg[:, None]
    x = double_swish(conv2d(x, W["conv1_w"], W["conv1_b"], 1))
    x = double_swish(conv2d(x, W["conv2_w"], W["conv2_b"], 2))
    x = double_swish(conv2d(x, W["conv3_w"], W["conv3_b"], 2))          # [N,128,1,18]
    n, c, t, f = x.shape
    x = x.transpose(0, 2, 1, 3).reshape(n, c * f)
    x = x @ W["out_w"].T + W["out_b"]
    return basic_norm(x.astype(np.float32), W["out_norm_eps"])


def encoder_step(W, seg, h, c):
    """seg [N,9,80]; h [L,N,D]; c [L,N,H] -> encoder_out [N,J], h', c'"""
    L = h.shape[0]
    x = encoder_embed(W, seg)
    nh, nc = np.empty_like(h), np.empty_like(c)
    for l in range(L):
        p = "l%d_" % l
        g = x @ W[p + "w_ih"].T + W[p + "b_ih"] + h[l] @ W[p + "w_hh"].T + W[p + "b_hh"]
        H = c.shape[2]
        i, f, gg, o = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
        cn = sigmoid(f) * c[l] + sigmoid(i) * np.tanh(gg)
        m = sigmoid(o) * np.tanh(cn)
        hn = (m @ W[p + "w_hr"].T).astype(np.float32)
        nh[l], nc[l] = hn, cn
        x = x + hn
        ff = double_swish((x @ W[p + "ff1_w"].T + W[p + "ff1_b"]).astype(np.float32))
        x = x + (ff @ W[p + "ff2_w"].T + W[p + "ff2_b"]).astype(np.float32)
        x = basic_norm(x.astype(np.float32), W[p + "norm_eps"])
    eo = (x @ W["enc_proj_w"].T + W["enc_proj_b"]).astype(np.float32)
    return eo, nh, nc


def decoder(W, ctx):
    """ctx int [N,2] -> [N,J]"""
    e = W["emb"][ctx]                                  # [N,2,E]
    k = W["dec_conv_w"][:, 0, :]                       # [E,2] depthwise
    y = np.maximum((e * k.T[None]).sum(1), 0.0).astype(np.float32)
    return (y @ W["dec_proj_w"].T + W["dec_proj_b"]).astype(np.float32)


def joiner(W, eo, do):
    return (np.tanh(eo + do) @ W["join_w"].T + W["join_b"]).astype(np.float32)
