// K1 log-mel: per-task math shared by the CUDA kernel (logmel.cu) and the host emulation used by
// the CPU test-suite (tests/csrc/logmel_host_emul.cu).  Everything here is __host__ __device__ and
// works on plain pointers, so the exact same code path is exercised with and without a GPU.
//
// Algorithm (reference audio/utils/audio_processor.py:48-52 == torchaudio spectrogram + mel scale):
//   a real 400-point DFT per frame, evaluated as ONE complex 200-point FFT of z[n] = x[2n] + i x[2n+1]
//   (200 = 8 x 25, the 25 again 5 x 5) followed by the even/odd split
//        X[k] = E[k] + W400^k O[k],   X[200-k] = conj(E[k] - W400^k O[k])
//   so only 201 bins are ever formed.  Power -> banded (<=16 tap) mel filters -> ln(. + 1e-9).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace lm {

constexpr int NFFT = 400, HOP = 160, NHALF = 200, NBINS = 201, NMEL = 80;
constexpr int NSAMP = 20000, NFRAMES = 126, PAD = 200;
constexpr int MAXTAPS = 16;
constexpr int GF = 4;                               // frames per warp group: 4 frames x 8 residues = one 25-point DFT per lane
constexpr int GROUPS = (NFRAMES + GF - 1) / GF;     // 32 groups per clip (the last one holds frames 124, 125)
constexpr int GSAMP = (GF - 1) * HOP + NFFT;        // 880 consecutive padded samples cover a group's four frames
constexpr int PLD = NBINS;                          // power row stride inside a group (odd)
constexpr int LLD = 127;              // leading dimension of the 80 x 126 log-mel tile

struct Plan {
    float win[NFFT];                  // hann[n] * 0.5 / sqrt(sum hann^2)   (0.5 = even/odd split factor)
    float2 tw200[8][25];              // W200^(r*k2), [k2][r]
    alignas(16) float mel_w[NMEL][MAXTAPS];   // the window's weights fb[mel_lo + j][m] (zero outside the filter)
    float2 tw400[101];                // W400^k, k = 0..100
    int mel_lo[NMEL];                 // first bin of the filter's 16-bin window: min(first non-zero bin, 201 - 16)
    int mel_nq[NMEL];                 // tap quads to sum: ceil((last non-zero bin + 1 - mel_lo) / 4), 1..4
    int status;                       // 0 ok, 1 = a filter had more than MAXTAPS taps
    int pad_;
};
static_assert(sizeof(Plan) % 16 == 0 && offsetof(Plan, mel_w) % 16 == 0, "Plan is copied with 16-byte vectors");

#define LM_HD __host__ __device__ __forceinline__

LM_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
LM_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
LM_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
LM_HD float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// Sample index of padded position p (0 .. NSAMP+2*PAD-1) under torch's "reflect" padding.
LM_HD int reflect_index(int p) {
    int i = p - PAD;
    if (i < 0) i = -i;
    if (i >= NSAMP) i = 2 * (NSAMP - 1) - i;
    return i;
}

// In-place forward 8-point DFT (e^{-2 pi i jk/8}), natural order in and out.
LM_HD void dft8(float2* v) {
    const float h = 0.70710678118654752440f;
    float2 a0 = cadd(v[0], v[4]), a1 = csub(v[0], v[4]);
    float2 a2 = cadd(v[2], v[6]), a3 = mul_neg_i(csub(v[2], v[6]));
    float2 a4 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
    float2 a6 = cadd(v[3], v[7]), a7 = mul_neg_i(csub(v[3], v[7]));
    float2 b0 = cadd(a0, a2), b2 = csub(a0, a2);          // even outputs 0,4 / 2,6 (even-index half)
    float2 b1 = cadd(a1, a3), b3 = csub(a1, a3);
    float2 b4 = cadd(a4, a6), b6 = mul_neg_i(csub(a4, a6));
    float2 b5 = cadd(a5, a7), b7 = csub(a5, a7);
    // odd half twiddles: W8^1 = h(1 - i), W8^3 = -h(1 + i)
    float2 t5 = make_float2(h * (b5.x + b5.y), h * (b5.y - b5.x));
    float2 t7 = make_float2(h * (b7.y - b7.x), -h * (b7.x + b7.y));
    v[0] = cadd(b0, b4); v[4] = csub(b0, b4);
    v[2] = cadd(b2, b6); v[6] = csub(b2, b6);
    v[1] = cadd(b1, t5); v[5] = csub(b1, t5);
    v[3] = cadd(b3, t7); v[7] = csub(b3, t7);
}

// Forward 5-point DFT of x0..x4 (stride 1 through references).
LM_HD void dft5(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4) {
    const float c1 = 0.30901699437494742410f;    // cos(2 pi / 5)
    const float c2 = -0.80901699437494742410f;   // cos(4 pi / 5)
    const float s1 = 0.95105651629515357212f;    // sin(2 pi / 5)
    const float s2 = 0.58778525229247312917f;    // sin(4 pi / 5)
    float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
    float2 m1 = make_float2(x0.x + c1 * t1.x + c2 * t2.x, x0.y + c1 * t1.y + c2 * t2.y);
    float2 m2 = make_float2(x0.x + c2 * t1.x + c1 * t2.x, x0.y + c2 * t1.y + c1 * t2.y);
    float2 u1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
    float2 u2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
    x0 = make_float2(x0.x + t1.x + t2.x, x0.y + t1.y + t2.y);
    x1 = make_float2(m1.x + u1.y, m1.y - u1.x);   // m1 - i u1
    x4 = make_float2(m1.x - u1.y, m1.y + u1.x);
    x2 = make_float2(m2.x + u2.y, m2.y - u2.x);
    x3 = make_float2(m2.x - u2.y, m2.y + u2.x);
}

// W25^(a*b) for a, b in 1..4 (row a-1, col b-1): exp(-2 pi i a b / 25).
struct W25 { float2 w[4][4]; };
LM_HD W25 w25_table() {
    W25 t;
    // cos / sin of 2 pi j / 25 for j = 0..16
    const float c[17] = {1.0f, 0.96858316112863108f, 0.87630668004386358f, 0.72896862742141155f,
                         0.53582679497899666f, 0.30901699437494742f, 0.06279051952931337f,
                         -0.18738131458572463f, -0.42577929156507272f, -0.63742398974868975f,
                         -0.80901699437494742f, -0.92977648588825146f, -0.99211470131447788f,
                         -0.99211470131447788f, -0.92977648588825146f, -0.80901699437494742f,
                         -0.63742398974868975f};
    const float s[17] = {0.0f, 0.24868988716485479f, 0.48175367410171532f, 0.68454710592868873f,
                         0.84432792550201508f, 0.95105651629515357f, 0.99802672842827156f,
                         0.98228725072868872f, 0.90482705246601958f, 0.77051324277578925f,
                         0.58778525229247313f, 0.36812455268467797f, 0.12533323356430426f,
                         -0.12533323356430426f, -0.36812455268467797f, -0.58778525229247313f,
                         -0.77051324277578925f};
#pragma unroll
    for (int a = 1; a <= 4; ++a)
#pragma unroll
        for (int b = 1; b <= 4; ++b) t.w[a - 1][b - 1] = make_float2(c[a * b], -s[a * b]);
    return t;
}

// Forward 25-point DFT: in y[r] (r = 0..24), out z[k] (k = 0..24), both natural order, in registers.
LM_HD void dft25(const float2* y, float2* z) {
    float2 a[5][5];   // a[n1][n2] = y[n1 + 5 n2]
#pragma unroll
    for (int n1 = 0; n1 < 5; ++n1)
#pragma unroll
        for (int n2 = 0; n2 < 5; ++n2) a[n1][n2] = y[n1 + 5 * n2];
    const W25 tw = w25_table();
#pragma unroll
    for (int n1 = 0; n1 < 5; ++n1) {
        dft5(a[n1][0], a[n1][1], a[n1][2], a[n1][3], a[n1][4]);       // over n2 -> index k2'
        if (n1 > 0) {
#pragma unroll
            for (int k2 = 1; k2 < 5; ++k2) a[n1][k2] = cmul(a[n1][k2], tw.w[n1 - 1][k2 - 1]);
        }
    }
#pragma unroll
    for (int k2 = 0; k2 < 5; ++k2) {
        dft5(a[0][k2], a[1][k2], a[2][k2], a[3][k2], a[4][k2]);       // over n1 -> index k1'
#pragma unroll
        for (int k1 = 0; k1 < 5; ++k1) z[k2 + 5 * k1] = a[k1][k2];
    }
}

// padded position p (0 .. NSAMP + 2 PAD - 1) -> sample under reflect padding; 0 beyond the padded signal (the two
// frames that pad the last group)
LM_HD float padded_sample(const float* __restrict__ wav, int p) {
    return p < NSAMP + 2 * PAD ? wav[reflect_index(p)] : 0.f;
}

// ---- stage A in registers: v[j] = the frame's sample pair at complex point r + 25 j (already reflect-padded),
// w16[2j], w16[2j+1] = the window at those two samples, tw7[k2-1] = W200^(r k2).  Out: the twiddled 8-point DFT,
// v[k2] = Y[k2*25 + r].
LM_HD void stage_a_regs(float2* v, const float* w16, const float2* tw7) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = make_float2(v[j].x * w16[2 * j], v[j].y * w16[2 * j + 1]);
    dft8(v);
#pragma unroll
    for (int k2 = 1; k2 < 8; ++k2) v[k2] = cmul(v[k2], tw7[k2 - 1]);
}

// ---- stage A: task (frame f, residue r): windowed load, 8-point DFTs, outer twiddle ------------
// Y layout per frame: Y[k2*25 + r] (float2).  `wav` is one clip (NSAMP floats), t the frame index.
LM_HD void stage_a(const float* __restrict__ wav, int t, int r, const float* __restrict__ win,
                   const float2* __restrict__ tw200 /*[8][25]*/, float2* __restrict__ Yf) {
    float2 v[8];
    const int base = HOP * t;               // padded position of the frame start
    const bool interior = (base >= PAD) && (base + NFFT - PAD <= NSAMP);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = 2 * (r + 25 * j);     // even sample of complex point r + 25 j
        float x0, x1;
        if (interior) {
            const float2 p = *reinterpret_cast<const float2*>(wav + (base - PAD + n));
            x0 = p.x; x1 = p.y;
        } else {
            x0 = wav[reflect_index(base + n)];
            x1 = wav[reflect_index(base + n + 1)];
        }
        v[j] = make_float2(x0 * win[n], x1 * win[n + 1]);
    }
    dft8(v);
    Yf[r] = v[0];
#pragma unroll
    for (int k2 = 1; k2 < 8; ++k2) Yf[k2 * 25 + r] = cmul(v[k2], tw200[k2 * 25 + r]);
}

// ---- stage B: task (frame f, k2): 25-point DFT over r; Z[k2 + 8 k1] ----------------------------
LM_HD void stage_b_load(const float2* __restrict__ Yf, int k2, float2* y) {
#pragma unroll
    for (int r = 0; r < 25; ++r) y[r] = Yf[k2 * 25 + r];
}
LM_HD void stage_b_store(float2* __restrict__ Zf, int k2, const float2* z) {
#pragma unroll
    for (int k1 = 0; k1 < 25; ++k1) Zf[k2 + 8 * k1] = z[k1];
}

// ---- stage C: task (frame f, k in 0..100): power of bins k and 200-k ---------------------------
// zk = Z[k], zn = Z[200 - k] (Z[0] for k = 0), tw = W400^k  ->  (|X[k]|^2, |X[200-k]|^2)
LM_HD float2 stage_c_pair(float2 zk, float2 zn, float2 tw) {
    const float2 E = make_float2(zk.x + zn.x, zk.y - zn.y);
    const float2 O = make_float2(zk.y + zn.y, zn.x - zk.x);
    const float2 T = cmul(tw, O);
    const float ar = E.x + T.x, ai = E.y + T.y;
    const float br = E.x - T.x, bi = E.y - T.y;
    return make_float2(ar * ar + ai * ai, br * br + bi * bi);
}
LM_HD void stage_c(const float2* __restrict__ Zf, int k, const float2* __restrict__ tw400,
                   float* __restrict__ Pf) {
    const float2 pw = stage_c_pair(Zf[k], Zf[k == 0 ? 0 : NHALF - k], tw400[k]);
    Pf[k] = pw.x;
    Pf[NHALF - k] = pw.y;                   // k == 100 writes the same bin twice with the same value
}

// ln(acc + 1e-9): on the device lg2.approx * ln 2 (relative error 2^-22 on values of magnitude <= 21: far inside the
// 1e-4 bar), on the host (CPU emulation in the test-suite) the C library's logf
LM_HD float log_eps(float acc) {
#ifdef __CUDA_ARCH__
    return __logf(acc + 1e-9f);
#else
    return logf(acc + 1e-9f);
#endif
}

// ---- stage D: task (frame f, mel m): banded filter + log --------------------------------------
// One filter's entry of the plan from the [201][80] filterbank: a 16-bin window that always lies inside the power row,
// so the kernel needs neither bounds checks nor predicates; zero weights around the filter add exact zeros.
// Returns false if the filter has more than MAXTAPS taps.
LM_HD bool plan_mel(const float* __restrict__ fb, int m, int* lo_out, int* nq_out, float* w16) {
    int lo = -1, hi = -1;
    for (int k = 0; k < NBINS; ++k)
        if (fb[k * NMEL + m] != 0.f) { if (lo < 0) lo = k; hi = k; }
    if (lo < 0) { lo = 0; hi = 0; }
    const bool ok = hi - lo + 1 <= MAXTAPS;
    if (!ok) hi = lo + MAXTAPS - 1;
    const int w0 = lo < NBINS - MAXTAPS ? lo : NBINS - MAXTAPS;
    *lo_out = w0;
    *nq_out = (hi + 1 - w0 + 3) / 4;
    for (int j = 0; j < MAXTAPS; ++j) w16[j] = (w0 + j >= lo && w0 + j <= hi) ? fb[(w0 + j) * NMEL + m] : 0.f;
    return ok;
}
LM_HD float stage_d(const float* __restrict__ Pf, int m, const int* __restrict__ mel_lo,
                    const int* __restrict__ mel_nq, const float* __restrict__ mel_w /*[NMEL][MAXTAPS]*/) {
    const float* pp = Pf + mel_lo[m];
    float acc = 0.f;
    for (int j = 0; j < 4 * mel_nq[m]; ++j) acc = fmaf(pp[j], mel_w[m * MAXTAPS + j], acc);
    return log_eps(acc);
}

}  // namespace lm
