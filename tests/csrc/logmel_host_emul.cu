// Host emulation of the K1 kernel's stages: the SAME __host__ __device__ functions as the CUDA
// kernel (multimodal_lipread_b200/csrc/logmel_core.cuh), driven task by task on the CPU.  Test-only:
// built and used by tests/test_logmel_host_emul.py so the FFT / split / mel math is checked against
// the oracle in the CPU suite, before the kernel ever sees a GPU.
#include "../../multimodal_lipread_b200/csrc/logmel_core.cuh"
#include <cmath>
#include <vector>

using namespace lm;

static void build_plan(const float* window, const float* fb, Plan& p) {
    double s = 0.0;
    for (int i = 0; i < NFFT; ++i) s += double(window[i]) * double(window[i]);
    const double norm = 0.5 / std::sqrt(s);
    for (int i = 0; i < NFFT; ++i) p.win[i] = float(double(window[i]) * norm);
    const double PI = 3.14159265358979323846;
    for (int k2 = 0; k2 < 8; ++k2)
        for (int r = 0; r < 25; ++r) {
            const double a = -2.0 * PI * double(r * k2) / 200.0;
            p.tw200[k2][r] = make_float2(float(std::cos(a)), float(std::sin(a)));
        }
    for (int k = 0; k <= 100; ++k) {
        const double a = -2.0 * PI * double(k) / 400.0;
        p.tw400[k] = make_float2(float(std::cos(a)), float(std::sin(a)));
    }
    p.status = 0;
    for (int m = 0; m < NMEL; ++m)
        if (!plan_mel(fb, m, &p.mel_lo[m], &p.mel_nq[m], p.mel_w[m])) p.status = 1;
}

// wav [20000] -> raw log-mel [80][126]; returns plan.status.  Driven the way the kernel's warps drive it: groups of
// four frames, 880 staged (reflect-padded) samples per group, stage A on registers with the lane's window / twiddle
// values, 25-point DFTs in place, powers formed from (Z[k], Z[200-k]) pairs, banded mel filters.
extern "C" int logmel_host_emul(const float* wav, const float* window, const float* fb, float* out) {
    Plan p;
    build_plan(window, fb, p);
    std::vector<float> stage(GSAMP);
    std::vector<float2> Z(NHALF);
    std::vector<float> P(NBINS + 2);
    for (int g = 0; g < GROUPS; ++g) {
        for (int i = 0; i < GSAMP; ++i) stage[i] = padded_sample(wav, GF * HOP * g + i);
        for (int f = 0; f < GF; ++f) {
            const int t = GF * g + f;
            if (t >= NFRAMES) break;
            const float2* xf = reinterpret_cast<const float2*>(stage.data() + HOP * f);
            for (int r = 0; r < 25; ++r) {
                float w16[16];
                float2 tw7[7], v[8];
                for (int j = 0; j < 8; ++j) {
                    w16[2 * j] = p.win[2 * (r + 25 * j)];
                    w16[2 * j + 1] = p.win[2 * (r + 25 * j) + 1];
                    v[j] = xf[r + 25 * j];
                }
                for (int k2 = 1; k2 < 8; ++k2) tw7[k2 - 1] = p.tw200[k2][r];
                stage_a_regs(v, w16, tw7);
                for (int k2 = 0; k2 < 8; ++k2) Z[k2 * 25 + r] = v[k2];
            }
            float2 y[8][25], z[8][25];
            for (int k2 = 0; k2 < 8; ++k2) { stage_b_load(Z.data(), k2, y[k2]); dft25(y[k2], z[k2]); }
            for (int k2 = 0; k2 < 8; ++k2) stage_b_store(Z.data(), k2, z[k2]);
            for (int k = 0; k <= 100; ++k) {
                const float2 pw = stage_c_pair(Z[k], Z[k == 0 ? 0 : NHALF - k], p.tw400[k]);
                P[k] = pw.x;
                P[NHALF - k] = pw.y;
            }
            for (int m = 0; m < NMEL; ++m)
                out[m * NFRAMES + t] = stage_d(P.data(), m, p.mel_lo, p.mel_nq, &p.mel_w[0][0]);
        }
    }
    return p.status;
}

// complex 25-point and 8-point DFT codelets exposed for direct checks
extern "C" void dft25_host(const float* in, float* out) {
    dft25(reinterpret_cast<const float2*>(in), reinterpret_cast<float2*>(out));
}
extern "C" void dft8_host(float* v) { dft8(reinterpret_cast<float2*>(v)); }
