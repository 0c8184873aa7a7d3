// =====================================================================================
//  vtgs_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
//  A plain C++ restatement of the splatting arithmetic on VTGaussian-SLAM's hot path,
//  used ONLY as the checker in tests/, __graft_entry__.smoke() and bench.py's
//  cpu_baseline / --impl reference legs.  Nothing in the product package
//  (vtgaussian_slam_b200/, diff_gaussian_rasterization/) may import, link or call it.
//
//  PARITY UNPINNED.  The algorithm restated here lives in a third-party dependency that
//  is absent from /root/reference: the pip module `diff_gaussian_rasterization`
//  from git+https://github.com/pengchongH/diff-gaussian-rasterization-w-depth-smallerGSradii.git
//  (reference requirements.txt:18, no pinned commit), lineage graphdeco-inria/
//  diff-gaussian-rasterization -> JonathonLuiten/...-w-depth -> this fork.  The
//  reference holds no tests, golden vectors or fixtures for this path (SURVEY.md 4,
//  8(c)), so this file restates the published upstream algorithm
//  and anchors on the reference's own call sites:
//     src/vtgaussian_slam.py:461,466,747  (Renderer(raster_settings=cam)(**rendervar))
//     utils/recon_helpers.py:14-26        (the 11 raster settings)
//     utils/slam_helpers.py:127-160,217-287,323-385 (what the inputs are)
//  The one delta the fork's name advertises ("smallerGSradii") is unknown offline:
//  the 3-sigma radius multiplier is therefore a runtime parameter
//  (VtgsCamera.radius_sigma_mult, default 3.0 = upstream).
//
//  Arithmetic contract (shared, as a written spec, with the CUDA kernels -- see
//  DESIGN.md "Arithmetic spec"): everything in the forward is IEEE fp32 with a fixed
//  operation order; a*b+c is fused ONLY where this file writes fmaf().  Compile with
//  -ffp-contract=off.  exp() is the polynomial vexpf() below (<= 1 ulp), not libm, so
//  that CPU and GPU take identical alpha / T-threshold decisions and every integer
//  output (radii, tiles_touched, sort keys, tile ranges, n_contrib) is bit-exact.
//  float->int casts saturate like CUDA's cvt.rzi.s32.f32 (the reference runs on CUDA).
// =====================================================================================
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../include/vtgs.h"

namespace {

// ---------------------------------------------------------------- arithmetic helpers
inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// CUDA float->int32 conversion semantics (round toward zero, saturate, NaN -> 0).
inline int f2i_sat(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (-2147483647 - 1);
    return (int)v;
}

// exp(x) for the blend: Cody-Waite reduction + degree-5 polynomial in r (Cephes expf
// coefficients), fixed fmaf order, scaling by integer exponent add.  Defined for all x
// by clamping to [-87, 88].
inline float vexpf(float x) {
    x = fminf(fmaxf(x, -87.0f), 88.0f);
    const float t = fmaf(x, 1.44269504088896341f, 12582912.0f);   // 1.5 * 2^23
    const float n = t - 12582912.0f;
    float r = fmaf(n, -0.693145751953125f, x);
    r = fmaf(n, -1.428606765330187045e-06f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    const float r2 = r * r;
    float e = fmaf(p, r2, r);
    e = e + 1.0f;
    const int ni = (int)n;
    return u2f(f2u(e) + ((uint32_t)ni << 23));
}

// m is the flat row-vector-convention matrix: row r of the transform is m[r], m[4+r], ...
inline float xform_row(const float* m, int r, float x, float y, float z) {
    return fmaf(m[8 + r], z, fmaf(m[4 + r], y, m[r] * x)) + m[12 + r];
}
inline float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
    return fmaf(a2, b2, fmaf(a1, b1, a0 * b0));
}

// Rotation matrix of quaternion q = (r, x, y, z), NOT re-normalised here (upstream
// leaves that to the caller: reference utils/slam_helpers.py:155 normalises).
inline void quat_to_R(const float* q, float R[9]) {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1.0f - 2.0f * fmaf(z, z, y * y);
    R[1] = 2.0f * fmaf(x, y, -(r * z));
    R[2] = 2.0f * fmaf(x, z, r * y);
    R[3] = 2.0f * fmaf(x, y, r * z);
    R[4] = 1.0f - 2.0f * fmaf(z, z, x * x);
    R[5] = 2.0f * fmaf(y, z, -(r * x));
    R[6] = 2.0f * fmaf(x, z, -(r * y));
    R[7] = 2.0f * fmaf(y, z, r * x);
    R[8] = 1.0f - 2.0f * fmaf(y, y, x * x);
}

struct Geom {
    float px, py, depth;
    float ca, cb, cc, op;           // conic (A, B, C) and opacity
    float cov_a, cov_b, cov_c;      // cov2D incl. low-pass (kept for the backward)
    int radius;
    int rminx, rminy, rmaxx, rmaxy; // tile rect, band-clipped in y
    uint32_t tiles;
};

struct Oracle {
    VtgsCamera cam{};
    int64_t N = 0;
    int C = 0;
    int gx = 0, gy = 0, row0 = 0, row1 = 0;
    std::vector<float> means3D, scales, rotations, opacities, colors;
    std::vector<Geom> g;
    std::vector<uint32_t> offsets;
    std::vector<uint64_t> keys;
    std::vector<uint32_t> vals;
    std::vector<uint32_t> ranges;   // [tiles][2]
    std::vector<float> final_T, out_color, out_depth;
    std::vector<uint32_t> n_contrib;
    uint64_t R = 0;
    uint64_t pair_tests = 0;        // S = sum n_contrib-style work counter (all tests done)
    uint64_t contributing = 0;      // K
};

void cov3d_from(const float* sc, float mod, const float* rot, float S[6]) {
    float Rm[9];
    quat_to_R(rot, Rm);
    const float s0 = mod * sc[0], s1 = mod * sc[1], s2 = mod * sc[2];
    const float a0 = s0 * s0, a1 = s1 * s1, a2 = s2 * s2;
    auto sig = [&](int i, int j) {
        return fmaf(a2 * Rm[3 * i + 2], Rm[3 * j + 2],
                    fmaf(a1 * Rm[3 * i + 1], Rm[3 * j + 1], (a0 * Rm[3 * i + 0]) * Rm[3 * j + 0]));
    };
    S[0] = sig(0, 0); S[1] = sig(0, 1); S[2] = sig(0, 2);
    S[3] = sig(1, 1); S[4] = sig(1, 2); S[5] = sig(2, 2);
}

// SURVEY.md Appendix A.1.
void preprocess_one(const Oracle& o, int64_t i, Geom& g) {
    const VtgsCamera& cam = o.cam;
    g = Geom{};
    const float* V = cam.viewmatrix;
    const float* P = cam.projmatrix;
    const float x = o.means3D[3 * i], y = o.means3D[3 * i + 1], z = o.means3D[3 * i + 2];
    const float tx = xform_row(V, 0, x, y, z);
    const float ty = xform_row(V, 1, x, y, z);
    const float tz = xform_row(V, 2, x, y, z);
    if (tz <= VTGS_NEAR_CULL) return;
    const float hx = xform_row(P, 0, x, y, z);
    const float hy = xform_row(P, 1, x, y, z);
    const float hw = xform_row(P, 3, x, y, z);
    const float pw = 1.0f / (hw + VTGS_EPS_W);
    const float ndcx = hx * pw, ndcy = hy * pw;

    float S[6];
    cov3d_from(&o.scales[3 * i], cam.scale_modifier, &o.rotations[4 * i], S);

    const float W = (float)cam.image_width, H = (float)cam.image_height;
    const float fx = W / (2.0f * cam.tanfovx);
    const float fy = H / (2.0f * cam.tanfovy);
    const float limx = VTGS_FRUSTUM_MULT * cam.tanfovx;
    const float limy = VTGS_FRUSTUM_MULT * cam.tanfovy;
    const float cx = fminf(limx, fmaxf(-limx, tx / tz)) * tz;
    const float cy = fminf(limy, fmaxf(-limy, ty / tz)) * tz;
    const float J00 = fx / tz;
    const float J02 = -(fx * cx) / (tz * tz);
    const float J11 = fy / tz;
    const float J12 = -(fy * cy) / (tz * tz);
    float m0[3], m1[3];
    for (int k = 0; k < 3; ++k) {
        m0[k] = fmaf(J02, V[4 * k + 2], J00 * V[4 * k + 0]);
        m1[k] = fmaf(J12, V[4 * k + 2], J11 * V[4 * k + 1]);
    }
    const float u0 = dot3(S[0], S[1], S[2], m0[0], m0[1], m0[2]);
    const float u1 = dot3(S[1], S[3], S[4], m0[0], m0[1], m0[2]);
    const float u2 = dot3(S[2], S[4], S[5], m0[0], m0[1], m0[2]);
    const float v0 = dot3(S[0], S[1], S[2], m1[0], m1[1], m1[2]);
    const float v1 = dot3(S[1], S[3], S[4], m1[0], m1[1], m1[2]);
    const float v2 = dot3(S[2], S[4], S[5], m1[0], m1[1], m1[2]);
    const float a = dot3(m0[0], m0[1], m0[2], u0, u1, u2) + VTGS_LOWPASS;
    const float b = dot3(m1[0], m1[1], m1[2], u0, u1, u2);
    const float c = dot3(m1[0], m1[1], m1[2], v0, v1, v2) + VTGS_LOWPASS;

    const float det = fmaf(a, c, -(b * b));
    if (det == 0.0f) return;
    const float inv = 1.0f / det;
    const float mid = 0.5f * (a + c);
    const float sq = sqrtf(fmaxf(VTGS_LAMBDA_FLOOR, fmaf(mid, mid, -det)));
    const float lam = fmaxf(mid + sq, mid - sq);
    const int radius = f2i_sat(ceilf(cam.radius_sigma_mult * sqrtf(lam)));
    // ndc2Pix is written with double literals upstream: evaluated in fp64, rounded once.
    const float px = (float)((((double)ndcx + 1.0) * (double)cam.image_width - 1.0) * 0.5);
    const float py = (float)((((double)ndcy + 1.0) * (double)cam.image_height - 1.0) * 0.5);
    const float rf = (float)radius;
    auto clampi = [](int v, int hi) { return std::min(hi, std::max(0, v)); };
    const int rminx = clampi(f2i_sat((px - rf) / 16.0f), o.gx);
    const int rminy = clampi(f2i_sat((py - rf) / 16.0f), o.gy);
    const int rmaxx = clampi(f2i_sat((((px + rf) + 16.0f) - 1.0f) / 16.0f), o.gx);
    const int rmaxy = clampi(f2i_sat((((py + rf) + 16.0f) - 1.0f) / 16.0f), o.gy);
    if ((rmaxx - rminx) * (rmaxy - rminy) == 0) return;

    g.px = px; g.py = py; g.depth = tz;
    g.ca = c * inv; g.cb = -b * inv; g.cc = a * inv; g.op = o.opacities[i];
    g.cov_a = a; g.cov_b = b; g.cov_c = c;
    g.radius = radius;
    g.rminx = rminx; g.rmaxx = rmaxx;
    g.rminy = std::max(rminy, o.row0);
    g.rmaxy = std::min(rmaxy, o.row1);
    const int hgt = std::max(0, g.rmaxy - g.rminy);
    g.tiles = (uint32_t)((rmaxx - rminx) * hgt);
}

inline float power_of(const Geom& g, float dx, float dy) {
    const float q = fmaf(g.ca * dx, dx, (g.cc * dy) * dy);
    return fmaf(-0.5f, q, -((g.cb * dx) * dy));
}

void run_forward(Oracle& o) {
    const VtgsCamera& cam = o.cam;
    const int W = cam.image_width, H = cam.image_height;
    o.gx = (W + 15) / 16; o.gy = (H + 15) / 16;
    o.row0 = 0; o.row1 = o.gy;
    if (cam.tile_row_end > cam.tile_row_begin) {
        o.row0 = std::max(0, cam.tile_row_begin);
        o.row1 = std::min(o.gy, cam.tile_row_end);
    }
    const int64_t N = o.N;
    const int C = o.C;
    o.g.resize(N);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) preprocess_one(o, i, o.g[i]);

    // A.2 binning: inclusive scan, duplicateWithKeys, stable sort, identifyTileRanges.
    o.offsets.resize(N);
    uint64_t acc = 0;
    for (int64_t i = 0; i < N; ++i) { acc += o.g[i].tiles; o.offsets[i] = (uint32_t)acc; }
    o.R = acc;
    std::vector<std::pair<uint64_t, uint32_t>> kv(o.R);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        const Geom& g = o.g[i];
        if (g.radius <= 0 || g.tiles == 0) continue;
        uint64_t off = (i == 0) ? 0 : o.offsets[i - 1];
        for (int ty = g.rminy; ty < g.rmaxy; ++ty)
            for (int tx = g.rminx; tx < g.rmaxx; ++tx) {
                uint64_t key = (uint64_t)(ty * o.gx + tx);
                key <<= 32;
                key |= f2u(g.depth);
                kv[off++] = {key, (uint32_t)i};
            }
    }
    std::stable_sort(kv.begin(), kv.end(),
                     [](const auto& a, const auto& b) { return a.first < b.first; });
    o.keys.resize(o.R); o.vals.resize(o.R);
    for (uint64_t j = 0; j < o.R; ++j) { o.keys[j] = kv[j].first; o.vals[j] = kv[j].second; }
    const int tiles = o.gx * o.gy;
    o.ranges.assign((size_t)tiles * 2, 0);
    for (uint64_t j = 0; j < o.R; ++j) {
        const uint32_t t = (uint32_t)(o.keys[j] >> 32);
        if (j == 0) o.ranges[2 * t] = 0;
        else {
            const uint32_t pt = (uint32_t)(o.keys[j - 1] >> 32);
            if (pt != t) { o.ranges[2 * pt + 1] = (uint32_t)j; o.ranges[2 * t] = (uint32_t)j; }
        }
        if (j == o.R - 1) o.ranges[2 * t + 1] = (uint32_t)o.R;
    }

    // A.3 forward blend.
    const size_t P = (size_t)W * H;
    o.final_T.assign(P, 1.0f);
    o.n_contrib.assign(P, 0);
    o.out_color.assign(P * C, 0.0f);
    o.out_depth.assign(P, 0.0f);
    for (int ch = 0; ch < C; ++ch) {
        const float bgc = ch < 3 ? cam.bg[ch] : 0.0f;
        for (size_t p = 0; p < P; ++p) o.out_color[ch * P + p] = bgc;   // rows outside the band
    }
    uint64_t tests = 0, contrib = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : tests, contrib)
    for (int tile = o.row0 * o.gx; tile < o.row1 * o.gx; ++tile) {
        const int tx0 = (tile % o.gx) * 16, ty0 = (tile / o.gx) * 16;
        const uint32_t rb = o.ranges[2 * tile], re = o.ranges[2 * tile + 1];
        std::vector<float> Cacc(C);
        for (int py = ty0; py < std::min(ty0 + 16, H); ++py)
            for (int px = tx0; px < std::min(tx0 + 16, W); ++px) {
                const float pxf = (float)px, pyf = (float)py;
                float T = 1.0f, D = 0.0f;
                std::fill(Cacc.begin(), Cacc.end(), 0.0f);
                uint32_t contributor = 0, last = 0;
                for (uint32_t j = rb; j < re; ++j) {
                    ++contributor;
                    const uint32_t id = o.vals[j];
                    const Geom& g = o.g[id];
                    const float dx = g.px - pxf, dy = g.py - pyf;
                    const float power = power_of(g, dx, dy);
                    if (power > 0.0f) continue;
                    const float alpha = fminf(VTGS_ALPHA_MAX, g.op * vexpf(power));
                    if (alpha < VTGS_ALPHA_MIN) continue;
                    const float test_T = T * (1.0f - alpha);
                    if (test_T < VTGS_T_MIN) break;
                    const float* col = &o.colors[(size_t)id * C];
                    for (int ch = 0; ch < C; ++ch) Cacc[ch] = fmaf(col[ch] * alpha, T, Cacc[ch]);
                    D = fmaf(g.depth * alpha, T, D);
                    T = test_T;
                    last = contributor;
                    ++contrib;
                }
                tests += contributor;
                const size_t pid = (size_t)py * W + px;
                o.final_T[pid] = T;
                o.n_contrib[pid] = last;
                for (int ch = 0; ch < C; ++ch)
                    o.out_color[ch * P + pid] = fmaf(T, ch < 3 ? cam.bg[ch] : 0.0f, Cacc[ch]);
                o.out_depth[pid] = D;
            }
    }
    o.pair_tests = tests;
    o.contributing = contrib;
}

// A.5 backward blend + A.6 backward preprocess.  Per-pair arithmetic in fp32 as upstream;
// the per-Gaussian sums (global float atomics of arbitrary order upstream) are
// accumulated in fp64 so that the oracle is the order-independent limit.
void run_backward(const Oracle& o, const float* dL_dpix, float* dL_dmeans2D, float* dL_dcolors,
                  float* dL_dopacity, float* dL_dmeans3D, float* dL_dscales, float* dL_drot,
                  float* dL_dconic_out) {
    const VtgsCamera& cam = o.cam;
    const int W = cam.image_width, H = cam.image_height, C = o.C;
    const size_t P = (size_t)W * H;
    const int64_t N = o.N;
    const int NG = 6 + C;   // mean2D(2) conic(3) opacity(1) colour(C)
    std::vector<double> acc((size_t)N * NG, 0.0);

#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = o.row0 * o.gx; tile < o.row1 * o.gx; ++tile) {
        const int tx0 = (tile % o.gx) * 16, ty0 = (tile / o.gx) * 16;
        const uint32_t rb = o.ranges[2 * tile], re = o.ranges[2 * tile + 1];
        if (re <= rb) continue;
        std::vector<double> loc((size_t)(re - rb) * NG, 0.0);
        std::vector<float> accum(C), lastc(C), dpix(C);
        for (int py = ty0; py < std::min(ty0 + 16, H); ++py)
            for (int px = tx0; px < std::min(tx0 + 16, W); ++px) {
                const size_t pid = (size_t)py * W + px;
                const float pxf = (float)px, pyf = (float)py;
                const float T_final = o.final_T[pid];
                float T = T_final;
                const uint32_t last = o.n_contrib[pid];
                uint32_t contributor = re - rb;
                std::fill(accum.begin(), accum.end(), 0.0f);
                std::fill(lastc.begin(), lastc.end(), 0.0f);
                float last_alpha = 0.0f;
                float bg_dot = 0.0f;
                for (int ch = 0; ch < C; ++ch) {
                    dpix[ch] = dL_dpix[ch * P + pid];
                    bg_dot += (ch < 3 ? cam.bg[ch] : 0.0f) * dpix[ch];
                }
                const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
                for (uint32_t j = re; j-- > rb;) {
                    --contributor;
                    if (contributor >= last) continue;
                    const uint32_t id = o.vals[j];
                    const Geom& g = o.g[id];
                    const float dx = g.px - pxf, dy = g.py - pyf;
                    const float power = power_of(g, dx, dy);
                    if (power > 0.0f) continue;
                    const float G = vexpf(power);
                    const float alpha = fminf(VTGS_ALPHA_MAX, g.op * G);
                    if (alpha < VTGS_ALPHA_MIN) continue;
                    T = T / (1.0f - alpha);
                    const float dchannel_dcolor = alpha * T;
                    float dL_dalpha = 0.0f;
                    double* L = &loc[(size_t)(j - rb) * NG];
                    const float* col = &o.colors[(size_t)id * C];
                    for (int ch = 0; ch < C; ++ch) {
                        const float c = col[ch];
                        accum[ch] = last_alpha * lastc[ch] + (1.0f - last_alpha) * accum[ch];
                        lastc[ch] = c;
                        dL_dalpha += (c - accum[ch]) * dpix[ch];
                        L[6 + ch] += (double)(dchannel_dcolor * dpix[ch]);
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
                    const float dL_dG = g.op * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * g.ca - gdy * g.cb;
                    const float dG_ddely = -gdy * g.cc - gdx * g.cb;
                    L[0] += (double)(dL_dG * dG_ddelx * ddelx_dx);
                    L[1] += (double)(dL_dG * dG_ddely * ddely_dy);
                    L[2] += (double)(-0.5f * gdx * dx * dL_dG);
                    L[3] += (double)(-0.5f * gdx * dy * dL_dG);
                    L[4] += (double)(-0.5f * gdy * dy * dL_dG);
                    L[5] += (double)(G * dL_dalpha);
                }
            }
        for (uint32_t j = rb; j < re; ++j) {
            const uint32_t id = o.vals[j];
            for (int k = 0; k < NG; ++k) {
                const double v = loc[(size_t)(j - rb) * NG + k];
                if (v != 0.0) {
#pragma omp atomic
                    acc[(size_t)id * NG + k] += v;
                }
            }
        }
    }

    const float* V = cam.viewmatrix;
    const float* Pm = cam.projmatrix;
    const float Wf = (float)W, Hf = (float)H;
    const float fx = Wf / (2.0f * cam.tanfovx), fy = Hf / (2.0f * cam.tanfovy);
    const float limx = VTGS_FRUSTUM_MULT * cam.tanfovx, limy = VTGS_FRUSTUM_MULT * cam.tanfovy;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        const double* A = &acc[(size_t)i * NG];
        dL_dmeans2D[3 * i] = (float)A[0]; dL_dmeans2D[3 * i + 1] = (float)A[1]; dL_dmeans2D[3 * i + 2] = 0.0f;
        for (int ch = 0; ch < C; ++ch) dL_dcolors[(size_t)i * C + ch] = (float)A[6 + ch];
        dL_dopacity[i] = (float)A[5];
        if (dL_dconic_out) { dL_dconic_out[3 * i] = (float)A[2]; dL_dconic_out[3 * i + 1] = (float)A[3]; dL_dconic_out[3 * i + 2] = (float)A[4]; }
        for (int k = 0; k < 3; ++k) { dL_dmeans3D[3 * i + k] = 0.0f; dL_dscales[3 * i + k] = 0.0f; }
        for (int k = 0; k < 4; ++k) dL_drot[4 * i + k] = 0.0f;
        const Geom& g = o.g[i];
        if (g.radius <= 0) continue;

        // ---- (i) conic -> cov2D
        const float a = g.cov_a, b = g.cov_b, c = g.cov_c;
        const float gxx = (float)A[2], gxy = (float)A[3], gyy = (float)A[4];
        const float denom = a * c - b * b;
        const float d2inv = 1.0f / (denom * denom + 0.0000001f);
        float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
        if (d2inv != 0.0f) {
            dL_da = d2inv * (-c * c * gxx + 2.0f * b * c * gxy + (denom - a * c) * gyy);
            dL_dc = d2inv * (-a * a * gyy + 2.0f * a * b * gxy + (denom - a * c) * gxx);
            dL_db = d2inv * 2.0f * (b * c * gxx - (denom + 2.0f * b * b) * gxy + a * b * gyy);
        }
        // ---- recompute the forward quantities
        const float x = o.means3D[3 * i], y = o.means3D[3 * i + 1], z = o.means3D[3 * i + 2];
        const float tx = xform_row(V, 0, x, y, z), ty = xform_row(V, 1, x, y, z), tz = xform_row(V, 2, x, y, z);
        const float txtz = tx / tz, tytz = ty / tz;
        const float cx = fminf(limx, fmaxf(-limx, txtz)) * tz;
        const float cy = fminf(limy, fmaxf(-limy, tytz)) * tz;
        const float xmul = (txtz < -limx || txtz > limx) ? 0.0f : 1.0f;
        const float ymul = (tytz < -limy || tytz > limy) ? 0.0f : 1.0f;
        const float J00 = fx / tz, J02 = -(fx * cx) / (tz * tz), J11 = fy / tz, J12 = -(fy * cy) / (tz * tz);
        float m0[3], m1[3];
        for (int k = 0; k < 3; ++k) {
            m0[k] = J02 * V[4 * k + 2] + J00 * V[4 * k + 0];
            m1[k] = J12 * V[4 * k + 2] + J11 * V[4 * k + 1];
        }
        float S[6];
        cov3d_from(&o.scales[3 * i], cam.scale_modifier, &o.rotations[4 * i], S);
        const float Sm[3][3] = {{S[0], S[1], S[2]}, {S[1], S[3], S[4]}, {S[2], S[4], S[5]}};
        // ---- (ii) cov2D -> cov3D (unique upper-triangular parameters)
        float dS[3][3];   // gradient wrt the FULL symmetric matrix entries (off-diagonals halved)
        for (int k = 0; k < 3; ++k)
            for (int l = 0; l < 3; ++l)
                dS[k][l] = m0[k] * m0[l] * dL_da + 0.5f * (m0[k] * m1[l] + m0[l] * m1[k]) * dL_db + m1[k] * m1[l] * dL_dc;
        // ---- cov2D -> M rows -> J -> t
        float Sm0[3], Sm1[3];
        for (int k = 0; k < 3; ++k) {
            Sm0[k] = Sm[k][0] * m0[0] + Sm[k][1] * m0[1] + Sm[k][2] * m0[2];
            Sm1[k] = Sm[k][0] * m1[0] + Sm[k][1] * m1[1] + Sm[k][2] * m1[2];
        }
        float dm0[3], dm1[3];
        for (int k = 0; k < 3; ++k) {
            dm0[k] = 2.0f * Sm0[k] * dL_da + Sm1[k] * dL_db;
            dm1[k] = 2.0f * Sm1[k] * dL_dc + Sm0[k] * dL_db;
        }
        // m0_k = J00 * Wr[0][k] + J02 * Wr[2][k], Wr[l][k] = V[4k + l]
        float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
        for (int k = 0; k < 3; ++k) {
            dJ00 += dm0[k] * V[4 * k + 0];
            dJ02 += dm0[k] * V[4 * k + 2];
            dJ11 += dm1[k] * V[4 * k + 1];
            dJ12 += dm1[k] * V[4 * k + 2];
        }
        const float tz2 = 1.0f / (tz * tz), tz3 = tz2 / tz;
        const float dtx = xmul * (-fx * tz2) * dJ02;
        const float dty = ymul * (-fy * tz2) * dJ12;
        const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.0f * fx * cx) * tz3 * dJ02 + (2.0f * fy * cy) * tz3 * dJ12;
        // dL/dp = Wr^T dL/dt
        float dmean[3];
        for (int k = 0; k < 3; ++k) dmean[k] = V[4 * k + 0] * dtx + V[4 * k + 1] * dty + V[4 * k + 2] * dtz;
        // ---- (iii) projection path: NDC xy -> p
        const float hx = xform_row(Pm, 0, x, y, z), hy = xform_row(Pm, 1, x, y, z), hw = xform_row(Pm, 3, x, y, z);
        const float mw = 1.0f / (hw + VTGS_EPS_W);
        const float mul1 = hx * mw * mw, mul2 = hy * mw * mw;
        const float g2x = (float)A[0], g2y = (float)A[1];
        for (int k = 0; k < 3; ++k)
            dmean[k] += (Pm[4 * k + 0] * mw - Pm[4 * k + 3] * mul1) * g2x + (Pm[4 * k + 1] * mw - Pm[4 * k + 3] * mul2) * g2y;
        for (int k = 0; k < 3; ++k) dL_dmeans3D[3 * i + k] = dmean[k];
        // ---- (iv) cov3D -> scale, rotation.  Sigma = sum_k s_k^2 r_k r_k^T, r_k = column k of R.
        float Rm[9];
        quat_to_R(&o.rotations[4 * i], Rm);
        const float mod = cam.scale_modifier;
        float D[3][3];   // dL/dR[i][k]
        for (int k = 0; k < 3; ++k) {
            const float sk = mod * o.scales[3 * i + k];
            float Gr[3];
            for (int r = 0; r < 3; ++r) Gr[r] = dS[r][0] * Rm[0 * 3 + k] + dS[r][1] * Rm[1 * 3 + k] + dS[r][2] * Rm[2 * 3 + k];
            const float rGr = Rm[0 * 3 + k] * Gr[0] + Rm[1 * 3 + k] * Gr[1] + Rm[2 * 3 + k] * Gr[2];
            dL_dscales[3 * i + k] = 2.0f * sk * rGr * mod;
            for (int r = 0; r < 3; ++r) D[r][k] = 2.0f * sk * sk * Gr[r];
        }
        const float qr = o.rotations[4 * i], qx = o.rotations[4 * i + 1], qy = o.rotations[4 * i + 2], qz = o.rotations[4 * i + 3];
        dL_drot[4 * i + 0] = 2.0f * (qz * (D[1][0] - D[0][1]) + qy * (D[0][2] - D[2][0]) + qx * (D[2][1] - D[1][2]));
        dL_drot[4 * i + 1] = 2.0f * (qy * (D[0][1] + D[1][0]) + qz * (D[0][2] + D[2][0]) + qr * (D[2][1] - D[1][2])) - 4.0f * qx * (D[1][1] + D[2][2]);
        dL_drot[4 * i + 2] = 2.0f * (qx * (D[0][1] + D[1][0]) + qr * (D[0][2] - D[2][0]) + qz * (D[1][2] + D[2][1])) - 4.0f * qy * (D[0][0] + D[2][2]);
        dL_drot[4 * i + 3] = 2.0f * (qr * (D[1][0] - D[0][1]) + qx * (D[0][2] + D[2][0]) + qy * (D[1][2] + D[2][1])) - 4.0f * qz * (D[0][0] + D[1][1]);
    }
}

}  // namespace

// ------------------------------------------------------------------------ C interface
extern "C" {

void* vtgso_create(void) { return new Oracle(); }
void vtgso_destroy(void* h) { delete (Oracle*)h; }

// Runs A.1-A.3 for `channels` precomputed colour channels.  Returns R.
uint64_t vtgso_forward(void* h, const VtgsCamera* cam, int64_t N, int32_t channels,
                       const float* means3D, const float* scales, const float* rotations,
                       const float* opacities, const float* colors) {
    Oracle& o = *(Oracle*)h;
    o.cam = *cam; o.N = N; o.C = channels;
    o.means3D.assign(means3D, means3D + 3 * N);
    o.scales.assign(scales, scales + 3 * N);
    o.rotations.assign(rotations, rotations + 4 * N);
    o.opacities.assign(opacities, opacities + N);
    o.colors.assign(colors, colors + (size_t)channels * N);
    run_forward(o);
    return o.R;
}

void vtgso_get_dims(void* h, int32_t* gx, int32_t* gy, uint64_t* R, uint64_t* pair_tests, uint64_t* contributing) {
    Oracle& o = *(Oracle*)h;
    *gx = o.gx; *gy = o.gy; *R = o.R; *pair_tests = o.pair_tests; *contributing = o.contributing;
}

// Any destination may be NULL.
void vtgso_get_geometry(void* h, int32_t* radii, uint32_t* tiles_touched, float* means2D, float* depths,
                        float* conic_opacity, uint32_t* offsets, int32_t* rects) {
    Oracle& o = *(Oracle*)h;
    for (int64_t i = 0; i < o.N; ++i) {
        const Geom& g = o.g[i];
        if (radii) radii[i] = g.radius;
        if (tiles_touched) tiles_touched[i] = g.tiles;
        if (means2D) { means2D[2 * i] = g.px; means2D[2 * i + 1] = g.py; }
        if (depths) depths[i] = g.depth;
        if (conic_opacity) { conic_opacity[4 * i] = g.ca; conic_opacity[4 * i + 1] = g.cb; conic_opacity[4 * i + 2] = g.cc; conic_opacity[4 * i + 3] = g.op; }
        if (offsets) offsets[i] = o.offsets[i];
        if (rects) { rects[4 * i] = g.rminx; rects[4 * i + 1] = g.rminy; rects[4 * i + 2] = g.rmaxx; rects[4 * i + 3] = g.rmaxy; }
    }
}

void vtgso_get_binning(void* h, uint64_t* keys, uint32_t* vals, uint32_t* ranges) {
    Oracle& o = *(Oracle*)h;
    if (keys) std::memcpy(keys, o.keys.data(), o.R * 8);
    if (vals) std::memcpy(vals, o.vals.data(), o.R * 4);
    if (ranges) std::memcpy(ranges, o.ranges.data(), o.ranges.size() * 4);
}

void vtgso_get_image(void* h, float* out_color, float* out_depth, float* final_T, uint32_t* n_contrib) {
    Oracle& o = *(Oracle*)h;
    if (out_color) std::memcpy(out_color, o.out_color.data(), o.out_color.size() * 4);
    if (out_depth) std::memcpy(out_depth, o.out_depth.data(), o.out_depth.size() * 4);
    if (final_T) std::memcpy(final_T, o.final_T.data(), o.final_T.size() * 4);
    if (n_contrib) std::memcpy(n_contrib, o.n_contrib.data(), o.n_contrib.size() * 4);
}

// dL_dcolors is [N, channels]; dL_dconic (optional) [N,3] exposes the intermediate.
void vtgso_backward(void* h, const float* dL_dout_color, float* dL_dmeans2D, float* dL_dcolors,
                    float* dL_dopacity, float* dL_dmeans3D, float* dL_dscales, float* dL_drotations,
                    float* dL_dconic) {
    run_backward(*(Oracle*)h, dL_dout_color, dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D,
                 dL_dscales, dL_drotations, dL_dconic);
}

float vtgso_expf(float x) { return vexpf(x); }

// Number of OpenMP threads of the oracle's loops (bench.py's CPU arm sets it to the cores it may use, whatever
// OMP_NUM_THREADS the launcher exported: torchrun sets it to 1).  Returns the value now in effect.
int vtgso_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

void vtgso_mark_visible(const VtgsCamera* cam, int64_t N, const float* means3D, uint8_t* present) {
    for (int64_t i = 0; i < N; ++i)
        present[i] = xform_row(cam->viewmatrix, 2, means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]) > VTGS_NEAR_CULL;
}

// -------------------------------------------------------------------------------------
// Front end of the fused path, restating the reference's pure-PyTorch builders with the
// spec'd elementary functions (so that integer outputs of the fused CUDA path can be
// compared bit-exactly):
//   transform_to_frame            utils/slam_helpers.py:323-385 (+ build_rotation,
//                                 utils/slam_external.py:25-42)
//   transformed_params2rendervar  utils/slam_helpers.py:127-160
//   get_depth_and_silhouette      utils/slam_helpers.py:217-234
// Outputs: means3D_cam[N,3], scales[N,3], rotations[N,4], opacities[N], colors6[N,6] =
// {r,g,b,z,1,z^2}.
// -------------------------------------------------------------------------------------
void vtgso_pose_matrix(const float* q_un, const float* t, float R[9]) {
    // F.normalize(q) then build_rotation's own re-normalisation
    const float n1 = sqrtf(fmaf(q_un[3], q_un[3], fmaf(q_un[2], q_un[2], fmaf(q_un[1], q_un[1], q_un[0] * q_un[0]))));
    const float d1 = fmaxf(n1, 1e-12f);
    float q[4] = {q_un[0] / d1, q_un[1] / d1, q_un[2] / d1, q_un[3] / d1};
    const float n2 = sqrtf(fmaf(q[3], q[3], fmaf(q[2], q[2], fmaf(q[1], q[1], q[0] * q[0]))));
    float qq[4] = {q[0] / n2, q[1] / n2, q[2] / n2, q[3] / n2};
    quat_to_R(qq, R);
    (void)t;
}

void vtgso_frontend(int64_t N, const float* means3D, const float* rgb, const float* unnorm_rot,
                    const float* logit_op, const float* log_scales, int32_t log_scales_dim,
                    const float* cam_q, const float* cam_t, const float* depth_row,
                    float* o_means, float* o_scales, float* o_rot, float* o_op, float* o_col6) {
    float R[9];
    vtgso_pose_matrix(cam_q, cam_t, R);
    // normalised camera quaternion (single F.normalize) for the anisotropic quat_mult
    const float n1 = sqrtf(fmaf(cam_q[3], cam_q[3], fmaf(cam_q[2], cam_q[2], fmaf(cam_q[1], cam_q[1], cam_q[0] * cam_q[0]))));
    const float d1 = fmaxf(n1, 1e-12f);
    const float cq[4] = {cam_q[0] / d1, cam_q[1] / d1, cam_q[2] / d1, cam_q[3] / d1};
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        const float x = means3D[3 * i], y = means3D[3 * i + 1], z = means3D[3 * i + 2];
        const float X = fmaf(R[2], z, fmaf(R[1], y, R[0] * x)) + cam_t[0];
        const float Y = fmaf(R[5], z, fmaf(R[4], y, R[3] * x)) + cam_t[1];
        const float Z = fmaf(R[8], z, fmaf(R[7], y, R[6] * x)) + cam_t[2];
        o_means[3 * i] = X; o_means[3 * i + 1] = Y; o_means[3 * i + 2] = Z;
        float q[4] = {unnorm_rot[4 * i], unnorm_rot[4 * i + 1], unnorm_rot[4 * i + 2], unnorm_rot[4 * i + 3]};
        if (log_scales_dim == 3) {
            const float n = sqrtf(fmaf(q[3], q[3], fmaf(q[2], q[2], fmaf(q[1], q[1], q[0] * q[0]))));
            const float d = fmaxf(n, 1e-12f);
            const float w2 = q[0] / d, x2 = q[1] / d, y2 = q[2] / d, z2 = q[3] / d;
            const float w1 = cq[0], x1 = cq[1], y1 = cq[2], z1 = cq[3];
            q[0] = w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2;
            q[1] = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2;
            q[2] = w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2;
            q[3] = w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2;
        }
        const float n = sqrtf(fmaf(q[3], q[3], fmaf(q[2], q[2], fmaf(q[1], q[1], q[0] * q[0]))));
        const float d = fmaxf(n, 1e-12f);
        for (int k = 0; k < 4; ++k) o_rot[4 * i + k] = q[k] / d;
        for (int k = 0; k < 3; ++k) {
            const float ls = log_scales[(size_t)i * log_scales_dim + (log_scales_dim == 3 ? k : 0)];
            o_scales[3 * i + k] = vexpf(ls);
        }
        o_op[i] = 1.0f / (1.0f + vexpf(-logit_op[i]));
        const float zc = fmaf(depth_row[2], Z, fmaf(depth_row[1], Y, depth_row[0] * X)) + depth_row[3];
        o_col6[6 * i + 0] = rgb[3 * i]; o_col6[6 * i + 1] = rgb[3 * i + 1]; o_col6[6 * i + 2] = rgb[3 * i + 2];
        o_col6[6 * i + 3] = zc; o_col6[6 * i + 4] = 1.0f; o_col6[6 * i + 5] = zc * zc;
    }
}

}  // extern "C"
