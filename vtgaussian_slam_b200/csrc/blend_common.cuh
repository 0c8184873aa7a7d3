// blend_common.cuh -- region lists and the lane-transposed test phase shared by the
// forward (K5') and backward (K6') blend kernels.
//
// Geometry of a tile block: 256 threads = 8 warps; warp w owns the 8x4-pixel region
// (w & 1, w >> 1) of the 16x16 tile and lane l the pixel (l & 7, l >> 3) of that region.
//
// View-tied Gaussians are ~1 px wide: a splat listed for a tile touches ~30% of the tile's
// eight regions and, inside a region it touches, ~25% of the 32 pixels.  Walking a tile
// list with one pixel per lane and the SAME splat in every lane (the upstream scheme) leaves
// 3/4 of the lanes idle in the expensive part (exp, blend / gradient terms).  The kernels here
// instead work on groups of 32 surviving splats of a region in three lane roles:
//   P1  lane = splat   : evaluate `power` for the 32 pixels (pixel coordinates are warp-uniform
//                        immediates; per-splat terms factor out) -> a 32x32 bit matrix
//                        "splat e may contribute to pixel p", transposed across the warp with a
//                        5-step butterfly so that every pixel lane gets its own splat mask;
//   P2  lane = pixel   : each lane walks ITS OWN mask in list order (different splats in
//                        different lanes), so the expensive part runs max-popcount times per
//                        group instead of once per splat;
//   P3  lane = splat   : (backward only) gather the per-(splat,pixel) terms P2 left in shared
//                        memory and reduce them per splat without any cross-lane shuffle.
// List positions (n_contrib), the alpha / T tests and the blend order per pixel are exactly the
// reference's, so every output is bit-identical to the one-splat-at-a-time formulation.
#pragma once
#include "common.cuh"

namespace vtgs {

// One group of 32 splats staged for P2 (pixel lanes pick arbitrary splats of the group).
struct GroupSmem {
    float4 a[32];          // {px, py, bits(1-based position in the tile list), opacity}
    float4 b[32];          // {A, B, C, pthr}
    float4 c[32];          // colours
};

// One entry of a region list, as registers of the lane that owns it in P1.
struct SplatRegs {
    float4 a, b, c;        // same packing as GroupSmem
};

// Region-list entry -> splat registers (gather of the 64-byte record; the 8 warps of a tile and the
// neighbouring tiles share these lines through L1 / L2).
__device__ __forceinline__ void load_splat(SplatRegs& r, bool valid, const GeomRecord* __restrict__ geom, uint2 ent) {
    if (valid) {
        const GeomRecord* rec = geom + ent.x;
        const float4 q0 = rec->q0, q1 = rec->q1;
        r.c = rec->q2;
        const float pthr = rec->q3.y;
        r.a = make_float4(q0.x, q0.y, __uint_as_float(ent.y), q1.w);
        r.b = make_float4(q1.x, q1.y, q1.z, pthr);
    }
}

// Shape of a warp's pixel region: REGION_W x REGION_H = 32 pixels, lane l = pixel (l % REGION_W, l / REGION_W).
// A 16x16 tile holds (16 / REGION_W) x (16 / REGION_H) = 8 regions, one per warp.
#ifndef VTGS_REGION_W
#define VTGS_REGION_W 8
#endif
constexpr int REGION_W = VTGS_REGION_W;
constexpr int REGION_H = 32 / REGION_W;
constexpr int REGIONS_X = 16 / REGION_W;
static_assert(REGION_W * REGION_H == 32 && 16 % REGION_W == 0 && 16 % REGION_H == 0, "region must tile 16x16 with 32 pixels");

// Which of the tile's 8 regions the box of a splat's alpha >= 1/255 ellipse touches (bit r = region r).
// tox, toy: pixel origin of the tile.
__device__ __forceinline__ uint32_t region_mask(const float4 q0, float tox, float toy) {
    const float x0 = q0.x - q0.z - tox, x1 = q0.x + q0.z - tox;
    const float y0 = q0.y - q0.w - toy, y1 = q0.y + q0.w - toy;
    uint32_t cm = 0, rmask = 0;
#pragma unroll
    for (int c = 0; c < REGIONS_X; ++c)
        if (x1 >= (float)(c * REGION_W) && x0 <= (float)(c * REGION_W + REGION_W - 1)) cm |= 1u << c;
#pragma unroll
    for (int r = 0; r < 16 / REGION_H; ++r)
        if (y1 >= (float)(r * REGION_H) && y0 <= (float)(r * REGION_H + REGION_H - 1)) rmask |= cm << (r * REGIONS_X);
    return rmask;
}

// Offset of a region's applied-contribution masks: one 32-bit word per (group of 32 splats, pixel lane).  A
// region list is shorter than its tile list (re - rb) but its last group is padded to 32, hence the +32.
__device__ __forceinline__ size_t mask_arena_base(uint32_t rb, uint32_t re, int tile, int warp) {
    return (size_t)8 * ((size_t)rb + (size_t)32 * tile) + (size_t)warp * ((size_t)(re - rb) + 32);
}

// 32x32 bit-matrix transpose across the warp: on entry lane e holds row e (bit p = column p),
// on return lane p holds column p (bit e = row e).
__device__ __forceinline__ uint32_t warp_transpose_bits(uint32_t x, int lane) {
#define VTGS_TSTEP(s, lo)                                                              \
    {                                                                                  \
        const uint32_t y = __shfl_xor_sync(VTGS_FULL_MASK, x, s);                      \
        x = (lane & s) ? ((x & ~(lo)) | ((y & ~(lo)) >> s)) : ((x & (lo)) | ((y & (lo)) << s)); \
    }
    VTGS_TSTEP(16, 0x0000FFFFu)
    VTGS_TSTEP(8, 0x00FF00FFu)
    VTGS_TSTEP(4, 0x0F0F0F0Fu)
    VTGS_TSTEP(2, 0x33333333u)
    VTGS_TSTEP(1, 0x55555555u)
#undef VTGS_TSTEP
    return x;
}

// P1: lane = splat (q0, q1 of this lane's splat; have == false for the tail of a partial group).
// emask: pixels of the region this splat may contribute to (power in [pthr, 0], in the spec'd
// arithmetic -- the same `power` P2 recomputes).  pmask (returned): for this lane AS A PIXEL, the
// splats of the group that may contribute to it.
__device__ __forceinline__ uint32_t p1_masks(bool have, const float4 q0, const float4 q1, float x0f, float y0f,
                                             int lane, uint32_t& emask) {
    const float pthr = q1.w;
    float dx[REGION_W], u[REGION_W], v[REGION_W], dy[REGION_H], wq[REGION_H];
#pragma unroll
    for (int c = 0; c < REGION_W; ++c) {
        dx[c] = fsub(q0.x, x0f + (float)c);
        u[c] = fmul(q1.x, dx[c]);
        v[c] = fmul(q1.y, dx[c]);
    }
#pragma unroll
    for (int r = 0; r < REGION_H; ++r) {
        dy[r] = fsub(q0.y, y0f + (float)r);
        wq[r] = fmul(fmul(q1.z, dy[r]), dy[r]);
    }
    uint32_t em = 0;
#pragma unroll
    for (int p = 0; p < 32; ++p) {
        const int c = p % REGION_W, r = p / REGION_W;
        const float q = ffma(u[c], dx[c], wq[r]);
        const float pw = ffma(-0.5f, q, -fmul(v[c], dy[r]));
        if (pw <= 0.0f && pw >= pthr) em |= 1u << p;
    }
    em = have ? em : 0u;
    emask = em;
    return warp_transpose_bits(em, lane);
}


// ---- chunked walk (forward K5' and backward K6') ---------------------------------------------------------------
// The splats of a region list that contribute to one pixel are spatially clustered along iso-depth lines, so inside
// ONE group of 32 list entries a few pixel lanes hold most of the work (measured on the C2 workload: the average lane
// has 6.4 contributions per group, the longest 17.4 -> 37 % useful lanes when the warp re-converges after every
// group).  Over several consecutive groups the per-pixel totals even out (62 % over 4 groups, 77 % over 8), so the
// kernels stage a CHUNK of GC groups in shared memory, run P1 for all of them, and then let every pixel lane walk
// its masks of the whole chunk without re-converging at group boundaries.
template <int GC>
struct ChunkSmem {
    float4 r0[GC * 32];      // {px, py, A, B}
    float4 r1[GC * 32];      // {C, opacity, c0, c1}
    float2 r2[GC * 32];      // {c2, c3}
    uint32_t pm[GC + 1][32]; // [group][pixel lane]: candidate mask from P1; after P2 the mask of APPLIED splats.
                             // The row after the chunk's last group holds a non-zero sentinel (ends the word search)
};

// P1 by rows, lane = splat: a conservative SUPERSET of the region's pixels whose spec'd `power` can be >= pthr.
// For a pixel row (dy fixed) power(dx) >= pt is the interval |dx + (B/A) dy| <= sqrt(-det dy^2 - 2 A pt) / A, so one
// square root per row yields the row's 8 mask bits (~22 instructions per row instead of ~6 per pixel).  P2 re-tests
// `power > 0` and `alpha < 1/255` exactly, so the margins only cost a few extra candidate pairs:
//   * pthr already sits 1e-3 below log(1/255 / opacity);
//   * 2e-6 * (A dx^2 + C dy^2 + 2 |B dx dy|) over the region bounds the rounding of the fp32 power and of `disc`
//     (needle-shaped splats far from their centre cancel large terms);
//   * 0.01 px covers the approximate reciprocal / square root.
// NaN / non-positive A degrade to "whole row" through fmaxf / fminf.
__device__ __forceinline__ uint32_t p1_rows(float sx, float sy, float A, float B, float C, float pthr, float x0f, float y0f) {
    const float ox = sx - x0f, oy = sy - y0f;
    const float dxm = fmaxf(fabsf(ox), fabsf(ox - (float)(REGION_W - 1)));
    const float dym = fmaxf(fabsf(oy), fabsf(oy - (float)(REGION_H - 1)));
    const float mag = fmaf(A * dxm, dxm, fmaf(C * dym, dym, 2.0f * fabsf(B) * dxm * dym));
    const float pt = fmaf(-2e-6f, mag, pthr - 1e-3f);
    const float iA = rcp_approx(A);
    const float BA = B * iA;
    const float det = fmaf(A, C, -B * B);
    const float k2 = -2.0f * A * pt;
    uint32_t em = 0u;
#pragma unroll
    for (int r = 0; r < REGION_H; ++r) {
        const float dy = oy - (float)r;
        const float disc = fmaf(-det * dy, dy, k2);
        const float h = sqrt_approx(fmaxf(disc, 0.0f)) * iA;
        const float ctr = fmaf(BA, dy, ox);
        const float lo = fmaxf(ctr - h - 0.01f, 0.0f), hi = fminf(ctr + h + 0.01f, (float)(REGION_W - 1));
        const int clo = __float2int_ru(lo), chi = __float2int_rd(hi);
        uint32_t row = (2u << chi) - (1u << clo);                  // bits clo..chi
        if (disc < 0.0f || chi < clo) row = 0u;
        em |= row << (r * REGION_W);
    }
    return em;
}

}  // namespace vtgs
