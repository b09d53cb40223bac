// Backward, TMA-staged persistent kernel with a tile-local gradient box: d(sum(color*g_color) + sum(depth*g_depth)) / d rgba.
//
// Planes are walked BACK TO FRONT with the forward's tile / ring / producer machinery (mpi_fwd_staged.cuh).  The forward, run
// in training mode, saved the transmittance T_i in front of every plane ([V,N,H,W]); with it one sweep suffices:
//     q_i = G.rgb_i + G_d*depth_i          d_i = q_i - R_i            (R_{N-1} = 0)
//     dL/d rgb_i = G * a_i * T_i           dL/d a_i = T_i * d_i       R_{i-1} = R_i + a_i * d_i
// which is autograd's  T_i q_i - (sum_{k>i} a_k q_k P_k) / (1 - a_i + 1e-10)  (cumprod_backward) without the division
// (1e-10 when a_i == 1) and without cancellation; grid_sampler_2d_backward then scatters every value through the four bilinear
// weights.
//
// Where the scatter goes (round 2).  Round 1 issued 16 red.global.add.f32 per (pixel, plane): 46 M REDG warp instructions per
// view, each wrapped in BSSY/BRA/BSYNC when predicated, 64-bit address arithmetic, shuffles for tap hand-over -- 900 SASS
// instructions per thread and plane, 28 % of the HBM roofline.  Measured on this part (profiles/r02_l2_reduce_probe.txt):
// shared-memory fp32 atomics are a CAS loop (5.8 cycles per warp instruction), native INTEGER shared atomics run at 1.0, and a
// dense, coalesced red.global of the finished sums is bound only by the DRAM read-modify-write of the gradient.  Hence:
//   * every (tile, plane) has a GRADIENT BOX in shared memory with exactly the layout of the staged plane box
//     ([row][channel][x], same index as the taps);
//   * consumers add their 16 contributions per pixel with `red.shared.add.s32` on FIXED-POINT values: c * 2^(26-e) rounded to
//     nearest (three FMAs against magic constants, no F2I), where 2^e bounds every contribution of the tile (computed from the
//     tile's upstream gradients, see tile_scale_exponent).  26 bits per contribution relative to that bound, exact (order
//     independent) integer sums: more repeatable than fp32 atomics, a few 1e-6 of the largest gradient in error;
//   * three FLUSHER warps convert a finished box to fp32 and add it to g_rgba with coalesced `red.global.add.v4.f32`
//     (skipping all-zero quads and texels outside the texture) and re-zero it, while the consumers fill the other box.
// Out-of-texture taps need no predication: they land in box cells that lie outside the texture, which the flush drops
// (padding_mode="zeros").  A warp whose footprints do not all lie inside the staged box (non-projective rays, oversized
// footprints, planes outside the exact-division range) samples global memory and scatters with red.global.add.f32 directly,
// as the direct kernel does: results never depend on the producer's footprint estimate.
#pragma once
#include "mpi_fwd_staged.cuh"

namespace gmpi {

// Build-time tunables, for A/B builds (profiles/README.md, round 2): 11 consumer + 4 flusher warps (64 x 22 tiles) is 1 % faster at
// 1024^2 and 27 % slower at 512^2; sweeping the box as one linear run of quads (all flusher lanes busy) is 3 % slower; waiting
// for the emptied gradient box only before the first scatter instead of before the sampling is 2.2 % faster.
#ifndef GMPI_BWD_CONS_WARPS
#define GMPI_BWD_CONS_WARPS 12
#endif
#ifndef GMPI_BWD_FLUSH_WARPS
#define GMPI_BWD_FLUSH_WARPS 3
#endif
#ifndef GMPI_BWD_LATE_WAIT
#define GMPI_BWD_LATE_WAIT 1
#endif
constexpr int kBwdConsWarps = GMPI_BWD_CONS_WARPS, kBwdFlushWarps = GMPI_BWD_FLUSH_WARPS;
constexpr int kBwdTileH = kPairs * kBwdConsWarps;                                   // 64 x 24 pixel tiles
constexpr int kBwdConsThreads = kBwdConsWarps * 32;
constexpr int kBwdThreads = (kBwdConsWarps + 1 + kBwdFlushWarps) * 32;              // 16 warps: 12 consumers, producer, 3 flushers
constexpr int kBwdStages = 2, kBwdBoxes = 2;
constexpr int kBwdMaxBH = (((kBwdTileH * 5) / 4 + 6 + kRowsPerOp - 1) / kRowsPerOp) * kRowsPerOp;
constexpr int kBwdPlaneFloats = kMaxBW * kBwdMaxBH * 4;
constexpr int kBwdTFloats = kTileW * kBwdTileH;
constexpr int kBwdStride = kBwdPlaneFloats + kBwdTFloats;
constexpr size_t kBwdSmem = (size_t)(kBwdStages * kBwdStride + kBwdBoxes * kBwdPlaneFloats) * 4 + (size_t)kMaxPlanesStaged * 32;
static_assert(kBwdSmem + 1024 <= 227 * 1024, "backward ring + gradient boxes + plane table must fit one SM");

struct BwdRing {
    static constexpr int kTileRows = kBwdTileH, kRingStages = kBwdStages, kBoxMaxH = kBwdMaxBH;
    static constexpr int kPlaneFloats = kBwdPlaneFloats, kStride = kBwdStride;
    static constexpr bool kReverse = true;       // planes back to front; each stage also carries the tile's saved transmittance
#ifndef GMPI_BWD_SLEEP
#define GMPI_BWD_SLEEP 1
#endif
    static constexpr bool kSleepPolls = GMPI_BWD_SLEEP != 0;
    static constexpr bool kWideFact = false;     // the backward keeps the 56..88-wide classes: its shared memory is full
    static constexpr bool kBinaryCopies = false; // expanded MPI: one 4-row copy per lane (see staged_producer)
    // factored MPI: colour box = 3 copies of 12 rows (row offset r * 3 * bw * 4 bytes must be a multiple of 128 for bw = 56..88,
    // i.e. r a multiple of 4; 18-row halves gave cudaErrorMisalignedAddress)
    static constexpr int kColourCopyRows = 12;
    static_assert(kBwdMaxBH % kColourCopyRows == 0 && kColourCopyRows % 4 == 0, "colour copies tile the box, 128-byte aligned");
};

struct GradPairs {
    f2 g0[kPairs], g1[kPairs], g2[kPairs], gs[kPairs];   // upstream colour gradient and g_depth * (ray . z_dir)
};

// what the flushers need to know about a finished gradient box
struct __align__(16) GradMeta {
    int bx0, by0;        // texel coordinates of box element [0][.][0]
    int rows, cls;       // staged rows (0: nothing in the box), width class (bw = kMinBW + cls * kBWStep)
    int plane;           // m * N + i
    float scale;         // alpha channel: 2^(e_a - kFixBits), fixed point -> fp32
    int mpi_bg;          // factored MPI: 2 * m + (1 if this plane's colour gradient goes to g_bg_rgb: the last plane)
    float scale_rgb;     // colour channels: 2^(e_rgb - kFixBitsRgb)
};
constexpr int kBwdAlphaOff = 3 * kMaxBW * kBwdMaxBH;     // factored MPI: alpha box behind the colour box (floats / ints)

// red.global.add.f32 without a return value
__device__ __forceinline__ void red_add(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Scatter one pixel's four channel gradients through its bilinear footprint straight into global memory
// (grid_sampler_2d_backward); the rare path of a warp whose footprints are not all inside the staged box.
// (The expanded instantiation passes one base pointer -- see sample_plane_direct for why -- the factored one four.)
__device__ __noinline__ void scatter_plane_global(float* __restrict__ gplane, size_t tex, int Wt, int Ht, int x0, int y0, float v0, float v1,
                                                  float v2, float v3, float w00, float w01, float w10, float w11) {
    const bool vx0 = (unsigned)x0 < (unsigned)Wt, vx1 = (unsigned)(x0 + 1) < (unsigned)Wt;
    const bool vy0 = (unsigned)y0 < (unsigned)Ht, vy1 = (unsigned)(y0 + 1) < (unsigned)Ht;
    float* b0 = gplane + ((long long)y0 * Wt + x0);
    const float v[4] = {v0, v1, v2, v3};
#pragma unroll
    for (int c = 0; c < 4; ++c, b0 += tex) {
        if (vx0 && vy0) red_add(b0, v[c] * w00);
        if (vx1 && vy0) red_add(b0 + 1, v[c] * w01);
        if (vx0 && vy1) red_add(b0 + Wt, v[c] * w10);
        if (vx1 && vy1) red_add(b0 + Wt + 1, v[c] * w11);
    }
}
__device__ __noinline__ void scatter_pixel_global(const GradChans gch, int Wt, int Ht, int x0, int y0, float v0, float v1,
                                                  float v2, float v3, float w00, float w01, float w10, float w11) {
    const bool vx0 = (unsigned)x0 < (unsigned)Wt, vx1 = (unsigned)(x0 + 1) < (unsigned)Wt;
    const bool vy0 = (unsigned)y0 < (unsigned)Ht, vy1 = (unsigned)(y0 + 1) < (unsigned)Ht;
    const long long o = (long long)y0 * Wt + x0;
    const float v[4] = {v0, v1, v2, v3};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float* b0 = gch.c[c] + o;
        if (vx0 && vy0) red_add(b0, v[c] * w00);
        if (vx1 && vy0) red_add(b0 + 1, v[c] * w01);
        if (vx0 && vy1) red_add(b0 + Wt, v[c] * w10);
        if (vx1 && vy1) red_add(b0 + Wt + 1, v[c] * w11);
    }
}

// Fixed point.  Every contribution c of a tile satisfies |c| <= 2 qmax, where
// qmax = max over the tile's pixels of |G_r| + |G_g| + |G_b| + |G_d (ray . z_dir)| * max_i |scale_i|:
//   |dL/d rgb contribution| = |G_c| a T w <= qmax;   |dL/d a contribution| = T |q - R| w <= 2 qmax   (rgb, a, T, w in [0,1];
//   R is a sub-convex combination of the q's).  With 2^e > 4 qmax (a factor 2 of slack for inputs that leave [0,1] by rounding)
// contributions are rounded to multiples of 2^(e - kFixBits): |c| 2^(kFixBits - e) < 2^(kFixBits - 1), i.e. 25 bits + sign per
// contribution (finer than the fp32 accumulation it replaces whenever the running sum is within 4x of the bound), and a texel
// may collect 2^(31 - kFixBits + 1) = 64 contributions of maximum size in int32 -- each pixel has one footprint per plane, so
// that takes a 8x8 minification... of the PIXEL grid onto one texel (scale < 1/4) at maximum gradient everywhere; wrap-around
// beyond that is the documented limit of this kernel (the direct kernel has none).
constexpr int kFixBits = 26, kFixSplit = 4;              // alpha: low kFixSplit bits come from the second conversion step
// The three colour channels have their own, tighter bound -- |dL/d rgb contribution| = |G_c| a T w <= gmax = max |G_c| over the
// tile, without the depth term and the factor 2 of the alpha bound -- and take the one-step conversion: 2^e_rgb > 2 gmax,
// contributions rounded to multiples of 2^(e_rgb - 22) (|c| 2^(22 - e_rgb) < 2^21: a factor 2 inside the 2^22 the magic constant
// allows, for inputs that leave [0,1] by rounding).
constexpr int kFixBitsRgb = 22;
constexpr float kMagicHi = 12582912.0f * 16.0f;          // 1.5 * 2^(23 + kFixSplit): ulp = 2^kFixSplit
constexpr int kMagicHiBits = 0x4b400000 + (kFixSplit << 23);
__device__ __forceinline__ int tile_scale_exponent(float qmax) {
    const float b = 4.0f * qmax;
    if (!(b > 0.0f)) return 0;
    int e = (int)((__float_as_uint(b) >> 23) & 0xffu) - 126;     // 2^e > b  (b = 1.m * 2^(E-127) < 2^(E-126))
    return max(-90, min(90, e));
}
// RN(x) for |x| < 2^(22 + kFixSplit), two pixels at once, as integers: the high part is read from the mantissa of x + 1.5 * 2^27
// (a multiple of 16), the exact remainder (|r| <= 8) from the mantissa of r + 1.5 * 2^23.  x = vF * w is never formed: both steps
// are FMAs on the exact product.
__device__ __forceinline__ void fix2(f2 vF, f2 w, int& ia, int& ib) {
    const f2 t1 = fma2(vF, w, splat(kMagicHi));
    const f2 nhi = fma2(t1, splat(-1.0f), splat(kMagicHi));      // -(high part), exact
    const f2 t2 = add2(fma2(vF, w, nhi), splat(kFloorMagic));    // remainder, rounded to an integer
    ia = ((__float_as_int(t1.x) - kMagicHiBits) << kFixSplit) + (__float_as_int(t2.x) - kFloorMagicBits);
    ib = ((__float_as_int(t1.y) - kMagicHiBits) << kFixSplit) + (__float_as_int(t2.y) - kFloorMagicBits);
}

// Fast body: four pixels (two packed pairs) from a staged box of compile-time width BW, contributions into the gradient box.
// Returns false (nothing sampled, R unchanged) if any footprint is not inside the box.
// One (pair, channel) of the scatter: the four bilinear contributions of two pixels into the gradient box.
template <int PITCH, bool kTwoStep>
__device__ __forceinline__ void scatter_channel(int* __restrict__ ga, int* __restrict__ gq, f2 vF, const f2 (&w)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int ca, cb;
        if (!kTwoStep) {     // RN(vF * w) read from the mantissa of one packed FMA
            const f2 t = fma2(vF, w[k], splat(kFloorMagic));
            ca = __float_as_int(t.x) - kFloorMagicBits; cb = __float_as_int(t.y) - kFloorMagicBits;
        } else {
            fix2(vF, w[k], ca, cb);
        }
        const int off = (k & 1) + (k >> 1) * PITCH;
        atomicAdd(ga + off, ca);
        atomicAdd(gq + off, cb);
    }
}

// Fast body: four pixels (two packed pairs) from a staged box of compile-time width BW: sample, form the gradients, add the 64
// fixed-point contributions to the gradient box `gb` (same layout as the staged box).  Returns false (nothing sampled, R
// unchanged, nothing added) if any footprint is not inside the box.
// AOFF as in sample_pairs: 0 = expanded stage, > 0 = factored (colour box [row][3][BW], alpha box [row][BW] at AOFF).
// Each pair is scattered as soon as it is formed: keeping both pairs' weights and values alive until a later scatter phase costs
// ~36 registers and spills.
template <int BW, int AOFF = 0>
__device__ __forceinline__ bool bwd_box_pairs(const float* __restrict__ sb, int* __restrict__ gb, int cx, int cy, int rows2, const CoordPairs& c,
                                              const f2 (&T)[kPairs], const GradPairs& G, f2 (&R)[kPairs], f2 Fs, f2 Fs_rgb,
                                              uint64_t* g_empty_bar, uint32_t g_empty_parity) {
    constexpr int RP = AOFF ? 3 * BW : 4 * BW, AP = AOFF ? BW : 4 * BW, A0 = AOFF ? AOFF : 3 * BW;
    const f2 m1 = splat(-1.0f), one = splat(1.0f);
    const f2 magic = splat(kFloorMagic), nmagic = splat(-kFloorMagic);
    f2 fx0[kPairs], fy0[kPairs];
    int idx[kPix], jdx[kPix];
    bool inbox = true;
#pragma unroll
    for (int P = 0; P < kPairs; ++P) {
        const f2 tx = add2_rm(c.ix[P], magic), ty = add2_rm(c.iy[P], magic);
        fx0[P] = add2(tx, nmagic);
        fy0[P] = add2(ty, nmagic);
        const int rxa = __float_as_int(tx.x) - cx, rxb = __float_as_int(tx.y) - cx;
        const int rya = __float_as_int(ty.x) - cy, ryb = __float_as_int(ty.y) - cy;
        inbox = inbox && (unsigned)rxa <= (unsigned)(BW - 2) && (unsigned)rxb <= (unsigned)(BW - 2) &&
                (unsigned)rya <= (unsigned)rows2 && (unsigned)ryb <= (unsigned)rows2;
        idx[2 * P] = rya * RP + rxa;
        idx[2 * P + 1] = ryb * RP + rxb;
        if (AOFF) { jdx[2 * P] = rya * AP + rxa + A0; jdx[2 * P + 1] = ryb * AP + rxb + A0; }     // separate alpha box
    }
    if (!__all_sync(0xffffffffu, inbox)) return false;   // warp-uniform
#pragma unroll
    for (int P = 0; P < kPairs; ++P) {
        const f2 wx1 = fma2(fx0[P], m1, c.ix[P]), wy1 = fma2(fy0[P], m1, c.iy[P]);
        const f2 wy0 = fma2(wy1, m1, one);
        f2 w[4];
        w[3] = mul2(wx1, wy1); w[2] = fma2(w[3], m1, wy1); w[1] = fma2(w[3], m1, wx1); w[0] = fma2(w[1], m1, wy0);
        const float* ta = sb + idx[2 * P];
        const float* tb = sb + idx[2 * P + 1];
#define GMPI_TAP(ch)                                                                                           \
    fma2(make_float2(ta[RP + ch * BW + 1], tb[RP + ch * BW + 1]), w[3],                                        \
         fma2(make_float2(ta[RP + ch * BW], tb[RP + ch * BW]), w[2],                                           \
              fma2(make_float2(ta[ch * BW + 1], tb[ch * BW + 1]), w[1], mul2(make_float2(ta[ch * BW], tb[ch * BW]), w[0]))))
        const f2 r = GMPI_TAP(0), g = GMPI_TAP(1), b = GMPI_TAP(2);
#undef GMPI_TAP
        const float* aa = AOFF ? sb + jdx[2 * P] : ta + A0;
        const float* ab = AOFF ? sb + jdx[2 * P + 1] : tb + A0;
        const f2 a = fma2(make_float2(aa[AP + 1], ab[AP + 1]), w[3],
                          fma2(make_float2(aa[AP], ab[AP]), w[2], fma2(make_float2(aa[1], ab[1]), w[1], mul2(make_float2(aa[0], ab[0]), w[0]))));
        const f2 q = fma2(G.g0[P], r, fma2(G.g1[P], g, fma2(G.g2[P], b, mul2(G.gs[P], c.sc[P]))));
        const f2 d = fma2(R[P], m1, q);                 // q - R
        const f2 wT = mul2(a, T[P]);
        R[P] = fma2(a, d, R[P]);
        // ---- scatter (grid_sampler_2d_backward), fixed point: colour channels one-step, alpha two-step ----
        // The gradient box is needed only from here on: the flushers get the first pair's sampling time to finish emptying it.
        if (GMPI_BWD_LATE_WAIT && P == 0) mbar_wait(g_empty_bar, g_empty_parity);
        const f2 wF = mul2(wT, Fs_rgb);                 // exact (power of two)
        int* ga = gb + idx[2 * P];
        int* gq = gb + idx[2 * P + 1];
        scatter_channel<RP, false>(ga, gq, mul2(G.g0[P], wF), w);
        scatter_channel<RP, false>(ga + BW, gq + BW, mul2(G.g1[P], wF), w);
        scatter_channel<RP, false>(ga + 2 * BW, gq + 2 * BW, mul2(G.g2[P], wF), w);
        scatter_channel<AP, true>(AOFF ? gb + jdx[2 * P] : ga + A0, AOFF ? gb + jdx[2 * P + 1] : gq + A0, mul2(mul2(T[P], d), Fs), w);   // dL/d alpha
    }
    return true;
}

// Flusher: rows fw, fw + 3, ... of a finished gradient box -> fp32 -> red.global.add.v4.f32, then zero.
template <int BW, bool FAC>
__device__ __forceinline__ void flush_box(int* __restrict__ gb, const GradMeta& gm, const RenderParams& p, size_t tex, int Ht, int Wt,
                                          int fw, int lane) {
    constexpr int Q = BW / 4;      // float4 quads per channel row
    // destination slabs of the four channels
    float* dst[4];
    if (FAC) {
        float* rgb = ((gm.mpi_bg & 1) ? p.g_bg_rgb : p.g_rgb) + (size_t)(gm.mpi_bg >> 1) * 3 * tex;
        dst[0] = rgb; dst[1] = rgb + tex; dst[2] = rgb + 2 * tex; dst[3] = p.g_alpha + (size_t)gm.plane * tex;
    } else {
        float* b = p.g_rgba + (size_t)gm.plane * 4 * tex;
        dst[0] = b; dst[1] = b + tex; dst[2] = b + 2 * tex; dst[3] = b + 3 * tex;
    }
    constexpr int kRowsPerWarp = (kBwdMaxBH + kBwdFlushWarps - 1) / kBwdFlushWarps;
    // BW quads per box row: channel ch, quad x4.  Expanded: [row][4][BW]; factored: [row][3][BW] + alpha box [row][BW].
    // All rows of this warp are loaded before any is processed: the flush is a latency chain otherwise (LDS -> test -> RED).
    for (int j = lane; j < BW; j += 32) {
        const int ch = j / Q, x4 = j - ch * Q;
        const int cell0 = FAC ? (ch < 3 ? j * 4 : kBwdAlphaOff + x4 * 4) : j * 4;          // int offset inside a row
        constexpr int kColPitch = FAC ? 3 * BW : 4 * BW;
        const int pitch = (FAC && ch == 3) ? BW : kColPitch;
        int4 a[kRowsPerWarp];
#pragma unroll
        for (int k = 0; k < kRowsPerWarp; ++k) {
            const int r = fw + k * kBwdFlushWarps;
            a[k] = *reinterpret_cast<const int4*>(gb + r * pitch + cell0);      // rows beyond gm.rows are never written: they read 0
        }
        const int tx = gm.bx0 + 4 * x4;
        const bool col_ok = (unsigned)tx < (unsigned)Wt;          // bx0 % 4 == 0 and Wt % 4 == 0: a quad is inside or outside as a whole
        float* d = (ch == 0 ? dst[0] : ch == 1 ? dst[1] : ch == 2 ? dst[2] : dst[3]) + tx;
        const float sc = ch < 3 ? gm.scale_rgb : gm.scale;
#pragma unroll
        for (int k = 0; k < kRowsPerWarp; ++k) {
            const int r = fw + k * kBwdFlushWarps;
            if ((a[k].x | a[k].y | a[k].z | a[k].w) == 0) continue;  // untouched halo (or a row beyond the box): nothing to add or clear
            *reinterpret_cast<int4*>(gb + r * pitch + cell0) = make_int4(0, 0, 0, 0);
            const int ty = gm.by0 + r;
            if (col_ok && (unsigned)ty < (unsigned)Ht)
                red_add_v4(d + (size_t)ty * Wt, (float)a[k].x * sc, (float)a[k].y * sc, (float)a[k].z * sc, (float)a[k].w * sc);
        }
    }
}

// ---- in-kernel zeroing of the gradient (opt-in: gmpi_debug_set_bwd_zero(1); GMPI_ZERO_GRAD defaults to stream memsets) -----------
// A memset of the gradient (16 B per texel-plane, 6.4 GB for the headline train step: 0.87 ms at the write peak) cannot overlap
// the persistent kernels -- they own every register of every SM, the memset kernel runs strictly before or after them
// (tools/zero_overlap_probe.py: forward + side-stream memset = the sum of the two).  This is the attempt to hide it in the
// backward, which is far from HBM-bound: the producer warp zeroes the gradient with plain 16-byte stores between the polls of
// its wait for a free ring stage, ONE MPI SLAB AHEAD of the tiles that add to it.  Every CTA owns 1/grid of each slab;
// zero_flags[m] counts the CTAs that are done with slab m.
// MEASURED (profiles/README.md, round 2): correct, deadlock-free, and 4.5 % SLOWER than the memsets (train step 6.56 vs 6.27 ms),
// whatever the pacing or the store's cache policy: 43 MB of stores per SM go through the same l1tex data path that the
// consumers' LDS/ATOMS keep 68 % busy, so the zeroing costs the kernel more than it costs a memset.  Kept as a tested option
// (a TMA bulk store from a zero page would bypass that path; the backward has no shared memory left for one).
// Protocol (any view order is safe, sorted-by-MPI views -- MPI.forward's layout -- are fast):
//   * before issuing the copies of a tile of MPI m the producer has zeroed, fenced and signalled its share of slabs 0..m;
//   * while it walks the tiles of MPI m it zeroes slab m + 1, eight stores at a time between the polls of its wait for a free
//     ring stage (at most zero_rate stores per stage on average); at the end of its walk, the rest;
//   * whoever adds to slab m with red.global (flushers; consumers on the rare generic path) first waits for zero_flags[m] ==
//     gridDim.x (one acquire load once the slab is known to be ready).
// No deadlock: a CTA only ever waits for flags of slabs <= its current tile's, every CTA signals those before it can block on its
// own ring, and all CTAs are co-resident (grid <= SMs, one CTA per SM).
struct GradZeroPacer {
    static constexpr bool kActive = true;
    float4* base;
    unsigned long long slab, off, end;      // float4 units: slab size; [off, end) = what is left of this CTA's share of slab `next`
    unsigned* flags;
    int M, rate, lane, next, allowed;       // next: first slab not yet signalled; slabs < allowed may be zeroed in the background
    int budget;                             // warp-wide stores this CTA may still issue in the current ring stage
    unsigned cta, n_cta;
    __device__ __forceinline__ void set_share() {
        if (next >= M) { off = end = 0; return; }
        const unsigned long long per = (slab + n_cta - 1) / n_cta;
        const unsigned long long lo = min(slab, per * cta), hi = min(slab, lo + per);
        off = (unsigned long long)next * slab + lo; end = (unsigned long long)next * slab + hi;
    }
    __device__ __forceinline__ void init(const RenderParams& p, int lane_) {
        base = p.zero_base; slab = p.zero_slab16; flags = p.zero_flags; M = flags ? p.M : 0; rate = p.zero_rate; lane = lane_;
        next = 0; allowed = 0; budget = 0; cta = blockIdx.x; n_cta = gridDim.x;
        set_share();
    }
    __device__ __forceinline__ void store8() {                     // eight warp-wide 512-byte stores
        float4* ptr = base + off + (unsigned)lane;
        const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (off + 256 <= end) {
#pragma unroll
            for (int k = 0; k < 8; ++k) ptr[32 * k] = z;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (off + 32 * k + (unsigned)lane < end) ptr[32 * k] = z;
        }
        off += 256;
    }
    __device__ __forceinline__ void signal() {
        __threadfence();
        __syncwarp();
        if (lane == 0) { __threadfence(); atomicAdd(flags + next, 1u); }
        ++next;
        set_share();
    }
    __device__ __forceinline__ void finish_through(int m) {
        while (next < M && next <= m) {
            while (off < end) store8();
            signal();
        }
    }
    __device__ __forceinline__ void before_tile(int m) { finish_through(m); allowed = max(allowed, m + 2); }
    __device__ __forceinline__ void new_stage() { budget = min(budget + rate, 8 * rate); }
    // between two polls of the producer's wait for a free stage: false = nothing to do (slab not yet allowed, or this stage's
    // budget is spent -- the stores must not crowd the copies out of the memory system, nor delay their issue by more than a chunk)
    __device__ __forceinline__ bool chunk() {
        if (budget <= 0 || next >= M || next >= allowed) return false;
        store8();
        budget -= 8;
        if (off >= end) signal();
        return true;
    }
    __device__ __forceinline__ void at_end() { finish_through(M - 1); }
};

// Called by every lane that is about to add to slab m with red.global: returns once all CTAs have zeroed their share of it.
__device__ __forceinline__ void wait_grad_zeroed(const RenderParams& p, int m) {
    if (!p.zero_flags) return;
    const unsigned* f = p.zero_flags + m;
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (v >= gridDim.x) return;
        __nanosleep(256);
    }
}

__device__ __forceinline__ void bwd_consumer_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kBwdConsThreads) : "memory"); }

template <bool kAlignCorners, bool kFactored>
__global__ void __launch_bounds__(kBwdThreads, 1)
mpi_bwd_box_kernel(const RenderParams p, const __grid_constant__ TmaMaps maps, const int tiles_x, const int tiles_y) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* s_buf = reinterpret_cast<float*>(smem_raw);                                          // rgba ring (+ transmittance boxes)
    int* s_grad = reinterpret_cast<int*>(smem_raw + (size_t)kBwdStages * kBwdStride * 4);       // gradient boxes
    PlaneConst* s_pc = reinterpret_cast<PlaneConst*>(smem_raw + (size_t)(kBwdStages * kBwdStride + kBwdBoxes * kBwdPlaneFloats) * 4);
    __shared__ StageMeta s_meta[kBwdStages];
    __shared__ GradMeta s_gmeta[kBwdBoxes];
    __shared__ __align__(8) uint64_t s_full[kBwdStages], s_empty[kBwdStages], g_full[kBwdBoxes], g_empty[kBwdBoxes];
    __shared__ TileWalk s_walk;
    __shared__ unsigned s_qmax[3], s_gmax[3];   // per-tile bounds (bits of non-negative floats), three slots in rotation
    __shared__ unsigned s_zmax;           // max_i |z_diff_i| of the current view's plane table

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        s_walk.init(tiles_x, p.H, p.V, (int)blockIdx.x, (int)gridDim.x, kBwdTileH, p.view_group);
        for (int s = 0; s < kBwdStages; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], kBwdConsWarps);
        }
        for (int s = 0; s < kBwdBoxes; ++s) {
            mbar_init(&g_full[s], kBwdConsWarps);
            mbar_init(&g_empty[s], kBwdFlushWarps);
        }
        s_qmax[0] = s_qmax[1] = s_qmax[2] = 0u;
        s_gmax[0] = s_gmax[1] = s_gmax[2] = 0u;
        s_zmax = 0u;
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < kBwdBoxes * kBwdPlaneFloats; i += kBwdThreads) s_grad[i] = 0;
    __syncthreads();

    const int Ht = p.Ht, Wt = p.Wt, N = p.N;
    const size_t tex = (size_t)Ht * Wt;

    if (warp == kBwdConsWarps) {
        // ================================ producer ================================
        if (lane == 0) tma_prefetch_desc(&maps.t);
        GradZeroPacer pacer;
        pacer.init(p, lane);
        staged_producer<kAlignCorners, BwdRing, kFactored>(p, maps, s_buf, s_meta, s_full, s_empty, &s_walk, lane, pacer);
    } else if (warp > kBwdConsWarps) {
        // ================================ flushers ================================
        const int fw = warp - kBwdConsWarps - 1;
        int f_box = 0;
        uint32_t f_phase = 0;
        int f_zeroed = -1;      // last slab known to be zeroed by all CTAs
        TileXY txy;
        for (int j = 0; s_walk.at(j, txy); ++j) {
            for (int ii = 0; ii < N; ++ii) {
                const int b = f_box;
                const uint32_t ph = f_phase;
                if (++f_box == kBwdBoxes) { f_box = 0; f_phase ^= 1u; }
                if (BwdRing::kSleepPolls) mbar_wait_sleep(&g_full[b], ph);
                else mbar_wait(&g_full[b], ph);
                const GradMeta gm = s_gmeta[b];
                int* gb = s_grad + b * kBwdPlaneFloats;
                if (gm.rows > 0) {
                    if ((gm.mpi_bg >> 1) != f_zeroed) { f_zeroed = gm.mpi_bg >> 1; wait_grad_zeroed(p, f_zeroed); }
                    switch (gm.cls) {     // warp-uniform
                        case 0: flush_box<56, kFactored>(gb, gm, p, tex, Ht, Wt, fw, lane); break;
                        case 1: flush_box<64, kFactored>(gb, gm, p, tex, Ht, Wt, fw, lane); break;
                        case 2: flush_box<72, kFactored>(gb, gm, p, tex, Ht, Wt, fw, lane); break;
                        case 3: flush_box<80, kFactored>(gb, gm, p, tex, Ht, Wt, fw, lane); break;
                        default: flush_box<88, kFactored>(gb, gm, p, tex, Ht, Wt, fw, lane); break;
                    }
                }
                __syncwarp();
                mbar_arrive_if(&g_empty[b], lane == 0);
            }
        }
    } else {
        // ================================ consumers ================================
        const float fWt = (float)Wt, fHt = (float)Ht;
        float hsx = 0.5f * (float)(Wt - 1), hsy = 0.5f * (float)(Ht - 1);
        int lane_ = lane;
        asm volatile("" : "+f"(hsx), "+f"(hsy), "+r"(lane_));     // opaque: not re-derived from the parameters in every plane iteration
        const size_t img = (size_t)p.H * p.W;
        int c_stage = 0, g_box = 0;
        uint32_t c_phase = 0, g_phase = 0;
        const float gscale = (p.options & GMPI_COLOR_MINUS1_1) ? 2.0f : 1.0f;   // upstream gradient is w.r.t. 2*color-1
        int v_table = -1;
        TileXY txy;
        for (int j = 0; s_walk.at(j, txy); ++j) {
            const int v = txy.v, px0 = txy.px0, py0 = txy.py0;
            const int m = __ldg(p.view2mpi + v);
            const float* e = p.eye + 3 * v;
            const float ev[3] = {__ldg(e), __ldg(e + 1), __ldg(e + 2)};
            const float zd[3] = {__ldg(p.z_dir + 3 * v), __ldg(p.z_dir + 3 * v + 1), __ldg(p.z_dir + 3 * v + 2)};
            if (v != v_table) {          // (view, plane) constants, once per view and CTA
                bwd_consumer_bar_sync();
                if (threadIdx.x == 0) s_zmax = 0u;
                bwd_consumer_bar_sync();
                float zm = 0.0f;
                for (int i = threadIdx.x; i < N; i += kBwdConsThreads) {
                    const PlaneConst pc = make_plane_const(p.dhw + ((size_t)m * N + i) * 3, ev[2]);
                    s_pc[i] = pc;
                    zm = fmaxf(zm, fabsf(pc.z_diff));
                }
                atomicMax(&s_zmax, __float_as_uint(zm));
                bwd_consumer_bar_sync();
                v_table = v;
            }
            // ---- per-pixel inputs (clamped into the image; the tile overhang carries zero upstream gradient) ----
            const float* rays = p.ray_dir + (size_t)v * 3 * img;
            RayConst rc[kPix];
            RayPairs rp;
            GradPairs G;
            float gq[kPix][4];
            bool rays_fast = (in_safe_range(ev[0]) || ev[0] == 0.0f) && (in_safe_range(ev[1]) || ev[1] == 0.0f);
            const float zmax = __uint_as_float(s_zmax);
            float qmax = 0.0f, gmax = 0.0f;
#pragma unroll
            for (int q = 0; q < kPix; ++q) {
                const int pxq = px0 + lane + 32 * (q & 1), pyq = py0 + kPairs * warp + (q >> 1);
                const bool valid = pxq < p.W && pyq < p.H;
                const size_t pix = (size_t)min(pyq, p.H - 1) * p.W + min(pxq, p.W - 1);
                const float* rd = rays + pix;
                rc[q] = make_ray_const(__ldg(rd), __ldg(rd + img), __ldg(rd + 2 * img), ev, zd);
                rays_fast = rays_fast && rc[q].fast && fabsf(rc[q].rx2) <= 0x1p40f && fabsf(rc[q].ry2) <= 0x1p40f;
                const float* gc = p.g_color + (size_t)v * 3 * img + pix;
                gq[q][0] = valid ? gscale * __ldg(gc) : 0.0f;
                gq[q][1] = valid ? gscale * __ldg(gc + img) : 0.0f;
                gq[q][2] = valid ? gscale * __ldg(gc + 2 * img) : 0.0f;
                gq[q][3] = (valid && p.g_depth) ? __ldg(p.g_depth + (size_t)v * img + pix) * rc[q].dz : 0.0f;
                qmax = fmaxf(qmax, fabsf(gq[q][0]) + fabsf(gq[q][1]) + fabsf(gq[q][2]) + fabsf(gq[q][3]) * (zmax * fabsf(rc[q].yrz)));
                gmax = fmaxf(gmax, fmaxf(fabsf(gq[q][0]), fmaxf(fabsf(gq[q][1]), fabsf(gq[q][2]))));
            }
            // ---- the tile's fixed-point scale: max over all consumer threads (three slots in rotation, see below) ----
            {
                const int slot = j % 3;
                for (int o = 16; o > 0; o >>= 1) {
                    qmax = fmaxf(qmax, __shfl_xor_sync(0xffffffffu, qmax, o));
                    gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
                }
                if (!(qmax < 0x1p100f)) qmax = 0x1p100f;             // inf/NaN upstream gradients: keep the exponents finite
                if (!(gmax < 0x1p100f)) gmax = 0x1p100f;
                if (lane == 0) { atomicMax(&s_qmax[slot], __float_as_uint(qmax)); atomicMax(&s_gmax[slot], __float_as_uint(gmax)); }
                if (threadIdx.x == 0) s_qmax[(j + 1) % 3] = s_gmax[(j + 1) % 3] = 0u;   // next tile's slots: their last readers passed the previous tile's barrier
                bwd_consumer_bar_sync();
                qmax = __uint_as_float(s_qmax[slot]);
                gmax = __uint_as_float(s_gmax[slot]);
            }
            const int e_fix = tile_scale_exponent(qmax);
            const f2 Fs = splat(__uint_as_float((unsigned)(127 + kFixBits - e_fix) << 23));      // 2^(kFixBits - e)
            const float inv_scale = __uint_as_float((unsigned)(127 - kFixBits + e_fix) << 23);    // 2^(e - kFixBits)
            const int e_rgb = tile_scale_exponent(0.5f * gmax);                                   // 2^e_rgb > 2 gmax
            const f2 Fs_rgb = splat(__uint_as_float((unsigned)(127 + kFixBitsRgb - e_rgb) << 23));
            const float inv_scale_rgb = __uint_as_float((unsigned)(127 - kFixBitsRgb + e_rgb) << 23);

            const bool idle = py0 + kPairs * warp >= p.H;      // warp-uniform: no row of this warp is inside the image
#pragma unroll
            for (int P = 0; P < kPairs; ++P) {
                rp.rx2[P] = make_float2(rc[2 * P].rx2, rc[2 * P + 1].rx2);
                rp.ry2[P] = make_float2(rc[2 * P].ry2, rc[2 * P + 1].ry2);
                rp.nrz[P] = make_float2(-rc[2 * P].rz, -rc[2 * P + 1].rz);
                rp.yrz[P] = make_float2(rc[2 * P].yrz, rc[2 * P + 1].yrz);
                G.g0[P] = make_float2(gq[2 * P][0], gq[2 * P + 1][0]);
                G.g1[P] = make_float2(gq[2 * P][1], gq[2 * P + 1][1]);
                G.g2[P] = make_float2(gq[2 * P][2], gq[2 * P + 1][2]);
                G.gs[P] = make_float2(gq[2 * P][3], gq[2 * P + 1][3]);
            }
            const f2 ex2 = splat(rc[0].ex2), ey2 = splat(rc[0].ey2), hsx2 = splat(hsx), hsy2 = splat(hsy);
            const bool warp_fast = __all_sync(0xffffffffu, rays_fast) && !idle;
            f2 R[kPairs];
#pragma unroll
            for (int P = 0; P < kPairs; ++P) R[P] = splat(0.0f);
            for (int ii = 0; ii < N; ++ii) {
                const int i = N - 1 - ii;
                const int s = c_stage, b = g_box;
                const uint32_t ph = c_phase, gph = g_phase;
                if (++c_stage == kBwdStages) { c_stage = 0; c_phase ^= 1u; }
                if (++g_box == kBwdBoxes) { g_box = 0; g_phase ^= 1u; }
                const PlaneConst pcc = s_pc[i];
                CoordPairs cc;
                if (warp_fast) coords_pairs<kAlignCorners>(pcc, rp, ex2, ey2, hsx2, hsy2, fWt, fHt, cc);
                mbar_wait(&s_full[s], ph);
                const StageMeta mt = s_meta[s];
                const float* sb = s_buf + s * kBwdStride;
                const int sel = mt.sel;
                f2 T[kPairs];      // transmittance saved by the forward, staged next to the plane tile: [kBwdTileH][kTileW]
#pragma unroll
                for (int P = 0; P < kPairs; ++P) {
                    const float* tr = sb + kBwdPlaneFloats + (kPairs * warp + P) * kTileW + lane;
                    T[P] = make_float2(tr[0], tr[32]);
                }
                // ---- gradient box of this (tile, plane): sample, wait until the flushers have emptied it, scatter ----
                if (!GMPI_BWD_LATE_WAIT) mbar_wait(&g_empty[b], gph ^ 1u);
                int* gb = s_grad + b * kBwdPlaneFloats;
                constexpr int AO = kFactored ? kBwdAlphaOff : 0;
                bool done = false;
                if (warp_fast) {   // warp-uniform, one-hot class
                    if (sel & (1 << 18)) done = bwd_box_pairs<72, AO>(sb, gb, mt.cx, mt.cy, mt.rows2, cc, T, G, R, Fs, Fs_rgb, &g_empty[b], gph ^ 1u);
                    else if (sel & (1 << 17)) done = bwd_box_pairs<64, AO>(sb, gb, mt.cx, mt.cy, mt.rows2, cc, T, G, R, Fs, Fs_rgb, &g_empty[b], gph ^ 1u);
                    else if (sel & (1 << 19)) done = bwd_box_pairs<80, AO>(sb, gb, mt.cx, mt.cy, mt.rows2, cc, T, G, R, Fs, Fs_rgb, &g_empty[b], gph ^ 1u);
                    else if (sel & (1 << 16)) done = bwd_box_pairs<56, AO>(sb, gb, mt.cx, mt.cy, mt.rows2, cc, T, G, R, Fs, Fs_rgb, &g_empty[b], gph ^ 1u);
                    else if (sel & (1 << 20)) done = bwd_box_pairs<88, AO>(sb, gb, mt.cx, mt.cy, mt.rows2, cc, T, G, R, Fs, Fs_rgb, &g_empty[b], gph ^ 1u);
                }
                const int cls = done ? 0 : -1;
                // a warp that did not use the box still waits: it must not arrive on g_full[b] a second time before the flushers
                // have taken the previous fill (the arrival counts of two fills would mix)
                if (GMPI_BWD_LATE_WAIT && !done) mbar_wait(&g_empty[b], gph ^ 1u);
                __syncwarp();
                mbar_arrive_if(&s_empty[s], lane_ == 0);    // taps and transmittance are consumed: hand the stage back
                if (warp == 0 && lane == 0) {               // (warp 0 always has a row inside the image)
                    GradMeta gm;
                    const int mode = (sel >> 8) & 3;
                    gm.bx0 = mt.cx - kFloorMagicBits; gm.by0 = mt.cy - kFloorMagicBits;
                    gm.rows = mode == 0 ? mt.rows2 + 2 : 0;
                    gm.cls = ((sel & 0xff) - kMinBW) / kBWStep;
                    gm.plane = m * N + i;
                    gm.scale = inv_scale;
                    gm.scale_rgb = inv_scale_rgb;
                    gm.mpi_bg = 2 * m + ((kFactored && p.g_bg_rgb && i == N - 1) ? 1 : 0);
                    s_gmeta[b] = gm;
                }
                __syncwarp();
                mbar_arrive_if(&g_full[b], lane_ == 0);
                if (cls < 0 && !idle) {
                    // ---- generic body (rare): per-pixel checks, sampling and scattering straight in global memory.  Also
                    // taken when the producer's corner-ray estimate says "nothing under the tile": a hint, never trusted ----
                    const float* plane = kFactored ? nullptr : p.rgba + ((size_t)m * N + i) * 4 * tex;
                    float* gplane = kFactored ? nullptr : p.g_rgba + ((size_t)m * N + i) * 4 * tex;
                    wait_grad_zeroed(p, m);       // (in-kernel GMPI_ZERO_GRAD: slab m must be zero before anything is added to it)
                    float* Rs = reinterpret_cast<float*>(R);
                    const float* Ts = reinterpret_cast<const float*>(T);
#pragma unroll
                    for (int q = 0; q < kPix; ++q) {
                        const int qq = (q & 1) + 2 * (q >> 1);      // R / T are stored as pairs: element (P, half) = 2 P + half
                        RayConst rg = rc[q];
                        rg.fast = false;
                        const TexCoord tc = plane_coord<kAlignCorners>(pcc, rg, hsx, hsy, fWt, fHt);
                        if (!coord_hits(tc.ix, tc.iy, fWt, fHt)) continue;
                        const float4 sv = sample_plane_any<kFactored>(p, plane, m, i, tex, tc.ix, tc.iy);
                        const float fx = floorf(tc.ix), fy = floorf(tc.iy);
                        const float wx1 = tc.ix - fx, wy1 = tc.iy - fy, wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
                        const float qv = fmaf(gq[q][0], sv.x, fmaf(gq[q][1], sv.y, fmaf(gq[q][2], sv.z, gq[q][3] * tc.scale)));
                        const float d = qv - Rs[qq];
                        const float w = sv.w * Ts[qq];
                        Rs[qq] = fmaf(sv.w, d, Rs[qq]);
                        if (kFactored)
                            scatter_pixel_global(grad_chans(p, m, i, tex), Wt, Ht, (int)fx, (int)fy, gq[q][0] * w, gq[q][1] * w, gq[q][2] * w,
                                                 Ts[qq] * d, wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1);
                        else
                            scatter_plane_global(gplane, tex, Wt, Ht, (int)fx, (int)fy, gq[q][0] * w, gq[q][1] * w, gq[q][2] * w, Ts[qq] * d,
                                                 wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1);
                    }
                }
            }
        }
    }
}

}  // namespace gmpi
