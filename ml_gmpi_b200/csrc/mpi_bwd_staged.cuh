// Backward, TMA-staged persistent variant: d(sum(color*g_color) + sum(depth*g_depth)) / d rgba.
//
// Same tile / ring / producer machinery as the forward (mpi_fwd_staged.cuh), planes walked BACK TO FRONT.  The forward,
// run in training mode, saved the transmittance T_i in front of every plane ([V,N,H,W]); with it one sweep suffices:
//     q_i = G.rgb_i + G_d*depth_i          d_i = q_i - R_i            (R_{N-1} = 0)
//     dL/d rgb_i = G * a_i * T_i           dL/d a_i = T_i * d_i       R_{i-1} = R_i + a_i * d_i
// which is autograd's  T_i q_i - (sum_{k>i} a_k q_k P_k) / (1 - a_i + 1e-10)  (cumprod_backward) without the division
// (1e-10 when a_i == 1) and without cancellation.  The four channels are sampled from the staged box (packed f32x2 math,
// as in the forward) and every value is scattered through the four bilinear weights with red.global.add.f32
// (grid_sampler_2d_backward); taps outside the texture are skipped (padding_mode="zeros").
#pragma once
#include "mpi_fwd_staged.cuh"

namespace gmpi {

struct GradPairs {
    f2 g0[kPairs], g1[kPairs], g2[kPairs], gs[kPairs];   // upstream colour gradient and g_depth * (ray . z_dir)
};

// red.global.add.f32 without a return value; the predicated form compiles to one predicated REDG (no branch).
__device__ __forceinline__ void red_add(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add_if(float* p, float v, bool ok) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q red.global.add.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"((int)ok) : "memory");
}

// Scatter one pixel's four channel gradients through its bilinear footprint with north-west texel (x0, y0)
// (grid_sampler_2d_backward).  kAllValid: the caller proved (warp-uniformly) that all four taps are inside the texture.
template <bool kAllValid>
__device__ __forceinline__ void scatter_pixel(float* __restrict__ gplane, size_t tex, int Wt, int Ht, int x0, int y0, const float (&v)[4],
                                              float w00, float w01, float w10, float w11) {
    float* b0 = gplane + ((long long)y0 * Wt + x0);
    float* b1 = b0 + Wt;
    if (kAllValid) {
#pragma unroll
        for (int c = 0; c < 4; ++c, b0 += tex, b1 += tex) {
            red_add(b0, v[c] * w00); red_add(b0 + 1, v[c] * w01);
            red_add(b1, v[c] * w10); red_add(b1 + 1, v[c] * w11);
        }
    } else {
        const bool vx0 = (unsigned)x0 < (unsigned)Wt, vx1 = (unsigned)(x0 + 1) < (unsigned)Wt;
        const bool vy0 = (unsigned)y0 < (unsigned)Ht, vy1 = (unsigned)(y0 + 1) < (unsigned)Ht;
#pragma unroll
        for (int c = 0; c < 4; ++c, b0 += tex, b1 += tex) {
            red_add_if(b0, v[c] * w00, vx0 && vy0); red_add_if(b0 + 1, v[c] * w01, vx1 && vy0);
            red_add_if(b1, v[c] * w10, vx0 && vy1); red_add_if(b1 + 1, v[c] * w11, vx1 && vy1);
        }
    }
}

// Fast body: four pixels (two packed pairs) from a staged box of compile-time width BW.  Returns false (nothing done) if
// any footprint is not inside the box.
template <int BW>
__device__ __forceinline__ bool bwd_pairs(const float* __restrict__ sb, int cx, int cy, int rows2, const CoordPairs& c,
                                          const f2 (&T)[kPairs], const GradPairs& G, f2 (&R)[kPairs], float* __restrict__ gplane,
                                          size_t tex, int Wt, int Ht) {
    const f2 m1 = splat(-1.0f), one = splat(1.0f);
    const f2 magic = splat(kFloorMagic), nmagic = splat(-kFloorMagic);
    const int bx0 = cx - kFloorMagicBits, by0 = cy - kFloorMagicBits;
    f2 fx0[kPairs], fy0[kPairs];
    int rxa[kPairs], rxb[kPairs], rya[kPairs], ryb[kPairs];
    bool inbox = true;
#pragma unroll
    for (int P = 0; P < kPairs; ++P) {
        const f2 tx = add2_rm(c.ix[P], magic), ty = add2_rm(c.iy[P], magic);
        fx0[P] = add2(tx, nmagic);
        fy0[P] = add2(ty, nmagic);
        rxa[P] = __float_as_int(tx.x) - cx; rxb[P] = __float_as_int(tx.y) - cx;
        rya[P] = __float_as_int(ty.x) - cy; ryb[P] = __float_as_int(ty.y) - cy;
        inbox = inbox && (unsigned)rxa[P] <= (unsigned)(BW - 2) && (unsigned)rxb[P] <= (unsigned)(BW - 2) &&
                (unsigned)rya[P] <= (unsigned)rows2 && (unsigned)ryb[P] <= (unsigned)rows2;
    }
    if (!__all_sync(0xffffffffu, inbox)) return false;   // warp-uniform: the body below uses full-mask shuffles
#pragma unroll
    for (int P = 0; P < kPairs; ++P) {
        const f2 wx1 = fma2(fx0[P], m1, c.ix[P]), wy1 = fma2(fy0[P], m1, c.iy[P]);
        const f2 wy0 = fma2(wy1, m1, one);
        const f2 w11 = mul2(wx1, wy1), w10 = fma2(w11, m1, wy1), w01 = fma2(w11, m1, wx1), w00 = fma2(w01, m1, wy0);
        const float* ta = sb + (rya[P] * (4 * BW) + rxa[P]);
        const float* tb = sb + (ryb[P] * (4 * BW) + rxb[P]);
#define GMPI_TAP(ch)                                                                                           \
    fma2(make_float2(ta[(4 + ch) * BW + 1], tb[(4 + ch) * BW + 1]), w11,                                       \
         fma2(make_float2(ta[(4 + ch) * BW], tb[(4 + ch) * BW]), w10,                                          \
              fma2(make_float2(ta[ch * BW + 1], tb[ch * BW + 1]), w01, mul2(make_float2(ta[ch * BW], tb[ch * BW]), w00))))
        const f2 r = GMPI_TAP(0), g = GMPI_TAP(1), b = GMPI_TAP(2), a = GMPI_TAP(3);
#undef GMPI_TAP
        const f2 q = fma2(G.g0[P], r, fma2(G.g1[P], g, fma2(G.g2[P], b, mul2(G.gs[P], c.sc[P]))));
        const f2 d = fma2(R[P], m1, q);                 // q - R
        const f2 w = mul2(a, T[P]);
        const f2 ga = mul2(T[P], d);
        R[P] = fma2(a, d, R[P]);
        const f2 gr = mul2(G.g0[P], w), gg = mul2(G.g1[P], w), gb = mul2(G.g2[P], w);
        const int xa = bx0 + rxa[P], ya = by0 + rya[P], xb = bx0 + rxb[P], yb = by0 + ryb[P];
        // interior footprints (the common case) need no per-tap range checks: decide once per warp
        const bool ok = (unsigned)xa < (unsigned)(Wt - 1) && (unsigned)ya < (unsigned)(Ht - 1) &&
                        (unsigned)xb < (unsigned)(Wt - 1) && (unsigned)yb < (unsigned)(Ht - 1);
        if (__all_sync(0xffffffffu, ok)) {
            // Tap combining: lanes hold consecutive pixels of one image row, so a lane's EAST texels (x0+1) are its right
            // neighbour's WEST texels whenever x0 advances by exactly one on the same texel row (the usual case at scale
            // ~1).  Hand those contributions over with a shuffle and drop the east atomics: 16 -> ~9 reds per pixel-plane.
            const unsigned full = 0xffffffffu;
            const int lane_id = threadIdx.x & 31;
            // (all shuffles are executed by all 32 lanes: no short-circuit around a *.sync)
            const int nxa = __shfl_down_sync(full, xa, 1), nya = __shfl_down_sync(full, ya, 1);
            const int nxb = __shfl_down_sync(full, xb, 1), nyb = __shfl_down_sync(full, yb, 1);
            const bool me_a = lane_id < 31 && nxa == xa + 1 && nya == ya;
            const bool me_b = lane_id < 31 && nxb == xb + 1 && nyb == yb;
            const int pma = __shfl_up_sync(full, (int)me_a, 1), pmb = __shfl_up_sync(full, (int)me_b, 1);
            const bool mw_a = lane_id > 0 && pma != 0;
            const bool mw_b = lane_id > 0 && pmb != 0;
            const f2 vals[4] = {gr, gg, gb, ga};
            const f2 takes = make_float2(mw_a ? 1.0f : 0.0f, mw_b ? 1.0f : 0.0f);
            float* a0 = gplane + ((long long)ya * Wt + xa);
            float* b0 = gplane + ((long long)yb * Wt + xb);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch, a0 += tex, b0 += tex) {
                f2 cw0 = mul2(vals[ch], w00), ce0 = mul2(vals[ch], w01), cw1 = mul2(vals[ch], w10), ce1 = mul2(vals[ch], w11);
                const float ra0 = __shfl_up_sync(full, ce0.x, 1), ra1 = __shfl_up_sync(full, ce1.x, 1);
                const float rb0 = __shfl_up_sync(full, ce0.y, 1), rb1 = __shfl_up_sync(full, ce1.y, 1);
                cw0 = fma2(make_float2(ra0, rb0), takes, cw0);        // branch-free: takes = 1 where the left neighbour hands over
                cw1 = fma2(make_float2(ra1, rb1), takes, cw1);
                red_add(a0, cw0.x); red_add(a0 + Wt, cw1.x);
                red_add_if(a0 + 1, ce0.x, !me_a); red_add_if(a0 + Wt + 1, ce1.x, !me_a);
                red_add(b0, cw0.y); red_add(b0 + Wt, cw1.y);
                red_add_if(b0 + 1, ce0.y, !me_b); red_add_if(b0 + Wt + 1, ce1.y, !me_b);
            }
        } else {
            const float va[4] = {gr.x, gg.x, gb.x, ga.x}, vb[4] = {gr.y, gg.y, gb.y, ga.y};
            scatter_pixel<false>(gplane, tex, Wt, Ht, xa, ya, va, w00.x, w01.x, w10.x, w11.x);
            scatter_pixel<false>(gplane, tex, Wt, Ht, xb, yb, vb, w00.y, w01.y, w10.y, w11.y);
        }
    }
    return true;
}

template <bool kAlignCorners>
__global__ void __launch_bounds__(kStagedThreads, 1)
mpi_bwd_staged_kernel(const RenderParams p, const __grid_constant__ TmaMaps maps, const int tiles_x, const int tiles_y) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* s_buf = reinterpret_cast<float*>(smem_raw);
    PlaneConst* s_pc = reinterpret_cast<PlaneConst*>(smem_raw + (size_t)kStages * kStageFloatsBwd * 4);
    __shared__ StageMeta s_meta[kStages];
    __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages];
    __shared__ TileWalk s_walk;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        s_walk.init(tiles_x, p.H, p.V, (int)blockIdx.x, (int)gridDim.x);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], kConsWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int Ht = p.Ht, Wt = p.Wt, N = p.N;
    const float fWt = (float)Wt, fHt = (float)Ht;
    const float hsx = 0.5f * (float)(Wt - 1), hsy = 0.5f * (float)(Ht - 1);
    const size_t img = (size_t)p.H * p.W;

    if (warp == kConsWarps) {
        if (lane == 0) tma_prefetch_desc(&maps.t);
        staged_producer<kAlignCorners, true>(p, maps, s_buf, s_meta, s_full, s_empty, &s_walk, lane);
    } else {
        int c_stage = 0;
        uint32_t c_phase = 0;
        const size_t tex = (size_t)Ht * Wt;
        const float gscale = (p.options & GMPI_COLOR_MINUS1_1) ? 2.0f : 1.0f;   // upstream gradient is w.r.t. 2*color-1
        int v_table = -1;
        TileXY txy;
        for (int j = 0; s_walk.at(j, txy); ++j) {
            const int v = txy.v, px0 = txy.px0, py0 = txy.py0;
            const int m = __ldg(p.view2mpi + v);
            const float* e = p.eye + 3 * v;
            const float ev[3] = {__ldg(e), __ldg(e + 1), __ldg(e + 2)};
            const float zd[3] = {__ldg(p.z_dir + 3 * v), __ldg(p.z_dir + 3 * v + 1), __ldg(p.z_dir + 3 * v + 2)};
            if (v != v_table) {
                consumer_bar_sync();
                for (int i = threadIdx.x; i < N; i += kConsThreads) s_pc[i] = make_plane_const(p.dhw + ((size_t)m * N + i) * 3, ev[2]);
                consumer_bar_sync();
                v_table = v;
            }
            if (py0 + kPairs * warp >= p.H) {      // no row of this warp is inside the image (partial bottom tile)
                consumer_idle_tile(s_full, s_empty, N, lane, c_stage, c_phase);
                continue;
            }
            const float* rays = p.ray_dir + (size_t)v * 3 * img;
            RayConst rc[kPix];
            RayPairs rp;
            GradPairs G;
            size_t pix[kPix];
            bool valid[kPix];
            float gq[kPix][4];
            bool rays_fast = (in_safe_range(ev[0]) || ev[0] == 0.0f) && (in_safe_range(ev[1]) || ev[1] == 0.0f);
#pragma unroll
            for (int q = 0; q < kPix; ++q) {
                const int pxq = px0 + lane + 32 * (q & 1), pyq = py0 + kPairs * warp + (q >> 1);
                valid[q] = pxq < p.W && pyq < p.H;
                pix[q] = (size_t)min(pyq, p.H - 1) * p.W + min(pxq, p.W - 1);
                const float* rd = rays + pix[q];
                rc[q] = make_ray_const(__ldg(rd), __ldg(rd + img), __ldg(rd + 2 * img), ev, zd);
                rays_fast = rays_fast && rc[q].fast && fabsf(rc[q].rx2) <= 0x1p40f && fabsf(rc[q].ry2) <= 0x1p40f;
                // pixels of the tile overhang carry zero upstream gradient: they scatter nothing
                const float* gc = p.g_color + (size_t)v * 3 * img + pix[q];
                gq[q][0] = valid[q] ? gscale * __ldg(gc) : 0.0f;
                gq[q][1] = valid[q] ? gscale * __ldg(gc + img) : 0.0f;
                gq[q][2] = valid[q] ? gscale * __ldg(gc + 2 * img) : 0.0f;
                gq[q][3] = (valid[q] && p.g_depth) ? __ldg(p.g_depth + (size_t)v * img + pix[q]) * rc[q].dz : 0.0f;
            }
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
            const bool warp_fast = __all_sync(0xffffffffu, rays_fast);
            f2 R[kPairs];
#pragma unroll
            for (int P = 0; P < kPairs; ++P) R[P] = splat(0.0f);
            for (int ii = 0; ii < N; ++ii) {
                const int i = N - 1 - ii;
                const int s = c_stage;
                const uint32_t ph = c_phase;
                if (++c_stage == kStages) { c_stage = 0; c_phase ^= 1u; }
                const PlaneConst pcc = s_pc[i];    // (prefetching it one plane ahead, as the forward does, only adds spills here)
                CoordPairs cc;
                if (warp_fast) coords_pairs<kAlignCorners>(pcc, rp, ex2, ey2, hsx2, hsy2, fWt, fHt, cc);
                float* gplane = p.g_rgba + ((size_t)m * N + i) * 4 * tex;
                mbar_wait(&s_full[s], ph);
                const StageMeta mt = s_meta[s];
                const float* sb = s_buf + s * kStageFloatsBwd;
                const int sel = mt.sel;
                f2 T[kPairs];      // transmittance saved by the forward, staged next to the plane tile: [kTileH][kTileW]
#pragma unroll
                for (int P = 0; P < kPairs; ++P) {
                    const float* tr = sb + kStageFloats + (kPairs * warp + P) * kTileW + lane;
                    T[P] = make_float2(tr[0], tr[32]);
                }
                bool done = false;
                if (warp_fast) {   // warp-uniform, one-hot class
                    if (sel & (1 << 18)) done = bwd_pairs<72>(sb, mt.cx, mt.cy, mt.rows2, cc, T, G, R, gplane, tex, Wt, Ht);
                    else if (sel & (1 << 17)) done = bwd_pairs<64>(sb, mt.cx, mt.cy, mt.rows2, cc, T, G, R, gplane, tex, Wt, Ht);
                    else if (sel & (1 << 19)) done = bwd_pairs<80>(sb, mt.cx, mt.cy, mt.rows2, cc, T, G, R, gplane, tex, Wt, Ht);
                    else if (sel & (1 << 16)) done = bwd_pairs<56>(sb, mt.cx, mt.cy, mt.rows2, cc, T, G, R, gplane, tex, Wt, Ht);
                    else if (sel & (1 << 20)) done = bwd_pairs<88>(sb, mt.cx, mt.cy, mt.rows2, cc, T, G, R, gplane, tex, Wt, Ht);
                }
                __syncwarp();
                mbar_arrive_if(&s_empty[s], lane == 0);     // the generic body below does not read the staged box
                if (!done) {
                    // ---- generic body (rare): per-pixel checks, sampling straight from global memory.  Also taken when the
                    // producer's corner-ray estimate says "nothing under the tile" (mode 1): that is a hint, never trusted ----
                    const float* plane = p.rgba + ((size_t)m * N + i) * 4 * tex;
                    float* Rs = reinterpret_cast<float*>(R);
                    const float* Ts = reinterpret_cast<const float*>(T);
#pragma unroll
                    for (int q = 0; q < kPix; ++q) {
                        RayConst rg = rc[q];
                        rg.fast = false;
                        const TexCoord tc = plane_coord<kAlignCorners>(pcc, rg, hsx, hsy, fWt, fHt);
                        if (!coord_hits(tc.ix, tc.iy, fWt, fHt)) continue;
                        const float4 sv = sample_plane_direct(plane, Ht, Wt, tc.ix, tc.iy);
                        const float fx = floorf(tc.ix), fy = floorf(tc.iy);
                        const float wx1 = tc.ix - fx, wy1 = tc.iy - fy, wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
                        const float qv = fmaf(gq[q][0], sv.x, fmaf(gq[q][1], sv.y, fmaf(gq[q][2], sv.z, gq[q][3] * tc.scale)));
                        const float d = qv - Rs[q];
                        const float w = sv.w * Ts[q];
                        Rs[q] = fmaf(sv.w, d, Rs[q]);
                        const float vv[4] = {gq[q][0] * w, gq[q][1] * w, gq[q][2] * w, Ts[q] * d};
                        scatter_pixel<false>(gplane, tex, Wt, Ht, (int)fx, (int)fy, vv, wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1);
                    }
                }
            }
        }
    }
}

}  // namespace gmpi
