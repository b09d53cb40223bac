// Forward, TMA-staged persistent variant (the fast path).
//
// One CTA per SM walks (tile, plane) pairs: a 64x32-pixel output tile, planes front to back.  A producer warp
// computes, from the tile's four corner rays, the texel footprint of the tile on the next plane and issues
// cp.async.bulk.tensor copies of exactly that footprint (all four channels, row-chunks of kRowsPerOp) into a
// 3-stage shared-memory ring; 16 consumer warps (4 pixels per thread) take their 16 bilinear taps per plane from
// shared memory (conflict-free: a warp reads 32 consecutive x of one row) and composite in registers.
// TMA's out-of-bounds zero fill implements padding_mode="zeros".  Every consumer thread verifies that its taps lie
// inside the staged box and otherwise samples global memory directly, so results never depend on the footprint
// estimate (arbitrary ray tensors stay correct, only slower).
#pragma once
#include "mpi_common.cuh"
#include "tma_utils.cuh"

namespace gmpi {

constexpr int kTileW = 64, kTileH = 32;
constexpr int kConsWarps = 16, kConsThreads = kConsWarps * 32, kStagedThreads = kConsThreads + 32;
constexpr int kStages = 3;
constexpr int kRowsPerOp = 4;
constexpr int kMaxBW = 80, kMaxBH = 44;                 // largest staged footprint (texels)
constexpr int kMinBW = 40, kBWStep = 8;   // multiples of 8: row pitch 4*bw = 0 mod 32 banks
constexpr int kNumMaps = (kMaxBW - kMinBW) / kBWStep + 1;
constexpr int kStageFloats = kMaxBW * kMaxBH * 4;
constexpr size_t kStagedSmem = (size_t)kStages * kStageFloats * 4;

struct TmaMaps {
    CUtensorMap m[kNumMaps];
};

// per-stage header written by the producer before it arms the full barrier
struct __align__(16) StageMeta {
    float fbx0, fby0;      // box origin (texel coordinates of smem element [0][.][0]) as floats
    float fbw2, fbh2;      // bw-2, rows-2: a footprint with north-west tap (rx, ry) fits iff 0<=rx<=bw-2, 0<=ry<=rows-2
    int bw;                // staged width (row pitch = 4*bw floats)
    int mode;              // 0 staged, 1 nothing to sample (footprint misses the texture), 2 sample global memory directly
    int pad0, pad1;
    PlaneConst pc;
};

// Rare path (a ray whose footprint is not in the staged box): sample the plane from global memory.  Out of line so
// that it does not cost registers in the hot loop.
__device__ __noinline__ float4 sample_plane_direct(const float* __restrict__ plane, int Ht, int Wt, float ix, float iy) {
    const size_t tex = (size_t)Ht * Wt;
    const Taps tp = make_taps(ix, iy, Ht, Wt);
    return make_float4(tap4(plane, tp), tap4(plane + tex, tp), tap4(plane + 2 * tex, tp), tap4(plane + 3 * tex, tp));
}

template <bool kAlignCorners>
__global__ void __launch_bounds__(kStagedThreads, 1)
mpi_fwd_staged_kernel(const RenderParams p, const __grid_constant__ TmaMaps maps, const int tiles_x, const int tiles_y) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* s_buf = reinterpret_cast<float*>(smem_raw);   // the ring starts the dynamic segment (1024-byte aligned)
    __shared__ StageMeta s_meta[kStages];
    __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
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
    const int tiles_per_view = tiles_x * tiles_y;
    const int n_tiles = tiles_per_view * p.V;

    if (warp == kConsWarps) {
        // ================================ producer warp ================================
        if (lane < kNumMaps) tma_prefetch_desc(&maps.m[lane]);
        uint32_t it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int v = t / tiles_per_view, tt = t - v * tiles_per_view;
            const int px0 = (tt % tiles_x) * kTileW, py0 = (tt / tiles_x) * kTileH;
            const int m = __ldg(p.view2mpi + v);
            const float* e = p.eye + 3 * v;
            const float ev[3] = {__ldg(e), __ldg(e + 1), __ldg(e + 2)};
            const float zd[3] = {0.f, 0.f, 1.f};
            // lanes 0..3 (replicated over the warp): the four corner pixels of the tile, clamped into the image
            const int cx = min(px0 + ((lane & 1) ? kTileW - 1 : 0), p.W - 1);
            const int cy = min(py0 + ((lane & 2) ? kTileH - 1 : 0), p.H - 1);
            const float* rd = p.ray_dir + (size_t)v * 3 * img + (size_t)cy * p.W + cx;
            const RayConst rc = make_ray_const(__ldg(rd), __ldg(rd + img), __ldg(rd + 2 * img), ev, zd);
            for (int i = 0; i < N; ++i, ++it) {
                const int s = it % kStages;
                const uint32_t ph = (it / kStages) & 1;
                const PlaneConst pc = make_plane_const(p.dhw + ((size_t)m * N + i) * 3, ev[2]);
                const TexCoord tc = plane_coord<kAlignCorners>(pc, rc, hsx, hsy, fWt, fHt);
                // footprint of the tile = bounding box of the corner coordinates (the pixel -> texel map is projective,
                // hence monotone along image rows and columns), +-1 texel of slack for rounding
                const bool finite = fabsf(tc.ix) < 1e9f && fabsf(tc.iy) < 1e9f;
                const bool all_finite = __all_sync(0xffffffffu, finite);
                const int fx = finite ? (int)floorf(tc.ix) : 0, fy = finite ? (int)floorf(tc.iy) : 0;
                const int xmin = __reduce_min_sync(0xffffffffu, fx), xmax = __reduce_max_sync(0xffffffffu, fx);
                const int ymin = __reduce_min_sync(0xffffffffu, fy), ymax = __reduce_max_sync(0xffffffffu, fy);
                const int bx0 = xmin - 1, by0 = ymin - 1;
                const int need_w = xmax - xmin + 4, need_h = ymax - ymin + 4;     // +1 east/south tap, +-1 slack
                int mode = 0;
                if (!all_finite || need_w > kMaxBW || need_h > kMaxBH) mode = 2;
                else if (bx0 > Wt - 1 || bx0 + need_w - 1 < 0 || by0 > Ht - 1 || by0 + need_h - 1 < 0) mode = 1;
                const int k = mode == 0 ? max(0, (need_w - kMinBW + kBWStep - 1) / kBWStep) : 0;
                const int bw = kMinBW + k * kBWStep;
                const int n_ops = mode == 0 ? (need_h + kRowsPerOp - 1) / kRowsPerOp : 0;
                const int rows = n_ops * kRowsPerOp;
                mbar_wait(&s_empty[s], ph ^ 1);
                if (lane == 0) {
                    StageMeta mt;
                    mt.fbx0 = (float)bx0; mt.fby0 = (float)by0;
                    mt.fbw2 = (float)(bw - 2); mt.fbh2 = (float)(rows - 2);
                    mt.bw = bw; mt.mode = mode; mt.pad0 = mt.pad1 = 0;
                    mt.pc = pc;
                    s_meta[s] = mt;
                    if (n_ops > 0) mbar_arrive_expect_tx(&s_full[s], (uint32_t)(rows * bw * 16));
                    else mbar_arrive(&s_full[s]);
                }
                __syncwarp();
                if (lane < n_ops) {
                    float* dst = s_buf + (size_t)s * kStageFloats + (size_t)lane * kRowsPerOp * 4 * bw;
                    tma_load_4d(dst, &maps.m[k], &s_full[s], bx0, 0, by0 + lane * kRowsPerOp, m * N + i);
                }
            }
        }
    } else {
        // ================================ consumer warps ================================
        // warp w owns rows 2w, 2w+1 of the tile; a lane owns x = lane and lane+32 on both rows
        const bool check_last = (p.options & GMPI_CHECK_LAST_PLANE) != 0;
        const bool minus1_1 = (p.options & GMPI_COLOR_MINUS1_1) != 0;
        uint32_t it = 0;
        uint32_t flag = 0;
        if (blockIdx.x == 0) {          // mpi.py:70: every plane distance against view 0's eye
            const float eye0_z = __ldg(p.eye + 2);
            for (int j = threadIdx.x; j < p.M * N; j += kConsThreads)
                if (!(__ldg(p.dhw + (size_t)j * 3) >= eye0_z)) flag |= GMPI_FLAG_PLANE_BEHIND_EYE;
        }
        const size_t tex = (size_t)Ht * Wt;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int v = t / tiles_per_view, tt = t - v * tiles_per_view;
            const int px0 = (tt % tiles_x) * kTileW, py0 = (tt / tiles_x) * kTileH;
            const int m = __ldg(p.view2mpi + v);
            const float* e = p.eye + 3 * v;
            const float ev[3] = {__ldg(e), __ldg(e + 1), __ldg(e + 2)};
            const float zd[3] = {__ldg(p.z_dir + 3 * v), __ldg(p.z_dir + 3 * v + 1), __ldg(p.z_dir + 3 * v + 2)};
            const float* rays = p.ray_dir + (size_t)v * 3 * img;
            RayConst rc[4];
            bool rays_fast = in_safe_range(ev[0]) || ev[0] == 0.0f;
            rays_fast = rays_fast && (in_safe_range(ev[1]) || ev[1] == 0.0f);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int px = min(px0 + lane + 32 * (q & 1), p.W - 1), py = min(py0 + 2 * warp + (q >> 1), p.H - 1);
                const float* rd = rays + (size_t)py * p.W + px;
                rc[q] = make_ray_const(__ldg(rd), __ldg(rd + img), __ldg(rd + 2 * img), ev, zd);
                rays_fast = rays_fast && rc[q].fast && fabsf(rc[q].rx2) <= 0x1p40f && fabsf(rc[q].ry2) <= 0x1p40f;
            }
            // warp-uniform: every ray of this warp is in the range where the reciprocal+FMA division is exact and
            // no coordinate can be NaN, so the per-plane body needs no per-pixel range checks
            const bool warp_fast = __all_sync(0xffffffffu, rays_fast);
            float T[4] = {1.f, 1.f, 1.f, 1.f}, cr[4] = {0.f, 0.f, 0.f, 0.f}, cg[4] = {0.f, 0.f, 0.f, 0.f},
                  cb[4] = {0.f, 0.f, 0.f, 0.f}, cws[4] = {0.f, 0.f, 0.f, 0.f};
            const float* plane = p.rgba + (size_t)m * N * 4 * tex;
            for (int i = 0; i < N; ++i, ++it, plane += 4 * tex) {
                const int s = it % kStages;
                const uint32_t ph = (it / kStages) & 1;
                mbar_wait(&s_full[s], ph);
                const StageMeta mt = s_meta[s];
                const float* sb = s_buf + s * kStageFloats;
                const int bw = mt.bw, bw4 = 4 * mt.bw;
                bool done = false;
                if (warp_fast && mt.pc.fast != 0.0f && mt.mode == 0) {
                    // ---- fast body: straight-line code for the four pixels ----
                    float ix[4], iy[4], sc[4], fx0[4], fy0[4], rx[4], ry[4];
                    bool inbox = true;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float sq = div_by_rcp(mt.pc.z_diff, rc[q].rz, rc[q].yrz);
                        const float X2 = __fadd_rn(rc[q].ex2, __fmul_rn(rc[q].rx2, sq));
                        const float Y2 = __fadd_rn(rc[q].ey2, __fmul_rn(rc[q].ry2, sq));
                        float u = div_by_rcp(X2, mt.pc.pw, mt.pc.ypw);
                        float vv = div_by_rcp(Y2, mt.pc.ph, mt.pc.yph);
                        if (kAlignCorners) {
                            ix[q] = __fmul_rn(__fadd_rn(u, 1.0f), hsx);
                            iy[q] = __fmul_rn(__fadd_rn(vv, 1.0f), hsy);
                        } else {
                            if (u >= -1.0f && u <= 1.0f) u = __fmul_rn(u, 0.95f);
                            if (vv >= -1.0f && vv <= 1.0f) vv = __fmul_rn(vv, 0.95f);
                            ix[q] = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(u, 1.0f), fWt), -1.0f), 0.5f);
                            iy[q] = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(vv, 1.0f), fHt), -1.0f), 0.5f);
                        }
                        sc[q] = sq;
                        fx0[q] = floorf(ix[q]); fy0[q] = floorf(iy[q]);
                        rx[q] = fx0[q] - mt.fbx0; ry[q] = fy0[q] - mt.fby0;
                        inbox = inbox && rx[q] >= 0.0f && rx[q] <= mt.fbw2 && ry[q] >= 0.0f && ry[q] <= mt.fbh2;
                    }
                    if (inbox) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float wx1 = ix[q] - fx0[q], wy1 = iy[q] - fy0[q];
                            const float wy0 = 1.0f - wy1;
                            const float w11 = wx1 * wy1, w10 = wy1 - w11, w01 = wx1 - w11, w00 = wy0 - w01;
                            const float* t0 = sb + ((int)ry[q] * bw4 + (int)rx[q]);     // [row][channel][x]
                            const float* t1 = t0 + bw4;
                            const float r = fmaf(t1[1], w11, fmaf(t1[0], w10, fmaf(t0[1], w01, t0[0] * w00)));
                            t0 += bw; t1 += bw;
                            const float g = fmaf(t1[1], w11, fmaf(t1[0], w10, fmaf(t0[1], w01, t0[0] * w00)));
                            t0 += bw; t1 += bw;
                            const float b = fmaf(t1[1], w11, fmaf(t1[0], w10, fmaf(t0[1], w01, t0[0] * w00)));
                            t0 += bw; t1 += bw;
                            const float a = fmaf(t1[1], w11, fmaf(t1[0], w10, fmaf(t0[1], w01, t0[0] * w00)));
                            const float w = a * T[q];                       // mpi.py:423
                            cr[q] = fmaf(w, r, cr[q]);                      // mpi.py:430
                            cg[q] = fmaf(w, g, cg[q]);
                            cb[q] = fmaf(w, b, cb[q]);
                            cws[q] = fmaf(w, sc[q], cws[q]);                // depth_i = scale * (ray . z_dir), mpi.py:150
                            T[q] -= w;   // T(1-a); the reference's +1e-10 changes any later weight by < 1e-10 absolute
                        }
                        done = true;
                    }
                }
                if (!done) {
                    // ---- generic body: per-pixel range / box checks, direct sampling when not staged ----
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const TexCoord tc = plane_coord<kAlignCorners>(mt.pc, rc[q], hsx, hsy, fWt, fHt);
                        const float fx = floorf(tc.ix), fy = floorf(tc.iy);
                        const float rxx = fx - mt.fbx0, ryy = fy - mt.fby0;
                        float r, g, b, a;
                        if (mt.mode == 0 && rxx >= 0.0f && rxx <= mt.fbw2 && ryy >= 0.0f && ryy <= mt.fbh2) {
                            const float wx1 = tc.ix - fx, wy1 = tc.iy - fy;
                            const float wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
                            const float w00 = wx0 * wy0, w01 = wx1 * wy0, w10 = wx0 * wy1, w11 = wx1 * wy1;
                            const float* t0 = sb + ((int)ryy * bw4 + (int)rxx);
                            const float* t1 = t0 + bw4;
                            r = fmaf(t1[1], w11, fmaf(t1[0], w10, fmaf(t0[1], w01, t0[0] * w00)));
                            g = fmaf(t1[bw + 1], w11, fmaf(t1[bw], w10, fmaf(t0[bw + 1], w01, t0[bw] * w00)));
                            b = fmaf(t1[2 * bw + 1], w11, fmaf(t1[2 * bw], w10, fmaf(t0[2 * bw + 1], w01, t0[2 * bw] * w00)));
                            a = fmaf(t1[3 * bw + 1], w11, fmaf(t1[3 * bw], w10, fmaf(t0[3 * bw + 1], w01, t0[3 * bw] * w00)));
                        } else if (mt.mode != 1 && coord_hits(tc.ix, tc.iy, fWt, fHt)) {
                            const float4 sv = sample_plane_direct(plane, Ht, Wt, tc.ix, tc.iy);
                            r = sv.x; g = sv.y; b = sv.z; a = sv.w;
                        } else {
                            continue;   // no texel under this ray on this plane: contributes exactly nothing
                        }
                        const float w = a * T[q];
                        cr[q] = fmaf(w, r, cr[q]);
                        cg[q] = fmaf(w, g, cg[q]);
                        cb[q] = fmaf(w, b, cb[q]);
                        cws[q] = fmaf(w, tc.scale, cws[q]);
                        T[q] -= w;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[s]);
                if (check_last && i == N - 1) {     // assert_not_out_of_last_plane, mpi.py:103-109 (once per tile)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const TexCoord tc = plane_coord<kAlignCorners>(mt.pc, rc[q], hsx, hsy, fWt, fHt);
                        if (!(tc.u >= -1.0f && tc.u <= 1.0f && tc.v >= -1.0f && tc.v <= 1.0f)) flag |= GMPI_FLAG_LAST_PLANE_OOB;
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int px = px0 + lane + 32 * (q & 1), py = py0 + 2 * warp + (q >> 1);
                if (px >= p.W || py >= p.H) continue;
                const size_t pix = (size_t)py * p.W + px;
                float o0 = cr[q], o1 = cg[q], o2 = cb[q];
                if (minus1_1) {
                    o0 = fmaf(2.0f, o0, -1.0f); o1 = fmaf(2.0f, o1, -1.0f); o2 = fmaf(2.0f, o2, -1.0f);
                }
                float* co = p.color + (size_t)v * 3 * img + pix;
                co[0] = o0; co[img] = o1; co[2 * img] = o2;
                p.depth[(size_t)v * img + pix] = cws[q] * rc[q].dz;
            }
        }
        if (flag) atomicOr(p.flags, flag);
    }
}

}  // namespace gmpi
