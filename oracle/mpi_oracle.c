/*
 * oracle/mpi_oracle.c -- CPU restatement of the GMPI multiplane-image render path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA kernels in
 * ml_gmpi_b200/csrc/.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it.  The product path never does: it has no
 * CPU fallback and fails loudly when the CUDA library is missing.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors of its own
 * (SURVEY.md section 4), so this restatement is pinned against outputs of the *unmodified
 * reference run in the build container* (oracle/make_golden.py -> tests/golden/ *.npz, checked
 * by tests/test_oracle_golden.py).
 *
 * What is restated (reference file:line, all under /root/reference):
 *   homography()            gmpi/core/mpi.py:26-153   ray/plane intersection, normalised
 *                                                      coords, F.grid_sample, z-depth
 *   MPI.forward compositing gmpi/core/mpi.py:411-436  T_i = prod_{j<i}(1-a_j+1e-10),
 *                                                      w_i = a_i T_i, color/depth sums
 *   MPI.old_forward         gmpi/core/mpi.py:280-304  back-to-front "over" (second oracle)
 *   autograd of the above   (torch: grid_sampler_2d_backward, cumprod_backward)
 * Third-party arithmetic restated from its published definition (torch is not under
 * /root/reference; reference pins pytorch=1.9.1, environment.yml:22; container has 2.11.0):
 *   F.grid_sample(mode="bilinear", padding_mode="zeros"), ATen/native/GridSampler.h
 *   grid_sampler_unnormalize: align_corners ? ((c+1)/2)*(size-1) : ((c+1)*size-1)/2.
 *
 * Every floating-point operation up to the texel coordinate (ix, iy) is written as a separate
 * fp32 statement (compile with -ffp-contract=off): the coordinate is amplified by the texture
 * size times the texel gradient, so the 1e-4 parity bar needs the reference's exact rounding
 * sequence there.  After (ix, iy) the arithmetic is well conditioned.
 *
 * Layouts (row-major, fp32): rgba [M,N,4,Ht,Wt], dhw [M,N,3], view2mpi [V] int32,
 * ray_dir [V,3,H,W], eye [V,3], z_dir [V,3], color [V,3,H,W], depth [V,1,H,W].
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__FAST_MATH__)
#error "the oracle must be built without -ffast-math (exact fp32 op sequence)"
#endif

#define NARROW_SCALE 0.95f /* mpi.py:23 ALIGN_CORNERS_FALSE_NARROW_SCALE */

typedef struct {
    float ix, iy;   /* un-normalised texel coordinate (grid_sampler_unnormalize) */
    float scale;    /* ray parameter of the intersection, mpi.py:76 */
    float u, v;     /* normalised coordinate handed to grid_sample, mpi.py:89-99 */
} coord_t;

/* mpi.py:67-99 + GridSampler.h unnormalize.  One (pixel, plane) pair. */
static inline coord_t plane_coord(float d, float ph, float pw, const float eye[3],
                                  float rx, float ry, float rz, int Ht, int Wt,
                                  int align_corners) {
    coord_t c;
    float z_diff = d - eye[2];          /* mpi.py:74 */
    float scale = z_diff / rz;          /* mpi.py:76 */
    float tx = rx * scale;              /* mpi.py:79 (mul, then add: two roundings) */
    float ty = ry * scale;
    float X = eye[0] + tx;
    float Y = eye[1] + ty;
    float X2 = 2.0f * X;                /* mpi.py:89-90: (2*x)/width */
    float Y2 = 2.0f * Y;
    float u = X2 / pw;
    float v = Y2 / ph;
    if (!align_corners) {                        /* mpi.py:95-99 */
        if (v >= -1.0f && v <= 1.0f) v = v * NARROW_SCALE;
        if (u >= -1.0f && u <= 1.0f) u = u * NARROW_SCALE;
    }
    float ix, iy;
    if (align_corners) {
        float a = u + 1.0f, b = v + 1.0f;
        float a2 = a / 2.0f, b2 = b / 2.0f;
        ix = a2 * (float)(Wt - 1);
        iy = b2 * (float)(Ht - 1);
    } else {
        float a = u + 1.0f, b = v + 1.0f;
        float am = a * (float)Wt, bm = b * (float)Ht;
        float as = am - 1.0f, bs = bm - 1.0f;
        ix = as / 2.0f;
        iy = bs / 2.0f;
    }
    c.ix = ix; c.iy = iy; c.scale = scale; c.u = u; c.v = v;
    return c;
}

typedef struct {
    int x0, y0;          /* north-west tap */
    float w[4];          /* nw, ne, sw, se bilinear weights */
    int ok[4];           /* tap inside the texture (padding_mode="zeros") */
} taps_t;

/* ATen grid_sampler_2d bilinear footprint. */
static inline taps_t bilinear_taps(float ix, float iy, int Ht, int Wt) {
    taps_t t;
    float fx0 = floorf(ix), fy0 = floorf(iy);
    float fx1 = fx0 + 1.0f, fy1 = fy0 + 1.0f;
    t.w[0] = (fx1 - ix) * (fy1 - iy);
    t.w[1] = (ix - fx0) * (fy1 - iy);
    t.w[2] = (fx1 - ix) * (iy - fy0);
    t.w[3] = (ix - fx0) * (iy - fy0);
    /* coordinates far outside (or NaN) never pass the range test below */
    int in_range = (ix > -2.0f) && (ix < (float)Wt + 1.0f) && (iy > -2.0f) && (iy < (float)Ht + 1.0f);
    if (!in_range) {
        t.x0 = t.y0 = -4;
        t.ok[0] = t.ok[1] = t.ok[2] = t.ok[3] = 0;
        return t;
    }
    t.x0 = (int)fx0; t.y0 = (int)fy0;
    int x1 = t.x0 + 1, y1 = t.y0 + 1;
    int xa = t.x0 >= 0 && t.x0 < Wt, xb = x1 >= 0 && x1 < Wt;
    int ya = t.y0 >= 0 && t.y0 < Ht, yb = y1 >= 0 && y1 < Ht;
    t.ok[0] = xa && ya; t.ok[1] = xb && ya; t.ok[2] = xa && yb; t.ok[3] = xb && yb;
    return t;
}

static inline float tap_sum(const float* ch, const taps_t* t, int Wt) {
    float acc = 0.0f;
    if (t->ok[0]) acc += ch[(size_t)t->y0 * Wt + t->x0] * t->w[0];
    if (t->ok[1]) acc += ch[(size_t)t->y0 * Wt + t->x0 + 1] * t->w[1];
    if (t->ok[2]) acc += ch[(size_t)(t->y0 + 1) * Wt + t->x0] * t->w[2];
    if (t->ok[3]) acc += ch[(size_t)(t->y0 + 1) * Wt + t->x0 + 1] * t->w[3];
    return acc;
}

/* flags, same bit meaning as include/gmpi_mpi_render.h */
#define FLAG_RGBA_RANGE 1u
#define FLAG_ALPHA_RANGE 2u
#define FLAG_LAST_PLANE_OOB 4u
#define FLAG_PLANE_BEHIND_EYE 8u

/*
 * MPI.forward, mpi.py:308-436.  Returns the flag word (0 = all reference asserts hold).
 * check_last_plane mirrors assert_not_out_of_last_plane (mpi.py:381-395, 103-109).
 * Rows of a view are split over `nthreads` pthreads (the image has no libgomp); per-pixel
 * arithmetic is independent of the split.
 */
typedef struct {
    const float *rgba, *dhw, *ray_dir, *eye, *z_dir;
    float *color, *depth;
    int m, v, N, Ht, Wt, H, W, align_corners, check_last_plane, row0, row1;
    uint32_t flags;
} fwd_job_t;

static void* fwd_rows(void* arg) {
    fwd_job_t* j = (fwd_job_t*)arg;
    const int N = j->N, Ht = j->Ht, Wt = j->Wt, H = j->H, W = j->W, v = j->v, m = j->m;
    const size_t tex = (size_t)Ht * Wt, img = (size_t)H * W;
    const float* e = j->eye + 3 * v;
    const float* zd = j->z_dir + 3 * v;
    uint32_t flags = 0;
    for (int py = j->row0; py < j->row1; ++py) {
        for (int px = 0; px < W; ++px) {
            const size_t p = (size_t)py * W + px;
            const float rx = j->ray_dir[((size_t)v * 3 + 0) * img + p];
            const float ry = j->ray_dir[((size_t)v * 3 + 1) * img + p];
            const float rz = j->ray_dir[((size_t)v * 3 + 2) * img + p];
            /* mpi.py:149 einsum("nchw,nc->nhw") */
            const float dist2depth = rx * zd[0] + ry * zd[1] + rz * zd[2];
            float T = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, cd = 0.f;
            for (int i = 0; i < N; ++i) {
                const float* pd = j->dhw + ((size_t)m * N + i) * 3;
                coord_t c = plane_coord(pd[0], pd[1], pd[2], e, rx, ry, rz, Ht, Wt, j->align_corners);
                if (j->check_last_plane && i == N - 1) {
                    if (!(c.u >= -1.0f) || !(c.u <= 1.0f) || !(c.v >= -1.0f) || !(c.v <= 1.0f))
                        flags |= FLAG_LAST_PLANE_OOB;
                }
                taps_t t = bilinear_taps(c.ix, c.iy, Ht, Wt);
                const float* base = j->rgba + ((size_t)m * N + i) * 4 * tex;
                float r = tap_sum(base, &t, Wt);
                float g = tap_sum(base + tex, &t, Wt);
                float b = tap_sum(base + 2 * tex, &t, Wt);
                float a = tap_sum(base + 3 * tex, &t, Wt);
                /* mpi.py:150-151,411: depth = scale*dist2depth; disp = 1/depth; depth = 1/disp */
                float dpt = c.scale * dist2depth;
                float disp = 1.0f / dpt;
                dpt = 1.0f / disp;
                float wgt = a * T;                 /* mpi.py:423 */
                cr += wgt * r; cg += wgt * g; cb += wgt * b; cd += wgt * dpt;   /* :430,:434 */
                T = T * ((1.0f - a) + 1e-10f);     /* mpi.py:421,423 */
            }
            j->color[((size_t)v * 3 + 0) * img + p] = cr;
            j->color[((size_t)v * 3 + 1) * img + p] = cg;
            j->color[((size_t)v * 3 + 2) * img + p] = cb;
            j->depth[(size_t)v * img + p] = cd;
        }
    }
    j->flags = flags;
    return NULL;
}

uint32_t gmpi_oracle_forward_mt(const float* rgba, const int32_t* view2mpi, const float* dhw,
                                const float* ray_dir, const float* eye, const float* z_dir,
                                float* color, float* depth, int M, int V, int N, int Ht, int Wt,
                                int H, int W, int align_corners, int check_last_plane,
                                int nthreads) {
    uint32_t flags = 0;
    (void)M;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > H) nthreads = H;
    if (nthreads > 256) nthreads = 256;
    for (int v = 0; v < V; ++v) {
        const int m = view2mpi[v];
        /* mpi.py:70: distance >= z_eye[0]  (the reference compares against view 0's eye only) */
        for (int i = 0; i < N; ++i)
            if (!(dhw[((size_t)m * N + i) * 3] >= eye[2])) flags |= FLAG_PLANE_BEHIND_EYE;
        fwd_job_t jobs[256];
        pthread_t th[256];
        for (int t = 0; t < nthreads; ++t) {
            fwd_job_t jb = {rgba, dhw, ray_dir, eye, z_dir, color, depth, m, v, N, Ht, Wt, H, W,
                            align_corners, check_last_plane,
                            (int)((long)H * t / nthreads), (int)((long)H * (t + 1) / nthreads), 0};
            jobs[t] = jb;
            if (nthreads == 1) fwd_rows(&jobs[t]);
            else pthread_create(&th[t], NULL, fwd_rows, &jobs[t]);
        }
        for (int t = 0; t < nthreads; ++t) {
            if (nthreads > 1) pthread_join(th[t], NULL);
            flags |= jobs[t].flags;
        }
    }
    return flags;
}

uint32_t gmpi_oracle_forward(const float* rgba, const int32_t* view2mpi, const float* dhw,
                             const float* ray_dir, const float* eye, const float* z_dir,
                             float* color, float* depth, int M, int V, int N, int Ht, int Wt,
                             int H, int W, int align_corners, int check_last_plane) {
    return gmpi_oracle_forward_mt(rgba, view2mpi, dhw, ray_dir, eye, z_dir, color, depth, M, V, N,
                                  Ht, Wt, H, W, align_corners, check_last_plane, 1);
}

/* MPI.old_forward, mpi.py:280-304: back-to-front "over", no epsilon.  Second oracle. */
void gmpi_oracle_forward_over(const float* rgba, const int32_t* view2mpi, const float* dhw,
                              const float* ray_dir, const float* eye, const float* z_dir,
                              float* color, float* depth, int M, int V, int N, int Ht, int Wt,
                              int H, int W, int align_corners) {
    const size_t tex = (size_t)Ht * Wt, img = (size_t)H * W;
    (void)M;
    for (int v = 0; v < V; ++v) {
        const int m = view2mpi[v];
        const float* e = eye + 3 * v;
        const float* zd = z_dir + 3 * v;
        for (int py = 0; py < H; ++py)
            for (int px = 0; px < W; ++px) {
                const size_t p = (size_t)py * W + px;
                const float rx = ray_dir[((size_t)v * 3 + 0) * img + p];
                const float ry = ray_dir[((size_t)v * 3 + 1) * img + p];
                const float rz = ray_dir[((size_t)v * 3 + 2) * img + p];
                const float dist2depth = rx * zd[0] + ry * zd[1] + rz * zd[2];
                float cr = 0.f, cg = 0.f, cb = 0.f, cd = 0.f;
                for (int i = N - 1; i >= 0; --i) {
                    const float* pd = dhw + ((size_t)m * N + i) * 3;
                    coord_t c = plane_coord(pd[0], pd[1], pd[2], e, rx, ry, rz, Ht, Wt, align_corners);
                    taps_t t = bilinear_taps(c.ix, c.iy, Ht, Wt);
                    const float* base = rgba + ((size_t)m * N + i) * 4 * tex;
                    float r = tap_sum(base, &t, Wt), g = tap_sum(base + tex, &t, Wt);
                    float b = tap_sum(base + 2 * tex, &t, Wt), a = tap_sum(base + 3 * tex, &t, Wt);
                    float dpt = 1.0f / (1.0f / (c.scale * dist2depth));
                    cr = a * r + (1.0f - a) * cr;      /* mpi.py:302 */
                    cg = a * g + (1.0f - a) * cg;
                    cb = a * b + (1.0f - a) * cb;
                    cd = a * dpt + (1.0f - a) * cd;    /* mpi.py:304 */
                }
                color[((size_t)v * 3 + 0) * img + p] = cr;
                color[((size_t)v * 3 + 1) * img + p] = cg;
                color[((size_t)v * 3 + 2) * img + p] = cb;
                depth[(size_t)v * img + p] = cd;
            }
    }
}

/*
 * Gradient of sum(color*g_color) + sum(depth*g_depth) w.r.t. rgba, as torch autograd produces it
 * for MPI.forward: the grid and the depth are built under no_grad (mpi.py:65,148), so only the
 * sampled rgba carries gradient.
 *   w_i = a_i P_i, P_i = prod_{j<i} s_j, s_j = 1 - a_j + 1e-10           (mpi.py:421-423)
 *   q_i = sum_c Gc * rgb_ic + Gd * depth_i
 *   dL/d rgb_ic = Gc * w_i
 *   dL/d a_i    = P_i q_i  -  ( sum_{k>i} a_k q_k P_k ) / s_i    (cumprod_backward: reversed
 *                 cumsum of grad*output divided by input; inputs are never exactly zero here)
 * then grid_sampler_2d_backward scatters each through the four bilinear weights.
 * g_depth may be NULL.  g_rgba must be zero-initialised by the caller; views accumulate.
 */
typedef struct {
    const float *rgba, *dhw, *ray_dir, *eye, *z_dir, *g_color, *g_depth;
    float* g_rgba;
    int m, v, N, Ht, Wt, H, W, align_corners, row0, row1, atomic;
} bwd_job_t;

/* g += x.  With several row bands in flight two threads may hit the same texel (bilinear footprints of neighbouring rows
 * overlap): compare-and-swap on the bit pattern.  The sum is then order-dependent at the last-ulp level only. */
static inline void grad_add(float* g, float x, int atomic) {
    if (!atomic) { *g += x; return; }
    uint32_t* u = (uint32_t*)g;
    uint32_t old = __atomic_load_n(u, __ATOMIC_RELAXED), neu;
    do {
        float f;
        memcpy(&f, &old, 4);
        f += x;
        memcpy(&neu, &f, 4);
    } while (!__atomic_compare_exchange_n(u, &old, neu, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}

static void* bwd_rows(void* arg) {
    bwd_job_t* j = (bwd_job_t*)arg;
    const int N = j->N, Ht = j->Ht, Wt = j->Wt, H = j->H, W = j->W, v = j->v, m = j->m;
    const size_t tex = (size_t)Ht * Wt, img = (size_t)H * W;
    (void)H;
    float* sa = (float*)malloc(sizeof(float) * N * 6);
    float *A = sa, *P = sa + N, *Q = sa + 2 * N, *SUF = sa + 3 * N, *S = sa + 4 * N, *GA = sa + 5 * N;
    taps_t* tp = (taps_t*)malloc(sizeof(taps_t) * N);
    const float* e = j->eye + 3 * v;
    const float* zd = j->z_dir + 3 * v;
    for (int py = j->row0; py < j->row1; ++py)
        for (int px = 0; px < W; ++px) {
            const size_t p = (size_t)py * W + px;
            const float rx = j->ray_dir[((size_t)v * 3 + 0) * img + p];
            const float ry = j->ray_dir[((size_t)v * 3 + 1) * img + p];
            const float rz = j->ray_dir[((size_t)v * 3 + 2) * img + p];
            const float dist2depth = rx * zd[0] + ry * zd[1] + rz * zd[2];
            const float G0 = j->g_color[((size_t)v * 3 + 0) * img + p];
            const float G1 = j->g_color[((size_t)v * 3 + 1) * img + p];
            const float G2 = j->g_color[((size_t)v * 3 + 2) * img + p];
            const float Gd = j->g_depth ? j->g_depth[(size_t)v * img + p] : 0.0f;
            float T = 1.0f;
            for (int i = 0; i < N; ++i) {
                const float* pd = j->dhw + ((size_t)m * N + i) * 3;
                coord_t c = plane_coord(pd[0], pd[1], pd[2], e, rx, ry, rz, Ht, Wt, j->align_corners);
                tp[i] = bilinear_taps(c.ix, c.iy, Ht, Wt);
                const float* base = j->rgba + ((size_t)m * N + i) * 4 * tex;
                float r = tap_sum(base, &tp[i], Wt), g = tap_sum(base + tex, &tp[i], Wt);
                float b = tap_sum(base + 2 * tex, &tp[i], Wt), a = tap_sum(base + 3 * tex, &tp[i], Wt);
                float dpt = 1.0f / (1.0f / (c.scale * dist2depth));
                A[i] = a; P[i] = T; S[i] = (1.0f - a) + 1e-10f;
                Q[i] = G0 * r + G1 * g + G2 * b + Gd * dpt;
                T = T * S[i];
            }
            float suf = 0.0f;   /* reversed cumsum, back to front */
            for (int i = N - 1; i >= 0; --i) {
                SUF[i] = suf;
                suf += A[i] * Q[i] * P[i];
            }
            for (int i = 0; i < N; ++i) GA[i] = P[i] * Q[i] - SUF[i] / S[i];
            for (int i = 0; i < N; ++i) {
                float* gb = j->g_rgba + ((size_t)m * N + i) * 4 * tex;
                const float wgt = A[i] * P[i];
                const float gch[4] = {G0 * wgt, G1 * wgt, G2 * wgt, GA[i]};
                const taps_t* t = &tp[i];
                for (int ch = 0; ch < 4; ++ch) {
                    float* gc = gb + (size_t)ch * tex;
                    if (t->ok[0]) grad_add(gc + (size_t)t->y0 * Wt + t->x0, gch[ch] * t->w[0], j->atomic);
                    if (t->ok[1]) grad_add(gc + (size_t)t->y0 * Wt + t->x0 + 1, gch[ch] * t->w[1], j->atomic);
                    if (t->ok[2]) grad_add(gc + (size_t)(t->y0 + 1) * Wt + t->x0, gch[ch] * t->w[2], j->atomic);
                    if (t->ok[3]) grad_add(gc + (size_t)(t->y0 + 1) * Wt + t->x0 + 1, gch[ch] * t->w[3], j->atomic);
                }
            }
        }
    free(sa);
    free(tp);
    return NULL;
}

/* Row bands of each view on `nthreads` pthreads (views one after another).  nthreads == 1 is the plain sequential sum. */
void gmpi_oracle_backward_mt(const float* rgba, const int32_t* view2mpi, const float* dhw,
                             const float* ray_dir, const float* eye, const float* z_dir,
                             const float* g_color, const float* g_depth, float* g_rgba, int M, int V,
                             int N, int Ht, int Wt, int H, int W, int align_corners, int nthreads) {
    (void)M;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > H) nthreads = H;
    if (nthreads > 256) nthreads = 256;
    for (int v = 0; v < V; ++v) {
        bwd_job_t jobs[256];
        pthread_t th[256];
        for (int t = 0; t < nthreads; ++t) {
            bwd_job_t jb = {rgba, dhw, ray_dir, eye, z_dir, g_color, g_depth, g_rgba, view2mpi[v], v, N, Ht, Wt, H, W,
                            align_corners, (int)((long)H * t / nthreads), (int)((long)H * (t + 1) / nthreads), nthreads > 1};
            jobs[t] = jb;
            if (nthreads == 1) bwd_rows(&jobs[t]);
            else pthread_create(&th[t], NULL, bwd_rows, &jobs[t]);
        }
        if (nthreads > 1)
            for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    }
}

void gmpi_oracle_backward(const float* rgba, const int32_t* view2mpi, const float* dhw,
                          const float* ray_dir, const float* eye, const float* z_dir,
                          const float* g_color, const float* g_depth, float* g_rgba, int M, int V,
                          int N, int Ht, int Wt, int H, int W, int align_corners) {
    gmpi_oracle_backward_mt(rgba, view2mpi, dhw, ray_dir, eye, z_dir, g_color, g_depth, g_rgba, M, V, N, Ht, Wt, H, W,
                            align_corners, 1);
}

/* Texel coordinates only (for bit-exactness tests of the coordinate stage): out [V,N,2,H,W]. */
void gmpi_oracle_coords(const int32_t* view2mpi, const float* dhw, const float* ray_dir,
                        const float* eye, float* out, int V, int N, int Ht, int Wt, int H, int W,
                        int align_corners) {
    const size_t img = (size_t)H * W;
    for (int v = 0; v < V; ++v) {
        const int m = view2mpi[v];
        for (int i = 0; i < N; ++i) {
            const float* pd = dhw + ((size_t)m * N + i) * 3;
            for (size_t p = 0; p < img; ++p) {
                coord_t c = plane_coord(pd[0], pd[1], pd[2], eye + 3 * v, ray_dir[((size_t)v * 3) * img + p],
                                        ray_dir[((size_t)v * 3 + 1) * img + p],
                                        ray_dir[((size_t)v * 3 + 2) * img + p], Ht, Wt, align_corners);
                out[(((size_t)v * N + i) * 2 + 0) * img + p] = c.ix;
                out[(((size_t)v * N + i) * 2 + 1) * img + p] = c.iy;
            }
        }
    }
}

/* rgba / alpha range checks: mpi_renderer.py:447-449, mpi.py:185-187 */
uint32_t gmpi_oracle_check_range(const float* rgba, size_t n_mpi_planes, int Ht, int Wt) {
    uint32_t flags = 0;
    const size_t tex = (size_t)Ht * Wt;
    for (size_t k = 0; k < n_mpi_planes; ++k)
        for (int ch = 0; ch < 4; ++ch) {
            const float* p = rgba + (k * 4 + ch) * tex;
            for (size_t i = 0; i < tex; ++i)
                if (!(p[i] >= 0.0f) || !(p[i] <= 1.0f)) flags |= (ch == 3 ? (FLAG_ALPHA_RANGE | FLAG_RGBA_RANGE) : FLAG_RGBA_RANGE);
        }
    return flags;
}

int gmpi_oracle_abi_version(void) { return 1; }
