/*
 * A host that is NOT Python: plain C99 against include/gmpi_mpi_render.h, nothing else (the library links the CUDA runtime
 * statically; the program needs neither the toolkit nor torch).  It renders a small MPI from HOST memory through
 * gmpi_mpi_render_fwd_host -- what `MPIRenderer.render` (gmpi/core/mpi_renderer.py:387-469) does for one view -- and checks
 * the result against closed-form answers of the reference's compositing rule (gmpi/core/mpi.py:411-436):
 *
 *   plane 0 opaque (the reference's own sanity mode, eval/prepare_fake_data.py:51-56)
 *       => colour = plane 0's colour, depth = plane 0's distance;
 *   constant alpha a on every plane, colour c_i per plane, identity pose
 *       => colour = sum_i a (1-a)^i c_i, depth = sum_i a (1-a)^i d_i        (T_i = prod_{j<i} (1 - a + 1e-10)).
 *
 *   gcc -std=c99 -Iinclude examples/host_render.c -o host_render -Lml_gmpi_b200 -lgmpi_mpi_render -Wl,-rpath,$PWD/ml_gmpi_b200 -lm
 *   ./host_render [device]          exit 0: rendered and verified; 2: wrong result; 3: the library reported an error
 *                                   (printed; e.g. no CUDA device)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "gmpi_mpi_render.h"

enum { N = 6, T = 32, S = 24 };                 /* planes, texture size, image size */

static float* falloc(size_t n) {
    float* p = (float*)calloc(n, sizeof(float));
    if (!p) { fprintf(stderr, "out of memory\n"); exit(1); }
    return p;
}

/* identity pose at the origin looking down +Z; pinhole rays through pixel centres, fov such that every ray stays on every plane */
static void make_rays(float* ray_dir) {
    const double focal = S / (2.0 * tan(0.5 * 12.6 * 3.14159265358979323846 / 180.0));
    for (int y = 0; y < S; ++y)
        for (int x = 0; x < S; ++x) {
            const double dx = (x + 0.5 - S / 2.0) / focal, dy = (y + 0.5 - S / 2.0) / focal;
            const double n = sqrt(dx * dx + dy * dy + 1.0);
            ray_dir[0 * S * S + y * S + x] = (float)(dx / n);
            ray_dir[1 * S * S + y * S + x] = (float)(dy / n);
            ray_dir[2 * S * S + y * S + x] = (float)(1.0 / n);
        }
}

static int render(const float* rgba, const float* dhw, const float* ray_dir, float* color, float* depth, int device) {
    const int32_t view2mpi[1] = {0};
    const float eye[3] = {0.f, 0.f, 0.f}, z_dir[3] = {0.f, 0.f, 1.f};
    uint32_t flags = 0;
    const int rc = gmpi_mpi_render_fwd_host(rgba, view2mpi, dhw, ray_dir, eye, z_dir, color, depth, &flags, 1, 1, N, T, T, S, S,
                                            GMPI_ALIGN_CORNERS | GMPI_CHECK_LAST_PLANE, device);
    if (rc != GMPI_OK) {
        fprintf(stderr, "gmpi_mpi_render_fwd_host failed (%d): %s\n", rc, gmpi_last_error());
        return 3;
    }
    if (flags != 0) {
        fprintf(stderr, "unexpected flag word 0x%x\n", (unsigned)flags);
        return 2;
    }
    return 0;
}

static int check(const char* what, const float* got, size_t n, float want, float tol) {
    double worst = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double e = fabs((double)got[i] - want);
        if (e > worst) worst = e;
    }
    printf("%-28s expected %.7f, max abs error %.3g\n", what, want, worst);
    return worst <= tol ? 0 : 2;
}

int main(int argc, char** argv) {
    const int device = argc > 1 ? atoi(argv[1]) : 0;
    const size_t tex = (size_t)T * T, img = (size_t)S * S;
    float* rgba = falloc((size_t)N * 4 * tex);
    float* dhw = falloc((size_t)N * 3);
    float* ray_dir = falloc(3 * img);
    float* color = falloc(3 * img);
    float* depth = falloc(img);
    float plane_rgb[N][3], d[N];
    int rc = 0;

    printf("gmpi ABI version %d\n", gmpi_abi_version());
    make_rays(ray_dir);
    for (int i = 0; i < N; ++i) {                                         /* planes near -> far, 0.5 x 0.5 metric size */
        d[i] = 0.95f + 0.034f * (float)i;
        dhw[3 * i] = d[i]; dhw[3 * i + 1] = 0.5f; dhw[3 * i + 2] = 0.5f;
        for (int c = 0; c < 3; ++c) plane_rgb[i][c] = 0.1f + 0.13f * (float)i + 0.05f * (float)c;
    }

    /* 1. plane 0 opaque */
    for (int i = 0; i < N; ++i)
        for (int c = 0; c < 4; ++c)
            for (size_t t = 0; t < tex; ++t) rgba[((size_t)i * 4 + c) * tex + t] = c < 3 ? plane_rgb[i][c] : 1.0f;
    if ((rc = render(rgba, dhw, ray_dir, color, depth, device)) != 0) return rc;
    for (int c = 0; c < 3; ++c) rc |= check(c == 0 ? "opaque: red" : c == 1 ? "opaque: green" : "opaque: blue", color + c * img, img, plane_rgb[0][c], 2e-6f);
    rc |= check("opaque: depth", depth, img, d[0], 2e-6f);

    /* 2. constant alpha 0.3 on every plane */
    {
        const double a = 0.3;
        double want[4] = {0, 0, 0, 0}, trans = 1.0;
        for (int i = 0; i < N; ++i) {
            for (int c = 0; c < 3; ++c) want[c] += a * trans * plane_rgb[i][c];
            want[3] += a * trans * d[i];
            trans *= 1.0 - a + 1e-10;
        }
        for (int i = 0; i < N; ++i)
            for (size_t t = 0; t < tex; ++t) rgba[((size_t)i * 4 + 3) * tex + t] = (float)a;
        const int r = render(rgba, dhw, ray_dir, color, depth, device);
        if (r == 3) return 3;
        rc |= r;
        for (int c = 0; c < 3; ++c) rc |= check(c == 0 ? "alpha 0.3: red" : c == 1 ? "alpha 0.3: green" : "alpha 0.3: blue", color + c * img, img, (float)want[c], 2e-6f);
        rc |= check("alpha 0.3: depth", depth, img, (float)want[3], 2e-6f);
    }
    gmpi_mpi_release_host_cache();
    free(rgba); free(dhw); free(ray_dir); free(color); free(depth);
    printf(rc == 0 ? "OK\n" : "MISMATCH\n");
    return rc == 0 ? 0 : 2;
}
