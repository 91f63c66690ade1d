/* tests/c_abi/abi_from_c.c -- the C ABI used from plain C (C11, no C++ and no Python in between), the way R's .C() or any other
 * FFI binds it: include/nngp_b200.h compiles as C, the library links with a C linker line, every argument is a pointer and
 * every call reports through *status.  Only host-side entry points are called, so this runs without a GPU
 * (tests/test_abi_cpu.py::test_abi_links_and_runs_from_plain_c compiles and runs it).
 * Prints one line per check; exit code 0 = all passed. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "nngp_b200.h"

static int failures = 0;
#define CHECK(cond, what) do { if (cond) printf("ok   %s\n", what); else { printf("FAIL %s\n", what); failures++; } } while (0)

int main(void)
{
    int major = -1, minor = -1, status = -1, count = -1;
    nngp_version(&major, &minor);
    CHECK(major >= 0 && minor >= 0, "nngp_version");
    nngp_device_count(&count, &status);
    CHECK(count >= 0, "nngp_device_count reports a count (0 without a GPU)");

    /* R's stream: set.seed(1); runif(3) = 0.2655087 0.3721239 0.5728534 (any R session; Vignette.md:137-139 prints 500 x these) */
    int rstate[625], seed = 1, n3 = 3;
    double u[3];
    nngp_rng_set_seed(&seed, rstate, &status);
    CHECK(status == NNGP_OK, "nngp_rng_set_seed");
    nngp_rng_runif(rstate, &n3, u, &status);
    CHECK(status == NNGP_OK && fabs(u[0] - 0.2655087) < 5e-8 && fabs(u[1] - 0.3721239) < 5e-8 && fabs(u[2] - 0.5728534) < 5e-8,
          "nngp_rng_runif reproduces R's set.seed(1); runif(3)");

    /* ordered nearest neighbours of 5 points on a line, m = 2: column-major n x (m + 1), 1-based, NA = INT_MIN */
    double locs[10] = {0.0, 10.0, 1.0, 9.0, 0.4, /* second coordinate */ 0, 0, 0, 0, 0};
    int n = 5, d = 2, m = 2, nn[15];
    nngp_host_find_ordered_nn(locs, &n, &d, &m, nn, &status);
    const int want[15] = {1, 2, 3, 4, 5, /* nearest previous */ NNGP_NA_INT, 1, 1, 2, 1, /* second nearest */ NNGP_NA_INT, NNGP_NA_INT, 2, 3, 3};
    CHECK(status == NNGP_OK && memcmp(nn, want, sizeof want) == 0, "nngp_host_find_ordered_nn (layout, 1-based indices, NA padding)");

    int coloring[5], n_colors = 0;
    nngp_host_greedy_coloring(nn, &n, &m, coloring, &n_colors, &status);
    int proper = status == NNGP_OK && n_colors >= 3;   /* every row of 3 sites is a clique of the moral graph */
    for (int i = 0; i < n && proper; i++)
        for (int a = 0; a <= m; a++)
            for (int b = a + 1; b <= m; b++) {
                int s = nn[i + n * a], t = nn[i + n * b];
                if (s != NNGP_NA_INT && t != NNGP_NA_INT && coloring[s - 1] == coloring[t - 1]) proper = 0;
            }
    CHECK(proper, "nngp_host_greedy_coloring (proper colouring of the moral graph)");

    /* errors come back as a status and a message, never as a crash */
    int bad_n = -1;
    nngp_host_find_ordered_nn(locs, &bad_n, &d, &m, nn, &status);
    char msg[256];
    int len = (int)sizeof msg;
    nngp_last_error(msg, &len);
    CHECK(status == NNGP_ERR_ARG && strstr(msg, "nngp_host_find_ordered_nn") != NULL, "bad argument -> NNGP_ERR_ARG + nngp_last_error");

    /* no CPU fallback: without a device a compute entry point must refuse */
    if (count == 0) {
        int ctx = -1, n_obs = 5, covfun = 0, device = 0, layout = NNGP_LAYOUT_MORTON, lm[5] = {1, 2, 3, 4, 5};
        nngp_ctx_create(&n, &d, &m, locs, nn, coloring, &n_obs, lm, &covfun, &device, &layout, &ctx, &status);
        nngp_last_error(msg, &len);
        CHECK(status == NNGP_ERR_CUDA && strstr(msg, "no CPU fallback") != NULL, "no GPU -> nngp_ctx_create refuses (NNGP_ERR_CUDA, no CPU fallback)");
    }
    return failures ? 1 : 0;
}
