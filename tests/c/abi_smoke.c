/* abi_smoke.c -- include/fovpt.h is a plain C header and libfovpt.so a plain C library: this C99 program (no C++,
 * no CUDA headers) builds a one-triangle scene through the ABI and traces two rays.  Without a CUDA device it checks
 * that context creation fails cleanly instead (exit code 0 in both cases, 1 on any contract violation). */
#include <stdio.h>
#include <string.h>

#include "fovpt.h"

int main(void)
{
    fovpt_config cfg;
    fovpt_default_config(&cfg);
    if (cfg.max_depth != 4 || cfg.inner_radius != 74 || cfg.outer_radius != 241 || cfg.spp_fovea != 32) { printf("bad defaults\n"); return 1; }
    fovpt_ctx* ctx = NULL;
    int rc = fovpt_create(&ctx, 0, NULL);
    if (rc != 0) {
        if (ctx != NULL) { printf("failed create must leave ctx NULL\n"); return 1; }
        printf("no CUDA device: fovpt_create failed cleanly (rc %d)\n", rc);
        return 0;
    }
    const float vtx[9] = {-1.f, -1.f, 0.f, 1.f, -1.f, 0.f, 0.f, 1.f, 0.f};
    const uint32_t idx[3] = {0, 1, 2};
    unsigned char material[104];
    memset(material, 0, sizeof(material));
    uint64_t handle = 0;
    if (fovpt_scene_begin(ctx) || fovpt_scene_add_mesh(ctx, vtx, 3, idx, 1, NULL, material, -1) != 0 || fovpt_scene_commit(ctx, &handle) || !handle) {
        printf("scene: %s\n", fovpt_last_error(ctx)); return 1;
    }
    const uint32_t bad[3] = {0, 1, 7};
    if (fovpt_scene_add_mesh(ctx, vtx, 3, bad, 1, NULL, material, -1) >= 0) { printf("out-of-range index accepted\n"); return 1; }
    const float rays[16] = {0.f, 0.f, -5.f, 0.01f, 0.f, 0.f, 1.f, 1e16f, 5.f, 5.f, -5.f, 0.01f, 0.f, 0.f, 1.f, 1e16f};
    int32_t tri[2]; float tuv[6];
    if (fovpt_trace(ctx, rays, 2, 0, tri, tuv)) { printf("trace: %s\n", fovpt_last_error(ctx)); return 1; }
    if (tri[0] != 0 || tri[1] != -1 || tuv[0] != 5.0f) { printf("wrong hits %d %d t=%g\n", tri[0], tri[1], tuv[0]); return 1; }
    fovpt_destroy(ctx);
    printf("C ABI ok: hit triangle %d at t=%g, miss %d\n", tri[0], tuv[0], tri[1]);
    return 0;
}
