/* oracle/orc_api.h -- C interface shared by the two CPU checkers.  TEST INFRASTRUCTURE ONLY:
 * nothing under oracle/ is linked into, imported by, or called from the product path
 * (csrc/, host/, the python package); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs load these libraries.
 *
 *   oracle/liboracle.so        kind "port"      : the restatement in oracle/fovpt_oracle.cpp
 *   oracle/_ref/libfovref.so   kind "reference" : the reference's own deviceProgram.cu, Disney.cuh,
 *                                                 Probe.cuh, maths.h, sample.h, random.h, helpers.h,
 *                                                 Probe.h, Model.cpp, sutil/Camera.cpp compiled
 *                                                 UNMODIFIED for the host (oracle/Makefile)
 * Both export exactly these symbols, so one python wrapper drives either.
 *
 * Launch parameters cross this interface as the reference's 248-byte LaunchParams blob
 * (PT_sv5_/LaunchParams.h:49-91); buffer pointers inside it are HOST pointers here.
 */
#ifndef ORC_API_H
#define ORC_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

typedef struct orc_opts {
    int32_t  max_depth;     /* break when depth >= max_depth (deviceProgram.cu:515; sv5 literal 4)   */
    float    tmin, tmax;    /* deviceProgram.cu:41-42 : 0.01, 1e16                                   */
    int32_t  jitter_yx;     /* 1: draw jitter.y before jitter.x (g++'s right-to-left evaluation of    *
                             * deviceProgram.cu:479); 0: x first (device order, the product's order)  */
    int32_t  threads;       /* 0 = all hardware threads                                              */
    int32_t  accumulate;    /* 0: sv5 (lerp commented out, :565-581); 1: vmv2 lerp                    *
                             *    (PT_sv4_vmv2/deviceProgram.cu:545-553) -- port only                */
    float    exposure;      /* deviceProgram.cu:586 : 16                                             */
    int32_t  seed_all_passes_with_subframe;  /* 0: sv5; 1: port-only extension used by C2/C3        */
    /* optional captures, one entry per launch index (row-major lw*lh); may be NULL */
    int32_t* primary_tri;   /* global triangle id of sample 0's primary hit, -1 miss, -2 ring-culled */
    float*   primary_tuv;   /* 3 floats: t,u,v                                                       */
    float*   primary_dir;   /* 3 floats: the sample-0 camera direction                               */
    /* port only (the reference ships with both disabled / empty; oracle/_ref ignores these fields) */
    int32_t  features;      /* 1: write frame.normal_buffer / color_buffer / albedo_buffer as the commented-out  *
                             * statements deviceProgram.cu:612-614 would (first-hit normal and albedo, :509-512,  *
                             * :541-542, :653-665); 0: sv5 as shipped                                             */
    int32_t  filter;        /* orc_render only.  1: bilinear reconstruction of the sub-sampled rings instead of   *
                             * the fillSize^2 box splat (the reference's Filter.h is an empty class; definition    *
                             * in DESIGN.md section 3); 0: box splat                                               */
} orc_opts;

typedef struct orc_fov_cfg {     /* three-pass parameters, SimplePathtracer.cpp:20-21,137-209 */
    int32_t  foveated;           /* 1: FOV_ON three passes ; 0: FOV_OFF single full-frame pass */
    int32_t  inner_radius, outer_radius;   /* 74, 241 */
    int32_t  spp_periphery, spp_mid, spp_fovea;   /* 8, 16, 32 */
    int32_t  spp_uniform;        /* FOV_OFF samples_per_launch (4 in sv5) */
} orc_fov_cfg;

const char* orc_kind(void);
orc_ctx*    orc_create(void);
void        orc_destroy(orc_ctx*);
const char* orc_last_error(orc_ctx*);

/* scene ------------------------------------------------------------------------------ */
int orc_add_texture(orc_ctx*, const uint32_t* rgba8, int w, int h);
int orc_add_mesh(orc_ctx*, const float* vtx, int nvtx, const uint32_t* idx, int ntri,
                 const float* texcoord_or_null, const void* material104, int texture_id);
int orc_load_obj(orc_ctx*, const char* path);          /* loadOBJ, Model.cpp:138-217 */
int orc_add_box(orc_ctx*, const void* material104, const float pos[3], const float ext[3]); /* addBox :219-291 */
int orc_mesh_count(orc_ctx*);
int orc_mesh_info(orc_ctx*, int mesh, int* nvtx, int* ntri, int* ntexcoord, int* texture_id);
int orc_mesh_get(orc_ctx*, int mesh, float* vtx, uint32_t* idx, float* texcoord, void* material104);
int orc_texture_count(orc_ctx*);
int orc_texture_info(orc_ctx*, int tex, int* w, int* h);
int orc_texture_get(orc_ctx*, int tex, uint32_t* rgba8);
int orc_commit(orc_ctx*);                               /* build the canonical BVH */

/* probe / camera --------------------------------------------------------------------- */
int orc_set_probe(orc_ctx*, int w, int h, const float* rgba, const float offset[3]); /* BuildCDF, Probe.h:29-77 */
int orc_get_probe(orc_ctx*, float* pdfX, float* cdfX, float* pdfY, float* cdfY);
int orc_camera_uvw(const float eye[3], const float lookat[3], const float up[3], float fovy, float aspect,
                   float uvw9[9]);                       /* sutil/Camera.cpp:34-46 */

/* rendering -------------------------------------------------------------------------- */
int orc_launch(orc_ctx*, void* launch_params248, unsigned lw, unsigned lh, const orc_opts*);
int orc_render(orc_ctx*, void* launch_params248, const orc_fov_cfg*, const orc_opts*);
/* counters: [0] radiance rays [1] shadow rays [2] closest node visits [3] closest tri tests
 *           [4] any-hit node visits [5] any-hit tri tests [6] samples */
void orc_get_counters(orc_ctx*, uint64_t out[8]);
void orc_reset_counters(orc_ctx*);

/* ray queries: rays = n x {ox,oy,oz,tmin, dx,dy,dz,tmax}; mode 0 closest, 1 any(front faces),
 * 2 closest brute force, 3 any brute force.  tri = global id / -1 (closest) or 1/0 (any). */
int orc_trace(orc_ctx*, const float* rays, int n, int mode, int32_t* tri, float* tuv);

/* unit hooks (pin the restatement function by function) -------------------------------- */
uint32_t orc_tea4(uint32_t a, uint32_t b);
void     orc_rnd_seq(uint32_t seed, int n, float* out, uint32_t* seed_out);
void     orc_random_seq(uint32_t seed, int n, uint32_t* u32_out, float* f_out);
/* BasisFromVector(N) + BSDFSample(Random(seed)) + BSDFEval + BSDFPdf on the sampled direction.
 * out10 = L.xyz, pdf, type, f.xyz, pdf_again, draws_consumed(as float) */
void     orc_bsdf_sample(const void* material104, const float albedo[3], float etaI, float etaO,
                         const float N[3], const float V[3], uint32_t seed, float out10[10]);
/* BSDFEval + BSDFPdf for a given L. out4 = f.xyz, pdf */
void     orc_bsdf_eval(const void* material104, const float albedo[3], float etaI, float etaO,
                       const float N[3], const float V[3], const float L[3], float out4[4]);
/* ProbeSample(Random(seed)) -> out7 = dir.xyz, color.xyz, pdf */
int      orc_probe_sample(orc_ctx*, uint32_t seed, float out7[7]);
/* ProbeDirToUV + ProbeEval -> out6 = u, v, rgba */
int      orc_probe_eval(orc_ctx*, const float dir[3], float out6[6]);
/* exposure * reinhard(white=1) * make_color -> packed RGBA8 (deviceProgram.cu:586,597) */
uint32_t orc_tonemap(const float accum[3], float exposure);

#ifdef __cplusplus
}
#endif
#endif
