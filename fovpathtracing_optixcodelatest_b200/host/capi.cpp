// capi.cpp -- flat C hooks over the C++ host layer so that tests/ (ctypes) can drive the SAME classes a C++
// application links against: Model / loadOBJ / addBox, ProbeData::BuildCDF, sutil::Camera, SampleRenderer.
// Not part of the drop-in boundary (that is include/fovpt.h + SimplePathtracer.h); test plumbing only.
#include <cstring>
#include <stdexcept>
#include <string>

#include "ImageDecode.h"
#include "MultiGpu.h"
#include "SimplePathtracer.h"

namespace { thread_local std::string g_err; }

#define GUARD(body) try { body; return 0; } catch (const std::exception& e) { g_err = e.what(); return -1; }

extern "C" {

const char* fovpth_last_error() { return g_err.c_str(); }
int fovpth_sizeof(int what) { return what == 0 ? (int)sizeof(LaunchParams) : what == 1 ? (int)sizeof(Material) : what == 2 ? (int)sizeof(Probe) : -1; }

// ---- Model
Model* fovpth_model_new() { return new Model; }
Model* fovpth_load_obj(const char* path) { try { return loadOBJ(path); } catch (const std::exception& e) { g_err = e.what(); return nullptr; } }
void fovpth_model_free(Model* m) { delete m; }
int fovpth_add_box(Model* m, const void* material104, const float* pos, const float* ext)
{
    Material mat; std::memcpy(&mat, material104, sizeof(mat));
    GUARD(addBox(m, mat, make_float3(pos[0], pos[1], pos[2]), make_float3(ext[0], ext[1], ext[2])));
}
// append a raw mesh (python scenes -> C++ Model)
int fovpth_add_mesh(Model* m, const float* vtx, int nv, const unsigned* idx, int nt, const float* tc, const void* material104, int texture_id)
{
    TriangleMesh* mesh = new TriangleMesh;
    for (int i = 0; i < nv; ++i) mesh->vertex.push_back(make_float3(vtx[3 * i], vtx[3 * i + 1], vtx[3 * i + 2]));
    for (int i = 0; i < nt; ++i) mesh->index.push_back(make_uint3(idx[3 * i], idx[3 * i + 1], idx[3 * i + 2]));
    if (tc) for (int i = 0; i < nv; ++i) mesh->texcoord.push_back(make_float2(tc[2 * i], tc[2 * i + 1]));
    std::memcpy(&mesh->material, material104, sizeof(Material));
    mesh->diffuseTextureID = texture_id;
    m->meshes.push_back(mesh);
    return (int)m->meshes.size() - 1;
}
int fovpth_add_texture(Model* m, const uint32_t* rgba8, int w, int h)
{
    Texture* t = new Texture; t->resolution = make_int2(w, h); t->pixel = new uint32_t[(size_t)w * h];
    std::memcpy(t->pixel, rgba8, (size_t)w * h * 4);
    m->textures.push_back(t);
    return (int)m->textures.size() - 1;
}
int fovpth_mesh_count(const Model* m) { return (int)m->meshes.size(); }
int fovpth_mesh_info(const Model* m, int i, int* nv, int* nt, int* ntc, int* tex)
{
    const TriangleMesh* k = m->meshes[i];
    *nv = (int)k->vertex.size(); *nt = (int)k->index.size(); *ntc = (int)k->texcoord.size(); *tex = k->diffuseTextureID;
    return 0;
}
int fovpth_mesh_get(const Model* m, int i, float* vtx, unsigned* idx, float* tc, void* material104)
{
    const TriangleMesh* k = m->meshes[i];
    if (vtx) std::memcpy(vtx, k->vertex.data(), k->vertex.size() * 12);
    if (idx) std::memcpy(idx, k->index.data(), k->index.size() * 12);
    if (tc) std::memcpy(tc, k->texcoord.data(), k->texcoord.size() * 8);
    if (material104) std::memcpy(material104, &k->material, sizeof(Material));
    return 0;
}
int fovpth_texture_count(const Model* m) { return (int)m->textures.size(); }
int fovpth_texture_get(const Model* m, int i, int* w, int* h, uint32_t* out)
{
    const Texture* t = m->textures[i];
    *w = t->resolution.x; *h = t->resolution.y;
    if (out) std::memcpy(out, t->pixel, (size_t)t->resolution.x * t->resolution.y * 4);
    return 0;
}

// decode one image file with the texture decoders; *w,*h always, pixels when `out` has room for them
int fovpth_decode_image(const char* path, int* w, int* h, uint32_t* out, size_t capacity)
{
    std::vector<uint32_t> px; std::string why;
    if (!fovpt_image::decode_file(path, *w, *h, px, &why)) { g_err = why; return -1; }
    if (out && capacity >= px.size()) std::memcpy(out, px.data(), px.size() * 4);
    return 0;
}

// ---- multi-GPU tile partition (MultiGpu.h); owner: tiles_y*tiles_x bytes, loads: n_ranks doubles; returns tiles_x | tiles_y << 16
int fovpth_partition(int w, int h, unsigned gx, unsigned gy, const fovpt_config* cfg, int n_ranks, int tile, int interleave, uint8_t* owner, double* loads)
{
    try {
        const fovpt_multigpu::TilePartition p = interleave ? fovpt_multigpu::interleaved(w, h, n_ranks, tile)
                                                           : fovpt_multigpu::partition(w, h, make_uint2(gx, gy), *cfg, n_ranks, tile);
        if (owner) std::memcpy(owner, p.owner.data(), p.owner.size());
        if (loads) {
            const std::vector<double> l = interleave ? fovpt_multigpu::loads_of(p, w, h, make_uint2(gx, gy), *cfg, n_ranks) : p.loads;
            std::memcpy(loads, l.data(), l.size() * sizeof(double));
        }
        return p.tiles_x | (p.tiles_y << 16);
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// ---- ProbeData / Camera
int fovpth_build_cdf(int w, int h, const float* rgba, float* pdfX, float* cdfX, float* pdfY, float* cdfY)
{
    ProbeData p; p.width = w; p.height = h; p.data = (Color*)rgba;
    p.BuildCDF();
    std::memcpy(pdfX, p.pdfValuesX, (size_t)w * h * 4); std::memcpy(cdfX, p.cdfValuesX, (size_t)w * h * 4);
    std::memcpy(pdfY, p.pdfValuesY, (size_t)h * 4); std::memcpy(cdfY, p.cdfValuesY, (size_t)h * 4);
    delete[] p.pdfValuesX; delete[] p.cdfValuesX; delete[] p.pdfValuesY; delete[] p.cdfValuesY;
    return 0;
}
int fovpth_camera_uvw(const float* eye, const float* lookat, const float* up, float fovy, float aspect, float* out9)
{
    sutil::Camera c(make_float3(eye[0], eye[1], eye[2]), make_float3(lookat[0], lookat[1], lookat[2]), make_float3(up[0], up[1], up[2]), fovy, aspect);
    float3 U, V, W; c.UVWFrame(U, V, W);
    const float o[9] = {U.x, U.y, U.z, V.x, V.y, V.z, W.x, W.y, W.z};
    std::memcpy(out9, o, sizeof(o));
    return 0;
}

// ---- SampleRenderer (needs a GPU)
SampleRenderer* fovpth_renderer_new(const Model* m, int device, const fovpt_config* cfg)
{
    try { return new SampleRenderer(m, device, cfg); } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void fovpth_renderer_free(SampleRenderer* r) { delete r; }
int fovpth_resize(SampleRenderer* r, int w, int h) { GUARD(r->resize(make_int2(w, h))); }
int fovpth_set_camera(SampleRenderer* r, const float* eye, const float* lookat, const float* up, float fovy)
{
    GUARD(r->setCamera(sutil::Camera(make_float3(eye[0], eye[1], eye[2]), make_float3(lookat[0], lookat[1], lookat[2]), make_float3(up[0], up[1], up[2]), fovy, 1.0f)));
}
int fovpth_set_probe(SampleRenderer* r, int w, int h, const float* rgba, int build_cdf)
{
    ProbeData p; p.width = w; p.height = h; p.data = (Color*)rgba;
    if (build_cdf) p.BuildCDF();
    int rc = 0;
    try { r->setProbe(p); } catch (const std::exception& e) { g_err = e.what(); rc = -1; }
    if (build_cdf) { delete[] p.pdfValuesX; delete[] p.cdfValuesX; delete[] p.pdfValuesY; delete[] p.cdfValuesY; }
    return rc;
}
void* fovpth_launch_params(SampleRenderer* r) { return &r->launchParams; }      // the public member, as main.cpp uses it
int fovpth_render(SampleRenderer* r) { GUARD(r->render()); }
int fovpth_render_subframes(SampleRenderer* r, int n) { GUARD(r->renderSubframes(n)); }
// type: sutil::CUDAOutputBufferType (0 CUDA_DEVICE, 1 GL_INTEROP -- throws in this headless build, 2 ZERO_COPY, 3 CUDA_P2P)
int fovpth_render_to_output_buffer(SampleRenderer* r, int type, uint32_t* host_out)
{
    GUARD({
        sutil::CUDAOutputBuffer<uint32_t> target((sutil::CUDAOutputBufferType)type, r->launchParams.frame.size.x, r->launchParams.frame.size.y);
        target.setStream(r->stream);
        r->render(target);
        std::memcpy(host_out, target.getHostPointer(), (size_t)target.width() * target.height() * 4);
    });
}
int fovpth_download_pixels(SampleRenderer* r, uint32_t* out) { GUARD(r->downloadPixels(out)); }
int fovpth_download_accum(SampleRenderer* r, float* out) { GUARD(r->downloadAccum((float4*)out)); }
int fovpth_save_checkpoint(SampleRenderer* r, const char* path) { GUARD(r->saveCheckpoint(path)); }
int fovpth_load_checkpoint(SampleRenderer* r, const char* path) { GUARD(r->loadCheckpoint(path)); }

}  // extern "C"
