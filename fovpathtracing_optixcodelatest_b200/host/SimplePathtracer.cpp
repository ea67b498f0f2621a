// SimplePathtracer.cpp -- SampleRenderer on top of the libfovpt C ABI (see SimplePathtracer.h).
#include "SimplePathtracer.h"

#include <cstdio>
#include <cstring>
#include <iostream>
#include <vector>
#include <stdexcept>

void SampleRenderer::check(int rc)
{
    if (rc != 0) throw std::runtime_error(std::string("libfovpt: ") + (ctx ? fovpt_last_error(ctx) : "context creation failed"));
}

SampleRenderer::SampleRenderer(const Model* model_) : SampleRenderer(model_, 0, nullptr) {}

SampleRenderer::SampleRenderer(const Model* model_, int device, const fovpt_config* cfg) : model(model_)
{
    std::memset(&launchParams, 0, sizeof(launchParams));          // the reference leaves it uninitialised
    launchParams.white = 1.0f;
    const int rc = fovpt_create(&ctx, device, nullptr);
    if (rc != 0) throw std::runtime_error("#osc: no CUDA capable devices found!");            // SimplePathtracer.cpp:320-321
    stream = (cudaStream_t)fovpt_stream(ctx);
    if (cfg) config = *cfg; else fovpt_default_config(&config);
    check(fovpt_set_config(ctx, &config));

    // buildAccel + createTextures + buildSBT (SimplePathtracer.cpp:602-799, 534-599): hand the meshes over,
    // libfovpt flattens them and builds the BVH on the device.
    check(fovpt_scene_begin(ctx));
    for (const Texture* t : model->textures) {
        if (fovpt_scene_add_texture(ctx, t->pixel, t->resolution.x, t->resolution.y) < 0) check(1);
    }
    for (const TriangleMesh* m : model->meshes) {
        const float* tc = m->texcoord.size() == m->vertex.size() && !m->texcoord.empty() ? &m->texcoord[0].x : nullptr;
        // a texture id that does not name a loaded texture (e.g. addBox's default 1, Model.h:19) means "no texture"
        const int tex = (m->diffuseTextureID >= 0 && m->diffuseTextureID < (int)model->textures.size()) ? m->diffuseTextureID : -1;
        const float* vtx = m->vertex.empty() ? nullptr : &m->vertex[0].x;
        const uint32_t* idx = m->index.empty() ? nullptr : &m->index[0].x;
        static const float dummy_v[3] = {0, 0, 0}; static const uint32_t dummy_i[3] = {0, 0, 0};
        if (fovpt_scene_add_mesh(ctx, vtx ? vtx : dummy_v, m->vertex.size(), idx ? idx : dummy_i, m->index.size(), tc, &m->material, tex) < 0) check(1);
    }
    unsigned long long handle = 0;
    check(fovpt_scene_commit(ctx, (uint64_t*)&handle));
    launchParams.traversable = handle;
}

SampleRenderer::~SampleRenderer() { fovpt_destroy(ctx); }

void SampleRenderer::setConfig(const fovpt_config& cfg)
{
    config = cfg;
    check(fovpt_set_config(ctx, &config));
}

void SampleRenderer::render()
{
    if (launchParams.frame.size.x == 0) return;                   // SimplePathtracer.cpp:81-82
    check(fovpt_render_frame(ctx, &launchParams));                 // all passes in one submission; advances subframe_index (:210)
    check(fovpt_sync(ctx));                                        // CUDA_SYNC_CHECK (:212)
}

void SampleRenderer::renderSubframes(int n)
{
    if (launchParams.frame.size.x == 0) return;
    check(fovpt_render_frames(ctx, &launchParams, n));
    check(fovpt_sync(ctx));
}

void SampleRenderer::render(sutil::CUDAOutputBuffer<uint32_t>& renderTarget)
{
    launchParams.frame.frame_buffer = (uchar4*)renderTarget.map();  // SimplePathtracer.cpp:218-219
    render();
    renderTarget.unmap();
    // The reference leaves frame_buffer pointing at the display buffer, which its main loop maps again every frame
    // (main.cpp:417-423).  Here the target may be gone by the next render() / downloadPixels(): fall back to the
    // buffer resize() allocated, and let the context forget the borrowed pointer.
    launchParams.frame.frame_buffer = ownFrameBuffer;
    fovpt_forget_frame(ctx);
}

void SampleRenderer::resize(const int2& newSize)
{
    if (newSize.x == 0 || newSize.y == 0) return;                  // window minimised (:231)
    void *accum = nullptr, *frame = nullptr;
    check(fovpt_resize(ctx, newSize.x, newSize.y, &accum, &frame));
    launchParams.frame.size = newSize;
    launchParams.frame.accum_buffer = (float4*)accum;
    launchParams.frame.frame_buffer = (uchar4*)frame;
    ownFrameBuffer = (uchar4*)frame;
    // the reference allocates its three denoiser buffers here (:245-260) and never writes them (deviceProgram.cu:612-614 are
    // commented out); here they exist only when fovpt_config.features asks for those stores
    launchParams.frame.color_buffer = launchParams.frame.normal_buffer = launchParams.frame.albedo_buffer = nullptr;
    fovpt_config k;
    check(fovpt_get_config(ctx, &k));
    if (k.features) {
        void *col = nullptr, *nrm = nullptr, *alb = nullptr;
        check(fovpt_feature_buffers(ctx, &col, &nrm, &alb));
        launchParams.frame.color_buffer = (float4*)col; launchParams.frame.normal_buffer = (float4*)nrm; launchParams.frame.albedo_buffer = (float4*)alb;
    }
    launchParams.viewportSize = newSize;
}

void SampleRenderer::downloadPixels(uint32_t h_pixels[])
{
    check(fovpt_download(ctx, h_pixels, (size_t)launchParams.frame.size.x * launchParams.frame.size.y));
}

void SampleRenderer::downloadAccum(float4 h_accum[])
{
    check(fovpt_download_accum(ctx, &h_accum[0].x, (size_t)launchParams.frame.size.x * launchParams.frame.size.y));
}

void SampleRenderer::saveCheckpoint(const std::string& path)
{
    const int w = launchParams.frame.size.x, h = launchParams.frame.size.y;
    std::vector<float4> accum((size_t)w * h);
    downloadAccum(accum.data());
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write checkpoint " + path);
    const int32_t head[3] = {w, h, (int32_t)launchParams.frame.subframe_index};
    const bool ok = std::fwrite("FOVPTCK1", 1, 8, f) == 8 && std::fwrite(head, sizeof(head), 1, f) == 1 &&
                    std::fwrite(accum.data(), sizeof(float4), accum.size(), f) == accum.size();
    std::fclose(f);
    if (!ok) throw std::runtime_error("short write to checkpoint " + path);
}

void SampleRenderer::loadCheckpoint(const std::string& path)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot read checkpoint " + path);
    char magic[8]; int32_t head[3] = {0, 0, 0};
    bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "FOVPTCK1", 8) == 0 && std::fread(head, sizeof(head), 1, f) == 1;
    if (ok && (head[0] != launchParams.frame.size.x || head[1] != launchParams.frame.size.y)) { std::fclose(f); throw std::runtime_error("checkpoint was written for another frame size"); }
    std::vector<float4> accum(ok ? (size_t)head[0] * head[1] : 0);
    ok = ok && std::fread(accum.data(), sizeof(float4), accum.size(), f) == accum.size();
    std::fclose(f);
    if (!ok) throw std::runtime_error("not a checkpoint file: " + path);
    check(fovpt_upload_accum(ctx, &accum[0].x, accum.size()));
    launchParams.frame.subframe_index = (unsigned)head[2];
}

void SampleRenderer::setCamera(const sutil::Camera& camera)
{
    if (launchParams.frame.size.y == 0) throw std::runtime_error("setCamera before resize: frame.size is 0");
    lastSetCamera = camera;
    lastSetCamera.setAspectRatio(launchParams.frame.size.x / float(launchParams.frame.size.y));        // :286
    lastSetCamera.UVWFrame(launchParams.camera.U, launchParams.camera.V, launchParams.camera.W);
    launchParams.camera.eye = lastSetCamera.eye();
}

void SampleRenderer::setProbe(const ProbeData& probe)
{
    if (!probe.valid) throw std::runtime_error("Probe Data is not valid");                             // Probe.h:104-105
    const float off[3] = {probe.offset.x, probe.offset.y, probe.offset.z};
    check(fovpt_probe_upload(ctx, probe.width, probe.height, &probe.data[0].x, probe.pdfValuesX, probe.cdfValuesX,
                             probe.pdfValuesY, probe.cdfValuesY, off, &launchParams.probe));
}

fovpt_stats SampleRenderer::stats()
{
    fovpt_stats s;
    check(fovpt_get_stats(ctx, &s));
    return s;
}
