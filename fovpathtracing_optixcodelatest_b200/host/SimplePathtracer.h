// SimplePathtracer.h -- class SampleRenderer, the drop-in boundary of the foveated path tracer.
// Same public surface as the reference's class (PT_sv5_/SimplePathtracer.h:45-189): constructor from
// a Model, render(), render(CUDAOutputBuffer&), resize, downloadPixels, setCamera, setProbe, and the
// PUBLIC members the reference's main.cpp pokes directly -- launchParams (gaze frame.c, subframe_index,
// frame.size; main.cpp:323,366-367,404,435,479) and stream (main.cpp:338).
//
// Everything the reference did through OptiX (context, module, program groups, pipeline, SBT, accel
// build, the three optixLaunch calls) happens behind the C ABI of libfovpt (include/fovpt.h); this
// class holds no GPU code.  Errors surface as std::runtime_error, the reference's convention
// (sutil::Exception derives from it, sutil/Exception.h:93-194).
#pragma once
#include <cuda_runtime.h>
#include <string>

#include "../../include/fovpt.h"
#include "CUDAOutputBuffer.h"
#include "Camera.h"
#include "LaunchParams.h"
#include "Model.h"
#include "Probe.h"

class SampleRenderer {
public:
    /*! performs all setup: context, scene upload, on-device BVH build (reference ctor, SimplePathtracer.cpp:42-75).
        The model is borrowed for the renderer's lifetime, as in the reference. */
    SampleRenderer(const Model* model);
    /*! extension: choose the CUDA device / pass non-default settings (the reference hard-codes both) */
    SampleRenderer(const Model* model, int device, const fovpt_config* config);
    ~SampleRenderer();
    SampleRenderer(const SampleRenderer&) = delete;
    SampleRenderer& operator=(const SampleRenderer&) = delete;

    /*! render one frame: the three foveation passes (or the single FOV_OFF pass), then a device sync
        (SimplePathtracer.cpp:77-214).  A no-op until resize() has been called (:81-82). */
    void render();
    /*! extension: n successive render() calls of a progressive loop (subframe_index advancing, main.cpp:402-423) as ONE
        submission -- the subframes share one wavefront (fovpt_render_frames); accumulation and frame buffer end up
        bit-identical to calling render() n times */
    void renderSubframes(int n);
    /*! render into an external display buffer (SimplePathtracer.cpp:216-226) */
    void render(sutil::CUDAOutputBuffer<uint32_t>& renderTarget);
    /*! resize the frame buffers; drops the accumulation (SimplePathtracer.cpp:228-274) */
    void resize(const int2& newSize);
    /*! blocking download of frame.size.x * frame.size.y RGBA8 pixels (SimplePathtracer.cpp:276-280) */
    void downloadPixels(uint32_t h_pixels[]);
    /*! float4 accumulation buffer (extension, for parity tests) */
    void downloadAccum(float4 h_accum[]);
    /*! extension: checkpoint / resume of a progressive render.  The accumulation buffer and frame.subframe_index are the only
        persistent state of the path (the reference never saves them); a run that loads a checkpoint continues
        bit-identically.  File: "FOVPTCK1", width, height, subframe_index (int32 each), then width*height float4. */
    void saveCheckpoint(const std::string& path);
    void loadCheckpoint(const std::string& path);
    /*! aspect ratio is taken from frame.size, so call after resize (SimplePathtracer.cpp:282-289) */
    void setCamera(const sutil::Camera& camera);
    /*! uploads data + CDF tables; throws "Probe Data is not valid" before BuildCDF (Probe.h:102-124) */
    void setProbe(const ProbeData& probe);

    /*! runtime form of the reference's compile-time switches (FOV_ON/OFF, radii, spp, depth ...) */
    void setConfig(const fovpt_config& cfg);
    const fovpt_config& getConfig() const { return config; }
    fovpt_stats stats();
    /*! the libfovpt context, for the entry points of include/fovpt.h that have no reference counterpart: the multi-GPU
        tile ownership and shared-frame calls (fovpt_set_tile_owner, fovpt_peer_frame_*), ray queries, BVH introspection */
    fovpt_ctx* context() const { return ctx; }

    // ---- public data, exactly as in the reference (SimplePathtracer.h:116,146)
    cudaStream_t stream = nullptr;
    LaunchParams launchParams;

protected:
    void check(int rc);
    fovpt_ctx* ctx = nullptr;
    fovpt_config config;
    const Model* model = nullptr;
    sutil::Camera lastSetCamera;
    uchar4* ownFrameBuffer = nullptr;      // what resize() allocated (render(target) redirects frame_buffer, :219)
};
