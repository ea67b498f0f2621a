// CUDAOutputBuffer.h -- the render target type of SampleRenderer::render(CUDAOutputBuffer&)
// (reference sutil/CUDAOutputBuffer.h:43-330): same enum, same member functions, same meaning of each flavour --
//   CUDA_DEVICE : pixels in device memory, getHostPointer() copies them to a host vector
//   ZERO_COPY   : pixels in mapped, portable pinned HOST memory; the kernels write them over PCIe, getHostPointer() is that
//                 memory itself (no copy) -- the flavour the reference recommends for multi-GPU without full NVLink
//   CUDA_P2P    : pixels in device memory of setDevice()'s GPU; getHostPointer() copies through that device
//   GL_INTEROP  : needs an OpenGL context on the display GPU; this headless library has none, so constructing one throws the
//                 error the reference throws when the active device is not the display device (:117-125)
// getPBO()/deletePBO() are windowing calls (glad / GL buffer objects) and are not provided.
#pragma once
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>
#include <vector>

namespace sutil {

enum class CUDAOutputBufferType { CUDA_DEVICE = 0, GL_INTEROP = 1, ZERO_COPY = 2, CUDA_P2P = 3 };

template <typename PIXEL_FORMAT>
class CUDAOutputBuffer {
public:
    CUDAOutputBuffer(CUDAOutputBufferType type, int width, int height) : m_type(type)
    {
        if (type == CUDAOutputBufferType::GL_INTEROP)
            throw std::runtime_error("GL interop is only available on display device: this is a headless build without an OpenGL "
                                     "context, use CUDA_DEVICE, ZERO_COPY or CUDA_P2P");
        resize(width, height);
    }
    ~CUDAOutputBuffer() { release(); }
    CUDAOutputBuffer(const CUDAOutputBuffer&) = delete;
    CUDAOutputBuffer& operator=(const CUDAOutputBuffer&) = delete;

    void setDevice(int device_idx) { if (device_idx != m_device_idx) { release(); m_width = m_height = 0; } m_device_idx = device_idx; }
    void setStream(cudaStream_t s) { m_stream = s; }

    void resize(int width, int height)
    {
        if (width < 1) width = 1;                                  // ensureMinimumSize (:168)
        if (height < 1) height = 1;
        if (width == m_width && height == m_height && (m_device_pixels || m_host_zcopy_pixels)) return;
        release();
        const size_t bytes = (size_t)width * height * sizeof(PIXEL_FORMAT);
        DeviceScope scope(m_device_idx);
        if (m_type == CUDAOutputBufferType::ZERO_COPY) {
            check(cudaHostAlloc((void**)&m_host_zcopy_pixels, bytes, cudaHostAllocPortable | cudaHostAllocMapped), "cudaHostAlloc");
            check(cudaHostGetDevicePointer((void**)&m_device_pixels, m_host_zcopy_pixels, 0), "cudaHostGetDevicePointer");
        } else {
            check(cudaMalloc((void**)&m_device_pixels, bytes), "cudaMalloc");
        }
        m_width = width; m_height = height;
        if (!m_host_pixels.empty()) m_host_pixels.resize((size_t)width * height);
    }
    PIXEL_FORMAT* map() { return m_device_pixels; }                // nothing to do for the three headless flavours (:236-259)
    void unmap() { DeviceScope scope(m_device_idx); check(cudaStreamSynchronize(m_stream), "cudaStreamSynchronize"); }   // :262-279
    int width() const { return m_width; }
    int height() const { return m_height; }
    CUDAOutputBufferType type() const { return m_type; }
    PIXEL_FORMAT* getHostPointer()
    {
        if (m_type == CUDAOutputBufferType::ZERO_COPY) return m_host_zcopy_pixels;
        m_host_pixels.resize((size_t)m_width * m_height);
        DeviceScope scope(m_device_idx);
        check(cudaMemcpy(m_host_pixels.data(), m_device_pixels, m_host_pixels.size() * sizeof(PIXEL_FORMAT), cudaMemcpyDeviceToHost), "cudaMemcpy");
        return m_host_pixels.data();
    }

private:
    // makeCurrent() of the reference, but restoring the caller's device afterwards
    struct DeviceScope {
        int prev = 0;
        explicit DeviceScope(int d) { cudaGetDevice(&prev); if (d != prev) cudaSetDevice(d); else prev = -1; }
        ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
    };
    void release()
    {
        if (!m_device_pixels && !m_host_zcopy_pixels) return;
        DeviceScope scope(m_device_idx);
        if (m_type == CUDAOutputBufferType::ZERO_COPY) cudaFreeHost(m_host_zcopy_pixels); else cudaFree(m_device_pixels);
        m_device_pixels = nullptr; m_host_zcopy_pixels = nullptr;
    }
    static void check(cudaError_t e, const char* what)
    {
        if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));   // sutil::Exception convention
    }
    CUDAOutputBufferType m_type;
    int m_width = 0, m_height = 0;
    PIXEL_FORMAT* m_device_pixels = nullptr;
    PIXEL_FORMAT* m_host_zcopy_pixels = nullptr;
    std::vector<PIXEL_FORMAT> m_host_pixels;
    cudaStream_t m_stream = nullptr;
    int m_device_idx = 0;
};

}  // namespace sutil
