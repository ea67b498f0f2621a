// CUDAOutputBuffer.h -- the render target type of SampleRenderer::render(CUDAOutputBuffer&)
// (reference sutil/CUDAOutputBuffer.h:43-120).  Only the CUDA_DEVICE flavour the reference's headless
// path needs: a plain device allocation with map()/unmap()/getHostPointer(); GL interop, zero-copy and
// P2P display buffers are windowing code and out of scope.
#pragma once
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>
#include <vector>

namespace sutil {

enum class CUDAOutputBufferType { CUDA_DEVICE = 0 };

template <typename PIXEL_FORMAT>
class CUDAOutputBuffer {
public:
    CUDAOutputBuffer(CUDAOutputBufferType, int width, int height) { resize(width, height); }
    ~CUDAOutputBuffer() { cudaFree(m_device_pixels); }
    CUDAOutputBuffer(const CUDAOutputBuffer&) = delete;
    CUDAOutputBuffer& operator=(const CUDAOutputBuffer&) = delete;

    void resize(int width, int height)
    {
        if (width < 1 || height < 1) throw std::runtime_error("CUDAOutputBuffer dimensions must be at least 1");
        if (width == m_width && height == m_height) return;
        cudaFree(m_device_pixels); m_device_pixels = nullptr;
        check(cudaMalloc((void**)&m_device_pixels, (size_t)width * height * sizeof(PIXEL_FORMAT)), "cudaMalloc");
        m_width = width; m_height = height;
        m_host_pixels.resize((size_t)width * height);
    }
    PIXEL_FORMAT* map() { return m_device_pixels; }
    void unmap() { if (m_stream) check(cudaStreamSynchronize(m_stream), "cudaStreamSynchronize"); }
    void setStream(cudaStream_t s) { m_stream = s; }
    int width() const { return m_width; }
    int height() const { return m_height; }
    PIXEL_FORMAT* getHostPointer()
    {
        check(cudaMemcpy(m_host_pixels.data(), m_device_pixels, m_host_pixels.size() * sizeof(PIXEL_FORMAT), cudaMemcpyDeviceToHost), "cudaMemcpy");
        return m_host_pixels.data();
    }

private:
    static void check(cudaError_t e, const char* what)
    {
        if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));   // sutil::Exception convention
    }
    int m_width = 0, m_height = 0;
    PIXEL_FORMAT* m_device_pixels = nullptr;
    std::vector<PIXEL_FORMAT> m_host_pixels;
    cudaStream_t m_stream = nullptr;
};

}  // namespace sutil
