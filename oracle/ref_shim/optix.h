// Host-side stand-in for the OptiX SDK's <optix.h>  --  TEST INFRASTRUCTURE ONLY.
// The OptiX SDK is not present in this image; the reference headers only need a
// handful of type names from it to parse (LaunchParams.h:19, CUDABuffer.h:20,
// sutil/Exception.h:32).  Nothing here ships in the product library.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
typedef unsigned long long OptixTraversableHandle;
typedef int OptixResult;
#define OPTIX_SUCCESS 0
inline const char* optixGetErrorName(OptixResult) { return "OPTIX_SHIM"; }
