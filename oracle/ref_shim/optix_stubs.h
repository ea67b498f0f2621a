// intentionally empty: CUDABuffer.h:21 includes it, nothing from it is used.
#pragma once
