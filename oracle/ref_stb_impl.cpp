// TEST INFRASTRUCTURE: the reference defines STB_IMAGE_IMPLEMENTATION in main.cpp:5-6, which
// oracle/_ref does not build; this unit instantiates the vendored decoder for Model.cpp.
#define STB_IMAGE_IMPLEMENTATION
#include "support/stb/stb_image.h"
