// ImageDecode.h -- texture file decoding for loadOBJ (the reference calls stbi_load(file, &w, &h, &comp, STBI_rgb_alpha),
// PT_sv5_/Model.cpp:105-107).  Own decoders, no third-party code: PNG (all colour types, 1-16 bit, tRNS, Adam7),
// TGA (true colour / colour mapped / grey, raw and RLE, 8-32 bit), JPEG (baseline, extended sequential and
// progressive Huffman, 8 bit, grey / YCbCr / RGB, any sampling factors, restart intervals), BMP (24/32 bit) and
// binary PPM.  Output is what the reference's decoder hands to loadTexture: RGBA8, one uint32 per texel
// (R in the low byte), top row first; the arithmetic that decides texel values (16 -> 8 bit reduction, 5-bit
// expansion, JPEG integer IDCT / chroma upsampling / YCbCr conversion) follows that decoder so textures are
// byte-identical (tests/test_host_cpp.py::test_texture_decoders_match_reference_loader).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace fovpt_image {

// false + *why on failure (unknown format, truncated file, unsupported variant); px is left empty then
bool decode_file(const std::string& path, int& w, int& h, std::vector<uint32_t>& px, std::string* why = nullptr);
bool decode_memory(const unsigned char* data, size_t size, int& w, int& h, std::vector<uint32_t>& px, std::string* why = nullptr);

}  // namespace fovpt_image
