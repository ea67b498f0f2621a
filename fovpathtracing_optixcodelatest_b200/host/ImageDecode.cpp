// ImageDecode.cpp -- see ImageDecode.h.  Replaces the stbi_load call of the reference's loadTexture
// (PT_sv5_/Model.cpp:105-107) with decoders written for this repository.
//
// Layout of this file: byte cursor -> inflate (RFC 1951) -> PNG -> TGA -> JPEG (ITU T.81) -> BMP / PPM -> dispatch.
// Value-deciding arithmetic is chosen to produce the texels the reference's decoder produces:
//   PNG  : 16-bit samples keep their high byte; 1/2/4-bit grey is scaled by 255/(2^d-1); a tRNS colour key is compared on
//          the full sample (16-bit: before the reduction); no gamma handling
//   TGA  : 15/16-bit true colour is 5-5-5 without alpha, expanded as c*255/31; 16-bit grey is (grey, alpha)
//   JPEG : 13-bit fixed-point LLM inverse DCT with the +128 level shift folded into the row pass, triangle-filter chroma
//          upsampling for 2x1 / 1x2 / 2x2, point sampling for other factors, 12.8 fixed-point YCbCr conversion with the
//          Cb->G term truncated to 16 fractional bits
#include "ImageDecode.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace fovpt_image {
namespace {

struct Fail { const char* why; };
[[noreturn]] void fail(const char* why) { throw Fail{why}; }

// ---------------------------------------------------------------- byte cursor
struct Bytes {
    const uint8_t* p; size_t n; size_t pos = 0;
    Bytes(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
    bool eof() const { return pos >= n; }
    unsigned u8() { if (pos >= n) fail("file is truncated"); return p[pos++]; }
    unsigned u8_or0() { return pos < n ? p[pos++] : 0u; }
    unsigned be16() { unsigned a = u8(); return (a << 8) | u8(); }
    unsigned be32() { unsigned a = be16(); return (a << 16) | be16(); }
    unsigned le16() { unsigned a = u8(); return a | (u8() << 8); }
    unsigned le32() { unsigned a = le16(); return a | (le16() << 16); }
    void skip(size_t k) { if (pos > n || k > n - pos) fail("file is truncated"); pos += k; }
    const uint8_t* take(size_t k) { if (pos > n || k > n - pos) fail("file is truncated"); const uint8_t* q = p + pos; pos += k; return q; }
};

inline uint32_t rgba(unsigned r, unsigned g, unsigned b, unsigned a) { return r | (g << 8) | (b << 16) | (a << 24); }

// a corrupt header must not turn into a multi-gigabyte allocation: 2^28 texels (1 GiB of RGBA8) is far beyond any texture
inline void check_size(long long w, long long h) { if (w <= 0 || h <= 0 || w * h > (1ll << 28)) fail("image dimensions out of range"); }

// ---------------------------------------------------------------- inflate
// Canonical prefix code: a 10-bit first-level table (symbol | length << 9) and, for longer codes, the
// count / offset walk over the code lengths.
struct Prefix {
    static constexpr int kFast = 10;
    uint16_t fast[1 << kFast];
    uint16_t count[16];
    uint16_t sorted[288];

    void build(const uint8_t* len, int n)
    {
        std::memset(count, 0, sizeof(count));
        std::memset(fast, 0, sizeof(fast));
        for (int i = 0; i < n; ++i) ++count[len[i]];
        count[0] = 0;
        int left = 1;
        for (int l = 1; l < 16; ++l) { left = (left << 1) - count[l]; if (left < 0) fail("zlib: over-subscribed prefix code"); }
        uint16_t offs[16]; offs[1] = 0;
        for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + count[l];
        uint16_t next[16]; unsigned code = 0;
        for (int l = 1; l < 16; ++l) { code = (code + count[l - 1]) << 1; next[l] = (uint16_t)code; }
        for (int i = 0; i < n; ++i) {
            const int l = len[i];
            if (!l) continue;
            sorted[offs[l]++] = (uint16_t)i;
            const unsigned c = next[l]++;
            if (l <= kFast) {
                unsigned rev = 0;
                for (int b = 0; b < l; ++b) rev |= ((c >> b) & 1u) << (l - 1 - b);
                for (unsigned k = rev; k < (1u << kFast); k += 1u << l) fast[k] = (uint16_t)(i | (l << 9));
            }
        }
    }
};

struct BitsLsb {
    const uint8_t* p; size_t n; size_t pos = 0; uint64_t buf = 0; int cnt = 0;
    BitsLsb(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
    void fill() { while (cnt <= 56 && pos < n) { buf |= (uint64_t)p[pos++] << cnt; cnt += 8; } }
    unsigned bits(int k)
    {
        if (cnt < k) { fill(); if (cnt < k) fail("zlib: stream is truncated"); }
        const unsigned v = (unsigned)(buf & ((1ull << k) - 1));
        buf >>= k; cnt -= k;
        return v;
    }
    int symbol(const Prefix& h)
    {
        if (cnt < 15) fill();
        const unsigned e = h.fast[buf & ((1u << Prefix::kFast) - 1)];
        if (e) {
            const int l = (int)(e >> 9);
            if (l > cnt) fail("zlib: stream is truncated");
            buf >>= l; cnt -= l;
            return (int)(e & 511u);
        }
        int code = 0, first = 0, index = 0;
        for (int l = 1; l < 16; ++l) {
            code |= (int)bits(1);
            const int c = h.count[l];
            if (code - c < first) return h.sorted[index + (code - first)];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        fail("zlib: invalid prefix code");
    }
};

void inflate_zlib(const uint8_t* src, size_t n, std::vector<uint8_t>& out, size_t expected)
{
    static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

    if (n < 2) fail("zlib: stream is truncated");
    const unsigned cmf = src[0], flg = src[1];
    if ((cmf & 15) != 8 || ((cmf << 8) | flg) % 31 != 0 || (flg & 32)) fail("zlib: bad header");
    BitsLsb in(src + 2, n - 2);
    out.clear(); out.reserve(expected);
    Prefix lit, dist;
    for (bool last = false; !last;) {
        last = in.bits(1) != 0;
        const unsigned type = in.bits(2);
        if (type == 0) {
            in.bits(in.cnt & 7);
            const unsigned len = in.bits(16), nlen = in.bits(16);
            if ((len ^ 0xffffu) != nlen) fail("zlib: stored block length check failed");
            for (unsigned i = 0; i < len; ++i) out.push_back((uint8_t)in.bits(8));
            continue;
        }
        if (type == 3) fail("zlib: reserved block type");
        uint8_t lens[320];
        if (type == 1) {
            for (int i = 0; i < 288; ++i) lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
            lit.build(lens, 288);
            for (int i = 0; i < 30; ++i) lens[i] = 5;
            dist.build(lens, 30);
        } else {
            const int nlit = (int)in.bits(5) + 257, ndist = (int)in.bits(5) + 1, ncl = (int)in.bits(4) + 4;
            if (nlit > 286 || ndist > 30) fail("zlib: bad code counts");
            uint8_t cl[19] = {0};
            for (int i = 0; i < ncl; ++i) cl[order[i]] = (uint8_t)in.bits(3);
            Prefix clh; clh.build(cl, 19);
            int i = 0;
            while (i < nlit + ndist) {
                const int s = in.symbol(clh);
                if (s < 16) { lens[i++] = (uint8_t)s; continue; }
                int rep; uint8_t v = 0;
                if (s == 16) { if (!i) fail("zlib: repeat without a previous length"); v = lens[i - 1]; rep = 3 + (int)in.bits(2); }
                else if (s == 17) rep = 3 + (int)in.bits(3);
                else rep = 11 + (int)in.bits(7);
                if (i + rep > nlit + ndist) fail("zlib: code lengths overflow");
                while (rep--) lens[i++] = v;
            }
            if (!lens[256]) fail("zlib: no end-of-block code");
            lit.build(lens, nlit);
            dist.build(lens + nlit, ndist);
        }
        for (;;) {
            const int s = in.symbol(lit);
            if (s < 256) { out.push_back((uint8_t)s); continue; }
            if (s == 256) break;
            if (s > 285) fail("zlib: bad length symbol");
            const unsigned len = len_base[s - 257] + in.bits(len_extra[s - 257]);
            const int ds = in.symbol(dist);
            if (ds > 29) fail("zlib: bad distance symbol");
            const size_t d = dist_base[ds] + in.bits(dist_extra[ds]);
            if (d > out.size()) fail("zlib: distance reaches before the start");
            const size_t at = out.size();
            out.resize(at + len);
            uint8_t* o = out.data() + at;
            const uint8_t* f = o - d;
            for (unsigned i = 0; i < len; ++i) o[i] = f[i];         // byte by byte: ranges may overlap
        }
    }
}

// ---------------------------------------------------------------- PNG
bool is_png(const uint8_t* d, size_t n) { static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 13, 10, 26, 10}; return n >= 8 && !std::memcmp(d, sig, 8); }

void decode_png(const uint8_t* d, size_t n, int& w, int& h, std::vector<uint32_t>& px)
{
    Bytes in(d, n); in.skip(8);
    unsigned depth = 0, ctype = 0, interlace = 0;
    bool have_hdr = false, have_key = false;
    uint8_t pal[256][4]; int npal = 0;
    for (auto& e : pal) { e[0] = e[1] = e[2] = 0; e[3] = 255; }
    unsigned key[3] = {0, 0, 0};
    std::vector<uint8_t> z;
    for (;;) {
        const unsigned len = in.be32(), tag = in.be32();
        if (tag == 0x49484452u) {                                    // IHDR
            if (len != 13) fail("PNG: bad IHDR");
            w = (int)in.be32(); h = (int)in.be32(); depth = in.u8(); ctype = in.u8();
            const unsigned comp = in.u8(), filt = in.u8(); interlace = in.u8();
            if (w <= 0 || h <= 0 || w > (1 << 24) || h > (1 << 24)) fail("PNG: bad image size");
            check_size(w, h);
            if (comp || filt || interlace > 1) fail("PNG: unknown compression / filter / interlace method");
            const bool ok = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                            ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16)) ||
                            (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8));
            if (!ok) fail("PNG: bad colour type / bit depth");
            have_hdr = true;
        } else if (!have_hdr) {
            fail("PNG: first chunk is not IHDR");
        } else if (tag == 0x504c5445u) {                             // PLTE
            if (len > 768 || len % 3) fail("PNG: bad PLTE");
            npal = (int)len / 3;
            for (int i = 0; i < npal; ++i) { pal[i][0] = (uint8_t)in.u8(); pal[i][1] = (uint8_t)in.u8(); pal[i][2] = (uint8_t)in.u8(); }
        } else if (tag == 0x74524e53u) {                             // tRNS
            if (ctype == 3) { if ((int)len > npal) fail("PNG: bad tRNS"); for (unsigned i = 0; i < len; ++i) pal[i][3] = (uint8_t)in.u8(); }
            else if (ctype == 0 && len == 2) { key[0] = in.be16(); have_key = true; }
            else if (ctype == 2 && len == 6) { key[0] = in.be16(); key[1] = in.be16(); key[2] = in.be16(); have_key = true; }
            else fail("PNG: bad tRNS");
        } else if (tag == 0x49444154u) {                             // IDAT
            const uint8_t* q = in.take(len);
            z.insert(z.end(), q, q + len);
        } else if (tag == 0x49454e44u) {                             // IEND
            break;
        } else {
            if (!(tag & 0x20000000u)) fail("PNG: unknown critical chunk");
            in.skip(len);
        }
        in.skip(4);                                                  // CRC (not verified, as in the reference's decoder)
    }
    if (!have_hdr || z.empty()) fail("PNG: no image data");
    if (ctype == 3 && !npal) fail("PNG: palette image without PLTE");

    const int chan = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : 4;
    const int bits_px = chan * (int)depth;
    const int bpp = std::max(1, bits_px / 8);                        // filter distance in bytes
    auto row_bytes = [&](int pw) { return ((size_t)pw * bits_px + 7) / 8; };
    size_t expected = 0;
    static const int x0[7] = {0, 4, 0, 2, 0, 1, 0}, y0[7] = {0, 0, 4, 0, 2, 0, 1}, dx[7] = {8, 8, 4, 4, 2, 2, 1}, dy[7] = {8, 8, 8, 4, 4, 2, 2};
    if (!interlace) expected = (row_bytes(w) + 1) * (size_t)h;
    else for (int p = 0; p < 7; ++p) { const int pw = (w - x0[p] + dx[p] - 1) / dx[p], ph = (h - y0[p] + dy[p] - 1) / dy[p]; if (pw > 0 && ph > 0) expected += (row_bytes(pw) + 1) * (size_t)ph; }
    std::vector<uint8_t> raw;
    inflate_zlib(z.data(), z.size(), raw, expected);
    if (raw.size() < expected) fail("PNG: not enough pixel data");

    px.assign((size_t)w * h, 0);
    const unsigned scale = depth == 1 ? 255u : depth == 2 ? 85u : depth == 4 ? 17u : 1u;
    if (depth <= 8) for (auto& k : key) k &= 255u;
    std::vector<uint8_t> prev, cur;
    // one pass (the whole image, or one Adam7 sub-image): undo the filters, expand to RGBA8
    auto pass = [&](const uint8_t* src, int pw, int ph, int ox, int oy, int sx, int sy) {
        const size_t rb = row_bytes(pw);
        prev.assign(rb, 0); cur.resize(rb);
        for (int y = 0; y < ph; ++y) {
            const unsigned ft = *src++;
            if (ft > 4) fail("PNG: unknown filter type");
            for (size_t i = 0; i < rb; ++i) {
                const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
                int pred = 0;
                switch (ft) {
                    case 1: pred = a; break;
                    case 2: pred = b; break;
                    case 3: pred = (a + b) >> 1; break;
                    case 4: { const int pp = a + b - c, pa = std::abs(pp - a), pb = std::abs(pp - b), pc = std::abs(pp - c);
                              pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); } break;
                    default: break;
                }
                cur[i] = (uint8_t)(src[i] + pred);
            }
            src += rb;
            uint32_t* out = px.data() + (size_t)(oy + y * sy) * w;
            // sample k of pixel x as (8-bit value, full value for the colour-key compare)
            auto sample = [&](int x, int k, unsigned& full) -> unsigned {
                if (depth == 16) { const uint8_t* q = &cur[((size_t)x * chan + k) * 2]; full = ((unsigned)q[0] << 8) | q[1]; return q[0]; }
                if (depth == 8) { full = cur[(size_t)x * chan + k]; return full; }
                const size_t bit = (size_t)x * depth;
                const unsigned v = (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1);
                full = v;
                return v;
            };
            for (int x = 0; x < pw; ++x) {
                unsigned f0, f1, f2, f3;
                uint32_t v;
                if (ctype == 3) {
                    const unsigned i = sample(x, 0, f0);
                    v = rgba(pal[i][0], pal[i][1], pal[i][2], pal[i][3]);       // indices past PLTE read (0,0,0,255)
                } else if (ctype == 0) {
                    unsigned g = sample(x, 0, f0);
                    if (depth < 8) g *= scale;
                    v = rgba(g, g, g, (have_key && f0 == key[0]) ? 0u : 255u);
                } else if (ctype == 4) {
                    const unsigned g = sample(x, 0, f0), a = sample(x, 1, f1);
                    v = rgba(g, g, g, a);
                } else if (ctype == 2) {
                    const unsigned r = sample(x, 0, f0), g = sample(x, 1, f1), b = sample(x, 2, f2);
                    v = rgba(r, g, b, (have_key && f0 == key[0] && f1 == key[1] && f2 == key[2]) ? 0u : 255u);
                } else {
                    const unsigned r = sample(x, 0, f0), g = sample(x, 1, f1), b = sample(x, 2, f2), a = sample(x, 3, f3);
                    v = rgba(r, g, b, a);
                }
                out[ox + x * sx] = v;
            }
            prev.swap(cur);
        }
        return src;
    };
    const uint8_t* src = raw.data();
    if (!interlace) pass(src, w, h, 0, 0, 1, 1);
    else for (int p = 0; p < 7; ++p) {
        const int pw = (w - x0[p] + dx[p] - 1) / dx[p], ph = (h - y0[p] + dy[p] - 1) / dy[p];
        if (pw > 0 && ph > 0) src = pass(src, pw, ph, x0[p], y0[p], dx[p], dy[p]);
    }
}

// ---------------------------------------------------------------- TGA
struct TgaHeader { unsigned idlen, cmap_type, type, cmap_first, cmap_len, cmap_bits, w, h, bpp, desc; };

bool read_tga_header(const uint8_t* d, size_t n, TgaHeader& t)
{
    if (n < 18) return false;
    t.idlen = d[0]; t.cmap_type = d[1]; t.type = d[2];
    t.cmap_first = d[3] | (d[4] << 8); t.cmap_len = d[5] | (d[6] << 8); t.cmap_bits = d[7];
    t.w = d[12] | (d[13] << 8); t.h = d[14] | (d[15] << 8); t.bpp = d[16]; t.desc = d[17];
    auto colour_bits = [](unsigned b) { return b == 8 || b == 15 || b == 16 || b == 24 || b == 32; };
    if (t.cmap_type > 1 || t.w < 1 || t.h < 1) return false;
    if (t.cmap_type == 1) return (t.type == 1 || t.type == 9) && colour_bits(t.cmap_bits) && (t.bpp == 8 || t.bpp == 16);
    return (t.type == 2 || t.type == 3 || t.type == 10 || t.type == 11) && colour_bits(t.bpp);
}

void decode_tga(const uint8_t* d, size_t n, int& w, int& h, std::vector<uint32_t>& px)
{
    TgaHeader t;
    if (!read_tga_header(d, n, t)) fail("TGA: bad header");
    Bytes in(d, n); in.skip(18 + t.idlen);
    const bool indexed = t.cmap_type == 1, rle = t.type >= 8, grey = (t.type & 7) == 3;
    if (indexed && t.cmap_len == 0) fail("TGA: colour-mapped image without a palette");
    check_size(t.w, t.h);
    const unsigned cbits = indexed ? t.cmap_bits : t.bpp;
    if (grey && cbits != 8 && cbits != 16) fail("TGA: unsupported grey depth");
    const unsigned cbytes = (cbits + 7) / 8;
    auto colour = [&](const uint8_t* q) -> uint32_t {
        switch (cbits) {
            case 8: return rgba(q[0], q[0], q[0], 255);
            case 15: case 16:
                if (grey) return rgba(q[0], q[0], q[0], q[1]);
                else { const unsigned v = q[0] | (q[1] << 8); return rgba(((v >> 10) & 31) * 255 / 31, ((v >> 5) & 31) * 255 / 31, (v & 31) * 255 / 31, 255); }
            case 24: return rgba(q[2], q[1], q[0], 255);
            default: return rgba(q[2], q[1], q[0], q[3]);
        }
    };
    std::vector<uint32_t> palette;
    if (indexed) {
        in.skip(t.cmap_first);                                       // the reference's decoder skips this many BYTES
        const uint8_t* q = in.take((size_t)t.cmap_len * cbytes);
        palette.resize(t.cmap_len);
        for (unsigned i = 0; i < t.cmap_len; ++i) palette[i] = colour(q + (size_t)i * cbytes);
    }
    w = (int)t.w; h = (int)t.h;
    px.resize((size_t)w * h);
    const unsigned pbytes = t.bpp / 8 + ((t.bpp & 7) ? 1 : 0);
    auto pixel = [&]() -> uint32_t {
        const uint8_t* q = in.take(pbytes);
        if (!indexed) return colour(q);
        const unsigned i = pbytes == 1 ? q[0] : (q[0] | (q[1] << 8));
        return palette[i < t.cmap_len ? i : 0];
    };
    const size_t total = (size_t)w * h;
    if (!rle) for (size_t i = 0; i < total; ++i) px[i] = pixel();
    else for (size_t i = 0; i < total;) {
        const unsigned c = in.u8(); size_t run = (c & 127) + 1;
        if (run > total - i) run = total - i;
        if (c & 128) { const uint32_t v = pixel(); for (size_t k = 0; k < run; ++k) px[i++] = v; }
        else for (size_t k = 0; k < run; ++k) px[i++] = pixel();
    }
    if (!(t.desc & 0x20))                                            // bottom-left origin: flip to top row first
        for (int y = 0; y < h / 2; ++y) std::swap_ranges(px.begin() + (size_t)y * w, px.begin() + (size_t)(y + 1) * w, px.begin() + (size_t)(h - 1 - y) * w);
}

// ---------------------------------------------------------------- JPEG
static const uint8_t kZigzag[64 + 15] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                         35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
                                         63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};   // padding: a corrupt run cannot index past the block

struct JHuff {
    uint16_t look[512];            // 9-bit lookahead: symbol | length << 8, 0xffff = longer code
    uint8_t  value[256];
    unsigned maxcode[18];          // left-aligned to 16 bits, exclusive
    int      delta[17];
    bool     set = false;

    void build(const uint8_t counts[16], const uint8_t* vals, int nvals)
    {
        std::memcpy(value, vals, (size_t)nvals);
        std::fill(look, look + 512, (uint16_t)0xffff);
        unsigned code = 0; int k = 0;
        for (int l = 1; l <= 16; ++l) {
            delta[l] = k - (int)code;
            if (code + counts[l - 1] > (1u << l)) fail("JPEG: bad Huffman table");      // over-subscribed: would index past the tables
            for (int i = 0; i < counts[l - 1]; ++i, ++k, ++code) {
                if (l <= 9) { const unsigned lo = code << (9 - l); for (unsigned j = 0; j < (1u << (9 - l)); ++j) look[lo + j] = (uint16_t)(vals[k] | (l << 8)); }
            }
            if (code > (1u << l)) fail("JPEG: bad Huffman table");
            maxcode[l] = code << (16 - l);
            code <<= 1;
        }
        maxcode[17] = 0xffffffffu;
        set = true;
    }
};

struct JComp {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int x = 0, y = 0, w2 = 0, h2 = 0, dc_pred = 0;
    std::vector<uint8_t> data;
    std::vector<int16_t> coeff;    // progressive: w2/8 x h2/8 blocks of 64
};

struct Jpeg {
    Bytes in;
    int w = 0, h = 0, ncomp = 0;
    bool progressive = false, jfif = false;
    int adobe_transform = -1, rgb_ids = 0;
    uint16_t dequant[4][64];
    JHuff hdc[4], hac[4];
    JComp comp[4];
    int h_max = 1, v_max = 1, mcu_x = 0, mcu_y = 0;
    int restart_interval = 0, todo = 0;
    int scan_n = 0, order[4];
    int spec_start = 0, spec_end = 63, succ_high = 0, succ_low = 0, eob_run = 0;
    // entropy-coded segment reader (MSB first, FF00 unstuffing, stops feeding at a marker)
    uint32_t buf = 0; int cnt = 0; int marker = -1; bool nomore = false;

    Jpeg(const uint8_t* d, size_t n) : in(d, n) {}

    void grow()
    {
        do {
            unsigned b = nomore ? 0u : in.u8_or0();
            if (b == 0xff) {
                unsigned c = in.u8_or0();
                while (c == 0xff) c = in.u8_or0();
                if (c != 0) { marker = (int)c; nomore = true; b = 0; }      // zeros from here on
            }
            buf |= b << (24 - cnt);
            cnt += 8;
        } while (cnt <= 24);
    }
    int decode(const JHuff& hf)
    {
        if (cnt < 16) grow();
        const unsigned e = hf.look[buf >> 23];
        if (e != 0xffff) {
            const int l = (int)(e >> 8);
            if (l > cnt) fail("JPEG: entropy data is truncated");
            buf <<= l; cnt -= l;
            return (int)(e & 255u);
        }
        const unsigned t = buf >> 16;
        int l = 10;
        while (t >= hf.maxcode[l]) ++l;
        if (l > 16 || l > cnt) fail("JPEG: bad Huffman code");
        const int c = (int)((buf >> (32 - l)) & ((1u << l) - 1)) + hf.delta[l];
        if (c < 0 || c > 255) fail("JPEG: bad Huffman code");
        buf <<= l; cnt -= l;
        return hf.value[c];
    }
    int getbits(int nb)
    {
        if (!nb) return 0;
        if (cnt < nb) grow();
        if (cnt < nb) fail("JPEG: entropy data is truncated");
        const int v = (int)(buf >> (32 - nb));
        buf <<= nb; cnt -= nb;
        return v;
    }
    int extend(int nb)             // T.81 F.2.2.1: nb magnitude bits -> signed value
    {
        const int v = getbits(nb);
        return v < (1 << (nb - 1)) ? v - (1 << nb) + 1 : v;
    }
    void reset_entropy()
    {
        buf = 0; cnt = 0; nomore = false; marker = -1; eob_run = 0;
        for (auto& c : comp) c.dc_pred = 0;
        todo = restart_interval ? restart_interval : 0x7fffffff;
    }

    // ---- block decoders
    void block_baseline(int16_t d[64], JComp& c)
    {
        std::memset(d, 0, 64 * sizeof(int16_t));
        const uint16_t* q = dequant[c.tq];
        const int t = decode(hdc[c.td]);
        if (t > 15) fail("JPEG: bad DC size");
        c.dc_pred += t ? extend(t) : 0;
        d[0] = (int16_t)(c.dc_pred * q[0]);
        const JHuff& ac = hac[c.ta];
        for (int k = 1; k < 64;) {
            const int rs = decode(ac), s = rs & 15, r = rs >> 4;
            if (!s) { if (rs != 0xf0) break; k += 16; }
            else { k += r; const int z = kZigzag[k++]; d[z] = (int16_t)(extend(s) * q[z]); }
        }
    }
    void block_dc(int16_t d[64], JComp& c)
    {
        if (spec_end != 0) fail("JPEG: DC scan with AC coefficients");
        if (succ_high == 0) {
            std::memset(d, 0, 64 * sizeof(int16_t));
            const int t = decode(hdc[c.td]);
            if (t > 15) fail("JPEG: bad DC size");
            c.dc_pred += t ? extend(t) : 0;
            d[0] = (int16_t)(c.dc_pred * (1 << succ_low));
        } else if (getbits(1)) {
            d[0] = (int16_t)(d[0] + (1 << succ_low));
        }
    }
    void block_ac(int16_t d[64], const JHuff& ac)
    {
        if (succ_high == 0) {
            if (eob_run) { --eob_run; return; }
            int k = spec_start;
            do {
                const int rs = decode(ac), s = rs & 15, r = rs >> 4;
                if (!s) {
                    if (r < 15) { eob_run = (1 << r) + getbits(r) - 1; break; }
                    k += 16;
                } else {
                    k += r;
                    d[kZigzag[k++]] = (int16_t)(extend(s) * (1 << succ_low));
                }
            } while (k <= spec_end);
            return;
        }
        const int bit = 1 << succ_low;
        auto refine = [&](int16_t& p) { if (getbits(1) && !(p & bit)) p = (int16_t)(p + (p > 0 ? bit : -bit)); };
        if (eob_run) {
            --eob_run;
            for (int k = spec_start; k <= spec_end; ++k) { int16_t& p = d[kZigzag[k]]; if (p) refine(p); }
            return;
        }
        int k = spec_start;
        do {
            const int rs = decode(ac);
            int s = rs & 15, r = rs >> 4;
            if (!s) {
                if (r < 15) { eob_run = (1 << r) + getbits(r) - 1; r = 64; }      // run to the end of the band, refining on the way
            } else {
                if (s != 1) fail("JPEG: bad refinement code");
                s = getbits(1) ? bit : -bit;
            }
            while (k <= spec_end) {
                int16_t& p = d[kZigzag[k++]];
                if (p) refine(p);
                else { if (!r) { p = (int16_t)s; break; } --r; }
            }
        } while (k <= spec_end);
    }

    // ---- inverse DCT: 8x8 coefficients -> 8x8 samples at out (row stride given)
    // (64-bit intermediates: a valid stream stays far inside 32 bits and gets the same values; a corrupt one cannot overflow)
    typedef long long wide;
    static inline uint8_t clamp8(wide v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
    struct Lines { wide e0, e1, e2, e3, o0, o1, o2, o3; };
    static inline Lines idct_1d(wide s0, wide s1, wide s2, wide s3, wide s4, wide s5, wide s6, wide s7)
    {
        constexpr wide c0_541 = 2217, c1_847 = -7567, c0_765 = 3135, c1_175 = 4816, c0_298 = 1223, c2_053 = 8410, c3_072 = 12586,
                      c1_501 = 6149, c0_899 = -3685, c2_562 = -10497, c1_961 = -8034, c0_390 = -1597;        // round(x * 4096)
        // even part
        const wide z = (s2 + s6) * c0_541, a2 = z + s6 * c1_847, a3 = z + s2 * c0_765;
        const wide a0 = (s0 + s4) * 4096, a1 = (s0 - s4) * 4096;
        Lines L;
        L.e0 = a0 + a3; L.e3 = a0 - a3; L.e1 = a1 + a2; L.e2 = a1 - a2;
        // odd part
        const wide p3 = s7 + s3, p4 = s5 + s1, p1 = s7 + s1, p2 = s5 + s3, p5 = (p3 + p4) * c1_175;
        const wide q1 = p5 + p1 * c0_899, q2 = p5 + p2 * c2_562, q3 = p3 * c1_961, q4 = p4 * c0_390;
        L.o3 = s1 * c1_501 + q1 + q4; L.o2 = s3 * c3_072 + q2 + q3; L.o1 = s5 * c2_053 + q2 + q4; L.o0 = s7 * c0_298 + q1 + q3;
        return L;
    }
    static void idct(uint8_t* out, int stride, const int16_t d[64])
    {
        wide tmp[64];
        for (int i = 0; i < 8; ++i) {                                // columns, 10 bits dropped
            Lines L = idct_1d(d[i], d[8 + i], d[16 + i], d[24 + i], d[32 + i], d[40 + i], d[48 + i], d[56 + i]);
            L.e0 += 512; L.e1 += 512; L.e2 += 512; L.e3 += 512;
            tmp[i] = (L.e0 + L.o3) >> 10;      tmp[56 + i] = (L.e0 - L.o3) >> 10;
            tmp[8 + i] = (L.e1 + L.o2) >> 10;  tmp[48 + i] = (L.e1 - L.o2) >> 10;
            tmp[16 + i] = (L.e2 + L.o1) >> 10; tmp[40 + i] = (L.e2 - L.o1) >> 10;
            tmp[24 + i] = (L.e3 + L.o0) >> 10; tmp[32 + i] = (L.e3 - L.o0) >> 10;
        }
        for (int i = 0; i < 8; ++i, out += stride) {                 // rows, 17 bits dropped, +128 folded into the rounding term
            const wide* v = tmp + 8 * i;
            Lines L = idct_1d(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
            const wide bias = 65536 + (128 << 17);
            L.e0 += bias; L.e1 += bias; L.e2 += bias; L.e3 += bias;
            out[0] = clamp8((L.e0 + L.o3) >> 17); out[7] = clamp8((L.e0 - L.o3) >> 17);
            out[1] = clamp8((L.e1 + L.o2) >> 17); out[6] = clamp8((L.e1 - L.o2) >> 17);
            out[2] = clamp8((L.e2 + L.o1) >> 17); out[5] = clamp8((L.e2 - L.o1) >> 17);
            out[3] = clamp8((L.e3 + L.o0) >> 17); out[4] = clamp8((L.e3 - L.o0) >> 17);
        }
    }

    // ---- header segments
    void read_dqt()
    {
        int len = (int)in.be16() - 2;
        while (len > 0) {
            const unsigned pq = in.u8(), prec = pq >> 4, t = pq & 15;
            if (prec > 1 || t > 3) fail("JPEG: bad quantisation table");
            for (int i = 0; i < 64; ++i) dequant[t][kZigzag[i]] = (uint16_t)(prec ? in.be16() : in.u8());
            len -= prec ? 129 : 65;
        }
        if (len) fail("JPEG: bad DQT length");
    }
    void read_dht()
    {
        int len = (int)in.be16() - 2;
        while (len > 0) {
            const unsigned tc = in.u8(), cls = tc >> 4, t = tc & 15;
            if (cls > 1 || t > 3) fail("JPEG: bad Huffman table id");
            uint8_t counts[16], vals[256]; int total = 0;
            for (auto& c : counts) { c = (uint8_t)in.u8(); total += c; }
            if (total > 256) fail("JPEG: bad Huffman table");
            for (int i = 0; i < total; ++i) vals[i] = (uint8_t)in.u8();
            (cls ? hac[t] : hdc[t]).build(counts, vals, total);
            len -= 17 + total;
        }
        if (len) fail("JPEG: bad DHT length");
    }
    void read_sof()
    {
        const unsigned len = in.be16();
        if (in.u8() != 8) fail("JPEG: only 8-bit samples are supported");
        h = (int)in.be16(); w = (int)in.be16(); ncomp = (int)in.u8();
        if (h <= 0 || w <= 0) fail("JPEG: bad image size");
        check_size(w, h);
        if (ncomp == 4) fail("JPEG: CMYK / YCCK images are not supported");
        if (ncomp != 1 && ncomp != 3) fail("JPEG: bad component count");
        if (len != 8u + 3u * ncomp) fail("JPEG: bad SOF length");
        static const char rgb[3] = {'R', 'G', 'B'};
        rgb_ids = 0;
        for (int i = 0; i < ncomp; ++i) {
            JComp& c = comp[i];
            c.id = (int)in.u8(); if (ncomp == 3 && c.id == rgb[i]) ++rgb_ids;
            const unsigned hv = in.u8(); c.h = hv >> 4; c.v = hv & 15; c.tq = (int)in.u8();
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) fail("JPEG: bad component");
            h_max = std::max(h_max, c.h); v_max = std::max(v_max, c.v);
        }
        for (int i = 0; i < ncomp; ++i) if (h_max % comp[i].h || v_max % comp[i].v) fail("JPEG: fractional sampling ratios are not supported");
        mcu_x = (w + 8 * h_max - 1) / (8 * h_max); mcu_y = (h + 8 * v_max - 1) / (8 * v_max);
        for (int i = 0; i < ncomp; ++i) {
            JComp& c = comp[i];
            c.x = (w * c.h + h_max - 1) / h_max; c.y = (h * c.v + v_max - 1) / v_max;
            c.w2 = mcu_x * c.h * 8; c.h2 = mcu_y * c.v * 8;
            c.data.assign((size_t)c.w2 * c.h2, 0);
            if (progressive) c.coeff.assign((size_t)c.w2 * c.h2, 0);
        }
    }
    void read_sos()
    {
        const unsigned len = in.be16();
        scan_n = (int)in.u8();
        if (scan_n < 1 || scan_n > ncomp || len != 6u + 2u * scan_n) fail("JPEG: bad SOS");
        for (int i = 0; i < scan_n; ++i) {
            const int id = (int)in.u8(); const unsigned tt = in.u8();
            int which = 0;
            while (which < ncomp && comp[which].id != id) ++which;
            if (which == ncomp) fail("JPEG: scan names an unknown component");
            comp[which].td = tt >> 4; comp[which].ta = tt & 15;
            if (comp[which].td > 3 || comp[which].ta > 3) fail("JPEG: bad table selector");
            order[i] = which;
        }
        spec_start = (int)in.u8(); spec_end = (int)in.u8();
        const unsigned a = in.u8(); succ_high = a >> 4; succ_low = a & 15;
        if (progressive) { if (spec_start > 63 || spec_end > 63 || spec_start > spec_end || succ_high > 13 || succ_low > 13) fail("JPEG: bad progressive scan"); }
        else { if (spec_start != 0 || succ_high || succ_low) fail("JPEG: bad sequential scan"); spec_end = 63; }
    }

    // between MCUs: count down the restart interval, resynchronise on RSTn; false = stop the scan
    bool mcu_done()
    {
        if (--todo > 0) return true;
        if (cnt < 24) grow();
        if (marker < 0xd0 || marker > 0xd7) return false;
        reset_entropy();
        return true;
    }
    void decode_scan()
    {
        reset_entropy();
        int16_t blk[64];
        auto one = [&](JComp& c, int bx, int by) {
            if (!progressive) {
                if (!hdc[c.td].set || !hac[c.ta].set) fail("JPEG: scan uses an undefined Huffman table");
                block_baseline(blk, c);
                idct(c.data.data() + (size_t)c.w2 * by * 8 + bx * 8, c.w2, blk);
            } else {
                int16_t* d = c.coeff.data() + 64 * ((size_t)bx + (size_t)by * (c.w2 >> 3));
                if (spec_start == 0) { if (!hdc[c.td].set && !succ_high) fail("JPEG: scan uses an undefined Huffman table"); block_dc(d, c); }
                else { if (!hac[c.ta].set) fail("JPEG: scan uses an undefined Huffman table"); block_ac(d, hac[c.ta]); }
            }
        };
        if (scan_n == 1) {
            JComp& c = comp[order[0]];
            const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
            for (int j = 0; j < bh; ++j) for (int i = 0; i < bw; ++i) { one(c, i, j); if (!mcu_done()) return; }
        } else {
            if (progressive && spec_start != 0) fail("JPEG: interleaved AC scan");
            for (int j = 0; j < mcu_y; ++j) for (int i = 0; i < mcu_x; ++i) {
                for (int k = 0; k < scan_n; ++k) {
                    JComp& c = comp[order[k]];
                    for (int y = 0; y < c.v; ++y) for (int x = 0; x < c.h; ++x) one(c, i * c.h + x, j * c.v + y);
                }
                if (!mcu_done()) return;
            }
        }
    }
    void finish_progressive()
    {
        for (int n = 0; n < ncomp; ++n) {
            JComp& c = comp[n];
            const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
            const uint16_t* q = dequant[c.tq];
            for (int j = 0; j < bh; ++j) for (int i = 0; i < bw; ++i) {
                int16_t* d = c.coeff.data() + 64 * ((size_t)i + (size_t)j * (c.w2 >> 3));
                for (int k = 0; k < 64; ++k) d[k] = (int16_t)(d[k] * q[k]);
                idct(c.data.data() + (size_t)c.w2 * j * 8 + i * 8, c.w2, d);
            }
        }
    }

    void parse()
    {
        if (in.u8() != 0xff || in.u8() != 0xd8) fail("JPEG: no SOI");
        bool have_frame = false, done = false;
        int m = -1;
        while (!done) {
            if (m < 0) {                                             // next marker
                unsigned b = in.u8();
                if (b != 0xff) continue;
                do b = in.u8(); while (b == 0xff);
                if (!b) continue;
                m = (int)b;
            }
            const int cur = m; m = -1;
            switch (cur) {
                case 0xd9: done = true; break;
                case 0xdb: read_dqt(); break;
                case 0xc4: read_dht(); break;
                case 0xdd: if (in.be16() != 4) fail("JPEG: bad DRI"); restart_interval = (int)in.be16(); break;
                case 0xc0: case 0xc1: case 0xc2:
                    if (have_frame) fail("JPEG: more than one frame");
                    progressive = cur == 0xc2; read_sof(); have_frame = true; break;
                case 0xc3: case 0xc5: case 0xc6: case 0xc7: case 0xc9: case 0xca: case 0xcb: case 0xcd: case 0xce: case 0xcf:
                    fail("JPEG: lossless / hierarchical / arithmetic coding is not supported");
                case 0xda:
                    if (!have_frame) fail("JPEG: scan before frame header");
                    read_sos(); decode_scan();
                    if (marker >= 0) m = marker;
                    else {                                           // skip to the next marker
                        while (!in.eof()) { if (in.u8() == 0xff) { unsigned b = in.u8_or0(); while (b == 0xff) b = in.u8_or0(); if (b && (b < 0xd0 || b > 0xd7)) { m = (int)b; break; } } }
                        if (m < 0) done = true;
                    }
                    break;
                case 0xe0: {
                    const unsigned len = in.be16();
                    const uint8_t* q = in.take(len - 2);
                    if (len >= 7 && !std::memcmp(q, "JFIF\0", 5)) jfif = true;
                } break;
                case 0xee: {
                    const unsigned len = in.be16();
                    const uint8_t* q = in.take(len - 2);
                    if (len >= 14 && !std::memcmp(q, "Adobe\0", 6)) adobe_transform = q[11];
                } break;
                default:
                    if ((cur >= 0xd0 && cur <= 0xd7) || cur == 0x01) break;            // stand-alone markers
                    { const unsigned len = in.be16(); if (len < 2) fail("JPEG: bad segment length"); in.skip(len - 2); }
            }
        }
        if (!have_frame) fail("JPEG: no frame");
        if (progressive) finish_progressive();
    }

    // ---- chroma upsampling: one output row of width w from the component rows `near` / `far`
    static const uint8_t* up_row(uint8_t* out, const uint8_t* near, const uint8_t* far, int n, int hs, int vs)
    {
        if (hs == 1 && vs == 1) return near;
        if (hs == 1 && vs == 2) { for (int i = 0; i < n; ++i) out[i] = (uint8_t)((3 * near[i] + far[i] + 2) >> 2); return out; }
        if (hs == 2 && vs == 1) {
            if (n == 1) { out[0] = out[1] = near[0]; return out; }
            out[0] = near[0]; out[1] = (uint8_t)((near[0] * 3 + near[1] + 2) >> 2);
            int i = 1;
            for (; i < n - 1; ++i) { const int t = 3 * near[i] + 2; out[2 * i] = (uint8_t)((t + near[i - 1]) >> 2); out[2 * i + 1] = (uint8_t)((t + near[i + 1]) >> 2); }
            out[2 * i] = (uint8_t)((3 * near[n - 2] + near[n - 1] + 2) >> 2); out[2 * i + 1] = near[n - 1];    // sic: the reference's decoder weights the last pair this way
            return out;
        }
        if (hs == 2 && vs == 2) {
            int t1 = 3 * near[0] + far[0];
            if (n == 1) { out[0] = out[1] = (uint8_t)((t1 + 2) >> 2); return out; }
            out[0] = (uint8_t)((t1 + 2) >> 2);
            for (int i = 1; i < n; ++i) {
                const int t0 = t1; t1 = 3 * near[i] + far[i];
                out[2 * i - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4); out[2 * i] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
            }
            out[2 * n - 1] = (uint8_t)((t1 + 2) >> 2);
            return out;
        }
        for (int i = 0; i < n; ++i) for (int j = 0; j < hs; ++j) out[i * hs + j] = near[i];
        return out;
    }

    void to_rgba(std::vector<uint32_t>& px)
    {
        const bool is_rgb = ncomp == 3 && (rgb_ids == 3 || (adobe_transform == 0 && !jfif));
        struct Up { int hs, vs, ystep, ypos, n; const uint8_t *line0, *line1; std::vector<uint8_t> buf; } up[3];
        for (int k = 0; k < ncomp; ++k) {
            Up& u = up[k]; const JComp& c = comp[k];
            u.hs = h_max / c.h; u.vs = v_max / c.v; u.ystep = u.vs >> 1; u.ypos = 0;
            u.n = (w + u.hs - 1) / u.hs;
            u.line0 = u.line1 = c.data.data();
            u.buf.assign((size_t)w + 8, 0);
        }
        px.resize((size_t)w * h);
        const int kCr_r = 5743 << 8, kCr_g = -(2925 << 8), kCb_g = -(1410 << 8), kCb_b = 7258 << 8;      // round(x * 4096) << 8
        for (int j = 0; j < h; ++j) {
            const uint8_t* row[3] = {nullptr, nullptr, nullptr};
            for (int k = 0; k < ncomp; ++k) {
                Up& u = up[k];
                const bool bottom = u.ystep >= (u.vs >> 1);
                row[k] = up_row(u.buf.data(), bottom ? u.line1 : u.line0, bottom ? u.line0 : u.line1, u.n, u.hs, u.vs);
                if (++u.ystep >= u.vs) { u.ystep = 0; u.line0 = u.line1; if (++u.ypos < comp[k].y) u.line1 += comp[k].w2; }
            }
            uint32_t* out = px.data() + (size_t)j * w;
            if (ncomp == 1) for (int i = 0; i < w; ++i) out[i] = rgba(row[0][i], row[0][i], row[0][i], 255);
            else if (is_rgb) for (int i = 0; i < w; ++i) out[i] = rgba(row[0][i], row[1][i], row[2][i], 255);
            else for (int i = 0; i < w; ++i) {
                const int yf = (row[0][i] << 20) + (1 << 19), cb = row[1][i] - 128, cr = row[2][i] - 128;
                const int r = (yf + cr * kCr_r) >> 20;
                const int g = (yf + cr * kCr_g + (int)((unsigned)(cb * kCb_g) & 0xffff0000u)) >> 20;
                const int b = (yf + cb * kCb_b) >> 20;
                out[i] = rgba(clamp8(r), clamp8(g), clamp8(b), 255);
            }
        }
    }
};

void decode_jpeg(const uint8_t* d, size_t n, int& w, int& h, std::vector<uint32_t>& px)
{
    Jpeg j(d, n);
    j.parse();
    w = j.w; h = j.h;
    j.to_rgba(px);
}

// ---------------------------------------------------------------- BMP (uncompressed 24 / 32 bit) and binary PPM
void decode_bmp(const uint8_t* d, size_t n, int& w, int& h, std::vector<uint32_t>& px)
{
    Bytes in(d, n);
    in.skip(10);
    const unsigned off = in.le32(), hsz = in.le32();
    if (hsz < 40) fail("BMP: unsupported header");
    w = (int)in.le32(); int hh = (int)in.le32(); in.skip(2);
    const unsigned bpp = in.le16(), comp = in.le32();
    const bool top_down = hh < 0; h = top_down ? -hh : hh;
    if ((bpp != 24 && bpp != 32) || comp != 0 || w <= 0 || h <= 0) fail("BMP: only uncompressed 24 / 32 bit images are supported");
    check_size(w, h);
    const size_t stride = ((size_t)w * (bpp / 8) + 3) & ~(size_t)3;
    in.pos = off;
    px.resize((size_t)w * h);
    unsigned any_alpha = 0;
    for (int y = 0; y < h; ++y) {
        const uint8_t* row = in.take(y == h - 1 ? (size_t)w * (bpp / 8) : stride);
        uint32_t* out = px.data() + (size_t)(top_down ? y : h - 1 - y) * w;
        for (int x = 0; x < w; ++x) {
            const uint8_t* p = row + (size_t)x * (bpp / 8);
            const unsigned a = bpp == 32 ? p[3] : 255u;
            any_alpha |= a;
            out[x] = rgba(p[2], p[1], p[0], a);
        }
    }
    if (!any_alpha) for (auto& v : px) v |= 0xff000000u;            // 32-bit files that leave the fourth byte unused
}

void decode_ppm(const uint8_t* d, size_t n, int& w, int& h, std::vector<uint32_t>& px)
{
    Bytes in(d, n); in.skip(2);
    auto next_int = [&]() {
        unsigned c = in.u8();
        for (;;) {
            if (c == '#') { while (c != '\n') c = in.u8(); }
            else if (c == ' ' || c == '\t' || c == '\r' || c == '\n') c = in.u8();
            else break;
        }
        if (c < '0' || c > '9') fail("PPM: bad header");
        int v = 0;
        while (c >= '0' && c <= '9') { v = v * 10 + (int)(c - '0'); if (v > (1 << 24)) fail("PPM: bad header"); c = in.u8(); }
        return v;                                                   // the one whitespace byte after the number is consumed
    };
    w = next_int(); h = next_int();
    if (next_int() != 255 || w <= 0 || h <= 0) fail("PPM: only 8-bit P6 files are supported");
    check_size(w, h);
    const uint8_t* q = in.take((size_t)w * h * 3);
    px.resize((size_t)w * h);
    for (size_t i = 0; i < px.size(); ++i) px[i] = rgba(q[3 * i], q[3 * i + 1], q[3 * i + 2], 255);
}

}  // namespace

bool decode_memory(const unsigned char* d, size_t n, int& w, int& h, std::vector<uint32_t>& px, std::string* why)
{
    try {
        TgaHeader t;
        if (is_png(d, n)) decode_png(d, n, w, h, px);
        else if (n >= 2 && d[0] == 'B' && d[1] == 'M') decode_bmp(d, n, w, h, px);
        else if (n >= 3 && d[0] == 0xff && d[1] == 0xd8 && d[2] == 0xff) decode_jpeg(d, n, w, h, px);
        else if (n >= 2 && d[0] == 'P' && d[1] == '6') decode_ppm(d, n, w, h, px);
        else if (read_tga_header(d, n, t)) decode_tga(d, n, w, h, px);
        else fail("unknown image format");
        return true;
    } catch (const Fail& f) {
        if (why) *why = f.why;
    } catch (const std::exception& e) {
        if (why) *why = e.what();
    }
    px.clear(); w = h = 0;
    return false;
}

bool decode_file(const std::string& path, int& w, int& h, std::vector<uint32_t>& px, std::string* why)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { if (why) *why = "cannot open file"; px.clear(); w = h = 0; return false; }
    std::vector<unsigned char> data;
    unsigned char chunk[1 << 16];
    for (size_t got; (got = std::fread(chunk, 1, sizeof(chunk), f)) > 0;) data.insert(data.end(), chunk, chunk + got);
    std::fclose(f);
    return decode_memory(data.data(), data.size(), w, h, px, why);
}

}  // namespace fovpt_image
