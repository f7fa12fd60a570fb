// chrono_b200_imageio.hpp -- frame files in, composite files out, for the CLI-compatible driver (tools/chrono_b200_cli.cpp).
//
// Mirrors what the reference delegates to the `image` crate around the path:
//   read_image   image::open (src/streams.rs:63-69, src/simple.rs:45, src/shake.rs:222-283) followed by as_flat_samples_u8():
//                8-bit RGB or RGBA frames. PPM (P6), PNG (zlib inflate + un-filtering, non-interlaced), JPEG (nvJPEG through the
//                library's chb_decode_jpeg). Anything else is the reference's "Unexpected format. Not an 8 bit image." error.
//   save_image   src/main.rs:520-571: format by (lower-cased) extension, `--quality` for jpg/jpeg (nvJPEG through chb_encode_jpeg),
//                PNG (zlib deflate), baseline uncompressed TIFF, 24/32-bit BMP, PPM; the output directory is created when missing.
// Host code only; links zlib (-lz). No pixel arithmetic happens here.
#pragma once
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "chrono_b200.hpp"

namespace chrono_b200 {

struct Image {
    int w = 0, h = 0, c = 3;
    std::vector<uint8_t> px;  // h rows of w * c interleaved bytes
};

inline std::string lower_extension(const std::string& path) {
    const size_t slash = path.find_last_of('/');
    const size_t dot = path.find_last_of('.');
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return "";
    std::string e = path.substr(dot + 1);
    std::transform(e.begin(), e.end(), e.begin(), [](unsigned char ch) { return (char)std::tolower(ch); });
    return e;
}
inline std::vector<uint8_t> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("Unable to open image " + path);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
inline void write_file(const std::string& path, const std::vector<uint8_t>& bytes) {
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("Unable to create output file " + path);
    f.write(reinterpret_cast<const char*>(bytes.data()), (std::streamsize)bytes.size());
    if (!f) throw std::runtime_error("Unable to write output file " + path);
}

// ---------------------------------------------------------------------------------------------------------------- PPM
inline Image decode_ppm(const std::vector<uint8_t>& d, const std::string& path) {
    size_t pos = 0;
    auto token = [&]() {
        while (pos < d.size()) {
            if (d[pos] == '#') { while (pos < d.size() && d[pos] != '\n') pos++; }
            else if (std::isspace(d[pos])) pos++;
            else break;
        }
        std::string t;
        while (pos < d.size() && !std::isspace(d[pos])) t.push_back((char)d[pos++]);
        return t;
    };
    if (token() != "P6") throw std::runtime_error("Unexpected format. Not a binary PPM (P6): " + path);
    Image im;
    im.w = std::atoi(token().c_str()); im.h = std::atoi(token().c_str());
    if (std::atoi(token().c_str()) != 255) throw std::runtime_error("Unexpected format. Not an 8 bit image.");
    pos++;  // the single whitespace byte after maxval
    const size_t need = (size_t)im.w * im.h * 3;
    if (im.w < 1 || im.h < 1 || d.size() < pos + need) throw std::runtime_error("Truncated PPM: " + path);
    im.px.assign(d.begin() + pos, d.begin() + pos + need);
    return im;
}
inline std::vector<uint8_t> encode_ppm(int w, int h, int c, const uint8_t* px) {
    std::ostringstream hd;
    hd << "P6\n" << w << " " << h << "\n255\n";
    const std::string s = hd.str();
    std::vector<uint8_t> out(s.begin(), s.end());
    out.reserve(out.size() + (size_t)w * h * 3);
    for (size_t i = 0; i < (size_t)w * h; i++) out.insert(out.end(), px + i * c, px + i * c + 3);  // PPM has no alpha
    return out;
}

// ---------------------------------------------------------------------------------------------------------------- PNG
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }

inline Image decode_png(const std::vector<uint8_t>& d, const std::string& path) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (d.size() < 8 || std::memcmp(d.data(), sig, 8) != 0) throw std::runtime_error("Not a PNG file: " + path);
    Image im;
    std::vector<uint8_t> idat;
    int bit_depth = 0, color_type = 0, interlace = 0;
    for (size_t pos = 8; pos + 12 <= d.size();) {
        const uint32_t len = be32(&d[pos]);
        const std::string type(reinterpret_cast<const char*>(&d[pos + 4]), 4);
        if (pos + 12 + len > d.size()) throw std::runtime_error("Truncated PNG: " + path);
        const uint8_t* body = &d[pos + 8];
        if (type == "IHDR") {
            im.w = (int)be32(body); im.h = (int)be32(body + 4);
            bit_depth = body[8]; color_type = body[9]; interlace = body[12];
        } else if (type == "IDAT") {
            idat.insert(idat.end(), body, body + len);
        } else if (type == "IEND") {
            break;
        }
        pos += 12 + len;
    }
    if (bit_depth != 8 || (color_type != 2 && color_type != 6)) throw std::runtime_error("Unexpected format. Not an 8 bit image.");  // RGB8 / RGBA8 only
    if (interlace != 0) throw std::runtime_error("Interlaced PNG is not supported: " + path);
    im.c = color_type == 2 ? 3 : 4;
    const size_t stride = (size_t)im.w * im.c;
    std::vector<uint8_t> raw((stride + 1) * im.h);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size())
        throw std::runtime_error("Corrupt PNG data: " + path);
    im.px.resize(stride * im.h);
    const int bpp = im.c;
    for (int y = 0; y < im.h; y++) {  // undo the per-row filters (PNG spec 9.2)
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t* in = &raw[(stride + 1) * y + 1];
        uint8_t* out = &im.px[stride * y];
        const uint8_t* up = y ? out - stride : nullptr;
        for (size_t x = 0; x < stride; x++) {
            const int a = x >= (size_t)bpp ? out[x - bpp] : 0, b = up ? up[x] : 0, cc = (up && x >= (size_t)bpp) ? up[x - bpp] : 0;
            int pred = 0;
            switch (ft) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: {
                    const int p = a + b - cc, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - cc);
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : cc);
                    break;
                }
                default: throw std::runtime_error("Corrupt PNG filter type: " + path);
            }
            out[x] = (uint8_t)(in[x] + pred);
        }
    }
    return im;
}
inline std::vector<uint8_t> encode_png(int w, int h, int c, const uint8_t* px) {
    const size_t stride = (size_t)w * c;
    std::vector<uint8_t> raw((stride + 1) * h);
    for (int y = 0; y < h; y++) {  // filter type 1 (Sub) rows compress well on photographs and need no second row
        uint8_t* o = &raw[(stride + 1) * y];
        const uint8_t* in = px + stride * y;
        o[0] = 1;
        for (size_t x = 0; x < stride; x++) o[1 + x] = (uint8_t)(in[x] - (x >= (size_t)c ? in[x - c] : 0));
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) throw std::runtime_error("PNG deflate failed");
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    auto chunk = [&](const char* type, const std::vector<uint8_t>& body) {
        put_be32(out, (uint32_t)body.size());
        const size_t start = out.size();
        out.insert(out.end(), type, type + 4);
        out.insert(out.end(), body.begin(), body.end());
        put_be32(out, (uint32_t)crc32(0L, &out[start], (uInt)(out.size() - start)));
    };
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, (uint32_t)w); put_be32(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(c == 4 ? 6 : 2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk("IHDR", ihdr);
    z.resize(zlen);
    chunk("IDAT", z);
    chunk("IEND", {});
    return out;
}

// ---------------------------------------------------------------------------------------------------------------- TIFF / BMP (writers)
inline std::vector<uint8_t> encode_tiff(int w, int h, int c, const uint8_t* px) {  // baseline, little endian, one uncompressed strip
    std::vector<uint8_t> out = {'I', 'I', 42, 0, 8, 0, 0, 0};
    auto le16 = [&](uint32_t v) { out.push_back(v & 0xff); out.push_back((v >> 8) & 0xff); };
    auto le32 = [&](uint32_t v) { le16(v & 0xffff); le16(v >> 16); };
    const int n_tags = c == 4 ? 11 : 10;
    const uint32_t ifd_end = 8 + 2 + 12 * n_tags + 4;
    const uint32_t bps_off = ifd_end, data_off = bps_off + 2 * c;
    auto tag = [&](uint32_t id, uint32_t type, uint32_t count, uint32_t value) { le16(id); le16(type); le32(count); if (type == 3 && count == 1) { le16(value); le16(0); } else le32(value); };
    le16(n_tags);
    tag(256, 4, 1, (uint32_t)w);            // ImageWidth
    tag(257, 4, 1, (uint32_t)h);            // ImageLength
    tag(258, 3, (uint32_t)c, bps_off);      // BitsPerSample -> 8,8,8[,8]
    tag(259, 3, 1, 1);                      // Compression: none
    tag(262, 3, 1, 2);                      // PhotometricInterpretation: RGB
    tag(273, 4, 1, data_off);               // StripOffsets
    tag(277, 3, 1, (uint32_t)c);            // SamplesPerPixel
    tag(278, 4, 1, (uint32_t)h);            // RowsPerStrip
    tag(279, 4, 1, (uint32_t)((size_t)w * h * c));  // StripByteCounts
    tag(284, 3, 1, 1);                      // PlanarConfiguration: chunky
    if (c == 4) tag(338, 3, 1, 2);          // ExtraSamples: unassociated alpha
    le32(0);
    for (int i = 0; i < c; i++) le16(8);
    out.insert(out.end(), px, px + (size_t)w * h * c);
    return out;
}
inline std::vector<uint8_t> encode_bmp(int w, int h, int c, const uint8_t* px) {
    const int bpp = c == 4 ? 4 : 3;
    const size_t row = ((size_t)w * bpp + 3) & ~(size_t)3;
    std::vector<uint8_t> out;
    auto le16 = [&](uint32_t v) { out.push_back(v & 0xff); out.push_back((v >> 8) & 0xff); };
    auto le32 = [&](uint32_t v) { le16(v & 0xffff); le16(v >> 16); };
    out.push_back('B'); out.push_back('M');
    le32((uint32_t)(54 + row * h)); le32(0); le32(54);
    le32(40); le32((uint32_t)w); le32((uint32_t)h); le16(1); le16(8 * bpp); le32(0); le32((uint32_t)(row * h)); le32(2835); le32(2835); le32(0); le32(0);
    std::vector<uint8_t> line(row, 0);
    for (int y = h - 1; y >= 0; y--) {  // bottom-up, BGR[A]
        const uint8_t* in = px + (size_t)y * w * c;
        for (int x = 0; x < w; x++) {
            line[x * bpp] = in[x * c + 2]; line[x * bpp + 1] = in[x * c + 1]; line[x * bpp + 2] = in[x * c];
            if (bpp == 4) line[x * bpp + 3] = in[x * c + 3];
        }
        out.insert(out.end(), line.begin(), line.end());
    }
    return out;
}

// ---------------------------------------------------------------------------------------------------------------- by extension
inline bool is_jpeg_path(const std::string& path) {
    const std::string e = lower_extension(path);
    return e == "jpg" || e == "jpeg";
}
// image::open(path) + as_flat_samples_u8(). JPEG needs the library (GPU decode).
inline Image read_image(const std::string& path, const Context* ctx = nullptr) {
    const std::vector<uint8_t> bytes = read_file(path);
    const std::string e = lower_extension(path);
    if (e == "ppm" || e == "pnm") return decode_ppm(bytes, path);
    if (e == "png") return decode_png(bytes, path);
    if (e == "jpg" || e == "jpeg") {
        if (!ctx) throw std::runtime_error("JPEG frames are decoded on the GPU: a Context is required");
        Image im;
        check(chb_decode_jpeg(ctx->raw(), bytes.data(), bytes.size(), nullptr, 0, 0, &im.w, &im.h));
        im.px.resize((size_t)im.w * im.h * 3);
        check(chb_decode_jpeg(ctx->raw(), bytes.data(), bytes.size(), im.px.data(), im.px.size(), (size_t)im.w * 3, &im.w, &im.h));
        return im;
    }
    throw std::runtime_error("The image format could not be determined: " + path);
}

// save_image (src/main.rs:520-571)
inline void save_image(const uint8_t* buffer, int w, int h, int c, const std::string& out_path, int quality, const Context* ctx) {
    const std::string ext = lower_extension(out_path);
    if (ext.empty()) throw std::runtime_error("Expects an extension for output file to determine image format.");
    const size_t slash = out_path.find_last_of('/');
    if (slash != std::string::npos && slash > 0) {
        const std::string parent = out_path.substr(0, slash);
        struct stat sb;
        if (stat(parent.c_str(), &sb) != 0 && mkdir(parent.c_str(), 0777) != 0) throw std::runtime_error("Unable to create output directory " + parent);
    }
    if (ext == "jpg" || ext == "jpeg") {
        if (!ctx) throw std::runtime_error("JPEG output is encoded on the GPU: a Context is required");
        std::vector<uint8_t> rgb;
        const uint8_t* src = buffer;
        if (c == 4) {  // the JPEG stream carries no alpha
            rgb.resize((size_t)w * h * 3);
            for (size_t i = 0; i < (size_t)w * h; i++) std::memcpy(&rgb[3 * i], buffer + 4 * i, 3);
            src = rgb.data();
        }
        std::vector<uint8_t> out((size_t)w * h * 3 + 65536);
        size_t len = 0;
        check(chb_encode_jpeg(ctx->raw(), src, w, h, (size_t)w * 3, quality, out.data(), out.size(), &len));
        out.resize(len);
        write_file(out_path, out);
    } else if (ext == "png") write_file(out_path, encode_png(w, h, c, buffer));
    else if (ext == "tif" || ext == "tiff") write_file(out_path, encode_tiff(w, h, c, buffer));
    else if (ext == "bmp") write_file(out_path, encode_bmp(w, h, c, buffer));
    else if (ext == "ppm" || ext == "pnm") write_file(out_path, encode_ppm(w, h, c, buffer));
    else throw std::runtime_error("Unable to save output file " + out_path + ": unsupported image format ." + ext);
}

}  // namespace chrono_b200
