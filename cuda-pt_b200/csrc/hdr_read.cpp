// hdr_read.cpp — Radiance RGBE (.hdr) images for the XML scene loader, tone-mapped to 8 bits the way the decoder the
// reference links does it when an .hdr file is loaded through its 8-bit entry point (src/impl/textures.cu:59,100 call
// stbi_load for every texture): RGBE -> float (mantissa * 2^(e - 136); grey = (r + g + b) * scale / 3), then
// byte = pow(value, 1 / 2.2) * 255 + 0.5 clamped (alpha: value * 255 + 0.5).  Flat and run-length-encoded scanlines.
// tests/test_host.py checks the bytes against the reference's decoder (`ref_dump imgdump`).
#include "image_io.h"

#include <cmath>
#include <cstring>
#include <stdexcept>

namespace cptb {

std::vector<uint8_t> decode_hdr(const std::vector<uint8_t> &file, const std::string &path, int want, int &width, int &height) {
    auto fail = [&](const std::string &why) -> void { throw std::runtime_error("image: HDR '" + path + "': " + why); };
    size_t pos = 0;
    auto line = [&]() {
        std::string s;
        while (pos < file.size() && file[pos] != '\n')
            s.push_back(char(file[pos++]));
        if (pos < file.size())
            pos++;
        return s;
    };
    const std::string magic = line();
    if (magic != "#?RADIANCE" && magic != "#?RGBE")
        fail("not a Radiance file");
    bool format_ok = false;
    while (true) {
        if (pos >= file.size())
            fail("truncated header");
        const std::string l = line();
        if (l.empty())
            break;
        if (l == "FORMAT=32-bit_rle_rgbe")
            format_ok = true;
    }
    if (!format_ok)
        fail("unsupported format (32-bit_rle_rgbe is)");
    const std::string dims = line();
    int w = 0, h = 0;
    if (sscanf(dims.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0)
        fail("unsupported orientation (only -Y h +X w)");
    width = w, height = h;
    const int n = want ? want : 3;
    std::vector<float> px(size_t(w) * h * n);
    auto convert = [&](float *o, const uint8_t *rgbe) {
        if (rgbe[3] != 0) {
            const float f1 = float(std::ldexp(1.0f, int(rgbe[3]) - (128 + 8)));
            if (n <= 2)
                o[0] = (rgbe[0] + rgbe[1] + rgbe[2]) * f1 / 3;
            else
                o[0] = rgbe[0] * f1, o[1] = rgbe[1] * f1, o[2] = rgbe[2] * f1;
            if (n == 2)
                o[1] = 1;
            if (n == 4)
                o[3] = 1;
        } else {
            if (n >= 3)
                o[0] = o[1] = o[2] = 0;
            else
                o[0] = 0;
            if (n == 2)
                o[1] = 1;
            if (n == 4)
                o[3] = 1;
        }
    };
    auto need = [&](size_t bytes) {
        if (pos + bytes > file.size())
            fail("truncated pixel data");
    };
    bool flat = w < 8 || w >= 32768;
    int j0 = 0;
    uint8_t first[4];
    bool have_first = false;
    std::vector<uint8_t> scan(size_t(w) * 4);
    for (int j = 0; j < h && !flat; j++) {
        need(4);
        const uint8_t c1 = file[pos], c2 = file[pos + 1], len = file[pos + 2];
        if (c1 != 2 || c2 != 2 || (len & 0x80)) {
            // not run-length encoded: these four bytes are the first pixel of a flat image (only legal on the first line)
            memcpy(first, &file[pos], 4);
            pos += 4;
            have_first = true;
            flat = true;
            j0 = j;
            break;
        }
        if (((int(len) << 8) | file[pos + 3]) != w)
            fail("corrupt scanline width");
        pos += 4;
        for (int k = 0; k < 4; k++) {
            int i = 0;
            while (i < w) {
                need(1);
                int count = file[pos++];
                if (count > 128) {
                    need(1);
                    const uint8_t value = file[pos++];
                    count -= 128;
                    if (count > w - i)
                        fail("corrupt run");
                    for (int z = 0; z < count; z++)
                        scan[size_t(i++) * 4 + k] = value;
                } else {
                    if (count == 0 || count > w - i)
                        fail("corrupt run");
                    need(size_t(count));
                    for (int z = 0; z < count; z++)
                        scan[size_t(i++) * 4 + k] = file[pos++];
                }
            }
        }
        for (int i = 0; i < w; i++)
            convert(&px[(size_t(j) * w + i) * n], &scan[size_t(i) * 4]);
    }
    if (flat) {
        if (j0 != 0)
            fail("flat data after run-length encoded scanlines");
        for (int j = 0; j < h; j++)
            for (int i = 0; i < w; i++) {
                uint8_t rgbe[4];
                if (have_first) {
                    memcpy(rgbe, first, 4);
                    have_first = false;
                } else {
                    need(4);
                    memcpy(rgbe, &file[pos], 4);
                    pos += 4;
                }
                convert(&px[(size_t(j) * w + i) * n], rgbe);
            }
    }
    // float -> 8 bit: gamma 1 / 2.2 on the colour channels, linear alpha
    std::vector<uint8_t> out(px.size());
    const int colour = (n & 1) ? n : n - 1;
    const float inv_gamma = 1.0f / 2.2f;
    for (size_t i = 0; i < size_t(w) * h; i++) {
        int k = 0;
        for (; k < colour; k++) {
            float z = float(std::pow(px[i * n + k] * 1.0f, inv_gamma)) * 255 + 0.5f;
            z = z < 0 ? 0 : (z > 255 ? 255 : z);
            out[i * n + k] = uint8_t(int(z));
        }
        if (k < n) {
            float z = px[i * n + k] * 255 + 0.5f;
            z = z < 0 ? 0 : (z > 255 ? 255 : z);
            out[i * n + k] = uint8_t(int(z));
        }
    }
    return out;
}

} // namespace cptb
