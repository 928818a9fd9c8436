// 8-bit image input for the XML scene loader: a small PNG reader on top of zlib's inflate.
//
// The reference decodes every texture with stb_image to 8 bits per channel, forcing 4 channels for colour
// maps and 1 channel for roughness / ior maps (src/impl/textures.cu:53-81,92-131).  To give the same
// texels for the same file this reader follows the PNG specification for decoding and stb_image's
// documented channel-conversion rules (grey = (77 r + 150 g + 29 b) >> 8, opaque alpha = 255, 16-bit
// samples keep their high byte).  JPEG files go to jpeg_read.cpp, Radiance .hdr files to hdr_read.cpp.
#include "image_io.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <zlib.h>

namespace cptb {

namespace {

uint32_t be32(const uint8_t *p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

std::vector<uint8_t> read_file(const std::string &path) {
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp)
        throw std::runtime_error("image: cannot open '" + path + "'");
    std::vector<uint8_t> buf;
    uint8_t tmp[65536];
    size_t n;
    while ((n = fread(tmp, 1, sizeof(tmp), fp)) > 0)
        buf.insert(buf.end(), tmp, tmp + n);
    fclose(fp);
    return buf;
}

std::vector<uint8_t> inflate_all(const std::vector<uint8_t> &src, size_t expected, const std::string &path) {
    std::vector<uint8_t> dst(expected);
    z_stream zs{};
    if (inflateInit(&zs) != Z_OK)
        throw std::runtime_error("image: zlib init failed");
    zs.next_in = const_cast<Bytef *>(src.data());
    zs.avail_in = uInt(src.size());
    zs.next_out = dst.data();
    zs.avail_out = uInt(dst.size());
    const int rc = inflate(&zs, Z_FINISH);
    const size_t produced = zs.total_out;
    inflateEnd(&zs);
    if ((rc != Z_STREAM_END && rc != Z_OK && rc != Z_BUF_ERROR) || produced < expected)
        throw std::runtime_error("image: corrupt PNG data stream in '" + path + "'");
    return dst;
}

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    if (pa <= pb && pa <= pc)
        return a;
    return pb <= pc ? b : c;
}

} // namespace

Image8 decode_png(const std::string &path) {
    const std::vector<uint8_t> file = read_file(path);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 || memcmp(file.data(), sig, 8) != 0)
        throw std::runtime_error("image: '" + path + "' is not a PNG file (only PNG textures are supported)");
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    bool have_hdr = false;
    for (size_t pos = 8; pos + 12 <= file.size();) {
        const uint32_t len = be32(&file[pos]);
        const char *tag = reinterpret_cast<const char *>(&file[pos + 4]);
        const uint8_t *body = &file[pos + 8];
        if (pos + 12 + size_t(len) > file.size())
            throw std::runtime_error("image: truncated PNG '" + path + "'");
        if (!memcmp(tag, "IHDR", 4)) {
            if (len < 13)
                throw std::runtime_error("image: bad PNG header in '" + path + "'");
            w = be32(body), h = be32(body + 4);
            depth = body[8], ctype = body[9], interlace = body[12];
            have_hdr = true;
        } else if (!memcmp(tag, "PLTE", 4)) {
            palette.assign(body, body + len);
        } else if (!memcmp(tag, "tRNS", 4)) {
            trns.assign(body, body + len);
        } else if (!memcmp(tag, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!memcmp(tag, "IEND", 4)) {
            break;
        }
        pos += 12 + size_t(len);
    }
    if (!have_hdr || w == 0 || h == 0 || w > (1u << 24) || h > (1u << 24))
        throw std::runtime_error("image: bad PNG dimensions in '" + path + "'");
    if (interlace > 1)
        throw std::runtime_error("image: bad PNG interlace method in '" + path + "'");
    int samples;
    switch (ctype) {
    case 0: samples = 1; break;
    case 2: samples = 3; break;
    case 3: samples = 1; break;
    case 4: samples = 2; break;
    case 6: samples = 4; break;
    default: throw std::runtime_error("image: bad PNG colour type in '" + path + "'");
    }
    if (!(depth == 8 || depth == 16 || (depth < 8 && (ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4))))
        throw std::runtime_error("image: unsupported PNG bit depth in '" + path + "'");
    const size_t bits_pp = size_t(samples) * depth, stride = (bits_pp * w + 7) / 8, bpp = bits_pp >= 8 ? bits_pp / 8 : 1;

    // undo the per-scanline filters of one (sub-)image (PNG specification, section 9)
    auto unfilter = [&](const uint8_t *raw, uint32_t pw, uint32_t ph, uint8_t *dst) {
        const size_t ps = (bits_pp * pw + 7) / 8;
        for (uint32_t y = 0; y < ph; y++) {
            const uint8_t ft = raw[y * (ps + 1)];
            const uint8_t *in = &raw[y * (ps + 1) + 1];
            uint8_t *cur = dst + size_t(y) * ps;
            const uint8_t *up = y ? cur - ps : nullptr;
            for (size_t i = 0; i < ps; i++) {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
                int v = in[i];
                switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: throw std::runtime_error("image: bad PNG filter in '" + path + "'");
                }
                cur[i] = uint8_t(v);
            }
        }
    };
    std::vector<uint8_t> pix(stride * h);
    if (interlace == 0) {
        std::vector<uint8_t> raw = inflate_all(idat, (stride + 1) * h, path);
        unfilter(raw.data(), w, h, pix.data());
    } else { // Adam7: seven reduced images, each filtered on its own, pixels scattered back to their places
        static const int xo[7] = {0, 4, 0, 2, 0, 1, 0}, yo[7] = {0, 0, 4, 0, 2, 0, 1}, xs[7] = {8, 8, 4, 4, 2, 2, 1}, ys[7] = {8, 8, 8, 4, 4, 2, 2};
        size_t total = 0;
        uint32_t pw[7], ph[7];
        for (int p = 0; p < 7; p++) {
            pw[p] = (w > uint32_t(xo[p])) ? (w - xo[p] + xs[p] - 1) / xs[p] : 0;
            ph[p] = (h > uint32_t(yo[p])) ? (h - yo[p] + ys[p] - 1) / ys[p] : 0;
            if (pw[p] && ph[p])
                total += ((bits_pp * pw[p] + 7) / 8 + 1) * ph[p];
        }
        std::vector<uint8_t> raw = inflate_all(idat, total, path);
        size_t off = 0;
        for (int p = 0; p < 7; p++) {
            if (!pw[p] || !ph[p])
                continue;
            const size_t ps = (bits_pp * pw[p] + 7) / 8;
            std::vector<uint8_t> sub(ps * ph[p]);
            unfilter(raw.data() + off, pw[p], ph[p], sub.data());
            off += (ps + 1) * ph[p];
            for (uint32_t j = 0; j < ph[p]; j++)
                for (uint32_t i = 0; i < pw[p]; i++) {
                    const size_t x = size_t(xo[p]) + size_t(i) * xs[p], y = size_t(yo[p]) + size_t(j) * ys[p];
                    if (bits_pp >= 8) {
                        memcpy(&pix[y * stride + x * bpp], &sub[size_t(j) * ps + size_t(i) * bpp], bpp);
                    } else {
                        const size_t sb = size_t(i) * depth, db = x * depth;
                        const unsigned v = (sub[size_t(j) * ps + sb / 8] >> (8 - depth - sb % 8)) & ((1u << depth) - 1);
                        uint8_t &d = pix[y * stride + db / 8];
                        const int shift = int(8 - depth - db % 8);
                        d = uint8_t((d & ~(((1u << depth) - 1) << shift)) | (v << shift));
                    }
                }
        }
    }

    // expand to 8-bit samples; palette images become RGB(A), a colour key becomes an alpha channel
    Image8 img;
    img.width = int(w), img.height = int(h);
    const size_t n = size_t(w) * h;
    if (ctype == 3) {
        const bool alpha = !trns.empty();
        img.channels = alpha ? 4 : 3;
        img.data.resize(n * img.channels);
        for (uint32_t y = 0; y < h; y++)
            for (uint32_t x = 0; x < w; x++) {
                const size_t bit = size_t(x) * depth;
                const uint8_t byte = pix[size_t(y) * stride + bit / 8];
                const unsigned idx = depth == 8 ? byte : (byte >> (8 - depth - bit % 8)) & ((1u << depth) - 1);
                uint8_t *o = &img.data[(size_t(y) * w + x) * img.channels];
                for (int k = 0; k < 3; k++)
                    o[k] = size_t(idx) * 3 + k < palette.size() ? palette[idx * 3 + k] : 0;
                if (alpha)
                    o[3] = idx < trns.size() ? trns[idx] : 255;
            }
        return img;
    }
    const bool key = !trns.empty() && (ctype == 0 || ctype == 2) && trns.size() >= size_t(samples) * 2;
    img.channels = samples + (key ? 1 : 0);
    img.data.resize(n * img.channels);
    const int scale = depth == 1 ? 255 : (depth == 2 ? 85 : (depth == 4 ? 17 : 1));
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            uint8_t *o = &img.data[(size_t(y) * w + x) * img.channels];
            bool is_key = key;
            for (int s = 0; s < samples; s++) {
                unsigned full;
                if (depth == 16) {
                    const uint8_t *p = &pix[size_t(y) * stride + (size_t(x) * samples + s) * 2];
                    full = (unsigned(p[0]) << 8) | p[1];
                    o[s] = p[0];
                } else if (depth == 8) {
                    full = pix[size_t(y) * stride + size_t(x) * samples + s];
                    o[s] = uint8_t(full);
                } else {
                    const size_t bit = size_t(x) * depth;
                    full = (pix[size_t(y) * stride + bit / 8] >> (8 - depth - bit % 8)) & ((1u << depth) - 1);
                    o[s] = uint8_t(full * scale);
                }
                if (key)
                    is_key = is_key && full == ((unsigned(trns[2 * s]) << 8) | trns[2 * s + 1]);
            }
            if (key)
                o[samples] = is_key ? 0 : 255;
        }
    return img;
}

std::vector<uint8_t> convert_channels(const Image8 &img, int want) {
    const size_t n = size_t(img.width) * img.height;
    const int have = img.channels;
    if (want == have)
        return img.data;
    std::vector<uint8_t> out(n * want);
    auto luma = [](int r, int g, int b) { return uint8_t((r * 77 + g * 150 + 29 * b) >> 8); };
    for (size_t i = 0; i < n; i++) {
        const uint8_t *s = &img.data[i * have];
        uint8_t *d = &out[i * want];
        const bool grey_src = have <= 2;
        const uint8_t a = (have == 2) ? s[1] : (have == 4 ? s[3] : 255);
        if (want <= 2) {
            d[0] = grey_src ? s[0] : luma(s[0], s[1], s[2]);
            if (want == 2)
                d[1] = a;
        } else {
            d[0] = s[0], d[1] = grey_src ? s[0] : s[1], d[2] = grey_src ? s[0] : s[2];
            if (want == 4)
                d[3] = a;
        }
    }
    return out;
}

std::vector<uint8_t> load_image8(const std::string &path, int want_channels, int &width, int &height) {
    {
        FILE *fp = fopen(path.c_str(), "rb");
        uint8_t magic[2] = {0, 0};
        const bool got = fp && fread(magic, 1, 2, fp) == 2;
        if (fp)
            fclose(fp);
        if (got && magic[0] == 0xFF && magic[1] == 0xD8)
            return decode_jpeg(read_file(path), path, want_channels, width, height);
        if (got && magic[0] == '#' && magic[1] == '?')
            return decode_hdr(read_file(path), path, want_channels, width, height);
    }
    const Image8 img = decode_png(path);
    width = img.width, height = img.height;
    return convert_channels(img, want_channels);
}

} // namespace cptb
