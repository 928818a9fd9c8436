// jpeg_read.cpp — Huffman JPEG (ITU-T T.81: baseline, extended sequential, progressive) decoder for the XML scene loader.
//
// The reference decodes its textures with stb_image (src/impl/textures.cu:27-29,59,100) and the kernels sample the
// decoded bytes, so "the same file gives the same texels" needs the same integer arithmetic after entropy decoding:
//   * dequantised coefficients are held in 16 bits,
//   * the inverse DCT is the 13-bit fixed-point Loeffler-Ligtenberg-Moschytz form (libjpeg's "islow" with constants scaled by
//     4096, column pass rounded at 2^9 and shifted by 10, row pass rounded at 2^16 (+128 level shift) and shifted by 17),
//   * chroma is upsampled with the 3:1 "triangle" filters ((3 near + far + 2) >> 2 vertically or horizontally,
//     (9 a + 3 b + 3 c + d + 8) >> 4 for 2x2),
//   * YCbCr -> RGB runs in 20-bit fixed point with the 1.402 / 0.71414 / 0.34414 / 1.772 coefficients rounded to 12 bits,
//   * a request for 1 or 2 channels from a colour file returns the luma plane itself.
// tests/test_host.py checks every case byte for byte against the reference's own decoder (`ref_dump imgdump`).
// Baseline, extended-sequential and progressive Huffman files; arithmetic-coded, lossless, 12-bit and 4-component files are
// refused.
#include "image_io.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace cptb {
namespace {

struct Huff {
    // canonical code tables: for every code length the first code and the index of its first symbol
    int mincode[17], maxcode[18], valptr[17];
    uint8_t vals[256];
    bool present = false;
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int w2 = 0, h2 = 0;        // allocated size (whole MCUs)
    int x = 0, y = 0;          // size in samples
    int dc_pred = 0;
    std::vector<uint8_t> data; // decoded plane
    std::vector<int16_t> coef; // progressive files: 64 coefficients per block of the padded plane, until the last scan
};

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Decoder {
    const std::vector<uint8_t> &buf;
    const std::string &path;
    size_t pos = 0;
    uint16_t quant[4][64];
    bool quant_present[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    Component comp[3];
    int ncomp = 0, width = 0, height = 0, hmax = 1, vmax = 1, restart_interval = 0;
    bool is_rgb = false, saw_jfif = false, progressive = false;
    int ss = 0, se = 63, ah = 0, al = 0, eobrun = 0; // progressive scan parameters
    int adobe_transform = -1;
    // bit reader
    uint32_t bitbuf = 0;
    int bitcnt = 0;
    bool hit_marker = false;

    [[noreturn]] void fail(const std::string &why) const { throw std::runtime_error("image: JPEG '" + path + "': " + why); }
    uint8_t byte() {
        if (pos >= buf.size())
            fail("truncated file");
        return buf[pos++];
    }
    int be16() {
        const int a = byte();
        return (a << 8) | byte();
    }

    // ---- entropy-coded segment: bytes with FF00 unstuffing; a marker stops the supply (zeros follow)
    void fill() {
        while (bitcnt <= 24) {
            int b = 0;
            if (!hit_marker && pos < buf.size()) {
                b = buf[pos++];
                if (b == 0xFF) {
                    int c = pos < buf.size() ? buf[pos] : 0xD9;
                    while (c == 0xFF && pos + 1 < buf.size()) // fill bytes
                        c = buf[++pos];
                    if (c == 0) {
                        pos++;
                    } else {
                        hit_marker = true; // leave pos at the marker code
                        b = 0;
                    }
                }
            }
            bitbuf |= uint32_t(b) << (24 - bitcnt);
            bitcnt += 8;
        }
    }
    int bits(int n) {
        if (n == 0)
            return 0;
        if (bitcnt < n)
            fill();
        const int v = int(bitbuf >> (32 - n));
        bitbuf <<= n;
        bitcnt -= n;
        return v;
    }
    int decode(const Huff &h) {
        if (!h.present)
            fail("scan uses a Huffman table that was not defined");
        int code = 0;
        for (int len = 1; len <= 16; len++) {
            code = (code << 1) | bits(1);
            if (h.maxcode[len] >= 0 && code <= h.maxcode[len] && code >= h.mincode[len])
                return h.vals[h.valptr[len] + code - h.mincode[len]];
        }
        fail("bad Huffman code");
    }
    static int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }

    void build(Huff &h, const uint8_t *counts, const uint8_t *vals, int total) {
        memcpy(h.vals, vals, size_t(total));
        int code = 0, k = 0;
        for (int len = 1; len <= 16; len++) {
            h.valptr[len] = k;
            h.mincode[len] = code;
            code += counts[len - 1];
            k += counts[len - 1];
            h.maxcode[len] = counts[len - 1] ? code - 1 : -1;
            code <<= 1;
        }
        h.present = true;
    }

    // ---- inverse DCT of one dequantised block -> 8x8 bytes
    static inline int f2f(float x) { return int(x * 4096 + 0.5); }
    static inline uint8_t clamp(int x) { return uint8_t(x < 0 ? 0 : (x > 255 ? 255 : x)); }
    struct Odd {
        int x0, x1, x2, x3, t0, t1, t2, t3;
    };
    static inline Odd idct1d(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7) {
        int p2 = s2, p3 = s6;
        int p1 = (p2 + p3) * f2f(0.5411961f);
        int t2 = p1 + p3 * f2f(-1.847759065f);
        int t3 = p1 + p2 * f2f(0.765366865f);
        p2 = s0, p3 = s4;
        int t0 = (p2 + p3) * 4096, t1 = (p2 - p3) * 4096;
        Odd r;
        r.x0 = t0 + t3, r.x3 = t0 - t3, r.x1 = t1 + t2, r.x2 = t1 - t2;
        t0 = s7, t1 = s5, t2 = s3, t3 = s1;
        p3 = t0 + t2;
        int p4 = t1 + t3;
        p1 = t0 + t3, p2 = t1 + t2;
        const int p5 = (p3 + p4) * f2f(1.175875602f);
        t0 = t0 * f2f(0.298631336f);
        t1 = t1 * f2f(2.053119869f);
        t2 = t2 * f2f(3.072711026f);
        t3 = t3 * f2f(1.501321110f);
        p1 = p5 + p1 * f2f(-0.899976223f);
        p2 = p5 + p2 * f2f(-2.562915447f);
        p3 = p3 * f2f(-1.961570560f);
        p4 = p4 * f2f(-0.390180644f);
        r.t3 = t3 + p1 + p4, r.t2 = t2 + p2 + p3, r.t1 = t1 + p2 + p4, r.t0 = t0 + p1 + p3;
        return r;
    }
    static void idct(const int16_t *d, uint8_t *out, int stride) {
        int val[64];
        for (int i = 0; i < 8; i++) {
            int *v = val + i;
            const int16_t *c = d + i;
            if (c[8] == 0 && c[16] == 0 && c[24] == 0 && c[32] == 0 && c[40] == 0 && c[48] == 0 && c[56] == 0) {
                const int dcterm = c[0] * 4;
                v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dcterm;
            } else {
                Odd r = idct1d(c[0], c[8], c[16], c[24], c[32], c[40], c[48], c[56]);
                r.x0 += 512, r.x1 += 512, r.x2 += 512, r.x3 += 512;
                v[0] = (r.x0 + r.t3) >> 10, v[56] = (r.x0 - r.t3) >> 10;
                v[8] = (r.x1 + r.t2) >> 10, v[48] = (r.x1 - r.t2) >> 10;
                v[16] = (r.x2 + r.t1) >> 10, v[40] = (r.x2 - r.t1) >> 10;
                v[24] = (r.x3 + r.t0) >> 10, v[32] = (r.x3 - r.t0) >> 10;
            }
        }
        for (int i = 0; i < 8; i++) {
            const int *v = val + 8 * i;
            uint8_t *o = out + size_t(i) * stride;
            Odd r = idct1d(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
            const int bias = 65536 + (128 << 17);
            r.x0 += bias, r.x1 += bias, r.x2 += bias, r.x3 += bias;
            o[0] = clamp((r.x0 + r.t3) >> 17), o[7] = clamp((r.x0 - r.t3) >> 17);
            o[1] = clamp((r.x1 + r.t2) >> 17), o[6] = clamp((r.x1 - r.t2) >> 17);
            o[2] = clamp((r.x2 + r.t1) >> 17), o[5] = clamp((r.x2 - r.t1) >> 17);
            o[3] = clamp((r.x3 + r.t0) >> 17), o[4] = clamp((r.x3 - r.t0) >> 17);
        }
    }

    void decode_block(Component &c, int bx, int by) {
        int16_t d[64];
        memset(d, 0, sizeof(d));
        const uint16_t *q = quant[c.tq];
        const int t = decode(dc[c.td]);
        if (t > 15)
            fail("bad DC category");
        const int diff = t ? extend(bits(t), t) : 0;
        c.dc_pred += diff;
        d[0] = int16_t(c.dc_pred * q[0]);
        for (int k = 1; k < 64;) {
            const int rs = decode(ac[c.ta]), r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (rs != 0xF0)
                    break; // end of block
                k += 16;
                continue;
            }
            k += r;
            if (k > 63)
                fail("coefficient index past the block");
            d[kZigzag[k]] = int16_t(extend(bits(s), s) * q[k]);
            k++;
        }
        idct(d, c.data.data() + size_t(by) * 8 * c.w2 + size_t(bx) * 8, c.w2);
    }

    void restart() {
        // end of a restart interval: the next thing in the stream is RSTn (or, after the last MCU, some other marker, which
        // is left for the segment parser); predictions and the bit reader start afresh
        bitbuf = 0, bitcnt = 0;
        size_t p = pos;
        bool at_marker = hit_marker; // pos already sits on a marker code
        if (!at_marker && p < buf.size() && buf[p] == 0xFF) {
            while (p < buf.size() && buf[p] == 0xFF)
                p++;
            at_marker = p < buf.size() && buf[p] != 0;
        }
        if (at_marker && p < buf.size() && buf[p] >= 0xD0 && buf[p] <= 0xD7) {
            pos = p + 1;
            hit_marker = false;
        } else if (at_marker) {
            pos = p;
            hit_marker = true;
        }
        for (int i = 0; i < ncomp; i++)
            comp[i].dc_pred = 0;
        eobrun = 0;
    }

    // ---- progressive scans (T.81 Annex G): coefficients accumulate in c.coef, spectral band [ss, se], successive
    // approximation bit positions ah (previous) / al (this scan)
    void decode_block_progressive(Component &c, int bx, int by) {
        int16_t *d = c.coef.data() + 64 * (size_t(by) * (c.w2 >> 3) + size_t(bx));
        if (ss == 0) { // DC scan
            if (ah == 0) {
                const int t = decode(dc[c.td]);
                if (t > 15)
                    fail("bad DC category");
                const int diff = t ? extend(bits(t), t) : 0;
                c.dc_pred += diff;
                d[0] = int16_t(c.dc_pred * (1 << al));
            } else if (bits(1)) {
                d[0] = int16_t(d[0] + int16_t(1 << al));
            }
            return;
        }
        if (ah == 0) { // AC first scan of this band
            if (eobrun) {
                --eobrun;
                return;
            }
            for (int k = ss; k <= se;) {
                const int rs = decode(ac[c.ta]), r = rs >> 4, s = rs & 15;
                if (s == 0) {
                    if (r < 15) {
                        eobrun = 1 << r;
                        if (r)
                            eobrun += bits(r);
                        --eobrun;
                        break;
                    }
                    k += 16;
                } else {
                    k += r;
                    if (k > 63)
                        fail("coefficient index past the block");
                    d[kZigzag[k]] = int16_t(extend(bits(s), s) * (1 << al));
                    k++;
                }
            }
            return;
        }
        // AC refinement: one more bit for every coefficient that is already non-zero, new +-1 coefficients in between
        const int16_t bit = int16_t(1 << al);
        auto refine = [&](int16_t &p) {
            if (bits(1) && (p & bit) == 0)
                p = int16_t(p > 0 ? p + bit : p - bit);
        };
        if (eobrun) {
            --eobrun;
            for (int k = ss; k <= se; k++) {
                int16_t &p = d[kZigzag[k]];
                if (p != 0)
                    refine(p);
            }
            return;
        }
        int k = ss;
        do {
            const int rs = decode(ac[c.ta]);
            int r = rs >> 4;
            const int s = rs & 15;
            int value = 0;
            if (s == 0) {
                if (r < 15) {
                    eobrun = (1 << r) - 1;
                    if (r)
                        eobrun += bits(r);
                    r = 64; // run to the end of the band
                }
            } else {
                if (s != 1)
                    fail("bad refinement code");
                value = bits(1) ? bit : -bit;
            }
            while (k <= se) {
                int16_t &p = d[kZigzag[k++]];
                if (p != 0) {
                    refine(p);
                } else {
                    if (r == 0) {
                        p = int16_t(value);
                        break;
                    }
                    --r;
                }
            }
        } while (k <= se);
    }

    void finish_progressive() {
        for (int i = 0; i < ncomp; i++) {
            Component &c = comp[i];
            const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
            const uint16_t *q = quant[c.tq];
            for (int by = 0; by < bh; by++)
                for (int bx = 0; bx < bw; bx++) {
                    int16_t *d = c.coef.data() + 64 * (size_t(by) * (c.w2 >> 3) + size_t(bx));
                    for (int k = 0; k < 64; k++)
                        d[kZigzag[k]] = int16_t(d[kZigzag[k]] * q[k]);
                    idct(d, c.data.data() + size_t(by) * 8 * c.w2 + size_t(bx) * 8, c.w2);
                }
        }
    }

    void scan(int nscan, const int *order) {
        bitbuf = 0, bitcnt = 0, hit_marker = false, eobrun = 0;
        for (int i = 0; i < ncomp; i++)
            comp[i].dc_pred = 0;
        int todo = restart_interval ? restart_interval : 0x7fffffff;
        if (nscan == 1) { // non-interleaved: the component's own blocks, row by row
            Component &c = comp[order[0]];
            const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
            for (int by = 0; by < bh; by++)
                for (int bx = 0; bx < bw; bx++) {
                    if (progressive)
                        decode_block_progressive(c, bx, by);
                    else
                        decode_block(c, bx, by);
                    if (--todo <= 0) {
                        restart();
                        todo = restart_interval;
                    }
                }
            return;
        }
        const int mcu_w = 8 * hmax, mcu_h = 8 * vmax;
        const int mx = (width + mcu_w - 1) / mcu_w, my = (height + mcu_h - 1) / mcu_h;
        for (int j = 0; j < my; j++)
            for (int i = 0; i < mx; i++) {
                for (int k = 0; k < nscan; k++) {
                    Component &c = comp[order[k]];
                    for (int y = 0; y < c.v; y++)
                        for (int x = 0; x < c.h; x++) {
                            if (progressive)
                                decode_block_progressive(c, i * c.h + x, j * c.v + y);
                            else
                                decode_block(c, i * c.h + x, j * c.v + y);
                        }
                }
                if (--todo <= 0) {
                    restart();
                    todo = restart_interval;
                }
            }
    }

    void parse() {
        if (buf.size() < 4 || buf[0] != 0xFF || buf[1] != 0xD8)
            fail("not a JPEG file");
        pos = 2;
        bool have_frame = false, decoded = false;
        while (true) {
            // next marker
            int m = byte();
            while (m != 0xFF)
                m = byte();
            while (m == 0xFF)
                m = byte();
            if (m == 0xD9)
                break;
            if (m == 0x01 || (m >= 0xD0 && m <= 0xD7))
                continue;
            const size_t seg = pos;
            const int len = be16();
            if (len < 2 || seg + size_t(len) > buf.size())
                fail("bad segment length");
            if (m == 0xDB) { // DQT
                while (pos < seg + size_t(len)) {
                    const int pq = byte(), prec = pq >> 4, id = pq & 15;
                    if (id > 3 || prec > 1)
                        fail("bad quantisation table");
                    for (int k = 0; k < 64; k++)
                        quant[id][k] = uint16_t(prec ? be16() : byte());
                    quant_present[id] = true;
                }
            } else if (m == 0xC4) { // DHT
                while (pos < seg + size_t(len)) {
                    const int tc = byte(), cls = tc >> 4, id = tc & 15;
                    if (cls > 1 || id > 3)
                        fail("bad Huffman table");
                    uint8_t counts[16];
                    int total = 0;
                    for (int k = 0; k < 16; k++)
                        counts[k] = byte(), total += counts[k];
                    if (total > 256 || pos + size_t(total) > buf.size())
                        fail("bad Huffman table");
                    build(cls ? ac[id] : dc[id], counts, &buf[pos], total);
                    pos += size_t(total);
                }
            } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) { // SOF0 / SOF1 / SOF2 (progressive)
                progressive = m == 0xC2;
                if (have_frame)
                    fail("more than one frame");
                if (byte() != 8)
                    fail("only 8-bit samples are supported");
                height = be16(), width = be16(), ncomp = byte();
                if (width <= 0 || height <= 0)
                    fail("empty image");
                if (ncomp != 1 && ncomp != 3)
                    fail("only greyscale and 3-component files are supported");
                int rgb = 0;
                for (int i = 0; i < ncomp; i++) {
                    Component &c = comp[i];
                    c.id = byte();
                    rgb += c.id == "RGB"[i];
                    const int hv = byte();
                    c.h = hv >> 4, c.v = hv & 15, c.tq = byte();
                    if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3)
                        fail("bad component");
                    hmax = std::max(hmax, c.h), vmax = std::max(vmax, c.v);
                }
                is_rgb = ncomp == 3 && rgb == 3;
                const int mx = (width + 8 * hmax - 1) / (8 * hmax), my = (height + 8 * vmax - 1) / (8 * vmax);
                for (int i = 0; i < ncomp; i++) {
                    Component &c = comp[i];
                    if (hmax % c.h || vmax % c.v)
                        fail("fractional sampling ratios are not supported");
                    c.x = (width * c.h + hmax - 1) / hmax, c.y = (height * c.v + vmax - 1) / vmax;
                    c.w2 = mx * c.h * 8, c.h2 = my * c.v * 8;
                    c.data.assign(size_t(c.w2) * c.h2, 0);
                    if (progressive)
                        c.coef.assign(size_t(c.w2) * c.h2, 0);
                }
                have_frame = true;
            } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
                fail("unsupported JPEG process (lossless, hierarchical or arithmetic coding)");
            } else if (m == 0xDD) {
                restart_interval = be16();
            } else if (m == 0xE0) {
                saw_jfif = len >= 7 && memcmp(&buf[pos], "JFIF", 5) == 0;
            } else if (m == 0xEE) {
                if (len >= 14 && memcmp(&buf[pos], "Adobe", 5) == 0)
                    adobe_transform = buf[pos + 11];
            } else if (m == 0xDA) { // SOS
                if (!have_frame)
                    fail("scan before frame header");
                const int ns = byte();
                if (ns < 1 || ns > ncomp)
                    fail("bad scan header");
                int order[3] = {0, 0, 0};
                for (int k = 0; k < ns; k++) {
                    const int id = byte(), tt = byte();
                    int which = -1;
                    for (int i = 0; i < ncomp; i++)
                        if (comp[i].id == id)
                            which = i;
                    if (which < 0)
                        fail("scan refers to an unknown component");
                    comp[which].td = tt >> 4, comp[which].ta = tt & 15;
                    if (comp[which].td > 3 || comp[which].ta > 3 || !quant_present[comp[which].tq])
                        fail("scan refers to an undefined table");
                    order[k] = which;
                }
                ss = byte(), se = byte();
                const int ah_al = byte();
                ah = ah_al >> 4, al = ah_al & 15;
                if (progressive) {
                    if (ss > 63 || se > 63 || ss > se || ah > 13 || al > 13 || (ss == 0 && se != 0) || (ss != 0 && ns != 1))
                        fail("bad progressive scan parameters");
                } else if (ss != 0 || se != 63 || ah_al != 0) {
                    fail("bad scan parameters");
                }
                pos = seg + size_t(len);
                scan(ns, order);
                decoded = true;
                // continue after the entropy-coded data: pos sits at the marker code (or inside the data if none was seen)
                if (hit_marker)
                    pos = pos > 0 ? pos - 1 : 0; // step back onto the 0xFF so the marker loop finds it
                continue;
            }
            pos = seg + size_t(len);
        }
        if (!have_frame || !decoded)
            fail("no image data");
        if (progressive)
            finish_progressive();
    }
};

inline int div4(int x) { return x >> 2; }
inline int div16(int x) { return x >> 4; }

// one output row of a component at full resolution
void resample_row(std::vector<uint8_t> &out, const uint8_t *near, const uint8_t *far, int w, int hs, int vs) {
    if (hs == 1 && vs == 1) {
        memcpy(out.data(), near, size_t(w));
    } else if (hs == 1 && vs == 2) {
        for (int i = 0; i < w; i++)
            out[i] = uint8_t(div4(3 * near[i] + far[i] + 2));
    } else if (hs == 2 && vs == 1) {
        const uint8_t *in = near;
        if (w == 1) {
            out[0] = out[1] = in[0];
            return;
        }
        out[0] = in[0];
        out[1] = uint8_t(div4(in[0] * 3 + in[1] + 2));
        int i;
        for (i = 1; i < w - 1; i++) {
            const int n = 3 * in[i] + 2;
            out[i * 2 + 0] = uint8_t(div4(n + in[i - 1]));
            out[i * 2 + 1] = uint8_t(div4(n + in[i + 1]));
        }
        out[i * 2 + 0] = uint8_t(div4(in[w - 2] * 3 + in[w - 1] + 2));
        out[i * 2 + 1] = in[w - 1];
    } else if (hs == 2 && vs == 2) {
        if (w == 1) {
            out[0] = out[1] = uint8_t(div4(3 * near[0] + far[0] + 2));
            return;
        }
        int t1 = 3 * near[0] + far[0];
        out[0] = uint8_t(div4(t1 + 2));
        for (int i = 1; i < w; i++) {
            const int t0 = t1;
            t1 = 3 * near[i] + far[i];
            out[i * 2 - 1] = uint8_t(div16(3 * t0 + t1 + 8));
            out[i * 2] = uint8_t(div16(3 * t1 + t0 + 8));
        }
        out[w * 2 - 1] = uint8_t(div4(t1 + 2));
    } else { // other ratios: nearest sample
        for (int i = 0; i < w; i++)
            for (int j = 0; j < hs; j++)
                out[i * hs + j] = near[i];
    }
}

inline int f2fixed(float x) { return int(x * 4096.0f + 0.5f) << 8; }

} // namespace

std::vector<uint8_t> decode_jpeg(const std::vector<uint8_t> &file, const std::string &path, int want, int &width, int &height) {
    Decoder d{file, path};
    d.parse();
    if (d.ncomp == 3 && d.adobe_transform == 0 && !d.saw_jfif)
        d.is_rgb = true; // Adobe marker: components are R, G, B, no colour transform
    width = d.width, height = d.height;
    const int n = want ? want : (d.ncomp >= 3 ? 3 : 1);
    const bool colour_out = d.ncomp == 3 && n >= 3;
    const int planes = (d.ncomp == 3 && n < 3 && !d.is_rgb) ? 1 : d.ncomp; // 1 / 2 channels from a YCbCr file: the luma plane
    std::vector<uint8_t> out(size_t(width) * height * n);
    struct Resample {
        const uint8_t *line0, *line1;
        int hs, vs, w_lores, ystep, ypos;
        std::vector<uint8_t> linebuf;
    } rs[3];
    for (int k = 0; k < planes; k++) {
        const Component &c = d.comp[k];
        rs[k].hs = d.hmax / c.h, rs[k].vs = d.vmax / c.v;
        rs[k].ystep = rs[k].vs >> 1;
        rs[k].w_lores = (width + rs[k].hs - 1) / rs[k].hs;
        rs[k].ypos = 0;
        rs[k].line0 = rs[k].line1 = c.data.data();
        rs[k].linebuf.assign(size_t(width) + 3 + size_t(rs[k].hs) * 2, 0);
    }
    const uint8_t *row[3] = {nullptr, nullptr, nullptr};
    for (int j = 0; j < height; j++) {
        for (int k = 0; k < planes; k++) {
            Resample &r = rs[k];
            const bool y_bot = r.ystep >= (r.vs >> 1);
            if (r.hs == 1 && r.vs == 1) {
                row[k] = y_bot ? r.line1 : r.line0;
            } else {
                resample_row(r.linebuf, y_bot ? r.line1 : r.line0, y_bot ? r.line0 : r.line1, r.w_lores, r.hs, r.vs);
                row[k] = r.linebuf.data();
            }
            if (++r.ystep >= r.vs) {
                r.ystep = 0;
                r.line0 = r.line1;
                if (++r.ypos < d.comp[k].y)
                    r.line1 += d.comp[k].w2;
            }
        }
        uint8_t *o = out.data() + size_t(j) * width * n;
        if (colour_out) {
            for (int i = 0; i < width; i++, o += n) {
                if (d.is_rgb) {
                    o[0] = row[0][i], o[1] = row[1][i], o[2] = row[2][i];
                } else {
                    const int y_fixed = (int(row[0][i]) << 20) + (1 << 19);
                    const int cb = int(row[1][i]) - 128, cr = int(row[2][i]) - 128;
                    int r = y_fixed + cr * f2fixed(1.40200f);
                    int g = y_fixed + cr * -f2fixed(0.71414f) + int(uint32_t(cb * -f2fixed(0.34414f)) & 0xffff0000u);
                    int b = y_fixed + cb * f2fixed(1.77200f);
                    r >>= 20, g >>= 20, b >>= 20;
                    o[0] = uint8_t(r < 0 ? 0 : (r > 255 ? 255 : r));
                    o[1] = uint8_t(g < 0 ? 0 : (g > 255 ? 255 : g));
                    o[2] = uint8_t(b < 0 ? 0 : (b > 255 ? 255 : b));
                }
                if (n == 4)
                    o[3] = 255;
            }
        } else if (d.ncomp == 3 && d.is_rgb) { // grey from an RGB-coded file
            for (int i = 0; i < width; i++, o += n) {
                o[0] = uint8_t((row[0][i] * 77 + row[1][i] * 150 + 29 * row[2][i]) >> 8);
                if (n == 2)
                    o[1] = 255;
            }
        } else {
            for (int i = 0; i < width; i++, o += n) {
                const uint8_t y = row[0][i];
                o[0] = y;
                if (n >= 3)
                    o[1] = o[2] = y;
                if (n == 2)
                    o[1] = 255;
                if (n == 4)
                    o[3] = 255;
            }
        }
    }
    return out;
}

} // namespace cptb
