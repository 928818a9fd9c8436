// Host-side check of the image decoders (image_io.cpp, jpeg_read.cpp): decode <image> to <channels> and compare byte for byte
// with the dump `oracle/_ref/ref_dump imgdump` made with the reference's own decoder.  Built and run by tests/test_host.py.
#include <cstdlib>
#include "image_io.h"
#include <cstdio>
#include <cstring>
int main(int argc, char **argv) {
    int w, h;
    try {
        std::vector<uint8_t> px = cptb::load_image8(argv[1], atoi(argv[2]), w, h);
        FILE *fp = fopen(argv[3], "rb"); int32_t hdr[4]; if (!fp || fread(hdr, 4, 4, fp) != 4) return 3;
        std::vector<uint8_t> ref(size_t(hdr[0]) * hdr[1] * hdr[3]); if (fread(ref.data(), 1, ref.size(), fp) != ref.size()) return 3; fclose(fp);
        size_t bad = 0, first = 0; int maxd = 0;
        if (w != hdr[0] || h != hdr[1] || px.size() != ref.size()) { printf("%s size mismatch %dx%d vs %dx%d\n", argv[1], w, h, hdr[0], hdr[1]); return 1; }
        for (size_t i = 0; i < px.size(); i++) if (px[i] != ref[i]) { if (!bad) first = i; bad++; int d = abs(int(px[i]) - int(ref[i])); if (d > maxd) maxd = d; }
        printf("%s req %s: %dx%d mismatching bytes %zu of %zu (first at %zu, max diff %d)\n", argv[1], argv[2], w, h, bad, px.size(), first, maxd);
        return bad != 0;
    } catch (const std::exception &e) { printf("%s: %s\n", argv[1], e.what()); return 2; }
}
