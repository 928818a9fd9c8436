// nvdb_read.cpp — see nvdb_read.h.
//
// File layout (NanoVDB 32.x, little endian), as written by nanovdb::io::writeGrid:
//   file header   16 B   magic "NanoVDB0" / "NanoVDB2", version (major << 21 | minor << 10 | patch), grid count, codec
//   per grid      176 B  metadata (sizes, type, class, bounding boxes, ..., name size, node counts, codec) + name
//   per grid      the grid buffer: GridData (672 B) | TreeData (64 B) | root | upper nodes | lower nodes | leaves
// Node layouts are 32 B aligned; V = sizeof(value) (4 for float, 12 for Vec3f), T = table entry (8 B, 16 B for Vec3f):
//   root   index bbox (24) | table size (4) | background, min, max (3V) | avg, stddev (8) -> padded to 32; then tiles of
//          32 B: key (8: x << 42 | y << 21 | z of origin >> 12) | child byte offset from the root (8, 0 = value tile) |
//          state (4) | value (V)
//   upper  bbox (24) | flags (8) | value mask (4096) | child mask (4096) | min, max (2V) | avg, stddev (8) -> 32 B
//          aligned | table[32^3] of T (child byte offset from this node, or tile value); child n covers 128^3 voxels at
//          ((n >> 10) & 31, (n >> 5) & 31, n & 31) * 128
//   lower  same with 512 B masks and table[16^3]; child n is a leaf at ((n >> 8) & 15, (n >> 4) & 15, n & 15) * 8
//   leaf   bbox min (12) | bbox dif (3) | flags (1) | value mask (64) | min, max (2V) | avg, stddev (8) -> 32 B aligned |
//          values[512] in x-major order ((i & 7) << 6 | (j & 7) << 3 | (k & 7))
#include "nvdb_read.h"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace cptb {
namespace {

constexpr uint64_t kMagicNumb = 0x304244566f6e614eULL, kMagicGrid = 0x314244566f6e614eULL, kMagicFile = 0x324244566f6e614eULL;

struct Reader {
    const std::vector<uint8_t> &buf;
    const std::string &path;
    [[noreturn]] void fail(const std::string &why) const { throw std::runtime_error("nvdb '" + path + "': " + why); }
    void need(uint64_t off, uint64_t bytes) const {
        if (off > buf.size() || bytes > buf.size() - off)
            fail("truncated file");
    }
    template <typename T> T get(uint64_t off) const {
        need(off, sizeof(T));
        T v;
        memcpy(&v, buf.data() + off, sizeof(T));
        return v;
    }
};

inline uint64_t align32(uint64_t x) { return (x + 31) & ~uint64_t(31); }
inline bool mask_on(const Reader &r, uint64_t mask_off, uint32_t n) { return (r.get<uint64_t>(mask_off + 8 * (n >> 6)) >> (n & 63)) & 1; }

struct Builder {
    const Reader &r;
    int ncomp;
    NvdbGrid &out;
    float bg[3];
    std::vector<std::vector<float>> pool; // bricks in discovery order; renumbered in table order at the end

    int64_t table_slot(int x, int y, int z) const { // brick slot of the brick whose origin is (x, y, z), or -1 outside
        const CptsGridChannel &c = out.channel;
        const int bx = (x - c.bmin[0]) >> 3, by = (y - c.bmin[1]) >> 3, bz = (z - c.bmin[2]) >> 3;
        if (x < c.bmin[0] || y < c.bmin[1] || z < c.bmin[2] || bx >= c.nb[0] || by >= c.nb[1] || bz >= c.nb[2])
            return -1;
        return (int64_t(bx) * c.nb[1] + by) * c.nb[2] + bz;
    }
    bool is_bg(const float *v) const {
        for (int k = 0; k < ncomp; k++)
            if (v[k] != bg[k])
                return false;
        return true;
    }
    void put_brick(int x, int y, int z, std::vector<float> &&values) {
        const int64_t slot = table_slot(x, y, z);
        if (slot < 0)
            return; // outside the index bounding box: such voxels read as background on both sides
        out.bricks[slot] = int32_t(pool.size());
        pool.push_back(std::move(values));
    }
    // constant tile: every brick of the region that overlaps the table
    void put_tile(int x, int y, int z, int dim, const float *value) {
        if (is_bg(value))
            return;
        const CptsGridChannel &c = out.channel;
        const int x1 = std::min(x + dim, c.bmin[0] + c.nb[0] * 8), y1 = std::min(y + dim, c.bmin[1] + c.nb[1] * 8),
                  z1 = std::min(z + dim, c.bmin[2] + c.nb[2] * 8);
        for (int i = std::max(x, c.bmin[0]); i < x1; i += 8)
            for (int j = std::max(y, c.bmin[1]); j < y1; j += 8)
                for (int k = std::max(z, c.bmin[2]); k < z1; k += 8) {
                    std::vector<float> v(size_t(512) * ncomp);
                    for (int n = 0; n < 512; n++)
                        memcpy(&v[size_t(n) * ncomp], value, sizeof(float) * ncomp);
                    put_brick(i, j, k, std::move(v));
                }
    }
    void leaf(uint64_t off, int x, int y, int z) {
        const uint64_t V = 4u * ncomp, values = off + align32(16 + 64 + 2 * V + 8);
        r.need(values, 512 * V);
        std::vector<float> v(size_t(512) * ncomp);
        memcpy(v.data(), r.buf.data() + values, 512 * V);
        put_brick(x, y, z, std::move(v));
    }
    // internal node with 2^(3*log2dim) children, each covering child_dim^3 voxels
    void internal(uint64_t off, int x, int y, int z, int log2dim, int child_dim) {
        const uint64_t V = 4u * ncomp, T = ncomp == 1 ? 8 : 16, mask_bytes = (uint64_t(1) << (3 * log2dim)) / 8;
        const uint64_t child_mask = off + 32 + mask_bytes, table = off + align32(32 + 2 * mask_bytes + 2 * V + 8);
        const uint32_t n_children = 1u << (3 * log2dim), m = (1u << log2dim) - 1;
        r.need(table, uint64_t(n_children) * T);
        for (uint32_t n = 0; n < n_children; n++) {
            const int cx = x + int((n >> (2 * log2dim)) & m) * child_dim, cy = y + int((n >> log2dim) & m) * child_dim,
                      cz = z + int(n & m) * child_dim;
            const uint64_t entry = table + uint64_t(n) * T;
            if (mask_on(r, child_mask, n)) {
                const uint64_t child = off + uint64_t(r.get<int64_t>(entry));
                if (child_dim == 8)
                    leaf(child, cx, cy, cz);
                else
                    internal(child, cx, cy, cz, 4, 8);
            } else {
                float value[3] = {0, 0, 0};
                r.need(entry, V);
                memcpy(value, r.buf.data() + entry, V);
                put_tile(cx, cy, cz, child_dim, value);
            }
        }
    }
};

} // namespace

void read_nvdb(const std::string &path, int ncomp, NvdbGrid &out) {
    std::vector<uint8_t> buf;
    {
        FILE *fp = fopen(path.c_str(), "rb");
        if (!fp)
            throw std::runtime_error("nvdb: cannot open '" + path + "'");
        fseek(fp, 0, SEEK_END);
        const long size = ftell(fp);
        fseek(fp, 0, SEEK_SET);
        buf.resize(size > 0 ? size_t(size) : 0);
        const bool ok = buf.empty() || fread(buf.data(), 1, buf.size(), fp) == buf.size();
        fclose(fp);
        if (!ok)
            throw std::runtime_error("nvdb: cannot read '" + path + "'");
    }
    const Reader r{buf, path};
    if (ncomp != 1 && ncomp != 3)
        r.fail("only float and Vec3f grids are read");
    // ---- file header + first grid's metadata
    const uint64_t magic = r.get<uint64_t>(0);
    if (magic != kMagicNumb && magic != kMagicFile)
        r.fail(magic == kMagicGrid ? "raw grid buffers are not supported, expected a .nvdb file" : "not a NanoVDB file");
    const uint32_t version = r.get<uint32_t>(8);
    if ((version >> 21) != 32)
        r.fail("unsupported format version " + std::to_string(version >> 21) + " (32 is)");
    if (r.get<uint16_t>(12) < 1)
        r.fail("file holds no grid");
    const uint64_t meta = 16;
    const uint64_t grid_size = r.get<uint64_t>(meta), file_size = r.get<uint64_t>(meta + 8);
    const uint32_t grid_type = r.get<uint32_t>(meta + 32), name_size = r.get<uint32_t>(meta + 136);
    const uint16_t codec = r.get<uint16_t>(meta + 168);
    if (codec != 0 || file_size != grid_size)
        r.fail("compressed grids (ZIP / BLOSC) are not supported; re-save with Codec::NONE");
    if (grid_type != (ncomp == 1 ? 1u : 6u))
        r.fail(std::string("expected a ") + (ncomp == 1 ? "float" : "Vec3f") + " grid, found grid type " + std::to_string(grid_type));
    // metadata of all grids precede the first grid buffer
    uint64_t g = 16;
    for (uint16_t i = 0, n = r.get<uint16_t>(12); i < n; i++)
        g += 176 + r.get<uint32_t>(g + 136);
    (void)name_size;
    r.need(g, grid_size);
    if (grid_size < 672 + 64)
        r.fail("grid buffer too small");
    const uint64_t gm = r.get<uint64_t>(g);
    if (gm != kMagicNumb && gm != kMagicGrid)
        r.fail("grid buffer without a NanoVDB magic number");
    // ---- GridData: affine map (float copies) and world bounding box
    CptsGridChannel &c = out.channel;
    memset(&c, 0, sizeof(c));
    const uint64_t map = g + 296;
    for (int k = 0; k < 9; k++)
        c.inv_mat[k] = r.get<float>(map + 36 + 4 * k);
    for (int k = 0; k < 3; k++)
        c.vec[k] = r.get<float>(map + 72 + 4 * k);
    for (int k = 0; k < 6; k++)
        out.world_bbox[k] = r.get<double>(g + 560 + 8 * k);
    // ---- tree -> root
    const uint64_t tree = g + 672, V = 4u * ncomp;
    const uint64_t root = tree + uint64_t(r.get<int64_t>(tree + 24));
    int32_t bbox[6];
    for (int k = 0; k < 6; k++)
        bbox[k] = r.get<int32_t>(root + 4 * k);
    const uint32_t table_size = r.get<uint32_t>(root + 24);
    Builder b{r, ncomp, out, {0, 0, 0}, {}};
    for (int k = 0; k < ncomp; k++)
        b.bg[k] = c.background[k] = r.get<float>(root + 28 + 4 * k);
    out.value_min = r.get<float>(root + 28 + V);
    out.value_max = r.get<float>(root + 28 + 2 * V);
    c.ncomp = ncomp;
    const bool empty = bbox[3] < bbox[0] || bbox[4] < bbox[1] || bbox[5] < bbox[2];
    for (int k = 0; k < 3; k++) {
        c.bmin[k] = empty ? 0 : (bbox[k] & ~7);
        c.nb[k] = empty ? 1 : ((bbox[3 + k] - c.bmin[k]) >> 3) + 1;
        if (c.nb[k] > 4096)
            r.fail("index bounding box larger than 32768 voxels per axis");
    }
    out.bricks.assign(size_t(c.nb[0]) * c.nb[1] * c.nb[2], -1);
    const uint64_t tiles = root + align32(24 + 4 + 3 * V + 8);
    for (uint32_t t = 0; t < table_size && !empty; t++) {
        const uint64_t tile = tiles + 32ull * t;
        const uint64_t key = r.get<uint64_t>(tile);
        const int64_t child = r.get<int64_t>(tile + 8);
        const int x = int32_t(uint32_t((key >> 42) & 0x1fffff) << 12), y = int32_t(uint32_t((key >> 21) & 0x1fffff) << 12),
                  z = int32_t(uint32_t(key & 0x1fffff) << 12);
        if (child != 0) {
            b.internal(root + uint64_t(child), x, y, z, 5, 128);
        } else {
            float value[3] = {0, 0, 0};
            for (int k = 0; k < ncomp; k++)
                value[k] = r.get<float>(tile + 20 + 4 * k);
            b.put_tile(x, y, z, 4096, value);
        }
    }
    // ---- canonical numbering: bricks in table order; all-background bricks are dropped
    out.data.clear();
    int32_t next = 0;
    for (auto &slot : out.bricks) {
        if (slot < 0)
            continue;
        const std::vector<float> &v = b.pool[slot];
        bool any = false;
        for (int n = 0; n < 512 && !any; n++)
            any = !b.is_bg(&v[size_t(n) * ncomp]);
        if (!any) {
            slot = -1;
            continue;
        }
        out.data.insert(out.data.end(), v.begin(), v.end());
        slot = next++;
    }
}

} // namespace cptb
