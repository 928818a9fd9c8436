// scene_xml.cpp — stand-alone loader for the reference's Mitsuba-like XML scene format
// (version 1.2) producing the same flat arrays the reference's `Scene` object hands to a
// renderer (behaviour followed: src/impl/scene.cu:58-1192, src/impl/camera_model.cu:51-153,
// src/impl/config.cpp:27-157, src/impl/object.cu:30-58, src/core/bvh.cuh:322-377 for the
// linear node format).  Own code throughout: a ~120-line XML subset parser, a Wavefront OBJ
// reader and a binned-SAH BVH builder.  The reference remains the loader of record in an
// integrated build; this one exists so the library, its tests and the
// benchmark are self-contained.
//
// Image textures and environment maps are decoded from PNG files (image_io.cpp).
// Not supported (clear errors, never silent): sphere shapes (reference default build is
// TRIANGLE_ONLY), JPEG / HDR images, grid (NanoVDB) media.
#include "image_io.h"
#include "scene_host.h"
#include "nvdb_read.h"

#include <dlfcn.h>

#include <algorithm>
#include <queue>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

namespace cptb {
namespace {

// ------------------------------------------------------------------------------------------------
// XML subset: elements, attributes, comments, <?...?> declarations, self-closing tags
// ------------------------------------------------------------------------------------------------
struct XNode {
    std::string tag;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<XNode> kids;
    const std::string *attr(const std::string &k) const {
        for (auto &a : attrs)
            if (a.first == k)
                return &a.second;
        return nullptr;
    }
    std::string attr_or(const std::string &k, const std::string &d) const {
        auto p = attr(k);
        return p ? *p : d;
    }
    std::vector<const XNode *> children(const std::string &t) const {
        std::vector<const XNode *> out;
        for (auto &c : kids)
            if (c.tag == t)
                out.push_back(&c);
        return out;
    }
    const XNode *first(const std::string &t) const {
        for (auto &c : kids)
            if (c.tag == t)
                return &c;
        return nullptr;
    }
};

struct XParser {
    const std::string &s;
    size_t i = 0;
    explicit XParser(const std::string &src) : s(src) {}
    [[noreturn]] void fail(const std::string &m) const {
        throw std::runtime_error("xml: " + m + " near offset " + std::to_string(i));
    }
    void skip_ws() {
        while (i < s.size() && isspace((unsigned char)s[i]))
            ++i;
    }
    bool starts(const char *t) const { return s.compare(i, strlen(t), t) == 0; }
    void skip_misc() { // whitespace, comments, declarations, text
        while (i < s.size()) {
            skip_ws();
            if (starts("<!--")) {
                size_t e = s.find("-->", i + 4);
                if (e == std::string::npos)
                    fail("unterminated comment");
                i = e + 3;
            } else if (starts("<?")) {
                size_t e = s.find("?>", i + 2);
                if (e == std::string::npos)
                    fail("unterminated declaration");
                i = e + 2;
            } else if (i < s.size() && s[i] != '<') {
                while (i < s.size() && s[i] != '<')
                    ++i; // character data is ignored
            } else
                break;
        }
    }
    std::string name() {
        size_t b = i;
        while (i < s.size() && (isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == '-' || s[i] == ':' || s[i] == '.'))
            ++i;
        if (b == i)
            fail("expected a name");
        return s.substr(b, i - b);
    }
    XNode element() {
        if (s[i] != '<')
            fail("expected '<'");
        ++i;
        XNode n;
        n.tag = name();
        while (true) {
            skip_ws();
            if (i >= s.size())
                fail("unexpected end inside tag");
            if (s[i] == '/') {
                if (!starts("/>"))
                    fail("malformed tag end");
                i += 2;
                return n;
            }
            if (s[i] == '>') {
                ++i;
                break;
            }
            std::string k = name();
            skip_ws();
            if (s[i] != '=')
                fail("expected '='");
            ++i;
            skip_ws();
            const char q = s[i];
            if (q != '"' && q != '\'')
                fail("expected a quoted value");
            size_t e = s.find(q, i + 1);
            if (e == std::string::npos)
                fail("unterminated attribute");
            n.attrs.emplace_back(k, s.substr(i + 1, e - i - 1));
            i = e + 1;
        }
        while (true) {
            skip_misc();
            if (i >= s.size())
                fail("unterminated element <" + n.tag + ">");
            if (starts("</")) {
                i += 2;
                std::string t = name();
                if (t != n.tag)
                    fail("mismatched closing tag </" + t + "> for <" + n.tag + ">");
                skip_ws();
                if (s[i] != '>')
                    fail("malformed closing tag");
                ++i;
                return n;
            }
            n.kids.push_back(element());
        }
    }
};

XNode parse_xml_file(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f)
        throw std::runtime_error("cannot open scene file '" + path + "'");
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string src = ss.str();
    XParser p(src);
    p.skip_misc();
    if (p.i >= src.size())
        throw std::runtime_error("xml: empty document '" + path + "'");
    return p.element();
}

// ------------------------------------------------------------------------------------------------
// value parsing (scene.cu:58-122)
// ------------------------------------------------------------------------------------------------
std::vector<float> parse_floats(const std::string &v) {
    std::string t = v;
    for (auto &c : t)
        if (c == ',')
            c = ' ';
    std::stringstream ss(t);
    std::vector<float> out;
    float f;
    while (ss >> f)
        out.push_back(f);
    return out;
}

void parse_color(const std::string &v, float out[4]) {
    out[3] = 1.f;
    if (!v.empty() && v[0] == '#') {
        unsigned c = 0;
        sscanf(v.c_str() + 1, "%x", &c);
        out[0] = float((c >> 16) & 0xFF) / 255.0f;
        out[1] = float((c >> 8) & 0xFF) / 255.0f;
        out[2] = float(c & 0xFF) / 255.0f;
        return;
    }
    auto vals = parse_floats(v);
    if (vals.size() >= 3)
        out[0] = vals[0], out[1] = vals[1], out[2] = vals[2];
    else if (vals.size() == 1)
        out[0] = out[1] = out[2] = vals[0];
    else
        throw std::runtime_error("scene: cannot parse colour '" + v + "'");
}

F3 parse_vec3(const std::string &v) {
    auto vals = parse_floats(v);
    if (vals.size() != 3)
        return F3{0, 0, 0};
    return F3{vals[0], vals[1], vals[2]};
}

const XNode *named_child(const XNode &parent, const std::string &tag, std::initializer_list<const char *> names) {
    for (auto c : parent.children(tag)) {
        auto n = c->attr("name");
        if (!n)
            continue;
        for (auto want : names)
            if (*n == want)
                return c;
    }
    return nullptr;
}
bool get_float(const XNode &parent, std::initializer_list<const char *> names, float &out) {
    if (auto c = named_child(parent, "float", names)) {
        out = float(atof(c->attr_or("value", "0").c_str()));
        return true;
    }
    return false;
}
bool get_int(const XNode &parent, std::initializer_list<const char *> names, int &out) {
    if (auto c = named_child(parent, "integer", names)) {
        out = atoi(c->attr_or("value", "0").c_str());
        return true;
    }
    return false;
}
bool get_bool(const XNode &parent, std::initializer_list<const char *> names, bool &out) {
    if (auto c = named_child(parent, "bool", names)) {
        const std::string v = c->attr_or("value", "false");
        out = (v == "true" || v == "1" || v == "True");
        return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// tables (reference: src/core/preset_params.cuh:20-83 — physical constants, Tungsten / LuxRender data)
// ------------------------------------------------------------------------------------------------
struct Metal {
    const char *name;
    float eta[3], k[3];
};
const Metal kMetals[] = {
    {"Au", {0.1431189557f, 0.3749570432f, 1.4424785571f}, {3.9831604247f, 2.3857207478f, 1.6032152899f}},
    {"Cr", {4.3696828663f, 2.9167024892f, 1.6547005413f}, {5.2064337956f, 4.2313645277f, 3.7549467933f}},
    {"Cu", {0.2004376970f, 0.9240334304f, 1.1022119527f}, {3.9129485033f, 2.4528477015f, 2.1421879552f}},
    {"Ag", {0.1552646489f, 0.1167232965f, 0.1383806959f}, {4.8283433224f, 3.1222459278f, 2.1469504455f}},
    {"Al", {1.6574599595f, 0.8803689579f, 0.5212287346f}, {9.2238691996f, 6.2695232477f, 4.8370012281f}},
    {"W", {4.3707029924f, 3.3002972445f, 2.9982666528f}, {3.5006778591f, 2.6048652781f, 2.2731930614f}},
    {"TiO2", {3.4566203131f, 2.8017076558f, 2.9051485020f}, {0.0001026662f, -0.0000897534f, 0.0006356902f}},
    {"Ni", {2.3672753521f, 1.6633583302f, 1.4670554172f}, {4.4988329911f, 3.0501643957f, 2.3454274399f}},
    {"MgO", {2.0895885542f, 1.6507224525f, 1.5948759692f}, {0.f, 0.f, 0.f}},
    {"Na", {0.0602665320f, 0.0561412435f, 0.0619909494f}, {3.1792906496f, 2.1124800781f, 1.5790940266f}},
    {"SiC", {3.1723450205f, 2.5259677964f, 2.4793623897f}, {0.0000007284f, -0.0000006859f, 0.0000100150f}},
    {"V", {4.2775126218f, 3.5131538236f, 2.7611257461f}, {3.4911844504f, 2.8893580874f, 3.1116965117f}},
    {"CuO", {3.2453822204f, 2.4496293965f, 2.1974114493f}, {0.5202739621f, 0.5707372756f, 0.7172250613f}},
    {"Hg", {2.3989314904f, 1.4400254917f, 0.9095512090f}, {6.3276269444f, 4.3719414152f, 3.4217899270f}},
    {"Ir", {3.0864098394f, 2.0821938440f, 1.6178866805f}, {5.5921510077f, 4.0671757150f, 3.2672611269f}},
};
struct Dielectric {
    const char *name;
    float a, b;
};
const Dielectric kDielectrics[] = {{"Diamond", 2.3840f, 12100.0f},    {"DiamondHigh", 2.3840f, 60500.0f},
                                   {"Silica", 1.4580f, 3540.0f},      {"Glass_BK7", 1.5046f, 4200.0f},
                                   {"Glass_BaF10", 1.6700f, 7430.0f}, {"Glass_SF10", 1.7280f, 13420.0f},
                                   {"Sapphire", 1.7522f, 5500.0f},    {"Water", 1.3244f, 3100.0f}};

constexpr float kPiF = 3.1415926535897f;
constexpr float kDeg2Rad = kPiF / 180.f;
constexpr float kAabbEps = 1e-5f;

// ------------------------------------------------------------------------------------------------
// OBJ reader (triangles; polygons are fan-triangulated)
// ------------------------------------------------------------------------------------------------
struct ObjMesh {
    std::array<std::vector<F3>, 3> v, n;
    std::array<std::vector<F2>, 3> uv;
    std::vector<bool> has_normal;
};

void load_obj(const std::string &path, ObjMesh &mesh) {
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp)
        throw std::runtime_error("scene: cannot open mesh '" + path + "'");
    std::vector<F3> vs, ns;
    std::vector<F2> ts;
    std::vector<char> line(1 << 16);
    struct Idx {
        int v, t, n;
    };
    auto fix = [](int idx, size_t count) { return idx > 0 ? idx - 1 : (idx < 0 ? int(count) + idx : -1); };
    while (fgets(line.data(), int(line.size()), fp)) {
        const char *p = line.data();
        while (*p == ' ' || *p == '\t')
            ++p;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            char *e;
            F3 v;
            v.x = float(strtod(p + 2, &e));
            v.y = float(strtod(e, &e));
            v.z = float(strtod(e, &e));
            vs.push_back(v);
        } else if (p[0] == 'v' && p[1] == 'n') {
            char *e;
            F3 v;
            v.x = float(strtod(p + 3, &e));
            v.y = float(strtod(e, &e));
            v.z = float(strtod(e, &e));
            ns.push_back(v);
        } else if (p[0] == 'v' && p[1] == 't') {
            char *e;
            F2 v;
            v.x = float(strtod(p + 3, &e));
            v.y = float(strtod(e, &e));
            ts.push_back(v);
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            std::vector<Idx> corner;
            const char *q = p + 2;
            while (*q) {
                while (*q == ' ' || *q == '\t')
                    ++q;
                if (*q == 0 || *q == '\n' || *q == '\r')
                    break;
                Idx id{0, 0, 0};
                char *e;
                id.v = int(strtol(q, &e, 10));
                q = e;
                if (*q == '/') {
                    ++q;
                    if (*q != '/') {
                        id.t = int(strtol(q, &e, 10));
                        q = e;
                    }
                    if (*q == '/') {
                        ++q;
                        id.n = int(strtol(q, &e, 10));
                        q = e;
                    }
                }
                corner.push_back(id);
            }
            for (size_t k = 1; k + 1 < corner.size(); k++) {
                const Idx tri[3] = {corner[0], corner[k], corner[k + 1]};
                bool hn = false;
                for (int j = 0; j < 3; j++) {
                    const int vi = fix(tri[j].v, vs.size()), ti = fix(tri[j].t, ts.size()), ni = fix(tri[j].n, ns.size());
                    if (vi < 0 || size_t(vi) >= vs.size()) {
                        fclose(fp);
                        throw std::runtime_error("scene: bad vertex index in '" + path + "'");
                    }
                    mesh.v[j].push_back(vs[vi]);
                    if (ni >= 0 && size_t(ni) < ns.size()) {
                        hn = true;
                        mesh.n[j].push_back(ns[ni]);
                    } else
                        mesh.n[j].push_back(F3{0, 0, 0});
                    mesh.uv[j].push_back(ti >= 0 && size_t(ti) < ts.size() ? ts[ti] : F2{0, 0});
                }
                mesh.has_normal.push_back(hn);
            }
        }
    }
    fclose(fp);
    // face normals where the file has none (scene.cu:644-656 computes them; we always use this triangle's own vertices)
    for (size_t i = 0; i < mesh.has_normal.size(); i++) {
        if (mesh.has_normal[i])
            continue;
        const F3 a = mesh.v[0][i], b = mesh.v[1][i], c = mesh.v[2][i];
        const F3 e1{b.x - a.x, b.y - a.y, b.z - a.z}, e2{c.x - a.x, c.y - a.y, c.z - a.z};
        F3 n{e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x};
        const float inv = 1.f / std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
        n = F3{n.x * inv, n.y * inv, n.z * inv};
        for (int j = 0; j < 3; j++)
            mesh.n[j][i] = n;
    }
}

// ------------------------------------------------------------------------------------------------
// binned-SAH BVH over padded triangle boxes, emitted in the reference's linear format
// ------------------------------------------------------------------------------------------------
struct Box {
    float lo[3], hi[3];
    void clear() {
        lo[0] = lo[1] = lo[2] = 1e30f;
        hi[0] = hi[1] = hi[2] = -1e30f;
    }
    void grow(const Box &b) {
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], b.lo[a]);
            hi[a] = std::max(hi[a], b.hi[a]);
        }
    }
    float area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0)
            return 0.f;
        return 2.f * (dx * dy + dy * dz + dx * dz);
    }
};
struct PrimRef {
    Box box;
    float ctr[3];
    int prim;
};
struct BuildNode {
    Box box;
    int left = -1, right = -1; // children (indices into the build array) or -1
    int base = 0, count = 0;
    int depth = 0;
};

struct Builder {
    std::vector<PrimRef> refs;
    std::vector<BuildNode> nodes;
    int max_leaf = 16;
    int max_depth = 0;
    static constexpr int kMaxBins = 64;
    int kBins = 16;
    int sweep_below = 0;    // exact (sorted sweep) SAH for nodes with at most this many references (0 = binned only)
    const int *prim_obj = nullptr; // object of each primitive: near the root, "these objects | those objects" is a split candidate
    int object_split_depth = 0;
    float object_bias = 1.f;
    float kTraverse = 0.2f; // SAH cost of one inner-node visit relative to one triangle test

    int build(int base, int count, int depth) {
        const int id = int(nodes.size());
        nodes.emplace_back();
        {
            BuildNode &n = nodes[id];
            n.base = base, n.count = count, n.depth = depth;
            n.box.clear();
            for (int i = base; i < base + count; i++)
                n.box.grow(refs[i].box);
        }
        max_depth = std::max(max_depth, depth);
        if (count <= 2)
            return id;
        Box cb;
        cb.clear();
        for (int i = base; i < base + count; i++)
            for (int a = 0; a < 3; a++) {
                cb.lo[a] = std::min(cb.lo[a], refs[i].ctr[a]);
                cb.hi[a] = std::max(cb.hi[a], refs[i].ctr[a]);
            }
        const float parent_area = std::max(nodes[id].box.area(), 1e-20f);
        float best_cost = 1e30f;
        int best_axis = -1, best_bin = -1;
        for (int axis = 0; axis < 3; axis++) {
            const float ext = cb.hi[axis] - cb.lo[axis];
            if (!(ext > 1e-7f))
                continue;
            Box bins[kMaxBins];
            int cnt[kMaxBins] = {0};
            for (auto &b : bins)
                b.clear();
            const float scale = float(kBins) / ext;
            for (int i = base; i < base + count; i++) {
                int b = std::min(kBins - 1, std::max(0, int((refs[i].ctr[axis] - cb.lo[axis]) * scale)));
                bins[b].grow(refs[i].box);
                cnt[b]++;
            }
            float right_area[kMaxBins];
            int right_cnt[kMaxBins];
            Box acc;
            acc.clear();
            int c = 0;
            for (int b = kBins - 1; b > 0; b--) {
                acc.grow(bins[b]);
                c += cnt[b];
                right_area[b] = acc.area();
                right_cnt[b] = c;
            }
            acc.clear();
            c = 0;
            for (int b = 0; b < kBins - 1; b++) {
                acc.grow(bins[b]);
                c += cnt[b];
                if (c == 0 || right_cnt[b + 1] == 0)
                    continue;
                const float cost = kTraverse + (acc.area() * float(c) + right_area[b + 1] * float(right_cnt[b + 1])) / parent_area;
                if (cost < best_cost)
                    best_cost = cost, best_axis = axis, best_bin = b;
            }
        }
        int mid = -1;
        if (count <= sweep_below) { // exact SAH: sort by centroid on each axis, evaluate every split
            float sweep_cost = 1e30f;
            int sweep_axis = -1, sweep_pos = -1;
            std::vector<float> ra(size_t(count) + 1);
            for (int axis = 0; axis < 3; axis++) {
                std::sort(refs.begin() + base, refs.begin() + base + count,
                          [axis](const PrimRef &a, const PrimRef &b) { return a.ctr[axis] < b.ctr[axis]; });
                Box acc;
                acc.clear();
                for (int i = count - 1; i > 0; i--) {
                    acc.grow(refs[base + i].box);
                    ra[i] = acc.area();
                }
                acc.clear();
                for (int i = 1; i < count; i++) {
                    acc.grow(refs[base + i - 1].box);
                    const float cost = kTraverse + (acc.area() * float(i) + ra[i] * float(count - i)) / parent_area;
                    if (cost < sweep_cost)
                        sweep_cost = cost, sweep_axis = axis, sweep_pos = i;
                }
            }
            if (sweep_axis >= 0 && (sweep_cost < float(count) || count > max_leaf)) {
                if (sweep_axis != 2)
                    std::sort(refs.begin() + base, refs.begin() + base + count,
                              [sweep_axis](const PrimRef &a, const PrimRef &b) { return a.ctr[sweep_axis] < b.ctr[sweep_axis]; });
                mid = base + sweep_pos;
            } else if (count <= max_leaf) {
                return id;
            }
        }
        if (mid < 0 && prim_obj && depth < object_split_depth && count > max_leaf) {
            // object candidate: references in input order (objects are contiguous there), split at an object boundary.
            // The greedy SAH estimate of the two sides is compared with the best spatial candidate.
            std::sort(refs.begin() + base, refs.begin() + base + count, [](const PrimRef &a, const PrimRef &b) { return a.prim < b.prim; });
            std::vector<float> ra(size_t(count) + 1);
            Box acc;
            acc.clear();
            for (int i = count - 1; i > 0; i--) {
                acc.grow(refs[base + i].box);
                ra[i] = acc.area();
            }
            acc.clear();
            float obj_cost = 1e30f;
            int obj_pos = -1;
            for (int i = 1; i < count; i++) {
                acc.grow(refs[base + i - 1].box);
                if (prim_obj[refs[base + i].prim] == prim_obj[refs[base + i - 1].prim])
                    continue;
                const float cost = kTraverse + (acc.area() * float(i) + ra[i] * float(count - i)) / parent_area;
                if (cost < obj_cost)
                    obj_cost = cost, obj_pos = i;
            }
            if (obj_pos > 0 && obj_cost < object_bias * best_cost)
                mid = base + obj_pos;
        }
        if (mid >= 0) {
        } else if (best_axis >= 0 && (best_cost < float(count) || count > max_leaf)) {
            const float ext = cb.hi[best_axis] - cb.lo[best_axis], scale = float(kBins) / ext, lo = cb.lo[best_axis];
            const int axis = best_axis, bin = best_bin;
            auto it = std::partition(refs.begin() + base, refs.begin() + base + count, [&](const PrimRef &r) {
                return std::min(kBins - 1, std::max(0, int((r.ctr[axis] - lo) * scale))) <= bin;
            });
            mid = int(it - refs.begin());
        } else if (count > max_leaf) { // degenerate centroids: median split
            mid = base + count / 2;
        }
        if (mid <= base || mid >= base + count) {
            if (count > max_leaf)
                mid = base + count / 2;
            else
                return id;
        }
        const int l = build(base, mid - base, depth + 1);
        const int r = build(mid, base + count - mid, depth + 1);
        nodes[id].left = l;
        nodes[id].right = r;
        return id;
    }
};


// ---- insertion-based optimisation of the finished tree (Bittner, Hapala, Havran 2013: "Fast insertion-based optimization
// of bounding volume hierarchies"): take a subtree out, re-insert it where the surface area of the tree grows least
// (branch-and-bound search from the root), most wasteful nodes first.  Greedy top-down SAH cannot undo an early split;
// this can.  Leaves and their primitive ranges are untouched; afterwards the references are re-packed in DFS leaf order.
struct TreeOptimizer {
    std::vector<BuildNode> &n;
    std::vector<int> parent;
    explicit TreeOptimizer(std::vector<BuildNode> &nodes) : n(nodes), parent(nodes.size(), -1) {
        for (size_t i = 0; i < n.size(); i++)
            if (n[i].left >= 0)
                parent[size_t(n[i].left)] = int(i), parent[size_t(n[i].right)] = int(i);
    }
    static Box merged(const Box &a, const Box &b) {
        Box r = a;
        r.grow(b);
        return r;
    }
    void refit_up(int i) {
        for (; i >= 0; i = parent[size_t(i)]) {
            Box b = n[size_t(n[size_t(i)].left)].box;
            b.grow(n[size_t(n[size_t(i)].right)].box);
            n[size_t(i)].box = b;
        }
    }
    double total_area() const {
        double a = 0;
        for (size_t i = 0; i < n.size(); i++)
            if (n[i].left >= 0)
                a += n[i].box.area();
        return a;
    }
    // returns true if the tree changed
    bool reinsert(int x) {
        const int p = parent[size_t(x)];
        if (p <= 0)
            return false; // root or child of the root
        const int g = parent[size_t(p)], s = n[size_t(p)].left == x ? n[size_t(p)].right : n[size_t(p)].left;
        // detach x and p: s takes p's place
        (n[size_t(g)].left == p ? n[size_t(g)].left : n[size_t(g)].right) = s;
        parent[size_t(s)] = g;
        refit_up(g);
        // branch and bound for the cheapest new sibling
        const Box xb = n[size_t(x)].box;
        const float xa = xb.area();
        struct Cand {
            float induced;
            int node;
            bool operator<(const Cand &o) const { return induced > o.induced; } // min-heap on the induced cost
        };
        std::priority_queue<Cand> pq;
        { // the root itself is not a candidate (node 0 stays the root): start from its children
            const float root_induced = merged(n[0].box, xb).area() - n[0].box.area();
            pq.push(Cand{root_induced, n[0].left});
            pq.push(Cand{root_induced, n[0].right});
        }
        float best_cost = 1e30f;
        int best = s;
        while (!pq.empty()) {
            const Cand c = pq.top();
            pq.pop();
            if (c.induced + xa >= best_cost)
                break;
            const BuildNode &nd = n[size_t(c.node)];
            const float direct = merged(nd.box, xb).area(), cost = c.induced + direct;
            if (cost < best_cost)
                best_cost = cost, best = c.node;
            const float child_induced = cost - nd.box.area();
            if (nd.left >= 0 && child_induced + xa < best_cost) {
                pq.push(Cand{child_induced, nd.left});
                pq.push(Cand{child_induced, nd.right});
            }
        }
        // p becomes the parent of (best, x) in best's place
        const int bp = parent[size_t(best)];
        n[size_t(p)].left = best, n[size_t(p)].right = x;
        parent[size_t(best)] = p, parent[size_t(x)] = p;
        parent[size_t(p)] = bp;
        if (bp >= 0)
            (n[size_t(bp)].left == best ? n[size_t(bp)].left : n[size_t(bp)].right) = p;
        refit_up(p);
        return best != s;
    }
    void run(int passes, float fraction) {
        for (int pass = 0; pass < passes; pass++) {
            // candidates: inner and leaf nodes, largest surface area first (the ones that cost the most when badly placed)
            std::vector<int> order;
            order.reserve(n.size());
            for (size_t i = 1; i < n.size(); i++)
                if (parent[i] > 0)
                    order.push_back(int(i));
            std::sort(order.begin(), order.end(), [&](int a, int b) {
                const float aa = n[size_t(a)].box.area(), ab = n[size_t(b)].box.area();
                return aa > ab || (aa == ab && a < b);
            });
            const size_t count = size_t(double(order.size()) * fraction);
            for (size_t k = 0; k < count; k++)
                reinsert(order[k]);
        }
    }
};

void optimize_tree(Builder &b, int passes, float fraction) {
    const double before = TreeOptimizer(b.nodes).total_area();
    std::vector<BuildNode> backup = b.nodes;
    {
        TreeOptimizer opt(b.nodes);
        opt.run(passes, fraction);
        if (!(opt.total_area() < before)) {
            b.nodes.swap(backup);
            return;
        }
    }
    // depths, and references re-packed in DFS leaf order (leaf ranges ascend along the reference-order traversal)
    std::vector<PrimRef> packed;
    packed.reserve(b.refs.size());
    int max_depth = 0;
    std::vector<std::pair<int, int>> stack(1, {0, 0});
    while (!stack.empty()) {
        const auto [id, depth] = stack.back();
        stack.pop_back();
        BuildNode &nd = b.nodes[size_t(id)];
        nd.depth = depth;
        max_depth = std::max(max_depth, depth);
        if (nd.left < 0) {
            const int nb = int(packed.size());
            for (int i = nd.base; i < nd.base + nd.count; i++)
                packed.push_back(b.refs[size_t(i)]);
            nd.base = nb;
        } else {
            stack.push_back({nd.right, depth + 1});
            stack.push_back({nd.left, depth + 1});
        }
    }
    if (max_depth > 40 || packed.size() != b.refs.size()) { // keep the traversal stack of the renderer safe
        b.nodes.swap(backup);
        return;
    }
    b.refs.swap(packed);
    b.max_depth = max_depth;
}

uint16_t f2h(float f) { // round-to-nearest-even half (see scene_host.cpp); local copy to keep this TU independent
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7FFFFFFFu;
    if (x >= 0x7F800000u)
        return uint16_t(sign | 0x7C00u | (x > 0x7F800000u ? 0x0200u : 0u));
    if (x >= 0x477FF000u)
        return uint16_t(sign | 0x7C00u);
    if (x < 0x33000001u)
        return uint16_t(sign);
    const int exp = int(x >> 23) - 127;
    const uint32_t man = (x & 0x7FFFFFu) | 0x800000u;
    const int shift = exp < -14 ? 13 + (-14 - exp) : 13;
    uint32_t hm = man >> shift;
    const uint32_t rem = man & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (hm & 1u)))
        hm++;
    const uint32_t out = exp < -14 ? hm : ((uint32_t(exp + 15) << 10) + (hm - 0x400u));
    return uint16_t(sign | out);
}

int linearize(const Builder &b, int id, int depth, int cache_levels, std::vector<F4> &nodes, std::vector<U4> &cache) {
    // DFS pre-order; inner nodes store -(subtree size), leaves their primitive count (bvh.cuh:322-377)
    const BuildNode &n = b.nodes[id];
    const size_t my = nodes.size() / 2, my_cache = cache.size();
    auto as_float = [](int32_t v) {
        float f;
        memcpy(&f, &v, 4);
        return f;
    };
    nodes.push_back(F4{n.box.lo[0], n.box.lo[1], n.box.lo[2], as_float(n.base)});
    nodes.push_back(F4{n.box.hi[0], n.box.hi[1], n.box.hi[2], as_float(n.count)});
    if (depth < cache_levels) {
        U4 c;
        c.x = uint32_t(f2h(n.box.lo[0])) | (uint32_t(f2h(n.box.hi[0])) << 16);
        c.y = uint32_t(f2h(n.box.lo[1])) | (uint32_t(f2h(n.box.hi[1])) << 16);
        c.z = uint32_t(f2h(n.box.lo[2])) | (uint32_t(f2h(n.box.hi[2])) << 16);
        c.w = (uint32_t(my) << 8) | 1u;
        cache.push_back(c);
    }
    if (n.left < 0)
        return 1;
    int sub = linearize(b, n.left, depth + 1, cache_levels, nodes, cache);
    sub += linearize(b, n.right, depth + 1, cache_levels, nodes, cache);
    nodes[2 * my + 1].w = as_float(-(sub + 1));
    if (depth < cache_levels) {
        const uint32_t skip = uint32_t(cache.size() - my_cache);
        cache[my_cache].w = (cache[my_cache].w & 0xFFFFFF00u) | (skip & 0xFFu);
    }
    return sub + 1;
}

// ---- spectral tables for the dispersion BSDF (ColorSpaceXYZ::init, src/impl/xyz.cu:808-830): CIE 1931 2-degree
// observer 360-830 nm and D65 300-830 nm at 1 nm.  The numbers ship as data/spectral_tables.bin next to the library
// (read back at run time from the reference's own device textures by the golden-vector harness, tests/golden/README.md);
// CPTB200_DATA overrides the directory.
std::string data_dir() {
    std::string dir;
    if (const char *e = getenv("CPTB200_DATA"))
        dir = e;
    else {
        Dl_info info{};
        if (dladdr(reinterpret_cast<const void *>(&data_dir), &info) && info.dli_fname) {
            dir = info.dli_fname;
            const size_t slash = dir.find_last_of('/');
            dir = (slash == std::string::npos ? std::string(".") : dir.substr(0, slash)) + "/data";
        }
    }
    return dir;
}

// emission colour of grid media by normalised temperature: 1024 x (r, g, b, 1), data/blackbody_table.bin next to the
// library (the table the reference ships as scene/data/blackbody.bin, carried over through a scene dump)
void load_blackbody_table(HostScene &out) {
    const std::string path = data_dir() + "/blackbody_table.bin";
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp)
        throw std::runtime_error("scene: grid media need the blackbody table, cannot open '" + path + "'");
    char magic[8];
    int32_t n[2] = {0, 0};
    bool ok = fread(magic, 1, 8, fp) == 8 && memcmp(magic, "CPTBBDY1", 8) == 0 && fread(n, 4, 2, fp) == 2 && n[0] > 0 && n[0] <= 65536;
    if (ok) {
        out.blackbody.resize(size_t(n[0]));
        ok = fread(out.blackbody.data(), sizeof(F4), out.blackbody.size(), fp) == out.blackbody.size();
    }
    fclose(fp);
    if (!ok)
        throw std::runtime_error("scene: malformed blackbody table file '" + path + "'");
}

void load_spectral_tables(HostScene &out) {
    const std::string path = data_dir() + "/spectral_tables.bin";
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp)
        throw std::runtime_error("scene: dispersion BSDF needs the spectral tables, cannot open '" + path + "'");
    char magic[8];
    int32_t n[2] = {0, 0};
    bool ok = fread(magic, 1, 8, fp) == 8 && memcmp(magic, "CPTSPEC1", 8) == 0 && fread(n, 4, 2, fp) == 2 && n[0] == 471 && n[1] == 531;
    if (ok) {
        out.cie_xyz.resize(size_t(n[0]));
        out.d65_spd.resize(size_t(n[1]));
        ok = fread(out.cie_xyz.data(), sizeof(F4), out.cie_xyz.size(), fp) == out.cie_xyz.size() &&
             fread(out.d65_spd.data(), sizeof(float), out.d65_spd.size(), fp) == out.d65_spd.size();
    }
    fclose(fp);
    if (!ok)
        throw std::runtime_error("scene: malformed spectral table file '" + path + "'");
}

// ---- image textures (src/impl/textures.cu:53-81,92-131,169-203) ----
struct TexInfo {
    std::string diff, spec, glos, rough1, rough2, normal;
    bool rough_ior = false;
};

// float4 texels = byte / 255 * scale + offset; normal maps use scale 2, offset -1
int add_texture4(HostScene &out, const std::string &path, bool normal_map) {
    HostTexture t;
    const std::vector<uint8_t> px = load_image8(path, 4, t.width, t.height);
    const float offset = normal_map ? -1.f : 0.f, scale = normal_map ? 2.f : 1.f;
    t.channels = 4;
    t.texels.resize(px.size());
    for (size_t i = 0; i < px.size(); i++)
        t.texels[i] = (static_cast<float>(px[i]) / 255.0f) * scale + offset;
    out.textures.push_back(std::move(t));
    return int(out.textures.size()) - 1;
}

float roughness_to_alpha_host(float roughness) { // cuda_utils.cuh:115-120 (host instantiation: libm logf)
    roughness = std::fmax(roughness, 1e-3f);
    const float x = std::log(roughness);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

// two single-channel maps packed as float2.  Upstream reads the second map one byte ahead (`data2[i + 1]`,
// textures.cu:122); that is reproduced, with the out-of-range read of the very last texel clamped
int add_texture2(HostScene &out, const std::string &path1, const std::string &path2, bool is_ior) {
    HostTexture t;
    const std::vector<uint8_t> a = load_image8(path1, 1, t.width, t.height);
    std::vector<uint8_t> b;
    if (path2.length() > 1) {
        int w2 = 0, h2 = 0;
        b = load_image8(path2, 1, w2, h2);
        if (w2 != t.width || h2 != t.height)
            throw std::runtime_error("scene: composed roughness images differ in size");
    }
    const float offset = is_ior ? 1.f : 0.f, scale = is_ior ? 1.5f : 1.f;
    t.channels = 2;
    t.texels.resize(a.size() * 2);
    for (size_t i = 0; i < a.size(); i++) {
        float v1 = static_cast<float>(a[i]) / 255.0f;
        float v2 = b.empty() ? v1 : static_cast<float>(b[std::min(i + 1, b.size() - 1)]) / 255.0f;
        v1 = v1 * scale + offset;
        v2 = v2 * scale + offset;
        t.texels[2 * i] = is_ior ? roughness_to_alpha_host(v1) : v1;
        t.texels[2 * i + 1] = is_ior ? roughness_to_alpha_host(v2) : v2;
    }
    out.textures.push_back(std::move(t));
    return int(out.textures.size()) - 1;
}

} // namespace

void load_xml(const std::string &path, HostScene &out) {
    const XNode root = parse_xml_file(path);
    if (root.tag != "scene")
        throw std::runtime_error("scene: root element must be <scene>");
    if (auto v = root.attr("version"))
        if (*v != "1.2")
            throw std::runtime_error("scene: version required: '1.2', got '" + *v + "'");
    std::string folder;
    {
        size_t pos = path.find_last_of("/\\");
        folder = pos == std::string::npos ? "" : path.substr(0, pos + 1);
    }
    out = HostScene();

    // ---- renderer / accelerator / film (config.cpp:27-157, scene.cu:975-1000)
    const XNode *rdr = root.first("renderer"), *acc = root.first("accelerator"), *sensor = root.first("sensor");
    if (!sensor)
        throw std::runtime_error("scene: <sensor> missing");
    int spp = 256, max_depth = 8, max_diffuse = 4, max_specular = 2, max_transmit = 8, max_volume = 3;
    bool have_depth = false;
    int rdr_type = 0;
    if (rdr) {
        const std::string t = rdr->attr_or("type", "pt");
        rdr_type = t == "wfpt" ? 1 : (t == "ptd" ? 8 : (t == "lt" ? 2 : (t == "depth" ? 4 : (t == "vpt" ? 6 : 0))));
        get_int(*rdr, {"sample_count"}, spp);
        have_depth = get_int(*rdr, {"max_bounce"}, max_depth);
        get_int(*rdr, {"max_diffuse"}, max_diffuse);
        get_int(*rdr, {"max_specular"}, max_specular);
        get_int(*rdr, {"max_transmit"}, max_transmit);
        get_int(*rdr, {"max_volume"}, max_volume);
    }
    if (!have_depth) // MaxDepthParams default constructor (max_depth.h:35-41)
        max_depth = std::max(std::max(8, 8), std::max(4, 2));
    int cache_level = 4, max_node_num = 16;
    bool use_sbvh = false;
    if (acc) {
        if (get_int(*acc, {"cache_level"}, cache_level))
            cache_level = std::max(std::min(cache_level, 8), 0);
        if (get_int(*acc, {"max_node_num"}, max_node_num))
            max_node_num = std::min(std::max(max_node_num, 1), 255);
        get_bool(*acc, {"use_sbvh"}, use_sbvh);
    }
    int width = 1024, height = 1024;
    bool gamma = true;
    if (auto film = sensor->first("film")) {
        get_int(*film, {"width"}, width);
        get_int(*film, {"height"}, height);
        get_bool(*film, {"gamma_correction"}, gamma);
    }

    // ---- texture declarations (scene.cu:662-700): id -> image paths
    std::map<std::string, TexInfo> tex_map;
    for (auto t : root.children("texture")) {
        TexInfo info;
        for (auto c : t->children("string")) {
            const std::string n = c->attr_or("name", ""), v = folder + c->attr_or("value", "");
            if (n == "diffuse" || n == "emission")
                info.diff = v;
            else if (n == "specular")
                info.spec = v;
            else if (n == "glossy" || n == "sigma_a")
                info.glos = v;
            else if (n == "rough1" || n == "roughness_1" || n == "ior")
                info.rough1 = v, info.rough_ior = n == "ior";
            else if (n == "rough2" || n == "roughness_2")
                info.rough2 = v, info.rough_ior = false;
            else if (n == "normal")
                info.normal = v;
            else
                throw std::runtime_error("scene: unsupported texture type '" + n + "'");
        }
        tex_map.emplace(t->attr_or("id", ""), info); // first declaration of an id wins (unordered_map::emplace)
    }
    // ---- phase functions and media (scene.cu:702-900): index 0 of each table is the dummy / vacuum entry
    std::map<std::string, int> phase_ids, medium_ids;
    std::vector<std::pair<int, std::array<float, 4>>> phases(1, {CPTS_PHASE_ISOTROPIC, {0, 0, 0, 0}});
    for (auto ph : root.children("phase")) {
        std::string type = ph->attr_or("type", "hg");
        if (type.empty())
            type = "hg";
        std::pair<int, std::array<float, 4>> rec{CPTS_PHASE_ISOTROPIC, {0, 0, 0, 0}}; // Vec4() = (0,0,0,0)
        if (type == "hg") {
            float g = 0.2f;
            if (auto f = ph->first("float")) // only the first <float> child is inspected upstream
                if (f->attr_or("name", "") == "g")
                    g = float(atof(f->attr_or("value", "0.2").c_str()));
            rec.first = CPTS_PHASE_HG, rec.second[0] = g;
        } else if (type == "hg-duo") {
            float g1 = 0.2f, g2 = 0.8f, w = 0.5f;
            if (auto f = ph->first("float")) { // upstream advances with NextSiblingElement("string"): one <float> only
                const std::string n = f->attr_or("name", "");
                const float v = float(atof(f->attr_or("value", "0").c_str()));
                if (n == "g1")
                    g1 = v;
                else if (n == "g2")
                    g2 = v;
                else if (n == "weight")
                    w = v;
            }
            rec.first = CPTS_PHASE_DUO_HG, rec.second = {g1, g2, w, 1.f}; // Vec4(g1, g2, weight): w = 1
        } else if (type == "rayleigh") {
            rec.first = CPTS_PHASE_RAYLEIGH;
        } // "isotropic", "sggx" and unknown types fall back to isotropic upstream
        phase_ids[ph->attr_or("id", "")] = int(phases.size());
        phases.push_back(rec);
    }
    {
        int idx = 1;
        for (auto md : root.children("medium")) {
            const std::string id = md->attr_or("id", ""), type = md->attr_or("type", "homogeneous");
            if (type != "homogeneous" && type != "grid")
                throw std::runtime_error("scene: medium '" + id + "' has an unknown type '" + type + "'");
            int phase_id = 0;
            if (auto r = md->first("ref"))
                if (r->attr_or("type", "") == "phase") {
                    auto it = phase_ids.find(r->attr_or("id", ""));
                    if (it != phase_ids.end())
                        phase_id = it->second;
                }
            if (phase_id == 0)
                throw std::runtime_error("scene: volume '" + id + "' has no valid phase function");
            float scale = 1.f;
            get_float(*md, {"scale", "scaler"}, scale);
            float sa[4] = {0, 0, 0, 1}, ss[4] = {1, 1, 1, 1};
            for (auto c : md->children("rgb")) {
                const std::string n = c->attr_or("name", "");
                if (n == "sigma_a")
                    parse_color(c->attr_or("value", "0"), sa);
                else if (n == "sigma_s")
                    parse_color(c->attr_or("value", "1"), ss);
            }
            if (out.media.empty())
                out.media.push_back(CptsMedium{}); // vacuum
            CptsMedium rec{};
            rec.type = CPTS_MEDIUM_HOMOGENEOUS, rec.phase_type = phases[phase_id].first;
            for (int k = 0; k < 4; k++) {
                rec.sigma_a[k] = sa[k] * scale, rec.sigma_s[k] = ss[k] * scale; // create_homogeneous_volume (vol_homo.cu:26-32)
                rec.phase[k] = phases[phase_id].second[k];
            }
            if (type == "grid") { // scene.cu:850-892, GridVolumeManager::push / to_gpu (vol_grid.cu:226-302)
                std::string den, alb, ems;
                for (auto c : md->children("string")) {
                    const std::string n = c->attr_or("name", "");
                    if (n == "density")
                        den = folder + c->attr_or("value", "");
                    else if (n == "albedo")
                        alb = folder + c->attr_or("value", "");
                    else if (n == "emission")
                        ems = folder + c->attr_or("value", "");
                }
                if (den.empty())
                    throw std::runtime_error("scene: grid medium '" + id + "' has no density grid");
                float albedo[4] = {1, 1, 1, 1}, temp_scale = 1.f, em_scale = 1.f;
                if (auto c = md->first("rgb")) // only the first <rgb> is looked at (scene.cu:866-868)
                    if (c->attr_or("name", "") == "albedo")
                        parse_color(c->attr_or("value", "1"), albedo);
                get_float(*md, {"temp-scale", "temp_scale"}, temp_scale);
                get_float(*md, {"emission-scale", "em-scale"}, em_scale);
                rec.type = CPTS_MEDIUM_GRID;
                for (int k = 0; k < 4; k++) // MediumParams(albedo, (emission scale, temperature scale, 0), phase, scale)
                    rec.sigma_a[k] = albedo[k];
                rec.sigma_s[0] = em_scale, rec.sigma_s[1] = temp_scale, rec.sigma_s[2] = 0.f, rec.sigma_s[3] = 1.f;
                rec.grid = int(out.grids.size());
                CptsGrid g{};
                g.medium = idx;
                g.scale = scale, g.temp_scale = temp_scale, g.em_scale = em_scale;
                auto append = [&](const std::string &file, int ncomp, int slot) -> NvdbGrid {
                    NvdbGrid ng;
                    read_nvdb(file, ncomp, ng);
                    ng.channel.brick_offset = out.grid_bricks.size();
                    ng.channel.data_offset = out.grid_data.size();
                    out.grid_bricks.insert(out.grid_bricks.end(), ng.bricks.begin(), ng.bricks.end());
                    out.grid_data.insert(out.grid_data.end(), ng.data.begin(), ng.data.end());
                    g.ch[slot] = ng.channel;
                    return ng;
                };
                {
                    const NvdbGrid d = append(den, 1, CPTS_GRID_DENSITY);
                    g.majorant = d.value_max; // Tree::extrema (vol_grid.cu:131-133)
                    for (int k = 0; k < 3; k++)
                        g.aabb_min[k] = float(d.world_bbox[k]), g.aabb_max[k] = float(d.world_bbox[3 + k]);
                    double sum = 0;
                    size_t cnt = 0;
                    for (float v : d.data)
                        if (v > 0)
                            sum += v, cnt++;
                    g.avg_density = cnt ? float(sum / double(cnt)) : 0.f;
                }
                if (!alb.empty())
                    append(alb, 3, CPTS_GRID_ALBEDO);
                if (!ems.empty())
                    append(ems, 1, CPTS_GRID_TEMPERATURE);
                out.grids.push_back(g);
            }
            out.media.push_back(rec);
            medium_ids[id] = idx++;
        }
    }

    // ---- BSDFs (scene.cu:184-359)
    std::map<std::string, int> bsdf_ids;
    for (auto b : root.children("brdf")) {
        CptsBsdf rec{};
        for (int k = 0; k < 5; k++)
            rec.tex[k] = -1;
        const std::string type = b->attr_or("type", "lambertian"), id = b->attr_or("id", "");
        bsdf_ids[id] = int(out.bsdfs.size());
        for (auto c : b->children("rgb")) {
            const std::string name = c->attr_or("name", "");
            float col[4];
            parse_color(c->attr_or("value", "0"), col);
            float *dst = name == "k_d" ? rec.k_d : (name == "k_s" ? rec.k_s : ((name == "k_g" || name == "sigma_a") ? rec.k_g : nullptr));
            if (dst)
                memcpy(dst, col, sizeof(col));
        }
        if (auto r = b->first("ref")) // only the first <ref> child is inspected upstream (scene.cu:193-236)
            if (r->attr_or("type", "") == "texture") {
                auto it = tex_map.find(r->attr_or("id", ""));
                if (it == tex_map.end())
                    throw std::runtime_error("scene: BSDF '" + id + "' references an unknown texture");
                const TexInfo &ti = it->second;
                if (!ti.diff.empty())
                    rec.tex[0] = add_texture4(out, ti.diff, false);
                if (!ti.spec.empty())
                    rec.tex[1] = add_texture4(out, ti.spec, false);
                if (!ti.glos.empty())
                    rec.tex[2] = add_texture4(out, ti.glos, false);
                if (!ti.normal.empty())
                    rec.tex[3] = add_texture4(out, ti.normal, true);
                if (!ti.rough1.empty())
                    rec.tex[4] = add_texture2(out, ti.rough1, ti.rough2, ti.rough_ior);
            }
        if (type == "lambertian") {
            rec.type = CPTS_BSDF_LAMBERTIAN, rec.flags = 0x01 | 0x10;
        } else if (type == "specular") {
            rec.type = CPTS_BSDF_SPECULAR, rec.flags = 0x02 | 0x10;
        } else if (type == "det-refraction") {
            rec.type = CPTS_BSDF_TRANSLUCENT, rec.flags = 0x02 | 0x20;
        } else if (type == "conductor-ggx") {
            rec.type = CPTS_BSDF_GGX_CONDUCTOR, rec.flags = 0x04 | 0x10;
            std::string metal = "Cu";
            if (auto s = b->first("string")) { // only the first <string> child is inspected upstream
                const std::string n = s->attr_or("name", "");
                if (n == "type" || n == "metal" || n == "conductor")
                    metal = s->attr_or("value", "Cu");
            }
            const Metal *mt = &kMetals[2];
            for (const auto &m : kMetals)
                if (metal == m.name)
                    mt = &m;
            float rx = 0.1f, ry = 0.1f;
            if (get_float(*b, {"roughness_x", "rough_x"}, rx))
                rx = std::min(std::max(rx, 0.001f), 1.f);
            if (get_float(*b, {"roughness_y", "rough_y"}, ry))
                ry = std::min(std::max(ry, 0.001f), 1.f);
            for (int k = 0; k < 3; k++)
                rec.extras[k] = mt->eta[k], rec.extras[3 + k] = mt->k[k];
            rec.extras[6] = rx, rec.extras[7] = ry;
        } else if (type == "plastic" || type == "plastic-forward") {
            rec.type = type == "plastic" ? CPTS_BSDF_PLASTIC : CPTS_BSDF_PLASTIC_FORWARD;
            rec.flags = type == "plastic" ? (0x02 | 0x01 | 0x10) : (0x02 | 0x20 | 0x10);
            rec.k_g[0] = rec.k_g[1] = rec.k_g[2] = 0.f, rec.k_g[3] = 1.f; // scene.cu:293
            float ts = 1.f, th = 0.f, ior = 1.33f;
            get_float(*b, {"trans_scaler"}, ts);
            get_float(*b, {"thickness"}, th);
            get_float(*b, {"ior"}, ior);
            bool pen = false;
            if (auto fb = b->first("bool")) // only the first <bool> child is inspected upstream
                if (fb->attr_or("name", "") == "penetrable")
                    pen = fb->attr_or("value", "false") == "true";
            rec.extras[0] = ior, rec.extras[1] = ts, rec.extras[2] = th, rec.extras[3] = pen ? 1.f : 0.f;
        } else if (type == "dispersion") {
            rec.type = CPTS_BSDF_DISPERSION, rec.flags = 0x01 | 0x20;
            std::string name = "Diamond";
            if (auto s = b->first("string")) {
                const std::string n = s->attr_or("name", "");
                if (n == "type" || n == "dielectric")
                    name = s->attr_or("value", "Diamond");
            }
            const Dielectric *d = &kDielectrics[0];
            for (const auto &c : kDielectrics)
                if (name == c.name)
                    d = &c;
            rec.extras[0] = d->a, rec.extras[1] = d->b;
        } else if (type == "forward") {
            rec.type = CPTS_BSDF_FORWARD, rec.flags = 0;
        } else {
            throw std::runtime_error("scene: unknown BSDF type '" + type + "'");
        }
        out.bsdfs.push_back(rec);
    }
    if (out.bsdfs.size() > 48)
        throw std::runtime_error("scene: too many BSDFs (max 48)");

    // ---- emitter names (scene.cu:361-369)
    std::map<std::string, int> emitter_ids;
    {
        int idx = 1;
        for (auto e : root.children("emitter"))
            emitter_ids[e->attr_or("id", "")] = idx++;
    }

    // ---- shapes (scene.cu:570-660, object.cu:30-58)
    struct Obj {
        int bsdf, emitter, prim_offset, prim_num;
        float inv_area;
    };
    std::vector<Obj> objs;
    std::array<std::vector<F3>, 3> V, N;
    std::array<std::vector<F2>, 3> UV;
    std::map<std::string, int> emitter_obj;
    std::vector<int> obj_medium;
    for (auto sh : root.children("shape")) {
        const std::string type = sh->attr_or("type", "obj");
        if (type != "obj")
            throw std::runtime_error("scene: shape type '" + type + "' is not supported (the reference's default "
                                     "build is TRIANGLE_ONLY, defines.cuh:26)");
        std::string filename;
        if (auto f = named_child(*sh, "string", {"filename"}))
            filename = folder + f->attr_or("value", "");
        int bsdf = -1, emitter = 0;
        for (auto r : sh->children("ref")) {
            const std::string t = r->attr_or("type", ""), id = r->attr_or("id", "");
            if (t == "material") {
                auto it = bsdf_ids.find(id);
                if (it == bsdf_ids.end())
                    throw std::runtime_error("scene: unknown material '" + id + "'");
                bsdf = it->second;
            } else if (t == "emitter") {
                auto it = emitter_ids.find(id);
                if (it == emitter_ids.end())
                    throw std::runtime_error("scene: unknown emitter '" + id + "'");
                emitter = it->second;
                emitter_obj[id] = int(objs.size());
            }
        }
        if (bsdf < 0)
            throw std::runtime_error("scene: object with no BSDF");
        // medium bound by this shape + "cullable" flag (parseObjectMediumRef, scene.cu:903-932)
        int medium_word = 0;
        for (auto r : sh->children("ref"))
            if (r->attr_or("type", "") == "medium") {
                auto it = medium_ids.find(r->attr_or("id", ""));
                if (it != medium_ids.end())
                    medium_word = it->second & 0xff;
            }
        if (medium_word) {
            if (auto fb = sh->first("bool")) // only the first <bool> child is inspected upstream
                if (fb->attr_or("name", "") == "cullable" && fb->attr_or("value", "false") == "true")
                    medium_word += 1 << 10;
        }
        obj_medium.push_back(medium_word);
        ObjMesh mesh;
        load_obj(filename, mesh);
        Obj o{bsdf, emitter, int(V[0].size()), int(mesh.v[0].size()), 0.f};
        float area = 0;
        for (size_t i = 0; i < mesh.v[0].size(); i++) {
            const F3 a = mesh.v[0][i], b = mesh.v[1][i], c = mesh.v[2][i];
            const float e1x = b.x - a.x, e1y = b.y - a.y, e1z = b.z - a.z, e2x = c.x - a.x, e2y = c.y - a.y, e2z = c.z - a.z;
            const float cx = std::fmaf(e1y, e2z, -e1z * e2y), cy = std::fmaf(e1z, e2x, -e1x * e2z),
                        cz = std::fmaf(e1x, e2y, -e1y * e2x);
            area += std::sqrt(std::fmaf(cx, cx, std::fmaf(cy, cy, cz * cz)));
        }
        area *= 0.5f;
        o.inv_area = 1.f / area;
        for (int k = 0; k < 3; k++) {
            V[k].insert(V[k].end(), mesh.v[k].begin(), mesh.v[k].end());
            N[k].insert(N[k].end(), mesh.n[k].begin(), mesh.n[k].end());
            UV[k].insert(UV[k].end(), mesh.uv[k].begin(), mesh.uv[k].end());
        }
        objs.push_back(o);
    }
    if (objs.empty() || V[0].empty())
        throw std::runtime_error("scene: no geometry");
    if (V[0].size() > 64000000)
        throw std::runtime_error("scene: too many primitives");

    // ---- emitters (scene.cu:371-481)
    int envmap_id = 0;
    {
        CptsEmitter null_e{};
        null_e.type = CPTS_EMIT_NULL, null_e.obj_ref = -1, null_e.tex = -1, null_e.Le[3] = 1.f;
        out.emitters.push_back(null_e);
        for (auto e : root.children("emitter")) {
            CptsEmitter rec{};
            rec.tex = -1, rec.obj_ref = -1;
            const std::string type = e->attr_or("type", ""), id = e->attr_or("id", "");
            float emission[4] = {0, 0, 0, 1}, scaler[4] = {1, 1, 1, 1};
            if (auto c = named_child(*e, "rgb", {"emission"}))
                parse_color(c->attr_or("value", "0"), emission);
            if (auto c = named_child(*e, "rgb", {"scaler"}))
                parse_color(c->attr_or("value", "1"), scaler);
            for (int k = 0; k < 3; k++)
                rec.Le[k] = emission[k] * scaler[k];
            rec.Le[3] = 1.f * scaler[0];
            if (auto r = e->first("ref")) // scene.cu:395-417
                if (r->attr_or("type", "") == "texture") {
                    auto it = tex_map.find(r->attr_or("id", ""));
                    if (it == tex_map.end())
                        throw std::runtime_error("scene: emitter '" + id + "' references an unknown texture");
                    if (it->second.diff.empty())
                        throw std::runtime_error("scene: the texture of emitter '" + id + "' has no 'emission' image");
                    rec.tex = add_texture4(out, it->second.diff, false);
                }
            auto bound = emitter_obj.find(id);
            if (type == "point") {
                rec.type = CPTS_EMIT_POINT;
                if (auto p = named_child(*e, "point", {"center", "pos"})) {
                    rec.pos[0] = float(atof(p->attr_or("x", "0").c_str()));
                    rec.pos[1] = float(atof(p->attr_or("y", "0").c_str()));
                    rec.pos[2] = float(atof(p->attr_or("z", "0").c_str()));
                }
            } else if (type == "area" || type == "area-spot") {
                rec.type = type == "area" ? CPTS_EMIT_AREA : CPTS_EMIT_AREA_SPOT;
                auto bind = named_child(*e, "string", {"bind_type"});
                if (!bind)
                    throw std::runtime_error("scene: bound primitive is not specified for area source '" + id + "'");
                if (bind->attr_or("value", "") == "sphere")
                    throw std::runtime_error("scene: sphere-bound emitters are not supported");
                rec.obj_ref = bound == emitter_obj.end() ? 0 : bound->second; // unordered_map operator[] default
                if (type == "area-spot") {
                    rec.cos_val = 0.99f;
                    if (auto f = e->first("float")) {
                        const std::string n = f->attr_or("name", "");
                        if (n == "half-angle" || n == "angle")
                            rec.cos_val = std::cos(float(atof(f->attr_or("value", "0").c_str())) * kDeg2Rad);
                    }
                }
            } else if (type == "envmap") { // scene.cu:460-478
                rec.type = CPTS_EMIT_ENVMAP;
                envmap_id = int(out.emitters.size());
                float scale = 1.f, azimuth = 0.f, zenith = 0.f;
                get_float(*e, {"scaler"}, scale);
                get_float(*e, {"azimuth"}, azimuth);
                get_float(*e, {"zenith"}, zenith);
                rec.Le[0] = rec.Le[1] = rec.Le[2] = 0.f, rec.Le[3] = 1.f; // Emitter(Vec4(0, 1), -1, false), emitter.cuh:320
                rec.scale = scale, rec.azimuth = azimuth * kDeg2Rad, rec.zenith = zenith * kDeg2Rad;
                if (rec.tex < 0)
                    throw std::runtime_error("scene: the texture for the environment map is empty");
            } else {
                throw std::runtime_error("scene: unknown emitter type '" + type + "'");
            }
            out.emitters.push_back(rec);
        }
        if (out.emitters.size() > 9)
            throw std::runtime_error("scene: at most 8 emitters are supported (c_emitter[9])");
    }

    // ---- camera (camera_model.cu:51-65,83-153)
    {
        float fov = 0, focal_dist = 1, aperture = 0, hsign = 1;
        get_float(*sensor, {"fov"}, fov);
        get_float(*sensor, {"aperture"}, aperture);
        get_float(*sensor, {"focal_distance", "focal_dist"}, focal_dist);
        F3 target{0, 0, 0}, origin{0, 0, 0}, up{0, 1, 0};
        if (auto tr = sensor->first("transform")) {
            auto la = tr->first("lookat");
            if (!la)
                throw std::runtime_error("scene: camera element contains no 'lookat' transform");
            target = parse_vec3(la->attr_or("target", "0 0 0"));
            origin = parse_vec3(la->attr_or("origin", "0 0 0"));
            if (auto u = la->attr("up"))
                up = parse_vec3(*u);
        }
        for (auto b : sensor->children("bool"))
            if (b->attr_or("name", "") == "hflip" && b->attr_or("value", "") == "true")
                hsign = -hsign;
        auto len2 = [](F3 v) { return std::fmaf(v.x, v.x, std::fmaf(v.y, v.y, v.z * v.z)); };
        auto normalize = [&](F3 v) {
            const float s = 1.0f / std::sqrt(len2(v));
            return F3{v.x * s, v.y * s, v.z * s};
        };
        const F3 diff{target.x - origin.x, target.y - origin.y, target.z - origin.z};
        const F3 fwd = normalize(diff);
        up = normalize(up);
        const F3 right = normalize(F3{std::fmaf(up.y, fwd.z, -up.z * fwd.y), std::fmaf(up.z, fwd.x, -up.x * fwd.z),
                                      std::fmaf(up.x, fwd.y, -up.y * fwd.x)});
        float *c = out.camera;
        // rows of R = (right.k, up.k, forward.k): the three vectors are the COLUMNS (so3.cuh:62-68)
        c[0] = right.x, c[1] = up.x, c[2] = fwd.x;
        c[3] = right.y, c[4] = up.y, c[5] = fwd.y;
        c[6] = right.z, c[7] = up.z, c[8] = fwd.z;
        c[9] = origin.x, c[10] = origin.y, c[11] = origin.z;
        const float fov_rad = fov / 180.f * kPiF;
        const float focal = 0.5f * float(width) / std::tan(.5f * fov_rad);
        c[CPTS_CAM_INV_FOCAL] = 1.f / focal;
        c[CPTS_CAM_HW] = float(width) * 0.5f;
        c[CPTS_CAM_HH] = float(height) * 0.5f;
        c[CPTS_CAM_SIGN_X] = hsign;
        c[CPTS_CAM_APERTURE] = aperture;
        c[CPTS_CAM_FOCAL_DIST] = focal_dist <= 0 ? std::sqrt(len2(diff)) : focal_dist;
    }

    // ---- BVH
    const size_t P0 = V[0].size();
    Builder bld;
    bld.max_leaf = max_node_num;
    if (const char *e = getenv("CPTB200_SAH_TRAVERSE")) // builder tuning knob (experiments)
        bld.kTraverse = float(atof(e));
    if (const char *e = getenv("CPTB200_SAH_BINS"))
        bld.kBins = std::min(Builder::kMaxBins, std::max(4, atoi(e)));
    if (const char *e = getenv("CPTB200_SAH_SWEEP"))
        bld.sweep_below = std::max(0, atoi(e));
    if (const char *e = getenv("CPTB200_SAH_OBJDEPTH"))
        bld.object_split_depth = std::max(0, atoi(e));
    if (const char *e = getenv("CPTB200_SAH_OBJBIAS"))
        bld.object_bias = float(atof(e));
    bld.refs.resize(P0);
    {
        // primitive -> object word (bvh.cu:85-92; no media, triangles only)
        std::vector<int> prim_obj(P0);
        for (size_t o = 0; o < objs.size(); o++)
            for (int i = 0; i < objs[o].prim_num; i++)
                prim_obj[size_t(objs[o].prim_offset) + i] = int(o);
        for (size_t i = 0; i < P0; i++) {
            PrimRef &r = bld.refs[i];
            r.prim = int(i);
            for (int a = 0; a < 3; a++) {
                const float p[3] = {(&V[0][i].x)[a], (&V[1][i].x)[a], (&V[2][i].x)[a]};
                r.box.lo[a] = std::min(p[0], std::min(p[1], p[2])) - kAabbEps;
                r.box.hi[a] = std::max(p[0], std::max(p[1], p[2])) + kAabbEps;
                r.ctr[a] = (p[0] + p[1] + p[2]) * 0.33333333333f;
            }
        }
        bld.prim_obj = prim_obj.data();
        // Greedy top-down SAH is sensitive to where the bins fall near the root (a big floor under a dense model): the
        // total SAH cost of the finished tree is bimodal in the bin count (vader: 9.62 with 16 / 48 / 64 bins, 8.99 with
        // 24 / 32 / 40; whiskey 7.26 vs 6.43) and traversal time follows it (-2.5 % / -11 %).  So the tree is built with
        // several bin counts in parallel, each one optimised (optimize_tree), and the cheapest result is kept
        // (CPTB200_SAH_BINS pins a single count).
        auto sah_total = [](const Builder &b) {
            const float root = std::max(b.nodes[0].box.area(), 1e-20f);
            double cost = 0;
            for (const BuildNode &n : b.nodes)
                cost += double(n.box.area()) / root * (n.left < 0 ? double(n.count) : double(b.kTraverse));
            return cost;
        };
        int opt_passes = 2;
        float opt_fraction = 1.f;
        if (const char *e = getenv("CPTB200_TREE_OPT"))
            opt_passes = atoi(e);
        if (const char *e = getenv("CPTB200_TREE_OPT_FRACTION"))
            opt_fraction = float(atof(e));
        auto build_one = [P0, opt_passes, opt_fraction](Builder &b) {
            b.nodes.reserve(2 * P0);
            b.build(0, int(P0), 0);
            if (opt_passes > 0 && P0 > 64)
                optimize_tree(b, opt_passes, opt_fraction);
        };
        if (!getenv("CPTB200_SAH_BINS") && P0 > 64) {
            const int candidates[4] = {16, 24, 32, 40};
            std::vector<Builder> trial(4, bld);
            std::vector<std::thread> workers;
            for (int c = 0; c < 4; c++) {
                trial[c].kBins = candidates[c];
                workers.emplace_back([&trial, &build_one, c]() { build_one(trial[c]); });
            }
            for (auto &w : workers)
                w.join();
            int best = 0;
            for (int c = 1; c < 4; c++)
                if (sah_total(trial[c]) < sah_total(trial[best]))
                    best = c;
            bld = std::move(trial[best]);
        } else {
            build_one(bld);
        }
        // The cached near-root nodes hold fp16 boxes rounded to nearest (bvh.cuh:182-251, reproduced by the kernels): the
        // bounds of a large axis-aligned triangle (a wall: 2e-5 thick) collapse to zero thickness there and are never
        // entered, so a LEAF inside the cached levels loses its hits - the holes of the reference's own cornell image.
        // Our trees keep every leaf below the cached levels (the optimised tree lifts big triangles close to the root).
        int min_leaf_depth = bld.max_depth;
        {
            std::vector<std::pair<int, int>> st(1, {0, 0});
            while (!st.empty()) {
                const auto [id, depth] = st.back();
                st.pop_back();
                const BuildNode &nd = bld.nodes[size_t(id)];
                if (nd.left < 0)
                    min_leaf_depth = std::min(min_leaf_depth, depth);
                else if (depth + 1 < min_leaf_depth)
                    st.push_back({nd.left, depth + 1}), st.push_back({nd.right, depth + 1});
            }
        }
        const int cache_levels = std::min(std::min(std::max(bld.max_depth - 1, 0), cache_level), min_leaf_depth);
        linearize(bld, 0, 0, cache_levels, out.nodes, out.cache_nodes);
        cache_level = cache_levels;
        // reorder primitives into leaf order, build the emitter primitive map (bvh.cu:325-401)
        std::vector<std::vector<int>> eprims(out.emitters.size() - 1);
        out.obj_idxs.resize(P0);
        for (int k = 0; k < 3; k++) {
            out.verts[k].resize(P0);
            out.norms[k].resize(P0);
            out.uvs[k].resize(P0);
        }
        for (size_t i = 0; i < P0; i++) {
            const int src = bld.refs[i].prim, o = prim_obj[src];
            out.obj_idxs[i] = (o & 0x000fffff) | (obj_medium[o] << 20); // bits 20-27 medium, bit 30 cullable (bvh.cu:85-97)
            for (int k = 0; k < 3; k++) {
                out.verts[k][i] = V[k][src];
                out.norms[k][i] = N[k][src];
                out.uvs[k][i] = UV[k][src];
            }
            if (objs[o].emitter > 0)
                eprims[objs[o].emitter - 1].push_back(int(i));
        }
        std::vector<int> offsets(1, 0);
        for (auto &e : eprims) {
            offsets.push_back(offsets.back() + int(e.size()));
            out.emitter_prims.insert(out.emitter_prims.end(), e.begin(), e.end());
        }
        for (auto &o : objs)
            if (o.emitter > 0)
                o.prim_offset = offsets[o.emitter - 1];
    }
    for (const auto &o : objs) {
        CptsObject rec;
        rec.bsdf_emitter = (o.bsdf << 16) | (o.emitter & 0xFFFF);
        rec.prim_offset = o.prim_offset;
        rec.prim_num = o.prim_num;
        rec.inv_area = o.inv_area;
        out.objects.push_back(rec);
    }

    int32_t *cfg = out.config;
    cfg[CPTS_CFG_WIDTH] = width, cfg[CPTS_CFG_HEIGHT] = height, cfg[CPTS_CFG_SPP] = spp;
    cfg[CPTS_CFG_MAX_DEPTH] = max_depth, cfg[CPTS_CFG_MAX_DIFFUSE] = max_diffuse;
    cfg[CPTS_CFG_MAX_SPECULAR] = max_specular, cfg[CPTS_CFG_MAX_TRANSMIT] = max_transmit;
    cfg[CPTS_CFG_GAMMA] = gamma, cfg[CPTS_CFG_RENDERER] = rdr_type;
    cfg[CPTS_CFG_NUM_BSDFS] = int(out.bsdfs.size()), cfg[CPTS_CFG_NUM_EMITTERS] = int(out.emitters.size()) - 1;
    cfg[CPTS_CFG_ENVMAP_ID] = envmap_id, cfg[CPTS_CFG_NUM_OBJECTS] = int(out.objects.size());
    cfg[CPTS_CFG_NUM_PRIMS] = int(P0), cfg[CPTS_CFG_NUM_NODES] = out.num_nodes();
    cfg[CPTS_CFG_NUM_CACHE] = int(out.cache_nodes.size()), cfg[CPTS_CFG_CACHE_LEVEL] = cache_level;
    cfg[CPTS_CFG_MAX_NODE_NUM] = max_node_num, cfg[CPTS_CFG_USE_SBVH] = 0;
    cfg[CPTS_CFG_MAX_VOLUME] = max_volume, cfg[CPTS_CFG_NUM_MEDIA] = out.media.empty() ? 0 : int(out.media.size()) - 1;
    cfg[CPTS_CFG_CAM_VOL] = 0; // Scene::cam_vol_id is never set by the loader (scene.cu:939)
    for (const auto &b : out.bsdfs)
        if (b.type == CPTS_BSDF_DISPERSION && out.cie_xyz.empty()) {
            load_spectral_tables(out);
            break;
        }
    if (!out.grids.empty() && out.blackbody.empty())
        load_blackbody_table(out);
    validate_scene(out);
}

} // namespace cptb
