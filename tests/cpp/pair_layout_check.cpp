// Host-side check of the pair layout (scene_host.cpp build_pair_layout) incl. leaf refinement: every primitive is
// referenced by exactly one leaf record, refined parts stay inside their reference leaf's range and box, the permutation
// is a permutation that only moves primitives inside reference leaves.  Built and run by tests/test_host.py.
#include "scene_host.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
using namespace cptb;
int main(int argc, char **argv) {
    if (argc < 3)
        return 2;
    HostScene sc;
    load_cpts(argv[1], sc);
    PairLayout pl;
    build_pair_layout(sc, pl, atoi(argv[2]));
    const int P = sc.num_prims(), N = sc.num_nodes();
    std::vector<int> cov(P, 0);
    int errors = 0, leaf_records = 0;
    auto node_i = [&](int k, int which) { int32_t v; memcpy(&v, which ? &sc.nodes[2 * k + 1].w : &sc.nodes[2 * k].w, 4); return v; };
    std::function<void(int)> walk = [&](int cur) {
        for (int c = 0; c < 2; c++) {
            const F4 lo = pl.wnodes[4 * cur + 2 * c], hi = pl.wnodes[4 * cur + 2 * c + 1];
            int32_t ref, end;
            memcpy(&ref, &lo.w, 4), memcpy(&end, &hi.w, 4);
            if (lo.x > hi.x)
                continue; // empty slot of the root
            if (ref >= 0) {
                walk(ref);
                continue;
            }
            const int packed = ~ref, base = packed & ((1 << 26) - 1), cnt = packed >> 26;
            leaf_records++;
            // the record must lie inside the reference leaf whose DFS index is end - 1, and inside its box
            const int k = end - 1;
            if (k < 0 || k >= N || node_i(k, 1) <= 0) { errors++; continue; }
            const int lb = node_i(k, 0), lc = node_i(k, 1);
            if (base < lb || base + cnt > lb + lc) errors++;
            const F4 nlo = sc.nodes[2 * k], nhi = sc.nodes[2 * k + 1];
            if (lo.x < nlo.x || lo.y < nlo.y || lo.z < nlo.z || hi.x > nhi.x || hi.y > nhi.y || hi.z > nhi.z) errors++;
            for (int i = base; i < base + cnt; i++)
                cov[i]++;
        }
    };
    walk(pl.root);
    for (int v : cov)
        errors += v != 1;
    if (!pl.perm.empty()) {
        std::vector<int> seen(P, 0);
        for (int i = 0; i < P; i++) {
            const int o = pl.perm[i];
            if (o < 0 || o >= P || seen[o]++) { errors++; continue; }
        }
        // moves stay inside reference leaves
        for (int k = 0; k < N; k++)
            if (node_i(k, 1) > 0)
                for (int i = node_i(k, 0); i < node_i(k, 0) + node_i(k, 1); i++)
                    if (pl.perm[i] < node_i(k, 0) || pl.perm[i] >= node_i(k, 0) + node_i(k, 1)) errors++;
    }
    // wide layouts (build_wide_layout): every leaf record of the pair tree appears exactly once with the same box / reference /
    // DFS end, every inner slot points to a later node (breadth-first numbering, node 0 = root), the stack bound is a bound
    int wide_errors = 0, wide_stack = 0, wide4_stack_all = 0;
    for (int width : {4, 8}) {
        WideLayout wl;
        build_wide_layout(pl, width, wl);
        if (width == 4)
            wide4_stack_all = wl.max_stack_all;
        if (wl.max_stack_all < wl.max_stack)
            wide_errors++;
        std::vector<int> wcov(P, 0);
        int leaves = 0;
        std::function<void(int, int)> wwalk = [&](int node, int stack) {
            int inner = 0;
            for (int c = 0; c < width; c++) {
                int32_t ref;
                memcpy(&ref, &wl.nodes[(size_t(node) * width + c) * 2].w, 4);
                inner += ref >= 0;
            }
            const int need = stack + (inner > 1 ? inner - 1 : 0);
            if (need > wl.max_stack) wide_errors++;
            wide_stack = need > wide_stack ? need : wide_stack;
            for (int c = 0; c < width; c++) {
                const F4 lo = wl.nodes[(size_t(node) * width + c) * 2], hi = wl.nodes[(size_t(node) * width + c) * 2 + 1];
                int32_t ref, end;
                memcpy(&ref, &lo.w, 4), memcpy(&end, &hi.w, 4);
                if (ref == ~0 && end == 0)
                    continue; // empty slot
                if (ref >= 0) {
                    if (ref <= node || ref >= wl.num_nodes) { wide_errors++; continue; }
                    wwalk(ref, need);
                    continue;
                }
                const int packed = ~ref, base = packed & ((1 << 26) - 1), cnt = packed >> 26;
                leaves++;
                const int k = end - 1;
                if (k < 0 || k >= N) { wide_errors++; continue; }
                for (int i = base; i < base + cnt; i++)
                    wcov[i]++;
            }
        };
        wwalk(0, 0);
        for (int v : wcov)
            wide_errors += v != 1;
        wide_errors += leaves != leaf_records;
        wide_errors += wl.leaf_slots != leaf_records;
    }
    errors += wide_errors;
    printf("prims %d leaf_records %d refined_leaves %d depth %d nesting_violations %d clipped_refs %d wide_errors %d wide_stack %d wide4_stack_all %d errors %d\n", P, leaf_records,
           pl.refined_leaves, pl.max_depth, pl.nesting_violations, pl.clipped_refs, wide_errors, wide_stack, wide4_stack_all, errors);
    return errors ? 1 : 0;
}
