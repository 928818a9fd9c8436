/*
 * pt — offline renderer front end over the C ABI (include/cpt_b200.h), the counterpart of the reference's
 * `pt <scene.xml>` (app/pt_renderer.cu:26-117): load the scene, render `sample_count` samples with the
 * termination rules of the scene's <renderer type>, write ./render.png (RGBA8, DeviceImage::export_cpu
 * semantics incl. the 1/2.1 display gamma, host_device.cuh:63-112, cuda_utils.cuh:111-113).
 *
 * Plain C on purpose: it is also the proof that the boundary needs nothing but a C compiler.
 *     usage: pt <scene.xml | scene.cpts> [out.png] [device]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "cpt_b200.h"
#include "cpts_format.h"

static void put32(uint8_t *p, uint32_t v) { p[0] = v >> 24, p[1] = v >> 16, p[2] = v >> 8, p[3] = v; }

static int chunk(FILE *fp, const char *tag, const uint8_t *data, uint32_t len) {
    uint8_t hdr[8];
    put32(hdr, len);
    memcpy(hdr + 4, tag, 4);
    uint32_t crc = crc32(0L, hdr + 4, 4);
    if (len)
        crc = crc32(crc, data, len);
    uint8_t tail[4];
    put32(tail, crc);
    return fwrite(hdr, 1, 8, fp) == 8 && (len == 0 || fwrite(data, 1, len, fp) == len) && fwrite(tail, 1, 4, fp) == 4;
}

/* 8-bit RGBA PNG, filter 0 on every scanline, one IDAT */
static int write_png(const char *path, const uint8_t *rgba, int w, int h) {
    const size_t stride = (size_t)w * 4 + 1;
    uint8_t *raw = (uint8_t *)malloc(stride * h);
    if (!raw)
        return 0;
    for (int y = 0; y < h; y++) {
        raw[y * stride] = 0;
        memcpy(raw + y * stride + 1, rgba + (size_t)y * w * 4, (size_t)w * 4);
    }
    uLongf zlen = compressBound(stride * h);
    uint8_t *z = (uint8_t *)malloc(zlen);
    int ok = z && compress2(z, &zlen, raw, stride * h, 6) == Z_OK;
    FILE *fp = ok ? fopen(path, "wb") : NULL;
    if (fp) {
        static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
        uint8_t ihdr[13];
        put32(ihdr, (uint32_t)w), put32(ihdr + 4, (uint32_t)h);
        ihdr[8] = 8, ihdr[9] = 6, ihdr[10] = 0, ihdr[11] = 0, ihdr[12] = 0;
        ok = fwrite(sig, 1, 8, fp) == 8 && chunk(fp, "IHDR", ihdr, 13) && chunk(fp, "IDAT", z, (uint32_t)zlen) && chunk(fp, "IEND", NULL, 0);
        ok = (fclose(fp) == 0) && ok;
    } else
        ok = 0;
    free(raw);
    free(z);
    return ok;
}

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "Input file not specified. Usage: %s <path to xml> [out.png] [device]\n", argv[0]);
        return 1;
    }
    const char *path = argv[1], *out = argc > 2 ? argv[2] : "render.png";
    const int device = argc > 3 ? atoi(argv[3]) : 0;
    const size_t n = strlen(path);
    cpt_b200_scene *scene = NULL;
    cpt_b200_renderer *r = NULL;
    printf("[SCENE] Loading scenes from '%s'\n", path);
    const int is_xml = n > 4 && strcmp(path + n - 4, ".xml") == 0;
    if ((is_xml ? cpt_b200_scene_load_xml(path, &scene) : cpt_b200_scene_load(path, &scene)) != 0) {
        fprintf(stderr, "[SCENE] %s\n", cpt_b200_last_error());
        return 2;
    }
    int32_t cfg[CPTS_CFG_COUNT];
    cpt_b200_scene_config(scene, cfg);
    if (cpt_b200_create(scene, device, 0, &r) != 0) {
        fprintf(stderr, "[RENDERER] %s\n", cpt_b200_last_error());
        cpt_b200_scene_free(scene);
        return 3;
    }
    char info[1024];
    cpt_b200_info(r, info, sizeof(info));
    printf("[RENDERER] Path tracer loaded: %s\n", info);
    printf("[RENDERER] Prepare to render the scene... [%d] bounces, [%d] SPP\n", cfg[CPTS_CFG_MAX_DEPTH], cfg[CPTS_CFG_SPP]);
    const int w = cfg[CPTS_CFG_WIDTH], h = cfg[CPTS_CFG_HEIGHT];
    uint8_t *rgba = (uint8_t *)malloc((size_t)w * h * 4);
    int rc = rgba ? cpt_b200_render_offline(r, cfg[CPTS_CFG_SPP], cfg[CPTS_CFG_GAMMA], rgba) : -1;
    if (rc != 0)
        fprintf(stderr, "[RENDERER] %s\n", rgba ? cpt_b200_last_error() : "out of memory");
    else if (!write_png(out, rgba, w, h)) {
        fprintf(stderr, "[IMAGE] failed to write '%s'\n", out);
        rc = 4;
    } else {
        uint64_t st[8];
        cpt_b200_stats(r, st);
        printf("[IMAGE] Image saved to `%s` (%llu samples, %llu rays, %.1f ms on the device)\n", out, (unsigned long long)st[0],
               (unsigned long long)(st[0] + st[1] + st[2]), (double)st[5] / 1000.0);
    }
    free(rgba);
    cpt_b200_destroy(r);
    cpt_b200_scene_free(scene);
    return rc ? 4 : 0;
}
