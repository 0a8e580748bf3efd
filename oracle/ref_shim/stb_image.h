/*
 * stb_image.h -- SHIM, test infrastructure only (oracle/_ref build).
 *
 * The reference includes "stb_image.h" (ray_voxel.cpp:28-29) but does not vendor it and
 * the header is not present in this image.  The reference only uses
 *   stbi_load(path, &w, &h, &channels, 1)   (ray_voxel.cpp:195)
 *   stbi_image_free(data)                   (ray_voxel.cpp:223)
 * so this shim provides exactly those two entry points for binary PNM files:
 *   P5 (gray, maxval <= 255): bytes returned untouched, as real stb_image does;
 *   P6 (RGB,  maxval <= 255): converted with stb_image's published integer luma
 *       y = (77 r + 150 g + 29 b) >> 8.
 * Anything else fails to load (returns NULL), which exercises the reference's
 * skip-on-load-error path (ray_voxel.cpp:470-473).  Written from scratch.
 */
#ifndef PVP_ORACLE_STB_IMAGE_SHIM_H
#define PVP_ORACLE_STB_IMAGE_SHIM_H

#include <stdio.h>
#include <stdlib.h>

typedef unsigned char stbi_uc;

static int pvp_shim_pnm_int(FILE *f, int *out)
{
    int c = fgetc(f);
    for (;;) {
        while (c == ' ' || c == '\t' || c == '\n' || c == '\r') c = fgetc(f);
        if (c == '#') { while (c != '\n' && c != EOF) c = fgetc(f); continue; }
        break;
    }
    if (c < '0' || c > '9') return 0;
    int v = 0;
    while (c >= '0' && c <= '9') { v = v * 10 + (c - '0'); c = fgetc(f); }
    *out = v; /* the single whitespace byte after the token has been consumed */
    return 1;
}

static stbi_uc *stbi_load(const char *filename, int *x, int *y, int *channels_in_file, int desired_channels)
{
    (void)desired_channels; /* the reference always asks for 1 */
    FILE *f = fopen(filename, "rb");
    if (!f) return NULL;
    int m0 = fgetc(f), m1 = fgetc(f);
    int w = 0, h = 0, maxv = 0;
    if (m0 != 'P' || (m1 != '5' && m1 != '6') || !pvp_shim_pnm_int(f, &w) || !pvp_shim_pnm_int(f, &h) ||
        !pvp_shim_pnm_int(f, &maxv) || w <= 0 || h <= 0 || maxv <= 0 || maxv > 255) {
        fclose(f);
        return NULL;
    }
    int comp = (m1 == '5') ? 1 : 3;
    size_t npx = (size_t)w * (size_t)h;
    stbi_uc *raw = (stbi_uc *)malloc(npx * (size_t)comp);
    if (!raw || fread(raw, (size_t)comp, npx, f) != npx) { free(raw); fclose(f); return NULL; }
    fclose(f);
    if (comp == 3) {
        stbi_uc *g = (stbi_uc *)malloc(npx);
        if (!g) { free(raw); return NULL; }
        for (size_t i = 0; i < npx; ++i)
            g[i] = (stbi_uc)((raw[3 * i] * 77 + raw[3 * i + 1] * 150 + raw[3 * i + 2] * 29) >> 8);
        free(raw);
        raw = g;
    }
    *x = w; *y = h; *channels_in_file = comp;
    return raw;
}

static void stbi_image_free(void *p) { free(p); }

#endif
