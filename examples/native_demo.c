/* Shapes to pixels through libvello_b200.so alone (no Rust, no Python): builds a small scene with the native front end
 * (include/vello_b200_scene.h), renders it on the GPU (include/vello_b200.h) and writes a binary PPM.
 *
 *   gcc -O2 -Iinclude examples/native_demo.c -Lvello_b200 -lvello_b200 -Wl,-rpath,$PWD/vello_b200 -lm -o native_demo
 *   ./native_demo out.ppm
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vello_b200_scene.h"

static const double IDENTITY[6] = {1, 0, 0, 1, 0, 0};

int main(int argc, char **argv) {
    const uint32_t W = 512, H = 384;
    vb_scene *scene = vb_scene_new();
    if (!scene) return 1;

    /* background: linear gradient over the whole frame */
    const uint8_t rect_verbs[] = {'M', 'L', 'L', 'L', 'Z'};
    const double rect_coords[] = {0, 0, 512, 0, 512, 384, 0, 384};
    const vb_path frame = {rect_verbs, 5, rect_coords};
    const vb_color_stop stops[] = {{0.0f, {0.05f, 0.10f, 0.30f, 1.0f}}, {1.0f, {0.80f, 0.35f, 0.10f, 1.0f}}};
    vb_brush sky;
    memset(&sky, 0, sizeof sky);
    sky.kind = VB_BRUSH_LINEAR;
    sky.geom[0] = 0; sky.geom[1] = 0; sky.geom[2] = 0; sky.geom[3] = 384;
    sky.stops = stops; sky.n_stops = 2; sky.premul_interp = 1;
    if (vb_scene_fill(scene, VB_FILL_NON_ZERO, IDENTITY, &sky, NULL, &frame)) return 2;

    /* a translucent blob (cubic Beziers) and a stroked zig-zag on top */
    const uint8_t blob_verbs[] = {'M', 'C', 'C', 'C', 'Z'};
    const double blob_coords[] = {120, 200, 120, 80, 300, 60, 340, 180, 380, 300, 260, 340, 200, 300, 160, 275, 120, 260, 120, 200};
    const vb_path blob = {blob_verbs, 5, blob_coords};
    vb_brush teal;
    memset(&teal, 0, sizeof teal);
    teal.kind = VB_BRUSH_SOLID;
    teal.color.r = 0.1f; teal.color.g = 0.8f; teal.color.b = 0.7f; teal.color.a = 0.6f;
    if (vb_scene_fill(scene, VB_FILL_NON_ZERO, IDENTITY, &teal, NULL, &blob)) return 2;
    const uint8_t zig_verbs[] = {'M', 'L', 'L', 'L', 'L'};
    const double zig_coords[] = {40, 340, 140, 300, 240, 350, 340, 290, 470, 345};
    const vb_path zig = {zig_verbs, 5, zig_coords};
    const vb_stroke pen = {9.0, VB_JOIN_ROUND, VB_CAP_ROUND, VB_CAP_ROUND, 4.0, NULL, 0, 0.0};
    vb_brush white;
    memset(&white, 0, sizeof white);
    white.kind = VB_BRUSH_SOLID;
    white.color.r = white.color.g = white.color.b = white.color.a = 1.0f;
    if (vb_scene_stroke(scene, &pen, IDENTITY, &white, NULL, &zig)) return 2;

    /* shapes through the path builder (kurbo's Shape::path_elements, tolerance 0.1 like vello's fills) */
    vb_pathbuf *pb = vb_pathbuf_new();
    vb_pathbuf_circle(pb, 410.0, 90.0, 42.0, 0.1);
    vb_brush sun = white;
    sun.color.r = 1.0f; sun.color.g = 0.85f; sun.color.b = 0.3f; sun.color.a = 0.9f;
    vb_path shape = vb_pathbuf_view(pb);
    if (vb_scene_fill(scene, VB_FILL_NON_ZERO, IDENTITY, &sun, NULL, &shape)) return 2;
    vb_pathbuf_clear(pb);
    vb_pathbuf_rounded_rect(pb, 30.0, 30.0, 190.0, 110.0, 18.0, 0.1);
    shape = vb_pathbuf_view(pb);
    static const double dashes[2] = {12.0, 6.0};
    const vb_stroke thin = {3.0, VB_JOIN_MITER, VB_CAP_BUTT, VB_CAP_BUTT, 4.0, dashes, 2, 0.0}; /* dashed: cut on the CPU like vello (kurbo::dash) */
    if (vb_scene_stroke(scene, &thin, IDENTITY, &white, NULL, &shape)) return 2;
    vb_pathbuf_free(pb);

    vb_renderer *r = NULL;
    vb_options opt;
    memset(&opt, 0, sizeof opt);
    opt.max_retries = 6;
    int rc = vb_renderer_new(&opt, &r);
    if (rc) { fprintf(stderr, "vb_renderer_new: %s\n", vb_strerror(rc)); return 3; } /* no GPU: fails loudly, there is no fallback */
    vb_params params;
    memset(&params, 0, sizeof params);
    params.base_color = 0xff000000u; /* opaque black, premultiplied RGBA8, r in the low byte */
    params.width = W; params.height = H; params.aa = 2; /* MSAA16 */
    uint8_t *pixels = malloc((size_t)W * H * 4);
    vb_frame_stats stats;
    rc = vb_render_scene(r, scene, &params, pixels, /*out_is_device=*/0, &stats);
    if (rc) { fprintf(stderr, "vb_render_scene: %s (%s)\n", vb_strerror(rc), vb_last_error(r)); return 4; }
    FILE *f = fopen(argc > 1 ? argv[1] : "native_demo.ppm", "wb");
    if (!f) return 5;
    fprintf(f, "P6\n%u %u\n255\n", W, H);
    for (size_t i = 0; i < (size_t)W * H; i++) fwrite(pixels + 4 * i, 1, 3, f);
    fclose(f);
    free(pixels);
    vb_renderer_free(r);
    vb_scene_free(scene);
    return 0;
}
