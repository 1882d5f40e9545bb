/*
 * oracle.h — C API of the CPU parity oracle (liboracle.so).
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement, in f64, of the
 * reference renderer's hot path (racer-tracer `Renderer::render` ->
 * `CpuRenderer::raytrace` -> `ray_color`).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product
 * (racer_tracer_b200/) never does.
 *
 * PARITY UNPINNED: the reference holds no golden vectors, known-answer tests
 * or fixtures for the render path (its whole test suite is four Vec3
 * arithmetic tests, src/vec3.rs:446-503, which tests/test_oracle_kat.py
 * replays), and it cannot be compiled here (no Rust toolchain).  The oracle is
 * pinned instead by hand-derived known-answer vectors written from the cited
 * lines, by the published Philox4x32-10 test vectors, and statistically by
 * the only outputs of the reference that exist: its published renders
 * (assets/*.png, README.md:26-49), which the oracle reproduces to 46 / 42 /
 * 34 dB on 60x60 block means (three_balls / clown / cornell_box at the
 * reference's 200 spp, colour means equal to three decimals) —
 * tests/test_reference_renders.py.  Bit-level parity stays unpinned.
 *
 * It consumes the same flat structs as the product (include/racer_cuda.h) so
 * both sides trace exactly the same scene, tree and camera.
 */
#ifndef RACER_ORACLE_H
#define RACER_ORACLE_H

#include "../include/racer_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* RNG back ends.  ORACLE_RNG_PHILOX consumes exactly the counter-based
 * streams the GPU consumes (DESIGN.md "RNG streams"); ORACLE_RNG_SEQUENTIAL is
 * an independent xoshiro256** stream per tile drawn in the reference's call
 * order (src/util.rs:9-23 call sites), for independent-stream checks. */
enum { ORACLE_RNG_PHILOX = 0, ORACLE_RNG_SEQUENTIAL = 1 };

typedef struct oracle_counters {
    uint64_t samples;
    uint64_t segments;          /* scene.hit calls from ray_color            */
    uint64_t node_tests;        /* Aabb::hit calls                           */
    uint64_t prim_tests[4];     /* obj_hit calls by RC_PRIM_*                */
    uint64_t prim_hits[4];      /* accepted hits by RC_PRIM_*                */
    uint64_t scatters[4];       /* Material::scatter calls by RC_MAT_*       */
    uint64_t tex_evals[4];      /* Texture::value calls by RC_TEX_*          */
    uint64_t background;        /* BackgroundColor::color calls              */
    uint64_t depth_exhausted;   /* paths that returned white at depth 0      */
    uint64_t rejection_iters;   /* iterations of the rejection samplers      */
    uint64_t noise_calls;       /* Perlin::noise calls                       */
} oracle_counters;

typedef struct oracle_options {
    int32_t rng;            /* ORACLE_RNG_*                                  */
    int32_t threads;        /* 0 = hardware concurrency                      */
    int32_t tiles_w;        /* config.render.num_threads_width  (10)         */
    int32_t tiles_h;        /* config.render.num_threads_height (10)         */
    int32_t sample_begin;   /* trace samples [sample_begin, +sample_count)   */
    int32_t sample_count;   /* 0 = params.samples                            */
    int32_t linear_sum;     /* 1: out_rgb = raw radiance sums (no scale_sqrt)*/
    int32_t counterfactual; /* ORACLE_CF_* bits; 0 = the reference's behaviour   */
} oracle_options;

/* Deliberate departures from the reference, for tests that must show a check can FAIL.
 * ORACLE_CF_PER_SAMPLE_U: the horizontal jitter is drawn per sample instead of once per pixel
 * (the opposite of quirk Q1, src/renderer/cpu.rs:35-36); sequential RNG only. */
enum { ORACLE_CF_PER_SAMPLE_U = 1 };

/* CpuRenderer::render for one image (src/renderer/cpu.rs:26-131).
 * out_rgb: width*height*3 doubles, row 0 = top.  counters may be NULL. */
int oracle_render(const rc_scene* scene, const rc_camera* camera, const rc_params* params,
                  const oracle_options* opt, double* out_rgb, oracle_counters* counters);

/* CpuRendererScaled::render for one image (src/renderer/cpu_scaled.rs:37-121): params->width/height =
 * screen, params->samples/max_depth = config.preview; one colour per scale_w x scale_h block, upscaled;
 * pixels beyond a tile's last whole block stay 0.  Philox streams are keyed by the block's index in the
 * (width / scale_w)-wide block grid. */
int oracle_render_preview(const rc_scene* scene, const rc_camera* camera, const rc_params* params,
                          const oracle_options* opt, int32_t scale_w, int32_t scale_h, double* out_rgb);

/* Primary-visibility AOV with the fixed jitter (pixel centre, lens centre):
 * RayImageData of src/renderer.rs:33-39,58-88. */
int oracle_primary_aov(const rc_scene* scene, const rc_camera* camera, const rc_params* params,
                       uint32_t* id, double* t, double* normal, double* point);

/* Radiance of single samples, Philox streams: out[3*i..] = ray_color of
 * sample `sample_idx[i]` of pixel `pixel_idx[i]` (un-accumulated). */
int oracle_sample_radiance(const rc_scene* scene, const rc_camera* camera, const rc_params* params,
                           int32_t n, const int32_t* pixel_idx, const int32_t* sample_idx,
                           double* out, int32_t* out_segments);

/* Camera::new derivation, src/camera.rs:196-234. */
void oracle_camera_new(const double look_from[3], const double look_at[3], const double scene_up[3],
                       double vfov, double aperture, double focus_distance, double aspect_ratio,
                       double time_a, double time_b, rc_camera* out);

/* ToneMap::tone_map per pixel (the files under src/tone_map/) and the RGBA packer of
 * src/image_action/png.rs:21-31. */
void oracle_tone_map(const rc_tone_map* tm, const double* rgb, int64_t n_pixels, double* out);
void oracle_quantise_rgba(const double* rgb, int64_t n_pixels, uint8_t* rgba);

/* ---- unit-level entry points for known-answer tests --------------------- */
void oracle_philox4x32(const uint32_t ctr[4], const uint32_t key[2], int32_t rounds, uint32_t out[4]);
void oracle_philox2x32(const uint32_t ctr[2], uint32_t key, int32_t rounds, uint32_t out[2]);
/* Aabb::hit, src/aabb.rs:42-59 */
int oracle_aabb_hit(const double bmin[3], const double bmax[3], const double origin[3],
                    const double dir[3], double t_min, double t_max);
/* obj_hit of one primitive of the scene; returns 1 on hit.
 * out = {t, px,py,pz, nx,ny,nz, u, v, front_face} */
int oracle_prim_hit(const rc_scene* scene, int32_t prim, const double origin[3], const double dir[3],
                    double t_min, double t_max, double out[10]);
/* closest hit through the BVH (or the linear list); returns prim index or -1 */
int oracle_scene_hit(const rc_scene* scene, const double origin[3], const double dir[3],
                     double t_min, double t_max, double out[10]);
double oracle_reflectance(double cosine, double refraction_index);      /* dialectric.rs:17-22 */
void oracle_reflect(const double v[3], const double n[3], double out[3]);  /* vec3.rs:412-414 */
void oracle_refract(const double uv[3], const double n[3], double etai_over_etat, double out[3]); /* vec3.rs:416-422 */
void oracle_texture_value(const rc_scene* scene, int32_t texture, double u, double v,
                          const double point[3], double out[3]);
double oracle_perlin_noise(const rc_perlin* p, const double point[3]);   /* noise.rs:57-96 */
double oracle_perlin_turbulence(const rc_perlin* p, const double point[3], int32_t depth); /* noise.rs:98-109 */
void oracle_background(const rc_scene* scene, const double dir[3], double out[3]);
/* Camera::get_ray with explicit lens sample (dx,dy in the unit disk) */
void oracle_get_ray(const rc_camera* cam, double u, double v, double dx, double dy,
                    double origin[3], double dir[3]);
/* Vec3 arithmetic replayed by the reference's own tests, src/vec3.rs:446-503:
 * op 0 add, 1 sub, 2 mul (element-wise), 3 div by scalar b[0]. */
void oracle_vec3_op(int32_t op, const double a[3], const double b[3], double out[3]);

#ifdef __cplusplus
}
#endif
#endif
