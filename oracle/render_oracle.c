/* TEST INFRASTRUCTURE ONLY — see breakout_oracle.h. Observation half of the oracle:
 * rasteriser (OUR spec — the reference's BreakoutDrawer::draw is unimplemented!(),
 * /root/reference/src/breakout-game/src/breakout_drawer.rs:22-28; geometry and colours are taken from the egui
 * shape drawer /root/reference/src/breakout-game/src/app_game_drawer.rs:21-88), grayscale (image 0.24.9
 * imageops::grayscale, call site /root/reference/src/_breakout-ml/src/breakout_environment.rs:193), the 4-slot
 * frame ring (/root/reference/src/_breakout-ml/src/util/frame_ring_buffer.rs) and BreakoutEnvironment
 * (/root/reference/src/_breakout-ml/src/breakout_environment.rs:131-207). No reference test covers any of this:
 * PARITY UNPINNED; pixels are "bit-exact against this definition".
 *
 * Raster spec (DESIGN.md "Raster spec"):
 *   - frame 84x84, background black; scaled coordinate = pos * 84 / 600 in f32 (multiply, then divide:
 *     app_game_drawer.rs:21-36);
 *   - a pixel (i, j) is covered by a shape iff its centre (i+0.5, j+0.5) is;
 *   - bricks: filled rect [min_s, max_s) DARK_GRAY (96,96,96)        (app_game_drawer.rs:78-88)
 *   - ball: stroked circle, radius r_s = 10*84/600, stroke width 2.0 px (NOT scaled), YELLOW (255,255,0):
 *       covered iff (r_s-1)^2 <= d2 <= (r_s+1)^2, d2 = dx*dx + dy*dy in f32  (app_game_drawer.rs:50-61)
 *   - paddle: filled rect [min_s, max_s) WHITE                        (app_game_drawer.rs:65-76)
 *   - draw order bricks, ball, paddle; later shapes overwrite          (app_game_drawer.rs:38-44)
 *   - no anti-aliasing.
 */
#include "breakout_oracle.h"
#include <string.h>

static inline float scale84(float pos) { return pos * 84.0f / 600.0f; }

/* conservative integer bounds around a scaled interval (only to skip pixels that cannot be covered; the
 * coverage decision itself is always the f32 centre test below) */
static inline int lo_bound(float v) { int i = (int)v - 2; return i < 0 ? 0 : i; }
static inline int hi_bound(float v, int n) { int i = (int)v + 3; return i > n ? n : (i < 0 ? 0 : i); }

static void fill_rect(uint8_t* rgb, orc_v2 mn, orc_v2 mx, uint8_t r, uint8_t g, uint8_t b) {
    float x0 = scale84(mn.x), y0 = scale84(mn.y), x1 = scale84(mx.x), y1 = scale84(mx.y);
    for (int j = lo_bound(y0); j < hi_bound(y1, ORC_FRAME_H); ++j) {
        float py = (float)j + 0.5f;
        if (!(py >= y0 && py < y1)) continue;
        for (int i = lo_bound(x0); i < hi_bound(x1, ORC_FRAME_W); ++i) {
            float px = (float)i + 0.5f;
            if (px >= x0 && px < x1) { uint8_t* p = rgb + 3 * (j * ORC_FRAME_W + i); p[0] = r; p[1] = g; p[2] = b; }
        }
    }
}

void orc_draw_rgb(const orc_mechanics* m, uint8_t* rgb) {
    memset(rgb, 0, 3 * ORC_FRAME_BYTES);
    for (int k = 0; k < m->n_bricks; ++k) fill_rect(rgb, m->bricks[k].min, m->bricks[k].max, 96, 96, 96);
    {
        float bx = scale84(m->ball_shape.center.x), by = scale84(m->ball_shape.center.y);
        float rs = scale84(m->ball_shape.radius);
        float r_out = rs + 1.0f, r_in = rs - 1.0f;
        float out2 = r_out * r_out, in2 = r_in * r_in;
        for (int j = lo_bound(by - r_out); j < hi_bound(by + r_out, ORC_FRAME_H); ++j) {
            float dy = ((float)j + 0.5f) - by;
            for (int i = lo_bound(bx - r_out); i < hi_bound(bx + r_out, ORC_FRAME_W); ++i) {
                float dx = ((float)i + 0.5f) - bx;
                float d2 = dx * dx + dy * dy;
                if (d2 <= out2 && d2 >= in2) { uint8_t* p = rgb + 3 * (j * ORC_FRAME_W + i); p[0] = 255; p[1] = 255; p[2] = 0; }
            }
        }
    }
    fill_rect(rgb, m->panel_shape.min, m->panel_shape.max, 255, 255, 255);
}

/* image 0.24.9 color.rs rgb_to_luma for u8: integer Rec.709, (2126 R + 7152 G + 722 B) / 10000 [recalled]. */
void orc_grayscale(const uint8_t* rgb, uint8_t* luma) {
    for (int p = 0; p < ORC_FRAME_BYTES; ++p) {
        uint32_t l = (2126u * rgb[3 * p] + 7152u * rgb[3 * p + 1] + 722u * rgb[3 * p + 2]) / 10000u;
        luma[p] = (uint8_t)l;
    }
}

void orc_frame_ring_new(orc_frame_ring* r) { memset(r, 0, sizeof *r); }
void orc_frame_ring_add(orc_frame_ring* r, const uint8_t* frame) {
    memcpy(r->buffer[r->next_slot], frame, ORC_FRAME_BYTES);
    r->next_slot = (r->next_slot + 1 == ORC_NUM_FRAMES) ? 0 : r->next_slot + 1;
}

void orc_env_reset_with(orc_env* e, float dir_x) {
    e->sticky_err |= e->state.mechanics.err;
    orc_mechanics_new(&e->state.mechanics, dir_x);
    orc_frame_ring_new(&e->state.frame_buffer);
    e->episode_step = 0;
    e->episode_return = 0.0f;
}

void orc_env_step(orc_env* e, int action, float* reward, int* done) {
    uint8_t rgb[3 * ORC_FRAME_BYTES];
    uint8_t luma[ORC_FRAME_BYTES];
    uint32_t prev_score = e->state.mechanics.score;
    orc_mechanics_time_step(&e->state.mechanics, action);   /* None/Left/Right map 1:1 (:155-161) */
    orc_draw_rgb(&e->state.mechanics, rgb);
    orc_grayscale(rgb, luma);
    orc_frame_ring_add(&e->state.frame_buffer, luma);
    *reward = (float)(e->state.mechanics.score - prev_score);
    *done = e->state.mechanics.finished;
    e->episode_step += 1;
    e->episode_return += *reward;
}

float orc_env_goal_mean(void) {
    orc_mechanics m; orc_mechanics_new(&m, -0.25f);
    return (float)(m.n_bricks - 1);
}

/* BreakoutState::to_multi_dim_array: tensor[x][y][hist] = frame[hist].get_pixel(x, y) as f32 */
void orc_state_to_f32_xyh(const orc_state* s, float* out) {
    for (int hist = 0; hist < ORC_NUM_FRAMES; ++hist)
        for (int y = 0; y < ORC_FRAME_H; ++y)
            for (int x = 0; x < ORC_FRAME_W; ++x)
                out[(x * ORC_FRAME_H + y) * ORC_NUM_FRAMES + hist] = (float)s->frame_buffer.buffer[hist][y * ORC_FRAME_W + x];
}
