/* TEST INFRASTRUCTURE ONLY — known-answer self test of the oracle.
 * (1) the 13 rstest cases the reference holds for this path: /root/reference/src/breakout-game/src/mechanics.rs
 *     :659-675 (left wall, exact), :677-693 (right wall, exact), :708-752 (rectangle; normal +-0.01, way +-0.1,
 *     0 <= approximation < 0.8);
 * (2) Random123 philox4x32-10 known answers; (3) brick layout; (4) dir_x range. Exit code 0 = all pass. */
#include "breakout_oracle.h"
#include "philox.h"
#include <math.h>
#include <stdio.h>

static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { ++fails; printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } while (0)

static void wall_case(int right, float cx, float cy, float r, float mx, float my, int some, float way, float nx) {
    orc_circle b = {{cx, cy}, r}; orc_v2 mv = {mx, my}; orc_contact_surface c; uint32_t err = 0;
    int got = right ? orc_collision_test_right_wall(&b, mv, &c, &err) : orc_collision_test_left_wall(&b, mv, &c, &err);
    CHECK(got == some, "wall some %d != %d", got, some);
    if (got && some) CHECK(c.way == way && c.approximation == 0.0f && c.surface_normal.x == nx && c.surface_normal.y == 0.0f,
                           "wall surface way=%g nx=%g", c.way, c.surface_normal.x);
    CHECK(err == 0, "err flags %u", err);
}
static void rect_case(float mx, float my, float x0, float y0, float x1, float y1, int some, float way, float nx, float ny) {
    orc_circle b = {{100.0f, 100.0f}, 5.0f}; orc_v2 mv = {mx, my}; orc_aabb a = {{x0, y0}, {x1, y1}};
    orc_contact_surface c; uint32_t err = 0;
    int got = orc_collision_check_with_rectangle(&b, mv, &a, &c, &err);
    CHECK(got == some, "rect(%g,%g) some %d != %d", mx, my, got, some);
    if (got && some) {
        CHECK(fabsf(c.surface_normal.x - nx) <= 0.01f, "normal.x %g vs %g", c.surface_normal.x, nx);
        CHECK(fabsf(c.surface_normal.y - ny) <= 0.01f, "normal.y %g vs %g", c.surface_normal.y, ny);
        CHECK(fabsf(c.way - way) <= 0.1f, "way %g vs %g", c.way, way);
        CHECK(c.approximation >= 0.0f && c.approximation < 0.8f, "approximation %g", c.approximation);
    }
    CHECK((err & ~ORC_ERR_DEGENERATE) == 0, "err flags %u", err);
}

int main(void) {
    /* mechanics.rs:660-662 */
    wall_case(0, 10.0f, 10.0f, 5.0f, -2.0f, 2.0f, 0, 0, 0);
    wall_case(0, 5.0f, 10.0f, 5.0f, -5.0f, 0.0f, 1, 0.0f, 1.0f);
    wall_case(0, 7.0f, 7.0f, 5.0f, -5.0f, 0.0f, 1, 2.0f, 1.0f);
    /* mechanics.rs:678-680 */
    wall_case(1, 600.0f - 10.0f, 10.0f, 5.0f, 2.0f, 2.0f, 0, 0, 0);
    wall_case(1, 600.0f - 5.0f, 10.0f, 5.0f, 5.0f, 0.0f, 1, 0.0f, -1.0f);
    wall_case(1, 600.0f - 7.0f, 7.0f, 5.0f, 5.0f, 0.0f, 1, 2.0f, -1.0f);
    /* mechanics.rs:709-722 */
    const float s = 0.70710678f;
    rect_case(10.0f, 0.0f, 150.0f, 90.0f, 170.0f, 110.0f, 0, 0, 0, 0);
    rect_case(5.0f, 0.0f, 110.0f, 90.0f, 130.0f, 110.0f, 1, 5.0f, -1.0f, 0.0f);
    rect_case(3.0f, -3.0f, 100.0f, 70.0f, 120.0f, 93.0f, 1, 2.83f, 0.0f, 1.0f);
    rect_case(-8.0f, -8.0f, 70.0f, 80.0f, 90.0f, 100.0f, 1, 7.07f, 1.0f, 0.0f);
    rect_case(-1.46f, -1.46f, 80.0f, 80.0f, 95.0f, 95.0f, 1, 2.07f, s, s);
    rect_case(-5.0f, -5.0f, 80.0f, 80.0f, 95.0f, 95.0f, 1, 2.07f, s, s);
    rect_case(-4.2f, -4.2f, 80.0f, 80.0f, 90.0f, 90.0f, 0, 0, 0, 0);

    /* Random123 kat_vectors: philox4x32 10 */
    {
        uint32_t out[4];
        uint32_t c0[4] = {0, 0, 0, 0}, k0[2] = {0, 0};
        orc_philox4x32_10(c0, k0, out);
        CHECK(out[0] == 0x6627e8d5u && out[1] == 0xe169c58du && out[2] == 0xbc57ac4cu && out[3] == 0x9b00dbd8u, "philox kat 0: %08x %08x %08x %08x", out[0], out[1], out[2], out[3]);
        uint32_t c1[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}, k1[2] = {0xffffffffu, 0xffffffffu};
        orc_philox4x32_10(c1, k1, out);
        CHECK(out[0] == 0x408f276du && out[1] == 0x41c83b0eu && out[2] == 0xa20bc7c6u && out[3] == 0x6d5451fdu, "philox kat 1: %08x %08x %08x %08x", out[0], out[1], out[2], out[3]);
        uint32_t c2[4] = {0x243f6a88u, 0x85a308d3u, 0x13198a2eu, 0x03707344u}, k2[2] = {0xa4093822u, 0x299f31d0u};
        orc_philox4x32_10(c2, k2, out);
        CHECK(out[0] == 0xd16cfe09u && out[1] == 0x94fdccebu && out[2] == 0x5001e420u && out[3] == 0x24126ea1u, "philox kat 2: %08x %08x %08x %08x", out[0], out[1], out[2], out[3]);
    }
    /* brick layout: 60 bricks, 3 rows x 20, x = 30+27k, rows [35,60] [62,87] [89,114] (mechanics.rs:67-95) */
    {
        orc_mechanics m; orc_mechanics_new(&m, -0.25f);
        CHECK(m.n_bricks == 60, "n_bricks %d", m.n_bricks);
        for (int i = 0; i < m.n_bricks; ++i) {
            int r = i / 20, k = i % 20;
            CHECK(m.bricks[i].min.x == 30.0f + 27.0f * k && m.bricks[i].max.x == 55.0f + 27.0f * k &&
                  m.bricks[i].min.y == 35.0f + 27.0f * r && m.bricks[i].max.y == 60.0f + 27.0f * r && m.brick_id[i] == i, "brick %d", i);
        }
        CHECK(orc_mechanics_brick_mask(&m) == 0x0FFFFFFFFFFFFFFFull, "mask");
        CHECK(orc_env_goal_mean() == 59.0f, "goal");
    }
    /* dir_x in [-0.35, -0.15) for extreme entropy */
    CHECK(orc_dir_x_from_bits(0u) == -0.35f, "dir_x low");
    CHECK(orc_dir_x_from_bits(0xFFFFFFFFu) < -0.15f && orc_dir_x_from_bits(0xFFFFFFFFu) > -0.1501f, "dir_x high %g", orc_dir_x_from_bits(0xFFFFFFFFu));
    printf(fails ? "oracle selftest: %d FAILED\n" : "oracle selftest: all passed\n", fails);
    return fails ? 1 : 0;
}
