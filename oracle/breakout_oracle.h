/* TEST INFRASTRUCTURE ONLY — CPU oracle for the Breakout env + replay hot path.
 *
 * A plain-C restatement of the reference's algorithm (bitmagier/q-learning). The Rust reference cannot be
 * compiled here (no cargo/rustc; renderer is unimplemented!(); the Breakout env crate is dead code), so this
 * is a "port"-kind oracle. PARITY PINNING: the only known-answer vectors the reference holds for this path are
 * the 13 rstest cases at src/breakout-game/src/mechanics.rs:659-722; they are checked in oracle/selftest.c and
 * tests/test_oracle.py. Full-trajectory behaviour, the rasteriser, the frame ring, the replay buffer and the
 * sampler have NO reference tests => "parity unpinned" for those.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use this code.
 * The product (q-learning_b200/) never includes, links or calls anything in oracle/.
 */
#ifndef QLC_BREAKOUT_ORACLE_H
#define QLC_BREAKOUT_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y; } orc_v2;
typedef struct { orc_v2 min, max; } orc_aabb;                 /* algebra_2d.rs:11-14 */
typedef struct { orc_v2 center; float radius; } orc_circle;   /* algebra_2d.rs:32-36 */
typedef struct { float way, approximation; orc_v2 surface_normal; } orc_contact_surface; /* algebra_2d.rs:38-44 */
typedef struct { int some; float dist; orc_v2 normal1, normal2; } orc_contact;           /* parry2d Contact */

#define ORC_MAX_BRICKS 60

/* sticky error flags: the reference panics (assert!) — we record and keep going. */
#define ORC_ERR_WALL_DISTANCE 1u   /* mechanics.rs:265,284,303 */
#define ORC_ERR_APPROX_RANGE  2u   /* mechanics.rs:511 */
#define ORC_ERR_RECURSION     4u   /* proceed_ball_with recursion deeper than ORC_MAX_RECURSION */
#define ORC_ERR_BISECTION     8u   /* binary_search_first_contact deeper than ORC_MAX_BISECTION */
#define ORC_ERR_DEGENERATE   16u   /* parry2d degenerate contact branch (centre on the box boundary) */
#define ORC_MAX_RECURSION 16
#define ORC_MAX_BISECTION 40

/* mechanics.rs:46-54 (AoS, Vec<Brick> with positional removal) */
typedef struct {
    orc_aabb bricks[ORC_MAX_BRICKS];
    uint8_t  brick_id[ORC_MAX_BRICKS];  /* instrumentation: original index 20*row+k of each live brick */
    int      n_bricks;
    orc_circle ball_shape;
    orc_v2   ball_direction;
    float    ball_speed_per_sec;
    orc_aabb panel_shape;
    float    panel_speed_per_sec;
    int      finished;
    uint32_t score;
    uint32_t err;                       /* sticky ORC_ERR_* */
} orc_mechanics;

enum { ORC_CONTROL_NONE = 0, ORC_CONTROL_LEFT = 1, ORC_CONTROL_RIGHT = 2 }; /* breakout_environment.rs:104-120 */

void orc_mechanics_new(orc_mechanics* m, float dir_x);                      /* mechanics.rs:57-116 */
void orc_mechanics_time_step(orc_mechanics* m, int control);                /* mechanics.rs:119-129 */
uint64_t orc_mechanics_brick_mask(const orc_mechanics* m);

/* exposed for the known-answer tests (mechanics.rs:659-752) */
int orc_collision_test_left_wall(const orc_circle* ball, orc_v2 mv, orc_contact_surface* out, uint32_t* err);
int orc_collision_test_right_wall(const orc_circle* ball, orc_v2 mv, orc_contact_surface* out, uint32_t* err);
int orc_collision_test_top_wall(const orc_circle* ball, orc_v2 mv, orc_contact_surface* out, uint32_t* err);
int orc_collision_check_with_rectangle(const orc_circle* ball, orc_v2 mv, const orc_aabb* aabb,
                                       orc_contact_surface* out, uint32_t* err);
orc_contact orc_contact_test_circle_aabb(const orc_circle* c, const orc_aabb* b, uint32_t* err);
float orc_dir_x_from_bits(uint32_t random_bits);                            /* rand 0.8.5 gen_range(-0.35..-0.15) */
float orc_reset_dir_x(uint64_t seed, uint32_t env_global_id, uint32_t episode);

/* ---------- observation: rasteriser + grayscale + 4-frame ring (render_oracle.c) ---------- */
#define ORC_FRAME_W 84
#define ORC_FRAME_H 84
#define ORC_FRAME_BYTES (ORC_FRAME_W * ORC_FRAME_H)
#define ORC_NUM_FRAMES 4                                                    /* breakout_environment.rs:15 */

void orc_draw_rgb(const orc_mechanics* m, uint8_t* rgb /* [84][84][3] */);  /* breakout_drawer.rs:22-28 (spec) */
void orc_grayscale(const uint8_t* rgb, uint8_t* luma);                      /* image 0.24 imageops::grayscale */

typedef struct {                                                            /* frame_ring_buffer.rs:7-13 */
    uint8_t buffer[ORC_NUM_FRAMES][ORC_FRAME_BYTES];
    int next_slot;
} orc_frame_ring;
void orc_frame_ring_new(orc_frame_ring* r);                                 /* frame_ring_buffer.rs:17-31 */
void orc_frame_ring_add(orc_frame_ring* r, const uint8_t* frame);           /* frame_ring_buffer.rs:53-63 */

/* BreakoutState / BreakoutEnvironment (breakout_environment.rs:24-28,131-207) */
typedef struct {
    orc_mechanics mechanics;
    orc_frame_ring frame_buffer;
} orc_state;

typedef struct {
    orc_state state;
    /* vectorised-driver bookkeeping (ours): */
    uint64_t seed;
    uint32_t env_global_id;
    uint32_t episode;        /* episodes started so far - 1 */
    uint32_t episode_step;   /* steps taken in the current episode */
    float    episode_return;
    uint32_t sticky_err;     /* OR of the error flags of all finished episodes (mechanics.err restarts at 0) */
} orc_env;

void  orc_env_reset_with(orc_env* e, float dir_x);                          /* breakout_environment.rs:177-180 */
void  orc_env_step(orc_env* e, int action, float* reward, int* done);       /* breakout_environment.rs:184-201 */
float orc_env_goal_mean(void);                                              /* breakout_environment.rs:203-206 */
void  orc_state_to_f32_xyh(const orc_state* s, float* out /* [84][84][4] */); /* breakout_environment.rs:42-54 */

#ifdef __cplusplus
}
#endif
#endif
