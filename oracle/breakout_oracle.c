/* TEST INFRASTRUCTURE ONLY — see breakout_oracle.h. Physics half of the oracle.
 *
 * Follows, function by function:
 *   /root/reference/src/breakout-game/src/mechanics.rs   (constants :12-44, setup :57-116, time_step :119-129,
 *        proceed_ball_with :137-184, check_collisions :186-214, Ball::* :257-444, ContactCandidates :485-539,
 *        Panel :551-588, speed helpers :612-649)
 *   /root/reference/src/breakout-game/src/algebra_2d.rs  (:16-30 AaBB, :46-60 vectors, :62-75 contact call)
 * and restates the third-party arithmetic those files call, whose sources are NOT under /root/reference
 * (versions pinned by /root/reference/src/Cargo.lock): emath 0.22.0 Vec2 {length,normalized,dot}, parry2d 0.13.8
 * query::contact(ball, cuboid) with nalgebra 0.32.6 norms, rand 0.8.5 gen_range::<f32>. Those restatements are
 * from the published sources as recalled and are pinned only by the rstest cases at mechanics.rs:659-752.
 *
 * Build with: gcc -O2 -ffp-contract=off -fno-fast-math (Rust never contracts a*b+c; f32::hypot/acos/round call
 * glibc hypotf/acosf/roundf on Linux).
 */
#include "breakout_oracle.h"
#include "philox.h"
#include <math.h>
#include <string.h>

/* ---- mechanics.rs:12-44 ---- */
static const float MODEL_GRID_LEN_X = 600.0f;
static const float MODEL_GRID_LEN_Y = 600.0f;
static const float CEILING_HEIGHT_Y = 0.0f;
static const float SPACE_GRANULARITY = 0.001f;
static const float TIME_GRANULARITY_SECS = 0.02f;       /* Duration::from_millis(20).as_secs_f32() */
static const float PANEL_LEN_X = 60.0f;
static const float PANEL_LEN_Y = 10.0f;
#define PANEL_CENTER_POS_Y (MODEL_GRID_LEN_Y - 30.0f)
static const float PANEL_MAX_SPEED_PER_SECOND = 160.0f;
static const float PANEL_CONTROL_ACCEL_PER_SECOND = 20.0f;
static const float PANEL_SLOW_DOWN_ACCEL_PER_SECOND = 7.0f;
static const float BRICK_EDGE_LEN = 25.0f;
static const float BRICKS_SETUP_SPACING = 2.0f;
#define BRICKS_SETUP_ROWS 3
static const float BALL_RADIUS = 10.0f;
#define BRICKS_SETUP_DISTANCE_LEFT_WALL (BALL_RADIUS * 3.0f)
#define BRICKS_SETUP_MIN_DISTANCE_RIGHT_WALL BRICKS_SETUP_DISTANCE_LEFT_WALL
static const float BRICKS_SETUP_FIRST_ROW_TOP_Y = 60.0f;
static const float BALL_SPEED_PER_SEC = 200.0f;
static const float CONTACT_PREDICTION = 0.8f;
static const float CONTACT_PENETRATION_LIMIT = 0.0f;
static const float FRAC_PI_2 = 1.57079632679489661923132169163975144f;

/* ---- emath 0.22 Vec2 / Pos2 (restated) ---- */
static inline orc_v2 v2(float x, float y) { orc_v2 r = {x, y}; return r; }
static inline orc_v2 v2_add(orc_v2 a, orc_v2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline orc_v2 v2_sub(orc_v2 a, orc_v2 b) { return v2(a.x - b.x, a.y - b.y); }
static inline orc_v2 v2_scale(orc_v2 a, float f) { return v2(a.x * f, a.y * f); }   /* Vec2 * f32 and f32 * Vec2 */
static inline float  v2_length(orc_v2 a) { return hypotf(a.x, a.y); }               /* Vec2::length = x.hypot(y) */
static inline float  v2_dot(orc_v2 a, orc_v2 b) { return a.x * b.x + a.y * b.y; }
static inline orc_v2 v2_normalized(orc_v2 a) {
    float len = v2_length(a);
    if (len <= 0.0f) return a;
    return v2(a.x / len, a.y / len);
}

/* ---- algebra_2d.rs ---- */
static inline orc_v2 aabb_center(const orc_aabb* b) {                                /* :16-18 */
    return v2((b->min.x + b->max.x) / 2.0f, (b->min.y + b->max.y) / 2.0f);
}
static inline orc_aabb aabb_translate(const orc_aabb* b, orc_v2 v) {                 /* :20-30 */
    orc_aabb r; r.min = v2_add(b->min, v); r.max = v2_add(b->max, v); return r;
}
static inline orc_v2 reflected_vector(orc_v2 v, orc_v2 n) {                          /* :46-52 */
    float f = 2.0f * v2_dot(v, n);
    return v2_sub(v, v2_scale(n, f));
}
static inline float vector_angle(orc_v2 a, orc_v2 b) {                               /* :54-60 */
    return acosf(v2_dot(v2_normalized(a), v2_normalized(b)));
}

/* parry2d 0.13.8 query::contact(pos_ball, Ball, pos_cuboid, Cuboid, prediction) for pure translations,
 * i.e. contact_ball_convex_polyhedron = contact_convex_polyhedron_ball(pos12^-1, cuboid, ball).flipped(),
 * with Cuboid -> Aabb::project_local_point_and_get_feature and nalgebra Unit::try_new_and_get.
 * Call site: algebra_2d.rs:62-75. */
orc_contact orc_contact_test_circle_aabb(const orc_circle* circle, const orc_aabb* aabb, uint32_t* err) {
    orc_contact out; memset(&out, 0, sizeof out);
    orc_v2 ac = aabb_center(aabb);
    orc_v2 he = v2((aabb->max.x - aabb->min.x) / 2.0f, (aabb->max.y - aabb->min.y) / 2.0f);
    /* ball centre in the cuboid's local frame: -(t_cuboid - t_ball) */
    float pt[2] = { -(ac.x - circle->center.x), -(ac.y - circle->center.y) };
    float mins[2] = { -he.x, -he.y }, maxs[2] = { he.x, he.y };
    float mins_pt[2], pt_maxs[2], shift[2];
    for (int i = 0; i < 2; ++i) {
        mins_pt[i] = mins[i] - pt[i];
        pt_maxs[i] = pt[i] - maxs[i];
        float a = mins_pt[i] > 0.0f ? mins_pt[i] : 0.0f;   /* sup(0) */
        float b = pt_maxs[i] > 0.0f ? pt_maxs[i] : 0.0f;
        shift[i] = a - b;
    }
    int inside = (shift[0] == 0.0f && shift[1] == 0.0f);
    float proj[2], fshift[2];                               /* fshift = the shift do_project_local_point returns */
    if (!inside) {
        proj[0] = pt[0] + shift[0]; proj[1] = pt[1] + shift[1];
        fshift[0] = shift[0]; fshift[1] = shift[1];
    } else {
        /* non-solid projection: push to the nearest face */
        float best = -3.40282347e+38f; int is_mins = 0, best_id = 0;
        for (int i = 0; i < 2; ++i) {
            if (mins_pt[i] < pt_maxs[i]) {
                if (pt_maxs[i] > best) { best_id = i; is_mins = 0; best = pt_maxs[i]; }
            } else if (mins_pt[i] > best) { best_id = i; is_mins = 1; best = mins_pt[i]; }
        }
        fshift[0] = 0.0f; fshift[1] = 0.0f;
        fshift[best_id] = is_mins ? best : -best;
        proj[0] = pt[0] + fshift[0]; proj[1] = pt[1] + fshift[1];
    }
    float vx = proj[0] - pt[0], vy = proj[1] - pt[1];
    float sq = vx * vx + vy * vy;                            /* nalgebra norm_squared */
    const float eps = 1.1920929e-7f;                         /* DEFAULT_EPSILON = f32::EPSILON */
    float dist; orc_v2 poly_normal;                          /* cuboid-side normal (points box -> ball when outside) */
    if (sq > eps * eps) {                                    /* Unit::try_new_and_get(proj - centre, eps) */
        float len = sqrtf(sq);
        orc_v2 dir = v2(vx / len, vy / len);
        if (inside) { dist = -len - circle->radius; poly_normal = dir; }
        else        { dist =  len - circle->radius; poly_normal = v2(-dir.x, -dir.y); }
    } else {
        /* degenerate: ball centre (within eps of) the box boundary. Unreachable in play (needs >= r
         * penetration) but exercised by the reference's own test case mechanics.rs:721, so restated in full:
         * Aabb::project_local_point_and_get_feature (2-D) -> Cuboid::feature_normal, falling back to
         * normalize(proj.point) and then to the y axis. Flagged (informational). */
        if (err) *err |= ORC_ERR_DEGENERATE;
        dist = -circle->radius;
        int nzero = 0, last_not_zero = 0;
        for (int i = 0; i < 2; ++i) { if (fshift[i] == 0.0f) nzero++; else last_not_zero = i; }
        int have = 0; float fnx = 0.0f, fny = 0.0f;
        if (nzero == 2) {
            for (int i = 0; i < 2 && !have; ++i) {
                if (proj[i] > maxs[i] - eps)       { have = 1; if (i == 0) fnx = 1.0f;  else fny = 1.0f; }
                else if (proj[i] <= mins[i] + eps) { have = 1; if (i == 0) fnx = -1.0f; else fny = -1.0f; }
            }
        } else if (nzero == 1) {
            float centre_i = (mins[last_not_zero] + maxs[last_not_zero]) / 2.0f;
            float sgn = (proj[last_not_zero] < centre_i) ? -1.0f : 1.0f;
            have = 1; if (last_not_zero == 0) fnx = sgn; else fny = sgn;
        } else {
            float cx0 = (mins[0] + maxs[0]) / 2.0f, cy0 = (mins[1] + maxs[1]) / 2.0f;
            float dx = (proj[0] < cx0) ? -1.0f : 1.0f, dy = (proj[1] < cy0) ? -1.0f : 1.0f;
            float n = sqrtf(dx * dx + dy * dy);
            have = 1; fnx = dx / n; fny = dy / n;
        }
        if (!have) {                                         /* FeatureId::Unknown -> Unit::try_new(proj.point) */
            float psq = proj[0] * proj[0] + proj[1] * proj[1];
            if (psq > eps * eps) { float n = sqrtf(psq); fnx = proj[0] / n; fny = proj[1] / n; }
            else { fnx = 0.0f; fny = 1.0f; }                 /* Vector::y_axis() */
        }
        poly_normal = v2(fnx, fny);
    }
    if (dist <= CONTACT_PREDICTION) {
        out.some = 1;
        out.dist = dist;
        /* after Contact::flipped(): normal1 = ball side (= -poly_normal), normal2 = cuboid side */
        out.normal1 = v2(-poly_normal.x, -poly_normal.y);
        out.normal2 = poly_normal;
    }
    return out;
}

/* ---- Ball (mechanics.rs:257-444) ---- */
int orc_collision_test_left_wall(const orc_circle* ball, orc_v2 mv, orc_contact_surface* out, uint32_t* err) {
    float wall_distance_x = ball->center.x - ball->radius;
    if (!(wall_distance_x >= 0.0f) && err) *err |= ORC_ERR_WALL_DISTANCE;
    if (wall_distance_x + mv.x > 0.0f) return 0;
    orc_v2 w = v2_scale(mv, wall_distance_x / fabsf(mv.x));
    out->way = v2_length(w); out->approximation = 0.0f; out->surface_normal = v2(1.0f, 0.0f);
    return 1;
}
int orc_collision_test_right_wall(const orc_circle* ball, orc_v2 mv, orc_contact_surface* out, uint32_t* err) {
    float wall_distance_x = MODEL_GRID_LEN_X - ball->center.x - ball->radius;
    if (!(wall_distance_x >= 0.0f) && err) *err |= ORC_ERR_WALL_DISTANCE;
    if (mv.x < wall_distance_x) return 0;
    orc_v2 w = v2_scale(mv, wall_distance_x / fabsf(mv.x));
    out->way = v2_length(w); out->approximation = 0.0f; out->surface_normal = v2(-1.0f, 0.0f);
    return 1;
}
int orc_collision_test_top_wall(const orc_circle* ball, orc_v2 mv, orc_contact_surface* out, uint32_t* err) {
    float wall_distance_y = ball->center.y - ball->radius - CEILING_HEIGHT_Y;
    if (!(wall_distance_y >= 0.0f) && err) *err |= ORC_ERR_WALL_DISTANCE;
    if (wall_distance_y + mv.y > 0.0f) return 0;
    orc_v2 w = v2_scale(mv, wall_distance_y / fabsf(mv.y));
    out->way = v2_length(w); out->approximation = 0.0f; out->surface_normal = v2(0.0f, 1.0f);
    return 1;
}

static float moved_distance_after_collision(float p, orc_v2 n1, orc_v2 mv) {        /* :351-358 */
    return p / (v2_dot(n1, mv) / v2_length(mv));
}

static orc_contact_surface binary_search_first_contact(const orc_circle* ball, orc_v2 mv, float start, float end,
                                                       const orc_aabb* aabb, int depth, uint32_t* err) { /* :361-389 */
    float m = (start + end) / 2.0f;
    orc_circle c; c.center = v2_add(ball->center, v2_scale(mv, m)); c.radius = ball->radius;
    orc_contact ct = orc_contact_test_circle_aabb(&c, aabb, err);
    if (depth >= ORC_MAX_BISECTION) {
        /* the reference would recurse without bound; flag and return what we have */
        if (err) *err |= ORC_ERR_BISECTION;
        orc_contact_surface s; s.way = v2_length(mv) * m; s.approximation = 0.0f;
        s.surface_normal = ct.some ? ct.normal2 : v2(0.0f, 1.0f);
        return s;
    }
    if (!ct.some) return binary_search_first_contact(ball, mv, m, end, aabb, depth + 1, err);
    if (ct.dist < -CONTACT_PENETRATION_LIMIT) return binary_search_first_contact(ball, mv, start, m, aabb, depth + 1, err);
    orc_contact_surface s; s.way = v2_length(mv) * m; s.approximation = ct.dist; s.surface_normal = ct.normal2;
    return s;
}

static int find_non_penetrating_collision(const orc_circle* ball, orc_v2 mv, const orc_aabb* aabb,
                                          orc_contact_surface* out, uint32_t* err) {      /* :337-443 */
    orc_circle c; c.center = v2_add(ball->center, mv); c.radius = ball->radius;
    orc_contact ct = orc_contact_test_circle_aabb(&c, aabb, err);
    if (!ct.some) return 0;
    if (ct.dist < -CONTACT_PENETRATION_LIMIT) {
        float x = moved_distance_after_collision(fabsf(ct.dist), ct.normal1, mv);
        float portion = 1.0f - x / v2_length(mv);
        orc_circle c2; c2.center = v2_add(ball->center, v2_scale(mv, portion)); c2.radius = ball->radius;
        orc_contact ct2 = orc_contact_test_circle_aabb(&c2, aabb, err);
        if (!ct2.some) { *out = binary_search_first_contact(ball, mv, portion, 1.0f, aabb, 0, err); return 1; }
        if (ct2.dist < -CONTACT_PENETRATION_LIMIT) { *out = binary_search_first_contact(ball, mv, 0.0f, portion, aabb, 0, err); return 1; }
        out->way = v2_length(mv) * portion; out->approximation = ct2.dist; out->surface_normal = ct2.normal2;
        return 1;
    }
    out->way = v2_length(mv); out->approximation = ct.dist; out->surface_normal = ct.normal2;
    return 1;
}

int orc_collision_check_with_rectangle(const orc_circle* ball, orc_v2 mv, const orc_aabb* aabb,
                                       orc_contact_surface* out, uint32_t* err) {          /* :318-335 */
    orc_contact_surface c;
    if (!find_non_penetrating_collision(ball, mv, aabb, &c, err)) return 0;
    if (fabsf(vector_angle(mv, c.surface_normal)) > FRAC_PI_2) { *out = c; return 1; }
    return 0;
}

/* ---- ContactCandidates (mechanics.rs:485-539) ---- */
typedef struct { float way, approximation; orc_v2 surface_normal; int brick_idx; /* -1 = None */ } cos_t;
typedef struct { cos_t s[4 + ORC_MAX_BRICKS]; int n; } candidates_t;

static inline float path_len(const cos_t* e) { return e->way + e->approximation; }

static void candidates_consider(candidates_t* cs, orc_contact_surface c, int brick_idx, uint32_t* err) { /* :496-516 */
    if (!(c.approximation >= -CONTACT_PENETRATION_LIMIT && c.approximation <= CONTACT_PREDICTION)) *err |= ORC_ERR_APPROX_RANGE;
    cos_t e; e.way = c.way; e.approximation = c.approximation; e.surface_normal = c.surface_normal; e.brick_idx = brick_idx;
    cs->s[cs->n++] = e;
    if (cs->n > 1) {
        float shortest = INFINITY;
        for (int i = 0; i < cs->n; ++i) { float l = path_len(&cs->s[i]); if (l < shortest) shortest = l; }
        int k = 0;
        for (int i = 0; i < cs->n; ++i) if (path_len(&cs->s[i]) <= shortest + SPACE_GRANULARITY) cs->s[k++] = cs->s[i];
        cs->n = k;
    }
}

static int candidates_effective(const candidates_t* cs, orc_contact_surface* out) {   /* :519-538 */
    if (cs->n == 0) return 0;
    if (cs->n == 1) { out->way = cs->s[0].way; out->approximation = cs->s[0].approximation; out->surface_normal = cs->s[0].surface_normal; return 1; }
    orc_v2 sum = v2(0.0f, 0.0f); float dsum = 0.0f, wsum = 0.0f;
    for (int i = 0; i < cs->n; ++i) sum = v2_add(sum, cs->s[i].surface_normal);
    for (int i = 0; i < cs->n; ++i) dsum = dsum + cs->s[i].approximation;
    for (int i = 0; i < cs->n; ++i) wsum = wsum + cs->s[i].way;
    out->surface_normal = v2_normalized(sum);
    out->approximation = dsum / (float)cs->n;
    out->way = wsum / (float)cs->n;
    return 1;
}

/* ---- BreakoutMechanics ---- */
void orc_mechanics_new(orc_mechanics* m, float dir_x) {                               /* :57-116 */
    memset(m, 0, sizeof *m);
    int n = 0;
    for (int row = 0; row < BRICKS_SETUP_ROWS; ++row) {
        float left_x = BRICKS_SETUP_DISTANCE_LEFT_WALL;
        float upper_y = BRICKS_SETUP_FIRST_ROW_TOP_Y + (float)row * (BRICK_EDGE_LEN + BRICKS_SETUP_SPACING);
        int k = 0;
        for (;;) {
            orc_aabb b;
            b.min = v2(left_x, upper_y - BRICK_EDGE_LEN);
            b.max = v2(left_x + BRICK_EDGE_LEN, upper_y);
            if (b.max.x >= MODEL_GRID_LEN_X - BRICKS_SETUP_MIN_DISTANCE_RIGHT_WALL) break;
            left_x = b.max.x + BRICKS_SETUP_SPACING;
            m->bricks[n] = b; m->brick_id[n] = (uint8_t)(20 * row + k); ++n; ++k;
        }
    }
    m->n_bricks = n;
    m->ball_shape.center = v2(MODEL_GRID_LEN_X * 0.5f, MODEL_GRID_LEN_Y * 0.5f);
    m->ball_shape.radius = BALL_RADIUS;
    m->ball_direction = v2(dir_x, -1.0f);
    m->ball_speed_per_sec = BALL_SPEED_PER_SEC;
    m->panel_shape.min = v2(MODEL_GRID_LEN_X / 2.0f - PANEL_LEN_X / 2.0f, PANEL_CENTER_POS_Y - PANEL_LEN_Y / 2.0f);
    m->panel_shape.max = v2(MODEL_GRID_LEN_X / 2.0f + PANEL_LEN_X / 2.0f, PANEL_CENTER_POS_Y + PANEL_LEN_Y / 2.0f);
    m->panel_speed_per_sec = 0.0f;
    m->finished = 0; m->score = 0; m->err = 0;
}

uint64_t orc_mechanics_brick_mask(const orc_mechanics* m) {
    uint64_t mask = 0;
    for (int i = 0; i < m->n_bricks; ++i) mask |= (uint64_t)1 << m->brick_id[i];
    return mask;
}

static void check_collisions(const orc_mechanics* m, orc_v2 mv, candidates_t* cs, uint32_t* err) { /* :186-214 */
    cs->n = 0;
    orc_contact_surface c;
    if (orc_collision_test_left_wall(&m->ball_shape, mv, &c, err))  candidates_consider(cs, c, -1, err);
    if (orc_collision_test_right_wall(&m->ball_shape, mv, &c, err)) candidates_consider(cs, c, -1, err);
    if (orc_collision_test_top_wall(&m->ball_shape, mv, &c, err))   candidates_consider(cs, c, -1, err);
    if (orc_collision_check_with_rectangle(&m->ball_shape, mv, &m->panel_shape, &c, err)) candidates_consider(cs, c, -1, err);
    for (int idx = 0; idx < m->n_bricks; ++idx)
        if (orc_collision_check_with_rectangle(&m->ball_shape, mv, &m->bricks[idx], &c, err)) candidates_consider(cs, c, idx, err);
}

static void proceed_ball_with(orc_mechanics* m, orc_v2 mv, int depth) {               /* :137-184 */
    if (v2_length(mv) < SPACE_GRANULARITY) return;
    candidates_t cs;
    check_collisions(m, mv, &cs, &m->err);

    /* remove hit bricks, highest index first (:150-162) */
    int hit[4 + ORC_MAX_BRICKS]; int nh = 0;
    for (int i = 0; i < cs.n; ++i) if (cs.s[i].brick_idx >= 0) hit[nh++] = cs.s[i].brick_idx;
    for (int i = 1; i < nh; ++i) { int v = hit[i], j = i - 1; while (j >= 0 && hit[j] > v) { hit[j + 1] = hit[j]; --j; } hit[j + 1] = v; }
    for (int i = nh - 1; i >= 0; --i) {
        int idx = hit[i];
        memmove(&m->bricks[idx], &m->bricks[idx + 1], (size_t)(m->n_bricks - idx - 1) * sizeof(orc_aabb));
        memmove(&m->brick_id[idx], &m->brick_id[idx + 1], (size_t)(m->n_bricks - idx - 1));
        m->n_bricks -= 1;
        m->score += 1;
    }

    orc_contact_surface col;
    if (candidates_effective(&cs, &col)) {
        orc_v2 collision_center_pos = v2_add(m->ball_shape.center, v2_scale(m->ball_direction, col.way));
        float remaining_distance = v2_length(mv) - col.way;
        orc_v2 reflected_direction = v2_normalized(reflected_vector(m->ball_direction, col.surface_normal));
        m->ball_shape.center = collision_center_pos;
        m->ball_direction = reflected_direction;
        orc_v2 remaining_mv = v2_scale(reflected_direction, remaining_distance);
        if (v2_length(remaining_mv) > 0.0f) {
            if (depth >= ORC_MAX_RECURSION) { m->err |= ORC_ERR_RECURSION; return; }
            proceed_ball_with(m, remaining_mv, depth + 1);
        }
    } else {
        m->ball_shape.center = v2_add(m->ball_shape.center, mv);
    }
}

static float granulate_speed(float s) { return roundf(s * 1000.0f) / 1000.0f; }      /* :612 */
static float decrease_speed(float s, float brk) {                                    /* :616-628 */
    if (s > 0.0f) return fmaxf(granulate_speed(s - brk), 0.0f);
    else if (s < 0.0f) return fmaxf(granulate_speed(s + brk), 0.0f);
    return 0.0f;
}
static float accelerate(float s, float a, float limit) {                             /* :631-649 */
    float v = s + a, r;
    if (fabsf(v) > limit) r = signbit(v) ? -limit : limit; else r = v;
    return granulate_speed(r);
}

static void panel_proceed(orc_mechanics* m) {                                        /* :571-587 */
    orc_aabb p = aabb_translate(&m->panel_shape, v2(m->panel_speed_per_sec * TIME_GRANULARITY_SECS, 0.0f));
    if (p.min.x <= 0.0f) { m->panel_shape = aabb_translate(&p, v2(-p.min.x, 0.0f)); m->panel_speed_per_sec = 0.0f; }
    else if (p.max.x >= MODEL_GRID_LEN_X) { m->panel_shape = aabb_translate(&p, v2(MODEL_GRID_LEN_X - p.max.x, 0.0f)); m->panel_speed_per_sec = 0.0f; }
    else m->panel_shape = p;
}
static void panel_process_input(orc_mechanics* m, int control) {                     /* :553-566 */
    switch (control) {
    case ORC_CONTROL_NONE:  m->panel_speed_per_sec = decrease_speed(m->panel_speed_per_sec, PANEL_SLOW_DOWN_ACCEL_PER_SECOND); break;
    case ORC_CONTROL_LEFT:  m->panel_speed_per_sec = accelerate(m->panel_speed_per_sec, -PANEL_CONTROL_ACCEL_PER_SECOND, PANEL_MAX_SPEED_PER_SECOND); break;
    case ORC_CONTROL_RIGHT: m->panel_speed_per_sec = accelerate(m->panel_speed_per_sec,  PANEL_CONTROL_ACCEL_PER_SECOND, PANEL_MAX_SPEED_PER_SECOND); break;
    }
}

void orc_mechanics_time_step(orc_mechanics* m, int control) {                        /* :119-129 */
    panel_proceed(m);
    /* Ball::move_vector :258 — ((normalized * speed) * dt) */
    orc_v2 mv = v2_scale(v2_scale(v2_normalized(m->ball_direction), m->ball_speed_per_sec), TIME_GRANULARITY_SECS);
    proceed_ball_with(m, mv, 0);
    if (m->ball_shape.center.y >= m->panel_shape.max.y || m->n_bricks == 0) m->finished = 1;   /* :131-135 */
    if (!m->finished) panel_process_input(m, control);
}

/* rand 0.8.5 UniformFloat<f32>::sample_single(-0.35, -0.15) fed with one u32 of entropy (mechanics.rs:103).
 * The entropy source itself (ThreadRng) is irreproducible, so the u32 is an explicit input. */
float orc_dir_x_from_bits(uint32_t random_bits) {
    const float low = -0.35f, high = -0.15f;
    float scale = high - low;
    uint32_t fb = 0x3F800000u | (random_bits >> 9);
    float value1_2; memcpy(&value1_2, &fb, 4);
    float value0_1 = value1_2 - 1.0f;
    float res = value0_1 * scale + low;
    return res;                                   /* res < high for every input (checked in selftest.c) */
}

float orc_reset_dir_x(uint64_t seed, uint32_t env_global_id, uint32_t episode) {
    uint32_t ctr[4] = { env_global_id, episode, 0u, ORC_STREAM_RESET };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t out[4];
    orc_philox4x32_10(ctr, key, out);
    return orc_dir_x_from_bits(out[0]);
}

/* batched rectangle sweep for fuzzing the device routine: in [n][9] = cx, cy, r, mvx, mvy, minx, miny, maxx, maxy;
 * out [n][6] = some, way, approximation, nx, ny, err (u32 bits in a float slot) */
void orc_collision_rect_batch(const float* in, float* out, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + (size_t)i * 9;
        orc_circle ball = {{a[0], a[1]}, a[2]}; orc_v2 mv = {a[3], a[4]}; orc_aabb box = {{a[5], a[6]}, {a[7], a[8]}};
        orc_contact_surface c; memset(&c, 0, sizeof c); uint32_t err = 0;
        int some = orc_collision_check_with_rectangle(&ball, mv, &box, &c, &err);
        float* o = out + (size_t)i * 6;
        o[0] = some ? 1.0f : 0.0f;
        o[1] = some ? c.way : 0.0f; o[2] = some ? c.approximation : 0.0f; o[3] = some ? c.surface_normal.x : 0.0f; o[4] = some ? c.surface_normal.y : 0.0f;
        memcpy(&o[5], &err, 4);
    }
}
