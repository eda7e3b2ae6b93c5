//! Pins the CPU oracle (oracle/*.c) against the REAL reference wherever a Rust toolchain exists. Two modes:
//!
//! `trace-dumper <dir_x> <actions.txt>` — per-step trace of the reference mechanics (breakout-game/src/mechanics.rs) for explicit
//! inputs. One line per step: step cx cy dx dy pad_min_x pad_max_x pad_speed (f32 bit patterns, hex) n_bricks score finished.
//!
//! `trace-dumper --contacts <inputs.txt>` — the third-party arithmetic under the collision code, at bit level: for every input
//! line `cx cy radius min_x min_y max_x max_y` (hex f32 bits; tests/golden/contact_inputs_v1.txt) the result of the SAME call
//! `contact_test_circle_aabb` makes (breakout-game/src/algebra_2d.rs:62-75 — that function sits in a private module, so its body
//! is repeated here verbatim in meaning): `parry2d::query::contact(ball at centre, Ball(r), cuboid at box centre, Cuboid(half
//! extents), prediction 0.8)`. One line per input: `0`, or `1 dist normal1.x normal1.y normal2.x normal2.y` (hex f32 bits).
//! Compare with `python tests/compare_rust_trace.py --contacts <out> <inputs.txt>` in the B200 repository.
use std::env;
use std::fs;

use breakout_game::mechanics::{BreakoutMechanics, GameInput, PanelControl};
use egui::Vec2;
use nalgebra::{Isometry2, Vector2};
use parry2d::query;
use parry2d::shape::{Ball, Cuboid};

const CONTACT_PREDICTION: f32 = 0.8; // breakout-game/src/mechanics.rs:42

fn bits(tok: &str) -> f32 { f32::from_bits(u32::from_str_radix(tok, 16).expect("hex f32 bits")) }

fn contacts(path: &str) {
    for line in fs::read_to_string(path).expect("inputs file").lines() {
        let t: Vec<&str> = line.split_whitespace().collect();
        if t.len() != 7 {
            continue;
        }
        let (cx, cy, r, min_x, min_y, max_x, max_y) = (bits(t[0]), bits(t[1]), bits(t[2]), bits(t[3]), bits(t[4]), bits(t[5]), bits(t[6]));
        // AaBB::center (algebra_2d.rs:17) and the half extents exactly as algebra_2d.rs:66-73 computes them
        let (acx, acy) = ((min_x + max_x) / 2.0, (min_y + max_y) / 2.0);
        let c = query::contact(
            &Isometry2::translation(cx, cy),
            &Ball::new(r),
            &Isometry2::translation(acx, acy),
            &Cuboid::new(Vector2::new((max_x - min_x) / 2.0, (max_y - min_y) / 2.0)),
            CONTACT_PREDICTION,
        )
        .expect("contact calculation failed");
        match c {
            None => println!("0"),
            Some(c) => println!(
                "1 {:08x} {:08x} {:08x} {:08x} {:08x}",
                c.dist.to_bits(),
                c.normal1.x.to_bits(),
                c.normal1.y.to_bits(),
                c.normal2.x.to_bits(),
                c.normal2.y.to_bits()
            ),
        }
    }
}

fn trace(dir_x: f32, actions_path: &str) {
    let actions: Vec<u8> = fs::read_to_string(actions_path).expect("actions file").split_whitespace().map(|t| t.parse().expect("action 0/1/2")).collect();
    let mut m = BreakoutMechanics::default();
    // the initial direction is the only random draw of the mechanics (mechanics.rs:103): make it an explicit input
    m.ball.direction = Vec2::new(dir_x, -1.0);
    for (step, a) in actions.iter().enumerate() {
        let control = match a {
            0 => PanelControl::None,
            1 => PanelControl::AccelerateLeft,
            2 => PanelControl::AccelerateRight,
            _ => panic!("value out of range"),
        };
        m.time_step(GameInput::action(control));
        println!(
            "{} {:08x} {:08x} {:08x} {:08x} {:08x} {:08x} {:08x} {} {} {}",
            step,
            m.ball.shape.center.x.to_bits(),
            m.ball.shape.center.y.to_bits(),
            m.ball.direction.x.to_bits(),
            m.ball.direction.y.to_bits(),
            m.panel.shape.min.x.to_bits(),
            m.panel.shape.max.x.to_bits(),
            m.panel.speed_per_sec.to_bits(),
            m.bricks.len(),
            m.score,
            m.finished as u8
        );
        if m.finished {
            break;
        }
    }
}

fn main() {
    let args: Vec<String> = env::args().collect();
    if args.len() == 3 && args[1] == "--contacts" {
        contacts(&args[2]);
    } else {
        trace(args[1].parse().expect("dir_x"), &args[2]);
    }
}
