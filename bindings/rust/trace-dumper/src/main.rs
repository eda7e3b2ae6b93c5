//! Per-step trace of the reference mechanics (breakout-game/src/mechanics.rs) for explicit inputs.
//! One line per step: step cx cy dx dy pad_min_x pad_max_x pad_speed (f32 bit patterns, hex) n_bricks score finished
use std::env;
use std::fs;

use breakout_game::mechanics::{BreakoutMechanics, GameInput, PanelControl};
use egui::Vec2;

fn main() {
    let args: Vec<String> = env::args().collect();
    let dir_x: f32 = args[1].parse().expect("dir_x");
    let actions: Vec<u8> = fs::read_to_string(&args[2]).expect("actions file").split_whitespace().map(|t| t.parse().expect("action 0/1/2")).collect();

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
