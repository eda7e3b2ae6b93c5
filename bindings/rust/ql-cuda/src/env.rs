//! Environment / Action / state (ql/src/prelude.rs:12-68, _breakout-ml/src/breakout_environment.rs:24-207).
use std::fmt::{Debug, Display, Formatter};
use std::rc::Rc;

use anyhow::Result;
use console_engine::screen::Screen;
use ql::prelude::{Action, DebugVisualizer, Environment, ModelActionType, QlError};

use crate::{check, ffi};

/// `Parameter::default().history_buffer_len` (self_driving_tf_q_learner.rs:59): how many steps a state handle must stay alive
/// for the learner's replay FIFO. 1 M frames of 84 x 84 u8 = 7 GB of the 180 GB of HBM.
pub const DEFAULT_HISTORY_BUFFER_LEN: usize = 1_000_000;

/// Owns the `qlc_env` (one env + its frame ring in HBM). Shared by the environment and every state handle.
pub(crate) struct Handle(pub(crate) *mut ffi::qlc_env);

impl Handle {
    pub(crate) fn time(&self) -> u64 {
        let mut t = 0u64;
        unsafe { ffi::qlc_env_time(self.0, &mut t) };
        t
    }
}

impl Drop for Handle {
    fn drop(&mut self) {
        unsafe { ffi::qlc_env_destroy(self.0) };
    }
}

#[derive(Debug, Clone, Copy, Hash, PartialEq, Eq)]
pub enum BreakoutAction {
    None,
    Left,
    Right,
}

impl Action for BreakoutAction {
    const ACTION_SPACE: ModelActionType = 3;

    fn numeric(&self) -> ModelActionType {
        match self {
            BreakoutAction::None => 0,
            BreakoutAction::Left => 1,
            BreakoutAction::Right => 2,
        }
    }

    fn try_from_numeric(value: ModelActionType) -> Result<Self> {
        match value {
            0 => Ok(BreakoutAction::None),
            1 => Ok(BreakoutAction::Left),
            2 => Ok(BreakoutAction::Right),
            _ => Err(QlError("value out of range".to_string()))?,
        }
    }
}

impl Display for BreakoutAction {
    fn fmt(&self, f: &mut Formatter<'_>) -> std::fmt::Result { write!(f, "{:?}", self) }
}

/// BreakoutState (breakout_environment.rs:24-28) as a handle: the observation of the env after `time` steps, `k` of them in
/// the current episode. `Clone` copies the handle; the four frames it names stay in the HBM frame ring and remain readable
/// for `history_buffer_len` further steps — as long as the reference's replay FIFO can still hold the `Rc` around it.
#[derive(Clone)]
pub struct CudaBreakoutState {
    pub(crate) env: Rc<Handle>,
    pub(crate) time: u64,
    pub(crate) k: u32,
    model_dims: [u64; 3],
}

impl CudaBreakoutState {
    pub(crate) fn new(env: Rc<Handle>, time: u64, k: u32, model_dims: [u64; 3]) -> Self { Self { env, time, k, model_dims } }

    pub fn model_dims(&self) -> &[u64] { &self.model_dims }

    pub(crate) fn raw(&self) -> ffi::qlc_obs_handle { ffi::qlc_obs_handle { time: self.time, k: self.k, env: 0 } }

    /// `[b][x][y][hist]` f32, value = u8 as f32 (breakout_environment.rs:56-77) for any mix of handles of ONE environment:
    /// one gather kernel, the stacks cross PCIe as u8 and are widened into the returned vector by the library.
    pub fn gather_f32(batch: &[&CudaBreakoutState]) -> Result<Vec<f32>> {
        if batch.is_empty() {
            return Ok(Vec::new());
        }
        let first = batch[0];
        let per = (first.model_dims[0] * first.model_dims[1] * first.model_dims[2]) as usize;
        let mut handles = Vec::with_capacity(batch.len());
        for s in batch {
            if !Rc::ptr_eq(&s.env, &first.env) {
                Err(QlError("states of different environments in one batch".to_string()))?
            }
            handles.push(s.raw());
        }
        let mut out = vec![0f32; batch.len() * per];
        check(unsafe {
            ffi::qlc_obs_gather_host(first.env.0, handles.as_ptr(), handles.len() as u32, ffi::QLC_LAYOUT_F32_BXYH, out.as_mut_ptr() as *mut _)
        })?;
        Ok(out)
    }
}

impl Debug for CudaBreakoutState {
    fn fmt(&self, f: &mut Formatter<'_>) -> std::fmt::Result { write!(f, "CudaBreakoutState {{ t={}, k={} }}", self.time, self.k) }
}

impl DebugVisualizer for CudaBreakoutState {
    /// describes the environment's *current* mechanics (a handle names frames, not ball coordinates)
    fn one_line_info(&self) -> String {
        let (mut cx, mut cy, mut pmin, mut pmax, mut bricks) = (0f32, 0f32, 0f32, 0f32, 0u64);
        let sh = ffi::qlc_state_host {
            ball_cx: &mut cx, ball_cy: &mut cy, ball_dx: std::ptr::null_mut(), ball_dy: std::ptr::null_mut(),
            pad_min_x: &mut pmin, pad_max_x: &mut pmax, pad_speed: std::ptr::null_mut(), bricks: &mut bricks,
            score: std::ptr::null_mut(), episode_step: std::ptr::null_mut(), episode: std::ptr::null_mut(),
            err: std::ptr::null_mut(), finished: std::ptr::null_mut(),
        };
        unsafe { ffi::qlc_env_read_state(self.env.0, &sh) };
        format!("Breakout [{} bricks, ball_pos: [{cx} {cy}], panel_pos: [{} 570]]", bricks.count_ones(), (pmin + pmax) / 2.0)
    }
    fn render_to_console(&self) -> Screen { todo!() } // as in the reference (breakout_environment.rs:91)
}

/// shard statistics (qlc_episode_stats)
pub type EpisodeStats = ffi::qlc_episode_stats;

/// One Breakout env on the GPU behind `ql::prelude::Environment` (breakout_environment.rs:131-207).
pub struct CudaBreakoutEnvironment {
    env: Rc<Handle>,
    state: CudaBreakoutState,
    model_dims: [u64; 3],
}

impl CudaBreakoutEnvironment {
    /// `BreakoutEnvironment::new(frame_size_x, frame_size_y)` (:139-153), on device 0, seed 0, with a frame ring long enough
    /// for `Parameter::default().history_buffer_len`. Panics like the reference constructor cannot fail: see `with_options`.
    pub fn new(frame_size_x: usize, frame_size_y: usize) -> Self {
        Self::with_options(frame_size_x, frame_size_y, DEFAULT_HISTORY_BUFFER_LEN, 0, 0).expect("qlc_env_create")
    }

    /// `history_buffer_len` = `Parameter::history_buffer_len` of the learner that will hold the handles (must be >= it).
    pub fn with_options(frame_size_x: usize, frame_size_y: usize, history_buffer_len: usize, seed: u64, device: i32) -> Result<Self> {
        let cfg = ffi::qlc_config {
            struct_size: std::mem::size_of::<ffi::qlc_config>() as u32,
            device,
            n_envs: 1,
            env_id_base: 0,
            frame_w: frame_size_x as u32,
            frame_h: frame_size_y as u32,
            seed,
            replay_capacity: history_buffer_len as u64,
            max_episode_steps: 0, // the learner counts the steps of an episode itself (:149)
            episode_window: 100,
            auto_reset: 0, // the learner resets: learn_episode :142
            reserved: 0,
        };
        let mut h: *mut ffi::qlc_env = std::ptr::null_mut();
        check(unsafe { ffi::qlc_env_create(&cfg, &mut h) })?;
        let env = Rc::new(Handle(h));
        let model_dims = [frame_size_x as u64, frame_size_y as u64, ffi::QLC_NUM_FRAMES as u64];
        let state = CudaBreakoutState::new(Rc::clone(&env), 0, 0, model_dims);
        Ok(Self { env, state, model_dims })
    }

    /// 1 while the episode runs, 0 once it is over: the reference game ends with the first miss (mechanics.rs:131-135).
    pub fn lives(&self) -> u8 {
        let mut l = 0u8;
        check(unsafe { ffi::qlc_env_lives_host(self.env.0, &mut l) }).expect("qlc_env_lives_host");
        l
    }

    /// OR of the sticky error flags (conditions on which the reference panics: mechanics.rs:145,265,284,303,511)
    pub fn error_flags(&self) -> u32 {
        let mut e = 0u32;
        check(unsafe { ffi::qlc_env_error_flags(self.env.0, &mut e) }).expect("qlc_env_error_flags");
        e
    }

    pub(crate) fn handle(&self) -> Rc<Handle> { Rc::clone(&self.env) }
}

impl Environment for CudaBreakoutEnvironment {
    type S = CudaBreakoutState;
    type A = BreakoutAction;

    fn reset(&mut self) {
        check(unsafe { ffi::qlc_env_reset(self.env.0, std::ptr::null(), std::ptr::null()) }).expect("qlc_env_reset");
        self.state = CudaBreakoutState::new(Rc::clone(&self.env), self.env.time(), 0, self.model_dims);
    }

    fn state(&self) -> &Self::S { &self.state }

    fn step(&mut self, action: BreakoutAction) -> (&CudaBreakoutState, f32, bool) {
        let a = action.numeric();
        let (mut reward, mut done) = (0f32, 0u8);
        check(unsafe { ffi::qlc_env_step_host(self.env.0, &a, 1, &mut reward, &mut done) }).expect("qlc_env_step_host");
        self.state = CudaBreakoutState::new(Rc::clone(&self.env), self.state.time + 1, self.state.k + 1, self.model_dims);
        (&self.state, reward, done != 0)
    }

    fn episode_reward_goal_mean(&self) -> f32 { unsafe { ffi::qlc_env_goal_mean() } }
}
