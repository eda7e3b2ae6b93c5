//! ReplayBuffer with the method set of ql-with-tensorflow/src/learn/replay_buffer.rs:53-146, backed by the HBM frame ring.
use std::rc::Rc;

use anyhow::Result;
use ql::prelude::{Action, QlError};

use crate::env::{BreakoutAction, CudaBreakoutEnvironment, CudaBreakoutState, Handle, StateKind};
use crate::{check, ffi};

pub struct BufferSample<const N: usize> {
    pub state: [Rc<CudaBreakoutState>; N],
    pub state_next: [Rc<CudaBreakoutState>; N],
    pub reward: [f32; N],
    pub action: [BreakoutAction; N],
    pub done: [bool; N],
}

pub struct ReplayBuffer {
    env: Rc<Handle>,
}

impl ReplayBuffer {
    /// `ReplayBuffer::new(step_buffer_len, episode_reward_buffer_len)`; the ring itself was sized at env creation.
    pub fn new(env: &CudaBreakoutEnvironment, step_buffer_len: usize, _episode_reward_buffer_len: usize) -> Result<Self> {
        let h = env.handle();
        let mut cap = 0u64;
        check(unsafe { ffi::qlc_replay_capacity(h.0, &mut cap) })?;
        if (cap as usize) < step_buffer_len {
            Err(QlError("the environment's replay ring is shorter than step_buffer_len".to_string()))?
        }
        Ok(Self { env: h })
    }

    pub fn len(&self) -> usize {
        let mut n = 0u64;
        unsafe { ffi::qlc_replay_len(self.env.0, &mut n) };
        n as usize
    }

    /// The step kernel already appended this transition on the device (frame + 4-byte record); the call keeps the
    /// learner's call site (self_driving_tf_q_learner.rs:177) unchanged and checks the handles.
    pub fn add(&mut self, _action: BreakoutAction, state: Rc<CudaBreakoutState>, state_next: Rc<CudaBreakoutState>, _reward: f32, _done: bool) {
        debug_assert!(state_next.time == self.env.time() && state.time + 1 == state_next.time);
    }

    pub fn add_episode_reward(&mut self, episode_reward: f32) { unsafe { ffi::qlc_stats_push(self.env.0, episode_reward) }; }

    pub fn avg_episode_reward(&self) -> f32 {
        let mut v = 0f32;
        check(unsafe { ffi::qlc_stats_mean(self.env.0, &mut v) }).expect("episode reward history is empty");
        v
    }

    pub fn min_episode_reward(&self) -> f32 {
        let mut v = 0f32;
        check(unsafe { ffi::qlc_stats_min(self.env.0, &mut v) }).expect("episode reward history is empty");
        v
    }

    /// histogram of the stored actions (what the learner's log derives from `actions()`, :242-245)
    pub fn action_counts(&self) -> [u64; 3] {
        let mut c = [0u64; 3];
        unsafe { ffi::qlc_replay_action_counts(self.env.0, c.as_mut_ptr()) };
        c
    }

    pub fn episode_rewards(&self) -> Vec<f32> {
        let mut n = 0u32;
        unsafe { ffi::qlc_stats_window(self.env.0, std::ptr::null_mut(), 0, &mut n) };
        let mut out = vec![0f32; n as usize];
        unsafe { ffi::qlc_stats_window(self.env.0, out.as_mut_ptr(), n, &mut n) };
        out
    }

    pub fn get_many<const N: usize>(&self, indices: &[usize; N]) -> Result<BufferSample<N>> {
        let idx: [u32; N] = indices.map(|i| i as u32);
        let (mut reward, mut action, mut done) = ([0f32; N], [0u8; N], [0u8; N]);
        check(unsafe {
            ffi::qlc_replay_gather_host(self.env.0, idx.as_ptr(), N as u32, ffi::QLC_LAYOUT_U8_BHYX, std::ptr::null_mut(), std::ptr::null_mut(),
                                        reward.as_mut_ptr(), action.as_mut_ptr(), done.as_mut_ptr())
        })?;
        let now = self.env.time();
        let mut act = [BreakoutAction::None; N];
        for i in 0..N {
            act[i] = BreakoutAction::try_from_numeric(action[i])?;
        }
        Ok(BufferSample {
            state: idx.map(|i| Rc::new(CudaBreakoutState::new(Rc::clone(&self.env), StateKind::ReplayState(i), now))),
            state_next: idx.map(|i| Rc::new(CudaBreakoutState::new(Rc::clone(&self.env), StateKind::ReplayNext(i), now))),
            reward,
            action: act,
            done: done.map(|d| d != 0),
        })
    }

    pub(crate) fn handle(&self) -> &Rc<Handle> { &self.env }
}

/// `generate_distinct_random_ids` (self_driving_tf_q_learner.rs:276-296): BATCH distinct uniform ids in `0..len`, drawn
/// on the device from the Philox stream (seed, call_index).
pub fn generate_distinct_random_ids<const BATCH_SIZE: usize>(replay: &ReplayBuffer, call_index: u64) -> Result<[usize; BATCH_SIZE]> {
    let mut idx = [0u32; BATCH_SIZE];
    check(unsafe { ffi::qlc_replay_sample_host(replay.handle().0, BATCH_SIZE as u32, call_index, idx.as_mut_ptr()) })?;
    Ok(idx.map(|i| i as usize))
}
