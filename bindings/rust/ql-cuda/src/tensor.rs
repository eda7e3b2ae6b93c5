//! `ToMultiDimArray<Tensor<f32>>` for the state handle (trait: ql-with-tensorflow/src/ml_model/model.rs:12-26; the reference's
//! impl for its pixel-owning BreakoutState: _breakout-ml/src/breakout_environment.rs:39-78). Same dims, same element order
//! `[b][x][y][hist]`, same values (`u8 as f32`, no scaling) — produced by one gather kernel instead of per-element `Tensor::set`.
use std::rc::Rc;

use ql_with_tensorflow::ml_model::model::ToMultiDimArray;
use tensorflow::Tensor;

use crate::env::CudaBreakoutState;

impl ToMultiDimArray<Tensor<f32>> for CudaBreakoutState {
    fn dims(&self) -> &[u64] { self.model_dims() }

    fn to_multi_dim_array(&self) -> Tensor<f32> {
        let values = CudaBreakoutState::gather_f32(&[self]).expect("qlc_obs_gather_host");
        Tensor::new(self.model_dims()).with_values(&values).expect("tensor shape")
    }

    fn batch_to_multi_dim_array<const N: usize>(batch: &[&Rc<Self>; N]) -> Tensor<f32> {
        let refs: Vec<&CudaBreakoutState> = batch.iter().map(|s| s.as_ref()).collect();
        let values = CudaBreakoutState::gather_f32(&refs).expect("qlc_obs_gather_host");
        let d = batch[0].model_dims();
        Tensor::new(&[N as u64, d[0], d[1], d[2]]).with_values(&values).expect("tensor shape")
    }
}
