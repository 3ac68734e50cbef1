// Same ten modules as the reference (src/inference_fp32_ops/mod.rs:1-10).
pub mod convolution_op;
pub mod dropout_op;
pub mod global_average_pool_op;
pub mod max_pool_op;
pub mod relu_op;
pub mod reshape_op;
pub mod softmax_op;
pub mod mul_op;
pub mod add_op;
pub mod concatenate_op;

use std::ptr;

use crate::device::{check, default_context, DeviceTensor, Store};

/// 4-D slot of a store entry (the reference's `map.get(..).unwrap().1.unwrap()`); panics like the reference when absent.
pub(crate) fn slot4(store: &Store, name: &str, what: &str) -> DeviceTensor {
    let map = store.lock().unwrap();
    map.get(name).and_then(|v| v.1.clone()).unwrap_or_else(|| panic!("{what}: {name} has no 4-D slot in the store"))
}

/// 2-D slot (mul_op.rs:16-19).
pub(crate) fn slot2(store: &Store, name: &str, what: &str) -> DeviceTensor {
    let map = store.lock().unwrap();
    map.get(name).and_then(|v| v.0.clone()).unwrap_or_else(|| panic!("{what}: {name} has no 2-D slot in the store"))
}

/// Runs one unary b200_* entry point that allocates its output.
pub(crate) fn unary(f: unsafe extern "C" fn(*mut b200rt_sys::b200_ctx, *const b200rt_sys::b200_tensor, *mut *mut b200rt_sys::b200_tensor) -> i32,
                    x: &DeviceTensor) -> DeviceTensor {
    let ctx = default_context();
    let mut y = ptr::null_mut();
    check(unsafe { f(ctx.raw(), x.raw(), &mut y) }).unwrap_or_else(|e| panic!("b200rt: {}", e.message));
    DeviceTensor::from_raw(ctx.clone(), y)
}
