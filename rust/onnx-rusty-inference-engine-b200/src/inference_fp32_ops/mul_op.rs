//! mul() (ONNX MatMul), mul_op.rs:11-32 -> b200_matmul (the convolution's tcgen05 path: rows as pixels).
use std::ptr;

use onnx_protobuf::NodeProto;

use super::slot2;
use crate::device::{check, default_context, DeviceTensor, Store};

pub fn mul(output_container: &Store, node: &NodeProto) {
    let a = slot2(output_container, &node.input[0], "MatMul");       // both operands from the 2-D slot, mul_op.rs:16-19
    let b = slot2(output_container, &node.input[1], "MatMul");
    let ctx = default_context();
    let mut y = ptr::null_mut();
    check(unsafe { b200rt_sys::b200_matmul(ctx.raw(), a.raw(), b.raw(), ptr::null(), &mut y) }).unwrap_or_else(|e| panic!("b200rt: {}", e.message));
    output_container.lock().unwrap().insert(node.output[0].clone(), (Some(DeviceTensor::from_raw(ctx.clone(), y)), None));
}
