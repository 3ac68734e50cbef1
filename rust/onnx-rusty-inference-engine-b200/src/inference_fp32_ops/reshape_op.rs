//! reshape(), reshape_op.rs:16-92 -> b200_reshape: shape must be an int64 initializer (:35-43), output always 2-D (:87),
//! elements keep the reference's NCHW memory order (:89).
use std::ptr;

use onnx_protobuf::{NodeProto, TensorProto, ValueInfoProto};

use super::slot4;
use crate::device::{check, default_context, DeviceTensor, Store};
use crate::inference_engine::utils::{already_into_initializer, get_stored_tensor, initializer_i64};

pub fn reshape(output_container: &Store,
               node: &NodeProto,
               model_inputs: &Vec<ValueInfoProto>,
               model_initializers: &Vec<TensorProto>) {
    let data = if already_into_initializer(model_initializers, &node.input[0]) {
        get_stored_tensor(0, node, model_inputs, model_initializers)
    } else {
        slot4(output_container, &node.input[0], "Reshape")
    };
    if !already_into_initializer(model_initializers, &node.input[1]) {
        panic!("Unable to retrieve Shape for Reshape operation");       // reshape_op.rs:42
    }
    let shape = initializer_i64(&node.input[1], model_initializers);
    let ctx = default_context();
    let mut y = ptr::null_mut();
    check(unsafe { b200rt_sys::b200_reshape(ctx.raw(), data.raw(), shape.as_ptr(), shape.len() as i32, &mut y) }).unwrap_or_else(|e| panic!("b200rt: {}", e.message));
    output_container.lock().unwrap().insert(node.output[0].clone(), (Some(DeviceTensor::from_raw(ctx.clone(), y)), None));
}
