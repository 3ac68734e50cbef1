//! drop_out(), dropout_op.rs:12-50 -> b200_dropout: `ratio` the only attribute, identity at inference (a zero-copy alias).
use std::ptr;

use onnx_protobuf::NodeProto;

use super::slot4;
use crate::device::{check, default_context, DeviceTensor, Store};

pub fn drop_out(output_container: &Store, node: &NodeProto) {
    let x = slot4(output_container, &node.input[0], "Dropout");
    let mut ratio: f32 = 0.5;
    for attr in &node.attribute {
        match attr.name.as_ref() {
            "ratio" => ratio = attr.f,
            _ => panic!("ATTRIBUTE NAME FOR DROP OUT NOT FOUND, {}", <String as AsRef<str>>::as_ref(&attr.name)),
        }
    }
    let ctx = default_context();
    let mut y = ptr::null_mut();
    check(unsafe { b200rt_sys::b200_dropout(ctx.raw(), x.raw(), ratio, &mut y) }).unwrap_or_else(|e| panic!("b200rt: {}", e.message));
    output_container.lock().unwrap().insert(node.output[0].clone(), (None, Some(DeviceTensor::from_raw(ctx.clone(), y))));
}
