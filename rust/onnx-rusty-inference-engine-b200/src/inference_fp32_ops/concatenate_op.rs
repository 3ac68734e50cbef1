//! concatenation(), concatenate_op.rs:11-41 -> b200_concat: exactly two 4-D inputs, `axis` the only attribute.
use std::ptr;

use onnx_protobuf::NodeProto;

use super::slot4;
use crate::device::{check, default_context, DeviceTensor, Store};

pub fn concatenation(output_container: &Store, node: &NodeProto) {
    let a = slot4(output_container, &node.input[0], "Concat");
    let b = slot4(output_container, &node.input[1], "Concat");
    let mut axis: i64 = 1;                                              // concatenate_op.rs:22
    for attr in &node.attribute {
        match attr.name.as_ref() {
            "axis" => axis = attr.i,
            _ => panic!("ATTRIBUTE NAME FOR CONCATENATE NOT FOUND, {}", <String as AsRef<str>>::as_ref(&attr.name)),
        }
    }
    let ctx = default_context();
    let mut y = ptr::null_mut();
    check(unsafe { b200rt_sys::b200_concat(ctx.raw(), a.raw(), b.raw(), axis, &mut y) }).unwrap_or_else(|e| panic!("b200rt: {}", e.message));
    output_container.lock().unwrap().insert(node.output[0].clone(), (None, Some(DeviceTensor::from_raw(ctx.clone(), y))));
}
