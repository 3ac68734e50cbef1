//! max_pool(), max_pool_op.rs:65-129 -> b200_maxpool2d (zero-fill padding, pads honoured only under NOTSET).
use std::ptr;

use b200rt_sys as sys;
use onnx_protobuf::{NodeProto, TensorProto, ValueInfoProto};

use crate::device::{check, default_context, DeviceTensor, Store};
use crate::inference_engine::utils::get_stored_tensor;

pub fn max_pool(output_container: &Store,
                node: &NodeProto,
                model_inputs: &Vec<ValueInfoProto>,
                model_initializers: &Vec<TensorProto>) {
    let x = output_container.lock().unwrap().get(&node.input[0]).map(|v| v.1.clone().expect("MaxPool input must be rank 4"))
        .unwrap_or_else(|| get_stored_tensor(0, node, model_inputs, model_initializers));
    let mut p = sys::b200_pool_params::default();
    p.auto_pad = sys::B200_PAD_VALID;                                  // default, max_pool_op.rs:88 (no pad promotion)
    for attr in &node.attribute {
        match attr.name.as_ref() {
            "auto_pad" => p.auto_pad = match std::str::from_utf8(&attr.s).unwrap() {
                "SAME_UPPER" => sys::B200_PAD_SAME_UPPER,
                "SAME_LOWER" => sys::B200_PAD_SAME_LOWER,
                "VALID" => sys::B200_PAD_VALID,
                "NOTSET" => sys::B200_PAD_NOTSET,                       // sic, max_pool_op.rs:96
                other => panic!("MaxPool Auto Pad specified not found: {}", other),
            },
            "kernel_shape" => { p.kernel = [attr.ints[0], attr.ints[1]]; }
            "pads" => { for (i, v) in attr.ints.iter().take(4).enumerate() { p.pads[i] = *v; } }
            "storage_order" => {}
            "strides" => { p.strides = [attr.ints[0], attr.ints[1]]; }
            _ => panic!("ATTRIBUTE NAME FOR MAX POOL NOT FOUND, {}", <String as AsRef<str>>::as_ref(&attr.name)),
        }
    }
    let ctx = default_context();
    let mut y = ptr::null_mut();
    check(unsafe { sys::b200_maxpool2d(ctx.raw(), x.raw(), &p, &mut y) }).unwrap_or_else(|e| panic!("b200rt: {}", e.message));
    output_container.lock().unwrap().insert(node.output[0].clone(), (None, Some(DeviceTensor::from_raw(ctx.clone(), y))));
}
